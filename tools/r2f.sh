# round 2, multi-GPU call: dist check under pytest (world = all GPUs of the box and below) + weak-scaling bench lines
# usage: bash tools/r2f.sh <ngpus> <tag>
n=$1; tag=$2
python -m pytest tests/test_dist_gpu.py -m gpu -q --tb=short -s > gpurun_out/${tag}_pytest_dist.log 2>&1
tail -25 gpurun_out/${tag}_pytest_dist.log
for f in gpurun_out/dist_check_world*.log; do [ -f "$f" ] && cp $f gpurun_out/${tag}_$(basename $f); done
bench() {  # name, nproc, extra args
  name=$1; np=$2; shift; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $np --steps 20 --warmup 5 "$@" \
      > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_bench_$name.json') if l.startswith('{')][-1])
    print('$name n=%d value %.0f Mpix/s step %.1f us | e2e %.0f Mpix/s %.1f us/step | roofline %s' % (d['n_gpus'], d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['e2e']['ms_per_step']*1e3, {k: (round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline') or {}).items() if k in ('kernel','frac','achieved')}))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${tag}_bench_$name.err').read()[-2500:])
PY
}
bench n$n $n
AGS_DIST_PROFILE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $n --steps 60 --warmup 5 --device-arm-only \
      > /dev/null 2> gpurun_out/${tag}_segments_n$n.err
grep "segments" gpurun_out/${tag}_segments_n$n.err | head -3
AGS_DIST_BARRIER=1 bench n${n}_barriers $n
if [ "$n" = "8" ]; then bench config3_n8_1frame 8 --frames-per-gpu 1; fi
