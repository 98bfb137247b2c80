# round 2, 8-GPU call: dist check (world 2/4/8) + weak-scaling lines N=8/N=4 + segments + config[3] (1 frame/GPU) + config[4] at N=8
tag=r2p
python -m pytest tests/test_dist_gpu.py -m gpu -q --tb=short -s > gpurun_out/${tag}_pytest_dist.log 2>&1
tail -12 gpurun_out/${tag}_pytest_dist.log
for f in gpurun_out/dist_check_world*.log; do [ -f "$f" ] && cp $f gpurun_out/${tag}_$(basename $f); done
bench() {  # name, nproc, extra args
  name=$1; np=$2; shift; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $np --steps 20 --warmup 5 "$@" \
      > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_bench_$name.json') if l.startswith('{')][-1])
    e=d['e2e']
    print('$name n=%d value %.0f Mpix/s step %.1f us | e2e %.0f Mpix/s %.1f us/step upd_ms %s symm_allocs %s | roofline %s' % (d['n_gpus'], d['value'], d['ms_per_step']*1e3, e['value'], e['ms_per_step']*1e3, e.get('host_ms_per_update'), e.get('symmetric_allocations'), {k: (round(v,3) if isinstance(v,float) else v) for k,v in (d.get('roofline') or {}).items() if k in ('kernel','frac','achieved')}))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${tag}_bench_$name.err').read()[-2500:])
PY
}
bench n8 8
AGS_DIST_PROFILE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 60 --warmup 5 --device-arm-only \
      > /dev/null 2> gpurun_out/${tag}_segments_n8.err
grep "segments" gpurun_out/${tag}_segments_n8.err | head -8
bench config3_n8_1frame 8 --frames-per-gpu 1
bench config5_n8_1frame 8 --config 5 --frames-per-gpu 1
bench n4 4
