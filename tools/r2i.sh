# round 2, session 2: host-path rewrite (numpy fetch / heap LPT / symmetric head-room) check on 2 GPUs
# usage: bash tools/r2i.sh <ngpus> <tag>
n=$1; tag=$2
python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/${tag}_pytest_gpu_full.log 2>&1
tail -6 gpurun_out/${tag}_pytest_gpu_full.log
for f in gpurun_out/dist_check_world*.log; do [ -f "$f" ] && cp $f gpurun_out/${tag}_$(basename $f); done
summ() {
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_bench_$1.json') if l.startswith('{')][-1])
    e=d['e2e']
    print('$1 n=%d value %.0f Mpix/s step %.1f us | e2e %.0f Mpix/s %.1f us/step upd_ms %s symm_allocs %s' % (d['n_gpus'], d['value'], d['ms_per_step']*1e3, e['value'], e['ms_per_step']*1e3, e.get('host_ms_per_update'), e.get('symmetric_allocations')))
    print('   kernels us:', {k: round(v['ms']*1e3,1) for k,v in (d.get('kernels') or {}).items()})
    if d.get('update'): print('   update:', {k: (round(v,2) if isinstance(v,float) else v) for k,v in d['update'].items() if k!='what'})
except Exception as e:
    print('$1 FAILED', e); print(open('gpurun_out/${tag}_bench_$1.err').read()[-2500:])
PY
}
python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; summ n1
if [ "$n" != "1" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 40 --warmup 5 \
      > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err; summ n$n
AGS_DIST_PROFILE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $n --steps 60 --warmup 5 --device-arm-only \
      > /dev/null 2> gpurun_out/${tag}_segments_n$n.err
grep "segments" gpurun_out/${tag}_segments_n$n.err | head -3
fi
