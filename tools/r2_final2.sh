# round 2, final 2-GPU check: dist check under pytest (world 2) + the N=2 weak-scaling line in the driver's shape
tag=r2final
python -m pytest tests/test_dist_gpu.py -m gpu -q --tb=short -s > gpurun_out/${tag}_pytest_dist_n2.log 2>&1
tail -6 gpurun_out/${tag}_pytest_dist_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 \
      > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_bench_n2.json') if l.startswith('{')][-1])
    e=d['e2e']
    print('n2 value %.0f Mpix/s step %.1f us | e2e %.0f Mpix/s %.1f us/step upd_ms %s cpu_ms %s dev_allocs %s symm_allocs %s' % (d['value'], d['ms_per_step']*1e3, e['value'], e['ms_per_step']*1e3, e.get("host_ms_per_update"), e.get("cpu_ms_per_update"), e.get("device_allocs_per_update"), e.get("symmetric_allocations")))
    print('   kernels us:', {k: round(v['ms']*1e3,1) for k,v in (d.get('kernels') or {}).items()})
except Exception as e:
    print('n2 FAILED', e); print(open('gpurun_out/${tag}_bench_n2.err').read()[-2500:])
PY
