python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 \
      > gpurun_out/r2final11_bench_n2.json 2> gpurun_out/r2final11_bench_n2.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2final11_bench_n2.json') if l.startswith('{')][-1])
e=d['e2e']
print('n2 step %.1f us | e2e %.1f us/step upd_ms %s cpu_ms %s dev_allocs %s' % (d['ms_per_step']*1e3, e['ms_per_step']*1e3, e.get("host_ms_per_update"), e.get("cpu_ms_per_update"), e.get("device_allocs_per_update")))
PY
