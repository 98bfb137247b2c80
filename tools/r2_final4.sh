tag=r2final
for i in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$i bench.py --gpus 2 --steps 20 --warmup 5 \
      > gpurun_out/${tag}_bench_n2_run$i.json 2> gpurun_out/${tag}_bench_n2_run$i.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_bench_n2_run$i.json') if l.startswith('{')][-1])
    e=d['e2e']
    print('n2 run $i value %.0f Mpix/s step %.1f us | e2e %.0f Mpix/s %.1f us/step upd_ms %s symm_allocs %s' % (d['value'], d['ms_per_step']*1e3, e['value'], e['ms_per_step']*1e3, e.get('host_ms_per_update'), e.get('symmetric_allocations')))
except Exception as e:
    print('n2 FAILED', e); print(open('gpurun_out/${tag}_bench_n2_run$i.err').read()[-2500:])
PY
done
