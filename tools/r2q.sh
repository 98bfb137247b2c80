# round 2 (2 GPUs): sharded prune-pass renders, update_utility, config[4] sweep lines with 4 frames per GPU at N=1/2
tag=r2q
python -m pytest tests/test_dist_gpu.py tests/test_train_gpu.py tests/test_mapops_gpu.py tests/test_fullsize_gpu.py -m gpu -q --tb=short -x > gpurun_out/${tag}_pytest.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
for f in gpurun_out/dist_check_world*.log; do [ -f "$f" ] && cp $f gpurun_out/${tag}_$(basename $f); done
summ() {
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_bench_$1.json') if l.startswith('{')][-1])
    e=d['e2e']
    print('$1 n=%d value %.0f Mpix/s step %.1f us | e2e %.0f Mpix/s %.1f us/step upd_ms %s | inst %s' % (d['n_gpus'], d['value'], d['ms_per_step']*1e3, e['value'], e['ms_per_step']*1e3, e.get('host_ms_per_update'), d['config'].get('instances_per_step')))
    print('   kernels us:', {k: round(v['ms']*1e3,1) for k,v in (d.get('kernels') or {}).items()})
except Exception as e:
    print('$1 FAILED', e); print(open('gpurun_out/${tag}_bench_$1.err').read()[-2500:])
PY
}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 \
      > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err; summ n2
python bench.py --config 5 --frames-per-gpu 4 --steps 20 --warmup 5 --no-cpu-baseline --no-update-profile > gpurun_out/${tag}_bench_config5_b4_n1.json 2> gpurun_out/${tag}_bench_config5_b4_n1.err; summ config5_b4_n1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --config 5 --frames-per-gpu 4 --steps 20 --warmup 5 \
      > gpurun_out/${tag}_bench_config5_b4_n2.json 2> gpurun_out/${tag}_bench_config5_b4_n2.err; summ config5_b4_n2
