# round 2: config[4] (1M Gaussians, 1920x1080, 1 frame per GPU) at N=1 and N=2 + the reference's own
# render_cuda_core over the drop-in (needs tests/_ref_tmp, shipped as untracked test input)
tag=r2n
python -m pytest tests/test_reference_callsite_gpu.py -m gpu -q --tb=short -s > gpurun_out/${tag}_pytest_reference_callsite.log 2>&1
tail -15 gpurun_out/${tag}_pytest_reference_callsite.log
summ() {
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_bench_$1.json') if l.startswith('{')][-1])
    e=d['e2e']
    print('$1 n=%d value %.0f Mpix/s step %.1f us | e2e %.0f Mpix/s %.1f us/step | inst %s' % (d['n_gpus'], d['value'], d['ms_per_step']*1e3, e['value'], e['ms_per_step']*1e3, d['config'].get('instances_per_step')))
    print('   kernels us:', {k: round(v['ms']*1e3,1) for k,v in (d.get('kernels') or {}).items()})
except Exception as e:
    print('$1 FAILED', e); print(open('gpurun_out/${tag}_bench_$1.err').read()[-2500:])
PY
}
python bench.py --config 5 --frames-per-gpu 1 --steps 20 --warmup 5 --no-cpu-baseline --no-update-profile > gpurun_out/${tag}_bench_config5_n1.json 2> gpurun_out/${tag}_bench_config5_n1.err; summ config5_n1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --config 5 --frames-per-gpu 1 --steps 20 --warmup 5 \
      > gpurun_out/${tag}_bench_config5_n2.json 2> gpurun_out/${tag}_bench_config5_n2.err; summ config5_n2
