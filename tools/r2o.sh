# round 2 (1 GPU): reference call site over the drop-in, config[4] line at N=1, vectorised vis-count check, config[2] update loop
tag=r2o
python -m pytest tests/test_reference_callsite_gpu.py -m gpu -q --tb=short -s > gpurun_out/${tag}_pytest_reference_callsite.log 2>&1
tail -15 gpurun_out/${tag}_pytest_reference_callsite.log
python -m pytest tests/test_loss_adam_gpu.py tests/test_train_gpu.py -m gpu -q --tb=short > gpurun_out/${tag}_pytest_loss_train.log 2>&1
tail -3 gpurun_out/${tag}_pytest_loss_train.log
python bench.py --config 5 --frames-per-gpu 1 --steps 20 --warmup 5 --no-cpu-baseline --no-update-profile > gpurun_out/${tag}_bench_config5_n1.json 2> gpurun_out/${tag}_bench_config5_n1.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_bench_config5_n1.json') if l.startswith('{')][-1])
    e=d['e2e']
    print('config5_n1 value %.0f Mpix/s step %.1f us | e2e %.0f Mpix/s %.1f us/step | inst %s' % (d['value'], d['ms_per_step']*1e3, e['value'], e['ms_per_step']*1e3, d['config'].get('instances_per_step')))
    print('   kernels us:', {k: round(v['ms']*1e3,1) for k,v in (d.get('kernels') or {}).items()})
except Exception as e:
    print('config5_n1 FAILED', e); print(open('gpurun_out/${tag}_bench_config5_n1.err').read()[-2500:])
PY
