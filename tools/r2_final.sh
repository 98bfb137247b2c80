# round 2, final single-GPU evidence: full GPU test run, the bench arms, ncu launch list + --set full of one training step
tag=r2final
python -m pytest tests -m gpu -q --tb=short > gpurun_out/${tag}_pytest_gpu_full.log 2>&1
tail -4 gpurun_out/${tag}_pytest_gpu_full.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_n1_steps20.json 2> gpurun_out/${tag}_bench_n1_steps20.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
python bench.py --impl gpu_naive --steps 30 --warmup 3 > gpurun_out/${tag}_bench_gpu_naive.json 2> gpurun_out/${tag}_bench_gpu_naive.err
python - <<PY
import json
for n in ['n1','n1_steps20','reference_arm','gpu_naive']:
    try:
        d=json.loads([l for l in open('gpurun_out/${tag}_bench_%s.json' % n) if l.startswith('{')][-1])
        k=d.get('kernels') or {}
        e=d.get('e2e') or {}
        print(n, 'value %.1f ms/step %.4f e2e %s' % (d['value'], d['ms_per_step'], e.get('ms_per_step')), ' '.join('%s=%.0f' % (a[:11], b['ms']*1e3) for a,b in k.items()),
              'roofline', {a: (round(b,3) if isinstance(b,float) else b) for a,b in (d.get('roofline') or {}).items() if a in ('kernel','achieved','frac','share_of_step')},
              'cpu', (d.get('cpu_baseline') or {}).get('value'), 'update', {a: round(b,2) for a,b in (d.get('update') or {}).items() if a.endswith('_ms') or a=='ms_per_keyframe'}, 'launches', d.get('gpu_launches'))
    except Exception as ex:
        print(n, 'FAILED', ex); print(open('gpurun_out/${tag}_bench_%s.err' % n).read()[-1200:])
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:"composite|project|scatter|loss|alloc|adam|clear|stage" -s 160 -c 33 --csv \
    --log-file gpurun_out/${tag}_launches_train_steps.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-arm-only > /dev/null 2> gpurun_out/${tag}_launches.err
ncu --set full --clock-control none --import-source on -k regex:"composite_bwd" -s 9 -c 1 -o gpurun_out/${tag}_ncu_composite_bwd \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-arm-only > /dev/null 2> gpurun_out/${tag}_ncu.err
ncu -i gpurun_out/${tag}_ncu_composite_bwd.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_composite_bwd_raw.csv 2>&1
ncu --set full --clock-control none -k regex:"composite_fwd" -s 20 -c 1 -o gpurun_out/${tag}_ncu_composite_fwd \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-arm-only > /dev/null 2>> gpurun_out/${tag}_ncu.err
ncu -i gpurun_out/${tag}_ncu_composite_fwd.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_composite_fwd_raw.csv 2>&1
ls -la gpurun_out | grep ${tag} | awk '{print $5, $9}'
