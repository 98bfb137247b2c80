# round 2, call D: retuned K1, 32x16 loss tile, PX2 at 8 CTAs/SM; ncu of one timed step (per-kernel summary)
tag=r2d
python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/${tag}_pytest_gpu_full.log 2>&1
tail -6 gpurun_out/${tag}_pytest_gpu_full.log
run() {  # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_bench_$name.json') if l.startswith('{')][-1])
    k=d['kernels']
    print('$name step %.1f us e2e %.1f us/step' % (d['ms_per_step']*1e3, d['e2e']['ms_per_step']*1e3), ' '.join('%s=%.0f' % (n[:11], k[n]['ms']*1e3) for n in k), 'launches', d['gpu_launches'], 'update', {a: round(b,2) for a,b in d.get('update',{}).items() if a.endswith('_ms') or a=='ms_per_keyframe'})
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${tag}_bench_$name.err').read()[-1500:])
PY
}
run default AGS_X=0
run px2mb8 AGS_B200_LIB=$PWD/variants/px2mb8.so
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_steps20.json 2> gpurun_out/${tag}_bench_steps20.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/${tag}_bench_steps20.json') if l.startswith('{')][-1])
print('steps20: step %.1f us, e2e %.1f us/step (%d updates, N end %d)' % (d['ms_per_step']*1e3, d['e2e']['ms_per_step']*1e3, d['e2e']['updates'], d['e2e']['gaussians_end']))"
ncu --set full --clock-control none --import-source on -k regex:"composite|project|scatter|loss|alloc|adam|clear" -s 155 -c 10 -o gpurun_out/${tag}_ncu_step \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-arm-only > /dev/null 2> gpurun_out/${tag}_ncu.err
ncu -i gpurun_out/${tag}_ncu_step.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_step_raw.csv 2>&1
ls -la gpurun_out | tail -6
