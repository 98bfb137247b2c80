# round 2, call E: folded-select butterfly in composite_bwd (RED=3) + new record layout
tag=r2e
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
run red3 AGS_BWD_RED=3
run red2 AGS_BWD_RED=2
