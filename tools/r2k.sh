# round 2: tensor-core reduction in composite_bwd (AGS_BWD_RED=4) -- parity + A/B timing
tag=r2k
AGS_BWD_RED=4 python -m pytest tests/test_rasterizer_gpu.py -m gpu -q --tb=short -x -s > gpurun_out/${tag}_pytest_red4.log 2>&1
tail -40 gpurun_out/${tag}_pytest_red4.log
run() {  # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_bench_$name.json') if l.startswith('{')][-1])
    k=d['kernels']
    print('$name step %.1f us e2e %.1f us/step' % (d['ms_per_step']*1e3, d['e2e']['ms_per_step']*1e3), ' '.join('%s=%.0f' % (n[:11], k[n]['ms']*1e3) for n in k), 'loss_last', d['config']['loss_last'])
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${tag}_bench_$name.err').read()[-1500:])
PY
}
run red3 AGS_BWD_RED=3
run red4 AGS_BWD_RED=4
run red4_minb5 AGS_BWD_RED=4 AGS_B200_LIB=$PWD/variants/mma5.so
