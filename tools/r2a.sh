# round 2, call A: GPU tests (new flip-aware full-size parity), backward PX sweep, ncu --set full of one whole step
tag=r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt
python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/${tag}_pytest_gpu_full.log 2>&1
tail -25 gpurun_out/${tag}_pytest_gpu_full.log
for px in 1 2 4; do
  AGS_BWD_PX=$px python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_px$px.json 2> gpurun_out/${tag}_bench_px$px.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${tag}_bench_px$px.json') if l.startswith('{')][-1])
k=d['kernels']
print('px$px step %.1f us e2e %.1f us' % (d['ms_per_step']*1e3, d['e2e']['ms_per_step']*1e3), ' '.join('%s=%.0f' % (n[:11], k[n]['ms']*1e3) for n in k), 'launches', d['gpu_launches'])
PY
done
ncu --set full --clock-control none --import-source on -k regex:"composite|project|scatter|loss|alloc|adam|tile_sort" -s 73 -c 11 -o gpurun_out/${tag}_ncu_step \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${tag}_ncu.err
ncu -i gpurun_out/${tag}_ncu_step.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_step_raw.csv 2>&1
ls -la gpurun_out | tail -12
