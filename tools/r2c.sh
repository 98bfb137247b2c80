# round 2, call C: full GPU test run + ncu --set full (with source) of the kernels that are off their estimate
tag=r2c
python -m pytest tests -m gpu -q --tb=short > gpurun_out/${tag}_pytest_gpu_full.log 2>&1
tail -12 gpurun_out/${tag}_pytest_gpu_full.log
# the first timed step starts after: workload build (8+22 renders x 4 of our render kernels + postprocess) and the warm-up steps;
# match kernels by name and take launches late in the run instead of counting
for k in loss_fused_kernel project_fwd_kernel composite_bwd_kernel composite_fwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 45 -c 1 -o gpurun_out/${tag}_ncu_$k \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${tag}_ncu_$k.err
  ncu -i gpurun_out/${tag}_ncu_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_${k}_raw.csv 2>&1
  ncu -i gpurun_out/${tag}_ncu_$k.ncu-rep --page source --csv > gpurun_out/${tag}_ncu_${k}_src.csv 2>&1
done
ls -la gpurun_out | tail -14
