# refresh the judged evidence: GPU tests, bench line, reference arm, ncu --set full of the dominant kernel
tag=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/${tag}_pytest_gpu.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_ref.err
ncu --set full --clock-control none --import-source on -k regex:"composite_bwd_kernel" -s 4 -c 1 -o gpurun_out/${tag}_ncu_full_composite_bwd \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/${tag}_ncu_full_composite_bwd.ncu-rep --page details > gpurun_out/${tag}_ncu_full_composite_bwd_summary.txt 2>&1
ncu -i gpurun_out/${tag}_ncu_full_composite_bwd.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_composite_bwd_raw.csv 2>&1
cat gpurun_out/${tag}_pytest_gpu.log; tail -c 300 gpurun_out/${tag}_bench_n1.json
