# round 2: GPU-class baseline arm (tests + timing)
tag=r2j
python -m pytest tests/test_gpu_naive_gpu.py -m gpu -q --tb=short -s > gpurun_out/${tag}_pytest_naive.log 2>&1
tail -30 gpurun_out/${tag}_pytest_naive.log
python bench.py --impl gpu_naive --steps 30 --warmup 3 > gpurun_out/${tag}_bench_gpu_naive.json 2> gpurun_out/${tag}_bench_gpu_naive.err
tail -c 1500 gpurun_out/${tag}_bench_gpu_naive.json; tail -5 gpurun_out/${tag}_bench_gpu_naive.err
