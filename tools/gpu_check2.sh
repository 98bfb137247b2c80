python bench.py --no-cpu-baseline --steps 300 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
python tools/profile_update.py > gpurun_out/s3_update.log 2>&1
python tools/profile_spawn.py > gpurun_out/s3_spawn.log 2>&1
python -m pytest tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -3
cat gpurun_out/s3_update.log gpurun_out/s3_spawn.log
