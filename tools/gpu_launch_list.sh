# ncu launch list of OUR kernels over the last training steps of a short bench run
tag=$1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"project_|alloc_kernel|scatter_kernel|tile_sort|composite_|loss_|zero_grads|adam_kernel|stage_cameras" -s 150 -c 36 --csv \
    --log-file gpurun_out/${tag}_launches_train_steps.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -c . gpurun_out/${tag}_launches_train_steps.csv
