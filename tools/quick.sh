# quick A/B on one GPU: bash tools/quick.sh <tag> ["ENV=.. ENV=.." ...]   (each quoted group = one run)
tag=$1; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  echo "== run $i: $envs"
  env $envs python bench.py --steps 100 --warmup 5 --quick 2> gpurun_out/${tag}_quick_$i.err | tee -a gpurun_out/${tag}_quick.log | tail -1
done
