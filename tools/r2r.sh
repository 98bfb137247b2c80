tag=r2r
seg() { # name env
  name=$1; shift
  env "$@" AGS_DIST_PROFILE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 60 --warmup 5 --quick \
      > gpurun_out/${tag}_quick_$name.json 2> gpurun_out/${tag}_segments_$name.err
  echo "== $name"; tail -1 gpurun_out/${tag}_quick_$name.json; grep "segments" gpurun_out/${tag}_segments_$name.err | head -2
}
seg side AGS_X=1
seg noside AGS_DIST_SIDE=0
seg side_again AGS_X=1
