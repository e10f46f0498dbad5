#!/bin/bash
# BASELINE configs[2] (ViT-L/16 r32, GLOBAL batch 2048, strong scaling at 1/2/4/8 GPUs) and configs[3] (ViT-H/14 r32 at 8 GPUs)
# on ONE 8-GPU box:  gpurun --gpus 8 -- bash tools/scale_c3c4.sh   -> gpurun_out/r02_scale_*.json
set -u
out=gpurun_out
run() {  # run <name> <gpu list> <nproc> <port> <bench args...>
  name=$1; gpus=$2; n=$3; port=$4; shift 4
  if [ "$n" = 1 ]; then
    CUDA_VISIBLE_DEVICES=$gpus timeout 400 python bench.py --gpus 1 --no-cpu-baseline "$@" > $out/r02_scale_$name.json 2> $out/r02_scale_$name.err
  else
    CUDA_VISIBLE_DEVICES=$gpus timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus $n "$@" > $out/r02_scale_$name.json 2> $out/r02_scale_$name.err
  fi
  echo "$name rc=$? $(head -c 260 $out/r02_scale_$name.json)"
}
run vitl_n8 0,1,2,3,4,5,6,7 8 29601 --config vitl16_r32 --global-batch 2048 --steps 10 --warmup 3
run vitl_n4 0,1,2,3 4 29602 --config vitl16_r32 --global-batch 2048 --steps 10 --warmup 3 &
run vitl_n2 4,5 2 29603 --config vitl16_r32 --global-batch 2048 --steps 10 --warmup 3 &
run vitl_n1 6 1 29604 --config vitl16_r32 --global-batch 2048 --steps 5 --warmup 3 &
wait
run vith_n8 0,1,2,3,4,5,6,7 8 29605 --config vith14_r32 --steps 10 --warmup 3
run vitb_n8 0,1,2,3,4,5,6,7 8 29606 --steps 20 --warmup 5
