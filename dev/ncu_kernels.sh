#!/bin/bash
# Full ncu captures of the hot hand-written kernels (one launch each). Usage: bash dev/ncu_kernels.sh <tag>
TAG=${1:-a}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'warp_photo_fwd_kernel' -s 3 -c 1 \
  -o gpurun_out/warpfwd_$TAG -f python dev/kernel_bench.py --what warp2 --reps 2 > gpurun_out/ncu_warpfwd_$TAG.log 2>&1
echo "ncu warp fwd exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'warp_photo_bwd_kernel' -s 2 -c 1 \
  -o gpurun_out/warpbwd_$TAG -f python dev/kernel_bench.py --what warp2 --reps 2 > gpurun_out/ncu_warpbwd_$TAG.log 2>&1
echo "ncu warp bwd exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv_wino_kernel|conv_wgrad_kernel' -s 90 -c 6 \
  -o gpurun_out/conv_$TAG -f python dev/kernel_bench.py --what conv --reps 2 > gpurun_out/ncu_conv_$TAG.log 2>&1
echo "ncu conv exit $?"
