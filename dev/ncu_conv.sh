#!/bin/bash
# ncu --set full of the Winograd kernel (forward) and the Winograd weight-gradient kernel on the motion-decoder
# level-4 layer (64->64 @96x320, bs32).  Usage: bash dev/ncu_conv.sh <tag>
TAG=${1:-a}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv_wino_kernel' -s 46 -c 1 \
  -o gpurun_out/wino_$TAG -f python dev/kernel_bench.py --what conv --reps 2 > gpurun_out/ncu_wino_$TAG.log 2>&1
echo "ncu wino exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv_wgrad_wino_kernel' -s 22 -c 1 \
  -o gpurun_out/wgrad_$TAG -f python dev/kernel_bench.py --what conv --reps 2 > gpurun_out/ncu_wgrad_$TAG.log 2>&1
echo "ncu wgrad exit $?"
