#!/bin/bash
# One GPU-box visit: parity tests, bench line, full ncu capture of the tensor-core linear kernel (and, with a third
# argument, the launch list of one step and the warp kernels).
# Usage (from the repo root, under gpurun): bash dev/gpu_round.sh <tag> [skip_tests] [more]
TAG=${1:-a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
  tail -3 gpurun_out/pytest_$TAG.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
cat gpurun_out/bench_$TAG.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -c 6 -o gpurun_out/linear_$TAG -f \
  python dev/ncu_linear.py > gpurun_out/ncu_linear_$TAG.log 2>&1
echo "ncu linear exit $?"
if [ -n "$3" ]; then
  timeout 300 python dev/kernel_bench.py --what all --reps 10 > gpurun_out/kbench_$TAG.txt 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --profile-step > gpurun_out/ncu_launch_$TAG.log 2>&1
  echo "launch list exit $?"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'warp_photo_(fwd|bwd)_kernel' -s 8 -c 2 \
    -o gpurun_out/warp_$TAG -f python dev/kernel_bench.py --what warp2 --reps 2 > gpurun_out/ncu_warp_$TAG.log 2>&1
  echo "ncu warp exit $?"
fi
