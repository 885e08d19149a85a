#!/bin/bash
# Run on the GPU box (under gpurun): tests, bench, ncu launch list, ncu full capture of the fused kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
fi
python bench.py --steps ${STEPS:-20} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --profile > gpurun_out/ncu_launch.log 2>&1
  tail -2 gpurun_out/ncu_launch.log
  ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 3 -c 1 -f -o gpurun_out/prof_fused \
      python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
  ls -la gpurun_out
fi
