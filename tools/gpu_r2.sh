#!/bin/bash
# Round-2 GPU session script (run under gpurun): tests, bench, optional ncu.  Everything goes to gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout ${TEST_TIMEOUT:-1500} python -m pytest ${PYTEST_ARGS:-tests} -m gpu -x -q > gpurun_out/pytest.log 2>&1; tail -25 gpurun_out/pytest.log
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 600 python bench.py --steps ${STEPS:-20} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
fi
if [ "${SKIP_NCU:-1}" != "1" ]; then
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --profile > gpurun_out/ncu_launch.log 2>&1
  tail -2 gpurun_out/ncu_launch.log
  ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 3 -c 1 -f -o gpurun_out/prof_fused \
      python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
fi
for x in ${EXTRA:-}; do timeout 600 python $x > gpurun_out/$(basename $x .py).log 2>&1; tail -30 gpurun_out/$(basename $x .py).log; done
ls -la gpurun_out | head -40
