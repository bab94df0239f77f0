#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench; everything logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 60 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -n 15 gpurun_out/smoke.log
timeout 1200 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"
tail -n 5 gpurun_out/bench.err
cat gpurun_out/bench.log
