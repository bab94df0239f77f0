#!/bin/bash
# Round-end measurement on one GPU: the default bench (with the reference arms) and the reference arm alone.
python bench.py 2>gpurun_out/final1_err.log | tail -1 > gpurun_out/r02_bench_final_1gpu.json; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/final1_ref_err.log | tail -1 > gpurun_out/r02_bench_final_reference_arm.json; echo "ref rc=$?"
