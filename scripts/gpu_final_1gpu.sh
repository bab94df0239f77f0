#!/bin/bash
# Round-end measurement on one GPU: the default bench (with the reference arms), the reference arm alone, then the ncu
# launch list of a short bench run (after that command exited 0 without ncu).
python bench.py 2>gpurun_out/final1_err.log | tail -1 > gpurun_out/r02_bench_final_1gpu.json; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/final1_ref_err.log | tail -1 > gpurun_out/r02_bench_final_reference_arm.json; echo "ref rc=$?"
python bench.py --steps 3 --warmup 3 --no-baselines --no-extras > gpurun_out/short.json 2>gpurun_out/short_err.log; echo "short rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/r02_launches_step_final.csv \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-extras > gpurun_out/ncu_stdout.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02_launches_step_final.csv
