#!/bin/bash
# One gpurun call: launch list of a short bench run + one full ncu capture of the hot kernels.
# (each ncu pass runs only after the identical plain command exited 0, per B200_PROFILING.md)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-baselines"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLIST:-3000} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-hash_fwd_kernel|mlp_tc_fwd_kernel|mlp_tc_bwd_kernel}" -s ${KSKIP:-9} -c ${KCOUNT:-5} -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/
