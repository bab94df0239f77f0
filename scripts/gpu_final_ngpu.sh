#!/bin/bash
# Round-end measurement on N GPUs of one box under a hard time limit: bash scripts/gpu_final_ngpu.sh N
N=${1:-8}
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-baselines 2>gpurun_out/final${N}_err.log | tail -1 > gpurun_out/r02_bench_final_${N}gpu.json
echo "rc=$?"
grep -i -E "NVLS|error|Traceback" gpurun_out/final${N}_err.log | head -5
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_final_${N}gpu.json").read())
print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["eager_ms_per_step"])
print(d.get("strong_scaling"))
print(d.get("render_sharded"))
print({k: (v.get("ms_per_step"), v.get("value"), v.get("error")) for k, v in d.get("workloads", {}).items()})
PY
