#!/bin/bash
# One-off check of the graphed training step under torchrun on 2 GPUs (run through gpurun --gpus 2), hard time limit.
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-baselines 2>gpurun_out/bench2_err.log | tail -1 > gpurun_out/bench_graph_2gpu.json
echo rc=$?
tail -3 gpurun_out/bench2_err.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_graph_2gpu.json").read())
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["eager_ms_per_step"], d["step_mode"][:12])
print(d.get("strong_scaling"))
print(d.get("render_sharded"))
print({k: (v.get("ms_per_step"), v.get("error")) for k, v in d.get("workloads", {}).items()})
PY
