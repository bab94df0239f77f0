"""Run `bench.py --no-baselines` once per value of an environment tuning knob (the kernels read their knobs once per
process, so each value needs its own process) and print value / ms_per_step / fused-kernel times side by side.

    python scripts/sweep_knobs.py PN_SCATTER_SPLIT 6 7 8 9          # backward: levels scattered by half 0
    python scripts/sweep_knobs.py PN_FWD_CTAS 3 4                    # fused forward: CTAs per SM
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
knob, values = sys.argv[1], sys.argv[2:]
rows = []
for v in values:
    env = dict(os.environ, **{knob: v})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-baselines"],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        fk = d.get("fused_kernels", {})
        rows.append({knob: v, "rays_per_s": d["value"], "ms_per_step": d["ms_per_step"],
                     "field_fwd_ms": fk.get("field_fwd", {}).get("ms"), "field_bwd_ms": fk.get("field_bwd", {}).get("ms"),
                     "render_mpix": d.get("render", {}).get("mpix_per_s")})
    except Exception as ex:
        rows.append({knob: v, "error": repr(ex), "stderr": r.stderr[-400:]})
    print(json.dumps(rows[-1]), flush=True)
