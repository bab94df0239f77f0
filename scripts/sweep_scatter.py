"""Sweep of the scatter's two knobs on the fused backward (dense and 10 %-dense cotangents): PN_SCATTER_DIRECT = run-head
count above which a warp issues its reductions directly, PN_SCATTER_MAXLEN = longest sub-run the shuffle tree reduces."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for direct, maxlen in ((20, 32), (12, 32), (6, 32), (0, 32), (20, 8), (20, 4), (12, 4), (20, 2), (6, 4)):
    for dens in ("1.0", "0.1"):
        env = dict(os.environ, PN_FIELD_BWD="v4", PN_SCATTER_DIRECT=str(direct), PN_SCATTER_MAXLEN=str(maxlen))
        r = subprocess.run(["timeout", "200", sys.executable, os.path.join(ROOT, "scripts", "bench_field_kernels.py"), "--density", dens],
                           env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(json.dumps({"direct": direct, "maxlen": maxlen, "density": dens, "S192_bwd_ms": round(d["S192"]["bwd_ms"], 3),
                              "S64_bwd_ms": round(d["S64"]["bwd_ms"], 3), "g0_abs": d["S192"]["check"]["g0_abs"],
                              "g15_abs": d["S192"]["check"]["g15_abs"]}), flush=True)
        except Exception as ex:
            print(json.dumps({"direct": direct, "maxlen": maxlen, "error": r.stderr[-300:]}), flush=True)
