"""The `render_t22_quantised` leg of bench.py on its own (frame-level Mpix/s at log2_hashmap 22 for fp32 tables,
fake-quant in the gather and u8 code tables)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import indoor_nerf_b200 as pn  # noqa: E402
from indoor_nerf_b200 import model as pmodel, synthetic  # noqa: E402

pn.set_mlp_mode(os.environ.get("POCKETNERF_MLP", "bf16"))
print(json.dumps(bench.render_t22_leg(pn, pmodel, synthetic, torch.device("cuda", 0))))
