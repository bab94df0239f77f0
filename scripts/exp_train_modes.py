"""Learning curves of the arithmetic modes on the test scene of tests/test_gpu_training.py (PSNR per window)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_gpu_training import _scene, _train  # noqa: E402

mode = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 600
scene = _scene()
seed = int(os.environ.get("PN_SEED", "0"))
l, p, _ = _train(mode, steps, scene, seed=seed)
w = [(20, 40), (40, 80), (80, 150), (150, 250), (250, 400), (400, 500), (500, 600)]
print(mode, "seed", seed, os.environ.get("PN_BF16_UNFUSED"), os.environ.get("PN_EXP"), " ".join("%.2f" % float(np.mean(p[a:b])) for a, b in w if b <= steps))
