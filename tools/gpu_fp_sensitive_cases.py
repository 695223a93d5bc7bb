"""Runs the reference's floating-point-sensitive application cases through the CUDA engine and
reports how many rows of the reference's golden table are reproduced to the printed 4 decimals
(run on a GPU box: python tools/gpu_fp_sensitive_cases.py). The CPU oracle reproduces all rows of
both (tests/test_oracle_golden.py); for the GPU the chaotic one is a statistical check only."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lethe_b200 import abi  # noqa: E402
from lethe_b200.prm import load_prm  # noqa: E402
from lethe_b200.solver import DEMSolver  # noqa: E402

d = os.path.join(ROOT, "tests", "golden", "apps")
with open(os.path.join(d, "final_positions.json")) as f:
    gold_all = json.load(f)
for case in ("insert_list_3d_default_velocities", "solid_surface"):
    solver = DEMSolver(load_prm(os.path.join(d, case + ".prm")), engine_factory=lambda cfg: abi.load_engine(cfg), prm_directory=d)
    ids, x, props = solver.solve()
    gold = np.array([r[3:6] for r in gold_all[case]])
    err = np.abs(x - gold).max(axis=1)
    print(json.dumps({"case": case, "rows": len(ids), "rows_equal_to_4_decimals": int((err <= 0.5e-4 + 1e-9).sum()),
                      "max_error": float(err.max()), "median_error": float(np.median(err))}))
