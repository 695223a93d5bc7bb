#!/usr/bin/env python
"""Generates lethe_b200/data/periodic_cell_<n>.npy: the disordered 3-periodic unit cell that
bench.py tiles into config 5 (64 M-sphere periodic box) and cuts into the drum bed of config 2.

    python tools/make_periodic_cell.py [n_side=40] [phi=0.64] [--gpu]

workloads.grow_periodic_cell does the work (dilute jittered lattice grown to the final diameter
while the DEM engine integrates the collisions, then relaxed). By default the CPU oracle runs it
(no GPU in the build container; ~2 min for 64 000 spheres); --gpu uses the CUDA engine. The file
stores positions in units of the sphere diameter d (x/d, float64) and the cell edge L/d."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from lethe_b200 import workloads  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_side = int(args[0]) if args else 40
    phi = float(args[1]) if len(args) > 1 else 0.64
    if "--gpu" in sys.argv:
        from lethe_b200 import abi

        make = lambda cfg: abi.load_engine(cfg, 0)  # noqa: E731
    else:
        from oracle import loader

        loader.build()
        make = lambda cfg: loader.oracle_engine(cfg)  # noqa: E731
    d = 0.002
    t0 = time.time()
    x, L = workloads.grow_periodic_cell(make, n_side=n_side, d=d, phi=phi, seed=19, log=print)
    print(f"{len(x)} spheres in {time.time() - t0:.1f} s")
    out = os.path.join(ROOT, "lethe_b200", "data", f"periodic_cell_{n_side ** 3}.npz")
    np.savez_compressed(out, x_over_d=x / d, L_over_d=np.asarray(L) / d, phi=phi, seed=19)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
