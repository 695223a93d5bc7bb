"""Where does the CUDA engine stop being BITWISE equal to the oracle? Lock-step windows of the
two-sphere case insert_list_3d_default_velocities (free flight, the particle-particle collision,
the first wall impact, rolling): before every step the GPU gets the oracle's exact state, after it
forces, torques, positions, velocities and angular velocities are compared bit by bit.
    python tools/gpu_bitwise_probe.py            (on a GPU box)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lethe_b200 import abi  # noqa: E402
from lethe_b200.prm import load_prm  # noqa: E402
from lethe_b200.solver import box_wall_faces, list_insertion  # noqa: E402
from oracle import loader  # noqa: E402

d = os.path.join(ROOT, "tests", "golden", "apps")
p = load_prm(os.path.join(d, "insert_list_3d_default_velocities.prm"))
cfg = p.to_config(store_forces=True)
g, o = abi.load_engine(cfg), loader.oracle_engine(cfg)
for e in (g, o):
    e.set_walls(box_wall_faces(p.mesh, p.outlet_boundaries, p.periodic))
    e.add_particles(*list_insertion(p))


def ulps(a, b):
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    ia, ib = a.view(np.int64), b.view(np.int64)
    return int(np.abs(ia - ib).max()) if a.size else 0


def resync():
    ids, x, props = o.get_particles()
    g.step_host(0, ids, np.ascontiguousarray(x), np.ascontiguousarray(props))


it = 0
for start, length, name in ((0, 300, "free flight"), (2350, 500, "particle-particle collision"), (11000, 700, "first wall impact"),
                            (25000, 300, "rolling on the wall")):
    if start > it:
        o.step(start - it)
        g.step(start - it)
        it = start
    worst = {}
    first = None
    for k in range(length):
        resync()
        g.step(1)
        o.step(1)
        it += 1
        _, fg, tg = g.get_forces()
        _, fo, to = o.get_forces()
        _, xg, pg = g.get_particles()
        _, xo, po = o.get_particles()
        diff = {"force": ulps(fg, fo), "torque": ulps(tg, to), "x": ulps(xg, xo), "v": ulps(pg[:, 3:6], po[:, 3:6]),
                "omega": ulps(pg[:, 6:9], po[:, 6:9])}
        if first is None and any(diff.values()):
            first = (it, dict(diff), fg.tolist(), fo.tolist(), tg.tolist(), to.tolist())
        for key, val in diff.items():
            worst[key] = max(worst.get(key, 0), val)
    print(name, "steps", start + 1, "-", start + length, "max ulp distance:", worst)
    if first:
        print("   first difference at iteration", first[0], first[1])
        print("   F gpu", first[2], "\n   F ora", first[3], "\n   T gpu", first[4], "\n   T ora", first[5])
