#!/bin/bash
# compute-sanitizer over a small lock-step parity run of the engine: memcheck (out-of-bounds / misaligned), racecheck (shared-memory
# hazards of k_step's queue / result buffer), initcheck (reads of uninitialised device memory)
mkdir -p gpurun_out
S=${1:-san}
cat > /tmp/san_case.py <<'P'
import sys; sys.path.insert(0, ".")
import numpy as np
from lethe_b200 import abi
from lethe_b200.solver import box_wall_faces
from tests.util import packing_parameters, random_packing
d = 0.005
for pp, rolling, periodic, asc in [("hertz_mindlin_limit_overlap", "constant", (0, 0, 0), False), ("hertz_JKR", "epsd", (1, 0, 0), False),
                                  ("hertz_mindlin_limit_overlap", "constant", (0, 0, 0), True)]:
    ids, x, props, extent = random_packing(8, d=d, spacing=0.98, jitter=0.08, poly=0.2, seed=7)
    p = packing_parameters(extent, d=d, pp_model=pp, rolling=rolling, periodic=periodic, surface_energy=0.05 if "JKR" in pp else 0.0)
    p.sparse_contacts = asc
    p.asc_granular_temperature_threshold = 1e-3
    e = abi.load_engine(p.to_config(store_forces=True))
    e.set_walls(box_wall_faces(p.mesh, p.outlet_boundaries, p.periodic))
    e.set_particles(ids, x, props)
    e.step(40)
    e.force_contact_search()
    e.step(5)
    e.synchronize_velocities()
    _, xx, _ = e.get_particles()
    print(pp, rolling, periodic, asc, "ok", e.get_stats().n_rebuilds, float(np.abs(xx).max()))
P
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_case.py > gpurun_out/sanitizer_${tool}_$S.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok " gpurun_out/sanitizer_${tool}_$S.log | tail -5
done
