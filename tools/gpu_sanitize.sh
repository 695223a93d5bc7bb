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
# the streamed host step (upload stages, partial step launches through block lists, segment download) and the
# device-side transfer order, with pageable rows and with page-locked rows written by the device
import os
import torch
os.environ.update(LETHE_DEM_HOST_PIPELINE_MIN_ROWS="1", LETHE_DEM_HOST_STAGES="3", LETHE_DEM_HOST_SEG_ROWS="256")
for periodic, zero_copy in [((0, 0, 0), False), ((1, 1, 1), True)]:
    os.environ["LETHE_DEM_HOST_ZEROCOPY"] = "1" if zero_copy else "0"
    ids, x, props, extent = random_packing(14, d=d, spacing=0.99, jitter=0.06, seed=3)
    props[:, 3:6] += np.random.default_rng(7).normal(0, 0.2, (len(ids), 3))
    p = packing_parameters(extent, d=d, g=(0, 0, 0) if periodic[0] else (0, 0, -9.81), periodic=periodic, cell=extent[0] / 7)
    e = abi.load_engine(p.to_config())
    if not periodic[0]:
        e.set_walls(box_wall_faces(p.mesh, p.outlet_boundaries, p.periodic))
    e.set_particles(ids, x, props)
    e.step(3)
    rid, rows = e.get_state_rows()
    if zero_copy:
        keep = torch.from_numpy(rows).pin_memory()
        rows = keep.numpy()
    e.step_host_state(1, rid, rows)
    for _ in range(60):
        e.step_host_state(1, None, rows)
    print("streamed host step", periodic, zero_copy, "ok", e.get_stats().n_rebuilds, e.host_pipeline_stats(), float(np.abs(rows).max()))
P
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_case.py > gpurun_out/sanitizer_${tool}_$S.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok " gpurun_out/sanitizer_${tool}_$S.log | tail -5
done
