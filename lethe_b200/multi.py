"""Host side of the slab decomposition (one process per GPU): cut planes on grid-cell
boundaries along one axis, particle ownership, communicator bootstrap.

Mirrors what p4est's space-filling-curve partition gives the reference for a box-like
domain (include/dem/dem.h:245-276): contiguous blocks of cells per rank, one ghost cell
layer; here 1-D slabs (two NVLink peers per GPU)."""
from __future__ import annotations

import numpy as np

from . import abi


def slab_bounds(n_cells_axis: int, world: int):
    """Equal-width cell ranges [lo, hi) per rank along the slab axis."""
    if n_cells_axis < 2 * world:
        raise abi.DEMError(f"{n_cells_axis} cell layers cannot be split into {world} slabs of >= 2 layers")
    edges = [(n_cells_axis * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def balanced_slab_bounds(cell_axis_index: np.ndarray, n_cells_axis: int, world: int):
    """Cut planes that give every rank about the same number of particles (the role of the
    reference's particle-weighted repartition, source/dem/load_balancing.cc)."""
    hist = np.bincount(cell_axis_index, minlength=n_cells_axis).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(hist)])
    total = cum[-1]
    edges = [0]
    for r in range(1, world):
        target = total * r / world
        e = int(np.searchsorted(cum, target))
        e = max(edges[-1] + 2, min(e, n_cells_axis - 2 * (world - r)))
        edges.append(e)
    edges.append(n_cells_axis)
    return [(edges[r], edges[r + 1]) for r in range(world)]


def owner_mask(x: np.ndarray, mesh, axis: int, lo: int, hi: int):
    h = mesh.cell_size[axis]
    c = np.floor((x[:, axis] - mesh.lo[axis]) / h).astype(np.int64)
    return (c >= lo) & (c < hi)


def create_slab_engine(workload, rank: int, world: int, device: int, dist=None, axis: int = 0, store_forces=False, balanced=True,
                       precision="f64"):
    """Engine of rank `rank` holding its slab of `workload`; returns (engine, n_local)."""
    p = workload.params
    mesh = p.mesh
    if balanced:
        h = mesh.cell_size[axis]
        ca = np.clip(np.floor((workload.x[:, axis] - mesh.lo[axis]) / h).astype(np.int64), 0, mesh.n[axis] - 1)
        bounds = balanced_slab_bounds(ca, mesh.n[axis], world)
    else:
        bounds = slab_bounds(mesh.n[axis], world)
    lo, hi = bounds[rank]
    cfg = p.to_config(store_forces=store_forces, slab=(axis, lo, hi), precision=precision)
    engine = abi.load_engine(cfg, device)
    if dist is not None:
        obj = [abi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        nccl_id = obj[0]
    else:
        raise abi.DEMError("create_slab_engine needs torch.distributed to bootstrap the NCCL communicator")
    engine.comm_init(rank, world, nccl_id)
    engine.set_walls(workload.faces)
    for m in workload.motions:
        engine.set_boundary_motion(*m)
    if getattr(workload, "floating_walls", None):
        fw = workload.floating_walls
        engine.set_floating_walls([w[0] for w in fw], [w[1] for w in fw], [w[2] for w in fw], [w[3] for w in fw])
    for sd in getattr(workload, "solids", []):
        engine.add_solid_surface(*sd)  # every rank holds every solid
    mask = owner_mask(workload.x, mesh, axis, lo, hi)
    engine.set_particles(workload.ids[mask], workload.x[mask], workload.props[mask])
    engine.slab = (axis, lo, hi)
    return engine, int(mask.sum())
