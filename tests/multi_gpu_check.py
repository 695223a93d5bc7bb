"""Run under torchrun on >= 2 GPUs: slab-decomposed run vs the single-domain CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lethe_b200 import multi, workloads  # noqa: E402
from oracle import loader  # noqa: E402


def gather_rows(arrs, world):
    objs = [None] * world
    dist.all_gather_object(objs, arrs)
    return objs


def run_case(name, w, checkpoints, rank, world, local_rank, tol=(1e-9, 1e-6), load_balance=None):
    """Slab-decomposed run vs the single-domain oracle (and, for reference, vs a single-GPU
    engine on rank 0) at several horizons: a decomposition bug shows up at the first
    checkpoint, chaotic growth of summation-order noise only at the late ones."""
    eng, n_local = multi.create_slab_engine(w, rank, world, local_rank, dist, store_forces=True, balanced=False)
    expect_recut = bool(load_balance) and load_balance[0] != "dynamic_with_sparse_contacts"
    if load_balance:
        eng.set_load_balancing(*load_balance)
        if load_balance[0] == "dynamic_with_sparse_contacts":
            # the weights of load_balancing_mobility_status.prm: particles of active / inactive cells count for 1/1000
            eng.set_load_balancing_weights(2000.0, 1000.0, 0.001, 0.001)
    slab0 = eng.get_slab()
    o = single = None
    if rank == 0:
        o = loader.oracle_engine(w.params.to_config(store_forces=True))
        w.install(o)
        from lethe_b200 import abi

        single = abi.load_engine(w.params.to_config(store_forces=True), local_rank)
        w.install(single)
    ok = True
    done = 0
    for k, steps in enumerate(checkpoints):
        eng.step(steps - done)
        ids, x, props = eng.get_particles()
        fid, f, t = eng.get_forces()
        pi, pj, _ = eng.get_pairs()
        st = eng.get_stats()
        allrows = gather_rows((ids, x, props, f, t, pi, pj, st.n_rebuilds), world)
        if rank == 0:
            o.step(steps - done)
            single.step(steps - done)
            oid, ox, op = o.get_particles()
            _, of, ot = o.get_forces()
            qi, qj, _ = o.get_pairs()
            _, sx, _ = single.get_particles()
            _, sf, _ = single.get_forces()
            gid = np.concatenate([r[0] for r in allrows])
            order = np.argsort(gid)
            gx = np.concatenate([r[1] for r in allrows])[order]
            gf = np.concatenate([r[3] for r in allrows])[order]
            gt = np.concatenate([r[4] for r in allrows])[order]
            pairs = set()
            for r in allrows:
                pairs.update(zip(r[5].tolist(), r[6].tolist()))
            opairs = set(zip(qi.tolist(), qj.tolist()))
            ok &= np.array_equal(gid[order], oid)
            ex = np.abs(gx - ox).max() / np.abs(ox).max()
            ef = np.abs(gf - of).max() / max(np.abs(of).max(), 1e-300)
            et = np.abs(gt - ot).max() / max(np.abs(ot).max(), 1e-300)
            sx_err = np.abs(sx - ox).max() / np.abs(ox).max()
            sf_err = np.abs(sf - of).max() / max(np.abs(of).max(), 1e-300)
            rebuilds = [r[7] for r in allrows]
            same_pairs = pairs == opairs
            print(f"[{name}] N={len(gid)} ranks={world} steps={steps} rebuilds={rebuilds} oracle_rebuilds={o.get_stats().n_rebuilds} "
                  f"pairs={len(pairs)} same_pairs={same_pairs} slab-vs-oracle dx={ex:.2e} dF={ef:.2e} dT={et:.2e} | "
                  f"1gpu-vs-oracle dx={sx_err:.2e} dF={sf_err:.2e}", flush=True)
            if k == 0:  # the bar: short horizon (trajectories are chaotic, BASELINE north_star)
                # a load-balance iteration searches on top of the displacement-triggered ones (dem.cc:383-457)
                same_rebuilds = load_balance is not None or all(r == o.get_stats().n_rebuilds for r in rebuilds)
                ok &= same_pairs and ex < tol[0] and ef < tol[1] and et < tol[1] and same_rebuilds
        done = steps
    if w.params.sparse_contacts:
        # adaptive sparse contacts across slabs: every rank classifies its own cell layers; merged, the statuses are the
        # single-domain oracle's
        mesh = w.params.mesh
        parts = gather_rows((eng.get_slab()[:2], eng.get_mobility_status()), world)
        if rank == 0:
            merged = np.full(mesh.n[0] * mesh.n[1] * mesh.n[2], -1)
            layer = np.arange(len(merged)) % mesh.n[0]  # slab axis 0
            for (lo, hi), st in parts:
                own = (layer >= lo) & (layer < hi)
                merged[own] = st[own]
            so = o.get_mobility_status()
            same = bool(np.array_equal(merged, so))
            print(f"[{name}] mobility status: {np.bincount(so, minlength=5).tolist()} cells per status in the oracle, slab runs agree: {same}", flush=True)
            ok &= same and len(set(np.unique(so).tolist()) & {1, 4}) == 2
    if load_balance and not expect_recut and rank == 0:
        print(f"[{name}] dynamic_with_sparse_contacts: slab of rank 0 {slab0[:2]} -> {eng.get_slab()[:2]}, repartitions {eng.get_slab()[2]}", flush=True)
    if expect_recut:
        # the cuts must have moved towards the particles and evened out the load
        slabs = gather_rows((slab0, eng.get_slab(), eng.n_particles()), world)
        if rank == 0:
            counts = [r[2] for r in slabs]
            moved = any(r[0][:2] != r[1][:2] for r in slabs)
            events = [r[1][2] for r in slabs]
            print(f"[{name}] load balancing: slabs {[r[0][:2] for r in slabs]} -> {[r[1][:2] for r in slabs]}, particles per rank {counts}, "
                  f"repartitions {events}", flush=True)
            ok &= moved and min(events) >= 1 and max(counts) <= 1.5 * sum(counts) / world
            ok &= all(slabs[r][1][1] == slabs[r + 1][1][0] for r in range(world - 1)) and slabs[0][1][0] == 0
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(flag.item())


def two_processor_golden(rank, world, local_rank):
    """tests/dem/particle_particle_contact_on_two_processors.cc on two GPUs: the two spheres live
    on different ranks (cut at y = 0, the reference's own 2-rank partition) and collide across the
    cut. The slab run must equal the single-domain oracle to rounding, and the reference's 2-rank
    golden to 3e-7 m (see test_contact_on_two_processors_golden)."""
    import json

    from lethe_b200 import abi
    from tests.util import GOLDEN, two_processor_contact_case

    p, kw, ids, x, props = two_processor_contact_case()
    lo, hi = multi.slab_bounds(p.mesh.n[1], world)[rank]
    eng = abi.load_engine(p.to_config(slab=(1, lo, hi), **kw), local_rank)
    obj = [abi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    eng.comm_init(rank, world, obj[0])
    mask = multi.owner_mask(x, p.mesh, 1, lo, hi)
    eng.set_particles(ids[mask], x[mask], props[mask])
    with open(os.path.join(GOLDEN, "unit_goldens.json")) as f:
        gold = json.load(f)["contact_on_two_processors_y"]
    o = None
    if rank == 0:
        o = loader.oracle_engine(p.to_config(**kw))
        o.set_particles(ids, x, props)
    done, worst_o, worst_g = 0, 0.0, 0.0
    for k, y in enumerate(gold):
        target = 10 * k + 1
        eng.step(target - done)
        rows = gather_rows(eng.get_particles(), world)
        if rank == 0:
            o.step(target - done)
            got = {int(i): xx for r in rows for i, xx in zip(r[0], r[1])}
            yo = o.get_particles()[1][0, 1]
            worst_o = max(worst_o, abs(got[0][1] - yo))
            worst_g = max(worst_g, abs(got[0][1] - y))
        done = target
    ok = True
    if rank == 0:
        print(f"[two-processor golden] ranks={world} max|y - oracle|={worst_o:.2e} max|y - reference 2-rank golden|={worst_g:.2e}", flush=True)
        ok = worst_o <= 1e-15 and worst_g <= 3e-7
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(flag.item())


def host_streamed_case(rank, world, local_rank):
    """lethe_dem_step_host_state on every rank of a slab-decomposed periodic box, one DEM step per call with the rank's
    host rows in the engine's transfer order: the streamed form (upload stages, partial step launches with the fused halo
    push, download under the upload; a call in which the ranks agree on a new list runs plain) returns bit for bit the
    rows of the plain form, through list rebuilds and migration."""
    w = workloads.periodic_box(cells=(max(16, 6 * world), 10, 10), spacing=1.0, jitter=0.03, vel_sigma=0.5)
    w.params.dynamic_contact_search_factor = 0.4
    w.props[:, 3] += 8.0  # the whole packing drifts along the slab axis: particles change owner at the rebuilds
    finals = []
    for streamed in (False, True):
        os.environ["LETHE_DEM_HOST_PIPELINE"] = "1" if streamed else "0"
        os.environ["LETHE_DEM_HOST_PIPELINE_MIN_ROWS"] = "1"
        os.environ["LETHE_DEM_HOST_STAGES"] = "3"
        os.environ["LETHE_DEM_HOST_SEG_ROWS"] = "256"
        eng, _ = multi.create_slab_engine(w, rank, world, local_rank, dist, balanced=False)
        eng.step(2)

        def rows_now():
            ids, rows = eng.get_state_rows()
            return ids, rows, True

        ids, rows, fresh = rows_now()
        seen = eng.get_stats().n_migrated
        for _ in range(150):
            eng.step_host_state(1, ids if fresh else None, rows)
            fresh = False
            now = eng.get_stats().n_migrated
            if now != seen:  # particles changed owner: the rank re-reads the rows it owns
                ids, rows, fresh = rows_now()
                seen = now
        st = eng.get_stats()
        finals.append((eng.get_particles(), st.n_rebuilds, st.n_migrated, eng.host_pipeline_stats()))
        eng.close()
    for k in ("LETHE_DEM_HOST_PIPELINE", "LETHE_DEM_HOST_PIPELINE_MIN_ROWS", "LETHE_DEM_HOST_STAGES", "LETHE_DEM_HOST_SEG_ROWS"):
        os.environ.pop(k, None)
    (pa, ra, ma, sa), (pb, rb, mb, sb) = finals
    same = all(np.array_equal(u, v) for u, v in zip(pa, pb)) and ra == rb and ma == mb
    ok = same and sa[0] == 0 and sb[0] > 50 and ra >= 3 and ma > 0
    print(f"[host-streamed] rank {rank}: rows identical {same}, rebuilds {ra}/{rb}, migrated {ma}/{mb}, streamed calls {sb[0]} of 150, plans {sb[1]}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ok = True
    # non-periodic axis: a short drum (walls, rotating boundary, gravity, constant rolling).
    # The spheres get random initial angular velocities: constant rolling resistance acts along
    # omega_ij / |omega_ij| (rolling_resistance_torque_models.h:64), which for spheres that all
    # start at rest is the direction of a rounding-level vector — a noise amplifier in the
    # reference too, and not what this check is about.
    w = workloads.drum(n_target=30000, radius=0.03, spacing=1.0, jitter=0.02)
    w.props[:, 6:9] = np.random.default_rng(5).normal(0.0, 5.0, (w.n, 3))
    w.params.dynamic_contact_search_factor = 0.1
    ok &= run_case("drum", w, (10, 40), rank, world, local_rank, tol=(1e-11, 1e-8))
    # periodic axis: particles migrate across slabs and wrap around
    w = workloads.periodic_box(cells=(max(12, 3 * world), 6, 6), spacing=1.0, jitter=0.03, vel_sigma=0.5)
    w.params.dynamic_contact_search_factor = 0.1
    ok &= run_case("periodic", w, (30, 120), rank, world, local_rank, tol=(1e-13, 1e-10))
    # polydisperse hopper: floating wall that opens during the run, outlet deletion at rebuilds
    w = workloads.hopper(n_target=20000, gate_open_time=0.0003, min_nx=max(16, 3 * world))
    w.props[:, 6:9] = np.random.default_rng(6).normal(0.0, 5.0, (w.n, 3))
    w.params.dynamic_contact_search_factor = 0.1
    ok &= run_case("hopper", w, (20, 60), rank, world, local_rank, tol=(1e-11, 1e-8))
    # solid surface: a tilted sheet sweeping through a box packing along the slab axis, so that
    # particles in contact with it (and with each other) change owner while they carry history
    w = workloads.box_packing(n_side=max(16, 4 * world), nz=8, spacing=1.02, jitter=0.05)
    w.props[:, 6:9] = np.random.default_rng(7).normal(0.0, 5.0, (w.n, 3))
    w.props[:, 3] = 2.0  # the whole bed drifts along x: steady migration across the cuts
    w.params.rolling_model = "constant"
    w.params.dynamic_contact_search_factor = 0.1
    hi = w.params.mesh.hi
    v, t = workloads.sheet_mesh(-0.05 * hi[0], 1.05 * hi[0], -0.05 * hi[1], 1.05 * hi[1], lambda x, y: 0.3 * hi[2] + 0.2 * x, 8)
    w.solids = [(v, t, (0.0, 0.0, 10.0), (0.0, 2.0, 0.0), (0.5 * hi[0], 0.5 * hi[1], 0.5 * hi[2]))]
    ok &= run_case("solid", w, (20, 60), rank, world, local_rank, tol=(1e-11, 1e-8))
    # load balancing: a bed heaped against the low-x wall (55 % of the box), equal-width slabs (the upper
    # ranks start empty), `dynamic` method checking every 5 iterations: the cut planes follow the particles while
    # the heap collapses, pairs and forces stay those of the single-domain oracle
    w = workloads.box_packing(n_side=max(24, 8 * world), nz=8, spacing=1.02, jitter=0.05)
    keep = w.x[:, 0] < 0.55 * w.params.mesh.hi[0]
    w.ids, w.x, w.props = w.ids[keep], w.x[keep], w.props[keep]
    w.props[:, 6:9] = np.random.default_rng(9).normal(0.0, 5.0, (w.n, 3))
    w.params.rolling_model = "constant"
    w.params.dynamic_contact_search_factor = 0.1
    ok &= run_case("load-balance", w, (20, 80), rank, world, local_rank, tol=(1e-11, 1e-8), load_balance=("dynamic", 0.2, 5))
    # adaptive sparse contacts: a bed at rest with an agitated top layer; the node-based status crosses the cuts
    w = workloads.box_packing(n_side=max(14, 5 * world), nz=20, spacing=1.0, jitter=0.02)
    rng = np.random.default_rng(3)
    top = w.x[:, 2] > 0.75 * w.x[:, 2].max()
    w.props[:, 3:9] = 0.0
    w.props[top, 3:6] = rng.normal(0.0, 1.0, (int(top.sum()), 3))
    w.params.sparse_contacts = True
    w.params.asc_granular_temperature_threshold = 0.02
    w.params.asc_solid_fraction_threshold = 0.3
    w.params.rolling_model = "constant"
    w.params.dynamic_contact_search_factor = 0.05
    ok &= run_case("sparse-contacts", w, (15, 45), rank, world, local_rank, tol=(1e-11, 1e-8))
    # the same bed with the load balanced by mobility-weighted particle counts
    ok &= run_case("sparse-contacts-lb", w, (15, 45), rank, world, local_rank, tol=(1e-11, 1e-8), load_balance=("dynamic_with_sparse_contacts", 0.05, 5))
    ok &= host_streamed_case(rank, world, local_rank)
    if world == 2:
        ok &= two_processor_golden(rank, world, local_rank)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
