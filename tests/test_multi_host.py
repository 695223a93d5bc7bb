"""Host-side logic of the slab decomposition (no GPU): partition properties and a
world_size-2 gloo run of the ownership / bootstrap plumbing."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from lethe_b200 import multi, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_cover_axis():
    for n, w in [(16, 2), (17, 4), (650, 8), (16, 8)]:
        b = multi.slab_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[r][1] == b[r + 1][0] for r in range(w - 1))
        assert all(hi - lo >= 2 for lo, hi in b)
    with pytest.raises(Exception):
        multi.slab_bounds(7, 4)


def test_balanced_bounds_and_ownership_partition():
    w = workloads.drum(n_target=20000, radius=0.03)
    mesh = w.params.mesh
    ca = np.floor((w.x[:, 0] - mesh.lo[0]) / mesh.cell_size[0]).astype(np.int64)
    for world in (2, 3, 8):
        b = multi.balanced_slab_bounds(ca, mesh.n[0], world)
        masks = [multi.owner_mask(w.x, mesh, 0, lo, hi) for lo, hi in b]
        total = np.sum(masks, axis=0)
        assert np.all(total == 1)  # every particle has exactly one owner
        counts = [int(m.sum()) for m in masks]
        assert max(counts) < 1.3 * w.n / world


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_gloo_world2_ownership_and_id_broadcast(tmp_path):
    script = tmp_path / "w2.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        from lethe_b200 import multi, workloads
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        w = workloads.periodic_box(cells=(8, 4, 4))
        mesh = w.params.mesh
        lo, hi = multi.slab_bounds(mesh.n[0], world)[rank]
        mask = multi.owner_mask(w.x, mesh, 0, lo, hi)
        obj = [os.urandom(128) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)   # the NCCL unique id travels the same way
        counts = [None] * world
        dist.all_gather_object(counts, (int(mask.sum()), obj[0][:8].hex(), lo, hi))
        if rank == 0:
            assert sum(c[0] for c in counts) == w.n, counts
            assert counts[0][1] == counts[1][1]
            assert counts[0][3] == counts[1][2]
            print("GLOO_OK")
        dist.destroy_process_group()
    """))
    port = _free_port()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert "GLOO_OK" in out.stdout, out.stdout + out.stderr


def test_periodic_box_generated_per_slab_equals_global_generation():
    """bench.py's config-5 workload is built one slab at a time (a 64 M-sphere job never exists on
    one host): the union of the per-rank pieces must be exactly the single-domain workload and
    every piece must lie inside its rank's slab."""
    kw = dict(cells=(12, 6, 6), spacing=1.0, jitter=0.03, vel_sigma=0.5)
    whole = workloads.periodic_box(**kw)
    for world in (2, 3, 4):
        parts = [workloads.periodic_box(slab=(r, world), **kw) for r in range(world)]
        assert all(p.n_global == whole.n for p in parts)
        ids = np.concatenate([p.ids for p in parts])
        order = np.argsort(ids)
        assert np.array_equal(ids[order], np.sort(whole.ids))
        ref = np.argsort(whole.ids)
        assert np.array_equal(np.concatenate([p.x for p in parts])[order], whole.x[ref])
        assert np.array_equal(np.concatenate([p.props for p in parts])[order], whole.props[ref])
        bounds = multi.slab_bounds(whole.params.mesh.n[0], world)
        for r, p in enumerate(parts):
            assert multi.owner_mask(p.x, p.params.mesh, 0, *bounds[r]).all()


def test_gloo_world2_slab_generation_and_forced_search_agreement(tmp_path):
    """world_size-2 gloo run of the host logic around the per-step agreement: each rank builds its
    slab, and the logical-or of a rank-local trigger (e.g. an insertion on one rank) reaches both."""
    script = tmp_path / "w2b.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        from lethe_b200 import multi, workloads
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        w = workloads.periodic_box(cells=(8, 4, 4), slab=(rank, world))
        n = torch.tensor([w.n]); dist.all_reduce(n)
        trigger = torch.tensor([0x80000000 if rank == 1 else 0], dtype=torch.int64)   # host bit of rank 1 only
        dist.all_reduce(trigger, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert int(n) == w.n_global == 8 * 4 * 4 * 4, (int(n), w.n_global)
            assert int(trigger) == 0x80000000
            print("GLOO_OK")
        dist.destroy_process_group()
    """))
    port = _free_port()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert "GLOO_OK" in out.stdout, out.stdout + out.stderr


def test_load_balance_cut_rule():
    """lethe_dem_balanced_cuts (host arithmetic of the load-balance step, exported by the CUDA library;
    no device): the cuts move towards the balanced histogram, never by more than max_shift, every slab
    keeps min_width layers, and repeated events converge to the balanced partition."""
    from lethe_b200 import abi

    rng = np.random.default_rng(4)
    n_layers, world = 64, 4
    # a bed that fills the first third of the axis: equal-width slabs leave ranks 2 and 3 empty
    hist = np.zeros(n_layers, np.uint64)
    hist[:22] = rng.integers(900, 1100, 22)
    cuts = np.array([0, 16, 32, 48, 64], np.int32)
    shifts = []
    for _ in range(12):
        new = abi.balanced_cuts(hist, cuts, n_layers)
        assert new[0] == 0 and new[-1] == n_layers
        assert np.all(np.diff(new) >= 2)
        # every cut stays strictly inside the two slabs it separated: owners only change between adjacent ranks
        assert np.all(new[1:-1] > cuts[:-2]) and np.all(new[1:-1] < cuts[2:])
        shifts.append(int(np.abs(new - cuts).sum()))
        cuts = new
    counts = [int(hist[cuts[r]:cuts[r + 1]].sum()) for r in range(world)]
    assert max(counts) - min(counts) <= 2 * int(hist.max()), (cuts, counts)  # within one layer of particles per cut
    assert shifts[-1] == 0  # a fixed point
    # an already balanced partition stays where it is
    flat = np.full(n_layers, 100, np.uint64)
    assert np.array_equal(abi.balanced_cuts(flat, np.array([0, 16, 32, 48, 64], np.int32), 8), [0, 16, 32, 48, 64])
    # the situation of the 4-GPU check: a heap in the first 6 of 17 layers, equal slabs -> the cuts close in on it
    heap = np.zeros(17, np.uint64)
    heap[:6] = [700, 600, 500, 450, 400, 233]
    cuts = np.array([0, 4, 8, 12, 17], np.int32)
    for _ in range(8):
        cuts = abi.balanced_cuts(heap, cuts, 17)
    # six occupied layers cannot feed four slabs of at least two layers: the best partition leaves the last rank empty
    assert np.array_equal(cuts, [0, 2, 4, 6, 17])
