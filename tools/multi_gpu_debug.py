"""Bisect slab-vs-oracle differences on 2 GPUs (debug tool, not a test)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lethe_b200 import multi, workloads  # noqa: E402
from oracle import loader  # noqa: E402


def case(name, w, steps, rank, world, local_rank):
    eng, _ = multi.create_slab_engine(w, rank, world, local_rank, dist, store_forces=True, balanced=False)
    o = None
    if rank == 0:
        o = loader.oracle_engine(w.params.to_config(store_forces=True))
        w.install(o)
    done = 0
    for s in steps:
        eng.step(s - done)
        ids, x, props = eng.get_particles()
        _, f, t = eng.get_forces()
        objs = [None] * world
        dist.all_gather_object(objs, (ids, x, f, t, eng.slab))
        if rank == 0:
            o.step(s - done)
            oid, ox, op = o.get_particles()
            _, of, ot = o.get_forces()
            gid = np.concatenate([r[0] for r in objs])
            order = np.argsort(gid)
            gx = np.concatenate([r[1] for r in objs])[order]
            gf = np.concatenate([r[2] for r in objs])[order]
            gt = np.concatenate([r[3] for r in objs])[order]
            owner = np.concatenate([np.full(len(r[0]), k) for k, r in enumerate(objs)])[order]
            fs = np.abs(of).max()
            err = np.abs(gf - of).max(axis=1) / fs
            terr = np.abs(gt - ot).max(axis=1) / max(np.abs(ot).max(), 1e-300)
            bad = np.argsort(-np.maximum(err, terr))[:6]
            h = w.params.mesh.cell_size[0]
            cut = w.params.mesh.lo[0] + objs[0][4][2] * h
            print(f"[{name}] rebuilds={eng.get_stats().n_rebuilds}/{o.get_stats().n_rebuilds} n={len(gid)}/{len(oid)} steps={s} dx={np.abs(gx - ox).max() / np.abs(ox).max():.2e} dF={err.max():.2e} dT={terr.max():.2e} "
                  f"n_bad(F>1e-9)={(err > 1e-9).sum()} n_bad(T>1e-9)={(terr > 1e-9).sum()} cut_x={cut:.5f} h={h:.5f}", flush=True)
            for b in bad:
                r = np.hypot(ox[b, 1], ox[b, 2])
                print(f"    id={oid[b]} owner={owner[b]} x-cut={(ox[b, 0] - cut) / h:+.2f}h r={r:.5f} Fo={of[b]} Fg={gf[b]} To={ot[b]} Tg={gt[b]}", flush=True)
        done = s


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def drum(**kw):
        w = workloads.drum(n_target=30000, radius=0.03, spacing=1.0, jitter=0.02)
        w.params.dynamic_contact_search_factor = 0.1
        for k, v in kw.items():
            setattr(w.params, k, v)
        return w

    w = drum()
    w.props[:, 6:9] = np.random.default_rng(5).normal(0.0, 5.0, (w.n, 3))
    case("drum-omega", w, (5, 6, 7, 8, 9, 10), rank, world, local_rank)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
