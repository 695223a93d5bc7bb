#!/usr/bin/env python
"""bench.py — particle-steps/s of the DEM time-step hot path on N B200s (one rank per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one DEM time step (contact-detection check [+ list rebuild when triggered],
particle-particle + particle-wall forces, velocity-Verlet) of the whole system.

Workload (config.workload):
  N = 1   BASELINE.json configs[1]: 3D rotating drum, 1M spheres, Hertz-Mindlin limit-overlap
          + constant rolling resistance, faceted rotating cylinder wall.
  N > 1   the same drum made N times longer (1M spheres per GPU, weak scaling), slab-decomposed
          along the drum axis with a ghost-halo exchange every step and particle migration at
          list rebuilds (NCCL).  `--workload periodic_box --n-per-gpu 8000000` runs config 5.

value     whole-job particle-steps/s with state resident in HBM (CUDA events, max over ranks)
e2e       the same through lethe_dem_step_host_state on every rank: pinned HOST rows (x, v, omega)
          of the owned particles uploaded, one step, rows downloaded, every step (the
          reference-facing per-step plugin call; PCIe-bound at ~55 GB/s per direction)
roofline  fused step kernel: algorithmic bytes (160+16+4*C+48*T per particle-step,
          C,T measured) / CUDA-event kernel time, against MEASURED_PEAKS.json hbm_gbs
cpu_baseline  the CPU oracle (port of the reference algorithm) on a bounded sample, 1 core
--impl reference  the oracle on all host cores (independent sub-domains, no halo cost) — the
          reference itself (deal.II + MPI) cannot be built in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "particle-steps/sec"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                parts = [p.strip() for p in out.split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def algorithmic_bytes(c_half, t_half, epsd=False):
    # SURVEY.md §8d / BASELINE.md §2
    return 160.0 + 16.0 + 4.0 * c_half + 48.0 * t_half * (2.0 if epsd else 1.0)


def make_workload(args, rank, world):
    from lethe_b200 import workloads

    if args.workload == "periodic_box":
        # cubic cells per direction for ~n_per_gpu*world particles: 4 per FCC cell
        per = args.n_per_gpu
        side = max(4, round((per / 4.0) ** (1.0 / 3.0)))
        return workloads.periodic_box(cells=(side * world, side, side), spacing=1.005, jitter=0.002,
                                      slab=(rank, world) if world > 1 else None)
    if args.workload == "hopper":
        return workloads.hopper(n_target=args.n_per_gpu * world, gate_open_time=1e-5 * (args.settle + 100))
    if args.workload in ("cohesive_jkr", "cohesive_dmt"):
        side = max(4, round((args.n_per_gpu / 1.41) ** (1.0 / 3.0)))
        return workloads.cohesive_box(side, model="hertz_JKR" if args.workload == "cohesive_jkr" else "DMT")
    if args.workload == "box_packing":
        side = max(4, round((args.n_per_gpu / 1.41) ** (1.0 / 3.0)))
        return workloads.box_packing(side, spacing=1.005, jitter=0.002)
    return workloads.drum(n_target=args.n_per_gpu * world, spacing=1.005, jitter=0.002)


_RESULT_FD = None


def emit_result(line):
    """The one JSON line of the contract, on the process's original stdout."""
    text = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _RESULT_FD is None:
        os.write(1, text)
    else:
        os.write(_RESULT_FD, text)


def run_reference(args):
    """CPU arm: the oracle on every host core, each core an independent sub-domain of the
    workload (what MPI ranks would own, minus the halo cost)."""
    from lethe_b200 import workloads
    from oracle import loader

    loader.build()
    cores = max(1, min(os.cpu_count() or 1, args.cpu_cores or 10**6))
    per_core = args.cpu_particles
    ws = []
    for c in range(cores):
        w = workloads.drum(n_target=per_core, seed=19 + c, spacing=1.005, jitter=0.002)
        e = loader.oracle_engine(w.params.to_config())
        w.install(e)
        ws.append((w, e))
    n_total = sum(w.n for w, _ in ws)

    def run_all(k):
        ts = [threading.Thread(target=e.step, args=(k,)) for _, e in ws]
        [t.start() for t in ts]
        [t.join() for t in ts]

    sub = args.cpu_substeps
    run_all(args.cpu_settle)
    for _ in range(args.warmup):
        run_all(sub)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_all(sub)
    dt = time.perf_counter() - t0
    value = n_total * sub * args.steps / dt
    sample = f"{cores} independent drum slices x {per_core} spheres (one oracle thread each), {sub} DEM steps per bench step, after {args.cpu_settle} settling steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D rotating drum slices, HM limit-overlap + constant rolling (CPU oracle = port of the reference algorithm; "
                               "the deal.II/MPI reference cannot be built here)", "particles": n_total},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_result(line)


def cpu_baseline_sample(args):
    from lethe_b200 import workloads
    from oracle import loader

    loader.build()
    w = workloads.drum(n_target=args.cpu_particles, spacing=1.005, jitter=0.002)
    e = loader.oracle_engine(w.params.to_config())
    w.install(e)
    e.step(args.cpu_settle)
    n_steps = args.cpu_steps
    t0 = time.perf_counter()
    e.step(n_steps)
    dt = time.perf_counter() - t0
    return {"value": w.n * n_steps / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": f"drum slice of {w.n} spheres (same material / models / dt), {n_steps} steps after {args.cpu_settle} settling steps, single thread, g++ -O2 no FMA"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="drum", choices=["drum", "periodic_box", "box_packing", "hopper", "cohesive_jkr", "cohesive_dmt"])
    ap.add_argument("--n-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--settle", type=int, default=-1, help="untimed settling steps before warm-up (-1: per workload)")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--cpu-particles", type=int, default=40_000)
    ap.add_argument("--cpu-steps", type=int, default=300)
    ap.add_argument("--cpu-settle", type=int, default=-1, help="settling steps of the CPU sample (-1: same as --settle)")
    ap.add_argument("--cpu-substeps", type=int, default=100)
    ap.add_argument("--cpu-cores", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.settle < 0:
        # enough for the bed to come to rest on its supports (loose lattices fall first)
        args.settle = {"drum": 3000, "hopper": 20000, "box_packing": 20000, "periodic_box": 500}.get(args.workload, 3000)
    if args.cpu_settle < 0:
        args.cpu_settle = args.settle

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries ONE JSON line (rank 0). Libraries write there too (NCCL's version banner when the
    # environment sets NCCL_DEBUG, torch.distributed notices): file descriptor 1 is pointed at stderr
    # for the whole run and the line goes to the saved descriptor at the end.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from lethe_b200 import abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the DEM engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w = make_workload(args, rank, world)
    cfg_params = w.params
    if world > 1:
        from lethe_b200 import multi

        # a workload generated per slab (periodic_box) is cut into equal-width slabs; otherwise the
        # cut planes balance the particle histogram
        engine, n_local = multi.create_slab_engine(w, rank, world, local_rank, dist, balanced=not hasattr(w, "n_global"))
    else:
        engine = abi.load_engine(cfg_params.to_config(), local_rank)
        w.install(engine)
        n_local = w.n
    n_global = getattr(w, "n_global", w.n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- settle + warm-up (untimed) ----
    engine.step(args.settle)
    # the contact lists are double-buffered across rebuilds: make sure both generations (and the
    # migration / halo buffers) have been allocated before the clock starts. A forced rebuild is
    # transparent to the physics (history is carried over).
    for _ in range(2):
        engine.force_contact_search()
        engine.step(1)
    engine.step(max(3, args.warmup))
    barrier()

    # ---- timed region: K steps, device-resident ----
    engine.enable_timers(True)
    engine.get_timers(reset=True)
    launches0 = engine.kernel_launches()
    st0 = engine.get_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    # CUDA events on the stream the engine launches on (torch.cuda.Event would only see
    # torch's current stream)
    engine.event_record(0)
    engine.step(args.steps)
    engine.event_record(1)
    elapsed = engine.event_elapsed_ms() * 1e-3
    barrier()
    clocks = sampler.stop()
    timers = engine.get_timers(reset=True)  # CUDA events on the engine's stream
    launches = engine.kernel_launches() - launches0
    st1 = engine.get_stats()
    if world > 1:
        tt = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed = float(tt.item())
    value = n_global * args.steps / elapsed

    # ---- C-bar, T-bar for the roofline (one extra step with the touching counter on) ----
    engine.enable_timers(False, count_touching=True)
    engine.step(1)
    st2 = engine.get_stats()
    engine.enable_timers(False, count_touching=False)
    n_now = max(1, st2.n_particles)
    if world > 1:
        c_half = 0.5 * st2.n_pair_entries / n_now
    else:
        c_half = st2.n_pair_entries / n_now
    t_half = st2.n_pairs_touching / n_now
    bytes_per_pstep = algorithmic_bytes(c_half, t_half, cfg_params.rolling_model == "epsd")
    peak, peak_src = measured_peak()
    k_launch = max(1, timers["step_kernel_launches"])
    k_ms = timers["step_kernel_ms"] / k_launch
    achieved = bytes_per_pstep * n_local / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
        "kernel": "k_step (fused pp+pw forces, history, velocity-Verlet)", "kernel_ms": k_ms, "peak_source": peak_src,
        "algorithmic_bytes_per_particle_step": bytes_per_pstep, "C_half": c_half, "T_half": t_half,
        "kernel_share_of_step": timers["step_kernel_ms"] / (1e3 * elapsed),
        "rebuild_ms_total": timers["rebuild_ms"], "rebuilds": timers["rebuild_launches"],
    }
    traffic_file = os.path.join(ROOT, "profiles", "k_step_traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                tj = json.load(f)
            roofline["traffic"] = tj.get("dram_bytes_per_launch")
            roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass

    # ---- e2e: per-step plugin call with pinned host rows ----
    # Every rank keeps host rows of the particles it owns — what a step changes: x, v, omega, 72 B
    # per particle — and hands them to lethe_dem_step_host_state every step (rows up, one step, rows
    # down). The row -> id table goes up with the first call and again whenever particles changed
    # owner in a rebuild (the rank then re-reads its owned rows, inside the timed region).
    def owned_rows():
        ids_, x_, props_ = engine.get_particles()
        state = np.ascontiguousarray(np.concatenate([x_, props_[:, 3:9]], axis=1))
        return [torch.from_numpy(ids_.copy()).pin_memory(), torch.from_numpy(state).pin_memory(), True]

    id_uploads = [0]

    def host_step(rows, n_steps, rebuilds_seen):
        hid, hstate, fresh = rows
        engine.step_host_state_ptr(n_steps, len(hid), hid.data_ptr() if fresh else 0, hstate.data_ptr())
        id_uploads[0] += len(hid) if fresh else 0
        rows[2] = False
        if world > 1:
            r = engine.get_stats().n_migrated  # particles changed owner: re-read the owned rows
            if r != rebuilds_seen:
                return owned_rows(), r
        return rows, rebuilds_seen

    rows = owned_rows()
    seen = engine.get_stats().n_migrated
    for _ in range(3):
        rows, seen = host_step(rows, 1, seen)
    barrier()
    t0 = time.perf_counter()
    n_moved = 0
    id_uploads[0] = 0
    for _ in range(args.e2e_steps):
        n_moved += len(rows[0])
        rows, seen = host_step(rows, 1, seen)
    barrier()
    dt = time.perf_counter() - t0
    n_id_rows = id_uploads[0]
    sub = 100
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        rows, seen = host_step(rows, sub, seen)
    barrier()
    dtb = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt, dtb], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dtb = float(tt[0].item()), float(tt[1].item())
        nn = torch.tensor([float(n_moved), float(n_id_rows)], device="cuda", dtype=torch.float64)
        dist.all_reduce(nn, op=dist.ReduceOp.SUM)
        n_moved, n_id_rows = int(nn[0].item()), int(nn[1].item())
    per_step = n_moved / max(1, args.e2e_steps)
    e2e = {
        "value": n_moved / dt, "unit": "particle-steps/s",
        "h2d_bytes_per_step": int(per_step * 72 + 4 * n_id_rows / max(1, args.e2e_steps)), "d2h_bytes_per_step": int(per_step * 72),
        "call": "lethe_dem_step_host_state(n_steps=1) on every rank: upload the x/v/omega rows of the owned particles (72 B each; "
                "the id table only when ownership changed), 1 DEM step, download the rows, every step",
        "batched": {"steps_per_call": sub, "value": n_global * sub * 3 / dtb, "unit": "particle-steps/s"},
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(args)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": w.description, "particles": int(n_global), "particles_per_gpu": int(n_global // world),
                "parallelism": f"slab{world}" if world > 1 else "single",
                "settle_steps": args.settle, "l2": "inputs larger than L2 (state+lists >> 126 MB)",
                "rebuilds_in_timed_region": int(st1.n_rebuilds - st0.n_rebuilds),
            },
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
        }
        emit_result(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
