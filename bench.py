#!/usr/bin/env python
"""bench.py — particle-steps/s of the DEM time-step hot path on N B200s (one rank per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

A bench "step" is `--dem-steps` (default 100) consecutive DEM time steps of the whole system — the
contact-detection check of every step, the contact-list rebuilds the displacement criterion
triggers (several per bench step: they are INSIDE the timed region, with the particle migration
and ghost exchange they imply at N > 1), particle-particle + particle-wall forces, velocity-Verlet.
The reference arm uses the same unit. particle-steps/s counts DEM steps.

Workload (config.workload):
  periodic (default)  BASELINE.json configs[4], the configuration the north-star target is quoted
          on: a FIXED 64 M-sphere 3-periodic box at every N (strong scaling), slab-decomposed along
          x at N > 1. Disordered packing: the committed 64 000-sphere unit cell
          (lethe_b200/data/periodic_cell_64000.npz, grown and relaxed by the DEM engine itself,
          solid fraction 0.64, 2.8 touching pairs and 5.8 list entries per sphere) tiled
          10 x 10 x 10, every sphere with its own Maxwellian velocity (sigma 0.1 m/s); material of
          applications_tests/lethe-particles/multiperiodic_collisions_3d.prm (elastic, g = 0).
  drum    configs[1]: 3D rotating drum, 1 M spheres per GPU (weak scaling), HM limit-overlap +
          constant rolling resistance, faceted rotating cylinder wall, disordered bed.
  hopper | box_packing | cohesive_jkr | cohesive_dmt | periodic_lattice: the other configs' workloads.

value     whole-job particle-steps/s with state resident in HBM (CUDA events on the engine's
          stream, max over ranks), rebuilds included
e2e       the same through lethe_dem_step_host_state on every rank: pinned HOST rows (x, v, omega)
          of the owned particles uploaded, ONE DEM step, rows downloaded, every DEM step (the
          reference-facing per-step plugin call; PCIe-bound). The host keeps its rows in the
          engine's transfer order (lethe_dem_get_state_rows, re-read when particles change owner), so
          the call streams: upload, partial step launches and download overlap
roofline  fused step kernel: algorithmic bytes (160+16+4*C+48*T per particle-step,
          C,T measured in this run) / CUDA-event kernel time, against MEASURED_PEAKS.json hbm_gbs
cpu_baseline  the CPU oracle (port of the reference algorithm) on a bounded sample, 1 core
--impl reference  the oracle on all host cores (independent sub-domains of the same workload, no
          halo cost) — the reference itself (deal.II + MPI) cannot be built in this image.
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
CELL_D = 0.002  # sphere diameter of the periodic workload (m)


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


def periodic_reps(n_particles, n_cell):
    """Tiles per direction of the unit cell for ~n_particles spheres (x gets the remainder)."""
    r = max(1, round((n_particles / n_cell) ** (1.0 / 3.0)))
    rx = max(1, round(n_particles / (n_cell * r * r)))
    return (rx, r, r)


def make_workload(args, rank, world):
    from lethe_b200 import workloads

    if args.workload == "periodic":
        xc, Lc, _ = workloads.load_periodic_cell()
        reps = periodic_reps(args.particles, len(xc))
        return workloads.periodic_packing(xc * CELL_D, Lc * CELL_D, reps, d=CELL_D, vel_sigma=args.vel_sigma,
                                          slab=(rank, world) if world > 1 else None)
    if args.workload == "periodic_lattice":
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
    return workloads.drum(n_target=args.n_per_gpu * world, bed="disordered")


def make_cpu_workload(args, n_target, seed):
    """The same workload at a size the CPU oracle steps in seconds."""
    from lethe_b200 import workloads

    if args.workload == "drum":
        return workloads.drum(n_target=n_target, seed=seed, bed="disordered")
    xc, Lc, _ = workloads.load_periodic_cell()
    reps = periodic_reps(n_target, len(xc))
    if n_target < len(xc):
        # a sub-box of the cell is not periodic: small samples (tests) use a small lattice box instead
        side = max(4, round((n_target / 4.0) ** (1.0 / 3.0)))
        return workloads.periodic_box(cells=(side, side, side), spacing=1.005, jitter=0.002, seed=seed)
    return workloads.periodic_packing(xc * CELL_D, Lc * CELL_D, reps, d=CELL_D, vel_sigma=args.vel_sigma, seed=seed)


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
    from oracle import loader

    loader.build()
    cores = max(1, min(os.cpu_count() or 1, args.cpu_cores or 10**6))
    per_core = args.cpu_particles
    ws = []
    for c in range(cores):
        w = make_cpu_workload(args, per_core, 19 + c)
        e = loader.oracle_engine(w.params.to_config())
        w.install(e)
        ws.append((w, e))
    n_total = sum(w.n for w, _ in ws)

    def run_all(k):
        ts = [threading.Thread(target=e.step, args=(k,)) for _, e in ws]
        [t.start() for t in ts]
        [t.join() for t in ts]

    sub = args.dem_steps
    run_all(args.cpu_settle)
    r0 = sum(e.get_stats().n_rebuilds for _, e in ws)
    for _ in range(args.warmup):
        run_all(sub)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_all(sub)
    dt = time.perf_counter() - t0
    value = n_total * sub * args.steps / dt
    sample = (f"{cores} independent sub-domains x {ws[0][0].n} spheres of the same workload (one oracle thread each), {sub} DEM steps per "
              f"bench step, after {args.cpu_settle} settling steps; {sum(e.get_stats().n_rebuilds for _, e in ws) - r0} list rebuilds in all")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak" if args.workload != "periodic" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": ws[0][0].description + " (CPU oracle = port of the reference algorithm; the deal.II/MPI reference cannot "
                               "be built here)", "particles": n_total, "dem_steps_per_step": sub},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_result(line)


def cpu_baseline_sample(args):
    from oracle import loader

    loader.build()
    w = make_cpu_workload(args, args.cpu_particles, 19)
    e = loader.oracle_engine(w.params.to_config())
    w.install(e)
    e.step(args.cpu_settle)
    n_steps = args.cpu_steps
    t0 = time.perf_counter()
    e.step(n_steps)
    dt = time.perf_counter() - t0
    return {"value": w.n * n_steps / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": f"{w.n} spheres of the same workload (same material / models / dt), {n_steps} steps after {args.cpu_settle} settling steps, single thread, g++ -O2 no FMA"}


def bind_to_gpu_numa_node(local_rank):
    """Host side of the e2e path: run this rank on the CPU cores next to its GPU (NVML's CPU affinity of the
    device) so that the pinned host rows it allocates afterwards are first-touched on that NUMA node and
    the per-step copies do not cross the socket interconnect (8 ranks x 2 directions otherwise share it).
    Returns a short description for the e2e record; never fails the run."""
    try:
        import pynvml

        pynvml.nvmlInit()
        index = local_rank
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            entries = [v.strip() for v in visible.split(",") if v.strip()]
            if local_rank < len(entries) and entries[local_rank].isdigit():
                index = int(entries[local_rank])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"rank bound to the {len(cpus)} cores NVML lists next to GPU {index}"
        return "NVML lists no cores for this GPU: not bound"
    except Exception as exc:  # noqa: BLE001
        return f"not bound ({type(exc).__name__})"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="periodic",
                    choices=["periodic", "drum", "periodic_lattice", "box_packing", "hopper", "cohesive_jkr", "cohesive_dmt"])
    ap.add_argument("--particles", type=int, default=64_000_000, help="periodic: total spheres of the FIXED box (strong scaling)")
    ap.add_argument("--n-per-gpu", type=int, default=1_000_000, help="other workloads: spheres per GPU (weak scaling)")
    ap.add_argument("--dem-steps", type=int, default=100, help="DEM time steps per bench step")
    ap.add_argument("--vel-sigma", type=float, default=0.1)
    ap.add_argument("--settle", type=int, default=-1, help="untimed settling steps before warm-up (-1: per workload)")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="DEM steps of the e2e leg (-1: by size)")
    ap.add_argument("--cpu-particles", type=int, default=64_000)
    ap.add_argument("--cpu-steps", type=int, default=300)
    ap.add_argument("--cpu-settle", type=int, default=-1, help="settling steps of the CPU sample (-1: per workload)")
    ap.add_argument("--cpu-cores", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="f64", choices=["f64", "mixed"],
                    help="f64: the reference's arithmetic (the headline); mixed: FP32 contact model between FP64 geometry and integration")
    args = ap.parse_args()
    if args.settle < 0:
        # enough for the bed to come to rest on its supports (loose lattices fall first); the periodic
        # packing only needs its kinetic energy to spread over the contacts
        args.settle = {"drum": 3000, "hopper": 20000, "box_packing": 20000, "periodic": 200, "periodic_lattice": 500}.get(args.workload, 3000)
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
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    t_setup = time.perf_counter()
    w = make_workload(args, rank, world)
    cfg_params = w.params
    if world > 1:
        from lethe_b200 import multi

        # a workload generated per slab (periodic) is cut into equal-width slabs; otherwise the
        # cut planes balance the particle histogram
        engine, n_local = multi.create_slab_engine(w, rank, world, local_rank, dist, balanced=not hasattr(w, "n_global"), precision=args.precision)
    else:
        engine = abi.load_engine(cfg_params.to_config(precision=args.precision), local_rank)
        w.install(engine)
        n_local = w.n
    n_global = getattr(w, "n_global", w.n)
    # the host copies of the initial state are not needed any more (6 GB at 64 M)
    w.ids = w.x = w.props = None
    S = max(1, args.dem_steps)

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
    for _ in range(max(3, args.warmup)):
        engine.step(S)
    barrier()
    setup_s = time.perf_counter() - t_setup

    # ---- timed region: K bench steps of S DEM steps, device-resident, rebuilds included ----
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
    for _ in range(args.steps):
        engine.step(S)
    engine.event_record(1)
    elapsed = engine.event_elapsed_ms() * 1e-3
    barrier()
    clocks = sampler.stop()
    timers = engine.get_timers(reset=True)  # CUDA events on the engine's stream
    launches = engine.kernel_launches() - launches0
    st1 = engine.get_stats()
    migrated = float(st1.n_migrated - st0.n_migrated)
    if world > 1:
        tt = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed = float(tt.item())
        mm = torch.tensor([migrated], device="cuda", dtype=torch.float64)
        dist.all_reduce(mm, op=dist.ReduceOp.SUM)
        migrated = float(mm.item()) / 2  # every migration is counted by the sender and by the receiver
    n_dem = args.steps * S
    value = n_global * n_dem / elapsed

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
    rebuilds = int(st1.n_rebuilds - st0.n_rebuilds)
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
        "kernel": "k_step (fused pp+pw forces, history, velocity-Verlet)", "kernel_ms": k_ms, "peak_source": peak_src,
        "algorithmic_bytes_per_particle_step": bytes_per_pstep, "algorithmic_bytes_per_launch": bytes_per_pstep * n_local,
        "C_half": c_half, "T_half": t_half,
        "kernel_share_of_step": timers["step_kernel_ms"] / (1e3 * elapsed),
        "rebuild_ms_total": timers["rebuild_ms"], "rebuilds": timers["rebuild_launches"],
        "rebuild_ms_each": timers["rebuild_ms"] / max(1, timers["rebuild_launches"]),
        "rebuild_share_of_step": timers["rebuild_ms"] / (1e3 * elapsed),
    }
    traffic_file = os.path.join(ROOT, "profiles", "k_step_traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                tj = json.load(f).get(args.workload)
            if tj:
                # ncu's figure is per launch on ITS particle count: scaled to this run's rows per launch
                roofline["traffic"] = tj["dram_bytes_per_particle_step"] * n_local
                roofline["traffic_source"] = tj.get("source")
                roofline["traffic_over_algorithmic"] = tj["dram_bytes_per_particle_step"] / bytes_per_pstep
        except Exception:
            pass

    # ---- e2e: per-step plugin call with pinned host rows ----
    # Every rank keeps host rows of the particles it owns — what a step changes: x, v, omega, 72 B
    # per particle — and hands them to lethe_dem_step_host_state every DEM step (rows up, one step,
    # rows down; 2 x 72 B per particle cross PCIe inside the timed region, every step). The row -> id
    # table goes up with the first call and again whenever particles changed owner in a rebuild (the
    # rank then re-reads its owned rows with lethe_dem_get_state_rows, inside the timed region).
    def owned_rows():
        # the host keeps its rows in the order the engine names for the transfer (by cell layer, cell-sorted inside a layer:
        # the kind of order a cell-by-cell walk of ParticleHandler gives), not by particle id
        # page-locked once, with head-room for the rows that migrate in: a re-read fills the same buffers
        n_rows = engine.n_particles()
        if pinned["cap"] < n_rows:
            pinned["cap"] = int(n_rows * 1.03) + 1024
            pinned["ids"] = torch.empty(pinned["cap"], dtype=torch.int32).pin_memory()
            pinned["state"] = torch.empty((pinned["cap"], 9), dtype=torch.float64).pin_memory()
        hid, hstate = pinned["ids"][:n_rows], pinned["state"][:n_rows]
        engine.get_state_rows(hid.numpy().view(np.uint32), hstate.numpy())
        return [hid, hstate, True]

    id_uploads = [0]
    pinned = {"cap": 0, "ids": None, "state": None}

    def host_step(rows, n_steps, migrated_seen):
        hid, hstate, fresh = rows
        engine.step_host_state_ptr(n_steps, len(hid), hid.data_ptr() if fresh else 0, hstate.data_ptr())
        id_uploads[0] += len(hid) if fresh else 0
        rows[2] = False
        if world > 1:
            r = engine.get_stats().n_migrated  # particles changed owner: re-read the owned rows
            if r != migrated_seen:
                return owned_rows(), r
        return rows, migrated_seen

    # as many DEM steps as one bench step (100): the leg then holds the list rebuilds (and, at N > 1, the migrations with
    # the re-read of the owned rows) the per-step call meets in steady state, not a lucky window between two of them
    e2e_steps = args.e2e_steps if args.e2e_steps > 0 else S
    rows = owned_rows()
    seen = engine.get_stats().n_migrated
    for _ in range(3):
        rows, seen = host_step(rows, 1, seen)
    barrier()
    t0 = time.perf_counter()
    n_moved = 0
    id_uploads[0] = 0
    for _ in range(e2e_steps):
        n_moved += len(rows[0])
        rows, seen = host_step(rows, 1, seen)
    barrier()
    dt = time.perf_counter() - t0
    n_id_rows = id_uploads[0]
    streamed = engine.host_pipeline_stats()
    barrier()
    t0 = time.perf_counter()
    rows, seen = host_step(rows, S, seen)
    barrier()
    dtb = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt, dtb], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dtb = float(tt[0].item()), float(tt[1].item())
        nn = torch.tensor([float(n_moved), float(n_id_rows)], device="cuda", dtype=torch.float64)
        dist.all_reduce(nn, op=dist.ReduceOp.SUM)
        n_moved, n_id_rows = int(nn[0].item()), int(nn[1].item())
    per_step = n_moved / max(1, e2e_steps)
    h2d = int(per_step * 72 + 4 * n_id_rows / max(1, e2e_steps))
    e2e = {
        "value": n_moved / dt, "unit": "particle-steps/s",
        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(per_step * 72), "dem_steps": e2e_steps,
        "pcie_gbs_per_direction_if_copy_bound": (h2d + per_step * 72) / world / (dt / max(1, e2e_steps)) / 1e9,
        "host_placement": numa,
        "call": "lethe_dem_step_host_state(n_steps=1) on every rank: upload the x/v/omega rows of the owned particles (72 B each; "
                "the id table only when ownership changed), 1 DEM step, download the rows, every DEM step; page-locked rows in the "
                "engine's transfer order, streamed (upload / partial step launches / download overlapped) when streamed_calls > 0",
        "batched": {"steps_per_call": S, "value": n_global * S / dtb, "unit": "particle-steps/s"},
        # rank 0's calls that ran as upload / partial step launches / download pipelined over the rows
        "streamed_calls": streamed[0], "stream_plans": streamed[1],
    }
    del rows

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(args)

    if rank == 0:
        strong = args.workload == "periodic"
        line = {
            "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == "f64" else "f64 state and accumulation, f32 contact model", "data": "synthetic",
            "config": {
                "workload": w.description, "particles": int(n_global), "particles_per_gpu": int(n_global // world),
                "parallelism": f"slab{world}" if world > 1 else "single",
                "dem_steps_per_step": S, "ms_per_dem_step": 1e3 * elapsed / n_dem,
                "settle_steps": args.settle, "l2": "inputs larger than L2 (state+lists >> 126 MB)",
                "rebuilds_in_timed_region": rebuilds, "dem_steps_per_rebuild": n_dem / max(1, rebuilds),
                "particles_migrated_in_timed_region": int(migrated), "setup_s": setup_s,
            },
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
        }
        emit_result(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
