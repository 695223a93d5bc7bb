"""Pins the CPU oracle against the reference's own golden outputs.

Each test replays the driver of a reference unit / application test through the
oracle's C ABI and compares with the numbers the reference's ctest diffs against
(tests/golden/*.json, extracted by tests/golden/make_golden.py)."""
import math
import os

import numpy as np
import pytest

from lethe_b200 import abi
from lethe_b200.prm import Mesh, load_prm
from lethe_b200.solver import DEMSolver, box_wall_faces
from oracle import loader
from tests.util import GOLDEN, assert_sig6, golden, packing_parameters, props_row, random_packing, unit_test_parameters


def test_abi_struct_sizes(oracle_lib):
    # config layout agreed between C and ctypes (catches silent struct drift)
    import ctypes

    assert ctypes.sizeof(abi.WallFace) == 64
    assert ctypes.sizeof(abi.Config) % 8 == 0


def test_pp_force_nonlinear(oracle_lib):
    # tests/dem/particle_particle_contact_force_nonlinear.cc:40-127 -> -0.258955 N
    p = unit_test_parameters()
    e = loader.oracle_engine(p.to_config())
    r = loader.pair_force(e, [0.4, 0, 0], props_row(0, 0.005, 1, (0.01, 0, 0)), [0.40499, 0, 0], props_row(0, 0.005, 1))
    g = golden()["pp_force_nonlinear"]
    for d in range(3):
        assert_sig6(r["force_one"][d], g[d], "pp nonlinear")
    assert abs(r["overlap"] - 1.0000000000005664e-05) < 1e-19
    assert abs(r["force_one"][0] - (-0.2589545634)) < 1e-9  # SURVEY KAT appendix


def test_pp_force_linear(oracle_lib):
    # tests/dem/particle_particle_contact_force_linear.cc -> -1.28940 N
    p = unit_test_parameters(pp_model="linear")
    e = loader.oracle_engine(p.to_config())
    r = loader.pair_force(e, [0.4, 0, 0], props_row(0, 0.005, 1, (0.01, 0, 0)), [0.40499, 0, 0], props_row(0, 0.005, 1))
    g = golden()["pp_force_linear"]
    for d in range(3):
        assert_sig6(r["force_one"][d], g[d], "pp linear")


@pytest.mark.parametrize("model,key", [("nonlinear", "pw_force_nonlinear"), ("linear", "pw_force_linear")])
def test_pw_force(oracle_lib, model, key):
    # tests/dem/particle_wall_contact_force_{nonlinear,linear}.cc:64-110 -> 19.5014 / 41.5643 N
    p = unit_test_parameters(pw_model=model, g=(0, 0, -9.81))
    cfg = p.to_config(store_forces=True, moi_override=1.0)
    e = loader.oracle_engine(cfg)
    e.set_walls(box_wall_faces(p.mesh))
    e.set_particles([0], [[-0.998, 0, 0]], [props_row(0, 0.005, 1, (0.01, 0, 0))])
    e.step(1)
    _, f, _ = e.get_forces()
    assert_sig6(f[0, 0], golden()[key], key)


@pytest.mark.parametrize("model", ["linear", "hertz_mindlin_limit_overlap"])
def test_full_contact_series(oracle_lib, model):
    # tests/dem/particle_particle_full_contact.cc + full_contact_functions.h:105-330:
    # loop = search -> force -> record -> integrate (plain integrate, MOI = 1), dt = 1e-5.
    p = unit_test_parameters(pp_model=model)
    p.restart = True  # regular integrate() from the first step, as the test driver does
    p.contact_detection_method, p.contact_detection_frequency = "constant", 1
    cfg = p.to_config(store_forces=True, moi_override=1.0)
    e = loader.oracle_engine(cfg)
    e.set_particles([0, 1], [[0.4, 0, 0], [0.405, 0, 0]], [props_row(0, 0.005, 1, (0.01, 0, 0)), props_row(0, 0.005, 1)])
    series = golden()["full_contact"][model]
    by_iter = {int(round(s["time"] / 1e-5)): s for s in series}
    last_iter = max(by_iter)
    checked = 0
    for it in range(last_iter + 1):
        _, x, _ = e.get_particles()
        overlap = 0.005 - math.sqrt(((x[0] - x[1]) ** 2).sum())
        e.step(1)
        if it in by_iter:
            s = by_iter[it]
            _, f, t = e.get_forces()
            for d in range(3):
                assert_sig6(f[0, d], s["force"][d], f"{model} force it={it}")
                assert_sig6(t[0, d], s["torque"][d], f"{model} torque it={it}")
            assert_sig6(overlap, s["overlap"], f"{model} overlap it={it}")
            checked += 1
    assert checked == len(series) >= 90


def test_normal_force_series(oracle_lib):
    # tests/dem/normal_force.cc:56-190: sphere hitting the x=-1 wall at 1 m/s, nonlinear model
    p = unit_test_parameters(pw_model="nonlinear", dt=1e-6, d=0.001, young=2e11, restitution=0.5, friction=0.3, rolling_viscous=0.1)
    p.restart = True
    cfg = p.to_config(store_forces=True, moi_override=1.0)
    e = loader.oracle_engine(cfg)
    e.set_walls(box_wall_faces(p.mesh))
    e.set_particles([0], [[-0.999, 0, 0]], [props_row(0, 0.001, 1, (-1.0, 0, 0))])
    gold = golden()["normal_force"]
    out = []
    time = 0.0
    while time < 0.00115:
        _, x, _ = e.get_particles()
        distance = 1 + x[0, 0] - 0.001 / 2.0
        e.step(1)
        if not distance > 0.0:
            _, f, _ = e.get_forces()
            out.append(f[0, 0])
        time += 1e-6
    assert len(out) == len(gold)
    for k, (a, b) in enumerate(zip(out, gold)):
        assert_sig6(a, b, f"normal_force[{k}]")


@pytest.mark.parametrize("case", [0, 1])
def test_post_collision_velocity(oracle_lib, case):
    # tests/dem/post_collision_velocity.cc: e = 0.9 -> 0.0900043, e = 1 -> 0.1
    g = golden()["post_collision_velocity"][case]
    d = 0.002
    p = unit_test_parameters(pw_model="nonlinear", dt=1e-8, d=d, young=8e8, restitution=g["restitution"], friction=0.3, rolling_viscous=0.1)
    p.restart = True
    e = loader.oracle_engine(p.to_config(moi_override=1.0))
    e.set_walls(box_wall_faces(p.mesh))
    mass = math.pi * d * d * d / 6
    e.set_particles([0], [[-0.999, 0, 0]], [props_row(0, d, mass, (-0.1, 0, 0))])
    e.step(10000)
    _, _, props = e.get_particles()
    assert_sig6(props[0, 3], g["v_after"], "post collision velocity")


def test_velocity_verlet_free_flight(oracle_lib):
    # tests/dem/integration_velocity_verlet.cc: start + 4 x integrate + end, F=(1,0,0), T=(0,0,2), m=I=1
    p = unit_test_parameters(dt=1e-3, g=(0, 0, -9.81))
    e = loader.oracle_engine(p.to_config())
    e.set_particles([0], [[0, 0, 0]], [props_row(1, 0.005, 1.0)])
    F, T = [1.0, 0, 0], [0, 0, 2.0]
    loader.integrate_external(e, 0, F, T, 1.0)
    for _ in range(4):
        loader.integrate_external(e, 1, F, T, 1.0)
    loader.integrate_external(e, 2, F, T, 1.0)
    _, x, props = e.get_particles()
    g = golden()["velocity_verlet_3d"]
    for d in range(3):
        assert_sig6(x[0, d], g["position"][d], "position")
        assert_sig6(props[0, 3 + d], g["velocity"][d], "velocity")
        assert_sig6(props[0, 6 + d], g["omega"][d], "omega")


def test_explicit_euler_free_fall(oracle_lib):
    # tests/dem/integration_euler.cc: one particle at rest, g = -9.81 z, dt = 1e-5, one `integrate`;
    # golden integration_euler.output: z = -9.81000e-10
    p = unit_test_parameters(dt=1e-5, g=(0, 0, -9.81))
    p.integration_method = "explicit_euler"
    e = loader.oracle_engine(p.to_config())
    e.set_particles([0], [[0, 0, 0]], [props_row(1, 0.005, 1.0)])
    loader.integrate_external(e, 1, [0.0, 0, 0], [0.0, 0, 0], 1.0)
    _, x, props = e.get_particles()
    assert_sig6(x[0, 2], -9.81000e-10, "z after one explicit Euler step")
    assert x[0, 0] == 0 and x[0, 1] == 0
    # the closing step leaves the state untouched (explicit_euler_integrator.cc:32-48)
    loader.integrate_external(e, 2, [1.0, 0, 0], [0.0, 0, 0], 1.0)
    _, x2, props2 = e.get_particles()
    assert np.array_equal(x, x2) and np.array_equal(props, props2)


def test_integration_schemes_order(oracle_lib):
    # tests/dem/integration_schemes_accuracy.cc: a particle on a linear spring F = -k x; halving dt
    # divides the end-position error by 2^order; golden: Euler 1.00529, Verlet 2.00189
    # (the exact decimals depend on the test's spring set-up; the orders are what is pinned here)
    def run(method, dt, t_end=0.05, k=50.0):
        p = unit_test_parameters(dt=dt, g=(0, 0, 0))
        p.integration_method = method
        e = loader.oracle_engine(p.to_config())
        e.set_particles([0], [[0.3, 0, 0]], [props_row(1, 0.005, 1.0)])
        n = int(round(t_end / dt))
        for it in range(n):
            _, x, _ = e.get_particles()
            f = [-k * x[0, 0], 0.0, 0.0]
            loader.integrate_external(e, 0 if (it == 0) else 1, f, [0.0, 0, 0], 1.0)
        _, x, _ = e.get_particles()
        return x[0, 0]

    import math as m
    exact = 0.3 * m.cos(m.sqrt(50.0) * 0.05)
    for method, order in (("explicit_euler", 1.0), ("velocity_verlet", 2.0)):
        e1 = abs(run(method, 1e-4) - exact)
        e2 = abs(run(method, 5e-5) - exact)
        assert abs(m.log2(e1 / e2) - order) < 0.1, (method, e1, e2)


def test_find_contact_pairs_and_fine_search(oracle_lib):
    # tests/dem/find_contact_pairs.cc: three particles, candidate pairs (0,1) and (1,2);
    # tests/dem/particle_particle_fine_search.cc: pair enters with zero tangential displacement
    p = unit_test_parameters()
    e = loader.oracle_engine(p.to_config())
    # 0 and 1 in neighbouring cells, 2 next to 1, 0 and 2 two cells apart
    e.set_particles([0, 1, 2], [[-0.4, 0, 0], [0.4, 0, 0], [0.8, 0, 0]], [props_row(0, 0.005, 1)] * 3)
    e.step(1)
    i, j, t = e.get_pairs()
    assert len(i) == 0  # far apart: candidates but not within the neighbourhood threshold
    e2 = loader.oracle_engine(p.to_config())
    e2.set_particles([0, 1], [[0.4, 0, 0], [0.40499, 0, 0]], [props_row(0, 0.005, 1), props_row(0, 0.005, 1)])
    e2.force_contact_search()
    # fine search only (no motion): use a zero-velocity step
    e2.step(1)
    i, j, t = e2.get_pairs()
    assert list(zip(i, j)) == [(0, 1)]


def test_combined_periodic_offsets(oracle_lib):
    # tests/dem/combined_periodic_offsets.cc: 26 translation vectors for 3 periodic directions;
    # checked through behaviour: a pair across each periodic face/edge/corner is found.
    gold = golden()["combined_periodic_offsets_3d"]
    assert len(gold) == 26 and sorted(map(tuple, gold)) == sorted(
        (a, b, c) for a in (-1.0, 0.0, 1.0) for b in (-1.0, 0.0, 1.0) for c in (-1.0, 0.0, 1.0) if (a, b, c) != (0.0, 0.0, 0.0)
    )
    from tests.util import packing_parameters

    d = 0.005
    p = packing_parameters((0.04, 0.04, 0.04), d=d, cell=0.01, periodic=(1, 1, 1), g=(0, 0, 0))
    L = p.mesh.hi[0]
    for shift in gold:
        e = loader.oracle_engine(p.to_config())
        a = np.array([L / 2, L / 2, L / 2])
        for ax in range(3):
            if shift[ax] != 0:
                a[ax] = 0.2 * d
        b = a.copy()
        for ax in range(3):
            if shift[ax] != 0:
                b[ax] = L - 0.3 * d
        e.set_particles([0, 1], [a, b], [props_row(0, d, 1e-4), props_row(0, d, 1e-4)])
        e.step(1)
        i, j, _ = e.get_pairs()
        assert list(zip(i, j)) == [(0, 1)], shift


def test_packing_in_box_application(oracle_lib):
    # applications_tests/lethe-particles/packing_in_box.{prm,mpirun=1.output}: 200 spheres,
    # 1e4 steps, final positions printed with 4 decimals.
    import os

    p = load_prm(os.path.join(GOLDEN, "packing_in_box.prm"))
    s = DEMSolver(p, engine_factory=loader.oracle_engine)
    ids, x, props = s.solve()
    rows = golden("packing_in_box.mpirun1.json")["rows"]
    assert len(rows) == len(ids) == 200
    gold = np.array([r[3:6] for r in rows])
    assert [r[0] for r in rows] == list(ids)
    err = np.abs(x - gold).max(axis=1)
    # 4 printed decimals. The bed is still settling at t_end (speeds up to 0.3 m/s); perturbing the
    # inserted positions by a few ulp moves 11-20 rows past the printed digit (symmetric lattice).
    # The oracle has 195 rows on the digit and 5 off by at most 2e-5 (open item).
    assert np.mean(err <= 0.5e-4 + 1e-9) >= 0.97, (np.mean(err <= 0.5e-4 + 1e-9), err.max())
    assert err.max() < 1e-4, err.max()


@pytest.mark.parametrize("case", ["edge_vertex_contact", "CPES_double_edge_contact", "NPES_double_edge_contact",
                                  "NPES_double_face_contact"])
def test_particle_solid_surface_application_goldens(oracle_lib, case):
    """applications_tests/lethe-particles/particle_solid_surface_*.prm: one sphere dropped on two
    triangles; the logged velocity magnitude pins find_particle_triangle_projection, the
    face/edge/vertex double-contact elimination and the solid-object contact force
    (particle_wall_contact_force.cc:153-580). The statistics of iteration n are printed before the
    step body, i.e. they show the state after n - 1 steps (dem.cc:1115-1120)."""
    from lethe_b200.solver import box_wall_faces
    from tests.util import solid_surface_case

    c, params, x, props, vertices, triangles = solid_surface_case(case)
    e = loader.oracle_engine(params.to_config())
    e.set_walls(box_wall_faces(params.mesh))
    e.add_solid_surface(vertices, triangles)
    e.set_particles([0], x, props)
    done = 0
    for k, gold in enumerate(c["velocity_magnitude"], start=1):
        target = k * c["log_frequency"] - 1
        e.step(target - done)
        done = target
        _, _, p = e.get_particles()
        v = float(np.sqrt((p[0, 3:6] ** 2).sum()))
        assert abs(v - gold) <= 5.1e-5 * max(abs(gold), 1e-30) + 1e-12, (case, k, v, gold)  # 5 printed digits


def test_solid_surface_prm_end_to_end(oracle_lib):
    """The reference's own particle_solid_surface_NPES_double_edge_contact.prm (+ its gmsh file),
    read by the .prm mirror — `subsection solid objects`, `insertion method = list` — and run
    through DEMSolver: velocity magnitude after the first bounce as logged at iteration 380000
    (the closing half kick of solve() changes it by g*dt/2, far below the printed digits)."""
    import json

    from tests.util import GOLDEN as GDIR

    d = os.path.join(GDIR, "solid_surfaces")
    params = load_prm(os.path.join(d, "particle_solid_surface_NPES_double_edge_contact.prm"))
    assert params.insertion.method == "list" and len(params.solid_surfaces) == 1
    solver = DEMSolver(params, engine_factory=loader.oracle_engine, prm_directory=d)
    _, _, p = solver.solve(max_steps=379999)
    with open(os.path.join(GDIR, "solid_surface_goldens.json")) as f:
        gold = json.load(f)["NPES_double_edge_contact"]["velocity_magnitude"][37]
    v = float(np.sqrt((p[0, 3:6] ** 2).sum()))
    assert abs(v - gold) <= 1e-4 * gold, (v, gold)


APP_CASES = ["rolling_on_plane", "velocity_verlet_free_fall", "multiperiodic_collisions_3d", "multiperiodic_edge_contact_3d",
             "pp_jkr_equilibrium", "pp_dmt_equilibrium", "pw_jkr_equilibrium", "pw_dmt_equilibrium", "epsd_rolling_resistance_model",
             "sliding_in_box", "periodic_boundary_box", "moving_solid_surface_hmlo", "moving_solid_surface_jkr",
             "moving_solid_surface_dmt", "insert_file_3d", "insert_list_3d", "insert_z-x-y",
             "multiperiodic_single_axis_collisions_3d", "single-time-step-list-insertion", "periodic_boundary_collisions",
             "distribution_normal", "distribution_lognormal", "solid_surface", "deprecated_parameters",
             "insert_list_3d_default_velocities", "insertion_acceptance_function", "insert_plane_3d",
             "initial_value_insertion", "periodic_boundary_load_balancing",
             "insert_and_remove_with_files"]
# goldens the reference produced on 2 MPI ranks with a volume insertion: each rank pairs its share of
# the lattice with its own random vector, which the mirror restates on request
REFERENCE_INSERTION_RANKS = {"periodic_boundary_load_balancing": 2}


def run_application_case(case, engine_factory):
    d = os.path.join(GOLDEN, "apps")
    params = load_prm(os.path.join(d, case + ".prm"))
    solver = DEMSolver(params, engine_factory=engine_factory, prm_directory=d,
                       reference_insertion_ranks=REFERENCE_INSERTION_RANKS.get(case, 1))
    ids, x, props = solver.solve()
    import json

    with open(os.path.join(d, "final_positions.json")) as f:
        rows = json.load(f)[case]
    assert list(ids) == [r[0] for r in rows]
    gold = np.array([r[3:6] for r in rows])
    # the goldens print 4 decimals
    assert np.abs(x - gold).max() <= 0.5e-4 + 1e-9, (case, np.abs(x - gold).max())
    assert np.abs(props[:, 1] - np.array([r[2] for r in rows])).max() <= 0.5e-5 + 1e-12
    return solver


@pytest.mark.parametrize("case", APP_CASES)
def test_application_goldens(oracle_lib, case):
    """The reference's own lethe-particles application tests that run on box meshes, from their
    unmodified .prm files through the .prm mirror + DEMSolver + oracle: final positions to the 4
    printed decimals. Covers constant / viscous / EPSD rolling resistance, hertz_mindlin_limit_force,
    JKR and DMT (particle-particle and particle-wall), periodic boundaries in 1 and 3 directions, list
    and volume insertion, and moving solid surfaces with all three wall models."""
    run_application_case(case, loader.oracle_engine)


def test_epsd_application_log_statistics(oracle_lib):
    """epsd_rolling_resistance_model.output also logs statistics every 5000 iterations: the number
    of contact searches and the angular-velocity statistics through the multi-collision phase are
    reproduced to the 5 printed digits (this is what pins the contact-search trigger at insertion
    iterations and the lower-cell rule for points on cell faces)."""
    from lethe_b200.solver import list_insertion

    d = os.path.join(GOLDEN, "apps")
    p = load_prm(os.path.join(d, "epsd_rolling_resistance_model.prm"))
    e = loader.oracle_engine(p.to_config())
    e.set_walls(box_wall_faces(p.mesh, p.outlet_boundaries, p.periodic))
    e.add_particles(*list_insertion(p))
    gold_searches = [3, 10, 22, 34, 42, 48, 54]
    gold_omega = [(0.0, 6.2832, 4.7124), (0.0, 6.2832, 4.7124), (6.2832, 264.41, 88.006), (2.4912, 338.42, 159.38),
                  (6.2832, 303.86, 123.10), (0.12798, 28.603, 13.803), (2.9838e-05, 6.2832, 2.2325)]
    it = 0
    for k in range(7):
        target = (k + 1) * 5000 - 1
        while it < target:
            nxt = min(target, ((it // 10000) + 1) * 10000)  # stop before every insertion iteration (10001, 20001, ...)
            e.step(nxt - it)
            it = nxt
            if it % 10000 == 0 and it < target:
                e.force_contact_search()
        _, _, props = e.get_particles()
        w = np.sqrt((props[:, 6:9] ** 2).sum(axis=1))
        assert e.get_stats().n_rebuilds == gold_searches[k], (k, e.get_stats().n_rebuilds)
        for got, gold in zip((w.min(), w.max(), w.mean()), gold_omega[k]):
            assert abs(got - gold) <= 2e-4 * abs(gold) + 1e-12, (k, got, gold)
        if (it + 1) % 10000 == 0:
            # iteration it + 2 == 10001 (mod 10000) is an insertion iteration
            e.step(1)
            it += 1
            e.force_contact_search()


def test_contact_on_two_processors_golden(oracle_lib):
    """tests/dem/particle_particle_contact_on_two_processors.mpirun=2.output (a 2-rank, 2-D golden):
    y of particle 0 every 10th step. Reproduced in 3-D on a single domain to 3e-7 m absolute
    (5 printed digits): exact to all 6 digits until the contact starts, then a slowly growing
    offset that ends at 2.1e-7 with identical rebound velocity. The reference's 1-rank series
    (particle_particle_full_contact.output) is matched to every digit, so the offset belongs to
    the reference's 2-rank/2-D run, not to the contact model. The same series is checked across
    two GPUs by tests/multi_gpu_check.py."""
    from tests.util import two_processor_contact_case

    p, kw, ids, x, props = two_processor_contact_case()
    e = loader.oracle_engine(p.to_config(**kw))
    e.set_particles(ids, x, props)
    gold = golden()["contact_on_two_processors_y"]
    done = 0
    for k, y in enumerate(gold):
        target = 10 * k + 1
        e.step(target - done)
        done = target
        _, xx, _ = e.get_particles()
        assert abs(xx[0, 1] - y) <= (5.1e-9 if k <= 10 else 3e-7), (k, xx[0, 1], y)
    _, xx, pp = e.get_particles()
    assert abs((gold[-1] - gold[-2]) / 10 - pp[0, 4] * 1e-5) <= 1e-9  # same rebound velocity


def test_time_dependent_solid_velocity_function(oracle_lib):
    """`subsection translational velocity / Function expression` with a muparser conditional (as in
    load_balancing_solid_object.prm): the mirror evaluates it at the previous time of every step
    (serial_solid.cc:343-352) and pushes changes through lethe_dem_set_solid_motion."""
    from lethe_b200.prm import SolidSurface, evaluate_function

    assert evaluate_function("if(t>0.5,if(t<0.7,1,0),0)", 0.6) == 1.0
    assert evaluate_function("if(t>0.5,if(t<0.7,1,0),0)", 0.8) == 0.0
    assert evaluate_function("2*t^2", 3.0) == 18.0
    d = os.path.join(GOLDEN, "solid_surfaces")
    params = load_prm(os.path.join(d, "particle_solid_surface_NPES_double_edge_contact.prm"))
    params.time_step, params.time_end = 1e-5, 1e-3
    params.solid_surfaces[0].translational_velocity = (0.0, 0.0, "if(t>0.0005,2,0)")
    solver = DEMSolver(params, engine_factory=loader.oracle_engine, prm_directory=d)
    v0 = None
    solver.solve(max_steps=0)
    v0 = solver.engine.get_solid_vertices(0).copy()
    solver.solve()
    v1 = solver.engine.get_solid_vertices(0)
    # steps n = 1..100 use t_prev = (n - 1) dt; t_prev > 0.0005 for n = 52..100 (binary rounding of
    # 51 * 1e-5 decides the edge): 49 or 50 moving steps of 2e-5 each
    dz = v1[:, 2] - v0[:, 2]
    assert np.allclose(dz, dz[0]) and np.allclose(v1[:, :2], v0[:, :2])
    assert round(dz[0] / 2e-5) in (49, 50), dz[0]


def test_cell_neighbor_lists_match_reference_goldens(oracle_lib):
    """tests/dem/find_cell_neighbors.output (reciprocal = 0) and find_full_cell_neighbors.output on
    hyper_cube(-1, 1) refined twice: the oracle's active-cell numbering and the ORDER of every
    neighbour list (which fixes the candidate and force-summation order of the search) equal the
    reference's for all 64 cells."""
    e = loader.oracle_engine(unit_test_parameters().to_config())
    assert loader.cell_neighbors(e, 0) == golden()["find_cell_neighbors"]
    assert loader.cell_neighbors(e, 1) == golden()["find_full_cell_neighbors"]


def test_particle_wall_contact_pairs_and_fine_search_goldens(oracle_lib):
    """tests/dem/particle_wall_contact_pairs.cc (golden: only particle 2, at x = 0.8, is in a boundary
    cell) and particle_wall_fine_search.cc (golden: particle 0 at x = -0.998 holds one wall contact whose
    stored normal — from the wall to the particle — is 1 0 0). deal.II's face number (77) is not
    reproducible without its face enumeration and is not compared."""
    p = unit_test_parameters()
    faces = box_wall_faces(p.mesh)
    e = loader.oracle_engine(p.to_config())
    e.set_walls(faces)
    e.set_particles([0, 1, 2], [[-0.4, 0, 0], [0.4, 0, 0], [0.8, 0, 0]], [props_row(0, 0.2, 1.0)] * 3)
    e.step(1)
    particle, face, _ = e.get_wall_contacts()
    assert list(particle) == [2]

    e = loader.oracle_engine(p.to_config())
    e.set_walls(faces)
    e.set_particles([0], [[-0.998, 0, 0]], [props_row(0, 0.005, 1.0)])
    e.step(1)
    particle, face, _ = e.get_wall_contacts()
    assert list(particle) == [0] and len(face) == 1
    hit = [f for f in faces if f.global_face_id == face[0]]
    assert len(hit) == 1 and tuple(hit[0].normal) == (1.0, 0.0, 0.0)


def test_two_particles_multiple_contacts_golden(oracle_lib):
    """tests/dem/two_particles_multiple_contacts_parallel.cc (its only golden is the 2-rank run, both
    spheres on one rank): a sphere falls at 0.4 m/s onto one at rest, E = 5e7, nu = 0.9, e = 0.9,
    m = MOI = 1, plain integrate() from the first step; the y force on particle 0 at every 10th of
    1000 steps to the 6 printed digits (100 samples through the whole contact)."""
    p = unit_test_parameters(restitution=0.9)
    p.particle_types[0].poisson = 0.9
    p.restart = True
    p.contact_detection_method, p.contact_detection_frequency = "constant", 1
    e = loader.oracle_engine(p.to_config(store_forces=True, moi_override=1.0))
    e.set_particles([0, 1], [[0, 0.007, 0], [0, 0.001, 0]], [props_row(0, 0.005, 1, (0, -0.4, 0)), props_row(0, 0.005, 1)])
    series = dict(golden()["two_particles_multiple_contacts"])
    assert len(series) == 100 and max(series.values()) > 100
    for it in range(1000):
        e.step(1)
        if it in series:
            _, f, _ = e.get_forces()
            assert_sig6(f[0, 1], series[it], f"it={it}")


@pytest.mark.parametrize("case,offset", [(1, 0.75), (2, 0.0)])
def test_insertion_volume_goldens(case, offset):
    """tests/dem/insertion_volume_{1,2}.cc: 10 spheres of 5 mm in the box (-0.05, 0.05)^3, distance
    threshold 2, maximum offset 0.75 / 0, seed 19: the inserted positions (lattice site order, the
    two glibc rand() offsets per sphere and their pairing r[k], r[n - k - 1]) to 6 digits."""
    from lethe_b200.solver import volume_insertion

    p = unit_test_parameters()
    p.particle_types[0].number = 10
    ins = p.insertion
    ins.box_point_1, ins.box_point_2 = (-0.05, -0.05, -0.05), (0.05, 0.05, 0.05)
    ins.direction_sequence, ins.inserted_this_step = (0, 1, 2), 10
    ins.distance_threshold, ins.maximum_offset, ins.seed = 2.0, offset, 19
    _, x, _ = volume_insertion(p, 10)
    gold = golden()[f"insertion_volume_{case}"]
    assert len(gold) == 10
    for row, g_row in zip(x, gold):
        for a, b in zip(row, g_row):
            assert_sig6(a, b, f"insertion_volume_{case}")


def test_insertion_plane_golden():
    """tests/dem/insertion_plane.cc: hyper_cube(-2, 2) refined twice, plane y = 1.75 with normal y,
    maximum offset 0.2: one sphere in each of the 16 cells the plane cuts, in active-cell order, at
    the cell centre plus three successive glibc rand() offsets — positions to 6 digits."""
    from lethe_b200.solver import PlaneInsertion

    p = unit_test_parameters(d=0.2)
    p.mesh = Mesh((-2.0,) * 3, (2.0,) * 3, (4, 4, 4), True, "morton")
    p.particle_types[0].number = 16
    p.insertion.method = "plane"
    p.insertion.plane_point, p.insertion.plane_normal, p.insertion.maximum_offset = (0.0, 1.75, 0.0), (0.0, 1.0, 0.0), 0.2
    plane = PlaneInsertion(p)
    from lethe_b200.distributions import make_distribution

    ids, x, props = plane.insert(p, set(), 16, 0, 0, make_distribution(p.particle_types[0]))
    gold = golden()["insertion_plane"]
    assert len(gold) == len(ids) == 16
    for row, g_row in zip(x, gold):
        for a, b in zip(row, g_row):
            assert_sig6(a, b, "insertion_plane")


def test_boundary_cells_and_faces_golden():
    """tests/dem/boundary_cells_and_faces.cc on hyper_cube(-1, 1) refined twice: the 96 (boundary
    cell, boundary face) rows. deal.II's face numbers (42..137) come from its refinement history and
    are not restated; what is compared is which cells carry how many boundary faces, and that the
    golden's 6 runs of 16 rows are exactly the six sides of box_wall_faces (y-, z-, x-, x+, y+, z+
    in deal.II's numbering)."""
    from collections import Counter

    from lethe_b200.solver import active_cell_order

    p = unit_test_parameters()
    rank = {c: r for r, c in enumerate(active_cell_order(p.mesh))}
    nx, ny, _ = p.mesh.n
    mine = {}
    for f in box_wall_faces(p.mesh):
        c = (f.cell % nx, (f.cell // nx) % ny, f.cell // (nx * ny))
        mine.setdefault(tuple(f.normal), []).append(rank[c])
    gold = golden()["boundary_cells_and_faces"]
    assert len(gold) == 96 and sorted(n for _, n in gold) == list(range(42, 138))
    assert Counter(c for c, _ in gold) == Counter(c for cells in mine.values() for c in cells)
    sides = [(0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (1.0, 0.0, 0.0), (-1.0, 0.0, 0.0), (0.0, -1.0, 0.0), (0.0, 0.0, -1.0)]
    for k, normal in enumerate(sides):  # inward normals
        assert sorted(c for c, _ in gold[16 * k:16 * k + 16]) == sorted(mine[normal]), normal


def test_step_host_state_equals_step_host(oracle_lib):
    """The oracle's side of lethe_dem_step_host_state: rows of (x, v, omega) with a cached id table
    advance exactly like lethe_dem_step_host rows."""
    from tests.util import packing_parameters, random_packing

    ids, x, props, extent = random_packing(5, d=0.005, spacing=0.99, seed=3)
    params = packing_parameters(extent, d=0.005)
    engines = []
    for _ in range(2):
        e = loader.oracle_engine(params.to_config())
        e.set_walls(box_wall_faces(params.mesh))
        e.set_particles(ids, x, props)
        engines.append(e)
    hx, hp = x.copy(), props.copy()
    rows = np.ascontiguousarray(np.concatenate([x, props[:, 3:9]], axis=1))
    for k in range(10):
        engines[0].step_host(1, ids, hx, hp)
        engines[1].step_host_state(1, ids if k == 0 else None, rows)
    assert np.array_equal(rows[:, :3], hx) and np.array_equal(rows[:, 3:], hp[:, 3:9])
    with pytest.raises(abi.DEMError):
        engines[1].step_host_state(1, None, rows[:-1].copy())


def test_cfd_dem_subcycle_external_loads(oracle_lib):
    """The CFD-DEM caller's pattern (fem-dem/cfd_dem_coupling.cc:1380-1540): constant
    fluid-particle loads over the DEM sub-iterations of a CFD step, the first sub-iteration an
    opening step, velocities synchronised at the end of the CFD step. Velocity Verlet is exact for
    constant loads: after K CFD steps x = a t^2 / 2, v = a t, omega = T t / I — and identical to
    the uninterrupted run."""
    p = unit_test_parameters(dt=1e-4, g=(0, 0, 0))
    mass, d = 2.0, 0.01
    force, torque = np.array([0.3, -0.2, 0.1]), np.array([1e-6, 2e-6, -3e-6])
    moi = 0.1 * mass * d * d

    def run(cfd_steps, sub, subcycle):
        e = loader.oracle_engine(p.to_config())
        e.set_walls(box_wall_faces(p.mesh))
        e.set_particles([7], [[0.0, 0.0, 0.0]], [props_row(0, d, mass)])
        e.set_external_loads([7], [force], [torque])
        for _ in range(cfd_steps):
            if subcycle:
                e.restart_integration()
            e.step(sub)
            if subcycle:
                e.synchronize_velocities()
        if not subcycle:
            e.synchronize_velocities()
        return e.get_particles()

    t = 5 * 20 * 1e-4
    for subcycle in (True, False):
        _, x, props = run(5, 20, subcycle)
        assert np.allclose(x[0], 0.5 * force / mass * t * t, rtol=1e-12, atol=0)
        assert np.allclose(props[0, 3:6], force / mass * t, rtol=1e-12, atol=0)
        assert np.allclose(props[0, 6:9], torque / moi * t, rtol=1e-12, atol=0)
    # cleared loads: the particle coasts
    e = loader.oracle_engine(p.to_config())
    e.set_walls(box_wall_faces(p.mesh))
    e.set_particles([7], [[0.0, 0.0, 0.0]], [props_row(0, d, mass)])
    e.set_external_loads([7], [force])
    e.set_external_loads([], [])
    e.step(10)
    _, x, _ = e.get_particles()
    assert np.array_equal(x[0], np.zeros(3))


def mobility_status_case(case, engine_factory, ranks=1):
    """Runs an adaptive-sparse-contacts application golden through the .prm mirror + DEMSolver and
    returns (solver, per-cell status in lexicographic order, golden status, golden log)."""
    import json

    d = os.path.join(GOLDEN, "apps")
    params = load_prm(os.path.join(d, case + ".prm"))
    assert params.sparse_contacts
    solver = DEMSolver(params, engine_factory=engine_factory, prm_directory=d, reference_insertion_ranks=ranks)
    logged = {}

    def log(iteration):
        st = solver.engine.get_stats()
        logged[iteration] = (st.n_rebuilds, st.v_min, st.v_max, st.v_sum / st.n_particles, st.omega_min, st.omega_max, st.omega_sum / st.n_particles)

    solver.solve(log_callback=log)
    solver.logged = logged
    with open(os.path.join(d, "mobility_status.json")) as f:
        gold = json.load(f)[case]
    m = params.mesh
    want = np.full(m.n[0] * m.n[1] * m.n[2], -1)
    for c in gold["cells"]:
        idx = [int(round((c["lo"][k] - m.lo[k]) / m.cell_size[k])) for k in range(3)]
        want[idx[0] + m.n[0] * (idx[1] + m.n[1] * idx[2])] = c["status"]
    assert (want >= 0).all()
    return solver, solver.engine.get_mobility_status(), want, gold["log"]


def test_mobility_status_application_golden(oracle_lib):
    """applications_tests/lethe-particles/mobility_status.{prm,output}: 132 spheres poured on a
    floating wall in a 3 x 16 x 1 grid with adaptive sparse contacts, 1000 steps, contact search at
    every step. The golden prints the mobility status of the 48 cells at the end
    (AdaptiveSparseContacts::identify_mobility_status, adaptive_sparse_contacts.cc:132-356) and the
    particle statistics every 100 iterations."""
    solver, got, want, log = mobility_status_case("mobility_status", loader.oracle_engine)
    assert solver.engine.get_stats().n_rebuilds == 1001  # `Contact list generation` total of the closing table
    assert np.array_equal(got, want), (got.reshape(16, 3), want.reshape(16, 3))
    # logged statistics (5 printed digits): searches so far, |v| and |omega| min / max / average
    for block in log:
        it = block["iteration"]
        if it == "synchronized":
            continue
        n_searches, vmin, vmax, vavg, wmin, wmax, wavg = solver.logged[it]
        assert n_searches == int(block["Contact list generation"][3]), (it, n_searches)
        for got_v, gold_v in zip((vmin, vmax, vavg, wmin, wmax, wavg), block["Velocity magnitude"][:3] + block["Angular velocity magnitude"][:3]):
            assert abs(got_v - gold_v) <= 5.1e-5 * abs(gold_v) + 1e-300, (it, got_v, gold_v)


def test_load_balancing_mobility_status_golden_first_block(oracle_lib):
    """applications_tests/lethe-particles/load_balancing_mobility_status.{prm,mpirun=2.output}: the same
    pour on 2 MPI ranks with `dynamic_with_sparse_contacts` load balancing. Each rank pairs its share
    of the insertion lattice with its own random vector (reference_insertion_ranks = 2); with that the
    single-domain oracle reproduces the statistics of the first logged block (iteration 100, before
    the first contact) to the 5 printed digits. From the first contacts on, the 2-rank run keeps one
    tangential-history copy per rank for pairs that straddle the cut (update_fine_search_candidates.cc:
    136-152 restarts it when a partner changes owner) and repartitions at iteration 400 (clearing all
    histories, dem_action_manager.h:223-233): a single-domain run cannot follow it further, so the later
    blocks and the final statuses of this golden are not asserted; the bottom of the bed (rows 0-3 of
    the 16) agrees anyway."""
    solver, got, want, log = mobility_status_case("load_balancing_mobility_status", loader.oracle_engine, ranks=2)
    block = log[0]
    assert block["iteration"] == 100
    n_searches, vmin, vmax, vavg, wmin, wmax, wavg = solver.logged[100]
    assert n_searches == int(block["Contact list generation"][3])
    for got_v, gold_v in zip((vmin, vmax, vavg, wmin, wmax, wavg), block["Velocity magnitude"][:3] + block["Angular velocity magnitude"][:3]):
        assert abs(got_v - gold_v) <= 5.1e-5 * abs(gold_v) + 1e-300, (got_v, gold_v)
    assert np.array_equal(got.reshape(16, 3)[:4], want.reshape(16, 3)[:4])


def _checkpoint_restart_run(engine_factory, tmp_path, tag):
    """packing_in_box for 300 steps, checkpoint, 200 more steps (A); a second solver restarts from
    the checkpoint files with `set restart = true` and runs the same 200 steps (B). Returns both
    final states and the state at the checkpoint."""
    import copy

    p = load_prm(os.path.join(GOLDEN, "packing_in_box.prm"))
    p.time_end = 500 * p.time_step
    a = DEMSolver(p, engine_factory=engine_factory)
    a.solve(max_steps=300)  # ends with the closing half kick, like a run whose time_end is here
    prefix = str(tmp_path / f"restart_{tag}")
    a.write_checkpoint(prefix)
    ids_c, x_c, props_c = a.engine.get_particles()
    pr = copy.deepcopy(p)
    pr.restart = True
    b = DEMSolver(pr, engine_factory=engine_factory)
    b.read_checkpoint(prefix)
    assert b.iteration_number == 300 and abs(b.current_time - 300 * p.time_step) < 1e-15
    ids_b0, x_b0, props_b0 = b.engine.get_particles()
    assert np.array_equal(ids_b0, ids_c) and np.array_equal(x_b0, x_c) and np.array_equal(props_b0, props_c)
    out_b = b.solve()
    assert b.iteration_number == 500
    return (ids_c, x_c, props_c), out_b, b


def test_checkpoint_restart_round_trip(oracle_lib, tmp_path):
    """write_checkpoint.cc / read_checkpoint.cc semantics on the oracle: the files carry the
    simulation control and the particles; the restarted run resumes at the checkpointed iteration with
    regular integrate() steps (dem.cc:1162-1171), a forced contact search and empty contact
    histories (dem_action_manager.h:185-200); the state right after reading is bit-identical to the
    state written."""
    (ids_c, x_c, props_c), (ids, x, props), solver = _checkpoint_restart_run(loader.oracle_engine, tmp_path, "oracle")
    assert np.array_equal(ids, ids_c)
    assert np.isfinite(x).all() and np.abs(x - x_c).max() > 0  # it moved on
    with open(str(tmp_path / "restart_oracle.simulationcontrol")) as f:
        text = f.read()
    assert text.startswith("Simulation control\n") and "Iter 300" in text  # the reference's text layout


def cfd_rows(props9, rng, scale_f, scale_t):
    """Rows of DEM::CFDDEMProperties (23 doubles): the 9 DEM properties + random fluid loads."""
    n = len(props9)
    rows = np.zeros((n, 23))
    rows[:, :9] = props9
    rows[:, 9:18] = rng.normal(size=(n, 9)) * scale_f[:, None]  # two-way, one-way, drag
    rows[:, 18:21] = rng.normal(size=(n, 3)) * scale_t[:, None]  # fem_torque
    rows[:, 21] = rng.uniform(0, 1, n)  # volumetric contribution: the fluid solver's, carried untouched
    rows[:, 22] = rng.uniform(0, 1, n)  # momentum transfer coefficient
    return rows


def test_cfd_property_rows(oracle_lib):
    """The 23-property particle record of the CFD-DEM solver (dem_properties.h:92-142) through
    lethe_dem_set_particles_cfd / update_loads_cfd / get_particles_cfd: identical, bit for bit, to
    feeding the 9 DEM properties and the summed loads separately — the sum taken in the reference's
    order (two_way + one_way) + drag (cfd_dem_coupling.cc:891-902) — and columns 9-22 come back untouched."""
    d = 0.005
    ids, x, props, extent = random_packing(6, d=d, spacing=0.98, jitter=0.08, poly=0.2, seed=5)
    params = packing_parameters(extent, d=d)
    rng = np.random.default_rng(11)
    rows = cfd_rows(props, rng, 20.0 * props[:, 2], props[:, 2] * d)
    a, b = loader.oracle_engine(params.to_config()), loader.oracle_engine(params.to_config())
    for e in (a, b):
        e.set_walls(box_wall_faces(params.mesh))
    a.set_particles_cfd(ids, x, rows)
    b.set_particles(ids, x, props)
    b.set_external_loads(ids, (rows[:, 9:12] + rows[:, 12:15]) + rows[:, 15:18], rows[:, 18:21])
    for e in (a, b):
        e.restart_integration()
        e.step(25)
        e.synchronize_velocities()
    order = np.argsort(ids)
    ia, xa, ra = a.get_particles_cfd(rows[order])
    ib, xb, pb = b.get_particles()
    assert np.array_equal(ia, ib) and np.array_equal(xa, xb) and np.array_equal(ra[:, :9], pb)
    assert np.array_equal(ra[:, 9:], rows[order][:, 9:])
    assert np.abs(xa - x[order]).max() > 0
    # new loads for the next CFD step
    rows2 = cfd_rows(props, rng, 10.0 * props[:, 2], props[:, 2] * d)
    a.update_loads_cfd(ids, rows2)
    b.set_external_loads(ids, (rows2[:, 9:12] + rows2[:, 12:15]) + rows2[:, 15:18], rows2[:, 18:21])
    for e in (a, b):
        e.restart_integration()
        e.step(10)
    assert np.array_equal(a.get_particles()[1], b.get_particles()[1])


def thermal_properties(real_young=65e9, conductivity=1.0):
    """lpp of tests/dem/particle_particle_heat_transfer.cc:62-72 (glass beads in air)."""
    from lethe_b200 import abi

    th = abi.ThermalProperties()
    for t in range(abi.MAX_TYPES):
        th.real_youngs_modulus[t] = real_young
        th.surface_roughness[t] = 25e-9
        th.surface_slope[t] = 0.078
        th.microhardness[t] = 9e9
        th.thermal_conductivity[t] = conductivity * (1 + 0.5 * t)
        th.thermal_accommodation[t] = 0.7
    th.thermal_conductivity_gas = 0.027
    th.dynamic_viscosity_gas = 1.85e-5
    th.specific_heat_gas = 1006
    th.specific_heats_ratio_gas = 1
    th.molecular_mean_free_path_gas = 68e-9
    return th


def heat_transfer_unit_case(engine_factory):
    """tests/dem/particle_particle_heat_transfer.cc: two 1 cm glass spheres, 10 um overlap, 600 K and 100 K."""
    p = unit_test_parameters(dt=0.1, g=(0, 0, 0))
    t = p.particle_types[0]
    t.young, t.poisson, t.restitution, t.friction, t.rolling_friction, t.rolling_viscous_damping = 65e9, 0.22, 0.8, 1.0, 0.02, 0.0
    t.diameter = 0.01
    p.rolling_model = "constant"
    e = engine_factory(p.to_config(store_forces=True, moi_override=1.0))
    e.enable_heat_transfer(thermal_properties())
    e.set_particles([0, 1], [[0, 0, 0], [0.00999, 0, 0]], [props_row(0, 0.01, 1), props_row(0, 0.01, 1)])
    e.set_temperatures([0, 1], [600, 100], [840, 840])
    e.step(1)
    return e


def test_particle_particle_heat_transfer_golden(oracle_lib):
    """DEM-MP: tests/dem/particle_particle_heat_transfer.output (-1.56046 J/s on particle one) through the
    oracle's contact loop, and the temperature step that follows (integrate_temperature)."""
    e = heat_transfer_unit_case(loader.oracle_engine)
    ids, temp, rate = e.get_temperatures()
    assert_sig6(rate[0], -1.56046)
    assert rate[1] == -rate[0]
    assert temp[0] == 600 + 0.1 * rate[0] * (1 / 1.0) * (1 / 840.0)
    assert temp[1] == 100 + 0.1 * rate[1] * (1 / 1.0) * (1 / 840.0)


def test_thermal_resistances_golden(oracle_lib):
    """tests/dem/particle_particle_thermal_resistances.output: contact radius and the five resistances of
    calculate_contact_thermal_conductance for two 1 cm spheres with 100 um overlap, E = 5 MPa simulated /
    65 GPa real. The normal force comes from the contact model, as in the reference test."""
    import ctypes

    p = unit_test_parameters(dt=0.001, g=(0, 0, 0))
    t = p.particle_types[0]
    t.young, t.poisson, t.restitution, t.friction, t.rolling_friction, t.diameter = 5e6, 0.22, 0.8, 1.0, 0.02, 0.01
    p.rolling_model = "constant"
    e = loader.oracle_engine(p.to_config(store_forces=True, moi_override=1.0))
    e.set_particles([0, 1], [[0, 0, 0], [0.0099, 0, 0]], [props_row(0, 0.01, 1), props_row(0, 0.01, 1)])
    e.step(1)
    _, f, _ = e.get_forces()
    normal_force_norm = float(np.sqrt((f[0] ** 2).sum()))
    nu = 0.22
    e_eff, e_real = 5e6 / (2.0 * (1.0 - nu * nu)), 65e9 / (2.0 * (1.0 - nu * nu))
    prandtl = 1.85e-5 * 1006 / 0.027
    m = 2.0 * (2.0 - 0.7) / 0.7 * (2.0 * 1) / (1.0 + 1) * 68e-9 / prandtl
    overlap = (0.005 + 0.005) - 0.0099
    vin = (ctypes.c_double * 13)(0.005, 0.005, e_eff, e_real, 25e-9, 0.078, 9e9, 1.0, 1.0, 0.027, m, overlap, normal_force_norm)
    out = (ctypes.c_double * 8)()
    oracle_lib.oracle_dem_thermal_resistances(vin, out)
    gold = [7.51931e-05, 6649.55, 2992.28, 100.023, 1089.98, 255.915, 339.705, 0.00294373]
    for got, want in zip(out, gold):
        assert_sig6(got, want)


def test_temperature_integration_order(oracle_lib):
    """tests/dem/integration_temperature.cc: dT/dt = -T / (m c_p) integrated explicitly is first order
    (golden: 1.00001). Driven through the engine with a partner that is too far to touch, the rate set
    by hand is not available, so the scheme is checked on its closed form: T_n = T_0 (1 - dt/(m c_p))^n."""
    for dt, n in ((0.01, 100), (0.005, 200)):
        T = 300.0
        for _ in range(n):
            T += dt * (-T + 0.0) * (1 / 1.0) * (1 / 840.0)
        assert abs(T - 300.0 * (1 - dt / 840.0) ** n) < 1e-10
    e1 = 300.0 * (1 - 0.01 / 840.0) ** 100 - 300.0 * math.exp(-1.0 / 840.0)
    e2 = 300.0 * (1 - 0.005 / 840.0) ** 200 - 300.0 * math.exp(-1.0 / 840.0)
    assert abs(math.log(e1 / e2) / math.log(2.0) - 1.0) < 1e-3
