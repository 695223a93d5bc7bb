"""The C++ host side (lethe_b200/host): `.prm` reader, DEMSolver mirror, lethe-particles-b200.

CPU part: the product binary's --dump-config must agree with the Python mirror's
to_config() for the reference's own parameter file, and the same host sources, compiled
here against the CPU oracle's identically-shaped ABI (test-only binary under tests/_build),
must reproduce applications_tests/lethe-particles/packing_in_box.mpirun=1.output.
GPU part: the product binary itself against the same golden."""
import json
import os
import subprocess

import numpy as np
import pytest

from lethe_b200.prm import load_prm
from lethe_b200.solver import box_wall_faces
from oracle import loader
from tests.util import GOLDEN, golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "lethe_b200", "host")
PRM = os.path.join(GOLDEN, "packing_in_box.prm")


def product_binary():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "lethe_b200", "csrc"), "-s", "-j8"])
    subprocess.check_call(["make", "-C", HOST, "-s"])
    return os.path.join(HOST, "lethe-particles-b200")


def oracle_host_binary():
    """The host sources compiled against the CPU oracle's identically-shaped ABI (test-only binary
    under tests/_build); rebuilt only when a source is newer."""
    loader.build()
    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, "lethe-particles-oracle")
    srcs = [os.path.join(HOST, f) for f in ("dem_parameters.cc", "dem_solver.cc", "lethe_particles_b200.cc")]
    odir = os.path.join(ROOT, "oracle")
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")] + [os.path.join(odir, "libdem_oracle.so"),
                                                                                            os.path.join(ROOT, "include", "lethe_dem.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-DLETHE_DEM_ABI_PREFIX=oracle_dem_", "-o", exe, *srcs, "-L" + odir,
                               "-ldem_oracle", "-Wl,-rpath," + odir])
    return exe


def parse_xyz(text):
    rows = []
    for line in text.splitlines():
        parts = line.split()
        if len(parts) == 6 and parts[0].isdigit():
            rows.append([int(parts[0]), int(parts[1])] + [float(v) for v in parts[2:]])
    return rows


def check_against_golden(rows):
    gold = golden("packing_in_box.mpirun1.json")["rows"]
    assert len(rows) == len(gold) == 200
    assert [r[0] for r in rows] == [g[0] for g in gold]
    err = np.abs(np.array([r[3:6] for r in rows]) - np.array([g[3:6] for g in gold])).max(axis=1)
    assert np.mean(err <= 1.01e-4) > 0.9 and err.max() < 0.2 * 0.005, (np.mean(err <= 1.01e-4), err.max())


def test_dump_config_matches_python_mirror():
    exe = product_binary()
    out = subprocess.run([exe, PRM, "--dump-config"], capture_output=True, text=True, check=True).stdout
    c = json.loads(out)
    p = load_prm(PRM)
    ref = p.to_config()
    for key in ("pp_model", "pw_model", "rolling_model", "detection", "contact_detection_frequency", "cell_order", "n_types", "restart"):
        assert c[key] == getattr(ref, key), key
    for key in ("dt", "neighborhood_threshold", "d_max", "smallest_contact_search_criterion", "dmt_cut_off_threshold",
                "f_coefficient_epsd", "young_wall", "friction_wall"):
        assert c[key] == getattr(ref, key), key
    for key in ("g", "grid_lo", "cell_size", "grid_n", "periodic"):
        assert list(c[key]) == list(getattr(ref, key)), key
    assert c["young"][0] == ref.young[0] and c["friction"][0] == ref.friction[0]
    assert c["n_wall_faces"] == len(box_wall_faces(p.mesh, p.outlet_boundaries, p.periodic))


def test_product_binary_fails_loudly_without_gpu_or_runs():
    """No CPU fallback: without a CUDA device the binary exits 1 with the reference's banner."""
    exe = product_binary()
    r = subprocess.run([exe, PRM, "--quiet"], capture_output=True, text=True)
    if r.returncode != 0:
        assert r.returncode == 1 and "no CUDA device" in r.stderr and "Aborting!" in r.stderr
    else:
        check_against_golden(parse_xyz(r.stdout))


def test_host_sources_against_oracle_reproduce_reference_golden(tmp_path):
    exe = oracle_host_binary()
    r = subprocess.run([exe, PRM, "--quiet"], capture_output=True, text=True, check=True)
    check_against_golden(parse_xyz(r.stdout))


@pytest.mark.gpu
def test_product_binary_reproduces_reference_golden_on_gpu(tmp_path):
    exe = product_binary()
    r = subprocess.run([exe, PRM, "--quiet"], capture_output=True, text=True, check=True)
    check_against_golden(parse_xyz(r.stdout))
    # and the log path (progression banner + statistics table) runs: the golden case logs every
    # 1e6 iterations (never, in 1e4 steps), so use a copy that logs every 2500
    with open(PRM) as f:
        text = f.read().replace("set log frequency    = 1000000", "set log frequency    = 2500")
    assert "= 2500" in text
    prm2 = tmp_path / "packing_in_box_log.prm"
    prm2.write_text(text)
    r2 = subprocess.run([exe, str(prm2)], capture_output=True, text=True, check=True)
    assert "Transient iteration:" in r2.stdout and "Contact list generation" in r2.stdout
    check_against_golden(parse_xyz(r2.stdout[r2.stdout.index("id, type"):]))


def test_host_sources_solid_surface_prm_against_oracle():
    """The reference's particle_solid_surface_NPES_double_edge_contact.prm through the C++ host
    (prm reader incl. `solid objects` and `insertion method = list`, gmsh reader, DEMSolver mirror,
    statistics log) linked to the CPU oracle: the logged "Velocity magnitude" column equals the
    reference's .output to its 5 printed digits at all 60 log lines."""
    import re

    exe = oracle_host_binary()
    prm = os.path.join(GOLDEN, "solid_surfaces", "particle_solid_surface_NPES_double_edge_contact.prm")
    r = subprocess.run([exe, prm], capture_output=True, text=True, check=True)
    got = [float(m) for m in re.findall(r"Velocity magnitude\s*\|\s*\S+\s*\|\s*(\S+)", r.stdout)]
    with open(os.path.join(GOLDEN, "solid_surface_goldens.json")) as f:
        gold = json.load(f)["NPES_double_edge_contact"]["velocity_magnitude"]
    assert len(got) == len(gold) == 60
    assert all(abs(a - b) <= 1.01e-4 * abs(b) + 1e-12 for a, b in zip(got, gold)), list(zip(got, gold))[:5]


@pytest.mark.parametrize("case", ["insert_file_3d", "epsd_rolling_resistance_model", "moving_solid_surface_hmlo", "sliding_in_box",
                                  "distribution_normal", "distribution_lognormal", "solid_surface", "deprecated_parameters",
                                  "insert_list_3d_default_velocities", "insertion_acceptance_function", "insert_plane_3d",
                                  "initial_value_insertion", "insert_and_remove_with_files"])
def test_host_sources_application_goldens_against_oracle(case):
    """More of the reference's application cases through the C++ host mirror (file / list / volume
    insertion, solid objects, EPSD) linked to the oracle: the printed final table equals the
    reference's .output to its 4 decimals."""
    exe = oracle_host_binary()
    r = subprocess.run([exe, os.path.join(GOLDEN, "apps", case + ".prm"), "--quiet"], capture_output=True, text=True, check=True)
    rows = parse_xyz(r.stdout)
    with open(os.path.join(GOLDEN, "apps", "final_positions.json")) as f:
        gold = json.load(f)[case]
    assert [r_[0] for r_ in rows] == [g[0] for g in gold]
    err = np.abs(np.array([r_[3:6] for r_ in rows]) - np.array([g[3:6] for g in gold])).max()
    assert err <= 1.01e-4, (case, err)  # both sides print 4 decimals
    derr = np.abs(np.array([r_[2] for r_ in rows]) - np.array([g[2] for g in gold])).max()
    assert derr <= 1.01e-5, (case, derr)  # diameters: 5 decimals


def test_host_sources_time_dependent_solid_velocity(tmp_path):
    """A solid-object velocity that depends on time (muparser conditional and power, as in
    load_balancing_solid_object.prm) through the C++ host's own expression evaluator
    (function_expression.h): same final table as the Python mirror on the same engine."""
    import shutil

    from lethe_b200.prm import load_prm
    from lethe_b200.solver import DEMSolver

    src = os.path.join(GOLDEN, "apps", "moving_solid_surface_hmlo.prm")
    case_dir = tmp_path / "apps"
    case_dir.mkdir()
    shutil.copy(os.path.join(GOLDEN, "square.msh"), tmp_path / "square.msh")
    with open(src) as f:
        text = f.read().replace("set Function expression = -0.1 ; 0 ; 0", "set Function expression = if(t < 0.2, -0.1, -0.4 * t^2) ; 0 ; 0.01*sin(2*pi*t)")
    assert "sin(2*pi*t)" in text
    text = text.replace("set time end         = 1", "set time end         = 0.5")
    prm = case_dir / "moving.prm"
    prm.write_text(text)
    r = subprocess.run([oracle_host_binary(), str(prm), "--quiet"], capture_output=True, text=True, check=True)
    rows = parse_xyz(r.stdout)
    solver = DEMSolver(load_prm(str(prm)), engine_factory=loader.oracle_engine, prm_directory=str(case_dir))
    ids, x, props = solver.solve()
    assert [r_[0] for r_ in rows] == list(ids) and len(ids) > 0
    assert np.abs(np.array([r_[3:6] for r_ in rows]) - x).max() <= 0.51e-4
    # and the motion did something: not the constant-velocity golden
    with open(os.path.join(GOLDEN, "apps", "final_positions.json")) as f:
        gold = np.array([g[3:6] for g in json.load(f)["moving_solid_surface_hmlo"]])
    assert np.abs(x - gold).max() > 1e-3


def test_function_expression_evaluator_matches_python_mirror(tmp_path):
    """lethe_b200/host/function_expression.h against the Python mirror's evaluator on the
    expression forms the reference's DEM parameter files use (and a few more): same values,
    constancy detected, malformed input refused."""
    from lethe_b200.prm import evaluate_function

    exe = str(tmp_path / "eval_expression")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "eval_expression.cc")])
    point, t = (0.01, 0.07, -0.02), 0.6
    cases = ["if( (x*x + z * z) < (0.025^2), if(y > 0.05, 1., -1), -1.)", "if(t>0.5,if(t<0.7,1,0),0)", "-0.4 * t^2", "0.01*sin(2*pi*t)",
             "2^-1 + 3*(x - y)/z", "-x^2", "exp(-t)*cos(pi*t) + sqrt(abs(z))", "min(x, y) + max(y, z)", "1e-3 + 2.5E2*t", "3", "tanh(t) - log(1 + t)"]
    out = subprocess.run([exe, *(repr(v) for v in point), repr(t), *cases], capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(out) == len(cases)
    for expr, line in zip(cases, out):
        value, constant = line.split()
        expected = evaluate_function(expr, t, point)
        assert abs(float(value) - expected) <= 1e-15 * max(1.0, abs(expected)), (expr, value, expected)
        assert int(constant) == (0 if any(c in expr.replace("exp", "").replace("max", "").replace("tanh", "") for c in "xyzt") else 1), expr
    bad = subprocess.run([exe, "0", "0", "0", "0", "2 +", "foo(1)", "if(1, 2)", "1 2"], capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(bad) == 4 and all(line.startswith("error:") for line in bad), bad
