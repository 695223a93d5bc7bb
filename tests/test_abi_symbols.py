"""The C-ABI shared library loads (no GPU needed for that) and exports every function that
include/lethe_dem.h declares; the oracle exports the same set under its own prefix; the
product package never references the oracle."""
import ctypes
import os
import re
import subprocess

from lethe_b200 import abi
from oracle import loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "lethe_dem.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lethe_dem_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("lethe_dem_create", "lethe_dem_destroy", "lethe_dem_set_particles", "lethe_dem_add_particles", "lethe_dem_set_walls",
                 "lethe_dem_set_floating_walls", "lethe_dem_set_boundary_motion", "lethe_dem_step", "lethe_dem_step_host", "lethe_dem_step_host_state", "lethe_dem_set_external_loads", "lethe_dem_restart_integration",
                 "lethe_dem_synchronize_velocities", "lethe_dem_force_contact_search", "lethe_dem_get_particles", "lethe_dem_get_pairs",
                 "lethe_dem_get_forces", "lethe_dem_get_stats", "lethe_dem_comm_init"):
        assert must in names


def test_cuda_library_loads_and_exports_every_declared_symbol():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "lethe_b200", "csrc"), "-s", "-j8"])
    lib = ctypes.CDLL(abi.CUDA_LIB)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    # the python binding covers the same set
    assert sorted("lethe_dem_" + s for s in abi.ABI_SYMBOLS) == declared_functions()


def test_oracle_exports_the_same_interface():
    loader.build()
    lib = ctypes.CDLL(loader.LIB)
    # host arithmetic of the slab decomposition / counters of the host<->device transfer schedule: no counterpart in a CPU oracle
    skip = {"lethe_dem_balanced_cuts", "lethe_dem_host_pipeline_stats", "lethe_dem_get_transfer_order", "lethe_dem_get_state_rows"}
    missing = [n for n in declared_functions() if n not in skip and not hasattr(lib, n.replace("lethe_dem_", "oracle_dem_"))]
    assert not missing, missing


def test_product_package_never_touches_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "lethe_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")) or f == "Makefile":
                text = open(os.path.join(d, f)).read()
                for ln in text.splitlines():
                    s = ln.strip()
                    if s.startswith(("#include", "import ", "from ")) and "oracle" in s:
                        bad.append((f, s))
                    if "-ldem_oracle" in s or "libdem_oracle" in s:
                        bad.append((f, s))
    assert not bad, bad
