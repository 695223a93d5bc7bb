"""Loads oracle/libdem_oracle.so through the same ctypes binding as the CUDA
engine.  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs — never by lethe_b200."""
import ctypes
import os
import subprocess

from lethe_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libdem_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "dem_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB


_lib = None


def library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
    return _lib


def oracle_engine(config, device=0):
    return abi.Engine(library(), "oracle_dem_", config, device)


def pair_force(engine, x1, p1, x2, p2, tangential=None, rolling=None):
    """One particle-particle contact evaluation (oracle_dem_pair_force)."""
    import numpy as np

    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x1, p1, x2, p2 = f(x1), f(p1), f(x2), f(p2)
    t = f(tangential if tangential is not None else np.zeros(3)).copy()
    r = f(rolling if rolling is not None else np.zeros(3)).copy()
    f1, t1, f2, t2 = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3)
    ov = ctypes.c_double()
    P = ctypes.POINTER(ctypes.c_double)
    fn = library().oracle_dem_pair_force
    fn.restype = ctypes.c_int
    rc = fn(engine._ctx, *[a.ctypes.data_as(P) for a in (x1, p1, x2, p2, t, r, f1, t1, f2, t2)], ctypes.byref(ov))
    assert rc == 0
    return dict(force_one=f1, torque_one=t1, force_two=f2, torque_two=t2, tangential=t, rolling=r, overlap=ov.value)


def integrate_external(engine, phase, force, torque, moi):
    import numpy as np

    P = ctypes.POINTER(ctypes.c_double)
    f = np.ascontiguousarray(force, dtype=np.float64)
    t = np.ascontiguousarray(torque, dtype=np.float64)
    fn = library().oracle_dem_integrate_external
    fn.restype = ctypes.c_int
    rc = fn(engine._ctx, ctypes.c_int(phase), f.ctypes.data_as(P), t.ctypes.data_as(P), ctypes.c_double(moi))
    assert rc == 0


def set_option(engine, name, value):
    """oracle_dem_set_option: test-only switches of the oracle (e.g. dmt_stale_scratch)."""
    fn = library().oracle_dem_set_option
    fn.restype = ctypes.c_int
    rc = fn(engine._ctx, name.encode(), ctypes.c_int(int(value)))
    assert rc == 0, name


def cell_neighbors(engine, which, stride=32):
    """Cell-neighbour lists of the oracle in active-cell numbering (test hook, see dem_oracle.cpp)."""
    import ctypes as C

    import numpy as np

    n_cells = int(np.prod(list(engine.config.grid_n)))
    out = np.full((n_cells, stride), -1, dtype=np.int32)
    fn = library().oracle_dem_cell_neighbors
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    fn.restype = C.c_int
    assert fn(engine._ctx, which, stride, out.ctypes.data) == 0
    return [[int(v) for v in row if v >= 0] for row in out]
