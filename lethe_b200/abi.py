"""ctypes binding of include/lethe_dem.h.

The binding is generic over the symbol prefix: the CUDA engine exports
``lethe_dem_*`` (lethe_b200/csrc/liblethe_dem_b200.so); the CPU oracle used by
the tests exports the same functions as ``oracle_dem_*``.  This module never
loads the oracle by itself and has no CPU fallback: `load_engine()` raises if
the CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MAX_TYPES = 5
MAX_FLOATING_WALLS = 9
N_PROPERTIES = 9
NCCL_ID_BYTES = 128

# enum values of include/lethe_dem.h
PP_MODELS = {
    "linear": 0,
    "hertz_mindlin_limit_force": 1,
    "hertz_mindlin_limit_overlap": 2,
    "hertz": 3,
    "hertz_JKR": 4,
    "DMT": 5,
}
PW_MODELS = {"linear": 0, "nonlinear": 1, "JKR": 2, "DMT": 3}
ROLLING_MODELS = {"none": 0, "constant": 1, "viscous": 2, "epsd": 3}
DETECTION = {"dynamic": 0, "constant": 1}
CELL_ORDER = {"lexicographic": 0, "morton": 1}

_d5 = C.c_double * MAX_TYPES
_d3 = C.c_double * 3
_i3 = C.c_int32 * 3


class Config(C.Structure):
    """struct lethe_dem_config"""

    _fields_ = [
        ("pp_model", C.c_int32),
        ("pw_model", C.c_int32),
        ("rolling_model", C.c_int32),
        ("integrator", C.c_int32),
        ("detection", C.c_int32),
        ("contact_detection_frequency", C.c_int32),
        ("cell_order", C.c_int32),
        ("store_forces", C.c_int32),
        ("dt", C.c_double),
        ("g", _d3),
        ("neighborhood_threshold", C.c_double),
        ("d_max", C.c_double),
        ("smallest_contact_search_criterion", C.c_double),
        ("dmt_cut_off_threshold", C.c_double),
        ("f_coefficient_epsd", C.c_double),
        ("moi_override", C.c_double),
        ("n_types", C.c_int32),
        ("restart", C.c_int32),
        ("young", _d5),
        ("poisson", _d5),
        ("restitution", _d5),
        ("friction", _d5),
        ("rolling_friction", _d5),
        ("rolling_viscous_damping", _d5),
        ("surface_energy", _d5),
        ("hamaker", _d5),
        ("young_wall", C.c_double),
        ("poisson_wall", C.c_double),
        ("restitution_wall", C.c_double),
        ("friction_wall", C.c_double),
        ("rolling_friction_wall", C.c_double),
        ("rolling_viscous_damping_wall", C.c_double),
        ("surface_energy_wall", C.c_double),
        ("hamaker_wall", C.c_double),
        ("grid_lo", _d3),
        ("cell_size", _d3),
        ("grid_n", _i3),
        ("periodic", _i3),
        ("slab_axis", C.c_int32),
        ("slab_lo", C.c_int32),
        ("slab_hi", C.c_int32),
        ("sparse_contacts", C.c_int32),
        ("asc_granular_temperature_threshold", C.c_double),
        ("asc_solid_fraction_threshold", C.c_double),
        ("precision", C.c_int32),
        ("pad2", C.c_int32),
    ]


class WallFace(C.Structure):
    """struct lethe_wall_face"""

    _fields_ = [
        ("cell", C.c_int32),
        ("boundary_id", C.c_uint32),
        ("global_face_id", C.c_uint32),
        ("pad", C.c_uint32),
        ("normal", _d3),
        ("point", _d3),
    ]


class Stats(C.Structure):
    """struct lethe_dem_stats"""

    _fields_ = [
        ("n_particles", C.c_uint64),
        ("n_rebuilds", C.c_uint64),
        ("n_steps", C.c_uint64),
        ("n_pair_entries", C.c_uint64),
        ("n_wall_entries", C.c_uint64),
        ("n_pairs_touching", C.c_uint64),
        ("v_min", C.c_double),
        ("v_max", C.c_double),
        ("v_sum", C.c_double),
        ("omega_min", C.c_double),
        ("omega_max", C.c_double),
        ("omega_sum", C.c_double),
        ("ke_trans_min", C.c_double),
        ("ke_trans_max", C.c_double),
        ("ke_trans_sum", C.c_double),
        ("ke_rot_min", C.c_double),
        ("ke_rot_max", C.c_double),
        ("ke_rot_sum", C.c_double),
        ("n_migrated", C.c_uint64),
    ]


# every symbol include/lethe_dem.h declares (suffix after the prefix)
ABI_SYMBOLS = [
    "create",
    "destroy",
    "last_error",
    "create_error",
    "set_particles",
    "add_particles",
    "n_particles",
    "get_particles",
    "set_walls",
    "set_floating_walls",
    "set_boundary_motion",
    "add_solid_surface",
    "set_solid_motion",
    "get_solid_vertices",
    "get_solid_contacts",
    "step",
    "synchronize_velocities",
    "force_contact_search",
    "step_host",
    "step_host_state",
    "host_pipeline_stats",
    "get_transfer_order",
    "get_state_rows",
    "set_external_loads",
    "restart_integration",
    "set_time",
    "enable_heat_transfer",
    "set_temperatures",
    "get_temperatures",
    "set_particles_cfd",
    "update_loads_cfd",
    "get_particles_cfd",
    "get_pairs",
    "get_wall_contacts",
    "get_forces",
    "get_stats",
    "get_mobility_status",
    "get_timers",
    "enable_timers",
    "kernel_launches",
    "event_record",
    "event_elapsed",
    "set_load_balancing",
    "set_load_balancing_weights",
    "get_slab",
    "balanced_cuts",
    "nccl_unique_id",
    "comm_init",
]

class ThermalProperties(C.Structure):
    """Mirror of lethe_dem_thermal_properties."""

    _fields_ = [
        ("real_youngs_modulus", C.c_double * MAX_TYPES),
        ("surface_roughness", C.c_double * MAX_TYPES),
        ("surface_slope", C.c_double * MAX_TYPES),
        ("microhardness", C.c_double * MAX_TYPES),
        ("thermal_conductivity", C.c_double * MAX_TYPES),
        ("thermal_accommodation", C.c_double * MAX_TYPES),
        ("thermal_conductivity_gas", C.c_double),
        ("dynamic_viscosity_gas", C.c_double),
        ("specific_heat_gas", C.c_double),
        ("specific_heats_ratio_gas", C.c_double),
        ("molecular_mean_free_path_gas", C.c_double),
    ]


N_CFD_PROPERTIES = 23
LOAD_BALANCE_METHODS = {"none": 0, "once": 1, "frequent": 2, "dynamic": 3, "dynamic_with_sparse_contacts": 4}

_p_u32 = C.POINTER(C.c_uint32)
_p_f64 = C.POINTER(C.c_double)
_p_u64 = C.POINTER(C.c_uint64)


class DEMError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the reference throws
    std::runtime_error / AssertThrow, applications/lethe-particles/dem.cc:151-177)."""


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, t):
    return a.ctypes.data_as(t)


class Engine:
    """Thin object wrapper over one ``lethe_dem_ctx``."""

    def __init__(self, lib: C.CDLL, prefix: str, config: Config, device: int = 0):
        self._lib = lib
        self._prefix = prefix
        self._ctx = C.c_void_p()
        self.config = config
        create = self._fn("create")
        create.restype = C.c_int
        create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]
        rc = create(C.byref(config), device, C.byref(self._ctx))
        if rc != 0:
            err = self._fn("create_error")
            err.restype = C.c_char_p
            raise DEMError(f"{prefix}create failed ({rc}): {err().decode()}")

    # -- plumbing --
    def _fn(self, name):
        return getattr(self._lib, self._prefix + name)

    def _call(self, name, *args):
        fn = self._fn(name)
        fn.restype = C.c_int
        rc = fn(self._ctx, *args)
        if rc != 0:
            le = self._fn("last_error")
            le.restype = C.c_char_p
            le.argtypes = [C.c_void_p]
            raise DEMError(f"{self._prefix}{name} failed ({rc}): {le(self._ctx).decode()}")

    def close(self):
        if self._ctx:
            d = self._fn("destroy")
            d.restype = None
            d.argtypes = [C.c_void_p]
            d(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state --
    def set_particles(self, ids, x, props, add=False):
        ids, x, props = _u32(ids), _f64(x).reshape(-1, 3), _f64(props).reshape(-1, N_PROPERTIES)
        assert len(ids) == len(x) == len(props)
        self._call(
            "add_particles" if add else "set_particles",
            C.c_uint64(len(ids)),
            _ptr(ids, _p_u32),
            _ptr(x, _p_f64),
            _ptr(props, _p_f64),
        )

    def add_particles(self, ids, x, props):
        self.set_particles(ids, x, props, add=True)

    def n_particles(self) -> int:
        n = C.c_uint64()
        self._call("n_particles", C.byref(n))
        return n.value

    def get_particles(self):
        n = self.n_particles()
        ids = np.empty(n, np.uint32)
        x = np.empty((n, 3), np.float64)
        props = np.empty((n, N_PROPERTIES), np.float64)
        out = C.c_uint64()
        self._call("get_particles", C.c_uint64(n), C.byref(out), _ptr(ids, _p_u32), _ptr(x, _p_f64), _ptr(props, _p_f64))
        return ids[: out.value], x[: out.value], props[: out.value]

    def set_walls(self, faces):
        arr = (WallFace * len(faces))(*faces)
        self._call("set_walls", C.c_uint64(len(faces)), arr)

    def set_floating_walls(self, points, normals, t_start, t_end):
        p, nrm = _f64(points).reshape(-1, 3), _f64(normals).reshape(-1, 3)
        t0, t1 = _f64(t_start), _f64(t_end)
        self._call("set_floating_walls", C.c_int32(len(p)), _ptr(p, _p_f64), _ptr(nrm, _p_f64), _ptr(t0, _p_f64), _ptr(t1, _p_f64))

    def set_boundary_motion(self, boundary_id, translational_velocity=(0, 0, 0), rotational_speed=0.0, rotational_vector=(0, 0, 0), point_on_axis=(0, 0, 0)):
        self._call(
            "set_boundary_motion",
            C.c_uint32(boundary_id),
            _d3(*translational_velocity),
            C.c_double(rotational_speed),
            _d3(*rotational_vector),
            _d3(*point_on_axis),
        )

    def add_solid_surface(self, vertices, triangles, translational_velocity=(0, 0, 0), angular_velocity=(0, 0, 0), center_of_rotation=(0, 0, 0)) -> int:
        """A triangle-mesh solid surface (SerialSolid<2,3>); returns its index."""
        v = _f64(vertices).reshape(-1, 3)
        t = _u32(triangles).reshape(-1, 3)
        idx = C.c_int32(-1)
        self._call("add_solid_surface", C.c_uint32(len(v)), _ptr(v, _p_f64), C.c_uint32(len(t)), _ptr(t, _p_u32),
                   _d3(*translational_velocity), _d3(*angular_velocity), _d3(*center_of_rotation), C.byref(idx))
        self._solid_sizes = getattr(self, "_solid_sizes", {})
        self._solid_sizes[idx.value] = len(v)
        return idx.value

    def set_solid_motion(self, solid, translational_velocity=(0, 0, 0), angular_velocity=(0, 0, 0)):
        self._call("set_solid_motion", C.c_int32(solid), _d3(*translational_velocity), _d3(*angular_velocity))

    def get_solid_vertices(self, solid):
        n = self._solid_sizes[solid]
        out = np.empty((n, 3), np.float64)
        self._call("get_solid_vertices", C.c_int32(solid), C.c_uint32(n), _ptr(out, _p_f64))
        return out

    def get_solid_contacts(self):
        n = C.c_uint64()
        self._call("get_solid_contacts", C.c_uint64(0), C.byref(n), None, None, None, None)
        p = np.empty(n.value, np.uint32)
        sd = np.empty(n.value, np.uint32)
        tr = np.empty(n.value, np.uint32)
        t = np.empty((n.value, 3), np.float64)
        self._call("get_solid_contacts", C.c_uint64(n.value), C.byref(n), _ptr(p, _p_u32), _ptr(sd, _p_u32), _ptr(tr, _p_u32), _ptr(t, _p_f64))
        return p, sd, tr, t

    # -- hot path --
    def step(self, n_steps=1):
        self._call("step", C.c_uint64(n_steps))

    def synchronize_velocities(self):
        self._call("synchronize_velocities")

    def force_contact_search(self, clear_tangential_displacement=False):
        self._call("force_contact_search", C.c_int(int(clear_tangential_displacement)))

    def step_host(self, n_steps, ids, x, props):
        """In-place on the (contiguous) numpy buffers x[n,3], props[n,9]."""
        assert ids.dtype == np.uint32 and x.dtype == np.float64 and props.dtype == np.float64
        assert x.flags.c_contiguous and props.flags.c_contiguous
        self._call("step_host", C.c_uint64(n_steps), C.c_uint64(len(ids)), _ptr(ids, _p_u32), _ptr(x, _p_f64), _ptr(props, _p_f64))

    def set_external_loads(self, ids, force, torque=None):
        """lethe_dem_set_external_loads (CFD-DEM fluid-particle interaction); ids=[] clears all."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        f = np.ascontiguousarray(force, dtype=np.float64).reshape(len(ids), 3) if len(ids) else np.zeros((0, 3))
        t = None if torque is None else np.ascontiguousarray(torque, dtype=np.float64).reshape(len(ids), 3)
        self._call("set_external_loads", C.c_uint64(len(ids)), _ptr(ids, _p_u32), _ptr(f, _p_f64), None if t is None else _ptr(t, _p_f64))

    def restart_integration(self):
        self._call("restart_integration")

    def step_host_state(self, n_steps, ids, state9):
        """lethe_dem_step_host_state: rows of (x, v, omega) up, n_steps, rows down in place;
        ids=None reuses the id table of the previous call."""
        assert state9.dtype == np.float64 and state9.flags.c_contiguous and state9.shape[1] == 9
        id_arg = None if ids is None else _ptr(np.ascontiguousarray(ids, dtype=np.uint32), _p_u32)
        self._call("step_host_state", C.c_uint64(n_steps), C.c_uint64(len(state9)), id_arg, _ptr(state9, _p_f64))

    def get_transfer_order(self):
        """Ids of the owned particles in the row order that overlaps the copies of step_host_state best."""
        n = self.n_particles()
        ids = np.empty(n, np.uint32)
        out = C.c_uint64()
        self._call("get_transfer_order", C.c_uint64(n), C.byref(out), _ptr(ids, _p_u32))
        return ids[: out.value]

    def get_state_rows(self, ids_out=None, state_out=None):
        """Ids and (x, v, omega) rows of the owned particles in transfer order (optionally into caller-owned arrays)."""
        n = self.n_particles()
        ids = np.empty(n, np.uint32) if ids_out is None else ids_out
        state = np.empty((n, 9), np.float64) if state_out is None else state_out
        assert len(ids) >= n and len(state) >= n and ids.dtype == np.uint32 and state.dtype == np.float64
        out = C.c_uint64()
        self._call("get_state_rows", C.c_uint64(n), C.byref(out), _ptr(ids, _p_u32), _ptr(state, _p_f64))
        return ids[: out.value], state[: out.value]

    def host_pipeline_stats(self):
        """(calls of step_host_state that took the streamed form, plans made, streamed calls that wrote the host rows directly)."""
        a, b, d = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._call("host_pipeline_stats", C.byref(a), C.byref(b), C.byref(d))
        return a.value, b.value, d.value

    def step_host_state_ptr(self, n_steps, n, id_ptr, state_ptr):
        """Same on raw host addresses (pinned buffers); id_ptr = 0 reuses the previous id table."""
        self._call("step_host_state", C.c_uint64(n_steps), C.c_uint64(n), C.c_void_p(id_ptr or None), C.c_void_p(state_ptr))

    def step_host_ptr(self, n_steps, n, id_ptr, x_ptr, props_ptr):
        """Same as step_host on raw host addresses (pinned buffers)."""
        self._call("step_host", C.c_uint64(n_steps), C.c_uint64(n), C.c_void_p(id_ptr), C.c_void_p(x_ptr), C.c_void_p(props_ptr))

    # -- taps --
    def get_pairs(self):
        n = C.c_uint64()
        self._call("get_pairs", C.c_uint64(0), C.byref(n), None, None, None)
        i = np.empty(n.value, np.uint32)
        j = np.empty(n.value, np.uint32)
        t = np.empty((n.value, 3), np.float64)
        self._call("get_pairs", C.c_uint64(n.value), C.byref(n), _ptr(i, _p_u32), _ptr(j, _p_u32), _ptr(t, _p_f64))
        return i, j, t

    def get_wall_contacts(self):
        n = C.c_uint64()
        self._call("get_wall_contacts", C.c_uint64(0), C.byref(n), None, None, None)
        p = np.empty(n.value, np.uint32)
        f = np.empty(n.value, np.uint32)
        t = np.empty((n.value, 3), np.float64)
        self._call("get_wall_contacts", C.c_uint64(n.value), C.byref(n), _ptr(p, _p_u32), _ptr(f, _p_u32), _ptr(t, _p_f64))
        return p, f, t

    def get_forces(self):
        n = self.n_particles()
        ids = np.empty(n, np.uint32)
        f = np.empty((n, 3), np.float64)
        t = np.empty((n, 3), np.float64)
        out = C.c_uint64()
        self._call("get_forces", C.c_uint64(n), C.byref(out), _ptr(ids, _p_u32), _ptr(f, _p_f64), _ptr(t, _p_f64))
        return ids[: out.value], f[: out.value], t[: out.value]

    def get_stats(self) -> Stats:
        st = Stats()
        self._call("get_stats", C.byref(st))
        return st

    def enable_timers(self, timers=True, count_touching=False):
        self._call("enable_timers", C.c_int(int(bool(timers)) | (int(bool(count_touching)) << 1)))

    def event_record(self, which: int):
        self._call("event_record", C.c_int(which))

    def event_elapsed_ms(self) -> float:
        ms = C.c_double()
        self._call("event_elapsed", C.byref(ms))
        return ms.value

    # -- CFD-DEM rows of 23 properties (DEM::CFDDEMProperties) --
    def set_particles_cfd(self, ids, x, props23):
        ids, x, props23 = _u32(ids), _f64(x).reshape(-1, 3), _f64(props23).reshape(-1, N_CFD_PROPERTIES)
        self._call("set_particles_cfd", C.c_uint64(len(ids)), _ptr(ids, _p_u32), _ptr(x, _p_f64), _ptr(props23, _p_f64))

    def update_loads_cfd(self, ids, props23):
        ids, props23 = _u32(ids), _f64(props23).reshape(-1, N_CFD_PROPERTIES)
        self._call("update_loads_cfd", C.c_uint64(len(ids)), _ptr(ids, _p_u32), _ptr(props23, _p_f64))

    def get_particles_cfd(self, props23):
        """Rows sorted by id; columns 0-8 of `props23` (one row per particle, in id order) are overwritten."""
        n = self.n_particles()
        ids = np.empty(n, np.uint32)
        x = np.empty((n, 3), np.float64)
        props23 = np.ascontiguousarray(props23, dtype=np.float64).reshape(n, N_CFD_PROPERTIES).copy()
        out = C.c_uint64()
        self._call("get_particles_cfd", C.c_uint64(n), C.byref(out), _ptr(ids, _p_u32), _ptr(x, _p_f64), _ptr(props23, _p_f64))
        return ids, x, props23

    # -- DEM-MP heat transfer --
    def enable_heat_transfer(self, properties: "ThermalProperties"):
        self._call("enable_heat_transfer", C.byref(properties))

    def set_temperatures(self, ids, temperature, specific_heat):
        ids, temperature, specific_heat = _u32(ids), _f64(temperature), _f64(specific_heat)
        assert len(ids) == len(temperature) == len(specific_heat)
        self._call("set_temperatures", C.c_uint64(len(ids)), _ptr(ids, _p_u32), _ptr(temperature, _p_f64), _ptr(specific_heat, _p_f64))

    def get_temperatures(self):
        """(ids, temperature, heat transfer rate of the last step), rows sorted by id."""
        n = self.n_particles()
        ids, temp, rate = np.empty(n, np.uint32), np.empty(n, np.float64), np.empty(n, np.float64)
        out = C.c_uint64()
        self._call("get_temperatures", C.c_uint64(n), C.byref(out), _ptr(ids, _p_u32), _ptr(temp, _p_f64), _ptr(rate, _p_f64))
        return ids[: out.value], temp[: out.value], rate[: out.value]

    def set_time(self, iteration_number: int, current_time: float):
        self._call("set_time", C.c_uint64(iteration_number), C.c_double(current_time))

    def get_mobility_status(self):
        """Per-cell mobility status (lexicographic cell index) as of the last contact search."""
        n = int(self.config.grid_n[0]) * int(self.config.grid_n[1]) * int(self.config.grid_n[2])
        out = np.empty(n, np.int32)
        self._call("get_mobility_status", C.c_uint64(n), _ptr(out, C.POINTER(C.c_int32)))
        return out

    def kernel_launches(self) -> int:
        n = C.c_uint64()
        self._call("kernel_launches", C.byref(n))
        return n.value

    def get_timers(self, reset=True):
        a, b = C.c_double(), C.c_double()
        na, nb = C.c_uint64(), C.c_uint64()
        self._call("get_timers", C.c_int(int(reset)), C.byref(a), C.byref(na), C.byref(b), C.byref(nb))
        return {"step_kernel_ms": a.value, "step_kernel_launches": na.value, "rebuild_ms": b.value, "rebuild_launches": nb.value}

    def set_load_balancing(self, method="dynamic", threshold=0.5, frequency=100):
        self._call("set_load_balancing", C.c_int(LOAD_BALANCE_METHODS[method]), C.c_double(threshold), C.c_int(frequency))

    def set_load_balancing_weights(self, particle_weight=2000.0, cell_weight=1000.0, active_weight_factor=1.0, inactive_weight_factor=1.0):
        self._call("set_load_balancing_weights", C.c_double(particle_weight), C.c_double(cell_weight), C.c_double(active_weight_factor),
                   C.c_double(inactive_weight_factor))

    def get_slab(self):
        """(lo, hi, n_repartitions): the cell layers along the slab axis this context owns now."""
        lo, hi, n = C.c_int32(), C.c_int32(), C.c_uint64()
        self._call("get_slab", C.byref(lo), C.byref(hi), C.byref(n))
        return lo.value, hi.value, n.value

    def comm_init(self, rank, world_size, nccl_id: bytes):
        buf = (C.c_uint8 * NCCL_ID_BYTES).from_buffer_copy(nccl_id)
        self._call("comm_init", C.c_int(rank), C.c_int(world_size), buf)


# LETHE_DEM_B200_LIB selects another build of the same CUDA library (kernel tuning experiments)
CUDA_LIB = os.environ.get("LETHE_DEM_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "liblethe_dem_b200.so")


def load_library(path: str = CUDA_LIB) -> C.CDLL:
    if not os.path.exists(path):
        raise DEMError(
            f"CUDA engine library {path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    return C.CDLL(path)


def load_engine(config: Config, device: int = 0) -> Engine:
    """Create a DEM context on CUDA device `device` (fails loudly without one)."""
    return Engine(load_library(), "lethe_dem_", config, device)


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = (C.c_uint8 * NCCL_ID_BYTES)()
    rc = lib.lethe_dem_nccl_unique_id(buf)
    if rc != 0:
        raise DEMError("lethe_dem_nccl_unique_id failed")
    return bytes(buf)


def balanced_cuts(histogram, cuts, max_shift, min_width=2):
    """lethe_dem_balanced_cuts: the cut planes a load-balance step moves a slab decomposition to
    (host arithmetic inside the CUDA library; no device needed)."""
    lib = load_library()
    hist = np.ascontiguousarray(histogram, dtype=np.uint64)
    cur = np.ascontiguousarray(cuts, dtype=np.int32)
    out = np.empty_like(cur)
    rc = lib.lethe_dem_balanced_cuts(C.c_int32(len(hist)), hist.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int32(len(cur) - 1),
                                     cur.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(max_shift), C.c_int32(min_width),
                                     out.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise DEMError("lethe_dem_balanced_cuts: invalid arguments")
    return out
