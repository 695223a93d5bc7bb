"""Shared helpers for the parity tests: parameter sets of the reference's unit
tests and synthetic packings fed identically to the oracle and the CUDA engine."""
import json
import math
import os

import numpy as np

from lethe_b200 import abi
from lethe_b200.prm import DEMParameters, Mesh, ParticleType

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name="unit_goldens.json"):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def sig6(x):
    """Round to deallog's 6 significant digits."""
    return float(f"{x:.6g}")


def assert_sig6(value, gold, what=""):
    # deallog prints 6 significant digits; allow one unit in the last printed digit
    if gold == 0.0:
        assert abs(value) < 5e-7, (what, value, gold)
        return
    tol = 10 ** (math.floor(math.log10(abs(gold))) - 5) * 1.0
    assert abs(value - gold) <= tol, (what, value, gold)


def unit_test_parameters(pp_model="hertz_mindlin_limit_overlap", pw_model="nonlinear", rolling="constant", dt=1e-5, d=0.005,
                         young=5e7, restitution=0.5, friction=0.5, rolling_friction=0.1, rolling_viscous=0.5, g=(0, 0, 0),
                         wall_rolling_viscous=0.1):
    """set_default_dem_parameters + the overrides of tests/dem/*.cc: hyper_cube(-1,1) refined twice."""
    p = DEMParameters()
    p.time_step = dt
    p.pp_model, p.pw_model, p.rolling_model = pp_model, pw_model, rolling
    p.g = g
    t = ParticleType(diameter=d, young=young, poisson=0.3, restitution=restitution, friction=friction,
                     rolling_friction=rolling_friction, rolling_viscous_damping=rolling_viscous, surface_energy=0.0, hamaker=0.0,
                     density=2500)
    p.particle_types = [t]
    p.young_wall, p.poisson_wall, p.restitution_wall, p.friction_wall = young, 0.3, restitution, friction
    p.rolling_friction_wall, p.rolling_viscous_damping_wall = rolling_friction, wall_rolling_viscous
    p.mesh = Mesh((-1.0,) * 3, (1.0,) * 3, (4, 4, 4), True, "morton")
    return p


def props_row(ptype, d, mass, v=(0, 0, 0), w=(0, 0, 0)):
    return [ptype, d, mass, *v, *w]


def random_packing(n_side, d=0.005, spacing=1.02, jitter=0.05, seed=19, poly=0.0, vel=0.05, density=1000.0, nz=None, n_types=1):
    """Jittered lattice of ~n_side^3 spheres with overlaps and random velocities:
    a dense synthetic state that exercises touching + near pairs immediately."""
    rng = np.random.default_rng(seed)
    nz = nz or n_side
    ii, jj, kk = np.meshgrid(np.arange(n_side), np.arange(n_side), np.arange(nz), indexing="ij")
    x = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(np.float64)
    x = (x + 0.5) * (d * spacing) + rng.uniform(-jitter, jitter, x.shape) * d
    n = len(x)
    diam = d * (1.0 - poly * rng.uniform(0, 1, n))
    props = np.zeros((n, 9))
    props[:, 0] = rng.integers(0, n_types, n)
    props[:, 1] = diam
    props[:, 2] = density * 4.0 / 3.0 * math.pi * (diam * 0.5) ** 3
    props[:, 3:6] = rng.normal(0, vel, (n, 3))
    props[:, 6:9] = rng.normal(0, vel / d * 0.2, (n, 3))
    ids = rng.permutation(n).astype(np.uint32)
    extent = np.array([n_side, n_side, nz]) * d * spacing
    return ids, x, props, extent


def packing_parameters(extent, d=0.005, cell=None, pp_model="hertz_mindlin_limit_overlap", pw_model="nonlinear", rolling="none",
                       dt=1e-5, g=(0, 0, -9.81), n_types=1, periodic=(0, 0, 0), surface_energy=0.0, hamaker=4e-19, young=1e6,
                       cell_order="lexicographic", search_factor=0.05):
    p = DEMParameters()
    p.time_step = dt
    p.pp_model, p.pw_model, p.rolling_model = pp_model, pw_model, rolling
    p.g = g
    p.dynamic_contact_search_factor = search_factor  # small: several list rebuilds within a short test
    p.particle_types = []
    for t in range(n_types):
        p.particle_types.append(ParticleType(diameter=d, young=young * (1 + t), poisson=0.3 - 0.05 * t, restitution=0.3 + 0.2 * t,
                                             friction=0.1 + 0.2 * t, rolling_friction=0.1 + 0.05 * t, rolling_viscous_damping=0.1 + 0.1 * t,
                                             surface_energy=surface_energy, hamaker=hamaker))
    p.young_wall, p.restitution_wall, p.friction_wall = young, 0.3, 0.1
    p.surface_energy_wall, p.hamaker_wall = surface_energy, hamaker
    cell = cell or 2 * d
    n = tuple(max(3, int(math.ceil(e / cell))) for e in extent)
    p.mesh = Mesh((0.0, 0.0, 0.0), tuple(n[i] * cell for i in range(3)), n, True, cell_order)
    from lethe_b200.prm import BoundaryCondition
    for ax in range(3):
        if periodic[ax]:
            p.boundary_conditions.append(BoundaryCondition(type="periodic", periodic_id_0=2 * ax, periodic_id_1=2 * ax + 1, periodic_direction=ax))
    return p


def solid_surface_case(name):
    """Parameters, particle row and (rotated) solid mesh of one
    applications_tests/lethe-particles/particle_solid_surface_<name> case
    (tests/golden/solid_surface_goldens.json, made by make_solid_goldens.py)."""
    with open(os.path.join(GOLDEN, "solid_surface_goldens.json")) as f:
        c = json.load(f)[name]
    p = DEMParameters()
    p.time_step = c["dt"]
    p.pp_model, p.pw_model, p.rolling_model = c["pp_model"], c["pw_model"], "constant"
    p.g = tuple(c["g"])
    p.neighborhood_threshold = c["neighborhood_threshold"]
    p.dynamic_contact_search_factor = c["search_factor"]
    p.particle_types = [ParticleType(diameter=c["diameter"], density=c["density"], young=c["young"], poisson=c["poisson"],
                                     restitution=c["restitution"], friction=c["friction"])]
    p.young_wall, p.poisson_wall = c["young_wall"], c["poisson_wall"]
    p.restitution_wall, p.friction_wall = c["restitution_wall"], c["friction_wall"]
    n = 2 ** c["refinement"]
    p.mesh = Mesh((c["box"][0],) * 3, (c["box"][1],) * 3, (n, n, n), False, "morton")
    d = c["diameter"]
    mass = c["density"] * 4.0 / 3.0 * math.pi * (d * 0.5) ** 3
    row = [0, d, mass, 0, 0, 0, 0, 0, 0]
    # GridTools::rotate(axis, angle): Rodrigues' rotation matrix applied to every vertex
    v = np.array(c["vertices"], dtype=np.float64)
    a = np.array(c["rotation_axis"], dtype=np.float64)
    a = a / np.linalg.norm(a)
    th = c["rotation_angle"]
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = math.cos(th) * np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * np.outer(a, a)
    v = v @ R.T
    return c, p, np.array([c["position"]]), np.array([row]), v, np.array(c["triangles"], dtype=np.uint32)


def two_processor_contact_case():
    """tests/dem/particle_particle_contact_on_two_processors.cc: two spheres (m = 1, MOI = 1) collide
    head-on across y = 0, the boundary between the two ranks' halves of a 4 x 4 mesh; search and
    `integrate` (no opening half step) every step. 3-D stand-in of the 2-D test: same physics in
    the z = 0 plane. Returns (parameters, config kwargs, ids, x, props)."""
    p = unit_test_parameters(young=5e7, restitution=0.9, friction=0.5, rolling_friction=0.1, rolling_viscous=0.5, dt=1e-5, g=(0, 0, 0))
    p.particle_types[0].poisson = 0.9
    p.contact_detection_method = "constant"
    p.contact_detection_frequency = 1
    p.restart = True  # the test calls integrate() from the first step
    ids = np.array([0, 1], dtype=np.uint32)
    x = np.array([[0.0, 0.003, 0.0], [0.0, -0.003, 0.0]])
    props = np.array([props_row(0, 0.005, 1.0, v=(0, -0.5, 0)), props_row(0, 0.005, 1.0, v=(0, 0.5, 0))], dtype=np.float64)
    return p, dict(moi_override=1.0), ids, x, props
