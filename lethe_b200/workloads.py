"""Synthetic workloads of BASELINE.json's configs: parameter sets,
wall tables and seeded initial states fed identically to the CUDA engine and to the
CPU oracle.  Host set-up only (the reference does this in insertion / mesh code).

    box_packing   config 1: monodisperse spheres in a box under gravity, HM limit-overlap
    drum          config 2: rotating drum, HM limit-overlap + constant rolling resistance,
                            faceted cylinder wall rotating about x
    periodic_box  config 5: 3-periodic box, Maxwellian velocities, g = 0 (slab-decomposable)
"""
from __future__ import annotations

import math

import numpy as np

from . import abi
from .prm import BoundaryCondition, DEMParameters, Mesh, ParticleType
from .solver import box_wall_faces


def fcc_points(lo, hi, a):
    """FCC lattice with nearest-neighbour distance `a` filling the box [lo, hi)."""
    c = a * math.sqrt(2.0)  # cubic cell edge
    n = [max(1, int(math.floor((hi[d] - lo[d]) / c))) for d in range(3)]
    i, j, k = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing="ij")
    base = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64)
    offs = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    pts = (base[:, None, :] + offs[None, :, :]).reshape(-1, 3) * c
    return pts + np.asarray(lo) + 0.25 * c


def make_props(n, d, density, rng, vel_sigma=0.0, types=None):
    props = np.zeros((n, abi.N_PROPERTIES))
    d = np.broadcast_to(np.asarray(d, dtype=np.float64), (n,))
    props[:, 0] = 0 if types is None else types
    props[:, 1] = d
    props[:, 2] = density * 4.0 / 3.0 * math.pi * (d * 0.5) ** 3
    if vel_sigma > 0:
        props[:, 3:6] = rng.normal(0.0, vel_sigma, (n, 3))
    return props


def cylinder_wall_faces(mesh: Mesh, radius, centre_yz, d_max, boundary_id=4, cap_ids=(0, 1)):
    """Faceted cylinder (axis x): every grid cell the cylinder surface passes through gets
    one plane tangent to the cylinder at the cell's azimuth — the uniform-grid analogue of
    the infinite planes BoundaryCellsInformation extracts from a Q1 `subdivided_cylinder`
    mesh (find_boundary_cells_information.cc:130-219). End caps are the x-faces."""
    nx, ny, nz = mesh.n
    h = mesh.cell_size
    faces = []
    cy, cz = centre_yz
    for k in range(nz):
        for j in range(ny):
            ys = (mesh.lo[1] + j * h[1] - cy, mesh.lo[1] + (j + 1) * h[1] - cy)
            zs = (mesh.lo[2] + k * h[2] - cz, mesh.lo[2] + (k + 1) * h[2] - cz)
            rmax = max(math.hypot(y, z) for y in ys for z in zs)
            ymin = 0.0 if ys[0] <= 0 <= ys[1] else min(abs(ys[0]), abs(ys[1]))
            zmin = 0.0 if zs[0] <= 0 <= zs[1] else min(abs(zs[0]), abs(zs[1]))
            rmin = math.hypot(ymin, zmin)
            if rmax < radius - d_max or rmin > radius:
                continue
            yc, zc = 0.5 * (ys[0] + ys[1]), 0.5 * (zs[0] + zs[1])
            rc = math.hypot(yc, zc)
            ey, ez = yc / rc, zc / rc
            for i in range(nx):
                f = abi.WallFace()
                f.cell = i + nx * (j + ny * k)
                f.boundary_id = boundary_id
                f.global_face_id = f.cell * 8 + 6
                f.normal[:] = [0.0, -ey, -ez]
                f.point[:] = [mesh.lo[0] + (i + 0.5) * h[0], cy + radius * ey, cz + radius * ez]
                faces.append(f)
    for side, i in ((0, 0), (1, nx - 1)):
        for k in range(nz):
            for j in range(ny):
                f = abi.WallFace()
                f.cell = i + nx * (j + ny * k)
                f.boundary_id = cap_ids[side]
                f.global_face_id = f.cell * 8 + side
                f.normal[:] = [1.0 if side == 0 else -1.0, 0.0, 0.0]
                f.point[:] = [mesh.lo[0] if side == 0 else mesh.hi[0], mesh.lo[1] + (j + 0.5) * h[1], mesh.lo[2] + (k + 0.5) * h[2]]
                faces.append(f)
    return faces


class Workload:
    def __init__(self, name, params, ids, x, props, faces, motions=(), description=""):
        self.name, self.params = name, params
        self.ids, self.x, self.props = ids, x, props
        self.faces, self.motions = faces, list(motions)
        self.description = description

    @property
    def n(self):
        return len(self.ids)

    def install(self, engine):
        engine.set_walls(self.faces)
        for m in self.motions:
            engine.set_boundary_motion(*m)
        engine.set_particles(self.ids, self.x, self.props)


def drum(n_target=1_000_000, d=0.003, radius=0.12, fill=0.45, seed=19, spacing=1.0, jitter=0.02):
    """Config 2 (examples/dem/3d-rotating-drum/rotating-drum.prm scaled to n_target):
    rho 2500, Y 1e7, nu 0.2, e 0.97, mu 0.85, rolling = constant mu_r 0.05, dt 1e-5,
    omega_wall = 1.2147 rad/s about x; a jittered FCC bed fills the lower part of the drum."""
    rng = np.random.default_rng(seed)
    a = d * spacing
    # bed: points of the lattice inside the circle (with clearance) and below the fill level
    area = fill * math.pi * radius**2
    per_volume = math.sqrt(2.0) / a**3
    length = n_target / (per_volume * area) * 1.02
    h = 1.5 * d
    ny = nz = int(math.ceil(2 * radius / h)) + 2
    nx = int(math.ceil(length / h))
    lo = (0.0, -0.5 * ny * h, -0.5 * nz * h)
    mesh = Mesh(lo, (nx * h, lo[1] + ny * h, lo[2] + nz * h), (nx, ny, nz), True, "lexicographic")
    length = nx * h
    # level z_top such that the circular segment below it has area `fill`
    zs = np.linspace(-radius, radius, 4001)
    seg = np.array([radius**2 * math.acos(-z / radius) + z * math.sqrt(max(radius**2 - z * z, 0.0)) for z in zs])
    z_top = float(zs[np.searchsorted(seg, area)])
    pts = fcc_points((0.5 * d, -radius, -radius), (length - 0.5 * d, radius, z_top), a)
    r = np.hypot(pts[:, 1], pts[:, 2])
    pts = pts[(r < radius - 0.55 * d) & (pts[:, 2] < z_top) & (pts[:, 0] > 0.55 * d) & (pts[:, 0] < length - 0.55 * d)]
    if len(pts) > n_target:
        pts = pts[np.argsort(pts[:, 0], kind="stable")[:n_target]]
    pts = pts + rng.uniform(-jitter, jitter, pts.shape) * d
    n = len(pts)
    p = DEMParameters()
    p.time_step = 1e-5
    p.pp_model, p.pw_model, p.rolling_model = "hertz_mindlin_limit_overlap", "nonlinear", "constant"
    p.g = (0.0, 0.0, -9.81)
    p.dynamic_contact_search_factor = 0.9
    p.neighborhood_threshold = 1.3
    p.particle_types = [ParticleType(diameter=d, density=2500, young=1e7, poisson=0.2, restitution=0.97, friction=0.85,
                                     rolling_friction=0.05, rolling_viscous_damping=0.1)]
    p.young_wall, p.poisson_wall, p.restitution_wall, p.friction_wall, p.rolling_friction_wall = 1e7, 0.2, 0.85, 0.85, 0.05
    p.mesh = mesh
    p.boundary_conditions = [BoundaryCondition(type="rotational", boundary_id=4, rotational_speed=1.2147,
                                               rotational_vector=(1.0, 0.0, 0.0), point_on_rotational_vector=(0.0, 0.0, 0.0))]
    faces = cylinder_wall_faces(mesh, radius, (0.0, 0.0), d)
    motions = [(4, (0.0, 0.0, 0.0), 1.2147, (1.0, 0.0, 0.0), (0.0, 0.0, 0.0))]
    ids = rng.permutation(n).astype(np.uint32)
    props = make_props(n, d, 2500, rng)
    desc = f"3D rotating drum, {n} spheres d={d * 1e3:g} mm, R={radius} L={length:.3f} m, HM limit-overlap + constant rolling, faceted cylinder wall"
    return Workload("drum", p, ids, pts, props, faces, motions, desc)


def box_packing(n_side=47, nz=None, d=0.005, seed=19, spacing=1.0, jitter=0.02, poly=0.0):
    """Config 1 (applications_tests/lethe-particles/packing_in_box.prm scaled up):
    ~n_side^2*nz spheres settling in a box under gravity, HM limit-overlap, nonlinear walls."""
    rng = np.random.default_rng(seed)
    nz = nz or n_side
    a = d * spacing
    L = (n_side * a, n_side * a, nz * a)
    pts = fcc_points((0.5 * d, 0.5 * d, 0.5 * d), (L[0], L[1], L[2]), a)
    ext = pts.max(axis=0) + 0.8 * d
    h = 2.0 * d
    n = tuple(int(math.ceil(ext[k] / h)) + (2 if k == 2 else 0) for k in range(3))
    mesh = Mesh((0.0, 0.0, 0.0), tuple(n[k] * h for k in range(3)), n, True, "lexicographic")
    pts = pts + rng.uniform(-jitter, jitter, pts.shape) * d
    m = len(pts)
    dd = d * (1.0 - poly * rng.uniform(0, 1, m))
    p = DEMParameters()
    p.time_step = 1e-5
    p.pp_model, p.pw_model, p.rolling_model = "hertz_mindlin_limit_overlap", "nonlinear", "none"
    p.g = (0.0, 0.0, -9.81)
    p.dynamic_contact_search_factor = 0.9
    p.particle_types = [ParticleType(diameter=d, density=1000, young=1e6, poisson=0.3, restitution=0.3, friction=0.1)]
    p.young_wall, p.restitution_wall, p.friction_wall = 1e6, 0.3, 0.1
    p.mesh = mesh
    ids = rng.permutation(m).astype(np.uint32)
    props = make_props(m, dd, 1000, rng)
    return Workload("box_packing", p, ids, pts, props, box_wall_faces(mesh), (), f"box packing, {m} spheres, HM limit-overlap")


def periodic_box(n_cells_side=32, d=0.005, seed=19, spacing=1.0, jitter=0.03, vel_sigma=0.1, cells=(None, None, None)):
    """Config 5: 3-periodic box filled with a jittered FCC lattice (solid fraction ~0.55-0.7),
    Maxwellian velocities, g = 0. `cells` = FCC cubic cells per direction (n^3*4 particles)."""
    rng = np.random.default_rng(seed)
    a = d * spacing
    c = a * math.sqrt(2.0)
    nc = [cells[k] or n_cells_side for k in range(3)]
    i, j, k = np.meshgrid(np.arange(nc[0]), np.arange(nc[1]), np.arange(nc[2]), indexing="ij")
    base = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64)
    offs = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    pts = (base[:, None, :] + offs[None, :, :]).reshape(-1, 3) * c + 0.25 * c
    L = [nc[k] * c for k in range(3)]
    pts = pts + rng.uniform(-jitter, jitter, pts.shape) * d
    n = len(pts)
    # grid: cells of ~1.5 d that tile the box exactly
    ng = tuple(max(3, int(math.floor(L[k] / (1.5 * d)))) for k in range(3))
    mesh = Mesh((0.0, 0.0, 0.0), tuple(L), ng, True, "lexicographic")
    p = DEMParameters()
    p.time_step = 1e-5
    p.pp_model, p.pw_model, p.rolling_model = "hertz_mindlin_limit_overlap", "nonlinear", "none"
    p.g = (0.0, 0.0, 0.0)
    p.dynamic_contact_search_factor = 0.9
    p.particle_types = [ParticleType(diameter=d, density=1000, young=1e6, poisson=0.3, restitution=0.9, friction=0.3)]
    p.mesh = mesh
    p.boundary_conditions = [BoundaryCondition(type="periodic", periodic_id_0=2 * ax, periodic_id_1=2 * ax + 1, periodic_direction=ax) for ax in range(3)]
    ids = rng.permutation(n).astype(np.uint32)
    props = make_props(n, d, 1000, rng, vel_sigma=vel_sigma)
    return Workload("periodic_box", p, ids, pts, props, [], (), f"3-periodic box, {n} spheres, HM limit-overlap, Maxwellian v")
