"""Synthetic workloads of BASELINE.json's configs: parameter sets,
wall tables and seeded initial states fed identically to the CUDA engine and to the
CPU oracle.  Host set-up only (the reference does this in insertion / mesh code).

    box_packing   config 1: monodisperse spheres in a box under gravity, HM limit-overlap
    drum          config 2: rotating drum, HM limit-overlap + constant rolling resistance,
                            faceted cylinder wall rotating about x
    hopper        config 3: wedge hopper, polydisperse spheres held by a floating wall that
                            opens at `gate_open_time`, outlet (particle deletion) below the slot
    cohesive_box  config 4: config 1's geometry with hertz_JKR or DMT cohesion (history heavy)
    periodic_box  config 5: 3-periodic box, Maxwellian velocities, g = 0 (slab-decomposable)
"""
from __future__ import annotations

import math
import os

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
    props[:, 2] = density * 4.0 / 3.0 * math.pi * ((d * 0.5) * (d * 0.5) * (d * 0.5))
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


def plane_wall_faces(mesh: Mesh, point, normal, boundary_id, cell_filter=None, face_no=7):
    """One face row (infinite plane `point`, inward unit `normal`) for every grid cell the
    plane passes through or that lies within one cell diagonal behind it — what
    BoundaryCellsInformation extracts for the boundary cells of an inclined mesh wall
    (find_boundary_cells_information.cc:130-219). `cell_filter(i, j, k)` restricts the cells."""
    nx, ny, nz = mesh.n
    h = mesh.cell_size
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    centre = np.stack([mesh.lo[0] + (i + 0.5) * h[0], mesh.lo[1] + (j + 0.5) * h[1], mesh.lo[2] + (k + 0.5) * h[2]], axis=-1)
    nrm = np.asarray(normal, dtype=np.float64)
    nrm = nrm / np.linalg.norm(nrm)
    dist = (centre - np.asarray(point)) @ nrm
    half_diag = 0.5 * math.sqrt(h[0] ** 2 + h[1] ** 2 + h[2] ** 2)
    sel = (dist < half_diag) & (dist > -half_diag)
    if cell_filter is not None:
        sel &= cell_filter(i, j, k)
    faces = []
    for ci, cj, ck in zip(i[sel].tolist(), j[sel].tolist(), k[sel].tolist()):
        f = abi.WallFace()
        f.cell = ci + nx * (cj + ny * ck)
        f.boundary_id = boundary_id
        f.global_face_id = f.cell * 8 + face_no
        f.normal[:] = [float(v) for v in nrm]
        f.point[:] = [float(v) for v in point]
        faces.append(f)
    return faces


class Workload:
    def __init__(self, name, params, ids, x, props, faces, motions=(), description="", floating_walls=()):
        self.name, self.params = name, params
        self.ids, self.x, self.props = ids, x, props
        self.faces, self.motions = faces, list(motions)
        self.description = description
        self.floating_walls = list(floating_walls)  # (point, normal, t_start, t_end)
        self.solids = []  # (vertices, triangles, translational velocity, angular velocity, centre of rotation)

    @property
    def n(self):
        return len(self.ids)

    def install(self, engine):
        engine.set_walls(self.faces)
        if self.floating_walls:
            engine.set_floating_walls([w[0] for w in self.floating_walls], [w[1] for w in self.floating_walls],
                                      [w[2] for w in self.floating_walls], [w[3] for w in self.floating_walls])
        for m in self.motions:
            engine.set_boundary_motion(*m)
        for sd in self.solids:
            engine.add_solid_surface(*sd)
        engine.set_particles(self.ids, self.x, self.props)


def sheet_mesh(x0, x1, y0, y1, z_of_xy, n):
    """An n x n grid of squares over [x0,x1] x [y0,y1], each split into two triangles, at height
    z_of_xy(x, y): a stand-in for the gmsh surfaces of `subsection solid objects`."""
    gx, gy = np.meshgrid(np.linspace(x0, x1, n + 1), np.linspace(y0, y1, n + 1), indexing="ij")
    vertices = np.stack([gx.ravel(), gy.ravel(), z_of_xy(gx.ravel(), gy.ravel())], axis=1)
    tris = []
    for i in range(n):
        for j in range(n):
            a, b, c, d = i * (n + 1) + j, (i + 1) * (n + 1) + j, (i + 1) * (n + 1) + j + 1, i * (n + 1) + j + 1
            tris += [[a, b, c], [a, c, d]]
    return vertices, np.asarray(tris, dtype=np.uint32)


CELL_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "periodic_cell_64000.npz")


def load_periodic_cell(path=CELL_FILE):
    """The committed disordered 3-periodic unit cell (tools/make_periodic_cell.py): positions in
    units of the sphere diameter and the cell edge L/d, solid fraction 0.64."""
    z = np.load(path)
    return np.ascontiguousarray(z["x_over_d"]), np.asarray(z["L_over_d"], dtype=np.float64), float(z["phi"])


def disordered_points(lo, hi, d, phi=0.60, cell=None):
    """Sphere centres of a disordered packing of solid fraction `phi` filling the box [lo, hi):
    the periodic unit cell, dilated from its own solid fraction to `phi` (which opens a gap between
    neighbours), repeated periodically and cut to the box."""
    xc, Lc, phi_c = cell if cell is not None else load_periodic_cell()
    scale = d * (phi_c / phi) ** (1.0 / 3.0)
    L = Lc * scale
    t0 = [int(math.floor(lo[k] / L[k])) for k in range(3)]
    t1 = [int(math.ceil(hi[k] / L[k])) for k in range(3)]
    out = []
    for i in range(t0[0], t1[0]):
        for j in range(t0[1], t1[1]):
            for k in range(t0[2], t1[2]):
                pts = xc * scale + np.array([i * L[0], j * L[1], k * L[2]])
                keep = np.all((pts >= np.asarray(lo)) & (pts < np.asarray(hi)), axis=1)
                out.append(pts[keep])
    return np.concatenate(out) if out else np.zeros((0, 3))


def drum(n_target=1_000_000, d=0.003, radius=0.12, fill=0.45, seed=19, spacing=1.0, jitter=0.02, bed="lattice"):
    """Config 2 (examples/dem/3d-rotating-drum/rotating-drum.prm scaled to n_target):
    rho 2500, Y 1e7, nu 0.2, e 0.97, mu 0.85, rolling = constant mu_r 0.05, dt 1e-5,
    omega_wall = 1.2147 rad/s about x. bed="disordered": the lower part of the drum is filled
    with a random packing (the periodic unit cell dilated to solid fraction 0.60 and cut to the
    bed; it settles onto the wall in the first steps); bed="lattice": a jittered FCC bed."""
    rng = np.random.default_rng(seed)
    a = d * spacing
    # bed: points of the lattice inside the circle (with clearance) and below the fill level
    area = fill * math.pi * radius**2
    per_volume = math.sqrt(2.0) / a**3
    if bed == "disordered":
        per_volume = 0.60 * 6.0 / (math.pi * d**3)
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
    if bed == "disordered":
        pts = disordered_points((0.5 * d, -radius, -radius), (length - 0.5 * d, radius, z_top), d)
    else:
        pts = fcc_points((0.5 * d, -radius, -radius), (length - 0.5 * d, radius, z_top), a)
    r = np.hypot(pts[:, 1], pts[:, 2])
    pts = pts[(r < radius - 0.55 * d) & (pts[:, 2] < z_top) & (pts[:, 0] > 0.55 * d) & (pts[:, 0] < length - 0.55 * d)]
    if len(pts) > n_target:
        pts = pts[np.argsort(pts[:, 0], kind="stable")[:n_target]]
    if bed != "disordered":
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
    desc = (f"3D rotating drum, {n} spheres d={d * 1e3:g} mm, R={radius} L={length:.3f} m, HM limit-overlap + constant rolling, "
            f"faceted cylinder wall, {'disordered (random-packing) bed' if bed == 'disordered' else 'jittered FCC bed'}")
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


def periodic_box(n_cells_side=32, d=0.005, seed=19, spacing=1.0, jitter=0.03, vel_sigma=0.1, cells=(None, None, None), slab=None):
    """Config 5: 3-periodic box filled with a jittered FCC lattice (solid fraction ~0.55-0.7),
    Maxwellian velocities, g = 0. `cells` = FCC cubic cells per direction (n^3*4 particles).
    The lattice is generated one x-layer of cells at a time from a per-layer seeded stream, so
    `slab=(rank, world)` can build just the layers of one rank's slab (equal-width slabs along
    x, as multi.slab_bounds) with the very same particles the full generation would give it:
    a 64 M-particle job never materialises on one host. ids are the global lattice index."""
    a = d * spacing
    c = a * math.sqrt(2.0)
    nc = [cells[k] or n_cells_side for k in range(3)]
    L = [nc[k] * c for k in range(3)]
    # grid: cells of ~1.5 d that tile the box exactly
    ng = tuple(max(3, int(math.floor(L[k] / (1.5 * d)))) for k in range(3))
    mesh = Mesh((0.0, 0.0, 0.0), tuple(L), ng, True, "lexicographic")
    layers = range(nc[0])
    x_lo, x_hi = -1.0, L[0] + 1.0
    if slab is not None:
        from .multi import slab_bounds

        lo, hi = slab_bounds(ng[0], slab[1])[slab[0]]
        hx = L[0] / ng[0]
        x_lo, x_hi = lo * hx, hi * hx
        margin = jitter * d + 1e-12
        layers = range(max(0, int(math.floor((x_lo - margin) / c)) - 1), min(nc[0], int(math.ceil((x_hi + margin) / c)) + 1))
    j, k = np.meshgrid(np.arange(nc[1]), np.arange(nc[2]), indexing="ij")
    offs = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    per_layer = nc[1] * nc[2] * 4
    xs, vs, idl = [], [], []
    for i in layers:
        rng = np.random.default_rng([seed, i])
        base = np.stack([np.full(j.size, float(i)), j.ravel(), k.ravel()], axis=1)
        pts = (base[:, None, :] + offs[None, :, :]).reshape(-1, 3) * c + 0.25 * c
        pts = pts + rng.uniform(-jitter, jitter, pts.shape) * d
        vel = rng.normal(0.0, vel_sigma, pts.shape) if vel_sigma > 0 else np.zeros_like(pts)
        gid = np.arange(per_layer, dtype=np.int64) + i * per_layer
        if slab is not None:
            # the same ownership rule as multi.owner_mask
            cx = np.floor((pts[:, 0] - mesh.lo[0]) / mesh.cell_size[0]).astype(np.int64)
            keep = (cx >= lo) & (cx < hi)
            pts, vel, gid = pts[keep], vel[keep], gid[keep]
        xs.append(pts)
        vs.append(vel)
        idl.append(gid)
    pts = np.concatenate(xs) if xs else np.zeros((0, 3))
    n = len(pts)
    n_global = per_layer * nc[0]
    p = DEMParameters()
    p.time_step = 1e-5
    p.pp_model, p.pw_model, p.rolling_model = "hertz_mindlin_limit_overlap", "nonlinear", "none"
    p.g = (0.0, 0.0, 0.0)
    p.dynamic_contact_search_factor = 0.9
    p.particle_types = [ParticleType(diameter=d, density=1000, young=1e6, poisson=0.3, restitution=0.9, friction=0.3)]
    p.mesh = mesh
    p.boundary_conditions = [BoundaryCondition(type="periodic", periodic_id_0=2 * ax, periodic_id_1=2 * ax + 1, periodic_direction=ax) for ax in range(3)]
    ids = np.concatenate(idl).astype(np.uint32) if idl else np.zeros(0, np.uint32)
    props = make_props(n, d, 1000, np.random.default_rng(seed))
    if n:
        props[:, 3:6] = np.concatenate(vs)
    w = Workload("periodic_box", p, ids, pts, props, [], (), f"3-periodic box, {n_global} spheres, HM limit-overlap, Maxwellian v")
    w.n_global = n_global
    return w


def _periodic_params(L, d, young=1e6, restitution=1.0, friction=0.3, density=1000.0, dt=1e-5, search_factor=0.9):
    """3-periodic cube of side L (mesh cells just above the neighbourhood radius 1.3 d), g = 0, HM limit-overlap."""
    ng = tuple(max(3, int(math.floor(L[k] / (1.3 * d * (1.0 + 1e-9))))) for k in range(3))
    p = DEMParameters()
    p.time_step = dt
    p.pp_model, p.pw_model, p.rolling_model = "hertz_mindlin_limit_overlap", "nonlinear", "none"
    p.g = (0.0, 0.0, 0.0)
    p.dynamic_contact_search_factor = search_factor
    p.neighborhood_threshold = 1.3
    p.particle_types = [ParticleType(diameter=d, density=density, young=young, poisson=0.3, restitution=restitution, friction=friction)]
    p.mesh = Mesh((0.0, 0.0, 0.0), tuple(L), ng, True, "lexicographic")
    p.boundary_conditions = [BoundaryCondition(type="periodic", periodic_id_0=2 * ax, periodic_id_1=2 * ax + 1, periodic_direction=ax) for ax in range(3)]
    return p


def grow_periodic_cell(make_engine, n_side=100, d=0.002, phi=0.64, seed=19, d0_frac=0.5, eps=0.004, steps_per_stage=40,
                       relax_steps=1500, vel_sigma=0.05, log=None):
    """A DISORDERED 3-periodic packing of n_side^3 monodisperse spheres at solid fraction `phi`.

    There is no gravity to settle a periodic box with, so the packing is made the way random
    packings are made in the literature (Lubachevsky-Stillinger style growth): the spheres start
    as a dilute gas (diameter d0_frac*d on a simple-cubic lattice with +-0.2 d random offsets and
    thermal velocities) and are grown geometrically, `eps` per stage, while the engine itself
    integrates their collisions with a dissipative, frictionless material (restitution 0.5, mu 0);
    `relax_steps` steps at the final size let the last overlaps relax. Only the public per-step
    plugin call is used (lethe_dem_step_host: rows with the new diameter up, steps, rows down).
    make_engine(config) -> engine (abi.load_engine on a GPU; the oracle in CPU tests of small cells).
    Returns (x[n,3] wrapped into the cell, L[3])."""
    rng = np.random.default_rng(seed)
    a = d * (math.pi / (6.0 * phi)) ** (1.0 / 3.0)
    L = [n_side * a] * 3
    p = _periodic_params(L, d, restitution=0.5, friction=0.0)
    eng = make_engine(p.to_config())
    i, j, k = np.meshgrid(np.arange(n_side), np.arange(n_side), np.arange(n_side), indexing="ij")
    x = (np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1) + 0.5) * a
    x = x + rng.uniform(-0.2, 0.2, x.shape) * d
    n = len(x)
    dc = d0_frac * d
    props = make_props(n, d, 1000, rng, vel_sigma=vel_sigma)  # the mass of the final sphere throughout
    props[:, 1] = dc
    ids = np.arange(n, dtype=np.uint32)
    x = np.ascontiguousarray(x)
    eng.set_particles(ids, x, props)
    stage = 0
    while dc < d:
        dc = min(d, dc * (1.0 + eps))
        props[:, 1] = dc
        eng.step_host(steps_per_stage, ids, x, props)
        stage += 1
        if log and stage % 20 == 0:
            st = eng.get_stats()
            log(f"grow stage {stage}: d/d_final {dc / d:.4f} pairs/particle {st.n_pair_entries / n:.2f} v_max {st.v_max:.3g} rebuilds {st.n_rebuilds}")
    if relax_steps:
        eng.step_host(relax_steps, ids, x, props)
    if log:
        st = eng.get_stats()
        log(f"grown: {n} spheres phi {phi} pairs/particle {st.n_pair_entries / n:.2f} v_max {st.v_max:.3g} rebuilds {st.n_rebuilds}")
    eng.close()
    Lc = np.asarray(L)
    return np.mod(x, Lc), L


def periodic_packing(cell_x, L_cell, reps=(1, 1, 1), d=0.002, seed=19, vel_sigma=0.1, slab=None, restitution=1.0, friction=0.0):
    """Config 5: the disordered cell of grow_periodic_cell tiled reps[0] x reps[1] x reps[2] times
    (a periodic tiling of a periodic packing is a packing of the larger box), every sphere with
    its own Maxwellian velocity (sigma `vel_sigma` m/s per component, seeded per tile) so that the
    copies diverge at once. Material of multiperiodic_collisions_3d.prm (Y 1e6, nu 0.3, restitution
    1, friction 0, g = 0): the thermal motion does not decay, so the list-rebuild rate is steady
    (with friction the bed is cold, and rebuild-free, after a few hundred steps). `slab=(rank, world)` keeps the
    particles of one rank's equal-width slab along x (the ownership rule of multi.owner_mask)."""
    nc = len(cell_x)
    L = [L_cell[k] * reps[k] for k in range(3)]
    p = _periodic_params(L, d, restitution=restitution, friction=friction)
    mesh = p.mesh
    lo = hi = None
    if slab is not None:
        from .multi import slab_bounds

        lo, hi = slab_bounds(mesh.n[0], slab[1])[slab[0]]
    xs, vs, idl = [], [], []
    for ti in range(reps[0]):
        if slab is not None:
            # tiles that cannot reach the slab are skipped without being generated
            hx = mesh.cell_size[0]
            if (ti + 1) * L_cell[0] < lo * hx - 1e-9 or ti * L_cell[0] > hi * hx + 1e-9:
                continue
        for tj in range(reps[1]):
            for tk in range(reps[2]):
                tile = (ti * reps[1] + tj) * reps[2] + tk
                rng = np.random.default_rng([seed, tile])
                pts = cell_x + np.array([ti * L_cell[0], tj * L_cell[1], tk * L_cell[2]])
                vel = rng.normal(0.0, vel_sigma, pts.shape)
                gid = np.arange(nc, dtype=np.int64) + tile * nc
                if slab is not None:
                    cx = np.floor((pts[:, 0] - mesh.lo[0]) / mesh.cell_size[0]).astype(np.int64)
                    keep = (cx >= lo) & (cx < hi)
                    pts, vel, gid = pts[keep], vel[keep], gid[keep]
                xs.append(pts)
                vs.append(vel)
                idl.append(gid)
    pts = np.ascontiguousarray(np.concatenate(xs)) if xs else np.zeros((0, 3))
    n = len(pts)
    n_global = nc * reps[0] * reps[1] * reps[2]
    props = make_props(n, d, 1000, np.random.default_rng(seed))
    if n:
        props[:, 3:6] = np.concatenate(vs)
    ids = np.concatenate(idl).astype(np.uint32) if idl else np.zeros(0, np.uint32)
    desc = (f"3-periodic box, {n_global} spheres d={d * 1e3:g} mm, disordered packing (grown cell of {nc} tiled "
            f"{reps[0]}x{reps[1]}x{reps[2]}), HM limit-overlap, Maxwellian v sigma={vel_sigma:g} m/s, g=0")
    w = Workload("periodic_packing", p, ids, pts, props, [], (), desc)
    w.n_global = n_global
    return w


def hopper(n_target=4_000_000, d=0.00224, sigma=0.1, slot=12.0, seed=19, spacing=1.0, jitter=0.02, gate_open_time=0.02,
           chute_cells=3, min_nx=16):
    """Config 3 (examples/dem/3d-rectangular-hopper/hopper.prm scaled): a wedge hopper whose two
    45-degree walls converge to a slot of width `slot`*d at z = 0; polydisperse spheres (normal
    distribution about d with relative std `sigma`, truncated at +-2.5 sigma as
    distributions.cc:169-189) rest on a floating wall closing the slot until `gate_open_time`
    (subsection floating walls), then discharge through a short chute whose bottom boundary is
    an outlet: particles that pass it are deleted at the next rebuild. Rolling = constant
    mu_r 0.1786 as the example. The hopper is long in x (the slab axis of multi-GPU runs)."""
    rng = np.random.default_rng(seed)
    dmax = d * (1.0 + 2.5 * sigma)
    a = dmax * spacing
    w = slot * d
    # wedge cross-section area up to height H: w H + H^2; pick H ~ 40 d_max and get the length from n_target
    H = 40.0 * dmax
    area = w * H + H * H
    per_volume = math.sqrt(2.0) / a**3
    length = n_target / (per_volume * area * 0.93)
    h = 1.35 * dmax
    if length < min_nx * h:  # small cases: keep the hopper long enough to be cut into slabs, make it lower
        length = min_nx * h
        area = n_target / (per_volume * length * 0.93)
        H = max(6.0 * dmax, 0.5 * (-w + math.sqrt(w * w + 4.0 * area)))
    half_w = 0.5 * w + H + 2 * h
    ny = int(math.ceil(2 * half_w / h))
    nzu = int(math.ceil((H + 2 * dmax) / h))
    nx = max(4, int(math.ceil(length / h)))
    lo = (0.0, -0.5 * ny * h, -chute_cells * h)
    mesh = Mesh(lo, (nx * h, lo[1] + ny * h, nzu * h), (nx, ny, chute_cells + nzu), True, "lexicographic")
    length = nx * h
    pts = fcc_points((0.5 * dmax, -half_w, 0.55 * dmax), (length - 0.5 * dmax, half_w, H), a)
    inside = (np.abs(pts[:, 1]) < 0.5 * w + pts[:, 2] - 0.8 * dmax) & (pts[:, 0] > 0.55 * dmax) & (pts[:, 0] < length - 0.55 * dmax)
    pts = pts[inside]
    if len(pts) > n_target:
        pts = pts[np.argsort(pts[:, 0], kind="stable")[:n_target]]
    pts = pts + rng.uniform(-jitter, jitter, pts.shape) * d
    n = len(pts)
    dd = np.clip(rng.normal(d, sigma * d, n), d * (1 - 2.5 * sigma), dmax)
    p = DEMParameters()
    p.time_step = 1e-5
    p.pp_model, p.pw_model, p.rolling_model = "hertz_mindlin_limit_overlap", "nonlinear", "constant"
    p.g = (0.0, 0.0, -9.81)
    p.dynamic_contact_search_factor = 0.9
    p.neighborhood_threshold = 1.3
    p.particle_types = [ParticleType(diameter=dmax, density=600, young=5e6, poisson=0.5, restitution=0.7, friction=0.5,
                                     rolling_friction=0.1786, rolling_viscous_damping=0.1)]
    p.young_wall, p.poisson_wall, p.restitution_wall, p.friction_wall, p.rolling_friction_wall = 5e6, 0.5, 0.7, 0.5, 0.1786
    p.mesh = mesh
    p.boundary_conditions = [BoundaryCondition(type="outlet", boundary_id=4)]  # z-min face of the chute
    p.floating_walls = [((0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 0.0, gate_open_time)]
    faces = box_wall_faces(mesh, p.outlet_boundaries)
    above = lambda i, j, k: k >= chute_cells  # noqa: E731  the inclined walls end at the slot lip
    s2 = math.sqrt(0.5)
    faces += plane_wall_faces(mesh, (0.0, -0.5 * w, 0.0), (0.0, s2, s2), 10, above, face_no=6)
    faces += plane_wall_faces(mesh, (0.0, 0.5 * w, 0.0), (0.0, -s2, s2), 11, above, face_no=7)
    ids = rng.permutation(n).astype(np.uint32)
    props = make_props(n, dd, 600, rng)
    desc = (f"wedge hopper discharge, {n} polydisperse spheres d={d * 1e3:g} mm +-{sigma:.0%}, HM limit-overlap + constant rolling, "
            f"floating wall opens at t={gate_open_time:g} s, outlet deletion")
    return Workload("hopper", p, ids, pts, props, faces, (), desc, p.floating_walls)


def cohesive_box(n_side=47, d=0.001, model="hertz_JKR", seed=19, spacing=0.995, jitter=0.01, surface_energy=None):
    """Config 4: a dense box packing with cohesion — `hertz_JKR` (surface energy 0.5 J/m^2, the
    range of pp_jkr_equilibrium.prm) or `DMT` (Hamaker 4e-19, cut-off 0.1 as
    pp_dmt_equilibrium.prm) — slightly compressed so that the coordination number is high and
    nearly every list entry carries history."""
    w = box_packing(n_side, d=d, seed=seed, spacing=spacing, jitter=jitter)
    p = w.params
    p.pp_model = model
    p.pw_model = "JKR" if model == "hertz_JKR" else "DMT"
    gamma = surface_energy if surface_energy is not None else (0.5 if model == "hertz_JKR" else 1e-4)
    t = p.particle_types[0]
    t.surface_energy, t.hamaker = gamma, 4e-19
    t.young, t.restitution, t.friction = 1e6, 0.5, 0.3
    p.surface_energy_wall, p.hamaker_wall = gamma, 4e-19
    p.dmt_cut_off_threshold = 0.1
    w.name = "cohesive_box"
    w.description = f"cohesive box packing, {w.n} spheres d={d * 1e3:g} mm, {model} (gamma={gamma:g} J/m^2), history heavy"
    return w
