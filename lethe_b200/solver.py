"""Host-side mirror of the reference's `DEMSolver` for the hot path
(source/dem/dem.cc:1064-1267): parameter set-up, wall table, insertion, the time
loop and the `test` output.  Everything per-step is delegated to the C ABI
(`lethe_dem_step`); this module only decides *when* to call it, exactly like
`DEMSolver::solve` around `execute_contact_detection_and_search` /
`compute_contact_forces` / `integrate`.

The insertion and output paths are host code in the reference too (north star:
"Host code remains C++ (… the insertion and output paths)"); they are restated
here in Python only so that the reference's application tests can be replayed.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np

from . import abi
from .prm import DEMParameters, Mesh

# boundary ids of a colorized hyper_cube / hyper_rectangle: 2*axis + side
_FACE_NORMALS = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]


def box_wall_faces(mesh: Mesh, outlet_boundaries=(), periodic=(0, 0, 0)):
    """Rows of BoundaryCellsInformation::find_boundary_cells_information
    (source/dem/find_boundary_cells_information.cc:130-219) for a uniform box
    mesh: one row per (boundary cell, boundary face) that is neither an outlet
    nor periodic; inward normal, face centre as point."""
    nx, ny, nz = mesh.n
    h = mesh.cell_size
    faces = []
    n = (nx, ny, nz)
    fid = 0
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                idx = (i, j, k)
                cell = i + nx * (j + ny * k)
                for axis in range(3):
                    for side in (0, 1):
                        if idx[axis] != (0 if side == 0 else n[axis] - 1):
                            continue
                        face_no = 2 * axis + side
                        bid = face_no if mesh.colorize else 0
                        unique = cell * 6 + face_no
                        if periodic[axis] or bid in outlet_boundaries:
                            continue
                        f = abi.WallFace()
                        f.cell = cell
                        f.boundary_id = bid
                        f.global_face_id = unique
                        nrm = _FACE_NORMALS[face_no]
                        f.normal[:] = [float(v) for v in nrm]
                        centre = [mesh.lo[d] + (idx[d] + 0.5) * h[d] for d in range(3)]
                        centre[axis] = mesh.lo[axis] if side == 0 else mesh.hi[axis]
                        f.point[:] = centre
                        faces.append(f)
                        fid += 1
    return faces


_libc = None


def _glibc_rand_container(n: int, maximum_range: float, seed: int):
    """create_random_number_container (include/core/utilities.h:1061-1073): one
    srand(seed*(i+1)) + rand() per element — glibc's generator."""
    global _libc
    if _libc is None:
        _libc = ctypes.CDLL("libc.so.6")
        _libc.rand.restype = ctypes.c_int
        _libc.srand.argtypes = [ctypes.c_uint]
    RAND_MAX = 2147483647
    out = np.empty(n)
    for i in range(n):
        _libc.srand(ctypes.c_uint((seed * (i + 1)) & 0xFFFFFFFF))
        out[i] = (float(_libc.rand()) / float(RAND_MAX)) * maximum_range
    return out


def volume_insertion(p: DEMParameters, n_insert: int, first_id: int = 0, particle_type: int = 0, distribution=None, n_ranks: int = 1):
    """InsertionVolume::insert / find_insertion_location
    (source/dem/insertion_volume.cc:43-206) + assign_particle_properties
    (source/dem/insertion.cc:60-121) on one rank. `distribution` is the particle type's size
    distribution object (its generator state carries over from one insertion to the next).
    `n_ranks` > 1 restates how the reference run on that many MPI processes pairs sites with random
    offsets (each process takes a contiguous share of the lattice and indexes its OWN random vector,
    insertion_volume.cc:256-283): only needed to reproduce its multi-rank goldens."""
    ins = p.insertion
    d_max = p.d_max
    t = p.particle_types[particle_type]
    if distribution is None:
        from .distributions import make_distribution

        distribution = make_distribution(t)
    n_dir = [0, 0, 0]
    for axis in ins.direction_sequence:
        n_dir[axis] = int((ins.box_point_2[axis] - ins.box_point_1[axis]) / (ins.distance_threshold * d_max))
    n_sites = n_dir[0] * n_dir[1] * n_dir[2]
    a0, a1, a2 = ins.direction_sequence

    def location(site, r1, r2):
        i0 = site % n_dir[a0]
        i1 = (site % (n_dir[a0] * n_dir[a1])) // n_dir[a0]
        i2 = site // (n_dir[a0] * n_dir[a1])
        out = [0.0, 0.0, 0.0]
        out[a0] = ins.box_point_1[a0] + ((i0 + 0.5) * ins.distance_threshold - r1) * d_max
        out[a1] = ins.box_point_1[a1] + ((i1 + 0.5) * ins.distance_threshold - r2) * d_max
        out[a2] = ins.box_point_1[a2] + ((i2 + 0.5) * ins.distance_threshold - r1) * d_max
        return out

    # set_filtered_index (insertion_volume.cc:207-351): the sites whose un-jittered location the
    # acceptance function accepts; the random offsets are drawn for the accepted sites only
    if ins.acceptance_function:
        from .prm import evaluate_function

        sites = [k for k in range(n_sites) if evaluate_function(ins.acceptance_function, 0.0, location(k, 0.0, 0.0)) > 0.0]
    else:
        sites = range(n_sites)
    n_valid = len(sites)
    n_insert = min(n_insert, n_valid)
    x = np.empty((n_insert, 3))
    share = n_sites // n_ranks
    k = 0
    for rank in range(n_ranks):
        first = n_sites - (n_sites - (n_ranks - 1) * share) if rank == n_ranks - 1 else rank * share
        last = n_sites if rank == n_ranks - 1 else (rank + 1) * share
        mine = [site for site in sites if first <= site < last]
        rnd = _glibc_rand_container(len(mine), ins.maximum_offset, ins.seed)
        for counter, site in enumerate(mine):
            if k >= n_insert:
                break
            x[k] = location(site, rnd[counter], rnd[len(mine) - counter - 1])
            k += 1
    props = np.zeros((n_insert, abi.N_PROPERTIES))
    d = np.abs(distribution.sample(n_insert))  # particle_size_sampling (insertion.cc:78-90)
    h = d * 0.5
    props[:, 0] = particle_type
    props[:, 1] = d
    props[:, 2] = t.density * 4.0 / 3.0 * math.pi * (h * h * h)
    props[:, 3:6] = ins.initial_velocity
    props[:, 6:9] = ins.initial_omega
    ids = np.arange(first_id, first_id + n_insert, dtype=np.uint32)
    return ids, x, props


def active_cell_order(mesh):
    """Lexicographic cell (i, j, k) triples of the uniform grid in deal.II's active-cell order:
    lexicographic for an unrefined subdivided grid, the z-order curve after refine_global."""
    nx, ny, nz = mesh.n
    cells = [(i, j, k) for k in range(nz) for j in range(ny) for i in range(nx)]
    if mesh.cell_order != "morton":
        return cells

    def key(c):
        out = 0
        for b in range(max(nx, ny, nz).bit_length()):
            out |= ((c[0] >> b) & 1) << (3 * b) | ((c[1] >> b) & 1) << (3 * b + 1) | ((c[2] >> b) & 1) << (3 * b + 2)
        return out

    return sorted(cells, key=key)


class PlaneInsertion:
    """InsertionPlane (source/dem/insertion_plane.cc): at every insertion iteration one particle
    in each cell cut by the plane that holds no particle, at the cell centre plus rand() * maximum
    offset / RAND_MAX per axis — glibc's rand() stream, never seeded by this method."""

    def __init__(self, p: DEMParameters):
        mesh, ins = p.mesh, p.insertion
        h = mesh.cell_size
        point, normal = np.asarray(ins.plane_point), np.asarray(ins.plane_normal)
        self.cells = []  # find_inplane_cells (:43-88), in std::set order = active-cell order
        for c in active_cell_order(mesh):
            lo = np.array([mesh.lo[d] + c[d] * h[d] for d in range(3)])
            ref = None
            for v in range(8):  # deal.II vertex order: x fastest
                vertex = lo + np.array([(v & 1) * h[0], ((v >> 1) & 1) * h[1], ((v >> 2) & 1) * h[2]])
                dist = float(np.dot(vertex - point, normal))
                if ref is None:
                    ref = dist
                elif ref * dist <= 0:
                    self.cells.append(c)
                    break
        self.centers = {c: tuple(mesh.lo[d] + (c[d] + 0.5) * h[d] for d in range(3)) for c in self.cells}
        self.maximum_range_for_randomness = ins.maximum_offset / float(2147483647)
        _glibc_rand_container(0, 0.0, 0)  # loads libc
        _libc.srand(1)  # the state of a process that never called srand

    def insert(self, p, occupied_cells, remaining, first_id, particle_type, distribution):
        """-> ids, x, props of this iteration's particles. `occupied_cells`: (i, j, k) of the cells
        particles are registered in (as of the last sort)."""
        empty = [c for c in self.cells if c not in occupied_cells]
        n_insert = min(len(empty), remaining)
        empty = empty[len(empty) - n_insert:]  # surplus cells are dropped from the front (:216-222)
        x = np.empty((n_insert, 3))
        for k, c in enumerate(empty):
            for d in range(3):
                x[k, d] = self.centers[c][d] + float(_libc.rand()) * self.maximum_range_for_randomness
        t = p.particle_types[particle_type]
        props = np.zeros((n_insert, abi.N_PROPERTIES))
        dp = np.abs(distribution.sample(n_insert))
        half = dp * 0.5
        props[:, 0] = particle_type
        props[:, 1] = dp
        props[:, 2] = t.density * 4.0 / 3.0 * math.pi * (half * half * half)
        props[:, 3:6] = p.insertion.initial_velocity
        props[:, 6:9] = p.insertion.initial_omega
        return np.arange(first_id, first_id + n_insert, dtype=np.uint32), x, props


def list_insertion(p: DEMParameters, first_id: int = 0, particle_type: int = 0):
    """InsertionList::insert (source/dem/insertion_list.cc): the listed positions, velocities and
    diameters, all at the first insertion step."""
    ins = p.insertion
    n = len(ins.list_x)
    t = p.particle_types[particle_type]
    x = np.stack([np.asarray(ins.list_x), np.asarray(ins.list_y), np.asarray(ins.list_z)], axis=1).astype(np.float64)
    d = np.asarray(ins.list_diameters if len(ins.list_diameters) == n else [t.diameter] * n, dtype=np.float64)
    props = np.zeros((n, abi.N_PROPERTIES))
    props[:, 0] = particle_type
    props[:, 1] = d
    props[:, 2] = t.density * 4.0 / 3.0 * math.pi * ((d * 0.5) * (d * 0.5) * (d * 0.5))
    if n:
        props[:, 3:6] = np.asarray(ins.list_velocity)
        props[:, 6:9] = np.asarray(ins.list_omega)
    return np.arange(first_id, first_id + n, dtype=np.uint32), x, props


def file_insertion(p: DEMParameters, path: str, n_max: int, first_id: int = 0, particle_type: int = 0):
    """InsertionFile::insert (source/dem/insertion_file.cc:27-130): one `;`-separated table
    `p_x; p_y; p_z; v_x; v_y; v_z; w_x; w_y; w_z; diameters;` per insertion, at most `n_max` rows."""
    with open(path) as f:
        lines = [ln for ln in f.read().splitlines() if ln.strip()]
    header = [h.strip() for h in lines[0].split(";") if h.strip()]
    rows = [[float(v) for v in ln.split(";") if v.strip()] for ln in lines[1:]]
    data = {h: np.array([r[k] for r in rows]) for k, h in enumerate(header)}
    n = min(n_max, len(rows))
    t = p.particle_types[particle_type]
    x = np.stack([data["p_x"], data["p_y"], data["p_z"]], axis=1)[:n]
    d = data["diameters"][:n]
    props = np.zeros((n, abi.N_PROPERTIES))
    props[:, 0] = particle_type
    props[:, 1] = d
    props[:, 2] = t.density * 4.0 / 3.0 * math.pi * ((d * 0.5) * (d * 0.5) * (d * 0.5))
    for k, key in enumerate(("v_x", "v_y", "v_z", "w_x", "w_y", "w_z")):
        props[:, 3 + k] = data[key][:n]
    return np.arange(first_id, first_id + n, dtype=np.uint32), x, props


class DEMSolver:
    """`DEMSolver<3, DEMProperties>` with the hot path behind the C ABI."""

    def __init__(self, parameters: DEMParameters, engine_factory=None, device: int = 0, store_forces=False, moi_override=0.0,
                 prm_directory=".", reference_insertion_ranks=1):
        self.parameters = parameters
        self.prm_directory = prm_directory  # mesh file names of the .prm are relative to it
        self.reference_insertion_ranks = reference_insertion_ranks  # see volume_insertion(n_ranks=)
        self.config = parameters.to_config(store_forces=store_forces, moi_override=moi_override)
        factory = engine_factory or (lambda cfg: abi.load_engine(cfg, device))
        self.engine = factory(self.config)
        self.iteration_number = 0
        self.current_time = 0.0
        self._remaining = [t.number for t in parameters.particle_types]
        self._current_type = 0
        self._next_id = 0
        self._file_id = 0
        self._solid_motion = []
        from .distributions import make_distribution

        self._distributions = [make_distribution(t) for t in parameters.particle_types]  # setup_distributions, rank 0
        # plane insertion asks which cells hold particles, i.e. where the last sort registered them:
        # the positions that sort saw are kept (one engine call per iteration while it is active)
        self._plane = PlaneInsertion(parameters) if parameters.insertion.method == "plane" else None
        self._track_registration = self._plane is not None or parameters.insertion.remove_particles
        self._registered_cells = set()
        self._registered = {}  # particle id -> (i, j, k) of the cell the last sort registered it in
        self._setup_boundaries()

    # DEMSolver::setup_functions_and_pointers / boundary_cell_object.build
    def _setup_boundaries(self):
        p = self.parameters
        if p.mesh.expand_particle_wall_contact_search:
            raise abi.DEMError("`expand particle-wall contact search` is not on the B200 path (box meshes do not need it)")
        self.engine.set_walls(box_wall_faces(p.mesh, p.outlet_boundaries, p.periodic))
        if p.floating_walls:
            pts, nrm, t0, t1 = zip(*p.floating_walls)
            self.engine.set_floating_walls(pts, nrm, t0, t1)
        # DEMSolver::setup_solid_objects (dem.cc:164-191) + SerialSolid::setup_triangulation
        # (serial_solid.cc:163-216: read, rotate, translate)
        for so in p.solid_surfaces:
            from .mesh_io import dealii_simplex_surface, read_msh_triangles

            if so.mesh_type == "dealii":
                vertices, triangles = dealii_simplex_surface(so.grid_type, so.grid_arguments, so.initial_refinement)
            else:
                path = so.mesh_file if os.path.isabs(so.mesh_file) else os.path.join(self.prm_directory, so.mesh_file)
                vertices, triangles = read_msh_triangles(path)
            a = np.asarray(so.rotation_axis, dtype=np.float64)
            a = a / np.linalg.norm(a)
            th = so.rotation_angle
            K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
            rot = math.cos(th) * np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * np.outer(a, a)
            vertices = vertices @ rot.T + np.asarray(so.translation)
            from .prm import solid_velocity_at

            tv, av = solid_velocity_at(so, 0.0)
            self._solid_motion.append((tv, av))
            self.engine.add_solid_surface(vertices, triangles, tv, av, so.center_of_rotation)
        for bc in p.boundary_conditions:
            if bc.type == "rotational":
                self.engine.set_boundary_motion(bc.boundary_id, (0, 0, 0), bc.rotational_speed, bc.rotational_vector, bc.point_on_rotational_vector)
            elif bc.type == "translational":
                self.engine.set_boundary_motion(bc.boundary_id, bc.translational_velocity, 0.0, (0, 0, 0), (0, 0, 0))

    # DEMSolver::insert_particles (dem.cc:484-506)
    def _insertion_due(self) -> bool:
        ins = self.parameters.insertion
        if ins.frequency == 0:
            return False
        return (self.iteration_number % ins.frequency) == 1 or self.iteration_number == 1

    def _insert(self):
        p = self.parameters
        if self._remaining[self._current_type] == 0 and self._current_type != len(p.particle_types) - 1:
            self._current_type += 1
        remaining = self._remaining[self._current_type]
        if remaining == 0:
            return
        if p.insertion.remove_particles:
            self._remove_particles_in_box()
        if p.insertion.method == "file":
            files = p.insertion.input_files
            path = files[self._file_id % len(files)]
            self._file_id += 1
            path = path if os.path.isabs(path) else os.path.join(self.prm_directory, path)
            ids, x, props = file_insertion(p, path, remaining, self._next_id, self._current_type)
        elif p.insertion.method == "list":
            ids, x, props = list_insertion(p, self._next_id, self._current_type)
        elif p.insertion.method == "plane":
            ids, x, props = self._plane.insert(p, self._registered_cells, remaining, self._next_id, self._current_type,
                                               self._distributions[self._current_type])
        else:
            n = min(p.insertion.inserted_this_step, remaining)
            ids, x, props = volume_insertion(p, n, self._next_id, self._current_type, self._distributions[self._current_type],
                                             n_ranks=self.reference_insertion_ranks)
        self.engine.add_particles(ids, x, props)
        self._next_id += len(ids)
        self._remaining[self._current_type] -= len(ids)

    def _cell_of(self, row):
        mesh = self.parameters.mesh
        h = mesh.cell_size
        return tuple(int(math.floor((row[d] - mesh.lo[d]) / h[d])) for d in range(3))

    def _remove_particles_in_box(self):
        """Insertion::find_cells_in_removing_box + remove_particles_in_box (insertion.cc:132-260):
        every particle REGISTERED in a cell whose 8 vertices are in the box goes, and of the cells
        with some vertices in the box those particles whose position is in the box. The C ABI has
        no removal call and needs none: the particle is moved out of the triangulation, which is
        how the reference itself loses particles, and the sort of this iteration drops it. New
        particles then take the lowest free ids again (get_next_free_particle_index)."""
        p, mesh = self.parameters, self.parameters.mesh
        lo, hi = np.asarray(p.insertion.removal_box_point_1), np.asarray(p.insertion.removal_box_point_2)
        h = mesh.cell_size
        ids, x, props = self.engine.get_particles()
        gone = []
        for k, pid in enumerate(ids):
            c = self._registered.get(int(pid), self._cell_of(x[k]))
            inside = [all(lo[d] <= mesh.lo[d] + (c[d] + ((v >> d) & 1)) * h[d] <= hi[d] for d in range(3)) for v in range(8)]
            if all(inside) or (any(inside) and bool(np.all((lo <= x[k]) & (x[k] <= hi)))):
                gone.append(k)
        if gone:
            far = np.ascontiguousarray(x[gone])
            far[:] = np.asarray(mesh.hi) + 10.0 * (np.asarray(mesh.hi) - np.asarray(mesh.lo))
            self.engine.step_host(0, np.ascontiguousarray(ids[gone]), far, np.ascontiguousarray(props[gone]))
            kept = np.delete(ids, gone)
            self._next_id = int(kept.max()) + 1 if len(kept) else 0

    def _is_at_end(self) -> bool:
        # SimulationControlTransient::is_at_end (simulation_control.cc:371-378)
        p = self.parameters
        margin = max(1e-6 * p.time_step, 1e-12 * p.time_end)
        return self.current_time >= (p.time_end - margin)

    def solve(self, max_steps=None, log_callback=None):
        """The `while (simulation_control->integrate())` loop + closing half step.
        `log_callback(iteration_number)` runs where report_statistics does, at the top of every
        iteration that logs (`log frequency`, dem.cc:1116-1118)."""
        pending = 0
        steps = 0
        while not self._is_at_end() and (max_steps is None or steps < max_steps):
            self.iteration_number += 1
            self.current_time += self.parameters.time_step
            steps += 1
            if log_callback is not None and self.iteration_number % max(1, self.parameters.log_frequency) == 0:
                # report_statistics comes first in the iteration (dem.cc:1116-1118): the state the
                # previous iterations left
                if pending:
                    self.engine.step(pending)
                    pending = 0
                log_callback(self.iteration_number)
            # SerialSolid::move_solid_triangulation evaluates the velocity functions at the previous
            # time (serial_solid.cc:343-352): push new values before the step that uses them
            t_prev = self.current_time - self.parameters.time_step
            for k, so in enumerate(self.parameters.solid_surfaces):
                from .prm import solid_velocity_at

                motion = solid_velocity_at(so, t_prev)
                if motion != self._solid_motion[k]:
                    if pending:
                        self.engine.step(pending)
                        pending = 0
                    self.engine.set_solid_motion(k, *motion)
                    self._solid_motion[k] = motion
            if self._insertion_due():
                # the insertion belongs to this iteration: flush earlier ones first
                if pending:
                    self.engine.step(pending)
                    pending = 0
                if any(self._remaining):
                    self._insert()
                # insert_particles calls action_manager->particle_insertion_step() at every insertion
                # iteration, whether or not particles are left to insert (dem.cc:494-500)
                self.engine.force_contact_search()
            pending += 1
            if self._track_registration:
                seen_ids, seen, _ = self.engine.get_particles()  # what a sort in this iteration registers
                searches = self.engine.get_stats().n_rebuilds
                self.engine.step(pending)
                pending = 0
                if self.engine.get_stats().n_rebuilds != searches:
                    self._registered = {int(pid): self._cell_of(row) for pid, row in zip(seen_ids, seen)}
                    self._registered_cells = set(self._registered.values())
        if pending:
            self.engine.step(pending)
        self.engine.synchronize_velocities()
        return self.engine.get_particles()

    # ---- checkpoint / restart (write_checkpoint.cc:30-88, read_checkpoint.cc:14-130) ----
    def write_checkpoint(self, prefix: str):
        """What the reference checkpoints for a DEM run: the simulation control (time, iteration; same
        text layout as its `<prefix>.simulationcontrol`), the particles (id, location, 9 properties) and
        the insertion counters. deal.II serialises the particles inside the p4est triangulation file;
        here they go to `<prefix>.particles_b200.npz`. Contact history is not checkpointed (neither does
        the reference: a restart clears it, dem_action_manager.h:185-200)."""
        ids, x, props = self.engine.get_particles()
        dt = self.parameters.time_step
        with open(prefix + ".simulationcontrol", "w") as f:
            f.write("Simulation control\n")
            for k in range(4):
                f.write(f"dt_{k} {dt!r}\n")
            f.write("CFL  0\n")
            f.write(f"Time {self.current_time!r}\n")
            f.write(f"Iter {self.iteration_number}\n")
        np.savez(prefix + ".particles_b200.npz", id=ids, x=x, props=props, remaining=np.asarray(self._remaining, np.int64),
                 next_id=np.int64(self._next_id), current_type=np.int64(self._current_type))

    def read_checkpoint(self, prefix: str):
        """Resume from write_checkpoint: needs `subsection restart / set restart = true` (the engine then
        continues with regular integrate() steps, dem.cc:1162-1171)."""
        if not self.parameters.restart:
            raise abi.DEMError("read_checkpoint needs `set restart = true`")
        with open(prefix + ".simulationcontrol") as f:
            fields = dict(line.split() for line in f.read().splitlines()[1:] if line.strip())
        data = np.load(prefix + ".particles_b200.npz")
        self.current_time = float(fields["Time"])
        self.iteration_number = int(fields["Iter"])
        self._remaining = [int(v) for v in data["remaining"]]
        self._next_id = int(data["next_id"])
        self._current_type = int(data["current_type"])
        self.engine.set_particles(data["id"], data["x"], data["props"])
        self.engine.set_time(self.iteration_number, self.current_time)

    def test_output(self) -> str:
        """finish_simulation with `subsection test / enable = true` (dem.cc:760-770)."""
        ids, x, props = self.engine.get_particles()
        lines = ["id, type, dp, x, y, z "]
        for i in range(len(ids)):
            lines.append(f"{ids[i]} {int(props[i, 0])} {props[i, 1]:.5f} {x[i, 0]:.4f} {x[i, 1]:.4f} {x[i, 2]:.4f}")
        return "\n".join(lines)
