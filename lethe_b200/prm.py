"""Reader for deal.II ParameterHandler `.prm` files, restricted to the keys the
DEM hot path honours (declared in the reference's
source/core/parameters_lagrangian.cc:13-215,918-1337,1405-1742 and
source/core/parameters.cc for `simulation control` / `mesh`).

The host keeps the reference's parameter interface: same subsection names, same
key spellings, same defaults.  `DEMParameters.to_config()` produces the
`lethe_dem_config` that crosses the C ABI.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

import numpy as np

from . import abi


def parse_prm(text: str) -> dict:
    """`set key = value` / `subsection name … end` → nested dict (keys verbatim)."""
    root: dict = {}
    stack = [root]
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        m = re.match(r"subsection\s+(.*)$", line)
        if m:
            name = m.group(1).strip()
            stack.append(stack[-1].setdefault(name, {}))
            continue
        if line == "end":
            if len(stack) == 1:
                raise ValueError("unbalanced `end` in prm")
            stack.pop()
            continue
        m = re.match(r"set\s+(.*?)\s*=\s*(.*)$", line)
        if m:
            stack[-1][m.group(1).strip()] = m.group(2).strip()
            continue
        raise ValueError(f"cannot parse prm line: {raw!r}")
    if len(stack) != 1:
        raise ValueError("unterminated subsection in prm")
    return root


def _floats(s: str, sep=","):
    return [float(t) for t in s.split(sep) if t.strip()]


def _bool(s: str) -> bool:
    return s.strip().lower() in ("true", "1", "yes", "on")


@dataclass
class ParticleType:
    size_distribution_type: str = "uniform"
    diameter: float = 0.001
    standard_deviation: float = 0.0
    number: int = 0
    density: float = 1000.0
    young: float = 1e6
    poisson: float = 0.3
    restitution: float = 0.1
    friction: float = 0.1
    rolling_friction: float = 0.1
    rolling_viscous_damping: float = 0.1
    surface_energy: float = 0.0
    hamaker: float = 4e-19
    seed: int = 1           # `distribution prn seed`
    min_cutoff: float = -1.0  # `minimum / maximum diameter cutoff` (< 0: average -+ 2.5 sigma)
    max_cutoff: float = -1.0


@dataclass
class Mesh:
    """Uniform hex grid equivalent of the deal.II mesh."""

    lo: tuple = (0.0, 0.0, 0.0)
    hi: tuple = (1.0, 1.0, 1.0)
    n: tuple = (1, 1, 1)
    colorize: bool = False
    cell_order: str = "lexicographic"
    expand_particle_wall_contact_search: bool = False

    @property
    def cell_size(self):
        return tuple((h - l) / n for l, h, n in zip(self.lo, self.hi, self.n))

    @property
    def minimal_cell_diameter(self):
        # GridTools::minimal_cell_diameter: largest diagonal of the smallest cell
        return math.sqrt(sum(h * h for h in self.cell_size))


@dataclass
class BoundaryCondition:
    type: str = "fixed_wall"
    boundary_id: int = 0
    rotational_speed: float = 0.0
    rotational_vector: tuple = (1.0, 0.0, 0.0)
    point_on_rotational_vector: tuple = (0.0, 0.0, 0.0)
    translational_velocity: tuple = (0.0, 0.0, 0.0)
    periodic_id_0: int = 0
    periodic_id_1: int = 0
    periodic_direction: int = 0


@dataclass
class SolidSurface:
    """`subsection solid objects / solid surfaces / solid object N`
    (source/core/parameters.cc, Parameters::RigidSolidObject; SerialSolid<2,3>)."""

    mesh_file: str = ""
    rotation_axis: tuple = (1.0, 0.0, 0.0)
    rotation_angle: float = 0.0
    translation: tuple = (0.0, 0.0, 0.0)
    # `Function expression`s of time: floats, or muparser strings evaluated by `solid_velocity_at`
    translational_velocity: tuple = (0.0, 0.0, 0.0)
    angular_velocity: tuple = (0.0, 0.0, 0.0)
    center_of_rotation: tuple = (0.0, 0.0, 0.0)


@dataclass
class Insertion:
    method: str = "volume"
    list_x: tuple = ()
    list_y: tuple = ()
    list_z: tuple = ()
    list_velocity: tuple = ()  # per particle (vx, vy, vz)
    list_omega: tuple = ()
    list_diameters: tuple = ()
    input_files: tuple = ()  # `insertion method = file`: `list of input files`
    inserted_this_step: int = 0
    frequency: int = 1
    box_point_1: tuple = (0.0, 0.0, 0.0)
    box_point_2: tuple = (1.0, 1.0, 1.0)
    distance_threshold: float = 1.0
    maximum_offset: float = 1.0
    seed: int = 1
    direction_sequence: tuple = (0, 1, 2)
    initial_velocity: tuple = (0.0, 0.0, 0.0)
    initial_omega: tuple = (0.0, 0.0, 0.0)


@dataclass
class DEMParameters:
    """Subset of DEMSolverParameters<3> (include/dem/dem_solver_parameters.h:15-92)."""

    dimension: int = 3
    time_step: float = 1.0
    time_end: float = 1.0
    log_frequency: int = 1
    output_frequency: int = 1
    # model parameters (defaults: parameters_lagrangian.cc:918-1099)
    contact_detection_method: str = "dynamic"
    contact_detection_frequency: int = 1
    dynamic_contact_search_factor: float = 0.8
    neighborhood_threshold: float = 1.3
    pp_model: str = "hertz_mindlin_limit_overlap"
    pw_model: str = "nonlinear"
    rolling_model: str = "constant"
    integration_method: str = "velocity_verlet"
    dmt_cut_off_threshold: float = 0.1
    f_coefficient: float = 0.0
    solver_type: str = "dem"
    # physical properties
    g: tuple = (0.0, 0.0, 0.0)
    particle_types: list = field(default_factory=lambda: [ParticleType()])
    young_wall: float = 1e6
    poisson_wall: float = 0.3
    restitution_wall: float = 0.1
    friction_wall: float = 0.1
    rolling_friction_wall: float = 0.1
    rolling_viscous_damping_wall: float = 0.1
    surface_energy_wall: float = 0.0
    hamaker_wall: float = 4e-19
    mesh: Mesh = field(default_factory=Mesh)
    insertion: Insertion = field(default_factory=Insertion)
    boundary_conditions: list = field(default_factory=list)
    floating_walls: list = field(default_factory=list)  # (point, normal, t_start, t_end)
    solid_surfaces: list = field(default_factory=list)  # SolidSurface
    restart: bool = False
    test_enabled: bool = False

    # ------------------------------------------------------------------
    @property
    def d_max(self) -> float:
        # setup_distributions: maximum_particle_diameter (dem.cc:149-159);
        # normal/lognormal PSDs are truncated at +-2.5 sigma in the reference
        from .distributions import make_distribution

        return max(make_distribution(t).max_diameter() for t in self.particle_types)

    @property
    def periodic(self):
        p = [0, 0, 0]
        for bc in self.boundary_conditions:
            if bc.type == "periodic":
                p[bc.periodic_direction] = 1
        return tuple(p)

    @property
    def outlet_boundaries(self):
        return {bc.boundary_id for bc in self.boundary_conditions if bc.type == "outlet"}

    def smallest_contact_search_criterion(self) -> float:
        # dem.cc:289-294
        d = self.d_max
        return min(
            self.mesh.minimal_cell_diameter - d * 0.5,
            self.dynamic_contact_search_factor * (self.neighborhood_threshold - 1) * d * 0.5,
        )

    def to_config(self, store_forces=False, moi_override=0.0, slab=None) -> abi.Config:
        if self.integration_method not in ("velocity_verlet", "explicit_euler"):
            raise abi.DEMError(f"unknown integration method {self.integration_method!r} (velocity_verlet|explicit_euler)")
        if self.solver_type != "dem":
            raise abi.DEMError("solver type dem_mp is out of scope")
        for name, table, val in (
            ("particle particle contact force method", abi.PP_MODELS, self.pp_model),
            ("particle wall contact force method", abi.PW_MODELS, self.pw_model),
            ("rolling resistance torque method", abi.ROLLING_MODELS, self.rolling_model),
            ("contact detection method", abi.DETECTION, self.contact_detection_method),
        ):
            if val not in table:
                raise abi.DEMError(f"invalid {name}: {val!r}")
        c = abi.Config()
        c.pp_model = abi.PP_MODELS[self.pp_model]
        c.pw_model = abi.PW_MODELS[self.pw_model]
        c.rolling_model = abi.ROLLING_MODELS[self.rolling_model]
        c.integrator = 1 if self.integration_method == "explicit_euler" else 0
        c.detection = abi.DETECTION[self.contact_detection_method]
        c.contact_detection_frequency = self.contact_detection_frequency
        c.cell_order = abi.CELL_ORDER[self.mesh.cell_order]
        c.store_forces = int(store_forces)
        c.dt = self.time_step
        c.g[:] = self.g
        c.neighborhood_threshold = self.neighborhood_threshold
        c.d_max = self.d_max
        c.smallest_contact_search_criterion = self.smallest_contact_search_criterion()
        c.dmt_cut_off_threshold = self.dmt_cut_off_threshold
        c.f_coefficient_epsd = self.f_coefficient
        c.moi_override = moi_override
        c.n_types = len(self.particle_types)
        c.restart = int(self.restart)
        for i, t in enumerate(self.particle_types):
            c.young[i] = t.young
            c.poisson[i] = t.poisson
            c.restitution[i] = t.restitution
            c.friction[i] = t.friction
            c.rolling_friction[i] = t.rolling_friction
            c.rolling_viscous_damping[i] = t.rolling_viscous_damping
            c.surface_energy[i] = t.surface_energy
            c.hamaker[i] = t.hamaker
        c.young_wall = self.young_wall
        c.poisson_wall = self.poisson_wall
        c.restitution_wall = self.restitution_wall
        c.friction_wall = self.friction_wall
        c.rolling_friction_wall = self.rolling_friction_wall
        c.rolling_viscous_damping_wall = self.rolling_viscous_damping_wall
        c.surface_energy_wall = self.surface_energy_wall
        c.hamaker_wall = self.hamaker_wall
        c.grid_lo[:] = self.mesh.lo
        c.cell_size[:] = self.mesh.cell_size
        c.grid_n[:] = self.mesh.n
        c.periodic[:] = self.periodic
        if slab is None:
            c.slab_axis, c.slab_lo, c.slab_hi = -1, 0, 0
        else:
            c.slab_axis, c.slab_lo, c.slab_hi = slab
        return c


def evaluate_function(expr, t: float) -> float:
    """Value at time t of one component of a deal.II `Function expression` (muparser syntax:
    `if(c, a, b)`, `^`, the usual functions, variable t; the solid-object velocity functions are
    evaluated at the centre of rotation, whose coordinates no reference case uses)."""
    if not isinstance(expr, str):
        return float(expr)
    import math
    import re as _re

    py = _re.sub(r"\bif\s*\(", "_if(", expr).replace("^", "**")
    env = {name: getattr(math, name) for name in ("sin", "cos", "tan", "exp", "log", "sqrt", "tanh", "atan", "asin", "acos")}
    env.update({"_if": lambda c, a, b: a if c else b, "pi": math.pi, "t": t, "x": 0.0, "y": 0.0, "z": 0.0, "abs": abs, "min": min, "max": max})
    return float(eval(py, {"__builtins__": {}}, env))


def solid_velocity_at(solid: "SolidSurface", t: float):
    return (tuple(evaluate_function(c, t) for c in solid.translational_velocity),
            tuple(evaluate_function(c, t) for c in solid.angular_velocity))


_ROLLING_ALIASES = {
    "no_resistance": "none",
    "constant_resistance": "constant",
    "viscous_resistance": "viscous",
    "epsd_resistance": "epsd",
}


def _parse_mesh(sec: dict) -> Mesh:
    if sec.get("type", "dealii") != "dealii":
        raise abi.DEMError("only `mesh type = dealii` uniform hex grids are on the B200 path (gmsh meshes are out of scope)")
    grid = sec.get("grid type", "hyper_cube")
    args = [a.strip() for a in sec.get("grid arguments", "-1 : 1 : false").split(":")]
    ref = int(sec.get("initial refinement", "0"))
    if grid == "hyper_cube":
        lo, hi = float(args[0]), float(args[1])
        colorize = _bool(args[2]) if len(args) > 2 else False
        n = (2**ref,) * 3
        m = Mesh((lo,) * 3, (hi,) * 3, n, colorize, "morton")
    elif grid in ("subdivided_hyper_rectangle", "hyper_rectangle"):
        if grid == "hyper_rectangle":
            reps = [1, 1, 1]
            p1, p2 = _floats(args[0]), _floats(args[1])
            colorize = _bool(args[2]) if len(args) > 2 else False
        else:
            reps = [int(v) for v in _floats(args[0])]
            p1, p2 = _floats(args[1]), _floats(args[2])
            colorize = _bool(args[3]) if len(args) > 3 else False
        n = tuple(r * 2**ref for r in reps)
        m = Mesh(tuple(p1), tuple(p2), n, colorize, "lexicographic" if ref == 0 else "morton")
    else:
        raise abi.DEMError(f"grid type {grid!r} is not a uniform hex grid; out of scope for the B200 path")
    m.expand_particle_wall_contact_search = _bool(sec.get("expand particle-wall contact search", "false"))
    return m


def parameters_from_prm(text: str) -> DEMParameters:
    d = parse_prm(text)
    p = DEMParameters()
    p.dimension = int(d.get("dimension", "3"))
    if p.dimension != 3:
        raise abi.DEMError("only dimension = 3 is on the B200 path")
    sc = d.get("simulation control", {})
    p.time_step = float(sc.get("time step", "1"))
    p.time_end = float(sc.get("time end", "1"))
    p.log_frequency = int(sc.get("log frequency", "1"))
    p.output_frequency = int(sc.get("output frequency", "1"))
    p.test_enabled = _bool(d.get("test", {}).get("enable", "false"))
    p.restart = _bool(d.get("restart", {}).get("restart", "false"))

    mp = d.get("model parameters", {})
    cd = mp.get("contact detection", {})
    p.contact_detection_method = cd.get("contact detection method", "dynamic")
    p.contact_detection_frequency = int(cd.get("frequency", "1"))
    p.dynamic_contact_search_factor = float(cd.get("dynamic contact search size coefficient", "0.8"))
    p.neighborhood_threshold = float(cd.get("neighborhood threshold", "1.3"))
    p.pp_model = mp.get("particle particle contact force method", "hertz_mindlin_limit_overlap")
    p.pw_model = mp.get("particle wall contact force method", "nonlinear")
    rolling = mp.get("rolling resistance torque method", "constant")
    p.rolling_model = _ROLLING_ALIASES.get(rolling, rolling)
    p.integration_method = mp.get("integration method", "velocity_verlet")
    p.dmt_cut_off_threshold = float(mp.get("dmt cut-off threshold", "0.1"))
    p.f_coefficient = float(mp.get("f coefficient", "0.0"))
    p.solver_type = mp.get("solver type", "dem")

    lp = d.get("lagrangian physical properties", {})
    if "g" in lp:
        p.g = tuple(_floats(lp["g"]))
    else:
        p.g = (float(lp.get("gx", "0")), float(lp.get("gy", "0")), float(lp.get("gz", "0")))
    n_types = int(lp.get("number of particle types", "1"))
    if n_types > abi.MAX_TYPES:
        raise abi.DEMError("at most 5 particle types")
    p.particle_types = []
    for i in range(n_types):
        s = lp.get(f"particle type {i}", {})
        t = ParticleType()
        t.size_distribution_type = s.get("size distribution type", "uniform")
        t.diameter = float(s.get("average diameter", s.get("diameter", "0.001")))
        t.standard_deviation = float(s.get("standard deviation", "0"))
        t.number = int(s.get("number of particles", s.get("number", "0")))
        t.density = float(s.get("density particles", "1000"))
        t.young = float(s.get("young modulus particles", "1000000"))
        t.poisson = float(s.get("poisson ratio particles", "0.3"))
        t.restitution = float(s.get("restitution coefficient particles", "0.1"))
        t.friction = float(s.get("friction coefficient particles", "0.1"))
        t.rolling_friction = float(s.get("rolling friction particles", "0.1"))
        t.rolling_viscous_damping = float(s.get("rolling viscous damping particles", "0.1"))
        t.surface_energy = float(s.get("surface energy particles", "0"))
        t.hamaker = float(s.get("hamaker constant particles", "4e-19"))
        t.seed = int(s.get("distribution prn seed", "1"))
        t.min_cutoff = float(s.get("minimum diameter cutoff", "-1"))
        t.max_cutoff = float(s.get("maximum diameter cutoff", "-1"))
        p.particle_types.append(t)
    p.young_wall = float(lp.get("young modulus wall", "1000000"))
    p.poisson_wall = float(lp.get("poisson ratio wall", "0.3"))
    p.restitution_wall = float(lp.get("restitution coefficient wall", "0.1"))
    p.friction_wall = float(lp.get("friction coefficient wall", "0.1"))
    p.rolling_friction_wall = float(lp.get("rolling friction wall", "0.1"))
    p.rolling_viscous_damping_wall = float(lp.get("rolling viscous damping wall", "0.1"))
    p.surface_energy_wall = float(lp.get("surface energy wall", "0"))
    p.hamaker_wall = float(lp.get("hamaker constant wall", "4e-19"))

    if "mesh" in d:
        p.mesh = _parse_mesh(d["mesh"])

    ii = d.get("insertion info", {})
    ins = Insertion()
    ins.method = ii.get("insertion method", "volume")
    ins.inserted_this_step = int(ii.get("inserted number of particles at each time step", "0"))
    ins.frequency = int(ii.get("insertion frequency", "1"))
    if "insertion box points coordinates" in ii:
        a, b = ii["insertion box points coordinates"].split(":")
        ins.box_point_1, ins.box_point_2 = tuple(_floats(a)), tuple(_floats(b))
    ins.distance_threshold = float(ii.get("insertion distance threshold", "1"))
    ins.maximum_offset = float(ii.get("insertion maximum offset", "1"))
    ins.seed = int(ii.get("insertion prn seed", "1"))
    if "insertion direction sequence" in ii:
        ins.direction_sequence = tuple(int(v) for v in _floats(ii["insertion direction sequence"]))
    if "initial velocity" in ii:
        ins.initial_velocity = tuple(_floats(ii["initial velocity"]))
    if "initial angular velocity" in ii:
        ins.initial_omega = tuple(_floats(ii["initial angular velocity"]))
    if ins.method == "list":
        # insertion_list.cc: explicit positions / velocities / diameters
        ins.list_x, ins.list_y, ins.list_z = (tuple(_floats(ii.get("list " + a, ""))) for a in "xyz")
        n = len(ins.list_x)

        def triple(prefix):
            cols = [list(_floats(ii.get(f"list {prefix} {a}", ""))) for a in "xyz"]
            cols = [c + [0.0] * (n - len(c)) for c in cols]
            return tuple(zip(*cols)) if n else ()

        ins.list_velocity, ins.list_omega = triple("velocity"), triple("omega")
        ins.list_diameters = tuple(_floats(ii.get("list diameters", "")))
    if ins.method == "file":
        ins.input_files = tuple(f.strip() for f in ii.get("list of input files", "particles.input").split(",") if f.strip())
    p.insertion = ins

    so = d.get("solid objects", {}).get("solid surfaces", {})
    for i in range(int(so.get("number of solids", "0"))):
        s = so.get(f"solid object {i}", {})
        m = s.get("mesh", {})
        if m.get("type", "gmsh") != "gmsh":
            raise abi.DEMError("solid surfaces: only `mesh type = gmsh` files are read")

        def constant_function(sub, default=(0.0, 0.0, 0.0)):
            expr = s.get(sub, {}).get("Function expression")
            if expr is None:
                return default
            out = []
            for v in expr.split(";"):
                try:
                    out.append(float(v))
                except ValueError:
                    out.append(v.strip())  # a function of t (muparser syntax), see evaluate_function
            return tuple(out)

        p.solid_surfaces.append(SolidSurface(
            mesh_file=m.get("file name", ""),
            rotation_axis=tuple(_floats(m.get("initial rotation axis", "1, 0, 0"))),
            rotation_angle=float(m.get("initial rotation angle", "0")),
            translation=tuple(_floats(m.get("initial translation", "0, 0, 0"))),
            translational_velocity=constant_function("translational velocity"),
            angular_velocity=constant_function("angular velocity"),
            center_of_rotation=tuple(_floats(s.get("center of rotation", "0, 0, 0")))))

    bcs = d.get("DEM boundary conditions", {})
    for i in range(int(bcs.get("number of boundary conditions", "0"))):
        s = bcs.get(f"boundary condition {i}", {})
        bc = BoundaryCondition()
        bc.type = s.get("type", "fixed_wall")
        bc.boundary_id = int(s.get("boundary id", "0"))
        bc.rotational_speed = float(s.get("rotational speed", "0"))
        if "rotational vector" in s:
            bc.rotational_vector = tuple(_floats(s["rotational vector"]))
        if "point on rotational vector" in s:
            bc.point_on_rotational_vector = tuple(_floats(s["point on rotational vector"]))
        bc.translational_velocity = (float(s.get("speed x", "0")), float(s.get("speed y", "0")), float(s.get("speed z", "0")))
        bc.periodic_id_0 = int(s.get("periodic id 0", "0"))
        bc.periodic_id_1 = int(s.get("periodic id 1", "0"))
        bc.periodic_direction = int(s.get("periodic direction", "0"))
        p.boundary_conditions.append(bc)

    fw = d.get("floating walls", {})
    for i in range(int(fw.get("number of floating walls", "0"))):
        s = fw.get(f"wall {i}", {})
        pt = s.get("point on wall", {})
        nv = s.get("normal vector", {})
        if isinstance(pt, dict):  # legacy nested form `subsection point on wall / set x = …`
            point = (float(pt.get("x", "0")), float(pt.get("y", "0")), float(pt.get("z", "0")))
            normal = (float(nv.get("nx", "0")), float(nv.get("ny", "0")), float(nv.get("nz", "0")))
        else:
            point, normal = tuple(_floats(pt)), tuple(_floats(nv))
        p.floating_walls.append((point, normal, float(s.get("start time", "0")), float(s.get("end time", "0"))))
    return p


def load_prm(path: str) -> DEMParameters:
    with open(path) as f:
        return parameters_from_prm(f.read())
