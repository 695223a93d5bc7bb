// dem_parameters.cc — see dem_parameters.h.
#include "dem_parameters.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>

namespace lethe_b200
{
  namespace
  {
    Vec3 to_vec3(const std::vector<double> &v, const char *what)
    {
      if (v.size() != 3)
        throw std::runtime_error(std::string("expected 3 comma-separated numbers for `") + what + "`");
      return {{v[0], v[1], v[2]}};
    }

    int lookup(const std::map<std::string, int> &table, const std::string &value, const char *key)
    {
      auto it = table.find(value);
      if (it == table.end())
        throw std::runtime_error(std::string("invalid `") + key + "`: " + value);
      return it->second;
    }

    // `subsection mesh`: only deal.II generated uniform hex grids are on the B200 path
    Mesh parse_mesh(const PrmSection &sec)
    {
      if (sec.get("type", "dealii") != "dealii")
        throw std::runtime_error("only `mesh type = dealii` uniform hex grids are on the B200 path (gmsh meshes are out of scope)");
      const std::string grid = sec.get("grid type", "hyper_cube");
      const auto args = PrmSection::split(sec.get("grid arguments", "-1 : 1 : false"), ':');
      const int ref = int(sec.get_int("initial refinement", 0));
      auto as_bool = [](const std::string &s) { return PrmSection::lower(s) == "true"; };
      Mesh m;
      if (grid == "hyper_cube")
        {
          const double lo = std::stod(args.at(0)), hi = std::stod(args.at(1));
          m.lo = {{lo, lo, lo}};
          m.hi = {{hi, hi, hi}};
          m.colorize = args.size() > 2 && as_bool(args[2]);
          m.n = {{1 << ref, 1 << ref, 1 << ref}};
          m.cell_order = LETHE_CELL_ORDER_MORTON; // refine_global enumerates children hierarchically
        }
      else if (grid == "subdivided_hyper_rectangle" || grid == "hyper_rectangle")
        {
          std::vector<double> reps{1, 1, 1};
          size_t k = 0;
          if (grid == "subdivided_hyper_rectangle")
            reps = PrmSection::split_doubles(args.at(k++), ',');
          m.lo = to_vec3(PrmSection::split_doubles(args.at(k), ','), "grid arguments");
          m.hi = to_vec3(PrmSection::split_doubles(args.at(k + 1), ','), "grid arguments");
          m.colorize = args.size() > k + 2 && as_bool(args[k + 2]);
          for (int d = 0; d < 3; ++d)
            m.n[d] = int(reps.at(d)) << ref;
          m.cell_order = ref == 0 ? LETHE_CELL_ORDER_LEXICOGRAPHIC : LETHE_CELL_ORDER_MORTON;
        }
      else
        throw std::runtime_error("grid type `" + grid + "` is not a uniform hex grid; out of scope for the B200 path");
      m.expand_particle_wall_contact_search = sec.get_bool("expand particle-wall contact search", false);
      return m;
    }
  } // namespace

  void SolidSurface::velocities_at(double t, Vec3 &translational, Vec3 &angular) const
  {
    FunctionExpression::Variables v;
    v.t = t;
    translational = translational_velocity;
    angular = angular_velocity;
    for (int c = 0; c < 3; ++c)
      {
        if (!translational_velocity_function[c].empty())
          translational[c] = translational_velocity_function[c](v);
        if (!angular_velocity_function[c].empty())
          angular[c] = angular_velocity_function[c](v);
      }
  }

  double Mesh::minimal_cell_diameter() const
  {
    const Vec3 h = cell_size();
    return std::sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]);
  }

  DEMParameters DEMParameters::from_prm_file(const std::string &path)
  {
    std::ifstream in(path);
    if (!in)
      throw std::runtime_error("cannot open parameter file " + path);
    DEMParameters p = from_prm(*parse_prm(in));
    const auto slash = path.find_last_of('/');
    p.prm_directory = slash == std::string::npos ? "." : path.substr(0, slash);
    return p;
  }

  DEMParameters DEMParameters::from_prm(const PrmSection &d)
  {
    DEMParameters p;
    if (d.get_int("dimension", 3) != 3)
      throw std::runtime_error("only dimension = 3 is on the B200 path");
    const PrmSection &sc = d.sub("simulation control");
    p.time_step = sc.get_double("time step", 1.);
    p.time_end = sc.get_double("time end", 1.);
    p.log_frequency = sc.get_int("log frequency", 1);
    p.output_frequency = sc.get_int("output frequency", 1);
    p.test_enabled = d.sub("test").get_bool("enable", false);
    p.restart = d.sub("restart").get_bool("restart", false);

    const PrmSection &mp = d.sub("model parameters");
    const PrmSection &cd = mp.sub("contact detection");
    p.contact_detection_method = cd.get("contact detection method", "dynamic");
    p.contact_detection_frequency = cd.get_int("frequency", 1);
    p.dynamic_contact_search_factor = cd.get_double("dynamic contact search size coefficient", 0.8);
    p.neighborhood_threshold = cd.get_double("neighborhood threshold", 1.3);
    p.pp_model = mp.get("particle particle contact force method", "hertz_mindlin_limit_overlap");
    p.pw_model = mp.get("particle wall contact force method", "nonlinear");
    p.rolling_model = mp.get("rolling resistance torque method", "constant");
    static const std::map<std::string, std::string> rolling_alias{{"no_resistance", "none"},
                                                                  {"constant_resistance", "constant"},
                                                                  {"viscous_resistance", "viscous"},
                                                                  {"epsd_resistance", "epsd"}};
    if (rolling_alias.count(p.rolling_model))
      p.rolling_model = rolling_alias.at(p.rolling_model);
    p.integration_method = mp.get("integration method", "velocity_verlet");
    p.dmt_cut_off_threshold = mp.get_double("dmt cut-off threshold", 0.1);
    p.f_coefficient = mp.get_double("f coefficient", 0.0);
    p.solver_type = mp.get("solver type", "dem");
    // `subsection adaptive sparse contacts` (parameters_lagrangian.cc:1059-1116)
    const PrmSection &asc = mp.sub("adaptive sparse contacts");
    p.sparse_contacts = asc.get("enable adaptive sparse contacts", "false") == "true";
    p.asc_granular_temperature_threshold = asc.get_double("granular temperature threshold", 1e-4);
    p.asc_solid_fraction_threshold = asc.get_double("solid fraction threshold", 0.4);
    if (asc.get("enable particle advection", "false") == "true")
      throw std::runtime_error("adaptive sparse contacts: particle advection (CFD-DEM) is out of scope");

    const PrmSection &lp = d.sub("lagrangian physical properties");
    if (lp.has("g"))
      p.g = to_vec3(lp.get_list("g"), "g");
    else
      p.g = {{lp.get_double("gx", 0), lp.get_double("gy", 0), lp.get_double("gz", 0)}};
    const long n_types = lp.get_int("number of particle types", 1);
    if (n_types < 1 || n_types > LETHE_DEM_MAX_TYPES)
      throw std::runtime_error("number of particle types must be 1..5");
    p.particle_types.clear();
    for (long i = 0; i < n_types; ++i)
      {
        const PrmSection &s = lp.sub("particle type " + std::to_string(i));
        ParticleType t;
        t.size_distribution_type = s.get("size distribution type", "uniform");
        t.average_diameter = s.has("average diameter") ? s.get_double("average diameter", 0.001) : s.get_double("diameter", 0.001);
        t.standard_deviation = s.get_double("standard deviation", 0);
        t.number_of_particles = s.has("number of particles") ? s.get_int("number of particles", 0) : s.get_int("number", 0);
        t.density = s.get_double("density particles", 1000);
        t.young_modulus = s.get_double("young modulus particles", 1e6);
        t.poisson_ratio = s.get_double("poisson ratio particles", 0.3);
        t.restitution_coefficient = s.get_double("restitution coefficient particles", 0.1);
        t.friction_coefficient = s.get_double("friction coefficient particles", 0.1);
        t.rolling_friction = s.get_double("rolling friction particles", 0.1);
        t.rolling_viscous_damping = s.get_double("rolling viscous damping particles", 0.1);
        t.surface_energy = s.get_double("surface energy particles", 0);
        t.hamaker_constant = s.get_double("hamaker constant particles", 4e-19);
        t.prn_seed = s.get_int("distribution prn seed", 1);
        t.min_cutoff = s.get_double("minimum diameter cutoff", -1);
        t.max_cutoff = s.get_double("maximum diameter cutoff", -1);
        p.particle_types.push_back(t);
      }
    p.young_wall = lp.get_double("young modulus wall", 1e6);
    p.poisson_wall = lp.get_double("poisson ratio wall", 0.3);
    p.restitution_wall = lp.get_double("restitution coefficient wall", 0.1);
    p.friction_wall = lp.get_double("friction coefficient wall", 0.1);
    p.rolling_friction_wall = lp.get_double("rolling friction wall", 0.1);
    p.rolling_viscous_damping_wall = lp.get_double("rolling viscous damping wall", 0.1);
    p.surface_energy_wall = lp.get_double("surface energy wall", 0);
    p.hamaker_wall = lp.get_double("hamaker constant wall", 4e-19);

    if (d.has_sub("mesh"))
      p.mesh = parse_mesh(d.sub("mesh"));

    const PrmSection &ii = d.sub("insertion info");
    InsertionInfo &ins = p.insertion;
    ins.method = ii.get("insertion method", "volume");
    ins.inserted_this_step = ii.get_int("inserted number of particles at each time step", 0);
    ins.frequency = ii.get_int("insertion frequency", 1);
    if (ii.has("insertion box points coordinates"))
      {
        const auto pts = PrmSection::split(ii.get("insertion box points coordinates", ""), ':');
        ins.box_point_1 = to_vec3(PrmSection::split_doubles(pts.at(0), ','), "insertion box points coordinates");
        ins.box_point_2 = to_vec3(PrmSection::split_doubles(pts.at(1), ','), "insertion box points coordinates");
      }
    ins.distance_threshold = ii.get_double("insertion distance threshold", 1);
    ins.maximum_offset = ii.get_double("insertion maximum offset", 1);
    ins.prn_seed = ii.get_int("insertion prn seed", 1);
    if (ii.sub("insertion acceptance function").has("Function expression"))
      ins.acceptance_function = FunctionExpression(ii.sub("insertion acceptance function").get("Function expression", ""));
    ins.remove_particles = ii.get_bool("remove particles", false);
    if (ii.has("removal box points coordinates"))
      {
        const auto pts = PrmSection::split(ii.get("removal box points coordinates", ""), ':');
        ins.removal_box_point_1 = to_vec3(PrmSection::split_doubles(pts.at(0), ','), "removal box points coordinates");
        ins.removal_box_point_2 = to_vec3(PrmSection::split_doubles(pts.at(1), ','), "removal box points coordinates");
      }
    if (ii.has("insertion plane point"))
      ins.plane_point = to_vec3(ii.get_list("insertion plane point"), "insertion plane point");
    if (ii.has("insertion plane normal vector"))
      ins.plane_normal = to_vec3(ii.get_list("insertion plane normal vector"), "insertion plane normal vector");
    if (ii.has("insertion direction sequence"))
      {
        const auto seq = ii.get_list("insertion direction sequence");
        for (int k = 0; k < 3; ++k)
          ins.direction_sequence[k] = int(seq.at(k));
      }
    if (ii.has("initial velocity"))
      ins.initial_velocity = to_vec3(ii.get_list("initial velocity"), "initial velocity");
    if (ii.has("initial angular velocity"))
      ins.initial_omega = to_vec3(ii.get_list("initial angular velocity"), "initial angular velocity");

    if (ins.method == "list")
      {
        ins.list_x = ii.get_list("list x");
        ins.list_y = ii.get_list("list y");
        ins.list_z = ii.get_list("list z");
        ins.list_vx = ii.get_list("list velocity x");
        ins.list_vy = ii.get_list("list velocity y");
        ins.list_vz = ii.get_list("list velocity z");
        ins.list_wx = ii.get_list("list omega x");
        ins.list_wy = ii.get_list("list omega y");
        ins.list_wz = ii.get_list("list omega z");
        ins.list_diameters = ii.get_list("list diameters");
      }

    if (ins.method == "file")
      for (const auto &f : PrmSection::split(ii.get("list of input files", "particles.input"), ','))
        if (!PrmSection::trim(f).empty())
          ins.input_files.push_back(PrmSection::trim(f));

    const PrmSection &so = d.sub("solid objects").sub("solid surfaces");
    for (long i = 0; i < so.get_int("number of solids", 0); ++i)
      {
        const PrmSection &s = so.sub("solid object " + std::to_string(i));
        const PrmSection &m = s.sub("mesh");
        if (m.get("type", "dealii") == "dealii" && m.get("simplex", "false") != "true")
          throw std::runtime_error("solid surfaces: a `type = dealii` mesh needs `simplex = true` (the contact search is on triangles)");
        SolidSurface sd;
        sd.mesh_file = m.get("file name", "");
        sd.mesh_type = m.get("type", "dealii");
        sd.grid_type = m.get("grid type", "hyper_cube");
        sd.grid_arguments = m.get("grid arguments", "-1 : 1 : false");
        sd.initial_refinement = m.get_int("initial refinement", 0);
        if (m.has("initial rotation axis"))
          sd.rotation_axis = to_vec3(m.get_list("initial rotation axis"), "initial rotation axis");
        sd.rotation_angle = m.get_double("initial rotation angle", 0);
        if (m.has("initial translation"))
          sd.translation = to_vec3(m.get_list("initial translation"), "initial translation");
        // `Function expression = a ; b ; c`: three expressions of t (muparser syntax)
        auto velocity_function = [&](const char *sub, std::array<FunctionExpression, 3> &fn, Vec3 &at_start) {
          const PrmSection &f = s.sub(sub);
          if (!f.has("Function expression"))
            return;
          const auto parts = PrmSection::split(f.get("Function expression", ""), ';');
          if (parts.size() != 3)
            throw std::runtime_error(std::string("expected 3 `;`-separated expressions for `") + sub + "`");
          for (int c = 0; c < 3; ++c)
            {
              fn[c] = FunctionExpression(PrmSection::trim(parts[c]));
              at_start[c] = fn[c](FunctionExpression::Variables());
            }
        };
        velocity_function("translational velocity", sd.translational_velocity_function, sd.translational_velocity);
        velocity_function("angular velocity", sd.angular_velocity_function, sd.angular_velocity);
        if (s.has("center of rotation"))
          sd.center_of_rotation = to_vec3(s.get_list("center of rotation"), "center of rotation");
        p.solid_surfaces.push_back(sd);
      }

    const PrmSection &bcs = d.sub("DEM boundary conditions");
    for (long i = 0; i < bcs.get_int("number of boundary conditions", 0); ++i)
      {
        const PrmSection &s = bcs.sub("boundary condition " + std::to_string(i));
        BoundaryCondition bc;
        bc.type = s.get("type", "fixed_wall");
        bc.boundary_id = unsigned(s.get_int("boundary id", 0));
        bc.rotational_speed = s.get_double("rotational speed", 0);
        if (s.has("rotational vector"))
          bc.rotational_vector = to_vec3(s.get_list("rotational vector"), "rotational vector");
        if (s.has("point on rotational vector"))
          bc.point_on_rotational_vector = to_vec3(s.get_list("point on rotational vector"), "point on rotational vector");
        bc.translational_velocity = {{s.get_double("speed x", 0), s.get_double("speed y", 0), s.get_double("speed z", 0)}};
        bc.periodic_id_0 = unsigned(s.get_int("periodic id 0", 0));
        bc.periodic_id_1 = unsigned(s.get_int("periodic id 1", 0));
        bc.periodic_direction = int(s.get_int("periodic direction", 0));
        p.boundary_conditions.push_back(bc);
      }

    const PrmSection &fw = d.sub("floating walls");
    const long n_fw = fw.get_int("number of floating walls", 0);
    if (n_fw > LETHE_DEM_MAX_FLOATING_WALLS)
      throw std::runtime_error("at most 9 floating walls");
    for (long i = 0; i < n_fw; ++i)
      {
        const PrmSection &s = fw.sub("wall " + std::to_string(i));
        FloatingWall w;
        if (s.has_sub("point on wall")) // legacy nested form: subsection point on wall / set x = …
          {
            const PrmSection &pt = s.sub("point on wall"), &nv = s.sub("normal vector");
            w.point = {{pt.get_double("x", 0), pt.get_double("y", 0), pt.get_double("z", 0)}};
            w.normal = {{nv.get_double("nx", 0), nv.get_double("ny", 0), nv.get_double("nz", 0)}};
          }
        else
          {
            w.point = to_vec3(s.get_list("point on wall"), "point on wall");
            w.normal = to_vec3(s.get_list("normal vector"), "normal vector");
          }
        w.time_start = s.get_double("start time", 0);
        w.time_end = s.get_double("end time", 0);
        p.floating_walls.push_back(w);
      }
    return p;
  }

  // NormalDistribution / LogNormalDistribution constructors (distributions.cc:23-165,231-290),
  // number-based weighting
  double ParticleType::max_diameter() const
  {
    if (size_distribution_type == "uniform")
      return average_diameter;
    if (max_cutoff >= 0)
      return max_cutoff;
    if (size_distribution_type == "lognormal")
      {
        const double sigma_ln = std::sqrt(std::log(1. + (standard_deviation / average_diameter) * (standard_deviation / average_diameter)));
        return std::exp((std::log(average_diameter) - 0.5 * sigma_ln * sigma_ln) + 2.5 * sigma_ln);
      }
    return average_diameter + 2.5 * standard_deviation;
  }
  double ParticleType::min_diameter() const
  {
    if (size_distribution_type == "uniform")
      return average_diameter;
    if (min_cutoff >= 0)
      return min_cutoff;
    if (size_distribution_type == "lognormal")
      {
        const double sigma_ln = std::sqrt(std::log(1. + (standard_deviation / average_diameter) * (standard_deviation / average_diameter)));
        return std::exp((std::log(average_diameter) - 0.5 * sigma_ln * sigma_ln) - 2.5 * sigma_ln);
      }
    return average_diameter - 2.5 * standard_deviation;
  }

  double DEMParameters::maximum_particle_diameter() const
  {
    // setup_distributions: the largest find_max_diameter() over the particle types
    double d = 0;
    for (const auto &t : particle_types)
      d = std::max(d, t.max_diameter());
    return d;
  }

  std::array<int, 3> DEMParameters::periodic_directions() const
  {
    std::array<int, 3> p{{0, 0, 0}};
    for (const auto &bc : boundary_conditions)
      if (bc.type == "periodic")
        p.at(bc.periodic_direction) = 1;
    return p;
  }

  std::vector<unsigned> DEMParameters::outlet_boundaries() const
  {
    std::vector<unsigned> out;
    for (const auto &bc : boundary_conditions)
      if (bc.type == "outlet")
        out.push_back(bc.boundary_id);
    return out;
  }

  double DEMParameters::smallest_contact_search_criterion() const
  {
    const double d = maximum_particle_diameter();
    return std::min(mesh.minimal_cell_diameter() - d * 0.5, dynamic_contact_search_factor * (neighborhood_threshold - 1) * d * 0.5);
  }

  lethe_dem_config DEMParameters::to_config(bool store_forces) const
  {
    if (integration_method != "velocity_verlet" && integration_method != "explicit_euler")
      throw std::runtime_error("unknown integration method `" + integration_method + "` (velocity_verlet|explicit_euler)");
    if (solver_type != "dem")
      throw std::runtime_error("solver type dem_mp is out of scope");
    static const std::map<std::string, int> pp{{"linear", LETHE_PP_LINEAR},
                                               {"hertz_mindlin_limit_force", LETHE_PP_HERTZ_MINDLIN_LIMIT_FORCE},
                                               {"hertz_mindlin_limit_overlap", LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP},
                                               {"hertz", LETHE_PP_HERTZ},
                                               {"hertz_JKR", LETHE_PP_HERTZ_JKR},
                                               {"DMT", LETHE_PP_DMT}};
    static const std::map<std::string, int> pw{
      {"linear", LETHE_PW_LINEAR}, {"nonlinear", LETHE_PW_NONLINEAR}, {"JKR", LETHE_PW_JKR}, {"DMT", LETHE_PW_DMT}};
    static const std::map<std::string, int> rolling{{"none", LETHE_ROLLING_NONE},
                                                    {"constant", LETHE_ROLLING_CONSTANT},
                                                    {"viscous", LETHE_ROLLING_VISCOUS},
                                                    {"epsd", LETHE_ROLLING_EPSD}};
    static const std::map<std::string, int> detection{{"dynamic", LETHE_DETECTION_DYNAMIC}, {"constant", LETHE_DETECTION_CONSTANT}};
    lethe_dem_config c;
    std::memset(&c, 0, sizeof(c));
    c.pp_model = lookup(pp, pp_model, "particle particle contact force method");
    c.pw_model = lookup(pw, pw_model, "particle wall contact force method");
    c.rolling_model = lookup(rolling, rolling_model, "rolling resistance torque method");
    c.integrator = integration_method == "explicit_euler" ? LETHE_INTEGRATOR_EXPLICIT_EULER : LETHE_INTEGRATOR_VELOCITY_VERLET;
    c.detection = lookup(detection, contact_detection_method, "contact detection method");
    c.contact_detection_frequency = int(contact_detection_frequency);
    c.cell_order = mesh.cell_order;
    c.store_forces = store_forces ? 1 : 0;
    c.dt = time_step;
    c.neighborhood_threshold = neighborhood_threshold;
    c.d_max = maximum_particle_diameter();
    c.smallest_contact_search_criterion = smallest_contact_search_criterion();
    c.dmt_cut_off_threshold = dmt_cut_off_threshold;
    c.f_coefficient_epsd = f_coefficient;
    c.moi_override = 0;
    c.n_types = int(particle_types.size());
    c.restart = restart ? 1 : 0;
    for (size_t i = 0; i < particle_types.size(); ++i)
      {
        const ParticleType &t = particle_types[i];
        c.young[i] = t.young_modulus;
        c.poisson[i] = t.poisson_ratio;
        c.restitution[i] = t.restitution_coefficient;
        c.friction[i] = t.friction_coefficient;
        c.rolling_friction[i] = t.rolling_friction;
        c.rolling_viscous_damping[i] = t.rolling_viscous_damping;
        c.surface_energy[i] = t.surface_energy;
        c.hamaker[i] = t.hamaker_constant;
      }
    c.young_wall = young_wall;
    c.poisson_wall = poisson_wall;
    c.restitution_wall = restitution_wall;
    c.friction_wall = friction_wall;
    c.rolling_friction_wall = rolling_friction_wall;
    c.rolling_viscous_damping_wall = rolling_viscous_damping_wall;
    c.surface_energy_wall = surface_energy_wall;
    c.hamaker_wall = hamaker_wall;
    const Vec3 h = mesh.cell_size();
    const auto per = periodic_directions();
    for (int d = 0; d < 3; ++d)
      {
        c.g[d] = g[d];
        c.grid_lo[d] = mesh.lo[d];
        c.cell_size[d] = h[d];
        c.grid_n[d] = mesh.n[d];
        c.periodic[d] = per[d];
      }
    c.slab_axis = -1;
    c.sparse_contacts = sparse_contacts ? 1 : 0;
    c.asc_granular_temperature_threshold = asc_granular_temperature_threshold;
    c.asc_solid_fraction_threshold = asc_solid_fraction_threshold;
    return c;
  }
} // namespace lethe_b200
