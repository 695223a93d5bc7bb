// dem_parameters.h — the subset of DEMSolverParameters<3>
// (include/dem/dem_solver_parameters.h:15-92) that the DEM hot path honours, read from a
// `.prm` with the reference's subsection names, key spellings and defaults
// (source/core/parameters_lagrangian.cc:13-215,918-1337,1405-1742; `simulation control`,
// `mesh`, `test`, `restart` from source/core/parameters.cc), and its translation to the
// lethe_dem_config that crosses the C ABI.
#pragma once

#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/lethe_dem.h"
#include "prm_file.h"

namespace lethe_b200
{
  using Vec3 = std::array<double, 3>;

  // Parameters::Lagrangian::LagrangianPhysicalProperties, one entry per `particle type N`
  struct ParticleType
  {
    std::string size_distribution_type = "uniform";
    double average_diameter = 0.001;
    double standard_deviation = 0.0;
    long number_of_particles = 0;
    double density = 1000.0;
    double young_modulus = 1e6;
    double poisson_ratio = 0.3;
    double restitution_coefficient = 0.1;
    double friction_coefficient = 0.1;
    double rolling_friction = 0.1;
    double rolling_viscous_damping = 0.1;
    double surface_energy = 0.0;
    double hamaker_constant = 4e-19;
    long prn_seed = 1;
    double min_cutoff = -1, max_cutoff = -1; // `minimum / maximum diameter cutoff` (< 0: +-2.5 sigma)
    // find_min_diameter / find_max_diameter of the type's Distribution (distributions.cc)
    double max_diameter() const;
    double min_diameter() const;
  };

  // Uniform hex grid equivalent of the deal.II triangulation (`subsection mesh`)
  struct Mesh
  {
    Vec3 lo{{0, 0, 0}}, hi{{1, 1, 1}};
    std::array<int, 3> n{{1, 1, 1}};
    bool colorize = false;
    int cell_order = LETHE_CELL_ORDER_LEXICOGRAPHIC;
    bool expand_particle_wall_contact_search = false;
    Vec3 cell_size() const { return {{(hi[0] - lo[0]) / n[0], (hi[1] - lo[1]) / n[1], (hi[2] - lo[2]) / n[2]}}; }
    double minimal_cell_diameter() const; // GridTools::minimal_cell_diameter
  };

  // Parameters::Lagrangian::BCDEM, one entry per `boundary condition N`
  struct BoundaryCondition
  {
    std::string type = "fixed_wall";
    unsigned boundary_id = 0;
    double rotational_speed = 0;
    Vec3 rotational_vector{{1, 0, 0}}, point_on_rotational_vector{{0, 0, 0}}, translational_velocity{{0, 0, 0}};
    unsigned periodic_id_0 = 0, periodic_id_1 = 0;
    int periodic_direction = 0;
  };

  struct FloatingWall
  {
    Vec3 point{{0, 0, 0}}, normal{{0, 0, 0}};
    double time_start = 0, time_end = 0;
  };

  // `subsection solid objects / solid surfaces / solid object N` (Parameters::RigidSolidObject,
  // SerialSolid<2,3>): gmsh triangle surface, initial rotation / translation, constant velocities
  struct SolidSurface
  {
    std::string mesh_file;
    Vec3 rotation_axis{{1, 0, 0}};
    double rotation_angle = 0;
    Vec3 translation{{0, 0, 0}}, translational_velocity{{0, 0, 0}}, angular_velocity{{0, 0, 0}}, center_of_rotation{{0, 0, 0}};
  };

  // Parameters::Lagrangian::InsertionInfo (volume and list methods)
  struct InsertionInfo
  {
    std::string method = "volume";
    // insertion_list.cc: explicit positions / velocities / diameters
    std::vector<double> list_x, list_y, list_z, list_vx, list_vy, list_vz, list_wx, list_wy, list_wz, list_diameters;
    std::vector<std::string> input_files; // `insertion method = file`: `list of input files`
    long inserted_this_step = 0;
    long frequency = 1;
    Vec3 box_point_1{{0, 0, 0}}, box_point_2{{1, 1, 1}};
    double distance_threshold = 1.0;
    double maximum_offset = 1.0;
    long prn_seed = 1;
    std::array<int, 3> direction_sequence{{0, 1, 2}};
    Vec3 initial_velocity{{0, 0, 0}}, initial_omega{{0, 0, 0}};
  };

  struct DEMParameters
  {
    // simulation control
    double time_step = 1.0, time_end = 1.0;
    long log_frequency = 1, output_frequency = 1;
    // model parameters
    std::string contact_detection_method = "dynamic";
    long contact_detection_frequency = 1;
    double dynamic_contact_search_factor = 0.8;
    double neighborhood_threshold = 1.3;
    std::string pp_model = "hertz_mindlin_limit_overlap";
    std::string pw_model = "nonlinear";
    std::string rolling_model = "constant";
    std::string integration_method = "velocity_verlet";
    std::string solver_type = "dem";
    double dmt_cut_off_threshold = 0.1;
    double f_coefficient = 0.0;
    // lagrangian physical properties
    Vec3 g{{0, 0, 0}};
    std::vector<ParticleType> particle_types{ParticleType()};
    double young_wall = 1e6, poisson_wall = 0.3, restitution_wall = 0.1, friction_wall = 0.1;
    double rolling_friction_wall = 0.1, rolling_viscous_damping_wall = 0.1, surface_energy_wall = 0.0, hamaker_wall = 4e-19;
    Mesh mesh;
    InsertionInfo insertion;
    std::vector<BoundaryCondition> boundary_conditions;
    std::vector<FloatingWall> floating_walls;
    std::vector<SolidSurface> solid_surfaces;
    std::string prm_directory = "."; // mesh file names are relative to the parameter file
    bool restart = false;
    bool test_enabled = false;

    static DEMParameters from_prm(const PrmSection &root);
    static DEMParameters from_prm_file(const std::string &path);

    double maximum_particle_diameter() const;           // dem.cc:149-159
    std::array<int, 3> periodic_directions() const;     // DEM boundary conditions of type periodic
    std::vector<unsigned> outlet_boundaries() const;
    double smallest_contact_search_criterion() const;   // dem.cc:289-294
    lethe_dem_config to_config(bool store_forces = false) const;
  };
} // namespace lethe_b200
