// dem_solver.h — host-side mirror of `DEMSolver<3, DEM::DEMProperties::PropertiesIndex>`
// (include/dem/dem.h, source/dem/dem.cc) with the per-step work behind the C ABI.
//
// What stays on the host is what the reference keeps on the host too: parameter set-up,
// the wall table (BoundaryCellsInformation for box meshes), insertion, the time loop's
// bookkeeping (when to insert, log, stop) and the `test` output. One iteration of
// `while (simulation_control->integrate())` — contact detection + search, pp/pw forces,
// integration, trigger reset (dem.cc:1115-1246) — is `DEMEngine::step`.
#pragma once

#include <iosfwd>
#include <memory>
#include <random>

#include "dem_engine.h"
#include "dem_parameters.h"

namespace lethe_b200
{
  // BoundaryCellsInformation::find_boundary_cells_information for a uniform box mesh
  // (find_boundary_cells_information.cc:130-219): one row per (boundary cell, boundary face)
  // that is neither an outlet nor periodic; inward normal, face centre as the point.
  std::vector<lethe_wall_face> box_wall_faces(const Mesh &mesh, const std::vector<unsigned> &outlet_boundaries,
                                              const std::array<int, 3> &periodic);

  // InsertionVolume::insert + assign_particle_properties on one rank, uniform diameters
  // (insertion_volume.cc:43-206, insertion.cc:60-121); jitter from glibc rand() exactly as
  // create_random_number_container does (include/core/utilities.h:1061-1073).
  // Distribution::particle_size_sampling (distributions.cc): uniform, normal, lognormal with the
  // reference's generator (std::mt19937(seed + rank), std::normal_distribution / lognormal_distribution)
  class SizeDistribution
  {
  public:
    SizeDistribution(const ParticleType &t, unsigned rank);
    std::vector<double> sample(long n);

  private:
    ParticleType type;
    std::mt19937 gen;
    std::normal_distribution<> normal;
    std::lognormal_distribution<> lognormal;
  };
  ParticleRows volume_insertion(const DEMParameters &p, long n_insert, uint32_t first_id, int particle_type, SizeDistribution &sizes);
  // (i, j, k) of the uniform grid's cells in deal.II's active-cell order: lexicographic for an
  // unrefined subdivided grid, the z-order curve after refine_global
  std::vector<std::array<int, 3>> active_cell_order(const Mesh &mesh);
  // InsertionPlane (insertion_plane.cc): one particle in every cell the plane cuts that holds no
  // particle, at the cell centre plus rand() * maximum offset / RAND_MAX per axis
  class PlaneInsertion
  {
  public:
    explicit PlaneInsertion(const DEMParameters &p);
    // `occupied`: linear index (i + nx (j + ny k)) -> whether particles are registered in the cell
    ParticleRows insert(const DEMParameters &p, const std::vector<char> &occupied, long remaining, uint32_t first_id, int particle_type,
                        SizeDistribution &sizes);

  private:
    std::vector<std::array<int, 3>> cells; // find_inplane_cells, in active-cell order
    double maximum_range_for_randomness = 0;
  };
  // InsertionList::insert (insertion_list.cc): the listed positions / velocities / diameters
  ParticleRows list_insertion(const DEMParameters &p, uint32_t first_id, int particle_type);
  // InsertionFile::insert (insertion_file.cc:27-130): one `;`-separated table per insertion
  ParticleRows file_insertion(const DEMParameters &p, const std::string &path, long n_max, uint32_t first_id, int particle_type);
  // GridIn::read_msh for a triangle surface (gmsh 4.1 / 2.2 ASCII): vertices in node order,
  // triangles in element order (SerialSolid::setup_triangulation, serial_solid.cc:163-175)
  void read_msh_triangles(const std::string &path, std::vector<double> &vertices3, std::vector<uint32_t> &triangles3);
  // `type = dealii`, `simplex = true` solid surfaces (serial_solid.cc:176-196): hyper_cube /
  // hyper_rectangle in the z = 0 plane, refined, every quadrilateral split into 8 triangles
  void dealii_simplex_surface(const std::string &grid_type, const std::string &grid_arguments, long initial_refinement,
                              std::vector<double> &vertices3, std::vector<uint32_t> &triangles3);

  class DEMSolverB200
  {
  public:
    DEMSolverB200(const DEMParameters &parameters, int device, std::ostream &log);
    // DEMSolver::solve (dem.cc:1064-1267)
    void solve();
    // finish_simulation with `subsection test / set enable = true`: print_xyz (dem.cc:760-770)
    void print_xyz(std::ostream &out);
    DEMEngine &get_engine() { return *engine; }

  private:
    void setup_boundaries();     // setup_functions_and_pointers + boundary_cell_object.build
    bool insertion_due() const;  // insert_particles (dem.cc:484-506)
    void insert_particles();
    void remove_particles_in_box(); // Insertion::remove_particles_in_box (insertion.cc:132-260)
    long cell_of(const double *x) const; // linear index of the cell around a point, -1 outside
    bool is_at_end() const;      // SimulationControlTransient::is_at_end
    bool is_verbose_iteration() const { return (iteration_number % parameters.log_frequency) == 0; }
    void print_progression();    // SimulationControlTransientDEM::print_progression
    void report_statistics();    // dem.cc:902-971

    DEMParameters parameters;
    std::unique_ptr<DEMEngine> engine;
    std::ostream &pcout;
    unsigned long iteration_number = 0;
    double current_time = 0;
    std::vector<long> remaining_particles; // per type
    int current_inserting_type = 0;
    uint32_t next_id = 0;
    size_t current_file_id = 0;
    std::vector<SizeDistribution> size_distributions; // one per particle type (setup_distributions)
    // plane insertion asks which cells hold particles, i.e. where the last sort registered them
    std::unique_ptr<PlaneInsertion> plane_insertion;
    std::vector<char> occupied_cells;
    std::vector<long> registered_cell; // particle id -> linear cell index of the last sort (-1: none)
    std::vector<std::pair<Vec3, Vec3>> solid_motion; // last velocities handed to the engine
    // contact_list statistics of report_statistics
    double list_min = 1e300, list_max = 0, list_total = 0;
  };
} // namespace lethe_b200
