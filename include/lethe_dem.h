/* lethe_dem.h — C ABI of the B200-native DEM time-step engine.
 *
 * This is the drop-in boundary for the `lethe-particles` hot path of
 * chaos-polymtl/lethe (reference citations are relative to /root/reference):
 *
 *   reference interface                                   replaced by
 *   ----------------------------------------------------  ---------------------------
 *   DEMSolver ctor + setup_parameters / factories          lethe_dem_create
 *     (source/dem/dem.cc:32-58,60-138,196-281;
 *      set_particle_particle_contact_force_model.cc:12-103;
 *      set_particle_wall_contact_force_model.cc:12-…)
 *   ParticleHandler insertion result (id, x, props[9])     lethe_dem_set_particles /
 *     (source/dem/insertion.cc:60-121,                       lethe_dem_add_particles
 *      include/core/dem_properties.h:54-76)
 *   BoundaryCellsInformation::build                        lethe_dem_set_walls /
 *     (find_boundary_cells_information.cc:26-94,130-219)     lethe_dem_set_floating_walls
 *   DEM boundary conditions (rotational / translational)   lethe_dem_set_boundary_motion
 *     (particle_wall_contact_force.cc:602-610,
 *      particle_wall_contact_force.h:199-246)
 *   one iteration of the `while (simulation_control->      lethe_dem_step
 *     integrate())` loop: execute_contact_detection_and_
 *     search + compute_contact_forces + integrate
 *     (source/dem/dem.cc:1115-1183,598-717)
 *   DEMSolver::synchronize_velocities (dem.cc:719-745)     lethe_dem_synchronize_velocities
 *   DEMActionManager::particle_insertion_step /            lethe_dem_force_contact_search
 *     restart_simulation (dem_action_manager.h:191-232)
 *   report_statistics reductions (dem.cc:902-971)          lethe_dem_get_stats
 *   print_xyz / VTU snapshot (dem.cc:760-770)              lethe_dem_get_particles
 *
 * All arithmetic is IEEE FP64; indices are 32-bit. Every function returns 0 on
 * success and a negative code on error; lethe_dem_last_error() gives the text.
 * There is no CPU fallback: creating a context without a CUDA device fails.
 *
 * The oracle (oracle/dem_oracle.cpp, test infrastructure only) exports the same
 * functions with the prefix `oracle_dem_` and the same structs, so that parity
 * tests drive both through identical calls.
 */
#ifndef LETHE_DEM_H
#define LETHE_DEM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LETHE_DEM_MAX_TYPES 5            /* parameters_lagrangian.h:325 */
#define LETHE_DEM_MAX_FLOATING_WALLS 9   /* parameters_lagrangian.cc:1405-1575 */
#define LETHE_DEM_MAX_BOUNDARY_MOTIONS 16
#define LETHE_DEM_N_PROPERTIES 9         /* dem_properties.h:54-76 */
#define LETHE_DEM_FLOATING_WALL_BOUNDARY_ID 100 /* particle_wall_fine_search.cc:150 */

/* Parameters::Lagrangian::ParticleParticleContactForceModel
 * (include/core/parameters_lagrangian.h:33-86) */
enum lethe_pp_model {
  LETHE_PP_LINEAR = 0,
  LETHE_PP_HERTZ_MINDLIN_LIMIT_FORCE = 1,
  LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP = 2,
  LETHE_PP_HERTZ = 3,
  LETHE_PP_HERTZ_JKR = 4,
  LETHE_PP_DMT = 5
};
/* ParticleWallContactForceModel */
enum lethe_pw_model {
  LETHE_PW_LINEAR = 0,
  LETHE_PW_NONLINEAR = 1,
  LETHE_PW_JKR = 2,
  LETHE_PW_DMT = 3
};
/* RollingResistanceMethod */
enum lethe_rolling_model {
  LETHE_ROLLING_NONE = 0,
  LETHE_ROLLING_CONSTANT = 1,
  LETHE_ROLLING_VISCOUS = 2,
  LETHE_ROLLING_EPSD = 3
};
/* `integration method = velocity_verlet|explicit_euler` (parameters_lagrangian.cc:1049,1314;
 * DEMSolver::set_integrator_type, dem.cc:260-279) */
enum lethe_integrator { LETHE_INTEGRATOR_VELOCITY_VERLET = 0, LETHE_INTEGRATOR_EXPLICIT_EULER = 1 };
enum lethe_detection { LETHE_DETECTION_DYNAMIC = 0, LETHE_DETECTION_CONSTANT = 1 };
/* Order in which the reference's FE mesh enumerates its active cells; it only
 * fixes which particle of a pair is "particle one" (history sign) and the
 * summation order of the oracle. 0: lexicographic (subdivided_hyper_rectangle),
 * 1: hierarchical z-order (hyper_cube + refine_global). */
enum lethe_cell_order { LETHE_CELL_ORDER_LEXICOGRAPHIC = 0, LETHE_CELL_ORDER_MORTON = 1 };

typedef struct lethe_dem_config {
  /* model selectors (`subsection model parameters`, parameters_lagrangian.cc:918-1337) */
  int32_t pp_model;
  int32_t pw_model;
  int32_t rolling_model;
  int32_t integrator;
  int32_t detection;                   /* contact detection method */
  int32_t contact_detection_frequency; /* `frequency` */
  int32_t cell_order;                  /* enum lethe_cell_order */
  int32_t store_forces;                /* debug tap: keep last-step F/T for lethe_dem_get_forces */
  double dt;                           /* `time step` */
  double g[3];                         /* `g` */
  double neighborhood_threshold;       /* `neighborhood threshold` (1.3) */
  double d_max;                        /* maximum_particle_diameter (dem.cc:149-159) */
  /* min(minimal_cell_diameter - d_max/2, coeff*(thr-1)*d_max/2), dem.cc:289-294;
   * computed by the host because it contains the FE mesh's minimal cell diameter. */
  double smallest_contact_search_criterion;
  double dmt_cut_off_threshold;        /* `dmt cut-off threshold` */
  double f_coefficient_epsd;           /* `f coefficient` */
  /* Test hook mirroring the reference's unit tests, which force MOI = 1
   * (tests/dem/full_contact_functions.h:135-141). <= 0: MOI = 0.1 m d^2 (dem.cc:1004-1011). */
  double moi_override;
  /* `subsection lagrangian physical properties` raw tables; effective pair
   * tables are derived with the reference formulas
   * (particle_particle_contact_force.h:1639-1746, particle_wall_contact_force.cc:588-694) */
  int32_t n_types;
  /* `subsection restart / set restart`: a restarted run resumes with regular
   * integrate() steps instead of integrate_start (dem.cc:1162-1171). */
  int32_t restart;
  double young[LETHE_DEM_MAX_TYPES];
  double poisson[LETHE_DEM_MAX_TYPES];
  double restitution[LETHE_DEM_MAX_TYPES];
  double friction[LETHE_DEM_MAX_TYPES];
  double rolling_friction[LETHE_DEM_MAX_TYPES];
  double rolling_viscous_damping[LETHE_DEM_MAX_TYPES];
  double surface_energy[LETHE_DEM_MAX_TYPES];
  double hamaker[LETHE_DEM_MAX_TYPES];
  double young_wall, poisson_wall, restitution_wall, friction_wall;
  double rolling_friction_wall, rolling_viscous_damping_wall;
  double surface_energy_wall, hamaker_wall;
  /* The reference's FE mesh when it is a uniform hex grid (hyper_cube /
   * subdivided_hyper_rectangle): particles are binned into these cells and
   * candidates are particles in vertex-sharing cells (find_cell_neighbors.cc:10-104). */
  double grid_lo[3];
  double cell_size[3];
  int32_t grid_n[3];
  int32_t periodic[3];                 /* periodic direction flags (DEM boundary conditions) */
  /* slab decomposition (multi-GPU): this context owns grid cells
   * [slab_lo, slab_hi) along slab_axis; -1/0/0 = whole grid. */
  int32_t slab_axis;
  int32_t slab_lo;
  int32_t slab_hi;
  /* `subsection adaptive sparse contacts` (parameters_lagrangian.cc; AdaptiveSparseContacts::
   * set_parameters, adaptive_sparse_contacts.h:176-192): != 0 enables the per-cell mobility status
   * (identify_mobility_status, adaptive_sparse_contacts.cc:132-356) at every contact search, the
   * status-aware broad searches (particle_particle_broad_search.cc:134-316,
   * particle_wall_broad_search.cc:212-336) and integration (velocity_verlet_integrator.cc:117-210,
   * 292-436). `advect particles` (CFD-DEM) is not supported. */
  int32_t sparse_contacts;
  double asc_granular_temperature_threshold; /* `granular temperature threshold` */
  double asc_solid_fraction_threshold;       /* `solid fraction threshold` */
  /* Arithmetic of the particle-particle contact model (enum lethe_precision). The reference is
   * FP64 throughout; LETHE_PRECISION_MIXED is this library's documented-bound fast mode:
   * positions, overlaps, relative velocities, force accumulation and integration stay FP64, the
   * contact model between them runs in FP32. Particle-wall and solid-surface contacts stay FP64. */
  int32_t precision;
  int32_t pad2;
} lethe_dem_config;

enum lethe_precision { LETHE_PRECISION_F64 = 0, LETHE_PRECISION_MIXED = 1 };

/* AdaptiveSparseContacts::mobility_status of a cell (adaptive_sparse_contacts.h:154-162) */
enum lethe_mobility_status {
  LETHE_MOBILITY_INACTIVE = 0,
  LETHE_MOBILITY_STATIC_ACTIVE = 1,
  LETHE_MOBILITY_MOBILE = 4,
  LETHE_MOBILITY_EMPTY_NODE = 5 /* nodes only */
};

/* One row of boundary_cells_info_struct (include/dem/boundary_cells_info_struct.h:20-38):
 * a boundary face of grid cell `cell` (lexicographic index ix + nx*(iy + ny*iz)),
 * treated as an infinite plane through `point` with inward unit `normal`. */
typedef struct lethe_wall_face {
  int32_t cell;
  uint32_t boundary_id;
  uint32_t global_face_id;
  uint32_t pad;
  double normal[3];
  double point[3];
} lethe_wall_face;

/* min / max / sum over particles, as printed by report_statistics (dem.cc:902-971) */
typedef struct lethe_dem_stats {
  uint64_t n_particles;
  uint64_t n_rebuilds;          /* contact_build_number */
  uint64_t n_steps;
  uint64_t n_pair_entries;      /* unordered pairs in the contact list */
  uint64_t n_wall_entries;
  uint64_t n_pairs_touching;    /* pairs with overlap > threshold at the last step (store_forces) */
  double v_min, v_max, v_sum;          /* |v| */
  double omega_min, omega_max, omega_sum;
  double ke_trans_min, ke_trans_max, ke_trans_sum;
  double ke_rot_min, ke_rot_max, ke_rot_sum;
  uint64_t n_migrated;          /* multi-GPU: particles this rank has sent to or received from its neighbours so far */
} lethe_dem_stats;

typedef struct lethe_dem_ctx lethe_dem_ctx;

/* --- lifetime --- */
int lethe_dem_create(const lethe_dem_config *config, int device, lethe_dem_ctx **out);
void lethe_dem_destroy(lethe_dem_ctx *ctx);
const char *lethe_dem_last_error(const lethe_dem_ctx *ctx);
/* Text of the error of the last failed lethe_dem_create (no ctx exists then). */
const char *lethe_dem_create_error(void);

/* --- state in / out (host buffers; rows laid out exactly as PropertiesIndex) --- */
int lethe_dem_set_particles(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id,
                            const double *x3, const double *props9);
int lethe_dem_add_particles(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id,
                            const double *x3, const double *props9);
int lethe_dem_n_particles(lethe_dem_ctx *ctx, uint64_t *n);
/* Fills up to n_max rows sorted by particle id; *n_out = number written. */
int lethe_dem_get_particles(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out,
                            uint32_t *id, double *x3, double *props9);
int lethe_dem_set_walls(lethe_dem_ctx *ctx, uint64_t n_faces, const lethe_wall_face *faces);
int lethe_dem_set_floating_walls(lethe_dem_ctx *ctx, int32_t n, const double *point3,
                                 const double *normal3, const double *t_start,
                                 const double *t_end);
int lethe_dem_set_boundary_motion(lethe_dem_ctx *ctx, uint32_t boundary_id,
                                  const double translational_velocity[3],
                                  double rotational_speed, const double rotational_vector[3],
                                  const double point_on_rotation_axis[3]);
/* Solid surfaces (triangle meshes; `subsection solid objects / solid surfaces`, SerialSolid<2,3>,
 * source/core/serial_solid.cc). Replaces DEMSolver::setup_solid_objects (dem.cc:164-191); the
 * per-step motion (move_solid_triangulation, serial_solid.cc:333-410), the background-mesh
 * mapping (map_solid_in_background_triangulation, :83-150, refreshed when a vertex has moved more
 * than 3^-1/2 of the cell diameter, find_contact_detection_step.cc:139-161), the candidate search
 * (particle_wall_broad_search.cc:129-215) and calculate_particle_solid_object_contact
 * (particle_wall_contact_force.cc:153-580, incl. the face/edge/vertex double-contact elimination)
 * then run inside lethe_dem_step. `triangles3` holds 3 vertex indices per triangle, in mesh
 * element order. Velocities are the values of the reference's velocity functions for the
 * coming steps; update them with lethe_dem_set_solid_motion when they depend on time. */
int lethe_dem_add_solid_surface(lethe_dem_ctx *ctx, uint32_t n_vertices, const double *vertices3,
                                uint32_t n_triangles, const uint32_t *triangles3,
                                const double translational_velocity[3],
                                const double angular_velocity[3],
                                const double center_of_rotation[3], int32_t *solid_index);
int lethe_dem_set_solid_motion(lethe_dem_ctx *ctx, int32_t solid_index,
                               const double translational_velocity[3],
                               const double angular_velocity[3]);
/* current vertex positions of a solid (3 doubles per vertex) */
int lethe_dem_get_solid_vertices(lethe_dem_ctx *ctx, int32_t solid_index, uint32_t n_max,
                                 double *vertices3);
/* debug tap: (particle id, solid, triangle) candidates in (id, solid, triangle) order + history */
int lethe_dem_get_solid_contacts(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out,
                                 uint32_t *particle_id, uint32_t *solid, uint32_t *triangle,
                                 double *tangential3);

/* --- the hot path --- */
int lethe_dem_step(lethe_dem_ctx *ctx, uint64_t n_steps);
int lethe_dem_synchronize_velocities(lethe_dem_ctx *ctx);
int lethe_dem_force_contact_search(lethe_dem_ctx *ctx, int clear_tangential_displacement);
/* The second consumer of the path, the CFD-DEM coupling (fem-dem/cfd_dem_coupling.cc:1380-1540):
 * - lethe_dem_set_external_loads = add_fluid_particle_interaction_force / _torque (:881-925):
 *   force3 / torque3 ([n][3], torque3 may be NULL) of the listed particles are added to the
 *   contact forces of every following step and of lethe_dem_synchronize_velocities, after the
 *   contact forces and before the integration, until set again; n = 0 clears all loads. In a
 *   slab-decomposed job every rank sets the loads of the particles it owns.
 * - lethe_dem_restart_integration: the next step is an opening step of the scheme again
 *   (integrate_start). CFD-DEM synchronises the velocities at the end of every CFD time step
 *   (lethe_dem_synchronize_velocities) and opens the next one this way (:1420-1431). */
int lethe_dem_set_external_loads(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id,
                                 const double *force3, const double *torque3);
int lethe_dem_restart_integration(lethe_dem_ctx *ctx);
/* Reference-facing per-step call with HOST buffers (what a patched DEMSolver
 * that keeps ParticleHandler on the host would call every iteration): uploads
 * x/props for the n particles (same ids as resident), runs n_steps, downloads
 * the updated rows into the same buffers. Contact history stays resident. */
int lethe_dem_step_host(lethe_dem_ctx *ctx, uint64_t n_steps, uint64_t n,
                        const uint32_t *id, double *x3, double *props9);
/* The same call moving only what a step changes: state9 = [n][9] rows of
 * x, y, z, v_x, v_y, v_z, omega_x, omega_y, omega_z (what Integrator::integrate writes,
 * velocity_verlet_integrator.cc:214-290); type, diameter and mass keep the values given at
 * insertion. id = the particle of every row; NULL = the table of the previous call (same n),
 * which is then not uploaded again. 72 bytes per particle each way instead of 100 / 96.
 * With n_steps = 1 on one GPU the call is STREAMED when nothing but the plain step runs in it: the rows go up in the
 * caller's order in a few contiguous stages, the step kernel is launched on the blocks of
 * particles whose own rows and listed neighbours have arrived, and rows go back down as soon as their blocks are
 * stepped, so the download runs under the upload (PCIe is full duplex). Results are bit-identical to the plain
 * call. LETHE_DEM_HOST_PIPELINE=0 switches it off, LETHE_DEM_HOST_STAGES (8) and LETHE_DEM_HOST_PIPELINE_MIN_ROWS
 * (262144) tune it. Page-locked host rows are needed for the copies to overlap. LETHE_DEM_HOST_ZEROCOPY=1 lets the
 * device write page-locked rows itself, block by block as the step finishes them (slower than the copy engines on the
 * PCIe Gen5 boxes measured; off by default). */
int lethe_dem_step_host_state(lethe_dem_ctx *ctx, uint64_t n_steps, uint64_t n,
                              const uint32_t *id, double *state9);
/* The ids of the owned particles in the row order that lets the streamed call overlap its copies best: by cell layer
 * along the axis with the most layers, cell-sorted inside a layer (the kind of order a cell-by-cell walk of
 * ParticleHandler gives). With rows in this order a block of particles is complete one or two layers after its own rows
 * have arrived, so upload and download run side by side almost from the first stage. Any other order is still correct.
 * A host re-reads it when it re-reads its rows (after insertion, or when particles changed owner). */
int lethe_dem_get_transfer_order(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *id);
/* The same order together with the rows themselves: id[k] and state9[k] = (x, v, omega) of the owned particles, ready to be
 * handed back to lethe_dem_step_host_state. What a host calls to (re)read the rows it keeps — at the start and whenever
 * lethe_dem_stats.n_migrated shows that particles changed owner. */
int lethe_dem_get_state_rows(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *state9);
/* How often lethe_dem_step_host_state took the streamed form, how many times its plan was made (once per list
 * rebuild or new id table) and how many of the streamed calls wrote the host rows directly from the device. */
int lethe_dem_host_pipeline_stats(lethe_dem_ctx *ctx, uint64_t *n_streamed_calls, uint64_t *n_plans, uint64_t *n_direct_calls);

/* The CFD-DEM particle record (DEM::CFDDEMProperties::PropertiesIndex, include/core/dem_properties.h:92-142:
 * 23 doubles per particle — the 9 DEM properties, then fem_force_two_way_coupling[3], fem_force_one_way_
 * coupling[3], fem_drag[3], fem_torque[3], volumetric_contribution, momentum_transfer_coefficient).
 * - lethe_dem_set_particles_cfd: lethe_dem_set_particles with rows of 23; the fluid loads of every particle
 *   become its external loads exactly as add_fluid_particle_interaction_force / _torque add them
 *   (cfd_dem_coupling.cc:881-925): force = (two_way + one_way) + drag per component, torque = fem_torque.
 * - lethe_dem_update_loads_cfd: the loads alone, from rows of 23 (what changes at every CFD time step).
 * - lethe_dem_get_particles_cfd: rows sorted by id; properties 0-8 are written, 9-22 are left as the caller
 *   holds them (they belong to the fluid solver). */
#define LETHE_DEM_N_CFD_PROPERTIES 23
int lethe_dem_set_particles_cfd(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id, const double *x3, const double *props23);
int lethe_dem_update_loads_cfd(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id, const double *props23);
int lethe_dem_get_particles_cfd(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *x3, double *props23);

/* DEM-MP heat transfer (`solver type = dem_mp`): particle-particle conduction through the
 * contact (calculate_contact_thermal_conductance, source/dem/particle_heat_transfer.cc:9-318, called from
 * execute_contact_calculation, particle_particle_contact_force.h:2065-2150, for every pair with a positive
 * overlap) and explicit temperature integration (integrate_temperature, multiphysics_integrator.cc:6-34).
 * The two extra properties of DEM::DEMMPProperties (T = 9, specific_heat = 10, dem_properties.h:150-166) are
 * kept per particle id. Raw per-type properties as in `subsection lagrangian physical properties`; the
 * effective pair tables follow set_multiphysic_properties (…contact_force.h:1755-1826). Heat exchange with
 * solid surfaces and heat sources are not built; single GPU. */
typedef struct lethe_dem_thermal_properties {
  double real_youngs_modulus[LETHE_DEM_MAX_TYPES];
  double surface_roughness[LETHE_DEM_MAX_TYPES];
  double surface_slope[LETHE_DEM_MAX_TYPES];
  double microhardness[LETHE_DEM_MAX_TYPES];
  double thermal_conductivity[LETHE_DEM_MAX_TYPES];
  double thermal_accommodation[LETHE_DEM_MAX_TYPES];
  double thermal_conductivity_gas;
  double dynamic_viscosity_gas;
  double specific_heat_gas;
  double specific_heats_ratio_gas;
  double molecular_mean_free_path_gas;
} lethe_dem_thermal_properties;
int lethe_dem_enable_heat_transfer(lethe_dem_ctx *ctx, const lethe_dem_thermal_properties *properties);
/* PropertiesIndex::T and ::specific_heat of the listed particles (set at insertion by the reference). */
int lethe_dem_set_temperatures(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id, const double *temperature,
                               const double *specific_heat);
/* Rows sorted by id: temperature now, and contact_outcome.heat_transfer_rate of the last step (J/s). */
int lethe_dem_get_temperatures(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *temperature,
                               double *heat_transfer_rate);

/* Restart (read_checkpoint.cc:14-130: simulation_control->read(prefix) restores the iteration number
 * and the time; DEMActionManager::restart_simulation, dem_action_manager.h:185-200, triggers the
 * contact search and clears every tangential history): a context created with config.restart = 1
 * and given the checkpointed particles resumes at this iteration / time with regular
 * integrate() steps. The contact history is not part of a reference checkpoint and starts at zero. */
int lethe_dem_set_time(lethe_dem_ctx *ctx, uint64_t iteration_number, double current_time);

/* --- debug taps / statistics --- */
/* Unordered pairs (i_id < j_id) of the contact list with the tangential
 * displacement oriented i -> j. */
int lethe_dem_get_pairs(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *i_id,
                        uint32_t *j_id, double *tangential3);
int lethe_dem_get_wall_contacts(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out,
                                uint32_t *particle_id, uint32_t *face_id, double *tangential3);
/* Force / torque applied to each particle in the last step (requires
 * config.store_forces); rows sorted by particle id. */
int lethe_dem_get_forces(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *id,
                         double *force3, double *torque3);
int lethe_dem_get_stats(lethe_dem_ctx *ctx, lethe_dem_stats *stats);
/* Mobility status of every grid cell (lexicographic index ix + nx*(iy + ny*iz)) as of the last
 * contact search: AdaptiveSparseContacts::get_mobility_status_vector
 * (adaptive_sparse_contacts.h:317-326), what the reference's `mobility_status` test prints
 * (dem.cc:771-783). All LETHE_MOBILITY_MOBILE when sparse contacts are disabled. */
int lethe_dem_get_mobility_status(lethe_dem_ctx *ctx, uint64_t n_cells, int32_t *status);
/* Device time (ms, CUDA events on the engine's stream) spent in the fused step
 * kernel and in list rebuilds since the last call with reset != 0. */
int lethe_dem_get_timers(lethe_dem_ctx *ctx, int reset, double *step_kernel_ms,
                         uint64_t *step_kernel_launches, double *rebuild_ms,
                         uint64_t *rebuild_launches);
/* flags: bit 0 = CUDA-event timers, bit 1 = count touching pairs in the step kernel
 * (lethe_dem_stats.n_pairs_touching; one warp-aggregated atomic per warp). */
int lethe_dem_enable_timers(lethe_dem_ctx *ctx, int flags);
/* Region timing on the engine's own stream: record event `which` (0 = start, 1 = stop);
 * lethe_dem_event_elapsed synchronises and returns the device time between them. */
int lethe_dem_event_record(lethe_dem_ctx *ctx, int which);
int lethe_dem_event_elapsed(lethe_dem_ctx *ctx, double *ms);
/* Number of kernels this library has launched in the calling process so far. */
int lethe_dem_kernel_launches(lethe_dem_ctx *ctx, uint64_t *n_launches);

/* --- multi-GPU (one ctx per GPU per process; slab decomposition) ---
 * Replaces particle_handler.update_ghost_particles (dem.cc:686),
 * sort_particles_into_subdomains_and_cells + exchange_ghost_particles (dem.cc:986-989) and the
 * Utilities::MPI::logical_or of find_contact_detection_step.cc:53-58. NCCL carries the rebuild-step
 * exchanges (migration incl. contact history, ghost ids); the per-step ghost refresh is written by
 * the step kernel into the neighbour GPU's memory (CUDA IPC over NVLink) and the per-step logical_or
 * is a 4-byte peer-memory agreement. Every rank must be given the same wall / floating-wall /
 * boundary-motion tables. Environment switches (read at lethe_dem_comm_init / lethe_dem_create):
 *   LETHE_DEM_HALO=nccl      per-step ghost refresh over ncclSend/ncclRecv instead of peer stores
 *   LETHE_DEM_AGREE=nccl     per-step agreement over ncclAllReduce instead of peer memory
 *   LETHE_DEM_NO_PIPELINE=1  synchronous flag check before every step (no speculative launch) */
/* `subsection load balancing` (parameters_lagrangian.cc:960-1040,1118-1170; LagrangianLoadBalancing,
 * load_balancing.cc:17-58; DEMSolver::load_balance, dem.cc:383-457) for the slab decomposition: at a
 * load-balance iteration every internal cut plane moves towards the position that balances the
 * particles-per-layer histogram (by less than the narrowest slab per event), the particles that change
 * owner migrate with their contact history, and the contact search runs. frequency = `load balance
 * step` (once), `frequency` (frequent) or `dynamic check frequency` (dynamic); threshold = `threshold`.
 */
enum lethe_load_balance_method {
  LETHE_LOAD_BALANCE_NONE = 0,
  LETHE_LOAD_BALANCE_ONCE = 1,
  LETHE_LOAD_BALANCE_FREQUENT = 2,
  LETHE_LOAD_BALANCE_DYNAMIC = 3,
  LETHE_LOAD_BALANCE_DYNAMIC_WITH_SPARSE_CONTACTS = 4 /* needs config.sparse_contacts */
};
int lethe_dem_set_load_balancing(lethe_dem_ctx *ctx, int method, double threshold, int frequency);
/* `particle weight` (2000), the constant of `cell weight function` (1000), `active weight factor`, `inactive weight factor`
 * (1.0) of dynamic_with_sparse_contacts: load of a rank = cells x cell weight + sum over its particles of particle weight x
 * the factor of the particle's cell status (load_balancing.cc:60-122,184-222); the same weights shape the re-cut histogram. */
int lethe_dem_set_load_balancing_weights(lethe_dem_ctx *ctx, double particle_weight, double cell_weight,
                                         double active_weight_factor, double inactive_weight_factor);
/* Cell layers [lo, hi) along the slab axis this context owns now, and how many repartitions it has seen. */
int lethe_dem_get_slab(lethe_dem_ctx *ctx, int32_t *lo, int32_t *hi, uint64_t *n_repartitions);
/* The cut-plane rule on its own (host arithmetic, no device): cuts / new_cuts have world + 1 entries,
 * cuts[0] = 0, cuts[world] = n_layers. */
int lethe_dem_balanced_cuts(int32_t n_layers, const uint64_t *histogram, int32_t world, const int32_t *cuts,
                            int32_t max_shift, int32_t min_width, int32_t *new_cuts);
#define LETHE_DEM_NCCL_ID_BYTES 128
int lethe_dem_nccl_unique_id(uint8_t id[LETHE_DEM_NCCL_ID_BYTES]);
int lethe_dem_comm_init(lethe_dem_ctx *ctx, int rank, int world_size,
                        const uint8_t id[LETHE_DEM_NCCL_ID_BYTES]);

#ifdef __cplusplus
}
#endif
#endif /* LETHE_DEM_H */
