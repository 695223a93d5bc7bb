// dem_solid.cuh — solid surfaces (triangle-mesh walls, SerialSolid<2,3>): device layout and
// kernel launch interface. Reference path:
//   SerialSolid::move_solid_triangulation              serial_solid.cc:333-410
//   SerialSolid::map_solid_in_background_triangulation serial_solid.cc:83-150
//   find_floating_mesh_mapping_step                    find_contact_detection_step.cc:139-161
//   particle_solid_surfaces_contact_search             particle_wall_broad_search.cc:129-215
//   particle_floating_mesh_fine_search                 particle_wall_fine_search.cc:170-210
//   calculate_particle_solid_object_contact            particle_wall_contact_force.cc:153-580
//   LetheGridTools::find_particle_triangle_projection  lethe_grid_tools.cc:1226-1450
//
// Layout: all solids share one vertex array and one triangle array (global triangle index =
// solid-major, so ascending index = the reference's (solid, triangle) iteration order). The
// mapping gives every background cell the sorted list of triangles a particle registered in it
// is a candidate of; a particle's row in the solid contact list is a copy of its cell's list
// plus the per-(particle, triangle) history. Per step: k_move_solids, then k_solid_contacts
// over the compact set of particles with a non-empty row (it leaves their summed force and
// torque for the fused step kernel to add), then k_step.
#pragma once

#include <vector>

#include "dem_kernels.cuh"

namespace dem
{
  constexpr int MAX_SOLIDS = 16;
  constexpr uint32_t SOLID_HIST_BIT = 0x80000000u;
  constexpr uint32_t SOLID_INDEX_MASK = 0x7fffffffu;
  constexpr int SOLID_MAX_CONTACTS = 24; // simultaneous triangle contacts of one sphere kept for the elimination

  struct SolidMotionDev
  {
    double translational_velocity[3];
    double angular_velocity[3];
    double center_of_rotation[3];
  };

  struct SolidSetView
  {
    double *vertices;             // [n_vertices][3]
    double *displacement;         // [n_vertices][3] since the last mapping
    const uint32_t *vertex_solid; // [n_vertices]
    const uint32_t *tri;          // [n_triangles][3] global vertex indices
    const uint32_t *tri_solid;    // [n_triangles]
    const uint32_t *es_start, *es_idx; // edge-sharing neighbours (CSR over triangles)
    const uint32_t *vs_start, *vs_idx; // vertex-sharing neighbours
    SolidMotionDev *motion;       // [n_solids]
    uint32_t n_vertices, n_triangles, n_solids;
  };

  struct SolidListView
  {
    uint32_t *row_start; // [n_rows + 1]
    uint32_t *entry;     // global triangle index | SOLID_HIST_BIT
    double *hist;        // [E][3]
    double *roll;        // [E][3]
  };

  struct SolidMoveParams
  {
    SolidSetView s;
    double dt;
    double criterion; // 3^-1/2 of the background cell diameter
    // trigger words shared with the step kernel (StepParams::flag_*); remap_host is set when a
    // vertex has moved further than `criterion` since the last mapping
    uint32_t *flag_local, *flag_host;
    uint32_t *remap_host;
    const uint32_t *flag_check;
    uint32_t flag_tag;
    int spec_check;
  };
  void launch_move_solids(const SolidMoveParams &p, cudaStream_t s);

  struct SolidBuildParams
  {
    const int32_t *cell_reg;        // particle -> registered cell
    const uint32_t *cell_tri_start; // [n_cells + 1] candidate triangles of a cell (sorted)
    const uint32_t *cell_tri;
    uint32_t n_rows;
    SolidListView old_list;
    const uint32_t *old_of_new;
    uint32_t n_old_rows;
    int clear_history;
    SolidListView new_list;
    uint32_t *counts; // [n_rows + 1]; after the fill: 1 where the row is not empty
    int use_roll;
    HistPayload pay; // history that arrived with particles that immigrated in this rebuild
    const uint8_t *mobility; // adaptive sparse contacts: per-cell status or nullptr
  };
  void launch_count_solid_rows(const SolidBuildParams &p, cudaStream_t s);
  void launch_fill_solid_rows(const SolidBuildParams &p, cudaStream_t s);

  struct SolidContactParams
  {
    SolidSetView s;
    SolidListView list;
    const uint32_t *active; // particles with a non-empty row
    uint32_t n_active;
    StateView in;
    double *force, *torque; // [n][3] summed solid-surface force / torque of the active particles
    uint32_t *overflow;     // set when a sphere touches more than SOLID_MAX_CONTACTS triangles
    int pw_model, rolling_model;
    double dt;
    const uint32_t *flag_check;
    uint32_t flag_tag;
    int spec_check;
    // an overflow also raises the contact-search trigger of this step (StepParams::flag_*), so that
    // the host reaches the rebuild — where the overflow is turned into an error — at the next step
    uint32_t *flag_local, *flag_host;
  };
  void launch_solid_contacts(const SolidContactParams &p, const MaterialTables &mt, cudaStream_t s);

  // Host side of map_solid_in_background_triangulation + particle_solid_surfaces_contact_search's
  // cell logic: per background cell the sorted, duplicate-free list of the triangles mapped to the
  // cell or to one of its vertex-sharing neighbours (find_full_cell_neighbors).
  void map_solids_on_host(const GridDesc &g, const double *vertices3, const uint32_t *tri3, uint32_t n_triangles,
                          std::vector<uint32_t> &cell_tri_start, std::vector<uint32_t> &cell_tri);
} // namespace dem
