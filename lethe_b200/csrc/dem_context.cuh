// dem_context.cuh — the engine context (all device buffers of one GPU) shared by
// dem_engine.cu (control flow + C ABI) and dem_multi.cu (slab exchange).
#pragma once

#include <algorithm>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "dem_kernels.cuh"
#include "dem_multi.cuh"
#include "dem_solid.cuh"

#define CU_TRY(call)                                                                                       \
  do                                                                                                       \
    {                                                                                                      \
      cudaError_t err__ = (call);                                                                          \
      if (err__ != cudaSuccess)                                                                            \
        {                                                                                                  \
          char buf__[512];                                                                                 \
          snprintf(buf__, sizeof(buf__), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
          throw std::runtime_error(buf__);                                                                 \
        }                                                                                                  \
    }                                                                                                      \
  while (0)

namespace dem
{
  template <class T> struct DevBuf
  {
    T *p = nullptr;
    size_t cap = 0;
    // when set, outgrown allocations are parked here instead of freed (a peer GPU may still have
    // them mapped through CUDA IPC; the owner of the list frees them once that cannot be)
    std::vector<void *> *retire = nullptr;
    ~DevBuf() { release(); }
    void release()
    {
      if (p)
        cudaFree(p);
      p = nullptr;
      cap = 0;
    }
    // grow to hold n elements; keep_n > 0 preserves the first keep_n elements
    void ensure(size_t n, size_t keep_n = 0, cudaStream_t s = 0, double growth = 1.25)
    {
      if (n <= cap)
        return;
      size_t ncap = std::max<size_t>(n, size_t(double(cap) * growth) + 16);
      T *np = nullptr;
      CU_TRY(cudaMalloc(&np, ncap * sizeof(T)));
      if (keep_n && p)
        {
          CU_TRY(cudaMemcpyAsync(np, p, std::min(keep_n, cap) * sizeof(T), cudaMemcpyDeviceToDevice, s));
          CU_TRY(cudaStreamSynchronize(s));
        }
      if (p)
        {
          if (retire)
            retire->push_back(p);
          else
            cudaFree(p);
        }
      p = np;
      cap = ncap;
    }
  };

  struct StateBufs
  {
    DevBuf<double4> pos, vel, omg;
    DevBuf<uint32_t> id;
    DevBuf<int32_t> cell_reg;
    dem::StateView view() { return dem::StateView{pos.p, vel.p, omg.p}; }
    void ensure(size_t n, size_t keep, cudaStream_t s)
    {
      pos.ensure(n, keep, s);
      vel.ensure(n, keep, s);
      omg.ensure(n, keep, s);
      id.ensure(n, keep, s);
      cell_reg.ensure(n, keep, s);
    }
  };

  struct ListBufs
  {
    DevBuf<uint32_t> row_start, col;
    DevBuf<double4> hist; // one 32-byte row per entry
    DevBuf<double> roll;
    DevBuf<uint8_t> img, rowl;
    uint32_t n_rows = 0;
    uint64_t n_entries = 0;
    dem::ListView view() { return dem::ListView{row_start.p, col.p, hist.p, roll.p, img.p, rowl.p}; }
  };
  struct WallListBufs
  {
    DevBuf<uint32_t> row_start, entry;
    DevBuf<double> hist, roll;
    uint32_t n_rows = 0;
    uint64_t n_entries = 0;
    dem::WallListView view() { return dem::WallListView{row_start.p, entry.p, hist.p, roll.p}; }
  };

  struct SolidListBufs
  {
    DevBuf<uint32_t> row_start, entry;
    DevBuf<double> hist, roll;
    uint32_t n_rows = 0;
    uint64_t n_entries = 0;
    dem::SolidListView view() { return dem::SolidListView{row_start.p, entry.p, hist.p, roll.p}; }
  };
} // namespace dem

using dem::DevBuf;
using dem::ListBufs;
using dem::StateBufs;
using dem::WallListBufs;

struct lethe_dem_ctx
{
  using GridDesc = dem::GridDesc;
  using MaterialTables = dem::MaterialTables;
  using BoundaryMotionDev = dem::BoundaryMotionDev;
  using FloatingWallsDev = dem::FloatingWallsDev;
  using StatsPartial = dem::StatsPartial;
  lethe_dem_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string error;

  GridDesc grid;
  MaterialTables mt;
  double thr2 = 0;

  // state: generation `cur` is current; the step kernel writes cur^1. ids / registered cells
  // only change at a rebuild and follow their own generation `cur_ids`.
  StateBufs st[2];
  int cur = 0;
  DevBuf<double> disp;
  uint32_t n_owned = 0; // particles owned (integrated) by this context
  uint32_t n_ghost = 0; // ghost copies stored behind the owned ones
  // id -> particle slot map of the current list generation; the previous generation's map is
  // kept as the history source of ghost copies (multi-GPU)
  DevBuf<uint32_t> slot_of_id, slot_of_id_old;
  uint32_t slot_map_size = 0, slot_map_size_old = 0;
  uint32_t old_n_owned = 0;
  // ghost runs (multi-GPU): [start,end) per curve rank for the copies received from the lower /
  // upper neighbour rank
  DevBuf<uint32_t> ghost_start[2], ghost_end[2];
  uint32_t n_ghost_run[2] = {0, 0};

  // multi-GPU rebuild in progress: first input index of the particles that immigrated in it
  // (0xffffffff: none) and the contact history that came with them (dem_multi.cu)
  uint32_t first_immigrant = 0xffffffffu;
  DevBuf<dem::HistRecord> pay;
  DevBuf<uint32_t> pay_start;
  uint32_t n_pay = 0;

  // lists (double-buffered across rebuilds: the old one is the history source)
  ListBufs lists[2];
  WallListBufs wlists[2];
  int cur_list = 0;

  // grid tables
  DevBuf<int32_t> cell_rank, cell_of_rank;
  DevBuf<uint32_t> cell_count, cell_start;
  DevBuf<uint32_t> key, slot, perm, old_of_new, counts, scan_tmp;
  // load balancing of a slab decomposition (lethe_dem_set_load_balancing; LagrangianLoadBalancing, load_balancing.cc)
  int lb_method = 0; // lethe_load_balance_method
  double lb_threshold = 0.5;
  int lb_frequency = 100000;
  double lb_particle_weight = 2000.0, lb_cell_weight = 1000.0, lb_active_factor = 1.0, lb_inactive_factor = 1.0;
  bool lb_recut_pending = false;
  uint64_t n_recuts = 0;
  // DEM-MP heat transfer (dem_kernels.cuh HeatParams): temperature / specific heat per particle id, rate per row
  bool thermal_enabled = false;
  dem::ThermalTables thermal_tables;
  DevBuf<double> temperature, specific_heat, heat_rate;
  size_t thermal_size = 0; // ids covered by temperature / specific_heat
  // adaptive sparse contacts (dem_kernels.cuh AscParams)
  bool asc_enabled = false;
  bool asc_reset = false;    // mobility_status_reset_trigger (dem_action_manager.h:128-134,61-75)
  bool asc_in_force = false; // the lists of the current generation were built with the mobility status
  DevBuf<uint8_t> asc_cell_status, asc_row_mobile;
  DevBuf<int> asc_node_status, asc_xbuf; // asc_xbuf: node planes / cell layers exchanged across the cuts
  DevBuf<uint32_t> nb_cand;    // candidate cache of the neighbour counting pass (NB_CACHE x n_owned)
  DevBuf<uint8_t> nb_cand_img; //   image codes beside it (periodic grids)

  // walls
  std::vector<lethe_wall_face> faces_host; // sorted by cell
  DevBuf<uint32_t> cell_face_start;
  DevBuf<double> face_normal, face_point;
  DevBuf<uint32_t> face_boundary;
  DevBuf<int32_t> face_motion;
  uint32_t n_faces = 0;
  std::vector<uint32_t> face_gid_host;
  struct Motion
  {
    uint32_t boundary_id;
    BoundaryMotionDev m;
  };
  std::vector<Motion> motions_host;
  DevBuf<BoundaryMotionDev> motions;
  FloatingWallsDev fw_host;
  DevBuf<FloatingWallsDev> fw_dev;
  DevBuf<uint32_t> cell_fw_mask;
  bool walls_dirty = true;

  // solid surfaces (dem_solid.cuh): topology and initial geometry on the host, current vertex
  // positions / centres of rotation on the device
  uint32_t n_solids = 0;
  std::vector<double> solid_vertices_host;
  std::vector<uint32_t> solid_tri_host, solid_tri_solid_host, solid_vertex_solid_host; // global vertex indices
  std::vector<uint32_t> solid_vertex_start, solid_tri_start;                          // per solid offsets (+ end)
  std::vector<dem::SolidMotionDev> solid_motion_host;
  uint32_t n_solid_vertices_dev = 0; // vertices whose current position lives on the device
  bool solids_dirty = false, solid_map_needed = false;
  DevBuf<double> solid_vertices, solid_disp, solid_force, solid_torque;
  DevBuf<uint32_t> solid_vertex_solid, solid_tri, solid_tri_solid, solid_es_start, solid_es_idx, solid_vs_start, solid_vs_idx;
  DevBuf<dem::SolidMotionDev> solid_motion;
  DevBuf<uint32_t> cell_tri_start, cell_tri, solid_active, solid_overflow;
  dem::SolidListBufs slists[2];
  uint32_t n_solid_active = 0;
  volatile uint32_t *h_remap = nullptr; // mapped pinned: a solid vertex moved past the mapping criterion
  uint32_t *d_remap = nullptr;

  // triggers / time (DEMActionManager + SimulationControl)
  uint64_t iteration_number = 0;
  double current_time = 0;
  bool contact_search_trigger = true;
  bool clear_history_trigger = false;
  uint64_t n_rebuilds = 0;
  uint64_t n_migrated = 0; // particles sent to / received from the neighbouring ranks so far
  // contact-detection trigger (StepParams::flag_*): the tag of the step that asked for a new
  // list, 0 = nobody. `h_flag` is mapped pinned memory (device alias `d_flag`) the host polls,
  // `flag_dev[0]` the device copy the next (speculative) launch checks, `flag_dev[1]` the
  // job-wide agreed flag of a multi-GPU run.
  volatile uint32_t *h_flag = nullptr;
  uint32_t *d_flag = nullptr;
  DevBuf<uint32_t> flag_dev;
  // pipelined stepping: step k+1 is queued before the host has seen step k's flag
  bool pipeline = true;
  cudaEvent_t step_done[2] = {nullptr, nullptr};
  uint64_t n_void_launches = 0;

  // debug taps
  DevBuf<double> force_out, torque_out;
  DevBuf<unsigned long long> touching;

  // staging for host rows
  DevBuf<uint32_t> stage_ids;
  // lethe_dem_step_host_state: the row -> particle id table of the caller's host rows, kept between
  // calls so that an unchanged table is not uploaded again
  // lethe_dem_set_external_loads: force / torque per particle ID (CFD-DEM fluid-particle interaction),
  // and the per-row sums handed to the step kernel (solid surfaces + external)
  DevBuf<double> ext_force, ext_torque, addend_force, addend_torque;
  uint32_t ext_size = 0;
  bool ext_enabled = false;
  bool open_next_step = false; // lethe_dem_restart_integration
  DevBuf<uint32_t> host_row_ids;
  uint64_t host_row_ids_n = 0;
  uint64_t host_row_ids_version = 0; // bumped whenever the caller hands over a new row -> id table
  // streamed host step: the plan that orders upload, partial step launches and download over space
  struct HostPipe
  {
    bool valid = false;
    uint64_t rebuild_gen = ~0ull, ids_version = ~0ull;
    uint32_t n_rows = 0, n_owned = 0, seg_rows = 0, n_seg = 0, n_stages = 0, n_blocks = 0;
    DevBuf<uint32_t> seg_up, seg_down, block_ready, up_list, down_list, block_list, row_of_slot;
    DevBuf<uint8_t> up_stage_of_slot;
    std::vector<uint32_t> up_off, down_off, block_off; // per stage: offsets into up_list / down_list / block_list
    // per stage: runs of consecutive host rows (first row, row count) moved by one copy each
    std::vector<std::vector<std::pair<uint64_t, uint64_t>>> up_runs, down_runs;
    cudaStream_t s_up = nullptr, s_down = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_down;
    cudaEvent_t ev_begin = nullptr;
    uint64_t n_calls = 0, n_plans = 0, n_zero_copy_calls = 0;
  } host_pipe;
  DevBuf<double> stage_x, stage_p;
  DevBuf<StatsPartial> stats_partials;

  // timers
  bool timers_enabled = false;
  bool count_touching = false;
  cudaEvent_t region_ev[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> event_pool;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_step, pending_rebuild;
  double step_ms = 0, rebuild_ms = 0;
  uint64_t step_launches = 0, rebuild_launches = 0;

  dem::MultiGpu multi;
};

