// dem_kernels.cuh — device data layout and kernel launch interface of the DEM engine.
//
// HBM layout (all arrays struct-of-arrays over the particle index p in cell-sorted
// order; cells are ordered along a Morton curve so that the 27-cell neighbourhood
// of a particle is a handful of short contiguous runs that stay L2-resident):
//
//   pos[p] = (x, y, z, d)         double4, 32 B = one sector per neighbour gather
//   vel[p] = (vx, vy, vz, mass)   double4
//   omg[p] = (wx, wy, wz, type)   double4  (type stored as double, like PropertiesIndex::type)
//   id[p], cell[p]                u32 / i32 (cell = grid cell at the last sort)
//   disp[p]                       f64, sum of dt*|v| since the last rebuild
//
//   row_start[p], col[e]          CSR contact list, FULL (each unordered pair appears in the
//                                 row of both particles); col bit 31 = "history is non-zero"
//   hist[e] (double4), roll[e][3] tangential displacement / EPSD spring torque of entry e in the
//                                 orientation row-particle -> col-particle. Both rows keep their own
//                                 copy; the arithmetic is exactly antisymmetric so the copies stay
//                                 bit-wise negatives of each other.
//   img[e]                        periodic image code of the neighbour (0 = none), periodic runs only
//   rowl[e]                       row(e) mod 32: lets the step kernel sweep the entries of 32 rows as one
//                                 flat, coalesced range without searching row_start
//
// pos/vel/omg are double-buffered: the fused step kernel reads generation g and writes g^1,
// so no thread ever sees a half-updated neighbour and one launch does forces + integration.
#pragma once

#include <cuda_runtime.h>

#include "dem_physics.cuh"

namespace dem
{
  constexpr uint32_t COL_HIST_BIT = 0x80000000u;
  constexpr uint32_t COL_INDEX_MASK = 0x7fffffffu;

  // wall entry word: bit31 floating wall, bit30 normal flipped (floating), bit29 history non-zero
  constexpr uint32_t WALL_FLOATING_BIT = 0x80000000u;
  constexpr uint32_t WALL_FLIPPED_BIT = 0x40000000u;
  constexpr uint32_t WALL_HIST_BIT = 0x20000000u;
  constexpr uint32_t WALL_INDEX_MASK = 0x1fffffffu;

  struct GridDesc
  {
    double lo[3];
    double h[3];
    double L[3]; // (lo + n*h) - lo, the periodic offset of each direction
    int n[3];
    int periodic[3];
    int n_cells;
    int slab_axis, slab_lo, slab_hi; // owned cell range along slab_axis (-1: everything)
  };

  struct FaceTable
  {
    // sorted by cell; cell_face_start has n_cells+1 entries
    const uint32_t *cell_face_start;
    const double *normal; // [n_faces][3]
    const double *point;  // [n_faces][3]
    const uint32_t *boundary_id;
    const int32_t *motion; // index into motions or -1
    uint32_t n_faces;
  };

  struct BoundaryMotionDev
  {
    double translational_velocity[3];
    double rotational_speed;
    double rotational_vector[3];
    double point_on_axis[3];
  };

  struct FloatingWallsDev
  {
    int n;
    double point[LETHE_DEM_MAX_FLOATING_WALLS][3];
    double normal[LETHE_DEM_MAX_FLOATING_WALLS][3];
    double t0[LETHE_DEM_MAX_FLOATING_WALLS], t1[LETHE_DEM_MAX_FLOATING_WALLS];
  };

  struct StateView
  {
    double4 *pos, *vel, *omg;
  };

  struct ListView
  {
    uint32_t *row_start;
    uint32_t *col;
    double4 *hist; // [E] one 32-byte row per entry: (x, y, z, unused) — a single 256-bit access, one DRAM sector
    double *roll; // [E][3] (EPSD only, else nullptr)
    uint8_t *img; // periodic only, else nullptr
    uint8_t *rowl; // [E] row of the entry modulo 32 (= lane of the step kernel's warp that owns it)
  };

  struct WallListView
  {
    uint32_t *row_start;
    uint32_t *entry;
    double *hist; // [W][3]
    double *roll; // [W][3]
  };

  enum StepPhase
  {
    PHASE_START = 0,   // integrate_start: half kick + full drift
    PHASE_REGULAR = 1, // integrate
    PHASE_END = 2      // integrate_end: half kick, no drift
  };

  // Fused halo push (multi-GPU): the step kernel stores the new state of the owned particles
  // that lie in the slab's boundary cell layers straight into the ghost slots of the
  // neighbouring GPU's OUT generation, through peer (NVLink) pointers — the per-step
  // update_ghost_particles (dem.cc:686) without a separate pack / send / receive.
  // Direction d: 0 = towards the lower neighbour, 1 = towards the upper one.
  struct HaloPush
  {
    double4 *pos[2], *vel[2], *omg[2]; // peer arrays of the generation this step writes (nullptr: nothing to push)
    const uint32_t *bits[2];           // per 32-row block: rows lying in the boundary layer
    const uint32_t *prefix[2];         // per 32-row block: boundary rows before the block
    uint32_t base[2];                  // first ghost slot of my run in the peer's arrays
  };

  struct StepParams
  {
    StateView in, out;
    ListView list;
    WallListView walls;
    const uint32_t *id;
    double *disp;
    // Contact-detection trigger (find_contact_detection_step.cc:29-58). A step whose largest
    // accumulated displacement exceeds `criterion` writes its non-zero `flag_tag` to
    // `flag_local` (device) and, when given, to `flag_host` (mapped pinned, single-GPU).
    // A SPECULATIVE launch (`spec_check`) is queued before the host has seen the previous
    // step's flag: it reads `flag_check` first and returns without touching anything when an
    // earlier step (any tag but its own) already asked for a new list.
    uint32_t *flag_local;
    uint32_t *flag_host;
    const uint32_t *flag_check;
    uint32_t flag_tag;
    int spec_check;
    HaloPush halo;
    unsigned long long *touching_counter; // debug (store_forces)
    const uint8_t *row_mobile;            // adaptive sparse contacts: per row, 0 = not integrated; nullptr = all mobile
    double *force_out, *torque_out;       // debug taps [N][3] or nullptr
    const double *solid_force, *solid_torque; // [N][3] solid-surface contacts of this step (dem_solid.cuh) or nullptr
    FaceTable faces;
    const BoundaryMotionDev *motions;
    const FloatingWallsDev *floating;
    uint32_t n_owned; // particles integrated by this rank
    // partial launch (streamed host step): only the 128-row blocks block_list[0 .. n_blocks_listed) are
    // stepped by this launch; nullptr = every block
    const uint32_t *block_list;
    uint32_t n_blocks_listed;
    int phase;
    int integrator; // lethe_integrator
    int mixed_precision; // pair model in float (lethe_precision)
    int pw_model;
    int rolling_model;
    int periodic_any;
    double dt;
    double g[3];
    double criterion;
    double moi_override;
    double L[3];
  };

  void launch_step(int pp_model, int rolling_model, const StepParams &p, const MaterialTables &mt, cudaStream_t stream);
  constexpr uint32_t STEP_BLOCK_ROWS = 128; // rows (particles) one thread block of the step kernel owns

  // ---- streamed host step: lethe_dem_step_host_state pipelined over space ----
  // The caller's host rows are cut into segments of seg_rows consecutive rows and go up in the caller's order, in n_stages
  // contiguous copies (segment g in stage seg_up[g] = g * n_stages / n_seg). A 128-row block of the step kernel is
  // ready in the latest upload stage of its own rows and of every particle its list rows name; a segment can go back to
  // the host after the latest ready stage of the blocks its rows belong to. Any host order that is coherent in space
  // (the engine's own cell-sorted order, the reference's cell-by-cell ParticleHandler order, slabs) makes most blocks
  // ready with their own rows; an incoherent order degenerates to "everything is ready in the last stage", the plain call.
  struct HostPlanParams
  {
    const uint32_t *row_ids; // [n_rows] particle id of every host row
    uint32_t n_rows, seg_rows;
    const uint32_t *slot_of_id;
    uint32_t map_size;
    ListView list;
    uint32_t n_owned;
    const uint32_t *seg_up;    // [n_seg] upload stage (host-made)
    uint32_t *seg_down;        // [n_seg] download stage, zeroed
    uint8_t *up_stage_of_slot; // [n_owned + ghosts], zeroed: particles without a host row are always there
    uint32_t *row_of_slot;     // [n_owned] host row of every particle slot, initialised to 0xffffffff (none)
    uint32_t *block_ready;     // [n_blocks], zeroed
  };
  // pass 0: row_of_slot, 1: up_stage_of_slot, 2: block_ready, 3: seg_down
  void launch_host_plan(const HostPlanParams &p, int pass, cudaStream_t s);
  // the rows of the segments seg_list[0 .. n_segs): host rows -> particle slots and back
  void launch_update_state_rows_segs(const uint32_t *seg_list, uint32_t n_segs, uint32_t seg_rows, const uint32_t *ids, const double *state9,
                                     uint32_t n, const uint32_t *slot_of_id, uint32_t slot_map_size, StateView st, cudaStream_t s);
  void launch_pack_state_rows_segs(const uint32_t *seg_list, uint32_t n_segs, uint32_t seg_rows, const uint32_t *ids, uint32_t n,
                                   const uint32_t *slot_of_id, uint32_t slot_map_size, StateView st, double *state9, cudaStream_t s);
  // Transfer order of the owned slots (lethe_dem_get_transfer_order): stable sort by cell layer along `axis` (slots not
  // registered in a cell first), the engine's cell-sorted order kept inside a layer. perm[k] = slot of row k.
  void transfer_order_perm(const int32_t *cell_reg, GridDesc grid, int axis, uint32_t n, uint32_t *perm, cudaStream_t s);
  // ids and (x, v, omega) rows of the slots perm[0 .. n), in that order
  void launch_pack_state_rows_perm(const uint32_t *perm, uint32_t n, StateView st, const uint32_t *id, uint32_t *ids_out, double *state9,
                                   cudaStream_t s);
  // the rows of the 128-slot blocks block_list[0 .. n_blocks) written straight into the caller's page-locked host rows
  // (state9 = device alias of the host buffer): rows of consecutive slots that are consecutive on the host leave as
  // full 128-byte lines
  void launch_pack_state_rows_blocks(const uint32_t *block_list, uint32_t n_blocks, const uint32_t *row_of_slot, uint32_t n_owned,
                                     StateView st, double *state9, cudaStream_t s);

  // ---- DEM-MP heat transfer (particle_heat_transfer.cc, multiphysics_integrator.cc) ----
  // effective pair tables of set_multiphysic_properties (particle_particle_contact_force.h:1755-1826)
  struct ThermalTables
  {
    double real_E[25], roughness[25], slope[25], microhardness[25], gas_m[25];
    double conductivity[5];
    double conductivity_gas;
  };
  struct HeatParams
  {
    StateView in;
    ListView list;
    const uint32_t *id;
    uint32_t n_owned;
    int periodic_any;
    double dt;
    double L[3];
    double *temperature;         // per particle id
    const double *specific_heat; // per particle id
    double *rate;                // per row: contact_outcome.heat_transfer_rate
  };
  // per row: conduction through every contact with a positive overlap, summed in list order
  void launch_heat_rates(int pp_model, const HeatParams &p, const MaterialTables &mt, const ThermalTables &th, cudaStream_t stream);
  // integrate_temperature with a zero heat source
  void launch_integrate_temperature(const HeatParams &p, cudaStream_t stream);

  // ---- rebuild ----
  struct BinParams
  {
    double4 *pos; // wrapped in place
    const int32_t *cell_reg;
    GridDesc grid;
    const int32_t *cell_rank; // lexicographic -> curve rank
    uint32_t *cell_count;     // [n_cells + 4], zeroed
    uint32_t *key, *slot;
    uint32_t n;
  };
  void launch_bin(const BinParams &p, cudaStream_t s);
  void launch_scatter_perm(const uint32_t *key, const uint32_t *slot, const uint32_t *cell_start, uint32_t *perm, uint32_t n,
                           cudaStream_t s);
  void launch_sort_cells(const uint32_t *cell_start, uint32_t n_buckets, uint32_t *perm, const uint32_t *id, cudaStream_t s);
  struct GatherParams
  {
    StateView in, out;
    const uint32_t *id_in;
    uint32_t *id_out;
    const uint32_t *perm;
    const uint32_t *key; // old index -> bucket (curve rank)
    const int32_t *cell_of_rank;
    int32_t *cell_reg_out;
    uint32_t *old_of_new;
    double *disp;
    uint32_t *slot_of_id;
    uint32_t n_new;
    // multi-GPU: particles at input index >= first_immigrant arrived from a neighbouring rank in
    // this rebuild; their "old" index is the ghost slot they had here (or none), looked up in
    // the id map of the outgoing list generation
    uint32_t first_immigrant; // 0xffffffff: none
    uint32_t n_listed;        // rows of the outgoing list: input indices in [n_listed, first_immigrant) were inserted since
    const uint32_t *old_slot_of_id;
    uint32_t old_map_size;
  };
  void launch_gather(const GatherParams &p, cudaStream_t s);

  // ---- multi-GPU helpers (dem_multi.cu drives them) ----
  struct MigrateRecord
  {
    double4 pos, vel, omg;
  };
  // Contact history that travels with a migrating particle (SURVEY.md §8e "variable payload incl.
  // history rows"): one record per touching entry of its particle-particle row and of its wall
  // row, so that a pair keeps its tangential displacement whichever rank ends up evaluating it
  // and an N-GPU run reproduces the single-domain one.
  struct HistRecord
  {
    uint32_t qid;   // id of the migrating particle
    uint32_t rid;   // partner id, or the wall key (WALL_FLOATING_BIT | index) of a wall record
    uint32_t flags; // HIST_REC_*
    uint32_t pad;
    double h[3];    // tangential displacement in the orientation qid -> rid
    double roll[3]; // EPSD rolling spring torque (zero otherwise)
  };
  constexpr uint32_t HIST_REC_PERIODIC = 1u, HIST_REC_WALL = 2u, HIST_REC_FLIPPED = 4u, HIST_REC_SOLID = 8u;
  struct HistPackParams
  {
    const uint32_t *send_slot; // emigrants of one direction: slot in the (pre-sort) particle arrays
    uint32_t n;
    const uint32_t *id;        // [owned + ghost] ids of the pre-sort arrangement
    ListView list;
    WallListView walls;
    uint32_t n_rows, n_wall_rows;
    // solid-surface contact rows (dem_solid.cuh; nullptr when there are no solids): entry = global
    // triangle index | bit 31 (history)
    const uint32_t *solid_row_start, *solid_entry;
    const double *solid_hist, *solid_roll;
    uint32_t n_solid_rows;
    int use_roll, use_img;
    uint32_t *counts;          // [n + 1] (count pass) / exclusive offsets (pack pass)
    HistRecord *out;
  };
  void launch_hist_count(const HistPackParams &p, cudaStream_t s);
  void launch_hist_pack(const HistPackParams &p, cudaStream_t s);
  // pay_start[qid] = first record of qid (records of one particle are contiguous)
  void launch_hist_index(const HistRecord *rec, uint32_t n, uint32_t *pay_start, uint32_t map_size, cudaStream_t s);
  struct HistPayload
  {
    const HistRecord *rec; // nullptr: nothing arrived
    const uint32_t *start; // by particle id, 0xffffffff = none
    uint32_t n, map_size;
    const uint32_t *id;    // ids of the new arrangement (owned + ghost)
  };
  // classify owned particles against the slab [slab_lo, slab_hi) along grid.slab_axis after the
  // periodic wrap; movers are appended to send buffers (dir 0 = lower neighbour, 1 = upper) and
  // marked cell_reg = -2 so that the following sort drops them.
  struct ClassifyParams
  {
    double4 *pos;
    const double4 *vel, *omg;
    const uint32_t *id;
    int32_t *cell_reg;
    GridDesc grid;
    uint32_t n;
    MigrateRecord *send_rec[2];
    uint32_t *send_id[2];
    uint32_t *send_slot[2];
    uint32_t *send_count; // [2] + [2] = far movers (error)
    uint32_t send_cap;
    int max_hop; // cell layers a particle may lie outside the slab and still go to the adjacent rank (1; more right after a re-cut)
  };
  void launch_classify(const ClassifyParams &p, cudaStream_t s);
  // particles per cell layer along the slab axis (load balancing): hist[layer] += 1 for every owned particle
  void launch_layer_histogram(const double4 *pos, GridDesc grid, uint32_t n, uint32_t *hist, cudaStream_t s);
  // New cut planes of a slab decomposition: cuts[0] = 0 < cuts[1] < ... < cuts[world] = n_layers. Every
  // internal cut moves towards the position that balances the particle histogram, by at most
  // max_shift layers and never out of the two slabs it separates, keeping every slab at least min_width layers wide. Pure host function (the
  // role of p4est's weighted repartition, load_balancing.cc:9-40, for slabs).
  void balanced_cuts(int n_layers, const uint64_t *hist, int world, const int32_t *cuts, int max_shift, int min_width, int32_t *new_cuts);
  void launch_append_records(const MigrateRecord *rec, const uint32_t *ids, uint32_t n, StateView st, uint32_t *id_out,
                             int32_t *cell_reg, double *disp, uint32_t base, cudaStream_t s);
  // flag owned particles (sorted) lying in the slab's boundary cell layer `layer` along the axis
  void launch_flag_layer(const int32_t *cell_reg, GridDesc grid, int layer_cell, uint32_t n, uint32_t *flags, cudaStream_t s);
  void launch_compact_indices(const uint32_t *flags, const uint32_t *offsets, uint32_t n, uint32_t *out, cudaStream_t s);
  // per block of 32 rows: ballot of `flags` and the exclusive prefix at the block's first row (HaloPush tables)
  void launch_halo_warp_table(const uint32_t *flags, const uint32_t *offsets, uint32_t n, uint32_t *bits, uint32_t *prefix,
                              cudaStream_t s);
  // word[2] = (consult ? word[0] : 0) | host_bits : the rank's contribution to the per-step agreement
  void launch_prepare_flag(uint32_t *flag_words, uint32_t host_bits, int consult, cudaStream_t s);
  // Per-step agreement over peer memory (the logical_or of find_contact_detection_step.cc:53-58
  // without a library collective): every rank stores (seq, its word) into slot `rank` of every
  // rank's mailbox (system-scope release), waits until its own mailbox holds `seq` from all
  // `world` ranks (acquire) and writes the maximum to flag_words[1] and to `host_out` (mapped
  // pinned memory the host reads after the event behind this kernel: no copy engine in the step). mailbox layout:
  // [parity of seq][world] u64 = seq << 32 | word. A rank that waits longer than ~2 min writes
  // 0xffffffff (the host turns that into an error instead of a hung GPU).
  void launch_agree(uint64_t *const *peer_mailbox, uint64_t *my_mailbox, int rank, int world, uint32_t seq, uint32_t *flag_words,
                    uint32_t host_bits, int consult, uint32_t *host_out, cudaStream_t s);
  void launch_gather_state(StateView st, const uint32_t *idx, uint32_t n, double4 *pos, double4 *vel, double4 *omg,
                           cudaStream_t s);
  void launch_gather_ids(const uint32_t *id, const uint32_t *idx, uint32_t n, uint32_t *out, cudaStream_t s);
  // ghost run bookkeeping: cell of each ghost, [start,end) per curve rank, old index via id map
  struct GhostRunParams
  {
    const double4 *pos; // ghost positions (run)
    GridDesc grid;
    const int32_t *cell_rank;
    uint32_t base; // index of the first ghost of the run in the particle arrays
    uint32_t n;
    int32_t *cell_reg;          // [base + g]
    uint32_t *start, *end;      // by curve rank, zero-initialised
    const uint32_t *id;         // [base + g]
    const uint32_t *old_slot_of_id;
    uint32_t old_map_size, old_n_owned;
    uint32_t *old_of_new;       // [base + g]
  };
  void launch_ghost_run(const GhostRunParams &p, cudaStream_t s);
  void launch_register_ids(const uint32_t *id, uint32_t base, uint32_t n, uint32_t *slot_of_id, uint32_t map_size, cudaStream_t s);

  struct NeighborParams
  {
    StateView st;
    const int32_t *cell_reg;
    const int32_t *cell_rank;
    const uint32_t *cell_start; // by curve rank, n_cells+1
    GridDesc grid;
    double thr2;
    uint32_t n_rows;  // rows built (owned particles)
    uint32_t n_total; // owned + ghost particles present in the cell lists
    // ghost copies (multi-GPU): up to two runs, each sorted by curve rank, stored behind the
    // owned particles; per-rank [start,end) tables or nullptr
    const uint32_t *ghost_start[2];
    const uint32_t *ghost_end[2];
    uint32_t old_n_owned; // ownership split of the old list's particle indices
    // old list (history source)
    ListView old_list;
    const uint32_t *old_of_new; // new index -> old index or 0xffffffff
    uint32_t n_old_rows;
    int clear_history;
    // new list
    ListView new_list;
    uint32_t *counts; // [n_rows+1]
    int use_roll, use_img;
    HistPayload pay; // history that arrived with immigrants
    // candidate cache written by the counting pass, read by the filling pass (or nullptr):
    // candidate k of row q at [k * n_rows + q], k < NB_CACHE
    uint32_t *cand;
    uint8_t *cand_img;
    // adaptive sparse contacts: per-cell mobility status (lexicographic cell index) or nullptr.
    // A pair is listed iff at least one of its two cells is mobile
    // (particle_particle_broad_search.cc:134-316).
    const uint8_t *mobility;
  };
  constexpr uint32_t NB_CACHE = 24;
  void launch_count_neighbors(const NeighborParams &p, cudaStream_t s);
  void launch_fill_neighbors(const NeighborParams &p, cudaStream_t s);

  struct WallBuildParams
  {
    StateView st;
    const int32_t *cell_reg;
    GridDesc grid;
    FaceTable faces;
    const FloatingWallsDev *floating;
    const uint32_t *cell_fw_mask; // per cell bit mask of floating walls whose boundary cells include it
    double time;
    uint32_t n_rows;
    WallListView old_list;
    const uint32_t *old_of_new;
    uint32_t n_old_rows;
    int clear_history;
    WallListView new_list;
    uint32_t *counts;
    int use_roll;
    HistPayload pay;
    const uint8_t *mobility; // adaptive sparse contacts: only particles of mobile cells get wall candidates, or nullptr
  };
  void launch_count_walls(const WallBuildParams &p, cudaStream_t s);
  void launch_fill_walls(const WallBuildParams &p, cudaStream_t s);

  // exclusive scan of n+1 u32 values (in[n] is ignored and treated as 0); out[n] = total.
  // tmp must hold scan_tmp_elems(n+1) u32.
  size_t scan_tmp_elems(size_t n);
  void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n_plus_1, uint32_t *tmp, cudaStream_t s);

  // ---- misc ----
  struct StatsPartial
  {
    double vmin, vmax, vsum, wmin, wmax, wsum, ktmin, ktmax, ktsum, krmin, krmax, krsum;
  };
  void launch_stats(StateView st, uint32_t n, double moi_override, StatsPartial *partials, uint32_t n_blocks, cudaStream_t s);
  constexpr uint32_t STATS_BLOCK = 256;

  // host-facing AoS <-> device SoA
  void launch_unpack_host_rows(const uint32_t *ids, const double *x3, const double *props9, uint32_t n, StateView st,
                               uint32_t *id_out, int32_t *cell_reg, double *disp, uint32_t base, cudaStream_t s);
  void launch_update_from_host_rows(const uint32_t *ids, const double *x3, const double *props9, uint32_t n,
                                    const uint32_t *slot_of_id, uint32_t slot_map_size, StateView st, cudaStream_t s);
  void launch_accumulate_displacement(const double4 *vel, double *disp, uint32_t n, double dt, double criterion, uint32_t *flag_local,
                                      uint32_t *flag_host, uint32_t tag, cudaStream_t s);
  void launch_compose_external_loads(const uint32_t *id, uint32_t n, const double *ext_force, const double *ext_torque,
                                     uint32_t ext_size, const double *solid_force, const double *solid_torque, double *force,
                                     double *torque, cudaStream_t s);
  void launch_scatter_external_loads(const uint32_t *ids, const double *force3, const double *torque3, uint32_t n, double *ext_force,
                                     double *ext_torque, uint32_t ext_size, cudaStream_t s);
  void launch_update_state_rows(const uint32_t *ids, const double *state9, uint32_t n, const uint32_t *slot_of_id,
                                uint32_t slot_map_size, StateView st, cudaStream_t s);
  void launch_pack_state_rows(const uint32_t *ids, uint32_t n, const uint32_t *slot_of_id, uint32_t slot_map_size, StateView st,
                              double *state9, cudaStream_t s);
  void launch_pack_host_rows(const uint32_t *ids, uint32_t n, const uint32_t *slot_of_id, uint32_t slot_map_size, StateView st,
                             double *x3, double *props9, cudaStream_t s);
  void launch_pack_all_rows(StateView st, const uint32_t *id, uint32_t n, uint32_t *ids_out, double *x3, double *props9,
                            cudaStream_t s);
  void launch_fill_u32(uint32_t *p, uint32_t v, size_t n, cudaStream_t s);
  // process-wide count of kernel launches issued by this library
  void count_launch(unsigned n = 1);
  unsigned long long launch_count();
  // ---- adaptive sparse contacts (AdaptiveSparseContacts::identify_mobility_status,
  // adaptive_sparse_contacts.cc:132-356, on the uniform grid: nodes = grid vertices) ----
  struct AscParams
  {
    StateView st;               // cell-sorted owned particles
    const uint32_t *cell_start; // by curve rank
    const int32_t *cell_rank;   // lexicographic cell -> curve rank
    const int32_t *cell_reg;    // per particle: lexicographic cell
    GridDesc grid;
    double granular_temperature_threshold, solid_fraction_threshold;
    uint8_t *cell_status; // [n_cells] lethe_mobility_status, 0xff = not assigned yet
    int *node_status;     // [(nx+1)(ny+1)(nz+1)]
    uint8_t *row_mobile;  // [n_rows] 1 = the particle's cell is mobile
    uint32_t n_rows;
    int owned_lo, owned_hi; // slab decomposition: only cell layers [owned_lo, owned_hi) along grid.slab_axis are this rank's to classify; -1 = all
  };
  constexpr uint8_t ASC_UNASSIGNED = 0xffu; // cell_status while the passes run
  constexpr uint8_t ASC_FOREIGN = 0xfeu;    // a cell another rank classifies (its status arrives with MultiGpu::asc_exchange_cells)
  // slab decomposition: one plane of nodes (axis index `index` in node coordinates) or one layer of cells across the slab
  // axis, gathered into / merged from a contiguous buffer that travels to the neighbour rank
  struct AscPlaneParams
  {
    GridDesc grid;
    int index;
    int *node_status;
    uint8_t *cell_status;
    int *buf;
  };
  // op 0: buf <- node plane, 1: node plane <- max(node plane, buf), 2: buf <- cell layer, 3: cell layer <- buf
  void launch_asc_plane(const AscPlaneParams &p, int op, cudaStream_t s);
  // particles per cell layer weighted by the mobility status of their cell (x1000 fixed point): load_balancing.cc:184-222
  void launch_layer_histogram_weighted(const double4 *pos, const int32_t *cell_reg, const uint8_t *cell_status, GridDesc grid, uint32_t n,
                                       uint32_t w_mobile, uint32_t w_active, uint32_t w_inactive, uint32_t *hist, cudaStream_t s);
  // pass 0: empty cells, 1: mobile by criteria, 2: mobile by neighbour, 3: active / inactive, 4: per-particle flag
  void launch_asc_pass(const AscParams &p, int pass, cudaStream_t s);
} // namespace dem
