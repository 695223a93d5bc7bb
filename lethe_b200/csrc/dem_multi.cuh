// dem_multi.cuh — slab-decomposed multi-GPU driver (one context per GPU per process).
#pragma once
#include <cstdint>

#include "dem_kernels.cuh"

struct lethe_dem_ctx;

namespace dem
{
  struct MultiGpuImpl;
  struct MultiGpu
  {
    MultiGpuImpl *impl = nullptr;
    bool enabled() const { return impl != nullptr; }
    static int unique_id(uint8_t *id128);
    void init(lethe_dem_ctx *c, int rank, int world, const uint8_t *id128);
    void shutdown();
    // rebuild step: migrate particles that left the slab, re-exchange ghosts, rebuild lists
    void rebuild_with_exchange(lethe_dem_ctx *c);
    // non-rebuild step: refresh the state of the ghost copies (update_ghost_particles)
    void refresh_ghosts(lethe_dem_ctx *c);
    // logical_or over ranks of the displacement flag written by the last step kernel
    bool any_rank_flag(lethe_dem_ctx *c);
    // logical_or over ranks of a host-side decision
    bool agree(lethe_dem_ctx *c, bool local);
    // LagrangianLoadBalancing::check_load_balance_{once,frequent,dynamic} (load_balancing.cc:17-58) at the
    // top of an iteration: true = this iteration repartitions (collective; same answer on every rank)
    bool load_balance_due(lethe_dem_ctx *c);
    // adaptive sparse contacts across slabs (identify_mobility_status runs per rank on its own cells; the reference's
    // mobility_at_nodes.update_ghost_values(), adaptive_sparse_contacts.cc:212,270,308): the node values on the two cut
    // planes are max-merged with the neighbours' after each node pass, and the statuses of the neighbours' boundary
    // cell layers arrive after the last one (the broad search needs them for pairs that straddle a cut)
    void asc_exchange_nodes(lethe_dem_ctx *c);
    void asc_exchange_cells(lethe_dem_ctx *c);

    // ---- fused halo (peer-memory) mode ----
    // true once every rank has mapped its neighbours' state arrays (CUDA IPC over NVLink): the
    // step kernel then pushes the boundary-layer state itself and the only per-step collective
    // left is the 4-byte agreement below, which doubles as the barrier between steps.
    bool fused() const;
    // queue on the stream: contribution = (consult ? my displacement flag : 0) | host_bits,
    // all-reduce(max) into the device word the speculative step kernel checks, copy to the host
    void post_agree(lethe_dem_ctx *c, uint32_t host_bits, bool consult);
    // wait for the agreement posted last and return it (0 = nobody asked for a new list)
    uint32_t wait_agree(lethe_dem_ctx *c);
    // peer pointers / tables of the halo push for a kernel that writes state generation `out_gen`
    void fill_halo(lethe_dem_ctx *c, int out_gen, HaloPush &h) const;
    const uint32_t *agreed_flag_dev(lethe_dem_ctx *c) const;
  };
  void engine_rebuild_local(lethe_dem_ctx *c);
  void engine_upload_walls(lethe_dem_ctx *c);
  void engine_rebuild_sort(lethe_dem_ctx *c);
  void engine_rebuild_lists(lethe_dem_ctx *c);
  void engine_identify_mobility_status(lethe_dem_ctx *c);
  void engine_mirror_ids(lethe_dem_ctx *c);
} // namespace dem
