// dem_multi.cuh — slab-decomposed multi-GPU driver (one context per GPU per process).
#pragma once
#include <cstdint>

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
  };
  void engine_rebuild_local(lethe_dem_ctx *c);
  void engine_upload_walls(lethe_dem_ctx *c);
  void engine_rebuild_sort(lethe_dem_ctx *c);
  void engine_rebuild_lists(lethe_dem_ctx *c);
  void engine_mirror_ids(lethe_dem_ctx *c);
} // namespace dem
