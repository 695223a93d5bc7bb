// dem_multi.cu — placeholder until the slab exchange lands.
#include <stdexcept>
#include "dem_multi.cuh"
namespace dem
{
  int MultiGpu::unique_id(uint8_t *) { return -1; }
  void MultiGpu::init(lethe_dem_ctx *, int, int, const uint8_t *) { throw std::runtime_error("multi-GPU exchange not built"); }
  void MultiGpu::shutdown() {}
  void MultiGpu::rebuild_with_exchange(lethe_dem_ctx *) {}
  void MultiGpu::refresh_ghosts(lethe_dem_ctx *) {}
} // namespace dem
