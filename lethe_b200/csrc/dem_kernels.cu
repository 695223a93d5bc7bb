// dem_kernels.cu — contact-list rebuild: cell binning along a Morton curve, counting sort,
// neighbour (fine) search with tangential-history carry-over, wall candidate lists, scans.
//
// Reference path replaced (the `if (check_contact_search())` branch of
// DEMSolver::execute_contact_detection_and_search, source/dem/dem.cc:631-683):
//   periodic wrap              periodic_boundaries_manipulator.cc:145-224,266-338
//   sort into cells            ParticleHandler::sort_particles_into_subdomains_and_cells (dem.cc:982-1016)
//   pp broad search            particle_particle_broad_search.cc:9-132,316-380,735-766
//   history reconciliation     update_fine_search_candidates.cc:9-212, update_local_particle_containers.cc:40-200
//   pp fine search             particle_particle_fine_search.cc:19-232
//   pw broad + fine search     particle_wall_broad_search.cc:8-125, particle_wall_fine_search.cc:18-166
//
// The reference's pair list after a rebuild is, as a set of unordered pairs,
//   { (i,j) in vertex-sharing cells : d2 < thr2 }  U  { old pairs still in vertex-sharing cells : d2 == thr2 }
// and a pair keeps its history iff it was in the list before (in the same container:
// non-periodic vs periodic). That is what k_count/k_fill_neighbors build directly.
#include <algorithm>
#include <atomic>
#include <vector>

#include <cub/device/device_radix_sort.cuh> // host-side plumbing only (transfer_order_perm); no step or rebuild kernel uses it

#include "dem_kernels.cuh"

namespace dem
{
  namespace
  {
    constexpr int SCAN_THREADS = 256;
    constexpr int SCAN_ITEMS = 8;
    constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

    __device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v)
    {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
        {
          const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
          if ((threadIdx.x & 31) >= o)
            v += t;
        }
      return v;
    }

    // tile-local exclusive scan (warp shuffles + one shared-memory hop), tile totals to sums[]
    __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t *in, uint32_t *out, size_t n, size_t n_valid,
                                                                 uint32_t *sums)
    {
      __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
      const size_t base = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_ITEMS;
      uint32_t v[SCAN_ITEMS];
      uint32_t local = 0;
#pragma unroll
      for (int k = 0; k < SCAN_ITEMS; ++k)
        {
          const size_t idx = base + k;
          v[k] = (idx < n_valid) ? in[idx] : 0u;
          local += v[k];
        }
      const uint32_t incl = warp_inclusive_scan(local);
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      if (lane == 31)
        warp_tot[warp] = incl;
      __syncthreads();
      if (warp == 0)
        {
          uint32_t t = (lane < SCAN_THREADS / 32) ? warp_tot[lane] : 0u;
          t = warp_inclusive_scan(t);
          if (lane < SCAN_THREADS / 32)
            warp_tot[lane] = t;
        }
      __syncthreads();
      uint32_t run = (incl - local) + (warp > 0 ? warp_tot[warp - 1] : 0u);
#pragma unroll
      for (int k = 0; k < SCAN_ITEMS; ++k)
        {
          const size_t idx = base + k;
          if (idx < n)
            out[idx] = run;
          run += v[k];
        }
      if (threadIdx.x == SCAN_THREADS - 1)
        sums[blockIdx.x] = warp_tot[SCAN_THREADS / 32 - 1];
    }

    __global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t *out, size_t n, const uint32_t *sums_scanned)
    {
      const uint32_t add = sums_scanned[blockIdx.x];
      const size_t base = size_t(blockIdx.x) * SCAN_TILE;
      for (int k = threadIdx.x; k < SCAN_TILE; k += SCAN_THREADS)
        {
          const size_t idx = base + k;
          if (idx < n)
            out[idx] += add;
        }
    }

    __device__ __forceinline__ int cell_of_point(const GridDesc &g, double x, double y, double z)
    {
      const double p[3] = {x, y, z};
      int idx[3];
#pragma unroll
      for (int d = 0; d < 3; ++d)
        {
          const double r = (p[d] - g.lo[d]) / g.h[d];
          double f = floor(r);
          // a point exactly on a face shared by two cells belongs to the lower one (the
          // reference's point location returns the first cell that contains the point; pinned by
          // epsd_rolling_resistance_model.output)
          if (r == f && f > 0.0)
            f -= 1.0;
          if (!(f >= 0.0) || !(f < double(g.n[d])))
            return -1;
          idx[d] = int(f);
        }
      return idx[0] + g.n[0] * (idx[1] + g.n[1] * idx[2]);
    }

    __global__ void __launch_bounds__(256) k_bin(const __grid_constant__ BinParams P)
    {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      if (p >= P.n)
        return;
      double4 x = P.pos[p];
      const GridDesc &g = P.grid;
      const int creg = P.cell_reg[p];
      if ((g.periodic[0] | g.periodic[1] | g.periodic[2]) && creg >= 0)
        {
          // check_and_move_particles: only particles registered in a cell on a periodic face
          const int c[3] = {creg % g.n[0], (creg / g.n[0]) % g.n[1], creg / (g.n[0] * g.n[1])};
          double xv[3] = {x.x, x.y, x.z};
#pragma unroll
          for (int d = 0; d < 3; ++d)
            {
              if (!g.periodic[d])
                continue;
              const double lo = g.lo[d];
              const double hi = g.lo[d] + g.n[d] * g.h[d];
              if (c[d] == 0)
                {
                  const double distance_with_face = (xv[d] - lo) * -1.0;
                  if (distance_with_face >= 0.0)
                    xv[d] += g.L[d];
                }
              if (c[d] == g.n[d] - 1)
                {
                  const double distance_with_face = (xv[d] - hi) * 1.0;
                  if (distance_with_face >= 0.0)
                    xv[d] += -g.L[d];
                }
            }
          x.x = xv[0];
          x.y = xv[1];
          x.z = xv[2];
          P.pos[p] = x;
        }
      const int lin = creg == -2 ? -1 : cell_of_point(g, x.x, x.y, x.z); // -2: migrated to a neighbour rank
      uint32_t bucket;
      if (lin < 0)
        bucket = uint32_t(g.n_cells); // left the triangulation: dropped
      else
        bucket = uint32_t(P.cell_rank[lin]);
      P.key[p] = bucket;
      P.slot[p] = atomicAdd(&P.cell_count[bucket], 1u);
    }

    __global__ void __launch_bounds__(256) k_scatter_perm(const uint32_t *key, const uint32_t *slot, const uint32_t *cell_start,
                                                          uint32_t *perm, uint32_t n)
    {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      if (p >= n)
        return;
      perm[cell_start[key[p]] + slot[p]] = p;
    }

    // Deterministic in-cell order: sort each cell's slice of perm by particle id (the atomic
    // slot order is not reproducible). Cells hold a handful of particles: insertion sort.
    __global__ void __launch_bounds__(256) k_sort_cells(const uint32_t *cell_start, uint32_t n_buckets, uint32_t *perm,
                                                        const uint32_t *id)
    {
      const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
      if (c >= n_buckets)
        return;
      const uint32_t s = cell_start[c], e = cell_start[c + 1];
      for (uint32_t a = s + 1; a < e; ++a)
        {
          const uint32_t pa = perm[a];
          const uint32_t ida = id[pa];
          uint32_t b = a;
          while (b > s && id[perm[b - 1]] > ida)
            {
              perm[b] = perm[b - 1];
              --b;
            }
          perm[b] = pa;
        }
    }

    __global__ void __launch_bounds__(256) k_gather(const __grid_constant__ GatherParams P)
    {
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= P.n_new)
        return;
      const uint32_t o = P.perm[q];
      P.out.pos[q] = P.in.pos[o];
      P.out.vel[q] = P.in.vel[o];
      P.out.omg[q] = P.in.omg[o];
      const uint32_t pid = P.id_in[o];
      P.id_out[q] = pid;
      P.cell_reg_out[q] = P.cell_of_rank[P.key[o]];
      uint32_t old = o;
      if (o >= P.first_immigrant)
        old = (P.old_slot_of_id && pid < P.old_map_size) ? P.old_slot_of_id[pid] : 0xffffffffu;
      else if (o >= P.n_listed)
        old = 0xffffffffu; // inserted after the outgoing list was built: no history anywhere
      P.old_of_new[q] = old;
      P.disp[q] = 0.0;
      P.slot_of_id[pid] = q;
    }

    // ---- neighbour search ----
    struct NbVisitor
    {
      uint32_t count;
    };

    // Visits every (neighbour r, image code) of particle q that belongs in the contact list.
    // The 27 stencil cells are resolved first, in three batches of independent loads (curve rank,
    // [start, end) of the owned run and of the ghost runs), so that the per-cell chain of dependent
    // loads rank -> start -> position is paid once instead of 27 times; the walk itself keeps the
    // stencil order (z, y, x ascending), which is the order of the list row.
    template <class F> __device__ __forceinline__ void for_each_neighbor(const NeighborParams &P, uint32_t q, F &&f)
    {
      const GridDesc &g = P.grid;
      const double4 pq = P.st.pos[q];
      const vec3 xq = v3(pq.x, pq.y, pq.z);
      const int lin = P.cell_reg[q];
      const int ci = lin % g.n[0], cj = (lin / g.n[0]) % g.n[1], ck = lin / (g.n[0] * g.n[1]);
      const uint32_t old_q = P.old_of_new ? P.old_of_new[q] : 0xffffffffu;
      // adaptive sparse contacts: particles of inactive cells are in no list; a pair needs a mobile cell
      const uint32_t my_status = P.mobility ? P.mobility[lin] : uint32_t(LETHE_MOBILITY_MOBILE);
      if (my_status == LETHE_MOBILITY_INACTIVE)
        return;
      // s = image shift (in units of L) that brings the neighbour cell next to mine:
      // my cell on the low face, neighbour wrapped to the high face -> s = -1.
      int n1[3][3]; // wrapped cell coordinate per axis and offset, -1: outside a non-periodic face
      int sh[3][3]; // image shift per axis and offset
      const int cc[3] = {ci, cj, ck};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int o = 0; o < 3; ++o)
          {
            int v = cc[a] + o - 1, sft = 0;
            if (v < 0 || v >= g.n[a])
              {
                if (!g.periodic[a])
                  v = -1;
                else
                  {
                    sft = v < 0 ? -1 : 1;
                    v -= sft * g.n[a];
                  }
              }
            n1[a][o] = v;
            sh[a][o] = sft;
          }
      uint32_t rank[27];
#pragma unroll
      for (int k = 0; k < 27; ++k)
        {
          const int ni = n1[0][k % 3], nj = n1[1][(k / 3) % 3], nk = n1[2][k / 9];
          const bool outside = ni < 0 || nj < 0 || nk < 0;
          const int nlin = outside ? 0 : ni + g.n[0] * (nj + g.n[1] * nk);
          rank[k] = outside ? 0xffffffffu : uint32_t(P.cell_rank[nlin]);
          if (!outside && my_status != LETHE_MOBILITY_MOBILE && P.mobility[nlin] != LETHE_MOBILITY_MOBILE)
            rank[k] = 0xffffffffu;
        }
      uint32_t cs[27], ce[27];
#pragma unroll
      for (int k = 0; k < 27; ++k)
        {
          cs[k] = rank[k] != 0xffffffffu ? P.cell_start[rank[k]] : 0u;
          ce[k] = rank[k] != 0xffffffffu ? P.cell_start[rank[k] + 1] : 0u;
        }
      const bool ghosts = P.ghost_start[0] || P.ghost_start[1];
#pragma unroll
      for (int k = 0; k < 27; ++k)
        {
          if (rank[k] == 0xffffffffu)
            continue;
          const int sx = sh[0][k % 3], sy = sh[1][(k / 3) % 3], sz = sh[2][k / 9];
          const uint32_t img = (sx | sy | sz) ? uint32_t(1 + (sx + 1) + 3 * (sy + 1) + 9 * (sz + 1)) : 0u;
          // owned particles of the cell, then the ghost copies of either neighbour rank
          for (int range = 0; range < (ghosts ? 3 : 1); ++range)
            {
              uint32_t s, e;
              if (range == 0)
                {
                  s = cs[k];
                  e = ce[k];
                }
              else
                {
                  if (!P.ghost_start[range - 1])
                    continue;
                  s = P.ghost_start[range - 1][rank[k]];
                  e = P.ghost_end[range - 1][rank[k]];
                }
              for (uint32_t r = s; r < e; ++r)
                {
                  if (r == q)
                    continue;
                  const double4 pr = P.st.pos[r];
                  const vec3 xr = v3(pr.x, pr.y, pr.z);
                  double d2;
                  if (img)
                    {
                      const vec3 shift = v3(sx * g.L[0], sy * g.L[1], sz * g.L[2]);
                      const int first = sx != 0 ? sx : (sy != 0 ? sy : sz);
                      // canonical orientation (see decode_image in dem_step.cu)
                      if (first < 0)
                        d2 = dist2(xq, xr + shift);
                      else
                        d2 = dist2(xr, xq + (-shift));
                    }
                  else
                    d2 = dist2(xq, xr);
                  bool in = d2 < P.thr2;
                  if (!in && d2 == P.thr2 && old_q != 0xffffffffu && old_q < P.n_old_rows)
                    {
                      // pairs sitting exactly on the threshold are neither inserted (<) nor
                      // erased (>): they survive iff they were already listed.
                      const uint32_t old_r = P.old_of_new[r];
                      for (uint32_t eo = P.old_list.row_start[old_q]; eo < P.old_list.row_start[old_q + 1]; ++eo)
                        if ((P.old_list.col[eo] & COL_INDEX_MASK) == old_r &&
                            ((P.use_img ? P.old_list.img[eo] != 0 : false) == (img != 0)))
                          in = true;
                    }
                  if (in)
                    f(r, img);
                }
            }
        }
    }

    // The counting pass keeps what it finds: candidate k of row q goes to cand[k * n_rows + q]
    // (coalesced across the rows of a warp), image code beside it, for the first NB_CACHE
    // candidates of a row. The filling pass then reads its row back instead of walking the
    // stencil a second time; rows with more candidates than the cache holds are walked again.
    __global__ void __launch_bounds__(128) k_count_neighbors(const __grid_constant__ NeighborParams P)
    {
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= P.n_rows)
        return;
      uint32_t count = 0;
      for_each_neighbor(P, q, [&](uint32_t r, uint32_t img) {
        if (P.cand && count < NB_CACHE)
          {
            P.cand[size_t(count) * P.n_rows + q] = r;
            if (P.use_img)
              P.cand_img[size_t(count) * P.n_rows + q] = uint8_t(img);
          }
        ++count;
      });
      P.counts[q] = count;
    }

    __global__ void __launch_bounds__(128) k_fill_neighbors(const __grid_constant__ NeighborParams P)
    {
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= P.n_rows)
        return;
      uint32_t e = P.new_list.row_start[q];
      const uint32_t n_mine = P.new_list.row_start[q + 1] - e;
      const uint32_t old_q = P.old_of_new ? P.old_of_new[q] : 0xffffffffu;
      const bool have_old = !P.clear_history && old_q != 0xffffffffu && old_q < P.n_old_rows;
      uint32_t o0 = 0, o1 = 0;
      if (have_old)
        {
          o0 = P.old_list.row_start[old_q];
          o1 = P.old_list.row_start[old_q + 1];
        }
      const uint32_t qid = P.pay.rec ? P.pay.id[q] : 0u;
      auto emit = [&](uint32_t r, uint32_t img) {
        uint32_t word = r;
        const uint32_t old_r = (!P.clear_history && P.old_of_new) ? P.old_of_new[r] : 0xffffffffu;
        bool found = false;
        // The history of a pair follows the pair, whichever rank evaluates it (the reference
        // restarts it from zero when a partner changes owner, update_fine_search_candidates.cc:
        // 136-152; here an N-GPU run reproduces the single-domain one instead).
        // (1) my own old row: the partner may have been owned or a ghost then
        if (have_old && old_r != 0xffffffffu)
          for (uint32_t eo = o0; eo < o1; ++eo)
            {
              const uint32_t oc = P.old_list.col[eo];
              if ((oc & COL_INDEX_MASK) != old_r)
                continue;
              // history only survives inside the same container (periodic vs not)
              const bool old_periodic = P.use_img ? (P.old_list.img[eo] != 0) : false;
              if (old_periodic != (img != 0))
                continue;
              found = true;
              if (oc & COL_HIST_BIT)
                {
                  word |= COL_HIST_BIT;
                  P.new_list.hist[e] = P.old_list.hist[eo];
                  if (P.use_roll)
                    for (int d = 0; d < 3; ++d)
                      P.new_list.roll[3 * size_t(e) + d] = P.old_list.roll[3 * size_t(eo) + d];
                }
              break;
            }
        // (2) I immigrated and was a ghost here: the partner's old row holds this rank's copy of
        // the pair, in the opposite orientation
        if (!found && !have_old && !P.clear_history && old_q != 0xffffffffu && old_r != 0xffffffffu && old_r < P.n_old_rows)
          for (uint32_t eo = P.old_list.row_start[old_r]; eo < P.old_list.row_start[old_r + 1]; ++eo)
            {
              const uint32_t oc = P.old_list.col[eo];
              if ((oc & COL_INDEX_MASK) != old_q)
                continue;
              const bool old_periodic = P.use_img ? (P.old_list.img[eo] != 0) : false;
              if (old_periodic != (img != 0))
                continue;
              found = true;
              if (oc & COL_HIST_BIT)
                {
                  word |= COL_HIST_BIT;
                  {
                    const double4 ho = P.old_list.hist[eo];
                    P.new_list.hist[e] = make_double4(-ho.x, -ho.y, -ho.z, 0.0);
                  }
                  if (P.use_roll)
                    for (int d = 0; d < 3; ++d)
                      P.new_list.roll[3 * size_t(e) + d] = -P.old_list.roll[3 * size_t(eo) + d];
                }
              break;
            }
        // (3) I immigrated and the pair lived on the rank I came from: its history came with me
        if (!found && !have_old && !P.clear_history && P.pay.rec && qid < P.pay.map_size)
          {
            const uint32_t rid = P.pay.id[r];
            for (uint32_t k = P.pay.start[qid]; k < P.pay.n && P.pay.rec[k].qid == qid; ++k)
              {
                const HistRecord &rec = P.pay.rec[k];
                if (rec.rid != rid || (rec.flags & (HIST_REC_WALL | HIST_REC_SOLID)) || ((rec.flags & HIST_REC_PERIODIC) != 0) != (img != 0))
                  continue;
                word |= COL_HIST_BIT;
                P.new_list.hist[e] = make_double4(rec.h[0], rec.h[1], rec.h[2], 0.0);
                if (P.use_roll)
                  for (int d = 0; d < 3; ++d)
                    P.new_list.roll[3 * size_t(e) + d] = rec.roll[d];
                break;
              }
          }
        P.new_list.col[e] = word;
        P.new_list.rowl[e] = uint8_t(q & 31u);
        if (P.use_img)
          P.new_list.img[e] = uint8_t(img);
        ++e;
      };
      if (P.cand && n_mine <= NB_CACHE)
        {
          for (uint32_t k = 0; k < n_mine; ++k)
            emit(P.cand[size_t(k) * P.n_rows + q], P.use_img ? uint32_t(P.cand_img[size_t(k) * P.n_rows + q]) : 0u);
        }
      else
        for_each_neighbor(P, q, emit);
    }

    // ---- wall candidates ----
    template <class F> __device__ __forceinline__ void for_each_wall(const WallBuildParams &P, uint32_t q, F &&f)
    {
      const int lin = P.cell_reg[q];
      // particle_wall_broad_search.cc:239-246,310-317: mobile cells only
      if (P.mobility && P.mobility[lin] != LETHE_MOBILITY_MOBILE)
        return;
      if (P.faces.n_faces)
        for (uint32_t k = P.faces.cell_face_start[lin]; k < P.faces.cell_face_start[lin + 1]; ++k)
          f(k);
      if (P.floating && P.cell_fw_mask)
        {
          const uint32_t mask = P.cell_fw_mask[lin];
          if (mask)
            for (int w = 0; w < P.floating->n; ++w)
              if ((mask >> w) & 1u)
                if (P.time >= P.floating->t0[w] && P.time <= P.floating->t1[w])
                  f(WALL_FLOATING_BIT | uint32_t(w));
        }
    }

    __global__ void __launch_bounds__(256) k_count_walls(const __grid_constant__ WallBuildParams P)
    {
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= P.n_rows)
        return;
      uint32_t count = 0;
      for_each_wall(P, q, [&](uint32_t) { ++count; });
      P.counts[q] = count;
    }

    __global__ void __launch_bounds__(256) k_fill_walls(const __grid_constant__ WallBuildParams P)
    {
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= P.n_rows)
        return;
      uint32_t e = P.new_list.row_start[q];
      const uint32_t old_q = P.old_of_new ? P.old_of_new[q] : 0xffffffffu;
      const bool have_old = old_q != 0xffffffffu && old_q < P.n_old_rows;
      uint32_t o0 = 0, o1 = 0;
      if (have_old)
        {
          o0 = P.old_list.row_start[old_q];
          o1 = P.old_list.row_start[old_q + 1];
        }
      const double4 pq = P.st.pos[q];
      for_each_wall(P, q, [&](uint32_t key) {
        uint32_t word = key;
        bool found = false;
        for (uint32_t eo = o0; eo < o1 && !found; ++eo)
          {
            const uint32_t oe = P.old_list.entry[eo];
            if ((oe & (WALL_FLOATING_BIT | WALL_INDEX_MASK)) != key)
              continue;
            found = true;
            // the (particle, face) pair stayed a candidate: the contact_info survives as is,
            // including a floating wall's side (update_fine_search_candidates.cc:163-197)
            word |= (oe & WALL_FLIPPED_BIT);
            if ((oe & WALL_HIST_BIT) && !P.clear_history)
              {
                word |= WALL_HIST_BIT;
                for (int d = 0; d < 3; ++d)
                  P.new_list.hist[3 * size_t(e) + d] = P.old_list.hist[3 * size_t(eo) + d];
                if (P.use_roll)
                  for (int d = 0; d < 3; ++d)
                    P.new_list.roll[3 * size_t(e) + d] = P.old_list.roll[3 * size_t(eo) + d];
              }
          }
        if (!found && !have_old && !P.clear_history && P.pay.rec && P.pay.id[q] < P.pay.map_size)
          {
            // the (particle, face) contact_info of an immigrant came with it
            const uint32_t qid = P.pay.id[q];
            for (uint32_t k = P.pay.start[qid]; k < P.pay.n && P.pay.rec[k].qid == qid; ++k)
              {
                const HistRecord &rec = P.pay.rec[k];
                if (!(rec.flags & HIST_REC_WALL) || rec.rid != key)
                  continue;
                found = true;
                word |= WALL_HIST_BIT | ((rec.flags & HIST_REC_FLIPPED) ? WALL_FLIPPED_BIT : 0u);
                for (int d = 0; d < 3; ++d)
                  P.new_list.hist[3 * size_t(e) + d] = rec.h[d];
                if (P.use_roll)
                  for (int d = 0; d < 3; ++d)
                    P.new_list.roll[3 * size_t(e) + d] = rec.roll[d];
                break;
              }
          }
        if (!found && (key & WALL_FLOATING_BIT))
          {
            // particle_floating_wall_fine_search (particle_wall_fine_search.cc:122-140): normal
            // flipped to the particle's side when the entry is created
            const uint32_t w = key & WALL_INDEX_MASK;
            const vec3 connecting_vector =
              v3(pq.x, pq.y, pq.z) - v3(P.floating->point[w][0], P.floating->point[w][1], P.floating->point[w][2]);
            const double ip =
              dot(connecting_vector, v3(P.floating->normal[w][0], P.floating->normal[w][1], P.floating->normal[w][2]));
            if (ip < 0)
              word |= WALL_FLIPPED_BIT;
          }
        P.new_list.entry[e] = word;
        ++e;
      });
    }

    // ---- statistics ----
    __device__ __forceinline__ double warp_min(double v)
    {
      for (int o = 16; o > 0; o >>= 1)
        v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
      return v;
    }
    __device__ __forceinline__ double warp_max(double v)
    {
      for (int o = 16; o > 0; o >>= 1)
        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
      return v;
    }
    __device__ __forceinline__ double warp_sum(double v)
    {
      for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
      return v;
    }

    __global__ void __launch_bounds__(STATS_BLOCK) k_stats(StateView st, uint32_t n, double moi_override, StatsPartial *partials)
    {
      __shared__ double sm[12][STATS_BLOCK / 32];
      double vals[12] = {DBL_MAX, 0, 0, DBL_MAX, 0, 0, DBL_MAX, 0, 0, DBL_MAX, 0, 0};
      for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
        {
          const double4 x = st.pos[p], v = st.vel[p], w = st.omg[p];
          const double v2 = v.x * v.x + v.y * v.y + v.z * v.z;
          const double w2 = w.x * w.x + w.y * w.y + w.z * w.z;
          const double moi = moi_override > 0 ? moi_override : 0.1 * v.w * x.w * x.w;
          const double q[4] = {sqrt(v2), sqrt(w2), 0.5 * v.w * v2, 0.5 * moi * w2};
          for (int k = 0; k < 4; ++k)
            {
              vals[3 * k] = fmin(vals[3 * k], q[k]);
              vals[3 * k + 1] = fmax(vals[3 * k + 1], q[k]);
              vals[3 * k + 2] += q[k];
            }
        }
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      for (int k = 0; k < 12; ++k)
        {
          const double r = (k % 3 == 0) ? warp_min(vals[k]) : (k % 3 == 1 ? warp_max(vals[k]) : warp_sum(vals[k]));
          if (lane == 0)
            sm[k][warp] = r;
        }
      __syncthreads();
      if (threadIdx.x == 0)
        {
          double out[12];
          for (int k = 0; k < 12; ++k)
            {
              double r = sm[k][0];
              for (int w = 1; w < STATS_BLOCK / 32; ++w)
                r = (k % 3 == 0) ? fmin(r, sm[k][w]) : (k % 3 == 1 ? fmax(r, sm[k][w]) : r + sm[k][w]);
              out[k] = r;
            }
          StatsPartial sp = {out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7], out[8], out[9], out[10], out[11]};
          partials[blockIdx.x] = sp;
        }
    }

    // ---- host rows <-> device SoA ----
    __global__ void __launch_bounds__(256) k_unpack_host_rows(const uint32_t *ids, const double *x3, const double *props9,
                                                              uint32_t n, StateView st, uint32_t *id_out, int32_t *cell_reg,
                                                              double *disp, uint32_t base)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const double *p = props9 + 9 * size_t(k);
      const uint32_t q = base + k;
      st.pos[q] = make_double4(x3[3 * size_t(k)], x3[3 * size_t(k) + 1], x3[3 * size_t(k) + 2], p[1]);
      st.vel[q] = make_double4(p[3], p[4], p[5], p[2]);
      st.omg[q] = make_double4(p[6], p[7], p[8], p[0]);
      id_out[q] = ids[k];
      cell_reg[q] = -1; // not registered in any cell yet (no periodic wrap before the first sort)
      disp[q] = 0.0;
    }

    __global__ void __launch_bounds__(256) k_update_from_host_rows(const uint32_t *ids, const double *x3, const double *props9,
                                                                   uint32_t n, const uint32_t *slot_of_id, uint32_t map_size,
                                                                   StateView st)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const uint32_t pid = ids[k];
      if (pid >= map_size)
        return;
      const uint32_t q = slot_of_id[pid];
      if (q == 0xffffffffu)
        return;
      const double *p = props9 + 9 * size_t(k);
      st.pos[q] = make_double4(x3[3 * size_t(k)], x3[3 * size_t(k) + 1], x3[3 * size_t(k) + 2], p[1]);
      st.vel[q] = make_double4(p[3], p[4], p[5], p[2]);
      st.omg[q] = make_double4(p[6], p[7], p[8], p[0]);
    }

    // After the closing half kick the velocities have changed: the displacement the NEXT
    // iteration's contact-detection check adds (find_contact_detection_step.cc:26-50, dt |v| with
    // the velocity it finds) is accumulated here, as the step kernel does at the end of a step.
    __global__ void __launch_bounds__(256) k_accumulate_displacement(const double4 *vel, double *disp, uint32_t n, double dt,
                                                                     double criterion, uint32_t *flag_local, uint32_t *flag_host,
                                                                     uint32_t tag)
    {
      const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n)
        return;
      const double4 v = vel[i];
      const double dsp = disp[i] + dt * sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
      disp[i] = dsp;
      if (dsp > criterion && *reinterpret_cast<volatile uint32_t *>(flag_local) != tag)
        {
          *reinterpret_cast<volatile uint32_t *>(flag_local) = tag;
          if (flag_host)
            *reinterpret_cast<volatile uint32_t *>(flag_host) = tag;
        }
    }

    // External (fluid-particle interaction) loads live per particle ID; the step kernel adds one
    // per-ROW force / torque pair to the contact sums: gather the rows' loads, on top of the
    // solid-surface sums of this step when there are any.
    __global__ void __launch_bounds__(256) k_compose_external_loads(const uint32_t *id, uint32_t n, const double *ext_force,
                                                                    const double *ext_torque, uint32_t ext_size,
                                                                    const double *solid_force, const double *solid_torque,
                                                                    double *force, double *torque)
    {
      const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n)
        return;
      const uint32_t pid = id[i];
      const bool have = pid < ext_size;
#pragma unroll
      for (int d = 0; d < 3; ++d)
        {
          const double f = have ? ext_force[3 * size_t(pid) + d] : 0.0;
          const double t = have ? ext_torque[3 * size_t(pid) + d] : 0.0;
          force[3 * size_t(i) + d] = solid_force ? solid_force[3 * size_t(i) + d] + f : f;
          torque[3 * size_t(i) + d] = solid_torque ? solid_torque[3 * size_t(i) + d] + t : t;
        }
    }

    __global__ void __launch_bounds__(256) k_scatter_external_loads(const uint32_t *ids, const double *force3, const double *torque3,
                                                                    uint32_t n, double *ext_force, double *ext_torque, uint32_t ext_size)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const uint32_t pid = ids[k];
      if (pid >= ext_size)
        return;
#pragma unroll
      for (int d = 0; d < 3; ++d)
        {
          ext_force[3 * size_t(pid) + d] = force3[3 * size_t(k) + d];
          ext_torque[3 * size_t(pid) + d] = torque3 ? torque3[3 * size_t(k) + d] : 0.0;
        }
    }

    // step_host_state: the 9 doubles that a step changes (x, v, omega); diameter, mass and type
    // in the .w lanes stay what add_particles / set_particles made them
    __global__ void __launch_bounds__(256) k_update_state_rows(const uint32_t *ids, const double *state9, uint32_t n,
                                                               const uint32_t *slot_of_id, uint32_t map_size, StateView st)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const uint32_t pid = ids[k];
      if (pid >= map_size)
        return;
      const uint32_t q = slot_of_id[pid];
      if (q == 0xffffffffu)
        return;
      const double *p = state9 + 9 * size_t(k);
      st.pos[q] = make_double4(p[0], p[1], p[2], st.pos[q].w);
      st.vel[q] = make_double4(p[3], p[4], p[5], st.vel[q].w);
      st.omg[q] = make_double4(p[6], p[7], p[8], st.omg[q].w);
    }

    __global__ void __launch_bounds__(256) k_pack_state_rows(const uint32_t *ids, uint32_t n, const uint32_t *slot_of_id,
                                                             uint32_t map_size, StateView st, double *state9)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const uint32_t pid = ids[k];
      if (pid >= map_size)
        return;
      const uint32_t q = slot_of_id[pid];
      if (q == 0xffffffffu)
        return;
      const double4 x = st.pos[q], v = st.vel[q], w = st.omg[q];
      double *p = state9 + 9 * size_t(k);
      p[0] = x.x, p[1] = x.y, p[2] = x.z;
      p[3] = v.x, p[4] = v.y, p[5] = v.z;
      p[6] = w.x, p[7] = w.y, p[8] = w.z;
    }


    // ---- streamed host step (HostPlanParams, dem_kernels.cuh) ----
    __device__ __forceinline__ uint32_t plan_slot_of_row(const HostPlanParams &P, uint32_t r)
    {
      const uint32_t pid = P.row_ids[r];
      if (pid >= P.map_size)
        return 0xffffffffu;
      const uint32_t q = P.slot_of_id[pid];
      return q < P.n_owned ? q : 0xffffffffu;
    }
    template <int PASS> __global__ void __launch_bounds__(256) k_host_plan(const __grid_constant__ HostPlanParams P)
    {
      const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
      if constexpr (PASS == 2)
        {
          // one thread per list row: the latest upload stage among the row's particle and its listed neighbours
          if (t >= P.n_owned)
            return;
          uint32_t m = P.up_stage_of_slot[t];
          for (uint32_t e = P.list.row_start[t]; e < P.list.row_start[t + 1]; ++e)
            m = max(m, uint32_t(P.up_stage_of_slot[P.list.col[e] & COL_INDEX_MASK]));
          m = __reduce_max_sync(__activemask(), m); // a warp = 32 consecutive rows of one 128-row block
          if ((threadIdx.x & 31u) == 0 && m)
            atomicMax(P.block_ready + t / STEP_BLOCK_ROWS, m);
        }
      else
        {
          if (t >= P.n_rows)
            return;
          const uint32_t q = plan_slot_of_row(P, t);
          if (q == 0xffffffffu)
            return;
          const uint32_t seg = t / P.seg_rows;
          if constexpr (PASS == 0)
            P.row_of_slot[q] = t;
          else if constexpr (PASS == 1)
            P.up_stage_of_slot[q] = uint8_t(P.seg_up[seg]);
          else
            atomicMax(P.seg_down + seg, P.block_ready[q / STEP_BLOCK_ROWS]);
        }
    }

    __global__ void __launch_bounds__(256) k_update_state_rows_segs(const uint32_t *seg_list, uint32_t seg_rows, const uint32_t *ids,
                                                                    const double *state9, uint32_t n, const uint32_t *slot_of_id,
                                                                    uint32_t map_size, StateView st)
    {
      const uint32_t k = seg_list[blockIdx.y] * seg_rows + blockIdx.x * blockDim.x + threadIdx.x;
      if (blockIdx.x * blockDim.x + threadIdx.x >= seg_rows || k >= n)
        return;
      const uint32_t pid = ids[k];
      if (pid >= map_size)
        return;
      const uint32_t q = slot_of_id[pid];
      if (q == 0xffffffffu)
        return;
      const double *p = state9 + 9 * size_t(k);
      st.pos[q] = make_double4(p[0], p[1], p[2], st.pos[q].w);
      st.vel[q] = make_double4(p[3], p[4], p[5], st.vel[q].w);
      st.omg[q] = make_double4(p[6], p[7], p[8], st.omg[q].w);
    }

    __global__ void __launch_bounds__(256) k_pack_state_rows_segs(const uint32_t *seg_list, uint32_t seg_rows, const uint32_t *ids, uint32_t n,
                                                                  const uint32_t *slot_of_id, uint32_t map_size, StateView st, double *state9)
    {
      const uint32_t k = seg_list[blockIdx.y] * seg_rows + blockIdx.x * blockDim.x + threadIdx.x;
      if (blockIdx.x * blockDim.x + threadIdx.x >= seg_rows || k >= n)
        return;
      const uint32_t pid = ids[k];
      if (pid >= map_size)
        return;
      const uint32_t q = slot_of_id[pid];
      if (q == 0xffffffffu)
        return;
      const double4 x = st.pos[q], v = st.vel[q], w = st.omg[q];
      double *p = state9 + 9 * size_t(k);
      p[0] = x.x, p[1] = x.y, p[2] = x.z;
      p[3] = v.x, p[4] = v.y, p[5] = v.z;
      p[6] = w.x, p[7] = w.y, p[8] = w.z;
    }

    // one thread block = one 128-slot block of the step kernel, one warp = 32 consecutive slots; the 32 x 9 doubles are
    // transposed through shared memory so that consecutive lanes store consecutive words of the host rows
    __global__ void __launch_bounds__(128) k_pack_state_rows_blocks(const uint32_t *block_list, const uint32_t *row_of_slot, uint32_t n_owned,
                                                                    StateView st, double *state9)
    {
      __shared__ double buf[4][32 * 9];
      __shared__ uint32_t rows[4][32];
      const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
      const uint32_t q = block_list[blockIdx.x] * STEP_BLOCK_ROWS + threadIdx.x;
      uint32_t r = 0xffffffffu;
      if (q < n_owned)
        {
          r = row_of_slot[q];
          const double4 x = st.pos[q], v = st.vel[q], w = st.omg[q];
          double *b = buf[warp] + 9 * lane;
          b[0] = x.x, b[1] = x.y, b[2] = x.z;
          b[3] = v.x, b[4] = v.y, b[5] = v.z;
          b[6] = w.x, b[7] = w.y, b[8] = w.z;
        }
      rows[warp][lane] = r;
      __syncwarp();
#pragma unroll
      for (uint32_t k = 0; k < 9; ++k)
        {
          const uint32_t t = 32 * k + lane, rl = t / 9, comp = t - 9 * rl;
          const uint32_t rr = rows[warp][rl];
          if (rr != 0xffffffffu)
            state9[9 * size_t(rr) + comp] = buf[warp][t];
        }
    }

    __global__ void __launch_bounds__(256) k_layer_keys(const int32_t *cell_reg, GridDesc g, int axis, uint32_t n, uint16_t *keys, uint32_t *vals)
    {
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= n)
        return;
      const int cell = cell_reg[q];
      int layer = 0;
      if (cell >= 0)
        layer = 1 + (axis == 0 ? cell % g.n[0] : (axis == 1 ? (cell / g.n[0]) % g.n[1] : cell / (g.n[0] * g.n[1])));
      keys[q] = uint16_t(min(layer, 65535));
      vals[q] = q;
    }

    __global__ void __launch_bounds__(256) k_pack_state_rows_perm(const uint32_t *perm, uint32_t n, StateView st, const uint32_t *id,
                                                                  uint32_t *ids_out, double *state9)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const uint32_t q = perm[k];
      const double4 x = st.pos[q], v = st.vel[q], w = st.omg[q];
      if (ids_out)
        ids_out[k] = id[q];
      if (state9)
        {
          double *p = state9 + 9 * size_t(k);
          p[0] = x.x, p[1] = x.y, p[2] = x.z;
          p[3] = v.x, p[4] = v.y, p[5] = v.z;
          p[6] = w.x, p[7] = w.y, p[8] = w.z;
        }
    }

    __device__ __forceinline__ void write_row(double4 x, double4 v, double4 w, double *x3, double *props9, size_t k)
    {
      x3[3 * k] = x.x;
      x3[3 * k + 1] = x.y;
      x3[3 * k + 2] = x.z;
      double *p = props9 + 9 * k;
      p[0] = w.w;
      p[1] = x.w;
      p[2] = v.w;
      p[3] = v.x;
      p[4] = v.y;
      p[5] = v.z;
      p[6] = w.x;
      p[7] = w.y;
      p[8] = w.z;
    }

    __global__ void __launch_bounds__(256) k_pack_host_rows(const uint32_t *ids, uint32_t n, const uint32_t *slot_of_id,
                                                            uint32_t map_size, StateView st, double *x3, double *props9)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const uint32_t pid = ids[k];
      if (pid >= map_size)
        return;
      const uint32_t q = slot_of_id[pid];
      if (q == 0xffffffffu)
        return;
      write_row(st.pos[q], st.vel[q], st.omg[q], x3, props9, k);
    }

    __global__ void __launch_bounds__(256) k_pack_all_rows(StateView st, const uint32_t *id, uint32_t n, uint32_t *ids_out,
                                                           double *x3, double *props9)
    {
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= n)
        return;
      ids_out[q] = id[q];
      write_row(st.pos[q], st.vel[q], st.omg[q], x3, props9, q);
    }

    // ---- multi-GPU helpers ----
    __global__ void __launch_bounds__(256) k_classify(const __grid_constant__ ClassifyParams P)
    {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      if (p >= P.n)
        return;
      const GridDesc &g = P.grid;
      double4 x = P.pos[p];
      const int creg = P.cell_reg[p];
      if ((g.periodic[0] | g.periodic[1] | g.periodic[2]) && creg >= 0)
        {
          // same periodic displacement as k_bin (idempotent: the wrapped particle is then inside)
          const int c[3] = {creg % g.n[0], (creg / g.n[0]) % g.n[1], creg / (g.n[0] * g.n[1])};
          double xv[3] = {x.x, x.y, x.z};
#pragma unroll
          for (int d = 0; d < 3; ++d)
            {
              if (!g.periodic[d])
                continue;
              const double lo = g.lo[d];
              const double hi = g.lo[d] + g.n[d] * g.h[d];
              if (c[d] == 0 && (xv[d] - lo) * -1.0 >= 0.0)
                xv[d] += g.L[d];
              if (c[d] == g.n[d] - 1 && (xv[d] - hi) * 1.0 >= 0.0)
                xv[d] += -g.L[d];
            }
          x.x = xv[0];
          x.y = xv[1];
          x.z = xv[2];
          P.pos[p] = x;
          P.cell_reg[p] = -1; // wrapped already: k_bin must not wrap it a second time
        }
      const int lin = cell_of_point(g, x.x, x.y, x.z);
      if (lin < 0)
        return; // left the domain: the sort drops it
      const int a = g.slab_axis;
      const int ca = a == 0 ? lin % g.n[0] : (a == 1 ? (lin / g.n[0]) % g.n[1] : lin / (g.n[0] * g.n[1]));
      if (ca >= g.slab_lo && ca < g.slab_hi)
        return;
      const int na = g.n[a];
      // layers below the slab's first / above its last one (periodic: the short way round)
      int below = g.slab_lo - ca, above = ca - (g.slab_hi - 1);
      if (g.periodic[a])
        {
          below = ((below % na) + na) % na;
          above = ((above % na) + na) % na;
        }
      int dir;
      if (below > 0 && below <= P.max_hop && (above <= 0 || below <= above))
        dir = 0;
      else if (above > 0 && above <= P.max_hop)
        dir = 1;
      else
        {
          atomicAdd(&P.send_count[2], 1u); // moved further than the neighbouring slab: not supported
          P.cell_reg[p] = -2;
          return;
        }
      const uint32_t k = atomicAdd(&P.send_count[dir], 1u);
      if (k < P.send_cap)
        {
          MigrateRecord r;
          r.pos = x;
          r.vel = P.vel[p];
          r.omg = P.omg[p];
          P.send_rec[dir][k] = r;
          P.send_id[dir][k] = P.id[p];
          P.send_slot[dir][k] = p;
        }
      P.cell_reg[p] = -2;
    }

    __global__ void __launch_bounds__(256) k_append_records(const MigrateRecord *rec, const uint32_t *ids, uint32_t n, StateView st,
                                                            uint32_t *id_out, int32_t *cell_reg, double *disp, uint32_t base)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const uint32_t q = base + k;
      st.pos[q] = rec[k].pos;
      st.vel[q] = rec[k].vel;
      st.omg[q] = rec[k].omg;
      id_out[q] = ids[k];
      cell_reg[q] = -1;
      disp[q] = 0.0;
    }

    __global__ void __launch_bounds__(256) k_layer_histogram(const double4 *pos, GridDesc g, uint32_t n, uint32_t *hist)
    {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      if (p >= n)
        return;
      const double4 x = pos[p];
      const int a = g.slab_axis;
      const double xa = a == 0 ? x.x : (a == 1 ? x.y : x.z);
      int layer = int(floor((xa - g.lo[a]) / g.h[a]));
      layer = min(max(layer, 0), g.n[a] - 1);
      atomicAdd(hist + layer, 1u);
    }

    __global__ void __launch_bounds__(256) k_flag_layer(const int32_t *cell_reg, GridDesc g, int layer_cell, uint32_t n, uint32_t *flags)
    {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      if (p > n)
        return;
      if (p == n)
        {
          flags[p] = 0;
          return;
        }
      const int lin = cell_reg[p];
      const int a = g.slab_axis;
      const int ca = a == 0 ? lin % g.n[0] : (a == 1 ? (lin / g.n[0]) % g.n[1] : lin / (g.n[0] * g.n[1]));
      flags[p] = ca == layer_cell ? 1u : 0u;
    }

    __global__ void __launch_bounds__(256) k_compact_indices(const uint32_t *flags, const uint32_t *offsets, uint32_t n, uint32_t *out)
    {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      if (p < n && flags[p])
        out[offsets[p]] = p;
    }

    __global__ void __launch_bounds__(256) k_gather_state(StateView st, const uint32_t *idx, uint32_t n, double4 *pos, double4 *vel,
                                                          double4 *omg)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const uint32_t p = idx[k];
      pos[k] = st.pos[p];
      vel[k] = st.vel[p];
      omg[k] = st.omg[p];
    }

    __global__ void __launch_bounds__(256) k_gather_ids(const uint32_t *id, const uint32_t *idx, uint32_t n, uint32_t *out)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k < n)
        out[k] = id[idx[k]];
    }

    __global__ void __launch_bounds__(256) k_ghost_run(const __grid_constant__ GhostRunParams P)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= P.n)
        return;
      const double4 x = P.pos[k];
      const int lin = cell_of_point(P.grid, x.x, x.y, x.z);
      const uint32_t q = P.base + k;
      P.cell_reg[q] = lin;
      // the run arrives sorted by (curve rank, id): cell boundaries are where the rank changes
      const uint32_t rank = lin >= 0 ? uint32_t(P.cell_rank[lin]) : 0xffffffffu;
      uint32_t prev = 0xffffffffu, next = 0xffffffffu;
      if (k > 0)
        {
          const double4 xp = P.pos[k - 1];
          const int lp = cell_of_point(P.grid, xp.x, xp.y, xp.z);
          prev = lp >= 0 ? uint32_t(P.cell_rank[lp]) : 0xffffffffu;
        }
      if (k + 1 < P.n)
        {
          const double4 xn = P.pos[k + 1];
          const int ln = cell_of_point(P.grid, xn.x, xn.y, xn.z);
          next = ln >= 0 ? uint32_t(P.cell_rank[ln]) : 0xffffffffu;
        }
      if (rank != 0xffffffffu)
        {
          if (k == 0 || prev != rank)
            P.start[rank] = q;
          if (k + 1 == P.n || next != rank)
            P.end[rank] = q + 1;
        }
      // history source: the particle's slot in the previous list generation, whether it was a
      // ghost here already or an owned particle that has just emigrated
      uint32_t old = 0xffffffffu;
      const uint32_t pid = P.id[q];
      if (P.old_slot_of_id && pid < P.old_map_size)
        old = P.old_slot_of_id[pid];
      P.old_of_new[q] = old;
    }

    // ---- adaptive sparse contacts ----
    __device__ __forceinline__ int asc_node(const GridDesc &g, int i, int j, int k)
    {
      // periodic directions share the nodes of their two faces (periodic_node_ids)
      if (g.periodic[0] && i == g.n[0])
        i = 0;
      if (g.periodic[1] && j == g.n[1])
        j = 0;
      if (g.periodic[2] && k == g.n[2])
        k = 0;
      return i + (g.n[0] + 1) * (j + (g.n[1] + 1) * k);
    }
    __device__ __forceinline__ bool asc_any_node(const AscParams &P, int lin, int value)
    {
      const GridDesc &g = P.grid;
      const int ci = lin % g.n[0], cj = (lin / g.n[0]) % g.n[1], ck = lin / (g.n[0] * g.n[1]);
      bool any = false;
      for (int v = 0; v < 8; ++v)
        any |= P.node_status[asc_node(g, ci + (v & 1), cj + ((v >> 1) & 1), ck + (v >> 2))] == value;
      return any;
    }
    // assign_mobility_status (adaptive_sparse_contacts.h:389-412): the prevailing (larger) status stays at a node
    __device__ __forceinline__ void asc_raise_nodes(const AscParams &P, int lin, int value)
    {
      const GridDesc &g = P.grid;
      const int ci = lin % g.n[0], cj = (lin / g.n[0]) % g.n[1], ck = lin / (g.n[0] * g.n[1]);
      for (int v = 0; v < 8; ++v)
        atomicMax(P.node_status + asc_node(g, ci + (v & 1), cj + ((v >> 1) & 1), ck + (v >> 2)), value);
    }

    // One thread per cell; the four passes are separate launches because each reads the node
    // values the previous one wrote. Within a pass the node value tested is never one the pass writes.
    template <int PASS> __global__ void __launch_bounds__(128) k_asc_cells(const __grid_constant__ AscParams P)
    {
      const int lin = blockIdx.x * blockDim.x + threadIdx.x;
      if (lin >= P.grid.n_cells)
        return;
      if (P.owned_lo >= 0)
        {
          const GridDesc &g = P.grid;
          const int a = g.slab_axis;
          const int ca = a == 0 ? lin % g.n[0] : (a == 1 ? (lin / g.n[0]) % g.n[1] : lin / (g.n[0] * g.n[1]));
          if (ca < P.owned_lo || ca >= P.owned_hi)
            {
              if (PASS == 0)
                P.cell_status[lin] = ASC_FOREIGN;
              return;
            }
        }
      const uint32_t rank = uint32_t(P.cell_rank[lin]);
      const uint32_t p0 = P.cell_start[rank], p1 = P.cell_start[rank + 1];
      if constexpr (PASS == 0)
        {
          // 1. empty cells: cell inactive, nodes empty
          if (p1 == p0)
            {
              P.cell_status[lin] = LETHE_MOBILITY_INACTIVE;
              asc_raise_nodes(P, lin, LETHE_MOBILITY_EMPTY_NODE);
            }
          else
            P.cell_status[lin] = ASC_UNASSIGNED;
          return;
        }
      if (P.cell_status[lin] != ASC_UNASSIGNED)
        return;
      if constexpr (PASS == 1)
        {
          // 2. mobile by criteria: granular temperature, solid fraction (calculate_granular_temperature_
          // and_solid_fraction, adaptive_sparse_contacts.cc:35-130), next to an empty cell
          const unsigned int n = p1 - p0;
          double solid_volume = 0.0;
          double va[3] = {0, 0, 0};
          for (uint32_t r = p0; r < p1; ++r)
            {
              const double4 x = P.st.pos[r], v = P.st.vel[r];
              va[0] += v.x;
              va[1] += v.y;
              va[2] += v.z;
              solid_volume += M_PI * (x.w * x.w * x.w) / (2.0 * 3);
            }
          const double inv = 1.0 / n;
          for (int d = 0; d < 3; ++d)
            va[d] *= inv;
          const double solid_fraction = solid_volume / (P.grid.h[0] * P.grid.h[1] * P.grid.h[2]);
          double fl[3] = {0, 0, 0};
          for (uint32_t r = p0; r < p1; ++r)
            {
              const double4 v = P.st.vel[r];
              const double f0 = v.x - va[0], f1 = v.y - va[1], f2 = v.z - va[2];
              fl[0] += f0 * f0;
              fl[1] += f1 * f1;
              fl[2] += f2 * f2;
            }
          double granular_temperature = 0.0;
          for (int d = 0; d < 3; ++d)
            {
              fl[d] /= n;
              granular_temperature += fl[d] / 3;
            }
          if (granular_temperature > P.granular_temperature_threshold || solid_fraction < P.solid_fraction_threshold ||
              asc_any_node(P, lin, LETHE_MOBILITY_EMPTY_NODE))
            {
              P.cell_status[lin] = LETHE_MOBILITY_MOBILE;
              asc_raise_nodes(P, lin, LETHE_MOBILITY_MOBILE);
            }
        }
      else if constexpr (PASS == 2)
        {
          // 3. the additional mobile layer: a node made mobile by pass 1; its other nodes become active
          if (asc_any_node(P, lin, LETHE_MOBILITY_MOBILE))
            {
              P.cell_status[lin] = LETHE_MOBILITY_MOBILE;
              asc_raise_nodes(P, lin, LETHE_MOBILITY_STATIC_ACTIVE);
            }
        }
      else
        // 4. the active layer; the rest is inactive
        P.cell_status[lin] = asc_any_node(P, lin, LETHE_MOBILITY_STATIC_ACTIVE) ? LETHE_MOBILITY_STATIC_ACTIVE : LETHE_MOBILITY_INACTIVE;
    }

    // the two in-plane axes of a plane / layer across the slab axis
    __device__ __forceinline__ void asc_plane_axes(int a, int &u, int &v)
    {
      u = a == 0 ? 1 : 0;
      v = a == 2 ? 1 : 2;
    }
    template <int OP> __global__ void __launch_bounds__(256) k_asc_plane(const __grid_constant__ AscPlaneParams P)
    {
      const GridDesc &g = P.grid;
      const int a = g.slab_axis;
      int u, v;
      asc_plane_axes(a, u, v);
      const bool nodes = OP < 2;
      const int nu = g.n[u] + (nodes ? 1 : 0), nv = g.n[v] + (nodes ? 1 : 0);
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= nu * nv)
        return;
      int c[3];
      c[a] = P.index;
      c[u] = t % nu;
      c[v] = t / nu;
      if (nodes)
        {
          const int node = asc_node(g, c[0], c[1], c[2]);
          if (OP == 0)
            P.buf[t] = P.node_status[node];
          else
            atomicMax(P.node_status + node, P.buf[t]); // periodic in-plane nodes alias each other
        }
      else
        {
          const int lin = c[0] + g.n[0] * (c[1] + g.n[1] * c[2]);
          if (OP == 2)
            P.buf[t] = int(P.cell_status[lin]);
          else
            P.cell_status[lin] = uint8_t(P.buf[t]);
        }
    }

    __global__ void __launch_bounds__(256) k_layer_histogram_weighted(const double4 *pos, const int32_t *cell_reg, const uint8_t *cell_status,
                                                                      GridDesc g, uint32_t n, uint32_t w_mobile, uint32_t w_active,
                                                                      uint32_t w_inactive, uint32_t *hist)
    {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      if (p >= n)
        return;
      const double4 x = pos[p];
      const int a = g.slab_axis;
      const double xa = a == 0 ? x.x : (a == 1 ? x.y : x.z);
      int layer = int(floor((xa - g.lo[a]) / g.h[a]));
      layer = min(max(layer, 0), g.n[a] - 1);
      const int lin = cell_reg[p];
      const uint8_t st = lin >= 0 ? cell_status[lin] : uint8_t(LETHE_MOBILITY_MOBILE);
      atomicAdd(hist + layer, st == LETHE_MOBILITY_MOBILE ? w_mobile : (st == LETHE_MOBILITY_STATIC_ACTIVE ? w_active : w_inactive));
    }

    __global__ void __launch_bounds__(256) k_asc_rows(const __grid_constant__ AscParams P)
    {
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q < P.n_rows)
        P.row_mobile[q] = P.cell_status[P.cell_reg[q]] == LETHE_MOBILITY_MOBILE ? 1 : 0;
    }

    __global__ void __launch_bounds__(256) k_register_ids(const uint32_t *id, uint32_t base, uint32_t n, uint32_t *slot_of_id,
                                                          uint32_t map_size)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const uint32_t pid = id[base + k];
      if (pid < map_size)
        slot_of_id[pid] = base + k;
    }

    __global__ void __launch_bounds__(256) k_fill_u32(uint32_t *p, uint32_t v, size_t n)
    {
      for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        p[i] = v;
    }

    inline unsigned blocks_for(size_t n, unsigned bs) { return unsigned((n + bs - 1) / bs); }
  } // namespace

  static std::atomic<unsigned long long> g_launches{0};
  void count_launch(unsigned n) { g_launches += n; }
  unsigned long long launch_count() { return g_launches.load(); }

  size_t scan_tmp_elems(size_t n)
  {
    size_t total = 0;
    size_t m = n;
    while (m > 1)
      {
        m = (m + SCAN_TILE - 1) / SCAN_TILE;
        total += 2 * (m + 1);
        if (m == 1)
          break;
      }
    return total + 8;
  }

  static void scan_rec(const uint32_t *in, uint32_t *out, size_t n, size_t n_valid, uint32_t *tmp, cudaStream_t s)
  {
    const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t *sums = tmp;
    uint32_t *sums_scanned = tmp + (tiles + 1);
    k_scan_tiles<<<unsigned(tiles), SCAN_THREADS, 0, s>>>(in, out, n, n_valid, sums);
    count_launch();
    if (tiles > 1)
      {
        scan_rec(sums, sums_scanned, tiles, tiles, tmp + 2 * (tiles + 1), s);
        k_scan_add<<<unsigned(tiles), SCAN_THREADS, 0, s>>>(out, n, sums_scanned);
        count_launch();
      }
  }

  void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n_plus_1, uint32_t *tmp, cudaStream_t s)
  {
    if (n_plus_1 == 0)
      return;
    scan_rec(in, out, n_plus_1, n_plus_1 - 1, tmp, s);
  }

  void launch_bin(const BinParams &p, cudaStream_t s)
  {
    if (p.n)
      {
        k_bin<<<blocks_for(p.n, 256), 256, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_scatter_perm(const uint32_t *key, const uint32_t *slot, const uint32_t *cell_start, uint32_t *perm, uint32_t n,
                           cudaStream_t s)
  {
    if (n)
      {
        k_scatter_perm<<<blocks_for(n, 256), 256, 0, s>>>(key, slot, cell_start, perm, n);
        count_launch();
      }
  }
  void launch_sort_cells(const uint32_t *cell_start, uint32_t n_buckets, uint32_t *perm, const uint32_t *id, cudaStream_t s)
  {
    if (n_buckets)
      {
        k_sort_cells<<<blocks_for(n_buckets, 256), 256, 0, s>>>(cell_start, n_buckets, perm, id);
        count_launch();
      }
  }
  void launch_gather(const GatherParams &p, cudaStream_t s)
  {
    if (p.n_new)
      {
        k_gather<<<blocks_for(p.n_new, 256), 256, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_count_neighbors(const NeighborParams &p, cudaStream_t s)
  {
    if (p.n_rows)
      {
        k_count_neighbors<<<blocks_for(p.n_rows, 128), 128, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_fill_neighbors(const NeighborParams &p, cudaStream_t s)
  {
    if (p.n_rows)
      {
        k_fill_neighbors<<<blocks_for(p.n_rows, 128), 128, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_count_walls(const WallBuildParams &p, cudaStream_t s)
  {
    if (p.n_rows)
      {
        k_count_walls<<<blocks_for(p.n_rows, 256), 256, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_fill_walls(const WallBuildParams &p, cudaStream_t s)
  {
    if (p.n_rows)
      {
        k_fill_walls<<<blocks_for(p.n_rows, 256), 256, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_stats(StateView st, uint32_t n, double moi_override, StatsPartial *partials, uint32_t n_blocks, cudaStream_t s)
  {
    k_stats<<<n_blocks, STATS_BLOCK, 0, s>>>(st, n, moi_override, partials);
    count_launch();
  }
  void launch_unpack_host_rows(const uint32_t *ids, const double *x3, const double *props9, uint32_t n, StateView st,
                               uint32_t *id_out, int32_t *cell_reg, double *disp, uint32_t base, cudaStream_t s)
  {
    if (n)
      {
        k_unpack_host_rows<<<blocks_for(n, 256), 256, 0, s>>>(ids, x3, props9, n, st, id_out, cell_reg, disp, base);
        count_launch();
      }
  }
  void launch_update_from_host_rows(const uint32_t *ids, const double *x3, const double *props9, uint32_t n,
                                    const uint32_t *slot_of_id, uint32_t slot_map_size, StateView st, cudaStream_t s)
  {
    if (n)
      {
        k_update_from_host_rows<<<blocks_for(n, 256), 256, 0, s>>>(ids, x3, props9, n, slot_of_id, slot_map_size, st);
        count_launch();
      }
  }
  void launch_pack_host_rows(const uint32_t *ids, uint32_t n, const uint32_t *slot_of_id, uint32_t slot_map_size, StateView st,
                             double *x3, double *props9, cudaStream_t s)
  {
    if (n)
      {
        k_pack_host_rows<<<blocks_for(n, 256), 256, 0, s>>>(ids, n, slot_of_id, slot_map_size, st, x3, props9);
        count_launch();
      }
  }
  void launch_accumulate_displacement(const double4 *vel, double *disp, uint32_t n, double dt, double criterion, uint32_t *flag_local,
                                      uint32_t *flag_host, uint32_t tag, cudaStream_t s)
  {
    if (n)
      {
        k_accumulate_displacement<<<blocks_for(n, 256), 256, 0, s>>>(vel, disp, n, dt, criterion, flag_local, flag_host, tag);
        count_launch();
      }
  }
  void launch_compose_external_loads(const uint32_t *id, uint32_t n, const double *ext_force, const double *ext_torque,
                                     uint32_t ext_size, const double *solid_force, const double *solid_torque, double *force,
                                     double *torque, cudaStream_t s)
  {
    if (n)
      {
        k_compose_external_loads<<<blocks_for(n, 256), 256, 0, s>>>(id, n, ext_force, ext_torque, ext_size, solid_force, solid_torque,
                                                                    force, torque);
        count_launch();
      }
  }
  void launch_scatter_external_loads(const uint32_t *ids, const double *force3, const double *torque3, uint32_t n, double *ext_force,
                                     double *ext_torque, uint32_t ext_size, cudaStream_t s)
  {
    if (n)
      {
        k_scatter_external_loads<<<blocks_for(n, 256), 256, 0, s>>>(ids, force3, torque3, n, ext_force, ext_torque, ext_size);
        count_launch();
      }
  }
  void launch_update_state_rows(const uint32_t *ids, const double *state9, uint32_t n, const uint32_t *slot_of_id,
                                uint32_t slot_map_size, StateView st, cudaStream_t s)
  {
    if (n)
      {
        k_update_state_rows<<<blocks_for(n, 256), 256, 0, s>>>(ids, state9, n, slot_of_id, slot_map_size, st);
        count_launch();
      }
  }
  void launch_pack_state_rows(const uint32_t *ids, uint32_t n, const uint32_t *slot_of_id, uint32_t slot_map_size, StateView st,
                              double *state9, cudaStream_t s)
  {
    if (n)
      {
        k_pack_state_rows<<<blocks_for(n, 256), 256, 0, s>>>(ids, n, slot_of_id, slot_map_size, st, state9);
        count_launch();
      }
  }
  void launch_host_plan(const HostPlanParams &p, int pass, cudaStream_t s)
  {
    const uint32_t n = pass == 2 ? p.n_owned : p.n_rows;
    if (!n)
      return;
    const unsigned blocks = blocks_for(n, 256);
    switch (pass)
      {
        case 0:
          k_host_plan<0><<<blocks, 256, 0, s>>>(p);
          break;
        case 1:
          k_host_plan<1><<<blocks, 256, 0, s>>>(p);
          break;
        case 2:
          k_host_plan<2><<<blocks, 256, 0, s>>>(p);
          break;
        default:
          k_host_plan<3><<<blocks, 256, 0, s>>>(p);
          break;
      }
    count_launch();
  }
  void launch_update_state_rows_segs(const uint32_t *seg_list, uint32_t n_segs, uint32_t seg_rows, const uint32_t *ids, const double *state9,
                                     uint32_t n, const uint32_t *slot_of_id, uint32_t slot_map_size, StateView st, cudaStream_t s)
  {
    // blockIdx.y = segment: at most 65535 of them per launch (the planner keeps the segment count far below that)
    if (n_segs)
      {
        k_update_state_rows_segs<<<dim3(blocks_for(seg_rows, 256), n_segs), 256, 0, s>>>(seg_list, seg_rows, ids, state9, n, slot_of_id,
                                                                                        slot_map_size, st);
        count_launch();
      }
  }
  void launch_pack_state_rows_segs(const uint32_t *seg_list, uint32_t n_segs, uint32_t seg_rows, const uint32_t *ids, uint32_t n,
                                   const uint32_t *slot_of_id, uint32_t slot_map_size, StateView st, double *state9, cudaStream_t s)
  {
    if (n_segs)
      {
        k_pack_state_rows_segs<<<dim3(blocks_for(seg_rows, 256), n_segs), 256, 0, s>>>(seg_list, seg_rows, ids, n, slot_of_id, slot_map_size,
                                                                                      st, state9);
        count_launch();
      }
  }
  void transfer_order_perm(const int32_t *cell_reg, GridDesc grid, int axis, uint32_t n, uint32_t *perm, cudaStream_t s)
  {
    if (!n)
      return;
    uint16_t *keys = nullptr, *keys_out = nullptr;
    uint32_t *vals = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    auto check = [](cudaError_t e) {
      if (e != cudaSuccess)
        throw std::runtime_error(std::string("transfer_order_perm: ") + cudaGetErrorString(e));
    };
    check(cudaMalloc(&keys, size_t(n) * 2));
    check(cudaMalloc(&keys_out, size_t(n) * 2));
    check(cudaMalloc(&vals, size_t(n) * 4));
    k_layer_keys<<<blocks_for(n, 256), 256, 0, s>>>(cell_reg, grid, axis, n, keys, vals);
    count_launch();
    // LSD radix sort: stable, so the cell-sorted slot order survives inside a layer
    check(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, vals, perm, int(n), 0, 16, s));
    check(cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16)));
    check(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, vals, perm, int(n), 0, 16, s));
    check(cudaStreamSynchronize(s));
    cudaFree(tmp);
    cudaFree(vals);
    cudaFree(keys_out);
    cudaFree(keys);
  }
  void launch_pack_state_rows_perm(const uint32_t *perm, uint32_t n, StateView st, const uint32_t *id, uint32_t *ids_out, double *state9,
                                   cudaStream_t s)
  {
    if (n)
      {
        k_pack_state_rows_perm<<<blocks_for(n, 256), 256, 0, s>>>(perm, n, st, id, ids_out, state9);
        count_launch();
      }
  }
  void launch_pack_state_rows_blocks(const uint32_t *block_list, uint32_t n_blocks, const uint32_t *row_of_slot, uint32_t n_owned,
                                     StateView st, double *state9, cudaStream_t s)
  {
    if (n_blocks)
      {
        k_pack_state_rows_blocks<<<n_blocks, 128, 0, s>>>(block_list, row_of_slot, n_owned, st, state9);
        count_launch();
      }
  }
  void launch_pack_all_rows(StateView st, const uint32_t *id, uint32_t n, uint32_t *ids_out, double *x3, double *props9,
                            cudaStream_t s)
  {
    if (n)
      {
        k_pack_all_rows<<<blocks_for(n, 256), 256, 0, s>>>(st, id, n, ids_out, x3, props9);
        count_launch();
      }
  }
  namespace
  {
    template <bool PACK> __global__ void __launch_bounds__(128) k_hist_rows(const __grid_constant__ HistPackParams P)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= P.n)
        return;
      const uint32_t p = P.send_slot[k];
      uint32_t n = 0;
      const uint32_t base = PACK ? P.counts[k] : 0u;
      const uint32_t qid = P.id[p];
      if (p < P.n_rows)
        for (uint32_t e = P.list.row_start[p]; e < P.list.row_start[p + 1]; ++e)
          {
            const uint32_t c = P.list.col[e];
            if (!(c & COL_HIST_BIT))
              continue;
            if (PACK)
              {
                HistRecord r;
                r.qid = qid;
                r.rid = P.id[c & COL_INDEX_MASK];
                r.flags = (P.use_img && P.list.img[e]) ? HIST_REC_PERIODIC : 0u;
                r.pad = 0;
                const double4 hh = P.list.hist[e];
                r.h[0] = hh.x;
                r.h[1] = hh.y;
                r.h[2] = hh.z;
                for (int d = 0; d < 3; ++d)
                  r.roll[d] = P.use_roll ? P.list.roll[3 * size_t(e) + d] : 0.0;
                P.out[base + n] = r;
              }
            ++n;
          }
      if (p < P.n_wall_rows)
        for (uint32_t w = P.walls.row_start[p]; w < P.walls.row_start[p + 1]; ++w)
          {
            const uint32_t we = P.walls.entry[w];
            if (!(we & WALL_HIST_BIT))
              continue;
            if (PACK)
              {
                HistRecord r;
                r.qid = qid;
                r.rid = we & (WALL_FLOATING_BIT | WALL_INDEX_MASK);
                r.flags = HIST_REC_WALL | ((we & WALL_FLIPPED_BIT) ? HIST_REC_FLIPPED : 0u);
                r.pad = 0;
                for (int d = 0; d < 3; ++d)
                  {
                    r.h[d] = P.walls.hist[3 * size_t(w) + d];
                    r.roll[d] = P.use_roll ? P.walls.roll[3 * size_t(w) + d] : 0.0;
                  }
                P.out[base + n] = r;
              }
            ++n;
          }
      if (P.solid_row_start && p < P.n_solid_rows)
        for (uint32_t w = P.solid_row_start[p]; w < P.solid_row_start[p + 1]; ++w)
          {
            const uint32_t se = P.solid_entry[w];
            if (!(se & 0x80000000u))
              continue;
            if (PACK)
              {
                HistRecord r;
                r.qid = qid;
                r.rid = se & 0x7fffffffu;
                r.flags = HIST_REC_SOLID;
                r.pad = 0;
                for (int d = 0; d < 3; ++d)
                  {
                    r.h[d] = P.solid_hist[3 * size_t(w) + d];
                    r.roll[d] = P.use_roll ? P.solid_roll[3 * size_t(w) + d] : 0.0;
                  }
                P.out[base + n] = r;
              }
            ++n;
          }
      if (!PACK)
        P.counts[k] = n;
    }
    __global__ void __launch_bounds__(256) k_hist_index(const HistRecord *rec, uint32_t n, uint32_t *pay_start, uint32_t map_size)
    {
      const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const uint32_t qid = rec[k].qid;
      if ((k == 0 || rec[k - 1].qid != qid) && qid < map_size)
        pay_start[qid] = k;
    }
  } // namespace
  void launch_hist_count(const HistPackParams &p, cudaStream_t s)
  {
    if (p.n)
      {
        k_hist_rows<false><<<blocks_for(p.n, 128), 128, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_hist_pack(const HistPackParams &p, cudaStream_t s)
  {
    if (p.n)
      {
        k_hist_rows<true><<<blocks_for(p.n, 128), 128, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_hist_index(const HistRecord *rec, uint32_t n, uint32_t *pay_start, uint32_t map_size, cudaStream_t s)
  {
    if (n)
      {
        k_hist_index<<<blocks_for(n, 256), 256, 0, s>>>(rec, n, pay_start, map_size);
        count_launch();
      }
  }
  void launch_layer_histogram(const double4 *pos, GridDesc grid, uint32_t n, uint32_t *hist, cudaStream_t s)
  {
    if (!n)
      return;
    k_layer_histogram<<<(n + 255) / 256, 256, 0, s>>>(pos, grid, n, hist);
    count_launch(1);
  }

  void balanced_cuts(int n_layers, const uint64_t *hist, int world, const int32_t *cuts, int max_shift, int min_width, int32_t *new_cuts)
  {
    std::vector<uint64_t> cum(size_t(n_layers) + 1, 0);
    for (int l = 0; l < n_layers; ++l)
      cum[l + 1] = cum[l] + hist[l];
    const uint64_t total = cum[n_layers];
    new_cuts[0] = 0;
    new_cuts[world] = n_layers;
    for (int r = 1; r < world; ++r)
      {
        // first layer boundary at which the particles below reach r / world of the total
        const double target = double(total) * r / world;
        int e = int(std::lower_bound(cum.begin(), cum.end(), target, [](uint64_t c, double t) { return double(c) < t; }) - cum.begin());
        // pick the nearer of the two boundaries around the target
        if (e > 0 && e <= n_layers && (target - double(cum[e - 1])) < (double(cum[std::min(e, n_layers)]) - target))
          --e;
        // a cut stays strictly inside the two slabs it separates (so that every particle that changes owner goes to an
        // adjacent rank), and moves by at most max_shift layers
        const int down = std::min(max_shift, cuts[r] - cuts[r - 1] - 1), up = std::min(max_shift, cuts[r + 1] - cuts[r] - 1);
        e = std::max(cuts[r] - std::max(down, 0), std::min(e, cuts[r] + std::max(up, 0)));
        e = std::max(e, new_cuts[r - 1] + min_width);
        e = std::min(e, n_layers - min_width * (world - r));
        new_cuts[r] = e;
      }
  }

  void launch_classify(const ClassifyParams &p, cudaStream_t s)
  {
    if (p.n)
      {
        k_classify<<<blocks_for(p.n, 256), 256, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_append_records(const MigrateRecord *rec, const uint32_t *ids, uint32_t n, StateView st, uint32_t *id_out,
                             int32_t *cell_reg, double *disp, uint32_t base, cudaStream_t s)
  {
    if (n)
      {
        k_append_records<<<blocks_for(n, 256), 256, 0, s>>>(rec, ids, n, st, id_out, cell_reg, disp, base);
        count_launch();
      }
  }
  void launch_flag_layer(const int32_t *cell_reg, GridDesc grid, int layer_cell, uint32_t n, uint32_t *flags, cudaStream_t s)
  {
    k_flag_layer<<<blocks_for(size_t(n) + 1, 256), 256, 0, s>>>(cell_reg, grid, layer_cell, n, flags);
    count_launch();
  }
  namespace
  {
    __global__ void __launch_bounds__(256) k_halo_warp_table(const uint32_t *flags, const uint32_t *offsets, uint32_t n,
                                                             uint32_t *bits, uint32_t *prefix)
    {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      const uint32_t b = __ballot_sync(0xffffffffu, p < n && flags[p] != 0u);
      if ((threadIdx.x & 31u) == 0u && p < n)
        {
          bits[p >> 5] = b;
          prefix[p >> 5] = offsets[p];
        }
    }
    __global__ void k_prepare_flag(uint32_t *w, uint32_t host_bits, int consult) { w[2] = (consult ? w[0] : 0u) | host_bits; }

    __global__ void __launch_bounds__(32) k_agree(uint64_t *const *peer_mailbox, uint64_t *my_mailbox, int rank, int world, uint32_t seq,
                                                  uint32_t *flag_words, uint32_t host_bits, int consult, uint32_t *host_out)
    {
      const int lane = threadIdx.x;
      const uint32_t word = (consult ? flag_words[0] : 0u) | host_bits;
      const int parity = int(seq & 1u);
      uint32_t got = 0;
      if (lane < world)
        {
          // the stores of the kernels before this one (halo pushes into the peers) are ordered
          // before this release; the peer's acquire below makes them visible to its next kernel
          uint64_t *dst = peer_mailbox[lane] + size_t(parity) * world + rank;
          const uint64_t v = (uint64_t(seq) << 32) | word;
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
          const uint64_t *src = my_mailbox + size_t(parity) * world + lane;
          uint64_t r;
          const long long t0 = clock64();
          for (;;)
            {
              asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(r) : "l"(src) : "memory");
              if (uint32_t(r >> 32) == seq)
                {
                  got = uint32_t(r);
                  break;
                }
              if (clock64() - t0 > 240000000000ll)
                {
                  got = 0xffffffffu;
                  break;
                }
            }
        }
      for (int o = 16; o; o >>= 1)
        got = max(got, __shfl_xor_sync(0xffffffffu, got, o));
      if (lane == 0)
        {
          flag_words[1] = got;
          *reinterpret_cast<volatile uint32_t *>(host_out) = got;
        }
    }
  } // namespace
  void launch_agree(uint64_t *const *peer_mailbox, uint64_t *my_mailbox, int rank, int world, uint32_t seq, uint32_t *flag_words,
                    uint32_t host_bits, int consult, uint32_t *host_out, cudaStream_t s)
  {
    k_agree<<<1, 32, 0, s>>>(peer_mailbox, my_mailbox, rank, world, seq, flag_words, host_bits, consult, host_out);
    count_launch();
  }
  void launch_halo_warp_table(const uint32_t *flags, const uint32_t *offsets, uint32_t n, uint32_t *bits, uint32_t *prefix,
                              cudaStream_t s)
  {
    if (n)
      {
        k_halo_warp_table<<<blocks_for(n, 256), 256, 0, s>>>(flags, offsets, n, bits, prefix);
        count_launch();
      }
  }
  void launch_prepare_flag(uint32_t *flag_words, uint32_t host_bits, int consult, cudaStream_t s)
  {
    k_prepare_flag<<<1, 1, 0, s>>>(flag_words, host_bits, consult);
    count_launch();
  }
  void launch_compact_indices(const uint32_t *flags, const uint32_t *offsets, uint32_t n, uint32_t *out, cudaStream_t s)
  {
    if (n)
      {
        k_compact_indices<<<blocks_for(n, 256), 256, 0, s>>>(flags, offsets, n, out);
        count_launch();
      }
  }
  void launch_gather_state(StateView st, const uint32_t *idx, uint32_t n, double4 *pos, double4 *vel, double4 *omg,
                           cudaStream_t s)
  {
    if (n)
      {
        k_gather_state<<<blocks_for(n, 256), 256, 0, s>>>(st, idx, n, pos, vel, omg);
        count_launch();
      }
  }
  void launch_gather_ids(const uint32_t *id, const uint32_t *idx, uint32_t n, uint32_t *out, cudaStream_t s)
  {
    if (n)
      {
        k_gather_ids<<<blocks_for(n, 256), 256, 0, s>>>(id, idx, n, out);
        count_launch();
      }
  }
  void launch_ghost_run(const GhostRunParams &p, cudaStream_t s)
  {
    if (p.n)
      {
        k_ghost_run<<<blocks_for(p.n, 256), 256, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_asc_plane(const AscPlaneParams &p, int op, cudaStream_t s)
  {
    const int a = p.grid.slab_axis;
    const int u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
    const int extra = op < 2 ? 1 : 0;
    const unsigned n = unsigned((p.grid.n[u] + extra) * (p.grid.n[v] + extra));
    const unsigned blocks = (n + 255) / 256;
    if (op == 0)
      k_asc_plane<0><<<blocks, 256, 0, s>>>(p);
    else if (op == 1)
      k_asc_plane<1><<<blocks, 256, 0, s>>>(p);
    else if (op == 2)
      k_asc_plane<2><<<blocks, 256, 0, s>>>(p);
    else
      k_asc_plane<3><<<blocks, 256, 0, s>>>(p);
    count_launch(1);
  }

  void launch_layer_histogram_weighted(const double4 *pos, const int32_t *cell_reg, const uint8_t *cell_status, GridDesc grid, uint32_t n,
                                       uint32_t w_mobile, uint32_t w_active, uint32_t w_inactive, uint32_t *hist, cudaStream_t s)
  {
    if (!n)
      return;
    k_layer_histogram_weighted<<<(n + 255) / 256, 256, 0, s>>>(pos, cell_reg, cell_status, grid, n, w_mobile, w_active, w_inactive, hist);
    count_launch(1);
  }

  void launch_asc_pass(const AscParams &p, int pass, cudaStream_t s)
  {
    const unsigned cells = unsigned((p.grid.n_cells + 127) / 128);
    if (pass == 0)
      k_asc_cells<0><<<cells, 128, 0, s>>>(p);
    else if (pass == 1)
      k_asc_cells<1><<<cells, 128, 0, s>>>(p);
    else if (pass == 2)
      k_asc_cells<2><<<cells, 128, 0, s>>>(p);
    else if (pass == 3)
      k_asc_cells<3><<<cells, 128, 0, s>>>(p);
    else if (p.n_rows)
      k_asc_rows<<<(p.n_rows + 255) / 256, 256, 0, s>>>(p);
    count_launch(1);
  }

  void launch_register_ids(const uint32_t *id, uint32_t base, uint32_t n, uint32_t *slot_of_id, uint32_t map_size, cudaStream_t s)
  {
    if (n)
      {
        k_register_ids<<<blocks_for(n, 256), 256, 0, s>>>(id, base, n, slot_of_id, map_size);
        count_launch();
      }
  }
  void launch_fill_u32(uint32_t *p, uint32_t v, size_t n, cudaStream_t s)
  {
    if (n)
      {
        k_fill_u32<<<unsigned(std::min<size_t>((n + 255) / 256, 148 * 16)), 256, 0, s>>>(p, v, n);
        count_launch();
      }
  }
} // namespace dem
