// dem_step.cu — the fused DEM step kernel: particle-particle forces over the contact
// list (with in-place tangential-history update), particle-wall forces, velocity-Verlet
// integration and the displacement trigger, ONE launch per time step.
//
// Work decomposition (warp-cooperative):
//   * a warp owns 32 consecutive particles (= 32 consecutive rows of the FULL contact list,
//     a contiguous range [E0,E1) of list entries);
//   * phase A, entry-parallel: the 32 lanes sweep [E0,E1) SWEEP x 32 entries at a time with
//     coalesced loads of col[]/rowl[] (software-pipelined one iteration ahead) and SWEEP
//     position gathers in flight per lane, test "in contact?" with a sqrt-free two-sided bound
//     (exact fall-back in the 1e-12 band around the threshold) and append the touching
//     entries, by ballot + prefix popcount, to a queue in shared memory;
//   * phase B, once the range is swept: rounds of 32 queued pairs, one per lane, all lanes
//     converged on the full contact model (no lane idles on a non-touching entry); the
//     per-pair force/torque lands in shared memory and every two rounds the owner lanes add
//     their pairs up in list order — a per-particle segment reduction, no atomics,
//     deterministic, the same sequence of additions as a serial walk of the row;
//   * walls, integrator and the displacement trigger run per owner lane afterwards.
// Measured dead ends (profiles/): prefetching the round operands at the start
// of the previous round, cp.async staging of the next or of the current round (DEM_STAGE), a deeper sweep
// pipeline (SWEEP 8; with 128-bit accesses 4 was the optimum, with 256-bit accesses it is 3), 3 / 5 / 6 resident
// blocks per SM, a 256-slot queue, evict-first hints on the list streams (DEM_STREAM), reading
// neighbours that belong to the warp's own 32 rows from shared memory instead of through L1
// (+6 %: divergence costs more than the gathers), an L2 persisting access-policy window on the
// position array being gathered (+10 %) — the kernel is bound by the issue latency
// of dependent FP64 chains at 16 warps/SM. What did pay: prefetch.global.L2 of the round
// operands as soon as the sweep finds a touching entry (-2 %).
// Neighbours are read from state generation g, results go to g^1.
//
// Reference path replaced (one iteration of source/dem/dem.cc:1134-1183):
//   calculate_particle_particle_contact / execute_contact_calculation
//     (particle_particle_contact_force.cc:38-100, …force.h:1838-2063)
//   calculate_particle_wall_contact (particle_wall_contact_force.cc:45-142, .h:166-264)
//   VelocityVerletIntegrator::integrate{,_start,_end} (velocity_verlet_integrator.cc:14-115,214-290)
//   displacement accumulation of find_particle_contact_detection_step
//     (find_contact_detection_step.cc:29-47) for the NEXT step's check.
#include <cstdlib>

#include "dem_kernels.cuh"

namespace dem
{
  namespace
  {
#ifndef DEM_STEP_WARPS
#define DEM_STEP_WARPS 4
#endif
    constexpr int STEP_WARPS = DEM_STEP_WARPS; // warps per block
#ifndef DEM_QUEUE
#define DEM_QUEUE 512
#endif
#ifndef DEM_PREFETCH
#define DEM_PREFETCH 1 // 0: off, 1: prefetch.global.L2, 2: prefetch.global.L1 of the round operands at sweep time (1 M drum: 0.326 -> 0.319 ms)
#endif
    constexpr int QUEUE = DEM_QUEUE; // touching entries a warp can queue before it has to drain (>= 32 * SWEEP)
    constexpr int RES_SLOTS = 64;  // evaluated pairs buffered per warp before the owners add them up (2 rounds)
#ifndef DEM_SWEEP
#define DEM_SWEEP 3 // with 256-bit row accesses three gathers in flight per lane beat four (fewer spills at the 128-register cap): 1 M drum 0.316 / 0.278 / 0.2715 / 0.291 ms for 1 / 2 / 3 / 4
#endif
#ifndef DEM_MIN_BLOCKS
#define DEM_MIN_BLOCKS 4
#endif
    constexpr int SWEEP = DEM_SWEEP;      // 32-entry blocks of the list swept per phase-A iteration (loads in flight)
    constexpr int STEP_MIN_BLOCKS = DEM_MIN_BLOCKS; // resident blocks per SM the register allocation is held to
#ifndef DEM_MIN_BLOCKS_MIXED
#define DEM_MIN_BLOCKS_MIXED 4
#endif
    constexpr int STEP_MIN_BLOCKS_MIXED = DEM_MIN_BLOCKS_MIXED; // the same for the mixed-precision instantiations

    __device__ __forceinline__ ParticleView make_view(double4 p, double4 v, double4 w)
    {
      ParticleView r;
      r.x = v3(p.x, p.y, p.z);
      r.d = p.w;
      r.v = v3(v.x, v.y, v.z);
      r.m = v.w;
      r.w = v3(w.x, w.y, w.z);
      r.type = static_cast<int>(static_cast<unsigned int>(w.w));
      return r;
    }


    // 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256). A double4 row is one 32-byte
    // sector; CUDA 12's double4 is only 16-byte aligned for the compiler, which therefore emits two
    // 128-bit requests per row — twice the L1 wavefronts for a gather that lands one line per lane.
    // All rows here are 32-byte aligned (cudaMalloc base + 32 * index).
#ifndef DEM_LD256
#define DEM_LD256 1
#endif
    __device__ __forceinline__ double4 ld_row(const double4 *p)
    {
#if DEM_LD256
      double4 r;
      asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
      return r;
#else
      return *p;
#endif
    }
    // the same for a row this thread may have written earlier in the launch (history): ordered with its stores
    __device__ __forceinline__ double4 ld_row_rw(const double4 *p)
    {
#if DEM_LD256
      double4 r;
#if DEM_STREAM
      asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
#else
      asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
#endif
      return r;
#else
      return *p;
#endif
    }
#ifndef DEM_STREAM
#define DEM_STREAM 0 // 1: evict-first loads / stores for the list streams (col, rowl, img, history rows)
#endif
#if DEM_STREAM
#define DEM_LDS(p) __ldcs(p)
#else
#define DEM_LDS(p) (*(p))
#endif
#ifndef DEM_STAGE
#define DEM_STAGE 0 // 1: the operands of a round are gathered by cp.async into shared memory (no registers held across the latency)
#endif
    // 32-byte row -> two conflict-free 16-byte halves in shared memory, asynchronously
    __device__ __forceinline__ void cp_async_row(double2 *lo, double2 *hi, const void *g)
    {
      const uint32_t a = uint32_t(__cvta_generic_to_shared(lo)), b = uint32_t(__cvta_generic_to_shared(hi));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(b), "l"(reinterpret_cast<const char *>(g) + 16) : "memory");
    }
    // history rows are touched once per step: streaming (evict-first) accesses keep them from displacing the gathered state
    __device__ __forceinline__ void st_row_stream(double4 *p, const double4 &r)
    {
#if DEM_STREAM && DEM_LD256
      asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.x), "d"(r.y), "d"(r.z), "d"(r.w) : "memory");
#elif DEM_LD256
      asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.x), "d"(r.y), "d"(r.z), "d"(r.w) : "memory");
#else
      *p = r;
#endif
    }
    __device__ __forceinline__ void st_row(double4 *p, const double4 &r)
    {
#if DEM_LD256
      asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.x), "d"(r.y), "d"(r.z), "d"(r.w) : "memory");
#else
      *p = r;
#endif
    }

    // img = 1 + (sx+1) + 3*(sy+1) + 9*(sz+1) for a neighbour seen through the periodic
    // image shifted by (sx*Lx, sy*Ly, sz*Lz); 0 = no image.
    __device__ __forceinline__ void decode_image(uint32_t img, const double *L, vec3 &shift, bool &i_am_two)
    {
      const int c = int(img) - 1;
      const int sx = c % 3 - 1, sy = (c / 3) % 3 - 1, sz = c / 9 - 1;
      shift = v3(sx * L[0], sy * L[1], sz * L[2]);
      // canonical orientation of a periodic pair = the reference's: particle one sits in the
      // cell on periodic boundary 0 (low side), particle two is translated by -L
      // (find_cell_neighbors.cc:147-170, particle_particle_fine_search.cc:193-228).
      const int first = sx != 0 ? sx : (sy != 0 ? sy : sz);
      i_am_two = first > 0;
    }

    // Canonical positions of particle one / two of the pair (row particle, neighbour).
    template <bool PERIODIC>
    __device__ __forceinline__ void pair_positions(const StepParams &P, uint32_t img, const double4 &pme, const double4 &pj, vec3 &x1,
                                                   vec3 &x2, double &dsum, bool &i_am_two)
    {
      i_am_two = false;
      if constexpr (PERIODIC)
        {
          if (img)
            {
              vec3 shift;
              decode_image(img, P.L, shift, i_am_two);
              if (!i_am_two)
                {
                  x1 = v3(pme.x, pme.y, pme.z);
                  x2 = v3(pj.x, pj.y, pj.z) + shift;
                  dsum = pme.w + pj.w;
                }
              else
                {
                  x1 = v3(pj.x, pj.y, pj.z);
                  x2 = v3(pme.x, pme.y, pme.z) + (-shift);
                  dsum = pj.w + pme.w;
                }
              return;
            }
        }
      x1 = v3(pme.x, pme.y, pme.z);
      x2 = v3(pj.x, pj.y, pj.z);
      dsum = pme.w + pj.w;
    }

    struct __align__(16) WarpScratch
    {
      double4 pos[32], vel[32], omg[32]; // state of the warp's own 32 particles
      double2 res[RES_SLOTS][3];         // per evaluated pair: force (3) and torque (3) on the row particle
      uint32_t q_e[QUEUE];               // queued touching entries: list position,
      uint32_t q_c[QUEUE];               //   col word (neighbour index | history bit),
      uint8_t q_owner[QUEUE];            //   row (lane) that owns it
      uint8_t res_owner[RES_SLOTS];      // row that owns res[k] (non-decreasing in k)
      uint8_t seg[32];                   // per owner: first buffer position of its run at the current flush
      double2 self[32];                  // per owner: R* and m* of a pair of two copies of itself (SelfPair)
      double disp[32];                   // per owner: accumulated displacement, loaded with the state
      uint32_t w0[32], w1[32];           // per owner: its range of the wall list
      uint32_t halo[4];                  // HaloPush bits / prefix of this 32-row block, per direction
#if DEM_STAGE
      double2 stg[4][2][32];             // round operands (neighbour pos / vel / omg, history row) landed by cp.async, lane-private
#endif
    };

    template <int MODEL, int ROLLING, bool PERIODIC, bool MIXED>
    __global__ void __launch_bounds__(32 * STEP_WARPS, MIXED ? STEP_MIN_BLOCKS_MIXED : STEP_MIN_BLOCKS) k_step(const __grid_constant__ StepParams P, const __grid_constant__ MaterialTables mt)
    {
      extern __shared__ __align__(16) unsigned char scratch_raw[]; // STEP_WARPS x WarpScratch (may exceed the 48 KB static limit)
      const uint32_t lane = threadIdx.x & 31u;
      WarpScratch &S = reinterpret_cast<WarpScratch *>(scratch_raw)[threadIdx.x >> 5];
      const uint32_t block = P.block_list ? P.block_list[blockIdx.x] : blockIdx.x; // partial launch: the listed blocks only
      const uint32_t warp_base = (block * STEP_WARPS + (threadIdx.x >> 5)) * 32u;
      if (warp_base >= P.n_owned)
        return; // whole warp
      // speculative launch: the flag of the steps before this one, read with everything else
      // the prologue loads so that its latency is not exposed on its own
      const uint32_t flag_seen = P.spec_check ? *reinterpret_cast<const volatile uint32_t *>(P.flag_check) : 0u;
      const uint32_t n_rows = min(32u, P.n_owned - warp_base);
      const uint32_t i = warp_base + lane;
      const bool valid = lane < n_rows;
      const double dt = P.dt;

      // ---------------- stage the warp's own rows ----------------
      // every per-owner operand of the whole kernel is requested here, in one batch of
      // independent loads: state, the displacement accumulator and the wall-list range
      const uint32_t E0 = P.list.row_start[warp_base], E1 = P.list.row_start[warp_base + n_rows];
      double4 own_pos = make_double4(0, 0, 0, 1), own_vel = make_double4(0, 0, 0, 1), own_omg = make_double4(0, 0, 0, 0);
      double own_disp = 0.;
      uint32_t own_w0 = 0, own_w1 = 0;
      if (valid)
        {
          own_pos = ld_row(P.in.pos + i);
          own_vel = ld_row(P.in.vel + i);
          own_omg = ld_row(P.in.omg + i);
          own_disp = P.disp[i];
          own_w0 = P.walls.row_start[i];
          own_w1 = P.walls.row_start[i + 1];
        }
      if (flag_seen != 0u && flag_seen != P.flag_tag)
        return; // an earlier step asked for a new list: this launch is void (whole grid)
      S.pos[lane] = own_pos;
      S.vel[lane] = own_vel;
      S.omg[lane] = own_omg;
      S.disp[lane] = own_disp;
      S.w0[lane] = own_w0;
      S.w1[lane] = own_w1;
      if (lane < 4)
        {
          const uint32_t d = lane & 1u;
          S.halo[lane] = P.halo.pos[d] ? (lane < 2 ? P.halo.bits[d][warp_base >> 5] : P.halo.prefix[d][warp_base >> 5]) : 0u;
        }
      // R* and m* of a pair whose two particles have my diameter and mass (bit-identical to what
      // pp_calculate_contact computes for such a pair, which is every pair of a monodisperse type)
      S.self[lane] = make_double2((own_pos.w * own_pos.w) / (2 * (own_pos.w + own_pos.w)), (own_vel.w * own_vel.w) / (own_vel.w + own_vel.w));
      __syncwarp();

      vec3 F = v3(0, 0, 0), T = v3(0, 0, 0);
      unsigned int touching_count = 0;
      uint32_t q_n = 0, n_res = 0; // warp-uniform

      // phase B: evaluate the queued pairs [first, first + cnt), cnt <= 32, one per lane
      auto process_round = [&](uint32_t first, uint32_t cnt) {
        if (lane < cnt)
          {
            const uint32_t slot = first + lane;
            const uint32_t e = S.q_e[slot];
            const uint32_t owner = S.q_owner[slot];
            const uint32_t c = S.q_c[slot];
            const uint32_t j = c & COL_INDEX_MASK;
            const double4 pme = S.pos[owner];
            // every operand of the pair is requested here, before the first use of any of them
            double4 *hp = P.list.hist + size_t(e);
#if DEM_STAGE
            cp_async_row(&S.stg[0][0][lane], &S.stg[0][1][lane], P.in.pos + j);
            cp_async_row(&S.stg[1][0][lane], &S.stg[1][1][lane], P.in.vel + j);
            cp_async_row(&S.stg[2][0][lane], &S.stg[2][1][lane], P.in.omg + j);
            if (c & COL_HIST_BIT)
              cp_async_row(&S.stg[3][0][lane], &S.stg[3][1][lane], hp);
            uint32_t img = 0;
            if constexpr (PERIODIC)
              img = P.list.img[e];
            asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
            const double2 pj0 = S.stg[0][0][lane], pj1 = S.stg[0][1][lane], vj0 = S.stg[1][0][lane], vj1 = S.stg[1][1][lane],
                          wj0 = S.stg[2][0][lane], wj1 = S.stg[2][1][lane];
            const double4 pj = make_double4(pj0.x, pj0.y, pj1.x, pj1.y), vj = make_double4(vj0.x, vj0.y, vj1.x, vj1.y),
                          wj = make_double4(wj0.x, wj0.y, wj1.x, wj1.y);
#else
            const double4 pj = ld_row(P.in.pos + j);
            const double4 vj = ld_row(P.in.vel + j);
            const double4 wj = ld_row(P.in.omg + j);
            uint32_t img = 0;
            if constexpr (PERIODIC)
              img = P.list.img[e];
#endif
            vec3 h = v3(0, 0, 0), rs = v3(0, 0, 0);
            if (c & COL_HIST_BIT)
              {
#if DEM_STAGE
                const double2 h0 = S.stg[3][0][lane], h1 = S.stg[3][1][lane];
                const double4 h4 = make_double4(h0.x, h0.y, h1.x, h1.y);
#else
                const double4 h4 = ld_row_rw(hp);
#endif
                h = v3(h4.x, h4.y, h4.z);
                if constexpr (ROLLING == LETHE_ROLLING_EPSD)
                  {
                    const double *rp = P.list.roll + 3 * size_t(e);
                    rs = v3(rp[0], rp[1], rp[2]);
                  }
              }
            const ParticleView me = make_view(pme, S.vel[owner], S.omg[owner]);
            const ParticleView other = make_view(pj, vj, wj);
            vec3 x1, x2;
            double dsum;
            bool i_am_two;
            pair_positions<PERIODIC>(P, img, pme, pj, x1, x2, dsum, i_am_two);
            const double distance = sqrt(dist2(x1, x2));
            const double normal_overlap = 0.5 * dsum - distance;
            const double2 self = S.self[owner];
            vec3 fc, tc; // what the row particle receives: F -= fc, T += tc
            if constexpr (!MIXED)
              {
                PairResult r;
                r.normal_force = r.tangential_force = r.torque_one = r.torque_two = r.rolling = v3(0, 0, 0);
                const SelfPair sp{self.x, self.y};
                vec3 n, vt;
                double vn;
                if (PERIODIC && i_am_two)
                  {
                    // evaluate the pair in its canonical orientation (one = neighbour) so that both
                    // owners of the pair run bit-identical arithmetic; my copy of the history is the
                    // negative of the canonical one.
                    ParticleView one = other, two = me;
                    one.x = x1;
                    h = -h;
                    rs = -rs;
                    pp_update_contact_information<double>(h, vt, vn, n, one, two, x2, distance, dt);
                    pp_calculate_contact<MODEL, ROLLING, double>(mt, h, rs, vt, vn, n, normal_overlap, dt, one, two, r, sp);
                    // apply_force_and_torque_on_local_particles, particle two (…force.h:565-569)
                    fc = -(r.normal_force + r.tangential_force);
                    tc = -r.torque_two - r.rolling;
                    h = -h;
                    rs = -rs;
                  }
                else
                  {
                    ParticleView one = me;
                    one.x = x1;
                    pp_update_contact_information<double>(h, vt, vn, n, one, other, x2, distance, dt);
                    pp_calculate_contact<MODEL, ROLLING, double>(mt, h, rs, vt, vn, n, normal_overlap, dt, one, other, r, sp);
                    // particle one (…force.h:561-567)
                    fc = r.normal_force + r.tangential_force;
                    tc = -r.torque_one + r.rolling;
                  }
              }
            else
              {
                // Mixed precision (config.precision = LETHE_PRECISION_MIXED): what cancels is done in
                // double — the centre distance and the overlap above, the contact vector and the
                // relative translational velocity below — and handed to the model as floats; the
                // model itself (stiffnesses, damping, Coulomb limit, rolling resistance, history
                // update) runs in float; the force and torque go back to double for the segment
                // reduction and the integration. Both owners of a pair still run the same
                // arithmetic on the same operands (canonical orientation), so action = reaction
                // holds to the last bit here too.
                PairResult_t<float> r;
                r.normal_force = r.tangential_force = r.torque_one = r.torque_two = r.rolling = v3t<float>(0, 0, 0);
                const SelfPair_t<float> sp{float(self.x), float(self.y)};
                const bool flip = PERIODIC && i_am_two;
                const ParticleView &a = flip ? other : me; // particle one of the canonical orientation
                const ParticleView &b = flip ? me : other;
                ParticleView_t<float> one, two;
                one.x = v3t<float>(0, 0, 0);
                one.d = float(a.d);
                one.m = float(a.m);
                one.v = to_float(a.v - b.v);
                one.w = to_float(a.w);
                one.type = a.type;
                two.x = to_float(x2 - x1);
                two.d = float(b.d);
                two.m = float(b.m);
                two.v = v3t<float>(0, 0, 0);
                two.w = to_float(b.w);
                two.type = b.type;
                vec3f hf = to_float(flip ? -h : h), rf = to_float(flip ? -rs : rs);
                vec3f n, vt;
                float vn;
                pp_update_contact_information<float>(hf, vt, vn, n, one, two, two.x, float(distance), float(dt));
                pp_calculate_contact<MODEL, ROLLING, float>(mt, hf, rf, vt, vn, n, float(normal_overlap), float(dt), one, two, r, sp);
                if (flip)
                  {
                    fc = to_double(-(r.normal_force + r.tangential_force));
                    tc = to_double(-r.torque_two - r.rolling);
                    h = to_double(-hf);
                    rs = to_double(-rf);
                  }
                else
                  {
                    fc = to_double(r.normal_force + r.tangential_force);
                    tc = to_double(-r.torque_one + r.rolling);
                    h = to_double(hf);
                    rs = to_double(rf);
                  }
              }
            st_row_stream(hp, make_double4(h.x, h.y, h.z, 0.0));
            if constexpr (ROLLING == LETHE_ROLLING_EPSD)
              {
                double *rp = P.list.roll + 3 * size_t(e);
                rp[0] = rs.x;
                rp[1] = rs.y;
                rp[2] = rs.z;
              }
            if (!(c & COL_HIST_BIT))
              P.list.col[e] = c | COL_HIST_BIT;
            S.res[n_res + lane][0] = make_double2(fc.x, fc.y);
            S.res[n_res + lane][1] = make_double2(fc.z, tc.x);
            S.res[n_res + lane][2] = make_double2(tc.y, tc.z);
            S.res_owner[n_res + lane] = uint8_t(owner);
          }
        n_res += cnt;
        __syncwarp();
      };

      // Owner lanes add up the buffered pair results. Queue order = list order, so res_owner is
      // non-decreasing and every owner's pairs are one run of the buffer, added in list order:
      // the same sequence of additions as a serial walk of the row.
      auto flush_results = [&]() {
        static_assert(RES_SLOTS == 64, "the run search below looks at two buffer positions per lane");
        // lanes inspect buffer positions lane and lane + 32; the first lane of every run (head)
        // publishes the run's start for its owner, the ballots give every run's end
        const uint32_t o0 = lane < n_res ? S.res_owner[lane] : 0xffu;
        const uint32_t o1 = lane + 32 < n_res ? S.res_owner[lane + 32] : 0xffu;
        const uint32_t p0 = __shfl_up_sync(0xffffffffu, o0, 1);
        const uint32_t p1 = __shfl_up_sync(0xffffffffu, o1, 1);
        const uint32_t last0 = __shfl_sync(0xffffffffu, o0, 31);
        const bool head0 = lane < n_res && (lane == 0 || p0 != o0);
        const bool head1 = lane + 32 < n_res && ((lane == 0 ? last0 : p1) != o1);
        const uint64_t heads = (uint64_t(__ballot_sync(0xffffffffu, head1)) << 32) | __ballot_sync(0xffffffffu, head0);
        const uint32_t present =
          __reduce_or_sync(0xffffffffu, (lane < n_res ? (1u << o0) : 0u) | (lane + 32 < n_res ? (1u << o1) : 0u));
        if (head0)
          S.seg[o0] = uint8_t(lane);
        if (head1)
          S.seg[o1] = uint8_t(32 + lane);
        __syncwarp();
        if ((present >> lane) & 1u)
          {
            const uint32_t s0 = S.seg[lane];
            const uint64_t above = s0 >= 63 ? 0ull : (heads & (~0ull << (s0 + 1)));
            const uint32_t end = above ? uint32_t(__ffsll((long long)above) - 1) : n_res;
            for (uint32_t k = s0; k < end; ++k)
              {
                const double2 a = S.res[k][0], b = S.res[k][1], cc = S.res[k][2];
                F = F - v3(a.x, a.y, b.x);
                T = T + v3(b.y, cc.x, cc.y);
              }
            touching_count += end - s0;
          }
        n_res = 0;
        __syncwarp();
      };

      // The warp alternates between (A) sweeping its list entries, which only queues the
      // touching ones, until the range is exhausted (or, for very dense rows, the queue is
      // full), and (B) draining the queue in rounds of 32 pairs. Keeping the two apart keeps
      // the sweep a tight streaming loop with many loads in flight and the (large) pair
      // evaluation a single instance in the instruction stream.
      uint32_t eb = E0;
      while (eb < E1)
        {
          // ---------------- phase A ----------------
          // rowl and (periodic) img travel with col, one sweep iteration ahead: ow_nx = rowl | img << 8
          uint32_t c_nx[SWEEP], ow_nx[SWEEP];
#pragma unroll
          for (int u = 0; u < SWEEP; ++u)
            {
              const uint32_t e = eb + 32u * u + lane;
              c_nx[u] = e < E1 ? DEM_LDS(P.list.col + e) : 0u;
              ow_nx[u] = e < E1 ? DEM_LDS(P.list.rowl + e) : 0u;
              if constexpr (PERIODIC)
                ow_nx[u] |= e < E1 ? uint32_t(DEM_LDS(P.list.img + e)) << 8 : 0u;
            }
          while (eb < E1 && q_n + 32u * SWEEP <= QUEUE)
            {
              uint32_t c[SWEEP], ow[SWEEP];
              double4 pj[SWEEP];
#pragma unroll
              for (int u = 0; u < SWEEP; ++u)
                {
                  c[u] = c_nx[u];
                  ow[u] = ow_nx[u];
                  const uint32_t e = eb + 32u * u + lane;
                  if (e < E1)
                    pj[u] = ld_row(P.in.pos + (c[u] & COL_INDEX_MASK));
                }
#pragma unroll
              for (int u = 0; u < SWEEP; ++u)
                {
                  const uint32_t e = eb + 32u * (SWEEP + u) + lane;
                  c_nx[u] = e < E1 ? DEM_LDS(P.list.col + e) : 0u;
                  ow_nx[u] = e < E1 ? DEM_LDS(P.list.rowl + e) : 0u;
                  if constexpr (PERIODIC)
                    ow_nx[u] |= e < E1 ? uint32_t(DEM_LDS(P.list.img + e)) << 8 : 0u;
                }
#pragma unroll
              for (int u = 0; u < SWEEP; ++u)
                {
                  const uint32_t e = eb + 32u * u + lane;
                  bool touching = false;
                  if (e < E1)
                    {
                      const double4 pme = S.pos[ow[u] & 31u];
                      vec3 x1, x2;
                      double dsum;
                      bool i_am_two;
                      pair_positions<PERIODIC>(P, ow[u] >> 8, pme, pj[u], x1, x2, dsum, i_am_two);
                      const double d2 = dist2(x1, x2);
                      // in contact  <=>  0.5*dsum - sqrt(d2) > threshold (…force.h:1868-1876). Decide
                      // without the square root when d2 is clearly on one side of (0.5*dsum - thr)^2.
                      const double hgap = 0.5 * dsum - mt.pp_force_threshold;
                      const double hh = hgap * hgap;
                      if (d2 < hh * (1.0 - 1e-12))
                        touching = true;
                      else if (d2 > hh * (1.0 + 1e-12))
                        touching = false;
                      else
                        touching = (0.5 * dsum - sqrt(d2)) > mt.pp_force_threshold;
#ifdef DEM_EXPERIMENT_HALF
                      // timing experiment only (wrong physics): the cost of evaluating each pair once
                      if ((c[u] & COL_INDEX_MASK) < warp_base + (ow[u] & 31u))
                        touching = false;
#endif
                      // contact_info.tangential_displacement.clear() (…force.h:2057-2063): dropping
                      // the flag is the clear; the 24 B are not touched.
                      if (!touching && (c[u] & COL_HIST_BIT))
                        P.list.col[e] = c[u] & COL_INDEX_MASK;
                    }
                  const uint32_t m = __ballot_sync(0xffffffffu, touching);
                  if (touching)
                    {
#if DEM_PREFETCH
                      {
                        // the pair will be evaluated a few hundred cycles from now: start moving its
                        // operands (neighbour velocity / angular velocity, history row) towards the SM
                        const uint32_t jj = c[u] & COL_INDEX_MASK;
                        const double4 *hq = P.list.hist + size_t(e);
#if DEM_PREFETCH == 1
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.in.vel + jj));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.in.omg + jj));
                        if (c[u] & COL_HIST_BIT)
                          asm volatile("prefetch.global.L2 [%0];" ::"l"(hq));
#else
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(P.in.vel + jj));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(P.in.omg + jj));
                        if (c[u] & COL_HIST_BIT)
                          asm volatile("prefetch.global.L1 [%0];" ::"l"(hq));
#endif
                      }
#endif
                      const uint32_t slot = q_n + __popc(m & ((1u << lane) - 1u));
                      S.q_e[slot] = e;
                      S.q_c[slot] = c[u];
                      S.q_owner[slot] = uint8_t(ow[u] & 31u);
                    }
                  q_n += __popc(m);
                }
              eb += 32u * SWEEP;
            }
          __syncwarp();
          // ---------------- phase B ----------------
          for (uint32_t first = 0; first < q_n; first += 32)
            {
              process_round(first, min(32u, q_n - first));
              if (n_res > RES_SLOTS - 32)
                flush_results();
            }
          if (n_res)
            flush_results();
          q_n = 0;
        }

      if (!valid)
        return;
      const double4 pi = S.pos[lane], vi = S.vel[lane], wi = S.omg[lane];
      const ParticleView me = make_view(pi, vi, wi);

      // ---------------- particle-wall contacts ----------------
#ifdef DEM_EXPERIMENT_NOWALLS
      const uint32_t w0 = 0, w1 = 0; // register-pressure experiment only (wrong physics)
#else
      const uint32_t w0 = S.w0[lane], w1 = S.w1[lane];
#endif
      for (uint32_t w = w0; w < w1; ++w)
        {
          const uint32_t we = P.walls.entry[w];
          const uint32_t idx = we & WALL_INDEX_MASK;
          vec3 wall_normal, point;
          int motion = -1;
          if (we & WALL_FLOATING_BIT)
            {
              wall_normal = v3(P.floating->normal[idx][0], P.floating->normal[idx][1], P.floating->normal[idx][2]);
              if (we & WALL_FLIPPED_BIT)
                wall_normal = -1 * wall_normal;
              point = v3(P.floating->point[idx][0], P.floating->point[idx][1], P.floating->point[idx][2]);
            }
          else
            {
              const double *np = P.faces.normal + 3 * size_t(idx), *pp = P.faces.point + 3 * size_t(idx);
              wall_normal = v3(np[0], np[1], np[2]);
              point = v3(pp[0], pp[1], pp[2]);
              motion = P.faces.motion[idx];
            }
          // particle_wall_contact_force.cc:72-88
          const vec3 point_to_particle_vector = me.x - point;
          const vec3 projected_vector = ((dot(point_to_particle_vector, wall_normal)) / (norm2(wall_normal))) * wall_normal;
          const double normal_overlap = ((me.d) * 0.5) - (norm(projected_vector));
          if (normal_overlap > mt.pw_force_threshold)
            {
              vec3 h = v3(0, 0, 0), rs = v3(0, 0, 0);
              if (we & WALL_HIST_BIT)
                {
                  const double *hp = P.walls.hist + 3 * size_t(w);
                  h = v3(hp[0], hp[1], hp[2]);
                  if (P.rolling_model == LETHE_ROLLING_EPSD)
                    {
                      const double *rp = P.walls.roll + 3 * size_t(w);
                      rs = v3(rp[0], rp[1], rp[2]);
                    }
                }
              // update_contact_information (particle_wall_contact_force.h:166-264)
              const vec3 normal_vector = -wall_normal;
              const vec3 contact_point = me.x + (0.5 * me.d) * normal_vector;
              vec3 bt = v3(0, 0, 0), br = v3(0, 0, 0), bp = v3(0, 0, 0);
              double bs = 0.;
              if (motion >= 0)
                {
                  const BoundaryMotionDev &m = P.motions[motion];
                  bt = v3(m.translational_velocity[0], m.translational_velocity[1], m.translational_velocity[2]);
                  bs = m.rotational_speed;
                  br = v3(m.rotational_vector[0], m.rotational_vector[1], m.rotational_vector[2]);
                  bp = v3(m.point_on_axis[0], m.point_on_axis[1], m.point_on_axis[2]);
                }
              vec3 vector_to_rotating_axis = contact_point - bp;
              vector_to_rotating_axis = vector_to_rotating_axis - (dot(vector_to_rotating_axis, br)) * br;
              const vec3 vrel =
                bt - me.v + cross(((-0.5 * me.d) * me.w), normal_vector) + cross(bs * br, vector_to_rotating_axis);
              const double vn = dot(vrel, normal_vector);
              const vec3 vt = vrel - (vn * normal_vector);
              h = h + vt * dt;
              WallResult r;
              r.normal_force = r.tangential_force = r.tangential_torque = r.rolling = v3(0, 0, 0);
              pw_calculate_contact(P.pw_model, P.rolling_model, mt, wall_normal, h, rs, vt, vn, normal_overlap, dt, me, r);
              // apply_force_and_torque (particle_wall_contact_force.h:506-522)
              const vec3 total_force = r.normal_force + r.tangential_force;
              F = F - total_force;
              T = T + (r.tangential_torque + r.rolling);
              double *hp = P.walls.hist + 3 * size_t(w);
              hp[0] = h.x;
              hp[1] = h.y;
              hp[2] = h.z;
              if (P.rolling_model == LETHE_ROLLING_EPSD)
                {
                  double *rp = P.walls.roll + 3 * size_t(w);
                  rp[0] = rs.x;
                  rp[1] = rs.y;
                  rp[2] = rs.z;
                }
              if (!(we & WALL_HIST_BIT))
                P.walls.entry[w] = we | WALL_HIST_BIT;
            }
          else if (we & WALL_HIST_BIT)
            P.walls.entry[w] = we & ~WALL_HIST_BIT;
        }

      // particle - solid surface contacts (calculate_particle_solid_object_contact, dem.cc:527-537),
      // evaluated by k_solid_contacts just before this kernel
      if (P.solid_force)
        {
          F = F + v3(P.solid_force[3 * size_t(i) + 0], P.solid_force[3 * size_t(i) + 1], P.solid_force[3 * size_t(i) + 2]);
          T = T + v3(P.solid_torque[3 * size_t(i) + 0], P.solid_torque[3 * size_t(i) + 1], P.solid_torque[3 * size_t(i) + 2]);
        }

      if (P.touching_counter && touching_count)
        atomicAdd(P.touching_counter, (unsigned long long)touching_count);
      if (P.force_out)
        {
          P.force_out[3 * size_t(i) + 0] = F.x;
          P.force_out[3 * size_t(i) + 1] = F.y;
          P.force_out[3 * size_t(i) + 2] = F.z;
          P.torque_out[3 * size_t(i) + 0] = T.x;
          P.torque_out[3 * size_t(i) + 1] = T.y;
          P.torque_out[3 * size_t(i) + 2] = T.z;
        }

      // ---------------- velocity-Verlet ----------------
      const double MOI = P.moi_override > 0 ? P.moi_override : 0.1 * me.m * me.d * me.d; // dem.cc:1004-1011
      vec3 v = me.v, x = me.x, om = me.w;
      const vec3 g = v3(P.g[0], P.g[1], P.g[2]);
      if (P.integrator == LETHE_INTEGRATOR_EXPLICIT_EULER)
        {
          // ExplicitEulerIntegrator::integrate (explicit_euler_integrator.cc:69-130); the opening
          // step is a regular step (:14-26), the closing one leaves the state untouched (:32-48)
          if (P.phase != PHASE_END)
            {
              const double mass_inverse = 1 / me.m;
              const double MOI_inverse = 1 / MOI;
              v.x = v.x + dt * (g.x + (F.x) * mass_inverse);
              v.y = v.y + dt * (g.y + (F.y) * mass_inverse);
              v.z = v.z + dt * (g.z + (F.z) * mass_inverse);
              x.x = x.x + dt * v.x;
              x.y = x.y + dt * v.y;
              x.z = x.z + dt * v.z;
              om.x = om.x + dt * (T.x * MOI_inverse);
              om.y = om.y + dt * (T.y * MOI_inverse);
              om.z = om.z + dt * (T.z * MOI_inverse);
            }
        }
      else if (P.phase == PHASE_REGULAR)
        {
          const vec3 dt_g = g * dt;
          const double dt_mass_inverse = dt / me.m;
          const double dt_MOI_inverse = dt / MOI;
          v.x = v.x + (dt_g.x + F.x * dt_mass_inverse);
          v.y = v.y + (dt_g.y + F.y * dt_mass_inverse);
          v.z = v.z + (dt_g.z + F.z * dt_mass_inverse);
          x.x = x.x + v.x * dt;
          x.y = x.y + v.y * dt;
          x.z = x.z + v.z * dt;
          om.x = om.x + T.x * dt_MOI_inverse;
          om.y = om.y + T.y * dt_MOI_inverse;
          om.z = om.z + T.z * dt_MOI_inverse;
        }
      else
        {
          const vec3 half_dt_g = 0.5 * g * dt;
          const double half_dt_mass_inverse = 0.5 * dt / me.m;
          const double half_dt_MOI_inverse = 0.5 * dt / MOI;
          v.x = v.x + (half_dt_g.x + F.x * half_dt_mass_inverse);
          v.y = v.y + (half_dt_g.y + F.y * half_dt_mass_inverse);
          v.z = v.z + (half_dt_g.z + F.z * half_dt_mass_inverse);
          om.x = om.x + T.x * half_dt_MOI_inverse;
          om.y = om.y + T.y * half_dt_MOI_inverse;
          om.z = om.z + T.z * half_dt_MOI_inverse;
          if (P.phase == PHASE_START)
            {
              x.x = x.x + v.x * dt;
              x.y = x.y + v.y * dt;
              x.z = x.z + v.z * dt;
            }
        }
      // adaptive sparse contacts (velocity_verlet_integrator.cc:117-210,292-436): only the particles of
      // mobile cells are integrated; the others keep their state, their force and torque are dropped
      if (P.row_mobile && !P.row_mobile[i])
        {
          v = me.v;
          x = me.x;
          om = me.w;
        }
      const double4 new_pos = make_double4(x.x, x.y, x.z, pi.w), new_vel = make_double4(v.x, v.y, v.z, vi.w),
                    new_omg = make_double4(om.x, om.y, om.z, wi.w);
      st_row(P.out.pos + i, new_pos);
      st_row(P.out.vel + i, new_vel);
      st_row(P.out.omg + i, new_omg);
      // fused halo push: boundary-layer rows also go to the neighbour GPU's ghost slots
#pragma unroll
      for (int d = 0; d < 2; ++d)
        {
          const uint32_t bits = S.halo[d];
          if ((bits >> lane) & 1u)
            {
              const uint32_t k = P.halo.base[d] + S.halo[2 + d] + __popc(bits & ((1u << lane) - 1u));
              st_row(P.halo.pos[d] + k, new_pos);
              st_row(P.halo.vel[d] + k, new_vel);
              st_row(P.halo.omg[d] + k, new_omg);
            }
        }

      // displacement for the next step's contact-detection check
      if (P.phase != PHASE_END)
        {
          const double dsp = S.disp[lane] + dt * sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
          P.disp[i] = dsp;
          if (dsp > P.criterion && *reinterpret_cast<volatile uint32_t *>(P.flag_local) != P.flag_tag)
            {
              *reinterpret_cast<volatile uint32_t *>(P.flag_local) = P.flag_tag;
              if (P.flag_host)
                *reinterpret_cast<volatile uint32_t *>(P.flag_host) = P.flag_tag;
            }
        }
    }

    template <int MODEL, int ROLLING>
    void launch_mr(const StepParams &p, const MaterialTables &mt, cudaStream_t stream)
    {
      if (p.n_owned == 0)
        return;
      constexpr uint32_t per_block = 32 * STEP_WARPS;
      static_assert(per_block == STEP_BLOCK_ROWS, "the streamed host step plans in blocks of STEP_BLOCK_ROWS rows");
      if (p.block_list && p.n_blocks_listed == 0)
        return;
      const dim3 block(per_block), grid(p.block_list ? p.n_blocks_listed : (p.n_owned + per_block - 1) / per_block);
      constexpr size_t smem = sizeof(WarpScratch) * STEP_WARPS;
      auto launch = [&](auto kernel) {
        static bool configured[64] = {}; // per instantiation (the lambda's operator() is a template) and device
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured[dev & 63])
          {
            // a failure here surfaces as a launch error, which the engine checks after every step launch
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            configured[dev & 63] = true;
          }
        kernel<<<grid, block, smem, stream>>>(p, mt);
      };
      if (p.mixed_precision)
        {
          if (p.periodic_any)
            launch(k_step<MODEL, ROLLING, true, true>);
          else
            launch(k_step<MODEL, ROLLING, false, true>);
        }
      else if (p.periodic_any)
        launch(k_step<MODEL, ROLLING, true, false>);
      else
        launch(k_step<MODEL, ROLLING, false, false>);
      count_launch();
    }

    template <int MODEL>
    void launch_m(int rolling, const StepParams &p, const MaterialTables &mt, cudaStream_t stream)
    {
      switch (rolling)
        {
          case LETHE_ROLLING_NONE:
            launch_mr<MODEL, LETHE_ROLLING_NONE>(p, mt, stream);
            break;
          case LETHE_ROLLING_CONSTANT:
            launch_mr<MODEL, LETHE_ROLLING_CONSTANT>(p, mt, stream);
            break;
          case LETHE_ROLLING_VISCOUS:
            launch_mr<MODEL, LETHE_ROLLING_VISCOUS>(p, mt, stream);
            break;
          default:
            launch_mr<MODEL, LETHE_ROLLING_EPSD>(p, mt, stream);
            break;
        }
    }
  } // namespace

  // Runtime -> template dispatch, the same pattern as set_particle_particle_contact_force_model
  // / set_rolling_resistance_model (set_particle_particle_contact_force_model.cc:12-103).
  // ------------------------------------------------------------ DEM-MP heat transfer ----
  namespace
  {
    // particle_heat_transfer.cc:9-146
    __device__ __forceinline__ double th_harmonic_mean(double a, double b) { return (2 * a * b / (a + b + DBL_MIN)); }
    __device__ __forceinline__ double th_corrected_contact_radius(double effective_radius, double effective_youngs_modulus,
                                                                  double effective_real_youngs_modulus, double normal_force_norm)
    {
      const double contact_radius = pow((3 * normal_force_norm * effective_radius) / (4 * effective_youngs_modulus), (1.0 / 3.0));
      return contact_radius * pow(effective_youngs_modulus / effective_real_youngs_modulus, 1.0 / 5.0);
    }
    __device__ __forceinline__ double th_macrocontact(double harmonic_conductivity, double contact_radius)
    {
      return 0.5 / (contact_radius * harmonic_conductivity + DBL_MIN);
    }
    __device__ __forceinline__ double th_microcontact(double slope, double roughness, double microhardness, double contact_radius_squared,
                                                      double harmonic_conductivity, double maximum_pressure)
    {
      return 1.184 / (M_PI * harmonic_conductivity * contact_radius_squared) * (roughness / slope) *
             pow(microhardness / (maximum_pressure + DBL_MIN), 0.96);
    }
    __device__ __forceinline__ double th_solid_macrogap(double radius, double thermal_conductivity, double contact_radius_squared)
    {
      return 0.25 * M_PI * radius / (M_PI * (radius * radius - contact_radius_squared) * thermal_conductivity);
    }
    __device__ __forceinline__ double th_gas_microgap(double roughness, double contact_radius_squared, double gas_parameter_m,
                                                      double thermal_conductivity_gas, double maximum_pressure, double microhardness)
    {
      const double x_1 = 2.0 * maximum_pressure / microhardness;
      const double x_2 = 0.03 * maximum_pressure / microhardness;
      if (x_1 >= 2.0 || x_1 <= 0.0)
        return INFINITY;
      const double a_1 = erfcinv(x_1); // boost::math::erfc_inv
      const double a_2 = erfcinv(x_2) - a_1;
      return (2.82842712475 * roughness * a_2) /
             (M_PI * thermal_conductivity_gas * contact_radius_squared * log(fabs(1 + a_2 / (a_1 + gas_parameter_m / (2.82842712475 * roughness)))));
    }
    __device__ __forceinline__ double th_gas_macrogap(double harmonic_radius, double thermal_conductivity_gas, double contact_radius_squared,
                                                      double gas_parameter_m)
    {
      const double A = 2. * sqrt(harmonic_radius * harmonic_radius - contact_radius_squared);
      const double S = 2. * harmonic_radius - contact_radius_squared / harmonic_radius + gas_parameter_m;
      return 2.0 / (M_PI * thermal_conductivity_gas * (S * log(S / (S - A)) - A));
    }
    // calculate_contact_thermal_conductance<particle-particle> (particle_heat_transfer.cc:148-318)
    __device__ double th_contact_conductance(double radius_one, double radius_two, double effective_youngs_modulus, double effective_real_youngs_modulus,
                                             double roughness, double slope, double microhardness, double thermal_conductivity_one,
                                             double thermal_conductivity_two, double thermal_conductivity_gas, double gas_parameter_m,
                                             double normal_overlap, double normal_force_norm)
    {
      const double harmonic_conductivity = th_harmonic_mean(thermal_conductivity_one, thermal_conductivity_two);
      const double harmonic_radius = th_harmonic_mean(radius_one, radius_two);
      const double contact_radius =
        th_corrected_contact_radius(harmonic_radius * 0.5, effective_youngs_modulus, effective_real_youngs_modulus, normal_force_norm);
      const double corrected_normal_overlap = normal_overlap * pow(effective_youngs_modulus / effective_real_youngs_modulus, 2.0 / 3.0);
      const double contact_radius_squared = contact_radius * contact_radius;
      const double maximum_pressure = (2.0 * effective_real_youngs_modulus * corrected_normal_overlap) / (M_PI * contact_radius + DBL_MIN);
      const double resistance_macrocontact = th_macrocontact(harmonic_conductivity, contact_radius);
      const double resistance_microcontact =
        th_microcontact(slope, roughness, microhardness, contact_radius_squared, harmonic_conductivity, maximum_pressure);
      double resistance_solid_macrogap = th_solid_macrogap(radius_one, thermal_conductivity_one, contact_radius_squared);
      resistance_solid_macrogap += th_solid_macrogap(radius_two, thermal_conductivity_two, contact_radius_squared);
      const double resistance_gas_microgap =
        th_gas_microgap(roughness, contact_radius_squared, gas_parameter_m, thermal_conductivity_gas, maximum_pressure, microhardness);
      const double resistance_gas_macrogap = th_gas_macrogap(harmonic_radius, thermal_conductivity_gas, contact_radius_squared, gas_parameter_m);
      return 1.0 / (resistance_macrocontact + 1.0 / (1.0 / resistance_microcontact + 1.0 / resistance_gas_microgap)) +
             1.0 / (resistance_solid_macrogap + resistance_gas_macrogap);
    }

    // One thread per row: the normal force of every touching pair is re-evaluated with the contact model of the step kernel
    // (it depends on positions and velocities only, not on the history), in the pair's canonical orientation, so both
    // owners of a pair get exactly opposite rates.
    template <int MODEL, bool PERIODIC>
    __global__ void __launch_bounds__(128) k_heat_rates(const __grid_constant__ HeatParams P, const __grid_constant__ MaterialTables mt,
                                                         const __grid_constant__ ThermalTables th)
    {
      const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= P.n_owned)
        return;
      const double4 pme = P.in.pos[i];
      const ParticleView me = make_view(pme, P.in.vel[i], P.in.omg[i]);
      const double temperature_me = P.temperature[P.id[i]];
      double rate = 0.0;
      for (uint32_t e = P.list.row_start[i]; e < P.list.row_start[i + 1]; ++e)
        {
          const uint32_t j = P.list.col[e] & COL_INDEX_MASK;
          const double4 pj = P.in.pos[j];
          vec3 x1, x2;
          double dsum;
          bool i_am_two = false;
          if constexpr (PERIODIC)
            {
              const uint32_t img = P.list.img[e];
              if (img)
                {
                  vec3 shift;
                  decode_image(img, P.L, shift, i_am_two);
                  if (!i_am_two)
                    {
                      x1 = v3(pme.x, pme.y, pme.z);
                      x2 = v3(pj.x, pj.y, pj.z) + shift;
                    }
                  else
                    {
                      x1 = v3(pj.x, pj.y, pj.z);
                      x2 = v3(pme.x, pme.y, pme.z) + (-shift);
                    }
                }
              else
                {
                  x1 = v3(pme.x, pme.y, pme.z);
                  x2 = v3(pj.x, pj.y, pj.z);
                }
            }
          else
            {
              x1 = v3(pme.x, pme.y, pme.z);
              x2 = v3(pj.x, pj.y, pj.z);
            }
          dsum = i_am_two ? pj.w + pme.w : pme.w + pj.w;
          const double distance = sqrt(dist2(x1, x2));
          const double normal_overlap = 0.5 * dsum - distance;
          if (!(normal_overlap > mt.pp_force_threshold) || !(normal_overlap > 0))
            continue;
          const ParticleView other = make_view(pj, P.in.vel[j], P.in.omg[j]);
          ParticleView one = i_am_two ? other : me, two = i_am_two ? me : other;
          one.x = x1;
          // R* and m* of two copies of particle one: what the model's equal-size shortcut reads (bit-identical to the general branch)
          const SelfPair sp{(one.d * one.d) / (2 * (one.d + one.d)), (one.m * one.m) / (one.m + one.m)};
          vec3 h = v3(0, 0, 0), rs = v3(0, 0, 0), n, vt;
          double vn;
          PairResult r;
          r.normal_force = r.tangential_force = r.torque_one = r.torque_two = r.rolling = v3(0, 0, 0);
          pp_update_contact_information<double>(h, vt, vn, n, one, two, x2, distance, P.dt);
          pp_calculate_contact<MODEL, LETHE_ROLLING_NONE, double>(mt, h, rs, vt, vn, n, normal_overlap, P.dt, one, two, r, sp);
          const int k = one.type * mt.n_types + two.type;
          const double conductance =
            th_contact_conductance(0.5 * one.d, 0.5 * two.d, mt.Y[k], th.real_E[k], th.roughness[k], th.slope[k], th.microhardness[k],
                                   th.conductivity[one.type], th.conductivity[two.type], th.conductivity_gas, th.gas_m[k], normal_overlap,
                                   norm(r.normal_force));
          // apply_heat_transfer_on_local_particles (particle_heat_transfer.cc:320-331): one gains G (T2 - T1), two loses it
          const double temperature_other = P.temperature[P.id[j]];
          const double t_one = i_am_two ? temperature_other : temperature_me, t_two = i_am_two ? temperature_me : temperature_other;
          const double heat_transfer_rate = conductance * (t_two - t_one);
          rate = i_am_two ? rate - heat_transfer_rate : rate + heat_transfer_rate;
        }
      P.rate[i] = rate;
    }

    __global__ void __launch_bounds__(256) k_integrate_temperature(const __grid_constant__ HeatParams P)
    {
      const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= P.n_owned)
        return;
      const uint32_t id = P.id[i];
      const double mass_inverse = 1 / P.in.vel[i].w;
      const double specific_heat_inverse = 1 / P.specific_heat[id];
      P.temperature[id] = P.temperature[id] + P.dt * (P.rate[i] + 0.0) * mass_inverse * specific_heat_inverse;
    }

    template <int MODEL> void launch_heat_m(const HeatParams &p, const MaterialTables &mt, const ThermalTables &th, cudaStream_t stream)
    {
      const unsigned blocks = (p.n_owned + 127) / 128;
      if (p.periodic_any)
        k_heat_rates<MODEL, true><<<blocks, 128, 0, stream>>>(p, mt, th);
      else
        k_heat_rates<MODEL, false><<<blocks, 128, 0, stream>>>(p, mt, th);
      count_launch();
    }
  } // namespace

  void launch_heat_rates(int pp_model, const HeatParams &p, const MaterialTables &mt, const ThermalTables &th, cudaStream_t stream)
  {
    if (p.n_owned == 0)
      return;
    switch (pp_model)
      {
        case LETHE_PP_LINEAR:
          launch_heat_m<LETHE_PP_LINEAR>(p, mt, th, stream);
          break;
        case LETHE_PP_HERTZ_MINDLIN_LIMIT_FORCE:
          launch_heat_m<LETHE_PP_HERTZ_MINDLIN_LIMIT_FORCE>(p, mt, th, stream);
          break;
        case LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP:
          launch_heat_m<LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP>(p, mt, th, stream);
          break;
        case LETHE_PP_HERTZ:
          launch_heat_m<LETHE_PP_HERTZ>(p, mt, th, stream);
          break;
        case LETHE_PP_HERTZ_JKR:
          launch_heat_m<LETHE_PP_HERTZ_JKR>(p, mt, th, stream);
          break;
        default:
          launch_heat_m<LETHE_PP_DMT>(p, mt, th, stream);
          break;
      }
  }

  void launch_integrate_temperature(const HeatParams &p, cudaStream_t stream)
  {
    if (p.n_owned == 0)
      return;
    k_integrate_temperature<<<(p.n_owned + 255) / 256, 256, 0, stream>>>(p);
    count_launch();
  }

  void launch_step(int pp_model, int rolling_model, const StepParams &p, const MaterialTables &mt, cudaStream_t stream)
  {
#ifdef DEM_BENCH_VARIANT
    // timing builds (tools/build_variant.sh): only the model of the bench workloads is instantiated
    if (pp_model != LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP)
      abort();
    launch_m<LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP>(rolling_model, p, mt, stream);
    return;
#else
    switch (pp_model)
      {
        case LETHE_PP_LINEAR:
          launch_m<LETHE_PP_LINEAR>(rolling_model, p, mt, stream);
          break;
        case LETHE_PP_HERTZ_MINDLIN_LIMIT_FORCE:
          launch_m<LETHE_PP_HERTZ_MINDLIN_LIMIT_FORCE>(rolling_model, p, mt, stream);
          break;
        case LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP:
          launch_m<LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP>(rolling_model, p, mt, stream);
          break;
        case LETHE_PP_HERTZ:
          launch_m<LETHE_PP_HERTZ>(rolling_model, p, mt, stream);
          break;
        case LETHE_PP_HERTZ_JKR:
          launch_m<LETHE_PP_HERTZ_JKR>(rolling_model, p, mt, stream);
          break;
        default:
          launch_m<LETHE_PP_DMT>(rolling_model, p, mt, stream);
          break;
      }
#endif
  }
} // namespace dem
