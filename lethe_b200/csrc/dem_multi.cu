// dem_multi.cu — slab decomposition over the GPUs of one node: one context per GPU per
// process, NCCL point-to-point between neighbouring slabs.
//
// Reference behaviour mirrored (MPI through deal.II):
//   every step      particle_handler.update_ghost_particles()                 (dem.cc:686)
//                   + logical_or of the contact-detection flag                (find_contact_detection_step.cc:53-58)
//   rebuild steps   sort_particles_into_subdomains_and_cells() (migration)
//                   + exchange_ghost_particles(true)                           (dem.cc:986-989)
// Cross-slab pairs are evaluated on both ranks and applied to the owned particle only, each
// rank keeping its own copy of the pair history (…contact_force.h:2022-2033); the history of a
// pair restarts from zero when one of its particles changes owner
// (update_fine_search_candidates.cc:136-152).
//
// Layout: ghost copies live behind the owned particles in the same SoA arrays, as two runs
// (from the lower / upper neighbour), each already sorted by (Morton cell rank, id) because
// the sender's particles are. The per-step refresh therefore needs no unpack kernel: the
// three state arrays are received straight into place.
//
// NCCL is loaded with dlopen so that single-GPU use has no dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "dem_context.cuh"

namespace dem
{
  namespace
  {
    struct NcclApi
    {
      void *handle = nullptr;
      ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
      ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
      ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
      ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
      ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
      ncclResult_t (*GroupStart)() = nullptr;
      ncclResult_t (*GroupEnd)() = nullptr;
      ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
      ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
      const char *(*GetErrorString)(ncclResult_t) = nullptr;
      bool load()
      {
        if (handle)
          return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"})
          {
            handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (handle)
              break;
          }
        if (!handle)
          return false;
#define LOAD(sym) *(void **)(&sym) = dlsym(handle, "nccl" #sym)
        LOAD(GetUniqueId);
        LOAD(CommInitRank);
        LOAD(CommDestroy);
        LOAD(Send);
        LOAD(Recv);
        LOAD(GroupStart);
        LOAD(GroupEnd);
        LOAD(AllReduce);
        LOAD(AllGather);
        LOAD(GetErrorString);
#undef LOAD
        return GetUniqueId && CommInitRank && Send && Recv && GroupStart && GroupEnd && AllReduce;
      }
    };
    NcclApi g_nccl;

    inline void nccl_check(ncclResult_t r, const char *what)
    {
      if (r != ncclSuccess)
        throw std::runtime_error(std::string(what) + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error"));
    }
#define NCCL_TRY(call) nccl_check((call), #call)
  } // namespace

  struct MultiGpuImpl
  {
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    int peer[2] = {-1, -1}; // lower / upper neighbour rank or -1
    // migration buffers
    DevBuf<MigrateRecord> send_rec[2], recv_rec[2];
    DevBuf<uint32_t> send_id[2], recv_id[2], send_slot[2];
    DevBuf<HistRecord> send_hist[2]; // contact history of the emigrants (HistRecord)
    DevBuf<uint32_t> hist_offsets[2];
    DevBuf<uint32_t> counters; // [4] device
    DevBuf<uint32_t> xcount;   // [4] device: counts exchanged with the peers
    // halo: indices of my boundary-layer particles per direction + packed send buffers
    DevBuf<uint32_t> flags, offsets, send_idx[2], send_ids[2];
    DevBuf<double4> send_pos[2], send_vel[2], send_omg[2];
    uint32_t n_send[2] = {0, 0};
    DevBuf<int> flag_dev; // [2]
    int *flag_host = nullptr;

    // ---- fused halo over peer memory ----
    bool want_fused = true;   // LETHE_DEM_HALO=nccl keeps the send/recv halo
    bool fused_ready = false; // every rank mapped its neighbours' state arrays
    struct Mapping
    {
      cudaIpcMemHandle_t handle;
      void *base;
      uint64_t epoch; // exchange at which the mapping was last referenced
    };
    std::vector<Mapping> mappings;
    double4 *peer_arr[2][2][3] = {}; // [direction][peer generation][pos, vel, omg]
    uint32_t peer_base[2] = {0, 0};
    int gen_xor[2] = {0, 0}; // my generation index -> the peer's (both flip in lockstep)
    DevBuf<uint32_t> halo_bits[2], halo_prefix[2];
    DevBuf<uint8_t> info_dev;
    uint64_t epoch = 0;
    std::vector<void *> retired;                           // outgrown state arrays (DevBuf::retire)
    std::vector<std::pair<void *, uint64_t>> graveyard;    // ... with the exchange they were retired at
    // per-step agreement over peer memory (launch_agree); NCCL all-reduce when not available
    bool want_mailbox = true; // LETHE_DEM_AGREE=nccl keeps the library collective
    bool mailbox_ready = false, mailbox_tried = false;
    uint64_t *mailbox = nullptr;
    DevBuf<uint64_t *> peer_mailbox;
    std::vector<void *> mailbox_maps;
    uint32_t agree_seq = 0;
    uint32_t *agreed_host = nullptr;                       // mapped pinned [2]
    uint32_t *agreed_host_dev = nullptr;                   // its device alias
    cudaEvent_t agreed_ev[2] = {nullptr, nullptr};
    int agreed_slot = 0;
  };

  namespace
  {
    // what a rank tells the neighbour that pushes into one of its ghost runs
    struct HaloInfo
    {
      cudaIpcMemHandle_t handle[2][3];
      uint64_t offset[2][3]; // of the array inside the exported allocation
      uint32_t cur;          // current state generation of the sender
      uint32_t base;         // first slot of the ghost run
      uint32_t ok;
      uint32_t pad;
    };
  } // namespace

  int MultiGpu::unique_id(uint8_t *id128)
  {
    if (!g_nccl.load())
      return -1;
    static_assert(sizeof(ncclUniqueId) == LETHE_DEM_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess)
      return -2;
    std::memcpy(id128, &id, sizeof(id));
    return 0;
  }

  void MultiGpu::init(lethe_dem_ctx *c, int rank, int world, const uint8_t *id128)
  {
    if (!g_nccl.load())
      throw std::runtime_error("libnccl.so.2 not found: multi-GPU needs NCCL");
    if (c->grid.slab_axis < 0 || c->grid.slab_axis > 2)
      throw std::runtime_error("config.slab_axis must be 0..2 for a multi-GPU context");
    if (c->grid.slab_hi - c->grid.slab_lo < 2)
      throw std::runtime_error("every slab needs at least 2 cell layers");
    if (impl)
      throw std::runtime_error("communicator already initialised");
    if (c->thermal_enabled)
      throw std::runtime_error("heat transfer runs on a single GPU (ghost temperatures are not exchanged)");
    MultiGpuImpl *m = new MultiGpuImpl();
    m->rank = rank;
    m->world = world;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    NCCL_TRY(g_nccl.CommInitRank(&m->comm, world, id, rank));
    const bool periodic = c->grid.periodic[c->grid.slab_axis] != 0;
    m->peer[0] = rank > 0 ? rank - 1 : (periodic && world > 1 ? world - 1 : -1);
    m->peer[1] = rank < world - 1 ? rank + 1 : (periodic && world > 1 ? 0 : -1);
    m->counters.ensure(8);
    m->xcount.ensure(8);
    m->flag_dev.ensure(2);
    CU_TRY(cudaHostAlloc(&m->flag_host, 2 * sizeof(int), cudaHostAllocDefault));
    if (const char *e = getenv("LETHE_DEM_HALO"))
      m->want_fused = std::strcmp(e, "nccl") != 0;
    if (const char *e = getenv("LETHE_DEM_AGREE"))
      m->want_mailbox = std::strcmp(e, "nccl") != 0;
    if (world < 2)
      m->want_fused = false;
    if (world > 32 || !g_nccl.AllGather)
      m->want_mailbox = false;
    if (m->want_fused)
      {
        CU_TRY(cudaHostAlloc(&m->agreed_host, 2 * sizeof(uint32_t), cudaHostAllocMapped));
        m->agreed_host[0] = m->agreed_host[1] = 0;
        CU_TRY(cudaHostGetDevicePointer(reinterpret_cast<void **>(&m->agreed_host_dev), m->agreed_host, 0));
        for (auto &ev : m->agreed_ev)
          CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        m->info_dev.ensure(4 * sizeof(HaloInfo));
        // state arrays a neighbour may have mapped are never freed under its feet
        for (int g = 0; g < 2; ++g)
          {
            c->st[g].pos.retire = &m->retired;
            c->st[g].vel.retire = &m->retired;
            c->st[g].omg.retire = &m->retired;
          }
      }
    impl = m;
    c->contact_search_trigger = true;
  }

  void MultiGpu::shutdown()
  {
    if (!impl)
      return;
    if (impl->comm && g_nccl.CommDestroy)
      g_nccl.CommDestroy(impl->comm);
    if (impl->flag_host)
      cudaFreeHost(impl->flag_host);
    if (impl->agreed_host)
      cudaFreeHost(impl->agreed_host);
    for (auto ev : impl->agreed_ev)
      if (ev)
        cudaEventDestroy(ev);
    for (auto &mp : impl->mappings)
      cudaIpcCloseMemHandle(mp.base);
    for (void *q : impl->mailbox_maps)
      cudaIpcCloseMemHandle(q);
    if (impl->mailbox)
      cudaFree(impl->mailbox);
    for (void *q : impl->retired)
      cudaFree(q);
    for (auto &g : impl->graveyard)
      cudaFree(g.first);
    delete impl;
    impl = nullptr;
  }

  bool MultiGpu::agree(lethe_dem_ctx *c, bool local)
  {
    MultiGpuImpl *m = impl;
    cudaStream_t s = c->stream;
    m->flag_host[0] = local ? 1 : 0;
    CU_TRY(cudaMemcpyAsync(m->flag_dev.p, m->flag_host, sizeof(int), cudaMemcpyHostToDevice, s));
    NCCL_TRY(g_nccl.AllReduce(m->flag_dev.p, m->flag_dev.p, 1, ncclInt32, ncclMax, m->comm, s));
    CU_TRY(cudaMemcpyAsync(m->flag_host, m->flag_dev.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    return m->flag_host[0] != 0;
  }

  bool MultiGpu::load_balance_due(lethe_dem_ctx *c)
  {
    MultiGpuImpl *m = impl;
    if (c->lb_method == LETHE_LOAD_BALANCE_NONE || m->world < 2)
      return false;
    const uint64_t it = c->iteration_number;
    const uint64_t freq = uint64_t(std::max(1, c->lb_frequency));
    if (c->lb_method == LETHE_LOAD_BALANCE_ONCE)
      return it == freq; // `step` of the once method
    if (c->lb_method == LETHE_LOAD_BALANCE_FREQUENT)
      return (it % freq) == 0;
    if ((it % freq) != 0)
      return false;
    if (c->lb_method == LETHE_LOAD_BALANCE_DYNAMIC_WITH_SPARSE_CONTACTS)
      {
        // check_load_balance_with_sparse_contacts (load_balancing.cc:60-122): cell weight per owned cell + particle weight x
        // mobility factor per particle; repartition if (max - min) load > threshold * total / ranks
        cudaStream_t s = c->stream;
        DevBuf<uint32_t> acc;
        acc.ensure(size_t(c->grid.n[c->grid.slab_axis]) + 4);
        CU_TRY(cudaMemsetAsync(acc.p, 0, (size_t(c->grid.n[c->grid.slab_axis]) + 4) * 4, s));
        if (c->asc_in_force)
          launch_layer_histogram_weighted(c->st[c->cur].pos.p, c->st[c->cur].cell_reg.p, c->asc_cell_status.p, c->grid, c->n_owned, 1000u,
                                          uint32_t(std::lround(1000.0 * c->lb_active_factor)), uint32_t(std::lround(1000.0 * c->lb_inactive_factor)),
                                          acc.p, s);
        else
          launch_layer_histogram(c->st[c->cur].pos.p, c->grid, c->n_owned, acc.p, s);
        std::vector<uint32_t> hh(size_t(c->grid.n[c->grid.slab_axis]));
        CU_TRY(cudaMemcpyAsync(hh.data(), acc.p, hh.size() * 4, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        double weighted = 0;
        for (uint32_t v : hh)
          weighted += v;
        if (c->asc_in_force)
          weighted /= 1000.0;
        const int a = c->grid.slab_axis;
        const double cells = double(c->grid.slab_hi - c->grid.slab_lo) * (double(c->grid.n_cells) / c->grid.n[a]);
        const double load = cells * c->lb_cell_weight + weighted * c->lb_particle_weight;
        double v[3] = {load, -load, load};
        DevBuf<double> dv;
        dv.ensure(3);
        CU_TRY(cudaMemcpyAsync(dv.p, v, 24, cudaMemcpyHostToDevice, s));
        NCCL_TRY(g_nccl.AllReduce(dv.p, dv.p, 2, ncclDouble, ncclMax, m->comm, s));
        NCCL_TRY(g_nccl.AllReduce(dv.p + 2, dv.p + 2, 1, ncclDouble, ncclSum, m->comm, s));
        CU_TRY(cudaMemcpyAsync(v, dv.p, 24, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        return (v[0] - (-v[1])) > c->lb_threshold * (v[2] / m->world);
      }
    // dynamic: (max - min) particles per rank > threshold * (local particles / ranks). The reference
    // evaluates this with each rank's own count on the right-hand side (load_balancing.cc:44-55);
    // here any rank that finds it true makes all of them repartition.
    cudaStream_t s = c->stream;
    uint32_t h[2] = {c->n_owned, ~c->n_owned};
    CU_TRY(cudaMemcpyAsync(m->xcount.p, h, 8, cudaMemcpyHostToDevice, s));
    NCCL_TRY(g_nccl.AllReduce(m->xcount.p, m->xcount.p, 2, ncclUint32, ncclMax, m->comm, s));
    CU_TRY(cudaMemcpyAsync(h, m->xcount.p, 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    const uint32_t n_max = h[0], n_min = ~h[1];
    const bool local = double(n_max - n_min) > c->lb_threshold * double(c->n_owned / uint32_t(m->world));
    return agree(c, local);
  }

  namespace
  {
    // buf layout: [0] what I send down, [1] what I send up, [2] received from below, [3] received from above, `count` ints each
    void asc_exchange(lethe_dem_ctx *c, MultiGpuImpl *m, int *buf, size_t count)
    {
      cudaStream_t s = c->stream;
      NCCL_TRY(g_nccl.GroupStart());
      if (m->peer[1] >= 0)
        NCCL_TRY(g_nccl.Send(buf + 1 * count, count, ncclInt32, m->peer[1], m->comm, s));
      if (m->peer[0] >= 0)
        NCCL_TRY(g_nccl.Recv(buf + 2 * count, count, ncclInt32, m->peer[0], m->comm, s));
      if (m->peer[0] >= 0)
        NCCL_TRY(g_nccl.Send(buf + 0 * count, count, ncclInt32, m->peer[0], m->comm, s));
      if (m->peer[1] >= 0)
        NCCL_TRY(g_nccl.Recv(buf + 3 * count, count, ncclInt32, m->peer[1], m->comm, s));
      NCCL_TRY(g_nccl.GroupEnd());
    }
  } // namespace

  void MultiGpu::asc_exchange_nodes(lethe_dem_ctx *c)
  {
    MultiGpuImpl *m = impl;
    cudaStream_t s = c->stream;
    const GridDesc &g = c->grid;
    const int a = g.slab_axis, u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
    const size_t count = size_t(g.n[u] + 1) * (g.n[v] + 1);
    c->asc_xbuf.ensure(4 * count);
    AscPlaneParams pp;
    pp.grid = g;
    pp.node_status = c->asc_node_status.p;
    pp.cell_status = c->asc_cell_status.p;
    // my lower cut plane is node index slab_lo, my upper one slab_hi (asc_node wraps it on a periodic axis)
    pp.index = g.slab_lo;
    pp.buf = c->asc_xbuf.p + 0 * count;
    launch_asc_plane(pp, 0, s);
    pp.index = g.slab_hi;
    pp.buf = c->asc_xbuf.p + 1 * count;
    launch_asc_plane(pp, 0, s);
    asc_exchange(c, m, c->asc_xbuf.p, count);
    if (m->peer[0] >= 0)
      {
        pp.index = g.slab_lo;
        pp.buf = c->asc_xbuf.p + 2 * count;
        launch_asc_plane(pp, 1, s);
      }
    if (m->peer[1] >= 0)
      {
        pp.index = g.slab_hi;
        pp.buf = c->asc_xbuf.p + 3 * count;
        launch_asc_plane(pp, 1, s);
      }
  }

  void MultiGpu::asc_exchange_cells(lethe_dem_ctx *c)
  {
    MultiGpuImpl *m = impl;
    cudaStream_t s = c->stream;
    const GridDesc &g = c->grid;
    const int a = g.slab_axis, u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2, na = g.n[a];
    const size_t count = size_t(g.n[u]) * g.n[v];
    c->asc_xbuf.ensure(4 * count);
    AscPlaneParams pp;
    pp.grid = g;
    pp.node_status = c->asc_node_status.p;
    pp.cell_status = c->asc_cell_status.p;
    pp.index = g.slab_lo; // my lowest layer goes down
    pp.buf = c->asc_xbuf.p + 0 * count;
    launch_asc_plane(pp, 2, s);
    pp.index = g.slab_hi - 1; // my highest layer goes up
    pp.buf = c->asc_xbuf.p + 1 * count;
    launch_asc_plane(pp, 2, s);
    asc_exchange(c, m, c->asc_xbuf.p, count);
    if (m->peer[0] >= 0)
      {
        pp.index = (g.slab_lo - 1 + na) % na; // the lower neighbour's highest layer
        pp.buf = c->asc_xbuf.p + 2 * count;
        launch_asc_plane(pp, 3, s);
      }
    if (m->peer[1] >= 0)
      {
        pp.index = g.slab_hi % na; // the upper neighbour's lowest layer
        pp.buf = c->asc_xbuf.p + 3 * count;
        launch_asc_plane(pp, 3, s);
      }
  }

  bool MultiGpu::any_rank_flag(lethe_dem_ctx *c)
  {
    CU_TRY(cudaStreamSynchronize(c->stream));
    return *c->h_flag != 0; // agree() is applied by the caller to the combined decision
  }

  namespace
  {
    // exchange `n_send[d]` -> `n_recv[d]` (one u32 per direction) with the neighbours
    void exchange_counts(lethe_dem_ctx *c, MultiGpuImpl *m, const uint32_t n_send[2], uint32_t n_recv[2])
    {
      cudaStream_t s = c->stream;
      uint32_t h[4] = {n_send[0], n_send[1], 0, 0};
      CU_TRY(cudaMemcpyAsync(m->xcount.p, h, 16, cudaMemcpyHostToDevice, s));
      // phase A: count of what goes up, received from below; phase B: the other way round
      // (see for_each_direction_pair for why the order matters when both peers are one rank)
      NCCL_TRY(g_nccl.GroupStart());
      if (m->peer[1] >= 0)
        NCCL_TRY(g_nccl.Send(m->xcount.p + 1, 1, ncclUint32, m->peer[1], m->comm, s));
      if (m->peer[0] >= 0)
        NCCL_TRY(g_nccl.Recv(m->xcount.p + 2, 1, ncclUint32, m->peer[0], m->comm, s));
      if (m->peer[0] >= 0)
        NCCL_TRY(g_nccl.Send(m->xcount.p + 0, 1, ncclUint32, m->peer[0], m->comm, s));
      if (m->peer[1] >= 0)
        NCCL_TRY(g_nccl.Recv(m->xcount.p + 3, 1, ncclUint32, m->peer[1], m->comm, s));
      NCCL_TRY(g_nccl.GroupEnd());
      CU_TRY(cudaMemcpyAsync(h, m->xcount.p, 16, cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      n_recv[0] = m->peer[0] >= 0 ? h[2] : 0;
      n_recv[1] = m->peer[1] >= 0 ? h[3] : 0;
    }

    // With world == 2 and a periodic axis both neighbours are the same rank: NCCL matches the
    // two send/recv pairs between the same peers in issue order, so rank 0 must post its
    // (dir 0, dir 1) operations in the opposite order of rank 1's. Directions are therefore
    // always issued as "send up / recv from below" first, then "send down / recv from above".
    template <class F> void for_each_direction_pair(MultiGpuImpl *m, F &&f)
    {
      // phase A: send towards upper (dir 1), receive from lower (dir 0)
      f(1, 0);
      // phase B: send towards lower (dir 0), receive from upper (dir 1)
      f(0, 1);
      (void)m;
    }
  } // namespace

  namespace
  {
    typedef int (*cuMemGetAddressRange_t)(unsigned long long *, size_t *, unsigned long long);

    // (allocation base, offset) of a device pointer: cudaMalloc may sub-allocate, and an IPC
    // handle always names the whole allocation
    bool allocation_base(const void *p, void **base, uint64_t *offset)
    {
      static cuMemGetAddressRange_t fn = nullptr;
      static bool tried = false;
      if (!tried)
        {
          tried = true;
          void *f = nullptr;
          cudaDriverEntryPointQueryResult q;
          if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<cuMemGetAddressRange_t>(f);
        }
      if (!fn)
        return false;
      unsigned long long b = 0;
      size_t sz = 0;
      if (fn(&b, &sz, (unsigned long long)(uintptr_t)p) != 0)
        return false;
      *base = reinterpret_cast<void *>(uintptr_t(b));
      *offset = uint64_t(uintptr_t(p)) - uint64_t(b);
      return true;
    }

    void *map_peer_allocation(MultiGpuImpl *m, const cudaIpcMemHandle_t &h)
    {
      for (auto &mp : m->mappings)
        if (std::memcmp(&mp.handle, &h, sizeof(h)) == 0)
          {
            mp.epoch = m->epoch;
            return mp.base;
          }
      void *base = nullptr;
      if (cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        {
          cudaGetLastError();
          return nullptr;
        }
      m->mappings.push_back({h, base, m->epoch});
      return base;
    }

    // Once: every rank maps every rank's agreement mailbox (launch_agree). Collective; the result
    // (usable or not) is the same on all ranks.
    void bootstrap_mailboxes(lethe_dem_ctx *c, MultiGpuImpl *m)
    {
      cudaStream_t s = c->stream;
      m->mailbox_tried = true;
      struct Slot
      {
        cudaIpcMemHandle_t handle;
        uint64_t offset;
        uint64_t ok;
      };
      const int W = m->world;
      Slot mine;
      std::memset(&mine, 0, sizeof(mine));
      bool ok = cudaMalloc(&m->mailbox, size_t(2) * W * sizeof(uint64_t)) == cudaSuccess;
      void *base = nullptr;
      if (ok)
        {
          CU_TRY(cudaMemsetAsync(m->mailbox, 0, size_t(2) * W * sizeof(uint64_t), s));
          ok = allocation_base(m->mailbox, &base, &mine.offset) && cudaIpcGetMemHandle(&mine.handle, base) == cudaSuccess;
        }
      if (!ok)
        cudaGetLastError();
      mine.ok = ok ? 1 : 0;
      DevBuf<uint8_t> buf;
      buf.ensure(size_t(W + 1) * sizeof(Slot));
      CU_TRY(cudaMemcpyAsync(buf.p + size_t(W) * sizeof(Slot), &mine, sizeof(Slot), cudaMemcpyHostToDevice, s));
      NCCL_TRY(g_nccl.AllGather(buf.p + size_t(W) * sizeof(Slot), buf.p, sizeof(Slot), ncclUint8, m->comm, s));
      std::vector<Slot> all(W);
      CU_TRY(cudaMemcpyAsync(all.data(), buf.p, size_t(W) * sizeof(Slot), cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      std::vector<uint64_t *> ptrs(W, nullptr);
      for (int r = 0; r < W; ++r)
        {
          if (!all[r].ok)
            ok = false;
          if (!ok)
            break;
          if (r == m->rank)
            {
              ptrs[r] = m->mailbox;
              continue;
            }
          void *b = nullptr;
          if (cudaIpcOpenMemHandle(&b, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
            {
              cudaGetLastError();
              ok = false;
              break;
            }
          m->mailbox_maps.push_back(b);
          ptrs[r] = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(b) + all[r].offset);
        }
      if (ok)
        {
          m->peer_mailbox.ensure(W);
          CU_TRY(cudaMemcpyAsync(m->peer_mailbox.p, ptrs.data(), size_t(W) * sizeof(uint64_t *), cudaMemcpyHostToDevice, s));
        }
      uint32_t all_ok = ok ? 1u : 0u;
      CU_TRY(cudaMemcpyAsync(m->xcount.p, &all_ok, 4, cudaMemcpyHostToDevice, s));
      NCCL_TRY(g_nccl.AllReduce(m->xcount.p, m->xcount.p, 1, ncclUint32, ncclMin, m->comm, s));
      CU_TRY(cudaMemcpyAsync(&all_ok, m->xcount.p, 4, cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s)); // every mailbox is zeroed and mapped before anyone posts
      m->mailbox_ready = all_ok != 0;
    }

    // Every rebuild: tell the two neighbours where their pushes must land (IPC handles of my six
    // state arrays, my generation, the first slot of the ghost run they fill) and map theirs.
    void exchange_halo_info(lethe_dem_ctx *c, MultiGpuImpl *m, const uint32_t g_recv[2])
    {
      cudaStream_t s = c->stream;
      if (m->want_mailbox && !m->mailbox_tried)
        bootstrap_mailboxes(c, m);
      ++m->epoch;
      HaloInfo mine[2], theirs[2];
      std::memset(mine, 0, sizeof(mine));
      std::memset(theirs, 0, sizeof(theirs));
      bool ok = true;
      HaloInfo base_info;
      std::memset(&base_info, 0, sizeof(base_info));
      for (int g = 0; g < 2 && ok; ++g)
        {
          double4 *arr[3] = {c->st[g].pos.p, c->st[g].vel.p, c->st[g].omg.p};
          for (int a = 0; a < 3 && ok; ++a)
            {
              void *b = nullptr;
              ok = arr[a] && allocation_base(arr[a], &b, &base_info.offset[g][a]) &&
                   cudaIpcGetMemHandle(&base_info.handle[g][a], b) == cudaSuccess;
            }
        }
      if (!ok)
        cudaGetLastError();
      base_info.cur = uint32_t(c->cur);
      base_info.ok = ok ? 1u : 0u;
      // run 0 is filled by the lower neighbour, run 1 by the upper one
      for (int r = 0; r < 2; ++r)
        {
          mine[r] = base_info;
          mine[r].base = c->n_owned + (r == 1 ? g_recv[0] : 0u);
        }
      uint8_t *dev = m->info_dev.p;
      CU_TRY(cudaMemcpyAsync(dev, mine, 2 * sizeof(HaloInfo), cudaMemcpyHostToDevice, s));
      NCCL_TRY(g_nccl.GroupStart());
      // phase A: my run-1 description goes up; from below comes the lower neighbour's run 1 (the
      // one I fill when pushing in direction 0). Phase B: the mirror image.
      if (m->peer[1] >= 0)
        NCCL_TRY(g_nccl.Send(dev + 1 * sizeof(HaloInfo), sizeof(HaloInfo), ncclUint8, m->peer[1], m->comm, s));
      if (m->peer[0] >= 0)
        NCCL_TRY(g_nccl.Recv(dev + 2 * sizeof(HaloInfo), sizeof(HaloInfo), ncclUint8, m->peer[0], m->comm, s));
      if (m->peer[0] >= 0)
        NCCL_TRY(g_nccl.Send(dev + 0 * sizeof(HaloInfo), sizeof(HaloInfo), ncclUint8, m->peer[0], m->comm, s));
      if (m->peer[1] >= 0)
        NCCL_TRY(g_nccl.Recv(dev + 3 * sizeof(HaloInfo), sizeof(HaloInfo), ncclUint8, m->peer[1], m->comm, s));
      NCCL_TRY(g_nccl.GroupEnd());
      CU_TRY(cudaMemcpyAsync(theirs, dev + 2 * sizeof(HaloInfo), 2 * sizeof(HaloInfo), cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      for (int d = 0; d < 2; ++d)
        {
          for (int g = 0; g < 2; ++g)
            for (int a = 0; a < 3; ++a)
              m->peer_arr[d][g][a] = nullptr;
          if (m->peer[d] < 0)
            continue;
          const HaloInfo &t = theirs[d];
          if (!t.ok)
            {
              ok = false;
              continue;
            }
          for (int g = 0; g < 2; ++g)
            for (int a = 0; a < 3; ++a)
              {
                void *b = map_peer_allocation(m, t.handle[g][a]);
                if (!b)
                  ok = false;
                else
                  m->peer_arr[d][g][a] = reinterpret_cast<double4 *>(static_cast<uint8_t *>(b) + t.offset[g][a]);
              }
          m->peer_base[d] = t.base;
          m->gen_xor[d] = int(t.cur ^ uint32_t(c->cur)) & 1;
        }
      // fused only if it works for every rank (the per-step protocol is collective)
      uint32_t all_ok = ok ? 1u : 0u;
      CU_TRY(cudaMemcpyAsync(m->xcount.p, &all_ok, 4, cudaMemcpyHostToDevice, s));
      NCCL_TRY(g_nccl.AllReduce(m->xcount.p, m->xcount.p, 1, ncclUint32, ncclMin, m->comm, s));
      CU_TRY(cudaMemcpyAsync(&all_ok, m->xcount.p, 4, cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      if (!all_ok && m->rank == 0 && !m->fused_ready && m->epoch == 1)
        fprintf(stderr, "[lethe_dem] CUDA IPC mapping of the neighbours' state arrays failed: halo stays on NCCL send/recv\n");
      m->fused_ready = all_ok != 0;
      // mappings nobody referenced for two exchanges belong to arrays their owner has outgrown;
      // outgrown arrays of mine that were retired two exchanges ago are mapped by nobody any more
      for (size_t k = 0; k < m->mappings.size();)
        if (m->mappings[k].epoch + 2 <= m->epoch)
          {
            cudaIpcCloseMemHandle(m->mappings[k].base);
            m->mappings[k] = m->mappings.back();
            m->mappings.pop_back();
          }
        else
          ++k;
      for (void *q : m->retired)
        m->graveyard.emplace_back(q, m->epoch);
      m->retired.clear();
      for (size_t k = 0; k < m->graveyard.size();)
        if (m->graveyard[k].second + 4 <= m->epoch)
          {
            cudaFree(m->graveyard[k].first);
            m->graveyard[k] = m->graveyard.back();
            m->graveyard.pop_back();
          }
        else
          ++k;
    }
  } // namespace

  bool MultiGpu::fused() const { return impl && impl->fused_ready; }

  const uint32_t *MultiGpu::agreed_flag_dev(lethe_dem_ctx *c) const { return c->flag_dev.p + 1; }

  void MultiGpu::post_agree(lethe_dem_ctx *c, uint32_t host_bits, bool consult)
  {
    MultiGpuImpl *m = impl;
    cudaStream_t s = c->stream;
    m->agreed_slot ^= 1;
    if (m->mailbox_ready)
      launch_agree(m->peer_mailbox.p, m->mailbox, m->rank, m->world, ++m->agree_seq, c->flag_dev.p, host_bits, consult ? 1 : 0,
                   m->agreed_host_dev + m->agreed_slot, s);
    else
      {
        launch_prepare_flag(c->flag_dev.p, host_bits, consult ? 1 : 0, s);
        NCCL_TRY(g_nccl.AllReduce(c->flag_dev.p + 2, c->flag_dev.p + 1, 1, ncclUint32, ncclMax, m->comm, s));
        CU_TRY(cudaMemcpyAsync(m->agreed_host + m->agreed_slot, c->flag_dev.p + 1, 4, cudaMemcpyDeviceToHost, s));
      }
    CU_TRY(cudaEventRecord(m->agreed_ev[m->agreed_slot], s));
  }

  uint32_t MultiGpu::wait_agree(lethe_dem_ctx *c)
  {
    MultiGpuImpl *m = impl;
    (void)c;
    CU_TRY(cudaEventSynchronize(m->agreed_ev[m->agreed_slot]));
    const uint32_t v = *reinterpret_cast<volatile uint32_t *>(m->agreed_host + m->agreed_slot);
    if (v == 0xffffffffu)
      throw std::runtime_error("multi-GPU step agreement timed out: a neighbouring rank did not reach this step");
    return v;
  }

  void MultiGpu::fill_halo(lethe_dem_ctx *c, int out_gen, HaloPush &h) const
  {
    std::memset(&h, 0, sizeof(h));
    const MultiGpuImpl *m = impl;
    if (!m || !m->fused_ready)
      return;
    (void)c;
    for (int d = 0; d < 2; ++d)
      {
        if (m->peer[d] < 0 || m->n_send[d] == 0)
          continue;
        const int pg = (out_gen ^ m->gen_xor[d]) & 1;
        h.pos[d] = m->peer_arr[d][pg][0];
        h.vel[d] = m->peer_arr[d][pg][1];
        h.omg[d] = m->peer_arr[d][pg][2];
        h.bits[d] = m->halo_bits[d].p;
        h.prefix[d] = m->halo_prefix[d].p;
        h.base[d] = m->peer_base[d];
      }
  }

  void MultiGpu::rebuild_with_exchange(lethe_dem_ctx *c)
  {
    MultiGpuImpl *m = impl;
    cudaStream_t s = c->stream;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (c->timers_enabled)
      {
        CU_TRY(cudaEventCreate(&ev0));
        CU_TRY(cudaEventCreate(&ev1));
        CU_TRY(cudaEventRecord(ev0, s));
      }
    // LETHE_DEM_TRACE=1: wall-clock of every phase of this exchange (stream synchronised at each mark)
    static const bool trace = std::getenv("LETHE_DEM_TRACE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    std::string trace_line;
    auto mark = [&](const char *what) {
      if (!trace)
        return;
      cudaStreamSynchronize(s);
      const auto now = std::chrono::steady_clock::now();
      char buf[64];
      snprintf(buf, sizeof buf, " %s %.2f", what, std::chrono::duration<double, std::milli>(now - t_last).count());
      trace_line += buf;
      t_last = now;
    };
    engine_upload_walls(c);
    mark("walls");

    // ---- 0. the id -> slot maps must cover every id of the job (ghosts, immigrants) ----
    {
      uint32_t want = c->slot_map_size;
      CU_TRY(cudaMemcpyAsync(m->xcount.p, &want, 4, cudaMemcpyHostToDevice, s));
      NCCL_TRY(g_nccl.AllReduce(m->xcount.p, m->xcount.p, 1, ncclUint32, ncclMax, m->comm, s));
      CU_TRY(cudaMemcpyAsync(&want, m->xcount.p, 4, cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      if (want > c->slot_map_size)
        {
          const size_t old = c->slot_map_size;
          c->slot_of_id.ensure(want, old, s, 1.0);
          launch_fill_u32(c->slot_of_id.p + old, 0xffffffffu, size_t(want) - old, s);
          c->slot_map_size = want;
        }
    }

    mark("idmap");
    // ---- 0b. load balancing: move the cut planes towards the balanced histogram ----
    // (the slab counterpart of triangulation.repartition() with particle weights, dem.cc:383-457).
    // A cut moves by less than the narrowest slab, so that every particle that changes owner goes to
    // an adjacent rank through the migration below; the contact history travels with it (the
    // reference clears every history at a load-balance step, dem_action_manager.h:223-233).
    int max_hop = 1;
    if (c->lb_recut_pending)
      {
        c->lb_recut_pending = false;
        const int a = c->grid.slab_axis, na = c->grid.n[a];
        DevBuf<uint32_t> hist_dev;
        hist_dev.ensure(size_t(na) + size_t(m->world) + 1);
        CU_TRY(cudaMemsetAsync(hist_dev.p, 0, (size_t(na) + size_t(m->world) + 1) * 4, s));
        if (c->lb_method == LETHE_LOAD_BALANCE_DYNAMIC_WITH_SPARSE_CONTACTS && c->asc_in_force)
          // particle weight x the factor of its cell's mobility status (load_balancing.cc:184-222), in thousandths
          launch_layer_histogram_weighted(c->st[c->cur].pos.p, c->st[c->cur].cell_reg.p, c->asc_cell_status.p, c->grid, c->n_owned, 1000u,
                                          uint32_t(std::lround(1000.0 * c->lb_active_factor)), uint32_t(std::lround(1000.0 * c->lb_inactive_factor)),
                                          hist_dev.p, s);
        else
          launch_layer_histogram(c->st[c->cur].pos.p, c->grid, c->n_owned, hist_dev.p, s);
        // every rank also publishes its lower cut in the slot behind the histogram
        const uint32_t lo = uint32_t(c->grid.slab_lo);
        CU_TRY(cudaMemcpyAsync(hist_dev.p + na + m->rank, &lo, 4, cudaMemcpyHostToDevice, s));
        NCCL_TRY(g_nccl.AllReduce(hist_dev.p, hist_dev.p, size_t(na) + size_t(m->world), ncclUint32, ncclSum, m->comm, s));
        std::vector<uint32_t> h(size_t(na) + size_t(m->world));
        CU_TRY(cudaMemcpyAsync(h.data(), hist_dev.p, h.size() * 4, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        std::vector<uint64_t> hist(h.begin(), h.begin() + na);
        std::vector<int32_t> cuts(size_t(m->world) + 1), new_cuts(size_t(m->world) + 1);
        for (int r = 0; r < m->world; ++r)
          cuts[r] = int32_t(h[size_t(na) + r]);
        cuts[m->world] = na;
        balanced_cuts(na, hist.data(), m->world, cuts.data(), na, 2, new_cuts.data());
        int largest_shift = 0;
        for (int r = 1; r < m->world; ++r)
          largest_shift = std::max(largest_shift, std::abs(new_cuts[r] - cuts[r]));
        c->grid.slab_lo = new_cuts[m->rank];
        c->grid.slab_hi = new_cuts[m->rank + 1];
        max_hop = largest_shift + 1;
        ++c->n_recuts;
      }

    // ---- 1. migration of the particles that left the slab ----
    const uint32_t n0 = c->n_owned;
    const uint32_t cap = std::max<uint32_t>(1024u, n0 / 8 + 1024u);
    for (int d = 0; d < 2; ++d)
      {
        m->send_rec[d].ensure(cap);
        m->send_id[d].ensure(cap);
        m->send_slot[d].ensure(cap);
      }
    CU_TRY(cudaMemsetAsync(m->counters.p, 0, 32, s));
    StateBufs &st = c->st[c->cur];
    ClassifyParams cp;
    cp.pos = st.pos.p;
    cp.vel = st.vel.p;
    cp.omg = st.omg.p;
    cp.id = st.id.p;
    cp.cell_reg = st.cell_reg.p;
    cp.grid = c->grid;
    cp.n = n0;
    for (int d = 0; d < 2; ++d)
      {
        cp.send_rec[d] = m->send_rec[d].p;
        cp.send_id[d] = m->send_id[d].p;
        cp.send_slot[d] = m->send_slot[d].p;
      }
    cp.send_count = m->counters.p;
    cp.send_cap = cap;
    cp.max_hop = max_hop;
    launch_classify(cp, s);
    uint32_t hc[4] = {0, 0, 0, 0};
    CU_TRY(cudaMemcpyAsync(hc, m->counters.p, 16, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    // every rank learns about an overflow on any rank BEFORE the exchanges start, so that all of
    // them fail together instead of one throwing and the others waiting in NCCL for ever
    if (c->multi.agree(c, hc[0] > cap || hc[1] > cap))
      throw std::runtime_error(hc[0] > cap || hc[1] > cap
                                 ? "migration buffer overflow: more than 1/8 of the slab left in one rebuild"
                                 : "migration buffer overflow on another rank");
    if (hc[2])
      fprintf(stderr, "[lethe_dem] rank %d: %u particles jumped over a whole slab and were dropped\n", m->rank, hc[2]);
    uint32_t n_send[2] = {m->peer[0] >= 0 ? hc[0] : 0, m->peer[1] >= 0 ? hc[1] : 0};
    uint32_t n_recv[2] = {0, 0};
    exchange_counts(c, m, n_send, n_recv);
    for (int d = 0; d < 2; ++d)
      {
        m->recv_rec[d].ensure(std::max<uint32_t>(n_recv[d], 1));
        m->recv_id[d].ensure(std::max<uint32_t>(n_recv[d], 1));
      }
    NCCL_TRY(g_nccl.GroupStart());
    for_each_direction_pair(m, [&](int ds, int dr) {
      if (m->peer[ds] >= 0 && n_send[ds])
        {
          NCCL_TRY(g_nccl.Send(m->send_rec[ds].p, size_t(n_send[ds]) * sizeof(MigrateRecord), ncclUint8, m->peer[ds], m->comm, s));
          NCCL_TRY(g_nccl.Send(m->send_id[ds].p, n_send[ds], ncclUint32, m->peer[ds], m->comm, s));
        }
      if (m->peer[dr] >= 0 && n_recv[dr])
        {
          NCCL_TRY(g_nccl.Recv(m->recv_rec[dr].p, size_t(n_recv[dr]) * sizeof(MigrateRecord), ncclUint8, m->peer[dr], m->comm, s));
          NCCL_TRY(g_nccl.Recv(m->recv_id[dr].p, n_recv[dr], ncclUint32, m->peer[dr], m->comm, s));
        }
    });
    NCCL_TRY(g_nccl.GroupEnd());

    mark("migrate");
    // ---- 1b. the contact history of the emigrants travels with them ----
    {
      const bool use_roll = c->cfg.rolling_model == LETHE_ROLLING_EPSD;
      const bool use_img = c->grid.periodic[0] || c->grid.periodic[1] || c->grid.periodic[2];
      ListBufs &l = c->lists[c->cur_list];
      WallListBufs &wl = c->wlists[c->cur_list];
      uint32_t h_send[2] = {0, 0}, h_recv[2] = {0, 0};
      HistPackParams hp[2];
      for (int d = 0; d < 2; ++d)
        {
          if (!n_send[d])
            continue;
          m->hist_offsets[d].ensure(size_t(n_send[d]) + 2);
          c->scan_tmp.ensure(scan_tmp_elems(size_t(n_send[d]) + 8));
          SolidListBufs &sl = c->slists[c->cur_list];
          const bool solids = c->n_solids > 0 && sl.n_rows > 0;
          hp[d] = HistPackParams{m->send_slot[d].p, n_send[d], st.id.p, l.view(), wl.view(), l.n_rows, wl.n_rows,
                                 solids ? sl.row_start.p : nullptr, sl.entry.p, sl.hist.p, sl.roll.p, solids ? sl.n_rows : 0u,
                                 use_roll ? 1 : 0, use_img ? 1 : 0, m->hist_offsets[d].p, nullptr};
          launch_hist_count(hp[d], s);
          exclusive_scan_u32(m->hist_offsets[d].p, m->hist_offsets[d].p, size_t(n_send[d]) + 1, c->scan_tmp.p, s);
          CU_TRY(cudaMemcpyAsync(&h_send[d], m->hist_offsets[d].p + n_send[d], 4, cudaMemcpyDeviceToHost, s));
        }
      CU_TRY(cudaStreamSynchronize(s));
      for (int d = 0; d < 2; ++d)
        if (h_send[d])
          {
            m->send_hist[d].ensure(h_send[d]);
            hp[d].out = m->send_hist[d].p;
            launch_hist_pack(hp[d], s);
          }
      exchange_counts(c, m, h_send, h_recv);
      c->n_pay = h_recv[0] + h_recv[1];
      if (c->n_pay)
        c->pay.ensure(c->n_pay);
      NCCL_TRY(g_nccl.GroupStart());
      for_each_direction_pair(m, [&](int ds, int dr) {
        if (m->peer[ds] >= 0 && h_send[ds])
          NCCL_TRY(g_nccl.Send(m->send_hist[ds].p, size_t(h_send[ds]) * sizeof(HistRecord), ncclUint8, m->peer[ds], m->comm, s));
        if (m->peer[dr] >= 0 && h_recv[dr])
          NCCL_TRY(g_nccl.Recv(c->pay.p + (dr == 1 ? h_recv[0] : 0), size_t(h_recv[dr]) * sizeof(HistRecord), ncclUint8, m->peer[dr],
                               m->comm, s));
      });
      NCCL_TRY(g_nccl.GroupEnd());
      if (c->n_pay)
        {
          c->pay_start.ensure(std::max<size_t>(c->slot_map_size, 1));
          launch_fill_u32(c->pay_start.p, 0xffffffffu, c->slot_map_size, s);
          launch_hist_index(c->pay.p, c->n_pay, c->pay_start.p, c->slot_map_size, s);
        }
    }

    const uint32_t n_in = n_recv[0] + n_recv[1];
    c->n_migrated += uint64_t(n_send[0]) + n_send[1] + n_in;
    c->first_immigrant = n0;
    if (n_in)
      {
        const size_t total = size_t(n0) + n_in;
        c->st[c->cur].ensure(total, n0, s);
        c->st[c->cur ^ 1].ensure(total, 0, s);
        c->disp.ensure(total, n0, s);
        uint32_t base = n0;
        for (int d = 0; d < 2; ++d)
          if (n_recv[d])
            {
              launch_append_records(m->recv_rec[d].p, m->recv_id[d].p, n_recv[d], c->st[c->cur].view(), c->st[c->cur].id.p,
                                    c->st[c->cur].cell_reg.p, c->disp.p, base, s);
              base += n_recv[d];
            }
        c->n_owned = uint32_t(total);
      }

    mark("history");
    // ---- 2. local sort (drops the particles sent away) ----
    engine_rebuild_sort(c);
    c->first_immigrant = 0xffffffffu;

    mark("sort");
    // ---- 3. ghost exchange: my boundary cell layers -> neighbours ----
    const uint32_t n = c->n_owned;
    StateBufs &sn = c->st[c->cur];
    m->flags.ensure(size_t(n) + 2);
    m->offsets.ensure(size_t(n) + 2);
    c->scan_tmp.ensure(scan_tmp_elems(size_t(n) + 8));
    const int layer[2] = {c->grid.slab_lo, c->grid.slab_hi - 1};
    for (int d = 0; d < 2; ++d)
      {
        m->n_send[d] = 0;
        if (m->peer[d] < 0)
          continue;
        launch_flag_layer(sn.cell_reg.p, c->grid, layer[d], n, m->flags.p, s);
        exclusive_scan_u32(m->flags.p, m->offsets.p, size_t(n) + 1, c->scan_tmp.p, s);
        uint32_t cnt = 0;
        CU_TRY(cudaMemcpyAsync(&cnt, m->offsets.p + n, 4, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        m->n_send[d] = cnt;
        m->send_idx[d].ensure(std::max<uint32_t>(cnt, 1));
        m->send_ids[d].ensure(std::max<uint32_t>(cnt, 1));
        m->send_pos[d].ensure(std::max<uint32_t>(cnt, 1));
        m->send_vel[d].ensure(std::max<uint32_t>(cnt, 1));
        m->send_omg[d].ensure(std::max<uint32_t>(cnt, 1));
        launch_compact_indices(m->flags.p, m->offsets.p, n, m->send_idx[d].p, s);
        launch_gather_ids(sn.id.p, m->send_idx[d].p, cnt, m->send_ids[d].p, s);
        if (m->want_fused)
          {
            m->halo_bits[d].ensure(size_t(n) / 32 + 2);
            m->halo_prefix[d].ensure(size_t(n) / 32 + 2);
            launch_halo_warp_table(m->flags.p, m->offsets.p, n, m->halo_bits[d].p, m->halo_prefix[d].p, s);
          }
      }
    uint32_t g_recv[2] = {0, 0};
    exchange_counts(c, m, m->n_send, g_recv);
    c->n_ghost_run[0] = g_recv[0];
    c->n_ghost_run[1] = g_recv[1];
    c->n_ghost = g_recv[0] + g_recv[1];
    const size_t total = size_t(n) + c->n_ghost;
    c->st[c->cur].ensure(total, n, s);
    c->st[c->cur ^ 1].ensure(total, n, s);
    c->old_of_new.ensure(std::max<size_t>(total, 1), n, s);
    // ids of the ghosts
    NCCL_TRY(g_nccl.GroupStart());
    for_each_direction_pair(m, [&](int ds, int dr) {
      if (m->peer[ds] >= 0 && m->n_send[ds])
        NCCL_TRY(g_nccl.Send(m->send_ids[ds].p, m->n_send[ds], ncclUint32, m->peer[ds], m->comm, s));
      if (m->peer[dr] >= 0 && g_recv[dr])
        NCCL_TRY(g_nccl.Recv(c->st[c->cur].id.p + n + (dr == 1 ? g_recv[0] : 0), g_recv[dr], ncclUint32, m->peer[dr], m->comm, s));
    });
    NCCL_TRY(g_nccl.GroupEnd());
    if (m->want_fused)
      exchange_halo_info(c, m, g_recv); // after the last (re)allocation of the state arrays in this rebuild
    refresh_ghosts(c); // positions / velocities of the new ghost set

    mark("ghosts");
    // ---- 4. ghost cell tables + history source ----
    const size_t n_cells = size_t(c->grid.n_cells);
    StateBufs &sg = c->st[c->cur];
    // ghost ids must be addressable in the id map
    for (int d = 0; d < 2; ++d)
      {
        if (!g_recv[d])
          continue;
        c->ghost_start[d].ensure(n_cells + 1);
        c->ghost_end[d].ensure(n_cells + 1);
        CU_TRY(cudaMemsetAsync(c->ghost_start[d].p, 0, (n_cells + 1) * 4, s));
        CU_TRY(cudaMemsetAsync(c->ghost_end[d].p, 0, (n_cells + 1) * 4, s));
        GhostRunParams gp;
        const uint32_t base = n + (d == 1 ? g_recv[0] : 0);
        gp.pos = sg.pos.p + base;
        gp.grid = c->grid;
        gp.cell_rank = c->cell_rank.p;
        gp.base = base;
        gp.n = g_recv[d];
        gp.cell_reg = sg.cell_reg.p;
        gp.start = c->ghost_start[d].p;
        gp.end = c->ghost_end[d].p;
        gp.id = sg.id.p;
        gp.old_slot_of_id = c->slot_map_size_old ? c->slot_of_id_old.p : nullptr;
        gp.old_map_size = c->slot_map_size_old;
        gp.old_n_owned = c->old_n_owned;
        gp.old_of_new = c->old_of_new.p;
        launch_ghost_run(gp, s);
      }
    if (c->n_ghost)
      launch_register_ids(sg.id.p, n, c->n_ghost, c->slot_of_id.p, c->slot_map_size, s);

    mark("tables");
    // ---- 4b. adaptive sparse contacts: mobility status of my cells, node values merged across the cuts ----
    engine_identify_mobility_status(c);
    // ---- 5. lists ----
    engine_rebuild_lists(c);
    c->n_pay = 0;
    engine_mirror_ids(c);
    mark("lists");
    if (trace)
      fprintf(stderr, "[lethe_dem trace] rank %d it %llu rebuild n_owned %u n_ghost %u fused %d mailbox %d ms:%s\n", m->rank,
              (unsigned long long)c->iteration_number, c->n_owned, c->n_ghost, int(m->fused_ready), int(m->mailbox_ready), trace_line.c_str());
    if (c->timers_enabled)
      {
        CU_TRY(cudaEventRecord(ev1, s));
        c->pending_rebuild.emplace_back(ev0, ev1);
        ++c->rebuild_launches;
      }
  }

  // update_ghost_particles: my boundary particles' state goes straight into the neighbours'
  // ghost runs (three arrays, no unpack kernel on the receiving side).
  void MultiGpu::refresh_ghosts(lethe_dem_ctx *c)
  {
    MultiGpuImpl *m = impl;
    cudaStream_t s = c->stream;
    StateBufs &st = c->st[c->cur];
    const uint32_t n = c->n_owned;
    for (int d = 0; d < 2; ++d)
      if (m->peer[d] >= 0 && m->n_send[d])
        launch_gather_state(st.view(), m->send_idx[d].p, m->n_send[d], m->send_pos[d].p, m->send_vel[d].p, m->send_omg[d].p, s);
    NCCL_TRY(g_nccl.GroupStart());
    for_each_direction_pair(m, [&](int ds, int dr) {
      if (m->peer[ds] >= 0 && m->n_send[ds])
        {
          const size_t cnt = size_t(m->n_send[ds]) * 4;
          NCCL_TRY(g_nccl.Send(m->send_pos[ds].p, cnt, ncclDouble, m->peer[ds], m->comm, s));
          NCCL_TRY(g_nccl.Send(m->send_vel[ds].p, cnt, ncclDouble, m->peer[ds], m->comm, s));
          NCCL_TRY(g_nccl.Send(m->send_omg[ds].p, cnt, ncclDouble, m->peer[ds], m->comm, s));
        }
      if (m->peer[dr] >= 0 && c->n_ghost_run[dr])
        {
          const size_t base = size_t(n) + (dr == 1 ? c->n_ghost_run[0] : 0);
          const size_t cnt = size_t(c->n_ghost_run[dr]) * 4;
          NCCL_TRY(g_nccl.Recv(st.pos.p + base, cnt, ncclDouble, m->peer[dr], m->comm, s));
          NCCL_TRY(g_nccl.Recv(st.vel.p + base, cnt, ncclDouble, m->peer[dr], m->comm, s));
          NCCL_TRY(g_nccl.Recv(st.omg.p + base, cnt, ncclDouble, m->peer[dr], m->comm, s));
        }
    });
    NCCL_TRY(g_nccl.GroupEnd());
  }
} // namespace dem
