// dem_engine.cu — host side of the B200 DEM engine: context, buffers, the per-step
// control flow of DEMSolver::solve and the C ABI of include/lethe_dem.h.
//
// Control flow mirrored (source/dem/dem.cc): one lethe_dem_step iteration ==
//   simulation_control->integrate()            (iteration_number++, time += dt)
//   execute_contact_detection_and_search()     (:598-688)  -> rebuild() when triggered
//   compute_contact_forces() + integrate()     (:690-717,1143-1181) -> ONE fused kernel
//   action_manager->reset_triggers()           (:1246)
// No CPU fallback exists: without a CUDA device lethe_dem_create fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "dem_context.cuh"

using namespace dem;

namespace
{
  std::string g_create_error;
}

namespace
{
  using Ctx = lethe_dem_ctx;

  inline double harmonic_mean(double a, double b) { return (2 * a * b / (a + b + DBL_MIN)); }

  // set_effective_properties (particle_particle_contact_force.h:1639-1746,
  // particle_wall_contact_force.cc:588-694), evaluated on the host with glibc
  void build_material_tables(const lethe_dem_config &c, MaterialTables &mt)
  {
    std::memset(&mt, 0, sizeof(mt));
    const int n = c.n_types;
    mt.n_types = n;
    for (int i = 0; i < n; ++i)
      {
        const double Yi = c.young[i], nui = c.poisson[i];
        for (int j = 0; j < n; ++j)
          {
            const int k = i * n + j;
            const double Yj = c.young[j], nuj = c.poisson[j];
            mt.Y[k] = (Yi * Yj) / ((Yj * (1.0 - nui * nui)) + (Yi * (1.0 - nuj * nuj)) + DBL_MIN);
            mt.G[k] = (Yi * Yj) / (2.0 * ((Yj * (2.0 - nui) * (1.0 + nui)) + (Yi * (2.0 - nuj) * (1.0 + nuj))) + DBL_MIN);
            const double rest = harmonic_mean(c.restitution[i], c.restitution[j]);
            mt.mu[k] = harmonic_mean(c.friction[i], c.friction[j]);
            mt.roll_visc[k] = harmonic_mean(c.rolling_viscous_damping[i], c.rolling_viscous_damping[j]);
            mt.roll_fric[k] = harmonic_mean(c.rolling_friction[i], c.rolling_friction[j]);
            mt.gamma[k] = c.surface_energy[i] + c.surface_energy[j] -
                          std::pow(std::sqrt(c.surface_energy[i]) - std::sqrt(c.surface_energy[j]), 2);
            mt.hamaker[k] = 0.5 * (c.hamaker[i] + c.hamaker[j]);
            const double lg = std::log(rest);
            mt.beta[k] = lg / std::sqrt(lg * lg + 9.8696);
          }
        const double Yw = c.young_wall, nuw = c.poisson_wall;
        mt.wY[i] = (Yi * Yw) / (Yw * (1. - nui * nui) + Yi * (1. - nuw * nuw) + DBL_MIN);
        mt.wG[i] = (Yi * Yw) / ((2. * Yw * (2. - nui) * (1. + nui)) + (2. * Yi * (2. - nuw) * (1. + nuw)) + DBL_MIN);
        const double rest = harmonic_mean(c.restitution[i], c.restitution_wall);
        mt.wmu[i] = harmonic_mean(c.friction[i], c.friction_wall);
        mt.wroll_fric[i] = harmonic_mean(c.rolling_friction[i], c.rolling_friction_wall);
        mt.wroll_visc[i] = harmonic_mean(c.rolling_viscous_damping[i], c.rolling_viscous_damping_wall);
        mt.wgamma[i] = c.surface_energy[i] + c.surface_energy_wall -
                       std::pow(std::sqrt(c.surface_energy[i]) - std::sqrt(c.surface_energy_wall), 2);
        mt.whamaker[i] = 0.5 * (c.hamaker[i] + c.hamaker_wall);
        const double lg = std::log(rest);
        mt.wbeta[i] = lg / std::sqrt((lg * lg) + 9.8696);
      }
    // get_force_calculation_threshold_distance (…force.h:504-529)
    mt.pp_force_threshold = 0.;
    if (c.pp_model == LETHE_PP_DMT)
      {
        const double maxA = *std::max_element(mt.hamaker, mt.hamaker + n * n);
        const double minG = *std::min_element(mt.gamma, mt.gamma + n * n);
        mt.pp_force_threshold = -std::sqrt(maxA / (12. * M_PI * minG * c.dmt_cut_off_threshold));
      }
    mt.pw_force_threshold = 0.;
    if (c.pw_model == LETHE_PW_DMT)
      {
        const double maxA = *std::max_element(mt.whamaker, mt.whamaker + n);
        const double minG = *std::min_element(mt.wgamma, mt.wgamma + n);
        mt.pw_force_threshold = -std::sqrt(maxA / (12. * M_PI * minG * c.dmt_cut_off_threshold));
      }
    mt.f_coefficient_epsd = c.f_coefficient_epsd;
  }

  // Morton (z-order) rank of every grid cell: the sort key of the particles.
  void build_cell_curve(const GridDesc &g, std::vector<int32_t> &rank_of_cell, std::vector<int32_t> &cell_of_rank)
  {
    const int nmax = std::max(g.n[0], std::max(g.n[1], g.n[2]));
    int bits = 0;
    while ((1 << bits) < nmax)
      ++bits;
    std::vector<std::pair<uint64_t, int32_t>> codes;
    codes.reserve(g.n_cells);
    for (int k = 0; k < g.n[2]; ++k)
      for (int j = 0; j < g.n[1]; ++j)
        for (int i = 0; i < g.n[0]; ++i)
          {
            uint64_t m = 0;
            for (int b = 0; b < bits; ++b)
              {
                m |= uint64_t((i >> b) & 1) << (3 * b);
                m |= uint64_t((j >> b) & 1) << (3 * b + 1);
                m |= uint64_t((k >> b) & 1) << (3 * b + 2);
              }
            codes.emplace_back(m, i + g.n[0] * (j + g.n[1] * k));
          }
    std::sort(codes.begin(), codes.end());
    rank_of_cell.assign(g.n_cells, 0);
    cell_of_rank.assign(g.n_cells, 0);
    for (int r = 0; r < g.n_cells; ++r)
      {
        cell_of_rank[r] = codes[r].second;
        rank_of_cell[codes[r].second] = r;
      }
  }

  cudaEvent_t get_event(Ctx *c)
  {
    if (!c->event_pool.empty())
      {
        cudaEvent_t e = c->event_pool.back();
        c->event_pool.pop_back();
        return e;
      }
    cudaEvent_t e;
    CU_TRY(cudaEventCreate(&e));
    return e;
  }
  void resolve_timers(Ctx *c)
  {
    if (c->pending_step.empty() && c->pending_rebuild.empty())
      return;
    CU_TRY(cudaStreamSynchronize(c->stream));
    for (auto &pr : c->pending_step)
      {
        float ms = 0;
        CU_TRY(cudaEventElapsedTime(&ms, pr.first, pr.second));
        c->step_ms += ms;
        c->event_pool.push_back(pr.first);
        c->event_pool.push_back(pr.second);
      }
    for (auto &pr : c->pending_rebuild)
      {
        float ms = 0;
        CU_TRY(cudaEventElapsedTime(&ms, pr.first, pr.second));
        c->rebuild_ms += ms;
        c->event_pool.push_back(pr.first);
        c->event_pool.push_back(pr.second);
      }
    c->pending_step.clear();
    c->pending_rebuild.clear();
  }

  void upload_walls(Ctx *c)
  {
    if (!c->walls_dirty)
      return;
    const int n_cells = c->grid.n_cells;
    // face table sorted by cell (stable: keeps the host's order inside a cell)
    std::vector<lethe_wall_face> f = c->faces_host;
    std::stable_sort(f.begin(), f.end(), [](const lethe_wall_face &a, const lethe_wall_face &b) { return a.cell < b.cell; });
    c->n_faces = uint32_t(f.size());
    std::vector<uint32_t> start(size_t(n_cells) + 1, 0);
    for (auto &face : f)
      start[size_t(face.cell) + 1]++;
    for (int k = 0; k < n_cells; ++k)
      start[k + 1] += start[k];
    std::vector<double> nrm(3 * f.size() + 3), pt(3 * f.size() + 3);
    std::vector<uint32_t> bid(f.size() + 1);
    std::vector<int32_t> mot(f.size() + 1, -1);
    c->face_gid_host.assign(f.size(), 0);
    for (size_t k = 0; k < f.size(); ++k)
      {
        for (int d = 0; d < 3; ++d)
          {
            nrm[3 * k + d] = f[k].normal[d];
            pt[3 * k + d] = f[k].point[d];
          }
        bid[k] = f[k].boundary_id;
        c->face_gid_host[k] = f[k].global_face_id;
        for (size_t m = 0; m < c->motions_host.size(); ++m)
          if (c->motions_host[m].boundary_id == f[k].boundary_id)
            mot[k] = int32_t(m);
      }
    c->cell_face_start.ensure(start.size());
    c->face_normal.ensure(nrm.size());
    c->face_point.ensure(pt.size());
    c->face_boundary.ensure(bid.size());
    c->face_motion.ensure(mot.size());
    CU_TRY(cudaMemcpyAsync(c->cell_face_start.p, start.data(), start.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(c->face_normal.p, nrm.data(), nrm.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(c->face_point.p, pt.data(), pt.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(c->face_boundary.p, bid.data(), bid.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(c->face_motion.p, mot.data(), mot.size() * 4, cudaMemcpyHostToDevice, c->stream));
    std::vector<BoundaryMotionDev> md(std::max<size_t>(1, c->motions_host.size()));
    for (size_t m = 0; m < c->motions_host.size(); ++m)
      md[m] = c->motions_host[m].m;
    c->motions.ensure(md.size());
    CU_TRY(cudaMemcpyAsync(c->motions.p, md.data(), md.size() * sizeof(BoundaryMotionDev), cudaMemcpyHostToDevice, c->stream));
    // floating walls: boundary cells = cells with a vertex closer than the maximal cell
    // diameter to the plane (find_boundary_cells_information.cc:653-703)
    c->fw_dev.ensure(1);
    CU_TRY(cudaMemcpyAsync(c->fw_dev.p, &c->fw_host, sizeof(FloatingWallsDev), cudaMemcpyHostToDevice, c->stream));
    if (c->fw_host.n > 0)
      {
        std::vector<uint32_t> mask(n_cells, 0);
        const GridDesc &g = c->grid;
        const double maxd = std::sqrt(g.h[0] * g.h[0] + g.h[1] * g.h[1] + g.h[2] * g.h[2]);
        for (int w = 0; w < c->fw_host.n; ++w)
          for (int cell = 0; cell < n_cells; ++cell)
            {
              const int ci[3] = {cell % g.n[0], (cell / g.n[0]) % g.n[1], cell / (g.n[0] * g.n[1])};
              for (int v = 0; v < 8; ++v)
                {
                  const double vx[3] = {g.lo[0] + (ci[0] + (v & 1)) * g.h[0], g.lo[1] + (ci[1] + ((v >> 1) & 1)) * g.h[1],
                                        g.lo[2] + (ci[2] + ((v >> 2) & 1)) * g.h[2]};
                  const double cv[3] = {vx[0] - c->fw_host.point[w][0], vx[1] - c->fw_host.point[w][1],
                                        vx[2] - c->fw_host.point[w][2]};
                  const double dist =
                    (cv[0] * c->fw_host.normal[w][0] + cv[1] * c->fw_host.normal[w][1]) + cv[2] * c->fw_host.normal[w][2];
                  if (std::fabs(dist) < maxd)
                    {
                      mask[cell] |= 1u << w;
                      break;
                    }
                }
            }
        c->cell_fw_mask.ensure(mask.size());
        CU_TRY(cudaMemcpyAsync(c->cell_fw_mask.p, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice, c->stream));
      }
    CU_TRY(cudaStreamSynchronize(c->stream));
    c->walls_dirty = false;
  }

  FaceTable face_table(Ctx *c)
  {
    return FaceTable{c->cell_face_start.p, c->face_normal.p, c->face_point.p, c->face_boundary.p, c->face_motion.p, c->n_faces};
  }

  uint32_t read_u32(Ctx *c, const uint32_t *dptr)
  {
    uint32_t v = 0;
    CU_TRY(cudaMemcpyAsync(&v, dptr, 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return v;
  }


  // ------------------------------------------------------ solid surfaces ----
  SolidSetView solid_view(Ctx *c)
  {
    return SolidSetView{c->solid_vertices.p, c->solid_disp.p, c->solid_vertex_solid.p, c->solid_tri.p, c->solid_tri_solid.p,
                        c->solid_es_start.p, c->solid_es_idx.p, c->solid_vs_start.p, c->solid_vs_idx.p, c->solid_motion.p,
                        uint32_t(c->solid_vertex_solid_host.size()), uint32_t(c->solid_tri_solid_host.size()), c->n_solids};
  }

  // Uploads topology (and the geometry of solids added since the last upload; solids already on
  // the device keep their current, moved vertices). SerialSolid::setup_containers
  // (serial_solid.cc:617-682) gives the edge- / vertex-sharing neighbour tables.
  void upload_solids(Ctx *c)
  {
    if (!c->solids_dirty)
      return;
    cudaStream_t s = c->stream;
    const uint32_t nv = uint32_t(c->solid_vertex_solid_host.size()), nt = uint32_t(c->solid_tri_solid_host.size());
    const uint32_t nv_old = c->n_solid_vertices_dev;
    c->solid_vertices.ensure(3 * size_t(nv), 3 * size_t(nv_old), s);
    c->solid_disp.ensure(3 * size_t(nv), 3 * size_t(nv_old), s);
    if (nv > nv_old)
      {
        CU_TRY(cudaMemcpyAsync(c->solid_vertices.p + 3 * size_t(nv_old), c->solid_vertices_host.data() + 3 * size_t(nv_old),
                               3 * size_t(nv - nv_old) * 8, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemsetAsync(c->solid_disp.p + 3 * size_t(nv_old), 0, 3 * size_t(nv - nv_old) * 8, s));
      }
    // motion records: new solids only (the centres of the old ones have moved on the device)
    const uint32_t ns_old = nv_old ? c->solid_vertex_solid_host[nv_old - 1] + 1 : 0;
    c->solid_motion.ensure(MAX_SOLIDS);
    CU_TRY(cudaMemcpyAsync(c->solid_motion.p + ns_old, c->solid_motion_host.data() + ns_old,
                           size_t(c->n_solids - ns_old) * sizeof(SolidMotionDev), cudaMemcpyHostToDevice, s));
    // neighbour tables over the global triangle index
    std::vector<std::vector<uint32_t>> cells_of_vertex(nv);
    for (uint32_t t = 0; t < nt; ++t)
      for (int k = 0; k < 3; ++k)
        cells_of_vertex[c->solid_tri_host[3 * size_t(t) + k]].push_back(t);
    std::vector<uint32_t> es_start(size_t(nt) + 1, 0), vs_start(size_t(nt) + 1, 0), es_idx, vs_idx;
    for (uint32_t t = 0; t < nt; ++t)
      {
        std::vector<uint32_t> around;
        for (int k = 0; k < 3; ++k)
          {
            const auto &l = cells_of_vertex[c->solid_tri_host[3 * size_t(t) + k]];
            around.insert(around.end(), l.begin(), l.end());
          }
        std::sort(around.begin(), around.end());
        around.erase(std::unique(around.begin(), around.end()), around.end());
        for (uint32_t n : around)
          {
            if (n == t)
              continue;
            int sharing = 0;
            for (int k = 0; k < 3; ++k)
              for (int m = 0; m < 3; ++m)
                if (c->solid_tri_host[3 * size_t(n) + k] == c->solid_tri_host[3 * size_t(t) + m])
                  {
                    ++sharing;
                    break;
                  }
            (sharing == 1 ? vs_idx : es_idx).push_back(n);
          }
        es_start[size_t(t) + 1] = uint32_t(es_idx.size());
        vs_start[size_t(t) + 1] = uint32_t(vs_idx.size());
      }
    auto up = [&](DevBuf<uint32_t> &d, const std::vector<uint32_t> &h) {
      d.ensure(std::max<size_t>(h.size(), 1));
      if (!h.empty())
        CU_TRY(cudaMemcpyAsync(d.p, h.data(), h.size() * 4, cudaMemcpyHostToDevice, s));
    };
    up(c->solid_vertex_solid, c->solid_vertex_solid_host);
    up(c->solid_tri, c->solid_tri_host);
    up(c->solid_tri_solid, c->solid_tri_solid_host);
    up(c->solid_es_start, es_start);
    up(c->solid_es_idx, es_idx);
    up(c->solid_vs_start, vs_start);
    up(c->solid_vs_idx, vs_idx);
    c->solid_overflow.ensure(1);
    CU_TRY(cudaMemsetAsync(c->solid_overflow.p, 0, 4, s));
    CU_TRY(cudaStreamSynchronize(s)); // the host vectors above go out of scope
    c->n_solid_vertices_dev = nv;
    c->solids_dirty = false;
  }

  // Rebuild-time part: (re)map the solids onto the background grid when asked
  // (find_floating_mesh_mapping_step / the first search), then every particle's candidate row.
  void rebuild_solid_lists(Ctx *c)
  {
    if (!c->n_solids)
      return;
    cudaStream_t s = c->stream;
    upload_solids(c);
    const bool use_roll = c->cfg.rolling_model == LETHE_ROLLING_EPSD;
    const uint32_t n_new = c->n_owned;
    CU_TRY(cudaStreamSynchronize(s));
    if (c->solid_map_needed || *c->h_remap)
      {
        const size_t nv = c->solid_vertex_solid_host.size();
        std::vector<double> v(3 * nv);
        CU_TRY(cudaMemcpyAsync(v.data(), c->solid_vertices.p, v.size() * 8, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        std::vector<uint32_t> start, tri;
        map_solids_on_host(c->grid, v.data(), c->solid_tri_host.data(), uint32_t(c->solid_tri_solid_host.size()), start, tri);
        c->cell_tri_start.ensure(start.size());
        c->cell_tri.ensure(std::max<size_t>(tri.size(), 1));
        CU_TRY(cudaMemcpyAsync(c->cell_tri_start.p, start.data(), start.size() * 4, cudaMemcpyHostToDevice, s));
        if (!tri.empty())
          CU_TRY(cudaMemcpyAsync(c->cell_tri.p, tri.data(), tri.size() * 4, cudaMemcpyHostToDevice, s));
        // reset_displacement_monitoring
        CU_TRY(cudaMemsetAsync(c->solid_disp.p, 0, 3 * nv * 8, s));
        CU_TRY(cudaStreamSynchronize(s));
        *c->h_remap = 0;
        c->solid_map_needed = false;
      }
    // cur_list has already been flipped by rebuild_lists: the old generation is cur_list ^ 1
    SolidListBufs &olds = c->slists[c->cur_list ^ 1];
    SolidListBufs &news = c->slists[c->cur_list];
    c->counts.ensure(size_t(n_new) + 2);
    c->scan_tmp.ensure(scan_tmp_elems(size_t(n_new) + 8));
    news.row_start.ensure(size_t(n_new) + 2);
    SolidBuildParams bp;
    bp.cell_reg = c->st[c->cur].cell_reg.p;
    bp.cell_tri_start = c->cell_tri_start.p;
    bp.cell_tri = c->cell_tri.p;
    bp.n_rows = n_new;
    bp.old_list = olds.view();
    bp.old_of_new = c->old_of_new.p;
    bp.n_old_rows = olds.n_rows;
    bp.mobility = c->asc_in_force ? c->asc_cell_status.p : nullptr;
    bp.clear_history = c->clear_history_trigger ? 1 : 0;
    bp.new_list = news.view();
    bp.counts = c->counts.p;
    bp.use_roll = use_roll ? 1 : 0;
    bp.pay = HistPayload{c->n_pay ? c->pay.p : nullptr, c->pay_start.p, c->n_pay, c->slot_map_size, c->st[c->cur].id.p};
    // a sphere touched more triangles than k_solid_contacts keeps for the elimination
    if (read_u32(c, c->solid_overflow.p))
      throw std::runtime_error("a particle touches more than 24 triangles of a solid surface at once: the mesh is too fine for "
                               "the particle size (SOLID_MAX_CONTACTS)");
    launch_count_solid_rows(bp, s);
    exclusive_scan_u32(c->counts.p, news.row_start.p, size_t(n_new) + 1, c->scan_tmp.p, s);
    const uint32_t n_entries = n_new ? read_u32(c, news.row_start.p + n_new) : 0;
    news.entry.ensure(std::max<size_t>(n_entries, 1));
    news.hist.ensure(std::max<size_t>(3 * size_t(n_entries), 1));
    if (use_roll)
      news.roll.ensure(std::max<size_t>(3 * size_t(n_entries), 1));
    bp.new_list = news.view();
    launch_fill_solid_rows(bp, s); // leaves counts[q] = row not empty
    news.n_rows = n_new;
    news.n_entries = n_entries;
    // compact set of particles that have candidates
    c->key.ensure(size_t(n_new) + 2);
    exclusive_scan_u32(c->counts.p, c->key.p, size_t(n_new) + 1, c->scan_tmp.p, s);
    c->n_solid_active = n_new ? read_u32(c, c->key.p + n_new) : 0;
    c->solid_active.ensure(std::max<size_t>(c->n_solid_active, 1));
    launch_compact_indices(c->counts.p, c->key.p, n_new, c->solid_active.p, s);
    c->solid_force.ensure(std::max<size_t>(3 * size_t(n_new), 1));
    c->solid_torque.ensure(std::max<size_t>(3 * size_t(n_new), 1));
    CU_TRY(cudaMemsetAsync(c->solid_force.p, 0, 3 * size_t(n_new) * 8, s));
    CU_TRY(cudaMemsetAsync(c->solid_torque.p, 0, 3 * size_t(n_new) * 8, s));
  }

  // ------------------------------------------------------------ rebuild ----
  // Phase 1: periodic wrap, binning, counting sort along the Morton curve, permutation of the
  // owned particles into cell order (drops particles that left the domain or migrated away).
  void rebuild_sort(Ctx *c)
  {
    cudaStream_t s = c->stream;
    const int n_cells = c->grid.n_cells;
    uint32_t n = c->n_owned;
    StateBufs &src = c->st[c->cur];
    StateBufs &dst = c->st[c->cur ^ 1];
    c->key.ensure(n + 1);
    c->slot.ensure(n + 1);
    c->perm.ensure(n + 1);
    c->cell_count.ensure(size_t(n_cells) + 8);
    c->cell_start.ensure(size_t(n_cells) + 8);
    c->scan_tmp.ensure(scan_tmp_elems(std::max<size_t>(size_t(n_cells) + 8, size_t(n) + 8)));
    CU_TRY(cudaMemsetAsync(c->cell_count.p, 0, (size_t(n_cells) + 8) * 4, s));
    BinParams bp{src.pos.p, src.cell_reg.p, c->grid, c->cell_rank.p, c->cell_count.p, c->key.p, c->slot.p, n};
    launch_bin(bp, s);
    // buckets: [0,n_cells) cells in curve order, n_cells = left the domain / migrated
    const size_t n_buckets = size_t(n_cells) + 3;
    exclusive_scan_u32(c->cell_count.p, c->cell_start.p, n_buckets + 1, c->scan_tmp.p, s);
    launch_scatter_perm(c->key.p, c->slot.p, c->cell_start.p, c->perm.p, n, s);
    // real cells only: the order inside the bucket of dropped particles (rank n_cells) is never used
    launch_sort_cells(c->cell_start.p, uint32_t(n_cells), c->perm.p, src.id.p, s);
    const uint32_t n_new = read_u32(c, c->cell_start.p + n_cells); // particles still inside the domain

    dst.ensure(std::max<size_t>(n_new, 1), 0, s);
    c->old_of_new.ensure(std::max<size_t>(n_new, 1));
    c->disp.ensure(std::max<size_t>(n_new, 1), 0, s);
    // the id map of the outgoing list generation becomes the "old" one
    std::swap(c->slot_of_id.p, c->slot_of_id_old.p);
    std::swap(c->slot_of_id.cap, c->slot_of_id_old.cap);
    c->slot_map_size_old = c->slot_map_size;
    c->slot_of_id.ensure(std::max<size_t>(c->slot_map_size, 1));
    if (c->slot_map_size)
      launch_fill_u32(c->slot_of_id.p, 0xffffffffu, c->slot_map_size, s);
    GatherParams gp{src.view(), dst.view(), src.id.p,       dst.id.p,       c->perm.p, c->key.p,
                    c->cell_of_rank.p, dst.cell_reg.p,      c->old_of_new.p, c->disp.p, c->slot_of_id.p, n_new,
                    c->first_immigrant, c->lists[c->cur_list].n_rows, c->slot_map_size_old ? c->slot_of_id_old.p : nullptr,
                    c->slot_map_size_old};
    launch_gather(gp, s);
    c->cur ^= 1;
    c->old_n_owned = c->lists[c->cur_list].n_rows;
    c->n_owned = n_new;
    c->n_ghost = 0;
    c->n_ghost_run[0] = c->n_ghost_run[1] = 0;
    // the other generation must be able to hold the step kernel's output
    c->st[c->cur ^ 1].ensure(std::max<size_t>(n_new, 1), 0, s);
  }

  // AdaptiveSparseContacts::identify_mobility_status at every contact search (dem.cc:639-644)
  void identify_mobility_status(Ctx *c)
  {
    c->asc_in_force = false;
    if (!c->asc_enabled)
      return;
    cudaStream_t s = c->stream;
    const size_t n_cells = size_t(c->grid.n_cells);
    c->asc_cell_status.ensure(n_cells);
    if (c->asc_reset)
      {
        // first iteration: every cell mobile, the regular searches and integration
        CU_TRY(cudaMemsetAsync(c->asc_cell_status.p, LETHE_MOBILITY_MOBILE, n_cells, s));
        return;
      }
    const size_t n_nodes = size_t(c->grid.n[0] + 1) * (c->grid.n[1] + 1) * (c->grid.n[2] + 1);
    c->asc_node_status.ensure(n_nodes);
    c->asc_row_mobile.ensure(std::max<size_t>(c->n_owned, 1));
    CU_TRY(cudaMemsetAsync(c->asc_node_status.p, 0, n_nodes * sizeof(int), s));
    StateBufs &stn = c->st[c->cur];
    AscParams ap;
    ap.st = stn.view();
    ap.cell_start = c->cell_start.p;
    ap.cell_rank = c->cell_rank.p;
    ap.cell_reg = stn.cell_reg.p;
    ap.grid = c->grid;
    ap.granular_temperature_threshold = c->cfg.asc_granular_temperature_threshold;
    ap.solid_fraction_threshold = c->cfg.asc_solid_fraction_threshold;
    ap.cell_status = c->asc_cell_status.p;
    ap.node_status = c->asc_node_status.p;
    ap.row_mobile = c->asc_row_mobile.p;
    ap.n_rows = c->n_owned;
    const bool slabs = c->multi.enabled();
    ap.owned_lo = slabs ? c->grid.slab_lo : -1;
    ap.owned_hi = slabs ? c->grid.slab_hi : -1;
    for (int pass = 0; pass < 5; ++pass)
      {
        launch_asc_pass(ap, pass, s);
        if (slabs && pass < 3)
          c->multi.asc_exchange_nodes(c); // mobility_at_nodes.update_ghost_values()
        if (slabs && pass == 3)
          c->multi.asc_exchange_cells(c);
      }
    c->asc_in_force = true;
  }

  // Phase 2: contact lists (particle-particle with history carry-over, particle-wall).
  void rebuild_lists(Ctx *c)
  {
    cudaStream_t s = c->stream;
    const bool use_roll = c->cfg.rolling_model == LETHE_ROLLING_EPSD;
    const bool use_img = c->grid.periodic[0] || c->grid.periodic[1] || c->grid.periodic[2];
    const uint32_t n_new = c->n_owned;
    StateBufs &stn = c->st[c->cur];

    ListBufs &oldl = c->lists[c->cur_list];
    ListBufs &newl = c->lists[c->cur_list ^ 1];
    c->counts.ensure(size_t(n_new) + 2);
    c->scan_tmp.ensure(scan_tmp_elems(size_t(n_new) + 8));
    newl.row_start.ensure(size_t(n_new) + 2);
    NeighborParams np;
    np.st = stn.view();
    np.cell_reg = stn.cell_reg.p;
    np.cell_rank = c->cell_rank.p;
    np.cell_start = c->cell_start.p;
    np.grid = c->grid;
    np.thr2 = c->thr2;
    np.n_rows = n_new;
    np.n_total = n_new + c->n_ghost;
    for (int k = 0; k < 2; ++k)
      {
        np.ghost_start[k] = c->n_ghost_run[k] ? c->ghost_start[k].p : nullptr;
        np.ghost_end[k] = c->n_ghost_run[k] ? c->ghost_end[k].p : nullptr;
      }
    np.old_n_owned = c->old_n_owned;
    np.old_list = oldl.view();
    np.old_of_new = c->old_of_new.p;
    np.n_old_rows = oldl.n_rows;
    np.clear_history = c->clear_history_trigger ? 1 : 0;
    np.new_list = newl.view();
    np.counts = c->counts.p;
    np.use_roll = use_roll;
    np.use_img = use_img;
    HistPayload pay{c->n_pay ? c->pay.p : nullptr, c->pay_start.p, c->n_pay, c->slot_map_size, stn.id.p};
    np.pay = pay;
    c->nb_cand.ensure(std::max<size_t>(size_t(NB_CACHE) * n_new, 1));
    if (use_img)
      c->nb_cand_img.ensure(std::max<size_t>(size_t(NB_CACHE) * n_new, 1));
    np.mobility = c->asc_in_force ? c->asc_cell_status.p : nullptr;
    np.cand = c->nb_cand.p;
    np.cand_img = use_img ? c->nb_cand_img.p : nullptr;
    launch_count_neighbors(np, s);
    exclusive_scan_u32(c->counts.p, newl.row_start.p, size_t(n_new) + 1, c->scan_tmp.p, s);
    const uint32_t n_entries = n_new ? read_u32(c, newl.row_start.p + n_new) : 0;
    newl.col.ensure(std::max<size_t>(n_entries, 1));
    newl.hist.ensure(std::max<size_t>(size_t(n_entries), 1));
    newl.rowl.ensure(std::max<size_t>(n_entries, 1));
    if (use_roll)
      newl.roll.ensure(std::max<size_t>(3 * size_t(n_entries), 1));
    if (use_img)
      newl.img.ensure(std::max<size_t>(n_entries, 1));
    np.new_list = newl.view();
    launch_fill_neighbors(np, s);
    newl.n_rows = n_new;
    newl.n_entries = n_entries;

    WallListBufs &oldw = c->wlists[c->cur_list];
    WallListBufs &neww = c->wlists[c->cur_list ^ 1];
    neww.row_start.ensure(size_t(n_new) + 2);
    WallBuildParams wp;
    wp.st = stn.view();
    wp.cell_reg = stn.cell_reg.p;
    wp.grid = c->grid;
    wp.faces = face_table(c);
    wp.floating = c->fw_host.n > 0 ? c->fw_dev.p : nullptr;
    wp.cell_fw_mask = c->fw_host.n > 0 ? c->cell_fw_mask.p : nullptr;
    wp.time = c->current_time;
    wp.n_rows = n_new;
    wp.old_list = oldw.view();
    wp.old_of_new = c->old_of_new.p;
    wp.n_old_rows = oldw.n_rows;
    wp.clear_history = c->clear_history_trigger ? 1 : 0;
    wp.new_list = neww.view();
    wp.counts = c->counts.p;
    wp.use_roll = use_roll;
    wp.pay = pay;
    wp.mobility = c->asc_in_force ? c->asc_cell_status.p : nullptr;
    launch_count_walls(wp, s);
    exclusive_scan_u32(c->counts.p, neww.row_start.p, size_t(n_new) + 1, c->scan_tmp.p, s);
    const uint32_t n_wall = n_new ? read_u32(c, neww.row_start.p + n_new) : 0;
    neww.entry.ensure(std::max<size_t>(n_wall, 1));
    neww.hist.ensure(std::max<size_t>(3 * size_t(n_wall), 1));
    if (use_roll)
      neww.roll.ensure(std::max<size_t>(3 * size_t(n_wall), 1));
    wp.new_list = neww.view();
    launch_fill_walls(wp, s);
    neww.n_rows = n_new;
    neww.n_entries = n_wall;

    c->cur_list ^= 1;
    rebuild_solid_lists(c);
    CU_TRY(cudaMemsetAsync(c->flag_dev.p, 0, 4 * sizeof(uint32_t), s));
    CU_TRY(cudaStreamSynchronize(s));
    *c->h_flag = 0; // no kernel in flight can touch the flag here
    ++c->n_rebuilds;
  }

  void rebuild(Ctx *c)
  {
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (c->timers_enabled)
      {
        ev0 = get_event(c);
        ev1 = get_event(c);
        CU_TRY(cudaEventRecord(ev0, c->stream));
      }
    upload_walls(c);
    rebuild_sort(c);
    identify_mobility_status(c);
    rebuild_lists(c);
    if (c->timers_enabled)
      {
        CU_TRY(cudaEventRecord(ev1, c->stream));
        c->pending_rebuild.emplace_back(ev0, ev1);
        ++c->rebuild_launches;
      }
  }

  // non-zero tag of a step: what its kernel leaves in the trigger flag
  inline uint32_t step_tag(uint64_t iteration) { return uint32_t(iteration % 0x7fffffffull) + 1u; }

  // DEM-MP: heat rates of this step's contacts and the temperature step (dem.cc:1134-1153), on the state and the
  // list the step kernel is about to read
  void heat_transfer_step(Ctx *c)
  {
    if (!c->thermal_enabled || !c->n_owned)
      return;
    cudaStream_t s = c->stream;
    if (c->thermal_size < c->slot_map_size)
      throw std::runtime_error("heat transfer: lethe_dem_set_temperatures has not covered every particle id");
    c->heat_rate.ensure(c->n_owned);
    HeatParams hp;
    hp.in = c->st[c->cur].view();
    hp.list = c->lists[c->cur_list].view();
    hp.id = c->st[c->cur].id.p;
    hp.n_owned = c->n_owned;
    hp.periodic_any = c->grid.periodic[0] || c->grid.periodic[1] || c->grid.periodic[2];
    hp.dt = c->cfg.dt;
    for (int d = 0; d < 3; ++d)
      hp.L[d] = c->grid.L[d];
    hp.temperature = c->temperature.p;
    hp.specific_heat = c->specific_heat.p;
    hp.rate = c->heat_rate.p;
    launch_heat_rates(c->cfg.pp_model, hp, c->mt, c->thermal_tables, s);
    launch_integrate_temperature(hp, s);
  }

  // the operands of one step kernel launch that do not depend on solids / external loads
  void fill_step_params(Ctx *c, int phase, bool speculative, StepParams &P)
  {
    cudaStream_t s = c->stream;
    std::memset(&P, 0, sizeof(P));
    P.in = c->st[c->cur].view();
    P.out = c->st[c->cur ^ 1].view();
    P.list = c->lists[c->cur_list].view();
    P.walls = c->wlists[c->cur_list].view();
    P.id = c->st[c->cur].id.p;
    P.disp = c->disp.p;
    P.flag_local = c->flag_dev.p;
    P.flag_host = c->d_flag;
    P.flag_check = c->flag_dev.p;
    P.flag_tag = step_tag(c->iteration_number);
    P.spec_check = speculative ? 1 : 0;
    if (c->multi.enabled())
      {
        c->multi.fill_halo(c, c->cur ^ 1, P.halo);
        P.flag_check = c->multi.agreed_flag_dev(c); // the job-wide agreement, never this step's own tag
      }
    // the first iteration after a reset integrates everything (check_mobility_status_reset)
    P.row_mobile = (c->asc_in_force && !c->asc_reset) ? c->asc_row_mobile.p : nullptr;
    if (c->cfg.store_forces || c->count_touching)
      {
        c->touching.ensure(1);
        CU_TRY(cudaMemsetAsync(c->touching.p, 0, sizeof(unsigned long long), s));
        P.touching_counter = c->touching.p;
      }
    if (c->cfg.store_forces)
      {
        c->force_out.ensure(std::max<size_t>(3 * size_t(c->n_owned), 1));
        c->torque_out.ensure(std::max<size_t>(3 * size_t(c->n_owned), 1));
        P.force_out = c->force_out.p;
        P.torque_out = c->torque_out.p;
      }
    P.faces = face_table(c);
    P.motions = c->motions.p;
    P.floating = c->fw_dev.p;
    P.n_owned = c->n_owned;
    P.phase = phase;
    P.integrator = c->cfg.integrator;
    P.mixed_precision = c->cfg.precision == LETHE_PRECISION_MIXED ? 1 : 0;
    P.pw_model = c->cfg.pw_model;
    P.rolling_model = c->cfg.rolling_model;
    P.periodic_any = c->grid.periodic[0] || c->grid.periodic[1] || c->grid.periodic[2];
    P.dt = c->cfg.dt;
    for (int d = 0; d < 3; ++d)
      {
        P.g[d] = c->cfg.g[d];
        P.L[d] = c->grid.L[d];
      }
    P.criterion = c->cfg.smallest_contact_search_criterion;
    P.moi_override = c->cfg.moi_override;
  }

  void launch_step_kernel(Ctx *c, int phase, bool speculative = false)
  {
    cudaStream_t s = c->stream;
    StepParams P;
    fill_step_params(c, phase, speculative, P);
    if (c->n_solids)
      {
        // dem.cc:1141-1147: move the solids (not in the closing half step, dem.cc:726-728), then
        // the particle - solid surface contacts of this step
        if (phase != PHASE_END)
          {
            SolidMoveParams mp;
            mp.s = solid_view(c);
            mp.dt = c->cfg.dt;
            mp.criterion = 0.57735026918962576451 *
                           std::sqrt((c->grid.h[0] * c->grid.h[0] + c->grid.h[1] * c->grid.h[1]) + c->grid.h[2] * c->grid.h[2]);
            mp.flag_local = P.flag_local;
            mp.flag_host = P.flag_host;
            mp.remap_host = c->d_remap;
            mp.flag_check = P.flag_check;
            mp.flag_tag = P.flag_tag;
            mp.spec_check = P.spec_check;
            launch_move_solids(mp, s);
          }
        SolidContactParams sp;
        sp.s = solid_view(c);
        sp.list = c->slists[c->cur_list].view();
        sp.active = c->solid_active.p;
        sp.n_active = c->n_solid_active;
        sp.in = P.in;
        sp.force = c->solid_force.p;
        sp.torque = c->solid_torque.p;
        sp.overflow = c->solid_overflow.p;
        sp.pw_model = c->cfg.pw_model;
        sp.rolling_model = c->cfg.rolling_model;
        sp.dt = c->cfg.dt;
        sp.flag_check = P.flag_check;
        sp.flag_tag = P.flag_tag;
        sp.spec_check = P.spec_check;
        sp.flag_local = P.flag_local;
        sp.flag_host = P.flag_host;
        launch_solid_contacts(sp, c->mt, s);
        P.solid_force = c->solid_force.p;
        P.solid_torque = c->solid_torque.p;
      }
    if (c->ext_enabled && c->n_owned)
      {
        // add_fluid_particle_interaction_force / _torque (cfd_dem_coupling.cc:881-925): after the
        // contact forces (solid surfaces included), before the integration
        c->addend_force.ensure(3 * size_t(c->n_owned));
        c->addend_torque.ensure(3 * size_t(c->n_owned));
        launch_compose_external_loads(P.id, c->n_owned, c->ext_force.p, c->ext_torque.p, c->ext_size, P.solid_force, P.solid_torque,
                                      c->addend_force.p, c->addend_torque.p, s);
        P.solid_force = c->addend_force.p;
        P.solid_torque = c->addend_torque.p;
      }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (c->timers_enabled)
      {
        ev0 = get_event(c);
        ev1 = get_event(c);
        CU_TRY(cudaEventRecord(ev0, s));
      }
    launch_step(c->cfg.pp_model, c->cfg.rolling_model, P, c->mt, s);
    if (c->timers_enabled)
      {
        CU_TRY(cudaEventRecord(ev1, s));
        c->pending_step.emplace_back(ev0, ev1);
        ++c->step_launches;
        if (c->pending_step.size() > 4096)
          resolve_timers(c);
      }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(c->step_done[c->iteration_number & 1], s));
    // ids / registered cells do not change in a step: both generations share them logically;
    // keep the other generation's copies in sync lazily (they are only read at rebuilds).
    c->cur ^= 1;
  }

  // ids and cell_reg live per generation; after a step the "current" generation flips, so
  // mirror them once per rebuild into both generations.
  void mirror_ids(Ctx *c)
  {
    StateBufs &a = c->st[c->cur], &b = c->st[c->cur ^ 1];
    const size_t n = size_t(c->n_owned) + c->n_ghost;
    if (!n)
      return;
    b.id.ensure(n);
    b.cell_reg.ensure(n);
    CU_TRY(cudaMemcpyAsync(b.id.p, a.id.p, n * 4, cudaMemcpyDeviceToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(b.cell_reg.p, a.cell_reg.p, n * 4, cudaMemcpyDeviceToDevice, c->stream));
  }

  // execute_contact_detection_and_search (dem.cc:598-688) + check functions (dem.cc:459-482)
  void contact_detection_and_search(Ctx *c)
  {
    bool search = c->contact_search_trigger;
    const uint64_t freq = uint64_t(std::max(1, c->cfg.contact_detection_frequency));
    if (c->cfg.detection == LETHE_DETECTION_CONSTANT)
      {
        if ((c->iteration_number % freq) == 0)
          search = true;
      }
    else if (!search && (c->iteration_number % freq) == 0)
      {
        // max displacement > criterion, evaluated by the previous step kernel
        if (c->multi.enabled())
          search = c->multi.any_rank_flag(c); // MPI logical_or (find_contact_detection_step.cc:53-58)
        else
          {
            CU_TRY(cudaStreamSynchronize(c->stream));
            if (*c->h_flag)
              search = true;
          }
      }
    if (c->multi.enabled())
      search = c->multi.agree(c, search); // logical_or over ranks; insertion on one rank triggers all
    if (search)
      {
        if (c->multi.enabled())
          c->multi.rebuild_with_exchange(c);
        else
          {
            rebuild(c);
            mirror_ids(c);
          }
      }
    else if (c->multi.enabled())
      c->multi.refresh_ghosts(c);
  }

  void drop_last_step_timer(Ctx *c);

  // Pipelined form of "check the flag, then step" for the common case (one GPU, dynamic
  // detection, nothing forced): the step kernel is queued SPECULATIVELY behind the previous
  // one, then the host waits for the previous kernel only and reads its flag. If that flag asks
  // for a new list, the queued launch has found the same flag on the device and has returned
  // without touching anything; the host takes the state flip back, rebuilds and launches the
  // step for real. The GPU therefore always has the next kernel queued and never waits for the
  // host's flag read; the sequence of kernels that DO something is exactly the unpipelined one.
  bool step_speculatively(Ctx *c, int phase)
  {
    if (!c->pipeline || c->multi.enabled() || c->contact_search_trigger || c->cfg.detection != LETHE_DETECTION_DYNAMIC || c->thermal_enabled)
      return false;
    const uint64_t freq = uint64_t(std::max(1, c->cfg.contact_detection_frequency));
    if ((c->iteration_number % freq) != 0 || phase != PHASE_REGULAR)
      return false;
    launch_step_kernel(c, phase, true);
    CU_TRY(cudaEventSynchronize(c->step_done[(c->iteration_number - 1) & 1])); // previous step's kernel
    const uint32_t seen = *c->h_flag;
    if (seen == 0u || seen == step_tag(c->iteration_number))
      return true; // nothing asked for a new list before this step: the launch is the step
    // void launch: undo its bookkeeping, rebuild, run the step
    c->cur ^= 1;
    ++c->n_void_launches;
    drop_last_step_timer(c);
    rebuild(c);
    mirror_ids(c);
    launch_step_kernel(c, phase, false);
    return true;
  }

  void drop_last_step_timer(Ctx *c)
  {
    if (c->timers_enabled && !c->pending_step.empty())
      {
        c->event_pool.push_back(c->pending_step.back().first);
        c->event_pool.push_back(c->pending_step.back().second);
        c->pending_step.pop_back();
        --c->step_launches;
      }
  }

  // Multi-GPU step with the fused halo (MultiGpu::fused): the ghost copies this step reads were
  // stored by the neighbours' previous step kernels, so the only collective of a step is the
  // 4-byte agreement "does anybody need a new list" — find_contact_detection_step.cc:53-58's
  // logical_or — which is also the barrier that orders those peer stores. It is queued on the
  // stream, the step kernel is queued speculatively behind it (void if the agreement is
  // non-zero), and the host reads the agreement while the kernel already runs.
  void step_fused_multi(Ctx *c, int phase)
  {
    const uint64_t freq = uint64_t(std::max(1, c->cfg.contact_detection_frequency));
    const bool on_check_iteration = (c->iteration_number % freq) == 0;
    const bool forced = c->contact_search_trigger || (c->cfg.detection == LETHE_DETECTION_CONSTANT && on_check_iteration);
    const bool consult = c->cfg.detection == LETHE_DETECTION_DYNAMIC && on_check_iteration;
    c->multi.post_agree(c, forced ? 0x80000000u : 0u, consult);
    const bool launched = !forced; // a rank that knows the answer does not speculate
    if (launched)
      launch_step_kernel(c, phase, true);
    if (c->multi.wait_agree(c) == 0u)
      return;
    if (launched)
      {
        c->cur ^= 1; // the launch was void on every rank
        ++c->n_void_launches;
        drop_last_step_timer(c);
      }
    c->multi.rebuild_with_exchange(c);
    launch_step_kernel(c, phase, false);
  }

  void one_step(Ctx *c)
  {
    c->iteration_number++;
    c->current_time += c->cfg.dt;
    const int phase = ((c->iteration_number <= 1 && !c->cfg.restart) || c->open_next_step) ? PHASE_START : PHASE_REGULAR;
    c->open_next_step = false;
    // DEMSolver::load_balance (dem.cc:383-457): a load-balance iteration repartitions and searches
    if (c->multi.enabled() && c->multi.load_balance_due(c))
      {
        c->lb_recut_pending = true;
        c->contact_search_trigger = true;
      }
    if (c->multi.fused() && c->pipeline)
      step_fused_multi(c, phase);
    else if (!step_speculatively(c, phase))
      {
        contact_detection_and_search(c);
        heat_transfer_step(c);
        launch_step_kernel(c, phase);
      }
    // reset_triggers (dem_action_manager.h:61-75): the iteration after a mobility-status reset
    // searches again, with the statuses identified from the velocities of the full step
    c->contact_search_trigger = c->asc_reset;
    c->asc_reset = false;
    c->clear_history_trigger = false;
  }

  template <class F> int guarded(Ctx *c, F &&f)
  {
    try
      {
        CU_TRY(cudaSetDevice(c->device));
        f();
        return 0;
      }
    catch (const std::exception &e)
      {
        c->error = e.what();
        return -1;
      }
  }

  void append_particles(Ctx *c, uint64_t n, const uint32_t *id, const double *x3, const double *props9)
  {
    if (n == 0)
      return;
    cudaStream_t s = c->stream;
    const size_t base = c->n_owned;
    const size_t total = base + n;
    if (total >= 0x7fffffffull)
      throw std::runtime_error("too many particles for 31-bit indices");
    for (uint64_t k = 0; k < n; ++k)
      {
        const double *p = props9 + 9 * k;
        if (!(p[0] >= 0.0) || !(p[0] < double(c->cfg.n_types)) || p[0] != std::floor(p[0]))
          throw std::runtime_error("particle type outside [0, number of particle types)");
        if (!(p[1] > 0.0) || !(p[2] > 0.0))
          throw std::runtime_error("particle diameter and mass must be positive");
      }
    c->st[c->cur].ensure(total, base, s);
    c->st[c->cur ^ 1].ensure(total, 0, s);
    c->disp.ensure(total, base, s);
    c->stage_ids.ensure(n);
    c->stage_x.ensure(3 * n);
    c->stage_p.ensure(9 * n);
    CU_TRY(cudaMemcpyAsync(c->stage_ids.p, id, n * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(c->stage_x.p, x3, 3 * n * 8, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(c->stage_p.p, props9, 9 * n * 8, cudaMemcpyHostToDevice, s));
    launch_unpack_host_rows(c->stage_ids.p, c->stage_x.p, c->stage_p.p, uint32_t(n), c->st[c->cur].view(), c->st[c->cur].id.p,
                            c->st[c->cur].cell_reg.p, c->disp.p, uint32_t(base), s);
    uint32_t max_id = 0;
    for (uint64_t k = 0; k < n; ++k)
      max_id = std::max(max_id, id[k]);
    if (size_t(max_id) + 1 > c->slot_map_size)
      {
        const size_t old = c->slot_map_size;
        c->slot_of_id.ensure(size_t(max_id) + 1, old, s, 1.5);
        launch_fill_u32(c->slot_of_id.p + old, 0xffffffffu, size_t(max_id) + 1 - old, s);
        c->slot_map_size = uint32_t(size_t(max_id) + 1);
      }
    // the new rows are addressable by id at once (lethe_dem_step_host* resolve their rows through
    // this map before the next rebuild renumbers the slots)
    launch_register_ids(c->st[c->cur].id.p, uint32_t(base), uint32_t(n), c->slot_of_id.p, c->slot_map_size, s);
    CU_TRY(cudaStreamSynchronize(s));
    c->n_owned = uint32_t(total);
    c->n_ghost = 0;
    // DEMActionManager::particle_insertion_step
    c->contact_search_trigger = true;
  }


  // ------------------------------------------------------------ streamed host step ----
  // lethe_dem_step_host_state moves 72 B per particle up and 72 B down around a step that takes a tenth of either copy.
  // PCIe is full duplex, but "upload everything, step, download everything" uses one direction at a time. The streamed
  // form takes the step apart: the host rows go up in the caller's order in a few contiguous stages, the
  // step kernel is launched on the 128-row blocks whose own rows and listed neighbours have all arrived
  // (StepParams::block_list), and a segment goes back down as soon as every block it has rows in has been stepped — so the
  // download of the first slabs runs under the upload of the later ones. The kernels that touch a particle, their
  // operands and their order per particle are those of the plain call: results are bit-identical
  // (test_streamed_host_step_is_bitwise_identical). The plan depends on the list and on the row -> id table only, so it
  // is made once per rebuild (HostPlanParams, dem_kernels.cuh). Calls that cannot be taken apart (a new list is due,
  // solids, external loads, sparse contacts, heat transfer, several ranks, more than one step) run the plain form.
  int env_int(const char *name, int fallback)
  {
    const char *e = std::getenv(name);
    return e && *e ? std::atoi(e) : fallback;
  }

  bool host_pipe_eligible(Ctx *c, uint64_t n_steps, uint64_t n)
  {
    const int wanted = env_int("LETHE_DEM_HOST_PIPELINE", 1); // read per call: tests switch it between calls
    const int min_rows = env_int("LETHE_DEM_HOST_PIPELINE_MIN_ROWS", 262144);
    if (!wanted || n_steps != 1 || n < uint64_t(std::max(1, min_rows)) || n >= 0xffffffffull)
      return false;
    if (c->thermal_enabled || c->n_solids || c->ext_enabled || c->asc_enabled || c->timers_enabled || c->cfg.store_forces ||
        c->count_touching)
      return false;
    // several ranks: only the fused-halo form of the step (peer stores + mailbox agreement) can be taken apart, and a
    // load-balance iteration is a collective decision of its own
    const bool multi = c->multi.enabled();
    if (multi && !(c->multi.fused() && c->pipeline && c->lb_method == 0))
      return false;
    if (c->contact_search_trigger || c->clear_history_trigger || c->open_next_step || c->lists[c->cur_list].n_rows != c->n_owned ||
        c->n_owned == 0)
      return false;
    const uint64_t it = c->iteration_number + 1;
    if (it <= 1 && !c->cfg.restart)
      return false; // the opening half step
    const uint64_t freq = uint64_t(std::max(1, c->cfg.contact_detection_frequency));
    if ((it % freq) == 0)
      {
        if (c->cfg.detection == LETHE_DETECTION_CONSTANT)
          return false;
        if (multi)
          return true; // the ranks agree on the flag inside the call (step_host_state_streamed)
        CU_TRY(cudaStreamSynchronize(c->stream));
        if (*c->h_flag)
          return false; // the previous step asked for a new list: the plain call rebuilds
      }
    return true;
  }

  void host_pipe_plan(Ctx *c, uint64_t n)
  {
    auto &hp = c->host_pipe;
    cudaStream_t s = c->stream;
    hp.valid = false;
    hp.rebuild_gen = c->n_rebuilds;
    hp.ids_version = c->host_row_ids_version;
    hp.n_rows = uint32_t(n);
    hp.n_owned = c->n_owned;
    ++hp.n_plans;
    hp.seg_rows = uint32_t(std::max(256, env_int("LETHE_DEM_HOST_SEG_ROWS", 4096)));
    while ((n + hp.seg_rows - 1) / hp.seg_rows > uint64_t(std::max(64, env_int("LETHE_DEM_HOST_MAX_SEGS", 4096))))
      hp.seg_rows *= 2;
    hp.n_seg = uint32_t((n + hp.seg_rows - 1) / hp.seg_rows);
    // stages: enough that the tail (the download of what only the last upload makes ready) is short, few enough that a
    // stage's copies and launches stay large against their fixed cost
    // (measured, B200 / PCIe Gen5: a stage costs about 0.1 ms of launch chain — scatter, partial step launch of less
    // than a few waves, pack — so it should carry some 0.4 ms of copy: 262144 rows; 1 M rows: 4 stages, 8 M rows and more: 32)
    const int auto_stages = int(std::min<uint64_t>(32, std::max<uint64_t>(2, (n + 131072) / 262144)));
    const int K = std::min(std::min(env_int("LETHE_DEM_HOST_STAGES", auto_stages), 120), int(hp.n_seg));
    if (K < 2)
      return;
    if (!hp.s_up)
      {
        CU_TRY(cudaStreamCreateWithFlags(&hp.s_up, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&hp.s_down, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&hp.ev_begin, cudaEventDisableTiming));
      }
    while (hp.ev_up.size() < size_t(K))
      {
        cudaEvent_t a, b;
        CU_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        hp.ev_up.push_back(a);
        hp.ev_down.push_back(b);
      }
    hp.n_stages = uint32_t(K);
    hp.n_blocks = (c->n_owned + STEP_BLOCK_ROWS - 1) / STEP_BLOCK_ROWS;
    const size_t n_slots = size_t(c->n_owned) + c->n_ghost;
    hp.seg_up.ensure(hp.n_seg);
    hp.seg_down.ensure(hp.n_seg);
    hp.block_ready.ensure(hp.n_blocks);
    hp.up_stage_of_slot.ensure(n_slots);
    hp.up_list.ensure(hp.n_seg);
    hp.down_list.ensure(hp.n_seg);
    hp.block_list.ensure(hp.n_blocks);
    hp.row_of_slot.ensure(std::max<size_t>(c->n_owned, 1));
    CU_TRY(cudaMemsetAsync(hp.row_of_slot.p, 0xff, size_t(c->n_owned) * 4, s));
    CU_TRY(cudaMemsetAsync(hp.seg_down.p, 0, size_t(hp.n_seg) * 4, s));
    CU_TRY(cudaMemsetAsync(hp.block_ready.p, 0, size_t(hp.n_blocks) * 4, s));
    CU_TRY(cudaMemsetAsync(hp.up_stage_of_slot.p, 0, n_slots, s));
    HostPlanParams P;
    std::memset(&P, 0, sizeof(P));
    P.row_ids = c->host_row_ids.p;
    P.n_rows = hp.n_rows;
    P.seg_rows = hp.seg_rows;
    P.slot_of_id = c->slot_of_id.p;
    P.map_size = c->slot_map_size;
    P.list = c->lists[c->cur_list].view();
    P.n_owned = c->n_owned;
    P.seg_up = hp.seg_up.p;
    P.seg_down = hp.seg_down.p;
    P.up_stage_of_slot = hp.up_stage_of_slot.p;
    P.row_of_slot = hp.row_of_slot.p;
    P.block_ready = hp.block_ready.p;
    launch_host_plan(P, 0, s);
    // upload stages: the caller's row order cut into K contiguous pieces, one copy each. (Ordering the segments
    // breadth-first over their contact adjacency instead was measured and is worse: on a 3-D packing the waves are
    // shells of hundreds of segments, so the readiness lag grows and the uploads fall apart into small copies.)
    std::vector<uint32_t> seg_up(hp.n_seg), seg_down(hp.n_seg), ready(hp.n_blocks);
    for (uint32_t g = 0; g < hp.n_seg; ++g)
      seg_up[g] = uint32_t(uint64_t(g) * uint64_t(K) / hp.n_seg);
    CU_TRY(cudaMemcpyAsync(hp.seg_up.p, seg_up.data(), size_t(hp.n_seg) * 4, cudaMemcpyHostToDevice, s));
    for (int pass = 1; pass < 4; ++pass)
      launch_host_plan(P, pass, s);
    CU_TRY(cudaMemcpyAsync(seg_down.data(), hp.seg_down.p, size_t(hp.n_seg) * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(ready.data(), hp.block_ready.p, size_t(hp.n_blocks) * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    // a segment none of whose rows names a particle of this context still travels both ways (its rows come back unchanged)
    for (uint32_t g = 0; g < hp.n_seg; ++g)
      seg_down[g] = std::min(std::max(seg_down[g], seg_up[g]), uint32_t(K - 1));
    // stable counting sorts by stage; consecutive segments of one stage move with one copy
    auto order = [&](const std::vector<uint32_t> &stage, std::vector<uint32_t> &off, std::vector<uint32_t> &list) {
      off.assign(size_t(K) + 1, 0);
      for (uint32_t v : stage)
        ++off[std::min(v, uint32_t(K - 1)) + 1];
      for (int k = 0; k < K; ++k)
        off[k + 1] += off[k];
      list.resize(stage.size());
      std::vector<uint32_t> at(off.begin(), off.end() - 1);
      for (uint32_t g = 0; g < stage.size(); ++g)
        list[at[std::min(stage[g], uint32_t(K - 1))]++] = g;
    };
    auto runs_of = [&](const std::vector<uint32_t> &off, const std::vector<uint32_t> &list,
                       std::vector<std::vector<std::pair<uint64_t, uint64_t>>> &runs) {
      runs.assign(size_t(K), {});
      for (int k = 0; k < K; ++k)
        for (uint32_t a = off[k]; a < off[k + 1];)
          {
            uint32_t b = a + 1;
            while (b < off[k + 1] && list[b] == list[b - 1] + 1)
              ++b;
            const uint64_t first = uint64_t(list[a]) * hp.seg_rows;
            const uint64_t end = std::min<uint64_t>(n, uint64_t(list[b - 1] + 1) * hp.seg_rows);
            runs[k].emplace_back(first, end - first);
            a = b;
          }
    };
    std::vector<uint32_t> up_list, down_list, block_list;
    order(seg_up, hp.up_off, up_list);
    order(seg_down, hp.down_off, down_list);
    order(ready, hp.block_off, block_list);
    runs_of(hp.up_off, up_list, hp.up_runs);
    runs_of(hp.down_off, down_list, hp.down_runs);
    CU_TRY(cudaMemcpyAsync(hp.up_list.p, up_list.data(), size_t(hp.n_seg) * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(hp.down_list.p, down_list.data(), size_t(hp.n_seg) * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(hp.block_list.p, block_list.data(), size_t(hp.n_blocks) * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaStreamSynchronize(s)); // the host vectors go out of scope
    hp.valid = true;
  }

  void step_host_state_streamed(Ctx *c, uint64_t n, double *state9)
  {
    auto &hp = c->host_pipe;
    cudaStream_t s = c->stream;
    const int K = int(hp.n_stages);
    c->stage_p.ensure(9 * n);
    c->iteration_number++;
    c->current_time += c->cfg.dt;
    if (c->multi.enabled())
      {
        // the one collective of a step (step_fused_multi): does any rank need a new list? It is also the barrier behind
        // which the neighbours' halo stores of the previous step are visible.
        const uint64_t freq = uint64_t(std::max(1, c->cfg.contact_detection_frequency));
        const bool consult = c->cfg.detection == LETHE_DETECTION_DYNAMIC && (c->iteration_number % freq) == 0;
        c->multi.post_agree(c, 0u, consult);
        if (c->multi.wait_agree(c) != 0u)
          {
            // plain form of this call: rows up, new list with migration, step, rows down
            CU_TRY(cudaMemcpyAsync(c->stage_p.p, state9, 9 * n * 8, cudaMemcpyHostToDevice, s));
            launch_update_state_rows(c->host_row_ids.p, c->stage_p.p, uint32_t(n), c->slot_of_id.p, c->slot_map_size, c->st[c->cur].view(), s);
            c->multi.rebuild_with_exchange(c);
            launch_step_kernel(c, PHASE_REGULAR, false);
            c->contact_search_trigger = false;
            c->clear_history_trigger = false;
            launch_pack_state_rows(c->host_row_ids.p, uint32_t(n), c->slot_of_id.p, c->slot_map_size, c->st[c->cur].view(), c->stage_p.p, s);
            CU_TRY(cudaMemcpyAsync(state9, c->stage_p.p, 9 * n * 8, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaStreamSynchronize(s));
            return;
          }
      }
    StepParams P;
    fill_step_params(c, PHASE_REGULAR, false, P);
    const StateView out = c->st[c->cur ^ 1].view();
    // page-locked host rows are addressable from the device: the stepped blocks' rows are then written straight into them
    // by the pack kernel of each stage (no staging, no per-segment wait for the last block of a segment)
    double *state9_dev = nullptr;
    if (env_int("LETHE_DEM_HOST_ZEROCOPY", 0)) // measured on B200 / PCIe Gen5: device stores into host memory reach about half the copy engines' rate
      {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, state9) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
          state9_dev = static_cast<double *>(attr.devicePointer);
        else
          cudaGetLastError();
      }
    CU_TRY(cudaEventRecord(hp.ev_begin, s));
    CU_TRY(cudaStreamWaitEvent(hp.s_up, hp.ev_begin, 0));
    CU_TRY(cudaStreamWaitEvent(hp.s_down, hp.ev_begin, 0));
    for (int k = 0; k < K; ++k)
      {
        for (const auto &r : hp.up_runs[k])
          CU_TRY(cudaMemcpyAsync(c->stage_p.p + 9 * r.first, state9 + 9 * r.first, 72 * r.second, cudaMemcpyHostToDevice, hp.s_up));
        CU_TRY(cudaEventRecord(hp.ev_up[k], hp.s_up));
        CU_TRY(cudaStreamWaitEvent(s, hp.ev_up[k], 0));
        launch_update_state_rows_segs(hp.up_list.p + hp.up_off[k], hp.up_off[k + 1] - hp.up_off[k], hp.seg_rows, c->host_row_ids.p,
                                      c->stage_p.p, uint32_t(n), c->slot_of_id.p, c->slot_map_size, P.in, s);
        P.block_list = hp.block_list.p + hp.block_off[k];
        P.n_blocks_listed = hp.block_off[k + 1] - hp.block_off[k];
        if (P.n_blocks_listed)
          launch_step(c->cfg.pp_model, c->cfg.rolling_model, P, c->mt, s);
        if (state9_dev)
          {
            launch_pack_state_rows_blocks(P.block_list, P.n_blocks_listed, hp.row_of_slot.p, c->n_owned, out, state9_dev, s);
            continue;
          }
        launch_pack_state_rows_segs(hp.down_list.p + hp.down_off[k], hp.down_off[k + 1] - hp.down_off[k], hp.seg_rows, c->host_row_ids.p,
                                    uint32_t(n), c->slot_of_id.p, c->slot_map_size, out, c->stage_p.p, s);
        CU_TRY(cudaEventRecord(hp.ev_down[k], s));
        CU_TRY(cudaStreamWaitEvent(hp.s_down, hp.ev_down[k], 0));
        for (const auto &r : hp.down_runs[k])
          CU_TRY(cudaMemcpyAsync(state9 + 9 * r.first, c->stage_p.p + 9 * r.first, 72 * r.second, cudaMemcpyDeviceToHost, hp.s_down));
      }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(c->step_done[c->iteration_number & 1], s));
    c->cur ^= 1;
    // reset_triggers (one_step)
    c->contact_search_trigger = false;
    c->clear_history_trigger = false;
    ++hp.n_calls;
    if (state9_dev)
      ++hp.n_zero_copy_calls;
    CU_TRY(cudaStreamSynchronize(s));
    CU_TRY(cudaStreamSynchronize(hp.s_down));
  }

  struct HostRows
  {
    std::vector<uint32_t> id;
    std::vector<double> x, p;
  };
  void download_rows(Ctx *c, HostRows &h)
  {
    const size_t n = c->n_owned;
    h.id.resize(n);
    h.x.resize(3 * n);
    h.p.resize(9 * n);
    if (!n)
      return;
    cudaStream_t s = c->stream;
    c->stage_ids.ensure(n);
    c->stage_x.ensure(3 * n);
    c->stage_p.ensure(9 * n);
    launch_pack_all_rows(c->st[c->cur].view(), c->st[c->cur].id.p, uint32_t(n), c->stage_ids.p, c->stage_x.p, c->stage_p.p, s);
    CU_TRY(cudaMemcpyAsync(h.id.data(), c->stage_ids.p, n * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(h.x.data(), c->stage_x.p, 3 * n * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(h.p.data(), c->stage_p.p, 9 * n * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
  }
} // namespace

// hooks used by dem_multi.cu
namespace dem
{
  void engine_rebuild_local(lethe_dem_ctx *c)
  {
    rebuild(c);
    mirror_ids(c);
  }
  void engine_upload_walls(lethe_dem_ctx *c) { upload_walls(c); }
  void engine_rebuild_sort(lethe_dem_ctx *c) { rebuild_sort(c); }
  void engine_rebuild_lists(lethe_dem_ctx *c) { rebuild_lists(c); }
  void engine_identify_mobility_status(lethe_dem_ctx *c) { identify_mobility_status(c); }
  void engine_mirror_ids(lethe_dem_ctx *c) { mirror_ids(c); }
  cudaEvent_t engine_get_event(lethe_dem_ctx *c) { return get_event(c); }
} // namespace dem

// =================================================================== C ABI ===
// ---- CFD-DEM rows of 23 properties (dem_properties.h:92-142) on top of the 9-property calls ----
namespace
{
  void split_cfd_rows(uint64_t n, const double *props23, std::vector<double> *props9, std::vector<double> &force3, std::vector<double> &torque3)
  {
    if (props9)
      props9->resize(9 * n);
    force3.resize(3 * n);
    torque3.resize(3 * n);
    for (uint64_t k = 0; k < n; ++k)
      {
        const double *p = props23 + 23 * k;
        if (props9)
          std::memcpy(props9->data() + 9 * k, p, 72);
        for (int d = 0; d < 3; ++d)
          {
            // add_fluid_particle_interaction_force (cfd_dem_coupling.cc:891-902): two_way + one_way + drag
            force3[3 * k + d] = (p[9 + d] + p[12 + d]) + p[15 + d];
            torque3[3 * k + d] = p[18 + d];
          }
      }
  }
} // namespace

extern "C" {

const char *lethe_dem_create_error(void) { return g_create_error.c_str(); }

int lethe_dem_create(const lethe_dem_config *config, int device, lethe_dem_ctx **out)
{
  if (!config || !out)
    {
      g_create_error = "null argument";
      return -1;
    }
  auto bad = [&](const char *m) {
    g_create_error = m;
    return -1;
  };
  if (config->n_types < 1 || config->n_types > LETHE_DEM_MAX_TYPES)
    return bad("n_types out of range (1..5)");
  if (config->grid_n[0] < 1 || config->grid_n[1] < 1 || config->grid_n[2] < 1)
    return bad("grid_n must be positive");
  if (config->integrator != LETHE_INTEGRATOR_VELOCITY_VERLET && config->integrator != LETHE_INTEGRATOR_EXPLICIT_EULER)
    return bad("unknown integrator (velocity_verlet or explicit_euler)");
  if (config->pp_model < 0 || config->pp_model > LETHE_PP_DMT || config->pw_model < 0 || config->pw_model > LETHE_PW_DMT ||
      config->rolling_model < 0 || config->rolling_model > LETHE_ROLLING_EPSD)
    return bad("invalid contact model selector");
  if (config->sparse_contacts && config->integrator == LETHE_INTEGRATOR_EXPLICIT_EULER)
    return bad("Adaptive sparse contacts are not supported with explicit Euler integrator, use Velocity Verlet integrator."); // explicit_euler_integrator.cc:157-159
  if (config->precision != LETHE_PRECISION_F64 && config->precision != LETHE_PRECISION_MIXED)
    return bad("unknown precision (LETHE_PRECISION_F64 or LETHE_PRECISION_MIXED)");
  if (!(config->dt > 0))
    return bad("dt must be positive");
  for (int d = 0; d < 3; ++d)
    if (config->periodic[d] && config->grid_n[d] < 3)
      return bad("a periodic direction needs at least 3 grid cells");
  if (double(config->grid_n[0]) * config->grid_n[1] * config->grid_n[2] > 2.0e9)
    return bad("grid too large");
  // report_cell_size_to_particle_diameter_ratio (include/dem/utilities.h:33-60, called from
  // dem.cc:1093-1096): the reference refuses a mesh whose smallest cell edge is below the largest
  // particle diameter. Its candidate set is "particles of vertex-sharing cells" whatever the
  // neighbourhood threshold, and the 27-cell stencil here reproduces exactly that set, so the
  // reference's own bound is the one enforced (a single layer of cells has no neighbour layer).
  for (int d = 0; d < 3; ++d)
    if (config->grid_n[d] > 1 && config->d_max > 0 && config->cell_size[d] < config->d_max * (1.0 - 1e-12))
      return bad("Minimum cell size is smaller than the maximum particle diameter. Consider coarsening the mesh to achieve a "
                 "ratio larger than 1");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    {
      g_create_error = std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e);
      return -2;
    }
  if (device < 0 || device >= n_dev)
    return bad("device index out of range");
  lethe_dem_ctx *c = nullptr;
  try
    {
      CU_TRY(cudaSetDevice(device));
      c = new lethe_dem_ctx();
      c->cfg = *config;
      c->device = device;
      CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      GridDesc &g = c->grid;
      for (int d = 0; d < 3; ++d)
        {
          g.lo[d] = config->grid_lo[d];
          g.h[d] = config->cell_size[d];
          g.n[d] = config->grid_n[d];
          g.periodic[d] = config->periodic[d];
          g.L[d] = (config->grid_lo[d] + config->grid_n[d] * config->cell_size[d]) - config->grid_lo[d];
        }
      g.n_cells = g.n[0] * g.n[1] * g.n[2];
      g.slab_axis = config->slab_axis;
      g.slab_lo = config->slab_lo;
      g.slab_hi = config->slab_hi;
      build_material_tables(*config, c->mt);
      c->asc_enabled = config->sparse_contacts != 0;
      c->asc_reset = c->asc_enabled; // set_sparse_contacts_enabled (dem_action_manager.h:128-134)
      c->thr2 = std::pow(config->neighborhood_threshold * config->d_max, 2); // dem.cc:156-159
      std::vector<int32_t> rank, inv;
      build_cell_curve(g, rank, inv);
      c->cell_rank.ensure(rank.size());
      c->cell_of_rank.ensure(inv.size());
      CU_TRY(cudaMemcpy(c->cell_rank.p, rank.data(), rank.size() * 4, cudaMemcpyHostToDevice));
      CU_TRY(cudaMemcpy(c->cell_of_rank.p, inv.data(), inv.size() * 4, cudaMemcpyHostToDevice));
      {
        void *hf = nullptr, *df = nullptr;
        CU_TRY(cudaHostAlloc(&hf, sizeof(uint32_t), cudaHostAllocMapped));
        c->h_flag = static_cast<volatile uint32_t *>(hf);
        *c->h_flag = 0;
        CU_TRY(cudaHostGetDevicePointer(&df, hf, 0));
        c->d_flag = static_cast<uint32_t *>(df);
      }
      {
        void *hf = nullptr, *df = nullptr;
        CU_TRY(cudaHostAlloc(&hf, sizeof(uint32_t), cudaHostAllocMapped));
        c->h_remap = static_cast<volatile uint32_t *>(hf);
        *c->h_remap = 0;
        CU_TRY(cudaHostGetDevicePointer(&df, hf, 0));
        c->d_remap = static_cast<uint32_t *>(df);
      }
      c->flag_dev.ensure(4);
      CU_TRY(cudaMemset(c->flag_dev.p, 0, 4 * sizeof(uint32_t)));
      for (auto &ev : c->step_done)
        CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      if (const char *e = getenv("LETHE_DEM_NO_PIPELINE"))
        c->pipeline = !(e[0] == '1');
      std::memset(&c->fw_host, 0, sizeof(c->fw_host));
      // empty lists so that a step with zero particles is well defined
      for (int k = 0; k < 2; ++k)
        {
          c->lists[k].row_start.ensure(2);
          c->wlists[k].row_start.ensure(2);
          CU_TRY(cudaMemset(c->lists[k].row_start.p, 0, 8));
          CU_TRY(cudaMemset(c->wlists[k].row_start.p, 0, 8));
        }
      c->walls_dirty = true;
    }
  catch (const std::exception &ex)
    {
      g_create_error = ex.what();
      delete c;
      return -3;
    }
  *out = c;
  return 0;
}

void lethe_dem_destroy(lethe_dem_ctx *c)
{
  if (!c)
    return;
  cudaSetDevice(c->device);
  if (c->stream)
    cudaStreamSynchronize(c->stream);
  c->multi.shutdown();
  for (auto &pr : c->pending_step)
    {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
  for (auto &pr : c->pending_rebuild)
    {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
  for (auto ev : c->event_pool)
    cudaEventDestroy(ev);
  for (auto ev : c->region_ev)
    if (ev)
      cudaEventDestroy(ev);
  for (auto ev : c->step_done)
    if (ev)
      cudaEventDestroy(ev);
  if (c->h_flag)
    cudaFreeHost(const_cast<uint32_t *>(c->h_flag));
  if (c->h_remap)
    cudaFreeHost(const_cast<uint32_t *>(c->h_remap));
  for (auto ev : c->host_pipe.ev_up)
    cudaEventDestroy(ev);
  for (auto ev : c->host_pipe.ev_down)
    cudaEventDestroy(ev);
  if (c->host_pipe.ev_begin)
    cudaEventDestroy(c->host_pipe.ev_begin);
  if (c->host_pipe.s_up)
    cudaStreamDestroy(c->host_pipe.s_up);
  if (c->host_pipe.s_down)
    cudaStreamDestroy(c->host_pipe.s_down);
  if (c->stream)
    cudaStreamDestroy(c->stream);
  delete c;
}

const char *lethe_dem_last_error(const lethe_dem_ctx *c) { return c ? c->error.c_str() : "null context"; }

int lethe_dem_set_particles(lethe_dem_ctx *c, uint64_t n, const uint32_t *id, const double *x3, const double *props9)
{
  return guarded(c, [&] {
    CU_TRY(cudaStreamSynchronize(c->stream));
    c->n_owned = 0;
    c->n_ghost = 0;
    for (int k = 0; k < 2; ++k)
      {
        c->lists[k].n_rows = 0;
        c->lists[k].n_entries = 0;
        c->wlists[k].n_rows = 0;
        c->wlists[k].n_entries = 0;
        c->slists[k].n_rows = 0;
        c->slists[k].n_entries = 0;
      }
    if (c->slot_map_size)
      launch_fill_u32(c->slot_of_id.p, 0xffffffffu, c->slot_map_size, c->stream);
    append_particles(c, n, id, x3, props9);
    c->contact_search_trigger = true;
  });
}

int lethe_dem_add_particles(lethe_dem_ctx *c, uint64_t n, const uint32_t *id, const double *x3, const double *props9)
{
  return guarded(c, [&] { append_particles(c, n, id, x3, props9); });
}

int lethe_dem_n_particles(lethe_dem_ctx *c, uint64_t *n)
{
  *n = c->n_owned;
  return 0;
}

int lethe_dem_get_particles(lethe_dem_ctx *c, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *x3, double *props9)
{
  return guarded(c, [&] {
    HostRows h;
    download_rows(c, h);
    const size_t n = h.id.size();
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h.id[a] < h.id[b]; });
    const uint64_t m = std::min<uint64_t>(n_max, n);
    for (uint64_t k = 0; k < m; ++k)
      {
        const uint32_t q = order[k];
        id[k] = h.id[q];
        std::memcpy(x3 + 3 * k, h.x.data() + 3 * size_t(q), 24);
        std::memcpy(props9 + 9 * k, h.p.data() + 9 * size_t(q), 72);
      }
    *n_out = m;
  });
}

int lethe_dem_set_walls(lethe_dem_ctx *c, uint64_t n_faces, const lethe_wall_face *faces)
{
  return guarded(c, [&] {
    for (uint64_t k = 0; k < n_faces; ++k)
      if (faces[k].cell < 0 || faces[k].cell >= c->grid.n_cells)
        throw std::runtime_error("wall face refers to a cell outside the grid");
    if (n_faces > WALL_INDEX_MASK)
      throw std::runtime_error("too many wall faces");
    c->faces_host.assign(faces, faces + n_faces);
    c->walls_dirty = true;
    c->contact_search_trigger = true;
    // face indices change: wall history cannot be matched any more
    for (int k = 0; k < 2; ++k)
      c->wlists[k].n_rows = 0;
  });
}

int lethe_dem_set_floating_walls(lethe_dem_ctx *c, int32_t n, const double *point3, const double *normal3, const double *t_start,
                                 const double *t_end)
{
  return guarded(c, [&] {
    if (n < 0 || n > LETHE_DEM_MAX_FLOATING_WALLS)
      throw std::runtime_error("at most 9 floating walls");
    c->fw_host.n = n;
    for (int w = 0; w < n; ++w)
      {
        for (int d = 0; d < 3; ++d)
          {
            c->fw_host.point[w][d] = point3[3 * w + d];
            c->fw_host.normal[w][d] = normal3[3 * w + d];
          }
        c->fw_host.t0[w] = t_start[w];
        c->fw_host.t1[w] = t_end[w];
      }
    c->walls_dirty = true;
    c->contact_search_trigger = true;
  });
}

int lethe_dem_set_boundary_motion(lethe_dem_ctx *c, uint32_t boundary_id, const double tv[3], double speed, const double axis[3],
                                  const double point[3])
{
  return guarded(c, [&] {
    lethe_dem_ctx::Motion m;
    m.boundary_id = boundary_id;
    for (int d = 0; d < 3; ++d)
      {
        m.m.translational_velocity[d] = tv[d];
        m.m.rotational_vector[d] = axis[d];
        m.m.point_on_axis[d] = point[d];
      }
    m.m.rotational_speed = speed;
    bool found = false;
    for (auto &e : c->motions_host)
      if (e.boundary_id == boundary_id)
        {
          e = m;
          found = true;
        }
    if (!found)
      {
        if (c->motions_host.size() >= LETHE_DEM_MAX_BOUNDARY_MOTIONS)
          throw std::runtime_error("too many boundary motions");
        c->motions_host.push_back(m);
      }
    c->walls_dirty = true;
    // the motion table is uploaded with the wall table at a rebuild: ask for one so that the
    // change applies from the next step on (as set_walls / set_floating_walls do)
    c->contact_search_trigger = true;
  });
}

int lethe_dem_add_solid_surface(lethe_dem_ctx *c, uint32_t n_vertices, const double *vertices3, uint32_t n_triangles,
                                const uint32_t *triangles3, const double tv[3], const double av[3], const double center[3],
                                int32_t *solid_index)
{
  return guarded(c, [&] {
    if (c->n_solids >= uint32_t(MAX_SOLIDS))
      throw std::runtime_error("too many solid surfaces");
    for (uint64_t k = 0; k < 3ull * n_triangles; ++k)
      if (triangles3[k] >= n_vertices)
        throw std::runtime_error("triangle refers to a vertex outside the solid");
    const uint32_t v0 = uint32_t(c->solid_vertex_solid_host.size());
    c->solid_vertices_host.resize(3 * size_t(v0));
    c->solid_vertices_host.insert(c->solid_vertices_host.end(), vertices3, vertices3 + 3 * size_t(n_vertices));
    c->solid_vertex_solid_host.insert(c->solid_vertex_solid_host.end(), n_vertices, c->n_solids);
    for (uint64_t k = 0; k < 3ull * n_triangles; ++k)
      c->solid_tri_host.push_back(v0 + triangles3[k]);
    c->solid_tri_solid_host.insert(c->solid_tri_solid_host.end(), n_triangles, c->n_solids);
    if (c->solid_vertex_start.empty())
      {
        c->solid_vertex_start.push_back(0);
        c->solid_tri_start.push_back(0);
      }
    c->solid_vertex_start.push_back(uint32_t(c->solid_vertex_solid_host.size()));
    c->solid_tri_start.push_back(uint32_t(c->solid_tri_solid_host.size()));
    SolidMotionDev m;
    for (int d = 0; d < 3; ++d)
      {
        m.translational_velocity[d] = tv[d];
        m.angular_velocity[d] = av[d];
        m.center_of_rotation[d] = center[d];
      }
    c->solid_motion_host.push_back(m);
    if (solid_index)
      *solid_index = int32_t(c->n_solids);
    ++c->n_solids;
    c->solids_dirty = true;
    // DEMActionManager::set_solid_objects_enabled: the first search maps the solids
    c->solid_map_needed = true;
    c->contact_search_trigger = true;
  });
}

int lethe_dem_set_solid_motion(lethe_dem_ctx *c, int32_t solid_index, const double tv[3], const double av[3])
{
  return guarded(c, [&] {
    if (solid_index < 0 || uint32_t(solid_index) >= c->n_solids)
      throw std::runtime_error("no such solid");
    for (int d = 0; d < 3; ++d)
      {
        c->solid_motion_host[solid_index].translational_velocity[d] = tv[d];
        c->solid_motion_host[solid_index].angular_velocity[d] = av[d];
      }
    const uint32_t n_on_device = c->n_solid_vertices_dev ? c->solid_vertex_solid_host[c->n_solid_vertices_dev - 1] + 1 : 0;
    if (uint32_t(solid_index) < n_on_device)
      {
        // velocities only: the centre of rotation on the device has moved with the solid
        SolidMotionDev *dst = c->solid_motion.p + solid_index;
        CU_TRY(cudaMemcpyAsync(dst->translational_velocity, tv, 24, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaMemcpyAsync(dst->angular_velocity, av, 24, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
      }
  });
}

int lethe_dem_get_solid_vertices(lethe_dem_ctx *c, int32_t solid_index, uint32_t n_max, double *vertices3)
{
  return guarded(c, [&] {
    if (solid_index < 0 || uint32_t(solid_index) >= c->n_solids)
      throw std::runtime_error("no such solid");
    upload_solids(c);
    const uint32_t v0 = c->solid_vertex_start[solid_index], v1 = c->solid_vertex_start[solid_index + 1];
    const uint32_t n = std::min(n_max, v1 - v0);
    CU_TRY(cudaMemcpyAsync(vertices3, c->solid_vertices.p + 3 * size_t(v0), 3 * size_t(n) * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
  });
}

int lethe_dem_get_solid_contacts(lethe_dem_ctx *c, uint64_t n_max, uint64_t *n_out, uint32_t *particle_id, uint32_t *solid,
                                 uint32_t *triangle, double *tangential3)
{
  return guarded(c, [&] {
    CU_TRY(cudaStreamSynchronize(c->stream));
    SolidListBufs &l = c->slists[c->cur_list];
    const size_t n = c->n_solids ? l.n_rows : 0, E = c->n_solids ? l.n_entries : 0;
    *n_out = E;
    if (n_max < E || E == 0)
      return;
    std::vector<uint32_t> rs(n + 1), ent(E), ids(n);
    std::vector<double> hist(3 * E);
    CU_TRY(cudaMemcpy(rs.data(), l.row_start.p, (n + 1) * 4, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(ent.data(), l.entry.p, E * 4, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(hist.data(), l.hist.p, 3 * E * 8, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(ids.data(), c->st[c->cur].id.p, n * 4, cudaMemcpyDeviceToHost));
    struct Row
    {
      uint32_t pid, solid, tri;
      double h[3];
    };
    std::vector<Row> rows;
    rows.reserve(E);
    for (size_t q = 0; q < n; ++q)
      for (uint32_t e = rs[q]; e < rs[q + 1]; ++e)
        {
          const uint32_t t = ent[e] & SOLID_INDEX_MASK;
          const uint32_t sd = c->solid_tri_solid_host[t];
          Row r{ids[q], sd, t - c->solid_tri_start[sd], {0, 0, 0}};
          if (ent[e] & SOLID_HIST_BIT)
            for (int d = 0; d < 3; ++d)
              r.h[d] = hist[3 * size_t(e) + d];
          rows.push_back(r);
        }
    std::sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) {
      return a.pid != b.pid ? a.pid < b.pid : (a.solid != b.solid ? a.solid < b.solid : a.tri < b.tri);
    });
    for (size_t k = 0; k < rows.size(); ++k)
      {
        particle_id[k] = rows[k].pid;
        solid[k] = rows[k].solid;
        triangle[k] = rows[k].tri;
        for (int d = 0; d < 3; ++d)
          tangential3[3 * k + d] = rows[k].h[d];
      }
  });
}

int lethe_dem_step(lethe_dem_ctx *c, uint64_t n_steps)
{
  return guarded(c, [&] {
    for (uint64_t s = 0; s < n_steps; ++s)
      one_step(c);
  });
}

// DEMSolver::synchronize_velocities (dem.cc:719-745)
int lethe_dem_synchronize_velocities(lethe_dem_ctx *c)
{
  return guarded(c, [&] {
    contact_detection_and_search(c);
    launch_step_kernel(c, PHASE_END);
    if (!c->multi.enabled() && c->cfg.detection == LETHE_DETECTION_DYNAMIC)
      {
        // a caller that goes on stepping after the synchronisation (CFD-DEM does, every CFD time
        // step): the next iteration's displacement check sees the synchronised velocities
        launch_accumulate_displacement(c->st[c->cur].view().vel, c->disp.p, c->n_owned, c->cfg.dt,
                                       c->cfg.smallest_contact_search_criterion, c->flag_dev.p, c->d_flag,
                                       step_tag(c->iteration_number), c->stream);
        CU_TRY(cudaEventRecord(c->step_done[c->iteration_number & 1], c->stream));
      }
    c->contact_search_trigger = c->asc_reset;
    c->asc_reset = false;
    c->clear_history_trigger = false;
  });
}


extern "C" int lethe_dem_set_particles_cfd(lethe_dem_ctx *c, uint64_t n, const uint32_t *id, const double *x3, const double *props23)
{
  std::vector<double> p9, f3, t3;
  split_cfd_rows(n, props23, &p9, f3, t3);
  const int rc = lethe_dem_set_particles(c, n, id, x3, p9.data());
  return rc ? rc : lethe_dem_set_external_loads(c, n, id, f3.data(), t3.data());
}

extern "C" int lethe_dem_update_loads_cfd(lethe_dem_ctx *c, uint64_t n, const uint32_t *id, const double *props23)
{
  std::vector<double> f3, t3;
  split_cfd_rows(n, props23, nullptr, f3, t3);
  return lethe_dem_set_external_loads(c, n, id, f3.data(), t3.data());
}

extern "C" int lethe_dem_get_particles_cfd(lethe_dem_ctx *c, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *x3, double *props23)
{
  std::vector<double> p9(9 * n_max);
  const int rc = lethe_dem_get_particles(c, n_max, n_out, id, x3, p9.data());
  if (rc)
    return rc;
  for (uint64_t k = 0; k < *n_out; ++k)
    std::memcpy(props23 + 23 * k, p9.data() + 9 * k, 72);
  return 0;
}

int lethe_dem_enable_heat_transfer(lethe_dem_ctx *c, const lethe_dem_thermal_properties *pr)
{
  return guarded(c, [&] {
    if (c->multi.enabled())
      throw std::runtime_error("heat transfer runs on a single GPU (ghost temperatures are not exchanged)");
    const int n = c->cfg.n_types;
    ThermalTables &t = c->thermal_tables;
    std::memset(&t, 0, sizeof(t));
    t.conductivity_gas = pr->thermal_conductivity_gas;
    // set_multiphysic_properties (particle_particle_contact_force.h:1755-1826)
    for (int i = 0; i < n; ++i)
      {
        const double real_youngs_modulus_i = pr->real_youngs_modulus[i], poisson_ratio_i = c->cfg.poisson[i];
        const double surface_roughness_i = pr->surface_roughness[i], surface_slope_i = pr->surface_slope[i];
        const double microhardness_i = pr->microhardness[i], thermal_accommodation_i = pr->thermal_accommodation[i];
        t.conductivity[i] = pr->thermal_conductivity[i];
        for (int j = 0; j < n; ++j)
          {
            const int k = i * n + j;
            const double real_youngs_modulus_j = pr->real_youngs_modulus[j], poisson_ratio_j = c->cfg.poisson[j];
            const double surface_roughness_j = pr->surface_roughness[j], surface_slope_j = pr->surface_slope[j];
            const double microhardness_j = pr->microhardness[j], thermal_accommodation_j = pr->thermal_accommodation[j];
            t.real_E[k] = (real_youngs_modulus_i * real_youngs_modulus_j) /
                          ((real_youngs_modulus_j * (1.0 - poisson_ratio_i * poisson_ratio_i)) +
                           (real_youngs_modulus_i * (1.0 - poisson_ratio_j * poisson_ratio_j)) + DBL_MIN);
            t.roughness[k] = std::sqrt(surface_roughness_i * surface_roughness_i + surface_roughness_j * surface_roughness_j);
            t.slope[k] = std::sqrt(surface_slope_i * surface_slope_i + surface_slope_j * surface_slope_j);
            t.microhardness[k] = (2 * microhardness_i * microhardness_j / (microhardness_i + microhardness_j + DBL_MIN));
            t.gas_m[k] = ((2. - thermal_accommodation_i) / thermal_accommodation_i + (2. - thermal_accommodation_j) / thermal_accommodation_j) *
                         (2. * pr->specific_heats_ratio_gas) / (1. + pr->specific_heats_ratio_gas) * pr->molecular_mean_free_path_gas /
                         (pr->dynamic_viscosity_gas * pr->specific_heat_gas / pr->thermal_conductivity_gas);
          }
      }
    c->thermal_enabled = true;
  });
}

int lethe_dem_set_temperatures(lethe_dem_ctx *c, uint64_t n, const uint32_t *id, const double *temperature, const double *specific_heat)
{
  return guarded(c, [&] {
    cudaStream_t s = c->stream;
    CU_TRY(cudaStreamSynchronize(s));
    size_t need = c->thermal_size;
    for (uint64_t k = 0; k < n; ++k)
      need = std::max<size_t>(need, size_t(id[k]) + 1);
    std::vector<double> ht(need, 0.), hc(need, 1.);
    if (c->thermal_size)
      {
        CU_TRY(cudaMemcpy(ht.data(), c->temperature.p, c->thermal_size * 8, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(hc.data(), c->specific_heat.p, c->thermal_size * 8, cudaMemcpyDeviceToHost));
      }
    for (uint64_t k = 0; k < n; ++k)
      {
        ht[id[k]] = temperature[k];
        hc[id[k]] = specific_heat[k];
      }
    c->temperature.ensure(need);
    c->specific_heat.ensure(need);
    CU_TRY(cudaMemcpy(c->temperature.p, ht.data(), need * 8, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(c->specific_heat.p, hc.data(), need * 8, cudaMemcpyHostToDevice));
    c->thermal_size = need;
  });
}

int lethe_dem_get_temperatures(lethe_dem_ctx *c, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *temperature, double *heat_transfer_rate)
{
  return guarded(c, [&] {
    cudaStream_t s = c->stream;
    CU_TRY(cudaStreamSynchronize(s));
    const size_t n = c->n_owned;
    std::vector<uint32_t> hid(n);
    std::vector<double> ht(c->thermal_size), hr(n, 0.);
    if (n)
      CU_TRY(cudaMemcpy(hid.data(), c->st[c->cur].id.p, n * 4, cudaMemcpyDeviceToHost));
    if (c->thermal_size)
      CU_TRY(cudaMemcpy(ht.data(), c->temperature.p, c->thermal_size * 8, cudaMemcpyDeviceToHost));
    if (n && c->heat_rate.cap >= n)
      CU_TRY(cudaMemcpy(hr.data(), c->heat_rate.p, n * 8, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return hid[a] < hid[b]; });
    const uint64_t m = std::min<uint64_t>(n_max, n);
    for (uint64_t k = 0; k < m; ++k)
      {
        const uint32_t q = order[k];
        id[k] = hid[q];
        temperature[k] = hid[q] < ht.size() ? ht[hid[q]] : 0.;
        heat_transfer_rate[k] = hr[q];
      }
    *n_out = m;
  });
}

int lethe_dem_set_time(lethe_dem_ctx *c, uint64_t iteration_number, double current_time)
{
  return guarded(c, [&] {
    c->iteration_number = iteration_number;
    c->current_time = current_time;
    c->contact_search_trigger = true; // restart_simulation(): search, with every history cleared
    c->clear_history_trigger = true;
  });
}

int lethe_dem_restart_integration(lethe_dem_ctx *c)
{
  c->open_next_step = true;
  return 0;
}

int lethe_dem_set_external_loads(lethe_dem_ctx *c, uint64_t n, const uint32_t *id, const double *force3, const double *torque3)
{
  return guarded(c, [&] {
    cudaStream_t s = c->stream;
    if (n == 0)
      {
        c->ext_enabled = false; // all loads cleared
        if (c->ext_size)
          {
            CU_TRY(cudaMemsetAsync(c->ext_force.p, 0, 3 * size_t(c->ext_size) * 8, s));
            CU_TRY(cudaMemsetAsync(c->ext_torque.p, 0, 3 * size_t(c->ext_size) * 8, s));
          }
        return;
      }
    uint32_t max_id = 0;
    for (uint64_t k = 0; k < n; ++k)
      max_id = std::max(max_id, id[k]);
    if (size_t(max_id) + 1 > c->ext_size)
      {
        const size_t old = c->ext_size, want = std::max<size_t>(size_t(max_id) + 1, c->slot_map_size);
        c->ext_force.ensure(3 * want, 3 * old, s, 1.0);
        c->ext_torque.ensure(3 * want, 3 * old, s, 1.0);
        CU_TRY(cudaMemsetAsync(c->ext_force.p + 3 * old, 0, 3 * (want - old) * 8, s));
        CU_TRY(cudaMemsetAsync(c->ext_torque.p + 3 * old, 0, 3 * (want - old) * 8, s));
        c->ext_size = uint32_t(want);
      }
    c->stage_ids.ensure(n);
    c->stage_x.ensure(3 * n);
    c->stage_p.ensure(9 * n);
    CU_TRY(cudaMemcpyAsync(c->stage_ids.p, id, n * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(c->stage_x.p, force3, 3 * n * 8, cudaMemcpyHostToDevice, s));
    if (torque3)
      CU_TRY(cudaMemcpyAsync(c->stage_p.p, torque3, 3 * n * 8, cudaMemcpyHostToDevice, s));
    launch_scatter_external_loads(c->stage_ids.p, c->stage_x.p, torque3 ? c->stage_p.p : nullptr, uint32_t(n), c->ext_force.p,
                                  c->ext_torque.p, c->ext_size, s);
    CU_TRY(cudaStreamSynchronize(s)); // the caller's buffers are free again
    c->ext_enabled = true;
  });
}

int lethe_dem_force_contact_search(lethe_dem_ctx *c, int clear_tangential_displacement)
{
  c->contact_search_trigger = true;
  if (clear_tangential_displacement)
    c->clear_history_trigger = true;
  return 0;
}

int lethe_dem_step_host_state(lethe_dem_ctx *c, uint64_t n_steps, uint64_t n, const uint32_t *id, double *state9)
{
  return guarded(c, [&] {
    cudaStream_t s = c->stream;
    if (n)
      {
        if (id)
          {
            c->host_row_ids.ensure(n);
            CU_TRY(cudaMemcpyAsync(c->host_row_ids.p, id, n * 4, cudaMemcpyHostToDevice, s));
            c->host_row_ids_n = n;
            ++c->host_row_ids_version;
          }
        else if (c->host_row_ids_n != n)
          throw std::runtime_error("step_host_state: id == NULL reuses the id table of the previous call, which had a different row count");
        if (host_pipe_eligible(c, n_steps, n))
          {
            auto &hp = c->host_pipe;
            if (!(hp.rebuild_gen == c->n_rebuilds && hp.ids_version == c->host_row_ids_version && hp.n_rows == n && hp.n_owned == c->n_owned))
              host_pipe_plan(c, n);
            if (hp.valid)
              {
                step_host_state_streamed(c, n, state9);
                return;
              }
          }
        c->stage_p.ensure(9 * n);
        CU_TRY(cudaMemcpyAsync(c->stage_p.p, state9, 9 * n * 8, cudaMemcpyHostToDevice, s));
        launch_update_state_rows(c->host_row_ids.p, c->stage_p.p, uint32_t(n), c->slot_of_id.p, c->slot_map_size, c->st[c->cur].view(), s);
      }
    for (uint64_t k = 0; k < n_steps; ++k)
      one_step(c);
    if (n)
      {
        launch_pack_state_rows(c->host_row_ids.p, uint32_t(n), c->slot_of_id.p, c->slot_map_size, c->st[c->cur].view(), c->stage_p.p, s);
        CU_TRY(cudaMemcpyAsync(state9, c->stage_p.p, 9 * n * 8, cudaMemcpyDeviceToHost, s));
      }
    CU_TRY(cudaStreamSynchronize(s));
  });
}

namespace
{
  // perm[k] = slot of row k of the transfer order (by cell layer along the axis with the most layers of this rank's part
  // of the grid, cell-sorted inside a layer), left in c->host_pipe.row_of_slot's sibling buffer `perm`
  void transfer_order(Ctx *c, DevBuf<uint32_t> &perm)
  {
    const GridDesc &g = c->grid;
    int axis = 0, best = 0;
    for (int d = 0; d < 3; ++d)
      {
        const int layers = (d == g.slab_axis && g.slab_lo >= 0) ? g.slab_hi - g.slab_lo : g.n[d];
        if (layers > best)
          best = layers, axis = d;
      }
    perm.ensure(std::max<size_t>(c->n_owned, 1));
    transfer_order_perm(c->st[c->cur].cell_reg.p, g, axis, c->n_owned, perm.p, c->stream);
  }
} // namespace

int lethe_dem_get_transfer_order(lethe_dem_ctx *c, uint64_t n_max, uint64_t *n_out, uint32_t *id)
{
  return guarded(c, [&] {
    const size_t n = c->n_owned;
    *n_out = n;
    if (n_max < n || n == 0)
      return;
    DevBuf<uint32_t> perm;
    transfer_order(c, perm);
    c->stage_ids.ensure(n);
    launch_pack_state_rows_perm(perm.p, uint32_t(n), c->st[c->cur].view(), c->st[c->cur].id.p, c->stage_ids.p, nullptr, c->stream);
    CU_TRY(cudaMemcpyAsync(id, c->stage_ids.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
  });
}

int lethe_dem_get_state_rows(lethe_dem_ctx *c, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *state9)
{
  return guarded(c, [&] {
    const size_t n = c->n_owned;
    *n_out = n;
    if (n_max < n || n == 0)
      return;
    DevBuf<uint32_t> perm;
    transfer_order(c, perm);
    c->stage_ids.ensure(n);
    c->stage_p.ensure(9 * n);
    launch_pack_state_rows_perm(perm.p, uint32_t(n), c->st[c->cur].view(), c->st[c->cur].id.p, c->stage_ids.p, c->stage_p.p, c->stream);
    CU_TRY(cudaMemcpyAsync(id, c->stage_ids.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(state9, c->stage_p.p, 9 * n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
  });
}

int lethe_dem_host_pipeline_stats(lethe_dem_ctx *c, uint64_t *n_streamed_calls, uint64_t *n_plans, uint64_t *n_direct_calls)
{
  if (n_direct_calls)
    *n_direct_calls = c->host_pipe.n_zero_copy_calls;
  if (n_streamed_calls)
    *n_streamed_calls = c->host_pipe.n_calls;
  if (n_plans)
    *n_plans = c->host_pipe.n_plans;
  return 0;
}

int lethe_dem_step_host(lethe_dem_ctx *c, uint64_t n_steps, uint64_t n, const uint32_t *id, double *x3, double *props9)
{
  return guarded(c, [&] {
    cudaStream_t s = c->stream;
    if (n)
      {
        c->stage_ids.ensure(n);
        c->stage_x.ensure(3 * n);
        c->stage_p.ensure(9 * n);
        CU_TRY(cudaMemcpyAsync(c->stage_ids.p, id, n * 4, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(c->stage_x.p, x3, 3 * n * 8, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(c->stage_p.p, props9, 9 * n * 8, cudaMemcpyHostToDevice, s));
        launch_update_from_host_rows(c->stage_ids.p, c->stage_x.p, c->stage_p.p, uint32_t(n), c->slot_of_id.p, c->slot_map_size,
                                     c->st[c->cur].view(), s);
      }
    for (uint64_t k = 0; k < n_steps; ++k)
      one_step(c);
    if (n)
      {
        launch_pack_host_rows(c->stage_ids.p, uint32_t(n), c->slot_of_id.p, c->slot_map_size, c->st[c->cur].view(), c->stage_x.p,
                              c->stage_p.p, s);
        CU_TRY(cudaMemcpyAsync(x3, c->stage_x.p, 3 * n * 8, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(props9, c->stage_p.p, 9 * n * 8, cudaMemcpyDeviceToHost, s));
      }
    CU_TRY(cudaStreamSynchronize(s));
  });
}

int lethe_dem_get_pairs(lethe_dem_ctx *c, uint64_t n_max, uint64_t *n_out, uint32_t *i_id, uint32_t *j_id, double *tangential3)
{
  return guarded(c, [&] {
    CU_TRY(cudaStreamSynchronize(c->stream));
    ListBufs &l = c->lists[c->cur_list];
    const size_t n = l.n_rows, E = l.n_entries;
    const size_t n_all = size_t(c->n_owned) + c->n_ghost;
    std::vector<uint32_t> rs(n + 1, 0), col(E), ids(n_all);
    std::vector<double> hist(4 * E); // 32-byte rows
    if (n)
      {
        CU_TRY(cudaMemcpy(rs.data(), l.row_start.p, (n + 1) * 4, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(ids.data(), c->st[c->cur].id.p, n_all * 4, cudaMemcpyDeviceToHost));
      }
    if (E)
      {
        CU_TRY(cudaMemcpy(col.data(), l.col.p, E * 4, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(hist.data(), l.hist.p, 4 * E * 8, cudaMemcpyDeviceToHost));
      }
    struct P
    {
      uint32_t i, j;
      double t[3];
    };
    std::vector<P> pairs;
    pairs.reserve(E / 2 + 1);
    for (size_t q = 0; q < n; ++q)
      for (uint32_t e = rs[q]; e < rs[q + 1]; ++e)
        {
          const uint32_t r = col[e] & COL_INDEX_MASK;
          const uint32_t a = ids[q], b = ids[r];
          // each unordered pair is listed from both rows; report the copy of the lower id
          // (ghost partners have no row here, so always report those)
          if (a < b || (r >= n && a > b))
            {
              P p;
              const bool has = (col[e] & COL_HIST_BIT) != 0;
              const double sgn = a < b ? 1.0 : -1.0;
              p.i = std::min(a, b);
              p.j = std::max(a, b);
              for (int d = 0; d < 3; ++d)
                p.t[d] = has ? sgn * hist[4 * size_t(e) + d] : 0.0;
              pairs.push_back(p);
            }
        }
    std::sort(pairs.begin(), pairs.end(), [](const P &a, const P &b) { return a.i != b.i ? a.i < b.i : a.j < b.j; });
    *n_out = pairs.size();
    const uint64_t m = std::min<uint64_t>(n_max, pairs.size());
    for (uint64_t k = 0; k < m; ++k)
      {
        i_id[k] = pairs[k].i;
        j_id[k] = pairs[k].j;
        if (tangential3)
          for (int d = 0; d < 3; ++d)
            tangential3[3 * k + d] = pairs[k].t[d];
      }
  });
}

int lethe_dem_get_wall_contacts(lethe_dem_ctx *c, uint64_t n_max, uint64_t *n_out, uint32_t *particle_id, uint32_t *face_id,
                                double *tangential3)
{
  return guarded(c, [&] {
    CU_TRY(cudaStreamSynchronize(c->stream));
    WallListBufs &l = c->wlists[c->cur_list];
    const size_t n = l.n_rows, W = l.n_entries;
    std::vector<uint32_t> rs(n + 1, 0), ent(W), ids(n);
    std::vector<double> hist(3 * W);
    if (n)
      {
        CU_TRY(cudaMemcpy(rs.data(), l.row_start.p, (n + 1) * 4, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(ids.data(), c->st[c->cur].id.p, n * 4, cudaMemcpyDeviceToHost));
      }
    if (W)
      {
        CU_TRY(cudaMemcpy(ent.data(), l.entry.p, W * 4, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(hist.data(), l.hist.p, 3 * W * 8, cudaMemcpyDeviceToHost));
      }
    struct P
    {
      uint32_t p, f;
      double t[3];
    };
    std::vector<P> all;
    for (size_t q = 0; q < n; ++q)
      for (uint32_t e = rs[q]; e < rs[q + 1]; ++e)
        {
          P p;
          p.p = ids[q];
          const uint32_t idx = ent[e] & WALL_INDEX_MASK;
          p.f = (ent[e] & WALL_FLOATING_BIT) ? (idx | 0x80000000u) : c->face_gid_host[idx];
          for (int d = 0; d < 3; ++d)
            p.t[d] = (ent[e] & WALL_HIST_BIT) ? hist[3 * size_t(e) + d] : 0.0;
          all.push_back(p);
        }
    std::sort(all.begin(), all.end(), [](const P &a, const P &b) { return a.p != b.p ? a.p < b.p : a.f < b.f; });
    *n_out = all.size();
    const uint64_t m = std::min<uint64_t>(n_max, all.size());
    for (uint64_t k = 0; k < m; ++k)
      {
        particle_id[k] = all[k].p;
        face_id[k] = all[k].f;
        if (tangential3)
          for (int d = 0; d < 3; ++d)
            tangential3[3 * k + d] = all[k].t[d];
      }
  });
}

int lethe_dem_get_forces(lethe_dem_ctx *c, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *force3, double *torque3)
{
  return guarded(c, [&] {
    if (!c->cfg.store_forces)
      throw std::runtime_error("store_forces not enabled in the config");
    CU_TRY(cudaStreamSynchronize(c->stream));
    const size_t n = c->n_owned;
    std::vector<uint32_t> ids(n);
    std::vector<double> f(3 * n), t(3 * n);
    if (n && c->force_out.p)
      {
        CU_TRY(cudaMemcpy(ids.data(), c->st[c->cur].id.p, n * 4, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(f.data(), c->force_out.p, 3 * n * 8, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(t.data(), c->torque_out.p, 3 * n * 8, cudaMemcpyDeviceToHost));
      }
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ids[a] < ids[b]; });
    const uint64_t m = std::min<uint64_t>(n_max, n);
    for (uint64_t k = 0; k < m; ++k)
      {
        const uint32_t q = order[k];
        id[k] = ids[q];
        for (int d = 0; d < 3; ++d)
          {
            force3[3 * k + d] = f[3 * size_t(q) + d];
            torque3[3 * k + d] = t[3 * size_t(q) + d];
          }
      }
    *n_out = m;
  });
}

int lethe_dem_get_stats(lethe_dem_ctx *c, lethe_dem_stats *st)
{
  return guarded(c, [&] {
    std::memset(st, 0, sizeof(*st));
    st->n_particles = c->n_owned;
    st->n_rebuilds = c->n_rebuilds;
    st->n_migrated = c->n_migrated;
    st->n_steps = c->iteration_number;
    st->n_pair_entries = c->lists[c->cur_list].n_entries / 2;
    st->n_wall_entries = c->wlists[c->cur_list].n_entries;
    if (c->multi.enabled())
      st->n_pair_entries = c->lists[c->cur_list].n_entries; // full-list entries (owned rows), not halved
    if ((c->cfg.store_forces || c->count_touching) && c->touching.p)
      {
        unsigned long long t = 0;
        CU_TRY(cudaMemcpyAsync(&t, c->touching.p, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        st->n_pairs_touching = t / 2;
      }
    const uint32_t n = c->n_owned;
    if (!n)
      return;
    const uint32_t nb = std::min<uint32_t>((n + STATS_BLOCK - 1) / STATS_BLOCK, 148 * 8);
    c->stats_partials.ensure(nb);
    launch_stats(c->st[c->cur].view(), n, c->cfg.moi_override, c->stats_partials.p, nb, c->stream);
    std::vector<StatsPartial> h(nb);
    CU_TRY(cudaMemcpyAsync(h.data(), c->stats_partials.p, nb * sizeof(StatsPartial), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    StatsPartial a = h[0];
    for (uint32_t b = 1; b < nb; ++b)
      {
        a.vmin = std::min(a.vmin, h[b].vmin);
        a.vmax = std::max(a.vmax, h[b].vmax);
        a.vsum += h[b].vsum;
        a.wmin = std::min(a.wmin, h[b].wmin);
        a.wmax = std::max(a.wmax, h[b].wmax);
        a.wsum += h[b].wsum;
        a.ktmin = std::min(a.ktmin, h[b].ktmin);
        a.ktmax = std::max(a.ktmax, h[b].ktmax);
        a.ktsum += h[b].ktsum;
        a.krmin = std::min(a.krmin, h[b].krmin);
        a.krmax = std::max(a.krmax, h[b].krmax);
        a.krsum += h[b].krsum;
      }
    st->v_min = a.vmin;
    st->v_max = a.vmax;
    st->v_sum = a.vsum;
    st->omega_min = a.wmin;
    st->omega_max = a.wmax;
    st->omega_sum = a.wsum;
    st->ke_trans_min = a.ktmin;
    st->ke_trans_max = a.ktmax;
    st->ke_trans_sum = a.ktsum;
    st->ke_rot_min = a.krmin;
    st->ke_rot_max = a.krmax;
    st->ke_rot_sum = a.krsum;
  });
}

int lethe_dem_enable_timers(lethe_dem_ctx *c, int flags)
{
  return guarded(c, [&] {
    resolve_timers(c);
    c->timers_enabled = (flags & 1) != 0;
    c->count_touching = (flags & 2) != 0;
  });
}

int lethe_dem_event_record(lethe_dem_ctx *c, int which)
{
  return guarded(c, [&] {
    if (which < 0 || which > 1)
      throw std::runtime_error("event index must be 0 or 1");
    if (!c->region_ev[which])
      CU_TRY(cudaEventCreate(&c->region_ev[which]));
    CU_TRY(cudaEventRecord(c->region_ev[which], c->stream));
  });
}

int lethe_dem_event_elapsed(lethe_dem_ctx *c, double *ms)
{
  return guarded(c, [&] {
    if (!c->region_ev[0] || !c->region_ev[1])
      throw std::runtime_error("record both events first");
    CU_TRY(cudaEventSynchronize(c->region_ev[1]));
    float f = 0;
    CU_TRY(cudaEventElapsedTime(&f, c->region_ev[0], c->region_ev[1]));
    *ms = f;
  });
}

int lethe_dem_kernel_launches(lethe_dem_ctx *, uint64_t *n_launches)
{
  *n_launches = dem::launch_count();
  return 0;
}

int lethe_dem_get_timers(lethe_dem_ctx *c, int reset, double *step_kernel_ms, uint64_t *step_kernel_launches, double *rebuild_ms,
                         uint64_t *rebuild_launches)
{
  return guarded(c, [&] {
    resolve_timers(c);
    *step_kernel_ms = c->step_ms;
    *step_kernel_launches = c->step_launches;
    *rebuild_ms = c->rebuild_ms;
    *rebuild_launches = c->rebuild_launches;
    if (reset)
      {
        c->step_ms = c->rebuild_ms = 0;
        c->step_launches = c->rebuild_launches = 0;
      }
  });
}

int lethe_dem_get_mobility_status(lethe_dem_ctx *c, uint64_t n_cells, int32_t *status)
{
  return guarded(c, [&] {
    if (n_cells != uint64_t(c->grid.n_cells))
      throw std::runtime_error("get_mobility_status: n_cells does not match the grid");
    if (!c->asc_enabled || c->asc_cell_status.cap < n_cells)
      {
        for (uint64_t k = 0; k < n_cells; ++k)
          status[k] = LETHE_MOBILITY_MOBILE;
        return;
      }
    std::vector<uint8_t> h(n_cells);
    CU_TRY(cudaMemcpyAsync(h.data(), c->asc_cell_status.p, n_cells, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    for (uint64_t k = 0; k < n_cells; ++k)
      status[k] = h[k];
  });
}

int lethe_dem_set_load_balancing(lethe_dem_ctx *c, int method, double threshold, int frequency)
{
  return guarded(c, [&] {
    if (method < LETHE_LOAD_BALANCE_NONE || method > LETHE_LOAD_BALANCE_DYNAMIC_WITH_SPARSE_CONTACTS)
      throw std::runtime_error("unknown load balance method (none|once|frequent|dynamic|dynamic_with_sparse_contacts)");
    if (method == LETHE_LOAD_BALANCE_DYNAMIC_WITH_SPARSE_CONTACTS && !c->asc_enabled)
      throw std::runtime_error("Invalid contact detection method: adaptive sparse contacts is not enabled while dynamic_with_sparse_contacts "
                               "is selected, use dynamic instead"); // parameters_lagrangian.cc:1160-1164
    c->lb_method = method;
    c->lb_threshold = threshold;
    c->lb_frequency = frequency;
  });
}

int lethe_dem_set_load_balancing_weights(lethe_dem_ctx *c, double particle_weight, double cell_weight, double active_weight_factor,
                                         double inactive_weight_factor)
{
  return guarded(c, [&] {
    c->lb_particle_weight = particle_weight;
    c->lb_cell_weight = cell_weight;
    c->lb_active_factor = active_weight_factor;
    c->lb_inactive_factor = inactive_weight_factor;
  });
}

int lethe_dem_get_slab(lethe_dem_ctx *c, int32_t *lo, int32_t *hi, uint64_t *n_repartitions)
{
  return guarded(c, [&] {
    const bool slab = c->grid.slab_axis >= 0;
    *lo = slab ? c->grid.slab_lo : 0;
    *hi = slab ? c->grid.slab_hi : c->grid.n[0];
    *n_repartitions = c->n_recuts;
  });
}

int lethe_dem_balanced_cuts(int32_t n_layers, const uint64_t *histogram, int32_t world, const int32_t *cuts, int32_t max_shift,
                            int32_t min_width, int32_t *new_cuts)
{
  if (n_layers < 1 || world < 1 || !histogram || !cuts || !new_cuts || n_layers < world * min_width)
    return -1;
  balanced_cuts(n_layers, histogram, world, cuts, max_shift, min_width, new_cuts);
  return 0;
}

int lethe_dem_nccl_unique_id(uint8_t id[LETHE_DEM_NCCL_ID_BYTES]) { return dem::MultiGpu::unique_id(id); }

int lethe_dem_comm_init(lethe_dem_ctx *c, int rank, int world_size, const uint8_t id[LETHE_DEM_NCCL_ID_BYTES])
{
  return guarded(c, [&] { c->multi.init(c, rank, world_size, id); });
}

} // extern "C"
