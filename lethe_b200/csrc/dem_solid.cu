// dem_solid.cu — solid surfaces (triangle-mesh walls): motion, candidate rows, contact force.
// See dem_solid.cuh for the reference map. Arithmetic rules as in dem_physics.cuh (FP64,
// -fmad=false, reference evaluation order) so that the CPU oracle and these kernels agree to
// the last bits of every branch decision of the triangle projection.
#include <algorithm>
#include <cmath>
#include <vector>

#include "dem_solid.cuh"

namespace dem
{
  namespace
  {
    enum TriangleContact
    {
      TC_FACE = 0,
      TC_EDGE = 1,
      TC_VERTEX = 2
    };

    // Eberly's closest point on a triangle as LetheGridTools evaluates it
    // (lethe_grid_tools.cc:1277-1434 / :1565-1697), quirks included (region 4: t = e / c).
    __host__ __device__ inline void closest_point_parameters(double a, double b, double c, double d, double e, double det, double &s,
                                                             double &t, int &indicator)
    {
      s = b * e - c * d;
      t = b * d - a * e;
      if (s + t <= det)
        {
          if (s < 0)
            {
              if (t < 0)
                {
                  indicator = TC_VERTEX; // region 4
                  if (d < 0)
                    {
                      t = 0;
                      if (-d >= a)
                        s = 1;
                      else
                        s = -d / a;
                    }
                  else
                    {
                      s = 0;
                      if (e >= 0)
                        t = 0;
                      else if (-e >= c)
                        t = 1;
                      else
                        t = e / c;
                    }
                }
              else
                {
                  indicator = TC_EDGE; // region 3
                  s = 0;
                  if (e >= 0)
                    t = 0;
                  else if (-e >= c)
                    t = 1;
                  else
                    t = -e / c;
                }
            }
          else if (t < 0)
            {
              indicator = TC_EDGE; // region 5
              t = 0;
              if (d >= 0)
                s = 0;
              else if (-d >= a)
                s = 1;
              else
                s = -d / a;
            }
          else
            {
              indicator = TC_FACE; // region 0
              const double inv_det = 1. / det;
              s *= inv_det;
              t *= inv_det;
            }
        }
      else
        {
          if (s < 0)
            {
              indicator = TC_VERTEX; // region 2
              const double tmp0 = b + d;
              const double tmp1 = c + e;
              if (tmp1 > tmp0)
                {
                  const double numer = tmp1 - tmp0;
                  const double denom = a - 2 * b + c;
                  if (numer >= denom)
                    s = 1;
                  else
                    s = numer / denom;
                  t = 1 - s;
                }
              else
                {
                  s = 0;
                  if (tmp1 <= 0)
                    t = 1;
                  else if (e >= 0)
                    t = 0;
                  else
                    t = -e / c;
                }
            }
          else if (t < 0)
            {
              indicator = TC_VERTEX; // region 6
              const double tmp0 = b + e;
              const double tmp1 = a + d;
              if (tmp1 > tmp0)
                {
                  const double numer = tmp1 - tmp0;
                  const double denom = a - 2 * b + c;
                  if (numer >= denom)
                    t = 1;
                  else
                    t = numer / denom;
                  s = 1 - t;
                }
              else
                {
                  t = 0;
                  if (tmp1 <= 0)
                    s = 1;
                  else if (d >= 0)
                    s = 0;
                  else
                    s = -d / a;
                }
            }
          else
            {
              indicator = TC_EDGE; // region 1
              const double numer = (c + e) - (b + d);
              if (numer <= 0)
                s = 0;
              else
                {
                  const double denom = a - 2 * b + c;
                  if (numer >= denom)
                    s = 1;
                  else
                    s = numer / denom;
                }
              t = 1 - s;
            }
        }
    }

    // ---- motion: SerialSolid::move_solid_triangulation (serial_solid.cc:333-410) ----
    __global__ void __launch_bounds__(128) k_move_solid_vertices(const __grid_constant__ SolidMoveParams P)
    {
      if (P.spec_check)
        {
          const uint32_t f = *reinterpret_cast<const volatile uint32_t *>(P.flag_check);
          if (f != 0u && f != P.flag_tag)
            return; // void launch: an earlier step asked for a new list
        }
      const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
      if (v >= P.s.n_vertices)
        return;
      const SolidMotionDev &m = P.s.motion[P.s.vertex_solid[v]];
      double *x = P.s.vertices + 3 * size_t(v);
      double *dsp = P.s.displacement + 3 * size_t(v);
      const vec3 center = v3(m.center_of_rotation[0], m.center_of_rotation[1], m.center_of_rotation[2]);
      const vec3 distance_vector = v3(x[0], x[1], x[2]) - center;
      vec3 local_velocity = v3(m.translational_velocity[0], m.translational_velocity[1], m.translational_velocity[2]);
      local_velocity = local_velocity + cross(v3(m.angular_velocity[0], m.angular_velocity[1], m.angular_velocity[2]), distance_vector);
      const vec3 vertex_displacement = P.dt * local_velocity;
      x[0] = x[0] + vertex_displacement.x;
      x[1] = x[1] + vertex_displacement.y;
      x[2] = x[2] + vertex_displacement.z;
      const vec3 nd = v3(dsp[0] + vertex_displacement.x, dsp[1] + vertex_displacement.y, dsp[2] + vertex_displacement.z);
      dsp[0] = nd.x;
      dsp[1] = nd.y;
      dsp[2] = nd.z;
      // find_floating_mesh_mapping_step, evaluated for the NEXT step's check
      if (fmax(fabs(nd.x), fmax(fabs(nd.y), fabs(nd.z))) > P.criterion)
        {
          *reinterpret_cast<volatile uint32_t *>(P.remap_host) = 1u;
          if (*reinterpret_cast<volatile uint32_t *>(P.flag_local) != P.flag_tag)
            {
              *reinterpret_cast<volatile uint32_t *>(P.flag_local) = P.flag_tag;
              if (P.flag_host)
                *reinterpret_cast<volatile uint32_t *>(P.flag_host) = P.flag_tag;
            }
        }
    }
    // center_of_rotation += translational_velocity * dt, after every vertex has used the old one
    __global__ void k_move_solid_centers(const __grid_constant__ SolidMoveParams P)
    {
      if (P.spec_check)
        {
          const uint32_t f = *reinterpret_cast<const volatile uint32_t *>(P.flag_check);
          // the vertex kernel of this very step may have raised the flag with this step's tag
          if (f != 0u && f != P.flag_tag)
            return;
        }
      const uint32_t k = threadIdx.x;
      if (k >= P.s.n_solids)
        return;
      SolidMotionDev &m = P.s.motion[k];
      for (int d = 0; d < 3; ++d)
        m.center_of_rotation[d] = m.center_of_rotation[d] + m.translational_velocity[d] * P.dt;
    }

    // ---- candidate rows: a particle's row is its cell's triangle list + history ----
    __global__ void __launch_bounds__(256) k_count_solid_rows(const __grid_constant__ SolidBuildParams P)
    {
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= P.n_rows)
        return;
      const int c = P.cell_reg[q];
      // adaptive sparse contacts: particles of mobile cells only (particle_wall_broad_search.cc:391-398)
      const bool listed = c >= 0 && (!P.mobility || P.mobility[c] == LETHE_MOBILITY_MOBILE);
      P.counts[q] = listed ? P.cell_tri_start[c + 1] - P.cell_tri_start[c] : 0u;
    }
    __global__ void __launch_bounds__(256) k_fill_solid_rows(const __grid_constant__ SolidBuildParams P)
    {
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= P.n_rows)
        return;
      const int c = P.cell_reg[q];
      uint32_t e = P.new_list.row_start[q];
      const uint32_t e_end = P.new_list.row_start[q + 1];
      const uint32_t old_q = P.old_of_new ? P.old_of_new[q] : 0xffffffffu;
      const bool have_old = !P.clear_history && old_q != 0xffffffffu && old_q < P.n_old_rows;
      uint32_t o0 = 0, o1 = 0;
      if (have_old)
        {
          o0 = P.old_list.row_start[old_q];
          o1 = P.old_list.row_start[old_q + 1];
        }
      // both rows are sorted by triangle index: merge
      uint32_t eo = o0;
      for (uint32_t k = c >= 0 ? P.cell_tri_start[c] : 0u; e < e_end; ++k, ++e)
        {
          const uint32_t t = P.cell_tri[k];
          uint32_t word = t;
          while (eo < o1 && (P.old_list.entry[eo] & SOLID_INDEX_MASK) < t)
            ++eo;
          // the (triangle, particle) pair stayed a candidate: its contact_info survives
          // (update_fine_search_candidates.cc:163-197)
          if (eo < o1 && (P.old_list.entry[eo] & SOLID_INDEX_MASK) == t && (P.old_list.entry[eo] & SOLID_HIST_BIT))
            {
              word |= SOLID_HIST_BIT;
              for (int d = 0; d < 3; ++d)
                P.new_list.hist[3 * size_t(e) + d] = P.old_list.hist[3 * size_t(eo) + d];
              if (P.use_roll)
                for (int d = 0; d < 3; ++d)
                  P.new_list.roll[3 * size_t(e) + d] = P.old_list.roll[3 * size_t(eo) + d];
            }
          else if (!have_old && !P.clear_history && P.pay.rec && P.pay.id[q] < P.pay.map_size)
            {
              // the contact_info of an immigrant came with it (HistRecord, HIST_REC_SOLID)
              const uint32_t qid = P.pay.id[q];
              for (uint32_t r = P.pay.start[qid]; r < P.pay.n && P.pay.rec[r].qid == qid; ++r)
                {
                  const HistRecord &rec = P.pay.rec[r];
                  if (!(rec.flags & HIST_REC_SOLID) || rec.rid != t)
                    continue;
                  word |= SOLID_HIST_BIT;
                  for (int d = 0; d < 3; ++d)
                    P.new_list.hist[3 * size_t(e) + d] = rec.h[d];
                  if (P.use_roll)
                    for (int d = 0; d < 3; ++d)
                      P.new_list.roll[3 * size_t(e) + d] = rec.roll[d];
                  break;
                }
            }
          P.new_list.entry[e] = word;
        }
      P.counts[q] = e_end > P.new_list.row_start[q] ? 1u : 0u;
    }

    // ---- contact force: calculate_particle_solid_object_contact ----
    struct Contact
    {
      uint32_t e;        // list entry
      uint32_t triangle; // global triangle index
      int indicator;
      double normal_overlap;
      vec3 normal; // triangle -> particle
    };

    __device__ inline bool in_csr(const uint32_t *start, const uint32_t *idx, uint32_t row, uint32_t x)
    {
      for (uint32_t k = start[row]; k < start[row + 1]; ++k)
        if (idx[k] == x)
          return true;
      return false;
    }

    __global__ void __launch_bounds__(64) k_solid_contacts(const __grid_constant__ SolidContactParams P, const __grid_constant__ MaterialTables mt)
    {
      if (P.spec_check)
        {
          const uint32_t f = *reinterpret_cast<const volatile uint32_t *>(P.flag_check);
          if (f != 0u && f != P.flag_tag)
            return;
        }
      const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
      if (a >= P.n_active)
        return;
      const uint32_t i = P.active[a];
      const double4 pi = P.in.pos[i], vi = P.in.vel[i], wi = P.in.omg[i];
      ParticleView me;
      me.x = v3(pi.x, pi.y, pi.z);
      me.d = pi.w;
      me.v = v3(vi.x, vi.y, vi.z);
      me.m = vi.w;
      me.w = v3(wi.x, wi.y, wi.z);
      me.type = static_cast<int>(static_cast<unsigned int>(wi.w));
      const double radius = me.d * 0.5;
      vec3 F = v3(0, 0, 0), T = v3(0, 0, 0);

      const uint32_t e0 = P.list.row_start[i], e1 = P.list.row_start[i + 1];
      uint32_t e = e0;
      // the reference handles one solid at a time: entries of one solid are contiguous
      while (e < e1)
        {
          const uint32_t solid = P.s.tri_solid[P.list.entry[e] & SOLID_INDEX_MASK];
          Contact rec[SOLID_MAX_CONTACTS];
          int n_rec = 0;
          uint32_t e_solid_end = e;
          for (; e_solid_end < e1 && P.s.tri_solid[P.list.entry[e_solid_end] & SOLID_INDEX_MASK] == solid; ++e_solid_end)
            {
              const uint32_t word = P.list.entry[e_solid_end];
              const uint32_t t = word & SOLID_INDEX_MASK;
              const double *q0 = P.s.vertices + 3 * size_t(P.s.tri[3 * size_t(t) + 0]);
              const double *q1 = P.s.vertices + 3 * size_t(P.s.tri[3 * size_t(t) + 1]);
              const double *q2 = P.s.vertices + 3 * size_t(P.s.tri[3 * size_t(t) + 2]);
              const vec3 p_0 = v3(q0[0], q0[1], q0[2]), p_1 = v3(q1[0], q1[1], q1[2]), p_2 = v3(q2[0], q2[1], q2[2]);
              // find_particle_triangle_projection (lethe_grid_tools.cc:1226-1450)
              const vec3 e_0 = p_1 - p_0, e_1 = p_2 - p_0;
              vec3 normal = cross(e_0, e_1);
              const double norm_normal = norm(normal);
              vec3 unit_normal = normal / norm_normal;
              const double ta = norm2(e_0), tb = dot(e_0, e_1), tc = norm2(e_1);
              const double det = ta * tc - tb * tb;
              const vec3 vector_to_plane = p_0 - me.x;
              if (dot(vector_to_plane, unit_normal) > 0)
                unit_normal = unit_normal * -1.0;
              // (sic) signed distance against a squared radius: never positive after the flip
              const double distance_squared = dot(vector_to_plane, unit_normal);
              bool touching = false;
              if (!(distance_squared > (radius * radius)))
                {
                  const double td = dot(e_0, vector_to_plane), te = dot(e_1, vector_to_plane);
                  double s, tt;
                  int indicator;
                  closest_point_parameters(ta, tb, tc, td, te, det, s, tt, indicator);
                  const vec3 pt_in_triangle = p_0 + s * e_0 + tt * e_1;
                  vec3 unit_normal_3d;
                  if (indicator == TC_FACE)
                    unit_normal_3d = unit_normal;
                  else
                    {
                      normal = me.x - pt_in_triangle;
                      unit_normal_3d = normal / norm(normal);
                    }
                  const double particle_triangle_distance = sqrt(dist2(me.x, pt_in_triangle));
                  const double normal_overlap = 0.5 * me.d - particle_triangle_distance;
                  if (normal_overlap > mt.pw_force_threshold)
                    {
                      touching = true;
                      if (n_rec < SOLID_MAX_CONTACTS)
                        {
                          rec[n_rec].e = e_solid_end;
                          rec[n_rec].triangle = t;
                          rec[n_rec].indicator = indicator;
                          rec[n_rec].normal_overlap = normal_overlap;
                          rec[n_rec].normal = unit_normal_3d;
                          ++n_rec;
                        }
                      else
                        {
                          *P.overflow = 1u;
                          if (P.flag_local)
                            *reinterpret_cast<volatile uint32_t *>(P.flag_local) = P.flag_tag;
                          if (P.flag_host)
                            *reinterpret_cast<volatile uint32_t *>(P.flag_host) = P.flag_tag;
                        }
                    }
                }
              // clear_contact_info: dropping the flag is the clear
              if (!touching && (word & SOLID_HIST_BIT))
                P.list.entry[e_solid_end] = t;
            }

          // double-contact elimination between connected triangles (:262-468)
          bool erased[SOLID_MAX_CONTACTS];
          for (int k = 0; k < n_rec; ++k)
            erased[k] = false;
          for (int c1 = 0; c1 < n_rec; ++c1)
            {
              if (erased[c1])
                continue;
              const uint32_t T1 = rec[c1].triangle;
              const int I1 = rec[c1].indicator;
              bool erase_contact_1 = false;
              for (int c2 = c1 + 1; c2 < n_rec; ++c2)
                {
                  if (erased[c2])
                    continue;
                  const uint32_t T2 = rec[c2].triangle;
                  const int I2 = rec[c2].indicator;
                  const bool es = in_csr(P.s.es_start, P.s.es_idx, T1, T2);
                  const bool vs = in_csr(P.s.vs_start, P.s.vs_idx, T1, T2);
                  if (!es && !vs)
                    continue; // disconnected triangles: both valid
                  if (I1 == TC_FACE)
                    {
                      if (I2 == TC_FACE)
                        continue;
                      if (I2 == TC_EDGE && vs)
                        continue;
                      erased[c2] = true;
                      continue;
                    }
                  if (I1 == TC_EDGE)
                    {
                      if (I2 == TC_FACE)
                        {
                          erase_contact_1 = true;
                          break;
                        }
                      if (I2 == TC_EDGE)
                        {
                          if (es)
                            erased[c2] = true;
                          continue;
                        }
                    }
                  if (I1 == TC_VERTEX)
                    {
                      if (I2 == TC_FACE)
                        {
                          erase_contact_1 = true;
                          break;
                        }
                      if (I2 == TC_EDGE)
                        {
                          if (vs)
                            {
                              erase_contact_1 = true;
                              break;
                            }
                          continue;
                        }
                      if (I2 == TC_VERTEX)
                        {
                          erased[c2] = true;
                          continue;
                        }
                    }
                }
              if (erase_contact_1)
                erased[c1] = true;
            }

          const SolidMotionDev &m = P.s.motion[solid];
          const vec3 translational_velocity = v3(m.translational_velocity[0], m.translational_velocity[1], m.translational_velocity[2]);
          const vec3 angular_velocity = v3(m.angular_velocity[0], m.angular_velocity[1], m.angular_velocity[2]);
          const vec3 center_of_rotation = v3(m.center_of_rotation[0], m.center_of_rotation[1], m.center_of_rotation[2]);
          for (int k = 0; k < n_rec; ++k)
            {
              const uint32_t ee = rec[k].e;
              const uint32_t word = P.list.entry[ee];
              if (erased[k])
                {
                  if (word & SOLID_HIST_BIT)
                    P.list.entry[ee] = word & SOLID_INDEX_MASK;
                  continue;
                }
              vec3 h = v3(0, 0, 0), rs = v3(0, 0, 0);
              double *hp = P.list.hist + 3 * size_t(ee);
              if (word & SOLID_HIST_BIT)
                {
                  h = v3(hp[0], hp[1], hp[2]);
                  if (P.rolling_model == LETHE_ROLLING_EPSD)
                    {
                      const double *rp = P.list.roll + 3 * size_t(ee);
                      rs = v3(rp[0], rp[1], rp[2]);
                    }
                }
              // update_particle_solid_object_contact_information (particle_wall_contact_force.h:283-331)
              const vec3 wall_normal = rec[k].normal;
              const vec3 normal_vector = -wall_normal;
              const double center_of_rotation_particle_distance = sqrt(dist2(center_of_rotation, me.x));
              const vec3 contact_relative_velocity =
                translational_velocity - me.v +
                cross((center_of_rotation_particle_distance * angular_velocity - 0.5 * me.d * me.w), normal_vector);
              const double vn = dot(contact_relative_velocity, normal_vector);
              const vec3 vt = contact_relative_velocity - vn * normal_vector;
              h = h + vt * P.dt;
              WallResult r;
              r.normal_force = r.tangential_force = r.tangential_torque = r.rolling = v3(0, 0, 0);
              pw_calculate_contact(P.pw_model, P.rolling_model, mt, wall_normal, h, rs, vt, vn, rec[k].normal_overlap, P.dt, me, r);
              const vec3 total_force = r.normal_force + r.tangential_force;
              F = F - total_force;
              T = T + (r.tangential_torque + r.rolling);
              hp[0] = h.x;
              hp[1] = h.y;
              hp[2] = h.z;
              if (P.rolling_model == LETHE_ROLLING_EPSD)
                {
                  double *rp = P.list.roll + 3 * size_t(ee);
                  rp[0] = rs.x;
                  rp[1] = rs.y;
                  rp[2] = rs.z;
                }
              if (!(word & SOLID_HIST_BIT))
                P.list.entry[ee] = word | SOLID_HIST_BIT;
            }
          e = e_solid_end;
        }
      P.force[3 * size_t(i) + 0] = F.x;
      P.force[3 * size_t(i) + 1] = F.y;
      P.force[3 * size_t(i) + 2] = F.z;
      P.torque[3 * size_t(i) + 0] = T.x;
      P.torque[3 * size_t(i) + 1] = T.y;
      P.torque[3 * size_t(i) + 2] = T.z;
    }

    inline unsigned blocks(size_t n, unsigned per) { return unsigned((n + per - 1) / per); }
  } // namespace

  void launch_move_solids(const SolidMoveParams &p, cudaStream_t s)
  {
    if (!p.s.n_vertices)
      return;
    k_move_solid_vertices<<<blocks(p.s.n_vertices, 128), 128, 0, s>>>(p);
    k_move_solid_centers<<<1, MAX_SOLIDS, 0, s>>>(p);
    count_launch(2);
  }
  void launch_count_solid_rows(const SolidBuildParams &p, cudaStream_t s)
  {
    if (p.n_rows)
      {
        k_count_solid_rows<<<blocks(p.n_rows, 256), 256, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_fill_solid_rows(const SolidBuildParams &p, cudaStream_t s)
  {
    if (p.n_rows)
      {
        k_fill_solid_rows<<<blocks(p.n_rows, 256), 256, 0, s>>>(p);
        count_launch();
      }
  }
  void launch_solid_contacts(const SolidContactParams &p, const MaterialTables &mt, cudaStream_t s)
  {
    if (p.n_active)
      {
        k_solid_contacts<<<blocks(p.n_active, 64), 64, 0, s>>>(p, mt);
        count_launch();
      }
  }

  // ---- host: mapping ----
  namespace
  {
    double host_point_triangle_distance(const double *p0, const double *p1, const double *p2, const double *pt)
    {
      // LetheGridTools::find_point_triangle_distance (lethe_grid_tools.cc:1536-1700), dim = 3
      double e0[3], e1[3], vp[3];
      for (int d = 0; d < 3; ++d)
        {
          e0[d] = p1[d] - p0[d];
          e1[d] = p2[d] - p0[d];
          vp[d] = p0[d] - pt[d];
        }
      auto dt3 = [](const double *x, const double *y) { return (x[0] * y[0] + x[1] * y[1]) + x[2] * y[2]; };
      const double a = dt3(e0, e0), b = dt3(e0, e1), c = dt3(e1, e1);
      const double det = a * c - b * b;
      const double d = dt3(e0, vp), e = dt3(e1, vp);
      double s, t;
      int ind;
      closest_point_parameters(a, b, c, d, e, det, s, t, ind);
      double acc = 0.0;
      for (int k = 0; k < 3; ++k)
        {
          const double q = (p0[k] + s * e0[k]) + t * e1[k];
          const double df = q - pt[k];
          acc = acc + df * df;
        }
      return std::sqrt(acc);
    }
  } // namespace

  void map_solids_on_host(const GridDesc &g, const double *vertices3, const uint32_t *tri3, uint32_t n_triangles,
                          std::vector<uint32_t> &cell_tri_start, std::vector<uint32_t> &cell_tri)
  {
    const int n_cells = g.n_cells;
    const double bg_cell_length = std::sqrt((g.h[0] * g.h[0] + g.h[1] * g.h[1]) + g.h[2] * g.h[2]);
    // (1) map_solid_in_background_triangulation: triangle t is mapped to every cell whose centre is
    // closer than one cell diameter; only cells near the triangle's bounding box can qualify
    std::vector<std::vector<uint32_t>> mapped(n_cells);
    for (uint32_t t = 0; t < n_triangles; ++t)
      {
        const double *p0 = vertices3 + 3 * size_t(tri3[3 * size_t(t)]);
        const double *p1 = vertices3 + 3 * size_t(tri3[3 * size_t(t) + 1]);
        const double *p2 = vertices3 + 3 * size_t(tri3[3 * size_t(t) + 2]);
        int lo[3], hi[3];
        for (int d = 0; d < 3; ++d)
          {
            const double mn = std::min(p0[d], std::min(p1[d], p2[d])) - bg_cell_length;
            const double mx = std::max(p0[d], std::max(p1[d], p2[d])) + bg_cell_length;
            lo[d] = std::max(0, int(std::floor((mn - g.lo[d]) / g.h[d])) - 1);
            hi[d] = std::min(g.n[d] - 1, int(std::floor((mx - g.lo[d]) / g.h[d])) + 1);
          }
        for (int k = lo[2]; k <= hi[2]; ++k)
          for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i)
              {
                const double center[3] = {g.lo[0] + (i + 0.5) * g.h[0], g.lo[1] + (j + 0.5) * g.h[1], g.lo[2] + (k + 0.5) * g.h[2]};
                if (host_point_triangle_distance(p0, p1, p2, center) < bg_cell_length)
                  mapped[i + g.n[0] * (j + g.n[1] * k)].push_back(t);
              }
      }
    // (2) particle_solid_surfaces_contact_search: the particles of a cell are candidates of the
    // triangles mapped to the cell and to its vertex-sharing neighbours
    cell_tri_start.assign(size_t(n_cells) + 1, 0);
    cell_tri.clear();
    std::vector<uint32_t> merged;
    for (int cell = 0; cell < n_cells; ++cell)
      {
        const int ci = cell % g.n[0], cj = (cell / g.n[0]) % g.n[1], ck = cell / (g.n[0] * g.n[1]);
        merged.clear();
        for (int dk = -1; dk <= 1; ++dk)
          for (int dj = -1; dj <= 1; ++dj)
            for (int di = -1; di <= 1; ++di)
              {
                const int i = ci + di, j = cj + dj, k = ck + dk;
                if (i < 0 || j < 0 || k < 0 || i >= g.n[0] || j >= g.n[1] || k >= g.n[2])
                  continue;
                const auto &m = mapped[i + g.n[0] * (j + g.n[1] * k)];
                merged.insert(merged.end(), m.begin(), m.end());
              }
        std::sort(merged.begin(), merged.end());
        merged.erase(std::unique(merged.begin(), merged.end()), merged.end());
        cell_tri.insert(cell_tri.end(), merged.begin(), merged.end());
        cell_tri_start[size_t(cell) + 1] = uint32_t(cell_tri.size());
      }
  }
} // namespace dem
