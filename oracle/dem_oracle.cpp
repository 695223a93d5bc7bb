// dem_oracle.cpp — CPU restatement of lethe-particles' DEM time step.
//
// TEST INFRASTRUCTURE ONLY. This file is the parity oracle for the CUDA engine
// in lethe_b200/csrc. Nothing in the product path may include, link or call it;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs use it.
//
// It restates, on plain arrays and without deal.II, the algorithm of the
// reference (citations relative to /root/reference). Parity is PINNED: the
// restatement reproduces the reference's own golden outputs
// (tests/dem/*.output, applications_tests/lethe-particles/packing_in_box*.output),
// see tests/test_oracle_golden.py and tests/golden/.
//
// Arithmetic rules that matter for the last bits:
//  * deal.II Tensor<1,3> / scalar multiplies by the reciprocal (inv = 1/s);
//  * dot products and squared norms are summed in component order;
//  * no FMA contraction (compile with -ffp-contract=off);
//  * the reference's constants are kept verbatim (0.66665, 1.8257, 1.3333, 9.8696).
//
// Container semantics: the reference keeps candidates and adjacency lists in
// ankerl::unordered_dense maps (insertion-ordered, erase = move last element into
// the hole). DenseRows below mimics that so that iteration order, and hence the
// floating-point summation order, is reference-like.

#include "../include/lethe_dem.h"

#include <algorithm>
#include <array>
#include <map>
#include <set>
#include <tuple>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace
{
  struct V3
  {
    double v[3];
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
  };
  inline V3 mk(double a, double b, double c) { return V3{{a, b, c}}; }
  inline V3 operator+(const V3 &a, const V3 &b) { return mk(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
  inline V3 operator-(const V3 &a, const V3 &b) { return mk(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
  inline V3 operator-(const V3 &a) { return mk(-a[0], -a[1], -a[2]); }
  inline V3 operator*(double s, const V3 &a) { return mk(s * a[0], s * a[1], s * a[2]); }
  inline V3 operator*(const V3 &a, double s) { return mk(a[0] * s, a[1] * s, a[2] * s); }
  // deal.II Tensor::operator/= for floating point: multiply by the reciprocal.
  inline V3 operator/(const V3 &a, double s)
  {
    const double inv = 1.0 / s;
    return mk(a[0] * inv, a[1] * inv, a[2] * inv);
  }
  inline double dot(const V3 &a, const V3 &b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
  inline double norm_square(const V3 &a) { return (a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]; }
  inline double norm(const V3 &a) { return std::sqrt(norm_square(a)); }
  inline V3 cross(const V3 &a, const V3 &b)
  {
    return mk(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
  }
  inline double distance_square(const V3 &a, const V3 &b)
  {
    double sum = 0.0;
    for (int i = 0; i < 3; ++i)
      {
        const double diff = a[i] - b[i];
        sum += diff * diff;
      }
    return sum;
  }
  inline double sq(double x) { return x * x; }
  inline double cube(double x) { return x * x * x; } // Utilities::fixed_power<3>: x * x^2 == x*x*x? see note

  // include/core/auxiliary_math_functions.h:17-22
  inline double harmonic_mean(double a, double b) { return (2 * a * b / (a + b + DBL_MIN)); }

  enum Prop { P_TYPE = 0, P_DP = 1, P_MASS = 2, P_VX = 3, P_VY = 4, P_VZ = 5, P_WX = 6, P_WY = 7, P_WZ = 8 };

  struct Particle
  {
    uint32_t id;
    V3 x;
    double p[9];
    int cell; // lexicographic grid cell the particle is registered in (last sort)
  };

  struct PPInfo
  {
    uint32_t two;
    V3 tangential_displacement;
    V3 rolling_resistance_spring_torque;
    V3 periodic_offset;
  };
  struct PPRow
  {
    uint32_t one;
    std::vector<PPInfo> second;
  };
  struct PWInfo
  {
    uint32_t face; // global_face_id or floating wall id
    V3 normal;
    V3 point;
    uint32_t boundary_id;
    V3 tangential_displacement;
    V3 rolling_resistance_spring_torque;
  };
  struct PWRow
  {
    uint32_t one;
    std::vector<PWInfo> second;
  };
  struct CandRow
  {
    uint32_t one;
    std::vector<uint32_t> c;
  };
  struct WCand
  {
    uint32_t face;
    V3 normal;
    V3 point;
    uint32_t boundary_id;
  };
  struct WCandRow
  {
    uint32_t one;
    std::vector<WCand> c;
  };

  // Insertion-ordered map id -> row with ankerl-style erase.
  template <class Row> struct DenseRows
  {
    std::vector<Row> rows;
    std::vector<int> pos; // id -> index in rows, -1 if absent
    void clear()
    {
      for (auto &r : rows)
        pos[r.one] = -1;
      rows.clear();
    }
    void ensure(uint32_t id)
    {
      if (pos.size() <= id)
        pos.resize(size_t(id) + 1, -1);
    }
    Row *find(uint32_t id)
    {
      if (id >= pos.size() || pos[id] < 0)
        return nullptr;
      return &rows[pos[id]];
    }
    Row &get_or_create(uint32_t id)
    {
      ensure(id);
      if (pos[id] < 0)
        {
          pos[id] = int(rows.size());
          rows.emplace_back();
          rows.back().one = id;
        }
      return rows[pos[id]];
    }
    // erase row index r; the last row moves into the hole (ankerl do_erase)
    void erase_at(size_t r)
    {
      pos[rows[r].one] = -1;
      if (r + 1 != rows.size())
        {
          rows[r] = std::move(rows.back());
          pos[rows[r].one] = int(r);
        }
      rows.pop_back();
    }
  };
  template <class T> inline void swap_erase(std::vector<T> &v, size_t k)
  {
    if (k + 1 != v.size())
      v[k] = std::move(v.back());
    v.pop_back();
  }

  struct BoundaryMotion
  {
    uint32_t boundary_id;
    V3 translational_velocity;
    double rotational_speed;
    V3 rotational_vector;
    V3 point_on_axis;
  };

  // SerialSolid<2,3> (source/core/serial_solid.cc) + the containers keyed by it
  // (include/dem/data_containers.h:147-177)
  struct Solid
  {
    std::vector<V3> vertices;
    std::vector<std::array<uint32_t, 3>> tri;
    std::vector<std::vector<uint32_t>> es_neighbors, vs_neighbors; // setup_containers (:617-682)
    V3 translational_velocity, angular_velocity, center_of_rotation;
    std::vector<V3> displacement_since_mapped;
    std::vector<std::pair<int, uint32_t>> mesh_info;           // (background cell, triangle), mapping order
    std::map<uint32_t, std::set<uint32_t>> candidates;         // triangle -> particle ids
    std::map<uint32_t, std::map<uint32_t, PWInfo>> in_contact; // triangle -> particle id -> contact_info
  };

  struct Oracle
  {
    lethe_dem_config cfg;
    std::string error;

    // grid
    int nx, ny, nz, n_cells;
    std::vector<int> cell_of_rank; // active-cell order -> lexicographic index
    std::vector<int> rank_of_cell;
    std::vector<std::vector<int>> cells_local_neighbor_list;          // [cell, n0, n1, ...] lexicographic ids
    std::vector<std::vector<int>> cells_local_periodic_neighbor_list; // [main, pn0, ...]
    std::vector<V3> combined_periodic_offsets;
    bool periodic_enabled;

    // particles
    std::vector<Particle> parts;               // iteration (= local index) order
    std::vector<std::vector<int>> cell_parts;  // per lexicographic cell: local indices in order
    std::vector<uint32_t> host_row_ids;        // step_host_state: id table of the previous call
    std::vector<V3> ext_force, ext_torque;     // set_external_loads: per particle id
    bool ext_enabled = false;
    bool open_next_step = false;               // restart_integration
    std::vector<int> slot_of_id;               // particle_container
    std::vector<V3> force, torque;
    std::vector<double> displacement, MOI;
    std::vector<V3> last_force, last_torque;

    // contact containers
    DenseRows<CandRow> local_candidates, periodic_candidates;
    DenseRows<PPRow> local_adjacent, periodic_adjacent;
    DenseRows<WCandRow> wall_candidates, fwall_candidates;
    DenseRows<PWRow> wall_in_contact, fwall_in_contact;

    // walls
    std::vector<lethe_wall_face> faces;
    std::vector<std::vector<int>> cell_faces; // per cell: indices into faces (in table order)
    int n_floating;
    V3 fw_point[LETHE_DEM_MAX_FLOATING_WALLS], fw_normal[LETHE_DEM_MAX_FLOATING_WALLS];
    double fw_t0[LETHE_DEM_MAX_FLOATING_WALLS], fw_t1[LETHE_DEM_MAX_FLOATING_WALLS];
    std::vector<std::vector<int>> fw_cells; // per floating wall: boundary cells (rank order)
    std::vector<BoundaryMotion> motions;

    // solid surfaces
    std::vector<Solid> solids;
    bool solid_object_search_trigger = false;

    // effective properties
    int n_types;
    std::vector<double> eY, eG, eRest, eMu, eRollVisc, eRollFric, eGamma, eHamaker, beta;
    std::vector<double> wY, wG, wRest, wMu, wRollVisc, wRollFric, wGamma, wHamaker, wbeta;
    double neighborhood_threshold_squared;
    double pp_force_threshold, pw_force_threshold;

    // time / triggers (DEMActionManager, dem_action_manager.h:395-412)
    uint64_t iteration_number;
    double current_time;
    bool contact_search_trigger;
    bool clear_tangential_displacement_trigger;
    uint64_t contact_build_number;
    uint64_t n_touching_last;
    // reference defect switch, see pp_execute_contact_calculation
    bool dmt_stale_scratch = true;

    // DEM-MP heat transfer (particle_heat_transfer.cc, multiphysics_integrator.cc)
    bool thermal_enabled = false;
    std::vector<double> th_real_E, th_roughness, th_slope, th_microhardness, th_gas_m; // n_types^2
    std::vector<double> th_conductivity;                                              // n_types
    double th_conductivity_gas = 0;
    std::vector<double> temperature, specific_heat; // per particle id
    std::vector<double> heat_transfer_rate;         // per local index (contact_outcome.heat_transfer_rate)
    std::vector<double> last_heat_transfer_rate;

    // adaptive sparse contacts (AdaptiveSparseContacts, adaptive_sparse_contacts.h)
    bool sparse_contacts_enabled = false;
    bool mobility_status_reset_trigger = false; // dem_action_manager.h:128-134
    std::vector<int> cell_mobility_status;      // per lexicographic cell
    std::vector<int> mobility_at_nodes;         // (nx+1)(ny+1)(nz+1) grid vertices
    // the status-aware searches / integration are in force (dem_action_manager.h:391-395)
    bool asc_active() const { return sparse_contacts_enabled && !mobility_status_reset_trigger; }
    bool cell_is_mobile(int cell) const { return !asc_active() || cell_mobility_status[cell] == LETHE_MOBILITY_MOBILE; }
  };

  // ---------------------------------------------------------------- grid ----
  inline int lin(const Oracle &o, int i, int j, int k) { return i + o.nx * (j + o.ny * k); }

  void build_cell_order(Oracle &o)
  {
    o.cell_of_rank.clear();
    o.rank_of_cell.assign(o.n_cells, -1);
    if (o.cfg.cell_order == LETHE_CELL_ORDER_MORTON)
      {
        // hyper_cube + refine_global: children of a hex are visited x-fastest,
        // recursively, i.e. Morton order with x the lowest bit.
        int nmax = std::max(o.nx, std::max(o.ny, o.nz));
        int bits = 0;
        while ((1 << bits) < nmax)
          ++bits;
        const uint64_t total = uint64_t(1) << (3 * bits);
        for (uint64_t m = 0; m < total; ++m)
          {
            int i = 0, j = 0, k = 0;
            for (int b = 0; b < bits; ++b)
              {
                i |= int((m >> (3 * b)) & 1) << b;
                j |= int((m >> (3 * b + 1)) & 1) << b;
                k |= int((m >> (3 * b + 2)) & 1) << b;
              }
            if (i < o.nx && j < o.ny && k < o.nz)
              o.cell_of_rank.push_back(lin(o, i, j, k));
          }
      }
    else
      {
        for (int c = 0; c < o.n_cells; ++c)
          o.cell_of_rank.push_back(c);
      }
    for (int r = 0; r < o.n_cells; ++r)
      o.rank_of_cell[o.cell_of_rank[r]] = r;
  }

  // cells touching lattice vertex (vi,vj,vk), sorted by active-cell rank
  // (GridTools::vertex_to_cell_map gives a std::set of cell iterators)
  void cells_at_vertex(const Oracle &o, int vi, int vj, int vk, std::vector<int> &out)
  {
    out.clear();
    for (int dk = -1; dk <= 0; ++dk)
      for (int dj = -1; dj <= 0; ++dj)
        for (int di = -1; di <= 0; ++di)
          {
            int i = vi + di, j = vj + dj, k = vk + dk;
            if (i < 0 || j < 0 || k < 0 || i >= o.nx || j >= o.ny || k >= o.nz)
              continue;
            out.push_back(lin(o, i, j, k));
          }
    std::sort(out.begin(), out.end(), [&](int a, int b) { return o.rank_of_cell[a] < o.rank_of_cell[b]; });
  }

  // find_cell_neighbors<dim,false> (source/dem/find_cell_neighbors.cc:10-104), one rank
  void find_cell_neighbors(Oracle &o)
  {
    o.cells_local_neighbor_list.clear();
    std::vector<char> total_cell_list(o.n_cells, 0);
    std::vector<int> vcells;
    for (int r = 0; r < o.n_cells; ++r)
      {
        const int cell = o.cell_of_rank[r];
        const int ci = cell % o.nx, cj = (cell / o.nx) % o.ny, ck = cell / (o.nx * o.ny);
        std::vector<int> local_neighbor_vector;
        local_neighbor_vector.push_back(cell);
        total_cell_list[cell] = 1;
        for (int vertex = 0; vertex < 8; ++vertex) // deal.II hex vertex v = i + 2j + 4k
          {
            cells_at_vertex(o, ci + (vertex & 1), cj + ((vertex >> 1) & 1), ck + ((vertex >> 2) & 1), vcells);
            for (int neighbor : vcells)
              {
                if (total_cell_list[neighbor])
                  continue;
                if (std::find(local_neighbor_vector.begin(), local_neighbor_vector.end(), neighbor) !=
                    local_neighbor_vector.end())
                  continue;
                local_neighbor_vector.push_back(neighbor);
              }
          }
        o.cells_local_neighbor_list.push_back(local_neighbor_vector);
      }
  }

  // find_cell_periodic_neighbors (find_cell_neighbors.cc:104-275) + get_periodic_neighbor_list
  // (:337-378), one rank. Main cells are the cells with a face on periodic boundary 0
  // (the low face of each periodic direction); they are visited in active-cell order.
  void find_cell_periodic_neighbors(Oracle &o)
  {
    o.cells_local_periodic_neighbor_list.clear();
    if (!o.periodic_enabled)
      return;
    const int n[3] = {o.nx, o.ny, o.nz};
    std::vector<char> total_cell_list(o.n_cells, 0);
    std::vector<int> vcells;
    for (int r = 0; r < o.n_cells; ++r)
      {
        const int cell = o.cell_of_rank[r];
        const int c[3] = {cell % o.nx, (cell / o.nx) % o.ny, cell / (o.nx * o.ny)};
        bool on_pb0 = false;
        for (int d = 0; d < 3; ++d)
          if (o.cfg.periodic[d] && c[d] == 0)
            on_pb0 = true;
        if (!on_pb0)
          continue;
        std::vector<int> vec;
        vec.push_back(cell);
        total_cell_list[cell] = 1;
        std::vector<int> periodic_neighbor_list;
        for (int vertex = 0; vertex < 8; ++vertex)
          {
            const int v[3] = {c[0] + (vertex & 1), c[1] + ((vertex >> 1) & 1), c[2] + ((vertex >> 2) & 1)};
            // coinciding vertices: every combination of periodic images, ascending vertex id
            std::vector<int> alt[3];
            for (int d = 0; d < 3; ++d)
              {
                alt[d].push_back(v[d]);
                if (o.cfg.periodic[d] && (v[d] == 0 || v[d] == n[d]))
                  alt[d].push_back(v[d] == 0 ? n[d] : 0);
                std::sort(alt[d].begin(), alt[d].end());
              }
            if (alt[0].size() * alt[1].size() * alt[2].size() == 1)
              continue;
            for (int wk : alt[2])
              for (int wj : alt[1])
                for (int wi : alt[0])
                  {
                    if (wi == v[0] && wj == v[1] && wk == v[2])
                      continue;
                    cells_at_vertex(o, wi, wj, wk, vcells);
                    for (int nb : vcells)
                      periodic_neighbor_list.push_back(nb);
                  }
          }
        for (int nb : periodic_neighbor_list)
          {
            if (total_cell_list[nb])
              continue;
            if (std::find(vec.begin(), vec.end(), nb) != vec.end())
              continue;
            vec.push_back(nb);
          }
        o.cells_local_periodic_neighbor_list.push_back(vec);
      }
  }

  // PeriodicBoundariesManipulator::compute_combined_periodic_offsets
  // (periodic_boundaries_manipulator.cc:229-263); offsets point from pb0 to pb1.
  void compute_combined_periodic_offsets(Oracle &o)
  {
    o.combined_periodic_offsets.clear();
    const int n[3] = {o.nx, o.ny, o.nz};
    for (int d = 0; d < 3; ++d)
      {
        if (!o.cfg.periodic[d])
          continue;
        V3 offset = mk(0, 0, 0);
        // point_on_periodic_face[d] - point_on_face[d] = hi - lo
        offset[d] = (o.cfg.grid_lo[d] + n[d] * o.cfg.cell_size[d]) - o.cfg.grid_lo[d];
        size_t current_size = o.combined_periodic_offsets.size();
        if (current_size == 0)
          {
            o.combined_periodic_offsets.push_back(offset);
            o.combined_periodic_offsets.push_back(-offset);
          }
        else
          {
            o.combined_periodic_offsets.push_back(offset);
            o.combined_periodic_offsets.push_back(-offset);
            for (size_t i = 0; i < current_size; ++i)
              {
                o.combined_periodic_offsets.push_back(o.combined_periodic_offsets[i] + offset);
                o.combined_periodic_offsets.push_back(o.combined_periodic_offsets[i] - offset);
              }
          }
      }
  }

  // --------------------------------------------------- effective properties --
  // ParticleParticleContactForce::set_effective_properties (…contact_force.h:1639-1746)
  // ParticleWallContactForce::set_effective_properties (…wall_contact_force.cc:588-694)
  void set_effective_properties(Oracle &o)
  {
    const lethe_dem_config &c = o.cfg;
    const int n = c.n_types;
    o.n_types = n;
    for (auto *v : {&o.eY, &o.eG, &o.eRest, &o.eMu, &o.eRollVisc, &o.eRollFric, &o.eGamma, &o.eHamaker, &o.beta})
      v->assign(size_t(n) * n, 0.0);
    for (auto *v : {&o.wY, &o.wG, &o.wRest, &o.wMu, &o.wRollVisc, &o.wRollFric, &o.wGamma, &o.wHamaker, &o.wbeta})
      v->assign(size_t(n), 0.0);
    for (int i = 0; i < n; ++i)
      {
        const double Yi = c.young[i], nui = c.poisson[i];
        for (int j = 0; j < n; ++j)
          {
            const int k = i * n + j;
            const double Yj = c.young[j], nuj = c.poisson[j];
            o.eY[k] = (Yi * Yj) / ((Yj * (1.0 - nui * nui)) + (Yi * (1.0 - nuj * nuj)) + DBL_MIN);
            o.eG[k] = (Yi * Yj) / (2.0 * ((Yj * (2.0 - nui) * (1.0 + nui)) + (Yi * (2.0 - nuj) * (1.0 + nuj))) + DBL_MIN);
            o.eRest[k] = harmonic_mean(c.restitution[i], c.restitution[j]);
            o.eMu[k] = harmonic_mean(c.friction[i], c.friction[j]);
            o.eRollVisc[k] = harmonic_mean(c.rolling_viscous_damping[i], c.rolling_viscous_damping[j]);
            o.eRollFric[k] = harmonic_mean(c.rolling_friction[i], c.rolling_friction[j]);
            o.eGamma[k] = c.surface_energy[i] + c.surface_energy[j] -
                          std::pow(std::sqrt(c.surface_energy[i]) - std::sqrt(c.surface_energy[j]), 2);
            o.eHamaker[k] = 0.5 * (c.hamaker[i] + c.hamaker[j]);
            const double lg = std::log(o.eRest[k]);
            o.beta[k] = lg / std::sqrt(lg * lg + 9.8696);
          }
        const double Yw = c.young_wall, nuw = c.poisson_wall;
        o.wY[i] = (Yi * Yw) / (Yw * (1. - nui * nui) + Yi * (1. - nuw * nuw) + DBL_MIN);
        o.wG[i] = (Yi * Yw) / ((2. * Yw * (2. - nui) * (1. + nui)) + (2. * Yi * (2. - nuw) * (1. + nuw)) + DBL_MIN);
        o.wRest[i] = harmonic_mean(c.restitution[i], c.restitution_wall);
        o.wMu[i] = harmonic_mean(c.friction[i], c.friction_wall);
        o.wRollFric[i] = harmonic_mean(c.rolling_friction[i], c.rolling_friction_wall);
        o.wRollVisc[i] = harmonic_mean(c.rolling_viscous_damping[i], c.rolling_viscous_damping_wall);
        o.wGamma[i] = c.surface_energy[i] + c.surface_energy_wall -
                      std::pow(std::sqrt(c.surface_energy[i]) - std::sqrt(c.surface_energy_wall), 2);
        o.wHamaker[i] = 0.5 * (c.hamaker[i] + c.hamaker_wall);
        const double lg = std::log(o.wRest[i]);
        o.wbeta[i] = lg / std::sqrt((lg * lg) + 9.8696);
      }
    // get_force_calculation_threshold_distance (…contact_force.h:504-529, wall .h equivalent)
    o.pp_force_threshold = 0.;
    if (c.pp_model == LETHE_PP_DMT)
      {
        const double maxA = *std::max_element(o.eHamaker.begin(), o.eHamaker.end());
        const double minG = *std::min_element(o.eGamma.begin(), o.eGamma.end());
        o.pp_force_threshold = -std::sqrt(maxA / (12. * M_PI * minG * c.dmt_cut_off_threshold));
      }
    o.pw_force_threshold = 0.;
    if (c.pw_model == LETHE_PW_DMT)
      {
        const double maxA = *std::max_element(o.wHamaker.begin(), o.wHamaker.end());
        const double minG = *std::min_element(o.wGamma.begin(), o.wGamma.end());
        o.pw_force_threshold = -std::sqrt(maxA / (12. * M_PI * minG * c.dmt_cut_off_threshold));
      }
    // dem.cc:156-159
    o.neighborhood_threshold_squared = std::pow(c.neighborhood_threshold * c.d_max, 2);
  }

  // -------------------------------------------------------------- sorting ----
  inline int cell_of_point(const Oracle &o, const V3 &x)
  {
    int idx[3];
    const int n[3] = {o.nx, o.ny, o.nz};
    for (int d = 0; d < 3; ++d)
      {
        const double r = (x[d] - o.cfg.grid_lo[d]) / o.cfg.cell_size[d];
        double f = std::floor(r);
        // a point exactly on a face shared by two cells belongs to the lower one: the reference's
        // point location returns the first (lowest-index) cell that contains the point. Pinned by
        // epsd_rolling_resistance_model.output, whose resting spheres sit on cell corners and keep
        // their wall-contact history only if they stay registered in the lower cell.
        if (r == f && f > 0.0)
          f -= 1.0;
        if (!(f >= 0.0) || !(f < double(n[d])))
          return -1; // left the triangulation: deal.II drops the particle
        idx[d] = int(f);
      }
    return lin(o, idx[0], idx[1], idx[2]);
  }

  // PeriodicBoundariesManipulator::execute_particles_displacement + check_and_move_particles
  // (periodic_boundaries_manipulator.cc:145-224,266-338): particles registered in a cell on
  // a periodic face that are on or beyond that face are translated by +-offset.
  void execute_particles_displacement(Oracle &o)
  {
    if (!o.periodic_enabled)
      return;
    const int n[3] = {o.nx, o.ny, o.nz};
    for (auto &p : o.parts)
      {
        if (p.cell < 0)
          continue;
        const int c[3] = {p.cell % o.nx, (p.cell / o.nx) % o.ny, p.cell / (o.nx * o.ny)};
        for (int d = 0; d < 3; ++d)
          {
            if (!o.cfg.periodic[d])
              continue;
            const double lo = o.cfg.grid_lo[d];
            const double hi = o.cfg.grid_lo[d] + n[d] * o.cfg.cell_size[d];
            const double offset = hi - lo;
            if (c[d] == 0)
              {
                // pb0 cell: outward normal -e_d, point on face lo
                const double distance_with_face = (p.x[d] - lo) * -1.0;
                if (distance_with_face >= 0.0)
                  p.x[d] += offset;
              }
            if (c[d] == n[d] - 1)
              {
                const double distance_with_face = (p.x[d] - hi) * 1.0;
                if (distance_with_face >= 0.0)
                  p.x[d] += -offset;
              }
          }
      }
  }

  // The cell a particle registered in `old_cell` belongs to after moving: it stays while it is inside
  // the old cell up to ParticleHandler's tolerance_inside_cell (1e-12 in unit-cell coordinates).
  inline int cell_after_move(const Oracle &o, const V3 &x, int old_cell)
  {
    if (old_cell >= 0)
      {
        const int idx[3] = {old_cell % o.nx, (old_cell / o.nx) % o.ny, old_cell / (o.nx * o.ny)};
        bool inside = true;
        for (int d = 0; d < 3; ++d)
          {
            const double u = (x[d] - (o.cfg.grid_lo[d] + idx[d] * o.cfg.cell_size[d])) / o.cfg.cell_size[d];
            inside = inside && u >= -1e-12 && u <= 1. + 1e-12;
          }
        if (inside)
          return old_cell;
      }
    return cell_of_point(o, x);
  }

  // DEMSolver::sort_particles_into_subdomains_and_cells (dem.cc:982-1016) ->
  // Particles::ParticleHandler::sort_particles_into_subdomains_and_cells of deal.II (external
  // dependency, >= 9.4 storage: one vector of particles per cell). The order of the particles
  // inside a cell fixes the order of the broad-search candidates, through it the insertion order of
  // the contact containers and so the floating-point summation order of the forces; chaotic cases
  // (solid_surface.output: 78 spheres, dozens of collisions each) only reproduce with the same order:
  //  1. cells in active-cell order, particles in cell order: those no longer inside go to a list;
  //  2. in list order each one is appended to the END of its new cell and its old slot invalidated;
  //  3. remove_particles walks the list BACKWARDS; each invalid slot is overwritten by the cell's
  //     LAST entry (swap-and-pop), so arrivals fill the holes of departures from the back.
  // Newly inserted particles were appended to their cells at insertion time, in insertion order.
  void sort_particles_into_subdomains_and_cells(Oracle &o)
  {
    std::vector<std::vector<int>> new_cells(o.n_cells);
    std::vector<char> lost(o.parts.size(), 0);
    size_t registered = 0;
    for (int c = 0; c < o.n_cells; ++c)
      {
        registered += o.cell_parts[c].size();
        new_cells[c] = o.cell_parts[c];
      }
    for (size_t s = registered; s < o.parts.size(); ++s)
      {
        if (o.parts[s].cell < 0)
          lost[s] = 1; // inserted outside of the triangulation: no cell is found for it
        else
          new_cells[o.parts[s].cell].push_back(int(s));
      }
    struct Out
    {
      int cell, index, destination;
    };
    std::vector<Out> out_of_cell;
    for (int r = 0; r < o.n_cells; ++r)
      {
        const int c = o.cell_of_rank[r];
        for (size_t k = 0; k < new_cells[c].size(); ++k)
          {
            const int destination = cell_after_move(o, o.parts[new_cells[c][k]].x, c);
            if (destination != c)
              out_of_cell.push_back({c, int(k), destination});
          }
      }
    for (const Out &m : out_of_cell)
      {
        const int s = new_cells[m.cell][m.index];
        if (m.destination >= 0)
          new_cells[m.destination].push_back(s);
        else
          lost[s] = 1;
      }
    for (auto m = out_of_cell.rbegin(); m != out_of_cell.rend(); ++m)
      {
        std::vector<int> &v = new_cells[m->cell];
        v[m->index] = v.back();
        v.pop_back();
      }
    std::vector<Particle> np;
    np.reserve(o.parts.size());
    o.cell_parts.assign(o.n_cells, std::vector<int>());
    for (int r = 0; r < o.n_cells; ++r)
      {
        const int c = o.cell_of_rank[r];
        for (int s : new_cells[c])
          {
            o.cell_parts[c].push_back(int(np.size()));
            np.push_back(o.parts[s]);
            np.back().cell = c;
          }
      }
    for (size_t s = 0; s < o.parts.size(); ++s)
      if (lost[s] && o.parts[s].id < o.slot_of_id.size())
        o.slot_of_id[o.parts[s].id] = -1;
    o.parts.swap(np);
    const size_t n = o.parts.size();
    o.force.assign(n, mk(0, 0, 0));
    o.torque.assign(n, mk(0, 0, 0));
    o.MOI.resize(n);
    for (size_t s = 0; s < n; ++s)
      {
        const double *p = o.parts[s].p;
        o.MOI[s] = o.cfg.moi_override > 0 ? o.cfg.moi_override : 0.1 * p[P_MASS] * p[P_DP] * p[P_DP];
      }
    o.displacement.assign(n, 0.);
    o.heat_transfer_rate.assign(n, 0.); // resize_interaction_containers
  }

  // update_particle_container (update_local_particle_containers.cc:11-38)
  void update_particle_container(Oracle &o)
  {
    std::fill(o.slot_of_id.begin(), o.slot_of_id.end(), -1);
    for (size_t s = 0; s < o.parts.size(); ++s)
      {
        const uint32_t id = o.parts[s].id;
        if (o.slot_of_id.size() <= id)
          o.slot_of_id.resize(size_t(id) + 1, -1);
        o.slot_of_id[id] = int(s);
      }
  }
  inline int slot(const Oracle &o, uint32_t id) { return id < o.slot_of_id.size() ? o.slot_of_id[id] : -1; }

  // ---------------------------------------------------------- broad search ---
  // store_candidates (particle_particle_broad_search.cc:735-766)
  inline void store_candidates(Oracle &o, DenseRows<CandRow> &cand, uint32_t main_id, const std::vector<int> &cellp,
                               size_t begin)
  {
    CandRow &row = cand.get_or_create(main_id);
    for (size_t k = begin; k < cellp.size(); ++k)
      row.c.push_back(o.parts[cellp[k]].id);
  }

  // ------------------------------------------------ adaptive sparse contacts ----
  // Grid vertex (node of the FE_Q(1) background DoF handler) index; periodic directions share
  // the nodes of their two faces (periodic_node_ids, adaptive_sparse_contacts.h:279-295: nodes
  // that are identity-constrained to each other exchange their status on every assignment).
  inline int node_index(const Oracle &o, int i, int j, int k)
  {
    if (o.cfg.periodic[0] && i == o.nx)
      i = 0;
    if (o.cfg.periodic[1] && j == o.ny)
      j = 0;
    if (o.cfg.periodic[2] && k == o.nz)
      k = 0;
    return i + (o.nx + 1) * (j + (o.ny + 1) * k);
  }
  inline void cell_nodes(const Oracle &o, int cell, int out[8])
  {
    const int ci = cell % o.nx, cj = (cell / o.nx) % o.ny, ck = cell / (o.nx * o.ny);
    for (int v = 0; v < 8; ++v)
      out[v] = node_index(o, ci + (v & 1), cj + ((v >> 1) & 1), ck + (v >> 2));
  }

  // AdaptiveSparseContacts::calculate_granular_temperature_and_solid_fraction
  // (adaptive_sparse_contacts.cc:35-130) for one non-empty cell
  void granular_temperature_and_solid_fraction(const Oracle &o, int cell, double &granular_temperature, double &solid_fraction)
  {
    const std::vector<int> &in_cell = o.cell_parts[cell];
    const unsigned int n_particles_in_cell = static_cast<unsigned int>(in_cell.size());
    double solid_volume = 0.0;
    // cell->measure() of an axis-aligned hexahedron
    const double cell_volume = o.cfg.cell_size[0] * o.cfg.cell_size[1] * o.cfg.cell_size[2];
    V3 velocity_cell_average = mk(0, 0, 0);
    V3 cell_velocity_fluctuation_squared_average = mk(0, 0, 0);
    for (int s : in_cell)
      {
        const double *p = o.parts[s].p;
        const double dp = p[P_DP];
        for (int d = 0; d < 3; ++d)
          velocity_cell_average[d] += p[P_VX + d];
        solid_volume += M_PI * (dp * dp * dp) / (2.0 * 3);
      }
    {
      // Tensor /= scalar
      const double inv = 1.0 / n_particles_in_cell;
      for (int d = 0; d < 3; ++d)
        velocity_cell_average[d] *= inv;
    }
    solid_fraction = solid_volume / cell_volume;
    for (int s : in_cell)
      {
        const double *p = o.parts[s].p;
        for (int d = 0; d < 3; ++d)
          {
            const double f = p[P_VX + d] - velocity_cell_average[d];
            cell_velocity_fluctuation_squared_average[d] += f * f;
          }
      }
    granular_temperature = 0.0;
    for (int d = 0; d < 3; ++d)
      {
        cell_velocity_fluctuation_squared_average[d] /= n_particles_in_cell;
        granular_temperature += cell_velocity_fluctuation_squared_average[d] / 3;
      }
  }

  // AdaptiveSparseContacts::identify_mobility_status (adaptive_sparse_contacts.cc:132-356);
  // `advect particles` (CFD-DEM) is not restated: inactive / static_active only.
  void identify_mobility_status(Oracle &o)
  {
    if (!o.sparse_contacts_enabled)
      return;
    o.cell_mobility_status.assign(o.n_cells, -1);
    if (o.mobility_status_reset_trigger)
      {
        o.cell_mobility_status.assign(o.n_cells, LETHE_MOBILITY_MOBILE);
        return;
      }
    o.mobility_at_nodes.assign(size_t(o.nx + 1) * (o.ny + 1) * (o.nz + 1), 0);
    int nodes[8];
    auto assign = [&](int cell, int cell_status, int node_status) {
      o.cell_mobility_status[cell] = cell_status;
      for (int v = 0; v < 8; ++v)
        o.mobility_at_nodes[nodes[v]] = std::max(node_status, o.mobility_at_nodes[nodes[v]]);
    };
    // 1. empty cells: cell inactive, nodes empty
    for (int cell = 0; cell < o.n_cells; ++cell)
      if (o.cell_parts[cell].empty())
        {
          cell_nodes(o, cell, nodes);
          assign(cell, LETHE_MOBILITY_INACTIVE, LETHE_MOBILITY_EMPTY_NODE);
        }
    // 2. mobile by criteria: granular temperature, solid fraction, next to an empty cell
    for (int cell = 0; cell < o.n_cells; ++cell)
      {
        if (o.cell_mobility_status[cell] != -1)
          continue;
        cell_nodes(o, cell, nodes);
        bool has_empty_node = false;
        for (int v = 0; v < 8; ++v)
          if (o.mobility_at_nodes[nodes[v]] == LETHE_MOBILITY_EMPTY_NODE)
            {
              has_empty_node = true;
              break;
            }
        double granular_temperature, solid_fraction;
        granular_temperature_and_solid_fraction(o, cell, granular_temperature, solid_fraction);
        if (granular_temperature > o.cfg.asc_granular_temperature_threshold || solid_fraction < o.cfg.asc_solid_fraction_threshold ||
            has_empty_node)
          assign(cell, LETHE_MOBILITY_MOBILE, LETHE_MOBILITY_MOBILE);
      }
    // 3. mobile by neighbour (a node flagged mobile by step 2): the additional mobile layer;
    //    its other nodes become active. Step 3 never creates a mobile node (max with active).
    for (int cell = 0; cell < o.n_cells; ++cell)
      {
        if (o.cell_mobility_status[cell] != -1)
          continue;
        cell_nodes(o, cell, nodes);
        bool has_mobile_node = false;
        for (int v = 0; v < 8; ++v)
          if (o.mobility_at_nodes[nodes[v]] == LETHE_MOBILITY_MOBILE)
            {
              has_mobile_node = true;
              break;
            }
        if (has_mobile_node)
          assign(cell, LETHE_MOBILITY_MOBILE, LETHE_MOBILITY_STATIC_ACTIVE);
      }
    // 4. the layer of active cells (a node flagged active); the rest is inactive
    for (int cell = 0; cell < o.n_cells; ++cell)
      {
        if (o.cell_mobility_status[cell] != -1)
          continue;
        cell_nodes(o, cell, nodes);
        bool has_active_nodes = false;
        for (int v = 0; v < 8; ++v)
          if (o.mobility_at_nodes[nodes[v]] == LETHE_MOBILITY_STATIC_ACTIVE)
            {
              has_active_nodes = true;
              break;
            }
        o.cell_mobility_status[cell] = has_active_nodes ? LETHE_MOBILITY_STATIC_ACTIVE : LETHE_MOBILITY_INACTIVE;
      }
  }

  // find_particle_particle_contact_pairs (particle_particle_broad_search.cc:9-132; with adaptive
  // sparse contacts :134-316), local part
  void find_particle_particle_contact_pairs(Oracle &o)
  {
    o.local_candidates.clear();
    const bool asc = o.asc_active();
    for (const auto &list : o.cells_local_neighbor_list)
      {
        const std::vector<int> &main = o.cell_parts[list[0]];
        const int main_status = asc ? o.cell_mobility_status[list[0]] : LETHE_MOBILITY_MOBILE;
        // inactive main cell: nothing; empty cells are inactive
        if (main_status == LETHE_MOBILITY_INACTIVE)
          continue;
        if (main.empty())
          continue;
        // pairs inside the main cell only if it is mobile
        if (main_status == LETHE_MOBILITY_MOBILE)
          for (size_t a = 0; a < main.size(); ++a)
            store_candidates(o, o.local_candidates, o.parts[main[a]].id, main, a + 1);
        for (size_t nb = 1; nb < list.size(); ++nb)
          {
            // an active main cell only pairs with mobile neighbours
            if (asc && main_status == LETHE_MOBILITY_STATIC_ACTIVE && o.cell_mobility_status[list[nb]] != LETHE_MOBILITY_MOBILE)
              continue;
            const std::vector<int> &other = o.cell_parts[list[nb]];
            for (size_t a = 0; a < main.size(); ++a)
              store_candidates(o, o.local_candidates, o.parts[main[a]].id, other, 0);
          }
      }
  }
  // find_particle_particle_periodic_contact_pairs (…broad_search.cc:316-380), local part
  void find_particle_particle_periodic_contact_pairs(Oracle &o)
  {
    o.periodic_candidates.clear();
    const bool asc = o.asc_active();
    for (const auto &list : o.cells_local_periodic_neighbor_list)
      {
        const std::vector<int> &main = o.cell_parts[list[0]];
        const int main_status = asc ? o.cell_mobility_status[list[0]] : LETHE_MOBILITY_MOBILE;
        if (main_status == LETHE_MOBILITY_INACTIVE)
          continue;
        if (main.empty())
          continue;
        for (size_t nb = 1; nb < list.size(); ++nb)
          {
            if (asc && main_status == LETHE_MOBILITY_STATIC_ACTIVE && o.cell_mobility_status[list[nb]] != LETHE_MOBILITY_MOBILE)
              continue;
            const std::vector<int> &other = o.cell_parts[list[nb]];
            for (size_t a = 0; a < main.size(); ++a)
              store_candidates(o, o.periodic_candidates, o.parts[main[a]].id, other, 0);
          }
      }
  }

  // find_particle_wall_contact_pairs + store_candidates
  // (particle_wall_broad_search.cc:8-56, particle_wall_broad_search.h:212-246).
  // boundary_cells_information is a std::map keyed by global_face_id -> ascending face id.
  void find_particle_wall_contact_pairs(Oracle &o, const std::vector<int> &faces_by_id)
  {
    o.wall_candidates.clear();
    for (int f : faces_by_id)
      {
        const lethe_wall_face &face = o.faces[f];
        if (face.cell < 0 || face.cell >= o.n_cells)
          continue;
        if (!o.cell_is_mobile(face.cell)) // particle_wall_broad_search.cc:239-246
          continue;
        for (int s : o.cell_parts[face.cell])
          {
            WCandRow &row = o.wall_candidates.get_or_create(o.parts[s].id);
            bool exists = false;
            for (auto &w : row.c)
              if (w.face == face.global_face_id)
                exists = true;
            if (!exists)
              row.c.push_back(WCand{face.global_face_id, mk(face.normal[0], face.normal[1], face.normal[2]),
                                    mk(face.point[0], face.point[1], face.point[2]), face.boundary_id});
          }
      }
  }

  // find_particle_floating_wall_contact_pairs (particle_wall_broad_search.cc:58-125)
  void find_particle_floating_wall_contact_pairs(Oracle &o, double simulation_time)
  {
    o.fwall_candidates.clear();
    for (int w = 0; w < o.n_floating; ++w)
      {
        if (!(simulation_time >= o.fw_t0[w] && simulation_time <= o.fw_t1[w]))
          continue;
        for (int cell : o.fw_cells[w])
          for (int s : o.cell_parts[cell])
            {
              if (!o.cell_is_mobile(cell)) // particle_wall_broad_search.cc:310-317
                continue;
              WCandRow &row = o.fwall_candidates.get_or_create(o.parts[s].id);
              bool exists = false;
              for (auto &c : row.c)
                if (c.face == uint32_t(w))
                  exists = true;
              if (!exists)
                row.c.push_back(WCand{uint32_t(w), mk(0, 0, 0), mk(0, 0, 0), LETHE_DEM_FLOATING_WALL_BOUNDARY_ID});
            }
      }
  }

  // BoundaryCellsInformation::find_boundary_cells_for_floating_walls
  // (find_boundary_cells_information.cc:653-703)
  void find_boundary_cells_for_floating_walls(Oracle &o, double maximum_cell_diameter)
  {
    o.fw_cells.assign(o.n_floating, std::vector<int>());
    for (int w = 0; w < o.n_floating; ++w)
      for (int r = 0; r < o.n_cells; ++r)
        {
          const int cell = o.cell_of_rank[r];
          const int c[3] = {cell % o.nx, (cell / o.nx) % o.ny, cell / (o.nx * o.ny)};
          bool in = false;
          for (int vertex = 0; vertex < 8 && !in; ++vertex)
            {
              V3 vx;
              vx[0] = o.cfg.grid_lo[0] + (c[0] + (vertex & 1)) * o.cfg.cell_size[0];
              vx[1] = o.cfg.grid_lo[1] + (c[1] + ((vertex >> 1) & 1)) * o.cfg.cell_size[1];
              vx[2] = o.cfg.grid_lo[2] + (c[2] + ((vertex >> 2) & 1)) * o.cfg.cell_size[2];
              const V3 connecting_vector = vx - o.fw_point[w];
              const double vertex_wall_distance = dot(connecting_vector, o.fw_normal[w]);
              if (std::fabs(vertex_wall_distance) < maximum_cell_diameter)
                in = true;
            }
          if (in)
            o.fw_cells[w].push_back(cell);
        }
  }

  // ---------------------------------------- update_fine_search_candidates ----
  // (update_fine_search_candidates.cc:9-212), local particle-particle flavour
  void update_fine_search_candidates_pp(DenseRows<PPRow> &pairs_in_contact, DenseRows<CandRow> &contact_candidates)
  {
    for (size_t r = 0; r < pairs_in_contact.rows.size();)
      {
        PPRow &row = pairs_in_contact.rows[r];
        const uint32_t particle_id = row.one;
        CandRow *cand = contact_candidates.find(particle_id);
        for (size_t k = 0; k < row.second.size();)
          {
            const uint32_t object_id = row.second[k].two;
            if (cand)
              {
                auto it = std::find(cand->c.begin(), cand->c.end(), object_id);
                if (it != cand->c.end())
                  {
                    cand->c.erase(it);
                    ++k;
                    continue;
                  }
              }
            CandRow *cand2 = contact_candidates.find(object_id);
            if (cand2)
              {
                auto it = std::find(cand2->c.begin(), cand2->c.end(), particle_id);
                if (it != cand2->c.end())
                  {
                    cand2->c.erase(it);
                    ++k;
                    continue;
                  }
              }
            swap_erase(row.second, k);
          }
        if (!row.second.empty())
          ++r;
        else
          pairs_in_contact.erase_at(r);
      }
  }
  // particle-wall / floating-wall flavour (:163-197)
  void update_fine_search_candidates_pw(DenseRows<PWRow> &pairs_in_contact, DenseRows<WCandRow> &contact_candidates)
  {
    for (size_t r = 0; r < pairs_in_contact.rows.size();)
      {
        PWRow &row = pairs_in_contact.rows[r];
        WCandRow *cand = contact_candidates.find(row.one);
        for (size_t k = 0; k < row.second.size();)
          {
            const uint32_t object_id = row.second[k].face;
            if (cand)
              {
                size_t q = 0;
                for (; q < cand->c.size(); ++q)
                  if (cand->c[q].face == object_id)
                    break;
                if (q < cand->c.size())
                  {
                    swap_erase(cand->c, q);
                    ++k;
                    continue;
                  }
              }
            swap_erase(row.second, k);
          }
        if (!row.second.empty())
          ++r;
        else
          pairs_in_contact.erase_at(r);
      }
  }

  // update_contact_container_iterators (update_local_particle_containers.cc:40-200)
  void update_contact_container_iterators_pp(Oracle &o, DenseRows<PPRow> &pairs)
  {
    for (size_t r = 0; r < pairs.rows.size();)
      {
        PPRow &row = pairs.rows[r];
        if (slot(o, row.one) < 0)
          {
            pairs.erase_at(r);
            continue;
          }
        for (size_t k = 0; k < row.second.size();)
          {
            if (slot(o, row.second[k].two) < 0)
              {
                swap_erase(row.second, k);
                continue;
              }
            if (o.clear_tangential_displacement_trigger)
              {
                row.second[k].tangential_displacement = mk(0, 0, 0);
                row.second[k].rolling_resistance_spring_torque = mk(0, 0, 0);
              }
            ++k;
          }
        ++r;
      }
  }
  void update_contact_container_iterators_pw(Oracle &o, DenseRows<PWRow> &pairs)
  {
    for (size_t r = 0; r < pairs.rows.size();)
      {
        if (slot(o, pairs.rows[r].one) < 0)
          {
            pairs.erase_at(r);
            continue;
          }
        ++r;
      }
  }

  // ----------------------------------------------------------- fine search ---
  // particle_particle_fine_search (particle_particle_fine_search.cc:19-232)
  void particle_particle_fine_search(Oracle &o, DenseRows<PPRow> &adjacent, DenseRows<CandRow> &candidates,
                                     bool periodic)
  {
    const double neighborhood_threshold = o.neighborhood_threshold_squared;
    for (auto &row : adjacent.rows)
      {
        if (row.second.empty())
          continue;
        const V3 one = o.parts[slot(o, row.one)].x;
        for (size_t k = 0; k < row.second.size();)
          {
            const V3 two = o.parts[slot(o, row.second[k].two)].x;
            if (!periodic)
              {
                const double square_distance = distance_square(one, two);
                if (square_distance > neighborhood_threshold)
                  swap_erase(row.second, k);
                else
                  ++k;
              }
            else
              {
                double min_square_distance = DBL_MAX;
                V3 nearest = mk(0, 0, 0);
                for (const V3 &t : o.combined_periodic_offsets)
                  {
                    const double d2 = distance_square(one, two + t);
                    if (d2 < min_square_distance)
                      {
                        min_square_distance = d2;
                        nearest = t;
                      }
                  }
                if (min_square_distance > neighborhood_threshold)
                  swap_erase(row.second, k);
                else
                  {
                    row.second[k].periodic_offset = nearest;
                    ++k;
                  }
              }
          }
      }
    for (auto &crow : candidates.rows)
      {
        if (crow.c.empty())
          continue;
        const V3 one = o.parts[slot(o, crow.one)].x;
        for (uint32_t two_id : crow.c)
          {
            const V3 two = o.parts[slot(o, two_id)].x;
            V3 offset = mk(0, 0, 0);
            bool add;
            if (!periodic)
              add = distance_square(one, two) < neighborhood_threshold;
            else
              {
                double min_square_distance = DBL_MAX;
                for (const V3 &t : o.combined_periodic_offsets)
                  {
                    const double d2 = distance_square(one, two + t);
                    if (d2 < min_square_distance)
                      {
                        min_square_distance = d2;
                        offset = t;
                      }
                  }
                add = min_square_distance < neighborhood_threshold;
              }
            if (add)
              {
                PPRow &row = adjacent.get_or_create(crow.one);
                bool exists = false;
                for (auto &e : row.second)
                  if (e.two == two_id)
                    exists = true;
                if (!exists)
                  row.second.push_back(PPInfo{two_id, mk(0, 0, 0), mk(0, 0, 0), offset});
              }
          }
      }
  }

  // particle_wall_fine_search (particle_wall_fine_search.cc:18-70)
  void particle_wall_fine_search(Oracle &o)
  {
    for (auto &crow : o.wall_candidates.rows)
      for (auto &c : crow.c)
        {
          PWRow &row = o.wall_in_contact.get_or_create(crow.one);
          bool exists = false;
          for (auto &e : row.second)
            if (e.face == c.face)
              exists = true;
          if (!exists)
            row.second.push_back(PWInfo{c.face, c.normal, c.point, c.boundary_id, mk(0, 0, 0), mk(0, 0, 0)});
        }
  }
  // particle_floating_wall_fine_search (particle_wall_fine_search.cc:72-166)
  void particle_floating_wall_fine_search(Oracle &o, double simulation_time)
  {
    for (auto &crow : o.fwall_candidates.rows)
      for (auto &c : crow.c)
        {
          const int w = int(c.face);
          if (!(simulation_time >= o.fw_t0[w] && simulation_time <= o.fw_t1[w]))
            continue;
          V3 normal_vector = o.fw_normal[w];
          const V3 connecting_vector = o.parts[slot(o, crow.one)].x - o.fw_point[w];
          const double ip = dot(connecting_vector, normal_vector);
          if (ip < 0)
            normal_vector = -1 * normal_vector;
          PWRow &row = o.fwall_in_contact.get_or_create(crow.one);
          bool exists = false;
          for (auto &e : row.second)
            if (e.face == c.face)
              exists = true;
          if (!exists)
            row.second.push_back(PWInfo{c.face, normal_vector, o.fw_point[w], LETHE_DEM_FLOATING_WALL_BOUNDARY_ID,
                                        mk(0, 0, 0), mk(0, 0, 0)});
        }
  }

  // ------------------------------------------------- rolling resistance (pp) --
  // rolling_resistance_torque_models.h:12-250, dispatch …contact_force.h:643-698
  V3 pp_rolling_resistance(const Oracle &o, double effective_r, const double *p1, const double *p2,
                           double rolling_friction_coeff, double rolling_viscous_damping_coeff, double dt,
                           double normal_spring_constant, double normal_force_norm, const V3 &n, V3 &cumulative)
  {
    switch (o.cfg.rolling_model)
      {
        case LETHE_ROLLING_NONE:
          return mk(0, 0, 0);
        case LETHE_ROLLING_CONSTANT:
          {
            const V3 w1 = mk(p1[P_WX], p1[P_WY], p1[P_WZ]), w2 = mk(p2[P_WX], p2[P_WY], p2[P_WZ]);
            const V3 omega_ij = w1 - w2;
            const V3 dir = omega_ij / (norm(omega_ij) + DBL_MIN);
            return (-rolling_friction_coeff * effective_r * normal_force_norm) * dir;
          }
        case LETHE_ROLLING_VISCOUS:
          {
            const V3 w1 = mk(p1[P_WX], p1[P_WY], p1[P_WZ]), w2 = mk(p2[P_WX], p2[P_WY], p2[P_WZ]);
            const V3 omega_ij = w1 - w2;
            const V3 dir = omega_ij / (norm(omega_ij) + DBL_MIN);
            const V3 v_omega = cross(w1, (p1[P_DP] * 0.5) * n) - cross(w2, (p2[P_DP] * 0.5) * (-n));
            return (-rolling_friction_coeff * effective_r * normal_force_norm * norm(v_omega)) * dir;
          }
        default: // EPSD
          {
            const double mu_r_times_R_e = rolling_friction_coeff * effective_r;
            V3 omega_ij;
            for (int d = 0; d < 3; ++d)
              omega_ij[d] = p1[P_WX + d] - p2[P_WX + d];
            const V3 omega_perp = omega_ij - dot(omega_ij, n) * n;
            const V3 delta_theta = dt * omega_perp;
            const double K_r = 2.25 * normal_spring_constant * sq(mu_r_times_R_e);
            cumulative = cumulative - K_r * delta_theta;
            const double M_r_max = mu_r_times_R_e * normal_force_norm;
            const double spring_norm = norm(cumulative);
            const double I_i = 1.4 * p1[P_MASS] * sq(0.5 * p1[P_DP]);
            const double I_j = 1.4 * p2[P_MASS] * sq(0.5 * p2[P_DP]);
            const double I_e = I_i * I_j / (I_i + I_j);
            const double C_r = rolling_viscous_damping_coeff * 2. * std::sqrt(I_e * K_r);
            if (spring_norm > M_r_max)
              {
                cumulative = cumulative * (M_r_max / spring_norm);
                return cumulative - (o.cfg.f_coefficient_epsd * C_r) * omega_perp;
              }
            return cumulative - C_r * omega_perp;
          }
      }
  }

  struct PPOut
  {
    V3 normal_force, tangential_force, t1, t2, rolling;
  };

  // Ferrari solution of the JKR contact-patch quartic (…contact_force.h:1356-1372).
  inline double jkr_contact_radius(double R, double overlap, double gamma, double Y, bool clamp_root1)
  {
    const double c0 = sq(R * overlap);
    const double c1 = -2. * sq(R) * M_PI * gamma / Y;
    const double c2 = -2. * overlap * R;
    const double P = -sq(c2) / 12. - c0;
    const double Q = -cube(c2) / 108. + c0 * c2 / 3. - sq(c1) * 0.125;
    double root1 = clamp_root1 ? std::max(0., (0.25 * sq(Q)) + (cube(P) / 27.)) : 0.25 * sq(Q) + cube(P) / 27.;
    const double U = std::cbrt(-0.5 * Q + std::sqrt(root1));
    const double s = -c2 * (5. / 6.) + U - P / (3. * U);
    const double w = std::sqrt(std::max(1e-16, c2 + 2. * s));
    const double lambda = 0.5 * c1 / w;
    const double root2 = std::max(1e-16, w * w - 4. * (c2 + s + lambda));
    return 0.5 * (w + std::sqrt(root2));
  }

  // calculate_contact for all particle-particle models (…contact_force.h:723-1544).
  // `out` is NOT reset here: the reference keeps these tensors across the pairs of
  // one row (…contact_force.h:1847-1853), which matters for the DMT non-contact branch.
  void pp_calculate_contact(const Oracle &o, int model, PPInfo &info, const V3 &vt, double vn, const V3 &n,
                            double overlap, double dt, const double *p1, const double *p2, PPOut &out)
  {
    const double d1 = p1[P_DP], d2 = p2[P_DP];
    const double effective_radius = (d1 * d2) / (2 * (d1 + d2));
    const double effective_mass = (p1[P_MASS] * p2[P_MASS]) / (p1[P_MASS] + p2[P_MASS]);
    const unsigned int t1 = static_cast<unsigned int>(p1[P_TYPE]);
    const unsigned int t2 = static_cast<unsigned int>(p2[P_TYPE]);
    const unsigned int k = t1 * o.n_types + t2;
    const double Y = o.eY[k], G = o.eG[k], beta = o.beta[k], mu = o.eMu[k];
    const double roll_visc = o.eRollVisc[k], roll_fric = o.eRollFric[k];

    if (model == LETHE_PP_DMT)
      {
        // calculate_DMT_contact (:1465-1544)
        constexpr double M_2PI = 2. * M_PI;
        const double gamma = o.eGamma[k], A = o.eHamaker[k];
        const double F_po = M_2PI * effective_radius * gamma;
        const double delta_0 = -std::sqrt(A * effective_radius / (6. * F_po));
        double cohesive_term;
        if (overlap > 0.)
          {
            cohesive_term = -F_po;
            pp_calculate_contact(o, LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP, info, vt, vn, n, overlap, dt, p1, p2, out);
          }
        else if (overlap > delta_0)
          {
            cohesive_term = -F_po;
            info.tangential_displacement = mk(0, 0, 0);
            info.rolling_resistance_spring_torque = mk(0, 0, 0);
          }
        else
          {
            cohesive_term = -A * effective_radius / (6. * sq(overlap));
            info.tangential_displacement = mk(0, 0, 0);
            info.rolling_resistance_spring_torque = mk(0, 0, 0);
          }
        out.normal_force = out.normal_force + cohesive_term * n;
        return;
      }

    if (model == LETHE_PP_LINEAR)
      {
        // calculate_linear_contact (:723-852)
        constexpr double characteristic_velocity = 1.0;
        const double kn = 1.0667 * std::sqrt(effective_radius) * Y *
                          std::pow((0.9375 * effective_mass * characteristic_velocity * characteristic_velocity /
                                    (std::sqrt(effective_radius) * Y)),
                                   0.2);
        const double kt = kn * 0.4;
        const double etan = -2 * beta * std::sqrt(effective_mass * kn);
        const double etat = etan * 0.6324555320336759;
        const double normal_force_value = kn * overlap + etan * vn;
        out.normal_force = normal_force_value * n;
        const V3 damping_tangential_force = etat * vt;
        out.tangential_force = (kt * info.tangential_displacement) + damping_tangential_force;
        const double coulomb_threshold = mu * normal_force_value;
        if (norm(out.tangential_force) > coulomb_threshold)
          {
            const V3 limited = coulomb_threshold * (out.tangential_force / (norm(out.tangential_force) + DBL_MIN));
            info.tangential_displacement = (limited - damping_tangential_force) / (kt + DBL_MIN);
            out.tangential_force = (kt * info.tangential_displacement) + damping_tangential_force;
          }
        out.t1 = cross(n, out.tangential_force * d1 * 0.5);
        out.t2 = out.t1 * d2 / d1;
        // note the swapped coefficient order in the reference (:841-851)
        out.rolling = pp_rolling_resistance(o, effective_radius, p1, p2, roll_visc, roll_fric, dt, kn,
                                            norm(out.normal_force), n, info.rolling_resistance_spring_torque);
        return;
      }

    const double radius_times_overlap_sqrt = std::sqrt(effective_radius * overlap);
    const double model_parameter_sn = 2.0 * Y * radius_times_overlap_sqrt;
    const double model_parameter_st = 8.0 * G * radius_times_overlap_sqrt;

    if (model == LETHE_PP_HERTZ_JKR)
      {
        // calculate_hertz_JKR_contact (:1304-1441)
        const double gamma = o.eGamma[k];
        const double a = jkr_contact_radius(effective_radius, overlap, gamma, Y, false);
        const double etan = -1.8257 * beta * std::sqrt(model_parameter_sn * effective_mass);
        const double kt = 8.0 * radius_times_overlap_sqrt * G;
        const double etat = etan * std::sqrt(model_parameter_st / model_parameter_sn);
        const double normal_force_coefficient =
          4. * cube(a) / (3. * effective_radius) * Y - std::sqrt(8. * M_PI * gamma * Y * cube(a));
        out.normal_force = (normal_force_coefficient + etan * vn) * n;
        out.tangential_force = kt * info.tangential_displacement + etat * vt;
        const double two_pull_off_force = 3. * M_PI * gamma * effective_radius;
        const double modified_coulomb_threshold = (normal_force_coefficient + two_pull_off_force) * mu;
        if (norm(out.tangential_force) > modified_coulomb_threshold)
          out.tangential_force =
            modified_coulomb_threshold * (out.tangential_force / (norm(out.tangential_force) + DBL_MIN));
        out.t1 = cross(n, out.tangential_force * d1 * 0.5);
        out.t2 = out.t1 * d2 / d1;
        const double kn = 0.66665 * model_parameter_sn;
        out.rolling = pp_rolling_resistance(o, effective_radius, p1, p2, roll_fric, roll_visc, dt, kn,
                                            norm(out.normal_force), n, info.rolling_resistance_spring_torque);
        return;
      }

    // hertz_mindlin_limit_overlap (:879-1010), limit_force (:1036-1151), hertz (:1176-1280)
    const double kn = 0.66665 * model_parameter_sn;
    const double etan = -1.8257 * beta * std::sqrt(model_parameter_sn * effective_mass);
    const double kt = 8.0 * G * radius_times_overlap_sqrt;
    const double normal_force_value = kn * overlap + etan * vn;
    out.normal_force = normal_force_value * n;
    const double coulomb_threshold_base = mu; // multiplied below, kept as in the reference
    if (model == LETHE_PP_HERTZ)
      {
        out.tangential_force = kt * info.tangential_displacement;
        const double coulomb_threshold = coulomb_threshold_base * normal_force_value;
        if (norm(out.tangential_force) > coulomb_threshold)
          out.tangential_force = coulomb_threshold * (out.tangential_force / (norm(out.tangential_force) + DBL_MIN));
      }
    else
      {
        const double etat = etan * std::sqrt(model_parameter_st / model_parameter_sn);
        const V3 damping_tangential_force = etat * vt;
        out.tangential_force = (kt * info.tangential_displacement) + damping_tangential_force;
        const double coulomb_threshold = coulomb_threshold_base * normal_force_value;
        if (norm(out.tangential_force) > coulomb_threshold)
          {
            if (model == LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP)
              {
                const V3 limited =
                  coulomb_threshold * (out.tangential_force / (norm(out.tangential_force) + DBL_MIN));
                info.tangential_displacement = (limited - damping_tangential_force) / (kt + DBL_MIN);
                out.tangential_force = (kt * info.tangential_displacement) + damping_tangential_force;
              }
            else
              out.tangential_force =
                coulomb_threshold * (out.tangential_force / (norm(out.tangential_force) + DBL_MIN));
          }
      }
    out.t1 = cross(n, out.tangential_force * d1 * 0.5);
    out.t2 = out.t1 * d2 / d1;
    out.rolling = pp_rolling_resistance(o, effective_radius, p1, p2, roll_fric, roll_visc, dt, kn,
                                        norm(out.normal_force), n, info.rolling_resistance_spring_torque);
  }

  // update_contact_information (…contact_force.h:223-298)
  inline void pp_update_contact_information(PPInfo &info, V3 &vt, double &vn, V3 &n, const double *p1,
                                            const double *p2, const V3 &x1, const V3 &x2, double dt)
  {
    const V3 contact_vector = x2 - x1;
    n = contact_vector / norm(contact_vector);
    V3 vrel = mk(p1[P_VX] - p2[P_VX], p1[P_VY] - p2[P_VY], p1[P_VZ] - p2[P_VZ]);
    const V3 w1 = mk(p1[P_WX], p1[P_WY], p1[P_WZ]), w2 = mk(p2[P_WX], p2[P_WY], p2[P_WZ]);
    vrel = vrel + cross(0.5 * (p1[P_DP] * w1 + p2[P_DP] * w2), n);
    vn = dot(vrel, n);
    vt = vrel - (vn * n);
    info.tangential_displacement = info.tangential_displacement + vt * dt;
    info.tangential_displacement =
      info.tangential_displacement - dot(info.tangential_displacement, n) * n;
  }

  // execute_contact_calculation<local / local_periodic> (…contact_force.h:1838-2063)
  // ---------------------------------------------------------- heat transfer ----

  // boost::math::erfc_inv restated: Newton on glibc's erfc, to the last bits of double
  double erfc_inv(double y)
  {
    // start value from the asymptotic / series form, then Newton: f = erfc(x) - y, f' = -2/sqrt(pi) exp(-x^2)
    double x;
    if (y >= 1.0)
      x = -std::sqrt(std::fabs(std::log((2.0 - y) * 0.8862269254527580136)));
    else
      x = std::sqrt(std::fabs(std::log(y * 0.8862269254527580136 + 1e-300)));
    if (y > 0.25 && y < 1.75)
      x = (1.0 - y) * 0.8862269254527580136;
    for (int it = 0; it < 60; ++it)
      {
        const double f = std::erfc(x) - y;
        const double dfdx = -1.1283791670955125739 * std::exp(-x * x);
        const double step = f / dfdx;
        x -= step;
        if (std::fabs(step) <= 1e-17 * std::fabs(x))
          break;
      }
    return x;
  }

  // particle_heat_transfer.cc:9-146
  double calculate_corrected_contact_radius(double effective_radius, double effective_youngs_modulus, double effective_real_youngs_modulus,
                                            double normal_force_norm)
  {
    const double contact_radius = std::pow((3 * normal_force_norm * effective_radius) / (4 * effective_youngs_modulus), (1.0 / 3.0));
    return contact_radius * std::pow(effective_youngs_modulus / effective_real_youngs_modulus, 1.0 / 5.0);
  }
  double calculate_macrocontact_resistance(double harmonic_conductivity, double contact_radius)
  {
    return 0.5 / (contact_radius * harmonic_conductivity + DBL_MIN);
  }
  double calculate_microcontact_resistance(double equivalent_surface_slope, double equivalent_surface_roughness, double effective_microhardness,
                                           double contact_radius_squared, double harmonic_conductivity, double maximum_pressure)
  {
    return 1.184 / (M_PI * harmonic_conductivity * contact_radius_squared) * (equivalent_surface_roughness / equivalent_surface_slope) *
           std::pow(effective_microhardness / (maximum_pressure + DBL_MIN), 0.96);
  }
  double calculate_solid_macrogap_resistance(double radius, double thermal_conductivity, double contact_radius_squared)
  {
    return 0.25 * M_PI * radius / (M_PI * (radius * radius - contact_radius_squared) * thermal_conductivity);
  }
  double calculate_interstitial_gas_microgap_resistance(double equivalent_surface_roughness, double contact_radius_squared, double gas_parameter_m,
                                                        double thermal_conductivity_gas, double maximum_pressure, double effective_microhardness)
  {
    const double x_1 = 2.0 * maximum_pressure / effective_microhardness;
    const double x_2 = 0.03 * maximum_pressure / effective_microhardness;
    if (x_1 >= 2.0 || x_1 <= 0.0)
      return INFINITY;
    const double a_1 = erfc_inv(x_1);
    const double a_2 = erfc_inv(x_2) - a_1;
    return (2.82842712475 * equivalent_surface_roughness * a_2) /
           (M_PI * thermal_conductivity_gas * contact_radius_squared *
            std::log(std::abs(1 + a_2 / (a_1 + gas_parameter_m / (2.82842712475 * equivalent_surface_roughness)))));
  }
  double calculate_interstitial_gas_macrogap_resistance(double harmonic_radius, double thermal_conductivity_gas, double contact_radius_squared,
                                                        double gas_parameter_m)
  {
    const double A = 2. * std::sqrt(harmonic_radius * harmonic_radius - contact_radius_squared);
    const double S = 2. * harmonic_radius - contact_radius_squared / harmonic_radius + gas_parameter_m;
    return 2.0 / (M_PI * thermal_conductivity_gas * (S * std::log(S / (S - A)) - A));
  }
  // calculate_contact_thermal_conductance<particle-particle> (particle_heat_transfer.cc:148-318); `parts` receives the
  // intermediate values the reference's unit test prints
  double calculate_contact_thermal_conductance(double radius_one, double radius_two, double effective_youngs_modulus,
                                               double effective_real_youngs_modulus, double equivalent_surface_roughness,
                                               double equivalent_surface_slope, double effective_microhardness, double thermal_conductivity_one,
                                               double thermal_conductivity_two, double thermal_conductivity_gas, double gas_parameter_m,
                                               double normal_overlap, double normal_force_norm, double *parts = nullptr)
  {
    const double harmonic_conductivity = harmonic_mean(thermal_conductivity_one, thermal_conductivity_two);
    const double harmonic_radius = harmonic_mean(radius_one, radius_two);
    const double contact_radius =
      calculate_corrected_contact_radius(harmonic_radius * 0.5, effective_youngs_modulus, effective_real_youngs_modulus, normal_force_norm);
    const double corrected_normal_overlap = normal_overlap * std::pow(effective_youngs_modulus / effective_real_youngs_modulus, 2.0 / 3.0);
    const double contact_radius_squared = contact_radius * contact_radius;
    const double maximum_pressure = (2.0 * effective_real_youngs_modulus * corrected_normal_overlap) / (M_PI * contact_radius + DBL_MIN);
    const double resistance_macrocontact = calculate_macrocontact_resistance(harmonic_conductivity, contact_radius);
    const double resistance_microcontact = calculate_microcontact_resistance(equivalent_surface_slope, equivalent_surface_roughness,
                                                                             effective_microhardness, contact_radius_squared,
                                                                             harmonic_conductivity, maximum_pressure);
    double resistance_solid_macrogap = calculate_solid_macrogap_resistance(radius_one, thermal_conductivity_one, contact_radius_squared);
    resistance_solid_macrogap += calculate_solid_macrogap_resistance(radius_two, thermal_conductivity_two, contact_radius_squared);
    const double resistance_gas_microgap = calculate_interstitial_gas_microgap_resistance(
      equivalent_surface_roughness, contact_radius_squared, gas_parameter_m, thermal_conductivity_gas, maximum_pressure, effective_microhardness);
    const double resistance_gas_macrogap =
      calculate_interstitial_gas_macrogap_resistance(harmonic_radius, thermal_conductivity_gas, contact_radius_squared, gas_parameter_m);
    const double thermal_conductance =
      1.0 / (resistance_macrocontact + 1.0 / (1.0 / resistance_microcontact + 1.0 / resistance_gas_microgap)) +
      1.0 / (resistance_solid_macrogap + resistance_gas_macrogap);
    if (parts)
      {
        parts[0] = contact_radius;
        parts[1] = resistance_macrocontact;
        parts[2] = resistance_microcontact;
        parts[3] = resistance_solid_macrogap;
        parts[4] = resistance_gas_microgap;
        parts[5] = resistance_gas_macrogap;
        parts[6] = 1.0 / thermal_conductance;
        parts[7] = thermal_conductance;
      }
    return thermal_conductance;
  }

  void pp_execute_contact_calculation(Oracle &o, PPRow &row, bool periodic, double dt)
  {
    if (row.second.empty())
      return;
    PPOut out;
    out.normal_force = out.tangential_force = out.t1 = out.t2 = out.rolling = mk(0, 0, 0);
    V3 n = mk(0, 0, 0), vt = mk(0, 0, 0);
    double vn = 0;
    const int s1 = slot(o, row.one);
    const double *p1 = o.parts[s1].p;
    const V3 x1 = o.parts[s1].x;
    for (auto &info : row.second)
      {
        const int s2 = slot(o, info.two);
        const double *p2 = o.parts[s2].p;
        const V3 x2 = periodic ? (o.parts[s2].x + info.periodic_offset) : o.parts[s2].x;
        const double normal_overlap = 0.5 * (p1[P_DP] + p2[P_DP]) - std::sqrt(distance_square(x1, x2));
        if (normal_overlap > o.pp_force_threshold)
          {
            // The reference declares the scratch tensors once per row (…force.h:1847-1853). Every
            // model overwrites them except the DMT non-contact branches (:1514-1532), which ADD the
            // cohesive term to whatever the previous pair of the row left behind and then apply the
            // stale tangential force / torques too. dmt_stale_scratch = true reproduces that;
            // false gives the evident intent (zero scratch per pair), which is what the CUDA
            // engine implements because the artefact depends on hash-map iteration order.
            if (!o.dmt_stale_scratch)
              out.normal_force = out.tangential_force = out.t1 = out.t2 = out.rolling = mk(0, 0, 0);
            pp_update_contact_information(info, vt, vn, n, p1, p2, x1, x2, dt);
            pp_calculate_contact(o, o.cfg.pp_model, info, vt, vn, n, normal_overlap, dt, p1, p2, out);
            // apply_force_and_torque_on_local_particles (:548-570)
            const V3 total_force = out.normal_force + out.tangential_force;
            o.force[s1] = o.force[s1] - total_force;
            o.force[s2] = o.force[s2] + total_force;
            o.torque[s1] = o.torque[s1] + (-out.t1 + out.rolling);
            o.torque[s2] = o.torque[s2] + (-out.t2 - out.rolling);
            ++o.n_touching_last;
          }
        else
          {
            info.tangential_displacement = mk(0, 0, 0);
            info.rolling_resistance_spring_torque = mk(0, 0, 0);
          }
        // DEM-MP (…contact_force.h:2065-2150): conduction through every contact with a positive overlap
        if (o.thermal_enabled && normal_overlap > 0)
          {
            const int t1 = int(p1[P_TYPE]), t2 = int(p2[P_TYPE]);
            const int k = t1 * o.n_types + t2;
            const uint32_t id1 = o.parts[s1].id, id2 = o.parts[s2].id;
            const double thermal_conductance = calculate_contact_thermal_conductance(
              0.5 * p1[P_DP], 0.5 * p2[P_DP], o.eY[k], o.th_real_E[k], o.th_roughness[k], o.th_slope[k], o.th_microhardness[k],
              o.th_conductivity[t1], o.th_conductivity[t2], o.th_conductivity_gas, o.th_gas_m[k], normal_overlap, norm(out.normal_force));
            // apply_heat_transfer_on_local_particles (particle_heat_transfer.cc:320-331)
            const double heat_transfer_rate = thermal_conductance * (o.temperature[id2] - o.temperature[id1]);
            o.heat_transfer_rate[s1] += heat_transfer_rate;
            o.heat_transfer_rate[s2] -= heat_transfer_rate;
          }
      }
  }

  // ------------------------------------------------------------ wall force ---
  const BoundaryMotion *find_motion(const Oracle &o, uint32_t boundary_id)
  {
    for (auto &m : o.motions)
      if (m.boundary_id == boundary_id)
        return &m;
    return nullptr;
  }

  // particle_wall_rolling_resistance_torque.h:12-226
  V3 pw_rolling_resistance(const Oracle &o, double R, const double *p, double rolling_friction_coeff,
                           double rolling_viscous_damping_coeff, double dt, double normal_spring_constant,
                           double normal_force_norm, const V3 &n, V3 &cumulative)
  {
    const V3 w = mk(p[P_WX], p[P_WY], p[P_WZ]);
    switch (o.cfg.rolling_model)
      {
        case LETHE_ROLLING_NONE:
          return mk(0, 0, 0);
        case LETHE_ROLLING_CONSTANT:
          {
            const double omega_value = norm(w);
            const V3 dir = w / (omega_value + DBL_MIN);
            return (-rolling_friction_coeff * R * normal_force_norm) * dir;
          }
        case LETHE_ROLLING_VISCOUS:
          {
            const double omega_value = norm(w);
            const V3 dir = w / (omega_value + DBL_MIN);
            const V3 v_omega = cross(w, R * n);
            return (-rolling_friction_coeff * R * normal_force_norm * norm(v_omega)) * dir;
          }
        default:
          {
            const double mu_r_times_R = rolling_friction_coeff * R;
            const V3 omega_perp = w - dot(w, n) * n;
            const V3 delta_theta = dt * omega_perp;
            const double K_r = 2.25 * normal_spring_constant * sq(mu_r_times_R);
            cumulative = cumulative - K_r * delta_theta;
            const double M_r_max = mu_r_times_R * normal_force_norm;
            const double spring_norm = norm(cumulative);
            const double I_e = 1.4 * p[P_MASS] * sq(R);
            const double C_r = rolling_viscous_damping_coeff * 2. * std::sqrt(I_e * K_r);
            if (spring_norm > M_r_max)
              {
                cumulative = cumulative * (M_r_max / spring_norm);
                return cumulative - (o.cfg.f_coefficient_epsd * C_r) * omega_perp;
              }
            return cumulative - C_r * omega_perp;
          }
      }
  }

  struct PWOut
  {
    V3 normal_force, tangential_force, tangential_torque, rolling;
  };

  // calculate_contact for particle-wall models (particle_wall_contact_force.h:610-1082)
  void pw_calculate_contact(const Oracle &o, int model, PWInfo &info, const V3 &vt, double vn, double overlap,
                            double dt, const double *p, PWOut &out)
  {
    const V3 normal_vector = -info.normal;
    const unsigned int type = static_cast<unsigned int>(p[P_TYPE]);
    const double Y = o.wY[type], G = o.wG[type], beta = o.wbeta[type], mu = o.wMu[type];
    const double roll_visc = o.wRollVisc[type], roll_fric = o.wRollFric[type];

    if (model == LETHE_PW_DMT)
      {
        constexpr double M_2PI = 2. * M_PI;
        const double R = 0.5 * p[P_DP];
        const double gamma = o.wGamma[type], A = o.wHamaker[type];
        const double F_po = M_2PI * R * gamma;
        const double delta_0 = -std::sqrt(A * R / (6. * F_po));
        double cohesive_term;
        if (overlap > 0.)
          {
            cohesive_term = -F_po;
            pw_calculate_contact(o, LETHE_PW_NONLINEAR, info, vt, vn, overlap, dt, p, out);
          }
        else if (overlap > delta_0)
          {
            cohesive_term = -F_po;
            info.tangential_displacement = mk(0, 0, 0);
            info.rolling_resistance_spring_torque = mk(0, 0, 0);
          }
        else
          {
            cohesive_term = -A * R / (6. * sq(overlap));
            info.tangential_displacement = mk(0, 0, 0);
            info.rolling_resistance_spring_torque = mk(0, 0, 0);
          }
        out.normal_force = out.normal_force + cohesive_term * normal_vector;
        return;
      }
    if (model == LETHE_PW_LINEAR)
      {
        const double R = p[P_DP] * 0.5;
        const double rp_sqrt = std::sqrt(R);
        const double kn = 1.0667 * rp_sqrt * Y * std::pow((0.9375 * p[P_MASS] * 1.0 * 1.0 / (rp_sqrt * Y)), 0.2);
        const double etan = 2 * beta * std::sqrt(p[P_MASS] * kn);
        const double kt = -kn * 0.4;
        const double etat = etan * 0.6324555320336759;
        out.normal_force = (kn * overlap + etan * vn) * normal_vector;
        out.tangential_force = (kt * info.tangential_displacement + etat * vt);
        const double coulomb_threshold = mu * norm(out.normal_force);
        if (norm(out.tangential_force) > coulomb_threshold)
          {
            out.tangential_force = coulomb_threshold * (out.tangential_force / norm(out.tangential_force));
            info.tangential_displacement = out.tangential_force / (kt + DBL_MIN);
          }
        out.tangential_torque = cross((R * normal_vector), -out.tangential_force);
        out.rolling = pw_rolling_resistance(o, R, p, roll_fric, roll_visc, dt, kn, norm(out.normal_force),
                                            info.normal, info.rolling_resistance_spring_torque);
        return;
      }
    if (model == LETHE_PW_JKR)
      {
        const double R = 0.5 * p[P_DP];
        const double gamma = o.wGamma[type];
        const double radius_times_overlap_sqrt = std::sqrt(R * overlap);
        const double sn = 2.0 * Y * radius_times_overlap_sqrt;
        const double st = 8.0 * G * radius_times_overlap_sqrt;
        const double a = jkr_contact_radius(R, overlap, gamma, Y, true);
        const double etan = 1.8257 * beta * std::sqrt(sn * p[P_MASS]);
        const double kt = -8.0 * G * radius_times_overlap_sqrt;
        const double etat = etan * std::sqrt(st / (sn + DBL_MIN));
        const double normal_force_norm =
          4. * Y * cube(a) / (3. * R) - std::sqrt(8. * M_PI * gamma * Y * cube(a)) + etan * vn;
        out.normal_force = normal_force_norm * normal_vector;
        const V3 damping_tangential_force = etat * vt;
        out.tangential_force = kt * info.tangential_displacement + damping_tangential_force;
        const double modified_coulomb_threshold = (normal_force_norm + 3. * M_PI * gamma * R) * mu;
        const double tangential_force_norm = norm(out.tangential_force);
        if (tangential_force_norm > modified_coulomb_threshold)
          {
            info.tangential_displacement =
              (modified_coulomb_threshold * (out.tangential_force / (tangential_force_norm + DBL_MIN)) -
               damping_tangential_force) /
              (kt + DBL_MIN);
            out.tangential_force = (kt * info.tangential_displacement) + damping_tangential_force;
          }
        out.tangential_torque = cross((R * normal_vector), -out.tangential_force);
        const double kn = 0.66665 * sn;
        out.rolling = pw_rolling_resistance(o, R, p, roll_fric, roll_visc, dt, kn, norm(out.normal_force),
                                            info.normal, info.rolling_resistance_spring_torque);
        return;
      }
    // nonlinear (:723-833)
    const double R = p[P_DP] * 0.5;
    const double radius_times_overlap_sqrt = std::sqrt(R * overlap);
    const double sn = 2.0 * Y * radius_times_overlap_sqrt;
    const double st = 8.0 * G * radius_times_overlap_sqrt;
    const double kn = 1.3333 * Y * radius_times_overlap_sqrt;
    const double etan = 1.8257 * beta * std::sqrt(sn * p[P_MASS]);
    const double kt = -8.0 * G * radius_times_overlap_sqrt + DBL_MIN;
    const double etat = etan * std::sqrt(st / sn);
    out.normal_force = (kn * overlap + etan * vn) * normal_vector;
    const V3 damping_tangential_force = etat * vt;
    out.tangential_force = kt * info.tangential_displacement + damping_tangential_force;
    const double coulomb_threshold = mu * norm(out.normal_force);
    const double tangential_force_norm = norm(out.tangential_force);
    if (tangential_force_norm > coulomb_threshold)
      {
        info.tangential_displacement =
          (coulomb_threshold * (out.tangential_force / (tangential_force_norm + DBL_MIN)) -
           damping_tangential_force) /
          (kt + DBL_MIN);
        out.tangential_force = (kt * info.tangential_displacement) + damping_tangential_force;
      }
    out.tangential_torque = cross((R * normal_vector), -out.tangential_force);
    out.rolling = pw_rolling_resistance(o, R, p, roll_fric, roll_visc, dt, kn, norm(out.normal_force), info.normal,
                                        info.rolling_resistance_spring_torque);
  }

  // calculate_particle_wall_contact (particle_wall_contact_force.cc:45-142) with
  // update_contact_information (particle_wall_contact_force.h:166-264)
  void calculate_particle_wall_contact(Oracle &o, DenseRows<PWRow> &pairs, double dt)
  {
    for (auto &row : pairs.rows)
      for (auto &info : row.second)
        {
          const int s = slot(o, row.one);
          const double *p = o.parts[s].p;
          const V3 x = o.parts[s].x;
          const V3 point_to_particle_vector = x - info.point;
          // find_projection (:~335)
          const V3 projected_vector =
            ((dot(point_to_particle_vector, info.normal)) / (norm_square(info.normal))) * info.normal;
          const double normal_overlap = ((p[P_DP]) * 0.5) - (norm(projected_vector));
          if (normal_overlap > o.pw_force_threshold)
            {
              const V3 normal_vector = -info.normal;
              const V3 particle_velocity = mk(p[P_VX], p[P_VY], p[P_VZ]);
              const V3 particle_angular_velocity = mk(p[P_WX], p[P_WY], p[P_WZ]);
              const V3 contact_point = x + (0.5 * p[P_DP]) * normal_vector;
              const BoundaryMotion *m = find_motion(o, info.boundary_id);
              const V3 bt = m ? m->translational_velocity : mk(0, 0, 0);
              const double bs = m ? m->rotational_speed : 0.;
              const V3 br = m ? m->rotational_vector : mk(0, 0, 0);
              const V3 bp = m ? m->point_on_axis : mk(0, 0, 0);
              V3 vector_to_rotating_axis = contact_point - bp;
              vector_to_rotating_axis = vector_to_rotating_axis - (dot(vector_to_rotating_axis, br)) * br;
              const V3 vrel = bt - particle_velocity +
                              cross(((-0.5 * p[P_DP]) * particle_angular_velocity), normal_vector) +
                              cross(bs * br, vector_to_rotating_axis);
              const double vn = dot(vrel, normal_vector);
              const V3 vt = vrel - (vn * normal_vector);
              info.tangential_displacement = info.tangential_displacement + vt * dt;

              PWOut out;
              out.normal_force = out.tangential_force = out.tangential_torque = out.rolling = mk(0, 0, 0);
              pw_calculate_contact(o, o.cfg.pw_model, info, vt, vn, normal_overlap, dt, p, out);
              // apply_force_and_torque (:506-522)
              const V3 total_force = out.normal_force + out.tangential_force;
              o.force[s] = o.force[s] - total_force;
              o.torque[s] = o.torque[s] + (out.tangential_torque + out.rolling);
            }
          else
            {
              info.tangential_displacement = mk(0, 0, 0);
              info.rolling_resistance_spring_torque = mk(0, 0, 0);
            }
        }
  }


  // ------------------------------------------------------- solid surfaces ----
  enum TriangleContact { TC_FACE = 0, TC_EDGE = 1, TC_VERTEX = 2, TC_NONE = 3 };

  // Eberly's closest point on a triangle, as both LetheGridTools functions evaluate it
  // (lethe_grid_tools.cc:1277-1434 and :1565-1697), quirks included (region 4: t = e / c).
  inline void closest_point_parameters(double a, double b, double c, double d, double e, double det, double &s, double &t,
                                       int &indicator)
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

  // LetheGridTools::find_point_triangle_distance (lethe_grid_tools.cc:1536-1700), dim = 3
  double find_point_triangle_distance(const V3 &p0, const V3 &p1, const V3 &p2, const V3 &point)
  {
    const V3 e_0 = p1 - p0, e_1 = p2 - p0;
    const double a = norm_square(e_0), b = dot(e_0, e_1), c = norm_square(e_1);
    const double det = a * c - b * b;
    const V3 vector_to_plane = p0 - point;
    const double d = dot(e_0, vector_to_plane), e = dot(e_1, vector_to_plane);
    double s, t;
    int ind;
    closest_point_parameters(a, b, c, d, e, det, s, t, ind);
    const V3 pt_in_triangle = p0 + s * e_0 + t * e_1;
    return std::sqrt(distance_square(pt_in_triangle, point));
  }

  // LetheGridTools::find_particle_triangle_projection (lethe_grid_tools.cc:1226-1450)
  bool find_particle_triangle_projection(const V3 &p0, const V3 &p1, const V3 &p2, const V3 &particle_position, double radius,
                                         V3 &projection, V3 &unit_normal_out, int &indicator)
  {
    const V3 e_0 = p1 - p0, e_1 = p2 - p0;
    V3 normal = cross(e_0, e_1);
    const double norm_normal = norm(normal);
    V3 unit_normal = normal / norm_normal;
    const double a = norm_square(e_0), b = dot(e_0, e_1), c = norm_square(e_1);
    const double det = a * c - b * b;
    const V3 vector_to_plane = p0 - particle_position;
    if (dot(vector_to_plane, unit_normal) > 0)
      unit_normal = unit_normal * -1.0;
    // (sic) a signed distance compared with a squared radius; it is never positive after the flip
    const double distance_squared = dot(vector_to_plane, unit_normal);
    if (distance_squared > (radius * radius))
      {
        indicator = TC_NONE;
        return false;
      }
    const double d = dot(e_0, vector_to_plane), e = dot(e_1, vector_to_plane);
    double s, t;
    closest_point_parameters(a, b, c, d, e, det, s, t, indicator);
    const V3 pt_in_triangle = p0 + s * e_0 + t * e_1;
    if (indicator == TC_FACE)
      unit_normal_out = unit_normal;
    else
      {
        normal = particle_position - pt_in_triangle;
        unit_normal_out = normal / norm(normal);
      }
    projection = pt_in_triangle;
    return true;
  }

  // SerialSolid::setup_containers (serial_solid.cc:617-682)
  void setup_solid_neighbors(Solid &sd)
  {
    const size_t nt = sd.tri.size();
    std::vector<std::set<uint32_t>> vertices_cell_map(sd.vertices.size());
    for (uint32_t t = 0; t < nt; ++t)
      for (int v = 0; v < 3; ++v)
        vertices_cell_map[sd.tri[t][v]].insert(t);
    sd.es_neighbors.assign(nt, {});
    sd.vs_neighbors.assign(nt, {});
    for (uint32_t t = 0; t < nt; ++t)
      {
        std::set<uint32_t> around;
        for (int v = 0; v < 3; ++v)
          around.insert(vertices_cell_map[sd.tri[t][v]].begin(), vertices_cell_map[sd.tri[t][v]].end());
        for (uint32_t n : around)
          {
            if (n == t)
              continue;
            unsigned int n_sharing_vertices = 0;
            for (int v = 0; v < 3; ++v)
              if (std::find(sd.tri[t].begin(), sd.tri[t].end(), sd.tri[n][v]) != sd.tri[t].end())
                n_sharing_vertices++;
            if (n_sharing_vertices == 1)
              sd.vs_neighbors[t].push_back(n);
            else
              sd.es_neighbors[t].push_back(n);
          }
      }
  }

  // SerialSolid::map_solid_in_background_triangulation (serial_solid.cc:83-150)
  void map_solid_in_background_triangulation(const Oracle &o, Solid &sd)
  {
    sd.mesh_info.clear();
    const double *h = o.cfg.cell_size;
    const double bg_cell_length = std::sqrt((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]);
    for (int r = 0; r < o.n_cells; ++r)
      {
        const int cell = o.cell_of_rank[r];
        const int ci = cell % o.nx, cj = (cell / o.nx) % o.ny, ck = cell / (o.nx * o.ny);
        const V3 bg_cell_center = mk(o.cfg.grid_lo[0] + (ci + 0.5) * h[0], o.cfg.grid_lo[1] + (cj + 0.5) * h[1],
                                     o.cfg.grid_lo[2] + (ck + 0.5) * h[2]);
        for (uint32_t t = 0; t < sd.tri.size(); ++t)
          {
            const double distance = find_point_triangle_distance(sd.vertices[sd.tri[t][0]], sd.vertices[sd.tri[t][1]],
                                                                 sd.vertices[sd.tri[t][2]], bg_cell_center);
            if (distance < bg_cell_length)
              sd.mesh_info.emplace_back(cell, t);
          }
      }
    for (auto &dsp : sd.displacement_since_mapped)
      dsp = mk(0, 0, 0);
  }

  // find_floating_mesh_mapping_step (find_contact_detection_step.cc:139-161); criterion dem.cc:296-308
  void find_floating_mesh_mapping_step(Oracle &o)
  {
    if (o.solids.empty())
      return;
    const double *h = o.cfg.cell_size;
    const double criterion = 0.57735026918962576451 * std::sqrt((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]);
    bool floating_mesh_requires_map = false;
    for (auto &sd : o.solids)
      {
        double displacement = 0.;
        for (auto &dsp : sd.displacement_since_mapped)
          for (int d = 0; d < 3; ++d)
            displacement = std::max(displacement, std::fabs(dsp[d]));
        floating_mesh_requires_map = floating_mesh_requires_map || displacement > criterion;
      }
    if (floating_mesh_requires_map)
      {
        o.solid_object_search_trigger = true;
        o.contact_search_trigger = true;
      }
  }

  // SerialSolid::move_solid_triangulation (serial_solid.cc:333-410)
  void move_solid_objects(Oracle &o)
  {
    const double time_step = o.cfg.dt;
    for (auto &sd : o.solids)
      {
        for (size_t v = 0; v < sd.vertices.size(); ++v)
          {
            const V3 distance_vector = sd.vertices[v] - sd.center_of_rotation;
            V3 local_velocity = sd.translational_velocity;
            local_velocity = local_velocity + cross(sd.angular_velocity, distance_vector);
            const V3 vertex_displacement = time_step * local_velocity;
            sd.vertices[v] = sd.vertices[v] + vertex_displacement;
            sd.displacement_since_mapped[v] = sd.displacement_since_mapped[v] + vertex_displacement;
          }
        sd.center_of_rotation = sd.center_of_rotation + sd.translational_velocity * time_step;
      }
  }

  // find_full_cell_neighbors (find_cell_neighbors.cc:293-333): the cell and its vertex-sharing cells
  void full_cell_neighbors(const Oracle &o, int cell, std::vector<int> &out)
  {
    out.clear();
    out.push_back(cell);
    const int ci = cell % o.nx, cj = (cell / o.nx) % o.ny, ck = cell / (o.nx * o.ny);
    for (int dk = -1; dk <= 1; ++dk)
      for (int dj = -1; dj <= 1; ++dj)
        for (int di = -1; di <= 1; ++di)
          {
            if (!di && !dj && !dk)
              continue;
            const int i = ci + di, j = cj + dj, k = ck + dk;
            if (i < 0 || j < 0 || k < 0 || i >= o.nx || j >= o.ny || k >= o.nz)
              continue;
            out.push_back(lin(o, i, j, k));
          }
  }

  // particle_solid_surfaces_contact_search (particle_wall_broad_search.cc:129-215)
  void particle_solid_surfaces_contact_search(Oracle &o)
  {
    std::vector<int> cell_list;
    for (auto &sd : o.solids)
      {
        sd.candidates.clear();
        for (auto &bt : sd.mesh_info)
          {
            full_cell_neighbors(o, bt.first, cell_list);
            for (int c : cell_list)
              {
                if (!o.cell_is_mobile(c)) // particle_wall_broad_search.cc:391-398
                  continue;
                for (int s : o.cell_parts[c])
                  sd.candidates[bt.second].insert(o.parts[s].id);
              }
          }
      }
  }

  // update_fine_search_candidates<particle_floating_mesh> (update_fine_search_candidates.cc:163-212)
  // + particle_floating_mesh_fine_search (particle_wall_fine_search.cc:170-210)
  void update_and_fine_search_solids(Oracle &o)
  {
    for (auto &sd : o.solids)
      {
        for (auto it = sd.in_contact.begin(); it != sd.in_contact.end();)
          {
            auto cand = sd.candidates.find(it->first);
            for (auto pit = it->second.begin(); pit != it->second.end();)
              {
                if (cand != sd.candidates.end())
                  {
                    auto f = cand->second.find(pit->first);
                    if (f != cand->second.end())
                      {
                        cand->second.erase(f);
                        ++pit;
                        continue;
                      }
                  }
                pit = it->second.erase(pit);
              }
            if (it->second.empty())
              it = sd.in_contact.erase(it);
            else
              ++it;
          }
        for (auto &c : sd.candidates)
          for (uint32_t pid : c.second)
            {
              PWInfo info;
              info.face = c.first;
              info.normal = info.point = mk(0, 0, 0);
              info.boundary_id = 0;
              info.tangential_displacement = info.rolling_resistance_spring_torque = mk(0, 0, 0);
              sd.in_contact[c.first].emplace(pid, info);
            }
      }
  }

  struct SolidContact
  {
    uint32_t triangle;
    double normal_overlap;
    int indicator;
    PWInfo *info;
  };

  // ParticleWallContactForce::calculate_particle_solid_object_contact
  // (particle_wall_contact_force.cc:153-580)
  void calculate_particle_solid_object_contact(Oracle &o, double dt)
  {
    for (size_t solid_counter = 0; solid_counter < o.solids.size(); ++solid_counter)
      {
        Solid &sd = o.solids[solid_counter];
        std::map<int, std::vector<SolidContact>> contact_record; // particle local index -> contacts in triangle order
        for (auto &tp : sd.in_contact)
          {
            const uint32_t t = tp.first;
            const V3 &p0 = sd.vertices[sd.tri[t][0]], &p1 = sd.vertices[sd.tri[t][1]], &p2 = sd.vertices[sd.tri[t][2]];
            for (auto &pi : tp.second)
              {
                PWInfo &contact_info = pi.second;
                const int s = slot(o, pi.first);
                if (s < 0)
                  continue;
                const double *pp = o.parts[s].p;
                V3 projection_point, normal_vector;
                int contact_indicator;
                if (!find_particle_triangle_projection(p0, p1, p2, o.parts[s].x, pp[P_DP] * 0.5, projection_point, normal_vector,
                                                       contact_indicator))
                  continue;
                const double particle_triangle_distance = std::sqrt(distance_square(o.parts[s].x, projection_point));
                const double normal_overlap = 0.5 * pp[P_DP] - particle_triangle_distance;
                if (normal_overlap > o.pw_force_threshold)
                  {
                    contact_info.normal = normal_vector;
                    contact_info.point = projection_point;
                    contact_info.boundary_id = uint32_t(solid_counter);
                    contact_record[s].push_back(SolidContact{t, normal_overlap, contact_indicator, &contact_info});
                  }
                else
                  {
                    contact_info.tangential_displacement = mk(0, 0, 0);
                    contact_info.rolling_resistance_spring_torque = mk(0, 0, 0);
                  }
              }
          }
        auto clear_contact_info = [](PWInfo &ci) {
          ci.tangential_displacement = mk(0, 0, 0);
          ci.rolling_resistance_spring_torque = mk(0, 0, 0);
        };
        auto has = [](const std::vector<uint32_t> &v, uint32_t x) { return std::find(v.begin(), v.end(), x) != v.end(); };
        for (auto &rec : contact_record)
          {
            const int particle_index = rec.first;
            std::vector<SolidContact> &R = rec.second;
            // double-contact elimination between connected triangles (:262-468)
            for (size_t c1 = 0; c1 < R.size();)
              {
                const uint32_t T1 = R[c1].triangle;
                const int I1 = R[c1].indicator;
                const auto &T1_es = sd.es_neighbors[T1];
                const auto &T1_vs = sd.vs_neighbors[T1];
                bool erase_contact_1 = false;
                size_t c2 = c1 + 1;
                while (c2 < R.size())
                  {
                    const uint32_t T2 = R[c2].triangle;
                    const int I2 = R[c2].indicator;
                    if (!has(T1_es, T2) && !has(T1_vs, T2))
                      {
                        ++c2;
                        continue;
                      }
                    if (I1 == TC_FACE)
                      {
                        if (I2 == TC_FACE)
                          {
                            ++c2;
                            continue;
                          }
                        if (I2 == TC_EDGE)
                          {
                            if (has(T1_vs, T2))
                              {
                                ++c2;
                                continue;
                              }
                            clear_contact_info(*R[c2].info);
                            R.erase(R.begin() + c2);
                            continue;
                          }
                        clear_contact_info(*R[c2].info);
                        R.erase(R.begin() + c2);
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
                            if (has(T1_es, T2))
                              {
                                clear_contact_info(*R[c2].info);
                                R.erase(R.begin() + c2);
                                continue;
                              }
                            else
                              {
                                ++c2;
                                continue;
                              }
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
                            if (has(T1_vs, T2))
                              {
                                erase_contact_1 = true;
                                break;
                              }
                            ++c2;
                            continue;
                          }
                        if (I2 == TC_VERTEX)
                          {
                            clear_contact_info(*R[c2].info);
                            R.erase(R.begin() + c2);
                            continue;
                          }
                      }
                    ++c2;
                  }
                if (erase_contact_1)
                  {
                    clear_contact_info(*R[c1].info);
                    R.erase(R.begin() + c1);
                    continue;
                  }
                ++c1;
              }
            const int s = particle_index;
            const double *pp = o.parts[s].p;
            for (auto &contact : R)
              {
                PWInfo &contact_info = *contact.info;
                // update_particle_solid_object_contact_information (particle_wall_contact_force.h:283-331)
                const V3 normal_vector = -contact_info.normal;
                const V3 particle_velocity = mk(pp[P_VX], pp[P_VY], pp[P_VZ]);
                const V3 particle_angular_velocity = mk(pp[P_WX], pp[P_WY], pp[P_WZ]);
                const double center_of_rotation_particle_distance = std::sqrt(distance_square(sd.center_of_rotation, o.parts[s].x));
                const V3 contact_relative_velocity =
                  sd.translational_velocity - particle_velocity +
                  cross((center_of_rotation_particle_distance * sd.angular_velocity - 0.5 * pp[P_DP] * particle_angular_velocity),
                        normal_vector);
                const double vn = dot(contact_relative_velocity, normal_vector);
                const V3 vt = contact_relative_velocity - vn * normal_vector;
                contact_info.tangential_displacement = contact_info.tangential_displacement + vt * dt;
                PWOut out;
                out.normal_force = out.tangential_force = out.tangential_torque = out.rolling = mk(0, 0, 0);
                pw_calculate_contact(o, o.cfg.pw_model, contact_info, vt, vn, contact.normal_overlap, dt, pp, out);
                const V3 total_force = out.normal_force + out.tangential_force;
                o.force[s] = o.force[s] - total_force;
                o.torque[s] = o.torque[s] + (out.tangential_torque + out.rolling);
              }
          }
      }
  }

  // ------------------------------------------------------------ integrator ---
  // ExplicitEulerIntegrator::integrate (explicit_euler_integrator.cc:69-130); integrate_start is
  // the same step (:14-26), integrate_end only zeroes force and torque (:32-48)
  void integrate_euler(Oracle &o)
  {
    const double dt = o.cfg.dt;
    for (size_t s = 0; s < o.parts.size(); ++s)
      {
        double *p = o.parts[s].p;
        const double mass_inverse = 1 / p[P_MASS];
        const double MOI_inverse = 1 / o.MOI[s];
        for (int d = 0; d < 3; ++d)
          {
            const double acceleration = o.cfg.g[d] + (o.force[s][d]) * mass_inverse;
            p[P_VX + d] += dt * acceleration;
            o.parts[s].x[d] += dt * p[P_VX + d];
            p[P_WX + d] += dt * (o.torque[s][d] * MOI_inverse);
          }
        o.force[s] = mk(0, 0, 0);
        o.torque[s] = mk(0, 0, 0);
      }
  }
  void zero_force_torque(Oracle &o)
  {
    for (size_t s = 0; s < o.parts.size(); ++s)
      {
        o.force[s] = mk(0, 0, 0);
        o.torque[s] = mk(0, 0, 0);
      }
  }

  // VelocityVerletIntegrator (velocity_verlet_integrator.cc:14-66,70-115,214-290)
  void integrate_start(Oracle &o)
  {
    if (o.cfg.integrator == LETHE_INTEGRATOR_EXPLICIT_EULER)
      return integrate_euler(o);
    const double dt = o.cfg.dt;
    const V3 g = mk(o.cfg.g[0], o.cfg.g[1], o.cfg.g[2]);
    const V3 half_dt_g = 0.5 * g * dt;
    for (size_t s = 0; s < o.parts.size(); ++s)
      {
        double *p = o.parts[s].p;
        const double half_dt_mass_inverse = 0.5 * dt / p[P_MASS];
        const double half_dt_MOI_inverse = 0.5 * dt / o.MOI[s];
        for (int d = 0; d < 3; ++d)
          {
            p[P_VX + d] += half_dt_g[d] + o.force[s][d] * half_dt_mass_inverse;
            p[P_WX + d] += o.torque[s][d] * half_dt_MOI_inverse;
          }
        for (int d = 0; d < 3; ++d)
          o.parts[s].x[d] += p[P_VX + d] * dt;
        o.force[s] = mk(0, 0, 0);
        o.torque[s] = mk(0, 0, 0);
      }
  }
  void integrate_end(Oracle &o)
  {
    if (o.cfg.integrator == LETHE_INTEGRATOR_EXPLICIT_EULER)
      return zero_force_torque(o);
    const double dt = o.cfg.dt;
    const V3 g = mk(o.cfg.g[0], o.cfg.g[1], o.cfg.g[2]);
    const V3 half_dt_g = 0.5 * g * dt;
    const bool asc = o.asc_active(); // velocity_verlet_integrator.cc:117-210
    for (size_t s = 0; s < o.parts.size(); ++s)
      {
        double *p = o.parts[s].p;
        if (asc && o.cell_mobility_status[o.parts[s].cell] != LETHE_MOBILITY_MOBILE)
          {
            o.force[s] = mk(0, 0, 0);
            o.torque[s] = mk(0, 0, 0);
            continue;
          }
        const double half_dt_mass_inverse = 0.5 * dt / p[P_MASS];
        const double half_dt_MOI_inverse = 0.5 * dt / o.MOI[s];
        for (int d = 0; d < 3; ++d)
          p[P_VX + d] += half_dt_g[d] + o.force[s][d] * half_dt_mass_inverse;
        for (int d = 0; d < 3; ++d)
          p[P_WX + d] += o.torque[s][d] * half_dt_MOI_inverse;
        o.force[s] = mk(0, 0, 0);
        o.torque[s] = mk(0, 0, 0);
      }
  }
  void integrate(Oracle &o)
  {
    if (o.cfg.integrator == LETHE_INTEGRATOR_EXPLICIT_EULER)
      return integrate_euler(o);
    const double dt = o.cfg.dt;
    const V3 g = mk(o.cfg.g[0], o.cfg.g[1], o.cfg.g[2]);
    const V3 dt_g = g * dt;
    // with adaptive sparse contacts (velocity_verlet_integrator.cc:292-436) only the particles of
    // mobile cells move; the others keep position and velocity, their force and torque are dropped
    const bool asc = o.asc_active();
    for (size_t s = 0; s < o.parts.size(); ++s)
      {
        double *p = o.parts[s].p;
        if (asc && o.cell_mobility_status[o.parts[s].cell] != LETHE_MOBILITY_MOBILE)
          {
            o.force[s] = mk(0, 0, 0);
            o.torque[s] = mk(0, 0, 0);
            continue;
          }
        const double dt_mass_inverse = dt / p[P_MASS];
        const double dt_MOI_inverse = dt / o.MOI[s];
        for (int d = 0; d < 3; ++d)
          p[P_VX + d] += dt_g[d] + o.force[s][d] * dt_mass_inverse;
        for (int d = 0; d < 3; ++d)
          o.parts[s].x[d] += p[P_VX + d] * dt;
        for (int d = 0; d < 3; ++d)
          p[P_WX + d] += o.torque[s][d] * dt_MOI_inverse;
        o.force[s] = mk(0, 0, 0);
        o.torque[s] = mk(0, 0, 0);
      }
  }

  // ------------------------------------------------------------ time step ----
  // find_particle_contact_detection_step (find_contact_detection_step.cc:9-59) with
  // check_contact_search_iteration_{dynamic,constant} (dem.cc:459-482)
  void contact_detection_iteration_check(Oracle &o)
  {
    const uint64_t freq = uint64_t(std::max(1, o.cfg.contact_detection_frequency));
    if (o.cfg.detection == LETHE_DETECTION_CONSTANT)
      {
        if ((o.iteration_number % freq) == 0)
          o.contact_search_trigger = true;
        return;
      }
    const bool parallel_update = (o.iteration_number % freq) == 0;
    if (o.contact_search_trigger)
      return;
    double max_displacement = 0.;
    const double dt = o.cfg.dt;
    for (size_t s = 0; s < o.parts.size(); ++s)
      {
        const double *p = o.parts[s].p;
        o.displacement[s] += dt * std::sqrt(p[P_VX] * p[P_VX] + p[P_VY] * p[P_VY] + p[P_VZ] * p[P_VZ]);
        max_displacement = std::max(max_displacement, o.displacement[s]);
      }
    const bool contact_detection_step = max_displacement > o.cfg.smallest_contact_search_criterion;
    if (parallel_update && contact_detection_step)
      o.contact_search_trigger = true;
  }

  // DEMSolver::execute_contact_detection_and_search (dem.cc:598-688)
  void execute_contact_detection_and_search(Oracle &o)
  {
    contact_detection_iteration_check(o);
    // solid objects: mapping onto the background mesh (dem.cc:605-628)
    find_floating_mesh_mapping_step(o);
    if (o.solid_object_search_trigger)
      for (auto &sd : o.solids)
        map_solid_in_background_triangulation(o, sd);
    if (!o.contact_search_trigger)
      return;
    execute_particles_displacement(o);
    sort_particles_into_subdomains_and_cells(o);
    identify_mobility_status(o); // dem.cc:639-644

    find_particle_particle_contact_pairs(o);
    if (o.periodic_enabled)
      find_particle_particle_periodic_contact_pairs(o);
    std::vector<int> faces_by_id(o.faces.size());
    for (size_t f = 0; f < o.faces.size(); ++f)
      faces_by_id[f] = int(f);
    std::stable_sort(faces_by_id.begin(), faces_by_id.end(),
                     [&](int a, int b) { return o.faces[a].global_face_id < o.faces[b].global_face_id; });
    find_particle_wall_contact_pairs(o, faces_by_id);
    if (o.n_floating > 0)
      find_particle_floating_wall_contact_pairs(o, o.current_time);
    if (!o.solids.empty())
      particle_solid_surfaces_contact_search(o);

    // DEMContactManager::update_contacts (dem_contact_manager.cc:52-144)
    update_fine_search_candidates_pp(o.local_adjacent, o.local_candidates);
    if (o.periodic_enabled)
      update_fine_search_candidates_pp(o.periodic_adjacent, o.periodic_candidates);
    update_fine_search_candidates_pw(o.wall_in_contact, o.wall_candidates);
    update_fine_search_candidates_pw(o.fwall_in_contact, o.fwall_candidates);

    // update_local_particles_in_cells (dem_contact_manager.cc:148-236)
    update_particle_container(o);
    update_contact_container_iterators_pp(o, o.local_adjacent);
    if (o.periodic_enabled)
      update_contact_container_iterators_pp(o, o.periodic_adjacent);
    update_contact_container_iterators_pw(o, o.wall_in_contact);
    update_contact_container_iterators_pw(o, o.fwall_in_contact);

    particle_particle_fine_search(o, o.local_adjacent, o.local_candidates, false);
    if (o.periodic_enabled)
      particle_particle_fine_search(o, o.periodic_adjacent, o.periodic_candidates, true);
    particle_wall_fine_search(o);
    if (o.n_floating > 0)
      particle_floating_wall_fine_search(o, o.current_time);
    if (!o.solids.empty())
      update_and_fine_search_solids(o);
    ++o.contact_build_number;
  }

  // DEMSolver::compute_contact_forces (dem.cc:690-717)
  void compute_contact_forces(Oracle &o)
  {
    o.n_touching_last = 0;
    const double dt = o.cfg.dt;
    for (auto &row : o.local_adjacent.rows)
      pp_execute_contact_calculation(o, row, false, dt);
    if (o.periodic_enabled)
      for (auto &row : o.periodic_adjacent.rows)
        pp_execute_contact_calculation(o, row, true, dt);
    calculate_particle_wall_contact(o, o.wall_in_contact, dt);
    if (o.n_floating > 0)
      calculate_particle_wall_contact(o, o.fwall_in_contact, dt);
    if (!o.solids.empty())
      calculate_particle_solid_object_contact(o, dt);
    // CFDDEMSolver::add_fluid_particle_interaction_force / _torque (cfd_dem_coupling.cc:881-925)
    if (o.ext_enabled)
      for (size_t s = 0; s < o.parts.size(); ++s)
        {
          const uint32_t id = o.parts[s].id;
          if (id < o.ext_force.size())
            {
              o.force[s] = o.force[s] + o.ext_force[id];
              o.torque[s] = o.torque[s] + o.ext_torque[id];
            }
        }
    if (o.cfg.store_forces)
      {
        o.last_force = o.force;
        o.last_torque = o.torque;
      }
  }

  // integrate_temperature (multiphysics_integrator.cc:6-34) with a zero heat source (dem.cc:1148-1152)
  void integrate_temperature(Oracle &o)
  {
    if (!o.thermal_enabled)
      return;
    o.heat_transfer_rate.resize(o.parts.size(), 0.);
    o.last_heat_transfer_rate = o.heat_transfer_rate;
    const double dt = o.cfg.dt;
    for (size_t s = 0; s < o.parts.size(); ++s)
      {
        const uint32_t id = o.parts[s].id;
        double &particle_heat_transfer_rate = o.heat_transfer_rate[s];
        const double particle_heat_source = 0.;
        const double mass_inverse = 1 / o.parts[s].p[P_MASS];
        const double specific_heat_inverse = 1 / o.specific_heat[id];
        o.temperature[id] += dt * (particle_heat_transfer_rate + particle_heat_source) * mass_inverse * specific_heat_inverse;
        particle_heat_transfer_rate = 0;
      }
  }

  void reset_triggers(Oracle &o)
  {
    // dem_action_manager.h:61-75: the iteration after a mobility-status reset searches again, with
    // the statuses identified from the velocities the full step produced
    o.contact_search_trigger = o.mobility_status_reset_trigger;
    o.mobility_status_reset_trigger = false;

    o.clear_tangential_displacement_trigger = false;
    o.solid_object_search_trigger = false;
  }

  void one_step(Oracle &o)
  {
    // SimulationControlTransient::integrate (simulation_control.cc:330-352)
    o.iteration_number++;
    o.current_time += o.cfg.dt;
    execute_contact_detection_and_search(o);
    move_solid_objects(o); // dem.cc:1141-1142
    compute_contact_forces(o);
    integrate_temperature(o); // dem.cc:1144-1153
    if ((o.iteration_number <= 1 && !o.cfg.restart) || o.open_next_step)
      integrate_start(o);
    else
      integrate(o);
    o.open_next_step = false;
    reset_triggers(o);
  }

  int fail(Oracle *o, const char *msg)
  {
    if (o)
      o->error = msg;
    return -1;
  }
  std::string g_create_error;
} // namespace

// ------------------------------------------------------------------ C API ---
extern "C" {

struct lethe_dem_ctx; // the oracle reuses the opaque handle type of lethe_dem.h

int oracle_dem_create(const lethe_dem_config *config, int /*device*/, lethe_dem_ctx **out)
{
  if (!config || !out)
    {
      g_create_error = "null argument";
      return -1;
    }
  if (config->n_types < 1 || config->n_types > LETHE_DEM_MAX_TYPES)
    {
      g_create_error = "n_types out of range";
      return -1;
    }
  if (config->grid_n[0] < 1 || config->grid_n[1] < 1 || config->grid_n[2] < 1)
    {
      g_create_error = "grid_n must be positive";
      return -1;
    }
  if (config->integrator != LETHE_INTEGRATOR_VELOCITY_VERLET && config->integrator != LETHE_INTEGRATOR_EXPLICIT_EULER)
    {
      g_create_error = "unknown integrator";
      return -1;
    }
  if (config->sparse_contacts && config->integrator == LETHE_INTEGRATOR_EXPLICIT_EULER)
    {
      // explicit_euler_integrator.cc:157-159
      g_create_error = "Adaptive sparse contacts are not supported with explicit Euler integrator, use Velocity Verlet integrator.";
      return -1;
    }
  Oracle *o = new Oracle();
  o->cfg = *config;
  o->nx = config->grid_n[0];
  o->ny = config->grid_n[1];
  o->nz = config->grid_n[2];
  o->n_cells = o->nx * o->ny * o->nz;
  o->periodic_enabled = config->periodic[0] || config->periodic[1] || config->periodic[2];
  o->n_floating = 0;
  o->iteration_number = 0;
  o->current_time = 0.;
  o->contact_search_trigger = true;
  o->clear_tangential_displacement_trigger = false;
  o->contact_build_number = 0;
  o->n_touching_last = 0;
  o->sparse_contacts_enabled = config->sparse_contacts != 0;
  o->mobility_status_reset_trigger = o->sparse_contacts_enabled; // set_sparse_contacts_enabled()
  build_cell_order(*o);
  find_cell_neighbors(*o);
  find_cell_periodic_neighbors(*o);
  compute_combined_periodic_offsets(*o);
  set_effective_properties(*o);
  o->cell_parts.assign(o->n_cells, std::vector<int>());
  o->cell_faces.assign(o->n_cells, std::vector<int>());
  *out = reinterpret_cast<lethe_dem_ctx *>(o);
  return 0;
}

void oracle_dem_destroy(lethe_dem_ctx *ctx) { delete reinterpret_cast<Oracle *>(ctx); }
const char *oracle_dem_last_error(const lethe_dem_ctx *ctx) { return reinterpret_cast<const Oracle *>(ctx)->error.c_str(); }
const char *oracle_dem_create_error(void) { return g_create_error.c_str(); }

int oracle_dem_add_particles(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id, const double *x3, const double *props9)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  for (uint64_t i = 0; i < n; ++i)
    {
      Particle p;
      p.id = id[i];
      p.x = mk(x3[3 * i], x3[3 * i + 1], x3[3 * i + 2]);
      std::memcpy(p.p, props9 + 9 * i, 9 * sizeof(double));
      p.cell = cell_of_point(*o, p.x);
      o->parts.push_back(p);
      o->force.push_back(mk(0, 0, 0));
      o->torque.push_back(mk(0, 0, 0));
      o->displacement.push_back(0.);
      o->MOI.push_back(0.1 * p.p[P_MASS] * p.p[P_DP] * p.p[P_DP]);
    }
  // DEMActionManager::particle_insertion_step (dem_action_manager.h:226-230)
  o->contact_search_trigger = true;
  return 0;
}

int oracle_dem_set_particles(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id, const double *x3, const double *props9)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  o->parts.clear();
  o->cell_parts.assign(o->n_cells, std::vector<int>());
  o->force.clear();
  o->torque.clear();
  o->displacement.clear();
  o->MOI.clear();
  o->local_adjacent.clear();
  o->periodic_adjacent.clear();
  o->wall_in_contact.clear();
  o->fwall_in_contact.clear();
  o->local_candidates.clear();
  o->periodic_candidates.clear();
  o->wall_candidates.clear();
  o->fwall_candidates.clear();
  return oracle_dem_add_particles(ctx, n, id, x3, props9);
}

int oracle_dem_n_particles(lethe_dem_ctx *ctx, uint64_t *n)
{
  *n = reinterpret_cast<Oracle *>(ctx)->parts.size();
  return 0;
}

int oracle_dem_get_particles(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *x3, double *props9)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  std::vector<int> order(o->parts.size());
  for (size_t s = 0; s < order.size(); ++s)
    order[s] = int(s);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return o->parts[a].id < o->parts[b].id; });
  uint64_t n = std::min<uint64_t>(n_max, order.size());
  for (uint64_t i = 0; i < n; ++i)
    {
      const Particle &p = o->parts[order[i]];
      id[i] = p.id;
      for (int d = 0; d < 3; ++d)
        x3[3 * i + d] = p.x[d];
      std::memcpy(props9 + 9 * i, p.p, 9 * sizeof(double));
    }
  *n_out = n;
  return 0;
}

int oracle_dem_set_walls(lethe_dem_ctx *ctx, uint64_t n_faces, const lethe_wall_face *faces)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  o->faces.assign(faces, faces + n_faces);
  o->contact_search_trigger = true;
  return 0;
}

int oracle_dem_set_floating_walls(lethe_dem_ctx *ctx, int32_t n, const double *point3, const double *normal3,
                                  const double *t_start, const double *t_end)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  if (n < 0 || n > LETHE_DEM_MAX_FLOATING_WALLS)
    return fail(o, "too many floating walls");
  o->n_floating = n;
  for (int w = 0; w < n; ++w)
    {
      o->fw_point[w] = mk(point3[3 * w], point3[3 * w + 1], point3[3 * w + 2]);
      o->fw_normal[w] = mk(normal3[3 * w], normal3[3 * w + 1], normal3[3 * w + 2]);
      o->fw_t0[w] = t_start[w];
      o->fw_t1[w] = t_end[w];
    }
  // GridTools::maximal_cell_diameter of the uniform grid (dem.cc:1106-1112 -> build)
  const double *h = o->cfg.cell_size;
  const double maximum_cell_diameter = std::sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]);
  find_boundary_cells_for_floating_walls(*o, maximum_cell_diameter);
  o->contact_search_trigger = true;
  return 0;
}

int oracle_dem_set_boundary_motion(lethe_dem_ctx *ctx, uint32_t boundary_id, const double tv[3], double speed,
                                   const double axis[3], const double point[3])
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  BoundaryMotion m;
  m.boundary_id = boundary_id;
  m.translational_velocity = mk(tv[0], tv[1], tv[2]);
  m.rotational_speed = speed;
  m.rotational_vector = mk(axis[0], axis[1], axis[2]);
  m.point_on_axis = mk(point[0], point[1], point[2]);
  for (auto &e : o->motions)
    if (e.boundary_id == boundary_id)
      {
        e = m;
        return 0;
      }
  o->motions.push_back(m);
  return 0;
}

int oracle_dem_add_solid_surface(lethe_dem_ctx *ctx, uint32_t n_vertices, const double *vertices3, uint32_t n_triangles,
                                 const uint32_t *triangles3, const double tv[3], const double av[3], const double center[3],
                                 int32_t *solid_index)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  Solid sd;
  for (uint32_t v = 0; v < n_vertices; ++v)
    sd.vertices.push_back(mk(vertices3[3 * v], vertices3[3 * v + 1], vertices3[3 * v + 2]));
  for (uint32_t t = 0; t < n_triangles; ++t)
    {
      for (int k = 0; k < 3; ++k)
        if (triangles3[3 * t + k] >= n_vertices)
          return fail(o, "triangle refers to a vertex outside the solid");
      sd.tri.push_back({triangles3[3 * t], triangles3[3 * t + 1], triangles3[3 * t + 2]});
    }
  sd.translational_velocity = mk(tv[0], tv[1], tv[2]);
  sd.angular_velocity = mk(av[0], av[1], av[2]);
  sd.center_of_rotation = mk(center[0], center[1], center[2]);
  sd.displacement_since_mapped.assign(n_vertices, mk(0, 0, 0));
  setup_solid_neighbors(sd);
  o->solids.push_back(sd);
  // DEMActionManager::set_solid_objects_enabled: the first search maps the solids
  o->solid_object_search_trigger = true;
  o->contact_search_trigger = true;
  if (solid_index)
    *solid_index = int32_t(o->solids.size()) - 1;
  return 0;
}

int oracle_dem_set_solid_motion(lethe_dem_ctx *ctx, int32_t solid_index, const double tv[3], const double av[3])
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  if (solid_index < 0 || size_t(solid_index) >= o->solids.size())
    return fail(o, "no such solid");
  o->solids[solid_index].translational_velocity = mk(tv[0], tv[1], tv[2]);
  o->solids[solid_index].angular_velocity = mk(av[0], av[1], av[2]);
  return 0;
}

int oracle_dem_get_solid_vertices(lethe_dem_ctx *ctx, int32_t solid_index, uint32_t n_max, double *vertices3)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  if (solid_index < 0 || size_t(solid_index) >= o->solids.size())
    return fail(o, "no such solid");
  const Solid &sd = o->solids[solid_index];
  for (uint32_t v = 0; v < std::min<size_t>(n_max, sd.vertices.size()); ++v)
    for (int d = 0; d < 3; ++d)
      vertices3[3 * v + d] = sd.vertices[v][d];
  return 0;
}

int oracle_dem_get_solid_contacts(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *particle_id, uint32_t *solid,
                                  uint32_t *triangle, double *tangential3)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  struct Row
  {
    uint32_t pid, solid, tri;
    V3 h;
  };
  std::vector<Row> rows;
  for (size_t sc = 0; sc < o->solids.size(); ++sc)
    for (auto &tp : o->solids[sc].in_contact)
      for (auto &pi : tp.second)
        rows.push_back(Row{pi.first, uint32_t(sc), tp.first, pi.second.tangential_displacement});
  std::sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) {
    return std::tie(a.pid, a.solid, a.tri) < std::tie(b.pid, b.solid, b.tri);
  });
  const uint64_t n = std::min<uint64_t>(n_max, rows.size());
  for (uint64_t k = 0; k < n; ++k)
    {
      particle_id[k] = rows[k].pid;
      solid[k] = rows[k].solid;
      triangle[k] = rows[k].tri;
      for (int d = 0; d < 3; ++d)
        tangential3[3 * k + d] = rows[k].h[d];
    }
  *n_out = rows.size();
  return 0;
}

// Test hook: the cell-neighbour lists in active-cell numbering. which = 0: find_cell_neighbors
// <dim, false> (each pair of neighbouring cells listed once, find_cell_neighbors.cc:10-104);
// which = 1: find_full_cell_neighbors (find_cell_neighbors.cc:293-333, the solid-surface search).
// out = rows of `stride` ints: [cell, neighbours..., -1 padding].
int oracle_dem_cell_neighbors(lethe_dem_ctx *ctx, int which, int stride, int32_t *out)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  for (int r = 0; r < o->n_cells; ++r)
    {
      std::vector<int> row;
      if (which == 0)
        {
          for (int c : o->cells_local_neighbor_list[r])
            row.push_back(o->rank_of_cell[c]);
        }
      else
        {
          // the cell, then the cells of its 8 vertices in vertex order, each vertex's cells in
          // active-cell order, first occurrence only
          const int cell = o->cell_of_rank[r];
          const int ci = cell % o->nx, cj = (cell / o->nx) % o->ny, ck = cell / (o->nx * o->ny);
          row.push_back(r);
          std::vector<int> vcells;
          for (int vertex = 0; vertex < 8; ++vertex)
            {
              cells_at_vertex(*o, ci + (vertex & 1), cj + ((vertex >> 1) & 1), ck + ((vertex >> 2) & 1), vcells);
              for (int nb : vcells)
                if (std::find(row.begin(), row.end(), o->rank_of_cell[nb]) == row.end())
                  row.push_back(o->rank_of_cell[nb]);
            }
        }
      if (int(row.size()) > stride)
        return fail(o, "stride too small");
      for (int k = 0; k < stride; ++k)
        out[size_t(r) * stride + k] = k < int(row.size()) ? row[k] : -1;
    }
  return 0;
}

int oracle_dem_step(lethe_dem_ctx *ctx, uint64_t n_steps)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  for (uint64_t s = 0; s < n_steps; ++s)
    one_step(*o);
  return 0;
}

// DEMSolver::synchronize_velocities (dem.cc:719-745)
int oracle_dem_synchronize_velocities(lethe_dem_ctx *ctx)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  execute_contact_detection_and_search(*o);
  compute_contact_forces(*o);
  integrate_end(*o);
  reset_triggers(*o);
  return 0;
}

int oracle_dem_force_contact_search(lethe_dem_ctx *ctx, int clear_tangential_displacement)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  o->contact_search_trigger = true;
  if (clear_tangential_displacement)
    o->clear_tangential_displacement_trigger = true;
  return 0;
}

int oracle_dem_step_host(lethe_dem_ctx *ctx, uint64_t n_steps, uint64_t n, const uint32_t *id, double *x3, double *props9)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  for (uint64_t i = 0; i < n; ++i)
    {
      const int s = slot(*o, id[i]);
      if (s < 0)
        continue;
      o->parts[s].x = mk(x3[3 * i], x3[3 * i + 1], x3[3 * i + 2]);
      std::memcpy(o->parts[s].p, props9 + 9 * i, 9 * sizeof(double));
    }
  oracle_dem_step(ctx, n_steps);
  for (uint64_t i = 0; i < n; ++i)
    {
      const int s = slot(*o, id[i]);
      if (s < 0)
        continue;
      for (int d = 0; d < 3; ++d)
        x3[3 * i + d] = o->parts[s].x[d];
      std::memcpy(props9 + 9 * i, o->parts[s].p, 9 * sizeof(double));
    }
  return 0;
}

// set_multiphysic_properties (particle_particle_contact_force.h:1755-1826)
int oracle_dem_enable_heat_transfer(lethe_dem_ctx *ctx, const lethe_dem_thermal_properties *pr)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  const int n = o->n_types;
  o->th_real_E.assign(n * n, 0.);
  o->th_roughness.assign(n * n, 0.);
  o->th_slope.assign(n * n, 0.);
  o->th_microhardness.assign(n * n, 0.);
  o->th_gas_m.assign(n * n, 0.);
  o->th_conductivity.assign(n, 0.);
  o->th_conductivity_gas = pr->thermal_conductivity_gas;
  for (int i = 0; i < n; ++i)
    {
      const double real_youngs_modulus_i = pr->real_youngs_modulus[i], poisson_ratio_i = o->cfg.poisson[i];
      const double surface_roughness_i = pr->surface_roughness[i], surface_slope_i = pr->surface_slope[i];
      const double microhardness_i = pr->microhardness[i], thermal_accommodation_i = pr->thermal_accommodation[i];
      o->th_conductivity[i] = pr->thermal_conductivity[i];
      for (int j = 0; j < n; ++j)
        {
          const int k = i * n + j;
          const double real_youngs_modulus_j = pr->real_youngs_modulus[j], poisson_ratio_j = o->cfg.poisson[j];
          const double surface_roughness_j = pr->surface_roughness[j], surface_slope_j = pr->surface_slope[j];
          const double microhardness_j = pr->microhardness[j], thermal_accommodation_j = pr->thermal_accommodation[j];
          o->th_real_E[k] = (real_youngs_modulus_i * real_youngs_modulus_j) /
                            ((real_youngs_modulus_j * (1.0 - poisson_ratio_i * poisson_ratio_i)) +
                             (real_youngs_modulus_i * (1.0 - poisson_ratio_j * poisson_ratio_j)) + DBL_MIN);
          o->th_roughness[k] = std::sqrt(surface_roughness_i * surface_roughness_i + surface_roughness_j * surface_roughness_j);
          o->th_slope[k] = std::sqrt(surface_slope_i * surface_slope_i + surface_slope_j * surface_slope_j);
          o->th_microhardness[k] = harmonic_mean(microhardness_i, microhardness_j);
          o->th_gas_m[k] = ((2. - thermal_accommodation_i) / thermal_accommodation_i + (2. - thermal_accommodation_j) / thermal_accommodation_j) *
                           (2. * pr->specific_heats_ratio_gas) / (1. + pr->specific_heats_ratio_gas) * pr->molecular_mean_free_path_gas /
                           (pr->dynamic_viscosity_gas * pr->specific_heat_gas / pr->thermal_conductivity_gas);
        }
    }
  o->thermal_enabled = true;
  return 0;
}

int oracle_dem_set_temperatures(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id, const double *temperature, const double *specific_heat)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  for (uint64_t k = 0; k < n; ++k)
    {
      if (id[k] >= o->temperature.size())
        {
          o->temperature.resize(size_t(id[k]) + 1, 0.);
          o->specific_heat.resize(size_t(id[k]) + 1, 1.);
        }
      o->temperature[id[k]] = temperature[k];
      o->specific_heat[id[k]] = specific_heat[k];
    }
  return 0;
}

int oracle_dem_get_temperatures(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *temperature,
                                double *heat_transfer_rate)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  std::vector<std::pair<uint32_t, size_t>> order;
  for (size_t s = 0; s < o->parts.size(); ++s)
    order.emplace_back(o->parts[s].id, s);
  std::sort(order.begin(), order.end());
  const uint64_t m = std::min<uint64_t>(n_max, order.size());
  for (uint64_t k = 0; k < m; ++k)
    {
      id[k] = order[k].first;
      temperature[k] = order[k].first < o->temperature.size() ? o->temperature[order[k].first] : 0.;
      heat_transfer_rate[k] = order[k].second < o->last_heat_transfer_rate.size() ? o->last_heat_transfer_rate[order[k].second] : 0.;
    }
  *n_out = m;
  return 0;
}

// the intermediate values tests/dem/particle_particle_thermal_resistances.cc prints (test hook)
int oracle_dem_thermal_resistances(const double *in13, double *out8)
{
  calculate_contact_thermal_conductance(in13[0], in13[1], in13[2], in13[3], in13[4], in13[5], in13[6], in13[7], in13[8], in13[9], in13[10],
                                        in13[11], in13[12], out8);
  return 0;
}

// CFD-DEM rows of 23 properties (dem_properties.h:92-142); the loads as add_fluid_particle_interaction_force /
// _torque add them (cfd_dem_coupling.cc:881-925)
int oracle_dem_set_particles(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id, const double *x3, const double *props9);
int oracle_dem_set_external_loads(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id, const double *force3, const double *torque3);
int oracle_dem_get_particles(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *x3, double *props9);
static void split_cfd_rows(uint64_t n, const double *props23, std::vector<double> *props9, std::vector<double> &force3,
                           std::vector<double> &torque3)
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
          force3[3 * k + d] = (p[9 + d] + p[12 + d]) + p[15 + d];
          torque3[3 * k + d] = p[18 + d];
        }
    }
}
int oracle_dem_set_particles_cfd(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id, const double *x3, const double *props23)
{
  std::vector<double> p9, f3, t3;
  split_cfd_rows(n, props23, &p9, f3, t3);
  const int rc = oracle_dem_set_particles(ctx, n, id, x3, p9.data());
  return rc ? rc : oracle_dem_set_external_loads(ctx, n, id, f3.data(), t3.data());
}
int oracle_dem_update_loads_cfd(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id, const double *props23)
{
  std::vector<double> f3, t3;
  split_cfd_rows(n, props23, nullptr, f3, t3);
  return oracle_dem_set_external_loads(ctx, n, id, f3.data(), t3.data());
}
int oracle_dem_get_particles_cfd(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *x3, double *props23)
{
  std::vector<double> p9(9 * n_max);
  const int rc = oracle_dem_get_particles(ctx, n_max, n_out, id, x3, p9.data());
  if (rc)
    return rc;
  for (uint64_t k = 0; k < *n_out; ++k)
    std::memcpy(props23 + 23 * k, p9.data() + 9 * k, 72);
  return 0;
}

// simulation_control->read(prefix) + DEMActionManager::restart_simulation (read_checkpoint.cc:47,
// dem_action_manager.h:185-200)
int oracle_dem_set_time(lethe_dem_ctx *ctx, uint64_t iteration_number, double current_time)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  o->iteration_number = iteration_number;
  o->current_time = current_time;
  o->contact_search_trigger = true;
  o->clear_tangential_displacement_trigger = true;
  return 0;
}

int oracle_dem_restart_integration(lethe_dem_ctx *ctx)
{
  reinterpret_cast<Oracle *>(ctx)->open_next_step = true;
  return 0;
}

int oracle_dem_set_external_loads(lethe_dem_ctx *ctx, uint64_t n, const uint32_t *id, const double *force3, const double *torque3)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  if (n == 0)
    {
      o->ext_enabled = false;
      o->ext_force.clear();
      o->ext_torque.clear();
      return 0;
    }
  for (uint64_t k = 0; k < n; ++k)
    {
      if (o->ext_force.size() <= id[k])
        {
          o->ext_force.resize(size_t(id[k]) + 1, mk(0, 0, 0));
          o->ext_torque.resize(size_t(id[k]) + 1, mk(0, 0, 0));
        }
      o->ext_force[id[k]] = mk(force3[3 * k], force3[3 * k + 1], force3[3 * k + 2]);
      o->ext_torque[id[k]] = torque3 ? mk(torque3[3 * k], torque3[3 * k + 1], torque3[3 * k + 2]) : mk(0, 0, 0);
    }
  o->ext_enabled = true;
  return 0;
}

int oracle_dem_step_host_state(lethe_dem_ctx *ctx, uint64_t n_steps, uint64_t n, const uint32_t *id, double *state9)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  if (id)
    o->host_row_ids.assign(id, id + n);
  else if (o->host_row_ids.size() != n)
    return fail(o, "step_host_state: id == NULL reuses the id table of the previous call, which had a different row count");
  for (uint64_t i = 0; i < n; ++i)
    {
      const int s = slot(*o, o->host_row_ids[i]);
      if (s < 0)
        continue;
      o->parts[s].x = mk(state9[9 * i], state9[9 * i + 1], state9[9 * i + 2]);
      std::memcpy(o->parts[s].p + P_VX, state9 + 9 * i + 3, 6 * sizeof(double));
    }
  oracle_dem_step(ctx, n_steps);
  for (uint64_t i = 0; i < n; ++i)
    {
      const int s = slot(*o, o->host_row_ids[i]);
      if (s < 0)
        continue;
      for (int d = 0; d < 3; ++d)
        state9[9 * i + d] = o->parts[s].x[d];
      std::memcpy(state9 + 9 * i + 3, o->parts[s].p + P_VX, 6 * sizeof(double));
    }
  return 0;
}

// Unordered pairs (i<j by id), tangential displacement oriented i -> j.
int oracle_dem_get_pairs(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *i_id, uint32_t *j_id,
                         double *tangential3)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  struct E
  {
    uint32_t i, j;
    V3 t;
  };
  std::vector<E> all;
  for (DenseRows<PPRow> *adj : {&o->local_adjacent, &o->periodic_adjacent})
    for (auto &row : adj->rows)
      for (auto &e : row.second)
        {
          if (row.one < e.two)
            all.push_back(E{row.one, e.two, e.tangential_displacement});
          else
            all.push_back(E{e.two, row.one, -e.tangential_displacement});
        }
  std::sort(all.begin(), all.end(), [](const E &a, const E &b) { return a.i != b.i ? a.i < b.i : a.j < b.j; });
  *n_out = all.size();
  const uint64_t n = std::min<uint64_t>(n_max, all.size());
  for (uint64_t k = 0; k < n; ++k)
    {
      i_id[k] = all[k].i;
      j_id[k] = all[k].j;
      if (tangential3)
        for (int d = 0; d < 3; ++d)
          tangential3[3 * k + d] = all[k].t[d];
    }
  return 0;
}

int oracle_dem_get_wall_contacts(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *particle_id,
                                 uint32_t *face_id, double *tangential3)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  struct E
  {
    uint32_t p, f;
    V3 t;
  };
  std::vector<E> all;
  for (auto &row : o->wall_in_contact.rows)
    for (auto &e : row.second)
      all.push_back(E{row.one, e.face, e.tangential_displacement});
  for (auto &row : o->fwall_in_contact.rows)
    for (auto &e : row.second)
      all.push_back(E{row.one, e.face | 0x80000000u, e.tangential_displacement});
  std::sort(all.begin(), all.end(), [](const E &a, const E &b) { return a.p != b.p ? a.p < b.p : a.f < b.f; });
  *n_out = all.size();
  const uint64_t n = std::min<uint64_t>(n_max, all.size());
  for (uint64_t k = 0; k < n; ++k)
    {
      particle_id[k] = all[k].p;
      face_id[k] = all[k].f;
      if (tangential3)
        for (int d = 0; d < 3; ++d)
          tangential3[3 * k + d] = all[k].t[d];
    }
  return 0;
}

int oracle_dem_get_forces(lethe_dem_ctx *ctx, uint64_t n_max, uint64_t *n_out, uint32_t *id, double *force3, double *torque3)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  if (!o->cfg.store_forces)
    return fail(o, "store_forces not enabled");
  std::vector<int> order(o->parts.size());
  for (size_t s = 0; s < order.size(); ++s)
    order[s] = int(s);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return o->parts[a].id < o->parts[b].id; });
  const uint64_t n = std::min<uint64_t>(n_max, order.size());
  for (uint64_t i = 0; i < n; ++i)
    {
      const int s = order[i];
      id[i] = o->parts[s].id;
      for (int d = 0; d < 3; ++d)
        {
          force3[3 * i + d] = size_t(s) < o->last_force.size() ? o->last_force[s][d] : 0.;
          torque3[3 * i + d] = size_t(s) < o->last_torque.size() ? o->last_torque[s][d] : 0.;
        }
    }
  *n_out = n;
  return 0;
}

int oracle_dem_get_stats(lethe_dem_ctx *ctx, lethe_dem_stats *st)
{
  st->n_migrated = 0; // single domain
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  std::memset(st, 0, sizeof(*st));
  st->n_particles = o->parts.size();
  st->n_rebuilds = o->contact_build_number;
  st->n_steps = o->iteration_number;
  for (DenseRows<PPRow> *adj : {&o->local_adjacent, &o->periodic_adjacent})
    for (auto &row : adj->rows)
      st->n_pair_entries += row.second.size();
  for (DenseRows<PWRow> *adj : {&o->wall_in_contact, &o->fwall_in_contact})
    for (auto &row : adj->rows)
      st->n_wall_entries += row.second.size();
  st->n_pairs_touching = o->n_touching_last;
  st->v_min = st->omega_min = st->ke_trans_min = st->ke_rot_min = DBL_MAX;
  for (size_t s = 0; s < o->parts.size(); ++s)
    {
      const double *p = o->parts[s].p;
      const double v2 = p[P_VX] * p[P_VX] + p[P_VY] * p[P_VY] + p[P_VZ] * p[P_VZ];
      const double w2 = p[P_WX] * p[P_WX] + p[P_WY] * p[P_WY] + p[P_WZ] * p[P_WZ];
      const double v = std::sqrt(v2), w = std::sqrt(w2);
      const double ket = 0.5 * p[P_MASS] * v2;
      const double ker = 0.5 * (0.1 * p[P_MASS] * p[P_DP] * p[P_DP]) * w2;
      st->v_min = std::min(st->v_min, v);
      st->v_max = std::max(st->v_max, v);
      st->v_sum += v;
      st->omega_min = std::min(st->omega_min, w);
      st->omega_max = std::max(st->omega_max, w);
      st->omega_sum += w;
      st->ke_trans_min = std::min(st->ke_trans_min, ket);
      st->ke_trans_max = std::max(st->ke_trans_max, ket);
      st->ke_trans_sum += ket;
      st->ke_rot_min = std::min(st->ke_rot_min, ker);
      st->ke_rot_max = std::max(st->ke_rot_max, ker);
      st->ke_rot_sum += ker;
    }
  if (o->parts.empty())
    st->v_min = st->omega_min = st->ke_trans_min = st->ke_rot_min = 0;
  return 0;
}

int oracle_dem_get_mobility_status(lethe_dem_ctx *ctx, uint64_t n_cells, int32_t *status)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  if (n_cells != uint64_t(o->n_cells))
    return fail(o, "get_mobility_status: n_cells does not match the grid");
  for (int c = 0; c < o->n_cells; ++c)
    status[c] = (o->sparse_contacts_enabled && !o->cell_mobility_status.empty()) ? o->cell_mobility_status[c] : LETHE_MOBILITY_MOBILE;
  return 0;
}

// a single domain has nothing to balance: the calls exist so that drivers run unchanged
int oracle_dem_set_load_balancing(lethe_dem_ctx *, int, double, int) { return 0; }
int oracle_dem_set_load_balancing_weights(lethe_dem_ctx *, double, double, double, double) { return 0; }
int oracle_dem_get_slab(lethe_dem_ctx *ctx, int32_t *lo, int32_t *hi, uint64_t *n_repartitions)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  *lo = 0;
  *hi = o->nx;
  *n_repartitions = 0;
  return 0;
}
int oracle_dem_enable_timers(lethe_dem_ctx *, int) { return 0; }
int oracle_dem_event_record(lethe_dem_ctx *, int) { return 0; }
int oracle_dem_event_elapsed(lethe_dem_ctx *, double *ms) { *ms = 0; return 0; }
int oracle_dem_nccl_unique_id(uint8_t *) { return -1; }
int oracle_dem_comm_init(lethe_dem_ctx *, int, int, const uint8_t *) { return -1; }
int oracle_dem_kernel_launches(lethe_dem_ctx *, uint64_t *n) { *n = 0; return 0; }
int oracle_dem_get_timers(lethe_dem_ctx *, int, double *a, uint64_t *na, double *b, uint64_t *nb)
{
  *a = *b = 0;
  *na = *nb = 0;
  return 0;
}

int oracle_dem_set_option(lethe_dem_ctx *ctx, const char *name, int value)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  if (std::string(name) == "dmt_stale_scratch")
    {
      o->dmt_stale_scratch = value != 0;
      return 0;
    }
  return fail(o, "unknown option");
}

// ---- single-contact taps used by the golden-vector tests (no grid needed) ----
// One particle-particle contact evaluation exactly as execute_contact_calculation
// does for one pair: returns force on particle one/two and torques; updates history.
int oracle_dem_pair_force(lethe_dem_ctx *ctx, const double *x1, const double *p1, const double *x2, const double *p2,
                          double *tangential3, double *rolling3, double *force_one3, double *torque_one3,
                          double *force_two3, double *torque_two3, double *overlap)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  PPInfo info;
  info.two = 1;
  info.tangential_displacement = mk(tangential3[0], tangential3[1], tangential3[2]);
  info.rolling_resistance_spring_torque = mk(rolling3[0], rolling3[1], rolling3[2]);
  info.periodic_offset = mk(0, 0, 0);
  const V3 a = mk(x1[0], x1[1], x1[2]), b = mk(x2[0], x2[1], x2[2]);
  const double normal_overlap = 0.5 * (p1[P_DP] + p2[P_DP]) - std::sqrt(distance_square(a, b));
  *overlap = normal_overlap;
  V3 f1 = mk(0, 0, 0), f2 = mk(0, 0, 0), t1 = mk(0, 0, 0), t2 = mk(0, 0, 0);
  if (normal_overlap > o->pp_force_threshold)
    {
      PPOut out;
      out.normal_force = out.tangential_force = out.t1 = out.t2 = out.rolling = mk(0, 0, 0);
      V3 n, vt;
      double vn;
      pp_update_contact_information(info, vt, vn, n, p1, p2, a, b, o->cfg.dt);
      pp_calculate_contact(*o, o->cfg.pp_model, info, vt, vn, n, normal_overlap, o->cfg.dt, p1, p2, out);
      const V3 total = out.normal_force + out.tangential_force;
      f1 = f1 - total;
      f2 = f2 + total;
      t1 = t1 + (-out.t1 + out.rolling);
      t2 = t2 + (-out.t2 - out.rolling);
    }
  else
    {
      info.tangential_displacement = mk(0, 0, 0);
      info.rolling_resistance_spring_torque = mk(0, 0, 0);
    }
  for (int d = 0; d < 3; ++d)
    {
      tangential3[d] = info.tangential_displacement[d];
      rolling3[d] = info.rolling_resistance_spring_torque[d];
      force_one3[d] = f1[d];
      force_two3[d] = f2[d];
      torque_one3[d] = t1[d];
      torque_two3[d] = t2[d];
    }
  return 0;
}

// Integrator taps for the free-flight golden (tests/dem/integration_velocity_verlet.cc):
// apply a constant external force / torque and an explicit MOI to every particle.
int oracle_dem_integrate_external(lethe_dem_ctx *ctx, int phase, const double *force3, const double *torque3, double moi)
{
  Oracle *o = reinterpret_cast<Oracle *>(ctx);
  for (size_t s = 0; s < o->parts.size(); ++s)
    {
      o->force[s] = mk(force3[0], force3[1], force3[2]);
      o->torque[s] = mk(torque3[0], torque3[1], torque3[2]);
      if (moi > 0)
        o->MOI[s] = moi;
    }
  if (phase == 0)
    integrate_start(*o);
  else if (phase == 1)
    integrate(*o);
  else
    integrate_end(*o);
  return 0;
}

} // extern "C"
