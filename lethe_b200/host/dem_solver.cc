// dem_solver.cc — see dem_solver.h.
#include "dem_solver.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <array>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

namespace lethe_b200
{
  std::vector<lethe_wall_face> box_wall_faces(const Mesh &mesh, const std::vector<unsigned> &outlets, const std::array<int, 3> &periodic)
  {
    static const double normals[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    const Vec3 h = mesh.cell_size();
    std::vector<lethe_wall_face> faces;
    for (int k = 0; k < mesh.n[2]; ++k)
      for (int j = 0; j < mesh.n[1]; ++j)
        for (int i = 0; i < mesh.n[0]; ++i)
          {
            const int idx[3] = {i, j, k};
            const int cell = i + mesh.n[0] * (j + mesh.n[1] * k);
            for (int axis = 0; axis < 3; ++axis)
              for (int side = 0; side < 2; ++side)
                {
                  if (idx[axis] != (side == 0 ? 0 : mesh.n[axis] - 1))
                    continue;
                  const int face_no = 2 * axis + side;
                  // colorized hyper_cube / hyper_rectangle: boundary id = 2*axis + side
                  const unsigned bid = mesh.colorize ? unsigned(face_no) : 0u;
                  if (periodic[axis] || std::find(outlets.begin(), outlets.end(), bid) != outlets.end())
                    continue;
                  lethe_wall_face f{};
                  f.cell = cell;
                  f.boundary_id = bid;
                  f.global_face_id = uint32_t(cell * 6 + face_no);
                  for (int d = 0; d < 3; ++d)
                    {
                      f.normal[d] = normals[face_no][d];
                      f.point[d] = mesh.lo[d] + (idx[d] + 0.5) * h[d];
                    }
                  f.point[axis] = side == 0 ? mesh.lo[axis] : mesh.hi[axis];
                  faces.push_back(f);
                }
          }
    return faces;
  }

  SizeDistribution::SizeDistribution(const ParticleType &t, unsigned rank)
    : type(t)
    , gen(unsigned(t.prn_seed) + rank)
    , normal(t.average_diameter, t.standard_deviation > 0 ? t.standard_deviation : 1.0)
  {
    if (t.size_distribution_type == "lognormal")
      {
        const double r = t.standard_deviation / t.average_diameter;
        const double sigma_ln = std::sqrt(std::log(1. + r * r));
        lognormal = std::lognormal_distribution<>(std::log(t.average_diameter) - 0.5 * sigma_ln * sigma_ln, sigma_ln);
      }
    else if (t.size_distribution_type != "uniform" && t.size_distribution_type != "normal")
      throw std::runtime_error("size distribution type `" + t.size_distribution_type + "` is not mirrored (uniform, normal, lognormal are)");
  }

  std::vector<double> SizeDistribution::sample(long n)
  {
    std::vector<double> out;
    out.reserve(size_t(std::max(0l, n)));
    if (type.size_distribution_type == "uniform")
      {
        out.assign(size_t(std::max(0l, n)), type.average_diameter);
        return out;
      }
    const double lo = type.min_diameter(), hi = type.max_diameter();
    while (long(out.size()) < n)
      {
        const double d = type.size_distribution_type == "normal" ? normal(gen) : lognormal(gen);
        if (d > lo && d < hi)
          out.push_back(d);
      }
    return out;
  }

  ParticleRows volume_insertion(const DEMParameters &p, long n_insert, uint32_t first_id, int particle_type, SizeDistribution &sizes)
  {
    const InsertionInfo &ins = p.insertion;
    const ParticleType &t = p.particle_types.at(particle_type);
    const double d_max = p.maximum_particle_diameter();
    long n_dir[3] = {0, 0, 0};
    for (int axis : ins.direction_sequence)
      n_dir[axis] = long((ins.box_point_2[axis] - ins.box_point_1[axis]) / (ins.distance_threshold * d_max));
    const long n_lattice = n_dir[0] * n_dir[1] * n_dir[2];
    const int a0 = ins.direction_sequence[0], a1 = ins.direction_sequence[1], a2 = ins.direction_sequence[2];
    auto location = [&](long site, double r1, double r2, double *out) {
      const long i0 = site % n_dir[a0], i1 = (site % (n_dir[a0] * n_dir[a1])) / n_dir[a0], i2 = site / (n_dir[a0] * n_dir[a1]);
      out[a0] = ins.box_point_1[a0] + ((i0 + 0.5) * ins.distance_threshold - r1) * d_max;
      out[a1] = ins.box_point_1[a1] + ((i1 + 0.5) * ins.distance_threshold - r2) * d_max;
      out[a2] = ins.box_point_1[a2] + ((i2 + 0.5) * ins.distance_threshold - r1) * d_max;
    };
    // set_filtered_index (insertion_volume.cc:207-351): the sites whose un-jittered location the
    // acceptance function accepts; the random offsets are drawn for the accepted sites only
    std::vector<long> sites;
    for (long k = 0; k < n_lattice; ++k)
      {
        if (!ins.acceptance_function.empty())
          {
            double x[3];
            location(k, 0., 0., x);
            FunctionExpression::Variables v;
            v.x = x[0], v.y = x[1], v.z = x[2];
            if (!(ins.acceptance_function(v) > 0.))
              continue;
          }
        sites.push_back(k);
      }
    const long n_sites = long(sites.size());
    n_insert = std::min(n_insert, n_sites);
    // one srand(seed * (i + 1)) + rand() per site
    std::vector<double> rnd(n_sites);
    for (long i = 0; i < n_sites; ++i)
      {
        srand(unsigned(ins.prn_seed * (i + 1)));
        rnd[i] = (double(rand()) / double(RAND_MAX)) * ins.maximum_offset;
      }
    ParticleRows rows;
    rows.id.resize(n_insert);
    rows.x.resize(3 * n_insert);
    rows.props.assign(size_t(LETHE_DEM_N_PROPERTIES) * n_insert, 0.0);
    const std::vector<double> diameters = sizes.sample(n_insert); // insertion.cc:78-90
    for (long k = 0; k < n_insert; ++k)
      {
        const double d = std::fabs(diameters[k]), half = d * 0.5;
        location(sites[k], rnd[k], rnd[n_sites - k - 1], &rows.x[3 * k]);
        double *pr = &rows.props[size_t(LETHE_DEM_N_PROPERTIES) * k];
        pr[0] = particle_type;
        pr[1] = d;
        pr[2] = t.density * 4.0 / 3.0 * M_PI * (half * half * half);
        for (int c = 0; c < 3; ++c)
          {
            pr[3 + c] = ins.initial_velocity[c];
            pr[6 + c] = ins.initial_omega[c];
          }
        rows.id[k] = first_id + uint32_t(k);
      }
    return rows;
  }

  std::vector<std::array<int, 3>> active_cell_order(const Mesh &mesh)
  {
    std::vector<std::array<int, 3>> cells;
    for (int k = 0; k < mesh.n[2]; ++k)
      for (int j = 0; j < mesh.n[1]; ++j)
        for (int i = 0; i < mesh.n[0]; ++i)
          cells.push_back({{i, j, k}});
    if (mesh.cell_order != LETHE_CELL_ORDER_MORTON)
      return cells;
    auto key = [](const std::array<int, 3> &c) {
      uint64_t out = 0;
      for (int b = 0; b < 21; ++b)
        out |= (uint64_t((c[0] >> b) & 1) << (3 * b)) | (uint64_t((c[1] >> b) & 1) << (3 * b + 1)) | (uint64_t((c[2] >> b) & 1) << (3 * b + 2));
      return out;
    };
    std::sort(cells.begin(), cells.end(), [&](const std::array<int, 3> &a, const std::array<int, 3> &b) { return key(a) < key(b); });
    return cells;
  }

  PlaneInsertion::PlaneInsertion(const DEMParameters &p)
  {
    const Mesh &mesh = p.mesh;
    const Vec3 h = mesh.cell_size();
    const Vec3 &point = p.insertion.plane_point, &normal = p.insertion.plane_normal;
    // find_inplane_cells (insertion_plane.cc:43-88): a sign change of the vertex distances
    for (const auto &c : active_cell_order(mesh))
      {
        double reference = 0;
        for (int v = 0; v < 8; ++v) // deal.II vertex order: x fastest
          {
            double distance = 0;
            for (int d = 0; d < 3; ++d)
              distance += ((mesh.lo[d] + (c[d] + ((v >> d) & 1)) * h[d]) - point[d]) * normal[d];
            if (v == 0)
              reference = distance;
            else if (reference * distance <= 0)
              {
                cells.push_back(c);
                break;
              }
          }
      }
    maximum_range_for_randomness = p.insertion.maximum_offset / double(RAND_MAX);
    srand(1); // the method never seeds: the stream of a process that has not called srand
  }

  ParticleRows PlaneInsertion::insert(const DEMParameters &p, const std::vector<char> &occupied, long remaining, uint32_t first_id,
                                      int particle_type, SizeDistribution &sizes)
  {
    const Mesh &mesh = p.mesh;
    const Vec3 h = mesh.cell_size();
    std::vector<std::array<int, 3>> empty;
    for (const auto &c : cells)
      if (!occupied.at(size_t(c[0] + mesh.n[0] * (c[1] + mesh.n[1] * c[2]))))
        empty.push_back(c);
    const long n_insert = std::min<long>(long(empty.size()), remaining);
    empty.erase(empty.begin(), empty.begin() + (long(empty.size()) - n_insert)); // surplus cells go from the front (:216-222)
    const ParticleType &t = p.particle_types.at(particle_type);
    ParticleRows rows;
    rows.id.resize(n_insert);
    rows.x.resize(3 * n_insert);
    rows.props.assign(size_t(LETHE_DEM_N_PROPERTIES) * n_insert, 0.0);
    for (long k = 0; k < n_insert; ++k)
      for (int d = 0; d < 3; ++d)
        rows.x[3 * k + d] = (mesh.lo[d] + (empty[k][d] + 0.5) * h[d]) + double(rand()) * maximum_range_for_randomness;
    const std::vector<double> diameters = sizes.sample(n_insert);
    for (long k = 0; k < n_insert; ++k)
      {
        const double d = std::fabs(diameters[k]), half = d * 0.5;
        double *pr = &rows.props[size_t(LETHE_DEM_N_PROPERTIES) * k];
        pr[0] = particle_type;
        pr[1] = d;
        pr[2] = t.density * 4.0 / 3.0 * M_PI * (half * half * half);
        for (int c = 0; c < 3; ++c)
          {
            pr[3 + c] = p.insertion.initial_velocity[c];
            pr[6 + c] = p.insertion.initial_omega[c];
          }
        rows.id[k] = first_id + uint32_t(k);
      }
    return rows;
  }

  DEMSolverB200::DEMSolverB200(const DEMParameters &prm, int device, std::ostream &log)
    : parameters(prm)
    , pcout(log)
  {
    const lethe_dem_config config = parameters.to_config();
    engine = std::make_unique<DEMEngine>(config, device);
    for (const auto &t : parameters.particle_types)
      {
        remaining_particles.push_back(t.number_of_particles);
        size_distributions.emplace_back(t, 0u);
      }
    if (parameters.insertion.method == "plane")
      plane_insertion = std::make_unique<PlaneInsertion>(parameters);
    if (plane_insertion || parameters.insertion.remove_particles)
      {
        occupied_cells.assign(size_t(parameters.mesh.n[0]) * parameters.mesh.n[1] * parameters.mesh.n[2], 0);
      }
    setup_boundaries();
  }

  void DEMSolverB200::setup_boundaries()
  {
    if (parameters.mesh.expand_particle_wall_contact_search)
      throw std::runtime_error("`expand particle-wall contact search` is not on the B200 path (box meshes do not need it)");
    engine->set_walls(box_wall_faces(parameters.mesh, parameters.outlet_boundaries(), parameters.periodic_directions()));
    if (!parameters.floating_walls.empty())
      {
        std::vector<double> pts, nrm, t0, t1;
        for (const auto &w : parameters.floating_walls)
          {
            pts.insert(pts.end(), w.point.begin(), w.point.end());
            nrm.insert(nrm.end(), w.normal.begin(), w.normal.end());
            t0.push_back(w.time_start);
            t1.push_back(w.time_end);
          }
        engine->set_floating_walls(pts, nrm, t0, t1);
      }
    // DEMSolver::setup_solid_objects (dem.cc:164-191); SerialSolid::setup_triangulation reads the
    // mesh, rotates it (GridTools::rotate(axis, angle)) and translates it (serial_solid.cc:163-216)
    for (const auto &so : parameters.solid_surfaces)
      {
        std::vector<double> v;
        std::vector<uint32_t> t;
        if (so.mesh_type == "dealii")
          dealii_simplex_surface(so.grid_type, so.grid_arguments, so.initial_refinement, v, t);
        else
          {
            const std::string path =
              (!so.mesh_file.empty() && so.mesh_file[0] == '/') ? so.mesh_file : parameters.prm_directory + "/" + so.mesh_file;
            read_msh_triangles(path, v, t);
          }
        Vec3 a = so.rotation_axis;
        const double an = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        for (auto &c : a)
          c /= an;
        const double co = std::cos(so.rotation_angle), si = std::sin(so.rotation_angle);
        const double K[3][3] = {{0, -a[2], a[1]}, {a[2], 0, -a[0]}, {-a[1], a[0], 0}};
        double R[3][3];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c)
            R[r][c] = co * (r == c ? 1.0 : 0.0) + si * K[r][c] + (1 - co) * a[r] * a[c];
        for (size_t k = 0; k < v.size() / 3; ++k)
          {
            const double x[3] = {v[3 * k], v[3 * k + 1], v[3 * k + 2]};
            for (int r = 0; r < 3; ++r)
              v[3 * k + r] = (R[r][0] * x[0] + R[r][1] * x[1] + R[r][2] * x[2]) + so.translation[r];
          }
        engine->add_solid_surface(v, t, so.translational_velocity.data(), so.angular_velocity.data(), so.center_of_rotation.data());
        solid_motion.push_back({so.translational_velocity, so.angular_velocity});
      }
    const double zero[3] = {0, 0, 0};
    for (const auto &bc : parameters.boundary_conditions)
      {
        if (bc.type == "rotational")
          engine->set_boundary_motion(bc.boundary_id, zero, bc.rotational_speed, bc.rotational_vector.data(),
                                      bc.point_on_rotational_vector.data());
        else if (bc.type == "translational")
          engine->set_boundary_motion(bc.boundary_id, bc.translational_velocity.data(), 0.0, zero, zero);
      }
  }

  ParticleRows list_insertion(const DEMParameters &p, uint32_t first_id, int particle_type)
  {
    const InsertionInfo &ins = p.insertion;
    const ParticleType &t = p.particle_types.at(particle_type);
    const size_t n = ins.list_x.size();
    auto at = [](const std::vector<double> &v, size_t k) { return k < v.size() ? v[k] : 0.0; };
    ParticleRows rows;
    for (size_t k = 0; k < n; ++k)
      {
        const double d = ins.list_diameters.size() == n ? ins.list_diameters[k] : t.average_diameter;
        rows.id.push_back(first_id + uint32_t(k));
        rows.x.insert(rows.x.end(), {ins.list_x[k], at(ins.list_y, k), at(ins.list_z, k)});
        const double h = d * 0.5;
        const double props[9] = {double(particle_type), d, t.density * 4.0 / 3.0 * M_PI * (h * h * h), at(ins.list_vx, k),
                                 at(ins.list_vy, k),    at(ins.list_vz, k), at(ins.list_wx, k), at(ins.list_wy, k), at(ins.list_wz, k)};
        rows.props.insert(rows.props.end(), props, props + 9);
      }
    return rows;
  }

  ParticleRows file_insertion(const DEMParameters &p, const std::string &path, long n_max, uint32_t first_id, int particle_type)
  {
    std::ifstream in(path);
    if (!in)
      throw std::runtime_error("cannot open insertion file " + path);
    std::string line;
    std::getline(in, line);
    std::vector<std::string> header;
    for (const auto &h : PrmSection::split(line, ';'))
      if (!PrmSection::trim(h).empty())
        header.push_back(PrmSection::trim(h));
    std::map<std::string, std::vector<double>> data;
    while (std::getline(in, line))
      {
        const auto v = PrmSection::split_doubles(line, ';');
        if (v.size() < header.size())
          continue;
        for (size_t k = 0; k < header.size(); ++k)
          data[header[k]].push_back(v[k]);
      }
    const ParticleType &t = p.particle_types.at(particle_type);
    const size_t n = std::min<size_t>(size_t(std::max(0l, n_max)), data["p_x"].size());
    ParticleRows rows;
    for (size_t k = 0; k < n; ++k)
      {
        const double d = data.at("diameters")[k], h = d * 0.5;
        rows.id.push_back(first_id + uint32_t(k));
        rows.x.insert(rows.x.end(), {data.at("p_x")[k], data.at("p_y")[k], data.at("p_z")[k]});
        const double props[9] = {double(particle_type), d, t.density * 4.0 / 3.0 * M_PI * (h * h * h), data.at("v_x")[k], data.at("v_y")[k],
                                 data.at("v_z")[k],     data.at("w_x")[k], data.at("w_y")[k], data.at("w_z")[k]};
        rows.props.insert(rows.props.end(), props, props + 9);
      }
    return rows;
  }

  // deal.II is an external dependency of the reference: restated are GridGenerator::hyper_cube /
  // hyper_rectangle<2,3> and the 2-D table of convert_hypercube_to_simplex_mesh (corners 0-3 in
  // lexicographic order, edge midpoints 4: x-low, 5: x-high, 6: y-low, 7: y-high, centre 8).
  void dealii_simplex_surface(const std::string &grid_type, const std::string &grid_arguments, long initial_refinement,
                              std::vector<double> &vertices3, std::vector<uint32_t> &triangles3)
  {
    const auto args = PrmSection::split(grid_arguments, ':');
    double p1[2], p2[2];
    if (grid_type == "hyper_cube")
      {
        p1[0] = p1[1] = std::stod(args.at(0));
        p2[0] = p2[1] = std::stod(args.at(1));
      }
    else if (grid_type == "hyper_rectangle")
      {
        const auto a = PrmSection::split_doubles(args.at(0), ','), b = PrmSection::split_doubles(args.at(1), ',');
        p1[0] = a.at(0), p1[1] = a.at(1), p2[0] = b.at(0), p2[1] = b.at(1);
      }
    else
      throw std::runtime_error("solid surfaces: dealii grid type `" + grid_type + "` is not generated here (hyper_cube, hyper_rectangle)");
    const long n = 1L << initial_refinement;
    const double hx = (p2[0] - p1[0]) / double(n), hy = (p2[1] - p1[1]) / double(n);
    static const int table[8][3] = {{0, 6, 4}, {8, 4, 6}, {8, 6, 5}, {1, 5, 6}, {2, 4, 7}, {8, 7, 4}, {8, 5, 7}, {3, 7, 5}};
    std::map<std::pair<long, long>, uint32_t> index; // half-cell lattice coordinates -> vertex
    auto vertex = [&](long i2, long j2) {
      auto it = index.find({i2, j2});
      if (it != index.end())
        return it->second;
      const uint32_t v = uint32_t(vertices3.size() / 3);
      index[{i2, j2}] = v;
      vertices3.push_back(p1[0] + 0.5 * double(i2) * hx);
      vertices3.push_back(p1[1] + 0.5 * double(j2) * hy);
      vertices3.push_back(0.0);
      return v;
    };
    std::vector<std::pair<long, long>> cells; // z-order of the refined quadrilaterals
    for (long k = 0; k < n * n; ++k)
      {
        long i = 0, j = 0;
        for (long b = 0; b < initial_refinement; ++b)
          {
            i |= ((k >> (2 * b)) & 1) << b;
            j |= ((k >> (2 * b + 1)) & 1) << b;
          }
        cells.push_back({i, j});
      }
    for (const auto &c : cells) // the quadrilateral mesh's own vertices come first
      for (long dj = 0; dj <= 2; dj += 2)
        for (long di = 0; di <= 2; di += 2)
          vertex(2 * c.first + di, 2 * c.second + dj);
    for (const auto &c : cells)
      {
        const long i = 2 * c.first, j = 2 * c.second;
        const uint32_t local[9] = {vertex(i, j),         vertex(i + 2, j),     vertex(i, j + 2),
                                   vertex(i + 2, j + 2), vertex(i, j + 1),     vertex(i + 2, j + 1),
                                   vertex(i + 1, j),     vertex(i + 1, j + 2), vertex(i + 1, j + 1)};
        for (const auto &t : table)
          for (int k = 0; k < 3; ++k)
            triangles3.push_back(local[t[k]]);
      }
  }

  void read_msh_triangles(const std::string &path, std::vector<double> &vertices3, std::vector<uint32_t> &triangles3)
  {
    std::ifstream in(path);
    if (!in)
      throw std::runtime_error("cannot open solid surface mesh " + path);
    std::map<std::string, std::vector<std::string>> sections;
    std::string line, name;
    while (std::getline(in, line))
      {
        line = PrmSection::trim(line);
        if (line.empty())
          continue;
        if (line[0] == '$')
          {
            name = line.rfind("$End", 0) == 0 ? "" : line.substr(1);
            continue;
          }
        if (!name.empty())
          sections[name].push_back(line);
      }
    auto numbers = [](const std::string &s) {
      std::vector<double> v;
      std::istringstream is(s);
      double x;
      while (is >> x)
        v.push_back(x);
      return v;
    };
    if (!sections.count("MeshFormat") || !sections.count("Nodes") || !sections.count("Elements"))
      throw std::runtime_error("not a gmsh ASCII file: " + path);
    const double version = numbers(sections["MeshFormat"][0]).at(0);
    std::map<long, std::array<double, 3>> nodes;
    std::vector<std::array<long, 3>> tris;
    const auto &nb = sections["Nodes"];
    const auto &eb = sections["Elements"];
    if (version >= 4.0)
      {
        size_t k = 1;
        for (long b = 0, nblocks = long(numbers(nb[0]).at(0)); b < nblocks; ++b)
          {
            const long n_in_block = long(numbers(nb[k]).at(3));
            for (long i = 0; i < n_in_block; ++i)
              {
                const auto xyz = numbers(nb[k + 1 + n_in_block + i]);
                nodes[long(numbers(nb[k + 1 + i]).at(0))] = {{xyz.at(0), xyz.at(1), xyz.at(2)}};
              }
            k += 1 + 2 * size_t(n_in_block);
          }
        k = 1;
        for (long b = 0, nblocks = long(numbers(eb[0]).at(0)); b < nblocks; ++b)
          {
            const auto head = numbers(eb[k]);
            const long etype = long(head.at(2)), n_in_block = long(head.at(3));
            for (long i = 0; i < n_in_block; ++i)
              if (etype == 2)
                {
                  const auto e = numbers(eb[k + 1 + i]);
                  tris.push_back({{long(e.at(1)), long(e.at(2)), long(e.at(3))}});
                }
            k += 1 + size_t(n_in_block);
          }
      }
    else
      {
        for (long i = 0, n = long(numbers(nb[0]).at(0)); i < n; ++i)
          {
            const auto v = numbers(nb[1 + i]);
            nodes[long(v.at(0))] = {{v.at(1), v.at(2), v.at(3)}};
          }
        for (long i = 0, n = long(numbers(eb[0]).at(0)); i < n; ++i)
          {
            const auto e = numbers(eb[1 + i]);
            if (long(e.at(1)) == 2)
              {
                const size_t o = 3 + size_t(e.at(2));
                tris.push_back({{long(e.at(o)), long(e.at(o + 1)), long(e.at(o + 2))}});
              }
          }
      }
    std::map<long, uint32_t> index;
    vertices3.clear();
    for (const auto &kv : nodes) // std::map: ascending node tag
      {
        index[kv.first] = uint32_t(index.size());
        vertices3.insert(vertices3.end(), kv.second.begin(), kv.second.end());
      }
    triangles3.clear();
    for (const auto &t : tris)
      for (long v : t)
        triangles3.push_back(index.at(v));
  }

  bool DEMSolverB200::insertion_due() const
  {
    const long f = parameters.insertion.frequency;
    if (f == 0)
      return false;
    return (iteration_number % f) == 1 || iteration_number == 1;
  }

  void DEMSolverB200::insert_particles()
  {
    const int last = int(parameters.particle_types.size()) - 1;
    if (remaining_particles[current_inserting_type] == 0 && current_inserting_type != last)
      ++current_inserting_type;
    const long remaining = remaining_particles[current_inserting_type];
    if (remaining == 0)
      return;
    if (parameters.insertion.remove_particles)
      remove_particles_in_box();
    const long n = std::min(parameters.insertion.inserted_this_step, remaining);
    ParticleRows rows;
    if (parameters.insertion.method == "file")
      {
        const auto &files = parameters.insertion.input_files;
        std::string path = files.at(current_file_id++ % files.size());
        if (path[0] != '/')
          path = parameters.prm_directory + "/" + path;
        rows = file_insertion(parameters, path, remaining, next_id, current_inserting_type);
      }
    else if (parameters.insertion.method == "list")
      rows = list_insertion(parameters, next_id, current_inserting_type);
    else if (parameters.insertion.method == "plane")
      rows = plane_insertion->insert(parameters, occupied_cells, remaining, next_id, current_inserting_type,
                                     size_distributions.at(current_inserting_type));
    else if (parameters.insertion.method != "volume")
      throw std::runtime_error("insertion method `" + parameters.insertion.method + "` is not mirrored by this host");
    else
      rows = volume_insertion(parameters, n, next_id, current_inserting_type, size_distributions.at(current_inserting_type));
    engine->add_particles(rows); // triggers the contact search (DEMActionManager::particle_insertion_step)
    next_id += uint32_t(rows.size());
    remaining_particles[current_inserting_type] -= long(rows.size());
  }

  long DEMSolverB200::cell_of(const double *x) const
  {
    const Mesh &mesh = parameters.mesh;
    const Vec3 h = mesh.cell_size();
    long c[3];
    for (int d = 0; d < 3; ++d)
      {
        c[d] = long(std::floor((x[d] - mesh.lo[d]) / h[d]));
        if (c[d] < 0 || c[d] >= mesh.n[d])
          return -1;
      }
    return c[0] + mesh.n[0] * (c[1] + mesh.n[1] * c[2]);
  }

  // Every particle REGISTERED in a cell whose 8 vertices are in the box goes, and of the cells with
  // some vertices in the box those particles whose position is in the box. The C ABI needs no
  // removal call: the particle is moved out of the triangulation — how the reference itself loses
  // particles — and the sort of this iteration drops it. New particles then take the lowest free
  // ids again (ParticleHandler::get_next_free_particle_index).
  void DEMSolverB200::remove_particles_in_box()
  {
    const Mesh &mesh = parameters.mesh;
    const Vec3 h = mesh.cell_size();
    const Vec3 &lo = parameters.insertion.removal_box_point_1, &hi = parameters.insertion.removal_box_point_2;
    const ParticleRows all = engine->get_particles();
    ParticleRows gone;
    uint32_t next_free = 0;
    for (size_t q = 0; q < all.size(); ++q)
      {
        long cell = all.id[q] < registered_cell.size() ? registered_cell[all.id[q]] : -1;
        if (cell < 0)
          cell = cell_of(&all.x[3 * q]);
        const long c[3] = {cell % mesh.n[0], (cell / mesh.n[0]) % mesh.n[1], cell / (long(mesh.n[0]) * mesh.n[1])};
        int vertices_inside = 0;
        for (int v = 0; v < 8 && cell >= 0; ++v)
          {
            bool inside = true;
            for (int d = 0; d < 3; ++d)
              {
                const double coordinate = mesh.lo[d] + (c[d] + ((v >> d) & 1)) * h[d];
                inside = inside && lo[d] <= coordinate && coordinate <= hi[d];
              }
            vertices_inside += inside;
          }
        bool position_inside = true;
        for (int d = 0; d < 3; ++d)
          position_inside = position_inside && lo[d] <= all.x[3 * q + d] && all.x[3 * q + d] <= hi[d];
        if (vertices_inside == 8 || (vertices_inside > 0 && position_inside))
          {
            gone.id.push_back(all.id[q]);
            for (int d = 0; d < 3; ++d)
              gone.x.push_back(mesh.hi[d] + 10.0 * (mesh.hi[d] - mesh.lo[d]));
            gone.props.insert(gone.props.end(), all.props.begin() + long(LETHE_DEM_N_PROPERTIES * q),
                              all.props.begin() + long(LETHE_DEM_N_PROPERTIES * (q + 1)));
          }
        else
          next_free = std::max(next_free, all.id[q] + 1);
      }
    if (gone.size())
      {
        engine->step_host(0, gone);
        next_id = next_free;
      }
  }

  bool DEMSolverB200::is_at_end() const
  {
    // simulation_control.cc:371-378
    const double margin = std::max(1e-6 * parameters.time_step, 1e-12 * parameters.time_end);
    return current_time >= (parameters.time_end - margin);
  }

  void DEMSolverB200::print_progression()
  {
    std::stringstream ss;
    ss << "Transient iteration: " << std::setw(8) << std::left << iteration_number << " Time: " << std::setw(8) << std::left
       << current_time << " Time step: " << std::setw(8) << std::left << parameters.time_step;
    const std::string line(ss.str().size() + 1, '*');
    pcout << '\n' << line << '\n' << ss.str() << '\n' << line << '\n';
  }

  void DEMSolverB200::report_statistics()
  {
    const lethe_dem_stats s = engine->stats();
    const double built = double(s.n_rebuilds) - list_total;
    list_max = std::max(built, list_max);
    list_min = std::min(built, list_min);
    list_total += built;
    const double list_avg = list_total / double(std::max<unsigned long>(iteration_number, 1)) * double(parameters.log_frequency);
    const double n = double(std::max<uint64_t>(s.n_particles, 1));
    auto row = [&](const char *name, double mn, double mx, double avg, double tot) {
      pcout << "| " << std::setw(30) << std::left << name << "| " << std::scientific << std::setprecision(6) << mn << " | " << mx
            << " | " << avg << " | " << tot << " |\n";
    };
    pcout << "| " << std::setw(30) << std::left << "Variable"
          << "| Min          | Max          | Average      | Total        |\n";
    row("Contact list generation", list_min, list_max, list_avg, list_total);
    row("Velocity magnitude", s.v_min, s.v_max, s.v_sum / n, s.v_sum);
    row("Angular velocity magnitude", s.omega_min, s.omega_max, s.omega_sum / n, s.omega_sum);
    row("Translational kinetic energy", s.ke_trans_min, s.ke_trans_max, s.ke_trans_sum / n, s.ke_trans_sum);
    row("Rotational kinetic energy", s.ke_rot_min, s.ke_rot_max, s.ke_rot_sum / n, s.ke_rot_sum);
    pcout << std::defaultfloat;
  }

  void DEMSolverB200::solve()
  {
    // Steps between two host-visible events (insertion, log line) are handed to the engine
    // in one call: the C ABI runs them back to back without returning to the host.
    uint64_t pending = 0;
    auto flush = [&] {
      if (pending)
        engine->step(pending);
      pending = 0;
    };
    while (!is_at_end())
      {
        ++iteration_number;
        current_time += parameters.time_step;
        if (is_verbose_iteration())
          {
            flush();
            print_progression();
            report_statistics();
          }
        // SerialSolid::move_solid_triangulation evaluates the velocity functions at the previous
        // time (serial_solid.cc:343-352): hand new values over before the step that uses them
        for (size_t k = 0; k < parameters.solid_surfaces.size(); ++k)
          {
            Vec3 tv, av;
            parameters.solid_surfaces[k].velocities_at(current_time - parameters.time_step, tv, av);
            if (tv != solid_motion[k].first || av != solid_motion[k].second)
              {
                flush();
                engine->set_solid_motion(int(k), tv.data(), av.data());
                solid_motion[k] = {tv, av};
              }
          }
        bool any_left = false;
        for (long r : remaining_particles)
          any_left = any_left || r > 0;
        if (insertion_due())
          {
            flush(); // the insertion belongs to this iteration: run the earlier ones first
            if (any_left)
              insert_particles();
            // action_manager->particle_insertion_step() at every insertion iteration, whether or
            // not particles are left to insert (dem.cc:494-500)
            engine->force_contact_search(false);
          }
        ++pending;
        if (plane_insertion || parameters.insertion.remove_particles)
          {
            // what a sort in this iteration registers: the positions it sees
            const ParticleRows seen = engine->get_particles();
            const uint64_t searches = engine->stats().n_rebuilds;
            flush();
            if (engine->stats().n_rebuilds != searches)
              {
                const Mesh &mesh = parameters.mesh;
                occupied_cells.assign(size_t(mesh.n[0]) * mesh.n[1] * mesh.n[2], 0);
                registered_cell.assign(registered_cell.size(), -1);
                for (size_t q = 0; q < seen.size(); ++q)
                  {
                    const long cell = cell_of(&seen.x[3 * q]);
                    if (seen.id[q] >= registered_cell.size())
                      registered_cell.resize(size_t(seen.id[q]) + 1, -1);
                    registered_cell[seen.id[q]] = cell;
                    if (cell >= 0)
                      occupied_cells[size_t(cell)] = 1;
                  }
              }
          }
      }
    flush();
    // closing half kick of the velocity-Verlet scheme (dem.cc:1249-1259)
    engine->synchronize_velocities();
  }

  void DEMSolverB200::print_xyz(std::ostream &out)
  {
    const ParticleRows r = engine->get_particles(); // sorted by id
    out << "id, type, dp, x, y, z \n";
    for (size_t k = 0; k < r.size(); ++k)
      {
        const double *p = &r.props[size_t(LETHE_DEM_N_PROPERTIES) * k];
        out << r.id[k] << " " << int(p[0]) << " " << std::fixed << std::setprecision(5) << p[1] << " " << std::setprecision(4)
            << r.x[3 * k] << " " << r.x[3 * k + 1] << " " << r.x[3 * k + 2] << "\n";
      }
    out << std::defaultfloat;
  }
} // namespace lethe_b200
