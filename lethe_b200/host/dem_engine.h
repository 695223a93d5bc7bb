// dem_engine.h — RAII C++ handle over the C ABI of include/lethe_dem.h.
// Error behaviour mirrors the reference: a failing call throws std::runtime_error, which
// the application's main() reports and turns into exit code 1
// (applications/lethe-particles/dem.cc:151-177). There is no CPU fallback behind it.
//
// LETHE_DEM_ABI_PREFIX lets the *tests* compile this same host code against the CPU
// oracle's identically-shaped `oracle_dem_*` symbols; product builds never define it.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lethe_dem.h"

#ifdef LETHE_DEM_ABI_PREFIX
#  define LETHE_DEM_CAT2(a, b) a##b
#  define LETHE_DEM_CAT(a, b) LETHE_DEM_CAT2(a, b)
#  define LETHE_DEM_FN(name) LETHE_DEM_CAT(LETHE_DEM_ABI_PREFIX, name)
extern "C" {
int LETHE_DEM_FN(create)(const lethe_dem_config *, int, lethe_dem_ctx **);
void LETHE_DEM_FN(destroy)(lethe_dem_ctx *);
const char *LETHE_DEM_FN(last_error)(const lethe_dem_ctx *);
const char *LETHE_DEM_FN(create_error)(void);
int LETHE_DEM_FN(set_particles)(lethe_dem_ctx *, uint64_t, const uint32_t *, const double *, const double *);
int LETHE_DEM_FN(add_particles)(lethe_dem_ctx *, uint64_t, const uint32_t *, const double *, const double *);
int LETHE_DEM_FN(n_particles)(lethe_dem_ctx *, uint64_t *);
int LETHE_DEM_FN(get_particles)(lethe_dem_ctx *, uint64_t, uint64_t *, uint32_t *, double *, double *);
int LETHE_DEM_FN(set_walls)(lethe_dem_ctx *, uint64_t, const lethe_wall_face *);
int LETHE_DEM_FN(set_floating_walls)(lethe_dem_ctx *, int32_t, const double *, const double *, const double *, const double *);
int LETHE_DEM_FN(set_boundary_motion)(lethe_dem_ctx *, uint32_t, const double *, double, const double *, const double *);
int LETHE_DEM_FN(add_solid_surface)(lethe_dem_ctx *, uint32_t, const double *, uint32_t, const uint32_t *, const double *, const double *,
                                    const double *, int32_t *);
int LETHE_DEM_FN(set_solid_motion)(lethe_dem_ctx *, int32_t, const double *, const double *);
int LETHE_DEM_FN(step)(lethe_dem_ctx *, uint64_t);
int LETHE_DEM_FN(step_host)(lethe_dem_ctx *, uint64_t, uint64_t, const uint32_t *, double *, double *);
int LETHE_DEM_FN(step_host_state)(lethe_dem_ctx *, uint64_t, uint64_t, const uint32_t *, double *);
int LETHE_DEM_FN(set_external_loads)(lethe_dem_ctx *, uint64_t, const uint32_t *, const double *, const double *);
int LETHE_DEM_FN(restart_integration)(lethe_dem_ctx *);
int LETHE_DEM_FN(synchronize_velocities)(lethe_dem_ctx *);
int LETHE_DEM_FN(force_contact_search)(lethe_dem_ctx *, int);
int LETHE_DEM_FN(get_stats)(lethe_dem_ctx *, lethe_dem_stats *);
}
#else
#  define LETHE_DEM_FN(name) lethe_dem_##name
#endif

namespace lethe_b200
{
  struct ParticleRows
  {
    std::vector<uint32_t> id;
    std::vector<double> x;     // [n][3]
    std::vector<double> props; // [n][9], PropertiesIndex order (dem_properties.h:54-76)
    size_t size() const { return id.size(); }
  };

  class DEMEngine
  {
  public:
    DEMEngine(const lethe_dem_config &config, int device)
    {
      if (LETHE_DEM_FN(create)(&config, device, &ctx) != 0)
        throw std::runtime_error(std::string("lethe_dem_create: ") + LETHE_DEM_FN(create_error)());
    }
    ~DEMEngine()
    {
      if (ctx)
        LETHE_DEM_FN(destroy)(ctx);
    }
    DEMEngine(const DEMEngine &) = delete;
    DEMEngine &operator=(const DEMEngine &) = delete;

    void set_particles(const ParticleRows &r) { check(LETHE_DEM_FN(set_particles)(ctx, r.size(), r.id.data(), r.x.data(), r.props.data())); }
    void add_particles(const ParticleRows &r) { check(LETHE_DEM_FN(add_particles)(ctx, r.size(), r.id.data(), r.x.data(), r.props.data())); }
    uint64_t n_particles()
    {
      uint64_t n = 0;
      check(LETHE_DEM_FN(n_particles)(ctx, &n));
      return n;
    }
    ParticleRows get_particles()
    {
      ParticleRows r;
      const uint64_t n = n_particles();
      r.id.resize(n);
      r.x.resize(3 * n);
      r.props.resize(LETHE_DEM_N_PROPERTIES * n);
      uint64_t got = 0;
      check(LETHE_DEM_FN(get_particles)(ctx, n, &got, r.id.data(), r.x.data(), r.props.data()));
      r.id.resize(got);
      r.x.resize(3 * got);
      r.props.resize(LETHE_DEM_N_PROPERTIES * got);
      return r;
    }
    void set_walls(const std::vector<lethe_wall_face> &faces) { check(LETHE_DEM_FN(set_walls)(ctx, faces.size(), faces.data())); }
    void set_floating_walls(const std::vector<double> &point3, const std::vector<double> &normal3, const std::vector<double> &t0,
                            const std::vector<double> &t1)
    {
      check(LETHE_DEM_FN(set_floating_walls)(ctx, int32_t(t0.size()), point3.data(), normal3.data(), t0.data(), t1.data()));
    }
    void set_boundary_motion(uint32_t boundary_id, const double v[3], double speed, const double axis[3], const double point[3])
    {
      check(LETHE_DEM_FN(set_boundary_motion)(ctx, boundary_id, v, speed, axis, point));
    }
    // DEMSolver::setup_solid_objects (dem.cc:164-191): one SerialSolid<2,3> per call
    int add_solid_surface(const std::vector<double> &vertices3, const std::vector<uint32_t> &triangles3, const double tv[3],
                          const double av[3], const double center[3])
    {
      int32_t index = -1;
      check(LETHE_DEM_FN(add_solid_surface)(ctx, uint32_t(vertices3.size() / 3), vertices3.data(), uint32_t(triangles3.size() / 3),
                                            triangles3.data(), tv, av, center, &index));
      return index;
    }
    void set_solid_motion(int solid, const double tv[3], const double av[3]) { check(LETHE_DEM_FN(set_solid_motion)(ctx, solid, tv, av)); }
    void step(uint64_t n_steps) { check(LETHE_DEM_FN(step)(ctx, n_steps)); }
    // overwrite the state of the listed particles, then n_steps steps (rows come back updated)
    void step_host(uint64_t n_steps, ParticleRows &r)
    {
      check(LETHE_DEM_FN(step_host)(ctx, n_steps, r.size(), r.id.data(), r.x.data(), r.props.data()));
    }
    void synchronize_velocities() { check(LETHE_DEM_FN(synchronize_velocities)(ctx)); }
    void force_contact_search(bool clear_tangential_displacement)
    {
      check(LETHE_DEM_FN(force_contact_search)(ctx, clear_tangential_displacement ? 1 : 0));
    }
    // rows of (x, v, omega) only; ids == nullptr reuses the id table of the previous call
    void step_host_state(uint64_t n_steps, uint64_t n, const uint32_t *ids, double *state9)
    {
      check(LETHE_DEM_FN(step_host_state)(ctx, n_steps, n, ids, state9));
    }
    // CFD-DEM: fluid-particle interaction loads per particle id (torque3 may be nullptr; n = 0 clears)
    void set_external_loads(uint64_t n, const uint32_t *ids, const double *force3, const double *torque3)
    {
      check(LETHE_DEM_FN(set_external_loads)(ctx, n, ids, force3, torque3));
    }
    void restart_integration() { check(LETHE_DEM_FN(restart_integration)(ctx)); }
    lethe_dem_stats stats()
    {
      lethe_dem_stats s;
      check(LETHE_DEM_FN(get_stats)(ctx, &s));
      return s;
    }

  private:
    void check(int rc)
    {
      if (rc != 0)
        throw std::runtime_error(LETHE_DEM_FN(last_error)(ctx));
    }
    lethe_dem_ctx *ctx = nullptr;
  };
} // namespace lethe_b200
