// dem_physics.cuh — per-contact physics of the DEM step as device functions.
//
// FP64 throughout, compiled with -fmad=false: the reference is built without
// FMA contraction and its goldens are sensitive to the last bit in the tensile
// tail of damped contacts, so every expression below keeps the
// reference's evaluation order: component-ordered dot products, vector / scalar
// as multiplication by the reciprocal (deal.II Tensor::operator/=), and the
// reference's literal constants. Reference: include/dem/particle_particle_contact_force.h,
// include/dem/rolling_resistance_torque_models.h, include/dem/particle_wall_contact_force.h,
// include/dem/particle_wall_rolling_resistance_torque.h (line ranges at each function).
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>

#include "../../include/lethe_dem.h"
#include "dem_math.cuh"

namespace dem
{
  struct vec3
  {
    double x, y, z;
  };
  __host__ __device__ __forceinline__ vec3 v3(double a, double b, double c) { return vec3{a, b, c}; }
  __host__ __device__ __forceinline__ vec3 operator+(vec3 a, vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
  __host__ __device__ __forceinline__ vec3 operator-(vec3 a, vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
  __host__ __device__ __forceinline__ vec3 operator-(vec3 a) { return v3(-a.x, -a.y, -a.z); }
  __host__ __device__ __forceinline__ vec3 operator*(double s, vec3 a) { return v3(s * a.x, s * a.y, s * a.z); }
  __host__ __device__ __forceinline__ vec3 operator*(vec3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
  // Tensor / scalar: reciprocal-multiply
  __host__ __device__ __forceinline__ vec3 operator/(vec3 a, double s)
  {
    const double inv = 1.0 / s;
    return v3(a.x * inv, a.y * inv, a.z * inv);
  }
  __host__ __device__ __forceinline__ double dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
  __host__ __device__ __forceinline__ double norm2(vec3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
  __host__ __device__ __forceinline__ double norm(vec3 a) { return sqrt(norm2(a)); }
  __host__ __device__ __forceinline__ vec3 cross(vec3 a, vec3 b)
  {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
  }
  // Point::distance_square: differences taken as (a - b), summed from zero
  __host__ __device__ __forceinline__ double dist2(vec3 a, vec3 b)
  {
    const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return ((0.0 + dx * dx) + dy * dy) + dz * dz;
  }
  __host__ __device__ __forceinline__ double sqr(double v) { return v * v; }
  __host__ __device__ __forceinline__ double cub(double v) { return v * v * v; }

  // Effective property tables (host-computed with the reference formulas,
  // particle_particle_contact_force.h:1639-1746, particle_wall_contact_force.cc:588-694).
  struct MaterialTables
  {
    int n_types;
    double Y[25], G[25], beta[25], mu[25], roll_fric[25], roll_visc[25], gamma[25], hamaker[25];
    double wY[5], wG[5], wbeta[5], wmu[5], wroll_fric[5], wroll_visc[5], wgamma[5], whamaker[5];
    double pp_force_threshold, pw_force_threshold;
    double f_coefficient_epsd;
  };

  struct ParticleView
  {
    vec3 x;
    double d;
    vec3 v;
    double m;
    vec3 w;
    int type;
  };

  // R* = d d / (2 (d + d)) and m* = m m / (m + m) of the particle that owns the list row,
  // computed once per particle per step: valid for every pair whose partner has bit-identical
  // diameter and mass (all pairs of a monodisperse type), where it saves two IEEE divisions.
  struct SelfPair
  {
    double effective_radius, effective_mass;
  };

  struct PairResult
  {
    vec3 normal_force, tangential_force, torque_one, torque_two, rolling;
  };

  // rolling_resistance_torque_models.h:12-250 (dispatch: …contact_force.h:643-698)
  template <int ROLLING>
  __device__ __forceinline__ vec3 pp_rolling(const MaterialTables &mt, double effective_r, const ParticleView &p1,
                                             const ParticleView &p2, double rolling_friction_coeff,
                                             double rolling_viscous_damping_coeff, double dt, double normal_spring_constant,
                                             double normal_force_norm, vec3 n, vec3 &cumulative)
  {
    if constexpr (ROLLING == LETHE_ROLLING_NONE)
      return v3(0, 0, 0);
    else if constexpr (ROLLING == LETHE_ROLLING_CONSTANT)
      {
        const vec3 omega_ij = p1.w - p2.w;
        const vec3 dir = omega_ij / (norm(omega_ij) + DBL_MIN);
        return (-rolling_friction_coeff * effective_r * normal_force_norm) * dir;
      }
    else if constexpr (ROLLING == LETHE_ROLLING_VISCOUS)
      {
        const vec3 omega_ij = p1.w - p2.w;
        const vec3 dir = omega_ij / (norm(omega_ij) + DBL_MIN);
        const vec3 v_omega = cross(p1.w, (p1.d * 0.5) * n) - cross(p2.w, (p2.d * 0.5) * (-n));
        return (-rolling_friction_coeff * effective_r * normal_force_norm * norm(v_omega)) * dir;
      }
    else
      {
        const double mu_r_times_R_e = rolling_friction_coeff * effective_r;
        const vec3 omega_ij = p1.w - p2.w;
        const vec3 omega_perp = omega_ij - dot(omega_ij, n) * n;
        const vec3 delta_theta = dt * omega_perp;
        const double K_r = 2.25 * normal_spring_constant * sqr(mu_r_times_R_e);
        cumulative = cumulative - K_r * delta_theta;
        const double M_r_max = mu_r_times_R_e * normal_force_norm;
        const double spring_norm = norm(cumulative);
        const double I_i = 1.4 * p1.m * sqr(0.5 * p1.d);
        const double I_j = 1.4 * p2.m * sqr(0.5 * p2.d);
        const double I_e = I_i * I_j / (I_i + I_j);
        const double C_r = rolling_viscous_damping_coeff * 2. * sqrt(I_e * K_r);
        if (spring_norm > M_r_max)
          {
            cumulative = cumulative * (M_r_max / spring_norm);
            return cumulative - (mt.f_coefficient_epsd * C_r) * omega_perp;
          }
        return cumulative - C_r * omega_perp;
      }
  }

  // Ferrari solution of the JKR contact-patch quartic (…contact_force.h:1356-1372;
  // the wall variant clamps root1 at 0, particle_wall_contact_force.h:~905)
  __device__ __forceinline__ double jkr_contact_radius(double R, double overlap, double gamma, double Y, bool clamp_root1)
  {
    const double c0 = sqr(R * overlap);
    const double c1 = -2. * sqr(R) * M_PI * gamma / Y;
    const double c2 = -2. * overlap * R;
    const double P = -sqr(c2) / 12. - c0;
    const double Q = -cub(c2) / 108. + c0 * c2 / 3. - sqr(c1) * 0.125;
    double root1 = clamp_root1 ? fmax(0., (0.25 * sqr(Q)) + (cub(P) / 27.)) : 0.25 * sqr(Q) + cub(P) / 27.;
    const double U = glibc_cbrt(-0.5 * Q + sqrt(root1)); // std::cbrt of the reference's host (dem_math.cuh)
    const double s = -c2 * (5. / 6.) + U - P / (3. * U);
    const double w = sqrt(fmax(1e-16, c2 + 2. * s));
    const double lambda = 0.5 * c1 / w;
    const double root2 = fmax(1e-16, w * w - 4. * (c2 + s + lambda));
    return 0.5 * (w + sqrt(root2));
  }

  // update_contact_information (…contact_force.h:223-298)
  __device__ __forceinline__ void pp_update_contact_information(vec3 &tangential_displacement, vec3 &vt, double &vn, vec3 &n,
                                                                const ParticleView &p1, const ParticleView &p2, vec3 x2,
                                                                double distance, double dt)
  {
    // `distance` = sqrt(dist2(p1.x, x2)) from the caller is bit-identical to norm(x2 - p1.x):
    // the component differences are exact negatives, their squares and the component-ordered
    // sum are the same, so the second square root is not taken.
    const vec3 contact_vector = x2 - p1.x;
    n = contact_vector / distance;
    vec3 vrel = p1.v - p2.v;
    vrel = vrel + cross(0.5 * (p1.d * p1.w + p2.d * p2.w), n);
    vn = dot(vrel, n);
    vt = vrel - (vn * n);
    tangential_displacement = tangential_displacement + vt * dt;
    tangential_displacement = tangential_displacement - dot(tangential_displacement, n) * n;
  }

  // calculate_*_contact for the particle-particle models (…contact_force.h:723-1544).
  // `r` must be zero-initialised by the caller.  NOTE (documented deviation):
  // in the DMT non-contact branch the reference adds the cohesive term onto whatever
  // the previous pair of the same row left in its scratch tensors
  // (…contact_force.h:1847-1853,1532); that order-dependent artefact is not
  // reproduced — the scratch is zero for every pair here.
  template <int MODEL, int ROLLING>
  __device__ __forceinline__ void pp_calculate_contact(const MaterialTables &mt, vec3 &tangential_displacement,
                                                       vec3 &rolling_spring_torque, vec3 vt, double vn, vec3 n, double overlap,
                                                       double dt, const ParticleView &p1, const ParticleView &p2, PairResult &r,
                                                       const SelfPair &self)
  {
    const double d1 = p1.d, d2 = p2.d;
    double effective_radius, effective_mass;
    if (d1 == d2 && p1.m == p2.m)
      {
        // same operands, same expression: bit-identical to the general branch
        effective_radius = self.effective_radius;
        effective_mass = self.effective_mass;
      }
    else
      {
        effective_radius = (d1 * d2) / (2 * (d1 + d2));
        effective_mass = (p1.m * p2.m) / (p1.m + p2.m);
      }
    const int k = p1.type * mt.n_types + p2.type;
    const double Y = mt.Y[k], G = mt.G[k], beta = mt.beta[k], mu = mt.mu[k];
    const double roll_visc = mt.roll_visc[k], roll_fric = mt.roll_fric[k];

    if constexpr (MODEL == LETHE_PP_DMT)
      {
        constexpr double M_2PI = 2. * M_PI;
        const double gamma = mt.gamma[k], A = mt.hamaker[k];
        const double F_po = M_2PI * effective_radius * gamma;
        const double delta_0 = -sqrt(A * effective_radius / (6. * F_po));
        double cohesive_term;
        if (overlap > 0.)
          {
            cohesive_term = -F_po;
            pp_calculate_contact<LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP, ROLLING>(mt, tangential_displacement, rolling_spring_torque,
                                                                                vt, vn, n, overlap, dt, p1, p2, r, self);
          }
        else if (overlap > delta_0)
          {
            cohesive_term = -F_po;
            tangential_displacement = v3(0, 0, 0);
            rolling_spring_torque = v3(0, 0, 0);
          }
        else
          {
            cohesive_term = -A * effective_radius / (6. * sqr(overlap));
            tangential_displacement = v3(0, 0, 0);
            rolling_spring_torque = v3(0, 0, 0);
          }
        r.normal_force = r.normal_force + cohesive_term * n;
      }
    else if constexpr (MODEL == LETHE_PP_LINEAR)
      {
        const double kn =
          1.0667 * sqrt(effective_radius) * Y * pow_0_2_cr((0.9375 * effective_mass * 1.0 * 1.0 / (sqrt(effective_radius) * Y)));
        const double kt = kn * 0.4;
        const double etan = -2 * beta * sqrt(effective_mass * kn);
        const double etat = etan * 0.6324555320336759;
        const double normal_force_value = kn * overlap + etan * vn;
        r.normal_force = normal_force_value * n;
        const vec3 damping_tangential_force = etat * vt;
        r.tangential_force = (kt * tangential_displacement) + damping_tangential_force;
        const double coulomb_threshold = mu * normal_force_value;
        if (norm(r.tangential_force) > coulomb_threshold)
          {
            const vec3 limited = coulomb_threshold * (r.tangential_force / (norm(r.tangential_force) + DBL_MIN));
            tangential_displacement = (limited - damping_tangential_force) / (kt + DBL_MIN);
            r.tangential_force = (kt * tangential_displacement) + damping_tangential_force;
          }
        r.torque_one = cross(n, r.tangential_force * d1 * 0.5);
        r.torque_two = r.torque_one * d2 / d1;
        // the reference passes the two rolling coefficients in swapped order here (:841-851)
        r.rolling = pp_rolling<ROLLING>(mt, effective_radius, p1, p2, roll_visc, roll_fric, dt, kn, norm(r.normal_force), n,
                                        rolling_spring_torque);
      }
    else
      {
        const double radius_times_overlap_sqrt = sqrt(effective_radius * overlap);
        const double model_parameter_sn = 2.0 * Y * radius_times_overlap_sqrt;
        const double model_parameter_st = 8.0 * G * radius_times_overlap_sqrt;
        if constexpr (MODEL == LETHE_PP_HERTZ_JKR)
          {
            const double gamma = mt.gamma[k];
            const double a = jkr_contact_radius(effective_radius, overlap, gamma, Y, false);
            const double etan = -1.8257 * beta * sqrt(model_parameter_sn * effective_mass);
            const double kt = 8.0 * radius_times_overlap_sqrt * G;
            const double etat = etan * sqrt(model_parameter_st / model_parameter_sn);
            const double normal_force_coefficient =
              4. * cub(a) / (3. * effective_radius) * Y - sqrt(8. * M_PI * gamma * Y * cub(a));
            r.normal_force = (normal_force_coefficient + etan * vn) * n;
            r.tangential_force = kt * tangential_displacement + etat * vt;
            const double two_pull_off_force = 3. * M_PI * gamma * effective_radius;
            const double modified_coulomb_threshold = (normal_force_coefficient + two_pull_off_force) * mu;
            if (norm(r.tangential_force) > modified_coulomb_threshold)
              r.tangential_force = modified_coulomb_threshold * (r.tangential_force / (norm(r.tangential_force) + DBL_MIN));
            r.torque_one = cross(n, r.tangential_force * d1 * 0.5);
            r.torque_two = r.torque_one * d2 / d1;
            const double kn = 0.66665 * model_parameter_sn;
            r.rolling = pp_rolling<ROLLING>(mt, effective_radius, p1, p2, roll_fric, roll_visc, dt, kn, norm(r.normal_force), n,
                                            rolling_spring_torque);
          }
        else
          {
            const double kn = 0.66665 * model_parameter_sn;
            const double etan = -1.8257 * beta * sqrt(model_parameter_sn * effective_mass);
            const double kt = 8.0 * G * radius_times_overlap_sqrt;
            const double normal_force_value = kn * overlap + etan * vn;
            r.normal_force = normal_force_value * n;
            const double coulomb_threshold = mu * normal_force_value;
            if constexpr (MODEL == LETHE_PP_HERTZ)
              {
                r.tangential_force = kt * tangential_displacement;
                if (norm(r.tangential_force) > coulomb_threshold)
                  r.tangential_force = coulomb_threshold * (r.tangential_force / (norm(r.tangential_force) + DBL_MIN));
              }
            else
              {
                const double etat = etan * sqrt(model_parameter_st / model_parameter_sn);
                const vec3 damping_tangential_force = etat * vt;
                r.tangential_force = (kt * tangential_displacement) + damping_tangential_force;
                if (norm(r.tangential_force) > coulomb_threshold)
                  {
                    if constexpr (MODEL == LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP)
                      {
                        const vec3 limited = coulomb_threshold * (r.tangential_force / (norm(r.tangential_force) + DBL_MIN));
                        tangential_displacement = (limited - damping_tangential_force) / (kt + DBL_MIN);
                        r.tangential_force = (kt * tangential_displacement) + damping_tangential_force;
                      }
                    else
                      r.tangential_force = coulomb_threshold * (r.tangential_force / (norm(r.tangential_force) + DBL_MIN));
                  }
              }
            r.torque_one = cross(n, r.tangential_force * d1 * 0.5);
            r.torque_two = r.torque_one * d2 / d1;
            r.rolling = pp_rolling<ROLLING>(mt, effective_radius, p1, p2, roll_fric, roll_visc, dt, kn, norm(r.normal_force), n,
                                            rolling_spring_torque);
          }
      }
  }

  // ------------------------------------------------------------------ walls ---
  struct WallResult
  {
    vec3 normal_force, tangential_force, tangential_torque, rolling;
  };

  // particle_wall_rolling_resistance_torque.h:12-226
  __device__ __forceinline__ vec3 pw_rolling(int rolling_model, const MaterialTables &mt, double R, const ParticleView &p,
                                             double rolling_friction_coeff, double rolling_viscous_damping_coeff, double dt,
                                             double normal_spring_constant, double normal_force_norm, vec3 n, vec3 &cumulative)
  {
    switch (rolling_model)
      {
        case LETHE_ROLLING_NONE:
          return v3(0, 0, 0);
        case LETHE_ROLLING_CONSTANT:
          {
            const double omega_value = norm(p.w);
            const vec3 dir = p.w / (omega_value + DBL_MIN);
            return (-rolling_friction_coeff * R * normal_force_norm) * dir;
          }
        case LETHE_ROLLING_VISCOUS:
          {
            const double omega_value = norm(p.w);
            const vec3 dir = p.w / (omega_value + DBL_MIN);
            const vec3 v_omega = cross(p.w, R * n);
            return (-rolling_friction_coeff * R * normal_force_norm * norm(v_omega)) * dir;
          }
        default:
          {
            const double mu_r_times_R = rolling_friction_coeff * R;
            const vec3 omega_perp = p.w - dot(p.w, n) * n;
            const vec3 delta_theta = dt * omega_perp;
            const double K_r = 2.25 * normal_spring_constant * sqr(mu_r_times_R);
            cumulative = cumulative - K_r * delta_theta;
            const double M_r_max = mu_r_times_R * normal_force_norm;
            const double spring_norm = norm(cumulative);
            const double I_e = 1.4 * p.m * sqr(R);
            const double C_r = rolling_viscous_damping_coeff * 2. * sqrt(I_e * K_r);
            if (spring_norm > M_r_max)
              {
                cumulative = cumulative * (M_r_max / spring_norm);
                return cumulative - (mt.f_coefficient_epsd * C_r) * omega_perp;
              }
            return cumulative - C_r * omega_perp;
          }
      }
  }

  // calculate_nonlinear_contact (particle_wall_contact_force.h:723-833)
  __device__ __forceinline__ void pw_nonlinear(int rolling_model, const MaterialTables &mt, vec3 wall_normal,
                                               vec3 &tangential_displacement, vec3 &rolling_spring_torque, vec3 vt, double vn,
                                               double overlap, double dt, const ParticleView &p, WallResult &r)
  {
    const vec3 normal_vector = -wall_normal;
    const int type = p.type;
    const double Y = mt.wY[type], G = mt.wG[type], beta = mt.wbeta[type], mu = mt.wmu[type];
    const double roll_visc = mt.wroll_visc[type], roll_fric = mt.wroll_fric[type];
    const double R = p.d * 0.5;
    const double rs = sqrt(R * overlap);
    const double sn = 2.0 * Y * rs;
    const double st = 8.0 * G * rs;
    const double kn = 1.3333 * Y * rs;
    const double etan = 1.8257 * beta * sqrt(sn * p.m);
    const double kt = -8.0 * G * rs + DBL_MIN;
    const double etat = etan * sqrt(st / sn);
    r.normal_force = (kn * overlap + etan * vn) * normal_vector;
    const vec3 damping_tangential_force = etat * vt;
    r.tangential_force = kt * tangential_displacement + damping_tangential_force;
    const double coulomb_threshold = mu * norm(r.normal_force);
    const double tangential_force_norm = norm(r.tangential_force);
    if (tangential_force_norm > coulomb_threshold)
      {
        tangential_displacement =
          (coulomb_threshold * (r.tangential_force / (tangential_force_norm + DBL_MIN)) - damping_tangential_force) / (kt + DBL_MIN);
        r.tangential_force = (kt * tangential_displacement) + damping_tangential_force;
      }
    r.tangential_torque = cross((R * normal_vector), -r.tangential_force);
    r.rolling =
      pw_rolling(rolling_model, mt, R, p, roll_fric, roll_visc, dt, kn, norm(r.normal_force), wall_normal, rolling_spring_torque);
  }

  // calculate_{linear,JKR,DMT}_contact + dispatch (particle_wall_contact_force.h:350-440,610-1082);
  // wall_normal is contact_info.normal_vector (wall -> particle). `r` zero-initialised by the caller.
  __device__ inline void pw_calculate_contact(int model, int rolling_model, const MaterialTables &mt, vec3 wall_normal,
                                              vec3 &tangential_displacement, vec3 &rolling_spring_torque, vec3 vt, double vn,
                                              double overlap, double dt, const ParticleView &p, WallResult &r)
  {
    const vec3 normal_vector = -wall_normal;
    const int type = p.type;
    const double Y = mt.wY[type], G = mt.wG[type], beta = mt.wbeta[type], mu = mt.wmu[type];
    const double roll_visc = mt.wroll_visc[type], roll_fric = mt.wroll_fric[type];
    if (model == LETHE_PW_DMT)
      {
        constexpr double M_2PI = 2. * M_PI;
        const double R = 0.5 * p.d;
        const double gamma = mt.wgamma[type], A = mt.whamaker[type];
        const double F_po = M_2PI * R * gamma;
        const double delta_0 = -sqrt(A * R / (6. * F_po));
        double cohesive_term;
        if (overlap > 0.)
          {
            cohesive_term = -F_po;
            pw_nonlinear(rolling_model, mt, wall_normal, tangential_displacement, rolling_spring_torque, vt, vn, overlap, dt, p, r);
          }
        else if (overlap > delta_0)
          {
            cohesive_term = -F_po;
            tangential_displacement = v3(0, 0, 0);
            rolling_spring_torque = v3(0, 0, 0);
          }
        else
          {
            cohesive_term = -A * R / (6. * sqr(overlap));
            tangential_displacement = v3(0, 0, 0);
            rolling_spring_torque = v3(0, 0, 0);
          }
        r.normal_force = r.normal_force + cohesive_term * normal_vector;
        return;
      }
    if (model == LETHE_PW_LINEAR)
      {
        const double R = p.d * 0.5;
        const double rp_sqrt = sqrt(R);
        const double kn = 1.0667 * rp_sqrt * Y * pow_0_2_cr((0.9375 * p.m * 1.0 * 1.0 / (rp_sqrt * Y)));
        const double etan = 2 * beta * sqrt(p.m * kn);
        const double kt = -kn * 0.4;
        const double etat = etan * 0.6324555320336759;
        r.normal_force = (kn * overlap + etan * vn) * normal_vector;
        r.tangential_force = (kt * tangential_displacement + etat * vt);
        const double coulomb_threshold = mu * norm(r.normal_force);
        if (norm(r.tangential_force) > coulomb_threshold)
          {
            r.tangential_force = coulomb_threshold * (r.tangential_force / norm(r.tangential_force));
            tangential_displacement = r.tangential_force / (kt + DBL_MIN);
          }
        r.tangential_torque = cross((R * normal_vector), -r.tangential_force);
        r.rolling = pw_rolling(rolling_model, mt, R, p, roll_fric, roll_visc, dt, kn, norm(r.normal_force), wall_normal,
                               rolling_spring_torque);
        return;
      }
    if (model == LETHE_PW_JKR)
      {
        const double R = 0.5 * p.d;
        const double gamma = mt.wgamma[type];
        const double rs = sqrt(R * overlap);
        const double sn = 2.0 * Y * rs;
        const double st = 8.0 * G * rs;
        const double a = jkr_contact_radius(R, overlap, gamma, Y, true);
        const double etan = 1.8257 * beta * sqrt(sn * p.m);
        const double kt = -8.0 * G * rs;
        const double etat = etan * sqrt(st / (sn + DBL_MIN));
        const double normal_force_norm = 4. * Y * cub(a) / (3. * R) - sqrt(8. * M_PI * gamma * Y * cub(a)) + etan * vn;
        r.normal_force = normal_force_norm * normal_vector;
        const vec3 damping_tangential_force = etat * vt;
        r.tangential_force = kt * tangential_displacement + damping_tangential_force;
        const double modified_coulomb_threshold = (normal_force_norm + 3. * M_PI * gamma * R) * mu;
        const double tangential_force_norm = norm(r.tangential_force);
        if (tangential_force_norm > modified_coulomb_threshold)
          {
            tangential_displacement =
              (modified_coulomb_threshold * (r.tangential_force / (tangential_force_norm + DBL_MIN)) - damping_tangential_force) /
              (kt + DBL_MIN);
            r.tangential_force = (kt * tangential_displacement) + damping_tangential_force;
          }
        r.tangential_torque = cross((R * normal_vector), -r.tangential_force);
        const double kn = 0.66665 * sn;
        r.rolling = pw_rolling(rolling_model, mt, R, p, roll_fric, roll_visc, dt, kn, norm(r.normal_force), wall_normal,
                               rolling_spring_torque);
        return;
      }
    pw_nonlinear(rolling_model, mt, wall_normal, tangential_displacement, rolling_spring_torque, vt, vn, overlap, dt, p, r);
  }
} // namespace dem
