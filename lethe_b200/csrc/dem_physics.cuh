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
  // 3-vectors in the scalar type the pair model runs in: double everywhere (reference arithmetic),
  // float for the model part of the mixed-precision step kernel.
  template <class T> struct vec3_t
  {
    T x, y, z;
  };
  using vec3 = vec3_t<double>;
  using vec3f = vec3_t<float>;
  __host__ __device__ __forceinline__ vec3 v3(double a, double b, double c) { return vec3{a, b, c}; }
  template <class T> __host__ __device__ __forceinline__ vec3_t<T> v3t(T a, T b, T c) { return vec3_t<T>{a, b, c}; }
  template <class T> __host__ __device__ __forceinline__ vec3_t<T> operator+(vec3_t<T> a, vec3_t<T> b) { return v3t<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
  template <class T> __host__ __device__ __forceinline__ vec3_t<T> operator-(vec3_t<T> a, vec3_t<T> b) { return v3t<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
  template <class T> __host__ __device__ __forceinline__ vec3_t<T> operator-(vec3_t<T> a) { return v3t<T>(-a.x, -a.y, -a.z); }
  template <class T> __host__ __device__ __forceinline__ vec3_t<T> operator*(T s, vec3_t<T> a) { return v3t<T>(s * a.x, s * a.y, s * a.z); }
  template <class T> __host__ __device__ __forceinline__ vec3_t<T> operator*(vec3_t<T> a, T s) { return v3t<T>(a.x * s, a.y * s, a.z * s); }
  __host__ __device__ __forceinline__ vec3 operator*(int s, vec3 a) { return v3(s * a.x, s * a.y, s * a.z); }
  // Tensor / scalar: reciprocal-multiply
  template <class T> __host__ __device__ __forceinline__ vec3_t<T> operator/(vec3_t<T> a, T s)
  {
    const T inv = T(1.0) / s;
    return v3t<T>(a.x * inv, a.y * inv, a.z * inv);
  }
  template <class T> __host__ __device__ __forceinline__ T dot(vec3_t<T> a, vec3_t<T> b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
  template <class T> __host__ __device__ __forceinline__ T norm2(vec3_t<T> a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
  __host__ __device__ __forceinline__ double norm(vec3 a) { return sqrt(norm2(a)); }
  __host__ __device__ __forceinline__ float norm(vec3f a) { return sqrtf(norm2(a)); }
  template <class T> __host__ __device__ __forceinline__ vec3_t<T> cross(vec3_t<T> a, vec3_t<T> b)
  {
    return v3t<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
  }
  // Point::distance_square: differences taken as (a - b), summed from zero
  __host__ __device__ __forceinline__ double dist2(vec3 a, vec3 b)
  {
    const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return ((0.0 + dx * dx) + dy * dy) + dz * dz;
  }
  template <class T> __host__ __device__ __forceinline__ T sqr(T v) { return v * v; }
  template <class T> __host__ __device__ __forceinline__ T cub(T v) { return v * v * v; }
  __host__ __device__ __forceinline__ vec3f to_float(vec3 a) { return vec3f{float(a.x), float(a.y), float(a.z)}; }
  __host__ __device__ __forceinline__ vec3 to_double(vec3f a) { return vec3{double(a.x), double(a.y), double(a.z)}; }

  // scalar functions / constants per precision
  template <class T> struct real_traits;
  template <> struct real_traits<double>
  {
    static __host__ __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __host__ __device__ __forceinline__ double max(double a, double b) { return fmax(a, b); }
    static __host__ __device__ __forceinline__ double cbrt(double x) { return glibc_cbrt(x); } // std::cbrt of the reference's host (dem_math.cuh)
    static __host__ __device__ __forceinline__ double pow_0_2(double x) { return pow_0_2_cr(x); }
    static __host__ __device__ __forceinline__ double tiny() { return DBL_MIN; }
  };
  template <> struct real_traits<float>
  {
    static __host__ __device__ __forceinline__ float sqrt(float x) { return ::sqrtf(x); }
    static __host__ __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
    static __host__ __device__ __forceinline__ float cbrt(float x) { return ::cbrtf(x); }
    static __host__ __device__ __forceinline__ float pow_0_2(float x) { return ::powf(x, 0.2f); }
    // the reference adds DBL_MIN to denominators that may be zero; in float that role is FLT_MIN's
    static __host__ __device__ __forceinline__ float tiny() { return FLT_MIN; }
  };

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

  template <class T> struct ParticleView_t
  {
    vec3_t<T> x;
    T d;
    vec3_t<T> v;
    T m;
    vec3_t<T> w;
    int type;
  };
  using ParticleView = ParticleView_t<double>;

  // R* = d d / (2 (d + d)) and m* = m m / (m + m) of the particle that owns the list row,
  // computed once per particle per step: valid for every pair whose partner has bit-identical
  // diameter and mass (all pairs of a monodisperse type), where it saves two IEEE divisions.
  template <class T> struct SelfPair_t
  {
    T effective_radius, effective_mass;
  };
  using SelfPair = SelfPair_t<double>;

  template <class T> struct PairResult_t
  {
    vec3_t<T> normal_force, tangential_force, torque_one, torque_two, rolling;
  };
  using PairResult = PairResult_t<double>;

  // rolling_resistance_torque_models.h:12-250 (dispatch: …contact_force.h:643-698)
  template <int ROLLING, class T>
  __device__ __forceinline__ vec3_t<T> pp_rolling(const MaterialTables &mt, T effective_r, const ParticleView_t<T> &p1,
                                                  const ParticleView_t<T> &p2, T rolling_friction_coeff,
                                                  T rolling_viscous_damping_coeff, T dt, T normal_spring_constant,
                                                  T normal_force_norm, vec3_t<T> n, vec3_t<T> &cumulative)
  {
    using R = real_traits<T>;
    if constexpr (ROLLING == LETHE_ROLLING_NONE)
      return v3t<T>(0, 0, 0);
    else if constexpr (ROLLING == LETHE_ROLLING_CONSTANT)
      {
        const vec3_t<T> omega_ij = p1.w - p2.w;
        const vec3_t<T> dir = omega_ij / (norm(omega_ij) + R::tiny());
        return (-rolling_friction_coeff * effective_r * normal_force_norm) * dir;
      }
    else if constexpr (ROLLING == LETHE_ROLLING_VISCOUS)
      {
        const vec3_t<T> omega_ij = p1.w - p2.w;
        const vec3_t<T> dir = omega_ij / (norm(omega_ij) + R::tiny());
        const vec3_t<T> v_omega = cross(p1.w, (p1.d * T(0.5)) * n) - cross(p2.w, (p2.d * T(0.5)) * (-n));
        return (-rolling_friction_coeff * effective_r * normal_force_norm * norm(v_omega)) * dir;
      }
    else
      {
        const T mu_r_times_R_e = rolling_friction_coeff * effective_r;
        const vec3_t<T> omega_ij = p1.w - p2.w;
        const vec3_t<T> omega_perp = omega_ij - dot(omega_ij, n) * n;
        const vec3_t<T> delta_theta = dt * omega_perp;
        const T K_r = T(2.25) * normal_spring_constant * sqr(mu_r_times_R_e);
        cumulative = cumulative - K_r * delta_theta;
        const T M_r_max = mu_r_times_R_e * normal_force_norm;
        const T spring_norm = norm(cumulative);
        const T I_i = T(1.4) * p1.m * sqr(T(0.5) * p1.d);
        const T I_j = T(1.4) * p2.m * sqr(T(0.5) * p2.d);
        // I_i * I_j ~ 1e-35 for 0.1 mm grains: the product is taken in double in either precision
        const T I_e = T(double(I_i) * double(I_j) / (double(I_i) + double(I_j)));
        const T C_r = rolling_viscous_damping_coeff * T(2.) * R::sqrt(I_e * K_r);
        if (spring_norm > M_r_max)
          {
            cumulative = cumulative * (M_r_max / spring_norm);
            return cumulative - (T(mt.f_coefficient_epsd) * C_r) * omega_perp;
          }
        return cumulative - C_r * omega_perp;
      }
  }

  // Ferrari solution of the JKR contact-patch quartic (…contact_force.h:1356-1372;
  // the wall variant clamps root1 at 0, particle_wall_contact_force.h:~905)
  template <class T> __device__ __forceinline__ T jkr_contact_radius(T R, T overlap, T gamma, T Y, bool clamp_root1)
  {
    using Rt = real_traits<T>;
    const T c0 = sqr(R * overlap);
    const T c1 = T(-2.) * sqr(R) * T(M_PI) * gamma / Y;
    const T c2 = T(-2.) * overlap * R;
    const T P = -sqr(c2) / T(12.) - c0;
    const T Q = -cub(c2) / T(108.) + c0 * c2 / T(3.) - sqr(c1) * T(0.125);
    T root1 = clamp_root1 ? Rt::max(T(0.), (T(0.25) * sqr(Q)) + (cub(P) / T(27.))) : T(0.25) * sqr(Q) + cub(P) / T(27.);
    const T U = Rt::cbrt(T(-0.5) * Q + Rt::sqrt(root1));
    const T s = -c2 * T(5. / 6.) + U - P / (T(3.) * U);
    const T w = Rt::sqrt(Rt::max(T(1e-16), c2 + T(2.) * s));
    const T lambda = T(0.5) * c1 / w;
    const T root2 = Rt::max(T(1e-16), w * w - T(4.) * (c2 + s + lambda));
    return T(0.5) * (w + Rt::sqrt(root2));
  }

  // update_contact_information (…contact_force.h:223-298)
  template <class T>
  __device__ __forceinline__ void pp_update_contact_information(vec3_t<T> &tangential_displacement, vec3_t<T> &vt, T &vn, vec3_t<T> &n,
                                                                const ParticleView_t<T> &p1, const ParticleView_t<T> &p2, vec3_t<T> x2,
                                                                T distance, T dt)
  {
    // `distance` = sqrt(dist2(p1.x, x2)) from the caller is bit-identical to norm(x2 - p1.x):
    // the component differences are exact negatives, their squares and the component-ordered
    // sum are the same, so the second square root is not taken.
    const vec3_t<T> contact_vector = x2 - p1.x;
    n = contact_vector / distance;
    vec3_t<T> vrel = p1.v - p2.v;
    vrel = vrel + cross(T(0.5) * (p1.d * p1.w + p2.d * p2.w), n);
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
  template <int MODEL, int ROLLING, class T>
  __device__ __forceinline__ void pp_calculate_contact(const MaterialTables &mt, vec3_t<T> &tangential_displacement,
                                                       vec3_t<T> &rolling_spring_torque, vec3_t<T> vt, T vn, vec3_t<T> n, T overlap,
                                                       T dt, const ParticleView_t<T> &p1, const ParticleView_t<T> &p2, PairResult_t<T> &r,
                                                       const SelfPair_t<T> &self)
  {
    using R = real_traits<T>;
    const T d1 = p1.d, d2 = p2.d;
    T effective_radius, effective_mass;
    if (d1 == d2 && p1.m == p2.m)
      {
        // same operands, same expression: bit-identical to the general branch
        effective_radius = self.effective_radius;
        effective_mass = self.effective_mass;
      }
    else
      {
        effective_radius = (d1 * d2) / (T(2) * (d1 + d2));
        effective_mass = (p1.m * p2.m) / (p1.m + p2.m);
      }
    const int k = p1.type * mt.n_types + p2.type;
    const T Y = T(mt.Y[k]), G = T(mt.G[k]), beta = T(mt.beta[k]), mu = T(mt.mu[k]);
    const T roll_visc = T(mt.roll_visc[k]), roll_fric = T(mt.roll_fric[k]);

    if constexpr (MODEL == LETHE_PP_DMT)
      {
        constexpr T M_2PI = T(2. * M_PI);
        const T gamma = T(mt.gamma[k]), A = T(mt.hamaker[k]);
        const T F_po = M_2PI * effective_radius * gamma;
        const T delta_0 = -R::sqrt(A * effective_radius / (T(6.) * F_po));
        T cohesive_term;
        if (overlap > T(0.))
          {
            cohesive_term = -F_po;
            pp_calculate_contact<LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP, ROLLING>(mt, tangential_displacement, rolling_spring_torque,
                                                                                vt, vn, n, overlap, dt, p1, p2, r, self);
          }
        else if (overlap > delta_0)
          {
            cohesive_term = -F_po;
            tangential_displacement = v3t<T>(0, 0, 0);
            rolling_spring_torque = v3t<T>(0, 0, 0);
          }
        else
          {
            cohesive_term = -A * effective_radius / (T(6.) * sqr(overlap));
            tangential_displacement = v3t<T>(0, 0, 0);
            rolling_spring_torque = v3t<T>(0, 0, 0);
          }
        r.normal_force = r.normal_force + cohesive_term * n;
      }
    else if constexpr (MODEL == LETHE_PP_LINEAR)
      {
        const T kn = T(1.0667) * R::sqrt(effective_radius) * Y *
                     R::pow_0_2((T(0.9375) * effective_mass * T(1.0) * T(1.0) / (R::sqrt(effective_radius) * Y)));
        const T kt = kn * T(0.4);
        const T etan = T(-2) * beta * R::sqrt(effective_mass * kn);
        const T etat = etan * T(0.6324555320336759);
        const T normal_force_value = kn * overlap + etan * vn;
        r.normal_force = normal_force_value * n;
        const vec3_t<T> damping_tangential_force = etat * vt;
        r.tangential_force = (kt * tangential_displacement) + damping_tangential_force;
        const T coulomb_threshold = mu * normal_force_value;
        if (norm(r.tangential_force) > coulomb_threshold)
          {
            const vec3_t<T> limited = coulomb_threshold * (r.tangential_force / (norm(r.tangential_force) + R::tiny()));
            tangential_displacement = (limited - damping_tangential_force) / (kt + R::tiny());
            r.tangential_force = (kt * tangential_displacement) + damping_tangential_force;
          }
        r.torque_one = cross(n, r.tangential_force * d1 * T(0.5));
        r.torque_two = r.torque_one * d2 / d1;
        // the reference passes the two rolling coefficients in swapped order here (:841-851)
        r.rolling = pp_rolling<ROLLING, T>(mt, effective_radius, p1, p2, roll_visc, roll_fric, dt, kn, norm(r.normal_force), n,
                                           rolling_spring_torque);
      }
    else
      {
        const T radius_times_overlap_sqrt = R::sqrt(effective_radius * overlap);
        const T model_parameter_sn = T(2.0) * Y * radius_times_overlap_sqrt;
        const T model_parameter_st = T(8.0) * G * radius_times_overlap_sqrt;
        if constexpr (MODEL == LETHE_PP_HERTZ_JKR)
          {
            const T gamma = T(mt.gamma[k]);
            // the quartic's coefficients ((R delta)^2 ~ 1e-18, their cubes ~ 1e-54) leave float's range:
            // the contact radius is solved in double in either precision
            const T a = T(jkr_contact_radius<double>(double(effective_radius), double(overlap), double(gamma), double(Y), false));
            const T etan = T(-1.8257) * beta * R::sqrt(model_parameter_sn * effective_mass);
            const T kt = T(8.0) * radius_times_overlap_sqrt * G;
            const T etat = etan * R::sqrt(model_parameter_st / model_parameter_sn);
            const T normal_force_coefficient =
              T(4.) * cub(a) / (T(3.) * effective_radius) * Y - R::sqrt(T(8.) * T(M_PI) * gamma * Y * cub(a));
            r.normal_force = (normal_force_coefficient + etan * vn) * n;
            r.tangential_force = kt * tangential_displacement + etat * vt;
            const T two_pull_off_force = T(3.) * T(M_PI) * gamma * effective_radius;
            const T modified_coulomb_threshold = (normal_force_coefficient + two_pull_off_force) * mu;
            if (norm(r.tangential_force) > modified_coulomb_threshold)
              r.tangential_force = modified_coulomb_threshold * (r.tangential_force / (norm(r.tangential_force) + R::tiny()));
            r.torque_one = cross(n, r.tangential_force * d1 * T(0.5));
            r.torque_two = r.torque_one * d2 / d1;
            const T kn = T(0.66665) * model_parameter_sn;
            r.rolling = pp_rolling<ROLLING, T>(mt, effective_radius, p1, p2, roll_fric, roll_visc, dt, kn, norm(r.normal_force), n,
                                               rolling_spring_torque);
          }
        else
          {
            const T kn = T(0.66665) * model_parameter_sn;
            const T etan = T(-1.8257) * beta * R::sqrt(model_parameter_sn * effective_mass);
            const T kt = T(8.0) * G * radius_times_overlap_sqrt;
            const T normal_force_value = kn * overlap + etan * vn;
            r.normal_force = normal_force_value * n;
            const T coulomb_threshold = mu * normal_force_value;
            if constexpr (MODEL == LETHE_PP_HERTZ)
              {
                r.tangential_force = kt * tangential_displacement;
                if (norm(r.tangential_force) > coulomb_threshold)
                  r.tangential_force = coulomb_threshold * (r.tangential_force / (norm(r.tangential_force) + R::tiny()));
              }
            else
              {
                const T etat = etan * R::sqrt(model_parameter_st / model_parameter_sn);
                const vec3_t<T> damping_tangential_force = etat * vt;
                r.tangential_force = (kt * tangential_displacement) + damping_tangential_force;
                if (norm(r.tangential_force) > coulomb_threshold)
                  {
                    if constexpr (MODEL == LETHE_PP_HERTZ_MINDLIN_LIMIT_OVERLAP)
                      {
                        const vec3_t<T> limited = coulomb_threshold * (r.tangential_force / (norm(r.tangential_force) + R::tiny()));
                        tangential_displacement = (limited - damping_tangential_force) / (kt + R::tiny());
                        r.tangential_force = (kt * tangential_displacement) + damping_tangential_force;
                      }
                    else
                      r.tangential_force = coulomb_threshold * (r.tangential_force / (norm(r.tangential_force) + R::tiny()));
                  }
              }
            r.torque_one = cross(n, r.tangential_force * d1 * T(0.5));
            r.torque_two = r.torque_one * d2 / d1;
            r.rolling = pp_rolling<ROLLING, T>(mt, effective_radius, p1, p2, roll_fric, roll_visc, dt, kn, norm(r.normal_force), n,
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
        const double a = jkr_contact_radius<double>(R, overlap, gamma, Y, true);
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
