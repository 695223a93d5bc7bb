// dem_math.cuh — the two libm functions of the hot path, restated so that the device computes the
// bits the reference computes with glibc on the host.
//
// The reference calls std::cbrt in the JKR contact-patch quartic (particle_particle_contact_force.h:
// 1356-1372, particle_wall_contact_force.h:~905) and std::pow(x, 0.2) in the linear model's spring
// constant (…force.h:723-760, particle_wall_contact_force.h:610-640). CUDA's cbrt / pow differ from
// glibc's in the last bit on some arguments, and the tensile tail of a damped contact amplifies
// that bit (SURVEY.md §8c rule 2). So:
//   * glibc_cbrt follows glibc's published algorithm for binary64 (sysdeps/ieee754/dbl-64/s_cbrt.c,
//     third-party, not part of /root/reference: frexp, a degree-6 polynomial start value, one
//     Halley step, the 2^(k/3) factor table, ldexp), operation for operation, without FMA
//     contraction — glibc's result is not the correctly rounded cube root, so matching it means
//     restating it;
//   * pow_0_2_cr is the correctly rounded x^c for c = the double nearest to 0.2 (= 1/5 + 1.11e-17):
//     glibc's pow is correctly rounded in all but astronomically rare cases (< 0.52 ulp), so the
//     correctly rounded value is what it returns. Start value from the device pow, one Newton
//     correction towards the fifth root with the residual x - y^5 evaluated in double-double
//     arithmetic (error-free products through fma), then the factor x^(c - 1/5) = 1 + (c - 1/5) ln x.
// tests/test_device_math.py compiles this header for the host and checks both against the libm
// of the machine bit for bit on a few million arguments of the ranges the models produce.
#pragma once

#include <cmath>

#if defined(__CUDACC__)
#define DEM_HD __host__ __device__ __forceinline__
#else
#define DEM_HD inline
#endif

namespace dem
{
  DEM_HD double glibc_cbrt(double x)
  {
    const double CBRT2 = 1.2599210498948731648, SQR_CBRT2 = 1.5874010519681994748;
    const double factor[5] = {1.0 / SQR_CBRT2, 1.0 / CBRT2, 1.0, CBRT2, SQR_CBRT2};
    int xe;
    const double xm = frexp(fabs(x), &xe);
    // zero, infinity, NaN
    if (xe == 0 && !(fabs(x) > 0.0 && fabs(x) < INFINITY))
      return x + x;
    const double u =
      (0.354895765043919860 +
       ((1.50819193781584896 -
         ((2.11499494167371287 -
           ((2.44693122563534430 - ((1.83469277483613086 - (0.784932344976639262 - 0.145263899385486377 * xm) * xm) * xm)) * xm)) *
          xm)) *
        xm));
    const double t2 = u * u * u;
    const double ym = u * (t2 + 2.0 * xm) / (2.0 * t2 + xm) * factor[2 + xe % 3];
    return ldexp(x > 0.0 ? ym : -ym, xe / 3);
  }

  // error-free product a * b = hi + lo
  DEM_HD void two_prod(double a, double b, double &hi, double &lo)
  {
    hi = a * b;
    lo = fma(a, b, -hi);
  }

  // correctly rounded pow(x, 0.2) for finite x > 0 (otherwise: what pow returns)
  DEM_HD double pow_0_2_cr(double x)
  {
    if (!(x > 0.0) || !(x < INFINITY))
      return pow(x, 0.2);
    const double y = pow(x, 0.2); // within a few ulp on either side
    // y^2, y^4, y^5 as double-double
    double p2h, p2l;
    two_prod(y, y, p2h, p2l);
    double p4h, p4l;
    two_prod(p2h, p2h, p4h, p4l);
    p4l = p4l + 2.0 * (p2h * p2l); // (p2h + p2l)^2, dropping p2l^2
    double p5h, p5l;
    two_prod(p4h, y, p5h, p5l);
    p5l = p5l + p4l * y;
    // residual x - y^5 (x - p5h is exact: the two are within a few ulp of each other)
    const double r = (x - p5h) - p5l;
    const double delta = r / (5.0 * p4h);
    // 0.2 as a double exceeds 1/5 by 2^-56 * 0.8: x^0.2 = x^(1/5) * (1 + 1.11e-17 ln x)
    const double delta_c = y * (1.1102230246251566e-17 * log(x));
    return y + (delta + delta_c);
  }
} // namespace dem
