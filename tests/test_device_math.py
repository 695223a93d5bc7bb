"""lethe_b200/csrc/dem_math.cuh — the device's restatement of the two libm functions on the hot
path (std::cbrt in the JKR quartic, std::pow(x, 0.2) in the linear model) — compiled for the host
and compared with the libm of this machine bit for bit.

glibc_cbrt restates glibc's algorithm and must agree on every argument. pow_0_2_cr is the correctly
rounded power; glibc's pow is within 0.52 ulp, i.e. correctly rounded except when the exact value
sits within ~5e-4 ulp of a rounding boundary: the two may differ by one ulp on < 0.5 % of the
arguments and never by more."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r"""
#include "%s/lethe_b200/csrc/dem_math.cuh"
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <random>
int main()
{
  std::mt19937_64 g(1);
  long bad_c = 0, bad_p = 0, far_p = 0;
  const long N = 2000000;
  for (long k = 0; k < N; ++k)
    {
      // magnitudes from 1e-30 to 1e+5: overlaps^2 ... stiffness ratios
      const double e = std::uniform_real_distribution<double>(-30, 5)(g);
      double x = std::exp(e * std::log(10.0)) * std::uniform_real_distribution<double>(0.5, 2)(g);
      if (k & 1)
        x = -x;
      const double a = dem::glibc_cbrt(x), b = std::cbrt(x);
      if (std::memcmp(&a, &b, 8))
        ++bad_c;
      const double xa = std::fabs(x);
      const double p = dem::pow_0_2_cr(xa), q = std::pow(xa, 0.2);
      if (std::memcmp(&p, &q, 8))
        {
          ++bad_p;
          int64_t ip, iq;
          std::memcpy(&ip, &p, 8);
          std::memcpy(&iq, &q, 8);
          if (ip - iq > 1 || iq - ip > 1)
            ++far_p;
        }
    }
  const double special[] = {0.0, -0.0, -8.0, 27.0, 1.0, 4.9e-324, 1.7e308};
  for (double z : special)
    {
      const double a = dem::glibc_cbrt(z), b = std::cbrt(z);
      if (std::memcmp(&a, &b, 8))
        ++bad_c;
    }
  printf("%%ld %%ld %%ld %%ld\n", N, bad_c, bad_p, far_p);
  return 0;
}
""" % ROOT


def test_device_cbrt_and_pow_match_libm():
    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    src, exe = os.path.join(build, "device_math.cpp"), os.path.join(build, "device_math")
    with open(src, "w") as f:
        f.write(SRC)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", exe, src])
    n, bad_cbrt, bad_pow, far_pow = map(int, subprocess.check_output([exe], text=True).split())
    assert bad_cbrt == 0, f"glibc_cbrt differs from std::cbrt on {bad_cbrt} of {n} arguments"
    assert far_pow == 0 and bad_pow < 0.005 * n, (bad_pow, far_pow, n)
