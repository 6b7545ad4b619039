// b2_bdp.cuh -- Bussi-Donadio-Parrinello stochastic velocity rescaling, evaluated ON THE DEVICE.
//
// The reference draws the noise on the host (std::mt19937 + std::uniform_real_distribution) after
// a blocking device->host copy of the temperature every step (ensemble_bdp.cu:69-101,
// svr_utilities.cuh:27-135).  Here one device thread owns the same generator state and runs the
// same published algorithm (G. Bussi et al., J. Chem. Phys. 126, 014101 (2007): resamplekin), so
// the step stays asynchronous AND reproduces the reference's -DDEBUG random stream (seed
// 12345678, ensemble_bdp.cu:31-32) draw for draw.  Bodies are host-compilable for tests/emu.
#pragma once
#include "b2_common.cuh"
#include <math.h>
#include <stdint.h>

// MT19937 (Matsumoto & Nishimura 1998) = std::mt19937, plus the state of the Box-Muller pair cache
struct B2BdpState {
  uint32_t mt[624];
  int idx;
  int iset;     // gasdev's cached second deviate (a function-local static in the reference)
  double gset;
  double factor; // last velocity scale factor (read by k_scale_by)
};

B2_HD void b2_mt_seed(B2BdpState& s, uint32_t seed)
{
  s.mt[0] = seed;
  for (int i = 1; i < 624; ++i)
    s.mt[i] = 1812433253u * (s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) + (uint32_t)i;
  s.idx = 624;
  s.iset = 0;
  s.gset = 0.0;
  s.factor = 1.0;
}

B2_HD uint32_t b2_mt_next(B2BdpState& s)
{
  if (s.idx >= 624) {
    for (int k = 0; k < 624; ++k) {
      const uint32_t y = (s.mt[k] & 0x80000000u) | (s.mt[(k + 1) % 624] & 0x7fffffffu);
      s.mt[k] = s.mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s.idx = 0;
  }
  uint32_t y = s.mt[s.idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// std::uniform_real_distribution<double>(0,1)(rng) with libstdc++'s generate_canonical<double,53>:
// two 32-bit draws, low word first, summed in double, divided by 2^64
B2_HD double b2_rand01(B2BdpState& s)
{
  const double lo = (double)b2_mt_next(s);
  const double hi = (double)b2_mt_next(s);
  double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
  if (r >= 1.0)
    r = 0.99999999999999988897769753748; // nextafter(1, 0)
  return r;
}

// Gaussian deviate, polar Box-Muller with one cached value (svr_utilities.cuh:27-50)
B2_HD double b2_gasdev(B2BdpState& s)
{
  if (s.iset == 0) {
    double v1, v2, rsq;
    do {
      v1 = 2.0 * b2_rand01(s) - 1.0;
      v2 = 2.0 * b2_rand01(s) - 1.0;
      rsq = v1 * v1 + v2 * v2;
    } while (rsq >= 1.0 || rsq == 0.0);
    const double fac = sqrt(-2.0 * log(rsq) / rsq);
    s.gset = v1 * fac;
    s.iset = 1;
    return v2 * fac;
  }
  s.iset = 0;
  return s.gset;
}

// Gamma deviate of integer order (svr_utilities.cuh:52-84)
B2_HD double b2_gamdev(B2BdpState& s, int ia)
{
  double x;
  if (ia < 6) {
    x = 1.0;
    for (int j = 1; j <= ia; ++j)
      x *= b2_rand01(s);
    return -log(x);
  }
  double e, y, am, sq;
  do {
    do {
      double v1, v2;
      do {
        v1 = b2_rand01(s);
        v2 = 2.0 * b2_rand01(s) - 1.0;
      } while (v1 * v1 + v2 * v2 > 1.0);
      y = v2 / v1;
      am = ia - 1;
      sq = sqrt(2.0 * am + 1.0);
      x = sq * y + am;
    } while (x <= 0.0);
    e = (1.0 + y * y) * exp(am * log(x / am) - sq * y);
  } while (b2_rand01(s) > e);
  return x;
}

// sum of nn squared Gaussian deviates (svr_utilities.cuh:86-104)
B2_HD double b2_sumnoises(B2BdpState& s, int nn)
{
  if (nn == 0)
    return 0.0;
  if (nn == 1) {
    const double rr = b2_gasdev(s);
    return rr * rr;
  }
  if (nn % 2 == 0)
    return 2.0 * b2_gamdev(s, nn / 2);
  const double rr = b2_gasdev(s);
  return 2.0 * b2_gamdev(s, (nn - 1) / 2) + rr * rr;
}

// new kinetic energy (svr_utilities.cuh:106-135)
B2_HD double b2_resamplekin(B2BdpState& s, double kk, double sigma, int ndeg, double taut)
{
  const double factor = taut > 0.1 ? exp(-1.0 / taut) : 0.0;
  const double rr = b2_gasdev(s);
  const double noise = b2_sumnoises(s, ndeg - 1);
  return kk + (1.0 - factor) * (sigma * (noise + rr * rr) / ndeg - kk) +
         2.0 * rr * sqrt(kk * sigma / ndeg * (1.0 - factor) * factor);
}

// velocity scale factor of one NVT step from the instantaneous temperature
// (Ensemble_BDP::integrate_nvt_bdp_2, ensemble_bdp.cu:91-100)
B2_HD double b2_bdp_factor(
  B2BdpState& s, double t_instant, int ndeg, double temperature, double temperature_coupling)
{
  const double ek = t_instant * ndeg * 8.617343e-5 * 0.5;
  const double sigma = ndeg * 8.617343e-5 * temperature * 0.5;
  const double ek_new = b2_resamplekin(s, ek, sigma, ndeg, temperature_coupling);
  return sqrt(ek_new / ek);
}
