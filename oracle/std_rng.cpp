// oracle/std_rng.cpp -- TEST INFRASTRUCTURE ONLY.
// The C++ standard library's own std::mt19937 + std::uniform_real_distribution<double>(0,1), i.e.
// exactly the generator objects the reference's BDP thermostat uses on the host
// (src/integrate/ensemble_bdp.cu:29-36, svr_utilities.cuh:29,54).  tests/ use it to pin the
// restatements of that stream (oracle_py.BdpOracle, gpumd_b200/csrc/b2_bdp.cuh).
#include <random>

extern "C" void stdrng_uniform01(unsigned seed, int n, double* out)
{
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> rand1(0, 1);
  for (int i = 0; i < n; ++i)
    out[i] = rand1(rng);
}

extern "C" void stdrng_raw(unsigned seed, int n, unsigned* out)
{
  std::mt19937 rng(seed);
  for (int i = 0; i < n; ++i)
    out[i] = (unsigned)rng();
}
