// b2_measure.cu -- heat-current autocorrelation (Green-Kubo) on top of the per-atom virial the force
// kernels produce: the consumer SURVEY.md 8f ranks fourth.
//
//   b200md_hac_create / _sample / _finish  <- HAC::preprocess / process / postprocess,
//       src/measure/hac.cu:32-280 (gpu_sum_heat :50-78, gpu_find_hac :111-170, find_rtc :173-181)
//
// sample(): every sample_interval steps the per-atom heat current (compute_heat.cu:32-90, our
// b200md_compute_heat) is summed into heat_all[nd + Nd*k], k = jx_in jx_out jy_in jy_out jz.
// finish(): hac[nc + Nc*k] = < J_k(0) (J_k(t) + J_partner(t)) > for the in/out pairs, < Jz(0) Jz(t) >
// for z (hac.cu:132-142), averaged over the Nd - nc available origins, and the running thermal
// conductivity by the trapezoidal rule with factor dt/2 / (k_B T^2 V) * KAPPA_UNIT_CONVERSION.
// The reductions are one block per component with a fixed summation order (deterministic).
#include "../../include/b200md.h"
#include "b2_host.h"
#include <new>
#include <vector>

namespace b2 {
namespace {

// one block per heat component: strided partial sums, then a tree -- the order of hac.cu:50-78
__global__ void __launch_bounds__(1024) k_sum_heat(int n, int heat_stride, int Nd, int nd, const double* g_heat, double* g_heat_all)
{
  __shared__ double s[1024];
  const int tid = threadIdx.x;
  double v = 0.0;
  for (int i = tid; i < n; i += 1024)
    v += g_heat[i + (size_t)heat_stride * blockIdx.x];
  s[tid] = v;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (tid < o)
      s[tid] += s[tid + o];
    __syncthreads();
  }
  if (tid == 0)
    g_heat_all[nd + (size_t)Nd * blockIdx.x] = s[0];
}

// block = correlation time nc; hac.cu:111-170
__global__ void __launch_bounds__(128) k_find_hac(int Nc, int Nd, const double* __restrict__ h, double* g_hac)
{
  __shared__ double s[5][128];
  const int tid = threadIdx.x, bid = blockIdx.x;
  double a[5] = {0, 0, 0, 0, 0};
  for (int i = tid; i + bid < Nd; i += 128) {
    const int j = i + bid;
    a[0] += h[i] * h[j] + h[i] * h[j + Nd];
    a[1] += h[i + Nd] * h[j + Nd] + h[i + Nd] * h[j];
    a[2] += h[i + 2 * Nd] * h[j + 2 * Nd] + h[i + 2 * Nd] * h[j + 3 * Nd];
    a[3] += h[i + 3 * Nd] * h[j + 3 * Nd] + h[i + 3 * Nd] * h[j + 2 * Nd];
    a[4] += h[i + 4 * Nd] * h[j + 4 * Nd];
  }
  for (int k = 0; k < 5; ++k)
    s[k][tid] = a[k];
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o)
      for (int k = 0; k < 5; ++k)
        s[k][tid] += s[k][tid + o];
    __syncthreads();
  }
  if (tid == 0)
    for (int k = 0; k < 5; ++k)
      g_hac[bid + (size_t)Nc * k] = s[k][0] / (Nd - bid);
}

} // namespace
} // namespace b2

using namespace b2;

struct b200md_hac {
  int number_of_steps = 0, sample_interval = 1, Nc = 0, Nd = 0;
  DevBuf<double> heat_all, heat, hac;
};

extern "C" {

int b200md_hac_create(int number_of_steps, int sample_interval, int Nc, b200md_hac** out)
{
  if (!out || number_of_steps <= 0 || sample_interval <= 0 || Nc <= 0 ||
      Nc > number_of_steps / sample_interval) {
    set_error("b200md_hac_create: need number_of_steps / sample_interval >= Nc > 0");
    return B200MD_ERR_ARG;
  }
  b200md_hac* p = new (std::nothrow) b200md_hac;
  if (!p) {
    set_error("out of host memory");
    return B200MD_ERR_ARG;
  }
  p->number_of_steps = number_of_steps;
  p->sample_interval = sample_interval;
  p->Nc = Nc;
  p->Nd = number_of_steps / sample_interval;
  if (p->heat_all.reserve((size_t)5 * p->Nd) != cudaSuccess || p->hac.reserve((size_t)5 * Nc) != cudaSuccess) {
    delete p;
    set_error("out of device memory (HAC)");
    return B200MD_ERR_CUDA;
  }
  cudaMemset(p->heat_all.p, 0, sizeof(double) * 5 * p->Nd);
  *out = p;
  return B200MD_OK;
}

void b200md_hac_destroy(b200md_hac* p) { delete p; }

int b200md_hac_sample(
  b200md_hac* p, int step, int n, int stride, const double* d_virial, const double* d_velocity,
  void* stream)
{
  if ((step + 1) % p->sample_interval != 0)
    return B200MD_OK;
  const int nd = (step + 1) / p->sample_interval - 1;
  if (nd < 0 || nd >= p->Nd)
    return B200MD_OK; // beyond the run length given at creation
  B2_CUDA(p->heat.reserve((size_t)5 * n));
  const int rc = b200md_compute_heat(n, stride, d_virial, d_velocity, p->heat.p, n, stream);
  if (rc != B200MD_OK)
    return rc;
  k_sum_heat<<<5, 1024, 0, (cudaStream_t)stream>>>(n, n, p->Nd, nd, p->heat.p, p->heat_all.p);
  B2_LAUNCHED();
  return B200MD_OK;
}

int b200md_hac_finish(
  b200md_hac* p, double time_step, double temperature, double volume, double* hac_out,
  double* rtc_out, void* stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  const int Nc = p->Nc;
  k_find_hac<<<Nc, 128, 0, st>>>(Nc, p->Nd, p->heat_all.p, p->hac.p);
  B2_LAUNCHED();
  std::vector<double> hac((size_t)5 * Nc);
  B2_CUDA(cudaMemcpyAsync(hac.data(), p->hac.p, sizeof(double) * 5 * Nc, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  const double K_B = 8.617343e-5, KAPPA_UNIT_CONVERSION = 1.573769e+5; // common.cuh:21,27
  const double dt = time_step * p->sample_interval;
  const double factor = dt * 0.5 / (K_B * temperature * temperature * volume) * KAPPA_UNIT_CONVERSION;
  for (int k = 0; k < 5; ++k) {
    double run = 0.0;
    for (int nc = 0; nc < Nc; ++nc) {
      const size_t idx = (size_t)Nc * k + nc;
      if (nc > 0)
        run += (hac[idx - 1] + hac[idx]) * factor; // find_rtc, hac.cu:173-181
      if (rtc_out)
        rtc_out[idx] = run;
      if (hac_out)
        hac_out[idx] = hac[idx];
    }
  }
  return B200MD_OK;
}

/* the recorded heat-current series heat_all[nd + Nd*k] (host copy; parity hook) */
int b200md_hac_series(b200md_hac* p, double* out, void* stream)
{
  B2_CUDA(cudaMemcpyAsync(out, p->heat_all.p, sizeof(double) * 5 * p->Nd, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  B2_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return B200MD_OK;
}

} // extern "C"
