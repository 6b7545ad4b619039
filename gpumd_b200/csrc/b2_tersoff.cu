// b2_tersoff.cu -- Tersoff-1989 potential (FP64) and the per-atom heat current of libb200md;
// C-ABI entry points b200md_tersoff_*, b200md_compute_heat.
#include "../../include/b200md.h"
#include "b2_host.h"
#include "b2_neighbor_host.h"
#include "b2_nep.cuh" // b2_body_unpack
#include "b2_tersoff.cuh"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace b2 {
namespace {

__global__ void __launch_bounds__(64) k_tersoff_partial(B2TersoffView P, B2Box box)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.n)
    b2_body_tersoff_partial(i, P, box);
}

__global__ void __launch_bounds__(128) k_tersoff_reduce(B2TersoffView P, B2Box box)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.n)
    b2_body_tersoff_reduce(i, P, box);
}

__global__ void __launch_bounds__(256) k_unpack_t(
  int n, const int* __restrict__ perm, const double* __restrict__ acc, double* pe, double* force,
  double* virial)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    b2_body_unpack(i, n, perm, acc, pe, force, virial);
}

__global__ void __launch_bounds__(256) k_heat(
  int n, int stride, const double* __restrict__ w, const double* __restrict__ v, double* heat,
  int hstride)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    b2_body_heat(i, stride, w, v, heat, hstride);
}

void fill(B2TersoffPara& t, const double* p)
{
  t.a = p[0];
  t.b = p[1];
  t.lambda = p[2];
  t.mu = p[3];
  t.beta = p[4];
  t.n = p[5];
  t.c2 = p[6] * p[6];
  t.d2 = p[7] * p[7];
  t.h = p[8];
  t.r1 = p[9];
  t.r2 = p[10];
  t.one_plus = 1.0 + t.c2 / t.d2;
  t.pi_factor = 3.14159265358979 / (t.r2 - t.r1); // PI, src/utilities/common.cuh:19
  t.mhn = -0.5 / p[5];
}

} // namespace
} // namespace b2

using namespace b2;

struct b200md_tersoff {
  int nt = 0, n = 0;
  double rc = 0.0;
  std::vector<std::string> symbols;
  Neighbor nb;
  DevBuf<int> nn, nl;
  DevBuf<double> f12, acc;
  B2TersoffView view;
};

#define B2_TRY(expr)        \
  do {                      \
    const int rc_ = (expr); \
    if (rc_ != B200MD_OK)   \
      return rc_;           \
  } while (0)

extern "C" {

int b200md_tersoff_create(const char* path, int num_atoms, b200md_tersoff** out)
{
  if (!path || !out || num_atoms <= 0) {
    set_error("b200md_tersoff_create: bad argument");
    return B200MD_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("b200md_tersoff_create: no CUDA device (libb200md has no CPU fallback)");
    return B200MD_ERR_CUDA;
  }
  FILE* fid = fopen(path, "r");
  if (!fid) {
    set_error(std::string("Failed to open ") + path);
    return B200MD_ERR_IO;
  }
  char name[64];
  int nt = 0;
  if (fscanf(fid, "%63s%d", name, &nt) != 2 || std::strcmp(name, "tersoff_1989") != 0 || nt < 1 ||
      nt > 2) {
    fclose(fid);
    set_error("tersoff_1989 potential file: first line must be 'tersoff_1989 <1|2> <symbols>'");
    return B200MD_ERR_ARG;
  }
  b200md_tersoff* p = new (std::nothrow) b200md_tersoff;
  if (!p) {
    fclose(fid);
    set_error("out of host memory");
    return B200MD_ERR_ARG;
  }
  p->nt = nt;
  for (int k = 0; k < nt; ++k) {
    if (fscanf(fid, "%63s", name) != 1) {
      fclose(fid);
      delete p;
      set_error("Reading error for Tersoff-1989 potential.");
      return B200MD_ERR_ARG;
    }
    p->symbols.push_back(name);
  }
  double para[23] = {0};
  const int want = nt == 1 ? 11 : 23;
  for (int k = 0; k < want; ++k) {
    if (fscanf(fid, "%lf", &para[k]) != 1) {
      fclose(fid);
      delete p;
      set_error("Reading error for Tersoff-1989 potential.");
      return B200MD_ERR_ARG;
    }
  }
  fclose(fid);
  B2TersoffView& P = p->view;
  std::memset(&P, 0, sizeof P);
  fill(P.p[0], para);
  p->rc = P.p[0].r2;
  if (nt == 2) { // mixing rules, tersoff1989.cu:100-130
    fill(P.p[1], para + 11);
    const double chi = para[22];
    B2TersoffPara& m = P.p[2];
    m.a = std::sqrt(P.p[0].a * P.p[1].a);
    m.b = std::sqrt(P.p[0].b * P.p[1].b) * chi;
    m.lambda = 0.5 * (P.p[0].lambda + P.p[1].lambda);
    m.mu = 0.5 * (P.p[0].mu + P.p[1].mu);
    m.r1 = std::sqrt(P.p[0].r1 * P.p[1].r1);
    m.r2 = std::sqrt(P.p[0].r2 * P.p[1].r2);
    m.pi_factor = 3.14159265358979 / (m.r2 - m.r1);
    p->rc = P.p[0].r2 > P.p[1].r2 ? P.p[0].r2 : P.p[1].r2;
  } else {
    P.p[1] = P.p[0];
    P.p[2] = P.p[0];
  }
  P.rc2 = (float)(p->rc * p->rc);
  p->n = num_atoms;
  const size_t N = (size_t)num_atoms;
  // neighbor.initialize(rc, num_atoms, 50), tersoff1989.cu:149
  const double rs = p->rc + 1.0;
  const int mn_skin = (int)(50 * rs * rs * rs / (p->rc * p->rc * p->rc));
  int rc = p->nb.init(num_atoms, p->rc, mn_skin);
  if (rc != B200MD_OK) {
    delete p;
    return rc;
  }
  if (p->nn.reserve(N) != cudaSuccess || p->nl.reserve(N * B2_TERSOFF_MAXL) != cudaSuccess ||
      p->f12.reserve(3 * N * B2_TERSOFF_MAXL) != cudaSuccess ||
      p->acc.reserve(13 * N) != cudaSuccess) {
    delete p;
    set_error("out of device memory");
    return B200MD_ERR_CUDA;
  }
  P.n = num_atoms;
  P.atoms = p->nb.atoms.p;
  P.nn_skin = p->nb.nn_skin.p;
  P.nl_skin = p->nb.nl_skin.p;
  P.nn = p->nn.p;
  P.nl = p->nl.p;
  P.f12 = p->f12.p;
  P.acc = p->acc.p;
  P.flags = p->nb.flags.p;
  cudaDeviceSynchronize();
  *out = p;
  return B200MD_OK;
}

void b200md_tersoff_destroy(b200md_tersoff* p) { delete p; }
double b200md_tersoff_rc(const b200md_tersoff* p) { return p->rc; }
const char* b200md_tersoff_symbol(const b200md_tersoff* p, int t)
{
  return (t >= 0 && t < p->nt) ? p->symbols[t].c_str() : "";
}
int b200md_tersoff_info(const b200md_tersoff* p, int what)
{
  if (what == 0)
    return p->nt;
  if (what == 6) {
    int bits = 0, rebuilds = 0;
    const_cast<b200md_tersoff*>(p)->nb.check(0, &bits, &rebuilds);
    return rebuilds;
  }
  return -1;
}

int b200md_tersoff_compute(
  b200md_tersoff* p, int n, const double h[9], const int pbc[3], const int* d_type,
  const double* d_position, double* d_potential, double* d_force, double* d_virial, void* stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  const B2Box box = make_box(h, pbc);
  B2_TRY(p->nb.update(box, d_type, d_position, n, st));
  p->n = n;
  p->view.n = n;
  k_tersoff_partial<<<grid_for(n, 64), 64, 0, st>>>(p->view, box);
  B2_LAUNCHED();
  k_tersoff_reduce<<<grid_for(n, 128), 128, 0, st>>>(p->view, box);
  B2_LAUNCHED();
  k_unpack_t<<<grid_for(n, 256), 256, 0, st>>>(
    n, p->nb.perm.p, p->acc.p, d_potential, d_force, d_virial);
  B2_LAUNCHED();
  return B200MD_OK;
}

int b200md_tersoff_invalidate(b200md_tersoff* p, int n_new, void* stream)
{
  B2_TRY(p->nb.invalidate(n_new, (cudaStream_t)stream));
  p->n = n_new;
  p->view.n = n_new;
  return B200MD_OK;
}

int b200md_tersoff_check(b200md_tersoff* p, void* stream)
{
  int bits = 0, rebuilds = 0;
  B2_TRY(p->nb.check((cudaStream_t)stream, &bits, &rebuilds));
  if (bits) {
    set_error("Tersoff neighbour capacity exceeded on the device (skin list or 32 local neighbours)");
    return B200MD_ERR_OVERFLOW;
  }
  return B200MD_OK;
}

int b200md_compute_heat(
  int n, int stride, const double* d_virial, const double* d_velocity, double* d_heat,
  int heat_stride, void* stream)
{
  k_heat<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
    n, stride, d_virial, d_velocity, d_heat, heat_stride);
  B2_LAUNCHED();
  return B200MD_OK;
}

} // extern "C"
