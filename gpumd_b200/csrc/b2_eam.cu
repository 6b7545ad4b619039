// b2_eam.cu -- analytic EAM potentials (Zhou-2004 alloy form, Dai-2006) of libb200md;
// C-ABI entry points b200md_eam_*.
#include "../../include/b200md.h"
#include "b2_eam.cuh"
#include "b2_host.h"
#include "b2_neighbor_host.h"
#include "b2_nep.cuh" // b2_body_unpack
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace b2 {
namespace {

__global__ void __launch_bounds__(128) k_eam_density(B2EamView P, B2Box box)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.n)
    b2_body_eam_density(i, P, box);
}

__global__ void __launch_bounds__(128) k_eam_force(B2EamView P, B2Box box)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.n)
    b2_body_eam_force(i, P, box);
}

__global__ void __launch_bounds__(256) k_unpack_e(
  int n, const int* __restrict__ perm, const double* __restrict__ acc, double* pe, double* force,
  double* virial)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    b2_body_unpack(i, n, perm, acc, pe, force, virial);
}

} // namespace
} // namespace b2

using namespace b2;

struct b200md_eam {
  int nt = 0, n = 0, model = 0;
  double rc = 0.0;
  std::vector<std::string> symbols;
  Neighbor nb;
  DevBuf<float> zp, Fp;
  DevBuf<double> acc;
  B2EamView view;
};

#define B2_TRY(expr)        \
  do {                      \
    const int rc_ = (expr); \
    if (rc_ != B200MD_OK)   \
      return rc_;           \
  } while (0)

extern "C" {

int b200md_eam_create(const char* path, int num_atoms, b200md_eam** out)
{
  if (!path || !out || num_atoms <= 0) {
    set_error("b200md_eam_create: bad argument");
    return B200MD_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("b200md_eam_create: no CUDA device (libb200md has no CPU fallback)");
    return B200MD_ERR_CUDA;
  }
  FILE* fid = fopen(path, "r");
  if (!fid) {
    set_error(std::string("Failed to open ") + path);
    return B200MD_ERR_IO;
  }
  char name[64];
  int nt = 0;
  if (fscanf(fid, "%63s%d", name, &nt) != 2) {
    fclose(fid);
    set_error("Reading error for EAM potential.");
    return B200MD_ERR_ARG;
  }
  int model = -1;
  if (std::strcmp(name, "eam_zhou_2004") == 0)
    model = 0;
  else if (std::strcmp(name, "eam_dai_2006") == 0)
    model = 1;
  if (model < 0 || nt < 1 || nt > 18 || (model == 1 && nt != 1)) {
    fclose(fid);
    set_error("EAM potential file: expected 'eam_zhou_2004 <1..18> ...' or 'eam_dai_2006 1 ...'");
    return B200MD_ERR_ARG;
  }
  b200md_eam* p = new (std::nothrow) b200md_eam;
  if (!p) {
    fclose(fid);
    set_error("out of host memory");
    return B200MD_ERR_ARG;
  }
  p->nt = nt;
  p->model = model;
  for (int k = 0; k < nt; ++k) {
    if (fscanf(fid, "%63s", name) != 1) {
      fclose(fid);
      delete p;
      set_error("Reading error for EAM potential.");
      return B200MD_ERR_ARG;
    }
    p->symbols.push_back(name);
  }
  B2EamView& P = p->view;
  std::memset(&P, 0, sizeof P);
  std::vector<float> zp((size_t)nt * EZ_COUNT, 0.0f);
  float rc = 0.0f;
  bool ok = true;
  if (model == 0) { // eam.cu:46-92
    for (int t = 0; t < nt && ok; ++t) {
      float x[21];
      for (int k = 0; k < 21; ++k)
        ok = ok && fscanf(fid, "%f", &x[k]) == 1;
      float* e = &zp[(size_t)t * EZ_COUNT];
      e[EZ_RE_INV] = 1.0f / x[0];
      e[EZ_FE] = x[1];
      e[EZ_RHO_E_INV] = 1.0f / x[2];
      e[EZ_RHO_S_INV] = 1.0f / x[3];
      e[EZ_ALPHA] = x[4];
      e[EZ_BETA] = x[5];
      e[EZ_A] = x[6];
      e[EZ_B] = x[7];
      e[EZ_KAPPA] = x[8];
      e[EZ_LAMBDA] = x[9];
      e[EZ_FN0] = x[10];
      e[EZ_FN1] = x[11];
      e[EZ_FN2] = x[12];
      e[EZ_FN3] = x[13];
      e[EZ_F0] = x[14];
      e[EZ_F1] = x[15];
      e[EZ_F2] = x[16];
      e[EZ_F3] = x[17];
      e[EZ_ETA] = x[18];
      e[EZ_FE_EMBED] = x[19];
      e[EZ_RC] = x[20];
      e[EZ_RHO_N] = x[2] * 0.85;
      e[EZ_RHO_0] = x[2] * 1.15;
      e[EZ_RHO_N_INV] = 1.0f / e[EZ_RHO_N];
      if (rc < x[20])
        rc = x[20];
    }
  } else { // eam.cu:94-122
    float x[9];
    for (int k = 0; k < 9; ++k)
      ok = ok && fscanf(fid, "%f", &x[k]) == 1;
    P.dA = x[0];
    P.dd = x[1];
    P.dc = x[2];
    P.dc0 = x[3];
    P.dc1 = x[4];
    P.dc2 = x[5];
    P.dc3 = x[6];
    P.dc4 = x[7];
    P.dB = x[8];
    rc = P.dc > P.dd ? P.dc : P.dd;
  }
  fclose(fid);
  if (!ok) {
    delete p;
    set_error("Reading error for EAM potential.");
    return B200MD_ERR_ARG;
  }
  p->rc = rc;
  p->n = num_atoms;
  const size_t N = (size_t)num_atoms;
  // neighbor.initialize(rc, number_of_atoms, 400), eam.cu:43 ("very safe for EAM"); bounded here by
  // what a sphere of rc+skin can hold at twice fcc-metal density
  const double rs = rc + 1.0;
  int mn = (int)(400 * rs * rs * rs / ((double)rc * rc * rc));
  const int mn_dense = (int)(4.19 * rs * rs * rs * 0.2) + 32;
  if (b2_tight_lists() && mn > mn_dense)
    mn = mn_dense;
  int r = p->nb.init(num_atoms, rc, mn);
  if (r != B200MD_OK) {
    delete p;
    return r;
  }
  if (p->zp.reserve(zp.size()) != cudaSuccess || p->Fp.reserve(N) != cudaSuccess ||
      p->acc.reserve(13 * N) != cudaSuccess) {
    delete p;
    set_error("out of device memory");
    return B200MD_ERR_CUDA;
  }
  cudaMemcpy(p->zp.p, zp.data(), sizeof(float) * zp.size(), cudaMemcpyHostToDevice);
  P.model = model;
  P.nt = nt;
  P.rc = rc;
  P.zp = p->zp.p;
  P.n = num_atoms;
  P.atoms = p->nb.atoms.p;
  P.nn_skin = p->nb.nn_skin.p;
  P.nl_skin = p->nb.nl_skin.p;
  P.Fp = p->Fp.p;
  P.acc = p->acc.p;
  cudaDeviceSynchronize();
  *out = p;
  return B200MD_OK;
}

void b200md_eam_destroy(b200md_eam* p) { delete p; }
double b200md_eam_rc(const b200md_eam* p) { return p->rc; }
const char* b200md_eam_symbol(const b200md_eam* p, int t)
{
  return (t >= 0 && t < p->nt) ? p->symbols[t].c_str() : "";
}
int b200md_eam_info(const b200md_eam* p, int what)
{
  if (what == 0)
    return p->nt;
  if (what == 6) {
    int bits = 0, rebuilds = 0;
    const_cast<b200md_eam*>(p)->nb.check(0, &bits, &rebuilds);
    return rebuilds;
  }
  return -1;
}

int b200md_eam_compute(
  b200md_eam* p, int n, const double h[9], const int pbc[3], const int* d_type,
  const double* d_position, double* d_potential, double* d_force, double* d_virial, void* stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  const B2Box box = make_box(h, pbc);
  B2_TRY(p->nb.update(box, d_type, d_position, n, st));
  p->n = n;
  p->view.n = n;
  k_eam_density<<<grid_for(n, 128), 128, 0, st>>>(p->view, box);
  B2_LAUNCHED();
  k_eam_force<<<grid_for(n, 128), 128, 0, st>>>(p->view, box);
  B2_LAUNCHED();
  k_unpack_e<<<grid_for(n, 256), 256, 0, st>>>(
    n, p->nb.perm.p, p->acc.p, d_potential, d_force, d_virial);
  B2_LAUNCHED();
  return B200MD_OK;
}

int b200md_eam_invalidate(b200md_eam* p, int n_new, void* stream)
{
  B2_TRY(p->nb.invalidate(n_new, (cudaStream_t)stream));
  p->n = n_new;
  p->view.n = n_new;
  return B200MD_OK;
}

int b200md_eam_check(b200md_eam* p, void* stream)
{
  int bits = 0, rebuilds = 0;
  B2_TRY(p->nb.check((cudaStream_t)stream, &bits, &rebuilds));
  if (bits) {
    set_error("EAM neighbour-list capacity exceeded on the device");
    return B200MD_ERR_OVERFLOW;
  }
  return B200MD_OK;
}

} // extern "C"
