// b2_host.cu -- error string, launch counter, box helper.
#include "../../include/b200md.h"
#include "b2_host.h"
#include <cmath>
#include <cstdio>

namespace b2 {

static thread_local std::string g_error;
long long g_launch_count = 0;

void set_error(const std::string& msg) { g_error = msg; }

bool cuda_ok(cudaError_t e, const char* what, const char* file, int line)
{
  if (e == cudaSuccess)
    return true;
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e), file, line, what);
  g_error = buf;
  return false;
}

B2Box make_box(const double h[9], const int pbc[3])
{
  B2Box b;
  double* c = b.h;
  for (int d = 0; d < 9; ++d)
    c[d] = h[d];
  for (int d = 0; d < 3; ++d)
    b.pbc[d] = pbc[d] ? 1 : 0;
  // adjugate / determinant
  c[9] = c[4] * c[8] - c[5] * c[7];
  c[10] = c[2] * c[7] - c[1] * c[8];
  c[11] = c[1] * c[5] - c[2] * c[4];
  c[12] = c[5] * c[6] - c[3] * c[8];
  c[13] = c[0] * c[8] - c[2] * c[6];
  c[14] = c[2] * c[3] - c[0] * c[5];
  c[15] = c[3] * c[7] - c[4] * c[6];
  c[16] = c[1] * c[6] - c[0] * c[7];
  c[17] = c[0] * c[4] - c[1] * c[3];
  const double det = c[0] * c[9] + c[1] * c[12] + c[2] * c[15];
  for (int d = 9; d < 18; ++d)
    c[d] /= det;
  b.volume = std::fabs(det);
  for (int d = 0; d < 3; ++d) {
    // area of the face spanned by the other two lattice vectors (columns of h)
    const int p = (d + 1) % 3, q = (d + 2) % 3;
    const double u[3] = {c[p], c[p + 3], c[p + 6]}, w[3] = {c[q], c[q + 3], c[q + 6]};
    const double s0 = u[1] * w[2] - u[2] * w[1], s1 = u[2] * w[0] - u[0] * w[2],
                 s2 = u[0] * w[1] - u[1] * w[0];
    b.thickness[d] = b.volume / std::sqrt(s0 * s0 + s1 * s1 + s2 * s2);
  }
  b.ortho = c[1] == 0 && c[2] == 0 && c[3] == 0 && c[5] == 0 && c[6] == 0 && c[7] == 0;
  for (int d = 0; d < 18; ++d)
    b.hf[d] = (float)c[d];
  return b;
}

StageProfiler::~StageProfiler()
{
  if (created)
    for (int a = 0; a < MAX_SAMPLES; ++a)
      for (int b = 0; b < MAX_STAGES; ++b) {
        cudaEventDestroy(ev[a][b][0]);
        cudaEventDestroy(ev[a][b][1]);
      }
}

void StageProfiler::enable(bool on)
{
  if (on && !created) {
    for (int a = 0; a < MAX_SAMPLES; ++a)
      for (int b = 0; b < MAX_STAGES; ++b) {
        cudaEventCreate(&ev[a][b][0]);
        cudaEventCreate(&ev[a][b][1]);
      }
    created = true;
  }
  enabled = on;
  steps = 0;
  for (int a = 0; a < MAX_SAMPLES; ++a)
    for (int b = 0; b < MAX_STAGES; ++b)
      used[a][b] = false;
}

void StageProfiler::read(float* ms_sum, int* counts)
{
  cudaDeviceSynchronize();
  for (int b = 0; b < MAX_STAGES; ++b) {
    ms_sum[b] = 0.0f;
    counts[b] = 0;
  }
  const int ns = steps < MAX_SAMPLES ? steps : MAX_SAMPLES;
  for (int a = 0; a < ns; ++a)
    for (int b = 0; b < MAX_STAGES; ++b)
      if (used[a][b]) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ev[a][b][0], ev[a][b][1]) == cudaSuccess) {
          ms_sum[b] += ms;
          counts[b] += 1;
        }
      }
}

} // namespace b2

extern "C" const char* b200md_last_error(void) { return b2::g_error.c_str(); }
extern "C" long long b200md_launch_count(void) { return b2::g_launch_count; }
