/*
 * oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin extern "C" shim over the reference's own vendored CPU implementation NEP_CPU
 * (/root/reference/tools/Miscellaneous/for_coding/for_perioidc_table/nep.{h,cpp}; class NEP3,
 * nep.h:95-112, NEP3::compute nep.cpp:2909-2991).  oracle/Makefile compiles that source WHERE IT
 * LIES together with this file into oracle/_ref/libnep_cpu_ref.so (git-ignored; it travels to the
 * GPU box).  No reference source is copied into this repository.  Used to (1) validate
 * oracle/nep_oracle.c and (2) time the reference's CPU path (`cpu_baseline.kind = "reference"`).
 *
 * Layouts on this boundary are GPUMD's (see oracle.h); the shim converts to NEP_CPU's
 * (box = ax,bx,cx,ay,by,cy,az,bz,cz -- identical to GPUMD's row-major h; virial rows
 * xx,xy,xz,yx,yy,yz,zx,zy,zz -> GPUMD's xx,yy,zz,xy,xz,yz,yx,zx,zy).
 */
#include "nep.h"
#include <cstring>
#include <string>
#include <vector>

extern "C" {

void* refnep_load(const char* path)
{
  NEP3* m = new NEP3(std::string(path));
  return (void*)m;
}

void refnep_free(void* h) { delete (NEP3*)h; }

int refnep_compute(
  void* handle, int N, const int* type, const double h9[9], const double* position, double* pe,
  double* force, double* virial)
{
  NEP3* m = (NEP3*)handle;
  std::vector<int> t(type, type + N);
  std::vector<double> box(h9, h9 + 9);
  std::vector<double> pos(position, position + 3 * (size_t)N);
  std::vector<double> p(N), f(3 * (size_t)N), v(9 * (size_t)N);
  m->compute(t, box, pos, p, f, v);
  if (pe)
    std::memcpy(pe, p.data(), sizeof(double) * N);
  if (force)
    std::memcpy(force, f.data(), sizeof(double) * 3 * (size_t)N);
  if (virial) {
    static const int map[9] = {0, 4, 8, 1, 2, 5, 3, 6, 7};
    for (int k = 0; k < 9; ++k)
      std::memcpy(virial + (size_t)k * N, v.data() + (size_t)map[k] * N, sizeof(double) * N);
  }
  return 0;
}
}
