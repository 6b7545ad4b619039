// b2_lj.cuh -- LJ pair loop body (replaces gpu_find_force, src/force/lj.cu:77-181).
// Walks the skin list and rejects by the type-pair cutoff exactly like the reference
// (lj.cu:117-130); half-force convention f12 = f2*r12/2, f21 = -f12 (lj.cu:133-139).
#pragma once
#include "b2_common.cuh"

struct B2LjView {
  int nt;
  const float* s6e4;  // [nt*nt]
  const float* s12e4; // [nt*nt]
  const float* rc2;   // [nt*nt]
  int n;
  const B2Atom* atoms;
  const int* nn_skin;
  const int* nl_skin;
  double* acc; // [13*n] sorted order
};

B2_HD void b2_body_lj(int i, const B2LjView& P, const B2Box& box)
{
  const size_t N = (size_t)P.n;
  const B2Geo geo = b2_geo(box);
  const B2Atom a1 = P.atoms[i];
  const int row = a1.type * P.nt;
  const int nn = P.nn_skin[i];
  float pe = 0.0f, fx = 0.0f, fy = 0.0f, fz = 0.0f;
  float vxx = 0.0f, vyy = 0.0f, vzz = 0.0f, vxy = 0.0f, vxz = 0.0f, vyz = 0.0f;
  // software pipeline: record of candidate k+1 and index of candidate k+2 in flight
  int jn = nn > 0 ? P.nl_skin[i] : i;
  B2Atom an = b2_load_atom(&P.atoms[jn]);
  int j2 = nn > 1 ? P.nl_skin[N + i] : i;
  for (int k = 0; k < nn; ++k) {
    const B2Atom a2 = an;
    an = b2_load_atom(&P.atoms[j2]);
    j2 = (k + 2 < nn) ? P.nl_skin[(size_t)(k + 2) * N + i] : i;
    float x12, y12, z12;
    b2_r12(geo, box, a1, a2, x12, y12, z12);
    const float d2 = b2_d2(x12, y12, z12);
    const int pair = row + a2.type;
    if (d2 >= B2_LDG(&P.rc2[pair]))
      continue;
    const float i2 = 1.0f / d2;
    const float i6 = i2 * i2 * i2;
    const float s6 = B2_LDG(&P.s6e4[pair]), s12 = B2_LDG(&P.s12e4[pair]);
    const float f2 = 6.0f * (s6 * i6 - s12 * 2.0f * i6 * i6) * i2; // (dU/dr)/r, lj.cu:67-75
    const float p2 = s12 * i6 * i6 - s6 * i6;
    // F_i += f12 - f21 = f2*r12 ;  W_i += r12 (x) f21 = -(f2/2) r12 (x) r12
    fx = fmaf(f2, x12, fx);
    fy = fmaf(f2, y12, fy);
    fz = fmaf(f2, z12, fz);
    const float h = -0.5f * f2;
    vxx = fmaf(x12 * x12, h, vxx);
    vyy = fmaf(y12 * y12, h, vyy);
    vzz = fmaf(z12 * z12, h, vzz);
    vxy = fmaf(x12 * y12, h, vxy);
    vxz = fmaf(x12 * z12, h, vxz);
    vyz = fmaf(y12 * z12, h, vyz);
    pe = fmaf(p2, 0.5f, pe);
  }
  double* a = P.acc + i;
  a[0] = pe;
  a[1 * N] = fx;
  a[2 * N] = fy;
  a[3 * N] = fz;
  a[4 * N] = vxx;
  a[5 * N] = vyy;
  a[6 * N] = vzz;
  a[7 * N] = vxy;
  a[8 * N] = vxz;
  a[9 * N] = vyz;
  a[10 * N] = vxy;
  a[11 * N] = vxz;
  a[12 * N] = vyz;
}
