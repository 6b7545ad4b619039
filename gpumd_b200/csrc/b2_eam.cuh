// b2_eam.cuh -- analytic EAM bodies (Zhou-2004 and Dai-2006), FP32 pair math like the reference
// (src/force/eam.cu:131-475).  Two gather-only passes over the skin list: density/embedding, then
// pair + embedding forces with F'(rho_j) gathered from the neighbour.  Unlike the reference no
// local list is written in pass 1 and re-read in pass 2 (eam.cu:315-334): both passes apply the
// same FP32 membership test `d12 < rc` to the cell-sorted skin list.
#pragma once
#include "b2_common.cuh"

// per-type Zhou-2004 record (24 floats): see b2_eam.cu for the order
enum {
  EZ_RE_INV, EZ_FE, EZ_RHO_E_INV, EZ_RHO_S_INV, EZ_RHO_N, EZ_RHO_N_INV, EZ_RHO_0, EZ_ALPHA, EZ_BETA,
  EZ_A, EZ_B, EZ_KAPPA, EZ_LAMBDA, EZ_FN0, EZ_FN1, EZ_FN2, EZ_FN3, EZ_F0, EZ_F1, EZ_F2, EZ_F3,
  EZ_ETA, EZ_FE_EMBED, EZ_RC, EZ_COUNT
};

struct B2EamView {
  int model; // 0 Zhou-2004, 1 Dai-2006
  int nt;
  float rc;
  const float* zp; // [nt * EZ_COUNT]
  float dA, dd, dc, dc0, dc1, dc2, dc3, dc4, dB; // Dai-2006
  int n;
  const B2Atom* atoms;
  const int* nn_skin;
  const int* nl_skin;
  float* Fp;   // [n]
  double* acc; // [13 * n]
};

B2_HD float b2_pow20_from_sq(float t)
{
  t *= t;
  t *= t * t * t * t;
  return t;
}

B2_HD void b2_zhou_f_fp(const float* e, float d, float& f, float& fp)
{
  const float rr = d * e[EZ_RE_INV];
  const float dl = rr - e[EZ_LAMBDA];
  const float t = b2_pow20_from_sq(dl * dl);
  f = e[EZ_FE] * expf(-e[EZ_BETA] * (rr - 1.0f)) / (1.0f + t);
  fp = -(f * e[EZ_RE_INV]) * (e[EZ_BETA] + 20.0f * t / dl / (1.0f + t));
}

B2_HD void b2_zhou_phi(const float* e, float d, float& phi, float& phip)
{
  const float rr = d * e[EZ_RE_INV];
  const float dk = rr - e[EZ_KAPPA], dl = rr - e[EZ_LAMBDA];
  const float t1 = b2_pow20_from_sq(dk * dk), t2 = b2_pow20_from_sq(dl * dl);
  const float p1 = 0.5f * e[EZ_A] * expf(-e[EZ_ALPHA] * (rr - 1.0f)) / (1.0f + t1);
  const float p2 = 0.5f * e[EZ_B] * expf(-e[EZ_BETA] * (rr - 1.0f)) / (1.0f + t2);
  phi = p1 - p2;
  phip = (p2 * e[EZ_RE_INV]) * (e[EZ_BETA] + 20.0f * t2 / dl / (1.0f + t2)) -
         (p1 * e[EZ_RE_INV]) * (e[EZ_ALPHA] + 20.0f * t1 / dk / (1.0f + t1));
}

B2_HD void b2_zhou_embed(const float* e, float rho, float& F, float& Fp)
{
  if (rho < e[EZ_RHO_N]) {
    const float x = rho * e[EZ_RHO_N_INV] - 1.0f;
    F = ((e[EZ_FN3] * x + e[EZ_FN2]) * x + e[EZ_FN1]) * x + e[EZ_FN0];
    Fp = ((3.0f * e[EZ_FN3] * x + 2.0f * e[EZ_FN2]) * x + e[EZ_FN1]) / e[EZ_RHO_N];
  } else if (rho < e[EZ_RHO_0]) {
    const float x = rho * e[EZ_RHO_E_INV] - 1.0f;
    F = ((e[EZ_F3] * x + e[EZ_F2]) * x + e[EZ_F1]) * x + e[EZ_F0];
    Fp = ((3.0f * e[EZ_F3] * x + 2.0f * e[EZ_F2]) * x + e[EZ_F1]) * e[EZ_RHO_E_INV];
  } else {
    const float x = rho * e[EZ_RHO_S_INV];
    const float xe = powf(x, e[EZ_ETA]);
    F = e[EZ_FE_EMBED] * (1.0f - e[EZ_ETA] * logf(x)) * xe;
    Fp = (e[EZ_ETA] / rho) * (F - e[EZ_FE_EMBED] * xe);
  }
}

// pass 1: rho_i = sum_j f_{t_j}(r_ij); F(rho_i), F'(rho_i)   (find_force_eam_step1, eam.cu:283-350)
B2_HD void b2_body_eam_density(int i, const B2EamView& P, const B2Box& box)
{
  const size_t N = (size_t)P.n;
  const B2Geo geo = b2_geo(box);
  const B2Atom a1 = P.atoms[i];
  const int nn = P.nn_skin[i];
  float rho = 0.0f;
  for (int k = 0; k < nn; ++k) {
    const int j = P.nl_skin[(size_t)k * N + i];
    const B2Atom a2 = b2_load_atom(&P.atoms[j]);
    float x12, y12, z12;
    b2_r12(geo, box, a1, a2, x12, y12, z12);
    const float d = sqrtf(b2_d2(x12, y12, z12));
    if (d < P.rc) {
      float f = 0.0f, fp;
      if (P.model == 0) {
        b2_zhou_f_fp(P.zp + a2.type * EZ_COUNT, d, f, fp);
      } else if (!(d > P.dd)) {
        const float t = (d - P.dd) * (d - P.dd);
        f = t + P.dB * P.dB * t * t;
      }
      rho += f;
    }
  }
  float F, Fp;
  if (P.model == 0) {
    b2_zhou_embed(P.zp + a1.type * EZ_COUNT, rho, F, Fp);
  } else {
    const float s = sqrtf(rho);
    F = -P.dA * s;
    Fp = -P.dA * 0.5f / s;
  }
  P.Fp[i] = Fp;
  P.acc[i] = (double)F;
}

// pass 2: pair + embedding forces (find_force_eam_step2, eam.cu:352-475)
B2_HD void b2_body_eam_force(int i, const B2EamView& P, const B2Box& box)
{
  const size_t N = (size_t)P.n;
  const B2Geo geo = b2_geo(box);
  const B2Atom a1 = P.atoms[i];
  const int t1 = a1.type;
  const int nn = P.nn_skin[i];
  const float Fp1 = P.Fp[i];
  const float* e1 = P.zp + t1 * EZ_COUNT;
  float pe = 0.0f, fx = 0.0f, fy = 0.0f, fz = 0.0f;
  float vxx = 0.0f, vyy = 0.0f, vzz = 0.0f, vxy = 0.0f, vxz = 0.0f, vyz = 0.0f;
  for (int k = 0; k < nn; ++k) {
    const int j = P.nl_skin[(size_t)k * N + i];
    const B2Atom a2 = b2_load_atom(&P.atoms[j]);
    float x12, y12, z12;
    b2_r12(geo, box, a1, a2, x12, y12, z12);
    const float d = sqrtf(b2_d2(x12, y12, z12));
    if (!(d < P.rc))
      continue;
    const float Fp2 = B2_LDG(&P.Fp[j]);
    float phi, phip, fp1, fp2;
    if (P.model == 0) {
      float f1;
      b2_zhou_phi(e1, d, phi, phip);
      b2_zhou_f_fp(e1, d, f1, fp1);
      fp2 = fp1;
      if (a2.type != t1) { // alloy mixing, eam.cu:196-218
        const float* e2 = P.zp + a2.type * EZ_COUNT;
        float ph2, pp2, f2;
        b2_zhou_phi(e2, d, ph2, pp2);
        b2_zhou_f_fp(e2, d, f2, fp2);
        const float f1i = 1.0f / f1, f2i = 1.0f / f2;
        const float ph1 = phi, pp1 = phip;
        phi = 0.5f * (ph1 * f2 * f1i + ph2 * f1 * f2i);
        phip = (pp1 * f2 + ph1 * (fp2 - f2 * fp1 * f1i)) * f1i;
        phip += (pp2 * f1 + ph2 * (fp1 - f1 * fp2 * f2i)) * f2i;
        phip *= 0.5f;
      }
    } else {
      if (d > P.dc) {
        phi = 0.0f;
        phip = 0.0f;
      } else {
        const float t = ((((P.dc4 * d + P.dc3) * d + P.dc2) * d + P.dc1) * d + P.dc0);
        const float dcut = d - P.dc;
        phi = 0.5f * dcut * dcut * t;
        phip = 2.0f * dcut * t;
        phip += (((4.0f * P.dc4 * d + 3.0f * P.dc3) * d + 2.0f * P.dc2) * d + P.dc1) * dcut * dcut;
        phip *= 0.5f;
      }
      if (d > P.dd) {
        fp1 = 0.0f;
      } else {
        const float t = 2.0f * (d - P.dd);
        fp1 = t * (1.0f + P.dB * P.dB * t * (d - P.dd));
      }
      fp2 = fp1;
    }
    const float dinv = 1.0f / d;
    phip *= dinv;
    fp1 *= dinv;
    fp2 *= dinv;
    const float c12 = phip + Fp1 * fp2; // f12 = r12 * c12
    const float c21 = phip + Fp2 * fp1; // f21 = -r12 * c21
    pe += phi;
    const float cs = c12 + c21;
    fx = fmaf(cs, x12, fx);
    fy = fmaf(cs, y12, fy);
    fz = fmaf(cs, z12, fz);
    vxx = fmaf(x12 * x12, -c21, vxx);
    vyy = fmaf(y12 * y12, -c21, vyy);
    vzz = fmaf(z12 * z12, -c21, vzz);
    vxy = fmaf(x12 * y12, -c21, vxy);
    vxz = fmaf(x12 * z12, -c21, vxz);
    vyz = fmaf(y12 * z12, -c21, vyz);
  }
  double* a = P.acc + i;
  a[0] += (double)pe;
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
