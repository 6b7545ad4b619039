/*
 * oracle/eam_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement (FP32 pair math, like the reference) of the analytic EAM potentials
 *   Zhou-2004 parameters            src/force/eam.cu:46-92
 *   Dai-2006 (Finnis-Sinclair)      src/force/eam.cu:94-122
 *   pair / density / embedding      src/force/eam.cu:131-280
 *   step 1 (density, F, F')         find_force_eam_step1, eam.cu:283-350
 *   step 2 (forces, virial)         find_force_eam_step2, eam.cu:352-475
 * para (Zhou): 21 numbers per type in file order; (Dai): 9 numbers.
 */
#include "oracle.h"
#include "oracle_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float re_inv, fe, rho_e_inv, rho_s_inv, rho_n, rho_n_inv, rho_0, alpha, beta, A, B, kappa, lambda,
    Fn0, Fn1, Fn2, Fn3, F0, F1, F2, F3, eta, Fe, rc;
} zhou_t;
typedef struct {
  float A, d, c, c0, c1, c2, c3, c4, B, rc;
} dai_t;

static float pw20(float t)
{ /* (x)^20 from x^2 as in eam.cu:136-138 */
  t *= t;
  t *= t * t * t * t;
  return t;
}

static void z_phi(const zhou_t* e, float d, float* phi, float* phip)
{
  float rr = d * e->re_inv;
  float t1 = pw20((rr - e->kappa) * (rr - e->kappa));
  float t2 = pw20((rr - e->lambda) * (rr - e->lambda));
  float p1 = 0.5f * e->A * expf(-e->alpha * (rr - 1.0f)) / (1.0f + t1);
  float p2 = 0.5f * e->B * expf(-e->beta * (rr - 1.0f)) / (1.0f + t2);
  *phi = p1 - p2;
  *phip = (p2 * e->re_inv) * (e->beta + 20.0f * t2 / (rr - e->lambda) / (1.0f + t2)) -
          (p1 * e->re_inv) * (e->alpha + 20.0f * t1 / (rr - e->kappa) / (1.0f + t1));
}

static void z_f_fp(const zhou_t* e, float d, float* f, float* fp)
{
  float rr = d * e->re_inv;
  float t = pw20((rr - e->lambda) * (rr - e->lambda));
  *f = e->fe * expf(-e->beta * (rr - 1.0f)) / (1.0f + t);
  *fp = -(*f * e->re_inv) * (e->beta + 20.0f * t / (rr - e->lambda) / (1.0f + t));
}

static void z_F(const zhou_t* e, float rho, float* F, float* Fp)
{
  if (rho < e->rho_n) {
    float x = rho * e->rho_n_inv - 1.0f;
    *F = ((e->Fn3 * x + e->Fn2) * x + e->Fn1) * x + e->Fn0;
    *Fp = ((3.0f * e->Fn3 * x + 2.0f * e->Fn2) * x + e->Fn1) / e->rho_n;
  } else if (rho < e->rho_0) {
    float x = rho * e->rho_e_inv - 1.0f;
    *F = ((e->F3 * x + e->F2) * x + e->F1) * x + e->F0;
    *Fp = ((3.0f * e->F3 * x + 2.0f * e->F2) * x + e->F1) * e->rho_e_inv;
  } else {
    float x = rho * e->rho_s_inv;
    float xe = powf(x, e->eta);
    *F = e->Fe * (1.0f - e->eta * logf(x)) * xe;
    *Fp = (e->eta / rho) * (*F - e->Fe * xe);
  }
}

int oracle_eam_compute(
  int model, int nt, const double* para, int N, const int* type, const double h[9],
  const int pbc[3], const double* pos, double* pe, double* force, double* virial)
{
  zhou_t z[18];
  dai_t dai;
  float rc = 0.0f;
  memset(z, 0, sizeof z);
  memset(&dai, 0, sizeof dai);
  if (model == 0) {
    if (nt < 1 || nt > 18)
      return -5;
    for (int t = 0; t < nt; ++t) {
      float x[21];
      for (int k = 0; k < 21; ++k)
        x[k] = (float)para[t * 21 + k];
      zhou_t* e = &z[t];
      e->re_inv = 1.0f / x[0];
      e->fe = x[1];
      e->rho_e_inv = 1.0f / x[2];
      e->rho_s_inv = 1.0f / x[3];
      e->alpha = x[4];
      e->beta = x[5];
      e->A = x[6];
      e->B = x[7];
      e->kappa = x[8];
      e->lambda = x[9];
      e->Fn0 = x[10];
      e->Fn1 = x[11];
      e->Fn2 = x[12];
      e->Fn3 = x[13];
      e->F0 = x[14];
      e->F1 = x[15];
      e->F2 = x[16];
      e->F3 = x[17];
      e->eta = x[18];
      e->Fe = x[19];
      e->rc = x[20];
      e->rho_n = x[2] * 0.85;
      e->rho_0 = x[2] * 1.15;
      e->rho_n_inv = 1.0f / e->rho_n;
      if (rc < e->rc)
        rc = e->rc;
    }
  } else {
    float x[9];
    for (int k = 0; k < 9; ++k)
      x[k] = (float)para[k];
    dai.A = x[0];
    dai.d = x[1];
    dai.c = x[2];
    dai.c0 = x[3];
    dai.c1 = x[4];
    dai.c2 = x[5];
    dai.c3 = x[6];
    dai.c4 = x[7];
    dai.B = x[8];
    dai.rc = dai.c > dai.d ? dai.c : dai.d;
    rc = dai.rc;
  }
  oracle_box b;
  oracle_box_init(&b, h, pbc);
  /* candidates: every atom within a slightly larger radius; membership is `d12 < rc` on the
     FP32 distance (eam.cu:320) */
  const int MN = 512;
  int* NN = (int*)calloc((size_t)N, sizeof(int));
  int* NL = (int*)malloc(sizeof(int) * (size_t)N * MN);
  int st = oracle_neighbor_list(N, h, pbc, pos, (double)rc + 0.01, NN, NL, MN);
  if (st != 0) {
    free(NN);
    free(NL);
    return st;
  }
  const double* x = pos;
  const double* y = pos + N;
  const double* zc = pos + 2 * N;
  float* Fp = (float*)calloc((size_t)N, sizeof(float));
  double* lpe = (double*)calloc((size_t)N, sizeof(double));
  for (int i = 0; i < N; ++i) { /* step 1 */
    float rho = 0.0f;
    for (int a = 0; a < NN[i]; ++a) {
      int j = NL[(size_t)i * MN + a];
      float r[3] = {(float)(x[j] - x[i]), (float)(y[j] - y[i]), (float)(zc[j] - zc[i])};
      oracle_mic_f32(&b, &r[0], &r[1], &r[2]);
      float d = sqrtf(oracle_d2_f32(r[0], r[1], r[2]));
      if (d < rc) {
        float f = 0.0f, fp;
        if (model == 0) {
          z_f_fp(&z[type[j]], d, &f, &fp);
        } else if (d <= dai.d) {
          float t = (d - dai.d) * (d - dai.d);
          f = t + dai.B * dai.B * t * t;
        }
        rho += f;
      }
    }
    float F, Fpi;
    if (model == 0) {
      z_F(&z[type[i]], rho, &F, &Fpi);
    } else {
      float s = sqrtf(rho);
      F = -dai.A * s;
      Fpi = -dai.A * 0.5f / s;
    }
    lpe[i] = F;
    Fp[i] = Fpi;
  }
  static const int map[9] = {0, 4, 8, 1, 2, 5, 3, 6, 7};
  for (int i = 0; i < N; ++i) { /* step 2 */
    int t1 = type[i];
    float sf[3] = {0, 0, 0}, sv[9] = {0}, sp = 0.0f;
    for (int a = 0; a < NN[i]; ++a) {
      int j = NL[(size_t)i * MN + a], t2 = type[j];
      float r[3] = {(float)(x[j] - x[i]), (float)(y[j] - y[i]), (float)(zc[j] - zc[i])};
      oracle_mic_f32(&b, &r[0], &r[1], &r[2]);
      float d = sqrtf(oracle_d2_f32(r[0], r[1], r[2]));
      if (!(d < rc))
        continue;
      float phi, phip, fp1, fp2, f1, f2;
      if (model == 0) {
        if (t1 == t2) {
          z_phi(&z[t1], d, &phi, &phip);
          z_f_fp(&z[t1], d, &f1, &fp1);
          fp2 = fp1;
        } else { /* eam.cu:196-218 */
          float ph1, pp1, ph2, pp2;
          z_phi(&z[t1], d, &ph1, &pp1);
          z_phi(&z[t2], d, &ph2, &pp2);
          z_f_fp(&z[t1], d, &f1, &fp1);
          z_f_fp(&z[t2], d, &f2, &fp2);
          float f1i = 1.0f / f1, f2i = 1.0f / f2;
          phi = 0.5f * (ph1 * f2 * f1i + ph2 * f1 * f2i);
          phip = (pp1 * f2 + ph1 * (fp2 - f2 * fp1 * f1i)) * f1i;
          phip += (pp2 * f1 + ph2 * (fp1 - f1 * fp2 * f2i)) * f2i;
          phip *= 0.5f;
        }
      } else {
        if (d > dai.c) {
          phi = 0.0f;
          phip = 0.0f;
        } else {
          float t = ((((dai.c4 * d + dai.c3) * d + dai.c2) * d + dai.c1) * d + dai.c0);
          phi = 0.5f * (d - dai.c) * (d - dai.c) * t;
          phip = 2.0f * (d - dai.c) * t;
          phip += (((4.0f * dai.c4 * d + 3.0f * dai.c3) * d + 2.0f * dai.c2) * d + dai.c1) *
                  (d - dai.c) * (d - dai.c);
          phip *= 0.5f;
        }
        if (d > dai.d) {
          fp1 = 0.0f;
        } else {
          float t = 2.0f * (d - dai.d);
          fp1 = t * (1.0f + dai.B * dai.B * t * (d - dai.d));
        }
        fp2 = fp1;
      }
      float dinv = 1.0f / d;
      phip *= dinv;
      fp1 *= dinv;
      fp2 *= dinv;
      float c12 = phip + Fp[i] * fp2, c21 = phip + Fp[j] * fp1;
      sp += phi;
      for (int dd = 0; dd < 3; ++dd) {
        float f12 = r[dd] * c12, f21 = -r[dd] * c21;
        sf[dd] += f12 - f21;
        for (int ee = 0; ee < 3; ++ee)
          sv[dd * 3 + ee] += r[dd] * (-r[ee] * c21);
      }
    }
    if (pe)
      pe[i] = lpe[i] + sp;
    if (force)
      for (int dd = 0; dd < 3; ++dd)
        force[(size_t)dd * N + i] = sf[dd];
    if (virial)
      for (int k = 0; k < 9; ++k)
        virial[(size_t)k * N + i] = sv[map[k]];
  }
  free(NN);
  free(NL);
  free(Fp);
  free(lpe);
  return 0;
}
