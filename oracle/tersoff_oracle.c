/*
 * oracle/tersoff_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement (FP64 throughout, like the reference) of
 *   Tersoff1989 parameters / mixing   src/force/tersoff1989.cu:31-131
 *   local neighbour filter (FP32)     src/force/neighbor.cu:699-737
 *   bond order b_ij, b'_ij            find_force_tersoff_step1, tersoff1989.cu:337-405
 *   partial forces dU_i/dr_ij         find_force_tersoff_step2, tersoff1989.cu:408-505
 *   F_i, W_i from f12 / f21           gpu_find_force_many_body (double), potential.cu:35-134
 *   heat current                      gpu_compute_heat, src/measure/compute_heat.cu:32-63
 * Pinned against the reference gpumd run on a B200 (tests/golden/refgpu_sp_si.npz).
 */
#include "oracle.h"
#include "oracle_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  double a, b, lambda, mu, beta, n, c2, d2, h, r1, r2, one_plus, pi_factor, mhn;
} ters_t;

static void ters_from(const double* p, ters_t* t)
{
  t->a = p[0];
  t->b = p[1];
  t->lambda = p[2];
  t->mu = p[3];
  t->beta = p[4];
  t->n = p[5];
  t->c2 = p[6] * p[6];
  t->d2 = p[7] * p[7];
  t->h = p[8];
  t->r1 = p[9];
  t->r2 = p[10];
  t->one_plus = 1.0 + t->c2 / t->d2;
  t->pi_factor = 3.14159265358979 / (t->r2 - t->r1); /* PI of common.cuh:19 */
  t->mhn = -0.5 / p[5];
}

static const ters_t* pick(const ters_t* t, int t1, int t2)
{
  if (t1 == 0 && t2 == 0)
    return &t[0];
  if (t1 == 1 && t2 == 1)
    return &t[1];
  return &t[2];
}

static void fc_fcp(const ters_t* t, double d, double* fc, double* fcp)
{
  if (d < t->r1) {
    *fc = 1.0;
    *fcp = 0.0;
  } else if (d < t->r2) {
    *fc = cos(t->pi_factor * (d - t->r1)) * 0.5 + 0.5;
    *fcp = -sin(t->pi_factor * (d - t->r1)) * t->pi_factor * 0.5;
  } else {
    *fc = 0.0;
    *fcp = 0.0;
  }
}

int oracle_tersoff_compute(
  int nt, const double* para, int N, const int* type, const double h[9], const int pbc[3],
  const double* pos, double* pe, double* force, double* virial)
{
  if (nt < 1 || nt > 2)
    return -5;
  ters_t t[3];
  ters_from(para, &t[0]);
  double rc = t[0].r2;
  if (nt == 2) {
    ters_from(para + 11, &t[1]);
    const double chi = para[22];
    memset(&t[2], 0, sizeof t[2]);
    t[2].a = sqrt(t[0].a * t[1].a); /* tersoff1989.cu:107-124 */
    t[2].b = sqrt(t[0].b * t[1].b) * chi;
    t[2].lambda = 0.5 * (t[0].lambda + t[1].lambda);
    t[2].mu = 0.5 * (t[0].mu + t[1].mu);
    t[2].r1 = sqrt(t[0].r1 * t[1].r1);
    t[2].r2 = sqrt(t[0].r2 * t[1].r2);
    t[2].pi_factor = 3.14159265358979 / (t[2].r2 - t[2].r1);
    rc = t[0].r2 > t[1].r2 ? t[0].r2 : t[1].r2;
  } else {
    t[1] = t[0];
    t[2] = t[0];
  }
  const int MN = 64;
  int* NN = (int*)calloc((size_t)N, sizeof(int));
  int* NL = (int*)malloc(sizeof(int) * (size_t)N * MN);
  /* local list: FP32 membership d2 < rc*rc of the (sorted) global list, neighbor.cu:720-735 */
  int st = oracle_neighbor_list(N, h, pbc, pos, rc, NN, NL, MN);
  if (st != 0) {
    free(NN);
    free(NL);
    return st;
  }
  oracle_box b;
  oracle_box_init(&b, h, pbc);
  const double* x = pos;
  const double* y = pos + N;
  const double* z = pos + 2 * N;
  double* B = (double*)calloc((size_t)N * MN, sizeof(double));
  double* BP = (double*)calloc((size_t)N * MN, sizeof(double));
  double* F12 = (double*)calloc((size_t)N * MN * 3, sizeof(double));
  double* lpe = (double*)calloc((size_t)N, sizeof(double));
  /* step 1 */
  for (int i = 0; i < N; ++i) {
    int t1 = type[i];
    for (int a = 0; a < NN[i]; ++a) {
      int j = NL[(size_t)i * MN + a];
      double x12 = x[j] - x[i], y12 = y[j] - y[i], z12 = z[j] - z[i];
      oracle_mic_f64(&b, &x12, &y12, &z12);
      double d12 = sqrt(x12 * x12 + y12 * y12 + z12 * z12);
      double zeta = 0.0;
      for (int c = 0; c < NN[i]; ++c) {
        int k = NL[(size_t)i * MN + c];
        if (k == j)
          continue;
        double x13 = x[k] - x[i], y13 = y[k] - y[i], z13 = z[k] - z[i];
        oracle_mic_f64(&b, &x13, &y13, &z13);
        double d13 = sqrt(x13 * x13 + y13 * y13 + z13 * z13);
        double cs = (x12 * x13 + y12 * y13 + z12 * z13) / (d12 * d13);
        double fc13, fcp13;
        fc_fcp(pick(t, t1, type[k]), d13, &fc13, &fcp13);
        const ters_t* ti = &t[t1];
        double tmp = ti->d2 + (cs - ti->h) * (cs - ti->h);
        zeta += fc13 * (ti->one_plus - ti->c2 / tmp);
      }
      const ters_t* ti = &t[t1];
      double bzn = pow(ti->beta * zeta, ti->n);
      double b12 = pow(1.0 + bzn, ti->mhn);
      if (zeta < 1.0e-16) {
        B[(size_t)i * MN + a] = 1.0;
        BP[(size_t)i * MN + a] = 0.0;
      } else {
        B[(size_t)i * MN + a] = b12;
        BP[(size_t)i * MN + a] = -b12 * bzn * 0.5 / ((1.0 + bzn) * zeta);
      }
    }
  }
  /* step 2 */
  for (int i = 0; i < N; ++i) {
    int t1 = type[i];
    const ters_t* ti = &t[t1];
    for (int a = 0; a < NN[i]; ++a) {
      int j = NL[(size_t)i * MN + a];
      const ters_t* p12 = pick(t, t1, type[j]);
      double x12 = x[j] - x[i], y12 = y[j] - y[i], z12 = z[j] - z[i];
      oracle_mic_f64(&b, &x12, &y12, &z12);
      double d12 = sqrt(x12 * x12 + y12 * y12 + z12 * z12), d12inv = 1.0 / d12;
      double fc12, fcp12;
      fc_fcp(p12, d12, &fc12, &fcp12);
      double fa12 = p12->b * exp(-p12->mu * d12), fap12 = -p12->mu * fa12;
      double fr12 = p12->a * exp(-p12->lambda * d12), frp12 = -p12->lambda * fr12;
      double b12 = B[(size_t)i * MN + a], bp12 = BP[(size_t)i * MN + a];
      double f3 = (fcp12 * (fr12 - b12 * fa12) + fc12 * (frp12 - b12 * fap12)) * d12inv;
      double f12[3] = {x12 * f3 * 0.5, y12 * f3 * 0.5, z12 * f3 * 0.5};
      lpe[i] += fc12 * (fr12 - b12 * fa12) * 0.5;
      for (int c = 0; c < NN[i]; ++c) {
        int k = NL[(size_t)i * MN + c];
        if (k == j)
          continue;
        const ters_t* p13 = pick(t, t1, type[k]);
        double x13 = x[k] - x[i], y13 = y[k] - y[i], z13 = z[k] - z[i];
        oracle_mic_f64(&b, &x13, &y13, &z13);
        double d13 = sqrt(x13 * x13 + y13 * y13 + z13 * z13);
        double fc13, fcp13;
        fc_fcp(p13, d13, &fc13, &fcp13);
        double fa13 = p13->b * exp(-p13->mu * d13);
        double bp13 = BP[(size_t)i * MN + c];
        double inv = 1.0 / (d12 * d13);
        double cs = (x12 * x13 + y12 * y13 + z12 * z13) * inv;
        double cs_dd = cs * d12inv * d12inv;
        double tmp = ti->d2 + (cs - ti->h) * (cs - ti->h);
        double g = ti->one_plus - ti->c2 / tmp;
        double gp = 2.0 * ti->c2 * (cs - ti->h) / (tmp * tmp);
        double ta = (-bp12 * fc12 * fa12 * fc13 - bp13 * fc13 * fa13 * fc12) * gp;
        double tb = -bp13 * fc13 * fa13 * fcp12 * g * d12inv;
        f12[0] += (x12 * tb + ta * (x13 * inv - x12 * cs_dd)) * 0.5;
        f12[1] += (y12 * tb + ta * (y13 * inv - y12 * cs_dd)) * 0.5;
        f12[2] += (z12 * tb + ta * (z13 * inv - z12 * cs_dd)) * 0.5;
      }
      for (int d = 0; d < 3; ++d)
        F12[((size_t)i * MN + a) * 3 + d] = f12[d];
    }
  }
  /* reduction */
  static const int map[9] = {0, 4, 8, 1, 2, 5, 3, 6, 7};
  for (int i = 0; i < N; ++i) {
    double f[3] = {0, 0, 0}, v[9] = {0};
    for (int a = 0; a < NN[i]; ++a) {
      int j = NL[(size_t)i * MN + a];
      double r[3] = {x[j] - x[i], y[j] - y[i], z[j] - z[i]};
      oracle_mic_f64(&b, &r[0], &r[1], &r[2]);
      int off = 0;
      for (int k = 0; k < NN[j]; ++k)
        if (NL[(size_t)j * MN + k] == i) {
          off = k;
          break;
        }
      const double* f12 = &F12[((size_t)i * MN + a) * 3];
      const double* f21 = &F12[((size_t)j * MN + off) * 3];
      for (int d = 0; d < 3; ++d) {
        f[d] += f12[d] - f21[d];
        for (int e = 0; e < 3; ++e)
          v[d * 3 + e] += r[d] * f21[e];
      }
    }
    if (pe)
      pe[i] = lpe[i];
    if (force)
      for (int d = 0; d < 3; ++d)
        force[(size_t)d * N + i] = f[d];
    if (virial)
      for (int k = 0; k < 9; ++k)
        virial[(size_t)k * N + i] = v[map[k]];
  }
  free(NN);
  free(NL);
  free(B);
  free(BP);
  free(F12);
  free(lpe);
  return 0;
}

void oracle_compute_heat(int N, const double* w, const double* vel, double* heat)
{
  /* virial rows: 0 xx 1 yy 2 zz 3 xy 4 xz 5 yz 6 yx 7 zx 8 zy (compute_heat.cu:70-76) */
  const size_t n = (size_t)N;
  for (int i = 0; i < N; ++i) {
    const double vx = vel[i], vy = vel[n + i], vz = vel[2 * n + i];
    heat[i] = w[i] * vx + w[3 * n + i] * vy;
    heat[n + i] = w[4 * n + i] * vz;
    heat[2 * n + i] = w[6 * n + i] * vx + w[n + i] * vy;
    heat[3 * n + i] = w[5 * n + i] * vz;
    heat[4 * n + i] = w[7 * n + i] * vx + w[8 * n + i] * vy + w[2 * n + i] * vz;
  }
}
