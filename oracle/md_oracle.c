/*
 * oracle/md_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatements of the non-NEP pieces of the hot path:
 *   LJ pair loop           src/force/lj.cu:67-181 (FP32 pair math over the FP32 cutoff sets)
 *   position wrap          src/force/force.cu:424-459
 *   velocity-Verlet        src/integrate/ensemble.cu:176-214
 *   thermo reduction       src/integrate/ensemble.cu:434-633
 * LJ has no golden vector in the reference's own tests.  It is pinned instead
 * against the output of the reference itself run on a B200 (tests/golden/refgpu_sp_lj.npz, made by
 * scripts/run_reference_gpumd.py with oracle/_ref/gpumd_ref) and against the closed form.
 */
#include "oracle.h"
#include "oracle_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int oracle_lj_compute(
  int nt, const double* para, int N, const int* type, const double h[9], const int pbc[3],
  const double* pos, double* pe, double* force, double* virial)
{
  if (nt < 1 || nt > 10) /* MAX_TYPE, lj.cuh:22 */
    return -5;
  float s6e4[10][10], s12e4[10][10], rc2[10][10];
  double rc = 0.0;
  for (int a = 0; a < nt; ++a)
    for (int b = 0; b < nt; ++b) { /* lj.cu:45-56 */
      double eps = para[(a * nt + b) * 3], sig = para[(a * nt + b) * 3 + 1],
             cut = para[(a * nt + b) * 3 + 2];
      s6e4[a][b] = (float)(pow(sig, 6.0) * eps * 4.0);
      s12e4[a][b] = (float)(pow(sig, 12.0) * eps * 4.0);
      rc2[a][b] = (float)(cut * cut);
      if (rc < cut)
        rc = cut;
    }
  oracle_box b;
  oracle_box_init(&b, h, pbc);
  for (int d = 0; d < 3; ++d)
    if (pbc[d] && b.thickness[d] < 2.0 * rc)
      return -2;
  oracle_cells c;
  if (oracle_cells_build(&c, &b, N, pos, rc) != 0)
    return -1;
  const double* x = pos;
  const double* y = pos + N;
  const double* z = pos + 2 * N;
  int* nl = (int*)malloc(sizeof(int) * 4096);
  for (int i = 0; i < N; ++i) {
    int around[27];
    int na = oracle_cells_around(&c, &b, c.cell_of[i], around);
    int cnt = 0;
    for (int a = 0; a < na; ++a)
      for (int k = c.start[around[a]]; k < c.start[around[a] + 1]; ++k)
        if (c.items[k] != i && cnt < 4096)
          nl[cnt++] = c.items[k];
    /* ascending neighbour index = the reference's summation order (neighbor.cuh:112-136) */
    for (int a = 1; a < cnt; ++a) {
      int v = nl[a], q = a - 1;
      while (q >= 0 && nl[q] > v) {
        nl[q + 1] = nl[q];
        --q;
      }
      nl[q + 1] = v;
    }
    int t1 = type[i];
    float sf[3] = {0, 0, 0}, sp = 0, sv[9] = {0};
    for (int k = 0; k < cnt; ++k) { /* lj.cu:113-160 */
      int j = nl[k], t2 = type[j];
      float r[3];
      r[0] = (float)(x[j] - x[i]);
      r[1] = (float)(y[j] - y[i]);
      r[2] = (float)(z[j] - z[i]);
      oracle_mic_f32(&b, &r[0], &r[1], &r[2]);
      float d2 = oracle_d2_f32(r[0], r[1], r[2]);
      if (d2 >= rc2[t1][t2])
        continue;
      float i2 = 1.0f / d2; /* find_p2_and_f2, lj.cu:67-75 */
      float i6 = i2 * i2 * i2;
      float f2 = 6.0f * (s6e4[t1][t2] * i6 - s12e4[t1][t2] * 2.0f * i6 * i6) * i2;
      float p2 = s12e4[t1][t2] * i6 * i6 - s6e4[t1][t2] * i6;
      float f12[3] = {f2 * r[0] * 0.5f, f2 * r[1] * 0.5f, f2 * r[2] * 0.5f};
      for (int dd = 0; dd < 3; ++dd) {
        sf[dd] += f12[dd] - (-f12[dd]);
        for (int ee = 0; ee < 3; ++ee)
          sv[dd * 3 + ee] += r[dd] * (-f12[ee]);
      }
      sp += p2 * 0.5f;
    }
    if (pe)
      pe[i] = sp;
    if (force)
      for (int dd = 0; dd < 3; ++dd)
        force[(size_t)dd * N + i] = sf[dd];
    if (virial) {
      static const int map[9] = {0, 4, 8, 1, 2, 5, 3, 6, 7};
      for (int k = 0; k < 9; ++k)
        virial[(size_t)k * N + i] = sv[map[k]];
    }
  }
  free(nl);
  oracle_cells_free(&c);
  return 0;
}

void oracle_apply_pbc(int N, const double h[9], const int pbc[3], double* pos)
{
  oracle_box b;
  oracle_box_init(&b, h, pbc);
  double* x = pos;
  double* y = pos + N;
  double* z = pos + 2 * N;
  for (int n = 0; n < N; ++n) {
    double s[3];
    s[0] = b.h[9] * x[n] + b.h[10] * y[n] + b.h[11] * z[n];
    s[1] = b.h[12] * x[n] + b.h[13] * y[n] + b.h[14] * z[n];
    s[2] = b.h[15] * x[n] + b.h[16] * y[n] + b.h[17] * z[n];
    for (int d = 0; d < 3; ++d)
      if (pbc[d]) {
        if (s[d] < 0.0)
          s[d] += 1.0;
        else if (s[d] > 1.0)
          s[d] -= 1.0;
      }
    x[n] = b.h[0] * s[0] + b.h[1] * s[1] + b.h[2] * s[2];
    y[n] = b.h[3] * s[0] + b.h[4] * s[1] + b.h[5] * s[2];
    z[n] = b.h[6] * s[0] + b.h[7] * s[1] + b.h[8] * s[2];
  }
}

void oracle_velocity_verlet(
  int is_step1, int N, double dt, const double* mass, double* pos, double* vel, const double* f)
{
  /* nvcc contracts `vx += ax * time_step_half` and `g_x[i] += vx * time_step`
     (ensemble.cu:203-212) into FMAs by default; restated explicitly so the device result can be
     compared bit for bit */
  const double half = dt * 0.5;
  for (int i = 0; i < N; ++i) {
    const double minv = 1.0 / mass[i];
    for (int d = 0; d < 3; ++d) {
      double v = vel[(size_t)d * N + i];
      const double a = f[(size_t)d * N + i] * minv;
      v = fma(a, half, v);
      vel[(size_t)d * N + i] = v;
      if (is_step1)
        pos[(size_t)d * N + i] = fma(v, dt, pos[(size_t)d * N + i]);
    }
  }
}

/* gpu_velocity_verlet with fixed / moving groups, src/integrate/ensemble.cu:111-174 */
void oracle_velocity_verlet_groups(
  int is_step1, int N, double dt, const double* mass, double* pos, double* vel, const double* f,
  const int* label, int fixed_group, int move_group, const double* move_velocity)
{
  const double half = dt * 0.5;
  for (int i = 0; i < N; ++i) {
    const double minv = 1.0 / mass[i];
    for (int d = 0; d < 3; ++d) {
      double v = vel[(size_t)d * N + i];
      const double a = f[(size_t)d * N + i] * minv;
      if (label[i] == fixed_group) {
        v = 0.0;
        vel[(size_t)d * N + i] = 0.0;
      } else if (label[i] == move_group) {
        v = move_velocity[d];
        vel[(size_t)d * N + i] = 0.0;
      } else {
        v = fma(a, half, v);
        vel[(size_t)d * N + i] = v;
      }
      if (is_step1)
        pos[(size_t)d * N + i] = fma(v, dt, pos[(size_t)d * N + i]);
    }
  }
}

void oracle_find_thermo(
  int N, int N_temperature, double volume, const double* mass, const double* pe,
  const double* vel, const double* virial, double* t)
{
  const double K_B = 8.617343e-5; /* utilities/common.cuh:21 */
  const double* vx = vel;
  const double* vy = vel + N;
  const double* vz = vel + 2 * N;
  double ke2 = 0, u = 0, s[6] = {0};
  for (int n = 0; n < N; ++n) {
    ke2 += (vx[n] * vx[n] + vy[n] * vy[n] + vz[n] * vz[n]) * mass[n];
    u += pe[n];
    s[0] += virial[n] + vx[n] * vx[n] * mass[n];
    s[1] += virial[(size_t)N + n] + vy[n] * vy[n] * mass[n];
    s[2] += virial[(size_t)2 * N + n] + vz[n] * vz[n] * mass[n];
    s[3] += virial[(size_t)3 * N + n] + vx[n] * vy[n] * mass[n];
    s[4] += virial[(size_t)4 * N + n] + vx[n] * vz[n] * mass[n];
    s[5] += virial[(size_t)5 * N + n] + vy[n] * vz[n] * mass[n];
  }
  t[0] = ke2 / (3.0 * N_temperature * K_B);
  t[1] = u;
  for (int k = 0; k < 6; ++k)
    t[2 + k] = s[k] / volume;
}
