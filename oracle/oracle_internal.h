/* oracle/oracle_internal.h -- TEST INFRASTRUCTURE ONLY (see oracle.h). */
#ifndef B200MD_ORACLE_INTERNAL_H
#define B200MD_ORACLE_INTERNAL_H

#include "oracle.h"
#include <stddef.h>

struct oracle_nep_model {
  int version, num_types, num_types_sq;
  int zbl_enabled, zbl_flexible, use_typewise_cutoff_zbl;
  float typewise_cutoff_zbl_factor, zbl_rc_inner, zbl_rc_outer;
  float zbl_para[550];
  int atomic_numbers[ORACLE_MAX_TYPES];
  float rc_radial[ORACLE_MAX_TYPES], rc_angular[ORACLE_MAX_TYPES];
  float rc_radial_max, rc_angular_max;
  int MN_radial, MN_angular;
  int n_max_radial, n_max_angular, basis_size_radial, basis_size_angular;
  int L_max, has_q_222, has_q_1111, num_L, dim_angular, dim, num_neurons;
  int num_para_ann, num_para, num_c_radial;
  float* parameters;
  const float *w0[ORACLE_MAX_TYPES], *b0[ORACLE_MAX_TYPES], *w1[ORACLE_MAX_TYPES];
  const float *b1, *c, *q_scaler;
};

typedef struct {
  double h[18];
  float hf[18];
  int pbc[3];
  int is_orthogonal;
  double thickness[3];
  double volume;
} oracle_box;

typedef struct {
  int nb[3];
  int* start;   /* ncell+1 */
  int* items;   /* N, ascending atom index inside each cell */
  int* cell_of; /* N */
} oracle_cells;

extern const double ORACLE_C3B[24];
extern const double ORACLE_C4B[5];
extern const double ORACLE_C5B[3];
extern const float ORACLE_COVALENT_RADIUS[94];

void oracle_box_init(oracle_box* b, const double h[9], const int pbc[3]);
void oracle_mic_f32(const oracle_box* b, float* x, float* y, float* z);
void oracle_mic_f64(const oracle_box* b, double* x, double* y, double* z);
float oracle_d2_f32(float x, float y, float z);
int oracle_cells_build(oracle_cells* c, const oracle_box* b, int N, const double* pos, double rc);
void oracle_cells_free(oracle_cells* c);
int oracle_cells_around(const oracle_cells* c, const oracle_box* b, int cell, int out[27]);

#endif
