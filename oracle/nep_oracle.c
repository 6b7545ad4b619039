/*
 * oracle/nep_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference NEP evaluation:
 *   parser            src/force/nep.cu:100-395, 402-434
 *   neighbour sets    src/force/nep.cu:436-486 (large box), nep_small_box.cuh:56-132 (small box)
 *   descriptor + ANN  src/force/nep.cu:488-659, src/utilities/nep_utilities.cuh:169-194,285-310
 *   radial force      src/force/nep.cu:661-772
 *   angular force     src/force/nep.cu:774-861, src/force/potential.cu:170-297
 *   ZBL               src/force/nep.cu:863-975, nep_utilities.cuh:433-508
 * The pair math lives in nep_oracle_body.inc, compiled once with REAL=float (the GPU
 * reference's arithmetic) and once with REAL=double (a high-precision truth).
 * Forces/virials are assembled in the "scatter" form the reference's small-box kernels use
 * (nep_small_box.cuh:398-618): F_i += f12, F_j -= f12, W_j -= r12 (x) f12, which is the same
 * sum as the large-box gather form (nep.cu:734-753, potential.cu:251-276) re-ordered.
 */
#include "oracle.h"
#include "oracle_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- spherical-harmonic normalisation constants (nep_utilities.cuh:18-46) ---- */
const double ORACLE_C3B[24] = {
  0.238732414637843, 0.119366207318922, 0.119366207318922, 0.099471839432435,
  0.596831036594608, 0.596831036594608, 0.149207759148652, 0.149207759148652,
  0.139260575205408, 0.104445431404056, 0.104445431404056, 1.044454314040563,
  1.044454314040563, 0.174075719006761, 0.174075719006761, 0.011190581936149,
  0.223811638722978, 0.223811638722978, 0.111905819361489, 0.111905819361489,
  1.566681471060845, 1.566681471060845, 0.195835183882606, 0.195835183882606};
const double ORACLE_C4B[5] = {
  -0.007499480826664, -0.134990654879954, 0.067495327439977, 0.404971964639861,
  -0.809943929279723};
const double ORACLE_C5B[3] = {0.026596810706114, 0.053193621412227, 0.026596810706114};

/* covalent radii for the typewise ZBL cutoff (nep_utilities.cuh:143-154) */
const float ORACLE_COVALENT_RADIUS[94] = {
  0.426667f, 0.613333f, 1.6f,     1.25333f, 1.02667f, 1.0f,     0.946667f, 0.84f,    0.853333f,
  0.893333f, 1.86667f,  1.66667f, 1.50667f, 1.38667f, 1.46667f, 1.36f,     1.32f,    1.28f,
  2.34667f,  2.05333f,  1.77333f, 1.62667f, 1.61333f, 1.46667f, 1.42667f,  1.38667f, 1.33333f,
  1.32f,     1.34667f,  1.45333f, 1.49333f, 1.45333f, 1.53333f, 1.46667f,  1.52f,    1.56f,
  2.52f,     2.22667f,  1.96f,    1.85333f, 1.76f,    1.65333f, 1.53333f,  1.50667f, 1.50667f,
  1.44f,     1.53333f,  1.64f,    1.70667f, 1.68f,    1.68f,    1.64f,     1.76f,    1.74667f,
  2.78667f,  2.34667f,  2.16f,    1.96f,    2.10667f, 2.09333f, 2.08f,     2.06667f, 2.01333f,
  2.02667f,  2.01333f,  2.0f,     1.98667f, 1.98667f, 1.97333f, 2.04f,     1.94667f, 1.82667f,
  1.74667f,  1.64f,     1.57333f, 1.54667f, 1.48f,    1.49333f, 1.50667f,  1.76f,    1.73333f,
  1.73333f,  1.81333f,  1.74667f, 1.84f,    1.89333f, 2.68f,    2.41333f,  2.22667f, 2.10667f,
  2.02667f,  2.04f,     2.05333f, 2.06667f};

static const char* ELEMENT_SYMBOLS[94] = {
  "H",  "He", "Li", "Be", "B",  "C",  "N",  "O",  "F",  "Ne", "Na", "Mg", "Al", "Si", "P",  "S",
  "Cl", "Ar", "K",  "Ca", "Sc", "Ti", "V",  "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge",
  "As", "Se", "Br", "Kr", "Rb", "Sr", "Y",  "Zr", "Nb", "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd",
  "In", "Sn", "Sb", "Te", "I",  "Xe", "Cs", "Ba", "La", "Ce", "Pr", "Nd", "Pm", "Sm", "Eu", "Gd",
  "Tb", "Dy", "Ho", "Er", "Tm", "Yb", "Lu", "Hf", "Ta", "W",  "Re", "Os", "Ir", "Pt", "Au", "Hg",
  "Tl", "Pb", "Bi", "Po", "At", "Rn", "Fr", "Ra", "Ac", "Th", "Pa", "U",  "Np", "Pu"};

/* ------------------------------------------------------------------ parser */

static int read_tokens(FILE* f, char tok[][64], int max_tok)
{
  char line[4096];
  if (!fgets(line, sizeof line, f)) {
    return -1;
  }
  int n = 0;
  char* save = NULL;
  for (char* p = strtok_r(line, " \t\r\n", &save); p && n < max_tok;
       p = strtok_r(NULL, " \t\r\n", &save)) {
    strncpy(tok[n], p, 63);
    tok[n][63] = 0;
    n++;
  }
  return n;
}

#define FAIL(...)                 \
  do {                            \
    fprintf(stderr, "oracle: ");  \
    fprintf(stderr, __VA_ARGS__); \
    fprintf(stderr, "\n");        \
    goto fail;                    \
  } while (0)

oracle_nep_model* oracle_nep_load(const char* path)
{
  FILE* f = fopen(path, "r");
  if (!f) {
    fprintf(stderr, "oracle: cannot open %s\n", path);
    return NULL;
  }
  oracle_nep_model* m = (oracle_nep_model*)calloc(1, sizeof *m);
  char tok[200][64];
  int nt = read_tokens(f, tok, 200);
  if (nt < 3)
    FAIL("first line needs >= 3 items");
  /* nep.cu:113-143 : accepted headers (potential models only here) */
  if (!strcmp(tok[0], "nep4")) {
    m->version = 4;
  } else if (!strcmp(tok[0], "nep4_zbl")) {
    m->version = 4;
    m->zbl_enabled = 1;
  } else if (!strcmp(tok[0], "nep5")) {
    m->version = 5;
  } else if (!strcmp(tok[0], "nep5_zbl")) {
    m->version = 5;
    m->zbl_enabled = 1;
  } else if (!strcmp(tok[0], "nep3")) { /* same maths, shared ANN; used by old fixtures only */
    m->version = 3;
  } else if (!strcmp(tok[0], "nep3_zbl")) {
    m->version = 3;
    m->zbl_enabled = 1;
  } else
    FAIL("unsupported NEP model %s", tok[0]);
  m->num_types = atoi(tok[1]);
  if (nt != 2 + m->num_types || m->num_types < 1 || m->num_types > ORACLE_MAX_TYPES)
    FAIL("bad type count on first line");
  for (int n = 0; n < m->num_types; ++n) { /* nep.cu:157-167 */
    int z = 0;
    for (int e = 0; e < 94; ++e) {
      if (!strcmp(tok[2 + n], ELEMENT_SYMBOLS[e])) {
        z = e + 1;
        break;
      }
    }
    m->atomic_numbers[n] = z;
  }
  m->typewise_cutoff_zbl_factor = 0.65f;
  if (m->zbl_enabled) { /* nep.cu:170-194 */
    nt = read_tokens(f, tok, 200);
    if (nt != 3 && nt != 4)
      FAIL("zbl line");
    m->zbl_rc_inner = (float)atof(tok[1]);
    m->zbl_rc_outer = (float)atof(tok[2]);
    if (m->zbl_rc_inner == 0 && m->zbl_rc_outer == 0) {
      m->zbl_flexible = 1;
    } else if (nt == 4) {
      m->typewise_cutoff_zbl_factor = (float)atof(tok[3]);
      m->use_typewise_cutoff_zbl = 1;
    }
  }
  nt = read_tokens(f, tok, 200); /* cutoff, nep.cu:197-237 */
  if (nt != 5 && nt != m->num_types * 2 + 3)
    FAIL("cutoff line");
  if (nt == 5) {
    for (int n = 0; n < m->num_types; ++n) {
      m->rc_radial[n] = (float)atof(tok[1]);
      m->rc_angular[n] = (float)atof(tok[2]);
    }
  } else {
    for (int n = 0; n < m->num_types; ++n) {
      m->rc_radial[n] = (float)atof(tok[1 + n * 2]);
      m->rc_angular[n] = (float)atof(tok[2 + n * 2]);
    }
  }
  for (int n = 0; n < m->num_types; ++n) {
    if (m->rc_radial[n] > m->rc_radial_max)
      m->rc_radial_max = m->rc_radial[n];
    if (m->rc_angular[n] > m->rc_angular_max)
      m->rc_angular_max = m->rc_angular[n];
  }
  m->MN_radial = (int)ceil(atoi(tok[nt - 2]) * 1.25);
  m->MN_angular = (int)ceil(atoi(tok[nt - 1]) * 1.25);

  nt = read_tokens(f, tok, 200); /* n_max */
  if (nt != 3)
    FAIL("n_max line");
  m->n_max_radial = atoi(tok[1]);
  m->n_max_angular = atoi(tok[2]);
  nt = read_tokens(f, tok, 200); /* basis_size */
  if (nt != 3)
    FAIL("basis_size line");
  m->basis_size_radial = atoi(tok[1]);
  m->basis_size_angular = atoi(tok[2]);
  nt = read_tokens(f, tok, 200); /* l_max, nep.cu:263-312 */
  if (nt < 4)
    FAIL("l_max line");
  m->L_max = atoi(tok[1]);
  m->has_q_222 = atoi(tok[2]);
  m->has_q_1111 = atoi(tok[3]);
  for (int k = 4; k < nt; ++k) {
    if (atoi(tok[k]) != 0)
      FAIL("has_q_112/123/233/134 invariants are outside this oracle's scope");
  }
  if (m->L_max < 1 || m->L_max > 4)
    FAIL("L_max must be 1..4 in this oracle");
  m->num_L = m->L_max + (m->has_q_222 ? 1 : 0) + (m->has_q_1111 ? 1 : 0);
  m->dim_angular = (m->n_max_angular + 1) * m->num_L;
  nt = read_tokens(f, tok, 200); /* ANN */
  if (nt != 3)
    FAIL("ANN line");
  m->num_neurons = atoi(tok[1]);
  m->dim = (m->n_max_radial + 1) + m->dim_angular;
  m->num_types_sq = m->num_types * m->num_types;
  /* nep.cu:332-350 (NEP3 shared one ANN between all types) */
  if (m->version == 3) {
    m->num_para_ann = (m->dim + 2) * m->num_neurons + 1;
  } else if (m->version == 4) {
    m->num_para_ann = (m->dim + 2) * m->num_neurons * m->num_types + 1;
  } else {
    m->num_para_ann = ((m->dim + 2) * m->num_neurons + 1) * m->num_types + 1;
  }
  int nbr = (m->n_max_radial + 1) * (m->basis_size_radial + 1);
  int nba = (m->n_max_angular + 1) * (m->basis_size_angular + 1);
  m->num_c_radial = m->num_types_sq * nbr;
  int num_para_descriptor = m->num_types_sq * (nbr + nba);
  m->num_para = m->num_para_ann + num_para_descriptor;
  m->parameters = (float*)malloc(sizeof(float) * (m->num_para + m->dim));
  for (int n = 0; n < m->num_para + m->dim; ++n) {
    nt = read_tokens(f, tok, 200);
    if (nt < 1)
      FAIL("parameter %d missing", n);
    m->parameters[n] = (float)atof(tok[0]);
  }
  /* pointers, nep.cu:402-434 */
  float* p = m->parameters;
  for (int t = 0; t < m->num_types; ++t) {
    if (t > 0 && m->version == 3) {
      p = m->parameters;
    }
    m->w0[t] = p;
    p += m->num_neurons * m->dim;
    m->b0[t] = p;
    p += m->num_neurons;
    m->w1[t] = p;
    p += m->num_neurons;
    if (m->version == 5)
      p += 1;
  }
  m->b1 = p;
  p += 1;
  m->c = p;
  m->q_scaler = m->parameters + m->num_para;
  if (m->zbl_flexible) { /* nep.cu:370-377 */
    int nz = m->num_types * (m->num_types + 1) / 2;
    for (int d = 0; d < 10 * nz; ++d) {
      nt = read_tokens(f, tok, 200);
      if (nt < 1)
        FAIL("zbl parameter %d missing", d);
      m->zbl_para[d] = (float)atof(tok[0]);
    }
  }
  fclose(f);
  return m;
fail:
  fclose(f);
  oracle_nep_free(m);
  return NULL;
}

void oracle_nep_free(oracle_nep_model* m)
{
  if (m) {
    free(m->parameters);
    free(m);
  }
}

int oracle_nep_info(const oracle_nep_model* m, int what)
{
  switch (what) {
    case 0: return m->num_types;
    case 1: return m->dim;
    case 2: return m->num_neurons;
    case 3: return m->n_max_radial;
    case 4: return m->n_max_angular;
    case 5: return m->basis_size_radial;
    case 6: return m->basis_size_angular;
    case 7: return m->L_max;
    case 8: return m->num_L;
    case 9: return m->MN_radial;
    case 10: return m->MN_angular;
    case 11: return m->zbl_enabled;
    case 12: return m->version;
    default: return -1;
  }
}
double oracle_nep_rc_radial_max(const oracle_nep_model* m) { return m->rc_radial_max; }
double oracle_nep_rc_angular_max(const oracle_nep_model* m) { return m->rc_angular_max; }

/* ------------------------------------------------------------------ box helpers */

void oracle_box_init(oracle_box* b, const double h[9], const int pbc[3])
{
  for (int d = 0; d < 9; ++d)
    b->h[d] = h[d];
  for (int d = 0; d < 3; ++d)
    b->pbc[d] = pbc[d];
  double* c = b->h;
  /* Box::get_inverse, src/model/box.cu:60-80 */
  c[9] = c[4] * c[8] - c[5] * c[7];
  c[10] = c[2] * c[7] - c[1] * c[8];
  c[11] = c[1] * c[5] - c[2] * c[4];
  c[12] = c[5] * c[6] - c[3] * c[8];
  c[13] = c[0] * c[8] - c[2] * c[6];
  c[14] = c[2] * c[3] - c[0] * c[5];
  c[15] = c[3] * c[7] - c[4] * c[6];
  c[16] = c[1] * c[6] - c[0] * c[7];
  c[17] = c[0] * c[4] - c[1] * c[3];
  double det = c[0] * (c[4] * c[8] - c[5] * c[7]) + c[1] * (c[5] * c[6] - c[3] * c[8]) +
               c[2] * (c[3] * c[7] - c[4] * c[6]);
  for (int n = 9; n < 18; ++n)
    c[n] /= det;
  b->volume = fabs(det);
  /* Box::get_area / thickness, src/model/box.cu:26-58,93-103 */
  for (int d = 0; d < 3; ++d) {
    int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
    double a[3] = {c[d1], c[d1 + 3], c[d1 + 6]};
    double e[3] = {c[d2], c[d2 + 3], c[d2 + 6]};
    double s1 = a[1] * e[2] - a[2] * e[1];
    double s2 = a[2] * e[0] - a[0] * e[2];
    double s3 = a[0] * e[1] - a[1] * e[0];
    b->thickness[d] = b->volume / sqrt(s1 * s1 + s2 * s2 + s3 * s3);
  }
  /* Box::set_is_orthogonal, src/model/box.cu:111-117 */
  b->is_orthogonal =
    c[1] == 0 && c[2] == 0 && c[3] == 0 && c[5] == 0 && c[6] == 0 && c[7] == 0;
  for (int d = 0; d < 18; ++d)
    b->hf[d] = (float)c[d];
}

/* float minimum image, src/model/box.cuh:84-129 */
void oracle_mic_f32(const oracle_box* b, float* x, float* y, float* z)
{
  const float* h = b->hf;
  if (b->is_orthogonal) {
    float lx2 = h[0] * 0.5f, ly2 = h[4] * 0.5f, lz2 = h[8] * 0.5f;
    if (b->pbc[0]) {
      if (*x < -lx2)
        *x += h[0];
      else if (*x > +lx2)
        *x -= h[0];
    }
    if (b->pbc[1]) {
      if (*y < -ly2)
        *y += h[4];
      else if (*y > +ly2)
        *y -= h[4];
    }
    if (b->pbc[2]) {
      if (*z < -lz2)
        *z += h[8];
      else if (*z > +lz2)
        *z -= h[8];
    }
  } else {
    /* fma nesting nvcc's default contraction gives `a*x + b*y + c*z`: fma(c,z, fma(a,x, b*y))
       (same DAG shape as oracle_d2_f32; checked on PTX) */
    float sx = fmaf(h[11], *z, fmaf(h[9], *x, h[10] * *y));
    float sy = fmaf(h[14], *z, fmaf(h[12], *x, h[13] * *y));
    float sz = fmaf(h[17], *z, fmaf(h[15], *x, h[16] * *y));
    if (b->pbc[0])
      sx -= nearbyintf(sx);
    if (b->pbc[1])
      sy -= nearbyintf(sy);
    if (b->pbc[2])
      sz -= nearbyintf(sz);
    *x = fmaf(h[2], sz, fmaf(h[0], sx, h[1] * sy));
    *y = fmaf(h[5], sz, fmaf(h[3], sx, h[4] * sy));
    *z = fmaf(h[8], sz, fmaf(h[6], sx, h[7] * sy));
  }
}

/* double minimum image, src/model/box.cuh:37-82 */
void oracle_mic_f64(const oracle_box* b, double* x, double* y, double* z)
{
  const double* h = b->h;
  if (b->is_orthogonal) {
    if (b->pbc[0]) {
      if (*x < -h[0] * 0.5)
        *x += h[0];
      else if (*x > +h[0] * 0.5)
        *x -= h[0];
    }
    if (b->pbc[1]) {
      if (*y < -h[4] * 0.5)
        *y += h[4];
      else if (*y > +h[4] * 0.5)
        *y -= h[4];
    }
    if (b->pbc[2]) {
      if (*z < -h[8] * 0.5)
        *z += h[8];
      else if (*z > +h[8] * 0.5)
        *z -= h[8];
    }
  } else {
    double sx = h[9] * *x + h[10] * *y + h[11] * *z;
    double sy = h[12] * *x + h[13] * *y + h[14] * *z;
    double sz = h[15] * *x + h[16] * *y + h[17] * *z;
    if (b->pbc[0])
      sx -= nearbyint(sx);
    if (b->pbc[1])
      sy -= nearbyint(sy);
    if (b->pbc[2])
      sz -= nearbyint(sz);
    *x = h[0] * sx + h[1] * sy + h[2] * sz;
    *y = h[3] * sx + h[4] * sy + h[5] * sz;
    *z = h[6] * sx + h[7] * sy + h[8] * sz;
  }
}

/* d2 exactly as nvcc contracts `x12*x12 + y12*y12 + z12*z12` with its default -fmad=true:
 * fma(z,z, fma(x,x, y*y)).  Verified on the PTX nvcc 12.9 emits for
 * src/force/neighbor.cu:146-150 / src/force/nep.cu:467-471 ("FP32 membership"). */
float oracle_d2_f32(float x, float y, float z) { return fmaf(z, z, fmaf(x, x, y * y)); }

/* ------------------------------------------------------------------ candidate cells */

/* Bins of perpendicular width >= rc in fractional space; atoms binned by floor(s*nb) with
 * periodic wrap.  Only used to enumerate CANDIDATE pairs; membership is decided by the
 * reference's FP32 test. */
int oracle_cells_build(oracle_cells* c, const oracle_box* b, int N, const double* pos, double rc)
{
  const double* x = pos;
  const double* y = pos + N;
  const double* z = pos + 2 * N;
  for (int d = 0; d < 3; ++d) {
    int nb = (int)floor(b->thickness[d] / rc);
    if (nb < 1)
      nb = 1;
    if (nb > 1024)
      nb = 1024;
    c->nb[d] = nb;
  }
  int ncell = c->nb[0] * c->nb[1] * c->nb[2];
  c->start = (int*)calloc((size_t)ncell + 1, sizeof(int));
  c->items = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
  c->cell_of = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
  if (!c->start || !c->items || !c->cell_of)
    return -1;
  for (int i = 0; i < N; ++i) {
    double s[3];
    s[0] = b->h[9] * x[i] + b->h[10] * y[i] + b->h[11] * z[i];
    s[1] = b->h[12] * x[i] + b->h[13] * y[i] + b->h[14] * z[i];
    s[2] = b->h[15] * x[i] + b->h[16] * y[i] + b->h[17] * z[i];
    int id[3];
    for (int d = 0; d < 3; ++d) {
      int k = (int)floor(s[d] * c->nb[d]);
      if (b->pbc[d]) {
        k %= c->nb[d];
        if (k < 0)
          k += c->nb[d];
      } else {
        if (k < 0)
          k = 0;
        if (k >= c->nb[d])
          k = c->nb[d] - 1;
      }
      id[d] = k;
    }
    int cell = id[0] + c->nb[0] * (id[1] + c->nb[1] * id[2]);
    c->cell_of[i] = cell;
    c->start[cell + 1]++;
  }
  for (int k = 0; k < ncell; ++k)
    c->start[k + 1] += c->start[k];
  int* fill = (int*)calloc((size_t)ncell, sizeof(int));
  for (int i = 0; i < N; ++i) { /* ascending i inside every cell */
    int cell = c->cell_of[i];
    c->items[c->start[cell] + fill[cell]++] = i;
  }
  free(fill);
  return 0;
}

void oracle_cells_free(oracle_cells* c)
{
  free(c->start);
  free(c->items);
  free(c->cell_of);
}

/* unique neighbouring cells (|offset|<=1 per direction, wrapped when periodic) of a cell */
int oracle_cells_around(const oracle_cells* c, const oracle_box* b, int cell, int out[27])
{
  int id[3] = {cell % c->nb[0], (cell / c->nb[0]) % c->nb[1], cell / (c->nb[0] * c->nb[1])};
  int lst[3][3], cnt[3];
  for (int d = 0; d < 3; ++d) {
    cnt[d] = 0;
    for (int o = -1; o <= 1; ++o) {
      int k = id[d] + o;
      if (b->pbc[d]) {
        k %= c->nb[d];
        if (k < 0)
          k += c->nb[d];
      } else if (k < 0 || k >= c->nb[d]) {
        continue;
      }
      int dup = 0;
      for (int q = 0; q < cnt[d]; ++q)
        dup |= (lst[d][q] == k);
      if (!dup)
        lst[d][cnt[d]++] = k;
    }
  }
  int n = 0;
  for (int a = 0; a < cnt[2]; ++a)
    for (int e = 0; e < cnt[1]; ++e)
      for (int g = 0; g < cnt[0]; ++g)
        out[n++] = lst[0][g] + c->nb[0] * (lst[1][e] + c->nb[1] * lst[2][a]);
  return n;
}

static int cmp_int(const void* a, const void* b)
{
  int x = *(const int*)a, y = *(const int*)b;
  return (x > y) - (x < y);
}

int oracle_neighbor_list(
  int N, const double h[9], const int pbc[3], const double* pos, double rc, int* NN, int* NL,
  int mn)
{
  oracle_box b;
  oracle_box_init(&b, h, pbc);
  for (int d = 0; d < 3; ++d) {
    if (pbc[d] && b.thickness[d] < 2.0 * rc) {
      fprintf(stderr, "oracle_neighbor_list: box too thin for minimum image\n");
      return -2;
    }
  }
  oracle_cells c;
  if (oracle_cells_build(&c, &b, N, pos, rc) != 0)
    return -1;
  const double* x = pos;
  const double* y = pos + N;
  const double* z = pos + 2 * N;
  const float rc2 = (float)(rc * rc); /* neighbor.cu:341 passes rc*rc (double) as a float arg */
  int status = 0;
  for (int i = 0; i < N && status == 0; ++i) {
    int around[27];
    int na = oracle_cells_around(&c, &b, c.cell_of[i], around);
    int cnt = 0;
    for (int a = 0; a < na; ++a) {
      for (int k = c.start[around[a]]; k < c.start[around[a] + 1]; ++k) {
        int j = c.items[k];
        if (j == i)
          continue;
        float x12 = (float)(x[j] - x[i]); /* neighbor.cu:146-148 */
        float y12 = (float)(y[j] - y[i]);
        float z12 = (float)(z[j] - z[i]);
        oracle_mic_f32(&b, &x12, &y12, &z12);
        if (oracle_d2_f32(x12, y12, z12) < rc2) {
          if (cnt >= mn) {
            status = -3;
            break;
          }
          NL[(size_t)i * mn + cnt++] = j;
        }
      }
    }
    qsort(NL + (size_t)i * mn, cnt, sizeof(int), cmp_int); /* neighbor.cuh:112-136 */
    NN[i] = cnt;
  }
  oracle_cells_free(&c);
  return status;
}

/* ------------------------------------------------------------------ two precisions */

#define REAL float
#define SUFFIX(name) name##_f32
#define R_SQRT sqrtf
#define R_COS cosf
#define R_SIN sinf
#define R_TANH tanhf
#define R_EXP expf
#define R_POW powf
#define ORACLE_F32 1
#include "nep_oracle_body.inc"
#undef REAL
#undef SUFFIX
#undef R_SQRT
#undef R_COS
#undef R_SIN
#undef R_TANH
#undef R_EXP
#undef R_POW
#undef ORACLE_F32

#define REAL double
#define SUFFIX(name) name##_f64
#define R_SQRT sqrt
#define R_COS cos
#define R_SIN sin
#define R_TANH tanh
#define R_EXP exp
#define R_POW pow
#define ORACLE_F32 0
#include "nep_oracle_body.inc"

int oracle_nep_compute(
  const oracle_nep_model* m, int precision, int N, const int* type, const double h[9],
  const int pbc[3], const double* position, double* pe, double* force, double* virial,
  double* q, int* NN_r, int* NL_r, int mn_r, int* NN_a, int* NL_a, int mn_a)
{
  if (precision == 32)
    return nep_compute_f32(
      m, N, type, h, pbc, position, pe, force, virial, q, NN_r, NL_r, mn_r, NN_a, NL_a, mn_a);
  if (precision == 64)
    return nep_compute_f64(
      m, N, type, h, pbc, position, pe, force, virial, q, NN_r, NL_r, mn_r, NN_a, NL_a, mn_a);
  return -10;
}
