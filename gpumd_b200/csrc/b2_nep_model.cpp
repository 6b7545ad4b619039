// b2_nep_model.cpp -- nep.txt parser (format: src/force/nep.cu:100-395 of the reference).
#include "b2_nep_model.h"
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace b2 {

// x = hi + lo with hi = x rounded to TF32 (10 explicit mantissa bits, nearest, ties away from
// zero -- what cvt.rna.tf32.f32 does) and lo = tf32(x - hi); same split as b2tc::split_tf32
static void split_tf32(float x, float& hi, float& lo)
{
  auto rna = [](float v) {
    uint32_t u;
    std::memcpy(&u, &v, 4);
    if ((u & 0x7F800000u) != 0x7F800000u)
      u = (u + 0x1000u) & 0xFFFFE000u;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
  };
  hi = rna(x);
  lo = rna(x - hi);
}


// covalent radii used by the typewise ZBL cutoff (values: nep_utilities.cuh:143-154)
const float COVALENT_RADIUS[94] = {
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

static const char* SYMBOLS[94] = {
  "H",  "He", "Li", "Be", "B",  "C",  "N",  "O",  "F",  "Ne", "Na", "Mg", "Al", "Si", "P",  "S",
  "Cl", "Ar", "K",  "Ca", "Sc", "Ti", "V",  "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge",
  "As", "Se", "Br", "Kr", "Rb", "Sr", "Y",  "Zr", "Nb", "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd",
  "In", "Sn", "Sb", "Te", "I",  "Xe", "Cs", "Ba", "La", "Ce", "Pr", "Nd", "Pm", "Sm", "Eu", "Gd",
  "Tb", "Dy", "Ho", "Er", "Tm", "Yb", "Lu", "Hf", "Ta", "W",  "Re", "Os", "Ir", "Pt", "Au", "Hg",
  "Tl", "Pb", "Bi", "Po", "At", "Rn", "Fr", "Ra", "Ac", "Th", "Pa", "U",  "Np", "Pu"};

int atomic_number_of(const std::string& s)
{
  for (int e = 0; e < 94; ++e)
    if (s == SYMBOLS[e])
      return e + 1;
  return 0;
}

static std::vector<std::string> next_tokens(std::ifstream& in)
{
  std::string line;
  std::vector<std::string> t;
  if (!std::getline(in, line))
    return t;
  std::istringstream ss(line);
  std::string w;
  while (ss >> w)
    t.push_back(w);
  return t;
}

static int padded_basis(int k1)
{
  if (k1 <= 9)
    return 9;
  if (k1 <= 13)
    return 13;
  if (k1 <= 17)
    return 17;
  return -1;
}

std::string NepModel::load(const char* path)
{
  std::ifstream in(path);
  if (!in.is_open())
    return std::string("Failed to open ") + path;
  auto tok = next_tokens(in);
  if (tok.size() < 3)
    return "The first line of nep.txt should have at least 3 items.";
  const std::string& head = tok[0];
  if (head == "nep4") {
    version = 4;
  } else if (head == "nep4_zbl") {
    version = 4;
    zbl_enabled = true;
  } else if (head == "nep5") {
    version = 5;
  } else if (head == "nep5_zbl") {
    version = 5;
    zbl_enabled = true;
  } else {
    // dipole / polarizability / temperature models are different Potential subclasses' business
    return head + " is an unsupported NEP model (libb200md handles nep4/nep5[_zbl] potentials).";
  }
  nt = std::atoi(tok[1].c_str());
  if (nt < 1 || nt > 94 || (int)tok.size() != 2 + nt)
    return "The first line of nep.txt should have num_types atom symbols.";
  for (int n = 0; n < nt; ++n) {
    symbols.push_back(tok[2 + n]);
    atomic_numbers.push_back(atomic_number_of(tok[2 + n]));
  }
  if (zbl_enabled) {
    tok = next_tokens(in);
    if (tok.size() != 3 && tok.size() != 4)
      return "This line should be zbl rc_inner rc_outer [zbl_factor].";
    zbl_rc_inner = (float)std::atof(tok[1].c_str());
    zbl_rc_outer = (float)std::atof(tok[2].c_str());
    if (zbl_rc_inner == 0 && zbl_rc_outer == 0) {
      zbl_flexible = true;
    } else if (tok.size() == 4) {
      zbl_typewise_factor = (float)std::atof(tok[3].c_str());
      zbl_typewise = true;
    }
  }
  tok = next_tokens(in);
  if (tok.size() != 5 && (int)tok.size() != nt * 2 + 3)
    return "cutoff should have 4 or num_types * 2 + 2 parameters.";
  rc_radial.resize(nt);
  rc_angular.resize(nt);
  for (int n = 0; n < nt; ++n) {
    const bool per_type = tok.size() != 5;
    rc_radial[n] = (float)std::atof(tok[per_type ? 1 + n * 2 : 1].c_str());
    rc_angular[n] = (float)std::atof(tok[per_type ? 2 + n * 2 : 2].c_str());
    rc_radial_max = std::fmax(rc_radial_max, rc_radial[n]);
    rc_angular_max = std::fmax(rc_angular_max, rc_angular[n]);
  }
  const int mnr = std::atoi(tok[tok.size() - 2].c_str());
  const int mna = std::atoi(tok[tok.size() - 1].c_str());
  if (mnr > 819)
    return "The maximum number of neighbors exceeds 819. Please reduce this value.";
  MN_radial = (int)std::ceil(mnr * 1.25);
  MN_angular = (int)std::ceil(mna * 1.25);

  tok = next_tokens(in);
  if (tok.size() != 3)
    return "This line should be n_max n_max_radial n_max_angular.";
  n_max_radial = std::atoi(tok[1].c_str());
  n_max_angular = std::atoi(tok[2].c_str());
  tok = next_tokens(in);
  if (tok.size() != 3)
    return "This line should be basis_size basis_size_radial basis_size_angular.";
  basis_size_radial = std::atoi(tok[1].c_str());
  basis_size_angular = std::atoi(tok[2].c_str());
  tok = next_tokens(in);
  if (tok.size() < 4)
    return "This line should be l_max l_max_3body has_q_222 has_q_1111 [...].";
  L_max = std::atoi(tok[1].c_str());
  has222 = std::atoi(tok[2].c_str()) ? 1 : 0;
  has1111 = std::atoi(tok[3].c_str()) ? 1 : 0;
  for (size_t k = 4; k < tok.size(); ++k)
    if (std::atoi(tok[k].c_str()) != 0)
      return "has_q_112/123/233/134 invariants are not supported by libb200md yet.";
  if (L_max != 4)
    return "libb200md supports l_max_3body = 4 only.";
  num_L = L_max + has222 + has1111;
  dim_angular = (n_max_angular + 1) * num_L;
  tok = next_tokens(in);
  if (tok.size() != 3)
    return "This line should be ANN num_neurons 0.";
  nneu = std::atoi(tok[1].c_str());
  dim = (n_max_radial + 1) + dim_angular;

  nr1 = n_max_radial + 1;
  na1 = n_max_angular + 1;
  kr1 = basis_size_radial + 1;
  ka1 = basis_size_angular + 1;
  K1R = padded_basis(kr1);
  K1A = padded_basis(ka1);
  if (K1R < 0 || K1A < 0 || kr1 < 2 || ka1 < 2)
    return "basis_size must be between 1 and 16.";
  KP = (K1R + 3) / 4 * 4;
  UST = nt * KP;
  DIMP = (dim + 15) / 16 * 16;
  if (DIMP > 128)
    return "descriptor dimension above 128 is not supported by libb200md.";

  const int ntsq = nt * nt;
  const int num_para_ann =
    version == 4 ? (dim + 2) * nneu * nt + 1 : ((dim + 2) * nneu + 1) * nt + 1;
  const int nbr = nr1 * kr1, nba = na1 * ka1;
  const int num_para = num_para_ann + ntsq * (nbr + nba);
  std::vector<float> para(num_para + dim);
  for (int n = 0; n < num_para + dim; ++n) {
    tok = next_tokens(in);
    if (tok.empty())
      return "nep.txt ended before all parameters were read.";
    para[n] = (float)std::atof(tok[0].c_str());
  }
  if (zbl_flexible) {
    const int nz = nt * (nt + 1) / 2;
    zbl_para.resize(10 * nz);
    for (int d = 0; d < 10 * nz; ++d) {
      tok = next_tokens(in);
      if (tok.empty())
        return "nep.txt ended before the flexible-ZBL table was read.";
      zbl_para[d] = (float)std::atof(tok[0].c_str());
    }
  }

  // ---- ANN weights, per type: w0[nneu][dim], b0[nneu], w1[nneu] (+1 bias for NEP5), then b1 ----
  w0p.assign((size_t)nt * nneu * DIMP, 0.0f);
  b0.assign((size_t)nt * nneu, 0.0f);
  w1.assign((size_t)nt * nneu, 0.0f);
  bias.assign(nt, 0.0f);
  size_t p = 0;
  for (int t = 0; t < nt; ++t) {
    for (int n = 0; n < nneu; ++n)
      for (int d = 0; d < dim; ++d)
        w0p[((size_t)t * nneu + n) * DIMP + d] = para[p++];
    for (int n = 0; n < nneu; ++n)
      b0[(size_t)t * nneu + n] = para[p++];
    for (int n = 0; n < nneu; ++n)
      w1[(size_t)t * nneu + n] = para[p++];
    if (version == 5)
      bias[t] = para[p++];
  }
  const float b1 = para[p++];
  for (int t = 0; t < nt; ++t)
    bias[t] = (version == 5) ? bias[t] + b1 : b1; // nep_utilities.cuh:193, 309
  // ---- expansion coefficients: file order [(n*(K+1)+k)*nt^2 + t1*nt + t2] ----
  c_r.assign((size_t)ntsq * nr1 * K1R, 0.0f);
  c_a.assign((size_t)ntsq * na1 * K1A, 0.0f);
  const float* cr = para.data() + num_para_ann;
  const float* ca = cr + (size_t)ntsq * nbr;
  for (int pair = 0; pair < ntsq; ++pair) {
    for (int n = 0; n < nr1; ++n)
      for (int k = 0; k < kr1; ++k)
        c_r[((size_t)pair * nr1 + n) * K1R + k] = cr[(size_t)(n * kr1 + k) * ntsq + pair];
    for (int n = 0; n < na1; ++n)
      for (int k = 0; k < ka1; ++k)
        c_a[((size_t)pair * na1 + n) * K1A + k] = ca[(size_t)(n * ka1 + k) * ntsq + pair];
  }
  // ---- the same coefficients padded to float4 granularity ----
  {
    const int kqa = (K1A + 3) / 4;
    nqr = (nr1 + 3) / 4;
    c_a4.assign((size_t)ntsq * na1 * kqa * 4, 0.0f);
    c_r4.assign((size_t)ntsq * nqr * K1R * 4, 0.0f);
    for (int pair = 0; pair < ntsq; ++pair) {
      for (int n = 0; n < na1; ++n)
        for (int k = 0; k < K1A; ++k)
          c_a4[((size_t)pair * na1 + n) * kqa * 4 + k] = c_a[((size_t)pair * na1 + n) * K1A + k];
      for (int n = 0; n < nr1; ++n)
        for (int k = 0; k < K1R; ++k)
          c_r4[(((size_t)pair * nqr + n / 4) * K1R + k) * 4 + n % 4] = c_r[((size_t)pair * nr1 + n) * K1R + k];
    }
  }
  // ---- tensor-core images of the hidden layer (layout documented in b2_nep_model.h) ----
  HN = (nneu + 15) / 16 * 16;
  DK = (dim + 7) / 8 * 8;
  DN = (dim + 15) / 16 * 16;
  K3 = (nr1 + 7) / 8 * 8;
  N3 = (nt * KP + 15) / 16 * 16;
  tc3_ok = N3 <= 256 && K3 <= 16; // k_mlp_tc stages the radial dU/dq from one 16-column chunk
  tc_img_floats = 2 * HN * DK + 2 * DN * HN + 2 * HN + (tc3_ok ? 2 * N3 * K3 : 0);
  {
    int cols = 32;
    while (cols < HN + DN || (tc3_ok && cols < N3))
      cols <<= 1;
    int ka = DK > HN ? DK : HN;
    if (tc3_ok && K3 > ka)
      ka = K3;
    const size_t smem = (size_t)tc_img_floats * 4 + 2 * 128 * (size_t)ka * 4 + 128 * (size_t)DK * 4 + 256;
    tc_ok = HN <= 256 && DN <= 256 && cols <= 512 && smem <= 200 * 1024;
  }
  if (tc_ok) {
    tc_img.assign((size_t)nt * tc_img_floats, 0.0f);
    for (int t = 0; t < nt; ++t) {
      float* b1_hi = tc_img.data() + (size_t)t * tc_img_floats;
      float* b1_lo = b1_hi + HN * DK;
      float* b2_hi = b1_lo + HN * DK;
      float* b2_lo = b2_hi + DN * HN;
      float* ib0 = b2_lo + DN * HN;
      float* iw1 = ib0 + HN;
      for (int n = 0; n < nneu; ++n) {
        for (int d = 0; d < dim; ++d) {
          float hi, lo;
          split_tf32(w0p[((size_t)t * nneu + n) * DIMP + d], hi, lo);
          const size_t o1 = (size_t)n * 4 + (size_t)(d / 4) * (HN * 4) + (d % 4); // row n, column d
          const size_t o2 = (size_t)d * 4 + (size_t)(n / 4) * (DN * 4) + (n % 4); // row d, column n
          b1_hi[o1] = hi;
          b1_lo[o1] = lo;
          b2_hi[o2] = hi;
          b2_lo[o2] = lo;
        }
        ib0[n] = b0[(size_t)t * nneu + n];
        iw1[n] = w1[(size_t)t * nneu + n];
      }
      if (tc3_ok) {
        float* b3_hi = iw1 + HN;
        float* b3_lo = b3_hi + N3 * K3;
        for (int t2 = 0; t2 < nt; ++t2)
          for (int n = 0; n < nr1; ++n)
            for (int k = 0; k < K1R; ++k) {
              float hi, lo;
              split_tf32(c_r[((size_t)(t * nt + t2) * nr1 + n) * K1R + k], hi, lo);
              const int row = t2 * KP + k; // column of the U row
              const size_t o = (size_t)row * 4 + (size_t)(n / 4) * (N3 * 4) + (n % 4);
              b3_hi[o] = hi;
              b3_lo[o] = lo;
            }
      }
    }
  }
  q_scaler.assign(DIMP, 0.0f);
  for (int d = 0; d < dim; ++d)
    q_scaler[d] = para[num_para + d];
  // ---- pair cutoffs, computed with the reference's FP32 expressions (nep.cu:473-474,530-531) ----
  rc_r.resize(ntsq);
  rcinv_r.resize(ntsq);
  rc2_r.resize(ntsq);
  rc_a.resize(ntsq);
  rcinv_a.resize(ntsq);
  rc2_a.resize(ntsq);
  for (int t1 = 0; t1 < nt; ++t1)
    for (int t2 = 0; t2 < nt; ++t2) {
      const int pr = t1 * nt + t2;
      volatile float r = (rc_radial[t1] + rc_radial[t2]) * 0.5f;
      volatile float a = (rc_angular[t1] + rc_angular[t2]) * 0.5f;
      rc_r[pr] = r;
      rc_a[pr] = a;
      volatile float ri = 1.0f / r, ai = 1.0f / a, r2 = r * r, a2 = a * a;
      rcinv_r[pr] = ri;
      rcinv_a[pr] = ai;
      rc2_r[pr] = r2;
      rc2_a[pr] = a2;
    }
  return "";
}

} // namespace b2
