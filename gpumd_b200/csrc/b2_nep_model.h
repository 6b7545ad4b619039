// b2_nep_model.h -- host-side NEP model: nep.txt parser and the flattened tables the kernels use.
// File format and parameter order follow the reference's NEP::NEP / update_potential
// (src/force/nep.cu:100-434); the table layouts are this library's own (see B2NepView).
#pragma once
#include <string>
#include <vector>

namespace b2 {

struct NepModel {
  int version = 4;
  int nt = 0;
  std::vector<std::string> symbols;
  std::vector<int> atomic_numbers;
  bool zbl_enabled = false, zbl_flexible = false, zbl_typewise = false;
  float zbl_rc_inner = 0.0f, zbl_rc_outer = 0.0f, zbl_typewise_factor = 0.65f;
  std::vector<float> zbl_para;
  std::vector<float> rc_radial, rc_angular; // per type
  float rc_radial_max = 0.0f, rc_angular_max = 0.0f;
  int MN_radial = 0, MN_angular = 0; // enlarged by 1.25 (nep.cu:234-235)
  int n_max_radial = 0, n_max_angular = 0, basis_size_radial = 0, basis_size_angular = 0;
  int L_max = 0, has222 = 0, has1111 = 0, num_L = 0, dim_angular = 0, dim = 0, nneu = 0;

  // derived sizes
  int nr1 = 0, na1 = 0, kr1 = 0, ka1 = 0, K1R = 0, K1A = 0, KP = 0, UST = 0, DIMP = 0;

  // flattened tables (see B2NepView for layouts)
  std::vector<float> rc_r, rcinv_r, rc2_r, rc_a, rcinv_a, rc2_a; // [nt*nt]
  std::vector<float> c_r, c_a, w0p, b0, w1, bias, q_scaler;
  // c_a / c_r padded for 128-bit loads: c_a4 [pair][n][(K1A+3)/4][4], c_r4 [pair][nq][K1R][4]
  // (n = 4*nq + lane of the vector), see B2NepView::c_a4 / c_r4
  std::vector<float> c_a4, c_r4;
  int nqr = 0;

  // tensor-core hidden layer (k_mlp_tc): per type one ready-to-copy shared-memory image
  //   [B1_hi | B1_lo | B2_hi | B2_lo | b0 | w1 | B3_hi | B3_lo]  (tc_img_floats floats, multiple of 4)
  // B1 = W0 as the [N = HN x K = DK] operand of Z = Q . W0^T, B2 = W0^T as the [N = DN x K = HN]
  // operand of dU/dq = C . W0, both split into TF32 hi/lo parts and stored K-major in 8 x 16-byte
  // core matrices: float offset of (row r, column k) = r*4 + (k/4)*(rows*4) + (k%4).
  // A third operand pair B3 = the radial expansion coefficients of the type, [N3 = nt*KP rows x
  // K3 = nr1 columns] with row t2*KP + k = c[t,t2,n,k]: U = FpR . B3^T, the pre-contracted radial
  // table, comes out of the same kernel (tc3_ok; otherwise k_utable does it).
  int HN = 0, DK = 0, DN = 0, K3 = 0, N3 = 0, tc_img_floats = 0;
  bool tc_ok = false;  // shapes fit the tcgen05 kernel (else the SIMT k_mlp is the only path)
  bool tc3_ok = false; // nt*KP <= 256: the U table is a third GEMM of k_mlp_tc
  std::vector<float> tc_img; // [nt][tc_img_floats]

  // returns empty string on success, else the error message
  std::string load(const char* path);
};

extern const float COVALENT_RADIUS[94];
int atomic_number_of(const std::string& symbol); // 0 if unknown

} // namespace b2
