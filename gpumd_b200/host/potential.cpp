// potential.cpp -- NEP_B200 / LJ_B200 / Tersoff1989_B200 / EAM_B200: forward the Potential virtuals to the C-ABI.
// Error convention of the reference: print and exit(1) (src/utilities/error.cuh:22-62).
#include "potential.h"
#include <cstdio>
#include <cstdlib>

void b2h_fail(const char* where)
{
  fprintf(stderr, "Error in %s:\n    %s\n", where, b200md_last_error());
  exit(1);
}

NEP_B200::NEP_B200(const char* file_potential, const int num_atoms)
{
  if (b200md_nep_create(file_potential, num_atoms, &handle_) != B200MD_OK)
    b2h_fail("NEP_B200");
  N1 = 0;
  N2 = num_atoms;
  num_atoms_ = num_atoms;
  nep_model_type = 0; // "potential" model (potential.cuh:33-34)
  rc = b200md_nep_rc(handle_);
  printf("Use the b200md NEP backend with %d atom type(s), D = %d, %d neurons.\n",
         b200md_nep_info(handle_, 0), b200md_nep_info(handle_, 1), b200md_nep_info(handle_, 2));
}

NEP_B200::~NEP_B200() { b200md_nep_destroy(handle_); }

void NEP_B200::compute(
  Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
  GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial)
{
  int pbc[3];
  b2h_pbc(box, pbc);
  // legacy default stream, like every reference kernel
  if (b200md_nep_compute(
        handle_, (int)type.size(), box.cpu_h, pbc, type.data(), position.data(), potential.data(),
        force.data(), virial.data(), nullptr) != B200MD_OK)
    b2h_fail("NEP_B200::compute");
  lists_current_ = false;
  // the reference looks at its neighbour counts every 1000 calls (nep.cu:1014-1034); same cadence
  // for the latched capacity check, which needs a synchronisation
  if (++num_calls_ % 1000 == 1)
    check();
}

void NEP_B200::set_accumulate(bool accumulate)
{
  if (b200md_nep_set_accumulate(handle_, accumulate ? 1 : 0) != B200MD_OK)
    b2h_fail("NEP_B200::set_accumulate");
}

void NEP_B200::export_radial()
{
  if (lists_current_)
    return;
  const int n = num_atoms_, mn = b200md_nep_info(handle_, 3);
  if ((int)NN_radial_.size() != n) {
    NN_radial_.resize(n);
    NL_radial_.resize((size_t)n * mn);
    row_major_.resize((size_t)n * mn);
  }
  // the C-ABI exports row-major rows in caller indices; the reference layout is column-major
  if (b200md_nep_export_neighbors(
        handle_, mn, NN_radial_.data(), row_major_.data(), 0, nullptr, nullptr, nullptr) != B200MD_OK ||
      b200md_transpose_int(n, mn, row_major_.data(), NL_radial_.data(), nullptr) != B200MD_OK)
    b2h_fail("NEP_B200::export_radial");
  lists_current_ = true;
}

const GPU_Vector<int>& NEP_B200::get_NN_radial_ptr()
{
  export_radial();
  return NN_radial_;
}

const GPU_Vector<int>& NEP_B200::get_NL_radial_ptr()
{
  export_radial();
  return NL_radial_;
}

int NEP_B200::type_of(const std::string& symbol) const
{
  for (int t = 0; t < b200md_nep_info(handle_, 0); ++t)
    if (symbol == b200md_nep_symbol(handle_, t))
      return t;
  return -1;
}

void NEP_B200::check()
{
  if (b200md_nep_check(handle_, nullptr) != B200MD_OK)
    b2h_fail("NEP_B200::check");
}

LJ_B200::LJ_B200(const char* file_potential, const int num_atoms)
{
  if (b200md_lj_create(file_potential, num_atoms, &handle_) != B200MD_OK)
    b2h_fail("LJ_B200");
  N1 = 0;
  N2 = num_atoms;
  rc = b200md_lj_rc(handle_);
  printf("Use the b200md LJ backend with %d atom type(s).\n", b200md_lj_info(handle_, 0));
}

LJ_B200::~LJ_B200() { b200md_lj_destroy(handle_); }

void LJ_B200::compute(
  Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
  GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial)
{
  int pbc[3];
  b2h_pbc(box, pbc);
  if (b200md_lj_compute(
        handle_, (int)type.size(), box.cpu_h, pbc, type.data(), position.data(), potential.data(),
        force.data(), virial.data(), nullptr) != B200MD_OK)
    b2h_fail("LJ_B200::compute");
  if (++num_calls_ % 1000 == 1)
    check(); // latched neighbour-capacity errors, the cadence NEP_B200 uses
}

int LJ_B200::type_of(const std::string& symbol) const
{
  for (int t = 0; t < b200md_lj_info(handle_, 0); ++t)
    if (symbol == b200md_lj_symbol(handle_, t))
      return t;
  return -1;
}

void LJ_B200::check()
{
  if (b200md_lj_check(handle_, nullptr) != B200MD_OK)
    b2h_fail("LJ_B200::check");
}

Tersoff1989_B200::Tersoff1989_B200(const char* file_potential, const int num_atoms)
{
  if (b200md_tersoff_create(file_potential, num_atoms, &handle_) != B200MD_OK)
    b2h_fail("Tersoff1989_B200");
  N1 = 0;
  N2 = num_atoms;
  rc = b200md_tersoff_rc(handle_);
  printf("Use the b200md Tersoff-1989 backend with %d atom type(s).\n", b200md_tersoff_info(handle_, 0));
}

Tersoff1989_B200::~Tersoff1989_B200() { b200md_tersoff_destroy(handle_); }

void Tersoff1989_B200::compute(
  Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
  GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial)
{
  int pbc[3];
  b2h_pbc(box, pbc);
  if (b200md_tersoff_compute(
        handle_, (int)type.size(), box.cpu_h, pbc, type.data(), position.data(), potential.data(),
        force.data(), virial.data(), nullptr) != B200MD_OK)
    b2h_fail("Tersoff1989_B200::compute");
  if (++num_calls_ % 1000 == 1)
    check(); // latched neighbour-capacity errors, the cadence NEP_B200 uses
}

int Tersoff1989_B200::type_of(const std::string& symbol) const
{
  for (int t = 0; t < b200md_tersoff_info(handle_, 0); ++t)
    if (symbol == b200md_tersoff_symbol(handle_, t))
      return t;
  return -1;
}

void Tersoff1989_B200::check()
{
  if (b200md_tersoff_check(handle_, nullptr) != B200MD_OK)
    b2h_fail("Tersoff1989_B200::check");
}

EAM_B200::EAM_B200(const char* file_potential, const int num_atoms)
{
  if (b200md_eam_create(file_potential, num_atoms, &handle_) != B200MD_OK)
    b2h_fail("EAM_B200");
  N1 = 0;
  N2 = num_atoms;
  rc = b200md_eam_rc(handle_);
  printf("Use the b200md EAM backend with %d atom type(s).\n", b200md_eam_info(handle_, 0));
}

EAM_B200::~EAM_B200() { b200md_eam_destroy(handle_); }

void EAM_B200::compute(
  Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
  GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial)
{
  int pbc[3];
  b2h_pbc(box, pbc);
  if (b200md_eam_compute(
        handle_, (int)type.size(), box.cpu_h, pbc, type.data(), position.data(), potential.data(),
        force.data(), virial.data(), nullptr) != B200MD_OK)
    b2h_fail("EAM_B200::compute");
  if (++num_calls_ % 1000 == 1)
    check(); // latched neighbour-capacity errors, the cadence NEP_B200 uses
}

int EAM_B200::type_of(const std::string& symbol) const
{
  for (int t = 0; t < b200md_eam_info(handle_, 0); ++t)
    if (symbol == b200md_eam_symbol(handle_, t))
      return t;
  return -1;
}

void EAM_B200::check()
{
  if (b200md_eam_check(handle_, nullptr) != B200MD_OK)
    b2h_fail("EAM_B200::check");
}
