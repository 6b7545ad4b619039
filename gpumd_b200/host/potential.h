// potential.h -- GPUMD's Potential plugin interface (src/force/potential.cuh:20-113, the members and
// the pure virtual the run loop uses) and the adapters that forward it to libb200md's C-ABI.
// Inside the reference tree these adapters derive from the reference's own class Potential
// (INTEGRATION.md); here a same-shaped base keeps the standalone executable self-contained.
#pragma once
#include "../../include/b200md.h"
#include <string>

#if defined(B200MD_IN_GPUMD)
// Compiled INSIDE the reference tree (oracle/Makefile.gpumd_b200, -I<reference>/src): the adapters
// derive from the reference's own class Potential and use its Box / GPU_Vector.
#include "force/potential.cuh"
#define B2H_EXTRA_VIRTUAL
#define B2H_EXTRA_OVERRIDE
#else
// Standalone driver (gpumd_b200/b200md): a same-shaped base keeps it self-contained.
#include "model.h"
#define B2H_EXTRA_VIRTUAL virtual
#define B2H_EXTRA_OVERRIDE override

class Potential
{
public:
  int N1 = 0;
  int N2 = 0;
  double rc = 0.0;
  int nep_model_type = -1;
  int ilp_flag = 0;
  virtual ~Potential() = default;
  virtual void compute(
    Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
    GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial) = 0;
  // the radial neighbour list of the last compute (potential.cuh:66-77; used by
  // ensemble_ti_liquid.cu:394-396): counts [N] and column-major indices [N * MN]
  virtual const GPU_Vector<int>& get_NN_radial_ptr()
  {
    static GPU_Vector<int> dummy_NN;
    return dummy_NN;
  }
  virtual const GPU_Vector<int>& get_NL_radial_ptr()
  {
    static GPU_Vector<int> dummy_NL;
    return dummy_NL;
  }
  // type index of an atomic symbol in this potential's header (read_xyz.cu:349-361); -1 if absent
  virtual int type_of(const std::string& symbol) const = 0;
  // latched device-side errors (neighbour capacity); exits like the reference on failure
  virtual void check() = 0;
};
#endif

inline void b2h_pbc(const Box& box, int out[3])
{
  out[0] = box.pbc_x;
  out[1] = box.pbc_y;
  out[2] = box.pbc_z;
}

// replaces class NEP : Potential, src/force/nep.cuh:27-184
class NEP_B200 : public Potential
{
public:
  NEP_B200(const char* file_potential, const int num_atoms);
  ~NEP_B200() override;
  using Potential::compute; // the reference's other overloads keep their defaults
  void compute(
    Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
    GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial) override;
  B2H_EXTRA_VIRTUAL int type_of(const std::string& symbol) const B2H_EXTRA_OVERRIDE;
  B2H_EXTRA_VIRTUAL void check() B2H_EXTRA_OVERRIDE;

  // NN_radial [N] and NL_radial [N * MN_radial, entry k of atom i at k*N + i] in the caller's atom
  // order, as the reference's NEP exposes them (nep.cuh:100-101, potential.cuh:66-77)
  const GPU_Vector<int>& get_NN_radial_ptr() override;
  const GPU_Vector<int>& get_NL_radial_ptr() override;
  // false: outputs are stored, not added (a driver with this single potential skips its zeroing pass)
  void set_accumulate(bool accumulate);

private:
  void export_radial();
  b200md_nep* handle_ = nullptr;
  int num_calls_ = 0;
  int num_atoms_ = 0;
  bool lists_current_ = false;
  GPU_Vector<int> NN_radial_, NL_radial_, row_major_;
};

// replaces class LJ : Potential, src/force/lj.cuh:31-49
class LJ_B200 : public Potential
{
public:
  LJ_B200(const char* file_potential, const int num_atoms);
  ~LJ_B200() override;
  using Potential::compute; // the reference's other overloads keep their defaults
  void compute(
    Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
    GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial) override;
  B2H_EXTRA_VIRTUAL int type_of(const std::string& symbol) const B2H_EXTRA_OVERRIDE;
  B2H_EXTRA_VIRTUAL void check() B2H_EXTRA_OVERRIDE;

private:
  b200md_lj* handle_ = nullptr;
  int num_calls_ = 0;
};

// replaces class Tersoff1989 : Potential, src/force/tersoff1989.cuh:22-60 (FP64 like the reference)
class Tersoff1989_B200 : public Potential
{
public:
  Tersoff1989_B200(const char* file_potential, const int num_atoms);
  ~Tersoff1989_B200() override;
  using Potential::compute; // the reference's other overloads keep their defaults
  void compute(
    Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
    GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial) override;
  B2H_EXTRA_VIRTUAL int type_of(const std::string& symbol) const B2H_EXTRA_OVERRIDE;
  B2H_EXTRA_VIRTUAL void check() B2H_EXTRA_OVERRIDE;

private:
  b200md_tersoff* handle_ = nullptr;
  int num_calls_ = 0;
};

// replaces class EAM : Potential, src/force/eam.cuh:44-75 (eam_zhou_2004, eam_dai_2006)
class EAM_B200 : public Potential
{
public:
  EAM_B200(const char* file_potential, const int num_atoms);
  ~EAM_B200() override;
  using Potential::compute; // the reference's other overloads keep their defaults
  void compute(
    Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
    GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial) override;
  B2H_EXTRA_VIRTUAL int type_of(const std::string& symbol) const B2H_EXTRA_OVERRIDE;
  B2H_EXTRA_VIRTUAL void check() B2H_EXTRA_OVERRIDE;

private:
  b200md_eam* handle_ = nullptr;
  int num_calls_ = 0;
};

[[noreturn]] void b2h_fail(const char* where);
