// potential.h -- GPUMD's Potential plugin interface (src/force/potential.cuh:20-113, the members and
// the pure virtual the run loop uses) and the adapters that forward it to libb200md's C-ABI.
// Inside the reference tree these adapters derive from the reference's own class Potential
// (INTEGRATION.md); here a same-shaped base keeps the standalone executable self-contained.
#pragma once
#include "../../include/b200md.h"
#include "model.h"

class Potential
{
public:
  int N1 = 0;
  int N2 = 0;
  double rc = 0.0;
  int nep_model_type = -1;
  virtual ~Potential() = default;
  virtual void compute(
    Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
    GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial) = 0;
  // type index of an atomic symbol in this potential's header (read_xyz.cu:349-361); -1 if absent
  virtual int type_of(const std::string& symbol) const = 0;
  // latched device-side errors (neighbour capacity); exits like the reference on failure
  virtual void check() = 0;
};

// replaces class NEP : Potential, src/force/nep.cuh:27-184
class NEP_B200 : public Potential
{
public:
  NEP_B200(const char* file_potential, const int num_atoms);
  ~NEP_B200() override;
  void compute(
    Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
    GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial) override;
  int type_of(const std::string& symbol) const override;
  void check() override;

private:
  b200md_nep* handle_ = nullptr;
  int num_calls_ = 0;
};

// replaces class LJ : Potential, src/force/lj.cuh:31-49
class LJ_B200 : public Potential
{
public:
  LJ_B200(const char* file_potential, const int num_atoms);
  ~LJ_B200() override;
  void compute(
    Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
    GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial) override;
  int type_of(const std::string& symbol) const override;
  void check() override;

private:
  b200md_lj* handle_ = nullptr;
};

// replaces class Tersoff1989 : Potential, src/force/tersoff1989.cuh:22-60 (FP64 like the reference)
class Tersoff1989_B200 : public Potential
{
public:
  Tersoff1989_B200(const char* file_potential, const int num_atoms);
  ~Tersoff1989_B200() override;
  void compute(
    Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
    GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial) override;
  int type_of(const std::string& symbol) const override;
  void check() override;

private:
  b200md_tersoff* handle_ = nullptr;
};

// replaces class EAM : Potential, src/force/eam.cuh:44-75 (eam_zhou_2004, eam_dai_2006)
class EAM_B200 : public Potential
{
public:
  EAM_B200(const char* file_potential, const int num_atoms);
  ~EAM_B200() override;
  void compute(
    Box& box, const GPU_Vector<int>& type, const GPU_Vector<double>& position,
    GPU_Vector<double>& potential, GPU_Vector<double>& force, GPU_Vector<double>& virial) override;
  int type_of(const std::string& symbol) const override;
  void check() override;

private:
  b200md_eam* handle_ = nullptr;
};

[[noreturn]] void b2h_fail(const char* where);
