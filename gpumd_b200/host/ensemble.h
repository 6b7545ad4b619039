// ensemble.h -- the Ensemble plugin interface (src/integrate/ensemble.cuh:26-157) and the NVE
// integrator (src/integrate/ensemble_nve.cu:31-95) on top of libb200md's kernels.
#pragma once
#include "../../include/b200md.h"
#include "model.h"
#include <vector>

class Ensemble
{
public:
  virtual ~Ensemble() = default;
  virtual void compute1(
    const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
    GPU_Vector<double>& thermo) = 0;
  virtual void compute2(
    const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
    GPU_Vector<double>& thermo) = 0;
  int type = 0;
  int fixed_group = -1;
  int move_group = -1;
  double temperature = 0.0;

protected:
  // Ensemble::velocity_verlet, ensemble.cu:348-397 (plain variant) and find_thermo, :636-673
  void velocity_verlet(const bool is_step1, const double time_step, Atom& atom);
  void find_thermo(const double volume, Atom& atom, GPU_Vector<double>& thermo);
  GPU_Vector<char> scratch_;
};

// replaces Ensemble_BER for type 1 (nvt_ber), src/integrate/ensemble_ber.cu:178-233
class Ensemble_BER_B200 : public Ensemble
{
public:
  Ensemble_BER_B200(int t, double T, double Tc)
  {
    type = t;
    temperature = T;
    temperature_coupling = Tc;
  }
  void compute1(
    const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
    GPU_Vector<double>& thermo) override;
  void compute2(
    const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
    GPU_Vector<double>& thermo) override;
  double temperature_coupling = 100.0;
};

// replaces Ensemble_BDP for type 4 (nvt_bdp), src/integrate/ensemble_bdp.cu:69-101,151-196; the
// generator lives on the device and follows the reference's -DDEBUG stream (seed 12345678)
class Ensemble_BDP_B200 : public Ensemble
{
public:
  Ensemble_BDP_B200(int t, int N, double T, double Tc, unsigned seed = 12345678u);
  ~Ensemble_BDP_B200() override;
  void compute1(
    const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
    GPU_Vector<double>& thermo) override;
  void compute2(
    const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
    GPU_Vector<double>& thermo) override;

private:
  b200md_bdp* bdp_ = nullptr;
};

// replaces Ensemble_NHC for type 2 (nvt_nhc), src/integrate/ensemble_nhc.cu:173-237
class Ensemble_NHC_B200 : public Ensemble
{
public:
  Ensemble_NHC_B200(int t, int N, double T, double Tc, double time_step);
  ~Ensemble_NHC_B200() override;
  void compute1(
    const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
    GPU_Vector<double>& thermo) override;
  void compute2(
    const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
    GPU_Vector<double>& thermo) override;

private:
  void thermostat(const double time_step, Box& box, Atom& atom, GPU_Vector<double>& thermo);
  b200md_nhc* nhc_ = nullptr;
};

class Ensemble_NVE_B200 : public Ensemble
{
public:
  explicit Ensemble_NVE_B200(int t = 0) { type = t; }
  void compute1(
    const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
    GPU_Vector<double>& thermo) override;
  void compute2(
    const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
    GPU_Vector<double>& thermo) override;
};
