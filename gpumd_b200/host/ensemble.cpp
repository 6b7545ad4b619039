// ensemble.cpp -- see ensemble.h.  Every kernel runs on the legacy default stream, like the
// reference's integrators.
#include "ensemble.h"
#include "potential.h"
#include <cstdio>
#include <cstdlib>

static void set_move(Ensemble& e, int mg, const double mv[3])
{
  e.move_group = mg;
  for (int d = 0; d < 3; ++d)
    e.move_velocity[d] = mv ? mv[d] : 0.0;
}

void B200_Integrator::b2_velocity_verlet(
  const Ensemble& e, const bool is_step1, const double time_step, const std::vector<Group>& group,
  Atom& atom)
{
  const int n = atom.number_of_atoms;
  const int* label = nullptr;
  if (e.fixed_group >= 0 || e.move_group >= 0) {
    if ((int)group.size() <= e.fixed_grouping_method) {
      fprintf(stderr, "Input Error:\n    fix / move need a grouping method in model.xyz.\n");
      exit(1);
    }
    label = group[e.fixed_grouping_method].label.data();
  }
  if (b200md_velocity_verlet_groups(
        is_step1 ? 1 : 0, n, n, time_step, atom.mass.data(), atom.position_per_atom.data(),
        atom.velocity_per_atom.data(), atom.force_per_atom.data(), label, e.fixed_group,
        e.move_group, e.move_velocity, nullptr) != B200MD_OK)
    b2h_fail("Ensemble::velocity_verlet");
}

void B200_Integrator::b2_find_thermo(
  const Ensemble& e, const double volume, const std::vector<Group>& group, Atom& atom,
  GPU_Vector<double>& thermo)
{
  const int n = atom.number_of_atoms;
  int n_temperature = n; // ensemble.cu:645-651
  if (e.fixed_group >= 0)
    n_temperature -= group[e.fixed_grouping_method].cpu_size[e.fixed_group];
  if (e.move_group >= 0)
    n_temperature -= group[e.move_grouping_method].cpu_size[e.move_group];
  const size_t need = (size_t)b200md_thermo_scratch_bytes(n);
  if (scratch_.size() != need)
    scratch_.resize(need, 0);
  if (b200md_find_thermo(
        n, n_temperature, volume, atom.mass.data(), atom.potential_per_atom.data(),
        atom.velocity_per_atom.data(), atom.virial_per_atom.data(), thermo.data(),
        scratch_.data(), nullptr) != B200MD_OK)
    b2h_fail("Ensemble::find_thermo");
}

void Ensemble_NVE_B200::compute1(
  const double time_step, const std::vector<Group>& group, Box&, Atom& atom, GPU_Vector<double>&)
{
  b2_velocity_verlet(*this, true, time_step, group, atom);
}

void Ensemble_NVE_B200::compute2(
  const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
  GPU_Vector<double>& thermo)
{
  b2_velocity_verlet(*this, false, time_step, group, atom);
  b2_find_thermo(*this, box.get_volume(), group, atom, thermo);
}

Ensemble_BER_B200::Ensemble_BER_B200(int t, int mg, const double mv[3], double T, double Tc)
{
  type = t;
  temperature = T;
  temperature_coupling = Tc;
  set_move(*this, mg, mv);
}

void Ensemble_BER_B200::compute1(
  const double time_step, const std::vector<Group>& group, Box&, Atom& atom, GPU_Vector<double>&)
{
  b2_velocity_verlet(*this, true, time_step, group, atom);
}

void Ensemble_BER_B200::compute2(
  const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
  GPU_Vector<double>& thermo)
{
  b2_velocity_verlet(*this, false, time_step, group, atom);
  b2_find_thermo(*this, box.get_volume(), group, atom, thermo);
  const int n = atom.number_of_atoms;
  if (b200md_berendsen_temperature(
        n, n, temperature, temperature_coupling, thermo.data(), atom.velocity_per_atom.data(),
        nullptr) != B200MD_OK)
    b2h_fail("Ensemble_BER_B200::compute2");
}

Ensemble_BDP_B200::Ensemble_BDP_B200(
  int t, int mg, const double mv[3], int N, double T, double Tc, unsigned seed)
{
  type = t;
  temperature = T;
  temperature_coupling = Tc;
  set_move(*this, mg, mv);
  if (b200md_bdp_create(N, T, Tc, seed, &bdp_) != B200MD_OK)
    b2h_fail("Ensemble_BDP_B200");
}

Ensemble_BDP_B200::~Ensemble_BDP_B200() { b200md_bdp_destroy(bdp_); }

void Ensemble_BDP_B200::compute1(
  const double time_step, const std::vector<Group>& group, Box&, Atom& atom, GPU_Vector<double>&)
{
  b2_velocity_verlet(*this, true, time_step, group, atom);
}

void Ensemble_BDP_B200::compute2(
  const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
  GPU_Vector<double>& thermo)
{
  b2_velocity_verlet(*this, false, time_step, group, atom);
  b2_find_thermo(*this, box.get_volume(), group, atom, thermo);
  const int n = atom.number_of_atoms;
  if (b200md_bdp_step(bdp_, n, n, thermo.data(), atom.velocity_per_atom.data(), nullptr) != B200MD_OK)
    b2h_fail("Ensemble_BDP_B200::compute2");
}

Ensemble_NHC_B200::Ensemble_NHC_B200(
  int t, int mg, const double mv[3], int N, double T, double Tc, double time_step_in)
{
  type = t;
  temperature = T;
  temperature_coupling = Tc;
  set_move(*this, mg, mv);
  if (b200md_nhc_create(N, T, Tc, time_step_in, &nhc_) != B200MD_OK)
    b2h_fail("Ensemble_NHC_B200");
}

Ensemble_NHC_B200::~Ensemble_NHC_B200() { b200md_nhc_destroy(nhc_); }

void Ensemble_NHC_B200::thermostat(
  const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
  GPU_Vector<double>& thermo)
{
  b2_find_thermo(*this, box.get_volume(), group, atom, thermo);
  const int n = atom.number_of_atoms;
  if (b200md_nhc_half_step(
        nhc_, n, n, time_step, thermo.data(), atom.velocity_per_atom.data(), nullptr) != B200MD_OK)
    b2h_fail("Ensemble_NHC_B200::thermostat");
}

void Ensemble_NHC_B200::compute1(
  const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
  GPU_Vector<double>& thermo)
{
  thermostat(time_step, group, box, atom, thermo);
  b2_velocity_verlet(*this, true, time_step, group, atom);
}

void Ensemble_NHC_B200::compute2(
  const double time_step, const std::vector<Group>& group, Box& box, Atom& atom,
  GPU_Vector<double>& thermo)
{
  b2_velocity_verlet(*this, false, time_step, group, atom);
  thermostat(time_step, group, box, atom, thermo);
}
