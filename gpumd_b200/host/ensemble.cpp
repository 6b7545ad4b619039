#include "ensemble.h"
#include "potential.h"

void Ensemble::velocity_verlet(const bool is_step1, const double time_step, Atom& atom)
{
  if (fixed_group != -1 || move_group != -1) {
    fprintf(stderr, "Input Error:\n    fixed/move groups are not supported by the b200md backend.\n");
    exit(1);
  }
  if (b200md_velocity_verlet(
        is_step1 ? 1 : 0, atom.number_of_atoms, time_step, atom.mass.data(),
        atom.position_per_atom.data(), atom.velocity_per_atom.data(), atom.force_per_atom.data(),
        nullptr) != B200MD_OK)
    b2h_fail("Ensemble::velocity_verlet");
}

void Ensemble::find_thermo(const double volume, Atom& atom, GPU_Vector<double>& thermo)
{
  const int n = atom.number_of_atoms;
  const size_t need = (size_t)b200md_thermo_scratch_bytes(n);
  if (scratch_.size() != need)
    scratch_.resize(need, 0);
  if (b200md_find_thermo(
        n, n, volume, atom.mass.data(), atom.potential_per_atom.data(),
        atom.velocity_per_atom.data(), atom.virial_per_atom.data(), thermo.data(),
        scratch_.data(), nullptr) != B200MD_OK)
    b2h_fail("Ensemble::find_thermo");
}

void Ensemble_NVE_B200::compute1(
  const double time_step, const std::vector<Group>&, Box&, Atom& atom, GPU_Vector<double>&)
{
  velocity_verlet(true, time_step, atom);
}

void Ensemble_NVE_B200::compute2(
  const double time_step, const std::vector<Group>&, Box& box, Atom& atom,
  GPU_Vector<double>& thermo)
{
  velocity_verlet(false, time_step, atom);
  find_thermo(box.get_volume(), atom, thermo);
}

void Ensemble_BER_B200::compute1(
  const double time_step, const std::vector<Group>&, Box&, Atom& atom, GPU_Vector<double>&)
{
  velocity_verlet(true, time_step, atom);
}

void Ensemble_BER_B200::compute2(
  const double time_step, const std::vector<Group>&, Box& box, Atom& atom,
  GPU_Vector<double>& thermo)
{
  velocity_verlet(false, time_step, atom);
  find_thermo(box.get_volume(), atom, thermo);
  const int n = atom.number_of_atoms;
  if (b200md_berendsen_temperature(
        n, n, temperature, temperature_coupling, thermo.data(), atom.velocity_per_atom.data(),
        nullptr) != B200MD_OK)
    b2h_fail("Ensemble_BER_B200::compute2");
}

Ensemble_BDP_B200::Ensemble_BDP_B200(int t, int N, double T, double Tc, unsigned seed)
{
  type = t;
  temperature = T;
  if (b200md_bdp_create(N, T, Tc, seed, &bdp_) != B200MD_OK)
    b2h_fail("Ensemble_BDP_B200");
}

Ensemble_BDP_B200::~Ensemble_BDP_B200() { b200md_bdp_destroy(bdp_); }

void Ensemble_BDP_B200::compute1(
  const double time_step, const std::vector<Group>&, Box&, Atom& atom, GPU_Vector<double>&)
{
  velocity_verlet(true, time_step, atom);
}

void Ensemble_BDP_B200::compute2(
  const double time_step, const std::vector<Group>&, Box& box, Atom& atom,
  GPU_Vector<double>& thermo)
{
  velocity_verlet(false, time_step, atom);
  find_thermo(box.get_volume(), atom, thermo);
  const int n = atom.number_of_atoms;
  if (b200md_bdp_step(bdp_, n, n, thermo.data(), atom.velocity_per_atom.data(), nullptr) != B200MD_OK)
    b2h_fail("Ensemble_BDP_B200::compute2");
}

Ensemble_NHC_B200::Ensemble_NHC_B200(int t, int N, double T, double Tc, double time_step)
{
  type = t;
  temperature = T;
  if (b200md_nhc_create(N, T, Tc, time_step, &nhc_) != B200MD_OK)
    b2h_fail("Ensemble_NHC_B200");
}

Ensemble_NHC_B200::~Ensemble_NHC_B200() { b200md_nhc_destroy(nhc_); }

void Ensemble_NHC_B200::thermostat(
  const double time_step, Box& box, Atom& atom, GPU_Vector<double>& thermo)
{
  find_thermo(box.get_volume(), atom, thermo);
  const int n = atom.number_of_atoms;
  if (b200md_nhc_half_step(
        nhc_, n, n, time_step, thermo.data(), atom.velocity_per_atom.data(), nullptr) != B200MD_OK)
    b2h_fail("Ensemble_NHC_B200::thermostat");
}

void Ensemble_NHC_B200::compute1(
  const double time_step, const std::vector<Group>&, Box& box, Atom& atom,
  GPU_Vector<double>& thermo)
{
  thermostat(time_step, box, atom, thermo);
  velocity_verlet(true, time_step, atom);
}

void Ensemble_NHC_B200::compute2(
  const double time_step, const std::vector<Group>&, Box& box, Atom& atom,
  GPU_Vector<double>& thermo)
{
  velocity_verlet(false, time_step, atom);
  thermostat(time_step, box, atom, thermo);
}
