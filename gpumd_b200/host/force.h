// force.h -- the Force driver (src/force/force.cuh:27-85): parse_potential + compute.
#pragma once
#include "potential.h"
#include <memory>
#include <vector>

class Force
{
public:
  // factory keyed on the first token of the potential file (force.cu:75-218); only the potentials
  // libb200md implements are accepted
  void parse_potential(const char* file_potential, const int num_atoms);
  // Force::compute, force.cu:771-985: wrap positions, zero the outputs, run the potential
  void compute(
    Box& box, GPU_Vector<double>& position_per_atom, GPU_Vector<int>& type,
    GPU_Vector<double>& potential_per_atom, GPU_Vector<double>& force_per_atom,
    GPU_Vector<double>& virial_per_atom);
  std::vector<std::unique_ptr<Potential>> potentials;

private:
  bool zero_first_ = true; // false when the single potential overwrites its outputs (NEP_B200)
};
