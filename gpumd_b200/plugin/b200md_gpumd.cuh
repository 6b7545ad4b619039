// b200md_gpumd.cuh -- the two hooks a GPUMD source tree calls to route its hot path through
// libb200md (applied to a COPY of the reference by oracle/patch_reference.py
// and built by oracle/Makefile.gpumd_b200):
//
//   Force::parse_potential   (src/force/force.cu:75-218)      -> b200md_make_potential
//   Integrate::initialize    (src/integrate/integrate.cu:76-280) -> b200md_make_ensemble
//
// Both return false -- and GPUMD proceeds with its own classes -- unless the environment variable
// GPUMD_B200 is set and libb200md implements the requested potential / ensemble.  The objects they
// create derive from the reference's OWN Potential / Ensemble (force/potential.cuh:20-113,
// integrate/ensemble.cuh:26-157): gpumd_b200/host/potential.cpp and ensemble.cpp compiled with
// -DB200MD_IN_GPUMD against the reference headers.
#pragma once
#include "force/potential.cuh"
#include "integrate/ensemble.cuh"
#include <memory>

bool b200md_make_potential(
  const char* potential_name, const char* potential_file, const int number_of_atoms,
  std::unique_ptr<Potential>& potential);

// target_pressure / pressure_coupling / deform*: the npt_ber parameters as Integrate holds them at
// initialize() (natural units); ignored for the other types
bool b200md_make_ensemble(
  const int type, const int move_group, const double* move_velocity, const int number_of_atoms,
  const double temperature, const double temperature_coupling, const double time_step,
  const double* target_pressure, const int num_target_pressure_components,
  const double* pressure_coupling, const int deform_x, const int deform_y, const int deform_z,
  const double* deform_rate, std::unique_ptr<Ensemble>& ensemble);
