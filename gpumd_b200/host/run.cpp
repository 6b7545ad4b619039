// run.cpp -- standalone driver `b200md`: reads GPUMD's own input files from the current directory
// (run.in keyword script, model.xyz extended XYZ, potential file) and runs the hot path through the
// adapter classes, writing thermo.out in the reference's format.
//
// Mirrors, for the keywords of the hot path only:
//   Run::Run / execute_run_in / parse_one_keyword   src/main_gpumd/run.cu:144-209, 343-575
//   Run::perform_a_run (the MD loop)                src/main_gpumd/run.cu:211-341
//   initialize_position (model.xyz)                 src/model/read_xyz.cu:145-425, 482-557
//   Velocity::initialize                            src/main_gpumd/velocity.cu:55-75, 312-347
//   Dump_Thermo                                     src/measure/dump_thermo.cu:57-129
// Keywords: potential, velocity <T> [seed <s>], ensemble nve | nvt_ber|nvt_nhc|nvt_bdp T T tau,
// time_step <fs>, dump_thermo <n>, run <n>.  Anything else is an input error (exit 1), as in the reference.
#include "ensemble.h"
#include "force.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>

static const double K_B = 8.617343e-5;                     // common.cuh:21
static const double TIME_UNIT_CONVERSION = 1.018051e+1;    // common.cuh:26
static const double PRESSURE_UNIT_CONVERSION = 1.602177e+2; // common.cuh:25

static const std::map<std::string, double> MASS_TABLE = {
  // subset of read_xyz.cu:36-142 covering the shipped potentials on the target configs
  {"H", 1.008},     {"C", 12.011},      {"N", 14.007},     {"O", 15.999},    {"Mg", 24.305},
  {"Al", 26.9815385}, {"Si", 28.085},   {"Ar", 39.948},    {"Ti", 47.867},   {"V", 50.9415},
  {"Cr", 51.9961},  {"Ni", 58.6934},    {"Cu", 63.546},    {"Zr", 91.224},   {"Mo", 95.95},
  {"Pd", 106.42},   {"Ag", 107.8682},   {"Te", 127.6},     {"Ba", 137.327},  {"Ta", 180.94788},
  {"W", 183.84},    {"Pt", 195.084},    {"Au", 196.966569}, {"Pb", 207.2}};

[[noreturn]] static void input_error(const std::string& msg)
{
  fprintf(stderr, "Input Error:\n    %s\n", msg.c_str());
  exit(1);
}

static std::string lower(std::string s)
{
  std::transform(s.begin(), s.end(), s.begin(), ::tolower);
  return s;
}

// value of key="..." or key=token on the comment line (case-insensitive key), "" if absent
static std::string header_value(const std::string& line, const std::string& key)
{
  const std::string low = lower(line);
  size_t p = low.find(key + "=");
  if (p == std::string::npos)
    return "";
  p += key.size() + 1;
  if (p < line.size() && line[p] == '"') {
    const size_t q = line.find('"', p + 1);
    return line.substr(p + 1, q - p - 1);
  }
  const size_t q = line.find_first_of(" \t", p);
  return line.substr(p, q == std::string::npos ? std::string::npos : q - p);
}

struct Model {
  Box box;
  Atom atom;
  bool has_velocity = false;
};

static void read_model(const char* path, Model& m)
{
  std::ifstream in(path);
  if (!in.is_open())
    input_error("Failed to open model.xyz.");
  std::string line;
  std::getline(in, line);
  const int N = std::atoi(line.c_str());
  if (N < 1)
    input_error("Number of atoms should be positive.");
  std::getline(in, line);
  std::string lat = header_value(line, "lattice");
  if (lat.empty())
    input_error("'lattice' is missing in the second line of model.xyz.");
  {
    std::istringstream ss(lat);
    double a[9];
    for (double& v : a)
      ss >> v;
    // stored transposed: lattice vectors become the columns of cpu_h (read_xyz.cu:203-219)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        m.box.cpu_h[r * 3 + c] = a[c * 3 + r];
  }
  const std::string pbc = header_value(line, "pbc");
  if (!pbc.empty()) {
    std::istringstream ss(pbc);
    std::string t[3];
    ss >> t[0] >> t[1] >> t[2];
    m.box.pbc_x = (t[0] == "T" || t[0] == "t");
    m.box.pbc_y = (t[1] == "T" || t[1] == "t");
    m.box.pbc_z = (t[2] == "T" || t[2] == "t");
  }
  // properties=species:S:1:pos:R:3[:mass:R:1][:vel:R:3] -> column offsets
  std::string props = lower(header_value(line, "properties"));
  if (props.empty())
    props = "species:s:1:pos:r:3";
  int off_species = -1, off_pos = -1, off_mass = -1, off_vel = -1, col = 0;
  {
    std::istringstream ss(props);
    std::string name, kind, width;
    while (std::getline(ss, name, ':') && std::getline(ss, kind, ':') && std::getline(ss, width, ':')) {
      const int w = std::atoi(width.c_str());
      if (name == "species")
        off_species = col;
      else if (name == "pos")
        off_pos = col;
      else if (name == "mass")
        off_mass = col;
      else if (name == "vel")
        off_vel = col;
      col += w;
    }
  }
  if (off_species < 0 || off_pos < 0)
    input_error("'species' or 'pos' is missing in properties.");
  Atom& a = m.atom;
  a.number_of_atoms = N;
  a.cpu_atom_symbol.resize(N);
  a.cpu_mass.resize(N);
  a.cpu_position_per_atom.resize((size_t)3 * N);
  a.cpu_velocity_per_atom.assign((size_t)3 * N, 0.0);
  m.has_velocity = off_vel >= 0;
  for (int n = 0; n < N; ++n) {
    if (!std::getline(in, line))
      input_error("model.xyz ended early.");
    std::istringstream ss(line);
    std::vector<std::string> tok;
    std::string w;
    while (ss >> w)
      tok.push_back(w);
    if ((int)tok.size() < col)
      input_error("Number of items in an atom line is too small.");
    a.cpu_atom_symbol[n] = tok[off_species];
    for (int d = 0; d < 3; ++d)
      a.cpu_position_per_atom[n + (size_t)N * d] = std::atof(tok[off_pos + d].c_str());
    if (off_mass >= 0) {
      a.cpu_mass[n] = std::atof(tok[off_mass].c_str());
    } else {
      auto it = MASS_TABLE.find(tok[off_species]);
      if (it == MASS_TABLE.end())
        input_error("Atom symbol " + tok[off_species] + " is not in the mass table.");
      a.cpu_mass[n] = it->second;
    }
    if (off_vel >= 0)
      for (int d = 0; d < 3; ++d) // A/fs -> natural units (read_xyz.cu:380-387)
        a.cpu_velocity_per_atom[n + (size_t)N * d] =
          std::atof(tok[off_vel + d].c_str()) * TIME_UNIT_CONVERSION;
  }
}

// Velocity::initialize: uniform random velocities, zero linear momentum, scale to T
// (velocity.cu:55-75,312-347; angular-momentum removal is skipped: periodic bulk systems only)
static void initialize_velocity(Atom& a, double temperature, bool use_seed, int seed)
{
  const int N = a.number_of_atoms;
  double* v = a.cpu_velocity_per_atom.data();
  if (use_seed) {
    for (int n = 0; n < N; ++n)
      for (int d = 0; d < 3; ++d) {
        srand((unsigned)seed + n * 3 + d);
        v[n + (size_t)N * d] = -1.0 + (rand() * 2.0) / RAND_MAX;
      }
  } else {
    for (int n = 0; n < N; ++n)
      for (int d = 0; d < 3; ++d)
        v[n + (size_t)N * d] = -1.0 + (rand() * 2.0) / RAND_MAX;
  }
  double p[3] = {0, 0, 0}, mt = 0;
  for (int n = 0; n < N; ++n) {
    mt += a.cpu_mass[n];
    for (int d = 0; d < 3; ++d)
      p[d] += a.cpu_mass[n] * v[n + (size_t)N * d];
  }
  double ke2 = 0;
  for (int n = 0; n < N; ++n)
    for (int d = 0; d < 3; ++d) {
      v[n + (size_t)N * d] -= p[d] / mt;
      ke2 += a.cpu_mass[n] * v[n + (size_t)N * d] * v[n + (size_t)N * d];
    }
  const double factor = std::sqrt(temperature / (ke2 / (3.0 * K_B * N)));
  for (size_t k = 0; k < (size_t)3 * N; ++k)
    v[k] *= factor;
}

class Run
{
public:
  Run()
  {
    read_model("model.xyz", model_);
    printf("Number of atoms is %d.\n", model_.atom.number_of_atoms);
    execute_run_in();
  }

private:
  Model model_;
  Force force_;
  std::unique_ptr<Ensemble> ensemble_;
  GPU_Vector<double> thermo_;
  std::vector<Group> group_;
  double time_step_ = 1.0 / TIME_UNIT_CONVERSION;
  int dump_thermo_ = 0;
  bool state_on_gpu_ = false;
  bool has_potential_ = false;
  long global_step_ = 0;

  void upload_state()
  {
    Atom& a = model_.atom;
    const int N = a.number_of_atoms;
    a.type.resize(N);
    a.type.copy_from_host(a.cpu_type.data());
    a.mass.resize(N);
    a.mass.copy_from_host(a.cpu_mass.data());
    a.position_per_atom.resize((size_t)3 * N);
    a.position_per_atom.copy_from_host(a.cpu_position_per_atom.data());
    a.velocity_per_atom.resize((size_t)3 * N);
    a.velocity_per_atom.copy_from_host(a.cpu_velocity_per_atom.data());
    a.force_per_atom.resize((size_t)3 * N, 0.0);
    a.virial_per_atom.resize((size_t)9 * N, 0.0);
    a.potential_per_atom.resize(N, 0.0);
    thermo_.resize(12, 0.0);
    state_on_gpu_ = true;
  }

  void execute_run_in()
  {
    std::ifstream in("run.in");
    if (!in.is_open())
      input_error("Failed to open run.in.");
    std::string line;
    while (std::getline(in, line)) {
      const size_t hash = line.find('#');
      if (hash != std::string::npos)
        line = line.substr(0, hash);
      std::istringstream ss(line);
      std::vector<std::string> tok;
      std::string w;
      while (ss >> w)
        tok.push_back(w);
      if (!tok.empty())
        parse_one_keyword(tok);
    }
  }

  void parse_one_keyword(const std::vector<std::string>& t)
  {
    Atom& a = model_.atom;
    if (t[0] == "potential") {
      if (t.size() != 2)
        input_error("potential should have 1 parameter.");
      force_.parse_potential(t[1].c_str(), a.number_of_atoms);
      a.cpu_type.resize(a.number_of_atoms);
      for (int n = 0; n < a.number_of_atoms; ++n) {
        const int ty = force_.potentials[0]->type_of(a.cpu_atom_symbol[n]);
        if (ty < 0)
          input_error("There is atom in model.xyz that is not allowed in the used potential.");
        a.cpu_type[n] = ty;
      }
      has_potential_ = true;
    } else if (t[0] == "velocity") {
      if (t.size() != 2 && t.size() != 4)
        input_error("velocity should have 1 or 3 parameters.");
      if (!model_.has_velocity) {
        const bool use_seed = t.size() == 4 && t[2] == "seed";
        initialize_velocity(a, std::atof(t[1].c_str()), use_seed, use_seed ? std::atoi(t[3].c_str()) : 0);
      }
    } else if (t[0] == "ensemble") {
      // Integrate::parse_ensemble, integrate.cu:406-432,569-600: nvt_* take T1 T2 tau_T/dt
      if (t.size() == 2 && t[1] == "nve") {
        ensemble_.reset(new Ensemble_NVE_B200(0));
      } else if (t.size() == 5 && (t[1] == "nvt_ber" || t[1] == "nvt_nhc" || t[1] == "nvt_bdp")) {
        const double T1 = std::atof(t[2].c_str()), T2 = std::atof(t[3].c_str());
        const double Tc = std::atof(t[4].c_str());
        if (T1 <= 0.0 || T2 <= 0.0)
          input_error("Temperatures should > 0.");
        if (T1 != T2)
          input_error("temperature ramps are not supported by the b200md backend.");
        if (Tc < 1.0)
          input_error("Temperature coupling should >= 1.");
        if (t[1] == "nvt_ber")
          ensemble_.reset(new Ensemble_BER_B200(1, T1, Tc));
        else if (t[1] == "nvt_bdp")
          ensemble_.reset(new Ensemble_BDP_B200(4, a.number_of_atoms, T1, Tc));
        else
          ensemble_.reset(new Ensemble_NHC_B200(2, a.number_of_atoms, T1, Tc, time_step_));
      } else {
        input_error("only 'ensemble nve', 'nvt_ber|nvt_nhc|nvt_bdp T T tau' are supported "
                    "by the b200md backend.");
      }
    } else if (t[0] == "time_step") {
      if (t.size() < 2)
        input_error("time_step should have at least 1 parameter.");
      time_step_ = std::atof(t[1].c_str()) / TIME_UNIT_CONVERSION; // run.cu:657
    } else if (t[0] == "dump_thermo") {
      dump_thermo_ = std::atoi(t[1].c_str());
    } else if (t[0] == "run") {
      perform_a_run(std::atoi(t[1].c_str()));
      dump_thermo_ = 0; // non-propagating keywords are reset after each run (run.cu:329-340)
    } else {
      input_error("'" + t[0] + "' is not a keyword supported by the b200md backend.");
    }
  }

  void write_thermo(FILE* fid)
  {
    // dump_thermo.cu:73-129: T, K, U, then sxx syy szz syz sxz sxy in GPa, then the box
    double th[8];
    thermo_.copy_to_host(th, 8);
    const int N = model_.atom.number_of_atoms;
    const double* h = model_.box.cpu_h;
    fprintf(fid, "%20.10e%20.10e%20.10e%20.10e%20.10e%20.10e%20.10e%20.10e%20.10e", th[0],
            1.5 * N * K_B * th[0], th[1], th[2] * PRESSURE_UNIT_CONVERSION,
            th[3] * PRESSURE_UNIT_CONVERSION, th[4] * PRESSURE_UNIT_CONVERSION,
            th[7] * PRESSURE_UNIT_CONVERSION, th[6] * PRESSURE_UNIT_CONVERSION,
            th[5] * PRESSURE_UNIT_CONVERSION);
    fprintf(fid, "%20.10e%20.10e%20.10e%20.10e%20.10e%20.10e%20.10e%20.10e%20.10e\n", h[0], h[3], h[6],
            h[1], h[4], h[7], h[2], h[5], h[8]);
  }

  void perform_a_run(int number_of_steps)
  {
    if (!has_potential_ || !ensemble_)
      input_error("'potential' and 'ensemble' must precede 'run'.");
    Atom& a = model_.atom;
    if (!state_on_gpu_)
      upload_state();
    FILE* fid = nullptr;
    if (dump_thermo_ > 0) {
      fid = fopen("thermo.out", "a");
      fprintf(fid, "# dump_thermo %d\n# format_version 1\n# num_atoms %d\n# dt_output %.10e fs\n",
              dump_thermo_, a.number_of_atoms, time_step_ * dump_thermo_ * TIME_UNIT_CONVERSION);
      fprintf(fid, "# columns T KE PE sxx syy szz syz sxz sxy ax ay az bx by bz cx cy cz\n");
    }
    // the first force evaluation is outside the timed loop, as in run.cu:217-248
    force_.compute(model_.box, a.position_per_atom, a.type, a.potential_per_atom, a.force_per_atom,
                   a.virial_per_atom);
    force_.potentials[0]->check();
    printf("Run %d steps.\n", number_of_steps);
    const auto t0 = std::chrono::high_resolution_clock::now();
    for (int step = 0; step < number_of_steps; ++step) {
      ensemble_->compute1(time_step_, group_, model_.box, a, thermo_);
      force_.compute(model_.box, a.position_per_atom, a.type, a.potential_per_atom, a.force_per_atom,
                     a.virial_per_atom);
      ensemble_->compute2(time_step_, group_, model_.box, a, thermo_);
      ++global_step_;
      if (fid && (step + 1) % dump_thermo_ == 0)
        write_thermo(fid);
    }
    B2H_CHECK(cudaDeviceSynchronize());
    force_.potentials[0]->check();
    const auto t1 = std::chrono::high_resolution_clock::now();
    const double sec = std::chrono::duration<double>(t1 - t0).count();
    printf("Time used for this run = %g second.\n", sec);
    printf("Speed of this run = %g atom*step/second.\n", a.number_of_atoms * (double)number_of_steps / sec);
    if (fid)
      fclose(fid);
  }
};

int main()
{
  printf("b200md: Blackwell-native MD hot path behind GPUMD's Potential/Ensemble surface.\n");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "Error:\n    no CUDA device found; b200md has no CPU fallback.\n");
    return 1;
  }
  Run run;
  printf("Finished running b200md.\n");
  return 0;
}
