/*
 * b200md.h -- C-ABI of libb200md.so: the Blackwell-native (sm_100a) MD hot path that drops in
 * behind GPUMD's Potential / Ensemble plugin surface.
 *
 * Every entry point states the reference interface it replaces (paths relative to the GPUMD
 * tree, commit d98135d).  Conventions on this boundary are GPUMD's own:
 *   - all `d_*` pointers are DEVICE pointers owned by the caller (GPUMD's GPU_Vector<T>);
 *     the library owns only its scratch (cell lists, neighbour lists, descriptor buffers);
 *   - per-atom arrays are SoA: position[3N] = x[N],y[N],z[N] (src/model/atom.cuh:21-52),
 *     virial[9N] = xx,yy,zz,xy,xz,yz,yx,zx,zy blocks of N (src/force/force.cu:859-861);
 *   - h[9] is Box::cpu_h[0..8] (src/model/box.cuh:18-35): row-major, lattice vectors as columns;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is what
 *     every reference kernel uses);
 *   - functions return 0 on success, non-zero on error; b200md_last_error() describes it.
 *     (The reference prints and exit(1)s, src/utilities/error.cuh:22-62; the C++ adapter in
 *     gpumd_b200/host does exactly that with these codes.)
 *   - work is enqueued asynchronously; capacity overflows detected on the device are latched
 *     and reported by b200md_*_check(), which synchronises the stream.
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef B200MD_H
#define B200MD_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200MD_OK 0
#define B200MD_ERR_ARG 1       /* bad argument / unsupported model */
#define B200MD_ERR_IO 2        /* cannot read potential file */
#define B200MD_ERR_CUDA 3      /* CUDA runtime error */
#define B200MD_ERR_SMALL_BOX 4 /* periodic thickness <= 2.5*(rc+1): reference's small-box path */
#define B200MD_ERR_OVERFLOW 5  /* neighbour / cell capacity exceeded on the device */

const char* b200md_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long b200md_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * NEP potential.  Replaces class NEP : Potential (src/force/nep.cuh:27-184):
 *   b200md_nep_create  <- NEP::NEP(const char* file_potential, const int num_atoms), nep.cu:100-395
 *   b200md_nep_compute <- NEP::compute(Box&, type, position, potential, force, virial),
 *                         nep.cu:1356-1389 -> compute_large_box nep.cu:978-1138
 *                         (outputs are ACCUMULATED with += exactly like nep.cu:653,755-770)
 * ------------------------------------------------------------------------------------------ */
typedef struct b200md_nep b200md_nep;

int b200md_nep_create(const char* nep_txt_path, int num_atoms, b200md_nep** out);
void b200md_nep_destroy(b200md_nep* p);

/* what: 0 num_types, 1 descriptor dim, 2 neurons, 3 MN_radial (enlarged), 4 MN_angular
 * (enlarged), 5 zbl enabled, 6 number of neighbour-list rebuilds so far */
int b200md_nep_info(const b200md_nep* p, int what);
double b200md_nep_rc(const b200md_nep* p); /* Potential::rc, src/force/potential.cuh:32 */
/* atomic symbol of type t (species -> type index mapping, src/model/read_xyz.cu:349-361) */
const char* b200md_nep_symbol(const b200md_nep* p, int t);

int b200md_nep_compute(
  b200md_nep* p, int n, const double h[9], const int pbc[3], const int* d_type,
  const double* d_position, double* d_potential, double* d_force, double* d_virial, void* stream);

/* Same call with HOST buffers (pageable or pinned): copies type/position to the device, runs the
 * path, and OVERWRITES the host arrays potential[n], force[3n], virial[9n] with the result
 * (synchronous).  This is the end-to-end entry bench.py times. */
int b200md_nep_compute_host(
  b200md_nep* p, int n, const double h[9], const int pbc[3], const int* type,
  const double* position, double* potential, double* force, double* virial);

/* Parity hook: the radial / angular neighbour sets of the LAST compute call, mapped back to the
 * caller's atom indices, ascending, row-major NL[i*mn + k] -- the quantity the reference builds
 * in find_neighbor_list_large_box (nep.cu:436-486).  Device pointers. */
int b200md_nep_export_neighbors(
  b200md_nep* p, int mn_r, int* d_NN_r, int* d_NL_r, int mn_a, int* d_NN_a, int* d_NL_a,
  void* stream);
/* Row-major rows NL[i*mn + k] -> the reference's column-major layout NL[k*n + i]
 * (NEP::get_NL_radial_ptr, nep.cuh:100-101 / potential.cuh:66-77).  Device pointers. */
int b200md_transpose_int(int n, int mn, const int* d_row_major, int* d_column_major, void* stream);
/* Parity hook: scaled descriptors q[d*N + i] (caller order) of the last compute call. */
int b200md_nep_export_descriptors(b200md_nep* p, float* d_q, void* stream);

/* synchronise and report latched device-side errors (B200MD_ERR_OVERFLOW) */
int b200md_nep_check(b200md_nep* p, void* stream);

/* Measurement hooks (no reference counterpart; the reference has no tracing).
 * b200md_nep_profile(p,1) makes every following b200md_nep_compute record CUDA events around each
 * stage on the launching stream (up to 256 calls); b200md_nep_profile_read synchronises and returns
 * the number of stages, filling ms_sum[stage] (total milliseconds) and counts[stage] (samples).
 * b200md_nep_mean_neighbors: mean skin / radial / angular list lengths of the last call. */
int b200md_nep_profile(b200md_nep* p, int enable);
int b200md_nep_profile_read(b200md_nep* p, int max_stages, float* ms_sum, int* counts);
const char* b200md_nep_stage_name(int stage);
int b200md_nep_mean_neighbors(b200md_nep* p, double out3[3]);

/* ------------------------------------------------------------------------------------------
 * LJ potential.  Replaces class LJ : Potential (src/force/lj.cuh:31-49):
 *   b200md_lj_create  <- LJ::LJ(FILE*, int num_types, int num_atoms), lj.cu:28-59
 *                        (the file is the whole potential file, first line "lj Nt sym...")
 *   b200md_lj_compute <- LJ::compute, lj.cu:184-219
 * ------------------------------------------------------------------------------------------ */
typedef struct b200md_lj b200md_lj;
int b200md_lj_create(const char* lj_txt_path, int num_atoms, b200md_lj** out);
void b200md_lj_destroy(b200md_lj* p);
double b200md_lj_rc(const b200md_lj* p);
int b200md_lj_info(const b200md_lj* p, int what); /* 0 num_types, 6 rebuild count */
const char* b200md_lj_symbol(const b200md_lj* p, int t);
int b200md_lj_compute(
  b200md_lj* p, int n, const double h[9], const int pbc[3], const int* d_type,
  const double* d_position, double* d_potential, double* d_force, double* d_virial, void* stream);
int b200md_lj_check(b200md_lj* p, void* stream);
int b200md_lj_invalidate(b200md_lj* p, int n_new, void* stream); /* see b200md_nep_invalidate */

/* ------------------------------------------------------------------------------------------
 * Tersoff-1989 potential, FP64.  Replaces class Tersoff1989 : Potential
 * (src/force/tersoff1989.cuh, tersoff1989.cu:31-586):
 *   b200md_tersoff_create  <- Tersoff1989::Tersoff1989(FILE*, int num_types, int num_atoms), :31-150
 *                             (whole potential file: "tersoff_1989 Nt sym..." + 11 numbers per type
 *                             [+ chi]); 1 or 2 types
 *   b200md_tersoff_compute <- Tersoff1989::compute, :508-586 (local list + step1 + step2 +
 *                             find_properties_many_body, potential.cu:35-134)
 * b200md_compute_heat      <- compute_heat / gpu_compute_heat, src/measure/compute_heat.cu:32-90:
 *                             d_heat[5 * heat_stride] = jx_in, jx_out, jy_in, jy_out, jz per atom
 * ------------------------------------------------------------------------------------------ */
typedef struct b200md_tersoff b200md_tersoff;
int b200md_tersoff_create(const char* path, int num_atoms, b200md_tersoff** out);
void b200md_tersoff_destroy(b200md_tersoff* p);
double b200md_tersoff_rc(const b200md_tersoff* p);
int b200md_tersoff_info(const b200md_tersoff* p, int what); /* 0 num_types, 6 rebuild count */
const char* b200md_tersoff_symbol(const b200md_tersoff* p, int t);
int b200md_tersoff_compute(
  b200md_tersoff* p, int n, const double h[9], const int pbc[3], const int* d_type,
  const double* d_position, double* d_potential, double* d_force, double* d_virial, void* stream);
int b200md_tersoff_invalidate(b200md_tersoff* p, int n_new, void* stream);
int b200md_tersoff_check(b200md_tersoff* p, void* stream);
int b200md_compute_heat(
  int n, int stride, const double* d_virial, const double* d_velocity, double* d_heat,
  int heat_stride, void* stream);

/* ------------------------------------------------------------------------------------------
 * Analytic EAM potentials.  Replaces class EAM : Potential (src/force/eam.cuh:20-90, eam.cu:28-580)
 * for `eam_zhou_2004` (1..18 types) and `eam_dai_2006` (1 type):
 *   b200md_eam_create  <- EAM::EAM(FILE*, char* name, int num_types, int num_atoms), eam.cu:28-44
 *   b200md_eam_compute <- EAM::compute, eam.cu:478-580 (step1 density/embedding, step2 forces)
 * ------------------------------------------------------------------------------------------ */
typedef struct b200md_eam b200md_eam;
int b200md_eam_create(const char* path, int num_atoms, b200md_eam** out);
void b200md_eam_destroy(b200md_eam* p);
double b200md_eam_rc(const b200md_eam* p);
int b200md_eam_info(const b200md_eam* p, int what); /* 0 num_types, 6 rebuild count */
const char* b200md_eam_symbol(const b200md_eam* p, int t);
int b200md_eam_compute(
  b200md_eam* p, int n, const double h[9], const int pbc[3], const int* d_type,
  const double* d_position, double* d_potential, double* d_force, double* d_virial, void* stream);
int b200md_eam_invalidate(b200md_eam* p, int n_new, void* stream);
int b200md_eam_check(b200md_eam* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Force::compute pre-steps (src/force/force.cu:771-801):
 *   b200md_apply_pbc        <- gpu_apply_pbc, force.cu:424-459
 *   b200md_zero_properties  <- initialize_properties, force.cu:314-333
 * ------------------------------------------------------------------------------------------ */
int b200md_apply_pbc(int n, const double h[9], const int pbc[3], double* d_position, void* stream);
int b200md_zero_properties(
  int n, double* d_potential, double* d_force, double* d_virial, void* stream);

/* ------------------------------------------------------------------------------------------
 * Integrator kernels behind Ensemble::compute1 / compute2 (src/integrate/ensemble.cuh:26-157):
 *   b200md_velocity_verlet <- Ensemble::velocity_verlet / gpu_velocity_verlet,
 *                             ensemble.cu:176-214,348-397 (the variant without groups)
 *   b200md_velocity_verlet_groups <- the variant with `fix` / `move` groups, ensemble.cu:111-174,
 *                             369-391: d_group_label = Group::label of the fixed grouping method
 *                             (device, n ints), fixed_group / move_group = -1 when unused,
 *                             move_velocity in natural units.  Atoms of either group keep zero
 *                             velocity; pass n_temperature = n - |fixed| - |move| to
 *                             b200md_find_thermo (ensemble.cu:645-651).
 *   b200md_find_thermo     <- Ensemble::find_thermo / gpu_find_thermo_instant_temperature,
 *                             ensemble.cu:434-673: d_thermo[0..7] = T,U,sxx,syy,szz,sxy,sxz,syz
 *   b200md_scale_velocity  <- Ensemble::scale_velocity_global, ensemble.cu:676-698
 * time_step is in GPUMD natural units (fs / 10.18051, src/utilities/common.cuh:26).
 * d_scratch for find_thermo: at least b200md_thermo_scratch_bytes(n) bytes of device memory.
 * ------------------------------------------------------------------------------------------ */
int b200md_velocity_verlet(
  int is_step1, int n, double time_step, const double* d_mass, double* d_position,
  double* d_velocity, const double* d_force, void* stream);
int b200md_velocity_verlet_groups(
  int is_step1, int n, int stride, double time_step, const double* d_mass, double* d_position,
  double* d_velocity, const double* d_force, const int* d_group_label, int fixed_group,
  int move_group, const double move_velocity[3], void* stream);
long long b200md_thermo_scratch_bytes(int n);
int b200md_find_thermo(
  int n, int n_temperature, double volume, const double* d_mass, const double* d_potential,
  const double* d_velocity, const double* d_virial, double* d_thermo8, void* d_scratch,
  void* stream);
int b200md_scale_velocity(int n, double factor, double* d_velocity, void* stream);

/* NVT thermostats (global velocity scaling).  `stride` as in the *_strided calls below (= n for a
 * single domain); d_thermo[0] must hold the instantaneous GLOBAL temperature (b200md_find_thermo,
 * all-reduced over ranks when sharded).  temperature_coupling = tau_T / time_step as in run.in.
 *   b200md_berendsen_temperature <- gpu_berendsen_temperature, src/integrate/ensemble_ber.cu:70-86
 *                                   (Ensemble_BER::compute2, :195-233)
 *   b200md_nhc_*                 <- Ensemble_NHC (src/integrate/ensemble_nhc.cu:31-50, 101-237):
 *                                   one call = find chain factor from d_thermo[0] + scale velocities,
 *                                   i.e. the thermostat half of integrate_nvt_nhc_1 / _2.  The chain
 *                                   is integrated on the device: no D2H copy per half step.
 *   b200md_bdp_*                 <- Ensemble_BDP, NVT branch (src/integrate/ensemble_bdp.cu:69-101,
 *                                   svr_utilities.cuh:27-135): stochastic velocity rescaling after
 *                                   the second half step.  The generator (std::mt19937 semantics)
 *                                   lives on the device; seed 12345678 reproduces the reference's
 *                                   -DDEBUG stream (ensemble_bdp.cu:31-32).  3*n_global must fit an
 *                                   int, as in the reference. */
int b200md_berendsen_temperature(
  int n, int stride, double temperature, double temperature_coupling, const double* d_thermo,
  double* d_velocity, void* stream);
typedef struct b200md_nhc b200md_nhc;
int b200md_nhc_create(
  long long n_global, double temperature, double temperature_coupling, double time_step,
  b200md_nhc** out);
void b200md_nhc_destroy(b200md_nhc* p);
int b200md_nhc_half_step(
  b200md_nhc* p, int n, int stride, double time_step, const double* d_thermo, double* d_velocity,
  void* stream);
typedef struct b200md_bdp b200md_bdp;
int b200md_bdp_create(
  long long n_global, double temperature, double temperature_coupling, unsigned seed,
  b200md_bdp** out);
void b200md_bdp_destroy(b200md_bdp* p);
int b200md_bdp_step(
  b200md_bdp* p, int n, int stride, const double* d_thermo, double* d_velocity, void* stream);

/* Heat-current autocorrelation, SURVEY.md 8f rank 4.  Replaces HAC::preprocess / process /
 * postprocess (src/measure/hac.cu:32-280):
 *   b200md_hac_create  <- preprocess: Nd = number_of_steps / sample_interval records of 5 components
 *   b200md_hac_sample  <- process: call every step; on (step+1) % sample_interval == 0 it runs
 *                         compute_heat (compute_heat.cu:32-90) and sums jx_in jx_out jy_in jy_out jz
 *   b200md_hac_finish  <- postprocess: hac[nc + Nc*k] (gpu_find_hac, :111-170) and the running thermal
 *                         conductivity rtc[nc + Nc*k] in W/mK (find_rtc, :173-181; factor dt/2/(kB T^2 V)
 *                         * 1.573769e5); host output arrays of 5*Nc doubles; time_step in natural units
 *   b200md_hac_series  <- the recorded heat-current series [nd + Nd*k] (5*Nd doubles, host) */
typedef struct b200md_hac b200md_hac;
int b200md_hac_create(int number_of_steps, int sample_interval, int Nc, b200md_hac** out);
void b200md_hac_destroy(b200md_hac* p);
int b200md_hac_sample(
  b200md_hac* p, int step, int n, int stride, const double* d_virial, const double* d_velocity,
  void* stream);
int b200md_hac_finish(
  b200md_hac* p, double time_step, double temperature, double volume, double* hac_out,
  double* rtc_out, void* stream);
int b200md_hac_series(b200md_hac* p, double* out, void* stream);

/* Langevin thermostats and the Berendsen barostat:
 *   b200md_langevin_*         <- curand states + gpu_langevin + momentum correction,
 *                                src/integrate/langevin_utilities.cuh:26-127; one apply() = one
 *                                Ensemble_LAN::integrate_nvt_lan_half (ensemble_lan.cu:92-124, c1 =
 *                                exp(-0.5/Tc)) or Ensemble_BAO::integrate_nvt_lan (ensemble_bao.cu:91-120,
 *                                c1 = exp(-1/Tc)); c2 = sqrt((1-c1^2) k_B T).  cuRAND XORWOW,
 *                                curand_init(seed, atom, 0): seed 1804289383 (glibc's first rand())
 *                                reproduces the stream of the reference's -DDEBUG build when model.xyz
 *                                carries the velocities.
 *   b200md_baoab_operator     <- gpu_operator_A (which = 0) / gpu_operator_B (which = 1),
 *                                ensemble_bao.cu:190-300; d_group_label may be NULL (fixed_group -1)
 *   b200md_berendsen_pressure <- Ensemble_BER::compute2, NPT branch (ensemble_ber.cu:88-172,237-285):
 *                                num_components 1 (isotropic), 3 (orthogonal, optional deform) or 6
 *                                (triclinic, Voigt order); target_pressure / pressure_coupling in
 *                                natural units as integrate.cu:1150-1153 leaves them; h[9] (HOST) is
 *                                updated in place, positions are scaled on the device.  Reads
 *                                d_thermo[2..7] back, i.e. synchronises the stream, as the reference. */
typedef struct b200md_langevin b200md_langevin;
int b200md_langevin_create(int n, unsigned long long seed, b200md_langevin** out);
void b200md_langevin_destroy(b200md_langevin* p);
int b200md_langevin_apply(
  b200md_langevin* p, int n, int stride, double c1, double c2, const double* d_mass,
  double* d_velocity, void* stream);
int b200md_baoab_operator(
  int which, int n, int stride, double time_step, const double* d_mass, double* d_position,
  double* d_velocity, const double* d_force, const int* d_group_label, int fixed_group, void* stream);
int b200md_berendsen_pressure(
  int n, int stride, int num_components, const double target_pressure[6],
  const double pressure_coupling[6], const int deform[3], const double deform_rate[3],
  const int pbc[3], double h[9], const double* d_thermo, double* d_position, void* stream);

/* ------------------------------------------------------------------------------------------
 * Spatial-domain sharding (replaces the hub-and-spoke scatter/gather of NEP_MULTIGPU,
 * src/force/nep_multigpu.cu:1249-1310,1552-1582,1764-1802).  A rank keeps ONE set of local SoA
 * arrays of length `stride` = owned + ghost atoms; the first n entries are owned.  The *_strided
 * variants act on the owned part in place; b200md_halo_pack gathers the positions a neighbour
 * domain needs (out[d*m + k] = position[d*stride + index[k]] + shift[d]) into a send buffer for
 * ncclSend/Recv.  Thermo: call b200md_find_thermo_strided with the GLOBAL n_temperature and
 * volume; the 8 outputs are then additive over ranks (ncclAllReduce SUM).
 * b200md_nep_invalidate: the caller changed the local atom set/order (migration): drop the cell
 * order and lists; n_new <= the num_atoms given at creation.
 * ------------------------------------------------------------------------------------------ */
int b200md_apply_pbc_strided(
  int n, int stride, const double h[9], const int pbc[3], double* d_position, void* stream);
int b200md_velocity_verlet_strided(
  int is_step1, int n, int stride, double time_step, const double* d_mass, double* d_position,
  double* d_velocity, const double* d_force, void* stream);
int b200md_find_thermo_strided(
  int n, int stride, int n_temperature, double volume, const double* d_mass,
  const double* d_potential, const double* d_velocity, const double* d_virial, double* d_thermo8,
  void* d_scratch, void* stream);
int b200md_halo_pack(
  int m, const int* d_index, int stride, const double* d_position, const double shift[3],
  double* d_out, void* stream);
int b200md_nep_invalidate(b200md_nep* p, int n_new, void* stream);
/* Only the atoms with caller index < n_owned receive energy / force / virial from the following
 * b200md_nep_compute calls (0 = all atoms, the default): in a spatial domain the rest are ghosts,
 * whose results the owner rank computes and the local rank would discard
 * (cf. nep_multigpu.cu:1764-1802, which copies back the owned range only). */
int b200md_nep_set_owned(b200md_nep* p, int n_owned);
/* accumulate = 1 (default): outputs are ADDED to the caller's arrays, the contract of Potential::compute
 * (nep.cu:653,755-770; Force::compute zeroes them first, force.cu:794-801).  accumulate = 0: they are
 * OVERWRITTEN, so a driver with a single potential can drop its zeroing pass and the final kernel its
 * read-modify-write (208 bytes per atom and step).  Small boxes (supercell path) always accumulate. */
int b200md_nep_set_accumulate(b200md_nep* p, int accumulate);
/* Atoms whose position (in the coordinates passed to b200md_nep_compute) lies outside [lo, hi) get
 * EMPTY radial / angular neighbour sets from the following calls: no descriptor, dU/dq or partial-
 * force work is spent on them, they only serve as neighbours of the others.  A spatial domain sets
 * this to its owned region grown by rc + skin: ghosts further away cannot be a neighbour of an owned
 * atom, so nothing an owned atom's force depends on is skipped (the reference recomputes descriptors
 * for its whole ghost layer, nep_multigpu.cu:1476-1545).  lo = hi = NULL switches it off. */
int b200md_nep_set_active_region(b200md_nep* p, const double lo[3], const double hi[3]);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core self-test (no reference counterpart): one CTA computes D[128 x N] = A[128 x K] .
 * B[N x K]^T on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32 split, FP32
 * accumulators in TMEM) exactly as the NEP hidden-layer kernel issues it.  Row-major device
 * arrays; 16 <= N <= 256, N % 16 == 0, K % 8 == 0; layout 0 / 1 = the two shared-memory operand
 * arrangements of b2_tc.cuh (1 is the hidden layer's).  Used by tests/ to validate the descriptor
 * and TMEM plumbing against a host GEMM.
 * ------------------------------------------------------------------------------------------- */
int b200md_tc_selftest(
  int layout, int N, int K, const float* d_A, const float* d_B, float* d_D, void* stream);

#ifdef __cplusplus
}
#endif
#endif
