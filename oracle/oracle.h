/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference GPUMD hot path, used as the
 * parity checker by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  Nothing in the product path (gpumd_b200/) may include,
 * link or call this.  Every function cites the reference file:line
 * (relative to /root/reference) whose behaviour it restates.
 *
 * Layout conventions are GPUMD's (src/model/atom.cuh:21-52):
 *   position[3N] = x[N],y[N],z[N];  force[3N] likewise;
 *   virial[9N]   = xx,yy,zz,xy,xz,yz,yx,zx,zy blocks of N (force.cu:859-861);
 *   box h[9]     = cpu_h[0..8] of src/model/box.cuh:18-35, i.e. row-major with
 *                  the lattice vectors a,b,c as COLUMNS (h[0],h[3],h[6] = a).
 */
#ifndef B200MD_ORACLE_H
#define B200MD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_TYPES 94

typedef struct oracle_nep_model oracle_nep_model;

/* nep.txt parser -- restates NEP::NEP, src/force/nep.cu:100-395. Returns NULL on error. */
oracle_nep_model* oracle_nep_load(const char* path);
void oracle_nep_free(oracle_nep_model* m);
/* query: 0 num_types, 1 dim, 2 num_neurons, 3 n_max_radial, 4 n_max_angular, 5 basis_r,
 * 6 basis_a, 7 L_max, 8 num_L, 9 MN_radial(enlarged), 10 MN_angular(enlarged), 11 zbl_enabled,
 * 12 version */
int oracle_nep_info(const oracle_nep_model* m, int what);
double oracle_nep_rc_radial_max(const oracle_nep_model* m);
double oracle_nep_rc_angular_max(const oracle_nep_model* m);

/*
 * Single-point NEP evaluation.  precision: 32 = FP32 pair math exactly as the GPU
 * reference (src/force/nep.cu:436-975, src/utilities/nep_utilities.cuh), 64 = all-FP64.
 * Outputs (any may be NULL): pe[N], force[3N], virial[9N] (overwritten, not accumulated),
 * q[dim*N] scaled descriptors (q[d*N+i]), NN_r[N], NL_r[N*mn_r] (row i holds its ascending
 * neighbour indices), NN_a[N], NL_a[N*mn_a].  mn_r / mn_a are the caller's row capacities.
 * Returns 0, or <0 on error (capacity exceeded, bad input).
 */
int oracle_nep_compute(
  const oracle_nep_model* m, int precision, int N, const int* type, const double h[9],
  const int pbc[3], const double* position, double* pe, double* force, double* virial,
  double* q, int* NN_r, int* NL_r, int mn_r, int* NN_a, int* NL_a, int mn_a);

/* Cutoff-neighbour sets with the reference's FP32 membership test (neighbor.cu:136-152):
 * all j != i with d2 < rc*rc, ascending j, row-major NL[i*mn + k].  Large-box (minimum image)
 * semantics only.  Returns 0 or <0. */
int oracle_neighbor_list(
  int N, const double h[9], const int pbc[3], const double* position, double rc, int* NN, int* NL,
  int mn);

/* LJ -- restates src/force/lj.cu:28-181.  para[(t1*nt+t2)*3 + {0,1,2}] = eps, sigma, cutoff. */
int oracle_lj_compute(
  int nt, const double* para, int N, const int* type, const double h[9], const int pbc[3],
  const double* position, double* pe, double* force, double* virial);

/* Tersoff-1989 (1 or 2 types), all FP64 -- restates src/force/tersoff1989.cu:31-586 with the
 * FP64 many-body reduction src/force/potential.cu:35-134.  para: per type 11 numbers
 * a b lambda mu beta n c d h r1 r2 (file order), then chi when nt == 2. */
int oracle_tersoff_compute(
  int nt, const double* para, int N, const int* type, const double h[9], const int pbc[3],
  const double* position, double* pe, double* force, double* virial);

/* Analytic EAM -- restates src/force/eam.cu:28-475.  model 0 = eam_zhou_2004 (21 numbers per
 * type, nt <= 18), model 1 = eam_dai_2006 (9 numbers, one type). */
int oracle_eam_compute(
  int model, int nt, const double* para, int N, const int* type, const double h[9],
  const int pbc[3], const double* position, double* pe, double* force, double* virial);

/* per-atom heat current J_i = W_i . v_i split as in gpu_compute_heat,
 * src/measure/compute_heat.cu:32-63: heat[5N] = jx_in, jx_out, jy_in, jy_out, jz */
void oracle_compute_heat(int N, const double* virial, const double* velocity, double* heat);

/* Force::compute's pre-step, src/force/force.cu:424-459: wrap positions into the box. */
void oracle_apply_pbc(int N, const double h[9], const int pbc[3], double* position);

/* gpu_velocity_verlet, src/integrate/ensemble.cu:176-214. */
void oracle_velocity_verlet(
  int is_step1, int N, double dt, const double* mass, double* position, double* velocity,
  const double* force);

/* gpu_velocity_verlet with `fix` / `move` groups, src/integrate/ensemble.cu:111-174. */
void oracle_velocity_verlet_groups(
  int is_step1, int N, double dt, const double* mass, double* position, double* velocity,
  const double* force, const int* group_label, int fixed_group, int move_group,
  const double* move_velocity);

/* gpu_find_thermo_instant_temperature, src/integrate/ensemble.cu:434-633:
 * thermo[0..7] = T, U, sxx, syy, szz, sxy, sxz, syz. */
void oracle_find_thermo(
  int N, int N_temperature, double volume, const double* mass, const double* pe,
  const double* velocity, const double* virial, double* thermo8);

#ifdef __cplusplus
}
#endif
#endif
