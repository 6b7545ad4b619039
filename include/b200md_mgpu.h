/*
 * b200md_mgpu.h -- C-ABI of libb200md_mgpu.so: spatial-domain sharding of the MD hot path over
 * several GPUs, host side in C++, device side CUDA + NCCL.
 *
 * Replaces NEP_MULTIGPU::compute (src/force/nep_multigpu.cu:1416-1803) and the single-GPU
 * integration around it: there GPU 0 owns all atoms, scatters positions / gathers forces through
 * blocking peer copies every step, rebuilds every list every step and integrates everything itself.
 * Here the box is cut into a Px x Py x Pz grid of BLOCKS (slabs = P x 1 x 1); every rank owns the
 * atoms of one block, integrates them where they live, and receives only the FP64 POSITIONS of the
 * ghost atoms within `halo` of its faces, by a staged (x, then y, then z) ncclSend/ncclRecv exchange
 * per force evaluation.  As in the reference (nep_multigpu.cuh:42-53) the halo of a many-body
 * potential is two cutoffs wide, so descriptors of the first ghost layer are recomputed locally and no
 * force travels back.  Migration and the rebuild of the ghost lists are device-side (flag kernels +
 * CUB compaction + packed NCCL messages) and happen only when some atom has moved further than
 * 0.7 * skin / 2 since the last one.  Thermo is an 8-double ncclAllReduce; thermostat state is
 * replicated.
 *
 * One GROUP = the domains hosted by this process:
 *   - distributed: one process per GPU (torchrun / mpirun style), exactly one domain per group,
 *     nccl_id = the 128 bytes of an ncclUniqueId shared by all ranks (b200md_mgpu_unique_id);
 *   - local: nccl_id = NULL and the group hosts ALL Px*Py*Pz domains on the current device, run in
 *     lock step with device-to-device copies instead of NCCL -- the same kernels, lists and
 *     ordering, testable on one GPU.
 * Functions return 0 or a B200MD_ERR_* code (b200md.h); b200md_mgpu_last_error() describes it.
 */
#ifndef B200MD_MGPU_H
#define B200MD_MGPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200md_mgpu b200md_mgpu;

typedef struct {
  double h[9];         /* global box, Box::cpu_h[0..8]; must be orthogonal */
  int pbc[3];          /* decomposed directions must be periodic */
  int grid[3];         /* Px, Py, Pz; world size = Px*Py*Pz */
  int rank;            /* distributed mode: this process's rank in [0, world) */
  const char* potential_file; /* nep*, lj, tersoff_1989, eam_zhou_2004 / eam_dai_2006 */
  double skin;         /* ghost-list skin in A (>= 1); migration when an atom moved 0.7*skin/2 */
  /* ensemble: 0 nve, 1 nvt_ber, 2 nvt_nhc, 4 nvt_bdp (the reference's type codes) */
  int ensemble;
  double temperature, temperature_coupling, time_step; /* time_step in natural units */
  unsigned bdp_seed;
  double capacity_factor; /* local arrays hold capacity_factor * (initial local atoms) + 1024 */
  int use_cuda_graph;     /* 1: replay each step as one CUDA graph (captured after migrations) */
} b200md_mgpu_config;

const char* b200md_mgpu_last_error(void);
/* fills 128 bytes with a fresh ncclUniqueId (call on rank 0, broadcast to the others) */
int b200md_mgpu_unique_id(char out128[128]);

int b200md_mgpu_create(const b200md_mgpu_config* cfg, const char* nccl_id128, b200md_mgpu** out);
void b200md_mgpu_destroy(b200md_mgpu* g);

/* Every rank passes the same GLOBAL host arrays (type[n], position[3n] SoA, mass[n], velocity[3n]
 * SoA or NULL); each domain keeps what it owns, then ghosts are exchanged and forces evaluated. */
int b200md_mgpu_distribute(
  b200md_mgpu* g, int n_global, const int* type, const double* position, const double* mass,
  const double* velocity);

/* nsteps velocity-Verlet steps (Run::perform_a_run's loop body, run.cu:259-295), migration
 * included.  Asynchronous apart from the displacement check every `check_every` steps. */
int b200md_mgpu_run(b200md_mgpu* g, int nsteps, int check_every);

/* the same, bracketed by CUDA events on the module's stream: *ms_out = device time of the nsteps
 * (synchronises before and after) */
int b200md_mgpu_run_timed(b200md_mgpu* g, int nsteps, int check_every, double* ms_out);

/* global thermo[0..7] = T, U, sxx, syy, szz, sxy, sxz, syz (synchronises) */
int b200md_mgpu_thermo(b200md_mgpu* g, double out8[8]);
/* global heat current jx_in, jx_out, jy_in, jy_out, jz (compute_heat.cu:32-90 summed, hac.cu:51,106) */
int b200md_mgpu_heat_current(b200md_mgpu* g, double out5[5]);

/* counters: what = 0 local domains, 1 owned atoms of domain k, 2 local (owned+ghost) atoms of
 * domain k, 3 migrations so far, 4 neighbour rebuilds of domain k, 5 global rank of domain k,
 * 6 kernels launched per step (this module's + libb200md's, counted on the last eager step) */
long long b200md_mgpu_info(b200md_mgpu* g, int what, int k);
/* owned atoms of local domain k in GLOBAL coordinates: id[n], position[3n], velocity[3n],
 * force[3n], potential[n], virial[9n] (host arrays; any may be NULL) */
int b200md_mgpu_get_owned(
  b200md_mgpu* g, int k, long long* id, double* position, double* velocity, double* force,
  double* potential, double* virial);
/* the LOCAL system of domain k as the potential sees it: type[n_loc], position[3*n_loc] in the local
 * frame (owned atoms first), local box h[9] and pbc[3] (host arrays; any may be NULL) */
int b200md_mgpu_get_local(
  b200md_mgpu* g, int k, int* type, double* position, double h_out[9], int pbc_out[3]);
/* per-phase device time of the last b200md_mgpu_run when profiling was on (ms per step):
 * 0 vv1+wrap, 1 halo, 2 force, 3 vv2+thermo; returns the number of phases written */
int b200md_mgpu_profile(b200md_mgpu* g, int enable, double* ms_out, int max_out);
/* latched device-side neighbour-capacity errors of every local domain */
int b200md_mgpu_check(b200md_mgpu* g);

#ifdef __cplusplus
}
#endif
#endif
