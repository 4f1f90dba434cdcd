/*
 * C-ABI of the B200-native VSVB energy engine (libvalence_b200.so).
 *
 * Two layers:
 *
 * (1) The reference's external Fortran entry points, with gfortran name mangling
 *     (lower case + trailing underscore, every argument by reference, default
 *     integer = int32, real(dp) = double).  A host that links or dlopen()s the
 *     reference's libvalence.so can switch to this library unchanged:
 *       valence_api_initialize_        replaces /root/reference/src/valence_api.F90:9-32
 *       valence_api_calculate_energy_  replaces /root/reference/src/valence_api.F90:37-108
 *       valence_api_finalize_          replaces /root/reference/src/valence_api.F90:114-134
 *       init_, getn_, calcsurface_, finalize_
 *                                      replace /root/reference/src/valence_api_nitrogen.F90:4-36
 *     (NITROGEN's USE_FORTRAN_PES interface, nitrogen/h2o/VSVB_STAT.job:8,29).
 *
 * (2) The engine seam: the argument list of the reference's internal hot routine
 *     vsvb_energy / guess_energy (/root/reference/src/valence.F90:309-345,1010-1011)
 *     lifted to a handle-based C interface, so a Fortran host can bind it with
 *     ISO_C_BINDING (INTEGRATION.md shows the interface block).
 *
 * Error behaviour: layer (1) follows the reference (xm_abort: print an `error`
 * line and stop the process, /root/reference/src/xm_module.F90:942-950); layer
 * (2) returns non-zero and keeps a message for vb_last_error().
 * There is no CPU fallback: without a usable CUDA device both layers fail.
 */
#ifndef VALENCE_B200_H
#define VALENCE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- (1) reference-compatible entry points ------------------------------------------------ */
/* Reads the input file named by argv[1] of the host process (valence_initialize_module.F90:49);
 * the environment variable VALENCE_INPUT, when set, names the file instead (for hosts such as
 * Python whose argv is not ours).  *info = 0 on return, like the reference.
 * Parallel runs: launch one process per GPU (torchrun / mpirun / srun); rank, size and local rank are read from the
 * launcher's environment, the ranks form an NCCL communicator and every later call is collective (the reference's
 * `call_mpi_init` / `comm` hand-off, valence_api.F90:22-27, has no MPI to act on here and is accepted unused). */
void valence_api_initialize_(int* info, int* call_mpi_init, int* comm);
/* x: 3*natom cartesians, atom-major, Angstrom.  v: total VSVB energy in Hartree incl. nuclear
 * repulsion.  Prints the reference's stdout lines and rewrites the `orbitals` file. */
void valence_api_calculate_energy_(double* x, double* v);
void valence_api_finalize_(int* call_mpi_finalize);
void init_(int* info);
void getn_(int* n);
void calcsurface_(double* x, double* v);   /* v in cm^-1 = Hartree * 219474.631 */
void finalize_(void);
/* geometry (Angstrom) as read from the input file; lets a driver evaluate "the input as is" */
const double* vb_api_input_coords(void);

/* ---- (2) engine seam ------------------------------------------------------------------------ */
typedef struct vb_engine vb_engine;

enum { VB_CNT_SCHWARZ_EREP = 0, VB_CNT_SCHWARZ_EXCH, VB_CNT_VALUE_EREP, VB_CNT_VALUE_EXCH, VB_CNT_INT2E,
       VB_CNT_SHELL_QUARTETS, VB_CNT_SHORTCUT, VB_CNT_ENTRIES, VB_CNT_N };

typedef struct vb_energy_result {
    double energy;      /* numerator / wfnorm + enucrep          (valence.F90:344) */
    double enucrep;     /* nuclear repulsion                     (tools_module.F90:19-40) */
    double numerator;   /* <Psi|H_el|Psi> as vsvb_energy returns (valence.F90:1062-1430) */
    double wfnorm;      /* <Psi|Psi>                             (valence.F90:1106) */
    double e1, e2;      /* one- and two-electron parts of the numerator */
    long long counters[VB_CNT_N];   /* screening / quartet counts with the reference's task semantics */
    long long n_entries, n_groups, n_pairgroups, n_tiles, n_tiles_mine;
    long long n_ao_quartets, n_prim_quartets;
    double flops_model;
    long long ref_shell_quartets;
    double t_total_ms, t_host_setup_ms, t_1e_ms, t_density_ms, t_diag_ms, t_tiles_ms;
    int launches, diag_launches, tile_launches;
    double min_pivot_ratio;
    long long h2d_bytes, d2h_bytes;   /* host<->device bytes copied by this call */
    double flops_transform;           /* FP64 tensor-core flops of the two density transforms of this rank's tile pass */
} vb_energy_result;

const char* vb_last_error(void);
int vb_engine_create(const char* input_path, int device, vb_engine** out);
void vb_engine_destroy(vb_engine* e);
int vb_engine_natom(const vb_engine* e);
int vb_engine_nelec(const vb_engine* e);
/* number of expansion terms of 1-based orbital iorb = order of the first_order_opt matrices (-1: no such orbital) */
int vb_engine_norbas(const vb_engine* e, int iorb);
/* One process per GPU of a node (replaces xm_propagate / xm_equalize*, /root/reference/src/xm_module.F90:711-910): the ranks
 * of the job form an NCCL communicator (libnccl is loaded at run time; the id travels through
 * /dev/shm/valence_b200_nccl_<key>, key NULL/"" = derived from the launcher's environment).  Afterwards vb_engine_energy,
 * vb_engine_first_order and vb_engine_run are COLLECTIVE: the tile pass is sharded, the packed accumulators are summed by
 * one all-reduce per energy (ham: one per orbital), every rank returns the full result.  vb_engine_attach_nccl adopts a
 * communicator of the host (an ncclComm_t) instead of creating one. */
int vb_engine_attach_comm(vb_engine* e, int rank, int nranks, const char* key);
int vb_engine_attach_nccl(vb_engine* e, int rank, int nranks, void* nccl_comm);
int vb_engine_set_coords(vb_engine* e, const double* x_angstrom);
/* guess_energy on one GPU */
int vb_engine_energy(vb_engine* e, vb_energy_result* out);
/* sharded form: every rank calls _partial(rank, nranks); the caller sums the vb_engine_accum_len()
 * doubles at vb_engine_accum_device() over ranks (one NCCL all-reduce); every rank calls _finish. */
int vb_engine_energy_partial(vb_engine* e, int rank, int nranks, vb_energy_result* out);
int vb_engine_energy_finish(vb_engine* e, vb_energy_result* out);
/* optional, ranks of ONE node, before _partial: build this rank's share of the host tables and publish it as the file
 * <prefix><rank> (use a /dev/shm path unique to the job); barrier; then _partial merges all shares instead of every
 * rank building everything (the reference's ranks likewise split the set-up work by task index, valence.F90:1162). */
int vb_engine_shard_tables(vb_engine* e, int rank, int nranks, const char* prefix);
/* first_order_opt matrices (valence.F90:527-764) of 1-based orbital iorb: ham/ovl receive the
 * norbas x norbas column-major matrices <Psi[chi_ib]|H_el|Psi[chi_jb]>, <Psi[chi_ib]|Psi[chi_jb]>
 * (numerators: not divided by the norm, no nuclear repulsion); cap = doubles available in each. */
int vb_engine_first_order(vb_engine* e, int iorb, double* ham, double* ovl, int cap, int* norbas, vb_energy_result* stats);
/* sharded form (one process per GPU): the two-electron part of ham is this rank's share of the tiles (the
 * one-electron part is added on rank 0 only); the caller sums ham over the ranks -- one all-reduce of
 * norbas^2 doubles, the counterpart of xm_equalize(ham) in first_order_opt (valence.F90:755-764); ovl is
 * complete on every rank. */
int vb_engine_first_order_sharded(vb_engine* e, int iorb, int rank, int nranks, double* ham, double* ovl, int cap, int* norbas,
                                  vb_energy_result* stats);
/* calculate_vsvb_energy (valence.F90:28-302): guess energy and, if the input asks for it, the
 * first-order orbital optimisation + spin-coupling optimisation of minimize_energy (:2744-2885) */
int vb_engine_run(vb_engine* e, int print, double* enucrep, double* guess_energy, double* total_energy, int* converged, int* iterations);
/* xm_output (xm_module.F90:464-577) without a GPU: writes `orbitals` (and `nelecwfn` when the input has several
 * spin couplings) for the wavefunction of an input file into the current directory, as the reference does after the
 * guess energy (valence.F90:192) and at convergence (:2882, converged != 0 adds the "converged to" line). */
int vb_write_wavefunction_files(const char* input_path, double energy, int converged);
double* vb_engine_accum_device(const vb_engine* e);
int vb_engine_accum_len(const vb_engine* e);
void* vb_engine_stream(const vb_engine* e);
/* debugging aid: per-tile energy partials of the last tile pass (returns their count) */
long long vb_engine_debug_tile_energies(const vb_engine* e, double* out, long long cap);
/* measured FP64 FMA peak of the device in TFLOP/s (roofline denominator of the ERI kernel) */
int vb_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
