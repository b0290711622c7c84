/* legolas_b200 — C ABI of the B200-native Legolas hot path.
 *
 * The reference (Legolas 2.0.6, pure Fortran) has no FFI for this path; the boundary
 * is two Fortran call sites in `program legolas`:
 *
 *   call build_matrices(matrix_B, matrix_A, settings, grid, background, physics)
 *        src/main.f08:68, src/matrices/mod_matrix_manager.f08:138-266
 *   call solve_evp(matrix_A, matrix_B, settings, omega, right_eigenvectors)
 *        src/main.f08:74 -> src/solvers/mod_solvers.f08:111-112
 *        -> src/solvers/arnoldi/smod_arpack_main.f08:67-77
 *        -> src/solvers/arnoldi/smod_arpack_shift_invert.f08:15-161
 *
 * A thin iso_c_binding shim (legolas_b200/fortran/) replaces those two
 * bodies with calls into this library.  Conventions: all pointers are caller-owned HOST
 * pointers unless a function name ends in `_device`; arrays are column-major;
 * `complex(dp)` is two interleaved doubles (re, im); Fortran `logical` is int32_t;
 * integers are int32_t; functions return 0 on success, a negative LGPU_E* code on an
 * argument / runtime error (message via lgpu_last_error) and never call exit().
 * A context is not re-entrant; different contexts may be used from different threads.
 * There is NO CPU fallback: every entry point fails with LGPU_ENOGPU without a device.
 */
#ifndef LEGOLAS_B200_H
#define LEGOLAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGPU_OK 0
#define LGPU_EINVAL (-1)   /* invalid argument (logger%error in the reference)       */
#define LGPU_ENOGPU (-2)   /* no CUDA device / CUDA runtime error                    */
#define LGPU_ESTATE (-3)   /* call order violated (e.g. solve before assemble)       */
#define LGPU_ENOMEM (-4)

#define LGPU_N_FIELDS 38

/* Slot order of the equilibrium / physics arrays sampled at grid%gaussian_grid.
 * Each slot is 4*(gridpts-1) doubles; a NULL slot means "identically zero"
 * (the reference's default `zero_func`, src/background/mod_background.f08:44-131). */
enum lgpu_field {
  LGPU_F_RHO0 = 0, LGPU_F_DRHO0, LGPU_F_T0, LGPU_F_DT0, LGPU_F_DDT0,
  LGPU_F_B01, LGPU_F_B02, LGPU_F_DB02, LGPU_F_DDB02, LGPU_F_B03, LGPU_F_DB03, LGPU_F_DDB03,
  LGPU_F_V01, LGPU_F_DV01, LGPU_F_DDV01, LGPU_F_V02, LGPU_F_DV02, LGPU_F_DDV02,
  LGPU_F_V03, LGPU_F_DV03, LGPU_F_DDV03,
  LGPU_F_G0, LGPU_F_ETA, LGPU_F_DETADT, LGPU_F_DETADR,
  LGPU_F_L0, LGPU_F_DLDT, LGPU_F_DLDRHO,           /* heatloss%get_L0 / get_dLdT / get_dLdrho */
  LGPU_F_TCPARA, LGPU_F_DTCPARADT, LGPU_F_TCPERP, LGPU_F_DTCPERPDRHO, LGPU_F_DTCPERPDT,
  LGPU_F_DTCPERPDB2, LGPU_F_TCPREFACTOR, LGPU_F_DTCPREFACTORDR,
  LGPU_F_HALLFACTOR, LGPU_F_INERTIAFACTOR
};

/* Scalars of settings_t / mod_equilibrium_params read by build_matrices and the
 * boundary manager (src/boundaries/mod_boundary_manager.f08:87-107). */
typedef struct {
  int32_t gridpts;            /* settings%grid%get_gridpts()                              */
  int32_t physics_type;       /* 0 "mhd" (8 eqs), 1 "hd" (5: rho,v1,v2,v3,T), 2 "hd-1d" (3: rho,v1,T);
                                 N = gridpts * 2 * nb_eqs and every vector / index crossing this ABI
                                 uses that numbering (src/settings/mod_settings.f08:69-86)        */
  int32_t geometry;           /* 0 Cartesian (eps=1, deps=0), 1 cylindrical (eps=x, deps=1)*/
  int32_t incompressible;     /* gamma := 1e12, src/settings/mod_physics_settings.f08:90  */
  int32_t flow, resistivity, cooling, heating, conduction, perpendicular_conduction;
  int32_t viscosity, viscous_heating, hall, electron_inertia, gravity;
  int32_t boundary_type;      /* 0 "wall", 1 "wall_weak"                                  */
  int32_t coaxial;            /* settings%grid%coaxial                                    */
  int32_t reserved;
  double k2, k3;              /* mod_equilibrium_params                                   */
  double gamma;               /* ratio of specific heats (ignored if incompressible)      */
  double viscosity_value;     /* settings%physics%viscosity%get_viscosity_value()         */
  double electron_fraction;   /* settings%physics%hall%get_electron_fraction()            */
  double gauss_nodes[4];      /* mod_global_variables.f08:31-43; all zero = 2.0.6 values   */
  double gauss_weights[4];
} lgpu_settings;

/* arpack_t after new_arpack_config (src/solvers/arnoldi/mod_arpack_type.f08:74-102):
 * the host has already validated nev/ncv/which and resolved the ncv/maxiter/tol defaults. */
typedef struct {
  int32_t nev, ncv, maxiter;
  char which[2];              /* "LM","SM","LR","SR","LI","SI"                            */
  int16_t pad;
  double tol;
  double sigma_re, sigma_im;
  int32_t refine_steps;       /* extra iterative-refinement sweeps per solve (default 0)  */
  int32_t reserved;
} lgpu_arnoldi;

/* What the host copies back into arpack_cfg%info / iparam(5,9,10,11) so that
 * parse_znaupd_info / parse_zneupd_info / parse_finished_stats work unchanged. */
typedef struct {
  int32_t info;               /* znaupd-style: 0 ok, 1 maxiter reached, 3 no shifts, -9 zero start vector */
  int32_t nconv;              /* iparam(5)                                                */
  int32_t n_op;               /* iparam(9)  OP*x applications                             */
  int32_t n_bx;               /* iparam(10) (always 0: bmat = "I")                        */
  int32_t n_reorth;           /* iparam(11) re-orthogonalisation passes                   */
  int32_t n_restart;          /* iparam(3) on exit: Arnoldi update iterations taken       */
  int32_t lu_info;            /* 0 or index (1-based block row) of a singular pivot block */
  int32_t reserved;
  double t_factor_ms, t_iter_ms, t_extract_ms;   /* device/host wall times of the phases  */
} lgpu_stats;

typedef struct lgpu_ctx lgpu_ctx;

int lgpu_create(lgpu_ctx** ctx, int32_t device, int32_t log_level);
int lgpu_destroy(lgpu_ctx* ctx);
const char* lgpu_last_error(const lgpu_ctx* ctx);
/* Run all work of this context on an existing CUDA stream (cudaStream_t); NULL = own stream. */
int lgpu_set_stream(lgpu_ctx* ctx, void* cuda_stream);
int lgpu_synchronize(lgpu_ctx* ctx);
/* Several contexts of one process on one device (one host thread per context, e.g. a parameter sweep that keeps a few
 * small units in flight per GPU - the reference's counterpart is the process pool of pylbo's runner,
 * post_processing/pylbo/automation/runner.py:202-211).  Some kernels wait for other CTAs of their own launch (the
 * fused Gram-Schmidt step has a device-wide barrier, the upper solve stages hand rows over through mailboxes), so the
 * library admits calls of different contexts concurrently only while the SMs they may hold waiting fit the device
 * together; a call that needs the whole device runs alone.  max_sms caps the CTAs of this context's Gram-Schmidt step
 * (default 0: one per SM, i.e. no other context runs beside it); a sweep with k units in flight sets it to SMs / k. */
int lgpu_set_sm_limit(lgpu_ctx* ctx, int32_t max_sms);

/* ---- replaces build_matrices (src/matrices/mod_matrix_manager.f08:138-266) -----------
 * base_grid: gridpts doubles; gauss_grid: 4*(gridpts-1) doubles; fields: LGPU_N_FIELDS
 * host pointers (NULL = zero).  A and B stay resident on the device in block-tridiagonal
 * form.  `_device` variant: the same pointers are device pointers (inputs already in HBM). */
int lgpu_assemble(lgpu_ctx* ctx, const lgpu_settings* settings, const double* base_grid,
                  const double* gauss_grid, const double* const fields[LGPU_N_FIELDS]);
int lgpu_assemble_device(lgpu_ctx* ctx, const lgpu_settings* settings, const double* base_grid,
                         const double* gauss_grid, const double* const fields[LGPU_N_FIELDS]);

/* matrix_t views of the device matrices (which: 0 = A, 1 = B).
 * export_coo: triplets in the reference's output order (rows ascending, per-row insertion
 * order; src/dataIO/mod_output.f08:476-508), 1-based; call with rows == NULL to get nnz.
 * export_blocks: raw (gridpts, 3, d, d) complex blocks [sub, diag, super], each d x d block
 * column-major, d = 2*nb_eqs.
 * import_coo: load host-assembled matrices instead of lgpu_assemble (solve_evp on an
 * arbitrary matrix_t pair whose half-bandwidth fits the block-tridiagonal envelope);
 * n must be a multiple of 16. */
int lgpu_matrix_dim(lgpu_ctx* ctx, int32_t* n);
int lgpu_export_coo(lgpu_ctx* ctx, int32_t which, int64_t* nnz, int32_t* rows, int32_t* cols,
                    double* vals_ri);
int lgpu_export_blocks(lgpu_ctx* ctx, int32_t which, double* blocks_ri);
int lgpu_import_coo(lgpu_ctx* ctx, int32_t which, int32_t n, int64_t nnz, const int32_t* rows,
                    const int32_t* cols, const double* vals_ri);

/* ---- pieces of solve_arpack_shift_invert, exposed for tests and other solvers ---------
 * factorize: A - sigma*B -> block-cyclic-reduction factors (replaces zgbtrf,
 *            src/solvers/mod_linear_systems.f08:102-127); lu_info as in lgpu_stats.
 * solve:     x = (A - sigma*B)^-1 rhs   (replaces zgbtrs, mod_linear_systems.f08:67-97)
 * matvec:    y = A*x or B*x             (replaces zgbmv, mod_banded_operations.f08:18-41)
 * apply_op:  y = (A - sigma*B)^-1 B x   (one reverse-communication step, shift_invert :98-104) */
int lgpu_factorize(lgpu_ctx* ctx, double sigma_re, double sigma_im, int32_t* lu_info);
int lgpu_solve(lgpu_ctx* ctx, const double* rhs_ri, double* x_ri, int32_t refine_steps);
int lgpu_matvec(lgpu_ctx* ctx, int32_t which, const double* x_ri, double* y_ri);
int lgpu_apply_op(lgpu_ctx* ctx, const double* x_ri, double* y_ri, int32_t refine_steps);
/* The same step with device-resident vectors, `repeat` times back to back on the context's stream (every application
 * reads x_dev and writes y_dev; x_dev != y_dev), as the Arnoldi driver issues it; *ms_per_application (may be NULL) =
 * CUDA-event time of the batch / repeat.  This is how the operator application is timed for the roofline: event pairs
 * around single launches serialise the programmatic dependent launches that overlap a kernel's prologue with its
 * predecessor. */
int lgpu_apply_op_device(lgpu_ctx* ctx, const double* x_dev, double* y_dev, int32_t refine_steps, int32_t repeat,
                         double* ms_per_application);

/* ---- replaces solve_arpack_shift_invert (smod_arpack_shift_invert.f08:15-161) ---------
 * resid0_ri: start vector (N complex) = zlarnv(2, [2022,9,30,179], N) from the host
 *            (mod_arpack_type.f08:194-211); omega_ri: nev complex, already back-transformed
 *            omega = sigma + 1/nu (:157), entries nconv..nev-1 are NaN; vr_ri: N x nev complex,
 *            ld = N, unit 2-norm Ritz vectors (may be NULL to skip the device->host copy). */
int lgpu_shift_invert(lgpu_ctx* ctx, const lgpu_arnoldi* cfg, const double* resid0_ri,
                      double* omega_ri, double* vr_ri, lgpu_stats* stats);
/* Same, but the start vector is taken from / Ritz vectors are left in device memory
 * (used for device-resident timing and for chaining into device-side consumers). */
int lgpu_shift_invert_device(lgpu_ctx* ctx, const lgpu_arnoldi* cfg, const double* resid0_dev,
                             double* omega_ri_host, double* vr_dev, lgpu_stats* stats);

/* ---- row N3 of the scope table: ARPACK "general" mode ---------------------------------------
 * Replaces solve_arpack_general (src/solvers/arnoldi/smod_arpack_general.f08:14-131), reached from
 * solve_evp with solver = "arnoldi", arpack_mode = "general" (smod_arpack_main.f08:57-65: mode = 1,
 * bmat = "I"): the standard problem OP x = omega x with OP = B^-1 A.  The reference re-factorises B
 * on every operator application (zgbsv, src/solvers/mod_linear_systems.f08:33-62); here B is
 * factorised once per call.  Same arguments and outputs as lgpu_shift_invert; cfg->sigma_* are
 * ignored and omega are ARPACK's Ritz values themselves (no back-transformation).  The resident
 * factorisation is dropped afterwards (lgpu_solve / lgpu_apply_op need a new lgpu_factorize). */
int lgpu_arnoldi_general(lgpu_ctx* ctx, const lgpu_arnoldi* cfg, const double* resid0_ri,
                         double* omega_ri, double* vr_ri, lgpu_stats* stats);

/* Page-locked host memory for callers that want the eigenvector read-back (N x nev complex,
 * 51 MB at the headline size) at full PCIe speed instead of the pageable-memory staging path.
 * Any host pointer is accepted by every entry point; these are optional. */
void* lgpu_host_alloc(size_t bytes);
void lgpu_host_free(void* ptr);

/* ---- rows N1 / N4 of the scope table (consumers of the factorisation and the matvecs) -------
 * residuals: res[k] = || A v_k - omega_k B v_k ||_2 / || omega_k v_k ||_2, 0 where omega_k is zero
 *            by the reference's is_zero rule (get_residual, src/dataIO/mod_output.f08:511-545);
 *            vr_ri: N x nev complex host, ld = N.
 * inverse_iteration: replaces inverse_iteration (src/solvers/smod_inverse_iteration.f08:16-205):
 *            LU of A - sigma B, then x <- normalised (A - sigma B)^-1 B x until
 *            || A x - ev B x || < |ev| tol with ev = x^H A x / x^H B x, at most maxiter solves
 *            (0 = the reference's default 100).  omega_ri: 1 complex; vr_ri: N complex, largest
 *            entry made real, or NULL; stats->info = 0 converged, 1 maxiter reached;
 *            stats->n_op = solves.  Deviation: the start vector is (A - sigma B)^-1 1 instead of
 *            LAPACK's U^-1 1 (the factors are not LAPACK's), and B is applied as stored, not
 *            through zhbmv's Hermitian completion of its upper triangle. */
int lgpu_residuals(lgpu_ctx* ctx, int32_t nev, const double* omega_ri, const double* vr_ri, double* res);
/* eigenfunctions: replaces base_ef_t%assemble (src/eigenfunctions/mod_base_efs.f08:35-61), i.e.
 *            assemble_eigenfunction + retransform_eigenfunction (mod_ef_assembly.f08:16-106) for all 8
 *            variables and the selected eigenvectors.  vr_ri: N x (max idx) complex host, ld = N;
 *            idxs: nsel 1-based column indices (idxs_to_assemble); out_ri: complex
 *            [nb_eqs][nsel][2*gridpts-1], i.e. quantities(:, i) of variable p (state-vector order) at
 *            ((p*nsel + i)*npts). */
int lgpu_eigenfunctions(lgpu_ctx* ctx, const double* vr_ri, int32_t nsel, const int32_t* idxs, double* out_ri);
int lgpu_inverse_iteration(lgpu_ctx* ctx, double sigma_re, double sigma_im, int32_t maxiter, double tol,
                           double* omega_ri, double* vr_ri, lgpu_stats* stats);

/* Host utility: LAPACK zlarnv(idist=2) (uniform (-1,1) re and im), bit-exact port of
 * dlaruv's 48-bit multiplicative congruential generator; iseed[4] updated in place. */
int lgpu_zlarnv(int32_t iseed[4], int32_t n, double* out_ri);

/* Launch / traffic accounting since the last call (for bench.py's gpu_launches claim). */
int lgpu_counters(lgpu_ctx* ctx, int64_t* kernel_launches, int32_t reset);
/* Optional per-kernel-class device timing: CUDA events are recorded on the context's stream
 * around every launch; lgpu_profile_read synchronises and returns the accumulated
 * milliseconds, launch counts and algorithmic bytes per class, in this order (LGPU_N_KINDS entries):
 * assemble, factor, matvec, fwd_stage0, fwd_stage, top_stage (the fused upper solve stages and the
 * top system), bwd_stage, bwd_stage0, dots, update, scale, gemm, cgs2_step (the fused
 * Gram-Schmidt step), other. */
#define LGPU_N_KINDS 14
int lgpu_set_profiling(lgpu_ctx* ctx, int32_t enable);
int lgpu_profile_read(lgpu_ctx* ctx, double* ms, int64_t* counts, double* algo_bytes,
                      int32_t nkinds, int32_t reset);
/* Device time (ms, CUDA events on the context's stream) of the most recent
 * assemble / factorize / arnoldi-loop / extraction phases. */
int lgpu_phase_times(lgpu_ctx* ctx, double* t_assemble_ms, double* t_factor_ms,
                     double* t_iter_ms, double* t_extract_ms);

#ifdef __cplusplus
}
#endif
#endif /* LEGOLAS_B200_H */
