/* ============================================================================
 * cnld_b200.h -- C ABI of libcnld_b200.so
 *
 * B200-native (sm_100a) implementation of cnl-dyna's frequency-response hot
 * path: assemble Z(k) (Galerkin single-layer Helmholtz, linear hats) -> form
 * G = K - w^2 M - j w B - 2 rho w^2 Z -> unpivoted LU -> solve all source
 * patches -> conj / boundary mask / patch average -> field pressure.
 *
 * Two groups of entry points:
 *   (A) cnld_b200_*  : batched, device-resident entry points for a frequency
 *       shard (not in the reference; this is what the sweep driver calls).
 *   (B) H2Lib-named  : the symbols cnl-dyna's Cython layer binds
 *       (cnld/h2lib/_*.pxd), same prototypes, host-visible struct prefixes
 *       identical to the ones Cython dereferences.  Declared in
 *       cnld_b200_h2lib.h.
 *
 * Conventions (both groups): real = double, field = C99 double _Complex
 * (two doubles re,im), uint = unsigned int (cnld/h2lib/_basic.pxd:3-6);
 * dense matrices column-major with leading dimension ld (include/amatrix.h:42-48);
 * all pointers are HOST pointers owned by the caller; the library owns device
 * memory until *_destroy.  Functions return 0 on success, non-zero on error
 * (message through cnld_b200_last_error()); no exceptions cross the ABI.
 * There is NO CPU fallback: without a CUDA device every compute entry fails.
 * ========================================================================== */
#ifndef CNLD_B200_H
#define CNLD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__cplusplus)
typedef struct { double re, im; } cnld_field; /* layout-identical to double _Complex */
#else
typedef double _Complex cnld_field;
#endif
typedef unsigned int cnld_uint;

typedef struct cnld_ctx cnld_ctx; /* opaque: mesh, quadrature tables, FEM data, device workspaces */

/* ---- library ---------------------------------------------------------- */
const char *cnld_b200_last_error(void);
const char *cnld_b200_version(void);
int cnld_b200_device_count(void);
/* number of kernel launches issued by this library since load (bench: gpu_launches) */
unsigned long long cnld_b200_launch_count(void);
/* Real FP64 flops executed by every ZGEMM launched so far in this process (host-side count:
 * 8 m n k for the 4M kernel, 6 m n k for the 3M kernel, tiles skipped by lower-only updates excluded). */
unsigned long long cnld_b200_dmma_flop_count(void);

/* ---- context: mesh + quadrature (frequency independent) ---------------
 * Replaces: Mesh.from_abstract -> Surface3d + prepare_surface3d
 *           (cnld/mesh.py:182-222,148; include/surface3d.h:45-71,101) and
 *           new_slp_helmholtz_bem3d(k, gr, q_reg, q_sing, LINEAR, LINEAR) ->
 *           build_singquad2d (include/helmholtzbem3d.h:81-83, singquad2d.h:179;
 *           call sites cnld/compressed_formats.py:629,675).
 * x[nv][3] vertex coordinates, t[nt][3] triangle vertex ids. */
cnld_ctx *cnld_b200_create(int device, cnld_uint nv, cnld_uint nt, const double *x,
                           const cnld_uint *t, cnld_uint q_reg, cnld_uint q_sing);
void cnld_b200_destroy(cnld_ctx *ctx);
int cnld_b200_set_streams(cnld_ctx *ctx, int nstreams); /* frequencies in flight, default 2 */

/* Gram determinants g[nt] (= 2*area) and normals n[nt][3] as computed on device
 * (prepare_surface3d, include/surface3d.h:101). Either pointer may be NULL. */
int cnld_b200_get_geometry(cnld_ctx *ctx, double *g, double *n);

/* ---- stage entry points (each mirrors one reference call) ----------------- */

/* assemble_bem3d_amatrix(bem, Z) (include/bem3d.h:3028; compressed_formats.py:634):
 * dense nv x nv Galerkin matrix of the Helmholtz SLP for wavenumber k = k_re + i k_im,
 * kernel exp(i k r)/(4 pi r).  Z is a host buffer, column-major, ld >= nv. */
int cnld_b200_assemble_slp_ll(cnld_ctx *ctx, double k_re, double k_im, cnld_field *Z,
                              cnld_uint ld);

/* bem->nearfield(ridx, cidx, bem, false, N) (include/bem3d.h:370-385): the
 * rows x cols sub-block of Z for the given vertex index sets. */
int cnld_b200_nearfield_slp_ll(cnld_ctx *ctx, double k_re, double k_im, const cnld_uint *ridx,
                               cnld_uint rows, const cnld_uint *cidx, cnld_uint cols,
                               cnld_field *N, cnld_uint ld);

/* lrdecomp_amatrix(a) (include/factorizations.h:173-182; compressed_formats.py:249):
 * in-place UNPIVOTED LU, L unit lower.  *info = 0 ok, k+1 if pivot k is zero/non-finite. */
int cnld_b200_lrdecomp(int device, cnld_uint n, cnld_field *a, cnld_uint ld, cnld_uint *info);
/* Same elimination for a complex SYMMETRIC matrix of which only the lower triangle is read
 * (a = L R with R = D L^T, half the flops); on exit a holds L and R exactly like lrdecomp's. */
int cnld_b200_lrdecomp_sym(int device, cnld_uint n, cnld_field *a, cnld_uint ld, cnld_uint *info);

/* lrsolve_amatrix_avector(false, LU, x) (include/factorizations.h:233;
 * compressed_formats.py:259-262) for nrhs right-hand sides at once. */
int cnld_b200_lrsolve(int device, cnld_uint n, const cnld_field *lu, cnld_uint ld,
                      cnld_field *x, cnld_uint nrhs, cnld_uint ldx);

/* addmul_amatrix(alpha, false, A, false, B, C) (cnld/h2lib/_amatrix.pxd:34) restricted to
 * alpha = +-1 and C = beta*C + alpha*A*B with beta in {0,1}: the FP64 tensor-core ZGEMM the
 * LU is built from, exposed so it can be parity-tested on its own. */
int cnld_b200_zgemm(int device, cnld_uint m, cnld_uint n, cnld_uint k, double alpha,
                    const cnld_field *A, cnld_uint lda, const cnld_field *B, cnld_uint ldb,
                    double beta, cnld_field *C, cnld_uint ldc);

/* ---- frequency sweep (the metric's unit of work) -----------------------
 * FEM operators of the whole array as ONE CSR pattern with three value arrays
 * (fem.array_mbk_spmatrix, cnld/fem.py:1289-1315: block = -(w^2) M - 1j w B + K). */
int cnld_b200_set_mbk(cnld_ctx *ctx, const int64_t *indptr, const cnld_uint *indices,
                      const double *K, const double *M, const double *B);
/* Patch loads F[nv x P] and patch averaging AVG[nv x P] in CSC
 * (fem.array_f_spmatrix / array_avg_spmatrix, cnld/fem.py:1255-1286) and the
 * boundary-node mask ob[nv] (mesh.on_boundary, cnld/mesh.py:262-266). */
int cnld_b200_set_loads(cnld_ctx *ctx, cnld_uint npatch, const int64_t *f_indptr,
                        const cnld_uint *f_idx, const double *f_val, const int64_t *a_indptr,
                        const cnld_uint *a_idx, const double *a_val, const uint8_t *ob);

/* generate_db.process for nf frequencies (cnld/scripts/generate_db.py:30-130):
 *   k = 2 pi f / c;  G = MBK(f) + (-(2 pi f)^2 * 2 rho) Z(k);  LU;  for every
 *   source patch s: x = conj(LU \ F[:,s]); x[ob] = 0; x_patch = AVG^T x.
 * x_patch[nf][P(src)][P(dst)], x_node[nf][P(src)][nv] (nullable), t_freq[nf]
 * device seconds per frequency (nullable). */
int cnld_b200_sweep_run(cnld_ctx *ctx, const double *freqs, cnld_uint nf, double c, double rho,
                        cnld_field *x_patch, cnld_field *x_node, double *t_freq);

/* Same work, inputs/outputs stay in HBM (device-resident timing for the bench's
 * `value`); results are kept in an internal device buffer and can be fetched with
 * cnld_b200_sweep_fetch afterwards. */
int cnld_b200_sweep_run_device(cnld_ctx *ctx, const double *freqs, cnld_uint nf, double c,
                               double rho);
int cnld_b200_sweep_fetch(cnld_ctx *ctx, cnld_uint nf, cnld_field *x_patch, cnld_field *x_node);

/* per-stage device time (seconds, CUDA events) accumulated over the last sweep_run*:
 * out[0]=assemble, [1]=factor, [2]=solve+epilogue (each summed over frequencies; they overlap
 * when several streams are in flight), [3]=kernel launches, [4]=device seconds of the whole
 * sweep (begin/end events joined across all streams).  out must hold 5 doubles. */
int cnld_b200_sweep_stage_times(cnld_ctx *ctx, double *out);

/* Page-lock a caller-owned host region in place (cudaHostRegister).  A later sweep_run whose x_node
 * buffer lies inside it copies node displacements straight into it (no staging buffer, no host
 * memcpy).  One region per context; ptr = NULL (or destroy) releases it.  The caller must keep the
 * memory alive until then. */
int cnld_b200_pin_host(cnld_ctx *ctx, void *ptr, size_t bytes);

/* Symmetric path (mode 1; default 0 = the reference's general elimination, solve_cholesky=False /
 * lrdecomp_amatrix in cnld/h2lib/factorizations_cy.pyx).  G = S + Delta, S = the complex symmetric
 * matrix given by G's lower triangle, Delta = the sparse strictly-upper asymmetry (the Sauter-Schwab
 * rules are not symmetric under swapping the two panels; plus any asymmetry of K, M, B).  S is
 * eliminated as S = L (D L^T) at half the flops, and x = G^-1 f is recovered by `refine_steps`
 * steps of e <- -S^-1 Delta e, x <- x + e.  |e_last| / |x| is measured on the device for every
 * frequency; a frequency whose estimated remaining relative error exceeds the tolerance
 * (default 1e-11) is redone with mode 0, so both modes return the same results to that tolerance.
 * symmetric_stats returns the mode and, for the last sweep, the number of frequencies redone and the
 * largest |e_last| / |x|; mbk_asym_nnz = entries of K, M, B that differ from their transposes. */
int cnld_b200_set_symmetric(cnld_ctx *ctx, int mode, int refine_steps);
int cnld_b200_set_symmetric_tol(cnld_ctx *ctx, double tol);

/* Translation classes (symmetric path only).  A CMUT array mesh is one membrane mesh translated to
 * every membrane position (cnld/mesh.py Mesh.from_abstract), and the kernel depends on x - y only:
 * the block of Z coupling membranes a and b depends on their offset alone.  When the mesh is
 * detected to be translated copies of its first connected component (identical connectivity,
 * vertices equal up to a translation to 32 ulp), one block per distinct offset is integrated and the
 * others are copied (opposite offsets: transposed).  Values differ from integrating every block only
 * through the rounding of the translated coordinates.  on = 1 by default; translation_info returns
 * whether the classes are in use, the number of components and of integrated off-diagonal blocks. */
int cnld_b200_set_translation_classes(cnld_ctx *ctx, int on);
int cnld_b200_translation_info(const cnld_ctx *ctx, unsigned *ncomp, unsigned *nclass);
/* Host-only view of the plan (tests): counts = {periodic, components, offset classes, work items,
 * copy items, vertices per component}; work items are {t_begin, t_end, s_begin, s_end} (row triangles
 * x column-triangle chunk, pairs s < t are integrated), copy items {dst_r0, dst_c0, src_r0, src_c0,
 * transposed} (vertex offsets of blocks of `vertices per component` rows and columns). */
int cnld_b200_translation_plan(cnld_uint nv, cnld_uint nt, const double *x, const cnld_uint *t, int allow,
                               cnld_uint *counts, cnld_uint *work, cnld_uint work_cap, cnld_uint *copy,
                               cnld_uint copy_cap);
int cnld_b200_symmetric_stats(const cnld_ctx *ctx, unsigned *mbk_asym_nnz, unsigned *redone, double *max_ratio);

/* Dominant-kernel timing for the roofline: when profiling is on, every trailing-update
 * ZGEMM launch of the LU is bracketed by CUDA events on its own stream.
 * out[0] = executed real flops (6 m n k per launch: 3M complex product), out[1] = seconds (sum of launch durations),
 * out[2] = number of launches, all over the last sweep_run*. */
int cnld_b200_set_profile(cnld_ctx *ctx, int on);
int cnld_b200_sweep_kernel_stats(cnld_ctx *ctx, double *out);

/* ---- field pressure ------------------------------------------------------
 * bem_pressure.mesh_pres_vector (cnld/bem_pressure.pyx:54-88) for nobs points and
 * one wavenumber: pvec[nobs][nv].  Kernel exp(-j k r)/(4 pi r), 3-point rule,
 * source z forced to 0, scale -(k c)^2 * 2 rho. */
int cnld_b200_pres_vector(cnld_ctx *ctx, const double *obs, cnld_uint nobs, double k, double c,
                          double rho, cnld_field *pvec);
/* Fused p[ik][isrc][iobs] = pvec(obs, ks[ik]) . disp[ik][isrc][:]  -- the loop of
 * array_patch_pres_imp_resp (cnld/bem_pressure.pyx:103-112) for all points at once. */
int cnld_b200_pressure(cnld_ctx *ctx, const double *obs, cnld_uint nobs, const double *ks,
                       cnld_uint nk, double c, double rho, const cnld_field *disp,
                       cnld_uint nsrc, cnld_field *p);

/* Many sources (nsrc = P source patches) at 10^5..10^6 observation points: the loop of
 * array_patch_pres_imp_resp (cnld/bem_pressure.pyx:103-112) for all points at once.  The observation
 * points stay on the device; per wavenumber exp(-j k r)/r is evaluated once per (point, source point),
 * folded to node space on chip and contracted with all displacement columns by one FP64 tensor-core
 * ZGEMM per chunk of points.  cnld_b200_pressure takes this path by itself when nsrc > 4. */
typedef struct cnld_pfield cnld_pfield;
cnld_pfield *cnld_b200_pfield_create(cnld_ctx *ctx, const double *obs, cnld_uint nobs, cnld_uint nsrc_max);
void cnld_b200_pfield_destroy(cnld_pfield *pf);
/* page-lock the caller's result buffer in place (D2H of a chunk then overlaps the next chunk) */
int cnld_b200_pfield_pin(cnld_pfield *pf, void *ptr, size_t bytes);
/* p[isrc][iobs] (HOST, nullable: results stay on the device) for one wavenumber, disp[nsrc][nv] HOST */
int cnld_b200_pfield_run(cnld_pfield *pf, double k, double c, double rho, const cnld_field *disp,
                         cnld_uint nsrc, cnld_field *p);
/* same, with the displacements of the previous run still on the device (device-resident timing) */
int cnld_b200_pfield_run_device(cnld_pfield *pf, double k, double c, double rho, cnld_uint nsrc);
/* results of the last run: all points (idx = NULL, p[nsrc][nobs]) or the listed ones (p[nsrc][nidx]) */
int cnld_b200_pfield_fetch(cnld_pfield *pf, cnld_uint nsrc, const cnld_uint *idx, cnld_uint nidx,
                           cnld_field *p);
/* out[7]: seconds of the last run in the pressure-vector kernel, in the contraction, in total (CUDA
 * events on the launch stream), chunks, observation points per chunk, triangle chunks of the mesh,
 * node-space writes per observation point */
int cnld_b200_pfield_info(cnld_pfield *pf, double *out);
/* Host-only view of the triangle chunks the pressure-vector kernel walks (tests): counts[6] = {chunks,
 * chunk-local node entries, max triangles per chunk, max nodes per chunk, entries that accumulate onto
 * an earlier chunk's write, vertices without a triangle}; order[nt] (nullable) = triangles in chunk order. */
int cnld_b200_pres_plan_stats(cnld_uint nv, cnld_uint nt, const double *x, const cnld_uint *t,
                              cnld_uint *counts, cnld_uint *order);

/* ---- H-matrix path (arrays too large to factor densely) ---------------------------------
 * Replaces build_bem3d_cluster(bem, clf, LINEAR) (include/bem3d.h:1493), build_strict_block /
 * build_nonstrict_block(root, root, &eta, admissible_*) (include/block.h:216-248, :78-140),
 * build_from_block_hmatrix (include/hmatrix.h:358), setup_hmatrix_aprx_paca_bem3d +
 * assemble_bem3d_hmatrix (include/bem3d.h:2627, :3056), addeval_hmatrix_avector
 * (include/hmatrix.h:419) and the Krylov solve (include/krylovsolvers.h:283;
 * scratch/freq_resp_db_krylov.py:38-79).  Call sites: cnld/compressed_formats.py:652-703. */
typedef struct cnld_hmat cnld_hmat;
/* admis: 0 = '2' (max diam <= eta dist, euclidean), 1 = 'max', 2 = 'sphere', 3 = '2min';
 * kmax: rank capacity of a low-rank leaf (0 -> 64). */
cnld_hmat *cnld_b200_hmat_create(cnld_ctx *ctx, cnld_uint clf, double eta, int admis, int strict,
                                 cnld_uint kmax);
void cnld_b200_hmat_destroy(cnld_hmat *h);
/* out[8]: clusters, leaves, admissible leaves, inadmissible leaves, pool capacity (entries),
 * stored entries (dense + k(m+n), after assembly), max rank, n */
int cnld_b200_hmat_info(cnld_hmat *h, double *out);
/* Translation classes of the leaves (mesh = translated copies of one membrane, see
 * cnld_b200_set_translation_classes): a leaf whose row and column vertices are the translates, by one
 * common lattice vector and in the same cluster order, of an earlier leaf's shares that leaf's data and
 * is not assembled.  out[4]: admissible leaves assembled, inadmissible leaves assembled, leaves sharing
 * another leaf's data, pool capacity actually allocated (entries). */
int cnld_b200_hmat_sharing(cnld_hmat *h, double *out);
/* tree export (any pointer may be NULL): perm[n] (cluster order -> DOF), per cluster son0/son1
 * (-1 = leaf), start/size into perm, bmin/bmax[.][3]; per leaf row/col cluster, admissibility, rank */
int cnld_b200_hmat_tree(cnld_hmat *h, cnld_uint *perm, int *son0, int *son1, cnld_uint *start,
                        cnld_uint *size, double *bmin, double *bmax, cnld_uint *leaf_rc,
                        cnld_uint *leaf_cc, uint8_t *leaf_adm, cnld_uint *leaf_rank);
/* fill every leaf for wavenumber k: dense near field, partial-pivot ACA (accuracy eps_aca) far field */
int cnld_b200_hmat_assemble(cnld_hmat *h, double k_re, double k_im, double eps_aca);
/* y[r][:] += alpha * Z x[r][:] for nrhs host vectors of length n (original DOF order) */
int cnld_b200_hmat_matvec(cnld_hmat *h, cnld_field alpha, const cnld_field *x, cnld_field *y,
                          cnld_uint nrhs);
/* Coarse space of the two-level GMRES preconditioner (optional; without it: block Jacobi).  W[n][m]
 * (node major, m <= 16): the first m in-vacuo eigenmodes of every membrane on its own nodes (zero on
 * clamped nodes); fmode[m]: their resonance frequencies (Hz), used to pick how many are needed at a
 * frequency.  The preconditioner is then  M^-1 = Q + P (I - G Q),  Q = W (W^H G W)^-1 W^H  with G W kept
 * on the device (one extra dense product per iteration, no extra H-matvec): the inter-membrane crosstalk
 * through the low modes is solved exactly, and the iteration count stops growing with the array size and
 * near the immersed resonances.  cnld_b200_hmat_coarse_used: functions per membrane in the last sweep. */
int cnld_b200_hmat_set_coarse(cnld_hmat *h, cnld_uint m, const double *W, const double *fmode);
int cnld_b200_hmat_coarse_used(cnld_hmat *h);
/* Low-rank leaves whose ACA stopped at the rank capacity kmax before reaching eps_aca (H2Lib's ACA keeps
 * growing the rank): count of the last assemble call, and summed over the last hsweep_run.  Non-zero
 * means the H-matrix is less accurate than eps_aca asks for -- raise kmax. */
int cnld_b200_hmat_capped(cnld_hmat *h, unsigned *last_assembly, unsigned *last_sweep);
int cnld_b200_hmat_todense(cnld_hmat *h, cnld_field *Z, cnld_uint ld);
/* device-resident timing of the matvec with nrhs vectors: seconds per call (CUDA events) */
int cnld_b200_hmat_bench_matvec(cnld_hmat *h, cnld_uint nrhs, int reps, double *seconds);
/* generate_db.process with the H-matrix + GMRES path: per frequency assemble Z_H(k), then solve
 * (MBK + (-2 rho w^2) Z_H) x = F[:, s] for all source patches by left-preconditioned restarted
 * GMRES (block-diagonal MBK^-1), `batch` right-hand sides in lock-step; residual tolerance tol.
 * Outputs as cnld_b200_sweep_run; iters[nf][P] (nullable) = Arnoldi steps per right-hand side. */
int cnld_b200_hsweep_run(cnld_ctx *ctx, cnld_hmat *h, const double *freqs, cnld_uint nf, double c,
                         double rho, double eps_aca, double tol, cnld_uint restart,
                         cnld_uint maxiter, cnld_uint batch, cnld_field *x_patch,
                         cnld_field *x_node, cnld_uint *iters, double *t_freq);

/* ---- impulse responses -------------------------------------------------
 * impulse_response.fft_to_fir / fft_to_sir (cnld/impulse_response.py:191-224) for nbatch one-sided spectra
 * s[nbatch][nf] on the frequency grid freqs[nf]: magnitude / unwrapped-phase interpolation by `interp`
 * (:161-175), conjugate mirror to nfft = 2 ((nf - 1) interp + 1) - 1 bins (:11-46), optional Kramers-Kronig
 * reconstruction (:115-125), real(ifft) * fs.  Outputs t[nfft] (nullable) and h[nbatch][nfft] (HOST).
 * Evaluated as one exactly-twiddled DFT GEMM on the FP64 tensor cores (csrc/impresp.cu). */
int cnld_b200_fft_to_fir(int device, cnld_uint nf, const double *freqs, cnld_uint nbatch, const cnld_field *s,
                         cnld_uint interp, int use_kkr, double *t, double *h);

/* ---- time-domain FIR convolution ----------------------------------------------
 * fir_conv_cy (cnld/simulation.pyx:51-77): x[j] = dt * sum_i sum_k fir[i][j][k + offset] p[nsample-1-k][i],
 * the convolution of FixedStepSolver (:400-410, :459-503).  The table fir[nsrc][ndest][nfir] stays on the
 * device; every convolution streams it once (HBM-bound). */
typedef struct cnld_fir cnld_fir;
cnld_fir *cnld_b200_fir_create(int device, cnld_uint nsrc, cnld_uint ndest, cnld_uint nfir, const double *fir);
void cnld_b200_fir_destroy(cnld_fir *f);
/* fir_conv_cy(fir, p, dt, offset): p[nsample][nsrc], x[ndest] (HOST) */
int cnld_b200_fir_conv(cnld_fir *f, const double *p, cnld_uint nsample, double dt, int offset, double *x);
/* resident stepping: push the finished step's pressures p_row[nsrc] into the device history; base = the
 * history convolved with offset 1 (`_fir_conv_base`); add = the current sample with offset 0 (`_fir_conv_add`) */
int cnld_b200_fir_reset(cnld_fir *f);
int cnld_b200_fir_push(cnld_fir *f, const double *p_row);
int cnld_b200_fir_base(cnld_fir *f, double dt, double *x);
int cnld_b200_fir_add(cnld_fir *f, const double *p_row, double dt, double *x);

/* ---- fixed-step time-domain solver (simulation.pyx:459-503) ---------------------
 * FixedStepSolver (cnld/simulation.pyx:297-555) with the table, the total-pressure history and all state arrays
 * resident on the device: a step is one pass over the table (the blind estimate = `_fir_conv_base`, :447-457 /
 * :400-404) followed by the fixed-point iterations on the current sample (:430-445, :466-498) in one CTA; no host
 * round trip between steps.
 *   fir[npatch][npatch][nfir]  table already resampled to the step dt (the constructor's interp1d, :311-314)
 *   v[nt][npatch]              drive on the solver's time grid (StateDB, :108-126); gap_eff[npatch]
 *   e0 = vacuum permittivity (scipy.constants.epsilon_0 of the caller), electrostatic pressure p_es_pp (:17-21)
 *   k, n, x0 contact spring (:80-88), lmbd contact damper (:91-99), atol / maxiter of the fixed point (:303)
 * create() sets the initial state (:335, current_step = 1).  The contact pressures reproduce np.vectorize's
 * output-type rule (truncation toward zero whenever patch 0 is out of contact), see csrc/firconv.cu. */
typedef struct cnld_td cnld_td;
cnld_td *cnld_b200_td_create(int device, cnld_uint npatch, cnld_uint nfir, const double *fir, cnld_uint nt, const double *v,
                             const double *gap_eff, double dt, double e0, double k, double n, double x0, double lmbd,
                             double atol, int maxiter);
void cnld_b200_td_destroy(cnld_td *t);
/* FixedStepSolver.reset (:516-527) */
int cnld_b200_td_reset(cnld_td *t);
/* nsteps x FixedStepSolver.step (:459-503); error if the time grid ends first.  Samples advance in blocks of up to 16
 * per pass over the table (the history before the block is convolved once for the whole block, the lags inside it per
 * sample); nsteps = 1 is the reference's scheme, one pass per sample.  Environment at create: CNLD_B200_TD_BLOCK=n
 * (1..16) block size, CNLD_B200_TD_MMA=0 history pass without the interleaved second copy of the table (blocks of 8),
 * CNLD_B200_TD_PDL=0 no programmatic dependent launches, CNLD_B200_TD_CLUSTER=0 step kernel on one CTA,
 * CNLD_B200_TD_OVERLAP=0 history passes in line instead of running ahead on a second stream,
 * CNLD_B200_TD_TRACE=1 phase marks for cnld_b200_td_trace. */
int cnld_b200_td_step(cnld_td *t, cnld_uint nsteps);
cnld_uint cnld_b200_td_current_step(const cnld_td *t);
/* taps a step's convolution reads: 1 + the last tap that is nonzero in any patch pair (a causal, Kramers-Kronig
 * table is exactly zero in its second half; those taps are never streamed) */
cnld_uint cnld_b200_td_taps(const cnld_td *t);
/* samples [i0, i1) of x, u, p_es, p_cont_spr, p_cont_dmp, p_tot -> HOST [i1 - i0][npatch] each (nullable) */
int cnld_b200_td_state(cnld_td *t, cnld_uint i0, cnld_uint i1, double *x, double *u, double *p_es, double *p_cont_spr,
                       double *p_cont_dmp, double *p_tot);
/* diagnostic, needs CNLD_B200_TD_TRACE=1 in the environment at create: %globaltimer marks [ns] of the step kernels of
 * samples [s0, s1), marks[s1 - s0][8]: kernel start, prologue done, dependency met, base ready, first fixed point done,
 * second fixed point done, state rows written, (unused) */
int cnld_b200_td_trace(cnld_td *t, cnld_uint s0, cnld_uint s1, unsigned long long *marks);
/* FixedStepSolver.error / .iters (:383-389) of the steps that produced samples [s0, s1), s0 >= 1 (HOST, nullable) */
int cnld_b200_td_info(cnld_td *t, cnld_uint s0, cnld_uint s1, double *error, int *iters);

/* ---- multi-GPU plumbing (one process per GPU) ------------------------------------------------------------
 * Replaces the reference's process pool + write lock (cnld/scripts/generate_db.py:229-242): frequencies are
 * sharded over ranks, nothing is exchanged inside a frequency; rank 0's frequency-independent setup is
 * broadcast once and the patch responses are gathered once, with NCCL over NVLink (libnccl.so.2 loaded at
 * run time, no PyTorch).  rank / world / addr / port come from the launcher (torchrun exports RANK,
 * WORLD_SIZE, LOCAL_RANK, MASTER_ADDR, MASTER_PORT; the port must be free on rank 0's host -- it carries the
 * NCCL unique id, one TCP connection per rank).  world = 1 needs neither NCCL nor a GPU: every call is a no-op.
 * All buffers are HOST buffers. */
typedef struct cnld_comm cnld_comm;
cnld_comm *cnld_b200_comm_create(int rank, int world, int device, const char *addr, int port);
void cnld_b200_comm_destroy(cnld_comm *comm);
int cnld_b200_comm_rank(const cnld_comm *comm);
int cnld_b200_comm_world(const cnld_comm *comm);
int cnld_b200_comm_bcast(cnld_comm *comm, void *buf, size_t bytes, int root);
/* every rank contributes `bytes` bytes; root receives world * bytes in rank order (recv may be NULL elsewhere) */
int cnld_b200_comm_gather(cnld_comm *comm, const void *send, size_t bytes, void *recv, int root);
/* element-wise maximum over the ranks, in place (timing: max over ranks); n = 0 is a barrier */
int cnld_b200_comm_allreduce_max(cnld_comm *comm, double *vals, int n);
int cnld_b200_comm_barrier(cnld_comm *comm);
/* out[3]: payload bytes this rank moved through NCCL, device seconds of those collectives, NCCL version */
int cnld_b200_comm_stats(cnld_comm *comm, double *out);

/* ---- database writer ---------------------------------------------------
 * Bulk appends to cnl-dyna's sqlite database (cnld/database.py:22-92 schema, :144-194 append_*;
 * cnld/scripts/generate_db.py:101-130 writes the same rows one source patch at a time through Python):
 * prepared statements on libsqlite3, one transaction per block of frequencies, straight from the arrays
 * cnld_b200_sweep_run returned.  x_patch[nf][P(src)][P(dst)], x_node[nf][P(src)][N] (nullable),
 * nodes[N][3], times[nf] (nullable), job_ids[nf] (nullable: progress rows marked complete, 1-based). */
int cnld_b200_db_append_block(const char *file, cnld_uint nf, const double *freqs, const double *wavenums,
                              cnld_uint P, const cnld_field *x_patch, const double *times, cnld_uint N,
                              const cnld_field *x_node, const double *nodes, const cnld_uint *job_ids, int fast);
/* drop the two frequency-response indexes before a bulk load / (re)build them after it.
 * With fast > 1, cnld_b200_db_append_block writes the node rows (cnld/database.py:38-45, 2.13 M per
 * frequency at the 4x4 array) and the patch rows (database.py:22-29) as b-tree pages in SQLite's file format instead of through SQL
 * (csrc/dbpages.cpp: records encoded by `fast` threads, pages appended to the file, interior levels
 * rebuilt; applies when the table has no index, i.e. inside the bulk bracket, and the file is a plain
 * rollback-journal database); cnld_b200_db_finish then generates the node index, already sorted, from the
 * descriptors of those appends when they cover the whole table, and lets SQLite build it otherwise.
 * Environment CNLD_B200_DB_MODE=sql disables the page path. */
int cnld_b200_db_begin_bulk(const char *file);
int cnld_b200_db_finish(const char *file);
/* Rows (source_patch, dest_patch, time, displacement) of patch_to_patch_imp_resp (cnld/database.py:47-53) from
 * h[P][P][nt] and times[nt] -- what cnld/scripts/generate_db.py:143-164 appends through pandas after the sweep.
 * Into an empty table of a plain rollback-journal database the rows and their (source_patch, dest_patch, time)
 * index are written as pages by `threads` threads; otherwise prepared statements in one transaction. */
int cnld_b200_db_append_imp_resp(const char *file, cnld_uint P, cnld_uint nt, const double *times, const double *h,
                                 int threads);
/* out[8]: node rows written as pages, node rows written through SQL, node indexes generated from
 * descriptors, node indexes built by CREATE INDEX, impulse-response rows written as pages / through SQL,
 * patch rows written as pages / through SQL (this process) */
int cnld_b200_db_counters(unsigned long long *out);

/* ---- microbenchmarks used for roofline denominators ----------------------- */
/* FP64 DMMA ZGEMM C(m x n) -= A(m x k) B(k x n) on device-resident random data;
 * returns achieved TFLOP/s (8 m n k flop) averaged over reps. */
int cnld_b200_bench_zgemm(int device, cnld_uint m, cnld_uint n, cnld_uint k, int reps,
                          double *tflops);
/* FP64 FMA-pipe peak probe (dependent-free DFMA chains), TFLOP/s */
int cnld_b200_bench_dfma(int device, int reps, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* CNLD_B200_H */
