/* svirl_b200 -- C ABI of the B200-native TDGL / CG hot path.
 *
 * This shared library replaces the pyCUDA layer of microsoft/svirl
 * (svirl/cuda/*.h kernels, svirl/parallel/{startup,reduction,utils}.py,
 * svirl/storage/arrays.py device side).  Plain pointers and sizes only: the
 * caller owns host (numpy) buffers, the library owns device memory.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure;
 *     svl_last_error() returns the message of the last failure on this thread.
 *   - not thread-safe per context; one context per process/GPU is the intended use.
 *   - host arrays are in the reference's flat device layout (x fastest):
 *       nodes  n = i + Nx*j              (svirl/cuda/common.h:75-81)
 *       edges  a: i + (Nx-1)*j ; b: Na + i + Nx*j, Na=(Nx-1)*Ny  (svirl/cuda/td.h:91-100)
 *       cells  i + (Nx-1)*j
 *     device storage is pitched planes; conversion happens in h2d/d2h.
 *   - `real` scalars are passed as double and rounded to the context dtype inside.
 *   - a NULL svl_buf* means "absent" exactly where the reference passes np.uintp(0).
 */
#ifndef SVIRL_B200_H
#define SVIRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svl_ctx svl_ctx;
typedef struct svl_buf svl_buf;

/* buffer kinds */
enum {
    SVL_NODE_C = 0, /* complex on nodes, Nx*Ny            (psi, jacobian_psi, directions)   */
    SVL_NODE_R = 1, /* real on nodes                      (spatial linear coefficient)      */
    SVL_EDGE   = 2, /* real on edges, a then b, Na+Nb     (vector potential and friends)    */
    SVL_CELL_R = 3, /* real on cells (Nx-1)*(Ny-1)        (magnetic field)                  */
    SVL_CELL_B = 4, /* 1 byte on cells                    (material tiling)                 */
    SVL_FLAT   = 5  /* contiguous array of n elements of elem_size bytes                    */
};

const char *svl_last_error(void);
int svl_version(void);

/* ---- lifecycle: replaces Startup.__init__/__del__ (svirl/parallel/startup.py:17-93).
 * dtype_bytes: 4 (float/complex64) or 8 (double/complex128).
 * dx, dy: as double(str(np.floatXX(dx))) -- the reference embeds the decimal repr and folds
 * 1/(dx*dx) in double (startup.py:54-60, td.h:34-35).
 * Slab decomposition (multi-GPU, new): this context owns node rows [j0, j1) of the global
 * Ny rows; single GPU: j0 = 0, j1 = Ny. */
int svl_create(svl_ctx **out, int device_id, int Nx, int Ny, double dx, double dy, int dtype_bytes,
               int j0, int j1);
int svl_destroy(svl_ctx *ctx);
int svl_synchronize(svl_ctx *ctx);
/* knobs: "psi_kernel" (0 = plain per-node, 1 = temporally blocked streaming, 2 = register-resident
 * tile kernel, default), "psi_k" (psi sweeps fused per launch), "tma" (0/1), "graphs" (0/1),
 * "a_kernel" (0 = per-node A sweep, 1..4 = tile kernel fusing sweep pairs in four layouts, default 2),
 * "cg_fused" (2 = two-pass CG iteration svl_cg_pass_a/_b, default where it applies; 1 = three-pass iteration
 * svl_cg_begin/_end; 0 = composition of the single kernels), "spin_timeout_ms" (bound of the spin waits on peer
 * GPUs, 0 = wait forever, default; env SVL_SPIN_TIMEOUT_MS); round 2b: "psi_links" (fp32 tile kernel: 1 = MUFU link
 * variables, default; 0 = polynomial sincos), "psi_patch" (1 = 2x4 node patch per thread, default; 0 = 1x8 column),
 * "psi_shape", "pipeline" (kappa = inf svl_td_run: next step's first launch pre-issued behind a device-side stop rule),
 * "pdl" (programmatic dependent launches: 0 off, 1 one GPU, 2 also slab batches, default), "slab_split", "slab_bnd"
 * (CTAs of the boundary launch of a slab batch: 0 = svl_slab_split_plan); the full table is in INTEGRATION.md section 6 */
int svl_set_option(svl_ctx *ctx, const char *name, int value);
/* host arithmetic behind option "slab_bnd" = 0: CTAs given to the nb boundary tiles of a slab batch when ni interior
 * tiles share `slots` resident CTAs (new; the reference is single-GPU) */
int svl_slab_split_plan(int nb, int ni, int slots);
int svl_get_stat(svl_ctx *ctx, const char *name, double *value);
/* self-test hook: the library's own sincos (used for every link variable exp(-i d A)) on n values */
int svl_debug_sincos(svl_ctx *ctx, size_t n, const double *x, double *s_out, double *c_out);
/* diagnostics for slab runs (option "trace" = N arms it): per psi-tile launch {first CTA start, last CTA
 * end, longest halo-flag wait, end of the last wait} in %globaltimer nanoseconds */
int svl_debug_trace(svl_ctx *ctx, unsigned long long *out, int max_records, int *n_out);
/* CUDA events on the context's launch stream (8 slots), for device-side timing */
int svl_event_record(svl_ctx *ctx, int slot);
int svl_event_elapsed_ms(svl_ctx *ctx, int slot0, int slot1, double *ms);

/* ---- buffers: replaces pycuda.gpuarray + cuda.memcpy_* (svirl/storage/arrays.py:520-587,
 * svirl/parallel/utils.py:21-27).  n is only used for SVL_FLAT. */
int svl_alloc(svl_ctx *ctx, int kind, size_t n, int elem_size, svl_buf **out);
int svl_free(svl_ctx *ctx, svl_buf *buf);
int svl_h2d(svl_ctx *ctx, svl_buf *dst, const void *src);   /* full array, reference flat layout */
int svl_d2h(svl_ctx *ctx, void *dst, const svl_buf *src);
int svl_d2d(svl_ctx *ctx, svl_buf *dst, const svl_buf *src);
int svl_fill_zero(svl_ctx *ctx, svl_buf *buf);
int svl_swap(svl_ctx *ctx, svl_buf *x, svl_buf *y);         /* swap storage of two same-kind buffers */
size_t svl_buf_size(const svl_buf *buf);                    /* elements in the flat layout */
/* row-slab transfers for fields too large for one host array (rows [r0, r1) of the plane;
 * part = 0 for nodes/cells/a, 1 for b) */
int svl_h2d_rows(svl_ctx *ctx, svl_buf *dst, int part, int r0, int r1, const void *src);
int svl_d2h_rows(svl_ctx *ctx, void *dst, const svl_buf *src, int part, int r0, int r1);

/* Host helper for seeded fields at scale: numpy's legacy Mersenne-Twister stream (np.random.seed / np.random.rand,
 * svirl/vars/vars.py:103-106).  key[624], *pos = RandomState(seed).get_state()[1:3]; skips `skip` doubles, then writes
 * n doubles; the state advances. */
int svl_mt19937_doubles(uint32_t *key, int *pos, unsigned long long skip, double *out, unsigned long long n);

/* psi0[n] = (1 - level*u1[n]) * exp(i*pi*level*(2*u2[n] - 1)) (svirl/vars/vars.py:106) for n values, written as
 * complex128 (complex_bytes 16) or complex64 (8); host arithmetic, bit-identical to the reference's numpy expression. */
int svl_seeded_psi(const double *u1, const double *u2, unsigned long long n, double level, void *out, int complex_bytes);

/* material tiling -> per-node flag plane (bits mm,mp,pm,pp: svirl/cuda/td.h:48-57).
 * mt == NULL: no tiling (all in-range cells are material). Must be called once before solving
 * and again whenever the tiling changes (mesh/grid.py:103-120). */
int svl_set_material(svl_ctx *ctx, const svl_buf *mt);

/* ---- TDGL (hot path 1) -------------------------------------------------------------- */

/* One Jacobi sweep: iterate_order_parameter_jacobi_step (svirl/cuda/td.h:5-133).
 * r_out = max over all nodes of max(|dRe|,|dIm|). */
int svl_td_psi_sweep(svl_ctx *ctx, double dt, double eps, const svl_buf *eps_field, const svl_buf *ab,
                     svl_buf *psi_rhs, const svl_buf *psi, svl_buf *psi_next, double langevin_c,
                     uint32_t jstep, uint32_t rand_t, double *r_out);
/* One Jacobi sweep: iterate_vector_potential_jacobi_step (svirl/cuda/td.h:311-463).
 * ab_phase is the kernel's `abi_ab_rhs` argument and may alias ab_next (quirk Q1). */
int svl_td_a_sweep(svl_ctx *ctx, double dt, double kappa2, double rho, double H, const svl_buf *psi,
                   const svl_buf *ab_phase, svl_buf *ab_rhs, const svl_buf *ab, svl_buf *ab_next,
                   double langevin_c, uint32_t jstep, uint32_t rand_t, double *r_out);

/* Whole psi-solve (svirl/solvers/td.py:157-218): <=1024 sweeps, stops after the first sweep
 * with r < stop_eps (exact reference rule int32(real(1e4 r/eps)) < 10000); psi is updated in
 * place (storage rotation, no copy).  sweeps_out = executed (= reference) sweep count. */
int svl_td_psi_solve(svl_ctx *ctx, double dt, double eps, const svl_buf *eps_field, const svl_buf *ab,
                     svl_buf *psi, double langevin_c, uint32_t rand_t, double stop_eps, int *sweeps_out);
/* Whole A-solve (svirl/solvers/td.py:252-325) including the link-phase aliasing quirk Q1. */
int svl_td_a_solve(svl_ctx *ctx, double dt, double kappa2, double rho, double H, const svl_buf *psi,
                   svl_buf *ab, double langevin_c, uint32_t rand_t, double stop_eps, int *sweeps_out);
/* Fixed vortices (svirl/vars/fixed_vortices.py, svirl/solvers/td.py:120-155, 207-216, 257-266, 319-325):
 *  svl_td_a_solve_ph  A-solve whose link phase comes from a separate, unperturbed buffer in every sweep;
 *  svl_edge_axpy_flat x[0:n_flat] += sign * y[0:n_flat] in the reference's packed edge order (xpy_r / xmy_r
 *                     launched with N = Nx*Ny on Na+Nb entries: quirk Q5, reproduced);
 *  svl_phase_lock     psi[n] <- |psi[n]| on a list of flat node indices (order_parameter_phase_lock). */
int svl_td_a_solve_ph(svl_ctx *ctx, double dt, double kappa2, double rho, double H, const svl_buf *psi,
                      const svl_buf *ab_phase, svl_buf *ab, double langevin_c, uint32_t rand_t, double stop_eps,
                      int *sweeps_out);
int svl_edge_axpy_flat(svl_ctx *ctx, svl_buf *x, const svl_buf *y, double sign, long long n_flat);
int svl_phase_lock(svl_ctx *ctx, svl_buf *psi, const svl_buf *lock_ns, int count);
/* Nt time steps of [psi-solve; A-solve if solveA] (svirl/solvers/td.py:342-367); rand_t is
 * incremented after every solve and returned.  sweeps[0..1] accumulate psi / A sweep counts. */
int svl_td_run(svl_ctx *ctx, int Nt, double dt, int solveA, double eps, const svl_buf *eps_field,
               double kappa2, double rho, double H, svl_buf *psi, svl_buf *ab, double langevin_psi,
               double langevin_A, uint32_t *rand_t, double stop_psi, double stop_A, long long *sweeps);

/* ---- CG (hot path 2) ---------------------------------------------------------------- */

/* free_energy_pseudodensity + gsum (svirl/cuda/observables.h:251-362, observables.py:124-149) */
int svl_free_energy(svl_ctx *ctx, double kappa2, double eps, const svl_buf *eps_field, double H,
                    const svl_buf *psi, const svl_buf *abei, const svl_buf *ab, double *E_out);
/* free_energy_jacobian_psi / _A (svirl/cuda/cg.h:16-121, 125-301); out is overwritten
 * (the reference zero-fills then accumulates, cg.py:128,165). */
int svl_jacobian_psi(svl_ctx *ctx, double kappa2, double eps, const svl_buf *eps_field, double H,
                     const svl_buf *psi, const svl_buf *abei, const svl_buf *ab, svl_buf *out);
int svl_jacobian_A(svl_ctx *ctx, double kappa2, double H, const svl_buf *psi, const svl_buf *abei,
                   const svl_buf *ab, svl_buf *out);
/* free_energy_conjgrad_coef_psi (cg.h:315-474): c[0..4]; _coef (cg.h:478-731): c[0..16] in the
 * kernel's flat order c00..c04,c10..c14,c20..c24,c30,c40.  Quirk Q11: eps_field is ignored. */
int svl_cg_coef_psi(svl_ctx *ctx, double kappa2, double eps, double H, const svl_buf *psi,
                    const svl_buf *dpsi, const svl_buf *abei, const svl_buf *ab, double *c5_out);
int svl_cg_coef(svl_ctx *ctx, double kappa2, double eps, double H, const svl_buf *psi, const svl_buf *dpsi,
                const svl_buf *abei, const svl_buf *ab, const svl_buf *dab, double *c17_out);
/* PR+ beta = max(sum g.(g-gp) / sum gp.gp, 0) (svirl/cuda/utils.h:13-70,140-146; cg.py:422-449);
 * works on SVL_NODE_C and SVL_EDGE.  nan -> 0 like the device fmax. */
int svl_cg_beta(svl_ctx *ctx, const svl_buf *g, const svl_buf *g_prev, double *beta_out);
/* z = alpha*x - y (axmy_c/_r, utils.h:97-114) and z = alpha*x + y (axpy_c/_r, utils.h:74-92) */
int svl_axmy(svl_ctx *ctx, const svl_buf *x, const svl_buf *y, svl_buf *z, double alpha);
int svl_axpy(svl_ctx *ctx, const svl_buf *x, const svl_buf *y, svl_buf *z, double alpha);

/* Fused CG iteration pieces (same arithmetic, fewer passes over HBM):
 *  svl_cg_begin: jacobians at (psi, ab) -> g; beta vs g_prev (if have_prev); d <- beta*d - g;
 *                17 (solveA) or 5 coefficients of the line-search polynomial -> c_out.
 *  svl_cg_end:   psi += alpha_psi*d_psi, ab += alpha_A*d_A; energy of the new state -> E_out.
 * The host keeps the reference's numpy/scipy line search between the two calls. */
int svl_cg_begin(svl_ctx *ctx, int solveA, int have_prev, double kappa2, double eps, const svl_buf *eps_field,
                 double H, const svl_buf *psi, const svl_buf *abei, const svl_buf *ab, svl_buf *g_psi,
                 svl_buf *g_psi_prev, svl_buf *d_psi, svl_buf *g_A, svl_buf *g_A_prev, svl_buf *d_A,
                 double *beta_inout /* [2]: psi, A */, double *c_out);
int svl_cg_end(svl_ctx *ctx, int solveA, double kappa2, double eps, const svl_buf *eps_field, double H,
               svl_buf *psi, const svl_buf *abei, svl_buf *ab, const svl_buf *d_psi, const svl_buf *d_A,
               double alpha_psi, double alpha_A, double *E_out);

/* Two-pass CG iteration (36R+2 bytes per node, the fused lower bound of the path; single GPU, no external potential):
 *  svl_cg_pass_a: [do_update: psi += alpha_psi*d_psi, ab += alpha_A*d_A, free energy of the new state -> E_out]
 *                 [do_grad: Jacobians at the (new) state -> g_psi, g_A IN PLACE (on entry they hold the previous
 *                  gradient when have_prev); have_prev: PR+ beta from the four sums -> beta_inout (also kept on the
 *                  device for pass b); otherwise beta_inout is taken as is (quirk Q6: beta persists across cg() calls)]
 *  svl_cg_pass_b: d <- beta*d - g; 17 (solveA) or 5 coefficients of the line-search polynomial -> c_out.
 * One iteration of svirl/solvers/cg.py:477-544 = pass_b, host line search, pass_a; the first pass_a of a cg() call
 * has do_update = 0, the last one may have do_grad = 0. */
int svl_cg_pass_a(svl_ctx *ctx, int solveA, int do_update, int do_grad, int have_prev, double kappa2, double eps,
                  const svl_buf *eps_field, double H, svl_buf *psi, const svl_buf *abei, svl_buf *ab,
                  const svl_buf *d_psi, const svl_buf *d_A, double alpha_psi, double alpha_A, svl_buf *g_psi,
                  svl_buf *g_A, double *beta_inout /* [2] */, double *E_out);
int svl_cg_pass_b(svl_ctx *ctx, int solveA, double kappa2, double eps, double H, const svl_buf *psi,
                  const svl_buf *abei, const svl_buf *ab, const svl_buf *g_psi, const svl_buf *g_A, svl_buf *d_psi,
                  svl_buf *d_A, double *c_out);

/* Opt-in host line search (replaces scipy.optimize.minimize(BFGS) of svirl/solvers/cg.py:378-419 and
 * the polyroots call of :227-235 when cfg.cg_line_search = 'native'): minimum of the coefficient polynomial in the
 * basin of (0, 0) by damped Newton on c / max|c|.  c: 17 (solveA) or 5 coefficients in the kernels' order. */
int svl_cg_line_search(const double *c, int solveA, double *alpha_out /* [2] */, int *iters_out);

/* ---- observables (svirl/cuda/observables.h:5-235) -------------------------------------- */
int svl_magnetic_field(svl_ctx *ctx, const svl_buf *abei, const svl_buf *ab, svl_buf *B_out);
int svl_current_density(svl_ctx *ctx, double kappa2, double H, const svl_buf *abei, const svl_buf *ab,
                        svl_buf *j_out);
int svl_supercurrent_density(svl_ctx *ctx, const svl_buf *psi, const svl_buf *abei, const svl_buf *ab,
                             svl_buf *js_out);
/* GPU pass of the vortex detector (svirl/observables/vortex_detector.py:55-72): writes the cell
 * indices n = i + (Nx-1) j (ascending) of every cell whose winding number v satisfies the
 * SUPERSET test |v| > 0.45 && |v - round(v)| < 0.15; the host re-tests candidates with the
 * reference's exact arithmetic.  count_out may exceed max_out (then call again). */
int svl_vortex_candidates(svl_ctx *ctx, double H, const svl_buf *psi, const svl_buf *ab, int64_t *cells_out,
                          double *v_out, size_t max_out, size_t *count_out);

/* ---- reductions (svirl/parallel/reduction.py:40-173): deterministic two-stage sums ------ */
int svl_sum(svl_ctx *ctx, const svl_buf *in, size_t n, double *out);
int svl_sum_v(svl_ctx *ctx, const svl_buf *in, size_t nv, int ne, double *out /* [ne] */);

/* ---- multi-GPU row slabs (new) -------------------------------------
 * One process per GPU; the context owns rows [j0, j1) (svl_create).  svl_slab_export writes
 * 144 bytes (the CUDA IPC handle of one arena that now holds the psi x3, a x3, b x3 planes and
 * the flag words, plus their 10 byte offsets) for the neighbours; svl_slab_connect opens the neighbours' handles (NULL at the ends of the chain).
 * From then on every psi / A launch pushes its boundary rows into the neighbours' halo rows by
 * direct peer stores and waits for theirs.  The MAX of the per-sweep residual slots over all
 * ranks goes through reduce_max_u64 (in place on n 64-bit values; exact, so the TD trajectory
 * does not depend on the slab count). */
int svl_slab_export(svl_ctx *ctx, svl_buf *psi, svl_buf *ab, void *handles_out);
int svl_slab_connect(svl_ctx *ctx, const void *lo_handles, int lo_j0, const void *hi_handles, int hi_j0);
int svl_slab_exchange(svl_ctx *ctx, svl_buf *buf);
/* Residual board (optional, replaces the callbacks below when connected): every rank exports one
 * 64-byte CUDA IPC handle, all ranks connect with the world x 64 bytes in rank order; the MAX of the
 * residual slots is then taken by one small kernel over peer memory (option "resid_board" = 0
 * switches back to the callbacks). */
int svl_slab_board_export(svl_ctx *ctx, void *handle_out);
int svl_slab_board_connect(svl_ctx *ctx, int rank, int world, const void *handles);
int svl_set_reduce_callback(svl_ctx *ctx, void (*reduce_max_u64)(unsigned long long *vals, int n));
/* same on device words, with the reduction enqueued on the context's stream (svl_get_stream) */
int svl_set_reduce_callback_device(svl_ctx *ctx, void (*reduce_max_dev)(unsigned long long *dvals, int n));
void *svl_get_stream(svl_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SVIRL_B200_H */
