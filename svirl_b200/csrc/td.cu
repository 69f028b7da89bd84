// TDGL time step: Jacobi sweeps of the psi and A equations and the solve drivers.
// Reference: svirl/cuda/td.h:5-133 (psi sweep), :311-463 (A sweep),
//            svirl/solvers/td.py:157-218, 252-325, 342-367 (drivers).
#include "common.cuh"

int svl_launch_psi_stream(svl_ctx *c, int K, double dt, double eps, const svl_buf *epsf, const svl_buf *ab,
                          const svl_buf *rhs, const svl_buf *psi, svl_buf *out, double lang_c, uint32_t rand_t,
                          unsigned long long *resid_slots);   // psi_stream.cu
int svl_psi_stream_fit_k(int K);
int svl_launch_psi_tile(svl_ctx *c, int K, double dt, double eps, const svl_buf *epsf, const svl_buf *ab,
                        const svl_buf *rhs, const svl_buf *psi, svl_buf *out, double lang_c, uint32_t rand_t,
                        unsigned long long *resid_slots);     // psi_tile.cu
int svl_td_small_run(svl_ctx *c, int Nt, double dt, int solveA, double eps, const svl_buf *epsf, double kappa2, double rho,
                     double H, svl_buf *psi, svl_buf *ab, double lang_psi, double lang_A, uint32_t *rand_t, double stop_psi,
                     double stop_A, long long *sweeps, bool *handled);   // td_small.cu
int svl_launch_a_tile(svl_ctx *c, int K, double dt, double kappa2, double rho, double H, const svl_buf *psi,
                      const svl_buf *rhs, const svl_buf *ab, svl_buf *out, double lang_c, uint32_t rand_t,
                      unsigned long long *resid_slots);       // a_tile.cu

// ----------------------------------------------------------------------------- plain psi sweep
// One thread per node; neighbours come through L1/L2.  NOISE: 0 none, 1 add + write back to rhs
// (kernel-level API at jstep 0), 2 add on every sweep without write back (solver: rhs may alias psi).
template <typename R, int NOISE>
__global__ void __launch_bounds__(256)
k_psi_sweep(Geo g, R dt, R eps, const R *__restrict__ epsf, const uint8_t *__restrict__ nf,
            const R *__restrict__ pa, const R *__restrict__ pb, typename V2<R>::type *rhs,
            const typename V2<R>::type *psi, typename V2<R>::type *out, R lang_c, uint32_t rand_t,
            unsigned long long *slot) {
    typedef typename V2<R>::type C;
    const R dx = (R)g.dx, dy = (R)g.dy, idx2 = (R)g.idx2, idy2 = (R)g.idy2;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = g.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    double r = 0.0;
    if (i < g.Nx && j < g.j1) {
        size_t n = g.at(i, j);
        unsigned f = nf[n];
        C p00 = psi[n];
        C nx;
        nx.x = 0; nx.y = 0;
        if (f) {
            C q = rhs[n];
            if (NOISE) {
                if (lang_c > (R)1.0e-32) {
                    uint32_t nn = (uint32_t)i + (uint32_t)g.Nx * (uint32_t)j;
                    q.x += lang_c * (rand_1<R>(nn, rand_t) - (R)0.5);
                    q.y += lang_c * (rand_2<R>(nn, rand_t) - (R)0.5);
                    if (NOISE == 1) rhs[n] = q;
                }
            }
            bool wW = f & (NF_MM | NF_MP), wE = f & (NF_PM | NF_PP);
            bool wS = f & (NF_MM | NF_PM), wN = f & (NF_MP | NF_PP);
            R e = epsf ? epsf[n] : eps;
            R sx = 0, sy = 0, tx = 0, ty = 0, s, c;
            if (wW) {   // U(-dx a[i-1,j]) psi[i-1,j] = (c + i s) psi
                C p = psi[n - 1];
                sincos_r<R>(dx * pa[n - 1], &s, &c);
                sx += c * p.x - s * p.y; sy += c * p.y + s * p.x;
            }
            if (wE) {   // U(dx a[i,j]) psi[i+1,j] = (c - i s) psi
                C p = psi[n + 1];
                sincos_r<R>(dx * pa[n], &s, &c);
                sx += c * p.x + s * p.y; sy += c * p.y - s * p.x;
            }
            if (wS) {
                C p = psi[n - g.P];
                sincos_r<R>(dy * pb[n - g.P], &s, &c);
                tx += c * p.x - s * p.y; ty += c * p.y + s * p.x;
            }
            if (wN) {
                C p = psi[n + g.P];
                sincos_r<R>(dy * pb[n], &s, &c);
                tx += c * p.x + s * p.y; ty += c * p.y - s * p.x;
            }
            R diag = (R)1.0 + dt * (q.x * q.x + q.y * q.y - e
                                    + (idx2 * (R)((int)wW + (int)wE) + idy2 * (R)((int)wS + (int)wN)));
            nx.x = (q.x + dt * (idx2 * sx + idy2 * tx)) / diag;
            nx.y = (q.y + dt * (idx2 * sy + idy2 * ty)) / diag;
        }
        out[n] = nx;
        r = fmax(fabs((double)(nx.x - p00.x)), fabs((double)(nx.y - p00.y)));
    }
    block_max_to_slot(r, slot);
}

template <typename R>
static int launch_psi_plain(svl_ctx *c, double dt, double eps, const svl_buf *epsf, const svl_buf *ab,
                            const svl_buf *rhs, const svl_buf *psi, svl_buf *out, double lang_c,
                            uint32_t rand_t, int noise, unsigned long long *slot) {
    typedef typename V2<R>::type C;
    const Geo &g = c->g;
    dim3 b(32, 8), gr((g.Nx + 31) / 32, (g.j1 - g.j0 + 7) / 8);
#define ARGS g, (R)dt, (R)eps, epsf ? (const R *)epsf->p[0] : nullptr, c->nf, (const R *)ab->p[0], \
             (const R *)ab->p[1], (C *)rhs->p[0], (const C *)psi->p[0], (C *)out->p[0], (R)lang_c, rand_t, slot
    if (noise == 0) k_psi_sweep<R, 0><<<gr, b, 0, c->stream>>>(ARGS);
    else if (noise == 1) k_psi_sweep<R, 1><<<gr, b, 0, c->stream>>>(ARGS);
    else k_psi_sweep<R, 2><<<gr, b, 0, c->stream>>>(ARGS);
#undef ARGS
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

int svl_launch_psi_sweep(svl_ctx *c, double dt, double eps, const svl_buf *epsf, const svl_buf *ab,
                         const svl_buf *rhs, const svl_buf *psi, svl_buf *out, double lang_c, uint32_t rand_t,
                         unsigned long long *slot) {
    int noise = lang_c > 1.0e-32 ? 2 : 0;
    if (c->rsize == 4) return launch_psi_plain<float>(c, dt, eps, epsf, ab, rhs, psi, out, lang_c, rand_t, noise, slot);
    return launch_psi_plain<double>(c, dt, eps, epsf, ab, rhs, psi, out, lang_c, rand_t, noise, slot);
}

// ----------------------------------------------------------------------------- A sweep
// One thread per node (i,j) updates the a-edge and the b-edge whose tail is that node.
// ph_a/ph_b (the reference's abi_ab_rhs) may alias oa/ob: each thread reads its own entry
// before it writes it and nobody else reads that entry (quirk Q1).
template <typename R, int NOISE>
__global__ void __launch_bounds__(256)
k_a_sweep(Geo g, R dt, R kappa2, R rho, R H, const uint8_t *__restrict__ nf,
          const typename V2<R>::type *__restrict__ psi, const R *ph_a, const R *ph_b, R *rhs_a, R *rhs_b,
          const R *a, const R *b, R *oa, R *ob, R lang_c, uint32_t rand_t, unsigned long long *slot) {
    typedef typename V2<R>::type C;
    const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy, idx2 = (R)g.idx2, idy2 = (R)g.idy2,
            idxy = (R)g.idxy;
    const R dt_rho = dt * rho, dtrk = dt_rho * kappa2;
    // the Jacobi diagonals are constants: multiply by their reciprocals (<= 1 ulp from the division)
    const R inv_da = (R)1.0 / ((R)1.0 + (R)2.0 * dtrk * idy2), inv_db = (R)1.0 / ((R)1.0 + (R)2.0 * dtrk * idx2);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = g.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    double r = 0.0;
    if (i < g.Nx && j < g.j1) {
        size_t n = g.at(i, j);
        const int P = g.P;
        unsigned f = nf[n];
        C p0 = psi[n];
        if (i < g.Nx - 1) {
            R q = rhs_a[n];
            if (NOISE) {
                if (lang_c > (R)1.0e-32) {
                    uint32_t ne = (uint32_t)i + (uint32_t)(g.Nx - 1) * (uint32_t)j;
                    q += lang_c * (rand_1<R>(ne, rand_t) - (R)0.5);
                    if (NOISE == 1) rhs_a[n] = q;
                }
            }
            R rh = 0, dd = 1;
            if (j == 0) { rh = (R)2.0 * kappa2 * H * idy; dd = 2; }
            else if (j + 1 == g.Ny) { rh = -(R)2.0 * kappa2 * H * idy; dd = 2; }
            R jl = 0;
            if (f & (NF_PM | NF_PP)) jl = idx * js_link<R, C>(p0, dx * ph_a[n], psi[n + 1]);
            R lo = 0, hi = 0;
            if (j > 0) lo = idy2 * a[n - P] - idxy * b[n - P] + idxy * b[n - P + 1];
            if (j + 1 < g.Ny) hi = idy2 * a[n + P] + idxy * b[n] - idxy * b[n + 1];
            R nx = (q + dt_rho * (jl + rh) + dtrk * dd * (lo + hi)) * inv_da;
            R old = a[n];
            oa[n] = nx;
            r = fabs((double)(nx - old));
        }
        if (j < g.Ny - 1) {
            R q = rhs_b[n];
            if (NOISE) {
                if (lang_c > (R)1.0e-32) {
                    uint32_t ne = (uint32_t)((size_t)(g.Nx - 1) * g.Ny) + (uint32_t)i + (uint32_t)g.Nx * (uint32_t)j;
                    q += lang_c * (rand_2<R>(ne, rand_t) - (R)0.5);
                    if (NOISE == 1) rhs_b[n] = q;
                }
            }
            R rh = 0, dd = 1;
            if (i == 0) { rh = -(R)2.0 * kappa2 * H * idx; dd = 2; }
            else if (i + 1 == g.Nx) { rh = (R)2.0 * kappa2 * H * idx; dd = 2; }
            R jl = 0;
            if (f & (NF_MP | NF_PP)) jl = idy * js_link<R, C>(p0, dy * ph_b[n], psi[n + P]);
            R lo = 0, hi = 0;
            if (i > 0) lo = idx2 * b[n - 1] - idxy * a[n - 1] + idxy * a[n - 1 + P];
            if (i + 1 < g.Nx) hi = idx2 * b[n + 1] + idxy * a[n] - idxy * a[n + P];
            R nx = (q + dt_rho * (jl + rh) + dtrk * dd * (lo + hi)) * inv_db;
            R old = b[n];
            ob[n] = nx;
            r = fmax(r, fabs((double)(nx - old)));
        }
    }
    block_max_to_slot(r, slot);
}

template <typename R>
static int launch_a(svl_ctx *c, double dt, double kappa2, double rho, double H, const svl_buf *psi,
                    const svl_buf *ph, const svl_buf *rhs, const svl_buf *ab, svl_buf *out, double lang_c,
                    uint32_t rand_t, int noise, unsigned long long *slot) {
    typedef typename V2<R>::type C;
    const Geo &g = c->g;
    dim3 b(32, 8), gr((g.Nx + 31) / 32, (g.j1 - g.j0 + 7) / 8);
#define ARGS g, (R)dt, (R)kappa2, (R)rho, (R)H, c->nf, (const C *)psi->p[0], (const R *)ph->p[0], (const R *)ph->p[1], \
             (R *)rhs->p[0], (R *)rhs->p[1], (const R *)ab->p[0], (const R *)ab->p[1], (R *)out->p[0], (R *)out->p[1],  \
             (R)lang_c, rand_t, slot
    if (noise == 0) k_a_sweep<R, 0><<<gr, b, 0, c->stream>>>(ARGS);
    else if (noise == 1) k_a_sweep<R, 1><<<gr, b, 0, c->stream>>>(ARGS);
    else k_a_sweep<R, 2><<<gr, b, 0, c->stream>>>(ARGS);
#undef ARGS
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

int svl_launch_a_sweep(svl_ctx *c, double dt, double kappa2, double rho, double H, const svl_buf *psi,
                       const svl_buf *ph, const svl_buf *rhs, const svl_buf *ab, svl_buf *out, double lang_c,
                       uint32_t rand_t, int noise, unsigned long long *slot) {
    if (c->rsize == 4) return launch_a<float>(c, dt, kappa2, rho, H, psi, ph, rhs, ab, out, lang_c, rand_t, noise, slot);
    return launch_a<double>(c, dt, kappa2, rho, H, psi, ph, rhs, ab, out, lang_c, rand_t, noise, slot);
}

// ----------------------------------------------------------------------------- stop rule
// Exact reference decision (td.h:124-132 + td.py:198-201): int32(real_t(1e4*r/eps) clamped at 1e8)
// and 1e-4*int < 1  <=>  int < 10000.
static bool stop_rule(const svl_ctx *c, double r, double eps) {
    double v;
    if (c->rsize == 4) v = (double)(float)(1.0e4 * r / (double)(float)eps);
    else v = 1.0e4 * r / eps;
    if (v > 1.0e8) v = 1.0e8;
    return (int)v < 10000;
}

// The same decision on the device (pipelined solves): go <- "the first sweep that meets the stop rule is sweep done-1",
// i.e. the batch that was just launched ends exactly where the reference stops.  IEEE double / float arithmetic as on
// the host, so both sides always agree.
__global__ void k_psi_gate(const unsigned long long *slots, int done, double eps, int f32, int *go) {
    int bad = 0;
    for (int s = threadIdx.x; s < done; s += 32) {
        const double r = __longlong_as_double((long long)slots[s]);
        double v;
        if (f32) v = (double)(float)(1.0e4 * r / (double)(float)eps);
        else v = 1.0e4 * r / eps;
        if (v > 1.0e8) v = 1.0e8;
        const bool stop = (int)v < 10000;
        if (stop != (s == done - 1)) bad = 1;
    }
    bad = __any_sync(0xffffffffu, bad);
    if (threadIdx.x == 0) *go = bad ? 0 : 1;
}

static inline double slot_value(unsigned long long bits) {
    double d;
    memcpy(&d, &bits, sizeof(d));
    return d;
}

// Residual slots [first, first+count) -> host, in two halves so that work can be enqueued between them: the device
// part (MAX over ranks, copy, event) and the host part (wait for the EVENT, not for the whole stream).
static bool resid_board(const svl_ctx *c) { return c->board_world > 1 && c->opt_resid_board && !c->opt_slab_nocomm; }
static int read_resid_enqueue(svl_ctx *c, int first, int count) {
    // slabs: MAX over ranks (bit patterns of non-negative doubles order like integers; exact)
    if (resid_board(c)) SVL_TRY(svl_board_allmax(c, first, count));                    // peer memory, no host, no NCCL
    else if (c->reduce_max_dev && !c->opt_slab_nocomm) c->reduce_max_dev(c->d_resid + first, count); // NCCL, enqueued on c->stream
    SVL_CHECK(cudaMemcpyAsync(c->h_resid + first, c->d_resid + first, (size_t)count * sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, c->stream));
    SVL_CHECK(cudaEventRecord(c->ev_go, c->stream));
    return 0;
}
static int read_resid_finish(svl_ctx *c, int first, int count) {
    SVL_CHECK(cudaEventSynchronize(c->ev_go));
    SVL_TRY(svl_peer_error(c));
    if (!resid_board(c) && c->reduce_max_u64 && !c->reduce_max_dev) c->reduce_max_u64(c->h_resid + first, count);
    return 0;
}
static int read_resid(svl_ctx *c, int first, int count) {
    SVL_TRY(read_resid_enqueue(c, first, count));
    return read_resid_finish(c, first, count);
}
static void resid_flip(svl_ctx *c) {
    c->resid_bank ^= 1;
    c->d_resid = c->d_resid_base + (size_t)c->resid_bank * SVL_MAX_SWEEPS;
    c->h_resid = c->h_resid_base + (size_t)c->resid_bank * SVL_MAX_SWEEPS;
}

// How many sweeps to launch before the next read-back.
// First batch: the previous solve's count, lowered by its last drop (counts fall monotonically
// while the system relaxes); 8 when there is no history.
static int first_batch(int last, int before) {
    if (last <= 0) return 8;
    int drop = before > last ? before - last : 0;
    int n = last - drop;
    return n < 1 ? 1 : n;
}
// Continuation: the Jacobi update norm decays geometrically, r_{s+1} ~ rho r_s, so the number
// of sweeps still missing is about log(eps/r)/log(rho); launch that many (bounded), then look again.
static int more_sweeps(double r_prev, double r_last, double eps) {
    if (!(r_last > 0.0) || !(r_prev > r_last)) return 2;
    double n = ceil(log(eps / r_last) / log(r_last / r_prev));
    if (!(n >= 1.0)) return 1;
    return n > 32.0 ? 32 : (int)n;
}

static int check_kinds(const svl_buf *psi, const svl_buf *ab, const svl_buf *epsf) {
    SVL_REQUIRE(psi && psi->kind == SVL_NODE_C, "psi must be SVL_NODE_C");
    SVL_REQUIRE(ab && ab->kind == SVL_EDGE, "ab must be SVL_EDGE");
    SVL_REQUIRE(!epsf || epsf->kind == SVL_NODE_R, "eps_field must be SVL_NODE_R");
    return 0;
}

// ----------------------------------------------------------------------------- single-sweep ABI
extern "C" int svl_td_psi_sweep(svl_ctx *c, double dt, double eps, const svl_buf *epsf, const svl_buf *ab,
                                svl_buf *rhs, const svl_buf *psi, svl_buf *out, double lang_c, uint32_t jstep,
                                uint32_t rand_t, double *r_out) {
    SVL_REQUIRE(c, "null context");
    SVL_TRY(check_kinds(psi, ab, epsf));
    SVL_REQUIRE(rhs && rhs->kind == SVL_NODE_C && out && out->kind == SVL_NODE_C, "rhs/out must be SVL_NODE_C");
    SVL_REQUIRE(out != psi, "psi_next must differ from psi");
    SVL_CHECK(cudaMemsetAsync(c->d_resid, 0, sizeof(unsigned long long), c->stream));
    int noise = (jstep == 0 && lang_c > 1.0e-32) ? 1 : 0;
    if (c->rsize == 4) SVL_TRY(launch_psi_plain<float>(c, dt, eps, epsf, ab, rhs, psi, out, lang_c, rand_t, noise, c->d_resid));
    else SVL_TRY(launch_psi_plain<double>(c, dt, eps, epsf, ab, rhs, psi, out, lang_c, rand_t, noise, c->d_resid));
    SVL_TRY(read_resid(c, 0, 1));
    if (r_out) *r_out = slot_value(c->h_resid[0]);
    return 0;
}

extern "C" int svl_td_a_sweep(svl_ctx *c, double dt, double kappa2, double rho, double H, const svl_buf *psi,
                              const svl_buf *ph, svl_buf *rhs, const svl_buf *ab, svl_buf *out, double lang_c,
                              uint32_t jstep, uint32_t rand_t, double *r_out) {
    SVL_REQUIRE(c, "null context");
    SVL_REQUIRE(psi && psi->kind == SVL_NODE_C, "psi must be SVL_NODE_C");
    SVL_REQUIRE(ph && rhs && ab && out && ph->kind == SVL_EDGE && rhs->kind == SVL_EDGE && ab->kind == SVL_EDGE &&
                out->kind == SVL_EDGE, "edge buffers required");
    SVL_REQUIRE(out != ab, "ab_next must differ from ab");
    SVL_CHECK(cudaMemsetAsync(c->d_resid, 0, sizeof(unsigned long long), c->stream));
    int noise = (jstep == 0 && lang_c > 1.0e-32) ? 1 : 0;
    SVL_TRY(svl_launch_a_sweep(c, dt, kappa2, rho, H, psi, ph, rhs, ab, out, lang_c, rand_t, noise, c->d_resid));
    SVL_TRY(read_resid(c, 0, 1));
    if (r_out) *r_out = slot_value(c->h_resid[0]);
    return 0;
}

// ----------------------------------------------------------------------------- psi solve
// Storage rotation instead of the reference's copy + ping-pong: B0 holds iterate 0 and is also
// the right-hand side; sweep s reads in(s) and writes out(s):
//   in(0) = B0, in(s odd) = S1, in(s even >= 2) = S2;   out(s even) = S1, out(s odd) = S2.
// The host reads the per-sweep residual maxima once per solve: sweeps are launched up to the
// count predicted from the previous time step, and the exact reference stop sweep is
// recovered afterwards (the input of the last sweep is still intact, so an overshoot by one
// is free; a larger overshoot replays from B0; an undershoot continues sweep by sweep).
struct PsiSolveArgs {
    double dt, eps; const svl_buf *epsf; const svl_buf *ab; double lang_c; uint32_t rand_t;
};

// Buffer rotation state: `cur` holds the newest iterate, the next launch writes to S[toggle].
struct PsiIter {
    svl_buf *B0, *S[2];
    svl_buf *cur, *prev;
    int toggle;
    int lastK;
    void reset() { cur = B0; prev = nullptr; toggle = 0; lastK = 0; }
};

// Launch sweeps [s0, s1).  With the temporally blocked kernel, runs of up to psi_k sweeps go into
// one launch; tail_single keeps the final sweep in a launch of its own so that the iterate before
// it survives (an overshoot by one sweep then costs nothing).
static int psi_launch_range(svl_ctx *c, const PsiSolveArgs &A, PsiIter &it, int s0, int s1, bool tail_single) {
    int s = s0;
    while (s < s1) {
        int K = 0;                                   // 0: plain per-node kernel (one sweep)
        if (c->opt_psi_kernel >= 1) {
            int left = s1 - s;
            if (tail_single && left > 1) left -= 1;
            int want = c->opt_psi_k < left ? c->opt_psi_k : left;
            if (c->opt_tma || c->opt_psi_kernel == 2) K = svl_psi_stream_fit_k(want);
            else K = want >= 4 ? 4 : 0;              // plain-load staging is built for K = 4 only
        }
        svl_buf *out = it.S[it.toggle];
        bool pushed = false;                         // the tile kernel waits for its halos and pushes its own
        if (K > 0 && c->opt_psi_kernel == 2) {
            SVL_TRY(svl_launch_psi_tile(c, K, A.dt, A.eps, A.epsf, A.ab, it.B0, it.cur, out, A.lang_c, A.rand_t, c->d_resid + s));
            pushed = true;
        } else if (K > 0) {
            SVL_TRY(svl_slab_wait(c));               // neighbours' halo rows of the input have arrived
            SVL_TRY(svl_launch_psi_stream(c, K, A.dt, A.eps, A.epsf, A.ab, it.B0, it.cur, out, A.lang_c, A.rand_t, c->d_resid + s));
        } else {
            K = 1;
            SVL_TRY(svl_slab_wait(c));
            SVL_TRY(svl_launch_psi_sweep(c, A.dt, A.eps, A.epsf, A.ab, it.B0, it.cur, out, A.lang_c, A.rand_t, c->d_resid + s));
        }
        if (!pushed) SVL_TRY(svl_slab_push_psi(c, out));   // my boundary rows -> neighbours' halos (peer stores)
        it.prev = it.cur; it.cur = out; it.toggle ^= 1; it.lastK = K;
        s += K;
    }
    return 0;
}

// K of the first launch of a solve that will run `upto` sweeps in its first batch (what psi_launch_range picks)
static int first_launch_K(svl_ctx *c, int upto) {
    int want = c->opt_psi_k < upto ? c->opt_psi_k : upto;
    return svl_psi_stream_fit_k(want);
}

// Pipelined solves (kappa = inf time stepping, option "pipeline").  The host learns the residuals of a batch ~10 us
// (one GPU) to ~25 us (slabs: rendezvous over the residual board) after the batch ends, and the GPU would idle until the
// next solve's first launch arrives.  Instead, right behind the batch the stream gets: the residual read-back, a
// one-warp kernel that evaluates the reference's stop rule on the device (k_psi_gate), and the NEXT solve's first
// launch, which does nothing unless the gate says that this batch ended exactly at the stop sweep (the common case:
// sweep counts are predicted from the previous step).  The host takes the same decision from the same numbers; on a
// miss the pre-issued launch was a no-op and the solve carries on as before (replay / continuation).
struct PsiNext { int allow; uint32_t rand_t; };

static int psi_solve(svl_ctx *c, double dt, double eps, const svl_buf *epsf, const svl_buf *ab, svl_buf *psi,
                     double lang_c, uint32_t rand_t, double stop_eps, int *sweeps_out, const PsiNext *next) {
    PsiIter it;
    it.B0 = psi;
    SVL_TRY(svl_scratch_node(c, 0, &it.S[0]));
    SVL_TRY(svl_scratch_node(c, 1, &it.S[1]));
    it.reset();
    PsiSolveArgs A = {dt, eps, epsf, ab, lang_c, rand_t};
    int done = 0, nstop = -1;
    const bool pre = c->spec_issued != 0;            // this solve's first launch is already in the stream
    if (pre) {
        resid_flip(c);                               // ... it wrote to the other bank (zeroed before the launch)
        it.prev = it.B0; it.cur = it.S[0]; it.toggle = 1; it.lastK = c->spec_K;
        c->spec_issued = 0;
    } else {
        SVL_CHECK(cudaMemsetAsync(c->d_resid, 0, SVL_MAX_SWEEPS * sizeof(unsigned long long), c->stream));
    }
    const bool can_pipe = next && next->allow && c->opt_pipeline && c->opt_psi_kernel == 2 &&
                          (!c->slab_on || c->opt_slab_nocomm || resid_board(c));
    bool hit = false;
    while (nstop < 0) {
        int upto;
        if (done == 0) upto = first_batch(c->pred_psi, c->pred_psi2);
        else upto = done + more_sweeps(done >= 2 ? slot_value(c->h_resid[done - 2]) : 0.0, slot_value(c->h_resid[done - 1]), stop_eps);
        if (upto > SVL_MAX_SWEEPS) upto = SVL_MAX_SWEEPS;
        int from = done;
        if (done == 0 && pre) {                      // sweeps [0, spec_K) are in flight; the batch size is the one
            from = it.lastK;                         // the issuer assumed (same prediction inputs)
            if (upto < from) upto = from;
        }
        SVL_TRY(psi_launch_range(c, A, it, from, upto, c->opt_psi_kernel == 0));
        SVL_TRY(read_resid_enqueue(c, done, upto - done));
        // ---- pre-issue the next solve's first launch behind the gate
        bool spec = false;
        unsigned long long save_epoch = c->epoch_psi, save_waited = c->waited;
        int specK = 0;
        if (can_pipe && upto < SVL_MAX_SWEEPS) {
            k_psi_gate<<<1, 32, 0, c->stream>>>(c->d_resid, upto, stop_eps, c->rsize == 4 ? 1 : 0, c->d_go);
            SVL_CHECK(cudaGetLastError());
            // if the gate opens, this solve ends with res = it.cur, nstop = upto: the state the next solve starts from
            SVL_TRY(svl_swap(c, psi, it.cur));
            specK = first_launch_K(c, first_batch(upto, c->pred_psi));
            resid_flip(c);
            SVL_CHECK(cudaMemsetAsync(c->d_resid, 0, SVL_MAX_SWEEPS * sizeof(unsigned long long), c->stream));
            c->spec_gate = c->d_go;
            int rc = svl_launch_psi_tile(c, specK, dt, eps, epsf, ab, psi, psi, it.S[0], lang_c, next->rand_t, c->d_resid);
            c->spec_gate = nullptr;
            resid_flip(c);
            if (rc) { svl_swap(c, psi, it.cur); return rc; }
            spec = true;
        }
        SVL_TRY(read_resid_finish(c, done, upto - done));
        for (int s = done; s < upto; s++)
            if (stop_rule(c, slot_value(c->h_resid[s]), stop_eps)) { nstop = s + 1; break; }
        done = upto;
        if (spec) {
            if (nstop == done) { hit = true; c->spec_issued = 1; c->spec_K = specK; c->stat_spec_hit += 1; }
            else {                                   // the gate stayed shut: undo the host-side bookkeeping of the no-op
                SVL_TRY(svl_swap(c, psi, it.cur));
                c->epoch_psi = save_epoch; c->waited = save_waited;
                c->stat_spec_miss += 1;
            }
        }
        if (nstop < 0 && done >= SVL_MAX_SWEEPS) nstop = SVL_MAX_SWEEPS;   // reference: loop exhausts, keeps last iterate
    }
    if (hit) {                                       // psi already holds the result; the next solve is under way
        c->pred_psi2 = c->pred_psi;
        c->pred_psi = nstop;
        c->stat_psi_sweeps += nstop;
        if (sweeps_out) *sweeps_out = nstop;
        return 0;
    }
    // `done` sweeps were executed; the reference stops after nstop <= done sweeps
    svl_buf *res = it.cur;
    if (nstop < done) {
        int first_of_last = done - it.lastK;            // first sweep index of the last launch
        if (nstop == first_of_last && it.prev != it.B0) {
            res = it.prev;                               // the input of the last launch is the answer
        } else if (nstop > first_of_last) {
            // overshoot inside the last launch: redo only that launch, shorter (its input is intact)
            c->stat_replays += 1;
            it.cur = it.prev; it.toggle ^= 1;
            SVL_TRY(psi_launch_range(c, A, it, first_of_last, nstop, false));
            res = it.cur;
        } else {
            c->stat_replays += 1;
            it.reset();
            SVL_TRY(psi_launch_range(c, A, it, 0, nstop, false));
            res = it.cur;
        }
    }
    SVL_TRY(svl_slab_wait(c));                       // halos of the result are complete before anyone reads them
    SVL_TRY(svl_swap(c, psi, res));
    c->pred_psi2 = c->pred_psi;
    c->pred_psi = nstop;
    c->stat_psi_sweeps += nstop;
    if (sweeps_out) *sweeps_out = nstop;
    return 0;
}

extern "C" int svl_td_psi_solve(svl_ctx *c, double dt, double eps, const svl_buf *epsf, const svl_buf *ab,
                                svl_buf *psi, double lang_c, uint32_t rand_t, double stop_eps, int *sweeps_out) {
    SVL_REQUIRE(c, "null context");
    SVL_TRY(check_kinds(psi, ab, epsf));
    c->spec_issued = 0;                              // a stand-alone solve never continues a pre-issued launch
    return psi_solve(c, dt, eps, epsf, ab, psi, lang_c, rand_t, stop_eps, sweeps_out, nullptr);
}

// ----------------------------------------------------------------------------- A solve
// The link phase of sweep s comes from iterate s - (s mod 2) (quirk Q1): for even s that is the
// sweep's own input, which is what lets a_tile.cu fuse sweeps (2m, 2m+1) into one launch.
struct ASolveArgs {
    double dt, kappa2, rho, H; const svl_buf *psi; double lang_c; uint32_t rand_t;
    const svl_buf *ph_fixed;      // fixed vortices: the link phase comes from this buffer in every sweep (no Q1 aliasing)
};

// Rotation state of the A solve.  `cur` holds iterate s, `even` the newest even iterate <= s
// (the link-phase source of sweeps s and s+1 when s is even: quirk Q1).  B0 (the caller's buffer)
// is iterate 0 and the right-hand side and is never written.
struct AIter {
    svl_buf *B0, *S1, *S2;
    svl_buf *cur, *even;
    int s;
    int l_s, l_K;                 // the last launch: first sweep, sweeps, and the state before it
    svl_buf *l_cur, *l_even;
    void reset() { cur = even = B0; s = 0; l_s = 0; l_K = 0; l_cur = l_even = B0; }
};

// Launch sweeps [it.s, upto).  Even sweeps go to the tile kernel (a_tile.cu), two per launch where
// possible; an odd sweep can only follow a lone even one (continuation after an undershoot) and
// takes the per-node kernel with the phase buffer aliasing the output, exactly like the reference.
static int a_advance(svl_ctx *c, const ASolveArgs &A, AIter &it, int upto) {
    const int noise = A.lang_c > 1.0e-32 ? 2 : 0;
    while (it.s < upto) {
        it.l_s = it.s; it.l_cur = it.cur; it.l_even = it.even;
        int K = 1;
        svl_buf *out;
        bool pushed = false;                         // the tile kernel waits for its halos and pushes its own
        if (A.ph_fixed) {
            // separate, unperturbed phase buffer (svirl/solvers/td.py:257-266): plain ping-pong of per-node sweeps
            out = it.cur == it.S1 ? it.S2 : it.S1;
            SVL_TRY(svl_slab_wait(c));
            SVL_TRY(svl_launch_a_sweep(c, A.dt, A.kappa2, A.rho, A.H, A.psi, A.ph_fixed, it.B0, it.cur, out, A.lang_c,
                                       A.rand_t, noise, c->d_resid + it.s));
            it.cur = it.even = out;
        } else if ((it.s & 1) == 0) {
            out = it.cur == it.S1 ? it.S2 : it.S1;
            if (c->opt_a_kernel >= 1) {
                K = upto - it.s >= 2 ? 2 : 1;
                SVL_TRY(svl_launch_a_tile(c, K, A.dt, A.kappa2, A.rho, A.H, A.psi, it.B0, it.cur, out, A.lang_c, A.rand_t,
                                          c->d_resid + it.s));
                pushed = true;
            } else {
                SVL_TRY(svl_slab_wait(c));
                SVL_TRY(svl_launch_a_sweep(c, A.dt, A.kappa2, A.rho, A.H, A.psi, it.cur, it.B0, it.cur, out, A.lang_c,
                                           A.rand_t, noise, c->d_resid + it.s));
            }
            it.cur = out;
            if (K == 2) it.even = out;
        } else {
            out = it.even == it.B0 ? (it.cur == it.S1 ? it.S2 : it.S1) : it.even;
            SVL_TRY(svl_slab_wait(c));
            SVL_TRY(svl_launch_a_sweep(c, A.dt, A.kappa2, A.rho, A.H, A.psi, it.even, it.B0, it.cur, out, A.lang_c, A.rand_t,
                                       noise, c->d_resid + it.s));
            it.cur = it.even = out;
        }
        if (!pushed) SVL_TRY(svl_slab_push_ab(c, out));
        it.s += K; it.l_K = K;
    }
    return 0;
}

static int a_solve(svl_ctx *c, double dt, double kappa2, double rho, double H, const svl_buf *psi, const svl_buf *ph_fixed,
                   svl_buf *ab, double lang_c, uint32_t rand_t, double stop_eps, int *sweeps_out);

extern "C" int svl_td_a_solve(svl_ctx *c, double dt, double kappa2, double rho, double H, const svl_buf *psi,
                              svl_buf *ab, double lang_c, uint32_t rand_t, double stop_eps, int *sweeps_out) {
    return a_solve(c, dt, kappa2, rho, H, psi, nullptr, ab, lang_c, rand_t, stop_eps, sweeps_out);
}

extern "C" int svl_td_a_solve_ph(svl_ctx *c, double dt, double kappa2, double rho, double H, const svl_buf *psi,
                                 const svl_buf *ab_phase, svl_buf *ab, double lang_c, uint32_t rand_t, double stop_eps,
                                 int *sweeps_out) {
    SVL_REQUIRE(ab_phase && ab_phase->kind == SVL_EDGE && ab_phase != ab, "ab_phase must be a distinct SVL_EDGE buffer");
    return a_solve(c, dt, kappa2, rho, H, psi, ab_phase, ab, lang_c, rand_t, stop_eps, sweeps_out);
}

static int a_solve(svl_ctx *c, double dt, double kappa2, double rho, double H, const svl_buf *psi, const svl_buf *ph_fixed,
                   svl_buf *ab, double lang_c, uint32_t rand_t, double stop_eps, int *sweeps_out) {
    SVL_REQUIRE(c, "null context");
    SVL_REQUIRE(psi && psi->kind == SVL_NODE_C, "psi must be SVL_NODE_C");
    SVL_REQUIRE(ab && ab->kind == SVL_EDGE, "ab must be SVL_EDGE");
    AIter it;
    it.B0 = ab;
    SVL_TRY(svl_scratch_edge(c, 0, &it.S1));
    SVL_TRY(svl_scratch_edge(c, 1, &it.S2));
    it.reset();
    ASolveArgs A = {dt, kappa2, rho, H, psi, lang_c, rand_t, ph_fixed};
    SVL_CHECK(cudaMemsetAsync(c->d_resid, 0, SVL_MAX_SWEEPS * sizeof(unsigned long long), c->stream));
    int done = 0, nstop = -1;
    while (nstop < 0) {
        int upto;
        if (done == 0) upto = first_batch(c->pred_A, c->pred_A2);
        else upto = done + more_sweeps(done >= 2 ? slot_value(c->h_resid[done - 2]) : 0.0, slot_value(c->h_resid[done - 1]), stop_eps);
        if (upto > SVL_MAX_SWEEPS) upto = SVL_MAX_SWEEPS;
        SVL_TRY(a_advance(c, A, it, upto));
        SVL_TRY(read_resid(c, done, upto - done));
        for (int s = done; s < upto; s++)
            if (stop_rule(c, slot_value(c->h_resid[s]), stop_eps)) { nstop = s + 1; break; }
        done = upto;
        if (nstop < 0 && done >= SVL_MAX_SWEEPS) nstop = SVL_MAX_SWEEPS;
    }
    svl_buf *res = it.cur;
    if (nstop < done) {
        if (nstop == it.l_s && nstop > 0) {
            res = it.l_cur;                              // the input of the last launch (still intact) is the answer
        } else if (nstop > it.l_s) {
            // the first sweep of the last pair converged: redo that launch as a single sweep
            c->stat_replays += 1;
            it.s = it.l_s; it.cur = it.l_cur; it.even = it.l_even;
            SVL_TRY(a_advance(c, A, it, nstop));
            res = it.cur;
        } else {
            c->stat_replays += 1;
            it.reset();
            SVL_TRY(a_advance(c, A, it, nstop));
            res = it.cur;
        }
    }
    SVL_TRY(svl_slab_wait(c));
    SVL_TRY(svl_swap(c, ab, res));
    c->pred_A2 = c->pred_A;
    c->pred_A = nstop;
    c->stat_A_sweeps += nstop;
    if (sweeps_out) *sweeps_out = nstop;
    return 0;
}

// ----------------------------------------------------------------------------- outer loop
extern "C" int svl_td_run(svl_ctx *c, int Nt, double dt, int solveA, double eps, const svl_buf *epsf, double kappa2,
                          double rho, double H, svl_buf *psi, svl_buf *ab, double lang_psi, double lang_A,
                          uint32_t *rand_t, double stop_psi, double stop_A, long long *sweeps) {
    SVL_REQUIRE(c && rand_t, "null argument");
    {   // small grids: the whole call in one launch of one thread-block cluster (option "graphs", td_small.cu)
        SVL_TRY(check_kinds(psi, ab, epsf));
        bool handled = false;
        SVL_TRY(svl_td_small_run(c, Nt, dt, solveA, eps, epsf, kappa2, rho, H, psi, ab, lang_psi, lang_A, rand_t, stop_psi,
                                 stop_A, sweeps, &handled));
        if (handled) return 0;
    }
    SVL_TRY(check_kinds(psi, ab, epsf));
    c->spec_issued = 0;
    for (int t = 0; t < Nt; t++) {
        int n = 0;
        PsiNext next = {(!solveA && t + 1 < Nt) ? 1 : 0, *rand_t + 1u};
        SVL_TRY(psi_solve(c, dt, eps, epsf, ab, psi, lang_psi, *rand_t, stop_psi, &n, &next));
        *rand_t += 1u;                       // td.py:204
        if (sweeps) sweeps[0] += n;
        if (solveA) {
            SVL_TRY(svl_td_a_solve(c, dt, kappa2, rho, H, psi, ab, lang_A, *rand_t, stop_A, &n));
            *rand_t += 1u;                   // td.py:313
            if (sweeps) sweeps[1] += n;
        }
    }
    return 0;
}

// ----------------------------------------------------------------------------- fixed-vortex helpers
// x[n] += sign * y[n] for the first n_flat entries of the reference's PACKED edge array (a then b):
// the reference launches xpy_r / xmy_r with N = Nx*Ny on an array of Na + Nb entries
// (svirl/solvers/td.py:128,151,260,322 -- SURVEY quirk Q5), so all of a and only the first
// N - Na = Ny entries of b are touched.  Reproduced on the pitched planes.
template <typename R>
__global__ void k_edge_axpy_flat(Geo g, R *xa, R *xb, const R *ya, const R *yb, R sign, long long n_flat) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = g.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.Nx || j >= g.j1) return;
    const size_t n = g.at(i, j);
    const long long Na = (long long)(g.Nx - 1) * g.Ny;
    if (i < g.Nx - 1 && (long long)i + (long long)(g.Nx - 1) * j < n_flat) xa[n] += sign * ya[n];
    if (j < g.Ny - 1 && Na + (long long)i + (long long)g.Nx * j < n_flat) xb[n] += sign * yb[n];
}

extern "C" int svl_edge_axpy_flat(svl_ctx *c, svl_buf *x, const svl_buf *y, double sign, long long n_flat) {
    SVL_REQUIRE(c && x && y && x->kind == SVL_EDGE && y->kind == SVL_EDGE, "two SVL_EDGE buffers required");
    dim3 b(32, 8), gr((c->g.Nx + 31) / 32, (c->g.j1 - c->g.j0 + 7) / 8);
    if (c->rsize == 4)
        k_edge_axpy_flat<float><<<gr, b, 0, c->stream>>>(c->g, (float *)x->p[0], (float *)x->p[1], (const float *)y->p[0],
                                                         (const float *)y->p[1], (float)sign, n_flat);
    else
        k_edge_axpy_flat<double><<<gr, b, 0, c->stream>>>(c->g, (double *)x->p[0], (double *)x->p[1], (const double *)y->p[0],
                                                          (const double *)y->p[1], sign, n_flat);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

// order_parameter_phase_lock (svirl/cuda/td.h:296-307): psi[n] <- |psi[n]| for the listed flat node
// indices (duplicates allowed: the operation is idempotent)
template <typename C>
__global__ void k_phase_lock(Geo g, C *psi, const int *ns, int count) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= count) return;
    const int n = ns[l], i = n % g.Nx, j = n / g.Nx;
    if (j < g.j0 || j >= g.j1) return;
    C *p = psi + g.at(i, j);
    C v = *p;
    v.x = sqrt(v.x * v.x + v.y * v.y);      // abs(complex): hypot without the scaling, like pycuda::abs for moderate values
    v.y = 0;
    *p = v;
}

extern "C" int svl_phase_lock(svl_ctx *c, svl_buf *psi, const svl_buf *lock_ns, int count) {
    SVL_REQUIRE(c && psi && psi->kind == SVL_NODE_C, "psi must be SVL_NODE_C");
    SVL_REQUIRE(lock_ns && lock_ns->kind == SVL_FLAT && lock_ns->esize == 4 && (size_t)count <= lock_ns->n,
                "lock_ns must be a flat int32 buffer of at least count entries");
    if (count <= 0) return 0;
    if (c->rsize == 4) k_phase_lock<float2><<<svl_nblocks(count, 128), 128, 0, c->stream>>>(c->g, (float2 *)psi->p[0], (const int *)lock_ns->p[0], count);
    else k_phase_lock<double2><<<svl_nblocks(count, 128), 128, 0, c->stream>>>(c->g, (double2 *)psi->p[0], (const int *)lock_ns->p[0], count);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}
