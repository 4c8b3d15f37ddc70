// Geometric-multigrid preconditioner for the CG solve of the hex-lattice problem (opt-in through
// smfem_pcg_use_multigrid).  The Jacobi-PCG of solver.cu is the path the north star names and stays the default; its
// iteration count grows like ne (875 at 100^3, 1725 at 200^3).  Here:
//   * levels: ne -> ceil(ne/2) -> ... down to <= 4; a coarse mesh takes every other node of the finer one, plus the last
//     one when ne is odd (so the inflated geometry and the z = 0 / z = 1 planes are followed), and its operator is
//     RE-ASSEMBLED with the library's own path (smfem_assemble + smfem_surface_mass with the same E, nu, beta): no Galerkin
//     triple products, no new pattern code;
//   * Dirichlet rows: a coarse dof is constrained iff the fine dof at the same node is (injection);
//   * smoother: Chebyshev in D^-1 A on [lmax/8, lmax] (lmax from a few power iterations), 2 steps before and after the
//     coarse correction; Jacobi with a fixed damping needed 68 iterations where Chebyshev(2) needs 17 (ne = 16, CPU prototype);
//   * transfers: trilinear interpolation P per displacement component, restriction P';
//   * coarsest level: 30 Chebyshev steps (the iteration count does not react to a better coarse solve);
//   * outer iteration: standard PCG, dot products by a fixed-grid two-stage reduction (bit-reproducible).
// The V-cycle is a fixed linear operator (same polynomial every time), symmetric in the sense PCG needs.
//
// Several GPUs (one process per GPU, z-slabs): the fine levels are DISTRIBUTED - coarse node plane K belongs to the rank
// that owns fine plane min(2K, ne_f), so restriction and interpolation need one ghost plane per side, like the SpMV.  Every
// vector that is read with ghost planes (smoother iterate x, residual, search direction p) lives in the rank's peer window;
// an exchange is  push (direct NVLink stores into the neighbours' ghost planes + release flag)  ->  wait  ->  consumer
// kernel  ->  acknowledge (the next push into that neighbour waits for it: the smoother has no all-reduce that would order
// the ranks).  Levels with ne <= 32 are REPLICATED: the restricted residual is all-gathered (stores into every peer's
// buffer) and each rank runs the small one-GPU hierarchy on the whole coarse problem, then interpolates its own slab - no
// scatter, and the existing one-GPU code is the coarse solver.  Dot products: local two-stage reduction + mailbox
// all-reduce, summed in rank order (bitwise identical on all ranks).  No host or NCCL call on the data path.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "smfem_internal.cuh"

namespace {

constexpr int NT = 256;
constexpr int REPLICATE_NE_DEFAULT = 32;  // levels with ne <= this run replicated on every rank (multi-GPU)
inline int replicate_ne() {  // env SMFEM_GMG_REPLICATE_NE: experiments / tests (must be the same on every rank)
    const char *e = std::getenv("SMFEM_GMG_REPLICATE_NE");
    return e ? std::atoi(e) : REPLICATE_NE_DEFAULT;
}

// ------------------------------------------------------------------------------------------------
// level plan: sizes and slabs of every level for every rank, from ne and the rank count alone
// ------------------------------------------------------------------------------------------------
struct GmgPlan {
    int nranks = 1;
    int ndist = 0;                 // distributed levels 0 .. ndist-1 (all levels when nranks == 1)
    int ne[16] = {};
    int k0[16][SMFEM_MAX_RANKS] = {}, k1[16][SMFEM_MAX_RANKS] = {};
    int ne_rep = 0;                // > 0: ne of the first replicated level
};

inline int fine_of_h(int I, int ne_f) { return 2 * I < ne_f ? 2 * I : ne_f; }

GmgPlan gmg_plan(int ne0, int nranks) {
    GmgPlan P;
    P.nranks = nranks;
    P.ne[0] = ne0;
    for (int r = 0; r < nranks; ++r) slab_range(ne0 + 1, r, nranks, P.k0[0][r], P.k1[0][r]);
    P.ndist = 1;
    int ne = ne0;
    while (ne > 4 && P.ndist < 12) {
        const int nec = (ne + 1) / 2, l = P.ndist;
        if (nranks > 1) {
            bool ok = nec > replicate_ne();
            int k0c[SMFEM_MAX_RANKS], k1c[SMFEM_MAX_RANKS];
            for (int r = 0; r < nranks && ok; ++r) {
                // coarse planes K with k0f <= fine_of(K) < k1f
                int a = 0, b = 0;
                while (a <= nec && fine_of_h(a, ne) < P.k0[l - 1][r]) ++a;
                b = a;
                while (b <= nec && fine_of_h(b, ne) < P.k1[l - 1][r]) ++b;
                k0c[r] = a;
                k1c[r] = b;
                if (b - a < 2) ok = false;
            }
            if (!ok) {
                P.ne_rep = nec;
                break;
            }
            for (int r = 0; r < nranks; ++r) {
                P.k0[l][r] = k0c[r];
                P.k1[l][r] = k1c[r];
            }
        } else {
            P.k0[l][0] = 0;
            P.k1[l][0] = nec + 1;
        }
        P.ne[l] = nec;
        P.ndist++;
        ne = nec;
    }
    return P;
}

inline int64_t plan_ncols(const GmgPlan &P, int l, int r) {
    const int64_t n1 = P.ne[l] + 1;
    return 3 * n1 * n1 * (P.k1[l][r] - P.k0[l][r] + 2);
}
// offsets (in doubles, from the start of the hierarchy's region = just behind p) of rank r's exchanged vectors
inline int64_t plan_x_off(const GmgPlan &P, int l, int r) {
    int64_t o = 0;
    for (int i = 0; i < l; ++i) o += 2 * plan_ncols(P, i, r);
    return o;
}
inline int64_t plan_res_off(const GmgPlan &P, int l, int r) { return plan_x_off(P, l, r) + plan_ncols(P, l, r); }
inline int64_t plan_gather_off(const GmgPlan &P, int r) { return plan_x_off(P, P.ndist, r); }
inline int64_t plan_gather_doubles(const GmgPlan &P) {  // the whole coarse lattice in ghost layout (n1 + 2 planes)
    const int64_t n1 = P.ne_rep + 1;
    return P.ne_rep > 0 ? 3 * n1 * n1 * (n1 + 2) : 0;
}

struct GmgLevel {
    smfem_mesh *mesh = nullptr;  // owned for levels > 0
    smfem_matrix *K = nullptr;   // owned for levels > 0
    Lattice L;
    int64_t nrows = 0, ncols = 0, ghost = 0;
    double *x = nullptr;    // ncols (ghost planes included: SpMV input)
    double *res = nullptr;  // ncols: masked residual with ghost planes (restriction input)
    bool own_vec = true;    // x / res from the allocator (one GPU) or inside the peer window (several)
    double *b = nullptr, *d = nullptr, *y = nullptr, *dinv = nullptr;  // nrows
    uint8_t *fixed = nullptr;  // nrows (level 0: the matrix's own mask, may be null)
    bool own_fixed = false;
    double lmax = 0;
    // where my first / last owned plane of x and res go: the lower neighbour's ghost_hi / the upper neighbour's ghost_lo
    double *x_lo = nullptr, *x_hi = nullptr, *res_lo = nullptr, *res_hi = nullptr;
};

struct Gmg {
    std::vector<GmgLevel> lev;
    GmgPlan plan;
    double *p = nullptr;   // ncols of level 0 (several GPUs: the matrix's own p inside the peer window)
    bool own_p = true;
    double *p_lo = nullptr, *p_hi = nullptr;
    double *r = nullptr;   // = lev[0].b (not owned): the residual is the V-cycle's right-hand side
    double *xs = nullptr;  // nrows of level 0
    double *partials = nullptr, *h_pinned = nullptr;
    int red_grid = 0;
    // communication (level-0 matrix's peer window); nranks == 1: unused
    CommView cv;
    PcgScalars *scal = nullptr;
    // replicated coarse hierarchy (several GPUs): a one-GPU Gmg on a shadow context over the whole coarse lattice
    Gmg *rep = nullptr;
    smfem_ctx *ctx1 = nullptr;
    smfem_mesh *rep_mesh = nullptr;
    smfem_matrix *rep_K = nullptr;
    Lattice rep_L;                     // my slab of the first replicated level (planes I contribute to the gather)
    double *gather = nullptr;          // whole coarse vector, inside my window
    double *gather_peer[SMFEM_MAX_RANKS] = {nullptr};
    double *rep_slab = nullptr;        // my planes of the coarse right-hand side / coordinates / mask before the gather
    double *rep_b_saved = nullptr;     // the replicated hierarchy's own level-0 b (restored before it is freed)
};

// Coarsening along one axis: ne_c = ceil(ne_f / 2); coarse node I sits on fine node min(2 I, ne_f) (for odd ne_f the last
// coarse element is a single fine element).  pw = weight of coarse node I in the linear interpolation at fine node i.
__device__ __forceinline__ int fine_of(int I, int ne_f) { return min(2 * I, ne_f); }
__device__ __forceinline__ double pw(int i, int I, int ne_f) {
    if (i == ne_f && (ne_f & 1)) return I == (ne_f + 1) / 2 ? 1.0 : 0.0;
    if (!(i & 1)) return I == (i >> 1) ? 1.0 : 0.0;
    return (I == (i >> 1) || I == (i >> 1) + 1) ? 0.5 : 0.0;
}

// owned node t of lattice L -> (i, j, k)
__device__ __forceinline__ void owned_ijk(const Lattice &L, int64_t m, int &i, int &j, int &k) {
    i = (int)(m % L.n1);
    j = (int)((m / L.n1) % L.n1);
    k = L.k0 + (int)(m / ((int64_t)L.n1 * L.n1));
}

__global__ void k_subsample_coords(Lattice Lc, Lattice Lf, const double *__restrict__ cf, double *__restrict__ cc) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)Lc.nown() * Lc.plane()) return;
    int i, j, k;
    owned_ijk(Lc, t, i, j, k);
    const int64_t nc = Lc.lnode(i, j, k), nf = Lf.lnode(fine_of(i, Lf.ne), fine_of(j, Lf.ne), fine_of(k, Lf.ne));
#pragma unroll
    for (int c = 0; c < 3; ++c) cc[3 * nc + c] = cf[3 * nf + c];
}

__global__ void k_inject_fixed(Lattice Lc, Lattice Lf, const uint8_t *__restrict__ ff, uint8_t *__restrict__ fc) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * (int64_t)Lc.nown() * Lc.plane()) return;
    const int c = (int)(t % 3);
    int i, j, k;
    owned_ijk(Lc, t / 3, i, j, k);
    const int64_t rf = 3 * (Lf.lnode(fine_of(i, Lf.ne), fine_of(j, Lf.ne), fine_of(k, Lf.ne)) - Lf.plane()) + c;  // owned row of the fine level
    fc[t] = ff ? ff[rf] : 0;
}

__global__ void k_dinv(int64_t n, const double *__restrict__ diag, const uint8_t *__restrict__ fixed, double *__restrict__ dinv) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dinv[t] = (fixed && fixed[t]) ? 0.0 : 1.0 / diag[t];
}

// b = (extra - K q_d) on the free rows, 0 on the constrained ones
__global__ void k_rhs(int64_t n, const double *__restrict__ y, const double *__restrict__ extra, const uint8_t *__restrict__ fixed,
                      double *__restrict__ r) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) r[t] = (fixed && fixed[t]) ? 0.0 : (extra ? extra[t] : 0.0) - (y ? y[t] : 0.0);
}

// Chebyshev from a zero guess: d = D^-1 b / theta, x = d
__global__ void k_cheb_first(int64_t n, int64_t ghost, const double *__restrict__ b, const double *__restrict__ dinv, double inv_theta,
                             double *__restrict__ x, double *__restrict__ d) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double v = dinv[t] * b[t] * inv_theta;
    d[t] = v;
    x[t + ghost] = v;
}

// r = D^-1 (b - A x) with y = A x;  d = c1 d + c2 r;  x += d
__global__ void k_cheb_step(int64_t n, int64_t ghost, const double *__restrict__ b, const double *__restrict__ y,
                            const double *__restrict__ dinv, double c1, double c2, double *__restrict__ x, double *__restrict__ d) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double r = dinv[t] * (b[t] - y[t]);
    const double v = (c1 != 0.0 ? c1 * d[t] : 0.0) + c2 * r;
    d[t] = v;
    x[t + ghost] += v;
}

// masked residual with room for ghost planes: res = b - A x on the free rows, 0 on the constrained ones
__global__ void k_residual(int64_t n, int64_t ghost, const double *__restrict__ b, const double *__restrict__ y,
                           const uint8_t *__restrict__ fixed, double *__restrict__ res) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) res[t + ghost] = (fixed && fixed[t]) ? 0.0 : b[t] - y[t];
}

// coarse rhs = P' res on the free coarse rows (owned coarse planes; fine planes fk-1 .. fk+1 incl. one ghost plane per side)
__global__ void k_restrict(Lattice Lc, Lattice Lf, const double *__restrict__ resf, const uint8_t *__restrict__ fixed_c, double *__restrict__ bc) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * (int64_t)Lc.nown() * Lc.plane()) return;
    if (fixed_c && fixed_c[t]) {
        bc[t] = 0.0;
        return;
    }
    const int c = (int)(t % 3);
    int I, J, Kz;
    owned_ijk(Lc, t / 3, I, J, Kz);
    const int nef = Lf.ne, n1f = Lf.n1;
    const int fi = fine_of(I, nef), fj = fine_of(J, nef), fk = fine_of(Kz, nef);
    double s = 0.0;
    for (int dk = -1; dk <= 1; ++dk) {
        const int k = fk + dk;
        if (k < 0 || k >= n1f) continue;
        const double wk = pw(k, Kz, nef);
        if (wk == 0.0) continue;
        for (int dj = -1; dj <= 1; ++dj) {
            const int j = fj + dj;
            if (j < 0 || j >= n1f) continue;
            const double wj = pw(j, J, nef);
            if (wj == 0.0) continue;
            for (int di = -1; di <= 1; ++di) {
                const int i = fi + di;
                if (i < 0 || i >= n1f) continue;
                const double wi = pw(i, I, nef);
                if (wi == 0.0) continue;
                s += wi * wj * wk * resf[3 * Lf.lnode(i, j, k) + c];
            }
        }
    }
    bc[t] = s;
}

// x_f += P x_c on the free fine rows (xc with ghost planes; Lc may be the whole replicated coarse lattice)
__global__ void k_prolong_add(Lattice Lf, Lattice Lc, const double *__restrict__ xc, const uint8_t *__restrict__ fixed_f, double *__restrict__ xf) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * (int64_t)Lf.nown() * Lf.plane()) return;
    if (fixed_f && fixed_f[t]) return;
    const int c = (int)(t % 3);
    int i, j, k;
    owned_ijk(Lf, t / 3, i, j, k);
    const int nef = Lf.ne;
    // parents along each axis: first parent a0 and whether there is a second one (weights 1 or 1/2 + 1/2)
    const bool li = (i == nef) && (nef & 1), lj = (j == nef) && (nef & 1), lk = (k == nef) && (nef & 1);
    const int i0 = li ? (nef + 1) / 2 : i >> 1, j0 = lj ? (nef + 1) / 2 : j >> 1, k0 = lk ? (nef + 1) / 2 : k >> 1;
    const int oi = (!li) && (i & 1), oj = (!lj) && (j & 1), ok = (!lk) && (k & 1);
    double s = 0.0;
    for (int dk = 0; dk <= ok; ++dk)
        for (int dj = 0; dj <= oj; ++dj)
            for (int di = 0; di <= oi; ++di) s += xc[3 * Lc.lnode(i0 + di, j0 + dj, k0 + dk) + c];
    xf[t + 3 * Lf.plane()] += s * (oi ? 0.5 : 1.0) * (oj ? 0.5 : 1.0) * (ok ? 0.5 : 1.0);
}

// fixed-grid dot product, stage 1 (stage 2: one block adds the partials in index order)
__global__ void __launch_bounds__(NT) k_dot1(int64_t n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ partials) {
    __shared__ double sh[NT];
    double s = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * NT + threadIdx.x; t < n; t += (int64_t)gridDim.x * NT) s += a[t] * b[t];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = NT / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(NT) k_dot2(int np, const double *__restrict__ partials, double *__restrict__ out) {
    __shared__ double sh[NT];
    double s = 0.0;
    for (int t = threadIdx.x; t < np; t += NT) s += partials[t];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = NT / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

// p = z + beta p   (both with ghost offset)
__global__ void k_update_p(int64_t n, int64_t ghost, const double *__restrict__ z, double beta, double *__restrict__ p) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t + ghost] = z[t + ghost] + beta * p[t + ghost];
}
// x += alpha p;  r -= alpha Ap (constrained rows stay 0)
__global__ void k_update_xr(int64_t n, int64_t ghost, double alpha, const double *__restrict__ p, const double *__restrict__ Ap,
                            const uint8_t *__restrict__ fixed, double *__restrict__ x, double *__restrict__ r) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (fixed && fixed[t]) return;
    x[t] += alpha * p[t + ghost];
    r[t] -= alpha * Ap[t];
}
// warm start (load stepping, examples/vector3D.jl:310-338: q is linear in d): x = w * x_prev, also as SpMV input
__global__ void k_warm(int64_t n, int64_t ghost, double w, const double *__restrict__ prev, double *__restrict__ x, double *__restrict__ p) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double v = w * prev[t];
    x[t] = v;
    p[t + ghost] = v;
}
__global__ void k_sub_masked(int64_t n, const double *__restrict__ y, const uint8_t *__restrict__ fixed, double *__restrict__ r) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && !(fixed && fixed[t])) r[t] -= y[t];
}
__global__ void k_scale_copy(int64_t n, int64_t ghost, double s, const double *__restrict__ y, const double *__restrict__ dinv,
                             double *__restrict__ x) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) x[t + ghost] = s * dinv[t] * y[t];
}
__global__ void k_final(int64_t n, int64_t ghost, const double *__restrict__ qd, const double *__restrict__ x,
                        const uint8_t *__restrict__ fixed, double *__restrict__ q) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) q[t] = (fixed && fixed[t]) ? (qd ? qd[t + ghost] : 0.0) : x[t] + 0.0;
}
// positive, non-smooth start vector of the power iteration; a function of the GLOBAL row so that it does not depend on the partition
__global__ void k_fill_test(int64_t n, int64_t ghost, int64_t row0, const double *__restrict__ dinv, double *__restrict__ x) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) x[t + ghost] = dinv[t] != 0.0 ? 1.0 + 0.37 * (double)((((unsigned long long)(t + row0)) * 2654435761ull) % 1000) / 1000.0 : 0.0;
}
__global__ void k_copy(int64_t n, const double *__restrict__ a, double *__restrict__ b) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) b[t] = a[t];
}
__global__ void k_u8_to_f64(int64_t n, const uint8_t *__restrict__ a, double *__restrict__ b) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) b[t] = a ? (double)a[t] : 0.0;
}
__global__ void k_f64_to_u8(int64_t n, const double *__restrict__ a, uint8_t *__restrict__ b) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) b[t] = a[t] != 0.0;
}

// ------------------------------------------------------------------------------------------------
// peer-memory exchanges (several GPUs)
// ------------------------------------------------------------------------------------------------
// Spin on a flag of the peer protocol; a peer that never answers (its host side threw, or the ranks took different paths)
// must surface as a launch failure, not as a hang: ~2^26 system-scope loads are tens of seconds.
#define SPIN_UNTIL(cond)                            \
    do {                                            \
        unsigned spins_ = 0;                        \
        while (!(cond)) {                           \
            if (++spins_ > (1u << 26)) __trap();    \
        }                                           \
    } while (0)

// halo push number s = hseq + 1: my first owned plane -> the lower neighbour's ghost_hi, my last owned plane -> the upper
// neighbour's ghost_lo.  Waits until both neighbours have consumed push s - 1 (acknowledgement flags), because nothing
// else orders the ranks inside a smoother.
__global__ void __launch_bounds__(NT) k_ghalo_push(int64_t plane_dofs, const double *__restrict__ first, const double *__restrict__ last,
                                                   double *__restrict__ dst_lo, double *__restrict__ dst_hi, PcgScalars *scal, CommView cv) {
    __shared__ bool s_last;
    const unsigned long long seq = scal->hseq + 1;
    const bool lo = cv.rank > 0, hi = cv.rank < cv.nranks - 1;
    if (threadIdx.x == 0) {
        if (lo) SPIN_UNTIL(ld_acquire_sys(&cv.self->gaflag[0]) + 1 >= seq);
        if (hi) SPIN_UNTIL(ld_acquire_sys(&cv.self->gaflag[1]) + 1 >= seq);
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane_dofs; i += stride) {
        if (lo) dst_lo[i] = first[i];
        if (hi) dst_hi[i] = last[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(&scal->ticketA, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        scal->ticketA = 0;
        __threadfence_system();
        if (lo) st_release_sys(&cv.peer[cv.rank - 1]->ghflag[1], seq);
        if (hi) st_release_sys(&cv.peer[cv.rank + 1]->ghflag[0], seq);
        scal->hseq = seq;
    }
}
// ... the consumer's side: my ghost planes hold the neighbours' push number hseq
__global__ void k_ghalo_wait(PcgScalars *scal, CommView cv) {
    const unsigned long long seq = scal->hseq;
    if (cv.rank > 0) SPIN_UNTIL(ld_acquire_sys(&cv.self->ghflag[0]) >= seq);
    if (cv.rank < cv.nranks - 1) SPIN_UNTIL(ld_acquire_sys(&cv.self->ghflag[1]) >= seq);
}
// ... and after the consumer kernel: the neighbours may overwrite my ghost planes
__global__ void k_ghalo_ack(PcgScalars *scal, CommView cv) {
    const unsigned long long seq = scal->hseq;
    __threadfence_system();
    if (cv.rank > 0) st_release_sys(&cv.peer[cv.rank - 1]->gaflag[1], seq);            // I am their upper neighbour
    if (cv.rank < cv.nranks - 1) st_release_sys(&cv.peer[cv.rank + 1]->gaflag[0], seq);  // I am their lower neighbour
}

// all-gather number s = gseq + 1: my n doubles -> offset `off` of EVERY rank's gather buffer
struct GatherPeers {
    double *buf[SMFEM_MAX_RANKS];
};
__global__ void __launch_bounds__(NT) k_gather_push(int64_t n, const double *__restrict__ src, int64_t off, GatherPeers gp, PcgScalars *scal, CommView cv) {
    __shared__ bool s_last;
    const unsigned long long seq = scal->gseq + 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = src[i];
        for (int q = 0; q < cv.nranks; ++q) gp.buf[q][off + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(&scal->ticketA, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        scal->ticketA = 0;
        __threadfence_system();
        for (int q = 0; q < cv.nranks; ++q) st_release_sys(&cv.peer[q]->gathflag[cv.rank], seq);
        scal->gseq = seq;
    }
}
__global__ void k_gather_wait(PcgScalars *scal, CommView cv) {
    const unsigned long long seq = scal->gseq;
    for (int q = 0; q < cv.nranks; ++q) SPIN_UNTIL(ld_acquire_sys(&cv.self->gathflag[q]) >= seq);
}

// all-reduce number s = rseq + 1 of one double (in place): publish to every rank's mailbox slot s & 3, add in rank order
__global__ void k_gallreduce(double *val, PcgScalars *scal, CommView cv) {
    const unsigned long long seq = scal->rseq + 1;
    const int slot = (int)(seq & 3ull);
    const double v = *val;
    for (int q = 0; q < cv.nranks; ++q) cv.peer[q]->gbox[slot][cv.rank][0] = v;
    __threadfence_system();
    for (int q = 0; q < cv.nranks; ++q) st_release_sys(&cv.peer[q]->gflag[slot][cv.rank], seq);
    double s = 0.0;
    for (int q = 0; q < cv.nranks; ++q) {
        SPIN_UNTIL(ld_acquire_sys(&cv.self->gflag[slot][q]) == seq);
        s += ld_volatile_f64(&cv.self->gbox[slot][q][0]);
    }
    *val = s;
    scal->rseq = seq;
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + NT - 1) / NT); }
inline bool multi(const Gmg *G) { return G->plan.nranks > 1; }

// exchange the ghost planes of a vector with ghost layout (call halo_done after the kernel that read them)
void halo(smfem_ctx *ctx, Gmg *G, const Lattice &L, const double *v, double *dst_lo, double *dst_hi) {
    if (!multi(G)) return;
    const int64_t pd = 3 * L.plane();
    int g = (int)((pd + NT - 1) / NT);
    if (g > ctx->sms * 2) g = ctx->sms * 2;
    LAUNCH(ctx, k_ghalo_push, g, NT, 0, pd, v + pd, v + (int64_t)L.nown() * pd, dst_lo, dst_hi, G->scal, G->cv);
    LAUNCH(ctx, k_ghalo_wait, 1, 1, 0, G->scal, G->cv);
}
void halo_done(smfem_ctx *ctx, Gmg *G) {
    if (!multi(G)) return;
    LAUNCH(ctx, k_ghalo_ack, 1, 1, 0, G->scal, G->cv);
}

// y = A x on a distributed level (x in the peer window when several GPUs are used)
void spmv_level(smfem_ctx *ctx, Gmg *G, GmgLevel &V, const double *x, double *x_lo, double *x_hi, double *y) {
    halo(ctx, G, V.L, x, x_lo, x_hi);
    spmv_device(ctx, V.K, x, y);
    halo_done(ctx, G);
}

// my planes of a vector of the first replicated level -> every rank's whole-lattice buffer
void gather_rep(smfem_ctx *ctx, Gmg *G, const double *slab) {
    const int64_t n = 3 * (int64_t)G->rep_L.nown() * G->rep_L.plane();
    const int64_t off = 3 * (int64_t)(G->rep_L.k0 + 1) * G->rep_L.plane();  // ghost layout of the whole lattice: plane k at k + 1
    GatherPeers gp;
    for (int q = 0; q < SMFEM_MAX_RANKS; ++q) gp.buf[q] = G->gather_peer[q];
    int g = (int)grid_for(n);
    if (g > ctx->sms * 2) g = ctx->sms * 2;
    if (g < 1) g = 1;
    LAUNCH(ctx, k_gather_push, g, NT, 0, n, slab, off, gp, G->scal, G->cv);
    LAUNCH(ctx, k_gather_wait, 1, 1, 0, G->scal, G->cv);
}

double dot(smfem_ctx *ctx, Gmg *G, int64_t n, const double *a, const double *b) {
    LAUNCH(ctx, k_dot1, G->red_grid, NT, 0, n, a, b, G->partials);
    LAUNCH(ctx, k_dot2, 1, NT, 0, G->red_grid, (const double *)G->partials, G->partials + G->red_grid);
    if (multi(G)) LAUNCH(ctx, k_gallreduce, 1, 1, 0, G->partials + G->red_grid, G->scal, G->cv);
    CUDA_CHECK(cudaMemcpyAsync(G->h_pinned, G->partials + G->red_grid, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return G->h_pinned[0];
}

// n Chebyshev steps on level l for A x = b (x_zero: the guess is 0 and x need not be read)
void smooth(smfem_ctx *ctx, Gmg *G, GmgLevel &V, int n, bool x_zero) {
    const double lmax = V.lmax, lmin = lmax / 8.0;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    const unsigned g = grid_for(V.nrows);
    for (int k = 0; k < n; ++k) {
        if (k == 0 && x_zero) {
            LAUNCH(ctx, k_cheb_first, g, NT, 0, V.nrows, V.ghost, (const double *)V.b, (const double *)V.dinv, 1.0 / theta, V.x, V.d);
            continue;
        }
        spmv_level(ctx, G, V, V.x, V.x_lo, V.x_hi, V.y);
        if (k == 0) {
            LAUNCH(ctx, k_cheb_step, g, NT, 0, V.nrows, V.ghost, (const double *)V.b, (const double *)V.y, (const double *)V.dinv, 0.0,
                   1.0 / theta, V.x, V.d);
        } else {
            const double rho_n = 1.0 / (2.0 * sigma - rho);
            LAUNCH(ctx, k_cheb_step, g, NT, 0, V.nrows, V.ghost, (const double *)V.b, (const double *)V.y, (const double *)V.dinv,
                   rho_n * rho, 2.0 * rho_n / delta, V.x, V.d);
            rho = rho_n;
        }
    }
}

bool g_in_rep = false;  // inside the replicated hierarchy's cycle (the debugging depth limit counts top-level levels only)
inline int debug_depth() {  // env SMFEM_GMG_DEBUG_DEPTH = d > 0: level d - 1 is treated as the coarsest (debugging)
    const char *e = std::getenv("SMFEM_GMG_DEBUG_DEPTH");
    return e ? std::atoi(e) : 0;
}

double dot(smfem_ctx *ctx, Gmg *G, int64_t n, const double *a, const double *b);
bool g_trace = false;  // gmg_apply_host with env SMFEM_GMG_TRACE: global norms after every stage (never inside a graph capture)
void trace(smfem_ctx *ctx, Gmg *G, int l, const char *what, int64_t n, const double *v) {
    if (!g_trace || g_in_rep) return;
    const double s = dot(ctx, G, n, v, v);
    if (G->cv.rank == 0) std::fprintf(stderr, "[gmg trace nranks=%d] level %d %-28s ||.||^2 = %.17g\n", G->plan.nranks, l, what, s);
}

void vcycle(smfem_ctx *ctx, Gmg *G, int l) {
    GmgLevel &V = G->lev[l];
    const int nl = (int)G->lev.size();
    trace(ctx, G, l, "rhs b", V.nrows, V.b);
    if (g_trace && !g_in_rep && G->cv.rank == 0) std::fprintf(stderr, "[gmg trace nranks=%d] level %d lmax = %.17g nrows(local) = %lld\n", G->plan.nranks, l, V.lmax, (long long)V.nrows);
    trace(ctx, G, l, "dinv", V.nrows, V.dinv);
    const bool last = (l == nl - 1);
    const int dd = g_in_rep ? 0 : debug_depth();
    if ((last && !G->rep) || (dd > 0 && l == dd - 1)) {
        smooth(ctx, G, V, nl == 1 ? 2 : 30, true);
        trace(ctx, G, l, "x after coarsest smoothing", V.nrows, V.x + V.ghost);
        return;
    }
    smooth(ctx, G, V, 2, true);
    trace(ctx, G, l, "x after pre-smoothing", V.nrows, V.x + V.ghost);
    spmv_level(ctx, G, V, V.x, V.x_lo, V.x_hi, V.y);
    trace(ctx, G, l, "y = A x", V.nrows, V.y);
    LAUNCH(ctx, k_residual, grid_for(V.nrows), NT, 0, V.nrows, V.ghost, (const double *)V.b, (const double *)V.y, (const uint8_t *)V.fixed, V.res);
    halo(ctx, G, V.L, V.res, V.res_lo, V.res_hi);
    if (!last) {
        GmgLevel &C = G->lev[l + 1];
        LAUNCH(ctx, k_restrict, grid_for(C.nrows), NT, 0, C.L, V.L, (const double *)V.res, (const uint8_t *)C.fixed, C.b);
        halo_done(ctx, G);
        vcycle(ctx, G, l + 1);
        halo(ctx, G, C.L, C.x, C.x_lo, C.x_hi);
        LAUNCH(ctx, k_prolong_add, grid_for(V.nrows), NT, 0, V.L, C.L, (const double *)C.x, (const uint8_t *)V.fixed, V.x);
        halo_done(ctx, G);
        trace(ctx, G, l, "x after prolongation", V.nrows, V.x + V.ghost);
    } else {
        // first replicated level: my planes of its right-hand side -> all-gather -> every rank solves the whole coarse problem
        Gmg *R = G->rep;
        GmgLevel &C = R->lev[0];
        const int64_t nslab = 3 * (int64_t)G->rep_L.nown() * G->rep_L.plane();
        // rows of my slab inside the whole-lattice mask
        const uint8_t *fx = C.fixed ? C.fixed + 3 * (int64_t)G->rep_L.k0 * G->rep_L.plane() : nullptr;
        LAUNCH(ctx, k_restrict, grid_for(nslab), NT, 0, G->rep_L, V.L, (const double *)V.res, fx, G->rep_slab);
        halo_done(ctx, G);
        gather_rep(ctx, G, G->rep_slab);  // -> C.b (= gather buffer + one ghost plane)
        g_in_rep = true;
        vcycle(G->ctx1, R, 0);
        g_in_rep = false;
        LAUNCH(ctx, k_prolong_add, grid_for(V.nrows), NT, 0, V.L, C.L, (const double *)C.x, (const uint8_t *)V.fixed, V.x);
    }
    smooth(ctx, G, V, 2, false);
}

void level_free(GmgLevel &V, bool owned) {
    if (V.own_vec) {
        dev_free(V.x);
        dev_free(V.res);
    }
    dev_free(V.b);
    dev_free(V.d);
    dev_free(V.y);
    dev_free(V.dinv);
    if (V.own_fixed) dev_free(V.fixed);
    if (owned) {
        if (V.K) smfem_matrix_free(V.K);
        if (V.mesh) smfem_mesh_free(V.mesh);
    }
}

void gmg_destroy(Gmg *G) {
    if (!G) return;
    if (G->rep) {
        if (G->rep_b_saved) G->rep->lev[0].b = G->rep_b_saved;  // it was pointed into the peer window
        gmg_destroy(G->rep);
        if (G->rep_K) smfem_matrix_free(G->rep_K);
        if (G->rep_mesh) smfem_mesh_free(G->rep_mesh);
    }
    if (G->ctx1) {
        G->ctx1->host_pool = nullptr;
        delete G->ctx1;
    }
    dev_free(G->rep_slab);
    for (size_t l = 0; l < G->lev.size(); ++l) level_free(G->lev[l], l > 0);
    if (G->own_p) dev_free(G->p);
    dev_free(G->xs);
    dev_free(G->partials);
    if (G->h_pinned) cudaFreeHost(G->h_pinned);
    delete G;
}

#define ABI_CHECK(call)                                                                   \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != SMFEM_OK) throw SmfemError(rc_, std::string(smfem_last_error()));      \
    } while (0)

// coarse operator on `mesh` with K's material and surface term (the library's own assembly path)
smfem_matrix *assemble_coarse(smfem_ctx *ctx, smfem_mesh *mesh, const smfem_matrix *K) {
    smfem_matrix *C = nullptr;
    ABI_CHECK(smfem_assemble(ctx, mesh, mesh->ne, 3, SMFEM_Q1, 3, K->Young, K->nu, &C));
    C->gmg_coarse = true;
    if (K->beta_total != 0.0) ABI_CHECK(smfem_surface_mass(ctx, C, mesh, nullptr, nullptr, 0, K->beta_total, 0));
    solver_alloc(ctx, C);
    return C;
}

Gmg *gmg_build(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh) {
    Gmg *G = new Gmg();
    try {
        G->plan = gmg_plan(K->lat.ne, ctx->nranks);
        const GmgPlan &P = G->plan;
        const int rank = ctx->rank;
        G->red_grid = ctx->sms * 4;
        G->partials = dev_alloc<double>(G->red_grid + 8);
        CUDA_CHECK(cudaMallocHost(&G->h_pinned, 64));
        double *region = nullptr;  // several GPUs: the hierarchy's part of my peer window (behind p)
        double *peer_region[SMFEM_MAX_RANKS] = {nullptr};
        if (multi(G)) {
            REQUIRE(K->comm_connected && K->gmg_region_doubles > 0, SMFEM_ERR_INVALID,
                    "multi-GPU multigrid: call smfem_comm_connect on K first");
            G->cv = K->comm;
            G->scal = K->scal;
            region = K->p + K->ncols_l;
            for (int q = 0; q < ctx->nranks; ++q) peer_region[q] = K->comm.peer_p[q] + plan_ncols(P, 0, q);
            // (the window was zeroed when it was allocated; it must NOT be cleared here: a faster peer may already be writing)
        }
        // ---- distributed levels (all levels on one GPU)
        GmgLevel V0;
        V0.mesh = mesh;
        V0.K = K;
        G->lev.push_back(V0);
        for (int l = 0; l < P.ndist; ++l) {
            if (l > 0) {
                GmgLevel Vn;
                G->lev.push_back(Vn);
            }
            GmgLevel &V = G->lev[l];
            if (l > 0) V.mesh = mesh_new_lattice(ctx, P.ne[l], P.k0[l][rank], P.k1[l][rank]);  // owned from here on (freed by gmg_destroy)
            V.L = V.mesh->lat;
            V.ghost = 3 * V.L.plane();
            V.nrows = 3 * (int64_t)V.L.nown() * V.L.plane();
            V.ncols = V.nrows + 2 * V.ghost;
            if (multi(G)) {
                V.own_vec = false;
                V.x = region + plan_x_off(P, l, rank);
                V.res = region + plan_res_off(P, l, rank);
                if (rank > 0) {
                    const int64_t hi_off = (int64_t)(P.k1[l][rank - 1] - P.k0[l][rank - 1] + 1) * V.ghost;  // their ghost_hi plane
                    V.x_lo = peer_region[rank - 1] + plan_x_off(P, l, rank - 1) + hi_off;
                    V.res_lo = peer_region[rank - 1] + plan_res_off(P, l, rank - 1) + hi_off;
                }
                if (rank < ctx->nranks - 1) {
                    V.x_hi = peer_region[rank + 1] + plan_x_off(P, l, rank + 1);
                    V.res_hi = peer_region[rank + 1] + plan_res_off(P, l, rank + 1);
                }
            } else {
                V.x = dev_alloc<double>(V.ncols);
                V.res = dev_alloc<double>(V.ncols);
                CUDA_CHECK(cudaMemsetAsync(V.x, 0, 8 * V.ncols, ctx->stream));
                CUDA_CHECK(cudaMemsetAsync(V.res, 0, 8 * V.ncols, ctx->stream));
            }
            if (l > 0) {
                const GmgLevel &F = G->lev[l - 1];
                CUDA_CHECK(cudaMemsetAsync(V.mesh->coords, 0, 8 * 3 * V.mesh->nNodes_l, ctx->stream));
                LAUNCH(ctx, k_subsample_coords, grid_for((int64_t)V.L.nown() * V.L.plane()), NT, 0, V.L, F.L, (const double *)F.mesh->coords,
                       V.mesh->coords);
                if (multi(G)) {  // ghost planes of the coarse coordinates: through the level's x buffer (same layout: 3 per node)
                    // OWNED planes only: a faster neighbour may already have pushed into my ghost planes
                    LAUNCH(ctx, k_copy, grid_for(V.nrows), NT, 0, V.nrows, (const double *)(V.mesh->coords + V.ghost), V.x + V.ghost);
                    halo(ctx, G, V.L, V.x, V.x_lo, V.x_hi);
                    LAUNCH(ctx, k_copy, grid_for(V.ncols), NT, 0, V.ncols, (const double *)V.x, V.mesh->coords);
                    halo_done(ctx, G);
                }
                if (std::getenv("SMFEM_GMG_TRACE")) {  // per-plane checksums of the coarse coordinates (global plane index)
                    std::vector<double> h(3 * V.mesh->nNodes_l);
                    CUDA_CHECK(cudaMemcpyAsync(h.data(), V.mesh->coords, 8 * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
                    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
                    for (int kk = V.L.k0 - 1; kk <= V.L.k1; ++kk) {
                        if (kk < 0 || kk >= V.L.n1) continue;
                        if (!(kk <= V.L.k0 + 1 || kk >= V.L.k1 - 2 || (kk >= 19 && kk <= 22))) continue;
                        double sum = 0;
                        const int64_t o = 3 * (int64_t)(kk - V.L.k0 + 1) * V.L.plane();
                        for (int64_t t = 0; t < 3 * V.L.plane(); ++t) sum += h[o + t] * (1.0 + 1e-3 * (t % 97));
                        std::fprintf(stderr, "[gmg trace nranks=%d rank %d] level %d coords plane %d checksum %.15g\n", ctx->nranks, ctx->rank, l, kk, sum);
                    }
                }
                V.K = assemble_coarse(ctx, V.mesh, K);
            }
            REQUIRE(V.K->nrows_l == V.nrows && V.K->ghost_cols == V.ghost, SMFEM_ERR_INVALID, "multigrid: level layout mismatch");
            V.b = dev_alloc<double>(V.nrows);
            V.d = dev_alloc<double>(V.nrows);
            V.y = dev_alloc<double>(V.nrows);
            V.dinv = dev_alloc<double>(V.nrows);
            if (l > 0) {
                V.fixed = dev_alloc<uint8_t>(V.nrows);
                V.own_fixed = true;
            }
        }
        // ---- replicated coarse hierarchy (several GPUs): the whole lattice of level ndist on every rank
        if (multi(G) && P.ne_rep > 0) {
            const GmgLevel &F = G->lev.back();
            const int nec = P.ne_rep;
            int a = 0;
            while (a <= nec && fine_of_h(a, F.L.ne) < F.L.k0) ++a;
            int b = a;
            while (b <= nec && fine_of_h(b, F.L.ne) < F.L.k1) ++b;
            G->rep_L.ne = nec;
            G->rep_L.n1 = nec + 1;
            G->rep_L.k0 = a;
            G->rep_L.k1 = b;  // may be empty (b == a): this rank then contributes nothing to the gathers
            G->gather = region + plan_gather_off(P, rank);
            for (int q = 0; q < ctx->nranks; ++q) G->gather_peer[q] = peer_region[q] + plan_gather_off(P, q);
            const int64_t nslab = 3 * (int64_t)G->rep_L.nown() * G->rep_L.plane();
            G->rep_slab = dev_alloc<double>(nslab + nslab / 8 + 16);  // + room for the slab's byte mask (gmg_refresh)
            // shadow context: same device, stream and events, but a "one GPU" view of the lattice
            G->ctx1 = new smfem_ctx(*ctx);
            G->ctx1->rank = 0;
            G->ctx1->nranks = 1;
            G->rep_mesh = mesh_new_lattice(G->ctx1, nec, 0, nec + 1);
            // coordinates: my planes (subsampled from the finest replicated-parent level) -> gather -> whole lattice
            if (nslab > 0) {
                // k_subsample_coords writes with ghost layout of rep_L: use a scratch of nodes_local size, then take the owned part
                double *tmp = dev_alloc<double>(3 * G->rep_L.nodes_local());
                LAUNCH(ctx, k_subsample_coords, grid_for((int64_t)G->rep_L.nown() * G->rep_L.plane()), NT, 0, G->rep_L, F.L,
                       (const double *)F.mesh->coords, tmp);
                LAUNCH(ctx, k_copy, grid_for(nslab), NT, 0, nslab, (const double *)(tmp + 3 * G->rep_L.plane()), G->rep_slab);
                CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
                dev_free(tmp);
            }
            gather_rep(ctx, G, G->rep_slab);
            CUDA_CHECK(cudaMemcpyAsync(G->rep_mesh->coords, G->gather, 8 * 3 * G->rep_mesh->nNodes_l, cudaMemcpyDeviceToDevice, ctx->stream));
            G->rep_K = assemble_coarse(G->ctx1, G->rep_mesh, K);
            G->rep = gmg_build(G->ctx1, G->rep_K, G->rep_mesh);
            // its level-0 right-hand side IS the gather buffer (owned rows start one ghost plane in)
            G->rep_b_saved = G->rep->lev[0].b;
            G->rep->lev[0].b = G->gather + 3 * G->rep_L.plane();
            G->rep->lev[0].fixed = dev_alloc<uint8_t>(G->rep->lev[0].nrows);
            G->rep->lev[0].own_fixed = true;
        }
        const GmgLevel &V = G->lev[0];
        if (multi(G)) {
            G->p = K->p;  // the matrix's search-direction vector, already in the peer window with its ghost planes
            G->own_p = false;
            if (rank > 0) G->p_lo = K->comm.peer_p[rank - 1] + (int64_t)(P.k1[0][rank - 1] - P.k0[0][rank - 1] + 1) * V.ghost;
            if (rank < ctx->nranks - 1) G->p_hi = K->comm.peer_p[rank + 1];
        } else {
            G->p = dev_alloc<double>(V.ncols);
        }
        G->r = G->lev[0].b;  // the V-cycle's level-0 right-hand side
        G->xs = dev_alloc<double>(V.nrows);
    } catch (...) {
        gmg_destroy(G);
        throw;
    }
    return G;
}

// Chebyshev bound of every level, once per hierarchy: lmax(D^-1 A) of the UNCONSTRAINED operator (power iteration from a
// positive, non-smooth vector, 10 % margin).  Constraining dofs takes a principal submatrix of D^-1/2 A D^-1/2, whose
// largest eigenvalue cannot be larger, so the bound holds for every Dirichlet set.
void gmg_bounds(smfem_ctx *ctx, Gmg *G) {
    for (size_t l = 0; l < G->lev.size(); ++l) {
        GmgLevel &V = G->lev[l];
        LAUNCH(ctx, k_dinv, grid_for(V.nrows), NT, 0, V.nrows, (const double *)V.K->diag, (const uint8_t *)nullptr, V.dinv);
        LAUNCH(ctx, k_fill_test, grid_for(V.nrows), NT, 0, V.nrows, V.ghost, V.K->row0, (const double *)V.dinv, V.x);
        double lam = 1.0;
        for (int it = 0; it < 15; ++it) {
            spmv_level(ctx, G, V, V.x, V.x_lo, V.x_hi, V.y);
            // x <- D^-1 A x, normalised (||x|| = 1 after the first pass, so the norm is the eigenvalue estimate)
            LAUNCH(ctx, k_scale_copy, grid_for(V.nrows), NT, 0, V.nrows, V.ghost, 1.0, (const double *)V.y, (const double *)V.dinv, V.x);
            const double nrm = std::sqrt(dot(ctx, G, V.nrows, V.x + V.ghost, V.x + V.ghost));
            REQUIRE(nrm > 0.0 && nrm == nrm, SMFEM_ERR_SINGULAR, "multigrid setup: D^-1 A x vanished (matrix without values?)");
            lam = nrm;
            LAUNCH(ctx, k_scale_copy, grid_for(V.nrows), NT, 0, V.nrows, V.ghost, 1.0 / nrm, (const double *)V.y, (const double *)V.dinv, V.x);
        }
        V.lmax = 1.1 * lam;
        if (!multi(G)) CUDA_CHECK(cudaMemsetAsync(V.x, 0, 8 * V.ncols, ctx->stream));  // several GPUs: every use rewrites x (owned) and its ghost planes
    }
    if (G->rep) gmg_bounds(G->ctx1, G->rep);
}

// masks and D^-1 of every level for the CURRENT boundary conditions (no host synchronisation)
void gmg_refresh(smfem_ctx *ctx, Gmg *G, smfem_matrix *K, const uint8_t *fixed0, bool have_fixed0) {
    for (size_t l = 0; l < G->lev.size(); ++l) {
        GmgLevel &V = G->lev[l];
        if (l == 0) {
            if (have_fixed0) V.fixed = const_cast<uint8_t *>(fixed0);  // level 0 of a replicated hierarchy keeps its own (gathered) mask
        } else {
            LAUNCH(ctx, k_inject_fixed, grid_for(V.nrows), NT, 0, V.L, G->lev[l - 1].L, (const uint8_t *)G->lev[l - 1].fixed, V.fixed);
        }
        LAUNCH(ctx, k_dinv, grid_for(V.nrows), NT, 0, V.nrows, (const double *)V.K->diag, (const uint8_t *)V.fixed, V.dinv);
    }
    if (G->rep) {
        // mask of the first replicated level: my planes injected from the last distributed level, all-gathered as doubles
        Gmg *R = G->rep;
        const GmgLevel &F = G->lev.back();
        const int64_t nslab = 3 * (int64_t)G->rep_L.nown() * G->rep_L.plane();
        if (nslab > 0) {
            uint8_t *tmp = reinterpret_cast<uint8_t *>(G->rep_slab + nslab);  // tail of the slab buffer (sized for it)
            LAUNCH(ctx, k_inject_fixed, grid_for(nslab), NT, 0, G->rep_L, F.L, (const uint8_t *)F.fixed, tmp);
            LAUNCH(ctx, k_u8_to_f64, grid_for(nslab), NT, 0, nslab, (const uint8_t *)tmp, G->rep_slab);
        }
        gather_rep(ctx, G, G->rep_slab);
        LAUNCH(ctx, k_f64_to_u8, grid_for(R->lev[0].nrows), NT, 0, R->lev[0].nrows, (const double *)R->lev[0].b, R->lev[0].fixed);
        gmg_refresh(G->ctx1, R, G->rep_K, nullptr, false);
    }
    (void)K;
}

}  // namespace

int64_t gmg_window_doubles(int ne, int rank, int nranks) {
    if (nranks <= 1) return 0;
    const GmgPlan P = gmg_plan(ne, nranks);
    return plan_gather_off(P, rank) + plan_gather_doubles(P);
}

// debugging / tests: z = M^-1 r, one application of the V-cycle preconditioner to a host vector (this rank's rows)
void gmg_apply_host(smfem_ctx *ctx, smfem_matrix *K, const double *r_host, double *z_host) {
    REQUIRE(K->gmg_on && K->gmg_mesh, SMFEM_ERR_INVALID, "enable the multigrid preconditioner first");
    solver_alloc(ctx, K);
    if (K->gmg_dirty || !K->gmg) {
        gmg_free(K);
        K->gmg = gmg_build(ctx, K, K->gmg_mesh);
        gmg_bounds(ctx, static_cast<Gmg *>(K->gmg));
        K->gmg_dirty = false;
    }
    Gmg *G = static_cast<Gmg *>(K->gmg);
    gmg_refresh(ctx, G, K, K->has_bc ? K->fixed : nullptr, true);
    GmgLevel &V = G->lev[0];
    g_trace = std::getenv("SMFEM_GMG_TRACE") != nullptr;
    CUDA_CHECK(cudaMemcpyAsync(G->r, r_host, 8 * V.nrows, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_rhs, grid_for(V.nrows), NT, 0, V.nrows, (const double *)nullptr, (const double *)G->r, (const uint8_t *)V.fixed, G->r);  // mask
    vcycle(ctx, G, 0);
    g_trace = false;
    CUDA_CHECK(cudaMemcpyAsync(z_host, V.x + V.ghost, 8 * V.nrows, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

void gmg_free(smfem_matrix *K) {
    if (K->gmg && K->sol_x == static_cast<Gmg *>(K->gmg)->xs) K->sol_x = nullptr;
    gmg_destroy(static_cast<Gmg *>(K->gmg));
    K->gmg = nullptr;
}

void gmg_enable(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, bool enable) {
    if (!enable) {
        gmg_free(K);
        K->gmg_on = false;
        return;
    }
    REQUIRE(K->structured && K->ndim == 3 && K->nDof == 3 && mesh && mesh->structured && mesh->lat.n1 == K->lat.n1 &&
                mesh->lat.k0 == K->lat.k0 && mesh->lat.k1 == K->lat.k1,
            SMFEM_ERR_UNSUPPORTED, "the multigrid preconditioner needs the hex-lattice matrix and its mesh");
    REQUIRE(K->values_ready && K->mat_known, SMFEM_ERR_INVALID, "assemble K (and add the surface term) before enabling multigrid");
    (void)ctx;
    gmg_free(K);
    K->gmg_mesh = mesh;
    K->gmg_on = true;
    K->gmg_dirty = true;
}

void gmg_pcg_solve(smfem_ctx *ctx, smfem_matrix *K, double rtol, int maxit, const double *rhs_extra, double *q_out, int *iters,
                   double *relres) {
    REQUIRE(K->values_ready, SMFEM_ERR_INVALID, "matrix has no values yet");
    solver_alloc(ctx, K);
    REQUIRE(K->comm_connected, SMFEM_ERR_INVALID, "multi-GPU: call smfem_comm_connect first");
    if (K->gmg_dirty || !K->gmg) {  // values or surface term changed since the hierarchy was built
        gmg_free(K);
        K->gmg = gmg_build(ctx, K, K->gmg_mesh);
        gmg_bounds(ctx, static_cast<Gmg *>(K->gmg));
        K->gmg_dirty = false;
    }
    Gmg *G = static_cast<Gmg *>(K->gmg);
    const double warm = K->sol_x ? K->warm_scale : 0.0;
    K->warm_scale = 0.0;  // applies to one solve
    CUDA_CHECK(cudaEventRecord(ctx->ev2, ctx->stream));
    gmg_refresh(ctx, G, K, K->has_bc ? K->fixed : nullptr, true);
    GmgLevel &V = G->lev[0];
    const int64_t n = V.nrows, gh = V.ghost;
    const unsigned g = grid_for(n);
    double *extra = nullptr;
    if (rhs_extra) {
        extra = dev_alloc<double>(n);
        CUDA_CHECK(cudaMemcpyAsync(extra, rhs_extra, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    const bool bc = K->has_bc && K->qd;
    if (bc) spmv_device(ctx, K, K->qd, V.y);  // q_d's ghost entries are filled locally: no exchange
    LAUNCH(ctx, k_rhs, g, NT, 0, n, (const double *)(bc ? V.y : nullptr), (const double *)extra, (const uint8_t *)V.fixed, G->r);
    CUDA_CHECK(cudaMemsetAsync(G->p, 0, 8 * V.ncols, ctx->stream));
    const double bnorm2 = dot(ctx, G, n, G->r, G->r);
    double res2 = bnorm2, rz_old = 0.0;
    if (warm != 0.0) {  // x0 = warm * previous solution (of either solver), r0 = b - A x0
        LAUNCH(ctx, k_warm, g, NT, 0, n, gh, warm, K->sol_x, G->xs, G->p);
        spmv_level(ctx, G, V, G->p, G->p_lo, G->p_hi, V.y);
        LAUNCH(ctx, k_sub_masked, g, NT, 0, n, (const double *)V.y, (const uint8_t *)V.fixed, G->r);
        res2 = dot(ctx, G, n, G->r, G->r);
    } else {
        CUDA_CHECK(cudaMemsetAsync(G->xs, 0, 8 * n, ctx->stream));
    }
    int it = 0;
    bool breakdown = false;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    if (bnorm2 > 0.0) {
        const double tol2 = rtol * rtol * bnorm2;
        // the V-cycle is ~130 launches, most of them on tiny coarse levels: captured once, replayed per iteration
        const int64_t l0 = ctx->launches, l1 = G->ctx1 ? G->ctx1->launches : 0;
        CUDA_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        vcycle(ctx, G, 0);
        const int64_t per_cycle = ctx->launches - l0 + (G->ctx1 ? G->ctx1->launches - l1 : 0);
        CUDA_CHECK(cudaStreamEndCapture(ctx->stream, &graph));
        ctx->launches = l0;  // captured, not launched; counted per replay below
        CUDA_CHECK(cudaGraphInstantiate(&gexec, graph, 0));
        while (it < maxit && res2 > tol2) {
            // z = M^-1 r: one V-cycle on level 0, whose right-hand side buffer is r itself (result in V.x)
            CUDA_CHECK(cudaGraphLaunch(gexec, ctx->stream));
            ctx->launches += per_cycle;
            const double rz = dot(ctx, G, n, G->r, V.x + gh);
            const double beta = it == 0 ? 0.0 : rz / rz_old;
            rz_old = rz;
            LAUNCH(ctx, k_update_p, g, NT, 0, n, gh, (const double *)V.x, beta, G->p);
            spmv_level(ctx, G, V, G->p, G->p_lo, G->p_hi, V.y);
            const double pAp = dot(ctx, G, n, G->p + gh, V.y);
            if (!(pAp > 0.0)) {
                breakdown = true;
                break;
            }
            LAUNCH(ctx, k_update_xr, g, NT, 0, n, gh, rz / pAp, (const double *)G->p, (const double *)V.y, (const uint8_t *)V.fixed, G->xs, G->r);
            res2 = dot(ctx, G, n, G->r, G->r);
            ++it;
            if (!(res2 == res2)) break;
        }
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    // q = q_d + C q_f on the owned rows (examples/vector3D.jl:322), and the TRUE residual ||(extra - K q)_free|| at exit
    double res2_true = res2;
    LAUNCH(ctx, k_final, g, NT, 0, n, gh, (const double *)(bc ? K->qd : nullptr), (const double *)G->xs, (const uint8_t *)V.fixed, G->p + gh);
    if (q_out) CUDA_CHECK(cudaMemcpyAsync(q_out, G->p + gh, 8 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (bnorm2 > 0.0 && !breakdown) {
        spmv_level(ctx, G, V, G->p, G->p_lo, G->p_hi, V.y);
        LAUNCH(ctx, k_rhs, g, NT, 0, n, (const double *)V.y, (const double *)extra, (const uint8_t *)V.fixed, V.d);
        res2_true = dot(ctx, G, n, V.d, V.d);
    }
    CUDA_CHECK(cudaEventRecord(ctx->ev3, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(ctx->ev3));
    CUDA_CHECK(cudaEventElapsedTime(&K->last_ms, ctx->ev2, ctx->ev3));
    K->last_iters = it;
    K->sol_x = G->xs;
    K->last_relres_rec = bnorm2 > 0 ? std::sqrt(res2 / bnorm2) : 0.0;
    K->last_relres_true = bnorm2 > 0 ? std::sqrt(res2_true / bnorm2) : 0.0;
    if (extra) dev_free(extra);
    if (iters) *iters = it;
    if (relres) *relres = K->last_relres_true;
    REQUIRE(!breakdown, SMFEM_ERR_SINGULAR, "PCG breakdown: p'Ap <= 0 (matrix not SPD on the free dofs; reference: SingularException)");
}
