// Geometric-multigrid preconditioner for the CG solve of the hex-lattice problem (opt-in through
// smfem_pcg_use_multigrid, single GPU).  The Jacobi-PCG of solver.cu is the path the north star names and stays the
// default; its iteration count grows like ne (875 at 100^3).  Here:
//   * levels: ne -> ceil(ne/2) -> ... down to <= 4; a coarse mesh takes every other node of the finer one, plus the last
//     one when ne is odd (so the inflated geometry and the z = 0 / z = 1 planes are followed), and its operator is RE-ASSEMBLED with the library's own path (smfem_assemble +
//     smfem_surface_mass with the same E, nu, beta): no Galerkin triple products, no new pattern code;
//   * Dirichlet rows: a coarse dof is constrained iff the fine dof at the same node is (injection);
//   * smoother: Chebyshev in D^-1 A on [lmax/8, lmax] (lmax from a few power iterations), 2 steps before and after the
//     coarse correction; Jacobi with a fixed damping needed 68 iterations where Chebyshev(2) needs 17 (ne = 16, CPU prototype);
//   * transfers: trilinear interpolation P per displacement component, restriction P';
//   * coarsest level: 30 Chebyshev steps (the iteration count does not react to a better coarse solve);
//   * outer iteration: standard PCG, dot products by a fixed-grid two-stage reduction (bit-reproducible).
// The V-cycle is a fixed linear operator (same polynomial every time), symmetric in the sense PCG needs.
#include <cmath>
#include <vector>

#include "smfem_internal.cuh"

namespace {

constexpr int NT = 256;

struct GmgLevel {
    smfem_mesh *mesh = nullptr;  // owned for levels > 0
    smfem_matrix *K = nullptr;   // owned for levels > 0
    Lattice L;
    int64_t nrows = 0, ncols = 0, ghost = 0;
    double *x = nullptr;  // ncols (ghost planes included: SpMV input)
    double *b = nullptr, *d = nullptr, *y = nullptr, *dinv = nullptr;  // nrows
    uint8_t *fixed = nullptr;  // nrows (level 0: the matrix's own mask, may be null)
    bool own_fixed = false;
    double lmax = 0;
};

struct Gmg {
    std::vector<GmgLevel> lev;
    double *p = nullptr;   // ncols of level 0
    double *r = nullptr;   // = lev[0].b (not owned): the residual is the V-cycle's right-hand side
    double *xs = nullptr;  // nrows of level 0
    double *partials = nullptr, *h_pinned = nullptr;
    int red_grid = 0;
};

// Coarsening along one axis: ne_c = ceil(ne_f / 2); coarse node I sits on fine node min(2 I, ne_f) (for odd ne_f the last
// coarse element is a single fine element).  pw = weight of coarse node I in the linear interpolation at fine node i.
__device__ __forceinline__ int fine_of(int I, int ne_f) { return min(2 * I, ne_f); }
__device__ __forceinline__ double pw(int i, int I, int ne_f) {
    if (i == ne_f && (ne_f & 1)) return I == (ne_f + 1) / 2 ? 1.0 : 0.0;
    if (!(i & 1)) return I == (i >> 1) ? 1.0 : 0.0;
    return (I == (i >> 1) || I == (i >> 1) + 1) ? 0.5 : 0.0;
}

__global__ void k_subsample_coords(Lattice Lc, Lattice Lf, const double *__restrict__ cf, double *__restrict__ cc) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)Lc.n1 * Lc.n1 * Lc.n1;
    if (t >= n) return;
    const int i = (int)(t % Lc.n1), j = (int)((t / Lc.n1) % Lc.n1), k = (int)(t / ((int64_t)Lc.n1 * Lc.n1));
    const int64_t nc = Lc.lnode(i, j, k), nf = Lf.lnode(fine_of(i, Lf.ne), fine_of(j, Lf.ne), fine_of(k, Lf.ne));
#pragma unroll
    for (int c = 0; c < 3; ++c) cc[3 * nc + c] = cf[3 * nf + c];
}

__global__ void k_inject_fixed(int n1c, int n1f, const uint8_t *__restrict__ ff, uint8_t *__restrict__ fc) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = 3 * (int64_t)n1c * n1c * n1c;
    if (t >= n) return;
    const int c = (int)(t % 3);
    const int64_t m = t / 3;
    const int i = (int)(m % n1c), j = (int)((m / n1c) % n1c), k = (int)(m / ((int64_t)n1c * n1c));
    const int nef = n1f - 1;
    const int64_t mf = ((int64_t)fine_of(k, nef) * n1f + fine_of(j, nef)) * n1f + fine_of(i, nef);
    fc[t] = ff ? ff[3 * mf + c] : 0;
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

// coarse rhs = P' (b - A x) on the free coarse rows; fine residual rows of constrained dofs do not contribute
__global__ void k_restrict(int n1c, int n1f, const double *__restrict__ bf, const double *__restrict__ yf,
                           const uint8_t *__restrict__ fixed_f, const uint8_t *__restrict__ fixed_c, double *__restrict__ bc) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = 3 * (int64_t)n1c * n1c * n1c;
    if (t >= n) return;
    if (fixed_c[t]) {
        bc[t] = 0.0;
        return;
    }
    const int c = (int)(t % 3);
    const int64_t m = t / 3;
    const int I = (int)(m % n1c), J = (int)((m / n1c) % n1c), Kz = (int)(m / ((int64_t)n1c * n1c));
    const int nef = n1f - 1;
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
                const int64_t rf = 3 * (((int64_t)k * n1f + j) * n1f + i) + c;
                if (fixed_f && fixed_f[rf]) continue;
                s += wi * wj * wk * (bf[rf] - yf[rf]);
            }
        }
    }
    bc[t] = s;
}

// x_f += P x_c on the free fine rows
__global__ void k_prolong_add(int n1f, int n1c, int64_t ghost_f, int64_t ghost_c, const double *__restrict__ xc,
                              const uint8_t *__restrict__ fixed_f, double *__restrict__ xf) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = 3 * (int64_t)n1f * n1f * n1f;
    if (t >= n) return;
    if (fixed_f && fixed_f[t]) return;
    const int c = (int)(t % 3);
    const int64_t m = t / 3;
    const int i = (int)(m % n1f), j = (int)((m / n1f) % n1f), k = (int)(m / ((int64_t)n1f * n1f));
    const int nef = n1f - 1;
    // parents along each axis: first parent a0 and whether there is a second one (weights 1 or 1/2 + 1/2)
    const bool li = (i == nef) && (nef & 1), lj = (j == nef) && (nef & 1), lk = (k == nef) && (nef & 1);
    const int i0 = li ? (nef + 1) / 2 : i >> 1, j0 = lj ? (nef + 1) / 2 : j >> 1, k0 = lk ? (nef + 1) / 2 : k >> 1;
    const int oi = (!li) && (i & 1), oj = (!lj) && (j & 1), ok = (!lk) && (k & 1);
    double s = 0.0;
    for (int dk = 0; dk <= ok; ++dk)
        for (int dj = 0; dj <= oj; ++dj)
            for (int di = 0; di <= oi; ++di) {
                const int64_t rc = 3 * (((int64_t)(k0 + dk) * n1c + (j0 + dj)) * n1c + (i0 + di)) + c;
                s += xc[rc + ghost_c];
            }
    xf[t + ghost_f] += s * (oi ? 0.5 : 1.0) * (oj ? 0.5 : 1.0) * (ok ? 0.5 : 1.0);
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
__global__ void k_fill_test(int64_t n, int64_t ghost, const double *__restrict__ dinv, double *__restrict__ x) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) x[t + ghost] = dinv[t] != 0.0 ? 1.0 + 0.37 * (double)((t * 2654435761ull) % 1000) / 1000.0 : 0.0;
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + NT - 1) / NT); }

double dot(smfem_ctx *ctx, Gmg *G, int64_t n, const double *a, const double *b) {
    LAUNCH(ctx, k_dot1, G->red_grid, NT, 0, n, a, b, G->partials);
    LAUNCH(ctx, k_dot2, 1, NT, 0, G->red_grid, (const double *)G->partials, G->partials + G->red_grid);
    CUDA_CHECK(cudaMemcpyAsync(G->h_pinned, G->partials + G->red_grid, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return G->h_pinned[0];
}

// n Chebyshev steps on level l for A x = b (x_zero: the guess is 0 and x need not be read)
void smooth(smfem_ctx *ctx, GmgLevel &V, int n, bool x_zero) {
    const double lmax = V.lmax, lmin = lmax / 8.0;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    const unsigned g = grid_for(V.nrows);
    for (int k = 0; k < n; ++k) {
        if (k == 0 && x_zero) {
            LAUNCH(ctx, k_cheb_first, g, NT, 0, V.nrows, V.ghost, (const double *)V.b, (const double *)V.dinv, 1.0 / theta, V.x, V.d);
            continue;
        }
        spmv_device(ctx, V.K, V.x, V.y);
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

void vcycle(smfem_ctx *ctx, Gmg *G, int l) {
    GmgLevel &V = G->lev[l];
    const int nl = (int)G->lev.size();
    if (l == nl - 1) {
        smooth(ctx, V, nl == 1 ? 2 : 30, true);
        return;
    }
    GmgLevel &C = G->lev[l + 1];
    smooth(ctx, V, 2, true);
    spmv_device(ctx, V.K, V.x, V.y);
    LAUNCH(ctx, k_restrict, grid_for(C.nrows), NT, 0, C.L.n1, V.L.n1, (const double *)V.b, (const double *)V.y, (const uint8_t *)V.fixed,
           (const uint8_t *)C.fixed, C.b);
    vcycle(ctx, G, l + 1);
    LAUNCH(ctx, k_prolong_add, grid_for(V.nrows), NT, 0, V.L.n1, C.L.n1, V.ghost, C.ghost, (const double *)C.x, (const uint8_t *)V.fixed, V.x);
    smooth(ctx, V, 2, false);
}

void level_free(GmgLevel &V, bool owned) {
    dev_free(V.x);
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
    for (size_t l = 0; l < G->lev.size(); ++l) level_free(G->lev[l], l > 0);
    dev_free(G->p);
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

Gmg *gmg_build(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh) {
    Gmg *G = new Gmg();
    try {
        G->red_grid = ctx->sms * 4;
        G->partials = dev_alloc<double>(G->red_grid + 8);
        CUDA_CHECK(cudaMallocHost(&G->h_pinned, 64));
        GmgLevel V0;
        V0.mesh = mesh;
        V0.K = K;
        G->lev.push_back(V0);
        int ne = K->lat.ne;
        while (ne > 4 && G->lev.size() < 12) {
            ne = (ne + 1) / 2;
            GmgLevel V;
            ABI_CHECK(smfem_meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3, &V.mesh));
            G->lev.push_back(V);  // owned from here on (freed by gmg_destroy)
            GmgLevel &C = G->lev.back();
            const GmgLevel &F = G->lev[G->lev.size() - 2];
            const int64_t nn = (int64_t)(ne + 1) * (ne + 1) * (ne + 1);
            LAUNCH(ctx, k_subsample_coords, grid_for(nn), NT, 0, C.mesh->lat, F.mesh->lat, (const double *)F.mesh->coords, C.mesh->coords);
            ABI_CHECK(smfem_assemble(ctx, C.mesh, ne, 3, SMFEM_Q1, 3, K->Young, K->nu, &C.K));
            if (K->beta_total != 0.0) ABI_CHECK(smfem_surface_mass(ctx, C.K, C.mesh, nullptr, nullptr, 0, K->beta_total, 0));
            solver_alloc(ctx, C.K);
        }
        for (size_t l = 0; l < G->lev.size(); ++l) {
            GmgLevel &V = G->lev[l];
            V.L = V.K->lat;
            V.nrows = V.K->nrows_l;
            V.ncols = V.K->ncols_l;
            V.ghost = V.K->ghost_cols;
            V.x = dev_alloc<double>(V.ncols);
            CUDA_CHECK(cudaMemsetAsync(V.x, 0, 8 * V.ncols, ctx->stream));
            V.b = dev_alloc<double>(V.nrows);
            V.d = dev_alloc<double>(V.nrows);
            V.y = dev_alloc<double>(V.nrows);
            V.dinv = dev_alloc<double>(V.nrows);
            if (l > 0) {
                V.fixed = dev_alloc<uint8_t>(V.nrows);
                V.own_fixed = true;
            }
        }
        const GmgLevel &V = G->lev[0];
        G->p = dev_alloc<double>(V.ncols);
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
        LAUNCH(ctx, k_fill_test, grid_for(V.nrows), NT, 0, V.nrows, V.ghost, (const double *)V.dinv, V.x);
        double lam = 1.0;
        for (int it = 0; it < 15; ++it) {
            spmv_device(ctx, V.K, V.x, V.y);
            // x <- D^-1 A x, normalised (||x|| = 1 after the first pass, so the norm is the eigenvalue estimate)
            LAUNCH(ctx, k_scale_copy, grid_for(V.nrows), NT, 0, V.nrows, V.ghost, 1.0, (const double *)V.y, (const double *)V.dinv, V.x);
            const double nrm = std::sqrt(dot(ctx, G, V.nrows, V.x + V.ghost, V.x + V.ghost));
            REQUIRE(nrm > 0.0 && nrm == nrm, SMFEM_ERR_SINGULAR, "multigrid setup: D^-1 A x vanished (matrix without values?)");
            lam = nrm;
            LAUNCH(ctx, k_scale_copy, grid_for(V.nrows), NT, 0, V.nrows, V.ghost, 1.0 / nrm, (const double *)V.y, (const double *)V.dinv, V.x);
        }
        V.lmax = 1.1 * lam;
        CUDA_CHECK(cudaMemsetAsync(V.x, 0, 8 * V.ncols, ctx->stream));
    }
}

// masks and D^-1 of every level for the CURRENT boundary conditions (no host synchronisation)
void gmg_refresh(smfem_ctx *ctx, Gmg *G, smfem_matrix *K) {
    for (size_t l = 0; l < G->lev.size(); ++l) {
        GmgLevel &V = G->lev[l];
        if (l == 0) V.fixed = K->has_bc ? K->fixed : nullptr;
        else
            LAUNCH(ctx, k_inject_fixed, grid_for(V.nrows), NT, 0, V.L.n1, G->lev[l - 1].L.n1, (const uint8_t *)G->lev[l - 1].fixed, V.fixed);
        LAUNCH(ctx, k_dinv, grid_for(V.nrows), NT, 0, V.nrows, (const double *)V.K->diag, (const uint8_t *)V.fixed, V.dinv);
    }
}

}  // namespace

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
    REQUIRE(ctx->nranks == 1, SMFEM_ERR_UNSUPPORTED, "the multigrid preconditioner is single-GPU (multi-GPU solves use Jacobi-PCG)");
    REQUIRE(K->structured && K->ndim == 3 && K->nDof == 3 && mesh && mesh->structured && mesh->lat.n1 == K->lat.n1, SMFEM_ERR_UNSUPPORTED,
            "the multigrid preconditioner needs the hex-lattice matrix and its mesh");
    REQUIRE(K->values_ready && K->mat_known, SMFEM_ERR_INVALID, "assemble K (and add the surface term) before enabling multigrid");
    gmg_free(K);
    K->gmg_mesh = mesh;
    K->gmg_on = true;
    K->gmg_dirty = true;
}

void gmg_pcg_solve(smfem_ctx *ctx, smfem_matrix *K, double rtol, int maxit, const double *rhs_extra, double *q_out, int *iters,
                   double *relres) {
    REQUIRE(K->values_ready, SMFEM_ERR_INVALID, "matrix has no values yet");
    solver_alloc(ctx, K);
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
    gmg_refresh(ctx, G, K);
    GmgLevel &V = G->lev[0];
    const int64_t n = V.nrows, gh = V.ghost;
    const unsigned g = grid_for(n);
    double *extra = nullptr;
    if (rhs_extra) {
        extra = dev_alloc<double>(n);
        CUDA_CHECK(cudaMemcpyAsync(extra, rhs_extra, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    const bool bc = K->has_bc && K->qd;
    if (bc) spmv_device(ctx, K, K->qd, V.y);
    LAUNCH(ctx, k_rhs, g, NT, 0, n, (const double *)(bc ? V.y : nullptr), (const double *)extra, (const uint8_t *)V.fixed, G->r);
    CUDA_CHECK(cudaMemsetAsync(G->p, 0, 8 * V.ncols, ctx->stream));
    const double bnorm2 = dot(ctx, G, n, G->r, G->r);
    double res2 = bnorm2, rz_old = 0.0;
    if (warm != 0.0) {  // x0 = warm * previous solution (of either solver), r0 = b - A x0
        LAUNCH(ctx, k_warm, g, NT, 0, n, gh, warm, K->sol_x, G->xs, G->p);
        spmv_device(ctx, K, G->p, V.y);
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
        const int64_t l0 = ctx->launches;
        CUDA_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        vcycle(ctx, G, 0);
        const int64_t per_cycle = ctx->launches - l0;
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
            spmv_device(ctx, K, G->p, V.y);
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
    if (q_out) {
        LAUNCH(ctx, k_final, g, NT, 0, n, gh, (const double *)(bc ? K->qd : nullptr), (const double *)G->xs, (const uint8_t *)V.fixed, V.y);
        CUDA_CHECK(cudaMemcpyAsync(q_out, V.y, 8 * n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_CHECK(cudaEventRecord(ctx->ev3, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(ctx->ev3));
    CUDA_CHECK(cudaEventElapsedTime(&K->last_ms, ctx->ev2, ctx->ev3));
    K->last_iters = it;
    K->sol_x = G->xs;
    if (extra) dev_free(extra);
    if (iters) *iters = it;
    if (relres) *relres = bnorm2 > 0 ? std::sqrt(res2 / bnorm2) : 0.0;
    REQUIRE(!breakdown, SMFEM_ERR_SINGULAR, "PCG breakdown: p'Ap <= 0 (matrix not SPD on the free dofs; reference: SingularException)");
}
