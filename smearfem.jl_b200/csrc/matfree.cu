// Matrix-free form of the solve's operator on the structured hex lattice (second half):
//     y = K_bar x = (K + beta b) x      without reading K's 12 bytes per nonzero
// K is what src/fem.jl:179-249 assembles (Ke = int B'DB, isotropic D, 2x2x2 Gauss rule), b the surface term of
// examples/vector3D.jl:193-262.  Per element and Gauss point, with g_a = J^-T grad N_a:
//     H = sum_b x_b g_b'   (3x3 displacement gradient),   sigma = lam tr(H) I + mu (H + H'),   y_a += w |det J| sigma g_a
// which is exactly sum_b (lam G_ab + mu G_ab' + mu tr(G_ab) I) x_b with G_ab = g_a g_b' (the identity the assembly kernels use).
// Traffic per application: coordinates + x + y = 72 B per node (0.07 GB at 100^3) instead of 2.0 GB of CSR values.
//
// One thread per element, 8 launches = the 8 parity colours of the lattice: elements of one colour share no node, so the
// contributions are folded into y with plain load-add-store in a fixed order (colour order) -> no atomics, bit-reproducible.
// Several ranks: a rank applies the element layers [k0-1, k1) of its slab and keeps the rows of its owned node planes; x carries
// the ghost planes exactly as for the CSR SpMV (same halo exchange), the ghost layer is recomputed instead of communicated.
#include <cmath>
#include <cstdlib>

#include "smfem_internal.cuh"

namespace {

struct MfArgs {
    Lattice L;
    const double *coords;  // local nodes with ghost planes, xyz contiguous
    const double *x;       // ncols_l (ghost planes included)
    double *y;             // owned rows
    double lam, mu, beta;
    double gpc;            // 1/sqrt(3): the 2-point rule on [-1, 1] (weights 1), src/fem.jl:21-31
    const PcgScalars *scal;
    int check_done;
    int cx, cy, cz;        // colour = parity of the element coordinates
    int nx, ny, nz, ez0;   // elements of this colour per axis (z: of this rank's layers, first layer ez0)
};

__global__ void __launch_bounds__(128) k_matfree_color(const __grid_constant__ MfArgs A) {
    if (A.check_done && A.scal->done) return;
    const Lattice &L = A.L;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)A.nx * A.ny * A.nz) return;
    const int tx = (int)(t % A.nx), ty = (int)((t / A.nx) % A.ny), tz = (int)(t / ((int64_t)A.nx * A.ny));
    const int ex = 2 * tx + A.cx, ey = 2 * ty + A.cy, ez = A.ez0 + 2 * tz;
    // the 8 nodes in natural order u = ox + 2 oy + 4 oz
    double X[8][3], U[8][3];
    int64_t ln[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        ln[u] = L.lnode(ex + (u & 1), ey + ((u >> 1) & 1), ez + (u >> 2));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            X[u][c] = A.coords[3 * ln[u] + c];
            U[u][c] = A.x[3 * ln[u] + c];
        }
    }
    double Y[8][3];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int c = 0; c < 3; ++c) Y[u][c] = 0.0;
#pragma unroll 1
    for (int gp = 0; gp < 8; ++gp) {
        const double xi = (gp & 1) ? A.gpc : -A.gpc, eta = (gp & 2) ? A.gpc : -A.gpc, zeta = (gp & 4) ? A.gpc : -A.gpc;
        const double Xf[2] = {1.0 - xi, 1.0 + xi}, Yf[2] = {1.0 - eta, 1.0 + eta}, Zf[2] = {0.125 * (1.0 - zeta), 0.125 * (1.0 + zeta)};
        // reference gradients d_u = (sx Y Z, X sy Z, X Y sz) / 8
        double d[8][3];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int ox = u & 1, oy = (u >> 1) & 1, oz = u >> 2;
            const double yz = Yf[oy] * Zf[oz], xz = Xf[ox] * Zf[oz], xy = 0.125 * Xf[ox] * Yf[oy];
            d[u][0] = ox ? yz : -yz;
            d[u][1] = oy ? xz : -xz;
            d[u][2] = oz ? xy : -xy;
        }
        // J[r][k] = d x_r / d xi_k (Jac = coords * dN, src/fem.jl:192);  G[i][k] = sum_u U_u,i d_u,k (reference displacement gradient)
        double J[9], G[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) J[q] = G[q] = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    J[r * 3 + k] += X[u][r] * d[u][k];
                    G[r * 3 + k] += U[u][r] * d[u][k];
                }
        double adj[9];  // adj[k][c]: J^-1 = adj / det (row k = reference direction, column c = physical direction)
        adj[0] = J[4] * J[8] - J[5] * J[7];
        adj[1] = J[2] * J[7] - J[1] * J[8];
        adj[2] = J[1] * J[5] - J[2] * J[4];
        adj[3] = J[5] * J[6] - J[3] * J[8];
        adj[4] = J[0] * J[8] - J[2] * J[6];
        adj[5] = J[2] * J[3] - J[0] * J[5];
        adj[6] = J[3] * J[7] - J[4] * J[6];
        adj[7] = J[1] * J[6] - J[0] * J[7];
        adj[8] = J[0] * J[4] - J[1] * J[3];
        const double det = J[0] * adj[0] + J[1] * adj[3] + J[2] * adj[6];
        const double f = 1.0 / fabs(det);  // w |det| / det^2 with w = 1 (src/fem.jl:196 uses abs(det))
        // H det = G adj   (H[i][c] = sum_k G[i][k] adj[k][c] / det)
        double H[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) H[i * 3 + c] = G[i * 3] * adj[c] + G[i * 3 + 1] * adj[3 + c] + G[i * 3 + 2] * adj[6 + c];
        const double tr = A.lam * (H[0] + H[4] + H[8]);
        double S[9];  // sigma det * f
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) S[i * 3 + c] = f * (A.mu * (H[i * 3 + c] + H[c * 3 + i]) + (i == c ? tr : 0.0));
        // y_a += sigma g_a,  g_a det = adj' d_a   ->   y_a,i += sum_k P[i][k] d_a,k,   P = S adj'
        double P[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) P[i * 3 + k] = S[i * 3] * adj[k * 3] + S[i * 3 + 1] * adj[k * 3 + 1] + S[i * 3 + 2] * adj[k * 3 + 2];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 3; ++i) Y[u][i] += P[i * 3] * d[u][0] + P[i * 3 + 1] * d[u][1] + P[i * 3 + 2] * d[u][2];
    }
    // fold into the owned rows (same-colour elements touch disjoint nodes: plain read-modify-write)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int k = ez + (u >> 2);
        if (k < L.k0 || k >= L.k1) continue;
        double *dst = A.y + 3 * (ln[u] - L.plane());  // owned row = local node minus the lower ghost plane
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[c] += Y[u][c];
    }
}

// Second form of the element kernel: the trilinear maps in MONOMIAL coordinates.  With phi = {1, xi, eta, zeta, xi eta, xi zeta,
// eta zeta, xi eta zeta} and c = W F / 8 (W = the 8 x 8 sign matrix of the corner shape functions, a 3-stage butterfly), the
// reference gradient of a nodal field F is
//     dF/dxi = c1 + c4 eta + c5 zeta + c7 eta zeta,   dF/deta = c2 + c4 xi + c6 zeta + c7 xi zeta,   dF/dzeta = c3 + c5 xi + c6 eta + c7 xi eta
// i.e. 9 FMA per column instead of 24, for the Jacobian (F = coordinates) and for the displacement gradient (F = x) alike, and
// the test side accumulates t_m += P[:, k] dphi_m/dxi_k in the same 7 monomials (36 instead of 72 operations per Gauss point),
// transformed back to the 8 nodes once per element (y = W' t / 8).  63 persistent doubles (c for X and U, t) instead of 96.
__device__ __forceinline__ void corner_to_monomial(const double (&F)[8][3], double (&c)[8][3]) {
    // natural corner order u = ox + 2 oy + 4 oz; N_u = (1 + sx xi)(1 + sy eta)(1 + sz zeta) / 8
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = F[u][i];
        // butterfly along x, y, z: sums -> even monomial power, differences -> odd
        double b[8], d[8];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            b[2 * p] = a[2 * p + 1] + a[2 * p];
            b[2 * p + 1] = a[2 * p + 1] - a[2 * p];
        }
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                d[4 * p + x] = b[4 * p + 2 + x] + b[4 * p + x];
                d[4 * p + 2 + x] = b[4 * p + 2 + x] - b[4 * p + x];
            }
        // index bits now: bit0 = x power, bit1 = y power; finish with z
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double lo = d[q], hi = d[4 + q];
            const double sum = 0.125 * (hi + lo), dif = 0.125 * (hi - lo);
            // monomial numbering m: 0:1 1:xi 2:eta 3:zeta 4:xi eta 5:xi zeta 6:eta zeta 7:xi eta zeta
            const int m_sum = q == 0 ? 0 : (q == 1 ? 1 : (q == 2 ? 2 : 4));
            const int m_dif = q == 0 ? 3 : (q == 1 ? 5 : (q == 2 ? 6 : 7));
            c[m_sum][i] = sum;
            c[m_dif][i] = dif;
        }
    }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_matfree_color2(const __grid_constant__ MfArgs A) {
    if (A.check_done && A.scal->done) return;
    const Lattice &L = A.L;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)A.nx * A.ny * A.nz) return;
    const int tx = (int)(t % A.nx), ty = (int)((t / A.nx) % A.ny), tz = (int)(t / ((int64_t)A.nx * A.ny));
    const int ex = 2 * tx + A.cx, ey = 2 * ty + A.cy, ez = A.ez0 + 2 * tz;
    const int64_t n0 = L.lnode(ex, ey, ez);
    const int64_t sy = L.n1, sz = L.plane();
    double cX[8][3], cU[8][3];
    {
        double F[8][3];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t ln = n0 + (u & 1) + ((u >> 1) & 1) * sy + (u >> 2) * sz;
#pragma unroll
            for (int c = 0; c < 3; ++c) F[u][c] = A.coords[3 * ln + c];
        }
        corner_to_monomial(F, cX);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t ln = n0 + (u & 1) + ((u >> 1) & 1) * sy + (u >> 2) * sz;
#pragma unroll
            for (int c = 0; c < 3; ++c) F[u][c] = A.x[3 * ln + c];
        }
        corner_to_monomial(F, cU);
    }
    double T[8][3];  // test-side accumulators in the monomials 1..7 (T[0] stays 0)
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int i = 0; i < 3; ++i) T[m][i] = 0.0;
#pragma unroll 1
    for (int gp = 0; gp < 8; ++gp) {
        const double xi = (gp & 1) ? A.gpc : -A.gpc, eta = (gp & 2) ? A.gpc : -A.gpc, zeta = (gp & 4) ? A.gpc : -A.gpc;
        const double xe = xi * eta, xz = xi * zeta, ez_ = eta * zeta;
        double J[9], G[9];  // [r][k]: d x_r / d xi_k and d u_r / d xi_k
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            J[r * 3 + 0] = cX[1][r] + cX[4][r] * eta + cX[5][r] * zeta + cX[7][r] * ez_;
            J[r * 3 + 1] = cX[2][r] + cX[4][r] * xi + cX[6][r] * zeta + cX[7][r] * xz;
            J[r * 3 + 2] = cX[3][r] + cX[5][r] * xi + cX[6][r] * eta + cX[7][r] * xe;
            G[r * 3 + 0] = cU[1][r] + cU[4][r] * eta + cU[5][r] * zeta + cU[7][r] * ez_;
            G[r * 3 + 1] = cU[2][r] + cU[4][r] * xi + cU[6][r] * zeta + cU[7][r] * xz;
            G[r * 3 + 2] = cU[3][r] + cU[5][r] * xi + cU[6][r] * eta + cU[7][r] * xe;
        }
        double adj[9];
        adj[0] = J[4] * J[8] - J[5] * J[7];
        adj[1] = J[2] * J[7] - J[1] * J[8];
        adj[2] = J[1] * J[5] - J[2] * J[4];
        adj[3] = J[5] * J[6] - J[3] * J[8];
        adj[4] = J[0] * J[8] - J[2] * J[6];
        adj[5] = J[2] * J[3] - J[0] * J[5];
        adj[6] = J[3] * J[7] - J[4] * J[6];
        adj[7] = J[1] * J[6] - J[0] * J[7];
        adj[8] = J[0] * J[4] - J[1] * J[3];
        const double det = J[0] * adj[0] + J[1] * adj[3] + J[2] * adj[6];
        const double f = 1.0 / fabs(det);
        double H[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) H[i * 3 + c] = G[i * 3] * adj[c] + G[i * 3 + 1] * adj[3 + c] + G[i * 3 + 2] * adj[6 + c];
        const double tr = A.lam * (H[0] + H[4] + H[8]);
        double S[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) S[i * 3 + c] = f * (A.mu * (H[i * 3 + c] + H[c * 3 + i]) + (i == c ? tr : 0.0));
        double P[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) P[i * 3 + k] = S[i * 3] * adj[k * 3] + S[i * 3 + 1] * adj[k * 3 + 1] + S[i * 3 + 2] * adj[k * 3 + 2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double p0 = P[i * 3], p1 = P[i * 3 + 1], p2 = P[i * 3 + 2];
            T[1][i] += p0;
            T[2][i] += p1;
            T[3][i] += p2;
            T[4][i] += p0 * eta + p1 * xi;
            T[5][i] += p0 * zeta + p2 * xi;
            T[6][i] += p1 * zeta + p2 * eta;
            T[7][i] += p0 * ez_ + p1 * xz + p2 * xe;
        }
    }
    // y_u = sum_m W[u][m] T[m] / 8 with W[u][m] = the sign product of monomial m at corner u; the same butterfly, transposed
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // z stage: (m_sum, m_dif) pairs as in corner_to_monomial
        const double s0 = T[0][i], s1 = T[1][i], s2 = T[2][i], s3 = T[4][i];   // even in zeta: 1, xi, eta, xi eta
        const double d0 = T[3][i], d1 = T[5][i], d2 = T[6][i], d3 = T[7][i];   // odd in zeta
        double lo[4] = {s0 - d0, s1 - d1, s2 - d2, s3 - d3}, hi[4] = {s0 + d0, s1 + d1, s2 + d2, s3 + d3};
        double yv[8];
#pragma unroll
        for (int oz = 0; oz < 2; ++oz) {
            const double *v = oz ? hi : lo;  // [1, xi, eta, xi eta] coefficients on this z face
            const double e0 = v[0] - v[2], e1 = v[1] - v[3], f0 = v[0] + v[2], f1 = v[1] + v[3];  // eta = -1 / +1
            yv[4 * oz + 0] = e0 - e1;
            yv[4 * oz + 1] = e0 + e1;
            yv[4 * oz + 2] = f0 - f1;
            yv[4 * oz + 3] = f0 + f1;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = ez + (u >> 2);
            if (k < L.k0 || k >= L.k1) continue;
            const int64_t ln = n0 + (u & 1) + ((u >> 1) & 1) * sy + (u >> 2) * sz;
            // exactly one contribution per address and launch (same-colour elements share no node), launches are ordered: the
            // reduction (RED.ADD.F64, no load, no wait) is as deterministic as a load-add-store and hides the latency of y
            atomicAdd(A.y + 3 * (ln - sz) + i, 0.125 * yv[u]);
        }
    }
}

// y += beta b x on the z = 0 / z = 1 faces (examples/vector3D.jl:193-262): one thread per node of a boundary plane gathers from
// its <= 4 faces; be = sum_g w_g |t1 x t2| N'N with the 2x2 rule, the same scalar mass for the three displacement components
__global__ void k_matfree_surface(const __grid_constant__ MfArgs A, int do_bottom, int do_top) {
    if (A.check_done && A.scal->done) return;
    const Lattice &L = A.L;
    const int n1 = L.n1;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)n1 * n1;
    int k;
    if (do_bottom && t < per) k = 0;
    else {
        if (do_bottom) t -= per;
        if (!do_top || t >= per) return;
        k = n1 - 1;
    }
    const int i = (int)(t % n1), j = (int)(t / n1);
    double acc[3] = {0, 0, 0};
    for (int fy = j - 1; fy <= j; ++fy)
        for (int fx = i - 1; fx <= i; ++fx) {
            if (fx < 0 || fy < 0 || fx >= L.ne || fy >= L.ne) continue;
            // face nodes in the reference's order (IEN[:, 1:4] / IEN[:, 5:8]): (0,0) (1,0) (1,1) (0,1)
            const int ox[4] = {0, 1, 1, 0}, oy[4] = {0, 0, 1, 1};
            double Xf[4][3], Uf[4][3];
            int a = 0;
            for (int b = 0; b < 4; ++b) {
                const int64_t nb = L.lnode(fx + ox[b], fy + oy[b], k);
                for (int c = 0; c < 3; ++c) {
                    Xf[b][c] = A.coords[3 * nb + c];
                    Uf[b][c] = A.x[3 * nb + c];
                }
                if (fx + ox[b] == i && fy + oy[b] == j) a = b;
            }
            for (int g = 0; g < 4; ++g) {
                const double xi = (ox[g] ? A.gpc : -A.gpc), eta = (oy[g] ? A.gpc : -A.gpc);  // Gauss points in the same corner order
                double N[4], dNx[4], dNe[4];
                for (int b = 0; b < 4; ++b) {
                    const double sx = ox[b] ? 1.0 : -1.0, sy = oy[b] ? 1.0 : -1.0;
                    N[b] = 0.25 * (1.0 + sx * xi) * (1.0 + sy * eta);
                    dNx[b] = 0.25 * sx * (1.0 + sy * eta);
                    dNe[b] = 0.25 * sy * (1.0 + sx * xi);
                }
                double t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0};
                for (int b = 0; b < 4; ++b)
                    for (int c = 0; c < 3; ++c) {
                        t1[c] += Xf[b][c] * dNx[b];
                        t2[c] += Xf[b][c] * dNe[b];
                    }
                const double nx = t1[1] * t2[2] - t1[2] * t2[1], ny = t1[2] * t2[0] - t1[0] * t2[2], nz = t1[0] * t2[1] - t1[1] * t2[0];
                const double w = sqrt(nx * nx + ny * ny + nz * nz) * N[a];
                for (int c = 0; c < 3; ++c) acc[c] += w * (N[0] * Uf[0][c] + N[1] * Uf[1][c] + N[2] * Uf[2][c] + N[3] * Uf[3][c]);
            }
        }
    double *dst = A.y + 3 * (L.lnode(i, j, k) - L.plane());
    for (int c = 0; c < 3; ++c) dst[c] += A.beta * acc[c];
}

// diag(K) without K (the Jacobi preconditioner of a CSR-less operator): per element and Gauss point, with g_a = J^-T grad N_a,
// K_aa[i][i] += w |det J| ((lam + mu) g_a,i^2 + mu |g_a|^2)   (= D11 g_i^2 + mu (|g|^2 - g_i^2), the diagonal of the assembly identity);
// same colouring / reduction stores as the operator
__global__ void __launch_bounds__(128) k_matfree_diag_color(const __grid_constant__ MfArgs A) {
    const Lattice &L = A.L;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)A.nx * A.ny * A.nz) return;
    const int tx = (int)(t % A.nx), ty = (int)((t / A.nx) % A.ny), tz = (int)(t / ((int64_t)A.nx * A.ny));
    const int ex = 2 * tx + A.cx, ey = 2 * ty + A.cy, ez = A.ez0 + 2 * tz;
    const int64_t n0 = L.lnode(ex, ey, ez);
    const int64_t sy = L.n1, sz = L.plane();
    double X[8][3], D[8][3];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t ln = n0 + (u & 1) + ((u >> 1) & 1) * sy + (u >> 2) * sz;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            X[u][c] = A.coords[3 * ln + c];
            D[u][c] = 0.0;
        }
    }
#pragma unroll 1
    for (int gp = 0; gp < 8; ++gp) {
        const double xi = (gp & 1) ? A.gpc : -A.gpc, eta = (gp & 2) ? A.gpc : -A.gpc, zeta = (gp & 4) ? A.gpc : -A.gpc;
        const double Xf[2] = {1.0 - xi, 1.0 + xi}, Yf[2] = {1.0 - eta, 1.0 + eta}, Zf[2] = {0.125 * (1.0 - zeta), 0.125 * (1.0 + zeta)};
        double d[8][3];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int ox = u & 1, oy = (u >> 1) & 1, oz = u >> 2;
            const double yz = Yf[oy] * Zf[oz], xz = Xf[ox] * Zf[oz], xy = 0.125 * Xf[ox] * Yf[oy];
            d[u][0] = ox ? yz : -yz;
            d[u][1] = oy ? xz : -xz;
            d[u][2] = oz ? xy : -xy;
        }
        double J[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) J[q] = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k) J[r * 3 + k] += X[u][r] * d[u][k];
        double adj[9];
        adj[0] = J[4] * J[8] - J[5] * J[7];
        adj[1] = J[2] * J[7] - J[1] * J[8];
        adj[2] = J[1] * J[5] - J[2] * J[4];
        adj[3] = J[5] * J[6] - J[3] * J[8];
        adj[4] = J[0] * J[8] - J[2] * J[6];
        adj[5] = J[2] * J[3] - J[0] * J[5];
        adj[6] = J[3] * J[7] - J[4] * J[6];
        adj[7] = J[1] * J[6] - J[0] * J[7];
        adj[8] = J[0] * J[4] - J[1] * J[3];
        const double det = J[0] * adj[0] + J[1] * adj[3] + J[2] * adj[6];
        const double f = 1.0 / fabs(det);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double g[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) g[c] = d[u][0] * adj[c] + d[u][1] * adj[3 + c] + d[u][2] * adj[6 + c];  // g_a det
            const double n2 = A.mu * (g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
#pragma unroll
            for (int c = 0; c < 3; ++c) D[u][c] += f * ((A.lam + A.mu) * g[c] * g[c] + n2);
        }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int k = ez + (u >> 2);
        if (k < L.k0 || k >= L.k1) continue;
        const int64_t ln = n0 + (u & 1) + ((u >> 1) & 1) * sy + (u >> 2) * sz;
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(A.y + 3 * (ln - sz) + c, D[u][c]);  // one contribution per address and launch
    }
}

// diag += beta diag(b): one thread per node of a boundary plane, the same face integrals as k_matfree_surface
__global__ void k_matfree_surface_diag(const __grid_constant__ MfArgs A, int do_bottom, int do_top) {
    const Lattice &L = A.L;
    const int n1 = L.n1;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)n1 * n1;
    int k;
    if (do_bottom && t < per) k = 0;
    else {
        if (do_bottom) t -= per;
        if (!do_top || t >= per) return;
        k = n1 - 1;
    }
    const int i = (int)(t % n1), j = (int)(t / n1);
    double acc = 0.0;
    for (int fy = j - 1; fy <= j; ++fy)
        for (int fx = i - 1; fx <= i; ++fx) {
            if (fx < 0 || fy < 0 || fx >= L.ne || fy >= L.ne) continue;
            const int ox[4] = {0, 1, 1, 0}, oy[4] = {0, 0, 1, 1};
            double Xf[4][3];
            int a = 0;
            for (int b = 0; b < 4; ++b) {
                const int64_t nb = L.lnode(fx + ox[b], fy + oy[b], k);
                for (int c = 0; c < 3; ++c) Xf[b][c] = A.coords[3 * nb + c];
                if (fx + ox[b] == i && fy + oy[b] == j) a = b;
            }
            for (int g = 0; g < 4; ++g) {
                const double xi = (ox[g] ? A.gpc : -A.gpc), eta = (oy[g] ? A.gpc : -A.gpc);
                double t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0}, Na = 0.0;
                for (int b = 0; b < 4; ++b) {
                    const double sx = ox[b] ? 1.0 : -1.0, sy = oy[b] ? 1.0 : -1.0;
                    const double N = 0.25 * (1.0 + sx * xi) * (1.0 + sy * eta), dNx = 0.25 * sx * (1.0 + sy * eta), dNe = 0.25 * sy * (1.0 + sx * xi);
                    if (b == a) Na = N;
                    for (int c = 0; c < 3; ++c) {
                        t1[c] += Xf[b][c] * dNx;
                        t2[c] += Xf[b][c] * dNe;
                    }
                }
                const double nx = t1[1] * t2[2] - t1[2] * t2[1], ny = t1[2] * t2[0] - t1[0] * t2[2], nz = t1[0] * t2[1] - t1[1] * t2[0];
                acc += sqrt(nx * nx + ny * ny + nz * nz) * Na * Na;
            }
        }
    double *dst = A.y + 3 * (L.lnode(i, j, k) - L.plane());
    for (int c = 0; c < 3; ++c) dst[c] += A.beta * acc;
}

// the boundary-plane rows of a slab read ghost planes written by the neighbours: wait for the halo flags (one thread)
__global__ void k_matfree_wait_halo(CommView cv, PcgScalars *scal, int check_done, unsigned long long halo_need) {
    if (check_done && scal->done) return;
    const unsigned long long need = halo_need ? halo_need : scal->it + 1;
    const unsigned long long t0 = global_timer_ns();
    unsigned spins = 0;
    if (cv.rank > 0)
        while (ld_acquire_sys(&cv.self->hflag[0]) < need)
            if (++spins > (1u << 28)) __trap();
    if (cv.rank < cv.nranks - 1)
        while (ld_acquire_sys(&cv.self->hflag[1]) < need)
            if (++spins > (1u << 28)) __trap();
    scal->t_wait_halo += global_timer_ns() - t0;  // wait accounting (smfem_pcg_wait_stats)
}

}  // namespace

void matfree_apply(smfem_ctx *ctx, smfem_matrix *K, const double *x, double *y, bool halo, bool check_done,
                   unsigned long long halo_need) {
    smfem_mesh *mesh = K->mf_mesh;
    REQUIRE(mesh && mesh->structured && K->structured && K->ndim == 3 && K->nDof == 3, SMFEM_ERR_UNSUPPORTED,
            "matrix-free operator: structured 3-D hex lattice with nDof = 3 only");
    REQUIRE(K->mat_known, SMFEM_ERR_INVALID, "matrix-free operator: material unknown (assemble K first)");
    const Lattice &L = K->lat;
    MfArgs A;
    A.L = L;
    A.coords = mesh->coords;
    A.x = x;
    A.y = y;
    const double f = K->Young / ((1 + K->nu) * (1 - 2 * K->nu));  // src/fem.jl:230
    A.lam = K->nu * f;
    A.mu = (1 - 2 * K->nu) / 2 * f;
    A.beta = K->beta_total;
    double xi[2], w[2];
    smfem_host_gauss(-1, 1, 2, xi, w);
    A.gpc = xi[1];
    A.scal = K->scal;
    A.check_done = check_done ? 1 : 0;
    if (halo && ctx->nranks > 1) LAUNCH(ctx, k_matfree_wait_halo, 1, 1, 0, K->comm, K->scal, A.check_done, halo_need);
    CUDA_CHECK(cudaMemsetAsync(y, 0, sizeof(double) * K->nrows_l, ctx->stream));
    const int l0 = L.k0 > 0 ? L.k0 - 1 : 0, l1 = L.k1 - 1 < L.ne - 1 ? L.k1 - 1 : L.ne - 1;  // element layers [l0, l1] of this rank
    const char *ver = std::getenv("SMFEM_MATFREE");  // "v1": the first (corner-form) element kernel, kept for A/B timing
    const bool v1 = ver && ver[0] == 'v' && ver[1] == '1';
    const int cfg = (ver && ver[0] == 'c') ? std::atoi(ver + 1) : 0;  // "c1".."c3": launch-shape experiments of the monomial kernel
    for (int c = 0; c < 8; ++c) {
        A.cx = c & 1;
        A.cy = (c >> 1) & 1;
        A.cz = c >> 2;
        A.nx = (L.ne - A.cx + 1) / 2;
        A.ny = (L.ne - A.cy + 1) / 2;
        // layers of parity cz (GLOBAL parity, so that the fold order does not depend on the partition) inside [l0, l1]
        A.ez0 = l0 + (((l0 & 1) != A.cz) ? 1 : 0);
        A.nz = A.ez0 > l1 ? 0 : (l1 - A.ez0) / 2 + 1;
        const int64_t n = (int64_t)A.nx * A.ny * A.nz;
        if (n <= 0) continue;
        if (v1) LAUNCH(ctx, k_matfree_color, (unsigned)((n + 127) / 128), 128, 0, A);
        else if (cfg == 1) LAUNCH(ctx, (k_matfree_color2<128, 4>), (unsigned)((n + 127) / 128), 128, 0, A);
        else if (cfg == 2) LAUNCH(ctx, (k_matfree_color2<64, 6>), (unsigned)((n + 63) / 64), 64, 0, A);
        else if (cfg == 3) LAUNCH(ctx, (k_matfree_color2<64, 8>), (unsigned)((n + 63) / 64), 64, 0, A);
        else LAUNCH(ctx, (k_matfree_color2<128, 3>), (unsigned)((n + 127) / 128), 128, 0, A);
    }
    const int bot = (L.k0 == 0), top = (L.k1 == L.n1);
    if (A.beta != 0.0 && (bot || top)) {
        const int64_t n = (int64_t)(bot + top) * L.n1 * L.n1;
        LAUNCH(ctx, k_matfree_surface, (unsigned)((n + 127) / 128), 128, 0, A, bot, top);
    }
}

// ---- the operator WITHOUT an assembled K (smfem_matfree_operator): only the diagonal is computed, for the Jacobi preconditioner
static MfArgs matfree_args(smfem_matrix *K, smfem_mesh *mesh, double *y) {
    MfArgs A;
    A.L = K->lat;
    A.coords = mesh->coords;
    A.x = nullptr;
    A.y = y;
    const double f = K->Young / ((1 + K->nu) * (1 - 2 * K->nu));  // src/fem.jl:230
    A.lam = K->nu * f;
    A.mu = (1 - 2 * K->nu) / 2 * f;
    A.beta = 0.0;
    double xi[2], w[2];
    smfem_host_gauss(-1, 1, 2, xi, w);
    A.gpc = xi[1];
    A.scal = nullptr;
    A.check_done = 0;
    return A;
}

void matfree_diag(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh) {
    const Lattice &L = K->lat;
    if (!K->diag) K->diag = dev_alloc<double>(K->nrows_l);
    CUDA_CHECK(cudaMemsetAsync(K->diag, 0, sizeof(double) * K->nrows_l, ctx->stream));
    MfArgs A = matfree_args(K, mesh, K->diag);
    const int l0 = L.k0 > 0 ? L.k0 - 1 : 0, l1 = L.k1 - 1 < L.ne - 1 ? L.k1 - 1 : L.ne - 1;
    for (int c = 0; c < 8; ++c) {
        A.cx = c & 1;
        A.cy = (c >> 1) & 1;
        A.cz = c >> 2;
        A.nx = (L.ne - A.cx + 1) / 2;
        A.ny = (L.ne - A.cy + 1) / 2;
        A.ez0 = l0 + (((l0 & 1) != A.cz) ? 1 : 0);
        A.nz = A.ez0 > l1 ? 0 : (l1 - A.ez0) / 2 + 1;
        const int64_t n = (int64_t)A.nx * A.ny * A.nz;
        if (n <= 0) continue;
        LAUNCH(ctx, k_matfree_diag_color, (unsigned)((n + 127) / 128), 128, 0, A);
    }
}

// K_bar = K + beta b for a CSR-less operator: only the diagonal and the operator's beta change
void matfree_add_surface(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, double beta) {
    const Lattice &L = K->lat;
    const int bot = (L.k0 == 0), top = (L.k1 == L.n1);
    if (beta != 0.0 && (bot || top)) {
        MfArgs A = matfree_args(K, mesh, K->diag);
        A.beta = beta;
        const int64_t n = (int64_t)(bot + top) * L.n1 * L.n1;
        LAUNCH(ctx, k_matfree_surface_diag, (unsigned)((n + 127) / 128), 128, 0, A, bot, top);
    }
    K->beta_total += beta;
    K->gmg_dirty = true;
}
