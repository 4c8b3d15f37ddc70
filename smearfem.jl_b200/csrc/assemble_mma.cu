// K1 + K3 for the structured hex lattice, nDof = 3, on the fp64 tensor-core path (DMMA.8x8x4): tiled, atomic-free,
// deterministic assembly.   replaces src/fem.jl:179-249 (element loop + COO scatter) and the value side of sparse(E,J,V) (:253)
//
// Identity used:  B'DB = lam G + mu G' + mu tr(G) I  per node pair, with  G_ab = sum_gp g_a g_b',
// g_b = sqrt(w |det J|) dN_b J^-1 (src/fem.jl:192-196).  Per element G (24 x 24) = g g' with g 24 x 8 (dofs x Gauss points):
// a rank-8 update, i.e. 3 x 3 tiles of an 8 x 8 x 8 product = 18 DMMA.8x8x4.  Lane (a = lane/4, q = lane%4) supplies
// g_a(gp q) and g_a(gp q+4) as BOTH the A and the B fragment (the product is symmetric), and receives the complete 3x3
// blocks G_ab for b = 2q, 2q+1: no operand ever goes through shared memory (the scalar kernel in assemble_tile.cu needs
// 27 LDS.64 per 72 DFMA and is bound by shared-memory wavefronts; DMMA runs at the DFMA rate - 37 TFLOP/s measured,
// tools/microbench/dmma.cu - so this trades nothing on the fp64 pipe).
//
// A CTA owns a tile of TX x TY node columns and marches up the z planes of its chunk.  Per element layer:
//   phase 1a (thread = element x Gauss point of the (TX+1) x (TY+1) footprint): Jacobian, adj(J) sign(det) sqrt(w/|det|)
//            -> 9 doubles per (element, gp) in shared memory.
//   rounds   the footprint elements are 4-coloured by the parity of (fx, fy); elements of one colour share no node, so the
//            warps of a round (one element each) add into the per-node staging area without conflicts; __syncthreads
//            between colours fixes the fold order (bit-reproducible).  Per element: phase 1b (own gradients from the
//            9 doubles: 18 FMA), 18 DMMA, material in registers, flush of the 2 blocks x 9 entries per lane.
//            Staging: block d = (dx,dy,dz) of a node lives in slot 3dx + 2dy + 9dz + 14 (29 slots x 9 doubles); with node
//            stride 263 and row stride == 10 (mod 16) the 16 lanes of a half-warp hit 16 distinct banks (brute-force
//            search over linear slot maps; the CSR-ordered layout [row][neighbour][j] cannot be made conflict-free).
//            An entry is plain-stored by the first element that reaches it (known in closed form), added to by the rest.
//   output   plane k is complete after layer k: lanes 0..26 walk the 3 x cz sections of 27 consecutive CSR entries of a
//            row: contiguous 216-byte stores of values (and column indices in the fused assembly); diagonal on the way.
#include <cstdlib>
#include <string>

#include "smfem_internal.cuh"
#include "tile_args.cuh"

namespace {

constexpr int SLOT_LX = 3, SLOT_LY = 2, SLOT_LZ = 9, SLOT_C0 = 14;

template <int TX_, int TY_, int NW_>
struct MTile {
    static constexpr int TX = TX_, TY = TY_, NW = NW_, NTH = NW_ * 32;
    static constexpr int EX = TX + 1, EY = TY + 1, NEL = EX * EY;
    static constexpr int PX = TX + 2, PY = TY + 2, PLANE = PX * PY * 3;  // node-plane coordinate buffer (with halo)
    static constexpr int SN = 263;                                       // node stride: SN - 9*3 == 12 (mod 16)
    static constexpr int SR = TX * SN + ((10 - (TX * SN) % 16) + 16) % 16;  // row stride: SR - 9*2 == 8 (mod 16)
    static constexpr int SP = TY * SR;                                   // staging plane (two of them: plane parity)
    static constexpr int ADJ = 10;                                       // doubles per (element, gp): 9 + pad (16-byte rows)
    static constexpr size_t SMEM_BYTES = sizeof(double) * (2 * SP + NEL * 8 * ADJ + 4 * PLANE);
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// old value of a staging entry; lanes that touch it first do not load
__device__ __forceinline__ double lds_unless(const double *p, bool first) {
    double v;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
    asm volatile(
        "{ .reg .pred pp; setp.eq.u32 pp, %2, 0; mov.f64 %0, 0d0000000000000000; @pp ld.shared.f64 %0, [%1]; }"
        : "=d"(v)
        : "r"(addr), "r"((unsigned)first)
        : "memory");
    return v;
}

template <class T, int MINB>
__global__ void __launch_bounds__(T::NTH, MINB) k_values_mma(const __grid_constant__ TileArgs A) {
    constexpr int TX = T::TX, TY = T::TY, NTH = T::NTH, EX = T::EX, NEL = T::NEL, ADJ = T::ADJ;
    extern __shared__ double smem[];
    double *stage = smem;                         // [2 plane parities][TY rows, stride SR][TX nodes, stride SN][29 slots][9]
    double *s_adj = smem + 2 * T::SP;             // [NEL][8 gp][ADJ]
    double *s_xyz = s_adj + NEL * 8 * ADJ;        // [4][PLANE] node-plane coordinate ring
    const Lattice &L = A.L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    int bid = blockIdx.x;
    const int tix = bid % A.tiles_x;
    bid /= A.tiles_x;
    const int tiy = bid % A.tiles_y;
    const int chunk_id = bid / A.tiles_y;
    const int X0 = tix * TX, Y0 = tiy * TY;
    const int zs = L.k0 + A.zb[chunk_id], ze = L.k0 + A.zb[chunk_id + 1];

    // ---- lane identity inside an element task: row node a (natural order ox + 2 oy + 4 oz), column nodes b = 2q, 2q+1
    const int a = lane >> 2, q = lane & 3;
    const int oxa = a & 1, oya = (a >> 1) & 1, oza = a >> 2;
    const int oyb = q & 1, ozb = q >> 1;  // oxb = i (the register index of the C fragment)
    const int dy = oyb - oya, dz = ozb - oza;
    int off[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) off[i] = 9 * (SLOT_LX * (i - oxa) + SLOT_LY * dy + SLOT_LZ * dz + SLOT_C0);
    // reference gradients of shape function a at this lane's two Gauss points g = q, q + 4 (reference order, src/fem.jl:174-176)
    //   dN_a/dxi = sx (1 + sy eta)(1 + sz zeta)/8   (src/fem.jl:63)
    double dN[2][3];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int g = q + 4 * u;
        const double xi = A.gp[g][0], eta = A.gp[g][1], zeta = A.gp[g][2];
        const double Xf = oxa ? 1.0 + xi : 1.0 - xi, Yf = oya ? 1.0 + eta : 1.0 - eta, Zf = 0.125 * (oza ? 1.0 + zeta : 1.0 - zeta);
        const double yz = Yf * Zf, xz = Xf * Zf, xy = 0.125 * Xf * Yf;
        dN[u][0] = oxa ? yz : -yz;
        dN[u][1] = oya ? xz : -xz;
        dN[u][2] = oza ? xy : -xy;
    }

    // closed-form CSR row starts: see k_struct_rowptr
    const int64_t S1 = 3 * (int64_t)L.n1 - 2;
    auto pre1 = [](int i) -> int64_t { return i == 0 ? 0 : 3 * (int64_t)i - 1; };
    const int64_t pairs_base = pre1(L.k0) * S1 * S1;

    // streamed coordinates (smfem_assemble_system): see k_values_tile
    int ready_upto = A.ready ? 0 : 0x7fffffff;
    auto wait_plane = [&](int p) {
        const int need = min(p + 2, min(L.k1 + 1, L.n1));
        unsigned spins = 0;
        while (ready_upto < need) {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(ready_upto) : "l"(A.ready) : "memory");
            if (++spins > (1u << 25)) __trap();
        }
    };

    // ---- phase 1a: adj(J) * sign(det) sqrt(w/|det|) per (element, Gauss point) of element layer `layer`
    auto phase1a = [&](int layer) {
        const double *P0 = s_xyz + (layer & 3) * T::PLANE, *P1 = s_xyz + ((layer + 1) & 3) * T::PLANE;
        for (int t = tid; t < NEL * 8; t += NTH) {
            const int e = t >> 3, g = t & 7;
            const int fy = e / EX, fx = e - fy * EX;
            const int ex = X0 - 1 + fx, ey = Y0 - 1 + fy;
            if (ex < 0 || ey < 0 || ex >= L.ne || ey >= L.ne) continue;
            double Xn[8][3];  // nodes in natural order u = ox + 2 oy + 4 oz
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int ox = u & 1, oy = (u >> 1) & 1, oz = u >> 2;
                const double *p = (oz ? P1 : P0) + 3 * ((fy + oy) * T::PX + fx + ox);
                Xn[u][0] = p[0];
                Xn[u][1] = p[1];
                Xn[u][2] = p[2];
            }
            const double xi = A.gp[g][0], eta = A.gp[g][1], zeta = A.gp[g][2];
            const double Xf[2] = {1.0 - xi, 1.0 + xi}, Yf[2] = {1.0 - eta, 1.0 + eta}, Zf[2] = {0.125 * (1.0 - zeta), 0.125 * (1.0 + zeta)};
            double J[9];  // J[r*3+k] = d x_r / d xi_k   (Jac = coords*dN, src/fem.jl:192), from the 12 edge differences
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double jx = 0, jy = 0, jz = 0;
#pragma unroll
                for (int t2 = 0; t2 < 4; ++t2) {
                    const int o1 = t2 & 1, o2 = t2 >> 1;
                    jx += (Yf[o1] * Zf[o2]) * (Xn[1 + 2 * o1 + 4 * o2][r] - Xn[2 * o1 + 4 * o2][r]);
                    jy += (Xf[o1] * Zf[o2]) * (Xn[o1 + 2 + 4 * o2][r] - Xn[o1 + 4 * o2][r]);
                    jz += (0.125 * Xf[o1] * Yf[o2]) * (Xn[o1 + 2 * o2 + 4][r] - Xn[o1 + 2 * o2][r]);
                }
                J[r * 3 + 0] = jx;
                J[r * 3 + 1] = jy;
                J[r * 3 + 2] = jz;
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
            const double sc = copysign(rsqrt(fabs(det)), det) * A.sw[g];
            double2 *dst = reinterpret_cast<double2 *>(s_adj + t * ADJ);
            dst[0] = make_double2(adj[0] * sc, adj[1] * sc);
            dst[1] = make_double2(adj[2] * sc, adj[3] * sc);
            dst[2] = make_double2(adj[4] * sc, adj[5] * sc);
            dst[3] = make_double2(adj[6] * sc, adj[7] * sc);
            dst[4] = make_double2(adj[8] * sc, 0.0);
        }
    };

    // ---- one footprint element (whole warp): phase 1b, 18 DMMA, material, flush into the staging planes
    auto element_task = [&](int fx, int fy, int layer, bool fl_bot, bool fl_top) {
        const int ex = X0 - 1 + fx, ey = Y0 - 1 + fy;
        if (ex < 0 || ey < 0 || ex >= L.ne || ey >= L.ne) return;  // warp-uniform
        double C[3][3][2];
#pragma unroll
        for (int ca = 0; ca < 3; ++ca)
#pragma unroll
            for (int cb = 0; cb < 3; ++cb) C[ca][cb][0] = C[ca][cb][1] = 0.0;
        if (!(A.skip & 2)) {
            double g[3][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const double *ad = s_adj + ((fy * EX + fx) * 8 + q + 4 * u) * ADJ;
                double m[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) m[t] = ad[t];
#pragma unroll
                for (int c = 0; c < 3; ++c) g[c][u] = dN[u][0] * m[c] + dN[u][1] * m[3 + c] + dN[u][2] * m[6 + c];
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                    for (int cb = 0; cb < 3; ++cb) dmma884(C[ca][cb][0], C[ca][cb][1], g[ca][u], g[cb][u]);
        }
        const int tx = fx - 1 + oxa, ty = fy - 1 + oya;
        const bool row_ok = tx >= 0 && ty >= 0 && tx < TX && ty < TY && X0 + tx < L.n1 && Y0 + ty < L.n1 && (oza ? fl_top : fl_bot);
        if (row_ok && !(A.skip & 4)) {
            // the other elements of this layer that hold the same node pair come earlier iff their footprint parity is lower
            const int exo = ex + (oxa ? 1 : -1), eyo = ey + (oya ? 1 : -1);
            const bool later_x = (fx & 1) && exo >= 0 && exo < L.ne;  // an x-neighbour element exists and is of an earlier colour
            const bool later_y = (fy & 1) && eyo >= 0 && eyo < L.ne;
            const bool later_z = oza == 0 && ozb == 0 && layer > 0;  // in-plane blocks of the lower plane: layer-1 was there first
            double *np = stage + ((layer + oza) & 1) * T::SP + ty * T::SR + tx * T::SN;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool first = !((i == oxa && later_x) || (dy == 0 && later_y) || later_z);
                double Kv[9];
                const double tr = C[0][0][i] + C[1][1][i] + C[2][2][i];
                const double mtr = A.mat.mu * tr, dm = A.mat.d11 - A.mat.mu;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const double gij = C[c][j][i], gji = C[j][c][i];
                        Kv[c * 3 + j] = (c == j) ? fma(dm, gij, mtr) : fma(A.mat.lam, gij, A.mat.mu * gji);
                    }
                double *p = np + off[i];
                double old[9];
#pragma unroll
                for (int m = 0; m < 9; ++m) old[m] = lds_unless(p + m, first);
#pragma unroll
                for (int m = 0; m < 9; ++m) p[m] = old[m] + Kv[m];
            }
        }
    };

    auto do_layer = [&](int layer, bool fl_bot, bool fl_top) {
        if (!(A.skip & 1)) phase1a(layer);
        __syncthreads();
#pragma unroll 1
        for (int cls = 0; cls < 4; ++cls) {
            const int cxp = cls & 1, cyp = cls >> 1;
            const int nxc = (T::EX - cxp + 1) / 2, nyc = (T::EY - cyp + 1) / 2;
#pragma unroll 1
            for (int idx = warp; idx < nxc * nyc; idx += T::NW) {
                const int iy = idx / nxc, ix = idx - iy * nxc;
                element_task(cxp + 2 * ix, cyp + 2 * iy, layer, fl_bot, fl_top);
            }
            __syncthreads();
        }
    };

    // ---- output of the completed node plane k: lanes 0..26 own (dx, dy, j) and walk the dz sections of the three rows
    const int o_r = lane / 3, o_j = lane - 3 * o_r;
    const int o_dy = o_r / 3 - 1, o_dx = o_r - 3 * (o_r / 3) - 1;                       // interior nodes (27 neighbours)
    const int o_src = 9 * (SLOT_LX * o_dx + SLOT_LY * o_dy + SLOT_C0 - SLOT_LZ) + o_j;  // staging offset of (dz = -1, c = 0)
    const int o_col = 3 * (o_dx + o_dy * L.n1) + o_j;
    auto output = [&](int k) {
        if (A.skip & 8) return;
        const double *sp = stage + (k & 1) * T::SP;
        const int lowz = k > 0, cz = 1 + lowz + (k < L.n1 - 1);
        const int64_t zbase = pre1(k) * S1 * S1 - pairs_base;
#pragma unroll 1
        for (int n = warp; n < TX * TY; n += T::NW) {
            const int ty = n / TX, tx = n - ty * TX;
            const int jx = X0 + tx, jy = Y0 + ty;
            if (jx >= L.n1 || jy >= L.n1) continue;
            const int lowx = jx > 0, lowy = jy > 0;
            const int cx = 1 + lowx + (jx < L.n1 - 1), cy = 1 + lowy + (jy < L.n1 - 1);
            const int64_t base = 9 * (zbase + (int64_t)cz * (pre1(jy) * S1 + (int64_t)cy * pre1(jx)));
            const int64_t row = (((int64_t)(k - L.k0) * L.n1 + jy) * L.n1 + jx) * 3;
            const double *nsrc = sp + ty * T::SR + tx * T::SN;
            if (cx * cy * cz == 27) {  // interior node: everything but the two base addresses is a lane constant
                if (lane < 27) {
                    const double *src = nsrc + o_src;
                    double *vp = A.val + base + lane;
                    double v[9];
#pragma unroll
                    for (int rz = 0; rz < 3; ++rz)
#pragma unroll
                        for (int c = 0; c < 3; ++c) v[rz * 3 + c] = src[9 * SLOT_LZ * rz + 3 * c];
#pragma unroll
                    for (int rz = 0; rz < 3; ++rz)
#pragma unroll
                        for (int c = 0; c < 3; ++c) vp[c * 81 + rz * 27] = v[rz * 3 + c];
                    if (A.colind) {
                        int32_t *cp = A.colind + base + lane;
                        const int32_t col0 = (int32_t)(L.lnode(jx, jy, k - 1) * 3) + o_col;
                        const int32_t zstep = (int32_t)(3 * L.plane());
#pragma unroll
                        for (int rz = 0; rz < 3; ++rz)
#pragma unroll
                            for (int c = 0; c < 3; ++c) cp[c * 81 + rz * 27] = col0 + rz * zstep;
                    }
                    if (o_dx == 0 && o_dy == 0) A.diag[row + o_j] = o_j == 0 ? v[3] : (o_j == 1 ? v[4] : v[5]);  // dz = 0, c = j
                }
                continue;
            }
            const int SL = 3 * cx * cy, TR = SL * cz;
            if (lane >= SL) continue;
            const int r = lane / 3, j = lane - 3 * r;
            const int ry = r / cx, rx = r - ry * cx;
            const int dxn = rx - lowx, dyn = ry - lowy;
            const double *src = nsrc + 9 * (SLOT_LX * dxn + SLOT_LY * dyn + SLOT_C0) + j;
            const bool on_diag_col = dxn == 0 && dyn == 0;
            for (int rz = 0; rz < cz; ++rz) {
                const int dzn = rz - lowz;
                const int32_t col = (int32_t)(L.lnode(jx + dxn, jy + dyn, k + dzn) * 3 + j);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double v = src[9 * SLOT_LZ * dzn + 3 * c];
                    const int64_t pos = base + (int64_t)c * TR + rz * SL + lane;
                    A.val[pos] = v;
                    if (A.colind) A.colind[pos] = col;
                    if (on_diag_col && dzn == 0 && c == j) A.diag[row + c] = v;
                }
            }
        }
    };

    wait_plane(zs + 1);
    stage_plane<T>(A, s_xyz, zs - 1, X0, Y0);
    stage_plane<T>(A, s_xyz, zs, X0, Y0);
    stage_plane<T>(A, s_xyz, zs + 1, X0, Y0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (zs >= 1) do_layer(zs - 1, false, true);  // the layer below the chunk: its upper node rows belong to plane zs

    for (int k = zs; k < ze; ++k) {
        if (k + 2 < L.n1 && k + 2 <= L.k1) wait_plane(k + 2);
        stage_plane<T>(A, s_xyz, k + 2, X0, Y0);  // lands during this layer; ring slot (k+2)&3 is free
        if (k < L.ne) do_layer(k, true, k + 1 < ze);
        output(k);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();  // staging plane k&1 is reused by layer k+1; coordinate plane k+2 has landed
    }
}

template <class T, int MINB>
void launch_mma(smfem_ctx *ctx, TileArgs &A, int nown) {
    static std::atomic<unsigned long long> attr_set{0};
    if (first_use_on_device(attr_set)) {
        CUDA_CHECK(cudaFuncSetAttribute(k_values_mma<T, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES));
    }
    A.tiles_x = (A.L.n1 + T::TX - 1) / T::TX;
    A.tiles_y = (A.L.n1 + T::TY - 1) / T::TY;
    const int ntiles = A.tiles_x * A.tiles_y;
    const std::vector<int> len = plan_chunks(ntiles, nown, ctx->sms * MINB);
    A.nchunks = (int)len.size();
    A.zb[0] = 0;
    for (int c = 0; c < A.nchunks; ++c) A.zb[c + 1] = A.zb[c] + len[c];
    const unsigned grid = (unsigned)(ntiles * A.nchunks);
    const int slot = (int)(ctx->asm_count % smfem_ctx::ASM_RING);
    if (!ctx->asm_ev[2 * slot]) {
        CUDA_CHECK(cudaEventCreate(&ctx->asm_ev[2 * slot]));
        CUDA_CHECK(cudaEventCreate(&ctx->asm_ev[2 * slot + 1]));
    }
    CUDA_CHECK(cudaEventRecord(ctx->asm_ev[2 * slot], ctx->stream));
    LAUNCH(ctx, (k_values_mma<T, MINB>), grid, T::NTH, T::SMEM_BYTES, A);
    CUDA_CHECK(cudaEventRecord(ctx->asm_ev[2 * slot + 1], ctx->stream));
    ctx->asm_count++;
}

}  // namespace

bool values_assemble_mma(smfem_ctx *ctx, TileArgs &A, int nown) {
    const char *e = std::getenv("SMFEM_TILE");
    const std::string v = e ? e : "";
    if (v == "mma84") launch_mma<MTile<8, 4, 16>, 1>(ctx, A, nown);
    else if (v == "mma75") launch_mma<MTile<7, 5, 12>, 1>(ctx, A, nown);
    else if (v == "mma44") launch_mma<MTile<4, 4, 8>, 2>(ctx, A, nown);
    else if (v.rfind("mma", 0) == 0) throw SmfemError(SMFEM_ERR_INVALID, "SMFEM_TILE=" + v + ": unknown DMMA tile (mma75, mma84, mma44)");
    else return false;
    return true;
}
