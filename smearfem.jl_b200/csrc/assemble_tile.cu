// K1 + K3 for the structured hex lattice, nDof = 3: tiled, atomic-free, deterministic gather assembly.
//   replaces src/fem.jl:179-249 (element loop + COO scatter) and the value side of sparse(E,J,V) (:253)
//
// A CTA owns a tile of TX x TY node columns and marches up the z planes of its chunk.
//   phase 1  (thread = element x Gauss point of the (TX+1)x(TY+1) element footprint of one layer):
//            Jacobian, inverse, |det| and the physical gradients of the 8 shape functions, scaled by
//            sqrt(w):  g_b = sqrt(w_gp |det J|) * dN_b J^-1   -> shared memory ring of two layers.
//            Every element layer is computed once per tile column (redundancy (TX+1)(TY+1)/(TX TY)).
//   phase 2  (thread = node x element slot s in 0..7): for the element at offset -s of the node (local
//            node a(s)) accumulate the 8 blocks  G_ab = sum_gp g_a g_b'  (72 fp64 accumulators, 9 FMA
//            per 6 shared loads... see DESIGN.md) in registers.
//   combine  8 conflict-free rounds (round b: the 8 slot-threads of a node hit 8 distinct neighbour
//            blocks) add the G blocks into a per-node staging area [27 neighbours][3x3] in shared memory.
//   output   lanes 0..26 of the warp own one neighbour each: apply the material once,
//            K_ab = lam G + mu G' + mu tr(G) I   (D(1,1) on the diagonal, src/fem.jl:230), and store the
//            node's three CSR rows; every stored entry of K is written exactly once (no memset, no
//            atomics, fold order fixed -> bit-reproducible).
#include <cmath>
#include <cstdlib>
#include <map>
#include <queue>
#include <tuple>
#include <mutex>
#include <vector>

#include "smfem_internal.cuh"
#include "tile_args.cuh"

namespace {

// Staging area of a node: 27 neighbour blocks of 3x3.  Block d = (dx,dy,dz) lives in slot c_slot[q] (q = (dz+1)*9 +
// (dy+1)*3 + (dx+1)), slot = r + 8k where the residue r in 0..7 is DISTINCT inside every 2x2x2 sub-cube of the 3x3x3
// neighbourhood (found by backtracking; at most 4 blocks share a residue -> 32 slots).  In a combine round the 8 slot
// threads of a node hit such a sub-cube, so their 8-byte banks 9*slot + m (mod 16) are pairwise distinct and, with the
// node stride == 8 (mod 16), disjoint from those of the second node of the half-warp: the read-modify-write rounds are
// bank-conflict free (they were exactly 2-way conflicted with the plain [27][9] layout: 148 vs 74 wavefronts per node).
// (the k of same-residue blocks is chosen so that the 27 output lanes also read with few conflicts: 27 vs ideal 18 wavefronts)
__constant__ unsigned char c_slot[27] = {5, 7, 4, 6, 3, 14, 2, 15, 1, 12, 10, 9, 17, 0, 13, 21, 20, 18, 29, 23, 28, 22, 11, 30, 26, 31, 25};
constexpr int STAGE_NODE = 32 * 9 + 8;  // 296 doubles per node (== 8 mod 16)

template <int TX_, int TY_>
struct Tile {
    static constexpr int TX = TX_, TY = TY_, NTH = TX_ * TY_ * 8;
    static constexpr int EX = TX + 1, EY = TY + 1, NEL = EX * EY;  // footprint elements per layer (45 for 8x4)
    // doubles per layer in the ring; +3: the two layers a half-warp reads (sz = 0/1) land in disjoint banks
    // (ncu: 2x excess shared wavefronts without the pad)
    static constexpr int PAD = (EX == 9) ? 3 : 8;  // 8x4 tile: banks {6,7,8,15,0,1}+3; 4x4 tile: {10,11,12,15,0,1}+8
    // element stride of the ring: 28 instead of 25 for the 4x4 tile makes the slot-dependent g_a loads 2-way instead of
    // 3-way conflicted (brute-force search over stride / pad; the g_b loads are conflict-free either way)
    static constexpr int NELP = (EX == 5) ? 28 : NEL;
    static constexpr int LAYER = 8 * 8 * 3 * NELP + PAD;
    static constexpr int PX = TX + 2, PY = TY + 2, PLANE = PX * PY * 3;  // node-plane coordinate buffer (with halo)
    // the staging area aliases the ring slot of the element layer that is dead after the main loop when it fits (4x4)
    static constexpr bool ALIAS = TX * TY * STAGE_NODE <= LAYER;
    // per-warp buffer for the CSR-ordered column indices of a run (<= 972 + 3 alignment entries), only where it still
    // allows 2 CTAs/SM (4x4 tile); the 8x4 tile reuses the staging region and waits for the bulk store instead
    static constexpr int COLBUF = ALIAS ? 976 : 0;  // int32 per warp
    static constexpr size_t SMEM_BYTES =
        sizeof(double) * (2 * LAYER + (ALIAS ? 0 : TX * TY * STAGE_NODE) + 8 * 8 * 3 + 8 + 4 * PLANE) + sizeof(int32_t) * COLBUF * (NTH / 32);
};

// element layer `layer` of the footprint -> ring slot.  g_b = dN_b adj(J) * sign(det) * sqrt(wp / |det|)
// ( = sqrt(wp |det|) * dN_b J^-1, src/fem.jl:192-196 ).  Register-only formulation of the Q1 gradients:
//   dN_b/dxi = sx_b (1 + sy_b eta)(1 + sz_b zeta)/8  (src/fem.jl:63)  ->  12 products per Gauss point, no table loads;
//   J(:,xi)  = sum over the 4 xi-edges of  YZ[oy][oz] * (x_{b+} - x_{b-})   (edge differences shared by the Gauss points);
//   the unscaled gradients dN adj(J) are formed while the rsqrt of |det| is in flight, then scaled.
template <class T>
__device__ __forceinline__ void phase1(const TileArgs &A, const double *s_gp, const double *s_sw, const double *s_xyz, double *S,
                                       int layer, int X0, int Y0) {
    constexpr int NEL = T::NEL, NELP = T::NELP, EX = T::EX, LAYER = T::LAYER, NTH = T::NTH;
    const Lattice &L = A.L;
    double *dst = S + (layer & 1) * LAYER;
    const double *P0 = s_xyz + (layer & 3) * T::PLANE, *P1 = s_xyz + ((layer + 1) & 3) * T::PLANE;
    // task = (element, pair of Gauss points).  Tasks are laid out in groups of RND lanes per Gauss-point pair (RND = NEL
    // rounded up to a warp multiple) so that the lanes of a warp store to consecutive e of ONE (gp, b, c) row of the
    // ring: no wrap-around bank conflicts (8x4 tile: 45 elements -> 64 lanes per pair, one pass of 256 threads).
    constexpr int RND = (NEL <= 25) ? NEL : (NEL + 31) / 32 * 32;  // 4x4 tile: dense mapping measured faster (3.1 vs 4 busy warps)
    for (int q = threadIdx.x; q < 4 * RND; q += NTH) {
        const int gpp = q / RND, e = q - gpp * RND;
        if (e >= NEL) continue;
        const int fy = e / EX, fx = e - fy * EX;
        const int ex = X0 - 1 + fx, ey = Y0 - 1 + fy;
        if (ex < 0 || ey < 0 || ex >= L.ne || ey >= L.ne) continue;
        // nodes in natural order u = ox + 2 oy + 4 oz
        double Xn[8][3];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int ox = u & 1, oy = (u >> 1) & 1, oz = u >> 2;
            const double *p = (oz ? P1 : P0) + 3 * ((fy + oy) * T::PX + fx + ox);
            Xn[u][0] = p[0];
            Xn[u][1] = p[1];
            Xn[u][2] = p[2];
        }
        double Ex[4][3], Ey[4][3], Ez[4][3];  // edge differences along xi / eta / zeta
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int o1 = t & 1, o2 = t >> 1;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                Ex[t][r] = Xn[1 + 2 * o1 + 4 * o2][r] - Xn[2 * o1 + 4 * o2][r];  // t = (oy, oz)
                Ey[t][r] = Xn[o1 + 2 + 4 * o2][r] - Xn[o1 + 4 * o2][r];          // t = (ox, oz)
                Ez[t][r] = Xn[o1 + 2 * o2 + 4][r] - Xn[o1 + 2 * o2][r];          // t = (ox, oy)
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int gp = 2 * gpp + h;
            const double xi = s_gp[3 * gp], eta = s_gp[3 * gp + 1], zeta = s_gp[3 * gp + 2];
            const double Xf[2] = {1.0 - xi, 1.0 + xi}, Yf[2] = {1.0 - eta, 1.0 + eta}, Zf[2] = {0.125 * (1.0 - zeta), 0.125 * (1.0 + zeta)};
            double YZ[4], XZ[4], XY[4];  // index t = o1 + 2 o2
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                YZ[t] = Yf[t & 1] * Zf[t >> 1];
                XZ[t] = Xf[t & 1] * Zf[t >> 1];
                XY[t] = 0.125 * Xf[t & 1] * Yf[t >> 1];
            }
            double J[9];  // J[r*3+k] = d x_r / d xi_k   (Jac = coords*dN, src/fem.jl:192)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                J[r * 3 + 0] = YZ[0] * Ex[0][r] + YZ[1] * Ex[1][r] + YZ[2] * Ex[2][r] + YZ[3] * Ex[3][r];
                J[r * 3 + 1] = XZ[0] * Ey[0][r] + XZ[1] * Ey[1][r] + XZ[2] * Ey[2][r] + XZ[3] * Ey[3][r];
                J[r * 3 + 2] = XY[0] * Ez[0][r] + XY[1] * Ez[1][r] + XY[2] * Ez[2][r] + XY[3] * Ez[3][r];
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
            const double sc = copysign(rsqrt(fabs(det)), det) * s_sw[gp];  // sign(det) sqrt(wp/|det|): long latency ...
            // ... overlapped with the unscaled gradients  t_u[c] = sum_k dN_u[k] adj[k][c]
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int ox = u & 1, oy = (u >> 1) & 1, oz = u >> 2;
                const double d0 = ox ? YZ[oy + 2 * oz] : -YZ[oy + 2 * oz];
                const double d1 = oy ? XZ[ox + 2 * oz] : -XZ[ox + 2 * oz];
                const double d2 = oz ? XY[ox + 2 * oy] : -XY[ox + 2 * oy];
                const int b = oz * 4 + (oy ? (ox ? 2 : 3) : (ox ? 1 : 0));  // reference local numbering (vector3D.jl:94-101)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    dst[((gp * 8 + b) * 3 + c) * NELP + e] = (d0 * adj[c] + d1 * adj[3 + c] + d2 * adj[6 + c]) * sc;
            }
        }
    }
}


template <class T, int MINB, int OUT>
__global__ void __launch_bounds__(T::NTH, MINB) k_values_tile(const __grid_constant__ TileArgs A) {
    constexpr int TX = T::TX, TY = T::TY, NTH = T::NTH, NEL = T::NELP /* ring stride */, EX = T::EX, LAYER = T::LAYER;
    extern __shared__ double smem[];
    double *S = smem;                             // [2][gp][b][c][e]
    double *stage_own = smem + 2 * LAYER;         // [node][32 slots][9] when not aliased
    double *s_dN = stage_own + (T::ALIAS ? 0 : TX * TY * STAGE_NODE);  // [gp][3]: Gauss-point coordinates (xi, eta, zeta)
    double *s_w = s_dN + 8 * 8 * 3;   // sqrt of the Gauss weights
    double *s_xyz = s_w + 8;          // [4][PLANE] node-plane coordinate ring
    int32_t *s_col = reinterpret_cast<int32_t *>(s_xyz + 4 * T::PLANE);  // [warp][COLBUF]
    const Lattice &L = A.L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < 8 * 3; t += NTH) s_dN[t] = (&A.gp[0][0])[t];
    if (tid < 8) s_w[tid] = sqrt(A.w[tid]);

    int bid = blockIdx.x;
    const int tix = bid % A.tiles_x;
    bid /= A.tiles_x;
    const int tiy = bid % A.tiles_y;
    const int chunk_id = bid / A.tiles_y;
    const int X0 = tix * TX, Y0 = tiy * TY;
    const int zs = L.k0 + A.zb[chunk_id], ze = L.k0 + A.zb[chunk_id + 1];

    // phase-2 identity of this thread
    const int s = lane & 7, nt = warp * 4 + (lane >> 3);
    const int sx = s & 1, sy = (s >> 1) & 1, sz = s >> 2;
    const int tx = nt % TX, ty = nt / TX;
    const int ix = X0 + tx, iy = Y0 + ty;
    const bool node_ok = ix < L.n1 && iy < L.n1;
    const int ex = ix - sx, ey = iy - sy;
    const bool el_xy_ok = node_ok && ex >= 0 && ey >= 0 && ex < L.ne && ey < L.ne;
    const int e = (ty - sy + 1) * EX + (tx - sx + 1);
    const int a = sz * 4 + ((sy << 1) | (sx ^ sy));  // local node number of this node inside element -s
    int soff[8];  // staging offset of this thread's target block in combine round beta
#pragma unroll
    for (int beta = 0; beta < 8; ++beta)
        soff[beta] = 9 * c_slot[((beta >> 2) - sz + 1) * 9 + (((beta >> 1) & 1) - sy + 1) * 3 + ((beta & 1) - sx + 1)];
    const int out_off = lane < 27 ? 9 * c_slot[lane] : 0;
    // closed-form CSR row starts (no dependent global load in the output phase): see k_struct_rowptr
    const int64_t S1 = 3 * (int64_t)L.n1 - 2;
    auto pre1 = [](int i) -> int64_t { return i == 0 ? 0 : 3 * (int64_t)i - 1; };
    auto cnt1 = [&](int i) -> int { return 1 + (i > 0) + (i < L.n1 - 1); };
    const int64_t pairs_base = pre1(L.k0) * S1 * S1;
    // streamed coordinates (smfem_assemble_system): wait until plane p AND p + 1 have landed - the 128-byte line that
    // straddles the end of plane p must not enter L1 with bytes of p + 1 that the copy engine has not written yet
    int ready_upto = A.ready ? 0 : 0x7fffffff;
    auto wait_plane = [&](int p) {
        const int need = min(p + 2, min(L.k1 + 1, L.n1));  // the slab ends with plane min(k1, n1-1)
        unsigned spins = 0;
        while (ready_upto < need) {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(ready_upto) : "l"(A.ready) : "memory");
            if (++spins > (1u << 25)) __trap();  // tens of seconds: the copies were queued before this kernel, so never in practice
        }
    };
    wait_plane(zs + 1);
    stage_plane<T>(A, s_xyz, zs - 1, X0, Y0);
    stage_plane<T>(A, s_xyz, zs, X0, Y0);
    stage_plane<T>(A, s_xyz, zs + 1, X0, Y0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    for (int k = zs; k < ze; ++k) {
        if (k + 2 < L.n1 && k + 2 <= L.k1) wait_plane(k + 2);
        stage_plane<T>(A, s_xyz, k + 2, X0, Y0);  // lands during this plane's phase 2; ring slot (k+2)&3 is free
        if (!(A.skip & 1)) {
            if (k == zs && k - 1 >= 0) phase1<T>(A, s_dN, s_w, s_xyz, S, k - 1, X0, Y0);
            if (k < L.ne) phase1<T>(A, s_dN, s_w, s_xyz, S, k, X0, Y0);
        }
        __syncthreads();

        // ---- phase 2: G_ab for b = 0..7 -------------------------------------------------------
        const int layer = k - sz;
        const bool el_ok = el_xy_ok && layer >= 0 && layer < L.ne;
        double G[8][9];
#pragma unroll
        for (int b = 0; b < 8; ++b)
#pragma unroll
            for (int m = 0; m < 9; ++m) G[b][m] = 0.0;
        if (el_ok && !(A.skip & 2)) {
            const double *Sb = S + (layer & 1) * LAYER + e;
#pragma unroll 2
            for (int gp = 0; gp < 8; ++gp) {
                const double *Sg = Sb + gp * (8 * 3 * NEL);
                const double ga0 = Sg[(a * 3 + 0) * NEL], ga1 = Sg[(a * 3 + 1) * NEL], ga2 = Sg[(a * 3 + 2) * NEL];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const double gb0 = Sg[(b * 3 + 0) * NEL], gb1 = Sg[(b * 3 + 1) * NEL], gb2 = Sg[(b * 3 + 2) * NEL];
                    G[b][0] += ga0 * gb0;
                    G[b][1] += ga0 * gb1;
                    G[b][2] += ga0 * gb2;
                    G[b][3] += ga1 * gb0;
                    G[b][4] += ga1 * gb1;
                    G[b][5] += ga1 * gb2;
                    G[b][6] += ga2 * gb0;
                    G[b][7] += ga2 * gb1;
                    G[b][8] += ga2 * gb2;
                }
            }
        }
        // ---- combine: round beta (natural order) -> neighbour offset d = beta - s, distinct for the 8 slot threads.
        // Block d is touched for the FIRST time in round beta_min(d) = max(d,0) per axis, i.e. by the lanes with
        // (s & beta) == 0: those plain-store (no zeroing pass, no load), everybody else read-modify-writes.
        // Threads whose element does not exist still take part with G = 0 so that every block gets initialised.
        // the staging area lives in the ring slot of layer k-1, which every thread has finished reading now
        double *stage = T::ALIAS ? S + ((k - 1) & 1) * LAYER : stage_own;
        double *my_stage = stage + nt * STAGE_NODE;
        if (T::ALIAS) __syncthreads();
        else __syncwarp();
#pragma unroll
        for (int beta = 0; beta < 8; ++beta) {
            const int obx = beta & 1, oby = (beta >> 1) & 1, obz = beta >> 2;
            const int b = obz * 4 + (oby ? (obx ? 2 : 3) : (obx ? 1 : 0));  // reference local numbering
            if (!(A.skip & 4)) {
                double *dst = my_stage + soff[beta];
                if (beta == 0) {
#pragma unroll
                    for (int m = 0; m < 9; ++m) dst[m] = G[b][m];
                } else {
                    // branch-free: a divergent store / read-modify-write pair would issue the 9 stores twice per round
                    // (first-touch lanes read a stale value that the select discards)
                    const bool first = (s & beta) == 0;
                    double old[9];
#pragma unroll
                    for (int m = 0; m < 9; ++m) old[m] = dst[m];
#pragma unroll
                    for (int m = 0; m < 9; ++m) dst[m] = first ? G[b][m] : old[m] + G[b][m];
                }
            }
            __syncwarp();
        }
        // ---- output: lane q < 27 owns neighbour q of each of the warp's 4 nodes -----------------------------
        // The warp's 4 nodes are consecutive in x, so their 12 CSR rows are ONE contiguous run of K.val / K.colind (<= 972 entries).
        // OUT 0: every lane stores its 3x3 block (values and column indices) straight to global memory.
        // OUT 2: values and column indices are first permuted into CSR order IN PLACE in the warp's staging region (node
        //        jn's final range [243 jn, 243 jn + 243) only overlaps the staging of nodes <= jn, which are consumed by
        //        then) and leave by TMA bulk stores (cp.async.bulk.global.shared::cta).
        // OUT 3 (default): values as in 0, column indices as CSR-ordered runs in a buffer of their own + TMA bulk store.
        // Measured at 100^3 (values / fused, ms): 0: 1.39 / 1.72, 2: 1.45 / 1.62, 3: 1.39 / 1.56.  The kernel is bound by
        // L1TEX wavefronts (shared memory AND global stores): the 4-byte index stores with a 12-byte lane stride are the
        // expensive ones; for the values the extra shared-memory round trip of 2 costs what the scattered stores cost.
        if (!(A.skip & 8)) {
            const int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
            const int jy = Y0 + (warp * 4) / TX, jx0 = X0 + (warp * 4) % TX;
            const int cy = 1 + (jy > 0) + (jy < L.n1 - 1), cz = 1 + (k > 0) + (k < L.n1 - 1);
            const bool row_ok = jy < L.n1 && jx0 < L.n1;
            const int64_t run_base =
                9 * (pre1(k) * S1 * S1 + (int64_t)cz * (pre1(jy) * S1 + (int64_t)cy * pre1(jx0)) - pairs_base);
            double *R = stage + warp * 4 * STAGE_NODE;
            if (OUT == 0 || OUT == 3) {
                int64_t base_n = run_base;
                if (lane < 27 && row_ok) {
                    for (int jn = 0; jn < 4; ++jn) {
                        const int n2 = warp * 4 + jn, jx = jx0 + jn;
                        if (jx >= L.n1) break;
                        const int cx = 1 + (jx > 0) + (jx < L.n1 - 1);
                        const int TR = 3 * cx * cy * cz;
                        const int nx = jx + dx, ny = jy + dy, nz = k + dz;
                        if (nx >= 0 && ny >= 0 && nz >= 0 && nx < L.n1 && ny < L.n1 && nz < L.n1) {
                            const int rank = ((dz + (k > 0)) * cy + (dy + (jy > 0))) * cx + (dx + (jx > 0));
                            const int64_t row = (((int64_t)(k - L.k0) * L.n1 + jy) * L.n1 + jx) * 3;
                            const int64_t base = ((A.skip & 32) ? (base_n & 1023) + 1024 * warp : base_n) + 3 * rank;  // 32: ablation
                            const double *g = stage + n2 * STAGE_NODE + out_off;
                            const double tr = g[0] + g[4] + g[8];
                            {
#pragma unroll
                                for (int c = 0; c < 3; ++c)
#pragma unroll
                                    for (int j = 0; j < 3; ++j) {
                                        const double gij = g[c * 3 + j], gji = g[j * 3 + c];
                                        const double v =
                                            (c == j) ? A.mat.d11 * gij + A.mat.mu * (tr - gij) : A.mat.lam * gij + A.mat.mu * gji;
                                        A.val[base + (int64_t)c * TR + j] = v;
                                        if (lane == 13 && c == j) A.diag[row + c] = v;
                                        if (OUT == 0 && A.colind)
                                            A.colind[base + (int64_t)c * TR + j] = (int32_t)(L.lnode(nx, ny, nz) * 3 + j);
                                    }
                            }
                        }
                        base_n += 3 * TR;
                    }
                }
                if (OUT == 3 && A.colind) {
                    // column indices: built in CSR order in the warp's staging region (dead now) and written by one TMA bulk store
                    base_n = __shfl_sync(0xffffffffu, base_n, 0);
                    const int run_len = (int)(base_n - run_base);
                    if (T::COLBUF)
                        emit_colind_run<true>(A, s_col + warp * T::COLBUF, run_base, run_len, row_ok, jx0, jy, k, cy, cz, dx, dy, dz, lane);
                    else
                        emit_colind_run<false>(A, reinterpret_cast<int32_t *>(R), run_base, run_len, row_ok, jx0, jy, k, cy, cz, dx, dy, dz, lane);
                }
            } else {
                const int par = (int)(run_base & 1);
                int run_len = 0;
                for (int jn = 0; jn < 4; ++jn) {
                    const int n2 = warp * 4 + jn, jx = jx0 + jn;
                    const bool n_ok = row_ok && jx < L.n1;
                    const int cx = 1 + (jx > 0) + (jx < L.n1 - 1);
                    const int TR = 3 * cx * cy * cz;
                    const int nx = jx + dx, ny = jy + dy, nz = k + dz;
                    const bool ok = n_ok && lane < 27 && nx >= 0 && ny >= 0 && nz >= 0 && nx < L.n1 && ny < L.n1 && nz < L.n1;
                    const int rank = ((dz + (k > 0)) * cy + (dy + (jy > 0))) * cx + (dx + (jx > 0));
                    double v[9];
                    if (ok) {
                        const double *g = stage + n2 * STAGE_NODE + out_off;
                        double gg[9];
#pragma unroll
                        for (int m = 0; m < 9; ++m) gg[m] = g[m];
                        const double tr = gg[0] + gg[4] + gg[8];
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                const double gij = gg[c * 3 + j], gji = gg[j * 3 + c];
                                v[c * 3 + j] = (c == j) ? A.mat.d11 * gij + A.mat.mu * (tr - gij) : A.mat.lam * gij + A.mat.mu * gji;
                            }
                        if (lane == 13) {
                            const int64_t row = (((int64_t)(k - L.k0) * L.n1 + jy) * L.n1 + jx) * 3;
                            A.diag[row] = v[0];
                            A.diag[row + 1] = v[4];
                            A.diag[row + 2] = v[8];
                        }
                    }
                    __syncwarp();  // node jn's staging has been read by every lane: its area may now be overwritten
                    if (ok) {
                        double *dst = R + par + run_len + 3 * rank;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int j = 0; j < 3; ++j) dst[c * TR + j] = v[c * 3 + j];
                    }
                    if (n_ok) run_len += 3 * TR;
                }
                {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && run_len > 0) {
                        const int64_t a0 = run_base + par, a1 = (run_base + run_len) & ~(int64_t)1;
                        if (par) A.val[run_base] = R[par];
                        if (a1 < run_base + run_len) A.val[a1] = R[par + (a1 - run_base)];
                        if (a1 > a0) {
                            const unsigned src = (unsigned)__cvta_generic_to_shared(R + 2 * par);
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(A.val + a0), "r"(src),
                                         "r"((unsigned)((a1 - a0) * 8))
                                         : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        }
                    }
                }
                if (A.colind) {  // fused assembly: the pattern's column indices take the same route through the region
                    emit_colind_run<false>(A, reinterpret_cast<int32_t *>(R), run_base, run_len, row_ok, jx0, jy, k, cy, cz, dx, dy, dz, lane);
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();  // staging + ring slot (k-1)&1 are reused by the next plane; coordinate plane k+2 has landed
    }
    if (OUT == 3 && T::COLBUF && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the last bulk store
}

}  // namespace

bool values_tile_enabled() {
    const char *e = std::getenv("SMFEM_VALUES");
    return !(e && std::string(e) == "atomic");
}

// Split the owned planes of a tile column into chunks (one CTA each).  The hardware hands CTAs to free slots in blockIdx
// order (chunk-major here), i.e. list scheduling; CTAs of equal length in a grid that is not a multiple of the resident
// slot count leave a tail (100^3 on 296 slots: 2704 CTAs of 26 planes = 9.13 waves -> 10).  Candidates: 1..MAX_CHUNKS
// chunks whose lengths decay geometrically (long chunks first, short ones fill the tail); the makespan of each is
// simulated with duration = planes + c0 (prologue: one extra element layer, pipeline fill) and the best one is kept
// (100^3: 55 + 30 + 16 planes, modelled makespan 242 plane-times instead of 270; ideal 231).
std::vector<int> plan_chunks(int ntiles, int nown, int slots) {
    if (const char *e = std::getenv("SMFEM_TILE_CHUNKS")) {  // experiments: explicit comma-separated lengths
        std::vector<int> len;
        int sum = 0;
        for (const char *p = e; *p;) {
            len.push_back(std::atoi(p));
            sum += len.back();
            while (*p && *p != ',') ++p;
            if (*p == ',') ++p;
        }
        if (sum == nown && (int)len.size() <= MAX_CHUNKS) return len;
    }
    static std::map<std::tuple<int, int, int>, std::vector<int>> cache;
    static std::mutex cache_mu;  // one process may assemble on several GPUs from several host threads (smfem_init_multi)
    std::lock_guard<std::mutex> cache_lock(cache_mu);
    const auto key = std::make_tuple(ntiles, nown, slots);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    double c0 = 1.0;
    if (const char *e = std::getenv("SMFEM_TILE_C0")) c0 = std::atof(e);
    const int minlen = 4;
    const double cap = std::max((double)minlen, (double)ntiles * nown / slots / 4);  // keep >= ~4 rounds of CTAs: the model is idealised
    std::vector<int> best;
    double best_t = 1e300;
    const double ratios[] = {1.0, 0.9, 0.8, 0.7, 0.6, 0.5};
    for (int nc = 1; nc <= MAX_CHUNKS && nc * minlen <= std::max(nown, minlen); ++nc)
        for (double r : ratios) {
            if (nc == 1 && r != 1.0) continue;
            // lengths proportional to r^i, at least minlen, summing to nown
            std::vector<int> len(nc, minlen);
            int rest = nown - nc * minlen;
            if (rest < 0) { len.assign(1, nown); rest = 0; }
            double wsum = 0;
            for (int i = 0; i < (int)len.size(); ++i) wsum += std::pow(r, i);
            int given = 0;
            for (int i = 0; i < (int)len.size(); ++i) {
                const int g = (int)std::floor(rest * std::pow(r, i) / wsum);
                len[i] += g;
                given += g;
            }
            for (int i = 0; given < rest; i = (i + 1) % (int)len.size(), ++given) ++len[i];
            if (len[0] > cap && (nc + 1) * minlen <= nown && nc < MAX_CHUNKS) continue;
            // list scheduling: all CTAs of a chunk have the same duration
            std::priority_queue<double, std::vector<double>, std::greater<double>> free_at;
            for (int i = 0; i < slots; ++i) free_at.push(0.0);
            double span = 0;
            for (int c = 0; c < (int)len.size(); ++c)
                for (int t = 0; t < ntiles; ++t) {
                    const double end = free_at.top() + len[c] + c0;
                    free_at.pop();
                    free_at.push(end);
                    span = std::max(span, end);
                }
            if (span < best_t - 1e-9) {
                best_t = span;
                best = len;
            }
        }
    cache[key] = best;
    return best;
}

template <class T, int MINB, int OUT>
static void launch_tile(smfem_ctx *ctx, TileArgs &A, int nown) {
    static std::atomic<unsigned long long> attr_set{0};
    if (first_use_on_device(attr_set)) {
        CUDA_CHECK(cudaFuncSetAttribute(k_values_tile<T, MINB, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES));
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
    LAUNCH(ctx, (k_values_tile<T, MINB, OUT>), grid, T::NTH, T::SMEM_BYTES, A);
    CUDA_CHECK(cudaEventRecord(ctx->asm_ev[2 * slot + 1], ctx->stream));
    ctx->asm_count++;
}

void tile_fill_args(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, Material mat, bool write_colind, const int *ready, TileArgs &A) {
    if (!K->diag) K->diag = dev_alloc<double>(K->nrows_l);
    A.L = mesh->lat;
    A.coords = mesh->coords;
    A.rowptr = K->rowptr;
    A.val = K->val;
    A.colind = write_colind ? K->colind : nullptr;
    A.ready = ready;
    A.diag = K->diag;
    A.mat = mat;
    {
        double xi[2], w[2];
        smfem_host_gauss(-1, 1, 2, xi, w);
        const int ix[8] = {0, 1, 1, 0, 0, 1, 1, 0}, iy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, iz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
        for (int g = 0; g < 8; ++g) {
            A.gp[g][0] = xi[ix[g]];
            A.gp[g][1] = xi[iy[g]];
            A.gp[g][2] = xi[iz[g]];
            A.w[g] = w[ix[g]] * w[iy[g]] * w[iz[g]];
            A.sw[g] = std::sqrt(A.w[g]);
        }
    }
    const char *sk = std::getenv("SMFEM_TILE_SKIP");
    A.skip = sk ? std::atoi(sk) : 0;
    const char *om = std::getenv("SMFEM_TILE_OUT");
    A.out_mode = om ? std::atoi(om) : 3;
}

void values_assemble_tile(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, Material mat, bool write_colind, const int *ready) {
    TileArgs A;
    tile_fill_args(ctx, mesh, K, mat, write_colind, ready, A);
    if (values_assemble_tile3(ctx, A, A.L.nown())) return;  // split-role layer-march kernel (assemble_tile3.cu), SMFEM_TILE=v3
    if (values_assemble_tile2(ctx, A, A.L.nown())) return;  // layer-march kernel (assemble_tile2.cu)
    if (values_assemble_mma(ctx, A, A.L.nown())) return;  // DMMA kernel (assemble_mma.cu), selected by SMFEM_TILE=mma*
    const char *e = std::getenv("SMFEM_TILE");
    const bool big = e && std::string(e) == "8x4";  // 256 threads, 1 CTA/SM; default 4x4: 128 threads, 2 CTAs/SM
#define SMFEM_TILE_CASE(M)                                              \
    case M:                                                             \
        if (big) launch_tile<Tile<8, 4>, 1, M>(ctx, A, A.L.nown());     \
        else launch_tile<Tile<4, 4>, 2, M>(ctx, A, A.L.nown());         \
        break;
    switch (A.out_mode) {
        SMFEM_TILE_CASE(2)
        SMFEM_TILE_CASE(3)
        default:
        SMFEM_TILE_CASE(0)
    }
#undef SMFEM_TILE_CASE
}
