// K1 + K3 for the structured hex lattice, nDof = 3, third formulation ("layer march, split roles"): the layer-march kernel of
// assemble_tile2.cu with the 72 accumulators of a (node column, element slot) pair divided between TWO threads, so that the
// kernel runs at 16 instead of 8 warps per SM (the occupancy experiment in profiles/r2_tile2_variants.txt: 4 -> 8 warps was
// worth x1.43):
//   replaces src/fem.jl:179-249 (element loop + COO scatter) and the value side of sparse(E,J,V) (:253)
// CTA = 4 x 8 node columns, 256 threads = 8 warps, 2 CTAs/SM.  Warps 0-3 ("same-plane" role) accumulate the dz = 0 blocks
// (a and b on the swept face; carried in registers from a layer's top sweep into the next layer's bottom sweep) and emit one
// level per layer; warps 4-7 ("other-plane" role) accumulate the dz = -1 / +1 blocks (a on the other face) and emit two levels
// per layer.  Each role loads the g_b of the swept face itself (the price: twice the shared-memory loads of the sweeps); phase 1,
// the shuffle combine, the staging layout and the closed-form CSR geometry are those of assemble_tile2.cu, the staging area
// holds one row of 4 nodes at a time (the two node rows of a warp are emitted one after the other) to fit two CTAs per SM.
// Same arithmetic in the same order as k_values_tile2: the results are bit-identical.
#include <cmath>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "smfem_internal.cuh"
#include "tile_args.cuh"

namespace {

// WPRv: warps per role (tile = 4 x 2 WPRv node columns, 64 WPRv threads).  FULLv: the staging area of a warp holds both of its
// node rows (8 nodes, as in assemble_tile2.cu) instead of one at a time.
//   <4, false>: 4 x 8 tile, 256 threads, 2 CTAs/SM = 16 warps/SM, <= 128 registers, half-size staging (shared memory is 6 KB short)
//   <3, true> : 4 x 6 tile, 192 threads, 2 CTAs/SM = 12 warps/SM, <= 168 registers, full staging
template <int WPRv, bool FULLv>
struct T3 {
    static constexpr int TX = 4, WPR = WPRv, TY = 2 * WPRv, NTH = 64 * WPRv, OPT = 15;
    static constexpr bool FULL = FULLv;
    static constexpr int EX = TX + 1, EY = TY + 1, NEL = EX * EY;  // footprint elements per layer (45 / 35)
    static constexpr int NELP = NEL;
    // pads in front of the rows of the face nodes q = 1, 2, 3: conflict-free own-node loads (see T2 in assemble_tile2.cu)
    static constexpr int P1 = NELP == 45 ? 2 : 0, P2 = NELP == 45 ? 4 : 0, P3 = NELP == 45 ? 4 : 14;
    static constexpr int FACE = 12 * NELP + P3;
    static constexpr int GS = ((2 * FACE + 15) / 16) * 16 + 2;
    static constexpr int LAYER = 8 * GS;
    static constexpr int PX = TX + 2, PY = TY + 2, PLANE = PX * PY * 3;
    static constexpr int SN = 84;
    static constexpr int STAGE_WARP = (FULLv ? 8 : 4) * SN;
    static constexpr size_t SMEM_BYTES = sizeof(double) * (LAYER + (NTH / 32) * STAGE_WARP + 8 * 3 + 8 + 4 * PLANE);
    __host__ __device__ static constexpr int boff(int b) {
        return (b >> 2) * FACE + (b & 3) * 3 * NELP + ((b & 3) == 0 ? 0 : ((b & 3) == 1 ? P1 : ((b & 3) == 2 ? P2 : P3)));
    }
    static constexpr int R1 = (boff(1) - 1) & 15, R2 = (boff(2) - EX - 1) & 15, R3 = (boff(3) - EX) & 15;
    static_assert(R1 % 4 == 0 && R2 % 4 == 0 && R3 % 4 == 0 && R1 && R2 && R3 && R1 != R2 && R1 != R3 && R2 != R3,
                  "face-row pads do not give a conflict-free own-node load");
    static_assert(NELP == 45 || NELP == 35, "pads are tabulated for these two tile heights");
};

template <class T>
__device__ __forceinline__ void phase1_layer(const TileArgs &A, const double *s_gp, const double *s_sw, const double *s_xyz, double *S,
                                             int layer, int X0, int Y0) {
    constexpr int NEL = T::NEL, NELP = T::NELP, EX = T::EX, NTH = T::NTH;
    const Lattice &L = A.L;
    const double *P0 = s_xyz + (layer & 3) * T::PLANE, *P1 = s_xyz + ((layer + 1) & 3) * T::PLANE;
    // task = (element, Gauss point): 360 tasks on 128 threads = 3 rounds at 94 % lane use (pairs of Gauss points sharing the
    // edge differences need 4 x 64 task slots: 70 %).  OPT 1: the 8 Gauss points of an element sit in consecutive lanes (their
    // coordinate loads are broadcasts, their stores go to 8 distinct bank pairs); else consecutive lanes = consecutive e
    for (int q = threadIdx.x; q < 8 * NEL; q += NTH) {
        const int gp = (T::OPT & 1) ? (q & 7) : q / NEL, e = (T::OPT & 1) ? (q >> 3) : q - gp * NEL;
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
        double Ex[4][3], Ey[4][3], Ez[4][3];  // edge differences along xi / eta / zeta
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int o1 = t & 1, o2 = t >> 1;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                Ex[t][r] = Xn[1 + 2 * o1 + 4 * o2][r] - Xn[2 * o1 + 4 * o2][r];
                Ey[t][r] = Xn[o1 + 2 + 4 * o2][r] - Xn[o1 + 4 * o2][r];
                Ez[t][r] = Xn[o1 + 2 * o2 + 4][r] - Xn[o1 + 2 * o2][r];
            }
        }
        {
            const double xi = s_gp[3 * gp], eta = s_gp[3 * gp + 1], zeta = s_gp[3 * gp + 2];
            const double Xf[2] = {1.0 - xi, 1.0 + xi}, Yf[2] = {1.0 - eta, 1.0 + eta}, Zf[2] = {0.125 * (1.0 - zeta), 0.125 * (1.0 + zeta)};
            double YZ[4], XZ[4], XY[4];
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
            const double sc = copysign(rsqrt(fabs(det)), det) * s_sw[gp];  // sign(det) sqrt(wp / |det|)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int ox = u & 1, oy = (u >> 1) & 1, oz = u >> 2;
                const double d0 = ox ? YZ[oy + 2 * oz] : -YZ[oy + 2 * oz];
                const double d1 = oy ? XZ[ox + 2 * oz] : -XZ[ox + 2 * oz];
                const double d2 = oz ? XY[ox + 2 * oy] : -XY[ox + 2 * oy];
                const int b = oz * 4 + (oy ? (ox ? 2 : 3) : (ox ? 1 : 0));  // reference local numbering (vector3D.jl:94-101)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    S[gp * T::GS + T::boff(b) + c * NELP + e] = (d0 * adj[c] + d1 * adj[3 + c] + d2 * adj[6 + c]) * sc;
            }
        }
    }
}

__device__ __forceinline__ double shfl_xor_f64(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

// One sweep over the 8 Gauss points: X[q] += g_a g_b(q)' for the 4 nodes b of one face of this thread's element (Sf points at the
// face's g_b rows, Sa at the row of the node a this thread's role pairs them with: the column's node on the swept face for the
// same-plane role, on the other face for the other-plane role)
template <class T3>
__device__ __forceinline__ void sweep1(const double *Sf, const double *Sa, double (&X)[4][9]) {
    constexpr int NELP = T3::NELP;
#pragma unroll 2
    for (int gp = 0; gp < 8; ++gp) {
        const double *Sg = Sf + gp * T3::GS;
        double gb[4][3], ga[3];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int c = 0; c < 3; ++c) gb[q][c] = Sg[T3::boff(q) + c * NELP];
#pragma unroll
        for (int c = 0; c < 3; ++c) ga[c] = Sa[gp * T3::GS + c * NELP];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) X[q][i * 3 + j] += ga[i] * gb[q][j];
    }
}

// Closed-form CSR geometry of the lattice (see k_struct_rowptr) and the per-thread constants of the output phase; none of
// them depends on the plane, so they are set up once per CTA.
struct NodeGeo {
    int n1, k0;
    int64_t S1, pairs_base;
    int soff[3];   // staging offset (3 * rank among the node's in-plane neighbours) of the thread's output blocks, -1: none
    int len_own;   // 3 cx cy of the thread's node: entries per CSR row and dz level
    int rowc[2];   // pre(jy) S1 + cy pre(jx0) for the warp's two node rows
    int crel27;    // interior nodes: column of section entry `lane` relative to 3 * node  (lane < 27)
    bool fast;     // all 8 nodes of the warp exist and are interior in x and y
    __device__ __forceinline__ int cnt(int i) const { return 1 + (i > 0) + (i < n1 - 1); }
    __device__ __forceinline__ int64_t pre(int i) const { return i == 0 ? 0 : 3 * (int64_t)i - 1; }
};

// Combine the 4 in-plane blocks X[q] of the 4 slot threads of every node onto the 9 neighbour blocks of ONE dz level (27 shuffles),
// apply the material, then - one node row of the warp at a time - park the level in the staging area in CSR order and copy it
// out: for each node and each of its 3 rows a section of len = 3 cx cy consecutive entries at  rowstart + lz * len.
// After the shuffles thread (sx, sy) holds   O0: (sy - sx, 0)  [not slot 3]   O1: (-sx, 1 - 2 sy)  [not slot 2]
//                                            O2: (1 - sx, 1 - 2 sy)  [not slot 1]
template <class T3>
__device__ __forceinline__ void emit_level3(const TileArgs &A, const NodeGeo &G, double (&X)[4][9], double *stage_w, int lane, int ix, int iy,
                                            int k, int dz, int jx0, int jy0) {
    const int slot = lane & 3, sx = slot & 1, sy = slot >> 1, nwl = lane >> 2;
    const Lattice &L = A.L;
    const int len = G.len_own;
    double V[3][9];  // the thread's three output blocks with the material applied (src/fem.jl:230-249); FULL: stored at once
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        double g[9];
#pragma unroll
        for (int m = 0; m < 9; ++m) {
            if (o == 0) {
                const double R1 = shfl_xor_f64(sy ? X[3][m] : X[1][m], 2);
                const double Y0 = (sy ? X[2][m] : X[0][m]) + R1;
                const double RA = shfl_xor_f64(Y0, 3);
                g[m] = (slot == 0) ? Y0 + RA : Y0;
            } else {
                const double Ya = sy ? X[0][m] : X[3][m];
                const double Yb = sy ? X[1][m] : X[2][m];
                const double RB = shfl_xor_f64(sx ? Yb : Ya, 1);
                g[m] = o == 1 ? ((slot == 0) ? Ya + RB : Ya) : ((slot == 3) ? Yb + RB : Yb);
            }
        }
        const double tr = g[0] + g[4] + g[8];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double gij = g[c * 3 + j], gji = g[j * 3 + c];
                const double v = (c == j) ? A.mat.d11 * gij + A.mat.mu * (tr - gij) : A.mat.lam * gij + A.mat.mu * gji;
                if (T3::FULL) {
                    if (G.soff[o] >= 0) stage_w[nwl * T3::SN + G.soff[o] + c * len + j] = v;
                } else {
                    V[o][c * 3 + j] = v;
                }
                if (o == 0 && c == j && dz == 0 && slot == 0 && G.soff[0] >= 0)
                    A.diag[(((int64_t)(k - L.k0) * L.n1 + iy) * L.n1 + ix) * 3 + c] = v;
            }
    }
    const int cz = G.cnt(k), lz = dz + (k > 0), nz = k + dz;
    const int64_t planeoff = 9 * (G.pre(k) * G.S1 * G.S1 - G.pairs_base);
    const int ll = lane < 27 ? lane : lane - 16;  // lanes >= 27 re-read lanes 11..15's words (no bank conflict with lanes 16..26)
#pragma unroll 1
    if (T3::FULL) __syncwarp();
    for (int yrow = 0; yrow < 2; ++yrow) {
        const double *stage_r = stage_w + (T3::FULL ? yrow * 4 * T3::SN : 0);  // this node row's 4 staged nodes
        if (!T3::FULL && (nwl >> 2) == yrow) {
            double *my = stage_w + (nwl & 3) * T3::SN;
#pragma unroll
            for (int o = 0; o < 3; ++o)
                if (G.soff[o] >= 0) {
                    double *dst = my + G.soff[o];
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int j = 0; j < 3; ++j) dst[c * len + j] = V[o][c * 3 + j];
                }
        }
        if (!T3::FULL) __syncwarp();
        const int jy = jy0 + yrow;
        if (G.fast) {
            // interior warp: 12 sections of 27 entries; all loads first, then the stores; lane = position inside the section
            double v[12];
#pragma unroll
            for (int r = 0; r < 12; ++r) v[r] = stage_r[(r / 3) * T3::SN + (r % 3) * 27 + ll];
            if (lane < 27) {
                const int32_t col0 = (int32_t)(L.lnode(jx0, jy, nz) * 3) + G.crel27;
                const int64_t g0 = planeoff + 9 * (int64_t)cz * G.rowc[yrow] + lz * 27 + lane;
                auto copy_out = [&](auto trc) {
                    constexpr int TRC = decltype(trc)::value;
                    const int TR = TRC ? TRC : 27 * cz;
                    double *vp = A.val + g0;
#pragma unroll
                    for (int r = 0; r < 12; ++r) vp[r * TR] = v[r];
                    if (A.colind) {
                        int32_t *cp = A.colind + g0;
#pragma unroll
                        for (int r = 0; r < 12; ++r) cp[r * TR] = col0 + 3 * (r / 3);
                    }
                };
                if (cz == 3) copy_out(std::integral_constant<int, 81>());
                else copy_out(std::integral_constant<int, 0>());
            }
        } else if (jy < L.n1) {
            const int cyr = G.cnt(jy);
            int64_t base = planeoff + 9 * (int64_t)cz * G.rowc[yrow];
#pragma unroll 1
            for (int node = 0; node < 4; ++node) {
                const int jx = jx0 + node;
                if (jx >= L.n1) break;
                const int cxn = G.cnt(jx), lenn = 3 * cxn * cyr, TR = lenn * cz;
                if (lane < lenn) {
                    const int blk = (lane * 11) >> 5, j = lane - 3 * blk;  // lane / 3 for lane < 32
                    const int dyr = cxn == 3 ? (blk * 11) >> 5 : blk >> 1;   // cxn is 2 or 3 (n1 >= 2)
                    const int dxr = blk - dyr * cxn;
                    const int32_t col = (int32_t)(L.lnode(jx + dxr - (jx > 0), jy + dyr - (jy > 0), nz) * 3) + j;
                    const double *src = stage_r + node * T3::SN + lane;
                    const int64_t g0 = base + lz * lenn + lane;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        A.val[g0 + (int64_t)c * TR] = src[c * lenn];
                        if (A.colind) A.colind[g0 + (int64_t)c * TR] = col;
                    }
                }
                base += 3 * TR;
            }
        }
        if (!T3::FULL) __syncwarp();  // the staging area is rewritten by the next node row
    }
    if (T3::FULL) __syncwarp();  // ... by the next level
}

template <class T3>
__global__ void __launch_bounds__(T3::NTH, 2) k_values_tile3(const __grid_constant__ TileArgs A) {
    constexpr int NTH = T3::NTH, EX = T3::EX, LAYER = T3::LAYER;
    extern __shared__ double smem[];
    double *S = smem;                               // [gp][b][c][e]: the resident element layer
    double *s_stage = S + LAYER;                    // [warp][4 nodes][SN]
    double *s_gp = s_stage + (NTH / 32) * T3::STAGE_WARP;  // [gp][3]
    double *s_w = s_gp + 8 * 3;                     // sqrt of the Gauss weights
    double *s_xyz = s_w + 8;                        // [4][PLANE] node-plane coordinate ring
    const Lattice &L = A.L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int role = warp / T3::WPR, wq = warp - role * T3::WPR;  // role 0: same-plane blocks (dz = 0), role 1: other-plane blocks (dz = -1, +1)
    for (int t = tid; t < 8 * 3; t += NTH) s_gp[t] = (&A.gp[0][0])[t];
    if (tid < 8) s_w[tid] = sqrt(A.w[tid]);

    int bid = blockIdx.x;
    const int tix = bid % A.tiles_x;
    bid /= A.tiles_x;
    const int tiy = bid % A.tiles_y;
    const int chunk_id = bid / A.tiles_y;
    const int X0 = tix * T3::TX, Y0 = tiy * T3::TY;
    const int zs = L.k0 + A.zb[chunk_id], ze = L.k0 + A.zb[chunk_id + 1];  // owned node planes [zs, ze) of this CTA

    // identity of this thread: node column (ix, iy), in-plane element slot (sx, sy)
    const int slot = lane & 3, sx = slot & 1, sy = slot >> 1, nwl = lane >> 2;
    const int tx = nwl & 3, ty = 2 * wq + (nwl >> 2);
    const int ix = X0 + tx, iy = Y0 + ty;
    const bool node_ok = ix < L.n1 && iy < L.n1;
    const int ex = ix - sx, ey = iy - sy;
    const bool el_ok = node_ok && ex >= 0 && ey >= 0 && ex < L.ne && ey < L.ne;
    const int e = (ty - sy + 1) * EX + (tx - sx + 1);
    const int aq = sy ? (sx ? 2 : 3) : (sx ? 1 : 0);  // in-plane reference number of this node inside its element
    const double *Se = S + e;
    const int boff_aq = aq == 0 ? T3::boff(0) : (aq == 1 ? T3::boff(1) : (aq == 2 ? T3::boff(2) : T3::boff(3)));
    double *stage_w = s_stage + warp * T3::STAGE_WARP;
    const int jx0 = X0, jy0 = Y0 + 2 * wq;
    NodeGeo G;
    G.n1 = L.n1;
    G.k0 = L.k0;
    G.S1 = 3 * (int64_t)L.n1 - 2;
    G.pairs_base = G.pre(L.k0) * G.S1 * G.S1;
    {
        const int cx = G.cnt(ix), cy = G.cnt(iy);
        G.len_own = node_ok ? 3 * cx * cy : 0;
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const int dx = o == 0 ? sy - sx : (o == 1 ? -sx : 1 - sx);
            const int dy = o == 0 ? 0 : 1 - 2 * sy;
            const bool held = o == 0 ? slot != 3 : (o == 1 ? slot != 2 : slot != 1);
            const int nx = ix + dx, ny = iy + dy;
            const bool ok = held && node_ok && nx >= 0 && ny >= 0 && nx < L.n1 && ny < L.n1;
            G.soff[o] = ok ? 3 * ((dy + (iy > 0)) * cx + (dx + (ix > 0))) : -1;
        }
#pragma unroll
        for (int yrow = 0; yrow < 2; ++yrow) {
            const int jy = min(jy0 + yrow, L.n1 - 1);
            G.rowc[yrow] = (int)(G.pre(jy) * G.S1 + (int64_t)G.cnt(jy) * G.pre(jx0 < L.n1 ? jx0 : 0));
        }
        const int blk = (lane * 11) >> 5, j = lane - 3 * blk, dyr = (blk * 11) >> 5, dxr = blk - 3 * dyr;
        G.crel27 = 3 * ((dyr - 1) * L.n1 + (dxr - 1)) + j;
        G.fast = jx0 >= 1 && jx0 + 3 <= L.n1 - 2 && jy0 >= 1 && jy0 + 1 <= L.n1 - 2;
    }

    // streamed coordinates (smfem_assemble_system): see assemble_tile.cu
    int ready_upto = A.ready ? 0 : 0x7fffffff;
    auto wait_plane = [&](int p) {
        const int need = min(p + 2, min(L.k1 + 1, L.n1));
        unsigned spins = 0;
        while (ready_upto < need) {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(ready_upto) : "l"(A.ready) : "memory");
            if (++spins > (1u << 25)) __trap();
        }
    };
    // coordinate staging: which word of a node plane this thread copies does not depend on the plane
    int stg_src;
    {
        const int t = tid;
        const int c = t % 3, n = t / 3, px = n % T3::PX, py = n / T3::PX;
        const int gx = X0 - 1 + px, gy = Y0 - 1 + py;
        stg_src = (t < T3::PLANE && gx >= 0 && gy >= 0 && gx < L.n1 && gy < L.n1) ? 3 * (gy * L.n1 + gx) + c : -1;
    }
    static_assert(T3::PLANE <= NTH, "one staged word per thread");
    auto stage = [&](int k) {
        if (k < 0 || k >= L.n1 || k > L.k1 || stg_src < 0) return;  // the slab holds planes k0-1 .. k1
        const double *src = A.coords + 3 * (int64_t)(k - L.k0 + 1) * L.n1 * L.n1;
        unsigned d = (unsigned)__cvta_generic_to_shared(s_xyz + (k & 3) * T3::PLANE + tid);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + stg_src) : "memory");
    };
    const int L0 = max(zs - 1, 0), L1 = min(ze - 1, L.ne - 1);  // element layers this CTA sweeps (inclusive)
    wait_plane(min(L0 + 1, L.n1 - 1));
    stage(L0);
    stage(L0 + 1);

    double X[4][9];  // role 0: the dz = 0 blocks (carried across layers);  role 1: the dz = -1 / +1 blocks of the current half-step
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int m = 0; m < 9; ++m) X[q][m] = 0.0;
    auto zero = [&]() {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int m = 0; m < 9; ++m) X[q][m] = 0.0;
    };

    for (int lay = L0; lay <= L1; ++lay) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();  // everybody is done with the previous layer in S; coordinate planes lay, lay + 1 have landed
        if (lay + 2 < L.n1 && lay + 2 <= L.k1) wait_plane(lay + 2);
        if (lay + 1 <= L1) stage(lay + 2);  // lands during this layer's sweeps
        phase1_layer<T3>(A, s_gp, s_w, s_xyz, S, lay, X0, Y0);
        __syncthreads();
        if (role == 0) {
            // bottom face: a and b on plane lay -> its dz = 0 blocks are complete;  top face: a and b on plane lay + 1, carried
            if (el_ok) sweep1<T3>(Se, Se + boff_aq, X);
            if (lay >= zs && lay < ze) emit_level3<T3>(A, G, X, stage_w, lane, ix, iy, lay, 0, jx0, jy0);
            zero();
            if (el_ok) sweep1<T3>(Se + T3::FACE, Se + T3::FACE + boff_aq, X);
        } else {
            // bottom face with a on the top face: dz = -1 blocks of plane lay + 1;  top face with a on the bottom face: dz = +1 of plane lay
            if (el_ok) sweep1<T3>(Se, Se + T3::FACE + boff_aq, X);
            if (lay + 1 >= zs && lay + 1 < ze) emit_level3<T3>(A, G, X, stage_w, lane, ix, iy, lay + 1, -1, jx0, jy0);
            zero();
            if (el_ok) sweep1<T3>(Se + T3::FACE, Se + boff_aq, X);
            if (lay >= zs && lay < ze) emit_level3<T3>(A, G, X, stage_w, lane, ix, iy, lay, +1, jx0, jy0);
            zero();
        }
    }
    // the top plane of the lattice has no element layer above it: its dz = 0 level is what the last top sweep left
    if (role == 0 && ze == L.n1 && L1 + 1 >= zs) emit_level3<T3>(A, G, X, stage_w, lane, ix, iy, L1 + 1, 0, jx0, jy0);
}

}  // namespace

template <class T>
static void launch_tile3(smfem_ctx *ctx, TileArgs &A, int nown) {
    static std::atomic<unsigned long long> attr_set{0};
    if (first_use_on_device(attr_set))
        CUDA_CHECK(cudaFuncSetAttribute(k_values_tile3<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES));
    A.tiles_x = (A.L.n1 + T::TX - 1) / T::TX;
    A.tiles_y = (A.L.n1 + T::TY - 1) / T::TY;
    const int ntiles = A.tiles_x * A.tiles_y;
    const std::vector<int> len = plan_chunks(ntiles, nown, ctx->sms * 2);
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
    LAUNCH(ctx, (k_values_tile3<T>), grid, T::NTH, T::SMEM_BYTES, A);
    CUDA_CHECK(cudaEventRecord(ctx->asm_ev[2 * slot + 1], ctx->stream));
    ctx->asm_count++;
}

// SMFEM_TILE = v3 (256 threads, 16 warps/SM, half staging) / v3b (192 threads, 12 warps/SM, full staging): the split-role
// layer-march kernels (returns false when not selected)
bool values_assemble_tile3(smfem_ctx *ctx, TileArgs &A, int nown) {
    const char *sel = std::getenv("SMFEM_TILE");
    const std::string m = sel ? sel : "";
    if (m == "v3") launch_tile3<T3<4, false>>(ctx, A, nown);
    else if (m == "v3b") launch_tile3<T3<3, true>>(ctx, A, nown);
    else return false;
    return true;
}
