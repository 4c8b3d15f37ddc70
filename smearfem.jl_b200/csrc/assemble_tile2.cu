// K1 + K3 for the structured hex lattice, nDof = 3, second formulation ("layer march"): same gather idea as
// assemble_tile.cu (atomic-free, every entry of K written exactly once, fixed fold order -> bit-reproducible), rearranged so
// that the kernel is no longer bound by shared-memory wavefronts:
//   replaces src/fem.jl:179-249 (element loop + COO scatter) and the value side of sparse(E,J,V) (:253)
//
// CTA = 4 x 8 node columns (128 threads, 2 CTAs/SM), marching up the ELEMENT LAYERS of its chunk; one layer of the
// (4+1) x (8+1) element footprint is resident in shared memory at a time.
//   phase 1   (thread = element x pair of Gauss points): g_b = sqrt(w|det J|) dN_b J^-1 for the 8 nodes -> S[gp][b][c][e].
//   passes    warp = 4 x 2 node columns x 4 in-plane element slots (sx, sy).  A thread works on ONE element of the layer for
//             BOTH nodes of its column that touch it (plane L through the element's bottom face, plane L+1 through the top
//             face).  Pass f (f = bottom / top face of b): 8 blocks G_ab = sum_gp g_a g_b' (72 fp64 accumulators), a in
//             {a_bot, a_top}, b in the 4 nodes of face f.  The 32 lanes of a warp touch only 15 distinct elements, which are
//             consecutive in the ring: every g_b load is ONE shared-memory wavefront (the 4x4x8-slot mapping of
//             assemble_tile.cu needed two), and g_a of the face being swept comes out of the loaded g_b registers.
//   carry     the dz = 0 blocks of plane L+1 (a_top, b top) stay in registers and keep accumulating in the next layer's
//             pass 1 (a_bot, b bottom): no shared-memory round trip across layers.
//   combine   the 4 slot threads of a node reduce their blocks onto the 9 in-plane neighbour blocks of one dz level with
//             27 64-bit shuffles (xor 2, xor 3, xor 1) instead of read-modify-write rounds through shared memory.
//   output    material applied in registers (K_ab = lam G + mu G' + mu tr(G) I; D(1,1) on the diagonal, src/fem.jl:230),
//             one dz level (<= 27 consecutive entries per CSR row) staged per warp in CSR order and copied out with
//             lane = position in the section: 216-byte contiguous stores of values and (fused assembly) column indices.
#include <cmath>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "smfem_internal.cuh"
#include "tile_args.cuh"

namespace {

// NTHv: threads per CTA (a warp owns 4 x 2 node columns: tile = 4 x 2 NTHv/32).  OPT bit 4: the node's own gradient on the swept
// face is loaded (conflict-free thanks to the pads) instead of selected out of the four loaded ones (18 selects per Gauss point);
// bit 8: the words of a coordinate plane a thread stages are computed once per CTA; bit 16: no register copy of the carried blocks
// before they are emitted (second instantiation of emit_level; costs registers, slower).  OPT bit 1: conflict-free shared-memory layout
// (rows of the 4 nodes of a face padded so that the per-slot g_a loads of a half-warp hit 16 distinct bank pairs; Gauss-point
// stride == 2 mod 16 and phase-1 tasks numbered element-major, so that the 8 Gauss points of an element read the same
// coordinates (broadcast) and store to 8 distinct bank pairs); bit 2: compile-time section strides on interior planes.
template <int NTHv, int OPTv>
struct T2 {
    static constexpr int TX = 4, NTH = NTHv, TY = 2 * (NTHv / 32), OPT = OPTv;
    static constexpr int EX = TX + 1, EY = TY + 1, NEL = EX * EY;  // 45 footprint elements per layer (4 x 8 tile)
    static constexpr int NELP = NEL;                               // element stride of a (b, c) row
    // pads (doubles) in front of the rows of the face nodes q = 1, 2, 3: with them the addresses a half-warp uses for "its own
    // node" (slot (sx, sy) -> node q(sx, sy) of element e - sx - EX sy) fall into 16 distinct bank pairs
    static constexpr int P1 = (OPT & 1) ? 2 : 0, P2 = (OPT & 1) ? 4 : 0, P3 = (OPT & 1) ? (NELP == 45 ? 4 : 12) : 0;
    static constexpr int FACE = 12 * NELP + P3;                    // doubles per face (4 nodes)
    static constexpr int GS = (OPT & 1) ? ((2 * FACE + 15) / 16) * 16 + 2 : 2 * FACE;  // Gauss-point stride
    static constexpr int LAYER = 8 * GS;                           // doubles: [gp][b][c][e]
    static constexpr int PX = TX + 2, PY = TY + 2, PLANE = PX * PY * 3;  // node-plane coordinate buffer (with halo)
    static constexpr int SN = 84;             // staging doubles per node and level (81 + pad; == 4 mod 16: see emit)
    static constexpr int STAGE_WARP = 8 * SN;  // a warp's 8 nodes
    static constexpr size_t SMEM_BYTES = sizeof(double) * (LAYER + (NTH / 32) * STAGE_WARP + 8 * 3 + 8 + 4 * PLANE);
    // offset of the rows of local node b (reference numbering, b = 4 face + q) inside a Gauss point's block
    __host__ __device__ static constexpr int boff(int b) {
        return (b >> 2) * FACE + (b & 3) * 3 * NELP + ((b & 3) == 0 ? 0 : ((b & 3) == 1 ? P1 : ((b & 3) == 2 ? P2 : P3)));
    }
    // slots (0,0), (1,0), (1,1), (0,1) own q = 0, 1, 2, 3 at element offsets 0, -1, -EX-1, -EX
    static constexpr int R1 = (boff(1) - 1) & 15, R2 = (boff(2) - EX - 1) & 15, R3 = (boff(3) - EX) & 15;
    static_assert(!(OPT & 1) || (R1 % 4 == 0 && R2 % 4 == 0 && R3 % 4 == 0 && R1 && R2 && R3 && R1 != R2 && R1 != R3 && R2 != R3),
                  "face-row pads do not give a conflict-free own-node load");
    static_assert(!(OPT & 1) || (GS % 16 == 2), "Gauss-point stride");
};

// element layer `layer` of the footprint -> S.  Same arithmetic as phase1 of assemble_tile.cu (register-only Q1 gradients,
// edge-form Jacobian, rsqrt overlapped with the unscaled gradients); src/fem.jl:192-196.
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

// One sweep over the 8 Gauss points for the 4 nodes b of one face of this thread's element (face 0: bottom = plane lay,
// face 1: top = plane lay + 1; Sf points at the face's g_b rows, So at the thread's own g_a row on the OTHER face).
//   Same[q]  += g_s g_b'   g_s = g_a of this column's node ON the swept face (one of the loaded g_b, selected by the slot):
//                          the dz = 0 blocks of plane lay + face
//   Other[q] += g_o g_b'   g_o = g_a of this column's node on the other face (loaded): the dz = 2 face - 1 blocks of
//                          plane lay + 1 - face                                    (q = in-plane reference number of b)
template <class T>
__device__ __forceinline__ void sweep(const double *Sf, const double *So, const double *Ss, int aq, double (&Same)[4][9], double (&Other)[4][9]) {
    constexpr int NELP = T::NELP;
#pragma unroll 2
    for (int gp = 0; gp < 8; ++gp) {
        const double *Sg = Sf + gp * T::GS;
        double gb[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int c = 0; c < 3; ++c) gb[q][c] = Sg[T::boff(q) + c * NELP];
        double gs[3], go[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // g_a on the swept face: one of the loaded g_b (3 x 6 selects), or - OPT 4 - its own conflict-free load
            gs[c] = (T::OPT & 4) ? Ss[gp * T::GS + c * NELP] : ((aq & 2) ? ((aq & 1) ? gb[3][c] : gb[2][c]) : ((aq & 1) ? gb[1][c] : gb[0][c]));
            go[c] = So[gp * T::GS + c * NELP];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    Same[q][i * 3 + j] += gs[i] * gb[q][j];
                    Other[q][i * 3 + j] += go[i] * gb[q][j];
                }
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

// Combine the 4 in-plane blocks X[q] of the 4 slot threads of every node onto the 9 neighbour blocks of ONE dz level, apply
// the material and park the level in the warp's staging area in CSR order; then copy it out: for each node and each of its
// 3 rows a section of len = 3 cx cy consecutive entries at  rowstart + lz * len.
// After the shuffles thread (sx, sy) holds   O0: (sy - sx, 0)  [not slot 3]   O1: (-sx, 1 - 2 sy)  [not slot 2]
//                                            O2: (1 - sx, 1 - 2 sy)  [not slot 1]
template <class T>
__device__ __forceinline__ void emit_level(const TileArgs &A, const NodeGeo &G, double (&X)[4][9], double *stage_w, int lane, int ix, int iy,
                                           int k, int dz, int jx0, int jy0) {
    const int slot = lane & 3, sx = slot & 1, sy = slot >> 1, nwl = lane >> 2;
    const Lattice &L = A.L;
    double *my = stage_w + nwl * T::SN;
    const int len = G.len_own;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        double g[9];
        if (!(A.skip & 4)) {
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
                    const double RB = shfl_xor_f64(sx ? Yb : Ya, 1);  // (issued twice, for o = 1 and o = 2: 9 extra shuffles, fewer live registers)
                    g[m] = o == 1 ? ((slot == 0) ? Ya + RB : Ya) : ((slot == 3) ? Yb + RB : Yb);
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < 9; ++m) g[m] = X[o][m] + X[3][m];
        }
        if (!(A.skip & (8 | 64)) && G.soff[o] >= 0) {
            const double tr = g[0] + g[4] + g[8];
            double *dst = my + G.soff[o];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const double gij = g[c * 3 + j], gji = g[j * 3 + c];
                    const double v = (c == j) ? A.mat.d11 * gij + A.mat.mu * (tr - gij) : A.mat.lam * gij + A.mat.mu * gji;
                    dst[c * len + j] = v;
                    if (o == 0 && c == j && dz == 0 && slot == 0)
                        A.diag[(((int64_t)(k - L.k0) * L.n1 + iy) * L.n1 + ix) * 3 + c] = v;
                }
        }
    }
    if (A.skip & 8) return;
    __syncwarp();
    // copy-out: lane = position inside the section
    const int cz = G.cnt(k), lz = dz + (k > 0), nz = k + dz;
    const int64_t planeoff = 9 * (G.pre(k) * G.S1 * G.S1 - G.pairs_base);
    const bool collapse = (A.skip & 32) != 0;  // ablation: same stores onto a small cache-resident window
    if (G.fast && !collapse) {
        // interior warp: 24 sections of 27 entries; all loads first, then the stores (the copy is latency-, not bandwidth-bound).
        // Lanes >= 27 re-read lanes 11..15's words: a different bank pair than anything lanes 16..26 touch (index 0 was a 2-way conflict)
        const int ll = (T::OPT & 1) ? (lane < 27 ? lane : lane - 16) : (lane < 27 ? lane : 0);
        double v[2][12];
#pragma unroll
        for (int yrow = 0; yrow < 2; ++yrow)
#pragma unroll
            for (int r = 0; r < 12; ++r)
                v[yrow][r] = (A.skip & 64) ? 1.0 : stage_w[(yrow * 4 + r / 3) * T::SN + (r % 3) * 27 + ll];
        if (lane < 27 && !(A.skip & 128)) {
            const int32_t colb = (int32_t)(L.lnode(jx0, jy0, nz) * 3) + G.crel27;
            // TRc: section stride = entries per row; a compile-time constant on interior planes (all offsets become immediates)
            auto copy_out = [&](auto trc) {
                constexpr int TRC = decltype(trc)::value;
                const int TR = TRC ? TRC : 27 * cz;
#pragma unroll
                for (int yrow = 0; yrow < 2; ++yrow) {
                    const int64_t g0 = planeoff + 9 * (int64_t)cz * G.rowc[yrow] + lz * 27 + lane;
                    double *vp = A.val + g0;
#pragma unroll
                    for (int r = 0; r < 12; ++r) vp[r * TR] = v[yrow][r];
                    if (A.colind) {
                        int32_t *cp = A.colind + g0;
                        const int32_t col0 = colb + yrow * 3 * L.n1;
#pragma unroll
                        for (int r = 0; r < 12; ++r) cp[r * TR] = col0 + 3 * (r / 3);
                    }
                }
            };
            if ((T::OPT & 2) && cz == 3) copy_out(std::integral_constant<int, 81>());
            else copy_out(std::integral_constant<int, 0>());
        }
    } else {
#pragma unroll 1
        for (int yrow = 0; yrow < 2; ++yrow) {
            const int jy = jy0 + yrow;
            if (jy >= L.n1) break;
            const int cyr = G.cnt(jy);
            int64_t base = planeoff + 9 * (int64_t)cz * (yrow ? G.rowc[1] : G.rowc[0]);
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
                    const double *src = stage_w + (yrow * 4 + node) * T::SN + lane;
                    const int64_t g0 = (collapse ? (base & 1023) + 2048 * (threadIdx.x >> 5) : base) + lz * lenn + lane;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        A.val[g0 + (int64_t)c * TR] = src[c * lenn];
                        if (A.colind) A.colind[g0 + (int64_t)c * TR] = col;
                    }
                }
                base += 3 * TR;
            }
        }
    }
    __syncwarp();  // the staging area is rewritten by the next level
}

template <class T, int MINB>
__global__ void __launch_bounds__(T::NTH, MINB) k_values_tile2(const __grid_constant__ TileArgs A) {
    using T2 = T;
    constexpr int NTH = T2::NTH, EX = T2::EX, LAYER = T2::LAYER;
    extern __shared__ double smem[];
    double *S = smem;                               // [gp][b][c][e]: the resident element layer
    double *s_stage = S + LAYER;                    // [warp][8 nodes][SN]
    double *s_gp = s_stage + (NTH / 32) * T2::STAGE_WARP;  // [gp][3]
    double *s_w = s_gp + 8 * 3;                     // sqrt of the Gauss weights
    double *s_xyz = s_w + 8;                        // [4][PLANE] node-plane coordinate ring
    const Lattice &L = A.L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < 8 * 3; t += NTH) s_gp[t] = (&A.gp[0][0])[t];
    if (tid < 8) s_w[tid] = sqrt(A.w[tid]);

    int bid = blockIdx.x;
    const int tix = bid % A.tiles_x;
    bid /= A.tiles_x;
    const int tiy = bid % A.tiles_y;
    const int chunk_id = bid / A.tiles_y;
    const int X0 = tix * T2::TX, Y0 = tiy * T2::TY;
    const int zs = L.k0 + A.zb[chunk_id], ze = L.k0 + A.zb[chunk_id + 1];  // owned node planes [zs, ze) of this CTA

    // identity of this thread: node column (ix, iy), in-plane element slot (sx, sy)
    const int slot = lane & 3, sx = slot & 1, sy = slot >> 1, nwl = lane >> 2;
    const int tx = nwl & 3, ty = 2 * warp + (nwl >> 2);
    const int ix = X0 + tx, iy = Y0 + ty;
    const bool node_ok = ix < L.n1 && iy < L.n1;
    const int ex = ix - sx, ey = iy - sy;
    const bool el_ok = node_ok && ex >= 0 && ey >= 0 && ex < L.ne && ey < L.ne;
    const int e = (ty - sy + 1) * EX + (tx - sx + 1);
    const int aq = sy ? (sx ? 2 : 3) : (sx ? 1 : 0);  // in-plane reference number of this node inside its element
    const double *Se = S + e;
    const int boff_aq = aq == 0 ? T2::boff(0) : (aq == 1 ? T2::boff(1) : (aq == 2 ? T2::boff(2) : T2::boff(3)));
    double *stage_w = s_stage + warp * T2::STAGE_WARP;
    const int jx0 = X0, jy0 = Y0 + 2 * warp;
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

    // coordinate staging: which words of a node plane this thread copies does not depend on the plane (OPT 8: computed once)
    constexpr int NSTG = (T2::PLANE + NTH - 1) / NTH;
    int stg_src[NSTG];
    if (T2::OPT & 8) {
#pragma unroll
        for (int q = 0; q < NSTG; ++q) {
            const int t = tid + q * NTH;
            const int c = t % 3, n = t / 3, px = n % T2::PX, py = n / T2::PX;
            const int gx = X0 - 1 + px, gy = Y0 - 1 + py;
            stg_src[q] = (t < T2::PLANE && gx >= 0 && gy >= 0 && gx < L.n1 && gy < L.n1) ? 3 * (gy * L.n1 + gx) + c : -1;
        }
    }
    auto stage = [&](int k) {
        if (!(T2::OPT & 8)) {
            stage_plane<T2>(A, s_xyz, k, X0, Y0);
            return;
        }
        if (k < 0 || k >= L.n1 || k > L.k1) return;  // the slab holds planes k0-1 .. k1
        const double *src = A.coords + 3 * (int64_t)(k - L.k0 + 1) * L.n1 * L.n1;
        double *dst = s_xyz + (k & 3) * T2::PLANE + tid;
#pragma unroll
        for (int q = 0; q < NSTG; ++q)
            if (stg_src[q] >= 0) {
                unsigned d = (unsigned)__cvta_generic_to_shared(dst + q * NTH);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + stg_src[q]) : "memory");
            }
    };
    const int L0 = max(zs - 1, 0), L1 = min(ze - 1, L.ne - 1);  // element layers this CTA sweeps (inclusive)
    wait_plane(min(L0 + 1, L.n1 - 1));
    stage(L0);
    stage(L0 + 1);

    double Same[4][9], Other[4][9];  // dz = 0 blocks (carried from a layer's top sweep into the next layer's bottom sweep) / dz = -+1 blocks
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int m = 0; m < 9; ++m) Same[q][m] = Other[q][m] = 0.0;

    // half-steps hs = 2 lay + face.  The top plane of the lattice has no element layer above it: one virtual half-step
    // (no sweep) emits its dz = 0 level.
    const int hs_end = 2 * L1 + 1 + (ze == L.n1 ? 1 : 0);
    for (int hs = 2 * L0; hs <= hs_end; ++hs) {
        const int lay = hs >> 1, face = hs & 1;
        const bool real = lay <= L1;
        if (face == 0 && real) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();  // everybody is done with the previous layer in S; coordinate planes lay, lay + 1 have landed
            if (lay + 2 < L.n1 && lay + 2 <= L.k1) wait_plane(lay + 2);
            if (lay + 1 <= L1) stage(lay + 2);  // lands during this layer's sweeps
            if (!(A.skip & 1)) phase1_layer<T2>(A, s_gp, s_w, s_xyz, S, lay, X0, Y0);
            __syncthreads();
        }
        if (real && el_ok && !(A.skip & 2))
            sweep<T2>(Se + face * T2::FACE, Se + (1 - face) * T2::FACE + boff_aq, Se + face * T2::FACE + boff_aq, aq, Same, Other);
        // face 0: Other = dz -1 of plane lay + 1, then Same = dz 0 of plane lay (complete);  face 1: Other = dz +1 of plane lay
        if (T2::OPT & 16) {  // the carried blocks are emitted from their own registers (no copy into Other; emit_level instantiated twice)
            const int po = lay + 1 - face, dzo = 2 * face - 1;
            if (po >= zs && po < ze) emit_level<T2>(A, G, Other, stage_w, lane, ix, iy, po, dzo, jx0, jy0);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int m = 0; m < 9; ++m) Other[q][m] = 0.0;
            if (face == 0) {
                if (lay >= zs && lay < ze) emit_level<T2>(A, G, Same, stage_w, lane, ix, iy, lay, 0, jx0, jy0);
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int m = 0; m < 9; ++m) Same[q][m] = 0.0;
            }
            continue;
        }
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
            int p = lay + 1 - face, dz = 2 * face - 1;
            if (rep == 1) {
                if (face != 0) break;
                p = lay;
                dz = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int m = 0; m < 9; ++m) {
                        Other[q][m] = Same[q][m];
                        Same[q][m] = 0.0;
                    }
            }
            if (p >= zs && p < ze) emit_level<T2>(A, G, Other, stage_w, lane, ix, iy, p, dz, jx0, jy0);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int m = 0; m < 9; ++m) Other[q][m] = 0.0;
        }
    }
}

}  // namespace

template <class T, int MINB>
static void launch_tile2(smfem_ctx *ctx, TileArgs &A, int nown) {
    static std::atomic<unsigned long long> attr_set{0};
    if (first_use_on_device(attr_set)) {
        CUDA_CHECK(cudaFuncSetAttribute(k_values_tile2<T, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES));
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
    // experiment knob: SMFEM_TILE_SMEM_PAD=<bytes> of extra dynamic shared memory per CTA lowers the number of resident CTAs
    // (occupancy sensitivity: how much slower is the kernel with 4 instead of 8 warps per SM?)
    size_t smem = T::SMEM_BYTES;
    if (const char *pad = std::getenv("SMFEM_TILE_SMEM_PAD")) {
        smem += (size_t)std::atol(pad);
        CUDA_CHECK(cudaFuncSetAttribute(k_values_tile2<T, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    CUDA_CHECK(cudaEventRecord(ctx->asm_ev[2 * slot], ctx->stream));
    LAUNCH(ctx, (k_values_tile2<T, MINB>), grid, T::NTH, smem, A);
    CUDA_CHECK(cudaEventRecord(ctx->asm_ev[2 * slot + 1], ctx->stream));
    ctx->asm_count++;
}

// The default structured value kernel; SMFEM_TILE = 4x4 / 8x4 / mma* select the earlier kernels (returns false then).
// SMFEM_TILE = v2 (default: 128-thread CTAs, OPT 15 = conflict-free layout + compile-time strides + own-gradient load + staging
// plan), v2base (the first layer-march version), v2l / v2i / v2li / v2g / v2p (subsets), v2s (64-thread CTAs, 4 per SM),
// v2e / v2all (carried blocks emitted from their own registers: 255 registers, slower): kept selectable for A/B timing
// (tools/time_tile2.py, profiles/r2_tile2_variants.txt); all give the same bits
bool values_assemble_tile2(smfem_ctx *ctx, TileArgs &A, int nown) {
    const char *sel = std::getenv("SMFEM_TILE");  // read per call: tests switch kernels inside one process
    const std::string m = (!sel || !sel[0]) ? "v2" : sel;
    if (m == "v2") launch_tile2<T2<128, 15>, 2>(ctx, A, nown);
    else if (m == "v2li") launch_tile2<T2<128, 3>, 2>(ctx, A, nown);
    else if (m == "v2g") launch_tile2<T2<128, 7>, 2>(ctx, A, nown);
    else if (m == "v2p") launch_tile2<T2<128, 11>, 2>(ctx, A, nown);
    else if (m == "v2e") launch_tile2<T2<128, 19>, 2>(ctx, A, nown);
    else if (m == "v2all") launch_tile2<T2<128, 31>, 2>(ctx, A, nown);
    else if (m == "v2base") launch_tile2<T2<128, 0>, 2>(ctx, A, nown);
    else if (m == "v2l") launch_tile2<T2<128, 1>, 2>(ctx, A, nown);
    else if (m == "v2i") launch_tile2<T2<128, 2>, 2>(ctx, A, nown);
    else if (m == "v2s") launch_tile2<T2<64, 3>, 4>(ctx, A, nown);
    else return false;
    return true;
}
