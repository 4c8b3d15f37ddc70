// Shared by the structured-lattice value kernels (assemble_tile.cu: scalar-FMA gather kernel, assemble_mma.cu: DMMA kernel).
#pragma once
#include <vector>

#include "smfem_internal.cuh"

struct Material {
    double d11, lam, mu;
};

constexpr int MAX_CHUNKS = 24;

struct TileArgs {
    Lattice L;
    const double *coords;
    const int64_t *rowptr;
    double *val;
    int32_t *colind;  // non-null: the output phase also writes the pattern's column indices (fused assembly)
    const int *ready;  // non-null: coordinate planes [0, *ready) have arrived (written by host->device copies while the kernel runs)
    double *diag;
    Material mat;
    int tiles_x, tiles_y, nchunks;
    int zb[MAX_CHUNKS + 1];  // chunk c of a tile column = owned planes [zb[c], zb[c+1]) (offsets from L.k0, longest first)
    int out_mode;  // output route of the tile kernel (env SMFEM_TILE_OUT): see the output phase
    int skip;  // ablation bitmask (env SMFEM_TILE_SKIP; profiling only): 1 phase 1, 2 main loop, 4 combine, 8 output,
               // 16 / 32: column-index / value stores collapsed onto a small cache-resident window (no DRAM traffic)
    double gp[8][3];     // the 8 Gauss points in the reference's order (src/fem.jl:174-176)
    double w[8];
    double sw[8];        // sqrt(w)
};

// asynchronous copy of node plane k (tile + 1-node halo, clipped to the lattice) into the coordinate ring
template <class T>
__device__ __forceinline__ void stage_plane(const TileArgs &A, double *s_xyz, int k, int X0, int Y0) {
    const Lattice &L = A.L;
    if (k < 0 || k >= L.n1 || k > L.k1) return;  // the slab holds planes k0-1 .. k1
    double *dst = s_xyz + (k & 3) * T::PLANE;
    for (int t = threadIdx.x; t < T::PLANE; t += T::NTH) {
        const int c = t % 3, n = t / 3;
        const int px = n % T::PX, py = n / T::PX;
        const int gx = X0 - 1 + px, gy = Y0 - 1 + py;
        if (gx < 0 || gy < 0 || gx >= L.n1 || gy >= L.n1) continue;
        const double *src = A.coords + 3 * L.lnode(gx, gy, k) + c;
        unsigned d = (unsigned)__cvta_generic_to_shared(dst + t);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    }
}

// column indices of the <= 4 nodes (jx0.., jy, k) of a warp in CSR order -> Ri[0 .. run_len)
__device__ __forceinline__ void write_colind_run(const TileArgs &A, int32_t *Ri, bool row_ok, int jx0, int jy, int k, int cy, int cz,
                                                 int dx, int dy, int dz, int lane) {
    const Lattice &L = A.L;
    int off = 0;
    for (int jn = 0; jn < 4; ++jn) {
        const int jx = jx0 + jn;
        if (!(row_ok && jx < L.n1)) break;
        const int cx = 1 + (jx > 0) + (jx < L.n1 - 1);
        const int TR = 3 * cx * cy * cz;
        const int nx = jx + dx, ny = jy + dy, nz = k + dz;
        if (lane < 27 && nx >= 0 && ny >= 0 && nz >= 0 && nx < L.n1 && ny < L.n1 && nz < L.n1) {
            const int rank = ((dz + (k > 0)) * cy + (dy + (jy > 0))) * cx + (dx + (jx > 0));
            const int32_t col = (int32_t)(L.lnode(nx, ny, nz) * 3);
            int32_t *dst = Ri + off + 3 * rank;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int j = 0; j < 3; ++j) dst[c * TR + j] = col + j;
        }
        off += 3 * TR;
    }
}

// ... and out to K.colind[run_base .. run_base + run_len) with one TMA bulk store (16-byte aligned middle; <= 3 + 3
// head / tail entries by plain stores).  Returns after the bulk copy has finished READING the region.
// MODE 0: wait for the bulk read before returning;  1 (DEFER): wait for EARLIER bulk reads first, return without waiting;
// 2: no wait at all (the caller drains its bulk groups itself)
template <int MODE>
__device__ __forceinline__ void emit_colind_run(const TileArgs &A, int32_t *Ri, int64_t run_base, int run_len, bool row_ok, int jx0,
                                                int jy, int k, int cy, int cz, int dx, int dy, int dz, int lane) {
    const int par4 = (int)(run_base & 3);  // the region mirrors the 16-byte phase of the destination
    if (A.skip & 16) run_base = par4 + 1024 * (threadIdx.x >> 5);  // ablation: same stores, collapsed onto a cache-resident window
    // DEFER: Ri is a buffer of its own; the bulk store of the previous plane has had a whole plane of compute to drain
    if (MODE == 1 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    write_colind_run(A, Ri + par4, row_ok, jx0, jy, k, cy, cz, dx, dy, dz, lane);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const int64_t a0 = (run_base + 3) & ~(int64_t)3, a1 = (run_base + run_len) & ~(int64_t)3;
    if (a1 > a0) {
        if (lane == 0) {
            const unsigned src = (unsigned)__cvta_generic_to_shared(Ri + par4 + (a0 - run_base));
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(A.colind + a0), "r"(src),
                         "r"((unsigned)((a1 - a0) * 4))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (MODE == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        } else if (lane < 4) {  // unaligned head (lanes 1..3) ...
            const int64_t t = run_base + (lane - 1);
            if (t < a0) A.colind[t] = Ri[par4 + (lane - 1)];
        } else if (lane < 7) {  // ... and tail (lanes 4..6)
            const int64_t t = a1 + (lane - 4);
            if (t < run_base + run_len) A.colind[t] = Ri[par4 + (t - run_base)];
        }
    } else {
        for (int t = lane; t < run_len; t += 32) A.colind[run_base + t] = Ri[par4 + t];
    }
}

// chunk planner and launch bookkeeping (assemble_tile.cu)
std::vector<int> plan_chunks(int ntiles, int nown, int slots);
void tile_fill_args(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, Material mat, bool write_colind, const int *ready, TileArgs &A);
// assemble_tile2.cu: the layer-march kernel; returns false when not selected (env SMFEM_TILE)
bool values_assemble_tile2(smfem_ctx *ctx, TileArgs &A, int nown);
// assemble_tile3.cu: the split-role layer-march kernel (SMFEM_TILE=v3)
bool values_assemble_tile3(smfem_ctx *ctx, TileArgs &A, int nown);
// assemble_mma.cu: returns false when the DMMA kernel is not selected (env SMFEM_TILE)
bool values_assemble_mma(smfem_ctx *ctx, TileArgs &A, int nown);
