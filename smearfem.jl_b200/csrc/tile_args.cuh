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

// chunk planner and launch bookkeeping (assemble_tile.cu)
std::vector<int> plan_chunks(int ntiles, int nown, int slots);
void tile_fill_args(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, Material mat, bool write_colind, const int *ready, TileArgs &A);
// assemble_mma.cu: returns false when the DMMA kernel is not selected (env SMFEM_TILE)
bool values_assemble_mma(smfem_ctx *ctx, TileArgs &A, int nown);
