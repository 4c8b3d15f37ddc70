// Solve side of the path: in-kernel Dirichlet conditions, CSR SpMV, Jacobi-preconditioned CG with
// fused vector kernels, halo exchange + dot-product all-reduce through peer memory (NVLink).
//   reference: examples/vector3D.jl:133-173 (setboundaryCond), :308-322 (solve; the dense inverse
//   of :318 is replaced by PCG on the same system K̄[free,free] q_f = -(K̄ q_d)[free]).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "smfem_internal.cuh"

// ------------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ld_stream_f64(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream_s32(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum (fixed tree); result valid in thread 0
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *smem /* NT/32 doubles */) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    double s = 0;
    if (threadIdx.x == 0)
        for (int i = 0; i < NT / 32; ++i) s += smem[i];
    return s;
}

// ------------------------------------------------------------------------------------------------
// All-reduce of up to 4 doubles through the peer windows: every rank stores its partial into slot
// [seq&3][rank] of EVERY rank's mailbox (direct NVLink stores), then releases a flag; readers spin
// on their own (local) mailbox and add the partials in rank order, so all ranks obtain bitwise
// identical sums.  Four slots (seq & 3), see CommHeader: no host barrier is needed anywhere, not even
// between consecutive solves.
// ------------------------------------------------------------------------------------------------
__device__ void allreduce_publish(const CommView &cv, unsigned long long seq, int nv, const double *vals) {
    int slot = (int)(seq & 3ull);
    for (int q = 0; q < cv.nranks; ++q) {
        CommHeader *h = cv.peer[q];
        for (int v = 0; v < nv; ++v) h->mbox[slot][cv.rank][v] = vals[v];
    }
    __threadfence_system();
    for (int q = 0; q < cv.nranks; ++q) st_release_sys(&cv.peer[q]->mflag[slot][cv.rank], seq);
}

// wait_ns (optional): accumulates the time this call spent waiting for the peers' flags
__device__ void allreduce_fetch(const CommView &cv, unsigned long long seq, int nv, double *out, unsigned long long *wait_ns = nullptr) {
    int slot = (int)(seq & 3ull);
    for (int v = 0; v < nv; ++v) out[v] = 0.0;
    const unsigned long long t0 = wait_ns ? global_timer_ns() : 0ull;
    for (int q = 0; q < cv.nranks; ++q) {
        unsigned spins = 0;
        while (ld_acquire_sys(&cv.self->mflag[slot][q]) != seq)
            if (++spins > (1u << 26)) __trap();  // a peer that died must surface as an error, not as a hang
        for (int v = 0; v < nv; ++v) out[v] += ld_volatile_f64(&cv.self->mbox[slot][q][v]);
    }
    if (wait_ns) *wait_ns += global_timer_ns() - t0;
}

// ------------------------------------------------------------------------------------------------
// Dirichlet data
// ------------------------------------------------------------------------------------------------
// examples/vector3D.jl:159-168: node with z == 0 -> q_d[3n] = 0, z == 1 -> q_d[3n] = -d (exact compares)
__global__ void k_dirichlet_zplanes(int64_t nNodes_l, const double *__restrict__ coords, int nDof, int64_t ghost_cols,
                                    int64_t nrows_l, double d, double *__restrict__ qd, uint8_t *__restrict__ fixed) {
    int64_t ln = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ln >= nNodes_l) return;
    double z = coords[3 * ln + 2];
    bool btm = (z == 0.0), top = (z == 1.0) && !btm;
    if (!(btm || top)) return;
    int64_t c = ln * nDof + 2;
    qd[c] = btm ? 0.0 : -d;
    int64_t r = c - ghost_cols;
    if (r >= 0 && r < nrows_l) fixed[r] = 1;
}

void dirichlet_zplanes(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, double d) {
    REQUIRE(K->ndim == 3 && K->nDof == 3, SMFEM_ERR_UNSUPPORTED, "setboundaryCond is 3-D / nDof=3 only (indexes coord[3], 3*nNode)");
    solver_alloc(ctx, K);
    CUDA_CHECK(cudaMemsetAsync(K->qd, 0, sizeof(double) * K->ncols_l, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(K->fixed, 0, K->nrows_l, ctx->stream));
    // NB the reference writes q_d[3*nNode] regardless of ID (examples/vector3D.jl:162,165)
    // ghost planes outside the domain carry z = 0 coordinates only when k0 == 0 / k1 == n1 planes are absent:
    // they are excluded by restricting the launch to in-domain local planes.
    int64_t first = 0, count = mesh->nNodes_l;
    if (mesh->structured) {
        const Lattice &L = mesh->lat;
        int lo = L.k0 - 1 < 0 ? 0 : L.k0 - 1, hi = L.k1 + 1 > L.n1 ? L.n1 : L.k1 + 1;
        first = (int64_t)(lo - (L.k0 - 1)) * L.plane();
        count = (int64_t)(hi - lo) * L.plane();
    }
    LAUNCH(ctx, k_dirichlet_zplanes, (unsigned)((count + 255) / 256), 256, 0, count, mesh->coords + 3 * first, K->nDof,
           K->ghost_cols - first * K->nDof, K->nrows_l, d, K->qd + first * K->nDof, K->fixed);
    K->has_bc = true;
}

__global__ void k_dirichlet_list(int64_t n, const int64_t *__restrict__ dofs, const double *__restrict__ vals,
                                 int64_t col_off_g, int64_t ncols_l, int64_t ghost_cols, int64_t nrows_l,
                                 double *__restrict__ qd, uint8_t *__restrict__ fixed) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int64_t c = dofs[t] - 1 - col_off_g;
    if (c < 0 || c >= ncols_l) return;
    qd[c] = vals[t];
    int64_t r = c - ghost_cols;
    if (r >= 0 && r < nrows_l) fixed[r] = 1;
}

void dirichlet_list(smfem_ctx *ctx, smfem_matrix *K, const int64_t *dofs, const double *vals, int64_t n) {
    solver_alloc(ctx, K);
    CUDA_CHECK(cudaMemsetAsync(K->qd, 0, sizeof(double) * K->ncols_l, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(K->fixed, 0, K->nrows_l, ctx->stream));
    if (n > 0) {
        for (int64_t i = 0; i < n; ++i)
            REQUIRE(dofs[i] >= 1 && dofs[i] <= K->m_g, SMFEM_ERR_INVALID, "set_dirichlet: dof id out of range");
        int64_t *d_dofs = dev_alloc<int64_t>(n);
        double *d_vals = dev_alloc<double>(n);
        CUDA_CHECK(cudaMemcpyAsync(d_dofs, dofs, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(d_vals, vals, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(ctx, k_dirichlet_list, (unsigned)((n + 255) / 256), 256, 0, n, d_dofs, d_vals, K->row0 - K->ghost_cols,
               K->ncols_l, K->ghost_cols, K->nrows_l, K->qd, K->fixed);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        dev_free(d_dofs);
        dev_free(d_vals);
    }
    K->has_bc = true;
}

// ------------------------------------------------------------------------------------------------
// K6: CSR SpMV.
// Variant 0 ("w3"): one warp streams the contiguous nonzeros of 3 consecutive rows (the three
// dofs of a node on the hex lattice: 3 x 81 = 243 values -> 95 % lane use instead of 81/96), with
// L1-bypassing streaming loads for val/colind and cached gathers of x.  Variant 1: warp per row.
// MODE bit 0: Dirichlet row mask; bit 1: fused dot(x_row, y) partial; bit 2: wait for halo flags.
// ------------------------------------------------------------------------------------------------
struct SpmvArgs {
    int64_t nrows, ghost_cols;
    const int64_t *rowptr;
    const int32_t *colind;
    const double *val;
    const double *x;  // ncols_l
    double *y;        // nrows
    const uint8_t *fixed;
    double *partials;
    PcgScalars *scal;
    CommView cv;
    int64_t rot;  // warp rotation so that boundary-plane rows run last
    int64_t row_begin, row_end;  // k_spmv_group: rows [row_begin, row_end) (multiples of the group size)
    int lat_n1;  // > 0: K is the hex-lattice matrix (nDof 3): the 81 columns of an interior node's rows are closed-form
    int check_done;                 // inside the PCG graph: return at once when scal->done is set
    unsigned long long halo_need;   // > 0: wait for this halo sequence number instead of scal->it + 1 (bench_spmv)
};
#define SPMV_DONE_CHECK(A)                                 \
    do {                                                   \
        if ((A).check_done && (A).scal->done) return;      \
    } while (0)
#define SPMV_HALO_NEED(A) ((A).halo_need ? (A).halo_need : (A).scal->it + 1)

template <int RPW, int MODE>
__global__ void __launch_bounds__(256) k_spmv(SpmvArgs A) {
    SPMV_DONE_CHECK(A);
    constexpr bool MASK = MODE & 1, DOT = MODE & 2, HALO = MODE & 4;
    __shared__ double s_red[8];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (A.nrows + RPW - 1) / RPW;
    int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double dot = 0.0;
    if (wid < nwarps) {
        if (HALO) wid = (wid + A.rot) % nwarps;
        const int64_t r0 = wid * RPW;
        const int nr = (int)((A.nrows - r0) < RPW ? (A.nrows - r0) : RPW);
        if (HALO && A.cv.nranks > 1) {
            // rows of the first / last owned plane read ghost planes written by the neighbours
            const unsigned long long need = SPMV_HALO_NEED(A);
            bool lo = (A.cv.rank > 0) && (r0 < A.cv.plane_dofs);
            bool hi = (A.cv.rank < A.cv.nranks - 1) && (r0 + nr > A.nrows - A.cv.plane_dofs);
            if (lane == 0) {
                if (lo)
                    while (ld_acquire_sys(&A.cv.self->hflag[0]) < need) {
                    }
                if (hi)
                    while (ld_acquire_sys(&A.cv.self->hflag[1]) < need) {
                    }
            }
            __syncwarp();
        }
        int64_t b[RPW + 1];
#pragma unroll
        for (int i = 0; i <= RPW; ++i) b[i] = A.rowptr[r0 + (i < nr ? i : nr)];
        double acc[RPW];
#pragma unroll
        for (int i = 0; i < RPW; ++i) acc[i] = 0.0;
#pragma unroll 4
        for (int64_t p = b[0] + lane; p < b[RPW]; p += 32) {
            double v = ld_stream_f64(A.val + p);
            int c = ld_stream_s32(A.colind + p);
            double prod = v * A.x[c];
            if (RPW == 1) {
                acc[0] += prod;
            } else {
#pragma unroll
                for (int i = 0; i < RPW; ++i)
                    if (p >= b[i] && p < b[i + 1]) acc[i] += prod;
            }
        }
#pragma unroll
        for (int i = 0; i < RPW; ++i) acc[i] = warp_sum(acc[i]);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < RPW; ++i)
                if (i < nr) {
                    double yv = acc[i];
                    if (MASK && A.fixed[r0 + i]) yv = 0.0;
                    A.y[r0 + i] = yv;
                    if (DOT) dot += yv * A.x[A.ghost_cols + r0 + i];
                }
        }
    }
    if (DOT) {
        double s = block_sum<256>(dot, s_red);
        if (threadIdx.x == 0) {
            A.partials[blockIdx.x] = s;
            __threadfence();
            unsigned t = atomicAdd(&A.scal->ticketB, 1u);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double v = 0.0;
            for (int64_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += ld_volatile_f64(A.partials + i);
            double tot = block_sum<256>(v, s_red);
            if (threadIdx.x == 0) {
                A.scal->ticketB = 0;
                allreduce_publish(A.cv, 2ull * A.scal->it + 1ull, 1, &tot);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Variant 2 ("stream", default): CSR-stream.  The nonzeros are cut into blocks of ~SPMV_CH entries
// aligned to row boundaries (blk_row[], built once per pattern).  A persistent CTA takes blocks
// cyclically; phase 1 streams val/colind of the whole block with perfectly coalesced L1-bypassing
// loads (16 independent loads in flight per thread), multiplies by the gathered x and parks the
// products in shared memory; phase 2 reduces each row from shared memory with one warp per row.
// Row length never matters for coalescing or lane use, the grid is ~6 CTAs/SM so the fused dot
// needs < 1k partials, and no atomics are involved.
// ------------------------------------------------------------------------------------------------
constexpr int SPMV_CH = 1776;      // target nonzeros per block: CH + SLACK + 2*3 alignment strangers <= 2048 = 2 quads/thread
constexpr int SPMV_SLACK = 256;    // max row length supported by the stream kernel
constexpr int SPMV_NT = 256;
constexpr int SPMV_MAXROWS = 512;  // max rows per block handled by the stream kernel (2 per thread)

__global__ void k_spmv_blockrows(int64_t nrows, const int64_t *__restrict__ rowptr, int nblk, int32_t *__restrict__ blk_row) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nblk) return;
    if (b == nblk) {
        blk_row[b] = (int32_t)nrows;
        return;
    }
    int64_t target = (int64_t)b * SPMV_CH;
    int64_t lo = 0, hi = nrows;  // first row r with rowptr[r] >= target
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rowptr[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    blk_row[b] = (int32_t)lo;
}

__global__ void k_max_rowlen(int64_t nrows, const int64_t *__restrict__ rowptr, int *__restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrows) atomicMax(out, (int)(rowptr[r + 1] - rowptr[r]));
}

__device__ __forceinline__ double2 ld_stream_v2f64(const double *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_stream_v4s32(const int32_t *p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(SPMV_NT, 4) k_spmv_stream(SpmvArgs A, const int32_t *__restrict__ blk_row, int nblk) {
    SPMV_DONE_CHECK(A);
    constexpr bool MASK = MODE & 1, DOT = MODE & 2, HALO = MODE & 4;
    // 128-bit streaming loads: a plain read stream with 8 B/lane loads tops out at 4.84 TB/s on B200, with
    // 16 B/lane at 6.86 TB/s (tools/microbench/peaks.cu).  A block [p0, p0+n) is read as aligned quads starting at
    // pa = p0 & ~3; the <= 3 leading and trailing strangers are multiplied too but never summed.
    constexpr int QPT = (SPMV_CH + SPMV_SLACK + 8 + 4 * SPMV_NT - 1) / (4 * SPMV_NT);  // quads per thread (2)
    constexpr int RPT = SPMV_MAXROWS / SPMV_NT;                                         // rows per thread in the epilogue
    __shared__ __align__(16) double s_prod[QPT * 4 * SPMV_NT];
    __shared__ double s_y[SPMV_MAXROWS];
    __shared__ int s_rp[SPMV_MAXROWS + 1];
    __shared__ double s_red[SPMV_NT / 32];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double dot = 0.0;
    // Software pipeline: while block i is reduced out of shared memory, the registers already hold block
    // i+1 (val, colind, its row pointers, Dirichlet flags and x_row) and the bounds of block i+2.
    auto block_of = [&](int b0) { return HALO ? (int)(((int64_t)b0 + A.rot) % nblk) : b0; };
    int b0 = blockIdx.x;
    int R0 = 0, R1 = 0, n = 0;
    int64_t p0 = 0;
    double2 rva[QPT], rvb[QPT];
    int4 rc[QPT];
    int rrp[RPT + 1];
    double rxr[RPT];
    bool rfx[RPT];
    auto load_bounds = [&](int bb, int &r0, int &r1, int64_t &q0, int &cnt) {
        if (bb < nblk) {
            const int b = block_of(bb);
            r0 = blk_row[b];
            r1 = blk_row[b + 1];
            q0 = A.rowptr[r0];
            cnt = (int)(A.rowptr[r1] - q0);
        } else {
            r0 = r1 = 0;
            q0 = 0;
            cnt = 0;
        }
    };
    auto issue_loads = [&](int r0, int r1, int64_t q0, int cnt) {
        const int64_t pa = q0 & ~(int64_t)3;
        const int cnt4 = (int)(q0 - pa) + cnt;  // entries from the aligned start
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            const int t = 4 * (tid + j * SPMV_NT);
            if (t < cnt4) {
                rc[j] = ld_stream_v4s32(A.colind + pa + t);
                rva[j] = ld_stream_v2f64(A.val + pa + t);
                rvb[j] = ld_stream_v2f64(A.val + pa + t + 2);
            }
        }
        const int nr = r1 - r0;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int t = tid + j * SPMV_NT;
            if (t < nr) {
                rrp[j] = (int)(A.rowptr[r0 + t] - pa);
                if (MASK) rfx[j] = A.fixed[r0 + t] != 0;
                if (DOT) rxr[j] = A.x[A.ghost_cols + r0 + t];
            }
        }
    };
    load_bounds(b0, R0, R1, p0, n);
    issue_loads(R0, R1, p0, n);
    int nR0, nR1, nn;
    int64_t np0;
    load_bounds(b0 + gridDim.x, nR0, nR1, np0, nn);
    for (; b0 < nblk; b0 += gridDim.x) {
        if (HALO && A.cv.nranks > 1) {
            const bool lo = (A.cv.rank > 0) && (R0 < A.cv.plane_dofs);
            const bool hi = (A.cv.rank < A.cv.nranks - 1) && (R1 > A.nrows - A.cv.plane_dofs);
            if ((lo || hi) && tid == 0) {
                const unsigned long long need = SPMV_HALO_NEED(A);
                if (lo)
                    while (ld_acquire_sys(&A.cv.self->hflag[0]) < need) {
                    }
                if (hi)
                    while (ld_acquire_sys(&A.cv.self->hflag[1]) < need) {
                    }
            }
            if (lo || hi) __syncthreads();
        }
        // registers -> gather x -> products (and the block's row pointers) into shared memory
        const int nr = R1 - R0, cR0 = R0;
        const int cnt4 = (int)(p0 & 3) + n;
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            const int t = 4 * (tid + j * SPMV_NT);
            if (t < cnt4) {
                const double x0 = A.x[rc[j].x], x1 = A.x[rc[j].y], x2 = A.x[rc[j].z], x3 = A.x[rc[j].w];
                *reinterpret_cast<double2 *>(s_prod + t) = make_double2(rva[j].x * x0, rva[j].y * x1);
                *reinterpret_cast<double2 *>(s_prod + t + 2) = make_double2(rvb[j].x * x2, rvb[j].y * x3);
            }
        }
        bool cfx[RPT];
        double cxr[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int t = tid + j * SPMV_NT;
            if (t < nr) s_rp[t] = rrp[j];
            cfx[j] = MASK ? rfx[j] : false;
            cxr[j] = DOT ? rxr[j] : 0.0;
        }
        if (tid == 0) s_rp[nr] = cnt4;
        __syncthreads();
        // prefetch the next block into registers, and the bounds of the one after
        R0 = nR0, R1 = nR1, p0 = np0, n = nn;
        issue_loads(R0, R1, p0, n);
        load_bounds(b0 + 2 * gridDim.x, nR0, nR1, np0, nn);
        // row sums out of shared memory: one warp per row
        for (int r = warp; r < nr; r += SPMV_NT / 32) {
            const int a = s_rp[r], e = s_rp[r + 1];
            double s = 0.0;
            for (int t = a + lane; t < e; t += 32) s += s_prod[t];
            s = warp_sum(s);
            if (lane == 0) s_y[r] = s;
        }
        __syncthreads();
        // epilogue: one thread per row, coalesced store of y, Dirichlet mask, fused p'Ap
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int t = tid + j * SPMV_NT;
            if (t < nr) {
                double yv = s_y[t];
                if (MASK && cfx[j]) yv = 0.0;
                A.y[cR0 + t] = yv;
                if (DOT) dot += yv * cxr[j];
            }
        }
    }
    if (DOT) {
        double s = block_sum<SPMV_NT>(dot, s_red);
        if (tid == 0) {
            A.partials[blockIdx.x] = s;
            __threadfence();
            unsigned t = atomicAdd(&A.scal->ticketB, 1u);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double v = 0.0;
            for (int i = tid; i < (int)gridDim.x; i += SPMV_NT) v += ld_volatile_f64(A.partials + i);
            double tot = block_sum<SPMV_NT>(v, s_red);
            if (tid == 0) {
                A.scal->ticketB = 0;
                allreduce_publish(A.cv, 2ull * A.scal->it + 1ull, 1, &tot);
            }
        }
    }
}

// also records the largest number of rows of any block (the stream kernel handles <= SPMV_MAXROWS)
__global__ void k_spmv_maxrows(int nblk, const int32_t *__restrict__ blk_row, int *__restrict__ out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nblk) atomicMax(out, blk_row[b + 1] - blk_row[b]);
}


// ------------------------------------------------------------------------------------------------
// Variant 3 ("tma", default): the CSR-stream kernel with the val / colind / rowptr streams moved by
// the TMA engine (cp.async.bulk global -> shared, completion on an mbarrier) instead of per-thread
// loads.  ncu on variant 2 showed the L1TEX sector pipeline 68 % busy (201 M sectors for 3 GB of DRAM
// data: 8 + 4 sectors per 32 nonzeros for val/colind on top of ~13 for the x gather); bulk copies
// bypass L1TEX, so it only serves the gather.  One producer warp runs STAGES blocks ahead of 8
// consumer warps; products are formed in place in the val stage, rows reduced by one warp each.
// ------------------------------------------------------------------------------------------------
constexpr int TMA_STAGES = 3;
constexpr int TMA_CAP = SPMV_CH + SPMV_SLACK + 16;  // entries per stage (2320)
constexpr int TMA_ROWCAP = SPMV_MAXROWS + 4;        // row pointers per stage
constexpr int TMA_NT = SPMV_NT + 32;                // 8 consumer warps + 1 producer warp

struct TmaStage {
    alignas(128) double val[TMA_CAP];
    alignas(128) int col[TMA_CAP];
    alignas(128) long long rp[TMA_ROWCAP];
};
struct TmaSmem {
    TmaStage st[TMA_STAGES];
    alignas(16) double y[SPMV_MAXROWS];
    alignas(16) int bounds[TMA_STAGES][8];  // R0, R1, off, n, rpoff
    alignas(16) long long p0[TMA_STAGES];
    alignas(8) unsigned long long full[TMA_STAGES], empty[TMA_STAGES];
    double red[SPMV_NT / 32];
    int last;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(TMA_NT, 2) k_spmv_tma(SpmvArgs A, const int32_t *__restrict__ blk_row, int nblk) {
    SPMV_DONE_CHECK(A);
    constexpr bool MASK = MODE & 1, DOT = MODE & 2, HALO = MODE & 4;
    constexpr int PER = (SPMV_CH + SPMV_SLACK + SPMV_NT - 1) / SPMV_NT;  // 9
    constexpr int RPT = SPMV_MAXROWS / SPMV_NT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TmaSmem &S = *reinterpret_cast<TmaSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nmine = (nblk - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // blocks of this CTA

    if (warp == SPMV_NT / 32) {
        // ------------------------------ producer warp ----------------------------------------------------------
        // The 32 lanes fetch the bounds of 32 consecutive blocks in parallel (the blk_row -> rowptr chain is two
        // dependent global loads, ~1.5 us: one lane doing them per block capped the ring at ~3.5 TB/s), then
        // lane 0 issues the bulk copies from the shuffled values as stages become free.
        auto fetch = [&](int i, int &R0, int &R1, long long &p0, int &n) {
            R0 = R1 = n = 0;
            p0 = 0;
            if (i < nmine) {
                const int b0 = blockIdx.x + i * gridDim.x;
                const int b = HALO ? (int)(((int64_t)b0 + A.rot) % nblk) : b0;
                R0 = blk_row[b];
                R1 = blk_row[b + 1];
                p0 = A.rowptr[R0];
                n = (int)(A.rowptr[R1] - p0);
            }
        };
        int nR0, nR1, nn;
        long long np0;
        fetch(lane, nR0, nR1, np0, nn);
        for (int base = 0; base < nmine; base += 32) {
            const int cR0 = nR0, cR1 = nR1, cn = nn;
            const long long cp0 = np0;
            fetch(base + 32 + lane, nR0, nR1, np0, nn);  // next batch in flight while this one is issued
            const int lim = min(32, nmine - base);
            for (int j = 0; j < lim; ++j) {
                const int R0 = __shfl_sync(0xffffffffu, cR0, j), R1 = __shfl_sync(0xffffffffu, cR1, j);
                const int n = __shfl_sync(0xffffffffu, cn, j);
                const long long p0 = __shfl_sync(0xffffffffu, cp0, j);
                if (lane == 0) {
                    const int i = base + j;
                    const int st = i % TMA_STAGES;
                    if (i >= TMA_STAGES) mbar_wait(&S.empty[st], ((i / TMA_STAGES) - 1) & 1);
                    const long long pa = p0 & ~3ll;  // 32 B / 16 B aligned sources
                    const int off = (int)(p0 - pa);
                    const unsigned cnt4 = (unsigned)((off + n + 3) & ~3);
                    const int ra = R0 & ~1;
                    const int rpoff = R0 - ra;
                    const unsigned nrp = (unsigned)((rpoff + (R1 - R0) + 1 + 1) & ~1);
                    S.bounds[st][0] = R0;
                    S.bounds[st][1] = R1;
                    S.bounds[st][2] = off;
                    S.bounds[st][3] = n;
                    S.bounds[st][4] = rpoff;
                    S.p0[st] = p0;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&S.full[st], cnt4 * 12u + nrp * 8u);
                    bulk_g2s(S.st[st].val, A.val + pa, cnt4 * 8u, &S.full[st]);
                    bulk_g2s(S.st[st].col, A.colind + pa, cnt4 * 4u, &S.full[st]);
                    bulk_g2s(S.st[st].rp, A.rowptr + ra, nrp * 8u, &S.full[st]);
                }
                __syncwarp();
            }
        }
        return;
    }

    // ---------------------------------- consumers: 8 warps ---------------------------------------------------
    double dot = 0.0;
    for (int i = 0; i < nmine; ++i) {
        const int st = i % TMA_STAGES;
        mbar_wait(&S.full[st], (i / TMA_STAGES) & 1);
        const int R0 = S.bounds[st][0], R1 = S.bounds[st][1], off = S.bounds[st][2], n = S.bounds[st][3],
                  rpoff = S.bounds[st][4];
        const long long p0 = S.p0[st];
        const int nr = R1 - R0;
        if (HALO && A.cv.nranks > 1) {
            const bool lo = (A.cv.rank > 0) && (R0 < A.cv.plane_dofs);
            const bool hi = (A.cv.rank < A.cv.nranks - 1) && (R1 > A.nrows - A.cv.plane_dofs);
            if ((lo || hi) && tid == 0) {
                const unsigned long long need = SPMV_HALO_NEED(A);
                if (lo)
                    while (ld_acquire_sys(&A.cv.self->hflag[0]) < need) {
                    }
                if (hi)
                    while (ld_acquire_sys(&A.cv.self->hflag[1]) < need) {
                    }
            }
            if (lo || hi) asm volatile("bar.sync 1, %0;" ::"n"(SPMV_NT) : "memory");
        }
        // row metadata for the epilogue (latency hidden behind the products + row sums)
        bool cfx[RPT];
        double cxr[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int t = tid + j * SPMV_NT;
            cfx[j] = false;
            cxr[j] = 0.0;
            if (t < nr) {
                if (MASK) cfx[j] = A.fixed[R0 + t] != 0;
                if (DOT) cxr[j] = A.x[A.ghost_cols + R0 + t];
            }
        }
        // products in place: val[t] *= x[col[t]]
        double *v = S.st[st].val + off;
        const int *c = S.st[st].col + off;
        double xv[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int t = tid + j * SPMV_NT;
            if (t < n) xv[j] = A.x[c[t]];
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int t = tid + j * SPMV_NT;
            if (t < n) v[t] *= xv[j];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(SPMV_NT) : "memory");
        const long long *rp = S.st[st].rp + rpoff;
        for (int r = warp; r < nr; r += SPMV_NT / 32) {
            const int a = (int)(rp[r] - p0), e = (int)(rp[r + 1] - p0);
            double s = 0.0;
            for (int t = a + lane; t < e; t += 32) s += v[t];
            s = warp_sum(s);
            if (lane == 0) S.y[r] = s;
        }
        // generic-proxy writes (in-place products) must be ordered before the TMA refills this stage
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(SPMV_NT) : "memory");
        if (tid == 0) mbar_arrive(&S.empty[st]);  // the stage may be refilled
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int t = tid + j * SPMV_NT;
            if (t < nr) {
                double yv = S.y[t];
                if (MASK && cfx[j]) yv = 0.0;
                A.y[R0 + t] = yv;
                if (DOT) dot += yv * cxr[j];
            }
        }
        // S.y is rewritten only after the next block's first consumer barrier: no extra sync needed
    }
    if (DOT) {
        double s = warp_sum(dot);
        if (lane == 0) S.red[warp] = s;
        asm volatile("bar.sync 1, %0;" ::"n"(SPMV_NT) : "memory");
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < SPMV_NT / 32; ++w) tot += S.red[w];
            A.partials[blockIdx.x] = tot;
            __threadfence();
            unsigned t = atomicAdd(&A.scal->ticketB, 1u);
            S.last = (t == gridDim.x - 1) ? 1 : 0;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(SPMV_NT) : "memory");
        if (S.last) {
            __threadfence();
            double vv = 0.0;
            for (int k = tid; k < (int)gridDim.x; k += SPMV_NT) vv += ld_volatile_f64(A.partials + k);
            vv = warp_sum(vv);
            asm volatile("bar.sync 1, %0;" ::"n"(SPMV_NT) : "memory");
            if (lane == 0) S.red[warp] = vv;
            asm volatile("bar.sync 1, %0;" ::"n"(SPMV_NT) : "memory");
            if (tid == 0) {
                double tot = 0.0;
                for (int w = 0; w < SPMV_NT / 32; ++w) tot += S.red[w];
                A.scal->ticketB = 0;
                allreduce_publish(A.cv, 2ull * A.scal->it + 1ull, 1, &tot);
            }
        }
    }
}

template <int MODE>
static void launch_spmv_tma(smfem_ctx *ctx, smfem_matrix *K, const SpmvArgs &A) {
    static std::atomic<unsigned long long> attr_set{0};
    if (first_use_on_device(attr_set)) {
        CUDA_CHECK(cudaFuncSetAttribute(k_spmv_tma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TmaSmem)));
    }
    int grid = ctx->sms * 2;
    if (grid > K->nblk) grid = K->nblk;
    LAUNCH(ctx, (k_spmv_tma<MODE>), grid, TMA_NT, sizeof(TmaSmem), A, (const int32_t *)K->blk_row, K->nblk);
}


// ------------------------------------------------------------------------------------------------
// Variant 4 ("group3", default when applicable): rows come in triples with identical column patterns (the
// three dofs of a node; checked once per pattern by k_check_group3).  One warp takes a triple: colind
// is read and x gathered ONCE per column slot and used for the three rows.  ncu on the other variants
// showed the L1TEX pipeline, not DRAM, as the limiter: per 32 nonzeros ~13 gather sectors + 12
// val/colind sectors; sharing the gather and colind halves the L1 wavefronts and cuts DRAM traffic from
// 12 to 9.33 B per nonzero (the matrix stays plain CSR; rows whose patterns differ use variant 2).
// Persistent grid: each warp walks triples with stride #warps, next triple's row pointer prefetched.
// ------------------------------------------------------------------------------------------------
__global__ void k_check_group3(int64_t ngroups, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                               int *__restrict__ bad) {
    int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (g >= ngroups) return;
    int64_t b0 = rowptr[3 * g], b1 = rowptr[3 * g + 1], b2 = rowptr[3 * g + 2], b3 = rowptr[3 * g + 3];
    int64_t L = b1 - b0;
    if (b2 - b1 != L || b3 - b2 != L) {
        if (lane == 0) *bad = 1;
        return;
    }
    for (int64_t s = lane; s < L; s += 32) {
        int c = colind[b0 + s];
        if (colind[b1 + s] != c || colind[b2 + s] != c) *bad = 1;
    }
}

template <int MODE, int G>
__global__ void __launch_bounds__(256, 3) k_spmv_group(SpmvArgs A) {
    SPMV_DONE_CHECK(A);
    // G = 3: row triples sharing one column pattern (k_check_group3);  G = 1: any CSR matrix, one row per warp step.
    constexpr bool MASK = MODE & 1, DOT = MODE & 2, HALO = MODE & 4;
    __shared__ double s_red[8];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31;
    const int64_t ngroups = A.row_end / G;  // groups [row_begin/G, row_end/G)
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t g = A.row_begin / G + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    double dot = 0.0;
    auto group_of = [&](int64_t gi) { return gi; };
    // register sets: [cur] is being reduced while [nxt] is in flight (warp-level software pipeline)
    int cc[3], nc[3];
    double cw[3][G], nw[3][G];
    int64_t cb0 = 0, nb0 = 0, fb0 = 0, fb1 = 0;  // cur / next / future row pointers
    int cL = 0, nL = 0;
    int64_t cg = 0, ng_ = 0, fg = 0;
    // hex lattice: a row triple of length 81 belongs to an interior node, whose columns are 3*(node + dz n1^2 + dy n1 + dx) + j
    // in slot order - no colind stream for ~94 % of the rows (0.33 of 2.47 GB at 100^3); other lengths read colind as usual
    int off[3] = {0, 0, 0};
    const bool lattice = (G == 3) && A.lat_n1 > 0;
    if (lattice) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int s = lane + 32 * j, d = s / 3;
            off[j] = 3 * (((d / 9 - 1) * A.lat_n1 + ((d / 3) % 3 - 1)) * A.lat_n1 + (d % 3 - 1)) + (s - 3 * d);
        }
    }
    auto issue = [&](int64_t gi, int64_t b0, int L, int (&c)[3], double (&w)[3][G]) {
        const double *__restrict__ v0 = A.val + b0;
        const int32_t *__restrict__ c0 = A.colind + b0;
        const bool closed = lattice && L == 81;
        const int cbase = (int)(A.ghost_cols + G * gi);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int s = lane + 32 * j;
            if (s < L) {
                c[j] = closed ? cbase + off[j] : ld_stream_s32(c0 + s);
#pragma unroll
                for (int q = 0; q < G; ++q) w[j][q] = ld_stream_f64(v0 + (int64_t)q * L + s);
            }
        }
    };
    // prologue: bounds of the first two groups, loads of the first
    if (g < ngroups) {
        cg = group_of(g);
        cb0 = A.rowptr[G * cg];
        cL = (int)(A.rowptr[G * cg + 1] - cb0);
        issue(cg, cb0, cL, cc, cw);
    }
    if (g + nwarps < ngroups) {
        fg = group_of(g + nwarps);
        fb0 = A.rowptr[G * fg];
        fb1 = A.rowptr[G * fg + 1];
    }
    for (; g < ngroups; g += nwarps) {
        const int64_t r0 = G * cg;
        if (HALO && A.cv.nranks > 1) {
            const bool lo = (A.cv.rank > 0) && (r0 < A.cv.plane_dofs);
            const bool hi = (A.cv.rank < A.cv.nranks - 1) && (r0 + G > A.nrows - A.cv.plane_dofs);
            if (lo || hi) {
                if (lane == 0) {
                    const unsigned long long need = SPMV_HALO_NEED(A);
                    const bool rec = blockIdx.x == 0 && threadIdx.x == 0;  // wait accounting: the first warp of the launch
                    const unsigned long long t0 = rec ? global_timer_ns() : 0ull;
                    if (lo)
                        while (ld_acquire_sys(&A.cv.self->hflag[0]) < need) {
                        }
                    if (hi)
                        while (ld_acquire_sys(&A.cv.self->hflag[1]) < need) {
                        }
                    if (rec) A.scal->t_wait_halo += global_timer_ns() - t0;
                }
                __syncwarp();
            }
        }
        // gathers for the current group (its colind registers have landed or are about to)
        double xv[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) xv[j] = (lane + 32 * j < cL) ? A.x[cc[j]] : 0.0;
        // next group: its bounds were loaded one iteration ago -> issue its streams now; fetch future bounds
        const bool have_next = g + nwarps < ngroups;
        if (have_next) {
            ng_ = fg;
            nb0 = fb0;
            nL = (int)(fb1 - fb0);
            issue(ng_, nb0, nL, nc, nw);
        }
        if (g + 2 * nwarps < ngroups) {
            fg = group_of(g + 2 * nwarps);
            fb0 = A.rowptr[G * fg];
            fb1 = A.rowptr[G * fg + 1];
        }
        double acc[G];
#pragma unroll
        for (int q = 0; q < G; ++q) acc[q] = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (lane + 32 * j < cL) {
#pragma unroll
                for (int q = 0; q < G; ++q) acc[q] += cw[j][q] * xv[j];
            }
        if (cL > 96) {  // long rows (> 32 neighbour nodes): plain tail loop
            const double *__restrict__ v0 = A.val + cb0;
            const int32_t *__restrict__ c0 = A.colind + cb0;
            for (int s = lane + 96; s < cL; s += 32) {
                const double x1 = A.x[c0[s]];
#pragma unroll
                for (int q = 0; q < G; ++q) acc[q] += v0[(int64_t)q * cL + s] * x1;
            }
        }
        // MASK: Dirichlet rows are NOT masked here any more: p = 0 on them, so p'Ap is unaffected, and
        // k_pcg_update_xr pins r = 0 where dinv == 0 (no per-group load of the flag in the hot loop).
        if (G == 3 && !DOT) {
            // three row sums with 6 instead of 15 double shuffles: halve the lanes AND the set of rows per step
            const bool up = lane & 16, up8 = lane & 8;
            const double t0 = __shfl_xor_sync(0xffffffffu, up ? acc[0] : acc[G - 1], 16);
            const double t1 = __shfl_xor_sync(0xffffffffu, acc[1 % G], 16);
            const double u0 = (up ? acc[G - 1] : acc[0]) + t0;  // lanes 0-15: row 0, lanes 16-31: row 2 (pair sums)
            const double u1 = up ? 0.0 : acc[1 % G] + t1;         // lanes 0-15: row 1
            const double t2 = __shfl_xor_sync(0xffffffffu, up8 ? u0 : u1, 8);
            double v = (up8 ? u1 : u0) + t2;                     // lanes 0-7 row 0, 8-15 row 1, 16-23 row 2, 24-31 nothing
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if ((lane & 7) == 0 && lane < 24) A.y[r0 + (lane >> 3)] = v;
        } else {
#pragma unroll
            for (int q = 0; q < G; ++q) acc[q] = warp_sum(acc[q]);  // xor-shuffles: every lane holds the row sums
            if (lane < G) {
                double yv = acc[0];
#pragma unroll
                for (int q = 1; q < G; ++q)
                    if (lane == q) yv = acc[q];
                A.y[r0 + lane] = yv;
            }
        }
        if (DOT) {
            // x_row (= p on the group's own rows) is among the gathered values: the diagonal columns
            const int dcol = (int)(A.ghost_cols + r0);
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (lane + 32 * j < cL) {
#pragma unroll
                    for (int q = 0; q < G; ++q)
                        if (cc[j] == dcol + q) dot += acc[q] * xv[j];
                }
            if (cL > 96) {
                const int32_t *__restrict__ c0 = A.colind + cb0;
                for (int s2 = lane + 96; s2 < cL; s2 += 32) {
                    const int c = c0[s2];
#pragma unroll
                    for (int q = 0; q < G; ++q)
                        if (c == dcol + q) dot += acc[q] * A.x[c];
                }
            }
        }
        // rotate the pipeline (a ping-pong of the two register sets instead of these copies was measured 40 % slower:
        // the compiler then serialises the streams of the next group behind the reduction)
        cg = ng_;
        cb0 = nb0;
        cL = nL;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            cc[j] = nc[j];
#pragma unroll
            for (int q = 0; q < G; ++q) cw[j][q] = nw[j][q];
        }
    }
    if (G > 1 && A.row_end == A.nrows && blockIdx.x == 0 && threadIdx.x < (unsigned)(A.nrows % G)) {  // leftover rows (never for G = 3 here)
        const int64_t r = (A.nrows / G) * G + threadIdx.x;
        double sacc = 0.0;
        for (int64_t p = A.rowptr[r]; p < A.rowptr[r + 1]; ++p) sacc += A.val[p] * A.x[A.colind[p]];
        if (MASK && A.fixed[r]) sacc = 0.0;
        A.y[r] = sacc;
        if (DOT) dot += sacc * A.x[A.ghost_cols + r];
    }
    if (DOT) {
        double s = block_sum<256>(dot, s_red);
        if (threadIdx.x == 0) {
            A.partials[blockIdx.x] = s;
            __threadfence();
            unsigned t = atomicAdd(&A.scal->ticketB, 1u);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double v = 0.0;
            for (int64_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += ld_volatile_f64(A.partials + i);
            double tot = block_sum<256>(v, s_red);
            if (threadIdx.x == 0) {
                A.scal->ticketB = 0;
                allreduce_publish(A.cv, 2ull * A.scal->it + 1ull, 1, &tot);
            }
        }
    }
}

template <int MODE, int G>
static void launch_spmv_group(smfem_ctx *ctx, smfem_matrix *K, const SpmvArgs &A) {
    static int per_sm = 0;
    if (per_sm == 0) {
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_spmv_group<MODE, G>, 256, 0));
        if (per_sm < 1) per_sm = 1;
    }
    LAUNCH(ctx, (k_spmv_group<MODE, G>), ctx->sms * per_sm, 256, 0, A);  // exactly the resident grid: no second wave
}

static void spmv_stream_setup(smfem_ctx *ctx, smfem_matrix *K) {
    if (K->blk_row || K->csr_less) return;  // a CSR-less operator has no rows to analyse
    int *d_max = dev_alloc<int>(1);
    CUDA_CHECK(cudaMemsetAsync(d_max, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, k_max_rowlen, (unsigned)((K->nrows_l + 255) / 256), 256, 0, K->nrows_l, (const int64_t *)K->rowptr, d_max);
    CUDA_CHECK(cudaMemcpyAsync(&K->max_rowlen, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    dev_free(d_max);
    K->nblk = (int)((K->nnz_l + SPMV_CH - 1) / SPMV_CH);
    if (K->nblk < 1) K->nblk = 1;
    K->blk_row = dev_alloc<int32_t>(K->nblk + 1);
    LAUNCH(ctx, k_spmv_blockrows, (unsigned)((K->nblk + 1 + 255) / 256), 256, 0, K->nrows_l, (const int64_t *)K->rowptr, K->nblk,
           K->blk_row);
    int *d_mr = dev_alloc<int>(1);
    int mr = 0;
    CUDA_CHECK(cudaMemsetAsync(d_mr, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, k_spmv_maxrows, (unsigned)((K->nblk + 255) / 256), 256, 0, K->nblk, (const int32_t *)K->blk_row, d_mr);
    CUDA_CHECK(cudaMemcpyAsync(&mr, d_mr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    dev_free(d_mr);
    K->stream_ok = (K->max_rowlen <= SPMV_SLACK && mr <= SPMV_MAXROWS);
    K->group3_ok = false;
    if (K->nDof == 3 && K->nrows_l % 3 == 0 && K->nrows_l > 0) {
        int *d_bad = dev_alloc<int>(1);
        int bad = 1;
        CUDA_CHECK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
        int64_t ng = K->nrows_l / 3;
        LAUNCH(ctx, k_check_group3, (unsigned)((ng * 32 + 255) / 256), 256, 0, ng, (const int64_t *)K->rowptr,
               (const int32_t *)K->colind, d_bad);
        CUDA_CHECK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        dev_free(d_bad);
        K->group3_ok = (bad == 0);
    }
}

static SpmvArgs make_spmv_args(smfem_matrix *K, const double *x, double *y) {
    SpmvArgs A;
    A.nrows = K->nrows_l;
    A.ghost_cols = K->ghost_cols;
    A.rowptr = K->rowptr;
    A.colind = K->colind;
    A.val = K->val;
    A.x = x;
    A.y = y;
    A.fixed = K->fixed;
    A.partials = K->partials;
    A.scal = K->scal;
    A.cv = K->comm;
    A.rot = 0;
    A.row_begin = 0;
    A.row_end = K->nrows_l;
    A.check_done = 0;
    A.halo_need = 0;
    static const bool lattice_cols = [] {
        const char *e = std::getenv("SMFEM_SPMV_LATTICE");
        return !(e && e[0] == '0');
    }();
    A.lat_n1 = (lattice_cols && K->structured && K->nDof == 3 && K->ndim == 3) ? K->lat.n1 : 0;
    return A;
}

template <int MODE>
static void launch_spmv(smfem_ctx *ctx, smfem_matrix *K, const SpmvArgs &A, int variant) {
    if (variant == 4 && K->group3_ok) {
        launch_spmv_group<MODE, 3>(ctx, K, A);
    } else if (variant == 4) {
        launch_spmv_group<MODE, 1>(ctx, K, A);  // same pipeline, one row per warp step
    } else if (variant == 3 && K->stream_ok) {
        launch_spmv_tma<MODE>(ctx, K, A);
    } else if ((variant == 2 || variant == 3 || variant == 4) && K->stream_ok) {
        int grid = ctx->sms * K->ctas_per_sm;
        if (grid > K->nblk) grid = K->nblk;
        LAUNCH(ctx, (k_spmv_stream<MODE>), grid, SPMV_NT, 0, A, (const int32_t *)K->blk_row, K->nblk);
    } else if (variant >= 1) {
        int64_t nw = A.nrows;
        unsigned grid = (unsigned)((nw * 32 + 255) / 256);
        LAUNCH(ctx, (k_spmv<1, MODE>), grid, 256, 0, A);
    } else {
        int64_t nw = (A.nrows + 2) / 3;
        unsigned grid = (unsigned)((nw * 32 + 255) / 256);
        LAUNCH(ctx, (k_spmv<3, MODE>), grid, 256, 0, A);
    }
}

// rotation that makes the rows of the boundary planes the LAST work of the grid (halo overlap)
static int64_t spmv_rotation(smfem_ctx *ctx, smfem_matrix *K, int variant) {
    if (ctx->nranks == 1 || K->nrows_l == 0) return 0;
    if (variant == 4 && K->group3_ok) return (K->comm.plane_dofs / 3) % (K->nrows_l / 3);
    if (variant == 4) return K->comm.plane_dofs % K->nrows_l;
    if ((variant == 2 || variant == 3 || variant == 4) && K->stream_ok)
        return (int64_t)((double)K->nblk * (double)K->comm.plane_dofs / (double)K->nrows_l) + 1;
    int64_t nw = (variant == 0) ? (K->nrows_l + 2) / 3 : K->nrows_l;
    int64_t per_plane = (variant == 0) ? (K->comm.plane_dofs + 2) / 3 : K->comm.plane_dofs;
    return per_plane % nw;
}

static int64_t spmv_grid(smfem_matrix *K, int variant) {
    int64_t nw = variant == 1 ? K->nrows_l : (K->nrows_l + 2) / 3;
    return (nw * 32 + 255) / 256;
}

// ------------------------------------------------------------------------------------------------
// K7: fused CG vector kernels (device-side scalars; no host round trip inside an iteration)
//   A: beta = rz_k / rz_{k-1};  p = D^-1 r + beta p;  boundary planes of p are also stored into
//      the neighbours' ghost planes (NVLink peer stores) and a flag is released     [halo push]
//   B: Ap = K p (masked), partial p'Ap -> all-reduce publish                         [k_spmv]
//   C: alpha = rz_k / p'Ap;  x += alpha p;  r -= alpha Ap;  partial r'D^-1 r, r'r -> publish
// ------------------------------------------------------------------------------------------------
constexpr int VEC_NT = 256;

__global__ void __launch_bounds__(VEC_NT)
k_pcg_update_p(int64_t n, int64_t ghost_cols, const double *__restrict__ r, const double *__restrict__ dinv,
               double *__restrict__ p, PcgScalars *scal, CommView cv, unsigned long long it_start_unused) {
    __shared__ double s_beta;
    __shared__ bool s_last, s_stop;
    if (scal->done) return;  // set by an earlier launch: the rest of the graph replay is a no-op
    const unsigned long long it = scal->it;
    if (threadIdx.x == 0) {
        const bool first = scal->spare != 0.0;  // spare != 0 marks the first iteration of a solve
        double v[3];
        allreduce_fetch(cv, 2ull * it, first ? 3 : 2, v, blockIdx.x == 0 ? &scal->t_wait_rz : nullptr);  // (rz_k, rr_k [, ||b||^2]) published by the previous C / init
        const double bn2 = first ? v[2] : scal->bnorm2;
        // the stopping test: every CTA of every rank evaluates it on bitwise identical numbers
        const bool stop = !(v[1] > scal->rtol2 * bn2) || scal->iters >= scal->maxit || scal->breakdown != 0;
        double rz_prev = scal->rzs[(it + 1ull) & 1ull];  // parity slots: [it-1]
        double beta = first ? 0.0 : v[0] / rz_prev;
        s_beta = beta;
        s_stop = stop;
        if (blockIdx.x == 0) {
            scal->rzs[it & 1ull] = v[0];
            scal->rr = v[1];
            if (first) scal->bnorm2 = v[2];
            scal->beta = beta;
            if (stop) {
                __threadfence();
                scal->done = 1;
            }
        }
    }
    __syncthreads();
    if (s_stop) return;
    const double beta = s_beta;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double *po = p + ghost_cols;
    const bool push_lo = cv.nranks > 1 && cv.rank > 0, push_hi = cv.nranks > 1 && cv.rank < cv.nranks - 1;
    double *dst_lo = push_lo ? cv.peer_p[cv.rank - 1] + cv.lo_dst_off : nullptr;  // their ghost_hi
    double *dst_hi = push_hi ? cv.peer_p[cv.rank + 1] : nullptr;                   // their ghost_lo (offset 0)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double z = r[i] * dinv[i];
        double pv = (beta == 0.0) ? z : z + beta * po[i];
        po[i] = pv;
        if (push_lo && i < cv.plane_dofs) dst_lo[i] = pv;
        if (push_hi && i >= n - cv.plane_dofs) dst_hi[i - (n - cv.plane_dofs)] = pv;
    }
    if (cv.nranks > 1) {
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
            if (push_lo) st_release_sys(&cv.peer[cv.rank - 1]->hflag[1], it + 1);
            if (push_hi) st_release_sys(&cv.peer[cv.rank + 1]->hflag[0], it + 1);
        }
    }
}

__global__ void __launch_bounds__(VEC_NT)
k_pcg_update_xr(int64_t n, int64_t ghost_cols, const double *__restrict__ p, const double *__restrict__ Ap,
                const double *__restrict__ dinv, double *__restrict__ x, double *__restrict__ r,
                double *__restrict__ partials, PcgScalars *scal, CommView cv) {
    __shared__ double s_alpha;
    __shared__ double s_red[VEC_NT / 32];
    __shared__ bool s_last;
    if (scal->done) return;
    const unsigned long long it = scal->it;
    if (threadIdx.x == 0) {
        double pAp;
        allreduce_fetch(cv, 2ull * it + 1ull, 1, &pAp, blockIdx.x == 0 ? &scal->t_wait_pap : nullptr);
        double rz = scal->rzs[it & 1ull];
        double alpha = 0.0;
        if (pAp > 0.0) alpha = rz / pAp;
        else if (rz != 0.0 && blockIdx.x == 0) scal->breakdown = 1;
        s_alpha = alpha;
        if (blockIdx.x == 0) {
            scal->alpha = alpha;
            scal->pAp = pAp;
        }
    }
    __syncthreads();
    const double alpha = s_alpha;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double *po = p + ghost_cols;
    double rz = 0.0, rr = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double di = dinv[i];
        double xv = x[i] + alpha * po[i];
        double rv = (di == 0.0) ? 0.0 : r[i] - alpha * Ap[i];  // Dirichlet rows (dinv == 0): residual pinned to 0
        x[i] = xv;
        r[i] = rv;
        rz += rv * rv * di;
        rr += rv * rv;
    }
    double s1 = block_sum<VEC_NT>(rz, s_red);
    double s2 = block_sum<VEC_NT>(rr, s_red);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = s1;
        partials[2 * blockIdx.x + 1] = s2;
        __threadfence();
        unsigned t = atomicAdd(&scal->ticketC, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double a = 0.0, b = 0.0;
        for (int64_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
            a += ld_volatile_f64(partials + 2 * i);
            b += ld_volatile_f64(partials + 2 * i + 1);
        }
        double t1 = block_sum<VEC_NT>(a, s_red);
        double t2 = block_sum<VEC_NT>(b, s_red);
        if (threadIdx.x == 0) {
            scal->ticketC = 0;
            scal->spare = 0.0;  // first-iteration marker cleared
            scal->iters += 1;
            double v[2] = {t1, t2};
            allreduce_publish(cv, 2ull * it + 2ull, 2, v);
            __threadfence();
            scal->it = it + 1;
        }
    }
}

// r0 = free ? (extra - K q_d) : 0 ; x0 = 0 ; D^-1 ; publishes (r'D^-1 r, r'r) and opens a new solve
__global__ void __launch_bounds__(VEC_NT)
k_pcg_init(int64_t n, const double *__restrict__ Kqd, const double *__restrict__ Kq0, const double *__restrict__ extra,
           const double *__restrict__ diag,
           const uint8_t *__restrict__ fixed, double *__restrict__ x, double *__restrict__ r, double *__restrict__ dinv,
           double *__restrict__ partials, PcgScalars *scal, CommView cv, double warm, double rtol2, unsigned maxit) {
    __shared__ double s_red[VEC_NT / 32];
    __shared__ bool s_last;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // Kqd = K (q_d + x0) gives the residual; Kq0 (warm start only) = K q_d gives the right-hand side b, whose norm
    // is the reference of the stopping test (with a good warm start ||r0|| is already at rounding level).
    double rz = 0.0, rr = 0.0, bb = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        bool fx = fixed[i] != 0;
        const double ex = extra ? extra[i] : 0.0;
        double rv = fx ? 0.0 : (ex - Kqd[i]);
        const double bv = fx ? 0.0 : (Kq0 ? ex - Kq0[i] : rv);
        bb += bv * bv;
        double di = (fx || diag[i] == 0.0) ? 0.0 : 1.0 / diag[i];
        x[i] = (warm != 0.0 && !fx) ? warm * x[i] : 0.0;  // warm start: x0 = warm * (previous solution on the free dofs)
        r[i] = rv;
        dinv[i] = di;
        rz += rv * rv * di;
        rr += rv * rv;
    }
    double s1 = block_sum<VEC_NT>(rz, s_red);
    double s2 = block_sum<VEC_NT>(rr, s_red);
    double s3 = block_sum<VEC_NT>(bb, s_red);
    if (threadIdx.x == 0) {
        partials[3 * blockIdx.x] = s1;
        partials[3 * blockIdx.x + 1] = s2;
        partials[3 * blockIdx.x + 2] = s3;
        __threadfence();
        unsigned t = atomicAdd(&scal->ticketC, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double a = 0.0, b = 0.0, c = 0.0;
        for (int64_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
            a += ld_volatile_f64(partials + 3 * i);
            b += ld_volatile_f64(partials + 3 * i + 1);
            c += ld_volatile_f64(partials + 3 * i + 2);
        }
        double t1 = block_sum<VEC_NT>(a, s_red);
        double t2 = block_sum<VEC_NT>(b, s_red);
        double t3 = block_sum<VEC_NT>(c, s_red);
        if (threadIdx.x == 0) {
            const unsigned long long it = scal->it;
            scal->ticketC = 0;
            scal->spare = 1.0;  // marks "first iteration": beta = 0
            scal->breakdown = 0;
            scal->rtol2 = rtol2;
            scal->maxit = maxit;
            scal->iters = 0;
            scal->done = 0;
            scal->t_wait_halo = scal->t_wait_rz = scal->t_wait_pap = 0ull;
            double v[3] = {t1, t2, t3};
            allreduce_publish(cv, 2ull * it + 2ull, 3, v);
            __threadfence();
            scal->it = it + 1;
        }
    }
}

// w = q_d + warm * x_prev on the OWNED entries only.  The ghost planes of w are the neighbours' to fill (halo push): a faster
// neighbour may already have pushed when this kernel runs, so they must not be touched here (with one rank they are never read).
__global__ void k_warm_vector(int64_t n, int64_t ghost_cols, int64_t ncols, const double *__restrict__ qd,
                              const double *__restrict__ x, double warm, double *__restrict__ w) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    (void)ncols;
    w[ghost_cols + r] = qd[ghost_cols + r] + warm * x[r];
}

// p'Ap as its own pass (50 MB of traffic, ~11 us at 100^3): fusing it into the SpMV cost 60 us there
__global__ void __launch_bounds__(VEC_NT)
k_pcg_dot(int64_t n, int64_t ghost_cols, const double *__restrict__ p, const double *__restrict__ Ap, double *__restrict__ partials,
          PcgScalars *scal, CommView cv) {
    __shared__ double s_red[VEC_NT / 32];
    __shared__ bool s_last;
    if (scal->done) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double *po = p + ghost_cols;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += po[i] * Ap[i];
    double s1 = block_sum<VEC_NT>(acc, s_red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s1;
        __threadfence();
        unsigned t = atomicAdd(&scal->ticketB, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double v = 0.0;
        for (int64_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += ld_volatile_f64(partials + i);
        double tot = block_sum<VEC_NT>(v, s_red);
        if (threadIdx.x == 0) {
            scal->ticketB = 0;
            allreduce_publish(cv, 2ull * scal->it + 1ull, 1, &tot);
        }
    }
}

// fetch the current global (rz, rr) into scal (so the host can read the residual)
__global__ void k_pcg_fetch(PcgScalars *scal, CommView cv, double *out3) {
    double v[3];
    allreduce_fetch(cv, 2ull * scal->it, 3, v);  // the third value is only meaningful right after k_pcg_init (||b||^2)
    out3[0] = v[0];
    out3[1] = v[1];
    out3[2] = v[2];
}

// true residual at exit:  sum over the free rows of (extra - K (q_d + x))^2  ->  published as sequence 2 it + 1 (the p'Ap
// slot of the iteration that never ran)
__global__ void __launch_bounds__(VEC_NT)
k_true_resid(int64_t n, const double *__restrict__ Kq, const double *__restrict__ extra, const uint8_t *__restrict__ fixed,
             double *__restrict__ partials, PcgScalars *scal, CommView cv) {
    __shared__ double s_red[VEC_NT / 32];
    __shared__ bool s_last;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double rv = fixed[i] ? 0.0 : (extra ? extra[i] : 0.0) - Kq[i];
        acc += rv * rv;
    }
    double s1 = block_sum<VEC_NT>(acc, s_red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s1;
        __threadfence();
        unsigned t = atomicAdd(&scal->ticketB, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double v = 0.0;
        for (int64_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += ld_volatile_f64(partials + i);
        double tot = block_sum<VEC_NT>(v, s_red);
        if (threadIdx.x == 0) {
            scal->ticketB = 0;
            allreduce_publish(cv, 2ull * scal->it + 1ull, 1, &tot);
        }
    }
}
// ... fetched by every rank; the iteration counter moves on so that the next solve / halo push uses fresh sequence numbers
__global__ void k_true_resid_fetch(PcgScalars *scal, CommView cv) {
    double v;
    allreduce_fetch(cv, 2ull * scal->it + 1ull, 1, &v);
    scal->rr_true = v;
    __threadfence();
    scal->it = scal->it + 1;
}

__global__ void k_final_q(int64_t n, int64_t ghost_cols, const double *__restrict__ qd, const double *__restrict__ x,
                          double *__restrict__ q) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) q[i] = qd[ghost_cols + i] + x[i];  // q = q_d + C q_f, examples/vector3D.jl:322
}

// halo push of an arbitrary vector with ghosts (bench_spmv)
__global__ void k_halo_push(int64_t n, int64_t ghost_cols, const double *__restrict__ v, PcgScalars *scal, CommView cv,
                            unsigned long long seq) {
    __shared__ bool s_last;
    const double *vo = v + ghost_cols;
    const bool push_lo = cv.rank > 0, push_hi = cv.rank < cv.nranks - 1;
    double *dst_lo = push_lo ? cv.peer_p[cv.rank - 1] + cv.lo_dst_off : nullptr;
    double *dst_hi = push_hi ? cv.peer_p[cv.rank + 1] : nullptr;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cv.plane_dofs; i += stride) {
        if (push_lo) dst_lo[i] = vo[i];
        if (push_hi) dst_hi[i] = vo[n - cv.plane_dofs + i];
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
        if (push_lo) st_release_sys(&cv.peer[cv.rank - 1]->hflag[1], seq);
        if (push_hi) st_release_sys(&cv.peer[cv.rank + 1]->hflag[0], seq);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int vec_grid(smfem_ctx *ctx, int64_t n) {
    int64_t g = (n + VEC_NT - 1) / VEC_NT;
    int64_t cap = (int64_t)ctx->sms * 8;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

void solver_alloc(smfem_ctx *ctx, smfem_matrix *K) {
    if (K->window) return;
    // multi-GPU hex lattice: room behind p for the vectors the multigrid hierarchy exchanges (sized from ne alone, so every
    // rank knows every peer's layout without communication; smfem_pcg_use_multigrid may be enabled at any time)
    K->gmg_region_doubles = (K->structured && K->ndim == 3 && K->nDof == 3 && ctx->nranks > 1 && !K->gmg_coarse)
                                ? gmg_window_doubles(K->lat.ne, ctx->rank, ctx->nranks) : 0;
    K->window_bytes = sizeof(CommHeader) + sizeof(double) * (size_t)(K->ncols_l + K->gmg_region_doubles);
    CUDA_CHECK(cudaMalloc(&K->window, K->window_bytes));
    CUDA_CHECK(cudaMemsetAsync(K->window, 0, K->window_bytes, ctx->stream));
    K->p = (double *)((char *)K->window + sizeof(CommHeader));
    K->x = dev_alloc<double>(K->nrows_l);
    K->r = dev_alloc<double>(K->nrows_l);
    K->Ap = dev_alloc<double>(K->nrows_l);
    K->dinv = dev_alloc<double>(K->nrows_l);
    K->qd = dev_alloc<double>(K->ncols_l);
    K->fixed = dev_alloc<uint8_t>(K->nrows_l);
    int64_t np = spmv_grid(K, 0);
    int64_t np1 = spmv_grid(K, 1);
    if (np1 > np) np = np1;
    if (np < 4 * (int64_t)ctx->sms * 8) np = 4 * (int64_t)ctx->sms * 8;  // k_pcg_init writes 3 partials per CTA
    K->partials = dev_alloc<double>(np + 16);
    K->partials_n = np;
    K->scal = dev_alloc<PcgScalars>(1);
    CUDA_CHECK(cudaMemsetAsync(K->scal, 0, sizeof(PcgScalars), ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(K->qd, 0, sizeof(double) * K->ncols_l, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(K->fixed, 0, K->nrows_l, ctx->stream));
    CUDA_CHECK(cudaMallocHost(&K->h_pinned, 256));
    static_assert(sizeof(PcgScalars) <= 256, "pinned read-back buffer");
    spmv_stream_setup(ctx, K);
    K->comm = CommView();
    K->comm.rank = ctx->rank;
    K->comm.nranks = ctx->nranks;
    K->comm.self = (CommHeader *)K->window;
    K->comm.plane_dofs = K->structured ? K->lat.plane() * K->nDof : 0;
    if (ctx->nranks == 1) {
        K->comm.peer[0] = K->comm.self;
        K->comm.peer_p[0] = K->p;
        K->comm_connected = true;
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

void solver_free(smfem_matrix *K) {
    if (K->pcg_graph) cudaGraphExecDestroy((cudaGraphExec_t)K->pcg_graph);
    K->pcg_graph = nullptr;
    for (int q = 0; q < SMFEM_MAX_RANKS; ++q)
        if (K->peer_maps[q]) {
            cudaIpcCloseMemHandle(K->peer_maps[q]);
            K->peer_maps[q] = nullptr;
        }
    if (K->window) cudaFree(K->window);
    K->window = nullptr;
    K->p = nullptr;
    dev_free(K->x);
    dev_free(K->r);
    dev_free(K->Ap);
    dev_free(K->dinv);
    dev_free(K->qd);
    dev_free(K->fixed);
    dev_free(K->partials);
    dev_free(K->blk_row);
    dev_free(K->scal);
    if (K->h_pinned) cudaFreeHost(K->h_pinned);
    K->h_pinned = nullptr;
}

void comm_export(smfem_ctx *ctx, smfem_matrix *K, void *handle_out) {
    solver_alloc(ctx, K);
    static_assert(sizeof(cudaIpcMemHandle_t) == SMFEM_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CUDA_CHECK(cudaIpcGetMemHandle(&h, K->window));
    std::memcpy(handle_out, &h, sizeof h);
}

void comm_connect(smfem_ctx *ctx, smfem_matrix *K, const void *handles) {
    solver_alloc(ctx, K);
    REQUIRE(K->structured, SMFEM_ERR_UNSUPPORTED, "multi-GPU needs a structured (slab-partitioned) mesh");
    REQUIRE(ctx->nranks <= SMFEM_MAX_RANKS, SMFEM_ERR_UNSUPPORTED, "at most 8 ranks");
    for (int q = 0; q < ctx->nranks; ++q) {
        void *base = K->window;
        if (q != ctx->rank) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, (const char *)handles + (size_t)q * SMFEM_IPC_HANDLE_BYTES, sizeof h);
            CUDA_CHECK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            K->peer_maps[q] = base;
        }
        K->comm.peer[q] = (CommHeader *)base;
        K->comm.peer_p[q] = (double *)((char *)base + sizeof(CommHeader));
    }
    if (ctx->rank > 0) {
        int k0, k1;
        slab_range(K->lat.n1, ctx->rank - 1, ctx->nranks, k0, k1);
        K->comm.lo_dst_off = (int64_t)(k1 - k0 + 1) * K->comm.plane_dofs;
    }
    K->comm_connected = true;
}

// The same connection inside ONE process (smfem_init_multi): the peers' windows are ordinary device pointers of this process;
// the caller has enabled peer access between the devices.  all_K[q] = rank q's matrix (its window already allocated).
void comm_connect_local(smfem_ctx *ctx, smfem_matrix *K, smfem_matrix *const *all_K, int n) {
    solver_alloc(ctx, K);
    REQUIRE(K->structured, SMFEM_ERR_UNSUPPORTED, "multi-GPU needs a structured (slab-partitioned) mesh");
    REQUIRE(n == ctx->nranks && n <= SMFEM_MAX_RANKS, SMFEM_ERR_INVALID, "comm_connect_local: one matrix per rank");
    for (int q = 0; q < n; ++q) {
        REQUIRE(all_K[q] && all_K[q]->window, SMFEM_ERR_INVALID, "comm_connect_local: a peer has no window yet (smfem_comm_prepare)");
        REQUIRE(all_K[q]->window_bytes >= sizeof(CommHeader), SMFEM_ERR_INVALID, "comm_connect_local: bad peer window");
        void *base = all_K[q]->window;
        K->comm.peer[q] = (CommHeader *)base;
        K->comm.peer_p[q] = (double *)((char *)base + sizeof(CommHeader));
    }
    if (ctx->rank > 0) {
        int k0, k1;
        slab_range(K->lat.n1, ctx->rank - 1, ctx->nranks, k0, k1);
        K->comm.lo_dst_off = (int64_t)(k1 - k0 + 1) * K->comm.plane_dofs;
    }
    K->comm_connected = true;
}

// y = K x without mask / fused dot.  halo: the rows of the first / last owned plane wait for the ghost planes pushed by the
// neighbours.  For the row-group kernel the SpMV is split into an interior launch (no flag code at all: the halo variant
// costs 46 us even on one GPU) followed by a boundary-plane launch that waits -- stream order gives the overlap of the
// NVLink transfer with the interior rows for free.  Other variants keep the single rotated launch.
static void spmv_apply(smfem_ctx *ctx, smfem_matrix *K, const double *x, double *y, bool halo, bool check_done = false,
                       unsigned long long halo_need = 0) {
    if (K->matfree_on) {  // the operator applied from the lattice coordinates (matfree.cu), same vector layout and halo protocol
        matfree_apply(ctx, K, x, y, halo, check_done, halo_need);
        return;
    }
    SpmvArgs A = make_spmv_args(K, x, y);
    A.check_done = check_done ? 1 : 0;
    A.halo_need = halo_need;
    const int variant = K->spmv_variant;
    if (!halo || ctx->nranks == 1) {
        launch_spmv<0>(ctx, K, A, variant);
        return;
    }
    if (variant != 4) {
        A.rot = spmv_rotation(ctx, K, variant);
        launch_spmv<4>(ctx, K, A, variant);
        return;
    }
    const int64_t pd = K->comm.plane_dofs, n = K->nrows_l;
    const int64_t lo = (ctx->rank > 0) ? pd : 0;                     // first owned plane needs ghost_lo
    const int64_t hi = (ctx->rank < ctx->nranks - 1) ? n - pd : n;   // last owned plane needs ghost_hi
    if (hi > lo) {
        A.row_begin = lo;
        A.row_end = hi;
        launch_spmv<0>(ctx, K, A, variant);
    }
    if (lo > 0) {
        A.row_begin = 0;
        A.row_end = lo < hi ? lo : (lo < n ? lo : n);
        launch_spmv<4>(ctx, K, A, variant);
    }
    if (hi < n) {
        A.row_begin = hi > lo ? hi : lo;  // a single owned plane is both first and last
        A.row_end = n;
        if (A.row_end > A.row_begin) launch_spmv<4>(ctx, K, A, variant);
    }
}

void spmv_device(smfem_ctx *ctx, smfem_matrix *K, const double *x, double *y) {
    solver_alloc(ctx, K);  // SpMV setup (row-triple check, stream blocks)
    spmv_apply(ctx, K, x, y, false);
}

// device-side barrier of the ranks (a dummy all-reduce with sequence 2 it + 1), then a fresh sequence number: no rank may
// push its next halo before every rank has finished reading the current one
__global__ void k_rank_barrier(PcgScalars *scal, CommView cv) {
    const double one = 1.0;
    double v;
    allreduce_publish(cv, 2ull * scal->it + 1ull, 1, &one);
    allreduce_fetch(cv, 2ull * scal->it + 1ull, 1, &v);
    __threadfence();
    scal->it = scal->it + 1;
}

// y = K x for host vectors of this rank's row slab.  Several ranks: collective; the ghost planes of x come from the
// neighbours through the peer window (same push / flag / wait as a CG iteration), so smfem_comm_connect must have run.
void spmv_host(smfem_ctx *ctx, smfem_matrix *K, const double *x, double *y) {
    REQUIRE(K->values_ready, SMFEM_ERR_INVALID, "matrix has no values yet");
    solver_alloc(ctx, K);
    REQUIRE(K->comm_connected, SMFEM_ERR_INVALID, "multi-GPU: call smfem_comm_connect first");
    if (ctx->nranks == 1) CUDA_CHECK(cudaMemsetAsync(K->p, 0, sizeof(double) * K->ncols_l, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(K->p + K->ghost_cols, x, sizeof(double) * K->nrows_l, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->nranks > 1) {
        PcgScalars *h = reinterpret_cast<PcgScalars *>(K->h_pinned);
        CUDA_CHECK(cudaMemcpyAsync(h, K->scal, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        int g = (int)((K->comm.plane_dofs + 255) / 256);
        if (g > ctx->sms * 4) g = ctx->sms * 4;
        LAUNCH(ctx, k_halo_push, g, 256, 0, K->nrows_l, K->ghost_cols, (const double *)K->p, K->scal, K->comm, h->it + 1);
        spmv_apply(ctx, K, K->p, K->Ap, true);
        LAUNCH(ctx, k_rank_barrier, 1, 1, 0, K->scal, K->comm);
    } else {
        spmv_apply(ctx, K, K->p, K->Ap, false);
    }
    CUDA_CHECK(cudaMemcpyAsync(y, K->Ap, sizeof(double) * K->nrows_l, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

__global__ void k_fill_pattern(int64_t n, double *__restrict__ v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = 1.0 + 1e-3 * (double)(i % 1024);
}

void bench_spmv(smfem_ctx *ctx, smfem_matrix *K, int variant, int reps, float *ms) {
    REQUIRE(K->values_ready, SMFEM_ERR_INVALID, "matrix has no values yet");
    REQUIRE(reps > 0, SMFEM_ERR_INVALID, "reps must be positive");
    solver_alloc(ctx, K);
    REQUIRE(K->comm_connected, SMFEM_ERR_INVALID, "multi-GPU: call smfem_comm_connect first");
    LAUNCH(ctx, k_fill_pattern, (unsigned)((K->ncols_l + 255) / 256), 256, 0, K->ncols_l, K->p);
    SpmvArgs A = make_spmv_args(K, K->p, K->Ap);
    A.rot = spmv_rotation(ctx, K, variant);
    // sequence numbers for the halo flags continue the solver's counter
    unsigned long long it0 = 0;
    CUDA_CHECK(cudaMemcpyAsync(&it0, &K->scal->it, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    int dbg_mode = 0;  // experiments: SMFEM_BENCH_MODE = 1 mask, 2 dot, 3 mask|dot, 7 solve mode (single GPU only)
    if (const char *e = std::getenv("SMFEM_BENCH_MODE")) dbg_mode = std::atoi(e);
    const int saved_variant = K->spmv_variant;
    const bool saved_mf = K->matfree_on;
    if (variant == 5) {  // the matrix-free operator (needs smfem_pcg_use_matrix_free to have named the mesh)
        REQUIRE(K->mf_mesh != nullptr, SMFEM_ERR_INVALID, "bench_spmv variant 5: call smfem_pcg_use_matrix_free first");
        K->matfree_on = true;
    } else {
        K->matfree_on = false;
        K->spmv_variant = variant;
    }
    int bench_seq = 0;
    auto one = [&](int j) {
        if (ctx->nranks == 1 && dbg_mode == 2) launch_spmv<2>(ctx, K, A, variant);
        else if (ctx->nranks == 1 && dbg_mode == 3) launch_spmv<3>(ctx, K, A, variant);
        else if (ctx->nranks == 1 && dbg_mode == 7) launch_spmv<7>(ctx, K, A, variant);
        else if (ctx->nranks > 1) {
            const unsigned long long seq = it0 + 1 + (unsigned long long)bench_seq++;  // every repetition really waits for its neighbours' push
            int g = (int)((K->comm.plane_dofs + 255) / 256);
            if (g > ctx->sms * 4) g = ctx->sms * 4;
            LAUNCH(ctx, k_halo_push, g, 256, 0, K->nrows_l, K->ghost_cols, (const double *)K->p, K->scal, K->comm, seq);
            spmv_apply(ctx, K, K->p, K->Ap, true, false, seq);
        } else {
            spmv_apply(ctx, K, K->p, K->Ap, false);
        }
        (void)j;
    };
    for (int j = 0; j < 3; ++j) one(j);
    CUDA_CHECK(cudaEventRecord(ctx->ev2, ctx->stream));
    for (int j = 0; j < reps; ++j) one(j);
    CUDA_CHECK(cudaEventRecord(ctx->ev3, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(ctx->ev3));
    float t = 0;
    CUDA_CHECK(cudaEventElapsedTime(&t, ctx->ev2, ctx->ev3));
    *ms = t / reps;
    K->spmv_variant = saved_variant;
    K->matfree_on = saved_mf;
    if (ctx->nranks > 1) {  // later pushes (solver: it + 1) must use larger sequence numbers than the ones used here
        unsigned long long it_new = it0 + (unsigned long long)bench_seq + 1;
        CUDA_CHECK(cudaMemcpyAsync(&K->scal->it, &it_new, 8, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
}

constexpr int PCG_CHUNK = 25;  // iterations per CUDA-graph replay (after convergence the rest of a replay is no-op launches)

// capture PCG_CHUNK iterations once per (matrix, SpMV variant); every later solve replays the instantiated graph.  One iteration:
//   update_p (+ convergence test, halo push) | SpMV (interior, then boundary planes after the halo flags) | p'Ap | update x, r
static cudaGraphExec_t pcg_graph(smfem_ctx *ctx, smfem_matrix *K, int variant) {
    if (K->pcg_graph && K->pcg_graph_variant == variant) return (cudaGraphExec_t)K->pcg_graph;
    if (K->pcg_graph) cudaGraphExecDestroy((cudaGraphExec_t)K->pcg_graph);
    K->pcg_graph = nullptr;
    const int64_t n = K->nrows_l;
    const int vg = vec_grid(ctx, n);
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    const int64_t l0 = ctx->launches;
    CUDA_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    for (int j = 0; j < PCG_CHUNK; ++j) {
        LAUNCH(ctx, k_pcg_update_p, vg, VEC_NT, 0, n, K->ghost_cols, (const double *)K->r, (const double *)K->dinv, K->p,
               K->scal, K->comm, 0ull);
        spmv_apply(ctx, K, K->p, K->Ap, true, /*check_done=*/true);
        LAUNCH(ctx, k_pcg_dot, vg, VEC_NT, 0, n, K->ghost_cols, (const double *)K->p, (const double *)K->Ap, K->partials,
               K->scal, K->comm);
        LAUNCH(ctx, k_pcg_update_xr, vg, VEC_NT, 0, n, K->ghost_cols, (const double *)K->p, (const double *)K->Ap,
               (const double *)K->dinv, K->x, K->r, K->partials, K->scal, K->comm);
    }
    K->pcg_graph_launches = ctx->launches - l0;
    CUDA_CHECK(cudaStreamEndCapture(ctx->stream, &graph));
    ctx->launches = l0;  // captured, not launched; counted per replay
    CUDA_CHECK(cudaGraphInstantiate(&gexec, graph, 0));
    cudaGraphDestroy(graph);
    K->pcg_graph = gexec;
    K->pcg_graph_variant = variant;
    return gexec;
}

void pcg_solve(smfem_ctx *ctx, smfem_matrix *K, double rtol, int maxit, const double *rhs_extra, double *q_out,
               int *iters, double *relres) {
    REQUIRE(K->values_ready, SMFEM_ERR_INVALID, "matrix has no values yet");
    solver_alloc(ctx, K);
    REQUIRE(K->comm_connected, SMFEM_ERR_INVALID, "multi-GPU: call smfem_comm_connect first");
    const int64_t n = K->nrows_l;
    const int variant = K->matfree_on ? 5 : K->spmv_variant;  // also the key of the cached iteration graph
    double *extra = nullptr;
    if (rhs_extra) {
        extra = dev_alloc<double>(n);
        CUDA_CHECK(cudaMemcpyAsync(extra, rhs_extra, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    const int vg = vec_grid(ctx, n);
    cudaGraphExec_t gexec = pcg_graph(ctx, K, variant);
    PcgScalars *h = reinterpret_cast<PcgScalars *>(K->h_pinned);
    auto read_scal = [&]() {
        CUDA_CHECK(cudaMemcpyAsync(h, K->scal, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    };
    auto halo_push = [&](const double *v, unsigned long long seq) {
        int g = (int)((K->comm.plane_dofs + 255) / 256);
        if (g > ctx->sms * 4) g = ctx->sms * 4;
        LAUNCH(ctx, k_halo_push, g, 256, 0, n, K->ghost_cols, v, K->scal, K->comm, seq);
    };
    CUDA_CHECK(cudaEventRecord(ctx->ev2, ctx->stream));
    const double warm = K->warm_scale;
    K->warm_scale = 0.0;  // applies to one solve
    if (warm == 0.0) {
        // K q_d (unmasked rows; q_d's ghost entries are filled locally, no exchange needed)
        spmv_apply(ctx, K, K->qd, K->Ap, false);
    } else {
        // warm start (load stepping, examples/vector3D.jl:310-338: q is linear in d):  r0 = extra - K (q_d + warm x_prev)
        if (K->sol_x && K->sol_x != K->x)  // the previous solve was the multigrid one: its solution lives in its own buffer
            CUDA_CHECK(cudaMemcpyAsync(K->x, K->sol_x, 8 * n, cudaMemcpyDeviceToDevice, ctx->stream));
        {  // b = extra - K q_d is still needed for the stopping test: K q_d -> r (r is rewritten by k_pcg_init)
            spmv_apply(ctx, K, K->qd, K->r, false);
        }
        LAUNCH(ctx, k_warm_vector, (unsigned)((K->ncols_l + 255) / 256), 256, 0, n, K->ghost_cols, K->ncols_l, (const double *)K->qd,
               (const double *)K->x, warm, K->p);
        if (ctx->nranks > 1) {
            read_scal();
            halo_push(K->p, h->it + 1);
        }
        spmv_apply(ctx, K, K->p, K->Ap, ctx->nranks > 1);
    }
    LAUNCH(ctx, k_pcg_init, vg, VEC_NT, 0, n, (const double *)K->Ap, (const double *)(warm != 0.0 ? K->r : nullptr),
           (const double *)extra, (const double *)K->diag, (const uint8_t *)K->fixed, K->x, K->r, K->dinv, K->partials, K->scal,
           K->comm, warm, rtol * rtol, (unsigned)maxit);
    // replay until the device-side test (k_pcg_update_p) has set `done`; the host only looks at the flag
    for (int64_t launched = 0;; launched += PCG_CHUNK) {
        CUDA_CHECK(cudaGraphLaunch(gexec, ctx->stream));
        ctx->launches += K->pcg_graph_launches;
        read_scal();
        if (h->done) break;
        REQUIRE(launched <= (int64_t)maxit + PCG_CHUNK, SMFEM_ERR_CUDA, "PCG: the device-side stopping test never fired (internal error)");
    }
    const int it = (int)h->iters;
    const double bnorm2 = h->bnorm2, res2 = h->rr;
    const unsigned brk = h->breakdown;
    // true residual  || (extra - K (q_d + x))_free ||  (the recursive one drifts over ~1e3 iterations); K->p, K->Ap are free now
    double res2_true = res2;
    {
        LAUNCH(ctx, k_warm_vector, (unsigned)((K->ncols_l + 255) / 256), 256, 0, n, K->ghost_cols, K->ncols_l, (const double *)K->qd,
               (const double *)K->x, 1.0, K->p);
        if (ctx->nranks > 1) halo_push(K->p, h->it + 1);
        spmv_apply(ctx, K, K->p, K->Ap, ctx->nranks > 1);
        LAUNCH(ctx, k_true_resid, vg, VEC_NT, 0, n, (const double *)K->Ap, (const double *)extra, (const uint8_t *)K->fixed, K->partials,
               K->scal, K->comm);
        LAUNCH(ctx, k_true_resid_fetch, 1, 1, 0, K->scal, K->comm);
        if (q_out) {  // q = q_d + C q_f on the owned rows is what the SpMV just read (examples/vector3D.jl:322)
            CUDA_CHECK(cudaMemcpyAsync(q_out, K->p + K->ghost_cols, 8 * n, cudaMemcpyDeviceToHost, ctx->stream));
        }
        read_scal();
        res2_true = h->rr_true;
    }
    CUDA_CHECK(cudaEventRecord(ctx->ev3, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(ctx->ev3));
    CUDA_CHECK(cudaEventElapsedTime(&K->last_ms, ctx->ev2, ctx->ev3));
    K->last_iters = it;
    K->last_wait_us[0] = 1e-3 * (double)h->t_wait_halo;
    K->last_wait_us[1] = 1e-3 * (double)h->t_wait_rz;
    K->last_wait_us[2] = 1e-3 * (double)h->t_wait_pap;
    K->sol_x = K->x;
    K->last_relres_rec = bnorm2 > 0 ? std::sqrt(res2 / bnorm2) : 0.0;
    K->last_relres_true = bnorm2 > 0 ? std::sqrt(res2_true / bnorm2) : 0.0;
    if (extra) dev_free(extra);
    if (iters) *iters = it;
    if (relres) *relres = K->last_relres_true;
    REQUIRE(brk == 0, SMFEM_ERR_SINGULAR, "PCG breakdown: p'Ap <= 0 (matrix not SPD on the free dofs; reference: SingularException)");
}
