// Assembly side of the path: sparsity pattern on device, element kernels, scatter, surface term,
// diagonal extraction, CSC export.
//   reference: src/fem.jl:135-256 (assemble_system), examples/vector3D.jl:175-264 (surface matrix)
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "smfem_internal.cuh"

// ------------------------------------------------------------------------------------------------
// Quadrature tables in the reference's Gauss-point order (src/fem.jl:161-164, :172-176), built on
// the host with the same code that backs the ABI's basis_function and uploaded once.
// ------------------------------------------------------------------------------------------------
__constant__ double c_dN3[8][8][3];  // [gp][node][d/dxi_d]
__constant__ double c_w3[8];
__constant__ double c_dN2[4][4][2];
__constant__ double c_N2[4][4];
__constant__ double c_w2[4];
__constant__ double c_dNq2[4][9][2];  // Q2 9-node quad gradients at the 2x2 Gauss points (src/fem.jl:90-110, :161-164)

void mesh_upload_tables() {
    double xi[2], w[2];
    smfem_host_gauss(-1, 1, 2, xi, w);
    const int ix[8] = {0, 1, 1, 0, 0, 1, 1, 0}, iy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, iz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    double dN3[8][8][3], w3[8], dN2[4][4][2], N2[4][4], w2[4], dNq2[4][9][2];
    for (int g = 0; g < 8; ++g) {
        double N[8], dN[24];
        int nn;
        smfem_host_basis(3, SMFEM_Q1, xi[ix[g]], xi[iy[g]], xi[iz[g]], N, dN, &nn);
        for (int a = 0; a < 8; ++a)
            for (int d = 0; d < 3; ++d) dN3[g][a][d] = dN[d * 8 + a];
        w3[g] = w[ix[g]] * w[iy[g]] * w[iz[g]];
    }
    for (int g = 0; g < 4; ++g) {
        double N[4], dN[8];
        int nn;
        smfem_host_basis(2, SMFEM_Q1, xi[ix[g]], xi[iy[g]], 0, N, dN, &nn);
        for (int a = 0; a < 4; ++a) {
            N2[g][a] = N[a];
            for (int d = 0; d < 2; ++d) dN2[g][a][d] = dN[d * 4 + a];
        }
        w2[g] = w[ix[g]] * w[iy[g]];
        double Nq[9], dNq[18];
        smfem_host_basis(2, SMFEM_Q2, xi[ix[g]], xi[iy[g]], 0, Nq, dNq, &nn);
        for (int a = 0; a < 9; ++a)
            for (int d = 0; d < 2; ++d) dNq2[g][a][d] = dNq[d * 9 + a];
    }
    CUDA_CHECK(cudaMemcpyToSymbol(c_dN3, dN3, sizeof dN3));
    CUDA_CHECK(cudaMemcpyToSymbol(c_w3, w3, sizeof w3));
    CUDA_CHECK(cudaMemcpyToSymbol(c_dN2, dN2, sizeof dN2));
    CUDA_CHECK(cudaMemcpyToSymbol(c_N2, N2, sizeof N2));
    CUDA_CHECK(cudaMemcpyToSymbol(c_w2, w2, sizeof w2));
    CUDA_CHECK(cudaMemcpyToSymbol(c_dNq2, dNq2, sizeof dNq2));
}

// ------------------------------------------------------------------------------------------------
// dof maps and connectivity views shared by the kernels
// ------------------------------------------------------------------------------------------------
struct DofMap {
    const int32_t *id;  // [comp][node] 0-based, or nullptr for dof = nDof*node + comp
    int64_t nNodes;
    int nDof;
    int64_t ghost_cols, nrows_l;
    __device__ __forceinline__ int64_t col(int64_t ln, int c) const {
        return id ? (int64_t)id[(int64_t)c * nNodes + ln] : ln * nDof + c;
    }
    __device__ __forceinline__ int64_t row(int64_t ln, int c) const {
        int64_t r = col(ln, c) - ghost_cols;
        return (r >= 0 && r < nrows_l) ? r : -1;
    }
};

struct Conn {
    const int32_t *ien;  // general: [a][e]
    int64_t nEl;         // elements this rank works on
    Lattice L;
    int structured;
    int layer0;  // structured: first element layer worked on
    __device__ __forceinline__ int64_t node(int64_t e, int a) const {
        if (!structured) return ien[(int64_t)a * nEl + e];
        int ne = L.ne;
        int ei = (int)(e % ne), ej = (int)((e / ne) % ne), ek = (int)(e / ((int64_t)ne * ne)) + layer0;
        // local node order of examples/vector3D.jl:94-101
        int ox = ((a & 3) == 1 || (a & 3) == 2), oy = ((a & 3) >= 2), oz = (a >> 2);
        return L.lnode(ei + ox, ej + oy, ek + oz);
    }
};

static DofMap make_dofmap(const smfem_mesh *mesh, const smfem_matrix *K) {
    DofMap d;
    d.id = mesh->structured ? nullptr : mesh->id;
    d.nNodes = mesh->nNodes_l;
    d.nDof = K->nDof;
    d.ghost_cols = K->ghost_cols;
    d.nrows_l = K->nrows_l;
    return d;
}

static Conn make_conn(const smfem_mesh *mesh) {
    Conn c;
    c.ien = mesh->ien;
    c.L = mesh->lat;
    c.structured = mesh->structured ? 1 : 0;
    c.layer0 = 0;
    c.nEl = mesh->nEl_g;
    if (mesh->structured) {
        int l0 = mesh->lat.k0 - 1 < 0 ? 0 : mesh->lat.k0 - 1;
        int l1 = mesh->lat.k1 < mesh->lat.ne ? mesh->lat.k1 : mesh->lat.ne;  // layers [l0,l1)
        c.layer0 = l0;
        c.nEl = (int64_t)(l1 - l0) * mesh->lat.ne * mesh->lat.ne;
    }
    return c;
}

// position of local column `c` inside local row `r` (binary search; rows are sorted ascending)
__device__ __forceinline__ int64_t csr_find(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                            int64_t r, int64_t c) {
    int64_t lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) >> 1;
        int32_t v = colind[mid];
        if (v == c) return mid;
        if (v < c) lo = mid + 1;
        else hi = mid - 1;
    }
    return -1;
}

// ------------------------------------------------------------------------------------------------
// K2 (structured): closed-form CSR pattern == Julia's sparse(E,J,V) pattern for the hex lattice.
// Along one axis node i has cnt1(i) in-range neighbours {i-1,i,i+1}; the neighbours of (i,j,k) in
// ascending global node order are the tensor product (k' slowest).  Each row of node m holds
// nDof * |nbrs(m)| columns: the dofs nDof*n + c' of its neighbours n, ascending.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cnt1(int i, int n1) { return 1 + (i > 0) + (i < n1 - 1); }
__device__ __forceinline__ int64_t pre1(int i) { return i == 0 ? 0 : 3 * (int64_t)i - 1; }
__device__ __forceinline__ int64_t pairs_before(int n1, int i, int j, int k) {
    int64_t S1 = 3 * (int64_t)n1 - 2;
    return pre1(k) * S1 * S1 + (int64_t)cnt1(k, n1) * (pre1(j) * S1 + (int64_t)cnt1(j, n1) * pre1(i));
}

__global__ void k_struct_rowptr(Lattice L, int nDof, int64_t nrows_l, int64_t *__restrict__ rowptr) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nrows_l) return;
    int64_t base = (int64_t)nDof * nDof * pairs_before(L.n1, 0, 0, L.k0);
    if (r == nrows_l) {
        int64_t S1 = 3 * (int64_t)L.n1 - 2;
        int64_t endp = (L.k1 >= L.n1) ? S1 * S1 * S1 : pairs_before(L.n1, 0, 0, L.k1);
        rowptr[r] = (int64_t)nDof * nDof * endp - base;
        return;
    }
    int64_t node = r / nDof;
    int c = (int)(r % nDof);
    int i = (int)(node % L.n1), j = (int)((node / L.n1) % L.n1), k = (int)(node / L.plane()) + L.k0;
    int cnt = cnt1(i, L.n1) * cnt1(j, L.n1) * cnt1(k, L.n1);
    rowptr[r] = (int64_t)nDof * nDof * pairs_before(L.n1, i, j, k) + (int64_t)c * nDof * cnt - base;
}

// one warp per owned node: lanes 0..26 compute the first column of one neighbour node each (the only
// div/mod work), park it in shared memory, then the warp streams the node's nDof rows (contiguous in
// colind) with coalesced stores.
template <int NDOF>
__global__ void __launch_bounds__(256) k_struct_colind(Lattice L, int64_t nOwnedNodes, const int64_t *__restrict__ rowptr,
                                                       int32_t *__restrict__ colind) {
    __shared__ int32_t s_nb[8][28];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t node = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (node >= nOwnedNodes) return;
    int i = (int)(node % L.n1), j = (int)((node / L.n1) % L.n1), k = (int)(node / L.plane()) + L.k0;
    int cx = cnt1(i, L.n1), cy = cnt1(j, L.n1), cz = cnt1(k, L.n1);
    int i0 = i - (i > 0), j0 = j - (j > 0), k0 = k - (k > 0);
    const int cnt = cx * cy * cz;
    if (lane < cnt) {
        int ai = lane % cx, aj = (lane / cx) % cy, ak = lane / (cx * cy);
        s_nb[w][lane] = (int32_t)(L.lnode(i0 + ai, j0 + aj, k0 + ak) * NDOF);
    }
    __syncwarp();
    const int T = NDOF * cnt;
    int32_t *dst = colind + rowptr[node * NDOF];
    for (int t = lane; t < NDOF * T; t += 32) {
        int s = t - ((t >= T) ? T : 0) - ((t >= 2 * T) ? T : 0);  // position inside its row (NDOF <= 3)
        int q = s / NDOF;
        dst[t] = s_nb[w][q] + (s - q * NDOF);
    }
}

// The same column indices in closed form WITHOUT reading rowptr, as a small persistent kernel (64-thread CTAs, < 64 registers)
// that fits on every SM beside the two resident CTAs of the value kernel (which leaves 4096 registers and ~30 KB of shared
// memory free): launched on the side stream, it fills the pattern while the value kernel computes (SMFEM_COLIND_SIDE=1).
// A warp walks x-lines of nodes; for interior nodes the 243 entries of the node's 3 rows are 3 node + rel(t), rel fixed per lane.
__global__ void __launch_bounds__(64) k_struct_colind_side(Lattice L, int32_t *__restrict__ colind) {
    const int lane = threadIdx.x & 31;
    const int wid = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nw = (int)((gridDim.x * blockDim.x) >> 5);
    const int n1 = L.n1;
    int rel[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int t = lane + 32 * m, s = t % 81, q = s / 3, jj = s - 3 * q;
        rel[m] = 3 * (((q / 9 - 1) * n1 + ((q / 3) % 3 - 1)) * n1 + (q % 3 - 1)) + jj;
    }
    const int64_t base = 9 * pairs_before(n1, 0, 0, L.k0);
    const int nlines = L.nown() * n1;
    for (int line = wid; line < nlines; line += nw) {
        const int k = L.k0 + line / n1, j = line - (line / n1) * n1;
        const int cy = cnt1(j, n1), cz = cnt1(k, n1);
        const bool inner_line = cy == 3 && cz == 3;
        int64_t start = 9 * pairs_before(n1, 0, j, k) - base;  // entries before node (0, j, k)
        for (int i = 0; i < n1; ++i) {
            const int cx = cnt1(i, n1);
            const int T = 3 * cx * cy * cz;
            int32_t *dst = colind + start;
            if (inner_line && cx == 3) {
                const int32_t c0 = (int32_t)(L.lnode(i, j, k) * 3);
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    if (m < 7 || lane < 243 - 224) dst[lane + 32 * m] = c0 + rel[m];
            } else {
                const int i0 = i - (i > 0), j0 = j - (j > 0), kk0 = k - (k > 0);
                for (int t = lane; t < 3 * T; t += 32) {
                    const int s = t % T, q = s / 3, jj = s - 3 * q;
                    const int ai = q % cx, aj = (q / cx) % cy, ak = q / (cx * cy);
                    dst[t] = (int32_t)(L.lnode(i0 + ai, j0 + aj, kk0 + ak) * 3) + jj;
                }
            }
            start += 3 * T;
        }
    }
}

static void matrix_alloc_pattern(smfem_matrix *K) {
    K->rowptr = dev_alloc<int64_t>(K->nrows_l + 1 + 8);  // +8: slack for 16 B-aligned bulk copies (TMA SpMV)
}

// sizes of the structured pattern in closed form + buffers; no kernel, no host synchronisation
void pattern_prepare_structured(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K) {
    const Lattice &L = mesh->lat;
    int nDof = K->nDof;
    K->structured = true;
    K->lat = L;
    K->m_g = (int64_t)nDof * mesh->nNodes_g;
    int64_t S1 = 3 * (int64_t)L.n1 - 2;
    K->nnz_g = (int64_t)nDof * nDof * S1 * S1 * S1;
    K->ghost_cols = L.plane() * nDof;
    K->nrows_l = (int64_t)L.nown() * L.plane() * nDof;
    K->ncols_l = K->nrows_l + 2 * K->ghost_cols;
    K->row0 = (int64_t)L.k0 * L.plane() * nDof;
    REQUIRE(K->ncols_l < (int64_t)INT32_MAX, SMFEM_ERR_UNSUPPORTED, "local dof count exceeds int32 column indices");
    if (K->rowptr != nullptr) return;
    // (row node, column node) pairs of the owned planes: 1-D prefix pre1(i) = #pairs of rows < i (see k_struct_rowptr)
    auto pre1 = [&](int i) -> int64_t { return i == 0 ? 0 : (i == L.n1 ? 3 * (int64_t)L.n1 - 2 : 3 * (int64_t)i - 1); };
    K->nnz_l = (int64_t)nDof * nDof * (pre1(L.k1) - pre1(L.k0)) * S1 * S1;
    matrix_alloc_pattern(K);
    K->colind = dev_alloc<int32_t>(K->nnz_l + 16);
    CUDA_CHECK(cudaMemsetAsync(K->colind + K->nnz_l, 0, 16 * sizeof(int32_t), ctx->stream));
}

void pattern_build_structured(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K) {
    pattern_prepare_structured(ctx, mesh, K);
    const Lattice &L = mesh->lat;
    int nDof = K->nDof;
    LAUNCH(ctx, k_struct_rowptr, (unsigned)((K->nrows_l + 1 + 255) / 256), 256, 0, L, nDof, K->nrows_l, K->rowptr);
    int64_t nOwned = (int64_t)L.nown() * L.plane();
    if (nDof == 3)
        LAUNCH(ctx, (k_struct_colind<3>), (unsigned)((nOwned * 32 + 255) / 256), 256, 0, L, nOwned, (const int64_t *)K->rowptr, K->colind);
    else
        LAUNCH(ctx, (k_struct_colind<1>), (unsigned)((nOwned * 32 + 255) / 256), 256, 0, L, nOwned, (const int64_t *)K->rowptr, K->colind);
}

// ------------------------------------------------------------------------------------------------
// K2 (general): arbitrary (IEN, ID).  node->element lists -> sorted unique node adjacency ->
// rows through the dof map, columns sorted ascending (Julia CSC has ascending row ids per column
// and the pattern is structurally symmetric).
// ------------------------------------------------------------------------------------------------
constexpr int MAX_ADJ = 64;   // max distinct neighbour nodes of a node (incl. itself)
constexpr int MAX_VAL = 16;   // max elements sharing a node

__global__ void k_n2e_count(const int32_t *__restrict__ ien, int64_t nEl, int nn, int *__restrict__ cnt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nEl * nn) return;
    atomicAdd(&cnt[ien[t]], 1);
}

__global__ void k_n2e_fill(const int32_t *__restrict__ ien, int64_t nEl, int nn, const int64_t *__restrict__ ptr,
                           int *__restrict__ cursor, int32_t *__restrict__ n2e) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nEl * nn) return;
    int node = ien[t];
    int e = (int)(t % nEl);
    int slot = atomicAdd(&cursor[node], 1);
    n2e[ptr[node] + slot] = e;
}

// thread per node: sorted unique neighbour list (two passes: count, then fill)
__global__ void k_node_adj(const int32_t *__restrict__ ien, int64_t nEl, int nn, int64_t nNodes,
                           const int64_t *__restrict__ n2e_ptr, const int32_t *__restrict__ n2e,
                           const int64_t *__restrict__ adj_ptr, int32_t *__restrict__ adj, int *__restrict__ adj_cnt,
                           int *__restrict__ err) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nNodes) return;
    int32_t loc[MAX_ADJ];
    int n = 0;
    for (int64_t q = n2e_ptr[m]; q < n2e_ptr[m + 1]; ++q) {
        int e = n2e[q];
        for (int a = 0; a < nn; ++a) {
            int32_t v = ien[(int64_t)a * nEl + e];
            int pos = 0;
            while (pos < n && loc[pos] < v) ++pos;
            if (pos < n && loc[pos] == v) continue;
            if (n >= MAX_ADJ) {
                *err = 1;
                return;
            }
            for (int s = n; s > pos; --s) loc[s] = loc[s - 1];
            loc[pos] = v;
            ++n;
        }
    }
    if (adj == nullptr) {
        adj_cnt[m] = n;
    } else {
        int64_t base = adj_ptr[m];
        for (int s = 0; s < n; ++s) adj[base + s] = loc[s];
    }
}

// row lengths: row of dof (m,c) has nDof*|adj(m)| entries
__global__ void k_gen_rowlen(DofMap D, const int64_t *__restrict__ adj_ptr, int64_t nNodes, int64_t ndof,
                             int64_t *__restrict__ rowlen, int *__restrict__ err) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nNodes * D.nDof) return;
    int64_t m = t / D.nDof;
    int c = (int)(t % D.nDof);
    int64_t r = D.col(m, c);
    if (r < 0 || r >= ndof) {
        *err = 2;
        return;
    }
    int64_t old = atomicExch((unsigned long long *)&rowlen[r], (unsigned long long)((adj_ptr[m + 1] - adj_ptr[m]) * D.nDof));
    if (old != 0) *err = 3;  // two (node, comp) pairs map to one dof: not a bijection
}

__global__ void k_gen_colind(DofMap D, const int64_t *__restrict__ adj_ptr, const int32_t *__restrict__ adj,
                             int64_t nNodes, const int64_t *__restrict__ rowptr, int32_t *__restrict__ colind) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nNodes * D.nDof) return;
    int64_t m = t / D.nDof;
    int c = (int)(t % D.nDof);
    int64_t r = D.col(m, c);
    int64_t base = rowptr[r];
    int n = 0;
    bool sorted = true;
    int32_t prev = -1;
    for (int64_t q = adj_ptr[m]; q < adj_ptr[m + 1]; ++q) {
        int32_t nb = adj[q];
        for (int cc = 0; cc < D.nDof; ++cc) {
            int32_t v = (int32_t)D.col(nb, cc);
            colind[base + n++] = v;
            sorted = sorted && (v > prev);
            prev = v;
        }
    }
    if (!sorted) {  // permuted ID maps: insertion sort of this row (rare, short rows)
        for (int a = 1; a < n; ++a) {
            int32_t v = colind[base + a];
            int b = a - 1;
            while (b >= 0 && colind[base + b] > v) {
                colind[base + b + 1] = colind[base + b];
                --b;
            }
            colind[base + b + 1] = v;
        }
    }
}

struct IntTo64 {
    __host__ __device__ int64_t operator()(int v) const { return (int64_t)v; }
};

// device exclusive prefix sum (CUB) of `n` values; `in` may be a transform iterator
template <class In>
static void exclusive_scan_dev(smfem_ctx *ctx, In in, int64_t *out, int64_t n) {
    void *tmp = nullptr;
    size_t bytes = 0;
    REQUIRE(n < (int64_t)INT32_MAX, SMFEM_ERR_UNSUPPORTED, "scan length exceeds int32");
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n, ctx->stream));
    tmp = dev_alloc<unsigned char>(bytes ? bytes : 1);
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n, ctx->stream));
    ctx->launches += 2;
    unsigned char *t8 = static_cast<unsigned char *>(tmp);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    dev_free(t8);
}

__global__ void k_max_int(int64_t n, const int *__restrict__ v, int *__restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) atomicMax(out, v[t]);
}

void pattern_build_general(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K) {
    REQUIRE(ctx->nranks == 1, SMFEM_ERR_UNSUPPORTED, "general (unstructured) meshes are single-GPU only");
    const int64_t nNodes = mesh->nNodes_g, nEl = mesh->nEl_g;
    const int nn = mesh->nn, nDof = K->nDof;
    K->structured = false;
    K->ghost_cols = 0;
    int64_t ndof = mesh->id ? mesh->ndof_id : nNodes * nDof;  // id == nullptr: standard map (or nDof == 1 raw node ids)
    REQUIRE(ndof == nNodes * nDof, SMFEM_ERR_UNSUPPORTED,
            "ID must be a bijection onto 1..nDof*nNodes (max(ID) != nDof*nNodes)");
    K->m_g = ndof;
    K->nrows_l = ndof;
    K->ncols_l = ndof;
    K->row0 = 0;
    REQUIRE(ndof < (int64_t)INT32_MAX, SMFEM_ERR_UNSUPPORTED, "dof count exceeds int32 column indices");

    // everything below runs on the device; the host only reads back three scalars (max valence, totals) and flags
    int *cnt = dev_alloc<int>(nNodes + 1), *cursor = dev_alloc<int>(nNodes + 1), *flags = dev_alloc<int>(4);
    int64_t *n2e_ptr = dev_alloc<int64_t>(nNodes + 1), *adj_ptr = dev_alloc<int64_t>(nNodes + 1);
    int32_t *n2e = nullptr, *adj = nullptr;
    int64_t *rowlen = nullptr;
    auto cleanup = [&] {
        dev_free(cnt);
        dev_free(cursor);
        dev_free(flags);
        dev_free(n2e_ptr);
        dev_free(adj_ptr);
        dev_free(n2e);
        dev_free(adj);
        dev_free(rowlen);
    };
    try {
        CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int) * (nNodes + 1), ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(cursor, 0, sizeof(int) * (nNodes + 1), ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(int) * 4, ctx->stream));
        int *err = flags, *d_maxval = flags + 1;
        const unsigned gE = (unsigned)((nEl * nn + 255) / 256), gN = (unsigned)((nNodes + 127) / 128);
        // node -> element lists
        LAUNCH(ctx, k_n2e_count, gE, 256, 0, (const int32_t *)mesh->ien, nEl, nn, cnt);
        LAUNCH(ctx, k_max_int, (unsigned)((nNodes + 255) / 256), 256, 0, nNodes, (const int *)cnt, d_maxval);
        exclusive_scan_dev(ctx, cub::TransformInputIterator<int64_t, IntTo64, const int *>(cnt, IntTo64()), n2e_ptr, nNodes + 1);
        int h[2] = {0, 0};
        CUDA_CHECK(cudaMemcpy(h, flags, sizeof(int) * 2, cudaMemcpyDeviceToHost));
        REQUIRE(h[1] <= MAX_VAL, SMFEM_ERR_UNSUPPORTED, "a node is shared by more than 16 elements");
        n2e = dev_alloc<int32_t>(nEl * nn);
        LAUNCH(ctx, k_n2e_fill, gE, 256, 0, (const int32_t *)mesh->ien, nEl, nn, (const int64_t *)n2e_ptr, cursor, n2e);
        // (element order inside a node's list is irrelevant: the adjacency below is sorted + unique)
        // node -> node adjacency: count pass, scan, fill pass
        int *adj_cnt = cnt;
        LAUNCH(ctx, k_node_adj, gN, 128, 0, (const int32_t *)mesh->ien, nEl, nn, nNodes, (const int64_t *)n2e_ptr,
               (const int32_t *)n2e, (const int64_t *)nullptr, (int32_t *)nullptr, adj_cnt, err);
        CUDA_CHECK(cudaMemsetAsync(adj_cnt + nNodes, 0, sizeof(int), ctx->stream));
        exclusive_scan_dev(ctx, cub::TransformInputIterator<int64_t, IntTo64, const int *>(adj_cnt, IntTo64()), adj_ptr, nNodes + 1);
        int64_t nadj = 0;
        CUDA_CHECK(cudaMemcpy(&nadj, adj_ptr + nNodes, 8, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(h, flags, sizeof(int), cudaMemcpyDeviceToHost));
        REQUIRE(h[0] == 0, SMFEM_ERR_UNSUPPORTED, "a node has more than 64 neighbour nodes");
        adj = dev_alloc<int32_t>(nadj);
        LAUNCH(ctx, k_node_adj, gN, 128, 0, (const int32_t *)mesh->ien, nEl, nn, nNodes, (const int64_t *)n2e_ptr,
               (const int32_t *)n2e, (const int64_t *)adj_ptr, adj, adj_cnt, err);
        // rows through the dof map
        DofMap D = make_dofmap(mesh, K);
        D.nrows_l = ndof;
        rowlen = dev_alloc<int64_t>(ndof + 1);
        CUDA_CHECK(cudaMemsetAsync(rowlen, 0, 8 * (ndof + 1), ctx->stream));
        const unsigned gR = (unsigned)((nNodes * nDof + 255) / 256);
        LAUNCH(ctx, k_gen_rowlen, gR, 256, 0, D, (const int64_t *)adj_ptr, nNodes, ndof, rowlen, err);
        matrix_alloc_pattern(K);
        exclusive_scan_dev(ctx, (const int64_t *)rowlen, K->rowptr, ndof + 1);
        CUDA_CHECK(cudaMemcpy(h, flags, sizeof(int), cudaMemcpyDeviceToHost));
        REQUIRE(h[0] == 0, SMFEM_ERR_INVALID, "ID is not a bijection onto 1..nDof*nNodes");
        CUDA_CHECK(cudaMemcpy(&K->nnz_l, K->rowptr + ndof, 8, cudaMemcpyDeviceToHost));
        K->nnz_g = K->nnz_l;
        K->colind = dev_alloc<int32_t>(K->nnz_l + 16);
        CUDA_CHECK(cudaMemsetAsync(K->colind + K->nnz_l, 0, 16 * sizeof(int32_t), ctx->stream));
        LAUNCH(ctx, k_gen_colind, gR, 256, 0, D, (const int64_t *)adj_ptr, (const int32_t *)adj, nNodes, (const int64_t *)K->rowptr,
               K->colind);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

// ------------------------------------------------------------------------------------------------
// Element colouring of a general mesh (north star: "graph-coloured ... scatter-add").
// Jones-Plassmann with fixed pseudo-random priorities: in every sweep an uncoloured element whose priority
// beats all its uncoloured neighbours (elements sharing a node) takes the smallest colour none of its
// coloured neighbours has.  The sweeps read the previous sweep's colours only (two buffers), so the result
// does not depend on thread timing; elements are then sorted by colour (stable radix sort).  The value
// kernel runs colour after colour (one launch each): a fixed fold order per matrix entry = bit-reproducible.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

__global__ void k_color_sweep(const int32_t *__restrict__ ien, int64_t nEl, int nn, const int64_t *__restrict__ n2e_ptr,
                              const int32_t *__restrict__ n2e, const int8_t *__restrict__ col_old, int8_t *__restrict__ col_new,
                              int *__restrict__ flags) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nEl) return;
    const int8_t c0 = col_old[e];
    if (c0 >= 0) {
        col_new[e] = c0;
        return;
    }
    const uint32_t pe = hash32((uint32_t)e);
    unsigned long long forbidden = 0;
    bool is_max = true;
    for (int a = 0; a < nn; ++a) {
        const int32_t node = ien[(int64_t)a * nEl + e];
        for (int64_t q = n2e_ptr[node]; q < n2e_ptr[node + 1]; ++q) {
            const int32_t f = n2e[q];
            if (f == e) continue;
            const int8_t cf = col_old[f];
            if (cf >= 0) {
                forbidden |= 1ull << cf;
            } else {
                const uint32_t pf = hash32((uint32_t)f);
                if (pf > pe || (pf == pe && f > e)) is_max = false;
            }
        }
    }
    if (!is_max) {
        col_new[e] = -1;
        flags[0] = 1;  // somebody is still uncoloured
        return;
    }
    const int c = __ffsll((long long)~forbidden) - 1;
    if (c < 0 || c >= 64) {
        flags[1] = 1;  // more than 64 colours needed
        col_new[e] = 0;
        return;
    }
    col_new[e] = (int8_t)c;
}

__global__ void k_color_hist(const int8_t *__restrict__ col, int64_t nEl, int *__restrict__ hist) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nEl) atomicAdd(&hist[col[e]], 1);
}

__global__ void k_iota32(int64_t n, int32_t *__restrict__ v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int32_t)i;
}

void mesh_color_elements(smfem_ctx *ctx, smfem_mesh *mesh) {
    if (mesh->ncolors != 0) return;
    REQUIRE(!mesh->structured && mesh->ien, SMFEM_ERR_INVALID, "element colouring is for general (IEN) meshes");
    const int64_t nNodes = mesh->nNodes_g, nEl = mesh->nEl_g;
    const int nn = mesh->nn;
    int *cnt = dev_alloc<int>(nNodes + 1), *cursor = dev_alloc<int>(nNodes + 1), *flags = dev_alloc<int>(2 + 64);
    int64_t *n2e_ptr = dev_alloc<int64_t>(nNodes + 1);
    int32_t *n2e = dev_alloc<int32_t>(nEl * nn), *ids = dev_alloc<int32_t>(nEl), *sorted = dev_alloc<int32_t>(nEl);
    int8_t *col[2] = {dev_alloc<int8_t>(nEl), dev_alloc<int8_t>(nEl)};
    uint8_t *keys_out = dev_alloc<uint8_t>(nEl);
    void *tmp = nullptr;
    auto cleanup = [&] {
        dev_free(cnt);
        dev_free(cursor);
        dev_free(flags);
        dev_free(n2e_ptr);
        dev_free(n2e);
        dev_free(ids);
        dev_free(col[0]);
        dev_free(col[1]);
        dev_free(keys_out);
        if (tmp) dev_cache_free(tmp);
    };
    try {
        CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int) * (nNodes + 1), ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(cursor, 0, sizeof(int) * (nNodes + 1), ctx->stream));
        const unsigned gE = (unsigned)((nEl * nn + 255) / 256), gEl = (unsigned)((nEl + 127) / 128);
        LAUNCH(ctx, k_n2e_count, gE, 256, 0, (const int32_t *)mesh->ien, nEl, nn, cnt);
        exclusive_scan_dev(ctx, cub::TransformInputIterator<int64_t, IntTo64, const int *>(cnt, IntTo64()), n2e_ptr, nNodes + 1);
        LAUNCH(ctx, k_n2e_fill, gE, 256, 0, (const int32_t *)mesh->ien, nEl, nn, (const int64_t *)n2e_ptr, cursor, n2e);
        CUDA_CHECK(cudaMemsetAsync(col[0], 0xFF, nEl, ctx->stream));
        int cur = 0, h[2] = {1, 0}, sweeps = 0;
        while (h[0]) {
            REQUIRE(++sweeps <= 1000, SMFEM_ERR_CUDA, "element colouring did not converge");
            CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(int) * 2, ctx->stream));
            LAUNCH(ctx, k_color_sweep, gEl, 128, 0, (const int32_t *)mesh->ien, nEl, nn, (const int64_t *)n2e_ptr, (const int32_t *)n2e,
                   (const int8_t *)col[cur], col[cur ^ 1], flags);
            CUDA_CHECK(cudaMemcpyAsync(h, flags, sizeof(int) * 2, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            cur ^= 1;
            if (h[1]) break;
        }
        if (h[1]) {  // a node of very high valence: keep the atomic scatter
            mesh->ncolors = -1;
            dev_free(sorted);
            cleanup();
            return;
        }
        // elements sorted by colour + colour offsets
        LAUNCH(ctx, k_iota32, (unsigned)((nEl + 255) / 256), 256, 0, nEl, ids);
        size_t tmp_bytes = 0;
        const uint8_t *keys_in = reinterpret_cast<const uint8_t *>(col[cur]);
        CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, (const int32_t *)ids, sorted, nEl, 0, 7, ctx->stream));
        tmp = dev_cache_alloc(tmp_bytes);
        CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, (const int32_t *)ids, sorted, nEl, 0, 7, ctx->stream));
        ctx->launches++;
        int *hist = flags + 2;
        CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(int) * 64, ctx->stream));
        LAUNCH(ctx, k_color_hist, gEl, 128, 0, (const int8_t *)col[cur], nEl, hist);
        int hh[64];
        CUDA_CHECK(cudaMemcpyAsync(hh, hist, sizeof(int) * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        int nc = 0;
        mesh->color_off[0] = 0;
        for (int c = 0; c < 64; ++c) {
            mesh->color_off[c + 1] = mesh->color_off[c] + hh[c];
            if (hh[c]) nc = c + 1;
        }
        REQUIRE(mesh->color_off[64] == nEl, SMFEM_ERR_CUDA, "element colouring lost elements");
        mesh->elist = sorted;
        mesh->ncolors = nc;
    } catch (...) {
        dev_free(sorted);
        cleanup();
        throw;
    }
    cleanup();
}

// ------------------------------------------------------------------------------------------------
// K1 + K3, general form: one thread per (element, local node a) computes the a-th block row of
//   Ke = sum_gp w * B'DB                                   (src/fem.jl:183-233)
// through the isotropic identity  K_ab = lam * G_ab + mu * G_ab' + mu tr(G_ab) I,
// G_ab = sum_gp w grad N_a grad N_b'  (material applied once after the Gauss loop), and adds it
// into the CSR values with fp64 atomics.  Used for unstructured meshes, 2-D and scalar problems;
// the structured hex path uses the tiled gather kernel in assemble_tile.cu.
// ------------------------------------------------------------------------------------------------
template <int NDIM>
__device__ __forceinline__ double jac_inv(const double *J, double *inv) {
    if (NDIM == 2) {
        double d = J[0] * J[3] - J[1] * J[2];
        double id = 1.0 / d;
        inv[0] = J[3] * id;
        inv[1] = -J[1] * id;
        inv[2] = -J[2] * id;
        inv[3] = J[0] * id;
        return d;
    } else {
        double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
        double d = J[0] * c00 + J[1] * c01 + J[2] * c02;
        double id = 1.0 / d;
        inv[0] = c00 * id;
        inv[1] = (J[2] * J[7] - J[1] * J[8]) * id;
        inv[2] = (J[1] * J[5] - J[2] * J[4]) * id;
        inv[3] = c01 * id;
        inv[4] = (J[0] * J[8] - J[2] * J[6]) * id;
        inv[5] = (J[2] * J[3] - J[0] * J[5]) * id;
        inv[6] = c02 * id;
        inv[7] = (J[1] * J[6] - J[0] * J[7]) * id;
        inv[8] = (J[0] * J[4] - J[1] * J[3]) * id;
        return d;
    }
}

struct Material {
    double d11, lam, mu;  // D(1,1), D(1,2), shear
};

// COLORED: the launch covers the elements elist[0 .. n_list) of ONE colour (no two share a node): every entry of K receives
// at most one add per launch, so the fold order is the colour order whatever the thread timing (the adds stay L2 reductions:
// a load-add-store round trip measured 2x slower than RED)
template <int NDIM, int NDOF, int NN = (1 << NDIM), bool COLORED = false>  // NN = 9: the reference's Q2 quad (2-D, scalar, 2x2 under-integration)
__global__ void __launch_bounds__(128)
k_values_atomic(Conn C, DofMap D, const double *__restrict__ coords, const int64_t *__restrict__ rowptr,
                const int32_t *__restrict__ colind, double *__restrict__ val, Material mat, const int32_t *__restrict__ elist = nullptr,
                int64_t n_list = 0) {
    constexpr int NGP = 1 << NDIM;
    __shared__ double s_dN[NGP][NN][NDIM];
    __shared__ double s_w[NGP];
    for (int t = threadIdx.x; t < NGP * NN * NDIM; t += blockDim.x) {
        int g = t / (NN * NDIM), rem = t % (NN * NDIM);
        s_dN[g][rem / NDIM][rem % NDIM] = (NDIM == 3) ? c_dN3[g][(rem / NDIM) % 8][rem % NDIM]
                                          : (NN == 9 ? c_dNq2[g][rem / NDIM][rem % NDIM] : c_dN2[g][(rem / NDIM) % 4][rem % NDIM]);
    }
    if (threadIdx.x < NGP) s_w[threadIdx.x] = (NDIM == 3) ? c_w3[threadIdx.x] : c_w2[threadIdx.x];
    __syncthreads();
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (COLORED ? n_list : C.nEl) * NN) return;
    int64_t e = COLORED ? (int64_t)elist[t / NN] : t / NN;
    int a = (int)(t % NN);
    int64_t nodes[NN];
    double X[NN][NDIM];
#pragma unroll
    for (int b = 0; b < NN; ++b) {
        nodes[b] = C.node(e, b);
#pragma unroll
        for (int d = 0; d < NDIM; ++d) X[b][d] = coords[nodes[b] * NDIM + d];
    }
    // does this thread own any row?  (ghost element layers only feed owned rows)
    int64_t rows[NDOF];
    bool any = false;
#pragma unroll
    for (int c = 0; c < NDOF; ++c) {
        rows[c] = D.row(nodes[a], c);
        any = any || rows[c] >= 0;
    }
    if (!any) return;
    constexpr int GS = (NDOF == 1) ? 1 : NDIM * NDIM;
    double G[NN][GS];
#pragma unroll
    for (int b = 0; b < NN; ++b)
#pragma unroll
        for (int s = 0; s < GS; ++s) G[b][s] = 0.0;
#pragma unroll 1
    for (int g = 0; g < NGP; ++g) {
        double J[NDIM * NDIM], inv[NDIM * NDIM];
#pragma unroll
        for (int r = 0; r < NDIM; ++r)
#pragma unroll
            for (int c = 0; c < NDIM; ++c) {
                double s = 0;
#pragma unroll
                for (int b = 0; b < NN; ++b) s += X[b][r] * s_dN[g][b][c];  // Jac = coords*dN, src/fem.jl:192
                J[r * NDIM + c] = s;
            }
        double w = s_w[g] * fabs(jac_inv<NDIM>(J, inv));  // :194-195
        double ga[NDIM];
#pragma unroll
        for (int c = 0; c < NDIM; ++c) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < NDIM; ++k) s += s_dN[g][a][k] * inv[k * NDIM + c];  // dNdX = dN*invJ, :196
            ga[c] = s * w;
        }
#pragma unroll
        for (int b = 0; b < NN; ++b) {
            double gb[NDIM];
#pragma unroll
            for (int c = 0; c < NDIM; ++c) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < NDIM; ++k) s += s_dN[g][b][k] * inv[k * NDIM + c];
                gb[c] = s;
            }
            if (NDOF == 1) {
                double s = 0;
#pragma unroll
                for (int c = 0; c < NDIM; ++c) s += ga[c] * gb[c];
                G[b][0] += s;
            } else {
#pragma unroll
                for (int i = 0; i < NDIM; ++i)
#pragma unroll
                    for (int j = 0; j < NDIM; ++j) G[b][(i * NDIM + j) % GS] += ga[i] * gb[j];
            }
        }
    }
    // material + scatter (src/fem.jl:236-249)
#pragma unroll 1
    for (int b = 0; b < NN; ++b) {
        if (NDOF == 1) {
            int64_t pos = csr_find(rowptr, colind, rows[0], D.col(nodes[b], 0));
            atomicAdd(&val[pos], G[b][0]);  // COLORED: the only add to this entry in the launch (RED is cheaper than load-add-store)
        } else {
            double tr = 0;
#pragma unroll
            for (int i = 0; i < NDIM; ++i) tr += G[b][(i * NDIM + i) % GS];
            // standard dof map: the NDOF columns of node b are consecutive and the NDOF rows of node a share one
            // pattern -> ONE search per node pair instead of NDOF^2
            int64_t rel0 = -1;
            if (D.id == nullptr && rows[0] >= 0) rel0 = csr_find(rowptr, colind, rows[0], D.col(nodes[b], 0)) - rowptr[rows[0]];
#pragma unroll
            for (int i = 0; i < NDIM; ++i) {
                if (rows[i % NDOF] < 0) continue;
#pragma unroll
                for (int j = 0; j < NDIM; ++j) {
                    double gij = G[b][(i * NDIM + j) % GS], gji = G[b][(j * NDIM + i) % GS];
                    double v = (i == j) ? mat.d11 * gij + mat.mu * (tr - gij) : mat.lam * gij + mat.mu * gji;
                    int64_t pos = (rel0 >= 0) ? rowptr[rows[i % NDOF]] + rel0 + j
                                              : csr_find(rowptr, colind, rows[i % NDOF], D.col(nodes[b], j % NDOF));
                    atomicAdd(&val[pos], v);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Gather form of the value assembly for GENERAL 3-D hex meshes with the standard dof map: every entry of K is
// written exactly once, by the warp that owns its row - no memset, no atomics, no colour passes (the scatter forms move each
// entry of K through L2 / DRAM once per contributing element: 8x for interior nodes, and K >> L2).
//   plan (once per mesh): the (element, local node) pairs of every node, sorted by element -> fold order = ascending element
//       index, the order in which sparse(E,J,V) folds duplicates (src/fem.jl:253); nodes are cut into warp tasks of <= 16 pairs.
//   element pass: A = sqrt(w |det J|) J^-1 per (element, Gauss point) -> 576 B per element of scratch (k_gather_elem);
//   kernel: lane = one (node a, element e) pair: the 8 blocks G_ab = sum_gp (dN_a A)(dN_b A)' of "its" element in registers (72
//       accumulators, no coordinates, no Jacobian: the next Gauss point's A is loaded while this one's 153 FMAs run), parked in a
//       per-warp staging area; then, node by node, the warp loads the node's
//       neighbour list (its row's columns) into shared memory, every pair looks up where its 8 nodes sit in that row, the 9
//       entries of every neighbour block are summed over the pairs in element order by all 32 lanes, and the warp streams the
//       node's 3 CSR rows out with the material applied - coalesced, contiguous (the rows of a node are adjacent).
// ------------------------------------------------------------------------------------------------

constexpr int GATHER_PAIRS = 16;  // (node, element) pairs per warp task: TWO lanes per pair (4 of the 8 blocks each)

__global__ void k_max_rowlen_nodes(int64_t nNodes, const int64_t *__restrict__ rowptr, int *__restrict__ out) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < nNodes) atomicMax(out, (int)((rowptr[3 * n + 1] - rowptr[3 * n]) / 3));
}

__global__ void k_gather_fill(const int32_t *__restrict__ ien, int64_t nEl, int nn, const int64_t *__restrict__ ptr,
                              int *__restrict__ cursor, int32_t *__restrict__ ent) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nEl * nn) return;
    const int node = ien[t];
    const int e = (int)(t % nEl), a = (int)(t / nEl);
    const int slot = atomicAdd(&cursor[node], 1);
    ent[ptr[node] + slot] = e * nn + a;
}

__global__ void k_gather_sort(int64_t nNodes, const int64_t *__restrict__ ptr, int32_t *__restrict__ ent, int *__restrict__ maxlen) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    const int64_t b = ptr[n];
    const int len = (int)(ptr[n + 1] - b);
    for (int i = 1; i < len; ++i) {  // insertion sort: the lists are short (8 for an interior hex node)
        const int32_t v = ent[b + i];
        int j = i - 1;
        while (j >= 0 && ent[b + j] > v) {
            ent[b + j + 1] = ent[b + j];
            --j;
        }
        ent[b + j + 1] = v;
    }
    atomicMax(maxlen, len);
}

void mesh_build_gather(smfem_ctx *ctx, smfem_mesh *mesh) {
    if (mesh->g_state != 0) return;
    const int64_t nNodes = mesh->nNodes_g, nEl = mesh->nEl_g;
    const int nn = mesh->nn;
    mesh->g_state = -1;
    if (nEl * nn >= (int64_t)INT32_MAX) return;
    int *cnt = dev_alloc<int>(nNodes + 1), *cursor = dev_alloc<int>(nNodes + 1), *d_max = dev_alloc<int>(1);
    int64_t *ptr = dev_alloc<int64_t>(nNodes + 1);
    int32_t *ent = dev_alloc<int32_t>(nEl * nn);
    auto fail = [&] {
        dev_free(cnt);
        dev_free(cursor);
        dev_free(d_max);
        dev_free(ptr);
        dev_free(ent);
    };
    try {
        CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int) * (nNodes + 1), ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(cursor, 0, sizeof(int) * (nNodes + 1), ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(d_max, 0, sizeof(int), ctx->stream));
        const unsigned gE = (unsigned)((nEl * nn + 255) / 256);
        LAUNCH(ctx, k_n2e_count, gE, 256, 0, (const int32_t *)mesh->ien, nEl, nn, cnt);
        exclusive_scan_dev(ctx, cub::TransformInputIterator<int64_t, IntTo64, const int *>(cnt, IntTo64()), ptr, nNodes + 1);
        LAUNCH(ctx, k_gather_fill, gE, 256, 0, (const int32_t *)mesh->ien, nEl, nn, (const int64_t *)ptr, cursor, ent);
        LAUNCH(ctx, k_gather_sort, (unsigned)((nNodes + 127) / 128), 128, 0, nNodes, (const int64_t *)ptr, ent, d_max);
        std::vector<int64_t> h_ptr((size_t)nNodes + 1);
        int h_max = 0;
        CUDA_CHECK(cudaMemcpyAsync(h_ptr.data(), ptr, sizeof(int64_t) * (nNodes + 1), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(&h_max, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        if (h_max > GATHER_PAIRS) {  // a node shared by more elements than a warp task holds: the scatter forms handle it
            fail();
            return;
        }
        std::vector<int32_t> task;  // greedy: whole nodes, <= GATHER_PAIRS pairs per warp task
        task.push_back(0);
        int64_t start = 0;
        for (int64_t n = 0; n < nNodes; ++n)
            if (h_ptr[n + 1] - start > GATHER_PAIRS) {
                task.push_back((int32_t)n);
                start = h_ptr[n];
            }
        task.push_back((int32_t)nNodes);
        mesh->g_ntasks = (int)task.size() - 1;
        mesh->g_task_node = dev_alloc<int32_t>(task.size());
        CUDA_CHECK(cudaMemcpyAsync(mesh->g_task_node, task.data(), sizeof(int32_t) * task.size(), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        fail();
        throw;
    }
    dev_free(cnt);
    dev_free(cursor);
    dev_free(d_max);
    mesh->g_ptr = ptr;
    mesh->g_ent = ent;
    mesh->g_state = 1;
}

constexpr int GATHER_STAGE = 73;  // doubles per lane in the staging area (72 + 1: odd stride, conflict-free)
__host__ __device__ inline size_t gather_warp_bytes(int max_slots) {
    // stage[16][73] f64 | red[max_slots][9] f64 | adj_all[16][max_slots] i32 | pnodes[16][8] i32   (16-byte aligned pieces)
    size_t b = sizeof(double) * (GATHER_PAIRS * GATHER_STAGE + (size_t)((max_slots + 1) & ~1) * 9) + 4 * GATHER_PAIRS * (size_t)((max_slots + 3) & ~3) +
               4 * GATHER_PAIRS * 8;
    return (b + 15) & ~(size_t)15;
}

// per (element, Gauss point): A = sqrt(w |det J|) J^-1 (src/fem.jl:192-196), so that g_a = dN_a A and G_ab += g_a g_b'.  The 8 (node,
// element) pairs of an element read it instead of re-evaluating the Jacobian 8 times (and the gather kernel needs no coordinates).
__global__ void __launch_bounds__(256) k_gather_elem(const int32_t *__restrict__ ien, int64_t nEl, const double *__restrict__ coords,
                                                     double *__restrict__ Ainv) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nEl * 8) return;
    const int64_t e = t >> 3;
    const int g = (int)(t & 7);
    double J[9], inv[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) J[q] = 0.0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {  // Jac = coords*dN, src/fem.jl:192
        const int64_t nb = ien[(int64_t)b * nEl + e];
        const double d0 = c_dN3[g][b][0], d1 = c_dN3[g][b][1], d2 = c_dN3[g][b][2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double x = coords[nb * 3 + r];
            J[r * 3 + 0] += x * d0;
            J[r * 3 + 1] += x * d1;
            J[r * 3 + 2] += x * d2;
        }
    }
    const double sw = sqrt(c_w3[g] * fabs(jac_inv<3>(J, inv)));  // :194-195
#pragma unroll
    for (int q = 0; q < 9; ++q) Ainv[t * 9 + q] = inv[q] * sw;
}

__global__ void __launch_bounds__(128, 4)
k_values_gather(const int32_t *__restrict__ ien, int64_t nEl, const double *__restrict__ Ainv, const int64_t *__restrict__ rowptr,
                const int32_t *__restrict__ colind, double *__restrict__ val, Material mat, const int64_t *__restrict__ g_ptr,
                const int32_t *__restrict__ g_ent, const int32_t *__restrict__ task_node, int ntasks, int max_slots) {
    constexpr int NN = 8, NGP = 8;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ double s_dN[NGP][NN][3];
    for (int t = threadIdx.x; t < NGP * NN * 3; t += blockDim.x) s_dN[t / (NN * 3)][(t % (NN * 3)) / 3][t % 3] = c_dN3[t / (NN * 3)][(t % (NN * 3)) / 3][t % 3];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int task = blockIdx.x * 4 + warp;
    if (task >= ntasks) return;
    unsigned char *mine_smem = s_dyn + warp * gather_warp_bytes(max_slots);
    double *stage = reinterpret_cast<double *>(mine_smem);        // [pair][73]: first the pair's 8 A matrices, then its 8 blocks
    double *red = stage + GATHER_PAIRS * GATHER_STAGE;             // [slot][9]: the node's row blocks with the material applied
    int32_t *adj_all = reinterpret_cast<int32_t *>(red + (size_t)((max_slots + 1) & ~1) * 9);  // [node of the task][slot]
    int32_t *pnodes = adj_all + GATHER_PAIRS * (size_t)((max_slots + 3) & ~3);  // [pair][8]: node ids of every pair's element (16-byte aligned)
    const int n_first = task_node[task], n_last = task_node[task + 1], nn_task = n_last - n_first;
    const int64_t base = g_ptr[n_first];
    const int cnt = (int)(g_ptr[n_last] - base);
    // two lanes per (node, element) pair: lane 2p + h accumulates the blocks b = 4h .. 4h + 3 of pair p (36 accumulators: ~120
    // registers, 16 warps per SM instead of 8)
    const int pair = lane >> 1, half = lane & 1;
    const bool active = pair < cnt;
    int a = 0;
    int64_t e = 0;
    if (active) {
        const int32_t ent = g_ent[base + pair];
        e = ent / NN;
        a = ent % NN;
    }
    // row starts / lengths and list bounds of the task's nodes: lane i <-> node n_first + i
    int64_t my_r0 = 0;
    int my_T = 0, my_p0 = 0, my_p1 = 0;
    if (lane < nn_task) {
        const int64_t n = n_first + lane;
        my_r0 = rowptr[3 * n];
        my_T = (int)(rowptr[3 * n + 1] - my_r0);
        my_p0 = (int)(g_ptr[n] - base);
        my_p1 = (int)(g_ptr[n + 1] - base);
    }
    if (active) {
        // the element's node ids (4 per lane) and its 8 A matrices (36 doubles per lane): asynchronous copies, all in flight at once
#pragma unroll
        for (int b = 0; b < 4; ++b) pnodes[pair * NN + 4 * half + b] = ien[(int64_t)(4 * half + b) * nEl + e];
        const double *Ae = Ainv + e * 72 + 36 * half;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + pair * GATHER_STAGE + 36 * half);
#pragma unroll 6
        for (int q = 0; q < 36; ++q) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * q), "l"(Ae + q) : "memory");
    }
    // the neighbour lists (= the columns of the nodes' first rows) of all nodes of the task, flattened over (node, slot)
    for (int idx0 = 0; idx0 < nn_task * max_slots; idx0 += 32) {  // uniform trip count: the shuffles need all lanes
        const int idx = idx0 + lane;
        const bool ok = idx < nn_task * max_slots;
        const int i = ok ? idx / max_slots : 0, sgm = idx - i * max_slots;
        const int64_t r0 = __shfl_sync(0xffffffffu, my_r0, i);
        const int nslots = __shfl_sync(0xffffffffu, my_T, i) / 3;
        if (ok && sgm < nslots) adj_all[idx] = colind[r0 + 3 * sgm] / 3;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    double G[4][9];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int m = 0; m < 9; ++m) G[b][m] = 0.0;
    if (active) {
#pragma unroll 1
        for (int g = 0; g < NGP; ++g) {
            double A[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) A[q] = stage[pair * GATHER_STAGE + g * 9 + q];
            double ga[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) ga[c] = s_dN[g][a][0] * A[c] + s_dN[g][a][1] * A[3 + c] + s_dN[g][a][2] * A[6 + c];  // sqrt(w |det|) dNdX, :196
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double *dn = s_dN[g][4 * half + b];
                double gb[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) gb[c] = dn[0] * A[c] + dn[1] * A[3 + c] + dn[2] * A[6 + c];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) G[b][i * 3 + j] += ga[i] * gb[j];
            }
        }
    }
    __syncwarp();  // both lanes of every pair have read the A matrices: the blocks may overwrite them
    if (active) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int m = 0; m < 9; ++m) stage[pair * GATHER_STAGE + (4 * half + b) * 9 + m] = G[b][m];
    }
    __syncwarp();
    for (int i = 0; i < nn_task; ++i) {
        const int p0 = __shfl_sync(0xffffffffu, my_p0, i), p1 = __shfl_sync(0xffffffffu, my_p1, i);  // this node's pairs [p0, p1)
        const int64_t r0 = __shfl_sync(0xffffffffu, my_r0, i);
        const int T = __shfl_sync(0xffffffffu, my_T, i);  // entries per row; the node's 3 rows are adjacent in K
        const int nslots = T / 3;
        // lane = neighbour slot: fold the blocks of the pairs whose element contains this neighbour, in ascending element order
        // (= the order sparse(E,J,V) folds duplicates); which local node it is inside the element: 8 compares per pair
        for (int sgm = lane; sgm < nslots; sgm += 32) {
            const int32_t want = adj_all[i * max_slots + sgm];
            double acc[9];
#pragma unroll
            for (int m = 0; m < 9; ++m) acc[m] = 0.0;
            for (int p = p0; p < p1; ++p) {
                const int4 lo4 = *reinterpret_cast<const int4 *>(pnodes + p * NN), hi4 = *reinterpret_cast<const int4 *>(pnodes + p * NN + 4);
                int b = -1;
                b = lo4.x == want ? 0 : b;
                b = lo4.y == want ? 1 : b;
                b = lo4.z == want ? 2 : b;
                b = lo4.w == want ? 3 : b;
                b = hi4.x == want ? 4 : b;
                b = hi4.y == want ? 5 : b;
                b = hi4.z == want ? 6 : b;
                b = hi4.w == want ? 7 : b;
                if (b >= 0) {
                    const double *src = stage + p * GATHER_STAGE + b * 9;
#pragma unroll
                    for (int m = 0; m < 9; ++m) acc[m] += src[m];
                }
            }
            // material on the summed block (src/fem.jl:236-249): red holds the 3 x 3 block of K for this neighbour
            const double tr = acc[0] + acc[4] + acc[8];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const double gij = acc[c * 3 + j], gji = acc[j * 3 + c];
                    red[sgm * 9 + c * 3 + j] = (c == j) ? mat.d11 * gij + mat.mu * (tr - gij) : mat.lam * gij + mat.mu * gji;
                }
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 3; ++c)  // row c of the node: entry s = 3 slot + j  <-  block[slot][c][j]; coalesced
            for (int sidx = lane; sidx < T; sidx += 32) {
                const int slot = sidx / 3, j = sidx - 3 * slot;
                val[r0 + (int64_t)c * T + sidx] = red[slot * 9 + c * 3 + j];
            }
        __syncwarp();
    }
}

void values_assemble_tile(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, Material mat, bool write_colind, const int *ready);  // assemble_tile.cu
bool values_tile_enabled();

void values_assemble(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, double Young, double nu, bool fuse_pattern, const int *ready) {
    Material mat;
    const int ndim = K->ndim, nDof = K->nDof;
    if (nDof == 2) {  // plane stress, src/fem.jl:217
        mat.d11 = Young / (1 - nu * nu);
        mat.lam = nu * Young / (1 - nu * nu);
        mat.mu = Young / (2 * (1 + nu));
    } else {  // src/fem.jl:230
        double f = Young / ((1 + nu) * (1 - 2 * nu));
        mat.d11 = (1 - nu) * f;
        mat.lam = nu * f;
        mat.mu = (1 - 2 * nu) / 2 * f;
    }
    K->Young = Young;
    K->nu = nu;
    K->beta_total = 0.0;
    K->mat_known = true;
    K->gmg_dirty = true;
    if (!K->val) {
        K->val = dev_alloc<double>(K->nnz_l + 16);
        CUDA_CHECK(cudaMemsetAsync(K->val + K->nnz_l, 0, 16 * sizeof(double), ctx->stream));
    }
    if (mesh->structured && ndim == 3 && nDof == 3 && values_tile_enabled()) {
        // fused assembly: colind is written by the tile kernel's output phase, rowptr in closed form by a small kernel.  That
        // kernel depends on nothing the tile kernel produces, so it runs on the (high-priority) side stream and fills the slots
        // the tile kernel's last wave leaves free instead of costing 0.03 ms of its own (SMFEM_ROWPTR_SIDE=0: in line, as before).
        static const bool side_ok = [] {
            const char *e = std::getenv("SMFEM_ROWPTR_SIDE");
            return !(e && e[0] == '0');
        }();
        const bool side = fuse_pattern && side_ok && ready == nullptr && ctx->copy_stream && ctx->ev_fork && ctx->ev_check;
        // SMFEM_COLIND_SIDE=1: the column indices come from the persistent side-stream kernel instead of the value kernel's output phase
        const char *cs_env = std::getenv("SMFEM_COLIND_SIDE");
        const bool colind_side = side && cs_env && cs_env[0] == '1';
        const unsigned rp_grid = (unsigned)((K->nrows_l + 1 + 255) / 256);
        if (fuse_pattern && !side) LAUNCH(ctx, k_struct_rowptr, rp_grid, 256, 0, mesh->lat, K->nDof, K->nrows_l, K->rowptr);
        if (side) {  // earlier work on the main stream may still read rowptr
            CUDA_CHECK(cudaEventRecord(ctx->ev_fork, ctx->stream));
            CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
        }
        if (colind_side) {  // launched first: one small CTA per SM, the value kernel's CTAs fill the rest
            k_struct_colind_side<<<ctx->sms, 64, 0, ctx->copy_stream>>>(mesh->lat, K->colind);
            ctx->launches++;
            CUDA_CHECK(cudaGetLastError());
        }
        values_assemble_tile(ctx, mesh, K, mat, fuse_pattern && !colind_side, ready);  // writes every entry and the diagonal: no memset, no extract_diag
        if (side) {
            k_struct_rowptr<<<rp_grid, 256, 0, ctx->copy_stream>>>(mesh->lat, K->nDof, K->nrows_l, K->rowptr);
            ctx->launches++;
            CUDA_CHECK(cudaGetLastError());
            CUDA_CHECK(cudaEventRecord(ctx->ev_check, ctx->copy_stream));
            CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_check, 0));
        }
        K->values_ready = true;
        return;
    } else {
        if (fuse_pattern) pattern_build_structured(ctx, mesh, K);  // no fused kernel for this element type: plain rebuild
        const char *ev0 = std::getenv("SMFEM_VALUES");
        if (!mesh->structured && mesh->ien && mesh->id == nullptr && ndim == 3 && nDof == 3 && mesh->nn == 8 && !(ev0 && ev0[0])) {
            // general hex mesh, standard dof map: gather form (SMFEM_VALUES=colored / atomic select the scatter forms)
            mesh_build_gather(ctx, mesh);
            int max_slots = 0;
            {
                int *d_m = dev_alloc<int>(1);
                CUDA_CHECK(cudaMemsetAsync(d_m, 0, sizeof(int), ctx->stream));
                LAUNCH(ctx, k_max_rowlen_nodes, (unsigned)((K->nrows_l / 3 + 255) / 256), 256, 0, K->nrows_l / 3, (const int64_t *)K->rowptr, d_m);
                CUDA_CHECK(cudaMemcpyAsync(&max_slots, d_m, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
                dev_free(d_m);
            }
            const size_t smem = 4 * gather_warp_bytes(max_slots);
            // (the per-element scratch is 576 B per element: beyond 8 GB the scatter forms take over)
            if (mesh->g_state == 1 && max_slots > 0 && max_slots < 255 && smem <= 110 * 1024 && mesh->nEl_g * 576 <= ((int64_t)8 << 30)) {
                static std::atomic<unsigned long long> attr_set{0};
                if (first_use_on_device(attr_set))
                    CUDA_CHECK(cudaFuncSetAttribute(k_values_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
                double *Ainv = dev_alloc<double>(mesh->nEl_g * 72);  // 576 B per element, read by its 8 (node, element) pairs
                LAUNCH(ctx, k_gather_elem, (unsigned)((mesh->nEl_g * 8 + 255) / 256), 256, 0, (const int32_t *)mesh->ien, mesh->nEl_g,
                       (const double *)mesh->coords, Ainv);
                LAUNCH(ctx, k_values_gather, (unsigned)((mesh->g_ntasks + 3) / 4), 128, smem, (const int32_t *)mesh->ien, mesh->nEl_g,
                       (const double *)Ainv, (const int64_t *)K->rowptr, (const int32_t *)K->colind, K->val, mat,
                       (const int64_t *)mesh->g_ptr, (const int32_t *)mesh->g_ent, (const int32_t *)mesh->g_task_node, mesh->g_ntasks, max_slots);
                K->values_ready = true;
                extract_diag(ctx, K);
                CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
                dev_free(Ainv);
                return;
            }
        }
        CUDA_CHECK(cudaMemsetAsync(K->val, 0, sizeof(double) * K->nnz_l, ctx->stream));
        Conn C = make_conn(mesh);
        DofMap D = make_dofmap(mesh, K);
        const int nn = mesh->nn;
        // general meshes: colour by colour without atomics (deterministic); SMFEM_VALUES=atomic keeps the atomic scatter
        const char *ev = std::getenv("SMFEM_VALUES");
        bool colored = !mesh->structured && mesh->ien && !(ev && std::string(ev) == "atomic");
        if (colored) {
            mesh_color_elements(ctx, mesh);
            colored = mesh->ncolors > 0;
        }
#define SMFEM_VALUES_CASE(ND, NDF, NNN)                                                                                              \
    if (colored) {                                                                                                                   \
        for (int c = 0; c < mesh->ncolors; ++c) {                                                                                    \
            const int64_t n_list = mesh->color_off[c + 1] - mesh->color_off[c];                                                      \
            if (n_list == 0) continue;                                                                                               \
            LAUNCH(ctx, (k_values_atomic<ND, NDF, NNN, true>), (unsigned)((n_list * nn + 127) / 128), 128, 0, C, D, mesh->coords,    \
                   K->rowptr, K->colind, K->val, mat, (const int32_t *)(mesh->elist + mesh->color_off[c]), n_list);                  \
        }                                                                                                                            \
    } else {                                                                                                                         \
        LAUNCH(ctx, (k_values_atomic<ND, NDF, NNN, false>), (unsigned)((C.nEl * nn + 127) / 128), 128, 0, C, D, mesh->coords,        \
               K->rowptr, K->colind, K->val, mat, (const int32_t *)nullptr, (int64_t)0);                                            \
    }
        if (ndim == 2 && nDof == 1 && nn == 9) {
            SMFEM_VALUES_CASE(2, 1, 9)
        } else if (ndim == 3 && nDof == 3) {
            SMFEM_VALUES_CASE(3, 3, 8)
        } else if (ndim == 3 && nDof == 1) {
            SMFEM_VALUES_CASE(3, 1, 8)
        } else if (ndim == 2 && nDof == 2) {
            SMFEM_VALUES_CASE(2, 2, 4)
        } else if (ndim == 2 && nDof == 1) {
            SMFEM_VALUES_CASE(2, 1, 4)
        } else {
            throw SmfemError(SMFEM_ERR_UNSUPPORTED, "assemble_system: supported (ndim,nDof) are (3,3) (3,1) (2,2) (2,1)");
        }
#undef SMFEM_VALUES_CASE
    }
    K->values_ready = true;
    extract_diag(ctx, K);
}

// ------------------------------------------------------------------------------------------------
// K4: b = int_{top U bottom} N'N dGamma  (examples/vector3D.jl:193-262) on K's pattern;
// K += beta*b (:308).  One thread per (face, local node a).
// ------------------------------------------------------------------------------------------------
__global__ void k_surface_mass(const int32_t *__restrict__ faces, int64_t nFaces, DofMap D,
                               const double *__restrict__ coords, const int64_t *__restrict__ rowptr,
                               const int32_t *__restrict__ colind, double *__restrict__ val, double *__restrict__ bval,
                               double beta) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nFaces * 4) return;
    int64_t f = t >> 2;
    int a = (int)(t & 3);
    int64_t nodes[4];
    double X[4][3];
    for (int b = 0; b < 4; ++b) {
        nodes[b] = faces[(int64_t)b * nFaces + f];
        for (int d = 0; d < 3; ++d) X[b][d] = coords[nodes[b] * 3 + d];
    }
    double acc[4] = {0, 0, 0, 0};
    for (int g = 0; g < 4; ++g) {
        double t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0};
        for (int b = 0; b < 4; ++b)
            for (int d = 0; d < 3; ++d) {
                t1[d] += X[b][d] * c_dN2[g][b][0];  // coords*dN, :224-225
                t2[d] += X[b][d] * c_dN2[g][b][1];
            }
        double cx = t1[1] * t2[2] - t1[2] * t2[1], cy = t1[2] * t2[0] - t1[0] * t2[2], cz = t1[0] * t2[1] - t1[1] * t2[0];
        double w = c_w2[g] * sqrt(cx * cx + cy * cy + cz * cz);  // :227-228
        for (int b = 0; b < 4; ++b) acc[b] += w * (c_N2[g][a] * c_N2[g][b]);  // be = M'M, :235
    }
    for (int c = 0; c < 3; ++c) {
        int64_t r = D.row(nodes[a], c);
        if (r < 0) continue;
        for (int b = 0; b < 4; ++b) {
            int64_t pos = csr_find(rowptr, colind, r, D.col(nodes[b], c));
            if (pos < 0) continue;
            if (bval) atomicAdd(&bval[pos], acc[b]);
            if (beta != 0.0) atomicAdd(&val[pos], beta * acc[b]);
        }
    }
}

void surface_mass(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, const int32_t *faces_dev, int64_t nFaces,
                  double beta, bool keep_b) {
    REQUIRE(K->ndim == 3 && K->nDof == 3, SMFEM_ERR_UNSUPPORTED,
            "apply_boundary_conditions: only the 3-D branch of the reference is executable");
    if (keep_b) {
        if (!K->bval) K->bval = dev_alloc<double>(K->nnz_l);
        CUDA_CHECK(cudaMemsetAsync(K->bval, 0, sizeof(double) * K->nnz_l, ctx->stream));
    }
    if (nFaces > 0) {
        DofMap D = make_dofmap(mesh, K);
        LAUNCH(ctx, k_surface_mass, (unsigned)((nFaces * 4 + 127) / 128), 128, 0, faces_dev, nFaces, D, mesh->coords,
               K->rowptr, K->colind, K->val, keep_b ? K->bval : (double *)nullptr, beta);
    }
    if (beta != 0.0) extract_diag(ctx, K);
    K->beta_total += beta;
    K->gmg_dirty = true;
}

// ------------------------------------------------------------------------------------------------
// K8: diagonal (Jacobi preconditioner)
// ------------------------------------------------------------------------------------------------
__global__ void k_extract_diag(int64_t nrows, int64_t ghost_cols, const int64_t *__restrict__ rowptr,
                               const int32_t *__restrict__ colind, const double *__restrict__ val,
                               double *__restrict__ diag) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    int64_t pos = csr_find(rowptr, colind, r, r + ghost_cols);
    diag[r] = pos >= 0 ? val[pos] : 0.0;
}

void extract_diag(smfem_ctx *ctx, smfem_matrix *K) {
    if (!K->diag) K->diag = dev_alloc<double>(K->nrows_l);
    LAUNCH(ctx, k_extract_diag, (unsigned)((K->nrows_l + 255) / 256), 256, 0, K->nrows_l, K->ghost_cols, K->rowptr,
           K->colind, K->val, K->diag);
}

// ------------------------------------------------------------------------------------------------
// Export as SparseMatrixCSC parts.  The pattern is structurally symmetric, so column j of the CSC
// has the row ids of CSR row j; the values are transposed through a search (nranks == 1) so that
// nzval really is K[i,j]; with nranks > 1 rows of other ranks are not resident and the slab of
// K' is returned instead (K is symmetric to rounding, ~1e-16 relative).
// ------------------------------------------------------------------------------------------------
__global__ void k_export_csc(int64_t nrows, int64_t ghost_cols, int64_t col_off_g, const int64_t *__restrict__ rowptr,
                             const int32_t *__restrict__ colind, const double *__restrict__ val, int transpose,
                             int64_t *__restrict__ colptr1, int64_t *__restrict__ rowval1, double *__restrict__ nzval) {
    int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (j > nrows) return;
    if (lane == 0) colptr1[j] = rowptr[j] + 1;
    if (j == nrows) return;
    for (int64_t p = rowptr[j] + lane; p < rowptr[j + 1]; p += 32) {
        int64_t c = colind[p];
        rowval1[p] = c + col_off_g + 1;
        if (val) {
            double v = val[p];
            if (transpose) {
                int64_t i = c - ghost_cols;  // local row of the mirror entry
                int64_t q = csr_find(rowptr, colind, i, j + ghost_cols);
                v = val[q];
            }
            nzval[p] = v;
        }
    }
}

void export_csc(smfem_ctx *ctx, smfem_matrix *K, int which, int64_t *colptr, int64_t *rowval, double *nzval) {
    const double *src = which == 1 ? K->bval : K->val;
    REQUIRE(which == 0 || which == 1 || which == 2, SMFEM_ERR_INVALID, "export: which must be 0 (K), 1 (b) or 2 (K, stored order)");
    REQUIRE(which != 1 || K->bval, SMFEM_ERR_INVALID, "export: surface matrix b was not kept");
    int64_t *d_colptr = dev_alloc<int64_t>(K->nrows_l + 1), *d_rowval = dev_alloc<int64_t>(K->nnz_l);
    double *d_nz = (nzval && src) ? dev_alloc<double>(K->nnz_l) : nullptr;
    int64_t col_off_g = K->row0 - K->ghost_cols;
    LAUNCH(ctx, k_export_csc, (unsigned)(((K->nrows_l + 1) * 32 + 255) / 256), 256, 0, K->nrows_l, K->ghost_cols,
           col_off_g, K->rowptr, K->colind, d_nz ? src : (const double *)nullptr, (ctx->nranks == 1 && which != 2) ? 1 : 0, d_colptr,
           d_rowval, d_nz);
    if (colptr) CUDA_CHECK(cudaMemcpyAsync(colptr, d_colptr, 8 * (K->nrows_l + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (rowval) CUDA_CHECK(cudaMemcpyAsync(rowval, d_rowval, 8 * K->nnz_l, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_nz) CUDA_CHECK(cudaMemcpyAsync(nzval, d_nz, 8 * K->nnz_l, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    dev_free(d_colptr);
    dev_free(d_rowval);
    dev_free(d_nz);
}
