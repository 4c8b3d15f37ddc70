// Internal structures of libsmearfem_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/smearfem_b200.h"

struct SmfemError : std::runtime_error {
    int code;
    SmfemError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(x)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (x);                                                                           \
        if (e_ != cudaSuccess)                                                                          \
            throw SmfemError(SMFEM_ERR_CUDA, std::string(#x) + " -> " + cudaGetErrorString(e_) + " (" + \
                                                 __FILE__ + ":" + std::to_string(__LINE__) + ")");     \
    } while (0)

#define REQUIRE(cond, code, msg)                      \
    do {                                              \
        if (!(cond)) throw SmfemError((code), (msg)); \
    } while (0)

constexpr int SMFEM_MAX_RANKS = 8;

struct smfem_ctx {
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // user stopwatch (smfem_timer_*)
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;  // internal timing (pcg_solve, bench_spmv)
    // smfem_assemble_system: host->device copies and the lattice check run on copy_stream beside the assembly
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_nodes = nullptr, ev_check = nullptr;
    int *h_flags = nullptr;  // pinned, 32 ints: [0..3] check verdicts, [8..15] watermarks of the streamed NodeList
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};  // staging buffers of the chunked lattice check (lattice_check.cu)
    void *host_pool = nullptr;                     // HostPool*, created on first use
    int64_t h2d_bytes = 0, d2h_bytes = 0;          // bulk transfers of the reference-facing calls (smfem_transfer_bytes)
    // event pairs around the last ASM_RING launches of the tile kernel (smfem_assembly_kernel_ms: the roofline's denominator)
    static constexpr int ASM_RING = 64;
    cudaEvent_t asm_ev[2 * ASM_RING] = {};
    int64_t asm_count = 0;
    int sms = 148;
    int64_t launches = 0;
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
};

// Counts every kernel this library launches (bench.py reports it as gpu_launches).
#define LAUNCH(ctx, kernel, grid, block, smem, ...)                        \
    do {                                                                   \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);   \
        (ctx)->launches++;                                                 \
        CUDA_CHECK(cudaGetLastError());                                    \
    } while (0)

// Device allocations go through a small per-process cache (exact-size free lists): the reference-facing
// calls (mesh_from_host -> assemble -> free) would otherwise spend more time in cudaMalloc/cudaFree of
// multi-GB buffers than in the kernels.  Cached blocks are returned to the driver by dev_cache_release().
void *dev_cache_alloc(size_t bytes);
void dev_cache_free(void *p);
void dev_cache_release();

template <class T>
inline T *dev_alloc(size_t n) {
    if (n == 0) n = 1;
    return static_cast<T *>(dev_cache_alloc(n * sizeof(T)));
}
template <class T>
inline void dev_free(T *&p) {
    if (p) dev_cache_free((void *)p);
    p = nullptr;
}

// ---------------------------------------------------------------------------------------------
// Structured hex lattice (examples/vector3D.jl:60-127): n1 = ne+1 nodes per axis, node id
// = k*n1^2 + j*n1 + i (x fastest).  A rank owns node planes [k0, k1); its LOCAL arrays always carry
// one ghost plane below and one above (unused at the ends of the domain), so the local plane of
// global plane k is  k - k0 + 1  and every rank has the same layout
//     [ ghost_lo | owned planes ... | ghost_hi ].
// ---------------------------------------------------------------------------------------------
struct Lattice {
    int n1 = 0, ne = 0;
    int k0 = 0, k1 = 0;  // owned planes [k0,k1)
    __host__ __device__ int64_t plane() const { return (int64_t)n1 * n1; }
    __host__ __device__ int nown() const { return k1 - k0; }
    __host__ __device__ int64_t nodes_local() const { return (int64_t)(k1 - k0 + 2) * n1 * n1; }
    // local node index (with ghosts) of global (i,j,k); k may be k0-1 .. k1
    __host__ __device__ int64_t lnode(int i, int j, int k) const {
        return ((int64_t)(k - k0 + 1) * n1 + j) * n1 + i;
    }
};

inline void slab_range(int n1, int rank, int nranks, int &k0, int &k1) {
    k0 = (int)((int64_t)n1 * rank / nranks);
    k1 = (int)((int64_t)n1 * (rank + 1) / nranks);
}

struct smfem_mesh {
    smfem_ctx *ctx = nullptr;
    int ndim = 3, nn = 8;
    bool structured = false;
    int64_t ne = 0;
    int64_t nNodes_g = 0, nEl_g = 0;
    Lattice lat;
    int64_t nNodes_l = 0;      // local nodes incl. ghost planes (== nNodes_g for general meshes)
    double *coords = nullptr;  // ndim x nNodes_l, xyz of a node contiguous (Julia NodeList layout)
    // general (unstructured) meshes only
    int32_t *ien = nullptr;  // [a][e], 0-based node ids (SoA like Julia's column-major IEN)
    int32_t *id = nullptr;   // [comp][node], 0-based dof ids; nullptr -> dof = nDof*node+comp
    int nDof_id = 0;
    bool std_id = false;     // ID is the standard node-major map (or absent): dof = nDof*node + comp
    int64_t ndof_id = 0;  // max(ID)
    // element colouring (general meshes; built on first use by mesh_color_elements): elements of one colour share no node
    int32_t *elist = nullptr;  // element ids sorted by colour
    int ncolors = 0;           // 0: not built, -1: not colourable with <= 64 colours (atomic scatter is used)
    int64_t color_off[65] = {};
    // gather plan (general 3-D hex meshes with the standard dof map; built on first use by mesh_build_gather): the (element, local
    // node) pairs of every node, sorted, and a partition of the nodes into warp tasks of <= 32 pairs
    int64_t *g_ptr = nullptr;        // nNodes + 1
    int32_t *g_ent = nullptr;        // nEl * nn entries  e * nn + a
    int32_t *g_task_node = nullptr;  // g_ntasks + 1: first node of every warp task
    int g_ntasks = 0;
    int g_state = 0;                 // 0: not built, 1: usable, -1: not usable (a node with more than 32 elements)
    // surface faces for general meshes are passed to smfem_surface_mass directly
};

// all-reduce mailboxes + halo flags, at the start of each rank's peer window
struct CommHeader {
    // slot = seq & 3: the (r'z, r'r) sums carry even sequence numbers, p'Ap the odd ones, so each kind alternates between two
    // slots.  A rank publishes seq + 4 only after it has fetched seq + 2 from every rank, and every rank publishes seq + 2 only
    // after all its reads of seq (stream order) - also across consecutive solves, which need no host barrier.
    double mbox[4][SMFEM_MAX_RANKS][4];
    unsigned long long mflag[4][SMFEM_MAX_RANKS];
    unsigned long long hflag[2];  // [0]: ghost_lo filled up to seq, [1]: ghost_hi
    unsigned long long pad[6];
    // multigrid-PCG (gmg.cu): its own halo flags / acknowledgements / all-gather flags / all-reduce mailbox, numbered by
    // the device counters hseq / gseq / rseq of PcgScalars (every rank runs the same sequence of exchanges)
    unsigned long long ghflag[2];   // [0]: my ghost_lo holds the lower neighbour's push number >= value, [1]: ghost_hi
    unsigned long long gaflag[2];   // [0]: the lower neighbour has consumed my push number <= value, [1]: the upper one
    unsigned long long gathflag[SMFEM_MAX_RANKS];
    unsigned long long gflag[4][SMFEM_MAX_RANKS];
    double gbox[4][SMFEM_MAX_RANKS][4];
};

struct PcgScalars {
    double rzs[2], pAp, alpha, beta, rr, bnorm2, spare;
    unsigned long long it;      // global iteration counter (never reset: sequence for flags)
    unsigned long long red;     // global all-reduce sequence counter
    unsigned int ticketA, ticketB, ticketC, breakdown;
    // per-solve control, written by k_pcg_init: the convergence test runs on the device (every rank sees bitwise identical
    // sums, so all ranks stop in the same iteration); once `done` is set the remaining kernels of a graph replay are no-ops
    double rtol2, rr_true;
    unsigned int iters, maxit, done, pad_;
    unsigned long long hseq, gseq, rseq;  // multigrid-PCG: halo pushes / all-gathers / all-reduces issued so far
    // Jacobi-PCG wait accounting of the current solve (ns of %globaltimer, CTA 0 only): time spent spinning on the neighbours'
    // halo flags inside the boundary-plane SpMV launches, and on the two all-reduce mailboxes (r'z / r'r before update_p, p'Ap
    // before update_xr).  Reset by k_pcg_init; smfem_pcg_wait_stats reports them.
    unsigned long long t_wait_halo, t_wait_rz, t_wait_pap;
};

struct CommView {  // passed by value to kernels
    int rank = 0, nranks = 1;
    CommHeader *self = nullptr;
    CommHeader *peer[SMFEM_MAX_RANKS] = {nullptr};
    double *peer_p[SMFEM_MAX_RANKS] = {nullptr};  // peers' p vectors (inside their windows)
    int64_t lo_dst_off = 0;  // offset in peer (rank-1)'s p of its ghost_hi plane
    int64_t plane_dofs = 0;  // dofs per node plane
};

struct smfem_matrix {
    smfem_ctx *ctx = nullptr;
    int ndim = 3, nDof = 3, nn = 8;
    bool structured = false;
    Lattice lat;
    int64_t m_g = 0, nnz_g = 0;
    int64_t row0 = 0, nrows_l = 0, ncols_l = 0, ghost_cols = 0, nnz_l = 0;
    int64_t *rowptr = nullptr;  // nrows_l+1, local offsets
    int32_t *colind = nullptr;  // local column index (into vectors with ghost planes)
    double *val = nullptr;
    double *bval = nullptr;  // surface matrix on the same pattern (only when kept)
    double *diag = nullptr;
    bool values_ready = false;
    // what the values were assembled from (the multigrid preconditioner re-assembles coarse operators with it)
    double Young = 0, nu = 0, beta_total = 0;
    bool mat_known = false;
    void *gmg = nullptr;            // Gmg* (gmg.cu), built at the first multigrid solve
    smfem_mesh *gmg_mesh = nullptr;  // not owned
    bool gmg_on = false, gmg_dirty = true;
    bool csr_less = false;            // smfem_matfree_operator: no rowptr / colind / val at all (the operator is applied matrix-free only)
    bool matfree_on = false;          // the solve's operator is applied matrix-free from mf_mesh's coordinates (matfree.cu)
    smfem_mesh *mf_mesh = nullptr;    // not owned
    bool gmg_coarse = false;          // a coarse-level operator owned by a multigrid hierarchy: no peer window region of its own
    int64_t gmg_region_doubles = 0;   // multi-GPU: doubles reserved behind p in the peer window for the hierarchy's exchanged vectors
    // Dirichlet
    uint8_t *fixed = nullptr;  // nrows_l
    double *qd = nullptr;      // ncols_l (ghost entries filled locally)
    bool has_bc = false;
    // solver workspace
    double *x = nullptr, *r = nullptr, *Ap = nullptr, *dinv = nullptr, *partials = nullptr;
    int64_t partials_n = 0;
    double *p = nullptr;  // ncols_l; lives inside `window`
    void *window = nullptr;
    size_t window_bytes = 0;
    PcgScalars *scal = nullptr;
    double *h_pinned = nullptr;
    CommView comm;
    void *peer_maps[SMFEM_MAX_RANKS] = {nullptr};
    bool comm_connected = false;
    const double *sol_x = nullptr;  // free part of the last solve's solution (device, nrows_l; q = q_d + x)
    double warm_scale = 0.0;  // next solve starts from warm_scale * (previous solution); reset after use
    void *pcg_graph = nullptr;  // cudaGraphExec_t of PCG_CHUNK iterations, instantiated once per (matrix, SpMV variant)
    int pcg_graph_variant = -1;
    int64_t pcg_graph_launches = 0;
    double last_relres_rec = 0, last_relres_true = 0;
    int spmv_variant = 4;  // 4 = row-triple (default; falls back to 2), 2 = CSR-stream, 3 = CSR-stream via TMA, 1 = warp/row, 0 = warp/3 rows
    int32_t *blk_row = nullptr;  // CSR-stream row blocks
    int nblk = 0, max_rowlen = 0, ctas_per_sm = 4;
    bool stream_ok = false, group3_ok = false;
    // stats
    float last_ms = 0, last_ms_spmv = 0;
    double last_wait_us[3] = {0, 0, 0};  // halo flags, r'z all-reduce, p'Ap all-reduce (CTA 0, whole solve)
    int last_iters = 0;
};

#ifdef __CUDACC__
// system-scope flag helpers of the peer-memory protocols (solver.cu, gmg.cu)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double ld_volatile_f64(const double *p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
#endif

// --- functions implemented across translation units ---------------------------------------------
void smfem_host_gauss(double a, double b, int n, double *xi, double *w);
void smfem_host_basis(int ndim, int func_class, double xi, double eta, double zeta, double *N, double *dN, int *nn);

void mesh_upload_tables();  // quadrature tables -> __constant__ (assemble.cu)
void mesh_generate_structured(smfem_ctx *ctx, smfem_mesh *m, double x0, double x1, double y0, double y1, double z0,
                              double z1);
void mesh_inflate(smfem_ctx *ctx, smfem_mesh *m, double x0, double x1, double y0, double y1);

constexpr int64_t LATTICE_CHUNK = 128 * 1024;  // Int64 entries per chunk of the hybrid lattice check (1 MiB)
bool lattice_check_hybrid(smfem_ctx *ctx, const Lattice &L, const int64_t *IEN, const int64_t *ID, int64_t nEl, int64_t nNodes, int ne,
                          int64_t *d_stage, int *d_flag);
void host_pool_destroy(smfem_ctx *ctx);
void pattern_prepare_structured(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K);
void pattern_build_structured(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K);
bool values_tile_enabled();
void pattern_build_general(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K);
void mesh_color_elements(smfem_ctx *ctx, smfem_mesh *mesh);  // assemble.cu
// ready (device int, optional): the coordinate planes [0, *ready) have been written; the tile kernel waits plane by plane
void values_assemble(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, double Young, double nu, bool fuse_pattern = false,
                     const int *ready = nullptr);
void surface_mass(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, const int32_t *faces_dev, int64_t nFaces,
                  double beta, bool keep_b);
void extract_diag(smfem_ctx *ctx, smfem_matrix *K);
void export_csc(smfem_ctx *ctx, smfem_matrix *K, int which, int64_t *colptr, int64_t *rowval, double *nzval);

void solver_alloc(smfem_ctx *ctx, smfem_matrix *K);
void solver_free(smfem_matrix *K);
void dirichlet_zplanes(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, double d);
void dirichlet_list(smfem_ctx *ctx, smfem_matrix *K, const int64_t *dofs, const double *vals, int64_t n);
void pcg_solve(smfem_ctx *ctx, smfem_matrix *K, double rtol, int maxit, const double *rhs_extra, double *q_out,
               int *iters, double *relres);
void spmv_host(smfem_ctx *ctx, smfem_matrix *K, const double *x, double *y);
void spmv_device(smfem_ctx *ctx, smfem_matrix *K, const double *x, double *y);  // device vectors: x ncols_l (ghost planes), y nrows_l
void project_nodes(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, const int64_t *ids, int64_t n, const double *cam,
                   double *nodes3d_out, double *nodes2d_out);  // postprocess.cu
// abi.cu: a hex-lattice mesh handle for an explicit slab of node planes [k0, k1) (coordinates allocated, not filled)
smfem_mesh *mesh_new_lattice(smfem_ctx *ctx, int64_t ne, int k0, int k1);
// gmg.cu: doubles a rank's peer window needs behind p for the multigrid hierarchy of the ne^3 lattice (0 on one GPU)
int64_t gmg_window_doubles(int ne, int rank, int nranks);
// gmg.cu: geometric-multigrid preconditioned CG (hex lattice; one GPU or z-slabs over several)
void gmg_enable(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, bool enable);
void gmg_free(smfem_matrix *K);
void gmg_apply_host(smfem_ctx *ctx, smfem_matrix *K, const double *r_host, double *z_host);
void gmg_pcg_solve(smfem_ctx *ctx, smfem_matrix *K, double rtol, int maxit, const double *rhs_extra, double *q_out, int *iters,
                   double *relres);
void bench_spmv(smfem_ctx *ctx, smfem_matrix *K, int variant, int reps, float *ms);
void comm_export(smfem_ctx *ctx, smfem_matrix *K, void *handle_out);
void comm_connect(smfem_ctx *ctx, smfem_matrix *K, const void *handles);
void comm_connect_local(smfem_ctx *ctx, smfem_matrix *K, smfem_matrix *const *all_K, int n);  // same process: raw peer pointers
void smfem_set_last_error(const char *msg);
void extract_borders(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, const int64_t *ids, int64_t n, const double *cam, int state,
                     int64_t ne, double *border_out, int64_t cap, int64_t *nborder, double *side2d_out);  // postprocess.cu
// matfree.cu: y = (K + beta b) x from the lattice coordinates (no CSR arrays read); same vector layout as the CSR SpMV
void matfree_apply(smfem_ctx *ctx, smfem_matrix *K, const double *x, double *y, bool halo, bool check_done, unsigned long long halo_need);
void matfree_diag(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh);                          // diag(K) from the coordinates
void matfree_add_surface(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, double beta);      // CSR-less K += beta b  // abi.cu: message for the calling thread's smfem_last_error()

// true exactly once per (call site, device): per-function attributes (dynamic shared memory limits) are per device, and one
// process may drive several GPUs (smfem_init_multi)
#include <atomic>
inline bool first_use_on_device(std::atomic<unsigned long long> &mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return !(mask.fetch_or(bit) & bit);
}
