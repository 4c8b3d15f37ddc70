// Mesh side of the path: host helper math (gaussian_quadrature / basis_function), device-side
// meshgrid + inflate_sphere, upload / validation of user meshes.
//   reference: src/fem.jl:21-31, :48-114; examples/vector3D.jl:10-130; src/PostProcess.jl:30-44
#include <cmath>
#include <cstring>

#include <mutex>
#include <unordered_map>

#include "smfem_internal.cuh"

// ------------------------------------------------------------------------------------------------
// caching device allocator (see smfem_internal.cuh)
// ------------------------------------------------------------------------------------------------
namespace {
std::mutex g_cache_mu;
std::unordered_map<void *, size_t> g_live;                    // ptr -> bytes (all blocks we handed out)
std::unordered_multimap<size_t, void *> g_free;               // bytes -> cached free blocks
size_t g_cached_bytes = 0;
constexpr size_t CACHE_MIN = (size_t)1 << 20;                 // only blocks >= 1 MiB are worth caching
constexpr size_t CACHE_MAX = (size_t)96 << 30;                // keep at most 96 GiB parked
}  // namespace

static size_t cache_key(size_t bytes) {  // blocks are only reused on the device they were allocated on
    int dev = 0;
    cudaGetDevice(&dev);
    return bytes ^ ((size_t)(dev + 1) << 56);
}

void *dev_cache_alloc(size_t bytes) {
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_free.find(cache_key(bytes));
        if (it != g_free.end()) {
            void *p = it->second;
            g_free.erase(it);
            g_cached_bytes -= bytes;
            g_live[p] = bytes;
            return p;
        }
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {  // out of memory: drop the cache and retry once
        cudaGetLastError();
        dev_cache_release();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess)
        throw SmfemError(SMFEM_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(bytes) + " bytes) -> " + cudaGetErrorString(e));
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_live[p] = bytes;
    return p;
}

void dev_cache_free(void *p) {
    size_t bytes = 0;
    bool park = false;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_live.find(p);
        if (it != g_live.end()) {
            bytes = it->second;
            g_live.erase(it);
            park = bytes >= CACHE_MIN && g_cached_bytes + bytes <= CACHE_MAX;
            if (park) g_cached_bytes += bytes;  // reserve the room now, publish the block below
        }
    }
    if (!park) {
        cudaFree(p);
        return;
    }
    // A context has two streams (main + copy stream) and several contexts may share a device, so a parked block may be handed
    // to a DIFFERENT stream than the one that last used it: drain the device first (what cudaFree would have done implicitly;
    // big blocks are freed at the end of a step, after its results were read).  Outside the lock: another rank's host thread
    // (smfem_init_multi) must be able to allocate while this device finishes.
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_free.emplace(cache_key(bytes), p);
}

void dev_cache_release() {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto &kv : g_free) cudaFree(kv.second);
    g_free.clear();
    g_cached_bytes = 0;
}

// ------------------------------------------------------------------------------------------------
// src/fem.jl:21-31 -- same expression order as the reference so its `==` tests hold bit-for-bit
// ------------------------------------------------------------------------------------------------
void smfem_host_gauss(double a, double b, int n, double *xi, double *w) {
    if (n == 2) {
        xi[0] = -(b - a) / (2 * std::sqrt(3.0)) + (b + a) / 2;
        xi[1] = (b - a) / (2 * std::sqrt(3.0)) + (b + a) / 2;
        w[0] = (b - a) / 2;
        w[1] = (b - a) / 2;
    } else if (n == 3) {
        xi[0] = -(b - a) / (2 * std::sqrt(5.0 / 3.0)) + (b + a) / 2;
        xi[1] = 0.0;  // the reference's literal 0 midpoint (src/fem.jl:27)
        xi[2] = (b - a) / (2 * std::sqrt(5.0 / 3.0)) + (b + a) / 2;
        w[0] = (b - a) / 2 * 5 / 9;
        w[1] = (b - a) / 2 * 8 / 9;
        w[2] = (b - a) / 2 * 5 / 9;
    } else {
        throw SmfemError(SMFEM_ERR_INVALID, "gaussian_quadrature: nGaussPoints must be 2 or 3 (reference: UndefVarError)");
    }
}

// ------------------------------------------------------------------------------------------------
// src/fem.jl:48-114.  dN is nn x ndim column-major (Julia Matrix); 1-D keeps the 1x2 row quirk.
// ------------------------------------------------------------------------------------------------
void smfem_host_basis(int ndim, int fc, double x, double e, double z, double *N, double *dN, int *nn_out) {
    if (fc == SMFEM_Q1 && ndim == 3) {
        const int nn = 8;
        const double Nv[8] = {(1 - x) * (1 - e) * (1 - z) / 8, (1 + x) * (1 - e) * (1 - z) / 8,
                              (1 + x) * (1 + e) * (1 - z) / 8, (1 - x) * (1 + e) * (1 - z) / 8,
                              (1 - x) * (1 - e) * (1 + z) / 8, (1 + x) * (1 - e) * (1 + z) / 8,
                              (1 + x) * (1 + e) * (1 + z) / 8, (1 - x) * (1 + e) * (1 + z) / 8};
        const double dx[8] = {-(1 - e) * (1 - z) / 8, (1 - e) * (1 - z) / 8,  (1 + e) * (1 - z) / 8, -(1 + e) * (1 - z) / 8,
                              -(1 - e) * (1 + z) / 8, (1 - e) * (1 + z) / 8,  (1 + e) * (1 + z) / 8, -(1 + e) * (1 + z) / 8};
        const double dy[8] = {-(1 - x) * (1 - z) / 8, -(1 + x) * (1 - z) / 8, (1 + x) * (1 - z) / 8, (1 - x) * (1 - z) / 8,
                              -(1 - x) * (1 + z) / 8, -(1 + x) * (1 + z) / 8, (1 + x) * (1 + z) / 8, (1 - x) * (1 + z) / 8};
        const double dz[8] = {-(1 - x) * (1 - e) / 8, -(1 + x) * (1 - e) / 8, -(1 + x) * (1 + e) / 8, -(1 - x) * (1 + e) / 8,
                              (1 - x) * (1 - e) / 8,  (1 + x) * (1 - e) / 8,  (1 + x) * (1 + e) / 8,  (1 - x) * (1 + e) / 8};
        for (int a = 0; a < nn; ++a) {
            N[a] = Nv[a];
            dN[a] = dx[a];
            dN[nn + a] = dy[a];
            dN[2 * nn + a] = dz[a];
        }
        *nn_out = nn;
    } else if (fc == SMFEM_Q1 && ndim == 2) {
        const int nn = 4;
        const double Nv[4] = {(1 - x) * (1 - e) / 4, (x + 1) * (1 - e) / 4, (1 + x) * (e + 1) / 4, (1 - x) * (1 + e) / 4};
        const double dx[4] = {-(1 - e) / 4, (1 - e) / 4, (e + 1) / 4, -(1 + e) / 4};
        const double dy[4] = {-(1 - x) / 4, -(x + 1) / 4, (1 + x) / 4, (1 - x) / 4};
        for (int a = 0; a < nn; ++a) {
            N[a] = Nv[a];
            dN[a] = dx[a];
            dN[nn + a] = dy[a];
        }
        *nn_out = nn;
    } else if (fc == SMFEM_Q1 && ndim == 1) {
        N[0] = 0.5 - 0.5 * x;
        N[1] = 0.5 + 0.5 * x;
        dN[0] = -0.5;  // 1x2 row matrix, src/fem.jl:75
        dN[1] = 0.5;
        *nn_out = 2;
    } else if (fc == SMFEM_Q2 && ndim == 2) {
        const int nn = 9;
        const double Nv[9] = {(1 - x) * x * (1 - e) * e / 4,          -x * (1 + x) * (1 - e) * e / 4,
                              x * (1 + x) * e * (1 + e) / 4,          -(1 - x) * x * e * (1 + e) / 4,
                              -(1 - x) * (1 + x) * (1 - e) * e / 2,   x * (1 + x) * (1 - e) * (1 + e) / 2,
                              (1 - x) * (1 + x) * e * (1 + e) / 2,    -(1 - x) * x * (1 - e) * (1 + e) / 2,
                              (1 - x) * (1 + x) * (1 - e) * (1 + e)};
        const double dx[9] = {(1 - 2 * x) * (1 - e) * e / 4,
                              -(1 + 2 * x) * (1 - e) * e / 4,
                              (1 + 2 * x) * e * (1 + e) / 4,
                              -(1 - 2 * x) * e * (1 + e) / 4,
                              x * (1 - e) * e,
                              (1 + 2 * x) * (1 - e) * (1 + e) / 2,
                              -x * e * (1 + e),
                              -(1 - 2 * x) * (1 - e) * (1 + e) / 2,
                              -2 * x * (1 - e) * (1 + e)};
        const double dy[9] = {(1 - x) * x * (1 - 2 * e) / 4,
                              -x * (1 + x) * (1 - 2 * e) / 4,
                              x * (1 + x) * (1 + 2 * e) / 4,
                              -(1 - x) * x * (1 + 2 * e) / 4,
                              -(1 - x) * (1 + x) * (1 - 2 * e) / 2,
                              -x * (1 + x) * e,
                              (1 - x) * (1 + x) * (1 + 2 * e) / 2,
                              (1 - x) * x * e,
                              -(1 - x) * (1 + x) * 2 * e};
        for (int a = 0; a < nn; ++a) {
            N[a] = Nv[a];
            dN[a] = dx[a];
            dN[nn + a] = dy[a];
        }
        *nn_out = nn;
    } else {
        throw SmfemError(SMFEM_ERR_UNSUPPORTED, "basis_function: the reference defines Q1 in 1/2/3-D and Q2 in 2-D only");
    }
}

// ------------------------------------------------------------------------------------------------
// examples/vector3D.jl:62-72: x = range(x0,x1,length=ne+1): correctly rounded x0+(x1-x0)*(i/ne),
// end points exact.  No FMA contraction, so the result is bit-identical to the oracle's.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double range_pt(double a, double b, int i, int ne) {
    if (i == 0) return a;
    if (i == ne) return b;
    return __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), __ddiv_rn((double)i, (double)ne)));
}

__global__ void k_meshgrid3d(Lattice lat, double x0, double x1, double y0, double y1, double z0, double z1,
                             double *__restrict__ coords) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = lat.nodes_local();
    if (t >= n) return;
    int i = (int)(t % lat.n1);
    int j = (int)((t / lat.n1) % lat.n1);
    int k = (int)(t / lat.plane()) + lat.k0 - 1;
    double x = 0, y = 0, z = 0;
    if (k >= 0 && k < lat.n1) {
        x = range_pt(x0, x1, i, lat.ne);
        y = range_pt(y0, y1, j, lat.ne);
        z = range_pt(z0, z1, k, lat.ne);
    }
    coords[3 * t + 0] = x;
    coords[3 * t + 1] = y;
    coords[3 * t + 2] = z;
}

void mesh_generate_structured(smfem_ctx *ctx, smfem_mesh *m, double x0, double x1, double y0, double y1, double z0,
                              double z1) {
    int64_t n = m->lat.nodes_local();
    LAUNCH(ctx, k_meshgrid3d, (unsigned)((n + 255) / 256), 256, 0, m->lat, x0, x1, y0, y1, z0, z1, m->coords);
}

// src/PostProcess.jl:30-44 (in place).  `scale ≈ 0.` with Julia's default atol=0 means scale == 0.
__global__ void k_inflate(int64_t n, int ndim, double cx, double cy, double *__restrict__ coords) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double dx = __dsub_rn(coords[ndim * t], cx), dy = __dsub_rn(coords[ndim * t + 1], cy);
    double scale = fmax(fabs(dx), fabs(dy));
    if (scale == 0.0) {
        coords[ndim * t] = 0.0;
        coords[ndim * t + 1] = 0.0;
    } else {
        double r = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        coords[ndim * t] = __ddiv_rn(__dmul_rn(scale, dx), r);
        coords[ndim * t + 1] = __ddiv_rn(__dmul_rn(scale, dy), r);
    }
}

void mesh_inflate(smfem_ctx *ctx, smfem_mesh *m, double x0, double x1, double y0, double y1) {
    double cx = 0.5 * (x0 + x1), cy = 0.5 * (y0 + y1);
    LAUNCH(ctx, k_inflate, (unsigned)((m->nNodes_l + 255) / 256), 256, 0, m->nNodes_l, m->ndim, cx, cy, m->coords);
}
