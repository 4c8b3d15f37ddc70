// extern "C" boundary of libsmearfem_b200.so (see include/smearfem_b200.h).  Every entry point
// catches C++ exceptions and converts them into status codes + a thread-local message.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "smfem_internal.cuh"

static thread_local std::string g_last_error;

void smfem_set_last_error(const char *msg) { g_last_error = msg ? msg : ""; }

template <class F>
static int guarded(F &&f) {
    try {
        f();
        return SMFEM_OK;
    } catch (const SmfemError &e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::exception &e) {
        g_last_error = std::string("internal error: ") + e.what();
        return SMFEM_ERR_INVALID;
    } catch (...) {
        g_last_error = "unknown internal error";
        return SMFEM_ERR_INVALID;
    }
}

#define NOTNULL(p) REQUIRE((p) != nullptr, SMFEM_ERR_INVALID, #p " must not be NULL")

// ------------------------------------------------------------------------------------------------
// device helpers for user meshes
// ------------------------------------------------------------------------------------------------
// converts Julia Int64 1-based indices to int32 0-based, flagging out-of-range entries
__global__ void k_convert_index(int64_t n, const int64_t *__restrict__ in, int64_t lo, int64_t hi,
                                int32_t *__restrict__ out, int *__restrict__ err) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int64_t v = in[t];
    if (v < lo || v > hi) {
        *err = 1;
        v = lo;
    }
    out[t] = (int32_t)(v - 1);
}

// is (IEN, ID) exactly what meshgrid(…, ne, 3) produces?  (examples/vector3D.jl:74, :94-101)
__global__ void k_check_lattice(int64_t nEl, int ne, const int64_t *__restrict__ IEN, int64_t nNodes, int nDof,
                                const int64_t *__restrict__ ID, int *__restrict__ mismatch) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int n1 = ne + 1;
    if (t < nEl * 8) {
        int64_t e = t % nEl;
        int a = (int)(t / nEl);
        int ei = (int)(e % ne), ej = (int)((e / ne) % ne), ek = (int)(e / ((int64_t)ne * ne));
        int ox = ((a & 3) == 1 || (a & 3) == 2), oy = ((a & 3) >= 2), oz = (a >> 2);
        int64_t want = ((int64_t)(ek + oz) * n1 + (ej + oy)) * n1 + (ei + ox) + 1;
        if (IEN[t] != want) mismatch[0] = 1;
    }
    if (ID && t < nNodes * nDof) {
        int64_t m = t % nNodes;
        int l = (int)(t / nNodes);
        if (ID[t] != nDof * m + l + 1) mismatch[1] = 1;
    }
}

// is ID the standard node-major map ID[m,l] = nDof*(m-1)+l (examples/vector3D.jl:74)?
__global__ void k_check_std_id(int64_t nNodes, int nDof, const int64_t *__restrict__ ID, int *__restrict__ mismatch) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nNodes * nDof) return;
    int64_t m = t % nNodes;
    int l = (int)(t / nNodes);
    if (ID[t] != nDof * m + l + 1) *mismatch = 1;
}

__global__ void k_max_i64(int64_t n, const int64_t *__restrict__ in, unsigned long long *__restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) atomicMax(out, (unsigned long long)in[t]);
}

// top / bottom faces of the structured lattice as local node ids [4][nFaces]
// (examples/vector3D.jl:102-113: bottom = local nodes 1-4 of layer 1, top = local nodes 5-8 of layer ne)
__global__ void k_struct_faces(Lattice L, int has_btm, int has_top, int32_t *__restrict__ faces, int64_t nFaces) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nFaces) return;
    int64_t per = (int64_t)L.ne * L.ne;
    bool top = has_btm ? (t >= per) : true;
    int64_t f = t % per;
    int fi = (int)(f % L.ne), fj = (int)(f / L.ne);
    int k = top ? L.ne : 0;
    (void)has_top;
    faces[0 * nFaces + t] = (int32_t)L.lnode(fi, fj, k);
    faces[1 * nFaces + t] = (int32_t)L.lnode(fi + 1, fj, k);
    faces[2 * nFaces + t] = (int32_t)L.lnode(fi + 1, fj + 1, k);
    faces[3 * nFaces + t] = (int32_t)L.lnode(fi, fj + 1, k);
}

__global__ void k_copy_slab_coords(Lattice L, const double *__restrict__ global, double *__restrict__ local) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.nodes_local()) return;
    int k = (int)(t / L.plane()) + L.k0 - 1;
    if (k < 0 || k >= L.n1) return;
    int64_t g = (int64_t)k * L.plane() + t % L.plane();
    local[3 * t] = global[3 * g];
    local[3 * t + 1] = global[3 * g + 1];
    local[3 * t + 2] = global[3 * g + 2];
}

static void set_lattice(smfem_ctx *ctx, smfem_mesh *m, int64_t ne) {
    m->structured = true;
    m->ne = ne;
    m->ndim = 3;
    m->nn = 8;
    m->lat.ne = (int)ne;
    m->lat.n1 = (int)ne + 1;
    slab_range(m->lat.n1, ctx->rank, ctx->nranks, m->lat.k0, m->lat.k1);
    REQUIRE(m->lat.k1 > m->lat.k0, SMFEM_ERR_INVALID, "mesh too small for this many ranks (need >= 1 node plane per rank)");
    m->nNodes_g = (int64_t)m->lat.n1 * m->lat.n1 * m->lat.n1;
    m->nEl_g = ne * ne * ne;
    m->nNodes_l = m->lat.nodes_local();
    REQUIRE(m->nNodes_l * 3 < (int64_t)INT32_MAX, SMFEM_ERR_UNSUPPORTED, "slab too large for int32 local indices");
}

smfem_mesh *mesh_new_lattice(smfem_ctx *ctx, int64_t ne, int k0, int k1) {
    smfem_mesh *m = new smfem_mesh();
    m->ctx = ctx;
    m->structured = true;
    m->ne = ne;
    m->ndim = 3;
    m->nn = 8;
    m->lat.ne = (int)ne;
    m->lat.n1 = (int)ne + 1;
    m->lat.k0 = k0;
    m->lat.k1 = k1;
    m->nNodes_g = (int64_t)m->lat.n1 * m->lat.n1 * m->lat.n1;
    m->nEl_g = ne * ne * ne;
    m->nNodes_l = m->lat.nodes_local();
    if (!(k1 > k0 && k0 >= 0 && k1 <= m->lat.n1)) {
        delete m;
        throw SmfemError(SMFEM_ERR_INVALID, "mesh_new_lattice: bad slab");
    }
    m->coords = dev_alloc<double>(3 * m->nNodes_l);
    return m;
}

extern "C" {

int smfem_abi_version(void) { return 2; }
const char *smfem_last_error(void) { return g_last_error.c_str(); }

int smfem_gaussian_quadrature(double a, double b, int n, double *xi, double *w) {
    return guarded([&] {
        NOTNULL(xi);
        NOTNULL(w);
        smfem_host_gauss(a, b, n, xi, w);
    });
}

int smfem_basis_function(int ndim, int func_class, double xi, double eta, double zeta, double *N, double *dN, int *nn) {
    return guarded([&] {
        NOTNULL(N);
        NOTNULL(dN);
        NOTNULL(nn);
        smfem_host_basis(ndim, func_class, xi, eta, zeta, N, dN, nn);
    });
}

int smfem_init(int device, int rank, int nranks, smfem_ctx **out) {
    return guarded([&] {
        NOTNULL(out);
        REQUIRE(nranks >= 1 && nranks <= SMFEM_MAX_RANKS && rank >= 0 && rank < nranks, SMFEM_ERR_INVALID,
                "smfem_init: need 0 <= rank < nranks <= 8");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw SmfemError(SMFEM_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
                                                 cudaGetErrorString(e));
        REQUIRE(device >= 0 && device < count, SMFEM_ERR_INVALID, "smfem_init: bad device ordinal");
        CUDA_CHECK(cudaSetDevice(device));
        smfem_ctx *c = new smfem_ctx();
        c->device = device;
        c->rank = rank;
        c->nranks = nranks;
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreate(&c->ev0));
        CUDA_CHECK(cudaEventCreate(&c->ev1));
        CUDA_CHECK(cudaEventCreate(&c->ev2));
        CUDA_CHECK(cudaEventCreate(&c->ev3));
        {  // kernels of the copy stream (lattice check) should slip in beside the assembly kernel
            int lo_prio = 0, hi_prio = 0;
            CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
            CUDA_CHECK(cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, hi_prio));
        }
        CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_nodes, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_check, cudaEventDisableTiming));
        CUDA_CHECK(cudaMallocHost(&c->h_flags, 32 * sizeof(int)));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_stage[0], cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_stage[1], cudaEventDisableTiming));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        c->sms = prop.multiProcessorCount;
        mesh_upload_tables();
        *out = c;
    });
}

int smfem_destroy(smfem_ctx *ctx) {
    return guarded([&] {
        if (!ctx) return;
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        dev_cache_release();
        if (ctx->flush_buf) cudaFree(ctx->flush_buf);
        cudaEventDestroy(ctx->ev0);
        cudaEventDestroy(ctx->ev1);
        cudaEventDestroy(ctx->ev2);
        cudaEventDestroy(ctx->ev3);
        cudaStreamSynchronize(ctx->copy_stream);
        host_pool_destroy(ctx);
        for (cudaEvent_t e : ctx->asm_ev)
            if (e) cudaEventDestroy(e);
        cudaEventDestroy(ctx->ev_stage[0]);
        cudaEventDestroy(ctx->ev_stage[1]);
        cudaEventDestroy(ctx->ev_fork);
        cudaEventDestroy(ctx->ev_nodes);
        cudaEventDestroy(ctx->ev_check);
        cudaStreamDestroy(ctx->copy_stream);
        cudaFreeHost(ctx->h_flags);
        cudaStreamDestroy(ctx->stream);
        delete ctx;
    });
}

int smfem_stream(smfem_ctx *ctx, void **stream_out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(stream_out);
        *stream_out = (void *)ctx->stream;
    });
}

int smfem_timer_start(smfem_ctx *ctx) {
    return guarded([&] {
        NOTNULL(ctx);
        CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    });
}

int smfem_timer_stop(smfem_ctx *ctx, float *ms_out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(ms_out);
        CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
        CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
        CUDA_CHECK(cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
    });
}

int smfem_sync(smfem_ctx *ctx) {
    return guarded([&] {
        NOTNULL(ctx);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    });
}

int smfem_launch_count(smfem_ctx *ctx, int64_t *count_out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(count_out);
        *count_out = ctx->launches;
    });
}

int smfem_flush_l2(smfem_ctx *ctx) {
    return guarded([&] {
        NOTNULL(ctx);
        if (!ctx->flush_buf) {
            ctx->flush_bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
            CUDA_CHECK(cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
        }
        CUDA_CHECK(cudaMemsetAsync(ctx->flush_buf, 0, ctx->flush_bytes, ctx->stream));
    });
}

int smfem_meshgrid(smfem_ctx *ctx, double x0, double x1, double y0, double y1, double z0, double z1, int64_t ne, int ndim,
                   smfem_mesh **out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(out);
        REQUIRE(ne >= 1 && ne <= 1200, SMFEM_ERR_INVALID, "meshgrid: need 1 <= ne <= 1200");
        REQUIRE(ndim == 3, SMFEM_ERR_UNSUPPORTED,
                "device meshgrid covers the 3-D branch (examples/vector3D.jl:60-127); build 2-D meshes on the host and "
                "pass them through smfem_mesh_from_host");
        CUDA_CHECK(cudaSetDevice(ctx->device));
        smfem_mesh *m = new smfem_mesh();
        m->ctx = ctx;
        set_lattice(ctx, m, ne);
        m->coords = dev_alloc<double>(3 * m->nNodes_l);
        mesh_generate_structured(ctx, m, x0, x1, y0, y1, z0, z1);
        *out = m;
    });
}

int smfem_mesh_from_host(smfem_ctx *ctx, const double *NodeList, const int64_t *IEN, const int64_t *ID, int64_t nNodes,
                         int64_t nEl, int nLocal, int ndim, int nDof, int64_t ne, smfem_mesh **out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(out);
        NOTNULL(NodeList);
        NOTNULL(IEN);
        REQUIRE(ndim == 2 || ndim == 3, SMFEM_ERR_UNSUPPORTED,
                "assemble_system: the reference's 1-D branch is not executable (src/fem.jl:75 vs :192)");
        REQUIRE(nLocal == (1 << ndim) || (ndim == 2 && nLocal == 9), SMFEM_ERR_UNSUPPORTED,
                "elements must be Q1 (2^ndim nodes) or the reference's 2-D Q2 quad (9 nodes)");
        REQUIRE(nNodes >= 1 && nEl >= 1 && nNodes < (int64_t)INT32_MAX / 4, SMFEM_ERR_INVALID, "bad mesh sizes");
        REQUIRE(nDof >= 1 && nDof <= 3, SMFEM_ERR_INVALID, "nDof must be 1, 2 or 3");
        REQUIRE(ID != nullptr || nDof == 1, SMFEM_ERR_INVALID, "ID is required when nDof > 1 (reference: MethodError on size(nothing,2))");
        CUDA_CHECK(cudaSetDevice(ctx->device));
        smfem_mesh *m = new smfem_mesh();
        m->ctx = ctx;
        m->ndim = ndim;
        m->nn = nLocal;
        m->ne = ne;
        m->nNodes_g = nNodes;
        m->nEl_g = nEl;
        // stage the Int64 arrays on the device
        int64_t *d_ien = dev_alloc<int64_t>(nEl * nLocal), *d_id = nullptr;
        CUDA_CHECK(cudaMemcpyAsync(d_ien, IEN, 8 * nEl * nLocal, cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d_bytes += 8 * nEl * nLocal + (ID ? 8 * nNodes * nDof : 0) + 8 * (int64_t)ndim * nNodes;
        if (ID) {
            d_id = dev_alloc<int64_t>(nNodes * nDof);
            CUDA_CHECK(cudaMemcpyAsync(d_id, ID, 8 * nNodes * nDof, cudaMemcpyHostToDevice, ctx->stream));
        }
        int *d_flag = dev_alloc<int>(4);  // [0] IEN != lattice, [1] ID != standard (lattice check), [2] range error, [3] ID != standard
        CUDA_CHECK(cudaMemsetAsync(d_flag, 0, 16, ctx->stream));
        bool lattice = false;
        if (ndim == 3 && ctx->nranks >= 1 && ne >= 1 && nEl == ne * ne * ne && nNodes == (ne + 1) * (ne + 1) * (ne + 1) &&
            (nDof == 3 || nDof == 1)) {
            int64_t nt = nEl * 8 > nNodes * nDof ? nEl * 8 : nNodes * nDof;
            LAUNCH(ctx, k_check_lattice, (unsigned)((nt + 255) / 256), 256, 0, nEl, (int)ne, (const int64_t *)d_ien, nNodes,
                   nDof, (const int64_t *)d_id, d_flag);
            int h[2] = {1, 1};
            CUDA_CHECK(cudaMemcpyAsync(h, d_flag, 8, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            lattice = (h[0] == 0 && h[1] == 0);
        }
        if (lattice) {
            set_lattice(ctx, m, ne);
            m->coords = dev_alloc<double>(3 * m->nNodes_l);
            CUDA_CHECK(cudaMemsetAsync(m->coords, 0, 8 * 3 * m->nNodes_l, ctx->stream));
            double *d_glob = dev_alloc<double>(3 * nNodes);
            CUDA_CHECK(cudaMemcpyAsync(d_glob, NodeList, 8 * 3 * nNodes, cudaMemcpyHostToDevice, ctx->stream));
            LAUNCH(ctx, k_copy_slab_coords, (unsigned)((m->nNodes_l + 255) / 256), 256, 0, m->lat, (const double *)d_glob,
                   m->coords);
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            dev_free(d_glob);
        } else {
            REQUIRE(ctx->nranks == 1, SMFEM_ERR_UNSUPPORTED, "unstructured meshes are single-GPU only");
            m->structured = false;
            m->nNodes_l = nNodes;
            m->coords = dev_alloc<double>((int64_t)ndim * nNodes);
            CUDA_CHECK(cudaMemcpyAsync(m->coords, NodeList, 8 * (int64_t)ndim * nNodes, cudaMemcpyHostToDevice, ctx->stream));
            m->ien = dev_alloc<int32_t>(nEl * nLocal);
            LAUNCH(ctx, k_convert_index, (unsigned)((nEl * nLocal + 255) / 256), 256, 0, nEl * nLocal, (const int64_t *)d_ien,
                   (int64_t)1, nNodes, m->ien, d_flag + 2);
            bool std_id = false;
            if (ID) {
                LAUNCH(ctx, k_check_std_id, (unsigned)((nNodes * nDof + 255) / 256), 256, 0, nNodes, nDof, (const int64_t *)d_id,
                       d_flag + 3);
                int h3 = 1;
                CUDA_CHECK(cudaMemcpyAsync(&h3, d_flag + 3, 4, cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
                std_id = (h3 == 0);  // standard node-major numbering: no ID table needed on the device (dof = nDof*node + comp)
            }
            if (ID && !std_id) {
                m->id = dev_alloc<int32_t>(nNodes * nDof);
                m->nDof_id = nDof;
                LAUNCH(ctx, k_convert_index, (unsigned)((nNodes * nDof + 255) / 256), 256, 0, nNodes * nDof,
                       (const int64_t *)d_id, (int64_t)1, (int64_t)INT32_MAX - 1, m->id, d_flag + 2);
                unsigned long long *d_max = (unsigned long long *)dev_alloc<int64_t>(1);
                CUDA_CHECK(cudaMemsetAsync(d_max, 0, 8, ctx->stream));
                LAUNCH(ctx, k_max_i64, (unsigned)((nNodes * nDof + 255) / 256), 256, 0, nNodes * nDof, (const int64_t *)d_id, d_max);
                CUDA_CHECK(cudaMemcpyAsync(&m->ndof_id, d_max, 8, cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
                dev_free(d_max);  // from dev_alloc: through the allocator (keeps its bookkeeping consistent)
            }
            if (ID && std_id) m->nDof_id = nDof;
            m->std_id = (ID == nullptr) || std_id;
            int h[4] = {0, 0, 0, 0};
            CUDA_CHECK(cudaMemcpyAsync(h, d_flag, 16, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            if (h[2]) {
                dev_free(d_ien);
                dev_free(d_id);
                dev_free(d_flag);
                smfem_mesh_free(m);
                throw SmfemError(SMFEM_ERR_INVALID, "IEN / ID entry out of range (reference: BoundsError)");
            }
        }
        dev_free(d_ien);
        dev_free(d_id);
        dev_free(d_flag);
        *out = m;
    });
}

int smfem_inflate_sphere(smfem_ctx *ctx, smfem_mesh *mesh, double x0, double x1, double y0, double y1) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh);
        mesh_inflate(ctx, mesh, x0, x1, y0, y1);
    });
}

int smfem_inflate_sphere_host(smfem_ctx *ctx, double *NodeList, int ndim, int64_t nNodes, double x0, double x1, double y0,
                              double y1) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(NodeList);
        REQUIRE((ndim == 2 || ndim == 3) && nNodes >= 1, SMFEM_ERR_INVALID, "inflate_sphere: bad sizes");
        smfem_mesh tmp;
        tmp.ndim = ndim;
        tmp.nNodes_l = nNodes;
        tmp.coords = dev_alloc<double>((int64_t)ndim * nNodes);
        try {
            CUDA_CHECK(cudaMemcpyAsync(tmp.coords, NodeList, 8 * (int64_t)ndim * nNodes, cudaMemcpyHostToDevice, ctx->stream));
            mesh_inflate(ctx, &tmp, x0, x1, y0, y1);
            CUDA_CHECK(cudaMemcpyAsync(NodeList, tmp.coords, 8 * (int64_t)ndim * nNodes, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        } catch (...) {
            dev_free(tmp.coords);
            throw;
        }
        dev_free(tmp.coords);
    });
}

int smfem_mesh_set_nodelist(smfem_ctx *ctx, smfem_mesh *mesh, const double *NodeList_global) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh);
        NOTNULL(NodeList_global);
        if (!mesh->structured) {
            CUDA_CHECK(cudaMemcpyAsync(mesh->coords, NodeList_global, 8 * (int64_t)mesh->ndim * mesh->nNodes_g,
                                       cudaMemcpyHostToDevice, ctx->stream));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            return;
        }
        // copy only the planes this rank holds
        const Lattice &L = mesh->lat;
        int lo = L.k0 - 1 < 0 ? 0 : L.k0 - 1, hi = L.k1 + 1 > L.n1 ? L.n1 : L.k1 + 1;
        int64_t first_l = (int64_t)(lo - (L.k0 - 1)) * L.plane();
        CUDA_CHECK(cudaMemcpyAsync(mesh->coords + 3 * first_l, NodeList_global + 3 * (int64_t)lo * L.plane(),
                                   8 * 3 * (int64_t)(hi - lo) * L.plane(), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    });
}

int smfem_mesh_info(smfem_mesh *mesh, int64_t *nNodes, int64_t *nEl, int *nLocal, int *ndim, int *structured,
                    int64_t *node0_owned, int64_t *nNodes_owned) {
    return guarded([&] {
        NOTNULL(mesh);
        if (nNodes) *nNodes = mesh->nNodes_g;
        if (nEl) *nEl = mesh->nEl_g;
        if (nLocal) *nLocal = mesh->nn;
        if (ndim) *ndim = mesh->ndim;
        if (structured) *structured = mesh->structured ? 1 : 0;
        if (node0_owned) *node0_owned = mesh->structured ? (int64_t)mesh->lat.k0 * mesh->lat.plane() : 0;
        if (nNodes_owned) *nNodes_owned = mesh->structured ? (int64_t)mesh->lat.nown() * mesh->lat.plane() : mesh->nNodes_g;
    });
}

int smfem_mesh_colors(smfem_ctx *ctx, smfem_mesh *mesh, int *ncolors, int64_t *color_sizes) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh);
        NOTNULL(ncolors);
        mesh_color_elements(ctx, mesh);
        *ncolors = mesh->ncolors;
        if (color_sizes)
            for (int c = 0; c < 64; ++c) color_sizes[c] = mesh->ncolors > 0 ? mesh->color_off[c + 1] - mesh->color_off[c] : 0;
    });
}

int smfem_mesh_export(smfem_ctx *ctx, smfem_mesh *mesh, double *NodeList_owned, int64_t *IEN, int64_t *ID, int64_t *IEN_top,
                      int64_t *IEN_btm) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh);
        if (NodeList_owned) {
            int64_t first = mesh->structured ? mesh->lat.plane() : 0;
            int64_t cnt = mesh->structured ? (int64_t)mesh->lat.nown() * mesh->lat.plane() : mesh->nNodes_g;
            CUDA_CHECK(cudaMemcpyAsync(NodeList_owned, mesh->coords + mesh->ndim * first, 8 * (int64_t)mesh->ndim * cnt,
                                       cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        }
        if (!(IEN || ID || IEN_top || IEN_btm)) return;
        REQUIRE(mesh->structured, SMFEM_ERR_UNSUPPORTED, "connectivity export is for meshgrid meshes (user meshes already own theirs)");
        // integer tables of examples/vector3D.jl:73-75, :94-113 (output formatting; tiny, host side)
        const int64_t ne = mesh->ne, n1 = ne + 1, nEl = ne * ne * ne, nN = n1 * n1 * n1, p = n1 * n1;
        if (ID)
            for (int64_t m = 0; m < nN; ++m)
                for (int l = 0; l < 3; ++l) ID[m + l * nN] = 3 * m + l + 1;
        int64_t nb = 0, nt = 0;
        for (int64_t k = 1; k <= ne; ++k)
            for (int64_t j = 1; j <= ne; ++j)
                for (int64_t i = 1; i <= ne; ++i) {
                    int64_t n = ((k - 1) * ne + (j - 1)) * ne + (i - 1);
                    int64_t v[8] = {(k - 1) * p + (j - 1) * n1 + i, (k - 1) * p + (j - 1) * n1 + i + 1,
                                    (k - 1) * p + j * n1 + i + 1,   (k - 1) * p + j * n1 + i,
                                    k * p + (j - 1) * n1 + i,       k * p + (j - 1) * n1 + i + 1,
                                    k * p + j * n1 + i + 1,         k * p + j * n1 + i};
                    if (IEN)
                        for (int a = 0; a < 8; ++a) IEN[n + a * nEl] = v[a];
                    if (k == 1) {
                        if (IEN_btm)
                            for (int a = 0; a < 4; ++a) IEN_btm[nb + a * ne * ne] = v[a];
                        ++nb;
                    } else if (k == ne) {
                        if (IEN_top)
                            for (int a = 0; a < 4; ++a) IEN_top[nt + a * ne * ne] = v[4 + a];
                        ++nt;
                    }
                }
        if (IEN_top && ne == 1) std::memset(IEN_top, 0, sizeof(int64_t) * 4);  // reference quirk: elseif never taken
    });
}

int smfem_mesh_free(smfem_mesh *mesh) {
    return guarded([&] {
        if (!mesh) return;
        dev_free(mesh->coords);
        dev_free(mesh->ien);
        dev_free(mesh->id);
        dev_free(mesh->elist);
        dev_free(mesh->g_ptr);
        dev_free(mesh->g_ent);
        dev_free(mesh->g_task_node);
        delete mesh;
    });
}

static smfem_matrix *new_matrix(smfem_ctx *ctx, smfem_mesh *mesh, int ndim, int nDof) {
    REQUIRE(ndim == mesh->ndim, SMFEM_ERR_INVALID, "ndim does not match the mesh (reference: DimensionMismatch)");
    REQUIRE((ndim == 3 && (nDof == 3 || nDof == 1)) || (ndim == 2 && (nDof == 2 || nDof == 1)), SMFEM_ERR_UNSUPPORTED,
            "supported (ndim,nDof): (3,3) (3,1) (2,2) (2,1)");
    REQUIRE(mesh->nn != 9 || nDof == 1, SMFEM_ERR_UNSUPPORTED, "9-node (Q2) elements: scalar problems only, as upstream");
    REQUIRE(mesh->structured || mesh->id != nullptr || mesh->nDof_id == nDof || nDof == 1, SMFEM_ERR_INVALID,
            "ID is required when nDof > 1");
    REQUIRE(mesh->structured || mesh->nDof_id == 0 || mesh->nDof_id == nDof || nDof == 1, SMFEM_ERR_INVALID,
            "size(ID,2) must equal nDof (src/fem.jl:238-242 is only consistent then)");
    smfem_matrix *K = new smfem_matrix();
    K->ctx = ctx;
    K->ndim = ndim;
    K->nDof = nDof;
    K->nn = mesh->nn;
    return K;
}

int smfem_pattern_build(smfem_ctx *ctx, smfem_mesh *mesh, int ndim, int nDof, smfem_matrix **K_out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh);
        NOTNULL(K_out);
        CUDA_CHECK(cudaSetDevice(ctx->device));
        smfem_matrix *K = new_matrix(ctx, mesh, ndim, nDof);
        try {
            if (mesh->structured) pattern_build_structured(ctx, mesh, K);
            else {
                // nDof == 1 uses raw node ids even if an ID was supplied (src/fem.jl:204-205)
                int32_t *saved = mesh->id;
                if (nDof == 1) mesh->id = nullptr;
                try {
                    pattern_build_general(ctx, mesh, K);
                } catch (...) {
                    mesh->id = saved;
                    throw;
                }
                mesh->id = saved;
            }
        } catch (...) {
            smfem_matrix_free(K);
            throw;
        }
        *K_out = K;
    });
}

int smfem_pattern_rebuild(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh);
        NOTNULL(K);
        REQUIRE(!K->csr_less, SMFEM_ERR_UNSUPPORTED, "this handle is a matrix-free operator without CSR arrays (smfem_matfree_operator)");
        REQUIRE(mesh->structured && K->structured, SMFEM_ERR_UNSUPPORTED, "pattern_rebuild: structured meshes only");
        pattern_build_structured(ctx, mesh, K);  // buffers exist: kernels only, no allocation, no host sync
    });
}

int smfem_reassemble(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, double Young, double nu) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh);
        NOTNULL(K);
        REQUIRE(!K->csr_less, SMFEM_ERR_UNSUPPORTED, "this handle is a matrix-free operator without CSR arrays (smfem_matfree_operator)");
        REQUIRE(mesh->structured && K->structured, SMFEM_ERR_UNSUPPORTED, "reassemble: structured meshes only");
        if (const char *dbg = std::getenv("SMFEM_DEBUG_CLEAR"); dbg && dbg[0] == '1') {  // tests: prove every entry is rewritten
            CUDA_CHECK(cudaMemsetAsync(K->colind, 0xFF, sizeof(int32_t) * K->nnz_l, ctx->stream));
            CUDA_CHECK(cudaMemsetAsync(K->rowptr, 0xFF, sizeof(int64_t) * (K->nrows_l + 1), ctx->stream));
            if (K->val) CUDA_CHECK(cudaMemsetAsync(K->val, 0xFF, sizeof(double) * K->nnz_l, ctx->stream));
        }
        values_assemble(ctx, mesh, K, Young, nu, /*fuse_pattern=*/true);
    });
}

int smfem_assemble_values(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, double Young, double nu) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh);
        NOTNULL(K);
        REQUIRE(!K->csr_less, SMFEM_ERR_UNSUPPORTED, "this handle is a matrix-free operator without CSR arrays (smfem_matfree_operator)");
        int32_t *saved = mesh->id;
        if (K->nDof == 1) mesh->id = nullptr;
        try {
            values_assemble(ctx, mesh, K, Young, nu);
        } catch (...) {
            mesh->id = saved;
            throw;
        }
        mesh->id = saved;
    });
}

int smfem_assemble(smfem_ctx *ctx, smfem_mesh *mesh, int64_t ne, int ndim, int func_class, int nDof, double Young, double nu,
                   smfem_matrix **K_out) {
    int rc = guarded([&] {
        NOTNULL(mesh);
        REQUIRE(func_class == SMFEM_Q1 || func_class == SMFEM_Q2, SMFEM_ERR_INVALID, "unknown FunctionClass");
        if (func_class == SMFEM_Q2)  // upstream Q2 exists in 2-D only and its nDof > 1 sizing assumes 2^ndim nodes (src/fem.jl:77-111, :143-145)
            REQUIRE(ndim == 2 && nDof == 1 && mesh->nn == 9, SMFEM_ERR_UNSUPPORTED,
                    "FunctionClass \"Q2\": only the 2-D scalar case with 9-node elements is executable upstream");
        else
            REQUIRE(mesh->nn == (1 << ndim), SMFEM_ERR_INVALID, "FunctionClass \"Q1\" needs 2^ndim nodes per element");
        // the reference loops 1:ne^ndim (src/fem.jl:179) and ignores size(IEN,1)
        int64_t want = 1;
        for (int d = 0; d < ndim; ++d) want *= ne;
        REQUIRE(want == mesh->nEl_g, SMFEM_ERR_INVALID, "ne^ndim must equal the number of elements of the mesh");
    });
    if (rc) return rc;
    if (ctx && K_out && mesh->structured && ndim == 3 && nDof == 3 && values_tile_enabled())
        // hex lattice: sizes in closed form, rowptr kernel + ONE tile kernel that writes column indices and values
        return guarded([&] {
            CUDA_CHECK(cudaSetDevice(ctx->device));
            smfem_matrix *K = new_matrix(ctx, mesh, ndim, nDof);
            try {
                pattern_prepare_structured(ctx, mesh, K);
                values_assemble(ctx, mesh, K, Young, nu, /*fuse_pattern=*/true);
            } catch (...) {
                smfem_matrix_free(K);
                throw;
            }
            *K_out = K;
        });
    rc = smfem_pattern_build(ctx, mesh, ndim, nDof, K_out);
    if (rc) return rc;
    rc = smfem_assemble_values(ctx, mesh, *K_out, Young, nu);
    if (rc) {
        smfem_matrix_free(*K_out);
        *K_out = nullptr;
    }
    return rc;
}

// assemble_system(ne, NodeList, IEN, ndim, FunctionClass, nDof, ID, Young, nu) with HOST arrays in ONE call (src/fem.jl:135).
// For a hex-lattice candidate (sizes match meshgrid's) the work is overlapped: the copy stream moves NodeList in pieces
// and the tile kernel on the main stream follows the arriving coordinate planes (speculative assembly), while IEN / ID (4.7x the bytes of NodeList) are checked
// against the lattice numbering - chunks from the front through PCIe + a check kernel, chunks from the back by host
// threads where they lie (lattice_check.cu).  The call returns when the check has passed (all host arrays have been read by
// then); the assembly may still be in flight on the context's stream, like after any other call.  If the check fails, the
// speculative result is discarded and the general path runs.
int smfem_assemble_system(smfem_ctx *ctx, const double *NodeList, const int64_t *IEN, const int64_t *ID, int64_t nNodes, int64_t nEl,
                          int nLocal, int64_t ne, int ndim, int func_class, int nDof, double Young, double nu, smfem_mesh **mesh_out,
                          smfem_matrix **K_out) {
    bool speculated = false, lattice_ok = false;
    int rc = guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh_out);
        NOTNULL(K_out);
        NOTNULL(NodeList);
        NOTNULL(IEN);
        *mesh_out = nullptr;
        *K_out = nullptr;
        const bool candidate = ndim == 3 && nDof == 3 && nLocal == 8 && func_class == SMFEM_Q1 && ID != nullptr && ne >= 1 &&
                               ne < 2000 && nEl == ne * ne * ne && nNodes == (ne + 1) * (ne + 1) * (ne + 1) && values_tile_enabled();
        if (!candidate) return;
        speculated = true;
        CUDA_CHECK(cudaSetDevice(ctx->device));
        smfem_mesh *m = new smfem_mesh();
        smfem_matrix *K = nullptr;
        int64_t *d_stage = nullptr;
        int *d_flag = nullptr;  // [0] mismatch seen by the device-side check, [1] number of coordinate planes that have arrived
        try {
            m->ctx = ctx;
            set_lattice(ctx, m, ne);
            const Lattice &L = m->lat;
            d_stage = dev_alloc<int64_t>(2 * LATTICE_CHUNK);
            d_flag = dev_alloc<int>(4);
            m->coords = dev_alloc<double>(3 * m->nNodes_l);
            cudaStream_t cs = ctx->copy_stream;
            // the copy stream starts after everything already queued on the main stream (buffers come from a cache)
            CUDA_CHECK(cudaEventRecord(ctx->ev_fork, ctx->stream));
            CUDA_CHECK(cudaStreamWaitEvent(cs, ctx->ev_fork, 0));
            CUDA_CHECK(cudaMemsetAsync(m->coords, 0, 8 * 3 * m->nNodes_l, cs));  // ghost planes outside the domain
            CUDA_CHECK(cudaMemsetAsync(d_flag, 0, 16, cs));
            CUDA_CHECK(cudaEventRecord(ctx->ev_nodes, cs));
            // NodeList: the planes this rank holds (k0-1 .. k1), in a few pieces with a watermark behind each
            // (queued BEFORE the kernel that waits for them is launched: a failing copy must not leave a spinning kernel behind)
            {
                const int lo = L.k0 - 1 < 0 ? 0 : L.k0 - 1, hi = L.k1 + 1 > L.n1 ? L.n1 : L.k1 + 1;
                const int parts = std::min(8, hi - lo);
                for (int sidx = 0; sidx < parts; ++sidx) {
                    const int p0 = lo + (int)((int64_t)(hi - lo) * sidx / parts), p1 = lo + (int)((int64_t)(hi - lo) * (sidx + 1) / parts);
                    CUDA_CHECK(cudaMemcpyAsync(m->coords + 3 * (int64_t)(p0 - (L.k0 - 1)) * L.plane(), NodeList + 3 * (int64_t)p0 * L.plane(),
                                               8 * 3 * (int64_t)(p1 - p0) * L.plane(), cudaMemcpyHostToDevice, cs));
                    ctx->h_flags[8 + sidx] = p1;
                    CUDA_CHECK(cudaMemcpyAsync(d_flag + 1, ctx->h_flags + 8 + sidx, 4, cudaMemcpyHostToDevice, cs));
                }
                ctx->h2d_bytes += 8 * 3 * (int64_t)(hi - lo) * L.plane();
            }
            // main stream: rowptr + fused tile kernel, which follows the arrival of the coordinate planes (d_flag[1])
            CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_nodes, 0));
            K = new_matrix(ctx, m, 3, 3);
            pattern_prepare_structured(ctx, m, K);
            values_assemble(ctx, m, K, Young, nu, /*fuse_pattern=*/true, d_flag + 1);
            // IEN / ID against the lattice numbering: PCIe + check kernel from the front, host threads from the back
            const bool host_ok = lattice_check_hybrid(ctx, L, IEN, ID, nEl, nNodes, (int)ne, d_stage, d_flag);
            ctx->h_flags[0] = 1;
            CUDA_CHECK(cudaMemcpyAsync(ctx->h_flags, d_flag, 4, cudaMemcpyDeviceToHost, cs));
            CUDA_CHECK(cudaEventRecord(ctx->ev_check, cs));
            CUDA_CHECK(cudaEventSynchronize(ctx->ev_check));
            lattice_ok = host_ok && ctx->h_flags[0] == 0;
            // the tile kernel may still read d_flag[1]: order the reuse of the block behind the work queued so far
            CUDA_CHECK(cudaEventRecord(ctx->ev_fork, ctx->stream));
            CUDA_CHECK(cudaStreamWaitEvent(cs, ctx->ev_fork, 0));
            if (!lattice_ok) CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        } catch (...) {
            cudaStreamSynchronize(ctx->copy_stream);
            cudaStreamSynchronize(ctx->stream);
            dev_free(d_stage);
            dev_free(d_flag);
            if (K) smfem_matrix_free(K);
            smfem_mesh_free(m);
            throw;
        }
        dev_free(d_stage);
        dev_free(d_flag);
        if (lattice_ok) {
            *mesh_out = m;
            *K_out = K;
        } else {
            smfem_matrix_free(K);
            smfem_mesh_free(m);
        }
    });
    if (rc) return rc;
    if (speculated && lattice_ok) return SMFEM_OK;
    // not a lattice (or not the fast element type): the two-step path
    rc = smfem_mesh_from_host(ctx, NodeList, IEN, ID, nNodes, nEl, nLocal, ndim, nDof, ne, mesh_out);
    if (rc) return rc;
    rc = smfem_assemble(ctx, *mesh_out, ne, ndim, func_class, nDof, Young, nu, K_out);
    if (rc) {
        smfem_mesh_free(*mesh_out);
        *mesh_out = nullptr;
    }
    return rc;
}

int smfem_matrix_info(smfem_matrix *K, int64_t *m, int64_t *n, int64_t *nnz, int64_t *row0, int64_t *nrows_local,
                      int64_t *nnz_local) {
    return guarded([&] {
        NOTNULL(K);
        if (m) *m = K->m_g;
        if (n) *n = K->m_g;
        if (nnz) *nnz = K->nnz_g;
        if (row0) *row0 = K->row0;
        if (nrows_local) *nrows_local = K->nrows_l;
        if (nnz_local) *nnz_local = K->nnz_l;
    });
}

int smfem_matrix_export_csc(smfem_ctx *ctx, smfem_matrix *K, int which, int64_t *colptr, int64_t *rowval, double *nzval) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        REQUIRE(!K->csr_less, SMFEM_ERR_UNSUPPORTED, "this handle is a matrix-free operator without CSR arrays (smfem_matfree_operator)");
        export_csc(ctx, K, which, colptr, rowval, nzval);
    });
}

int smfem_matrix_diag(smfem_ctx *ctx, smfem_matrix *K, double *diag_local) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        NOTNULL(diag_local);
        REQUIRE(K->diag != nullptr, SMFEM_ERR_INVALID, "matrix has no values yet");
        CUDA_CHECK(cudaMemcpyAsync(diag_local, K->diag, 8 * K->nrows_l, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        ctx->d2h_bytes += 8 * K->nrows_l;
    });
}

int smfem_assembly_kernel_ms(smfem_ctx *ctx, int last_n, float *avg_ms, int *n_used) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(avg_ms);
        int n = (int)std::min<int64_t>(ctx->asm_count, smfem_ctx::ASM_RING);
        if (last_n > 0 && last_n < n) n = last_n;
        double sum = 0;
        for (int i = 0; i < n; ++i) {
            const int slot = (int)((ctx->asm_count - 1 - i) % smfem_ctx::ASM_RING);
            CUDA_CHECK(cudaEventSynchronize(ctx->asm_ev[2 * slot + 1]));
            float ms = 0;
            CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->asm_ev[2 * slot], ctx->asm_ev[2 * slot + 1]));
            sum += ms;
        }
        *avg_ms = n ? (float)(sum / n) : 0.f;
        if (n_used) *n_used = n;
    });
}

int smfem_transfer_bytes(smfem_ctx *ctx, int64_t *h2d, int64_t *d2h) {
    return guarded([&] {
        NOTNULL(ctx);
        if (h2d) *h2d = ctx->h2d_bytes;
        if (d2h) *d2h = ctx->d2h_bytes;
    });
}

// K_bar = K + beta*b must not alias K (examples/vector3D.jl:308 keeps K): a device-side copy of pattern, values and diagonal;
// Dirichlet data, solver workspace, peer window and multigrid hierarchy are per matrix and start empty in the copy
int smfem_matrix_clone(smfem_ctx *ctx, smfem_matrix *K, smfem_matrix **out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        NOTNULL(out);
        REQUIRE(K->csr_less || (K->rowptr && K->colind), SMFEM_ERR_INVALID, "matrix has no pattern yet");
        smfem_matrix *C = new smfem_matrix();
        C->csr_less = K->csr_less;
        C->matfree_on = K->matfree_on;
        C->mf_mesh = K->mf_mesh;
        C->ctx = ctx;
        C->ndim = K->ndim;
        C->nDof = K->nDof;
        C->nn = K->nn;
        C->structured = K->structured;
        C->lat = K->lat;
        C->m_g = K->m_g;
        C->nnz_g = K->nnz_g;
        C->row0 = K->row0;
        C->nrows_l = K->nrows_l;
        C->ncols_l = K->ncols_l;
        C->ghost_cols = K->ghost_cols;
        C->nnz_l = K->nnz_l;
        C->values_ready = K->values_ready;
        C->Young = K->Young;
        C->nu = K->nu;
        C->beta_total = K->beta_total;
        C->mat_known = K->mat_known;
        C->spmv_variant = K->spmv_variant;
        auto dup = [&](auto *&dst, const auto *src, int64_t n) {
            using T = std::remove_const_t<std::remove_pointer_t<decltype(src)>>;
            if (!src) return;
            dst = dev_alloc<T>(n);
            CUDA_CHECK(cudaMemcpyAsync(dst, src, sizeof(T) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
        };
        if (!K->csr_less) {
            dup(C->rowptr, (const int64_t *)K->rowptr, K->nrows_l + 1 + 8);
            dup(C->colind, (const int32_t *)K->colind, K->nnz_l + 16);
            dup(C->val, (const double *)K->val, K->nnz_l + 16);
        }
        dup(C->bval, (const double *)K->bval, K->nnz_l);
        dup(C->diag, (const double *)K->diag, K->nrows_l);
        *out = C;
    });
}

int smfem_matrix_free(smfem_matrix *K) {
    return guarded([&] {
        if (!K) return;
        gmg_free(K);
        solver_free(K);
        dev_free(K->rowptr);
        dev_free(K->colind);
        dev_free(K->val);
        dev_free(K->bval);
        dev_free(K->diag);
        delete K;
    });
}

int smfem_surface_mass(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, const int64_t *IEN_top, const int64_t *IEN_btm,
                       int64_t nFaces, double beta, int keep_b) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        NOTNULL(mesh);
        REQUIRE(K->values_ready, SMFEM_ERR_INVALID, "assemble K before adding the surface term");
        if (K->csr_less) {  // matrix-free operator: the term lives in the operator (beta) and in the diagonal
            REQUIRE(!IEN_top && !IEN_btm && !keep_b && mesh->structured, SMFEM_ERR_UNSUPPORTED,
                    "matrix-free operator: the surface term is the lattice's top / bottom faces (no explicit face lists, no stored b)");
            matfree_add_surface(ctx, K, mesh, beta);
            return;
        }
        int32_t *faces = nullptr;
        int64_t nf = 0;
        if (IEN_top || IEN_btm) {
            REQUIRE(IEN_top && IEN_btm && nFaces >= 1, SMFEM_ERR_INVALID, "pass both IEN_top and IEN_btm (nFaces x 4)");
            REQUIRE(ctx->nranks == 1, SMFEM_ERR_UNSUPPORTED, "explicit face lists are single-GPU only");
            // bottom faces first, then top (examples/vector3D.jl:245-246); [4][2*nFaces], 0-based local ids
            nf = 2 * nFaces;
            std::vector<int32_t> h(4 * nf);
            int64_t ghost_nodes = mesh->structured ? mesh->lat.plane() : 0;
            for (int a = 0; a < 4; ++a)
                for (int64_t f = 0; f < nFaces; ++f) {
                    int64_t vb = IEN_btm[f + a * nFaces], vt = IEN_top[f + a * nFaces];
                    REQUIRE(vb >= 1 && vb <= mesh->nNodes_g && vt >= 1 && vt <= mesh->nNodes_g, SMFEM_ERR_INVALID,
                            "IEN_top / IEN_btm entry out of range (reference: BoundsError)");
                    h[(int64_t)a * nf + f] = (int32_t)(vb - 1 + ghost_nodes);
                    h[(int64_t)a * nf + nFaces + f] = (int32_t)(vt - 1 + ghost_nodes);
                }
            faces = dev_alloc<int32_t>(4 * nf);
            CUDA_CHECK(cudaMemcpyAsync(faces, h.data(), 4 * 4 * nf, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        } else {
            REQUIRE(mesh->structured, SMFEM_ERR_INVALID, "user meshes need explicit IEN_top / IEN_btm");
            const Lattice &L = mesh->lat;
            int has_btm = (L.k0 == 0), has_top = (L.k1 == L.n1);
            nf = (int64_t)(has_btm + has_top) * L.ne * L.ne;
            if (nf > 0) {
                faces = dev_alloc<int32_t>(4 * nf);
                LAUNCH(ctx, k_struct_faces, (unsigned)((nf + 255) / 256), 256, 0, L, has_btm, has_top, faces, nf);
            }
        }
        try {
            surface_mass(ctx, K, mesh, faces, nf, beta, keep_b != 0);
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        } catch (...) {
            dev_free(faces);
            throw;
        }
        dev_free(faces);
    });
}

int smfem_set_dirichlet_zplanes(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, double d) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        NOTNULL(mesh);
        dirichlet_zplanes(ctx, K, mesh, d);
    });
}

int smfem_set_dirichlet(smfem_ctx *ctx, smfem_matrix *K, const int64_t *dofs, const double *values, int64_t n) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        REQUIRE(n == 0 || (dofs && values), SMFEM_ERR_INVALID, "dofs / values must not be NULL");
        dirichlet_list(ctx, K, dofs, values, n);
    });
}

int smfem_pcg_solve(smfem_ctx *ctx, smfem_matrix *K, double rtol, int maxit, const double *rhs_extra, double *q_out,
                    int *iters_out, double *relres_out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        REQUIRE(rtol > 0 && maxit >= 0, SMFEM_ERR_INVALID, "need rtol > 0, maxit >= 0");
        if (K->gmg_on) gmg_pcg_solve(ctx, K, rtol, maxit, rhs_extra, q_out, iters_out, relres_out);
        else pcg_solve(ctx, K, rtol, maxit, rhs_extra, q_out, iters_out, relres_out);
    });
}

int smfem_pcg_use_multigrid(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, int enable) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        if (enable) NOTNULL(mesh);
        gmg_enable(ctx, K, mesh, enable != 0);
    });
}

int smfem_pcg_apply_preconditioner(smfem_ctx *ctx, smfem_matrix *K, const double *r, double *z) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        NOTNULL(r);
        NOTNULL(z);
        gmg_apply_host(ctx, K, r, z);
    });
}

int smfem_project_nodes(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, const int64_t *node_ids, int64_t n,
                        const double *CameraMatrix, double *nodes3d_out, double *nodes2d_out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh);
        project_nodes(ctx, mesh, K, node_ids, n, CameraMatrix, nodes3d_out, nodes2d_out);
    });
}

int smfem_extract_borders(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, const int64_t *side_node_ids, int64_t n,
                          const double *CameraMatrix, int state, int64_t ne, double *BorderPoints_out, int64_t capacity,
                          int64_t *nBorder_out, double *SideNodes2D_out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh);
        extract_borders(ctx, mesh, K, side_node_ids, n, CameraMatrix, state, ne, BorderPoints_out, capacity, nBorder_out, SideNodes2D_out);
    });
}

int smfem_pcg_set_warm_start(smfem_matrix *K, double scale) {
    return guarded([&] {
        NOTNULL(K);
        REQUIRE(K->x != nullptr || scale == 0.0, SMFEM_ERR_INVALID, "warm start needs a previous solve on this matrix");
        K->warm_scale = scale;
    });
}

int smfem_spmv_host(smfem_ctx *ctx, smfem_matrix *K, const double *x, double *y) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        NOTNULL(x);
        NOTNULL(y);
        spmv_host(ctx, K, x, y);
    });
}

int smfem_bench_spmv(smfem_ctx *ctx, smfem_matrix *K, int variant, int reps, float *ms_per_spmv) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        NOTNULL(ms_per_spmv);
        bench_spmv(ctx, K, variant, reps, ms_per_spmv);
    });
}

// Jacobi-PCG of the last solve: microseconds CTA 0 spent waiting for the neighbours' halo flags and for the two all-reduces
int smfem_pcg_wait_stats(smfem_matrix *K, double *halo_us, double *allreduce_rz_us, double *allreduce_pap_us) {
    return guarded([&] {
        NOTNULL(K);
        if (halo_us) *halo_us = K->last_wait_us[0];
        if (allreduce_rz_us) *allreduce_rz_us = K->last_wait_us[1];
        if (allreduce_pap_us) *allreduce_pap_us = K->last_wait_us[2];
    });
}

int smfem_set_spmv_variant(smfem_matrix *K, int variant) {
    return guarded([&] {
        NOTNULL(K);
        REQUIRE(variant >= 0 && variant <= 4, SMFEM_ERR_INVALID, "unknown SpMV variant");
        K->spmv_variant = variant;
    });
}

int smfem_pcg_stats(smfem_matrix *K, float *ms_total, float *ms_spmv_est, int *iters) {
    return guarded([&] {
        NOTNULL(K);
        if (ms_total) *ms_total = K->last_ms;
        if (ms_spmv_est) *ms_spmv_est = K->last_ms_spmv;
        if (iters) *iters = K->last_iters;
    });
}

int smfem_comm_export(smfem_ctx *ctx, smfem_matrix *K, void *handle_out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        NOTNULL(handle_out);
        comm_export(ctx, K, handle_out);
    });
}

// The operator of assemble_system (src/fem.jl:135-256) for the hex lattice WITHOUT the assembled matrix: a handle that carries
// the slab layout, the material, diag(K) (computed from the coordinates, for the Jacobi preconditioner) and nothing else.
int smfem_matfree_operator(smfem_ctx *ctx, smfem_mesh *mesh, double Young, double nu, smfem_matrix **K_out) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(mesh);
        NOTNULL(K_out);
        CUDA_CHECK(cudaSetDevice(ctx->device));
        REQUIRE(mesh->structured && mesh->ndim == 3, SMFEM_ERR_UNSUPPORTED, "matrix-free operator: structured 3-D hex lattice only");
        smfem_matrix *K = new_matrix(ctx, mesh, 3, 3);
        try {
            const Lattice &L = mesh->lat;
            K->structured = true;
            K->csr_less = true;
            K->lat = L;
            K->m_g = 3 * mesh->nNodes_g;
            const int64_t S1 = 3 * (int64_t)L.n1 - 2;
            K->nnz_g = 9 * S1 * S1 * S1;  // what the assembled matrix would hold (reported by smfem_matrix_info; nnz_local = 0)
            K->nnz_l = 0;
            K->ghost_cols = L.plane() * 3;
            K->nrows_l = (int64_t)L.nown() * L.plane() * 3;
            K->ncols_l = K->nrows_l + 2 * K->ghost_cols;
            K->row0 = (int64_t)L.k0 * L.plane() * 3;
            REQUIRE(K->ncols_l < (int64_t)INT32_MAX, SMFEM_ERR_UNSUPPORTED, "local dof count exceeds int32 column indices");
            K->Young = Young;
            K->nu = nu;
            K->beta_total = 0.0;
            K->mat_known = true;
            matfree_diag(ctx, K, mesh);
            K->values_ready = true;
            K->mf_mesh = mesh;
            K->matfree_on = true;
        } catch (...) {
            smfem_matrix_free(K);
            throw;
        }
        *K_out = K;
    });
}

// SURVEY 8(f) row 3, second half: the solve's operator applied matrix-free (matfree.cu)
int smfem_pcg_use_matrix_free(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, int enable) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        REQUIRE(enable || !K->csr_less, SMFEM_ERR_INVALID, "this operator has no assembled matrix to fall back to (smfem_matfree_operator)");
        if (enable) {
            NOTNULL(mesh);
            REQUIRE(mesh->structured && K->structured && K->ndim == 3 && K->nDof == 3, SMFEM_ERR_UNSUPPORTED,
                    "matrix-free operator: structured 3-D hex lattice with nDof = 3 only");
            REQUIRE(mesh->lat.n1 == K->lat.n1 && mesh->lat.k0 == K->lat.k0 && mesh->lat.k1 == K->lat.k1, SMFEM_ERR_INVALID,
                    "matrix-free operator: the mesh is not the lattice K was assembled on");
            REQUIRE(K->mat_known, SMFEM_ERR_INVALID, "matrix-free operator: assemble K first (material, diagonal for the preconditioner)");
            K->mf_mesh = mesh;
        }
        K->matfree_on = enable != 0;
    });
}

int smfem_comm_prepare(smfem_ctx *ctx, smfem_matrix *K) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        solver_alloc(ctx, K);
    });
}

int smfem_comm_connect_local(smfem_ctx *ctx, smfem_matrix *K, smfem_matrix *const *all_K, int n) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        NOTNULL(all_K);
        comm_connect_local(ctx, K, all_K, n);
    });
}

int smfem_comm_connect(smfem_ctx *ctx, smfem_matrix *K, const void *all_handles) {
    return guarded([&] {
        NOTNULL(ctx);
        NOTNULL(K);
        NOTNULL(all_handles);
        comm_connect(ctx, K, all_handles);
    });
}

}  // extern "C"
