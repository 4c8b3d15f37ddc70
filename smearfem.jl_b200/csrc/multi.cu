// One host process, several GPUs.  The reference's host is ONE Julia process (examples/vector3D.jl:266-345 `main()`), so the
// drop-in boundary must be drivable from a single thread of a single process: smfem_multi owns one context per GPU (rank r of
// n, z-slab r of the lattice) and one worker thread per GPU; every smfem_multi_* entry point runs the per-rank C-ABI call on
// all workers concurrently and returns when all are done.  The peer windows (halo planes, all-reduce mailboxes, multigrid
// exchange region) are connected with cudaDeviceEnablePeerAccess + raw device pointers: CUDA IPC handles cannot be opened
// inside the exporting process.  Kernels, protocols and results are exactly those of the one-process-per-GPU mode.
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "smfem_internal.cuh"

struct smfem_multi {
    int n = 0;
    std::vector<smfem_ctx *> ctx;
    std::vector<int> device;
    // one persistent worker per rank (a CUDA device is bound per host thread; collective calls block in all ranks at once)
    struct Worker {
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        std::function<int()> job;
        bool has_job = false, done = false, quit = false;
        int status = SMFEM_OK;
        std::string err;
    };
    std::vector<Worker *> w;
};
struct smfem_multi_mesh {
    std::vector<smfem_mesh *> h;
};
struct smfem_multi_matrix {
    std::vector<smfem_matrix *> h;
    bool connected = false;
};

namespace {

void worker_main(smfem_multi::Worker *w, int device) {
    cudaSetDevice(device);
    for (;;) {
        std::function<int()> job;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->has_job || w->quit; });
            if (w->quit) return;
            job = std::move(w->job);
            w->has_job = false;
        }
        int st = job();
        std::string err = st == SMFEM_OK ? std::string() : std::string(smfem_last_error());
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->status = st;
            w->err = std::move(err);
            w->done = true;
        }
        w->cv.notify_all();
    }
}

// f(rank) on every worker at once; first failing rank's status + message are reported on the calling thread
int run_all(smfem_multi *m, const std::function<int(int)> &f) {
    for (int r = 0; r < m->n; ++r) {
        smfem_multi::Worker *w = m->w[r];
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->job = [f, r] { return f(r); };
            w->has_job = true;
            w->done = false;
        }
        w->cv.notify_all();
    }
    int status = SMFEM_OK;
    for (int r = 0; r < m->n; ++r) {
        smfem_multi::Worker *w = m->w[r];
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->status != SMFEM_OK && status == SMFEM_OK) {
            status = w->status;
            smfem_set_last_error(("rank " + std::to_string(r) + ": " + w->err).c_str());
        }
    }
    return status;
}

int fail(int code, const char *msg) {
    smfem_set_last_error(msg);
    return code;
}

}  // namespace

extern "C" {

int smfem_init_multi(int n_gpus, const int *devices, smfem_multi **out) {
    if (!out) return fail(SMFEM_ERR_INVALID, "out must not be NULL");
    if (n_gpus < 1 || n_gpus > SMFEM_MAX_RANKS) return fail(SMFEM_ERR_INVALID, "smfem_init_multi: need 1 <= n_gpus <= 8");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(SMFEM_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    smfem_multi *m = new smfem_multi();
    m->n = n_gpus;
    m->ctx.assign(n_gpus, nullptr);
    for (int r = 0; r < n_gpus; ++r) {
        const int d = devices ? devices[r] : r;
        if (d < 0 || d >= count) {
            delete m;
            return fail(SMFEM_ERR_INVALID, "smfem_init_multi: bad device ordinal (fewer GPUs than ranks?)");
        }
        m->device.push_back(d);
    }
    for (int r = 0; r < n_gpus; ++r) {
        auto *w = new smfem_multi::Worker();
        w->th = std::thread(worker_main, w, m->device[r]);
        m->w.push_back(w);
    }
    int st = run_all(m, [m](int r) {
        int s = smfem_init(m->device[r], r, m->n, &m->ctx[r]);
        if (s != SMFEM_OK) return s;
        for (int q = 0; q < m->n; ++q) {  // peer access in both directions (every rank runs this loop for its own device)
            if (m->device[q] == m->device[r]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->device[r], m->device[q]);
            if (!can) {
                smfem_set_last_error("smfem_init_multi: GPUs without peer access (NVLink / PCIe P2P) cannot share a solve");
                return (int)SMFEM_ERR_UNSUPPORTED;
            }
            cudaError_t pe = cudaDeviceEnablePeerAccess(m->device[q], 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
                smfem_set_last_error(cudaGetErrorString(pe));
                return (int)SMFEM_ERR_CUDA;
            }
            cudaGetLastError();
        }
        return (int)SMFEM_OK;
    });
    if (st != SMFEM_OK) {
        smfem_multi_destroy(m);
        return st;
    }
    *out = m;
    return SMFEM_OK;
}

int smfem_multi_destroy(smfem_multi *m) {
    if (!m) return SMFEM_OK;
    run_all(m, [m](int r) {
        int s = m->ctx[r] ? smfem_destroy(m->ctx[r]) : SMFEM_OK;
        m->ctx[r] = nullptr;
        return s;
    });
    for (auto *w : m->w) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
        }
        w->cv.notify_all();
        w->th.join();
        delete w;
    }
    delete m;
    return SMFEM_OK;
}

int smfem_multi_size(smfem_multi *m, int *n_gpus) {
    if (!m || !n_gpus) return fail(SMFEM_ERR_INVALID, "NULL argument");
    *n_gpus = m->n;
    return SMFEM_OK;
}

int smfem_multi_rank_handles(smfem_multi *m, smfem_multi_mesh *mesh, smfem_multi_matrix *K, int rank, smfem_ctx **ctx_out,
                             smfem_mesh **mesh_out, smfem_matrix **K_out) {
    if (!m || rank < 0 || rank >= m->n) return fail(SMFEM_ERR_INVALID, "bad rank");
    if (ctx_out) *ctx_out = m->ctx[rank];
    if (mesh_out) *mesh_out = mesh ? mesh->h[rank] : nullptr;
    if (K_out) *K_out = K ? K->h[rank] : nullptr;
    return SMFEM_OK;
}

int smfem_multi_sync(smfem_multi *m) {
    if (!m) return fail(SMFEM_ERR_INVALID, "NULL argument");
    return run_all(m, [m](int r) { return smfem_sync(m->ctx[r]); });
}

int smfem_multi_meshgrid(smfem_multi *m, double x0, double x1, double y0, double y1, double z0, double z1, int64_t ne, int ndim,
                         smfem_multi_mesh **out) {
    if (!m || !out) return fail(SMFEM_ERR_INVALID, "NULL argument");
    auto *mm = new smfem_multi_mesh();
    mm->h.assign(m->n, nullptr);
    int st = run_all(m, [=](int r) { return smfem_meshgrid(m->ctx[r], x0, x1, y0, y1, z0, z1, ne, ndim, &mm->h[r]); });
    if (st != SMFEM_OK) {
        smfem_multi_mesh_free(m, mm);
        return st;
    }
    *out = mm;
    return SMFEM_OK;
}

int smfem_multi_inflate_sphere(smfem_multi *m, smfem_multi_mesh *mesh, double x0, double x1, double y0, double y1) {
    if (!m || !mesh) return fail(SMFEM_ERR_INVALID, "NULL argument");
    return run_all(m, [=](int r) { return smfem_inflate_sphere(m->ctx[r], mesh->h[r], x0, x1, y0, y1); });
}

int smfem_multi_assemble(smfem_multi *m, smfem_multi_mesh *mesh, int64_t ne, int ndim, int func_class, int nDof, double Young,
                         double nu, smfem_multi_matrix **K_out) {
    if (!m || !mesh || !K_out) return fail(SMFEM_ERR_INVALID, "NULL argument");
    auto *K = new smfem_multi_matrix();
    K->h.assign(m->n, nullptr);
    int st = run_all(m, [=](int r) { return smfem_assemble(m->ctx[r], mesh->h[r], ne, ndim, func_class, nDof, Young, nu, &K->h[r]); });
    if (st != SMFEM_OK) {
        smfem_multi_matrix_free(m, K);
        return st;
    }
    *K_out = K;
    return SMFEM_OK;
}

// assemble_system(ne, NodeList, IEN, ndim, FunctionClass, nDof, ID, Young, nu) (src/fem.jl:135) with the reference's HOST arrays,
// read concurrently by all ranks (each streams / verifies the part its slab uses)
int smfem_multi_assemble_system(smfem_multi *m, const double *NodeList, const int64_t *IEN, const int64_t *ID, int64_t nNodes,
                                int64_t nEl, int nLocal, int64_t ne, int ndim, int func_class, int nDof, double Young, double nu,
                                smfem_multi_mesh **mesh_out, smfem_multi_matrix **K_out) {
    if (!m || !mesh_out || !K_out) return fail(SMFEM_ERR_INVALID, "NULL argument");
    auto *mm = new smfem_multi_mesh();
    auto *K = new smfem_multi_matrix();
    mm->h.assign(m->n, nullptr);
    K->h.assign(m->n, nullptr);
    int st = run_all(m, [=](int r) {
        return smfem_assemble_system(m->ctx[r], NodeList, IEN, ID, nNodes, nEl, nLocal, ne, ndim, func_class, nDof, Young, nu, &mm->h[r],
                                     &K->h[r]);
    });
    if (st != SMFEM_OK) {
        smfem_multi_matrix_free(m, K);
        smfem_multi_mesh_free(m, mm);
        return st;
    }
    *mesh_out = mm;
    *K_out = K;
    return SMFEM_OK;
}

int smfem_multi_reassemble(smfem_multi *m, smfem_multi_mesh *mesh, smfem_multi_matrix *K, double Young, double nu) {
    if (!m || !mesh || !K) return fail(SMFEM_ERR_INVALID, "NULL argument");
    return run_all(m, [=](int r) { return smfem_reassemble(m->ctx[r], mesh->h[r], K->h[r], Young, nu); });
}

int smfem_multi_matrix_info(smfem_multi *m, smfem_multi_matrix *K, int64_t *mrows, int64_t *ncols, int64_t *nnz) {
    if (!m || !K) return fail(SMFEM_ERR_INVALID, "NULL argument");
    int64_t row0, nrl, nnzl;
    return smfem_matrix_info(K->h[0], mrows, ncols, nnz, &row0, &nrl, &nnzl);
}

// K_bar = K + beta * b over the lattice's top / bottom faces (examples/vector3D.jl:175-264, :308); mesh = the coordinates b is
// integrated over
int smfem_multi_surface_mass(smfem_multi *m, smfem_multi_matrix *K, smfem_multi_mesh *mesh, double beta) {
    if (!m || !mesh || !K) return fail(SMFEM_ERR_INVALID, "NULL argument");
    return run_all(m, [=](int r) { return smfem_surface_mass(m->ctx[r], K->h[r], mesh->h[r], nullptr, nullptr, 0, beta, 0); });
}

int smfem_multi_set_dirichlet_zplanes(smfem_multi *m, smfem_multi_matrix *K, smfem_multi_mesh *mesh, double d) {
    if (!m || !mesh || !K) return fail(SMFEM_ERR_INVALID, "NULL argument");
    return run_all(m, [=](int r) { return smfem_set_dirichlet_zplanes(m->ctx[r], K->h[r], mesh->h[r], d); });
}

// the peer windows of the n matrices, by raw pointer (first solve; idempotent)
static int multi_connect(smfem_multi *m, smfem_multi_matrix *K) {
    if (K->connected || m->n == 1) return SMFEM_OK;
    int st = run_all(m, [=](int r) { return smfem_comm_prepare(m->ctx[r], K->h[r]); });
    if (st != SMFEM_OK) return st;
    st = run_all(m, [=](int r) { return smfem_comm_connect_local(m->ctx[r], K->h[r], K->h.data(), m->n); });
    if (st == SMFEM_OK) K->connected = true;
    return st;
}

int smfem_multi_pcg_use_multigrid(smfem_multi *m, smfem_multi_matrix *K, smfem_multi_mesh *mesh, int enable) {
    if (!m || !K) return fail(SMFEM_ERR_INVALID, "NULL argument");
    int st = multi_connect(m, K);
    if (st != SMFEM_OK) return st;
    return run_all(m, [=](int r) { return smfem_pcg_use_multigrid(m->ctx[r], K->h[r], mesh ? mesh->h[r] : nullptr, enable); });
}

int smfem_multi_pcg_set_warm_start(smfem_multi *m, smfem_multi_matrix *K, double scale) {
    if (!m || !K) return fail(SMFEM_ERR_INVALID, "NULL argument");
    for (int r = 0; r < m->n; ++r) {
        int st = smfem_pcg_set_warm_start(K->h[r], scale);
        if (st != SMFEM_OK) return st;
    }
    return SMFEM_OK;
}

// q = inv(C' K_bar C) C' (rhs - K_bar q_d) + q_d (examples/vector3D.jl:315-322) with GLOBAL host vectors: rank r reads / writes
// its row slab [row0_r, row0_r + nrows_r)
int smfem_multi_pcg_solve(smfem_multi *m, smfem_multi_matrix *K, double rtol, int maxit, const double *rhs_extra_global,
                          double *q_global, int *iters, double *relres) {
    if (!m || !K) return fail(SMFEM_ERR_INVALID, "NULL argument");
    int st = multi_connect(m, K);
    if (st != SMFEM_OK) return st;
    std::vector<int> it(m->n, 0);
    std::vector<double> rel(m->n, 0.0);
    st = run_all(m, [&, m, K](int r) {
        int64_t mm, nn, nnz, row0, nrl, nnzl;
        int s = smfem_matrix_info(K->h[r], &mm, &nn, &nnz, &row0, &nrl, &nnzl);
        if (s != SMFEM_OK) return s;
        return smfem_pcg_solve(m->ctx[r], K->h[r], rtol, maxit, rhs_extra_global ? rhs_extra_global + row0 : nullptr,
                               q_global ? q_global + row0 : nullptr, &it[r], &rel[r]);
    });
    if (iters) *iters = it[0];
    if (relres) *relres = rel[0];
    return st;
}

int smfem_multi_pcg_stats(smfem_multi *m, smfem_multi_matrix *K, float *ms_total_max, int *iters) {
    if (!m || !K) return fail(SMFEM_ERR_INVALID, "NULL argument");
    float mx = 0;
    int it = 0;
    for (int r = 0; r < m->n; ++r) {
        float a = 0, b = 0;
        int st = smfem_pcg_stats(K->h[r], &a, &b, &it);
        if (st != SMFEM_OK) return st;
        mx = a > mx ? a : mx;
    }
    if (ms_total_max) *ms_total_max = mx;
    if (iters) *iters = it;
    return SMFEM_OK;
}

// SparseMatrixCSC(K) for the whole matrix: rank r's column slab lands at its offset (K is symmetric: CSC of the column slab ==
// transposed CSR of the row slab); colptr has m + 1 entries, rowval / nzval nnz entries, 1-based
int smfem_multi_matrix_export_csc(smfem_multi *m, smfem_multi_matrix *K, int which, int64_t *colptr, int64_t *rowval, double *nzval) {
    if (!m || !K || !colptr || !rowval || !nzval) return fail(SMFEM_ERR_INVALID, "NULL argument");
    std::vector<int64_t> row0(m->n), nrl(m->n), nnzl(m->n), off(m->n + 1, 0);
    for (int r = 0; r < m->n; ++r) {
        int64_t mm, nn, nnz;
        int st = smfem_matrix_info(K->h[r], &mm, &nn, &nnz, &row0[r], &nrl[r], &nnzl[r]);
        if (st != SMFEM_OK) return st;
        off[r + 1] = off[r] + nnzl[r];
    }
    // every rank writes its own slab of the three arrays; the local colptr (nrl + 1 entries, starting at 1) is shifted afterwards
    std::vector<std::vector<int64_t>> cp(m->n);
    int st = run_all(m, [&, m, K](int r) {
        cp[r].assign(nrl[r] + 1, 0);
        return smfem_matrix_export_csc(m->ctx[r], K->h[r], which, cp[r].data(), rowval + off[r], nzval + off[r]);
    });
    if (st != SMFEM_OK) return st;
    for (int r = 0; r < m->n; ++r)
        for (int64_t i = 0; i <= nrl[r]; ++i)
            if (i < nrl[r] || r == m->n - 1) colptr[row0[r] + i] = cp[r][i] + off[r];
    return SMFEM_OK;
}

int smfem_multi_matrix_free(smfem_multi *m, smfem_multi_matrix *K) {
    if (!K) return SMFEM_OK;
    if (m)
        run_all(m, [=](int r) {
            if (K->h[r]) smfem_matrix_free(K->h[r]);
            return (int)SMFEM_OK;
        });
    delete K;
    return SMFEM_OK;
}

int smfem_multi_mesh_free(smfem_multi *m, smfem_multi_mesh *mesh) {
    if (!mesh) return SMFEM_OK;
    if (m)
        run_all(m, [=](int r) {
            if (mesh->h[r]) smfem_mesh_free(mesh->h[r]);
            return (int)SMFEM_OK;
        });
    delete mesh;
    return SMFEM_OK;
}

}  // extern "C"
