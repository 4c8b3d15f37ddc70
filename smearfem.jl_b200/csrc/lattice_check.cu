// Is (IEN, ID) exactly what meshgrid(..., ne, 3) produces (examples/vector3D.jl:74, :94-101)?  Host arrays, two checkers.
//
// On a hex lattice the connectivity is redundant: the assembly needs only the coordinates.  But the reference interface
// hands over IEN (8 Int64 per element) and ID (3 Int64 per node) - 4.7x the bytes of NodeList - and they have to be
// verified, not trusted.  The arrays are cut into chunks; the calling thread streams chunks from the FRONT through PCIe
// (two staging buffers, check kernel behind each copy) while a small pool of host threads verifies chunks from the BACK
// where they lie, until the two fronts meet.  Whichever side is faster on the machine at hand does most of the work; the
// verdict is exact either way (every entry is compared by one of the two checkers).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>

#include <sched.h>

#include "smfem_internal.cuh"

namespace {

// device side: entries [off, off+len) of IEN (which = 0, column-major nEl x 8) or ID (which = 1, nNodes x 3)
__global__ void k_check_lattice_range(const int64_t *__restrict__ chunk, int which, int64_t off, int64_t len, int64_t nEl, int ne,
                                      int64_t nNodes, int *__restrict__ mismatch) {
    // small blocks, few registers: these CTAs have to fit beside the resident CTAs of the tile kernel (2 x 128 threads x 246
    // registers per SM) to run while the assembly is in flight
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = off + i;
    int64_t want;
    if (which == 0) {
        const int n1 = ne + 1;
        const int64_t e = t % nEl;
        const int a = (int)(t / nEl);
        const int ei = (int)(e % ne), ej = (int)((e / ne) % ne), ek = (int)(e / ((int64_t)ne * ne));
        const int ox = ((a & 3) == 1 || (a & 3) == 2), oy = ((a & 3) >= 2), oz = (a >> 2);
        want = ((int64_t)(ek + oz) * n1 + (ej + oy)) * n1 + (ei + ox) + 1;
    } else {
        want = 3 * (t % nNodes) + t / nNodes + 1;
    }
    if (chunk[i] != want) *mismatch = 1;
    }
}

// host side, same predicate, one division per lattice row
bool host_check_range(const int64_t *arr, int which, int64_t t0, int64_t t1, int64_t nEl, int ne, int64_t nNodes) {
    const int n1 = ne + 1;
    int64_t t = t0;
    while (t < t1) {
        int64_t run, base, step;
        if (which == 0) {
            const int64_t e = t % nEl;
            const int a = (int)(t / nEl);
            const int64_t layer = (int64_t)ne * ne;
            const int ek = (int)(e / layer), ej = (int)((e % layer) / ne), ei = (int)(e % ne);
            const int ox = ((a & 3) == 1 || (a & 3) == 2), oy = ((a & 3) >= 2), oz = (a >> 2);
            base = ((int64_t)(ek + oz) * n1 + (ej + oy)) * n1 + (ei + ox) + 1;
            run = std::min<int64_t>(ne - ei, t1 - t);
            step = 1;
        } else {
            const int64_t m = t % nNodes, l = t / nNodes;
            base = 3 * m + l + 1;
            run = std::min<int64_t>(nNodes - m, t1 - t);
            step = 3;
        }
        const int64_t *p = arr + t;
        uint64_t acc = 0;
        for (int64_t i = 0; i < run; ++i) acc |= (uint64_t)(p[i] ^ (base + step * i));
        if (acc) return false;
        t += run;
    }
    return true;
}

// CPUs this process may really use: affinity mask and cgroup quota (containers often show all cores of the host in
// hardware_concurrency() but throttle beyond the quota - measured: 8 checker threads on such a box are 2.5x slower than 4).
// The ranks of one box share its cores: each takes its 1/nranks share (8 ranks x 4 checkers on a 32-core host left no
// core for the callers, and the end-to-end step scaled at 0.45).
int default_host_threads(int nranks) {
    int n = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = std::min(n, CPU_COUNT(&set));
    if (FILE *f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {  // cgroup v2: "<quota> <period>" or "max <period>"
        long long q = 0, per = 0;
        if (std::fscanf(f, "%lld %lld", &q, &per) == 2 && q > 0 && per > 0) n = std::min<long long>(n, (q + per - 1) / per);
        std::fclose(f);
    } else if (FILE *g = std::fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) {  // cgroup v1
        long long q = -1, per = 100000;
        if (std::fscanf(g, "%lld", &q) != 1) q = -1;
        std::fclose(g);
        if (FILE *h = std::fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) {
            if (std::fscanf(h, "%lld", &per) != 1) per = 100000;
            std::fclose(h);
        }
        if (q > 0 && per > 0) n = std::min<long long>(n, (q + per - 1) / per);
    }
    if (nranks > 1) n = n / nranks;
    return std::max(1, std::min(4, n - 1));  // the calling thread drives the PCIe side
}

}  // namespace

// A few sleeping threads per context; run(f) executes f(worker) on all of them and returns immediately, wait() joins the round.
struct HostPool {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::function<void(int)> job;
    uint64_t round = 0;
    int pending = 0;
    bool stop = false;

    explicit HostPool(int n) {
        for (int w = 0; w < n; ++w)
            threads.emplace_back([this, w] {
                uint64_t seen = 0;
                for (;;) {
                    std::function<void(int)> f;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return stop || round != seen; });
                        if (stop) return;
                        seen = round;
                        f = job;
                    }
                    f(w);
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        if (--pending == 0) cv_done.notify_all();
                    }
                }
            });
    }
    void run(std::function<void(int)> f) {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = std::move(f);
            pending = (int)threads.size();
            ++round;
        }
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto &t : threads) t.join();
    }
};

void host_pool_destroy(smfem_ctx *ctx) {
    delete static_cast<HostPool *>(ctx->host_pool);
    ctx->host_pool = nullptr;
}

// Queues the PCIe side on ctx->copy_stream (the verdict of that side lands in d_flag[0]) and runs the host side to
// completion.  Returns false as soon as the host side has seen a mismatch; true means "host side clean" - the caller still
// has to read d_flag after the copy stream has drained.  d_stage: 2 * LATTICE_CHUNK device words.
bool lattice_check_hybrid(smfem_ctx *ctx, const Lattice &L, const int64_t *IEN, const int64_t *ID, int64_t nEl, int64_t nNodes, int ne,
                          int64_t *d_stage, int *d_flag) {
    struct Chunk {
        int which;
        int64_t off, len;
    };
    std::vector<Chunk> chunks;
    // this rank answers for what its slab uses: the element layers k0-1 .. k1-1 and the node planes k0-1 .. k1 (the whole
    // arrays on one GPU).  A rank whose part is not meshgrid's numbering leaves the lattice path on its own.
    auto add_range = [&](int which, int64_t b, int64_t e) {
        for (int64_t o = b; o < e; o += LATTICE_CHUNK) chunks.push_back({which, o, std::min<int64_t>(LATTICE_CHUNK, e - o)});
    };
    {
        const int64_t layer = (int64_t)ne * ne, plane = (int64_t)(ne + 1) * (ne + 1);
        const int64_t l0 = std::max(L.k0 - 1, 0), l1 = std::min(L.k1, ne);          // element layers [l0, l1)
        const int64_t p0 = std::max(L.k0 - 1, 0), p1 = std::min(L.k1 + 1, ne + 1);  // node planes [p0, p1)
        for (int a = 0; a < 8; ++a) add_range(0, a * nEl + l0 * layer, a * nEl + l1 * layer);
        for (int l = 0; l < 3; ++l) add_range(1, l * nNodes + p0 * plane, l * nNodes + p1 * plane);
    }
    std::atomic<int64_t> lo{0}, hi{(int64_t)chunks.size()};
    std::atomic<int> bad{0};

    int nthreads = 0;
    if (const char *e = std::getenv("SMFEM_HOST_THREADS")) nthreads = std::atoi(e);
    else nthreads = default_host_threads(ctx->nranks);
    HostPool *pool = static_cast<HostPool *>(ctx->host_pool);
    if (nthreads > 0 && (!pool || (int)pool->threads.size() != nthreads)) {
        delete pool;
        pool = new HostPool(nthreads);
        ctx->host_pool = pool;
    }
    if (nthreads > 0)
        pool->run([&](int) {
            for (;;) {
                int64_t h = hi.load();
                if (h <= lo.load() || bad.load()) return;
                if (!hi.compare_exchange_weak(h, h - 1)) continue;
                const Chunk &c = chunks[h - 1];
                if (!host_check_range(c.which ? ID : IEN, c.which, c.off, c.off + c.len, nEl, ne, nNodes)) bad.store(1);
            }
        });
    // PCIe side, from the front; at most two chunks in flight
    cudaStream_t cs = ctx->copy_stream;
    int slot = 0;
    bool used[2] = {false, false};
    try {
        for (;;) {
            int64_t l = lo.load();
            if (l >= hi.load() || bad.load()) break;
            if (!lo.compare_exchange_weak(l, l + 1)) continue;
            const Chunk &c = chunks[l];
            if (used[slot]) CUDA_CHECK(cudaEventSynchronize(ctx->ev_stage[slot]));
            int64_t *dst = d_stage + (int64_t)slot * LATTICE_CHUNK;
            CUDA_CHECK(cudaMemcpyAsync(dst, (c.which ? ID : IEN) + c.off, 8 * c.len, cudaMemcpyHostToDevice, cs));
            k_check_lattice_range<<<(unsigned)((c.len + 255) / 256), 64, 0, cs>>>(dst, c.which, c.off, c.len, nEl, ne, nNodes, d_flag);
            ctx->launches++;
            CUDA_CHECK(cudaGetLastError());
            CUDA_CHECK(cudaEventRecord(ctx->ev_stage[slot], cs));
            used[slot] = true;
            slot ^= 1;
            ctx->h2d_bytes += 8 * c.len;
        }
    } catch (...) {
        bad.store(1);
        if (nthreads > 0) pool->wait();
        throw;
    }
    if (nthreads > 0) pool->wait();
    return bad.load() == 0;
}
