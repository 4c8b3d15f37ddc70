// Device side of the example's per-load-step post-processing, so that only the projected
// border nodes cross PCIe after a solve instead of the whole displacement field:
//   motion       = [q[ID[:,1]] q[ID[:,2]] q[ID[:,3]]]'           examples/vector3D.jl:325
//   NodeList_new = NodeListCylinder + motion                      examples/vector3D.jl:327
//   back_project(NodeList_new[:, ids], CameraMatrix)              src/PostProcess.jl:131-152 (first step of extract_borders, :62-64)
//   extract_borders(NodeList_new, CameraMatrix, BorderNodesList, state, ne)   src/PostProcess.jl:60-117: projection of the side nodes
//       on the device; "init": per-layer left / right extremes and the sorted top / bottom arcs on the device; "update": convex
//       hull (LazySets.convex_hull = Andrew's monotone chain) of the projected points
// Spline fitting / plotting stay on the host with the unchanged PostProcess.jl.
#include <algorithm>
#include <utility>
#include <vector>

#include "smfem_internal.cuh"

namespace {

__global__ void k_project_nodes(int64_t n, const int64_t *__restrict__ ids, int64_t nNodes, int64_t ghost_nodes, int64_t ghost_cols,
                                const double *__restrict__ coords, const int32_t *__restrict__ id, const double *__restrict__ qd,
                                const double *__restrict__ x, const double *__restrict__ cam, double *__restrict__ out3,
                                double *__restrict__ out2, int *__restrict__ err) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t m = ids[t] - 1;  // 1-based node id, as Julia passes it
    if (m < 0 || m >= nNodes) {
        *err = 1;
        return;
    }
    double p[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int64_t dof = id ? (int64_t)id[(int64_t)c * nNodes + m] : 3 * m + c;  // ID[m, c] (0-based)
        const double q = (qd ? qd[ghost_cols + dof] : 0.0) + (x ? x[dof] : 0.0);   // q = q_d + C q_f, examples/vector3D.jl:322
        p[c] = coords[3 * (m + ghost_nodes) + c] + q;
        if (out3) out3[3 * t + c] = p[c];
    }
    if (!out2) return;
    // R = [1 0 0; 0 0 1; 0 -1 0], t = [0; -0.5; 2]   (src/PostProcess.jl:134-137)
    const double X = p[0] + 0.0, Y = p[2] + -0.5, Z = -p[1] + 2.0;
    const double nx = X / Z, ny = Y / Z, nz = Z / Z;  // :141-145
    // NodeListProj = CameraMatrix' * NodeListNorm, rows 1:2   (:147-149); cam is column-major 3 x 3
    out2[2 * t + 0] = cam[0] * nx + cam[1] * ny + cam[2] * nz;
    out2[2 * t + 1] = cam[3] * nx + cam[4] * ny + cam[5] * nz;
}

// ---- extract_borders, state "init" (src/PostProcess.jl:67-100) ------------------------------------------------------
// one CTA per layer of side nodes: first index of the minimal / maximal projected x (Julia's argmin / argmax)
__global__ void __launch_bounds__(256) k_border_layers(int szSide, const double *__restrict__ p2, int *__restrict__ minNode,
                                                       int *__restrict__ maxNode) {
    __shared__ double s_v[256];
    __shared__ int s_i[256];
    const int layer = blockIdx.x, tid = threadIdx.x;
    for (int pass = 0; pass < 2; ++pass) {  // 0: argmin, 1: argmax
        double bv = 0.0;
        int bi = -1;
        for (int t = tid; t < szSide; t += 256) {
            const double v = p2[2 * ((int64_t)layer * szSide + t)];
            const bool better = bi < 0 || (pass == 0 ? v < bv : v > bv);  // strict: the first index wins ties (t ascends per thread)
            if (better) {
                bv = v;
                bi = t;
            }
        }
        s_v[tid] = bv;
        s_i[tid] = bi;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if (tid < w) {
                const int oi = s_i[tid + w];
                const double ov = s_v[tid + w];
                const int mi = s_i[tid];
                const double mv = s_v[tid];
                const bool take = oi >= 0 && (mi < 0 || (pass == 0 ? ov < mv : ov > mv) || (ov == mv && oi < mi));
                if (take) {
                    s_v[tid] = ov;
                    s_i[tid] = oi;
                }
            }
            __syncthreads();
        }
        if (tid == 0) (pass == 0 ? minNode : maxNode)[layer] = layer * szSide + s_i[0];
        __syncthreads();
    }
}

// single CTA: BorderPoints = [Left | sort(Top) | reverse(Right) | reverse(sort(Bottom))]  (:96-99); sort = sortslices(dims=2):
// columns in lexicographic order (x, then y), by ranking (the arcs hold <= szSide points)
__global__ void __launch_bounds__(1024) k_border_assemble(int nLayers, int szSide, const double *__restrict__ p2,
                                                         const int *__restrict__ minNode, const int *__restrict__ maxNode,
                                                         double *__restrict__ border, int *__restrict__ counts) {
    __shared__ int s_nt, s_nb;
    const int tid = threadIdx.x;
    const int topBase = (nLayers - 1) * szSide;
    const double yTop = p2[2 * (int64_t)minNode[nLayers - 1] + 1], yBot = p2[2 * (int64_t)minNode[0] + 1];
    const bool has_bottom = nLayers > 1;  // `elseif Layers == 1`: with a single layer only the top branch runs
    auto in_top = [&](int t) { return p2[2 * (int64_t)(topBase + t) + 1] > yTop; };
    auto in_bot = [&](int t) { return has_bottom && p2[2 * (int64_t)t + 1] < yBot; };
    if (tid == 0) {
        int nt = 0, nb = 0;
        for (int t = 0; t < szSide; ++t) {
            nt += in_top(t);
            nb += in_bot(t);
        }
        s_nt = nt;
        s_nb = nb;
        counts[0] = 2 * nLayers + nt + nb;
        counts[1] = nt;
        counts[2] = nb;
    }
    __syncthreads();
    const int nt = s_nt, nb = s_nb;
    for (int l = tid; l < nLayers; l += blockDim.x) {
        border[2 * l] = p2[2 * (int64_t)minNode[l]];
        border[2 * l + 1] = p2[2 * (int64_t)minNode[l] + 1];
        const int dst = nLayers + nt + (nLayers - 1 - l);  // reverse(RightborderNodes)
        border[2 * dst] = p2[2 * (int64_t)maxNode[l]];
        border[2 * dst + 1] = p2[2 * (int64_t)maxNode[l] + 1];
    }
    for (int t = tid; t < szSide; t += blockDim.x) {
        for (int arc = 0; arc < 2; ++arc) {
            const bool mine = arc == 0 ? in_top(t) : in_bot(t);
            if (!mine) continue;
            const int64_t me = arc == 0 ? topBase + t : t;
            const double x = p2[2 * me], y = p2[2 * me + 1];
            int rank = 0;
            for (int u = 0; u < szSide; ++u) {
                if (!(arc == 0 ? in_top(u) : in_bot(u))) continue;
                const int64_t ot = arc == 0 ? topBase + u : u;
                const double ox = p2[2 * ot], oy = p2[2 * ot + 1];
                rank += (ox < x) || (ox == x && (oy < y || (oy == y && u < t)));
            }
            const int dst = arc == 0 ? nLayers + rank : 2 * nLayers + nt + (nb - 1 - rank);  // reverse(BottomLayer)
            border[2 * dst] = x;
            border[2 * dst + 1] = y;
        }
    }
}

}  // namespace

void project_nodes(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, const int64_t *ids, int64_t n, const double *cam,
                   double *nodes3d_out, double *nodes2d_out) {
    REQUIRE(ctx->nranks == 1, SMFEM_ERR_UNSUPPORTED, "project_nodes is single-GPU");
    REQUIRE(mesh->ndim == 3 && (!K || K->nDof == 3), SMFEM_ERR_UNSUPPORTED, "back_project needs 3-D nodes (src/PostProcess.jl:134)");
    REQUIRE(n >= 0 && (n == 0 || ids), SMFEM_ERR_INVALID, "node id list missing");
    REQUIRE(!nodes2d_out || cam, SMFEM_ERR_INVALID, "CameraMatrix missing");
    if (n == 0) return;
    int64_t *d_ids = dev_alloc<int64_t>(n);
    double *d_cam = dev_alloc<double>(9), *d3 = nodes3d_out ? dev_alloc<double>(3 * n) : nullptr,
           *d2 = nodes2d_out ? dev_alloc<double>(2 * n) : nullptr;
    int *d_err = dev_alloc<int>(1);
    auto cleanup = [&] {
        dev_free(d_ids);
        dev_free(d_cam);
        dev_free(d3);
        dev_free(d2);
        dev_free(d_err);
    };
    try {
        CUDA_CHECK(cudaMemcpyAsync(d_ids, ids, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
        if (cam) CUDA_CHECK(cudaMemcpyAsync(d_cam, cam, 72, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(d_err, 0, 4, ctx->stream));
        const int64_t ghost_nodes = mesh->structured ? mesh->lat.plane() : 0;
        const bool solved = K && K->sol_x;
        LAUNCH(ctx, k_project_nodes, (unsigned)((n + 255) / 256), 256, 0, n, (const int64_t *)d_ids, mesh->nNodes_g, ghost_nodes,
               K ? K->ghost_cols : 0, (const double *)mesh->coords, (const int32_t *)(mesh->structured ? nullptr : mesh->id),
               (const double *)(solved && K->has_bc ? K->qd : nullptr), (const double *)(solved ? K->sol_x : nullptr), (const double *)d_cam,
               d3, d2, d_err);
        int err = 0;
        CUDA_CHECK(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (d3) CUDA_CHECK(cudaMemcpyAsync(nodes3d_out, d3, 24 * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (d2) CUDA_CHECK(cudaMemcpyAsync(nodes2d_out, d2, 16 * n, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        REQUIRE(err == 0, SMFEM_ERR_INVALID, "node id out of range (reference: BoundsError)");
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

// extract_borders (src/PostProcess.jl:60-117) for the node list ids = BorderNodesList[1] (1-based), on the coordinates displaced
// by the last solve.  state 0 = "init" (needs ne), 1 = "update".  border_out: 2 x cap column-major; *nborder = columns written.
void extract_borders(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, const int64_t *ids, int64_t n, const double *cam, int state,
                     int64_t ne, double *border_out, int64_t cap, int64_t *nborder, double *side2d_out) {
    REQUIRE(state == 0 || state == 1, SMFEM_ERR_INVALID, "extract_borders: state must be init (0) or update (1) (reference: UndefVarError)");
    REQUIRE(n > 0 && ids && cam && border_out && nborder, SMFEM_ERR_INVALID, "extract_borders: missing argument");
    std::vector<double> side((size_t)2 * n);
    if (state == 1) {
        project_nodes(ctx, mesh, K, ids, n, cam, nullptr, side.data());
        // LazySets.convex_hull (monotone chain): sort by (x, y); lower hull, upper hull; pop while the turn is not strictly
        // counter-clockwise; vertices counter-clockwise from the lexicographically smallest point
        std::vector<std::pair<double, double>> pts((size_t)n);
        for (int64_t i = 0; i < n; ++i) pts[i] = {side[2 * i], side[2 * i + 1]};
        std::sort(pts.begin(), pts.end());
        std::vector<std::pair<double, double>> hull;
        if (n <= 2) {
            for (auto &p : pts)
                if (hull.empty() || hull.back() != p) hull.push_back(p);
        } else {
            auto turn = [](const std::pair<double, double> &o, const std::pair<double, double> &a, const std::pair<double, double> &b) {
                return (a.first - o.first) * (b.second - o.second) - (a.second - o.second) * (b.first - o.first);
            };
            std::vector<std::pair<double, double>> lower, upper;
            for (int64_t i = 0; i < n; ++i) {
                while (lower.size() >= 2 && turn(lower[lower.size() - 2], lower.back(), pts[i]) <= 0.0) lower.pop_back();
                lower.push_back(pts[i]);
            }
            for (int64_t i = n - 1; i >= 0; --i) {
                while (upper.size() >= 2 && turn(upper[upper.size() - 2], upper.back(), pts[i]) <= 0.0) upper.pop_back();
                upper.push_back(pts[i]);
            }
            hull.assign(lower.begin(), lower.end() - 1);
            hull.insert(hull.end(), upper.begin(), upper.end() - 1);
        }
        REQUIRE((int64_t)hull.size() <= cap, SMFEM_ERR_INVALID, "extract_borders: border buffer too small");
        for (size_t i = 0; i < hull.size(); ++i) {
            border_out[2 * i] = hull[i].first;
            border_out[2 * i + 1] = hull[i].second;
        }
        *nborder = (int64_t)hull.size();
        if (side2d_out) std::copy(side.begin(), side.end(), side2d_out);
        return;
    }
    REQUIRE(ne >= 0, SMFEM_ERR_INVALID, "extract_borders: Number of elements must be provided");
    const int64_t nLayers = ne + 1;
    const int64_t szSide = n / nLayers;  // :73 (integer division, as the reference)
    REQUIRE(szSide >= 1 && szSide < (1 << 30), SMFEM_ERR_INVALID, "extract_borders: fewer side nodes than layers");
    REQUIRE(2 * nLayers + 2 * szSide <= cap, SMFEM_ERR_INVALID, "extract_borders: border buffer too small (need 2 (ne + 1) + 2 n / (ne + 1) columns)");
    // the projection stays on the device; only the border and (optionally) SideNodes2D come back
    int64_t *d_ids = dev_alloc<int64_t>(n);
    double *d_cam = dev_alloc<double>(9), *d2 = dev_alloc<double>(2 * n), *d_border = dev_alloc<double>(2 * (2 * nLayers + 2 * szSide));
    int *d_min = dev_alloc<int>(2 * nLayers + 4), *d_max = d_min + nLayers, *d_counts = d_max + nLayers;
    auto cleanup = [&] {
        dev_free(d_ids);
        dev_free(d_cam);
        dev_free(d2);
        dev_free(d_border);
        dev_free(d_min);
    };
    try {
        CUDA_CHECK(cudaMemcpyAsync(d_ids, ids, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(d_cam, cam, 72, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(d_counts, 0, 16, ctx->stream));
        const int64_t ghost_nodes = mesh->structured ? mesh->lat.plane() : 0;
        const bool solved = K && K->sol_x;
        LAUNCH(ctx, k_project_nodes, (unsigned)((n + 255) / 256), 256, 0, n, (const int64_t *)d_ids, mesh->nNodes_g, ghost_nodes,
               K ? K->ghost_cols : 0, (const double *)mesh->coords, (const int32_t *)(mesh->structured ? nullptr : mesh->id),
               (const double *)(solved && K->has_bc ? K->qd : nullptr), (const double *)(solved ? K->sol_x : nullptr), (const double *)d_cam,
               (double *)nullptr, d2, d_counts + 3);
        LAUNCH(ctx, k_border_layers, (unsigned)nLayers, 256, 0, (int)szSide, (const double *)d2, d_min, d_max);
        LAUNCH(ctx, k_border_assemble, 1, 1024, 0, (int)nLayers, (int)szSide, (const double *)d2, (const int *)d_min, (const int *)d_max,
               d_border, d_counts);
        int counts[4] = {0, 0, 0, 0};
        CUDA_CHECK(cudaMemcpyAsync(counts, d_counts, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        REQUIRE(counts[3] == 0, SMFEM_ERR_INVALID, "node id out of range (reference: BoundsError)");
        CUDA_CHECK(cudaMemcpyAsync(border_out, d_border, 16 * (size_t)counts[0], cudaMemcpyDeviceToHost, ctx->stream));
        if (side2d_out) CUDA_CHECK(cudaMemcpyAsync(side2d_out, d2, 16 * n, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        *nborder = counts[0];
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}
