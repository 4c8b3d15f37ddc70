// Device side of the example's per-load-step post-processing, so that only the projected
// border nodes cross PCIe after a solve instead of the whole displacement field:
//   motion       = [q[ID[:,1]] q[ID[:,2]] q[ID[:,3]]]'           examples/vector3D.jl:325
//   NodeList_new = NodeListCylinder + motion                      examples/vector3D.jl:327
//   back_project(NodeList_new[:, ids], CameraMatrix)              src/PostProcess.jl:131-152 (first step of extract_borders, :62-64)
// The convex hull / spline fitting / plotting that follow stay on the host with the unchanged PostProcess.jl.
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
