/*
 * smearfem_b200.h -- C ABI of libsmearfem_b200.so: the B200-native (sm_100a) implementation of
 * smearFEM.jl's assembly + solve hot path.
 *
 * The reference has no FFI of its own: its boundary is the Julia call surface
 *   src/smearFEM.jl:3      export assemble_system, gaussian_quadrature, basis_function
 *   examples/vector3D.jl   meshgrid (:10), setboundaryCond (:133), apply_boundary_conditions (:175),
 *                          solve idiom (:308-322);  src/PostProcess.jl:30 inflate_sphere
 * Every entry point below names the reference function (file:line, relative to the reference
 * repository) it stands in for; INTEGRATION.md shows the Julia `ccall` binding for each.
 *
 * Conventions
 *   - plain C: opaque handles, raw pointers, sizes; no C++/torch types cross this boundary.
 *   - every function returns an int status (SMFEM_OK == 0); smfem_last_error() gives the message of
 *     the last failure on the calling thread.  No exception or exit() crosses the ABI.
 *   - HOST arrays use Julia's layouts verbatim: NodeList Float64 ndim x nNodes column-major
 *     (xyz of a node contiguous); IEN Int64 nEl x nLocal column-major, 1-based; ID Int64
 *     nNodes x nDof column-major, 1-based; sparse results are SparseMatrixCSC parts
 *     (colptr, rowval, nzval), Int64, 1-based.
 *   - one context == one GPU (rank); the slab partition is by z node planes.
 *     Multi-GPU runs are either one process per GPU (torchrun; peer windows exchanged as CUDA IPC
 *     handles) or ONE process driving all GPUs through smfem_init_multi (last section).  There is
 *     no CPU fallback: without a CUDA device smfem_init fails with SMFEM_ERR_CUDA.
 *   - calls are blocking w.r.t. the host unless stated; handles are not thread-safe.
 */
#ifndef SMEARFEM_B200_H
#define SMEARFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMFEM_OK 0
#define SMFEM_ERR_INVALID 1     /* bad argument (Julia: DimensionMismatch / BoundsError / ArgumentError) */
#define SMFEM_ERR_CUDA 2        /* CUDA runtime failure, including "no device" */
#define SMFEM_ERR_UNSUPPORTED 3 /* reference feature outside the built scope (e.g. 1-D, Q2 with nDof>1) */
#define SMFEM_ERR_SINGULAR 4    /* PCG breakdown: p'Ap <= 0 (Julia: SingularException from inv) */

#define SMFEM_Q1 1
#define SMFEM_Q2 2

typedef struct smfem_ctx smfem_ctx;
typedef struct smfem_mesh smfem_mesh;
typedef struct smfem_matrix smfem_matrix;

int smfem_abi_version(void);
const char *smfem_last_error(void);

/* ---- host-side helpers (pure functions; same expression order as the reference) ------------- */

/* gaussian_quadrature(a,b,nGaussPoints)           src/fem.jl:21-31.  n in {2,3}; else SMFEM_ERR_INVALID
 * (the reference leaves xi undefined -> UndefVarError). xi, w: n doubles each. */
int smfem_gaussian_quadrature(double a, double b, int n, double *xi, double *w);

/* basis_function(xi,eta,zeta,FunctionClass)       src/fem.jl:48-114.  ndim = number of coordinates
 * given (1,2,3).  N: nn doubles; dN: nn x ndim COLUMN-major (Julia Matrix), except the reference's
 * 1-D quirk where dN is the 1x2 row [-0.5 0.5] (src/fem.jl:75).  *nn returns the node count. */
int smfem_basis_function(int ndim, int func_class, double xi, double eta, double zeta, double *N, double *dN,
                         int *nn);

/* ---- context --------------------------------------------------------------------------------- */

/* device: CUDA ordinal.  rank/nranks: position of this process in the z-slab partition (0/1 for a
 * single GPU).  Creates the stream all library kernels run on. */
int smfem_init(int device, int rank, int nranks, smfem_ctx **out);
int smfem_destroy(smfem_ctx *ctx);
/* raw cudaStream_t of the context (for callers that want to record their own events) */
int smfem_stream(smfem_ctx *ctx, void **stream_out);
/* CUDA-event stopwatch on the context's stream (what bench.py times kernels with) */
int smfem_timer_start(smfem_ctx *ctx);
int smfem_timer_stop(smfem_ctx *ctx, float *ms_out); /* records, synchronises, returns elapsed ms */
int smfem_sync(smfem_ctx *ctx);
/* counters: kernels launched by this library on this context since init (bench.py: gpu_launches) */
int smfem_launch_count(smfem_ctx *ctx, int64_t *count_out);
/* write `bytes` of device scratch (L2 flush between timed iterations) */
int smfem_flush_l2(smfem_ctx *ctx);

/* ---- mesh ------------------------------------------------------------------------------------ */

/* meshgrid(x0,x1,y0,y1,z0,z1,ne,ndim)              examples/vector3D.jl:10-130.
 * Generated ON DEVICE; with nranks > 1 only this rank's z-slab (owned node planes plus one ghost
 * plane per side) is materialised.  ndim in {2,3}. */
int smfem_meshgrid(smfem_ctx *ctx, double x0, double x1, double y0, double y1, double z0, double z1, int64_t ne,
                   int ndim, smfem_mesh **out);

/* An arbitrary user mesh as passed to assemble_system (src/fem.jl:135; coords gathered at :180,
 * ids at :243-244).  Host arrays in Julia layout; copied to the device; index ranges validated.
 * ID may be NULL for nDof == 1 (the reference then uses raw node ids, src/fem.jl:204-205).
 * If (IEN, ID) are exactly what meshgrid would produce for `ne`, the structured fast path is
 * taken; otherwise the general (unstructured) path.  nranks must be 1. */
int smfem_mesh_from_host(smfem_ctx *ctx, const double *NodeList, const int64_t *IEN, const int64_t *ID,
                         int64_t nNodes, int64_t nEl, int nLocal, int ndim, int nDof, int64_t ne,
                         smfem_mesh **out);

/* inflate_sphere(NodeList,x0,x1,y0,y1)             src/PostProcess.jl:30-44 (in place, on device) */
int smfem_inflate_sphere(smfem_ctx *ctx, smfem_mesh *mesh, double x0, double x1, double y0, double y1);

/* same, on a HOST NodeList (ndim x nNodes, mutated in place like the reference): upload, kernel, download */
int smfem_inflate_sphere_host(smfem_ctx *ctx, double *NodeList, int ndim, int64_t nNodes, double x0, double x1,
                              double y0, double y1);

/* Overwrite node coordinates from a GLOBAL host NodeList (ndim x nNodes_global); each rank takes
 * its slab.  (Robustness inputs: jittered meshes) */
int smfem_mesh_set_nodelist(smfem_ctx *ctx, smfem_mesh *mesh, const double *NodeList_global);

/* sizes: global node/element counts, nodes per element, this rank's first owned node (0-based) and
 * owned node count */
int smfem_mesh_info(smfem_mesh *mesh, int64_t *nNodes, int64_t *nEl, int *nLocal, int *ndim, int *structured,
                    int64_t *node0_owned, int64_t *nNodes_owned);
/* Element colouring of a general (IEN) mesh, built on first use and cached in the mesh: elements of one colour share no
 * node, so the value scatter of assemble_system (src/fem.jl:236-249) runs colour by colour (one launch per colour, at most
 * one add per entry and launch) and every entry of K is folded in a fixed order (bit-reproducible).  ncolors = -1: more than 64 colours would be needed, the
 * atomic scatter is used.  color_sizes (optional): 64 entries, elements per colour.  SMFEM_ERR_INVALID on lattice meshes
 * (their tiled gather kernel needs no colouring). */
int smfem_mesh_colors(smfem_ctx *ctx, smfem_mesh *mesh, int *ncolors, int64_t *color_sizes);
/* Export in Julia layout (caller allocates from smfem_mesh_info; any pointer may be NULL).
 * NodeList: ndim x nNodes_owned of THIS rank's owned nodes; IEN/ID/IEN_top/IEN_btm: global arrays
 * (structured meshes regenerate them; only sensible for small ne, rank 0). */
int smfem_mesh_export(smfem_ctx *ctx, smfem_mesh *mesh, double *NodeList_owned, int64_t *IEN, int64_t *ID,
                      int64_t *IEN_top, int64_t *IEN_btm);
int smfem_mesh_free(smfem_mesh *mesh);

/* ---- assembly -------------------------------------------------------------------------------- */

/* assemble_system(ne,NodeList,IEN,ndim,FunctionClass,nDof,ID,Young,nu)   src/fem.jl:135-256.
 * Builds the sparsity pattern on device (bit-exact with Julia's sparse(E,J,V), src/fem.jl:253:
 * explicit zeros kept, rows ascending per column) and the values (fp64).  nDof==1: scalar Laplace
 * (:199-208); nDof==2: plane stress (:210-217); nDof==3: 3-D isotropic (:218-230).  func_class SMFEM_Q2 is accepted
 * where upstream's Q2 is executable: ndim == 2, nDof == 1, 9-node elements, 2x2 Gauss rule (:77-111, :183).
 * The result stays on the device; this rank holds the rows of its owned nodes. */
int smfem_assemble(smfem_ctx *ctx, smfem_mesh *mesh, int64_t ne, int ndim, int func_class, int nDof, double Young,
                   double nu, smfem_matrix **K_out);
/* assemble_system with the reference's HOST arguments in one call (src/fem.jl:135: ne, NodeList, IEN, ndim, FunctionClass,
 * nDof, ID, Young, nu): smfem_mesh_from_host + smfem_assemble.  When the sizes match meshgrid's hex lattice (examples/
 * vector3D.jl:10) the transfers and the work overlap: NodeList is copied first and the assembly starts from the
 * coordinates alone, while IEN and ID (4.7x the bytes) are checked against the lattice numbering - chunks from the front
 * through PCIe and a check kernel on a second stream, chunks from the back by host threads (env SMFEM_HOST_THREADS, default
 * min(4, usable cores - 1); 0 = PCIe only; with several ranks each one checks the part its slab uses); a failed check discards the speculative result and takes the general path.  Returns once
 * every host array has been read; the matrix may still be in flight on the context's stream (any later call orders behind it).
 * Outputs: the mesh handle (needed by surface_mass / set_dirichlet_zplanes) and the matrix handle. */
int smfem_assemble_system(smfem_ctx *ctx, const double *NodeList, const int64_t *IEN, const int64_t *ID, int64_t nNodes,
                          int64_t nEl, int nLocal, int64_t ne, int ndim, int func_class, int nDof, double Young, double nu,
                          smfem_mesh **mesh_out, smfem_matrix **K_out);
/* Average duration (CUDA events on the context's stream) of the last `last_n` launches (<= 64; 0 = all recorded) of the
 * structured assembly kernel k_values_tile - the dominant kernel of smfem_assemble / smfem_reassemble; bench.py's
 * roofline denominator.  Synchronises with the last of those launches. */
int smfem_assembly_kernel_ms(smfem_ctx *ctx, int last_n, float *avg_ms, int *n_used);
/* cumulative bytes moved host->device / device->host by the bulk transfers of the calls above (bench.py's e2e accounting) */
int smfem_transfer_bytes(smfem_ctx *ctx, int64_t *h2d, int64_t *d2h);
/* The two halves of smfem_assemble, for timing them separately. */
int smfem_pattern_build(smfem_ctx *ctx, smfem_mesh *mesh, int ndim, int nDof, smfem_matrix **K_out);
int smfem_assemble_values(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, double Young, double nu);
/* Re-run the pattern kernels into K's existing buffers (no allocation, asynchronous): lets the
 * bench time pattern build + values back to back with CUDA events.  Structured meshes only. */
int smfem_pattern_rebuild(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K);
/* Full assembly (pattern AND values) into K's existing buffers, asynchronous, no allocation: rowptr in
 * closed form + the tiled value kernel, whose output phase also writes colind (fused; one pass over K).
 * This is what smfem_assemble runs after allocating, and what bench.py times as a step. */
int smfem_reassemble(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, double Young, double nu);

/* m, n: global dims; nnz: global stored entries; row0/nrows_local/nnz_local: this rank's slab */
int smfem_matrix_info(smfem_matrix *K, int64_t *m, int64_t *n, int64_t *nnz, int64_t *row0, int64_t *nrows_local,
                      int64_t *nnz_local);
/* SparseMatrixCSC parts of this rank's COLUMN slab [row0, row0+nrows_local) (the whole matrix
 * when nranks == 1).  colptr: nrows_local+1 entries, 1-based, relative to this slab's first stored
 * entry; rowval: global 1-based row ids; nzval: K[row, col] (true transpose of the device CSR).
 * which: 0 = K as assembled (+ anything added in place), 1 = the surface matrix b of
 * smfem_surface_mass on K's pattern, 2 = K's values exactly as stored (the device's row slab read as columns, i.e. K'
 * without the transposing search: what which = 0 returns when nranks > 1; for bit-comparisons of 1-GPU and N-GPU runs). */
int smfem_matrix_export_csc(smfem_ctx *ctx, smfem_matrix *K, int which, int64_t *colptr, int64_t *rowval,
                            double *nzval);
int smfem_matrix_diag(smfem_ctx *ctx, smfem_matrix *K, double *diag_local);
/* Device-side copy of K (pattern, values, diagonal): `K_bar = K + beta*b` of examples/vector3D.jl:308 leaves K itself intact, so
 * the host mirrors clone before they add the surface term in place.  Dirichlet data / solver state are not copied. */
int smfem_matrix_clone(smfem_ctx *ctx, smfem_matrix *K, smfem_matrix **K_out);
int smfem_matrix_free(smfem_matrix *K);

/* apply_boundary_conditions(ne,NodeList,IEN,IEN_top,IEN_btm,ndim,FunctionClass,ID,nDof)
 *                                                  examples/vector3D.jl:175-264  and
 * K_bar = K + beta*b                               examples/vector3D.jl:308.
 * Computes b = int_{top U bottom} N'N on K's pattern (kept for export as `which`=1 when keep_b != 0;
 * costs a second value array) and, if beta != 0, adds beta*b into K in place.  Structured meshes use their own top/bottom faces;
 * IEN_top/IEN_btm (host, Int64, nFaces x 4 column-major, 1-based) override them when non-NULL. */
int smfem_surface_mass(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, const int64_t *IEN_top,
                       const int64_t *IEN_btm, int64_t nFaces, double beta, int keep_b);

/* ---- Dirichlet conditions + solve ------------------------------------------------------------ */

/* setboundaryCond(NodeList,ne,ndim,FunctionClass,d,nDof)      examples/vector3D.jl:133-173:
 * z-dof of every node with z == 0 -> 0, with z == 1 -> -d (exact compares, :161-166). */
int smfem_set_dirichlet_zplanes(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, double d);
/* general form: global 1-based dof ids with prescribed values (replaces any earlier set) */
int smfem_set_dirichlet(smfem_ctx *ctx, smfem_matrix *K, const int64_t *dofs, const double *values, int64_t n);

/* K_free = C'K̄C ; q_f = K_free^{-1} C'(-K̄ q_d) ; q = q_d + C q_f      examples/vector3D.jl:315-322,
 * with the reference's dense inverse replaced by Jacobi-preconditioned CG on the masked operator
 * (Dirichlet rows/columns masked inside the SpMV kernel; K itself is not modified).
 * rhs_extra (host, nrows_local, may be NULL) is added to the right-hand side (manufactured-solution
 * tests).  q_out (host, nrows_local doubles, may be NULL) receives this rank's slab of q.
 * Stops when ||r||_2 <= rtol*||b||_2 or after maxit iterations. */
int smfem_pcg_solve(smfem_ctx *ctx, smfem_matrix *K, double rtol, int maxit, const double *rhs_extra,
                    double *q_out, int *iters_out, double *relres_out);

/* Per-load-step post-processing of the example on the device (examples/vector3D.jl:325-329): for the listed nodes
 * (1-based ids, e.g. BorderNodesList[1]) returns  NodeList_new[:, ids] = NodeList[:, ids] + motion[:, ids]  with
 * motion = q[ID]' taken from K's last smfem_pcg_solve (K = NULL or not solved yet: motion = 0), and its projection
 * back_project(NodeList_new[:, ids], CameraMatrix) (src/PostProcess.jl:131-152: R = [1 0 0; 0 0 1; 0 -1 0],
 * t = [0; -0.5; 2], perspective divide, CameraMatrix' * p, rows 1:2).  CameraMatrix: 3 x 3 column-major.  Outputs are
 * 3 x n and 2 x n column-major, either may be NULL.  One GPU. */
int smfem_project_nodes(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, const int64_t *node_ids, int64_t n,
                        const double *CameraMatrix, double *nodes3d_out, double *nodes2d_out);
/* extract_borders(NodeList_new, CameraMatrix, BorderNodesList, state, ne) (src/PostProcess.jl:60-117) on the coordinates displaced by
 * K's last solve: side_node_ids = BorderNodesList[1] (1-based, layer by layer as meshgrid lists them).  state 0 = "init": per layer
 * the first node of minimal / maximal projected x, the top / bottom arcs above / below the left extreme sorted lexicographically
 * (sortslices), BorderPoints = [Left | Top | reverse(Right) | reverse(Bottom)] -- all on the device, only the border crosses PCIe.
 * state 1 = "update": convex hull of the projected side nodes in LazySets.convex_hull's convention (Andrew's monotone chain:
 * counter-clockwise from the lexicographically smallest point, collinear points dropped; LazySets is not vendored with the
 * reference, so this convention is restated from its documentation).  BorderPoints_out: 2 x capacity column-major (init needs
 * 2 (ne + 1) + 2 n / (ne + 1) columns at most, update n); *nBorder_out = columns written; SideNodes2D_out (2 x n) may be NULL.
 * fit_curve / plotting remain host code (PostProcess.jl unchanged).  One GPU. */
int smfem_extract_borders(smfem_ctx *ctx, smfem_mesh *mesh, smfem_matrix *K, const int64_t *side_node_ids, int64_t n,
                          const double *CameraMatrix, int state, int64_t ne, double *BorderPoints_out, int64_t capacity,
                          int64_t *nBorder_out, double *SideNodes2D_out);
/* Opt-in: the following smfem_pcg_solve calls on K use CG preconditioned by a geometric multigrid V-cycle instead of
 * Jacobi (examples/vector3D.jl:315-322 solved in ~20 instead of ~9 ne iterations).  Hex-lattice matrices on one GPU only
 * (SMFEM_ERR_UNSUPPORTED otherwise); `mesh` is K's mesh and must outlive the solves.  Coarse levels (ceil(ne/2), ... down to
 * <= 4) take every other node of the finer mesh and are re-assembled with K's own E, nu and surface term; the hierarchy
 * is built at the first solve and rebuilt after a re-assembly.  enable = 0 returns to Jacobi-PCG. */
int smfem_pcg_use_multigrid(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, int enable);
/* z = M^-1 r: one application of the multigrid V-cycle to a host vector (this rank's rows in, this rank's rows out; constrained
 * rows are masked).  For tests of the preconditioner itself (linearity, symmetry, 1-GPU == N-GPU); collective over the ranks. */
/* Opt-in (structured 3-D hex lattice, nDof 3): the following smfem_pcg_solve / smfem_spmv_host / multigrid
 * fine-level products apply K_bar = K + beta*b MATRIX-FREE from `mesh`'s coordinates (72 B per node of traffic instead of 12 B per
 * nonzero); K must have been assembled on that mesh (its diagonal is the Jacobi preconditioner, its material / beta define the
 * operator).  Same vector layout, halo protocol and Dirichlet handling as the CSR SpMV; bit-reproducible (8 colour passes, no
 * atomics).  smfem_bench_spmv variant 5 times it.  `mesh` must stay alive while the option is on. */
int smfem_pcg_use_matrix_free(smfem_ctx *ctx, smfem_matrix *K, smfem_mesh *mesh, int enable);
/* The same operator WITHOUT ever assembling K: a matrix handle that holds the slab layout, the material and diag(K) (computed
 * from the coordinates, for the Jacobi preconditioner) but no rowptr / colind / val - 0 bytes per nonzero instead of 12, so one
 * GPU solves meshes whose assembled matrix would not fit.  Accepted by smfem_surface_mass (lattice faces; the term enters the
 * operator's beta and the diagonal), smfem_set_dirichlet[_zplanes], smfem_pcg_use_multigrid (coarse levels are assembled as usual),
 * smfem_pcg_solve, smfem_spmv_host, smfem_bench_spmv (variant 5), smfem_comm_* / the multi layer, smfem_project_nodes,
 * smfem_extract_borders, smfem_matrix_diag / _info / _clone; calls that need CSR arrays (export, reassemble, …) return
 * SMFEM_ERR_UNSUPPORTED.  `mesh` must outlive the handle. */
int smfem_matfree_operator(smfem_ctx *ctx, smfem_mesh *mesh, double Young, double nu, smfem_matrix **K_out);
int smfem_pcg_apply_preconditioner(smfem_ctx *ctx, smfem_matrix *K, const double *r, double *z);
/* Load stepping (examples/vector3D.jl:310-338: the same K̄ solved for 50 prescribed displacements d; q is exactly
 * linear in d): the NEXT smfem_pcg_solve on K starts from scale * (previous solution on the free dofs) instead of 0.
 * Typical use: set_dirichlet_zplanes(d_new); set_warm_start(d_new / d_old); pcg_solve(...) -> 0-2 iterations. */
int smfem_pcg_set_warm_start(smfem_matrix *K, double scale);

/* y = K x with host vectors of this rank's row slab (tests, manufactured right-hand sides).  Several ranks: collective, after
 * smfem_comm_connect; the ghost planes of x travel through the peer window like in a CG iteration. */
int smfem_spmv_host(smfem_ctx *ctx, smfem_matrix *K, const double *x, double *y);
/* Time `reps` back-to-back device-resident SpMVs (x = deterministic pattern, halo exchange
 * included when nranks > 1); returns average ms per SpMV measured with CUDA events on the
 * context's stream.  variant: SpMV kernel variant (0 = default). */
int smfem_bench_spmv(smfem_ctx *ctx, smfem_matrix *K, int variant, int reps, float *ms_per_spmv);
/* choose the SpMV kernel the solver uses: 0 = warp per 3 rows (default), 1 = warp per row */
int smfem_set_spmv_variant(smfem_matrix *K, int variant);
/* statistics of the last smfem_pcg_solve on this matrix */
int smfem_pcg_stats(smfem_matrix *K, float *ms_total, float *ms_spmv_est, int *iters);
/* Wait accounting of the last Jacobi-PCG solve on K (several ranks): microseconds the first CTA spent spinning on the neighbours'
 * halo flags (boundary-plane SpMV launches / the matrix-free operator's halo wait) and on the two mailbox all-reduces of an
 * iteration (r'z, r'r before the search-direction update; p'Ap before the x / r update), summed over the solve (%globaltimer).
 * Measurement only; any pointer may be NULL. */
int smfem_pcg_wait_stats(smfem_matrix *K, double *halo_us, double *allreduce_rz_us, double *allreduce_pap_us);

/* ---- multi-GPU peer window (one process per GPU) ----------------------------------------------
 * After smfem_pattern_build / smfem_assemble each rank creates its communication window (halo
 * planes of the search direction + all-reduce mailboxes), exports a CUDA IPC handle, the host side
 * exchanges the handles with torch.distributed (all_gather), and every rank maps its peers.
 * Afterwards the CG kernels write halo planes and dot-product partials directly into peer memory
 * over NVLink; no host or NCCL call is on the data path.  SMFEM_IPC_HANDLE_BYTES per rank. */
#define SMFEM_IPC_HANDLE_BYTES 64
int smfem_comm_export(smfem_ctx *ctx, smfem_matrix *K, void *handle_out);
int smfem_comm_connect(smfem_ctx *ctx, smfem_matrix *K, const void *all_handles /* nranks*64 bytes */);

/* ---- one process, several GPUs ----------------------------------------------------------------
 * The reference host is ONE Julia process (examples/vector3D.jl:266-345 `main()`), so the path must also be drivable from a
 * single thread of a single process.  smfem_init_multi creates one context per GPU (rank r of n_gpus = z-slab r of the lattice,
 * on device devices[r]; devices == NULL: 0 .. n_gpus-1) and one worker thread per GPU.  Every smfem_multi_* call below runs the
 * per-rank call of the same name on all ranks concurrently and returns when all have finished (the first failing rank's status;
 * smfem_last_error() names the rank).  The peer windows are connected inside the process with cudaDeviceEnablePeerAccess + raw
 * device pointers (smfem_comm_connect_local) - CUDA IPC handles cannot be opened by the process that exported them.  Kernels,
 * halo / all-reduce protocols and results are those of the one-process-per-GPU mode (bit-identical; tests/test_gpu_multi.py).
 * Host vectors (rhs_extra, q) and exported CSC arrays are GLOBAL: rank r reads / writes its row slab of them. */
typedef struct smfem_multi smfem_multi;
typedef struct smfem_multi_mesh smfem_multi_mesh;
typedef struct smfem_multi_matrix smfem_multi_matrix;
int smfem_init_multi(int n_gpus, const int *devices, smfem_multi **out);
int smfem_multi_destroy(smfem_multi *m);
int smfem_multi_size(smfem_multi *m, int *n_gpus);
int smfem_multi_sync(smfem_multi *m);
/* the per-rank handles behind the multi handles (any of the outputs may be NULL), for calls this section does not wrap */
int smfem_multi_rank_handles(smfem_multi *m, smfem_multi_mesh *mesh, smfem_multi_matrix *K, int rank, smfem_ctx **ctx_out,
                             smfem_mesh **mesh_out, smfem_matrix **K_out);
/* meshgrid / inflate_sphere (examples/vector3D.jl:10-130, src/PostProcess.jl:30-44): every rank generates its slab on its device */
int smfem_multi_meshgrid(smfem_multi *m, double x0, double x1, double y0, double y1, double z0, double z1, int64_t ne, int ndim,
                         smfem_multi_mesh **out);
int smfem_multi_inflate_sphere(smfem_multi *m, smfem_multi_mesh *mesh, double x0, double x1, double y0, double y1);
/* assemble_system (src/fem.jl:135-256) from a device-resident mesh / from the reference's host arrays (all ranks read the same
 * arrays concurrently, each the part its slab uses) */
int smfem_multi_assemble(smfem_multi *m, smfem_multi_mesh *mesh, int64_t ne, int ndim, int func_class, int nDof, double Young,
                         double nu, smfem_multi_matrix **K_out);
int smfem_multi_assemble_system(smfem_multi *m, const double *NodeList, const int64_t *IEN, const int64_t *ID, int64_t nNodes,
                                int64_t nEl, int nLocal, int64_t ne, int ndim, int func_class, int nDof, double Young, double nu,
                                smfem_multi_mesh **mesh_out, smfem_multi_matrix **K_out);
int smfem_multi_reassemble(smfem_multi *m, smfem_multi_mesh *mesh, smfem_multi_matrix *K, double Young, double nu);
int smfem_multi_matrix_info(smfem_multi *m, smfem_multi_matrix *K, int64_t *mrows, int64_t *ncols, int64_t *nnz);
/* K + beta*b (examples/vector3D.jl:175-264, :308), setboundaryCond (:133-173), solve (:315-322) */
int smfem_multi_surface_mass(smfem_multi *m, smfem_multi_matrix *K, smfem_multi_mesh *mesh, double beta);
int smfem_multi_set_dirichlet_zplanes(smfem_multi *m, smfem_multi_matrix *K, smfem_multi_mesh *mesh, double d);
int smfem_multi_pcg_use_multigrid(smfem_multi *m, smfem_multi_matrix *K, smfem_multi_mesh *mesh, int enable);
int smfem_multi_pcg_set_warm_start(smfem_multi *m, smfem_multi_matrix *K, double scale);
int smfem_multi_pcg_solve(smfem_multi *m, smfem_multi_matrix *K, double rtol, int maxit, const double *rhs_extra_global,
                          double *q_global, int *iters, double *relres);
int smfem_multi_pcg_stats(smfem_multi *m, smfem_multi_matrix *K, float *ms_total_max, int *iters);
/* SparseMatrixCSC(K) of the WHOLE matrix: colptr m+1, rowval / nzval nnz entries, 1-based (what sparse(E,J,V) returned, src/fem.jl:253) */
int smfem_multi_matrix_export_csc(smfem_multi *m, smfem_multi_matrix *K, int which, int64_t *colptr, int64_t *rowval, double *nzval);
int smfem_multi_matrix_free(smfem_multi *m, smfem_multi_matrix *K);
int smfem_multi_mesh_free(smfem_multi *m, smfem_multi_mesh *mesh);
/* building blocks of the in-process connection: allocate rank's window / map the peers' windows by pointer (all_K[q] = rank q's matrix) */
int smfem_comm_prepare(smfem_ctx *ctx, smfem_matrix *K);
int smfem_comm_connect_local(smfem_ctx *ctx, smfem_matrix *K, smfem_matrix *const *all_K, int n);

#ifdef __cplusplus
}
#endif
#endif /* SMEARFEM_B200_H */
