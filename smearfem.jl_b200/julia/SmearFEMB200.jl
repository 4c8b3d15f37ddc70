# SmearFEMB200.jl -- the reference-side binding: a thin `ccall` shim over libsmearfem_b200.so that keeps
# smearFEM.jl's call surface (src/smearFEM.jl:3-4 exports + the example-script functions of
# examples/vector3D.jl).  Julia is not installed in the build image, so this file is NOT exercised by
# the test-suite; the Python mirror (smearfem.jl_b200/__init__.py) binds the very same symbols with the
# same argument order and is what the parity tests run.  Keep the two in sync with include/smearfem_b200.h.
#
# Usage (drop-in for the hot path).  smearFEM itself exports assemble_system / gaussian_quadrature / basis_function / greet_fem
# (src/smearFEM.jl:3-4), so this module EXPORTS NOTHING that clashes: name what you take from it -- an explicit import wins over
# smearFEM's implicit exports, and PostProcess (write_scene, fit_curve, plots) keeps coming from smearFEM unchanged:
#     using smearFEM
#     using SmearFEMB200: assemble_system, apply_boundary_conditions, setboundaryCond, solve, inflate_sphere, meshgrid
#     K  = assemble_system(ne, NodeList, IEN, ndim, "Q1", nDof, ID, Young, ν)      # device-resident
#     b  = apply_boundary_conditions(ne, NodeList, IEN, IEN_top, IEN_btm, ndim, "Q1", ID)
#     K̄  = K + β*b                                # a new device matrix (K stays K); add_surface_mass!(K, β*b) is the in-place form
#     q_d, C = setboundaryCond(NodeList, ne, ndim, "Q1", d, nDof)
#     q  = solve(K̄, q_d, C)                 # replaces inv(Matrix(C'K̄C)) * C'(-K̄ q_d); q = q_d + C q_f
#     SparseMatrixCSC(K)                    # materialise Julia's CSC when really needed
module SmearFEMB200

using SparseArrays, LinearAlgebra

# non-clashing names only (see the usage note above); everything else is reached as SmearFEMB200.name or by explicit import
export B200SparseMatrix, SurfaceMatrix, solve
export use_multigrid!, use_matrix_free!, project_nodes, extract_borders_device, element_colors

const LIB = get(ENV, "SMEARFEM_B200_LIB", joinpath(@__DIR__, "..", "libsmearfem_b200.so"))
const Q1, Q2 = Cint(1), Cint(2)

struct SmfemError <: Exception
    code::Cint
    msg::String
end
Base.showerror(io::IO, e::SmfemError) = print(io, "smearfem_b200 [status $(e.code)]: ", e.msg)

last_error() = unsafe_string(ccall((:smfem_last_error, LIB), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : throw(SmfemError(rc, last_error()))
fclass(s::AbstractString) = s == "Q1" ? Q1 : s == "Q2" ? Q2 : throw(ArgumentError("FunctionClass $s"))

# ---- context: this module drives ONE GPU (device SMEARFEM_B200_DEVICE, default 0).  A Julia process that wants all GPUs of the box
# uses the `Multi` submodule at the end of this file (smfem_init_multi: one process, n GPUs).  Launching several Julia processes
# with RANK / WORLD_SIZE is rejected here: the peer-window handle exchange of that mode needs a launcher-side all-gather
# (smearfem.jl_b200/distributed.py does it with torch.distributed) that this shim does not carry.
mutable struct Context
    h::Ptr{Cvoid}
end
const CTX = Ref{Union{Nothing,Context}}(nothing)
function context()
    if CTX[] === nothing
        parse(Int, get(ENV, "WORLD_SIZE", "1")) == 1 ||
            error("SmearFEMB200: WORLD_SIZE > 1 is not supported by the Julia shim; use SmearFEMB200.Multi (one process, n GPUs)")
        h = Ref{Ptr{Cvoid}}(C_NULL)
        dev = parse(Cint, get(ENV, "SMEARFEM_B200_DEVICE", "0"))
        rank = Cint(0); nr = Cint(1)
        check(ccall((:smfem_init, LIB), Cint, (Cint, Cint, Cint, Ptr{Ptr{Cvoid}}), dev, rank, nr, h))
        c = Context(h[])
        finalizer(c -> ccall((:smfem_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), c)
        CTX[] = c
    end
    return CTX[]
end

greet_fem() = println("Hello, I am the FEM module")          # src/fem.jl:4-6

# ---- src/fem.jl:21-31 -------------------------------------------------------------------------------------
function gaussian_quadrature(a, b, nGaussPoints=2)
    ξ = zeros(max(nGaussPoints, 1)); w = zeros(max(nGaussPoints, 1))
    check(ccall((:smfem_gaussian_quadrature, LIB), Cint, (Cdouble, Cdouble, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                a, b, nGaussPoints, ξ, w))
    return ξ, w
end

# ---- src/fem.jl:48-114 ------------------------------------------------------------------------------------
function basis_function(ξ, η=nothing, ζ=nothing, FunctionClass="Q1")
    ndim = isnothing(η) ? 1 : isnothing(ζ) ? 2 : 3
    N = zeros(9); ΔN = zeros(27); nn = Ref{Cint}(0)
    check(ccall((:smfem_basis_function, LIB), Cint,
                (Cint, Cint, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
                ndim, fclass(FunctionClass), ξ, something(η, 0.0), something(ζ, 0.0), N, ΔN, nn))
    n = Int(nn[])
    ndim == 1 && return N[1:n], reshape(ΔN[1:2], 1, 2)                 # src/fem.jl:75 (1x2 row)
    return N[1:n], reshape(ΔN[1:n*ndim], n, ndim)
end

# ---- mesh handle ------------------------------------------------------------------------------------------
mutable struct Mesh
    h::Ptr{Cvoid}
end
free!(m::Mesh) = (m.h != C_NULL && ccall((:smfem_mesh_free, LIB), Cint, (Ptr{Cvoid},), m.h); m.h = C_NULL)

function mesh_from_host(NodeList::Matrix{Float64}, IEN::Matrix{Int64}, ID, ndim, nDof, ne)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    idp = ID === nothing ? Ptr{Int64}(C_NULL) : pointer(ID)
    GC.@preserve NodeList IEN ID begin
        check(ccall((:smfem_mesh_from_host, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Cint, Cint, Cint, Int64, Ptr{Ptr{Cvoid}}),
                    context().h, NodeList, IEN, idp, size(NodeList, 2), size(IEN, 1), size(IEN, 2), ndim,
                    ID === nothing ? nDof : size(ID, 2), ne, h))
    end
    m = Mesh(h[]); finalizer(free!, m); return m
end

# ---- examples/vector3D.jl:10-130 (3-D branch; coordinates generated on the device) -------------------------
function meshgrid(x0, x1, y0, y1, z0, z1, ne, ndim)
    ndim == 3 || error("SmearFEMB200.meshgrid: use smearFEM's host meshgrid for ndim = 2")
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:smfem_meshgrid, LIB), Cint,
                (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Int64, Cint, Ptr{Ptr{Cvoid}}),
                context().h, x0, x1, y0, y1, z0, z1, ne, ndim, h))
    m = Mesh(h[])
    nN, nEl = (ne + 1)^3, ne^3
    NodeList = zeros(3, nN); IEN = zeros(Int64, nEl, 8); ID = zeros(Int64, nN, 3)
    IEN_top = zeros(Int64, ne^2, 4); IEN_btm = zeros(Int64, ne^2, 4)
    check(ccall((:smfem_mesh_export, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
                context().h, m.h, NodeList, IEN, ID, IEN_top, IEN_btm))
    free!(m)
    Border = Int[]; Bottom = Int[]; Top = Int[]                        # :76-82 (visualisation lists)
    mm = 1
    for k in 1:ne+1, j in 1:ne+1, i in 1:ne+1
        if i == 1 || i == ne + 1 || j == 1 || j == ne + 1
            push!(Border, mm)
        elseif k == 1
            push!(Bottom, mm)
        elseif k == ne + 1
            push!(Top, mm)
        end
        mm += 1
    end
    return NodeList, IEN, ID, IEN_top, IEN_btm, [Border, Bottom, Top]
end

# ---- src/PostProcess.jl:30-44 (in place) ----------------------------------------------------------------------
function inflate_sphere(NodeList::Matrix{Float64}, x0, x1, y0, y1)
    check(ccall((:smfem_inflate_sphere_host, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Int64, Cdouble, Cdouble, Cdouble, Cdouble),
                context().h, NodeList, size(NodeList, 1), size(NodeList, 2), x0, x1, y0, y1))
    return NodeList
end

# ---- src/fem.jl:135-256 ------------------------------------------------------------------------------------
mutable struct B200SparseMatrix            # stands in for SparseMatrixCSC{Float64,Int64}; lives in HBM
    h::Ptr{Cvoid}
    mesh::Mesh
end
free!(K::B200SparseMatrix) = (K.h != C_NULL && ccall((:smfem_matrix_free, LIB), Cint, (Ptr{Cvoid},), K.h); K.h = C_NULL)

function assemble_system(ne, NodeList, IEN, ndim, FunctionClass="Q1", nDof=1, ID=nothing, Young=1, ν=0.3)
    NL = Matrix{Float64}(NodeList); IENm = Matrix{Int64}(IEN)
    IDm = nDof > 1 ? Matrix{Int64}(ID) : nothing
    idp = IDm === nothing ? Ptr{Int64}(C_NULL) : pointer(IDm)
    mh = Ref{Ptr{Cvoid}}(C_NULL); kh = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve NL IENm IDm begin      # one call: transfers, lattice check and assembly overlap inside the library
        check(ccall((:smfem_assemble_system, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Cint, Int64, Cint, Cint, Cint, Cdouble, Cdouble,
                     Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}),
                    context().h, NL, IENm, idp, size(NL, 2), size(IENm, 1), size(IENm, 2), ne, ndim, fclass(FunctionClass), nDof,
                    Young, ν, mh, kh))
    end
    mesh = Mesh(mh[]); finalizer(free!, mesh)
    K = B200SparseMatrix(kh[], mesh); finalizer(free!, K); return K
end

function info(K::B200SparseMatrix)
    m = Ref{Int64}(0); n = Ref{Int64}(0); nz = Ref{Int64}(0); row0 = Ref{Int64}(0); nrl = Ref{Int64}(0); nzl = Ref{Int64}(0)
    # (no splatting here: ccall is a special form, its argument count must be visible in the syntax)
    check(ccall((:smfem_matrix_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
                K.h, m, n, nz, row0, nrl, nzl))
    return (m=m[], n=n[], nnz=nz[], row0=row0[], nrows_local=nrl[], nnz_local=nzl[])
end
Base.size(K::B200SparseMatrix) = (i = info(K); (Int(i.m), Int(i.n)))
SparseArrays.nnz(K::B200SparseMatrix) = Int(info(K).nnz)

function SparseArrays.SparseMatrixCSC(K::B200SparseMatrix; which=0)      # what `sparse(E,J,V)` returned (src/fem.jl:253)
    i = info(K)
    colptr = zeros(Int64, i.nrows_local + 1); rowval = zeros(Int64, i.nnz_local); nzval = zeros(i.nnz_local)
    check(ccall((:smfem_matrix_export_csc, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}),
                context().h, K.h, which, colptr, rowval, nzval))
    return SparseMatrixCSC(Int(i.m), Int(i.nrows_local), colptr, rowval, nzval)
end

# ---- examples/vector3D.jl:175-264 and :308 ----------------------------------------------------------------------
struct SurfaceMatrix                         # b = ∫ NᵀN over top ∪ bottom, kept lazy
    mesh::Mesh                               # the NodeList apply_boundary_conditions was given (examples/vector3D.jl:306), on the device
    IEN_top::Matrix{Int64}
    IEN_btm::Matrix{Int64}
    β::Float64
end
function apply_boundary_conditions(ne, NodeList, IEN, IEN_top, IEN_btm, ndim, FunctionClass, ID, nDof=3)
    ndim == 3 || error("apply_boundary_conditions: only the 3-D branch of the reference is executable")
    # b is integrated over THIS NodeList (which need not be the one K was assembled on)
    m = mesh_from_host(Matrix{Float64}(NodeList), Matrix{Int64}(IEN), ID === nothing ? nothing : Matrix{Int64}(ID), ndim, nDof, ne)
    return SurfaceMatrix(m, Matrix{Int64}(IEN_top), Matrix{Int64}(IEN_btm), 1.0)
end
Base.:*(β::Number, b::SurfaceMatrix) = SurfaceMatrix(b.mesh, b.IEN_top, b.IEN_btm, b.β * β)
function clone(K::B200SparseMatrix)                                        # device-side copy: pattern, values, diagonal
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:smfem_matrix_clone, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), context().h, K.h, h))
    K2 = B200SparseMatrix(h[], K.mesh); finalizer(free!, K2); return K2
end
function add_surface_mass!(K::B200SparseMatrix, b::SurfaceMatrix)          # K += β*b in place (saves the copy of `+`)
    check(ccall((:smfem_surface_mass, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Int64, Cdouble, Cint),
                context().h, K.h, b.mesh.h, b.IEN_top, b.IEN_btm, size(b.IEN_top, 1), b.β, 0))
    return K
end
# K += β*b over the lattice's own top / bottom faces (no face lists; also valid for a matrix-free operator)
function add_surface_mass!(K::B200SparseMatrix, β::Real)
    check(ccall((:smfem_surface_mass, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Int64, Cdouble, Cint),
                context().h, K.h, K.mesh.h, C_NULL, C_NULL, 0, Float64(β), 0))
    return K
end
# K̄ = K + β*b (examples/vector3D.jl:308): a NEW device matrix, K stays K as in the reference
Base.:+(K::B200SparseMatrix, b::SurfaceMatrix) = add_surface_mass!(clone(K), b)

# ---- examples/vector3D.jl:133-173 (host data preparation, verbatim semantics) ------------------------------------
function setboundaryCond(NodeList, ne, ndim, FunctionClass, d, nDof=1)
    q_d = zeros(nDof * (ne + 1)^ndim, 1)
    C = sparse(I, ndim * (ne + 1)^ndim, ndim * (ne + 1)^ndim)
    rCol = Int[]
    for nNode in 1:size(NodeList, 2)
        z = NodeList[3, nNode]
        if z == 0
            q_d[3*nNode] = 0; push!(rCol, 3 * nNode)
        elseif z == 1
            q_d[3*nNode] = -d; push!(rCol, 3 * nNode)
        end
    end
    return q_d, C[:, setdiff(1:size(C, 2), rCol)]
end

# ---- examples/vector3D.jl:315-322 ---------------------------------------------------------------------------------
function solve(K̄::B200SparseMatrix, q_d, C; rtol=1e-12, maxit=20000, warm_scale=0.0)
    # warm_scale != 0: start from warm_scale * previous solution (load stepping, examples/vector3D.jl:310-338: q ∝ d)
    warm_scale != 0 && check(ccall((:smfem_pcg_set_warm_start, LIB), Cint, (Ptr{Cvoid}, Cdouble), K̄.h, warm_scale))
    ndof = size(C, 1)
    free = rowvals(C)                                   # C = I[:, free]
    fixed = Int64.(setdiff(1:ndof, free)); vals = Float64.(q_d[fixed])
    check(ccall((:smfem_set_dirichlet, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Cdouble}, Int64),
                context().h, K̄.h, fixed, vals, length(fixed)))
    q = zeros(info(K̄).nrows_local); it = Ref{Cint}(0); rel = Ref{Cdouble}(0)
    check(ccall((:smfem_pcg_solve, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Ptr{Cdouble}),
                context().h, K̄.h, rtol, maxit, C_NULL, q, it, rel))
    return q
end

# ---- opt-in: CG preconditioned by a geometric multigrid V-cycle for the following solve(K̄, ...) calls (hex lattice, one GPU)
function use_multigrid!(K̄::B200SparseMatrix, enable::Bool=true)
    check(ccall((:smfem_pcg_use_multigrid, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint), context().h, K̄.h, K̄.mesh.h, enable))
    return K̄
end

# ---- opt-in: apply K̄ matrix-free from the mesh coordinates inside the following solves (hex lattice, nDof = 3)
function use_matrix_free!(K̄::B200SparseMatrix, enable::Bool=true)
    check(ccall((:smfem_pcg_use_matrix_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint), context().h, K̄.h, K̄.mesh.h, enable))
    return K̄
end

# ---- the lattice operator WITHOUT an assembled matrix (diag(K) only): solve / use_multigrid! / add_surface_mass!(K, β) work on it,
#      SparseMatrixCSC(K) and K + β*b (explicit face lists) do not
function matrix_free_operator(NodeList, IEN, ID, ne, Young, ν)
    mesh = mesh_from_host(Matrix{Float64}(NodeList), Matrix{Int64}(IEN), Matrix{Int64}(ID), 3, 3, ne)
    kh = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:smfem_matfree_operator, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Ptr{Cvoid}}), context().h, mesh.h, Young, ν, kh))
    K = B200SparseMatrix(kh[], mesh); finalizer(free!, K); return K
end

# ---- extract_borders(NodeList_new, CameraMatrix, BorderNodesList, state, ne) (src/PostProcess.jl:60-117) on the device, with
#      NodeList_new = NodeList + motion of the last solve; returns (BorderPoints, SideNodes2D) like the reference
function extract_borders_device(K̄::B200SparseMatrix, CameraMatrix, BorderNodesList, state::AbstractString, ne=nothing)
    state in ("init", "update") || error("extract_borders: state must be \"init\" or \"update\"")
    state == "init" && ne === nothing && error("Number of elements must be provided")
    ids = Vector{Int64}(vec(BorderNodesList[1])); cam = Matrix{Float64}(CameraMatrix); n = length(ids)
    cap = state == "update" ? n : 2 * (ne + 1) + 2 * (n ÷ (ne + 1)) + 2
    border = zeros(2, cap); side = zeros(2, n); nb = Ref{Int64}(0)
    check(ccall((:smfem_extract_borders, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Cdouble}, Cint, Int64, Ptr{Cdouble}, Int64, Ptr{Int64}, Ptr{Cdouble}),
                context().h, K̄.mesh.h, K̄.h, ids, n, cam, state == "init" ? 0 : 1, ne === nothing ? -1 : ne, border, cap, nb, side))
    return border[:, 1:nb[]], side
end

# ---- examples/vector3D.jl:325-329 on the device: NodeList_new[:, ids] and back_project(NodeList_new[:, ids], CameraMatrix)
#      (src/PostProcess.jl:131-152) with the displacement of the last solve; feed the result to the unchanged
#      extract_borders / fit_curve in place of their own back_project call
function project_nodes(K̄::B200SparseMatrix, ids, CameraMatrix)
    idv = Vector{Int64}(vec(ids)); cam = Matrix{Float64}(CameraMatrix)
    p3 = zeros(3, length(idv)); p2 = zeros(2, length(idv))
    check(ccall((:smfem_project_nodes, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                context().h, K̄.mesh.h, K̄.h, idv, length(idv), cam, p3, p2))
    return p3, p2
end

# ---- element colouring of a general mesh (deterministic scatter): number of colours and elements per colour
function element_colors(K::B200SparseMatrix)
    nc = Ref{Cint}(0); sizes = zeros(Int64, 64)
    check(ccall((:smfem_mesh_colors, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Ptr{Int64}), context().h, K.mesh.h, nc, sizes))
    return Int(nc[]), sizes[1:max(nc[], 0)]
end

# ---- one Julia process, n GPUs (smfem_init_multi): the reference's main() shape, K distributed in z-slabs ----------------------
module Multi
using ..SmearFEMB200: LIB, check, fclass
mutable struct MultiContext
    h::Ptr{Cvoid}
    n::Int
end
function MultiContext(n_gpus::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:smfem_init_multi, LIB), Cint, (Cint, Ptr{Cint}, Ptr{Ptr{Cvoid}}), n_gpus, C_NULL, h))
    c = MultiContext(h[], n_gpus)
    finalizer(c -> ccall((:smfem_multi_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), c)
    return c
end
mutable struct MultiMatrix
    ctx::MultiContext
    h::Ptr{Cvoid}
    mesh::Ptr{Cvoid}
end
# assemble_system(ne, NodeList, IEN, ndim, FunctionClass, nDof, ID, Young, ν) with the reference's host arrays on all GPUs
function assemble_system(c::MultiContext, ne, NodeList, IEN, ndim, FunctionClass="Q1", nDof=1, ID=nothing, Young=1, ν=0.3)
    NL = Matrix{Float64}(NodeList); IENm = Matrix{Int64}(IEN); IDm = nDof > 1 ? Matrix{Int64}(ID) : nothing
    idp = IDm === nothing ? Ptr{Int64}(C_NULL) : pointer(IDm)
    mh = Ref{Ptr{Cvoid}}(C_NULL); kh = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve NL IENm IDm begin
        check(ccall((:smfem_multi_assemble_system, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Cint, Int64, Cint, Cint, Cint, Cdouble, Cdouble,
                     Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}),
                    c.h, NL, IENm, idp, size(NL, 2), size(IENm, 1), size(IENm, 2), ne, ndim, fclass(FunctionClass), nDof, Young, ν, mh, kh))
    end
    return MultiMatrix(c, kh[], mh[])
end
add_surface_mass!(K::MultiMatrix, β) = (check(ccall((:smfem_multi_surface_mass, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble), K.ctx.h, K.h, K.mesh, β)); K)
use_multigrid!(K::MultiMatrix, on::Bool=true) = (check(ccall((:smfem_multi_pcg_use_multigrid, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint), K.ctx.h, K.h, K.mesh, on)); K)
# the example's Dirichlet data (z = 0: u_z = 0, z = 1: u_z = -d; examples/vector3D.jl:133-173) and solve (:315-322); q is global
function solve(K::MultiMatrix, d; rtol=1e-12, maxit=20000)
    check(ccall((:smfem_multi_set_dirichlet_zplanes, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble), K.ctx.h, K.h, K.mesh, d))
    m = Ref{Int64}(0); n = Ref{Int64}(0); nnz = Ref{Int64}(0)
    check(ccall((:smfem_multi_matrix_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), K.ctx.h, K.h, m, n, nnz))
    q = zeros(m[]); it = Ref{Cint}(0); rel = Ref{Cdouble}(0)
    check(ccall((:smfem_multi_pcg_solve, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Ptr{Cdouble}),
                K.ctx.h, K.h, rtol, maxit, C_NULL, q, it, rel))
    return q
end
end # module Multi

end # module
