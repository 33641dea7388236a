# PAMG.jl — thin `ccall` shim from PartitionedArrays.jl objects to libpamg.so (include/pamg.h).
#
# STATUS: NOT EXECUTED.  Julia is not installed in the build image and PartitionedArrays.jl is not
# vendored in the reference snapshot (/root/reference holds README.md:1-2 and LICENSE only), so this
# file is marshalling only and every accessor name below is [RECALL-UNVERIFIED] (SURVEY.md App. A).
# All logic lives below the C ABI, where it is exercised by tests/ through the same entry points
# (parallel_amg_b200/_lib.py is the ctypes twin of this file, function for function).
#
# Usage (debug or MPI backend, one part per GPU of one node):
#   S = PAMG.setup(A)                 # A::PSparseMatrix (assembled, symmetric)  ~ setup(amg(), x, A, b)
#   PAMG.solve!(x, S, b)              # AMG-preconditioned CG to rtol            ~ solve!(x, S, b)
#   PAMG.vcycle!(z, S, r)             # one preconditioner application           ~ ldiv!(z, P, r)
#   PAMG.finalize!(S)
module PAMG

using PartitionedArrays
using SparseArrays

const lib = get(ENV, "PAMG_LIB", joinpath(@__DIR__, "..", "parallel_amg_b200", "libpamg.so"))

struct Options            # mirrors pamg_options (include/pamg.h); filled by pamg_default_options
    struct_size::Int32; eps_strength::Float64; coarse_size::Int32; max_levels::Int32
    smoother::Int32; omega_jacobi::Float64; nu_pre::Int32; nu_post::Int32; cheb_degree::Int32
    cheb_lo_frac::Float64; cheb_hi_frac::Float64; spmv_format::Int32; use_graph::Int32
    lanes_per_row::Int32; tail_rows::Int32; sell_sigma::Int32; sell_rows_per_thread::Int32; fuse_halo::Int32
end

mutable struct Setup
    ctx::Ptr{Cvoid}
    nparts::Int
    nown::Vector{Int}
end

function check(ctx, st)
    st == 0 && return
    msg = ctx == C_NULL ? "" : unsafe_string(ccall((:pamg_last_error, lib), Cstring, (Ptr{Cvoid},), ctx))
    error("pamg status $st: $msg")          # -5 (PAMG_ERR_NOTCONV) is handled by the caller
end

default_options() = (r = Ref{Options}(); ccall((:pamg_default_options, lib), Cvoid, (Ref{Options},), r); r[])

# CSR of the own rows of one part with GLOBAL 0-based column ids.  Julia's SparseMatrixCSC of a
# symmetric matrix is a valid CSR of the same matrix, so colptr/rowval are reused as rowptr/col.
function part_rows(Aloc::SparseMatrixCSC, rows, cols)
    o2g = Int64.(own_to_global(rows)) .- 1
    l2g = Int64.(local_to_global(cols)) .- 1
    nown = length(o2g)
    At = Aloc                                  # symmetric: columns of A == rows of A
    rowptr = Int64.(At.colptr[1:nown+1]) .- 1
    colgid = [l2g[j] for j in At.rowval[1:rowptr[end]]]
    (o2g, rowptr, colgid, Float64.(At.nzval[1:rowptr[end]]))
end

# nullspace: optional n x k matrix (global row order) of near-nullspace vectors, e.g. the 6 rigid-body modes from
# `nullspace_linear_elasticity`; block_size: DOFs per node (3 for 3-D elasticity).
function setup(A::PSparseMatrix; devices = nothing, opts::Options = default_options(), nullspace = nothing, block_size = 1)
    rows, cols = partition(axes(A, 1)), partition(axes(A, 2))
    np = length(rows)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(C_NULL, ccall((:pamg_create, lib), Cint, (Int32, Ref{Ptr{Cvoid}}), np, r))
    ctx = r[]
    nown = zeros(Int, np)
    # debug backend: every part is visible here.  MPI backend: gather the parts to every rank first
    # (the host setup is replicated and deterministic), then drive only the local part below.
    parts = collect(zip(collect(local_values(A)), collect(rows), collect(cols)))
    for (p, (Aloc, ri, ci)) in enumerate(parts)
        o2g, rowptr, colgid, val = part_rows(Aloc, ri, ci)
        nown[p] = length(o2g)
        check(ctx, ccall((:pamg_set_part_rows, lib), Cint,
                         (Ptr{Cvoid}, Int32, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                         ctx, p - 1, length(o2g), o2g, rowptr, colgid, val))
    end
    if nullspace !== nothing
        B = permutedims(Matrix{Float64}(nullspace))           # row-major n x k for the C side
        check(ctx, ccall((:pamg_set_near_nullspace, lib), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}),
                         ctx, block_size, size(nullspace, 2), B))
    end
    check(ctx, ccall((:pamg_setup, lib), Cint, (Ptr{Cvoid}, Ref{Options}), ctx, Ref(opts)))
    local_parts = Int32.(0:np-1)
    devs = devices === nothing ? Int32.(0:np-1) : Int32.(devices)   # one part per GPU
    check(ctx, ccall((:pamg_device_init, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}),
                     ctx, np, local_parts, devs))
    Setup(ctx, np, nown)
end

own_ptrs(v::PVector) = [pointer(o) for o in collect(own_values(v))]   # contiguous own blocks

function solve!(x::PVector, S::Setup, b::PVector; rtol = 1e-8, maxiter = 200, precond = true)
    iters = Ref{Int32}(0)
    hist = zeros(Float64, maxiter + 2)
    bp, xp = own_ptrs(b), own_ptrs(x)
    st = GC.@preserve b x ccall((:pamg_pcg, lib), Cint,
        (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Float64, Int32, Int32, Ref{Int32}, Ptr{Float64}),
        S.ctx, bp, xp, rtol, maxiter, precond ? 1 : 0, iters, hist)
    st == -5 || check(S.ctx, st)
    wait(consistent!(x))                       # ghosts of the returned PVector (host side)
    (iterations = Int(iters[]), converged = st == 0, residuals = hist[1:iters[]+1])
end

function vcycle!(z::PVector, S::Setup, r::PVector)
    rp, zp = own_ptrs(r), own_ptrs(z)
    GC.@preserve r z check(S.ctx, ccall((:pamg_vcycle, lib), Cint,
        (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}), S.ctx, rp, zp))
    z
end

# device halo exchange on a level-0 vector given as LOCAL values (own then ghost): consistent!(v) |> wait
function device_consistent!(v::PVector, S::Setup)
    lp = [pointer(l) for l in collect(local_values(v))]
    GC.@preserve v check(S.ctx, ccall((:pamg_consistent, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Ptr{Float64}}), S.ctx, 0, lp))
    v
end

function device_assemble!(v::PVector, S::Setup)
    lp = [pointer(l) for l in collect(local_values(v))]
    GC.@preserve v check(S.ctx, ccall((:pamg_assemble, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Ptr{Float64}}), S.ctx, 0, lp))
    v
end

finalize!(S::Setup) = (ccall((:pamg_destroy, lib), Cvoid, (Ptr{Cvoid},), S.ctx); S.ctx = C_NULL; nothing)

end # module
