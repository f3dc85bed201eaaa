# TlpB200.jl -- Julia glue for the B200 KKT backend (UNTESTED in the build container: no Julia runtime
# there; the same C ABI is exercised exhaustively from Python, see tests/ and INTEGRATION.md).
#
# Structure mirrors the in-tree backends, e.g. src/KKT/Cholmod/cholmod.jl:1-65, spd.jl, sqd.jl:
# a Backend tag, a solver struct, and methods for setup / update! / solve! / backend / linear_system.
#
#   using Tulip, TlpB200
#   m = Tulip.Model{Float64}()
#   Tulip.set_parameter(m, "KKT_Backend", TlpB200.Backend())          # tulip_julia_api.jl:221-222
#   Tulip.set_parameter(m, "KKT_System",  Tulip.KKT.K1())              # or K2()
module TlpB200

using LinearAlgebra
using SparseArrays

import Tulip
import Tulip.KKT: AbstractKKTBackend, AbstractKKTSolver, AbstractKKTSystem, K1, K2
import Tulip.KKT: setup, update!, solve!, backend, linear_system

const libtlpb200 = get(ENV, "TLPB200_LIB", joinpath(@__DIR__, "..", "libtlpb200.so"))

# error codes of include/tlpb200.h
const OK, NOT_POSDEF, OOM, BAD_ARG, CUDA_ERR, INTERNAL = 0, 1, 2, 3, 4, 5

"""
    Backend(; device=0, ordering=1, piece_width=128, small_elems=4096, use_graph=true, ozaki_ncol=0, ...)

B200 (sm_100a) supernodal direct solver for the K1 / K2 systems.  Options are fields of the tag,
like `TlpKrylov.Backend` (src/KKT/Krylov/krylov.jl:41-44).
"""
Base.@kwdef struct Backend <: AbstractKKTBackend
    device::Int32 = 0
    ordering::Int32 = 1
    piece_width::Int32 = 128
    small_elems::Int32 = 4096
    relax_always::Int32 = 8
    use_graph::Bool = true
    dense_col_threshold::Int32 = 0   # K1 dense-column Schur path: 0 auto, < 0 off
    dense_solve_ncol::Int32 = 0      # supernodes with >= this many columns use the dense-solve sweeps (0 = 384)
    ozaki_ncol::Int32 = 0            # K1: supernodes with >= this many columns run their far Schur updates on the
                                     # tcgen05 int8 tensor-core path (0 = 1024, < 0 = FP64 DMMA path everywhere)
    refine_steps::Int32 = 0          # iterative-refinement steps inside solve! (0 = the reference's single solve)
end

# layout must match `tlpb200_options`
struct COptions
    ordering::Int32; device::Int32; piece_width::Int32; small_elems::Int32
    relax_always::Int32; use_graph::Int32; analyze_only::Int32
    rank::Int32; nranks::Int32; dense_col_threshold::Int32
    dense_solve_ncol::Int32; ozaki_ncol::Int32; refine_steps::Int32
    reserved::NTuple{3,Int32}
end

mutable struct B200Solver{S<:AbstractKKTSystem} <: AbstractKKTSolver{Float64}
    m::Int
    n::Int
    A::SparseMatrixCSC{Float64,Int}     # borrowed for the solver's lifetime (spd.jl:19)
    handle::Ptr{Cvoid}

    function B200Solver{S}(A::SparseMatrixCSC{Float64,Int}, sys::Int, b::Backend) where {S}
        m, n = size(A)
        opt = Ref(COptions(b.ordering, b.device, b.piece_width, b.small_elems, b.relax_always,
                           b.use_graph ? 1 : 0, 0, 0, 1, b.dense_col_threshold, b.dense_solve_ncol, b.ozaki_ncol,
                           b.refine_steps, ntuple(_ -> Int32(0), 3)))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:tlpb200_create, libtlpb200), Cint,
                   (Ref{Ptr{Cvoid}}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Cint, Cint, Ref{COptions}),
                   h, m, n, A.colptr, A.rowval, A.nzval, 1, sys, opt)     # index_base = 1
        kkt = new{S}(m, n, A, h[])
        # the KKT API has no close(): release device memory from a finalizer (SURVEY 8b)
        finalizer(k -> (k.handle != C_NULL && ccall((:tlpb200_destroy, libtlpb200), Cvoid, (Ptr{Cvoid},), k.handle); k.handle = C_NULL), kkt)
        rc == OK || _throw(rc, kkt)
        return kkt
    end
end

_errmsg(kkt) = unsafe_string(ccall((:tlpb200_last_error, libtlpb200), Cstring, (Ptr{Cvoid},), kkt.handle))

function _throw(rc, kkt)
    # only PosDefException / ZeroPivotException trigger the regularisation bump (HSD/step.jl:40);
    # OutOfMemoryError -> Trm_MemoryLimit (HSD.jl:327); anything else aborts the solve (HSD.jl:333-335)
    rc == NOT_POSDEF && throw(PosDefException(0))
    rc == OOM && throw(OutOfMemoryError())
    rc == BAD_ARG && throw(DimensionMismatch(_errmsg(kkt)))
    error("TlpB200 error $rc: $(_errmsg(kkt))")
end

backend(::B200Solver) = unsafe_string(ccall((:tlpb200_backend_name, libtlpb200), Cstring, ()))
linear_system(::B200Solver{K1}) = "Normal equations (K1)"
linear_system(::B200Solver{K2}) = "Augmented system (K2)"

# src/KKT/Cholmod/cholmod.jl:65: convert whatever matrix type to SparseMatrixCSC
setup(A, system::AbstractKKTSystem, b::Backend) = setup(convert(SparseMatrixCSC{Float64,Int}, A), system, b)
setup(A::SparseMatrixCSC{Float64,Int}, ::K1, b::Backend) = B200Solver{K1}(A, 1, b)
setup(A::SparseMatrixCSC{Float64,Int}, ::K2, b::Backend) = B200Solver{K2}(A, 2, b)
# KKTOptions.System defaults to DefaultKKTSystem (KKT.jl:51), "currently equivalent to K2" (KKT.jl:134-141); without this
# method the generic convert fallback above would call itself for ever on a SparseMatrixCSC{Float64,Int} (ADVICE r1).
# The Python mirror maps Default to K2 in the same way (kkt.py).
setup(A::SparseMatrixCSC{Float64,Int}, ::Tulip.KKT.DefaultKKTSystem, b::Backend) = setup(A, K2(), b)
setup(A::SparseMatrixCSC{Float64,Int}, system::AbstractKKTSystem, ::Backend) =
    error("TlpB200: unsupported KKT system $(typeof(system)); use K1() or K2()")

function update!(kkt::B200Solver, θ::AbstractVector{Float64}, regP::AbstractVector{Float64}, regD::AbstractVector{Float64})
    m, n = kkt.m, kkt.n
    # same checks and messages as spd.jl:26-34
    length(θ) == n || throw(DimensionMismatch("length(θ)=$(length(θ)) but KKT solver has n=$n."))
    length(regP) == n || throw(DimensionMismatch("length(regP)=$(length(regP)) but KKT solver has n=$n"))
    length(regD) == m || throw(DimensionMismatch("length(regD)=$(length(regD)) but KKT solver has m=$m"))
    bad = Ref{Int64}(-1)
    rc = ccall((:tlpb200_update, libtlpb200), Cint,
               (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}),
               kkt.handle, θ, regP, regD, bad)
    rc == OK || _throw(rc, kkt)
    return nothing
end

function solve!(dx::Vector{Float64}, dy::Vector{Float64}, kkt::B200Solver, ξp::Vector{Float64}, ξd::Vector{Float64})
    rc = ccall((:tlpb200_solve, libtlpb200), Cint,
               (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int64, Int64),
               kkt.handle, dx, dy, ξp, ξd, 1, kkt.n, kkt.m)
    rc == OK || _throw(rc, kkt)
    return nothing
end

end # module
