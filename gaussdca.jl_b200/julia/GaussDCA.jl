# GaussDCA.jl -- drop-in replacement for src/GaussDCA.jl of carlobaldassi/GaussDCA.jl.
#
# Same module name, same exports (reference src/GaussDCA.jl:3), same gDCA keyword signature (:8-16)
# and printrank methods (:67-74).  Lines :18-23 of the reference (argument check, FASTA parse via
# DCAUtils, optional dedup) are kept verbatim in behaviour and stay on the host; everything from the
# encoded alignment (:24) to the ranking (:44) is ONE ccall into libgdca_b200.so (include/gdca_b200.h).
#
# Julia is not installed in the build image of this repository, so this file is checked statically
# (tests/test_host_cpu.py::test_julia_wrapper_is_consistent_with_header) and mirrored call-for-call by
# the executable Python host layer gaussdca.jl_b200/api.py.  See INTEGRATION.md.
#
# Licence: the keyword signature of gDCA, check_arguments and the three printrank methods are taken from
# GaussDCA.jl (Copyright (C) Carlo Baldassi and contributors), which is free software under the GNU General Public
# License, version 3 or (at your option) any later version (reference LICENSE.md, COPYING).  This file is therefore
# distributed under the same terms: GPL-3.0-or-later, WITHOUT ANY WARRANTY; see <https://www.gnu.org/licenses/>.
module GaussDCA

export gDCA, printrank

using LinearAlgebra, Printf
using DCAUtils: read_fasta_alignment, remove_duplicate_sequences   # host I/O only (src/GaussDCA.jl:20-23)

# ---------------------------------------------------------------------------------------------
# library handle
# ---------------------------------------------------------------------------------------------
const libgdca = get(ENV, "GDCA_B200_LIB", joinpath(@__DIR__, "..", "libgdca_b200.so"))

# gdca_status_t
const GDCA_OK = Int32(0)
const GDCA_ERR_INVALID_ARG = Int32(1)
const GDCA_ERR_Q_TOO_BIG = Int32(2)
const GDCA_ERR_NOT_SPD = Int32(3)

# gdca_stats_t (include/gdca_b200.h) -- isbits mirror, field for field
struct GdcaStats
    L::Int64
    M::Int64
    n::Int64
    q::Int32
    posdef_info::Int32
    theta::Float64
    thresh::Int64
    meff::Float64
    ident_sum::UInt64
    theta_passes::Int32
    reserved::Int32
    ms_h2d::Float32
    ms_pack::Float32
    ms_theta::Float32
    ms_weights::Float32
    ms_cov::Float32
    ms_chol::Float32
    ms_inv::Float32
    ms_score::Float32
    ms_apc::Float32
    ms_rank::Float32
    ms_d2h::Float32
    ms_total::Float32
end
GdcaStats() = GdcaStats(0, 0, 0, 0, 0, 0.0, 0, 0.0, 0, 0, 0, ntuple(_ -> 0.0f0, 12)...)

# GDCA_B200_DEVICES="0,1,2,3" (or a count, "4" = devices 0..3) selects the GPUs; default: GDCA_B200_DEVICE (one ordinal) or GPU 0.
# With more than one device the context leads a device group (gdca_create_multi): gDCA(filename) then runs on all of them --
# one host-to-device copy of Z, NVLink broadcast, sharded sweep / covariance / inversion -- and returns the same bits.
function devices_from_env()
    v = strip(get(ENV, "GDCA_B200_DEVICES", ""))
    isempty(v) && return Int32[parse(Int32, get(ENV, "GDCA_B200_DEVICE", "0"))]
    occursin(",", v) || return Int32.(0:parse(Int, v)-1)
    return Int32[parse(Int32, x) for x in split(v, ",") if !isempty(strip(x))]
end

mutable struct Context
    handle::Ptr{Cvoid}
    function Context(devices::Vector{Int32} = devices_from_env())
        h = Ref{Ptr{Cvoid}}(C_NULL)
        st = length(devices) == 1 ?
            ccall((:gdca_create, libgdca), Int32, (Ref{Ptr{Cvoid}}, Int32), h, devices[1]) :
            ccall((:gdca_create_multi, libgdca), Int32, (Ref{Ptr{Cvoid}}, Ptr{Int32}, Int32), h, devices, length(devices))
        if st != GDCA_OK
            msg = unsafe_string(ccall((:gdca_last_error, libgdca), Cstring, (Ptr{Cvoid},), C_NULL))
            error("gdca_create(devices=$devices) failed: $msg")     # no CPU fallback by design
        end
        ctx = new(h[])
        finalizer(c -> (c.handle != C_NULL && ccall((:gdca_destroy, libgdca), Cvoid, (Ptr{Cvoid},), c.handle); c.handle = C_NULL), ctx)
        return ctx
    end
end

const default_ctx = Ref{Union{Nothing,Context}}(nothing)
function context()
    default_ctx[] === nothing && (default_ctx[] = Context())
    return default_ctx[]::Context
end

last_error(ctx::Context) = unsafe_string(ccall((:gdca_last_error, libgdca), Cstring, (Ptr{Cvoid},), ctx.handle))

# ---------------------------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------------------------
function gDCA(
        filename::AbstractString;
        pseudocount::Real = 0.8,
        θ = :auto,
        max_gap_fraction::Real = 0.9,
        score::Symbol = :frob,
        min_separation::Integer = 5,
        remove_dups::Bool = false
    )

    check_arguments(filename, pseudocount, θ, max_gap_fraction, score, min_separation)

    Z = read_fasta_alignment(filename, max_gap_fraction)      # Matrix{Int8}, N x M, one sequence per column
    if remove_dups
        Z, _ = remove_duplicate_sequences(Z)
    end
    N, M = size(Z)

    # q = Int(maximum(Z)), the q >= 32 error, weights, frequencies, pseudocount, C, inv(cholesky(C)),
    # block scores, APC and the ranking (src/GaussDCA.jl:25-44) all happen inside gdca_run.
    ctx = context()
    len = ccall((:gdca_ranking_length, libgdca), Int64, (Int64, Int64), N, min_separation)
    R = Vector{Tuple{Int,Int,Float64}}(undef, len)            # 24-byte isbits rows == gdca_rank_t
    stats = Ref(GdcaStats())
    st = ccall((:gdca_run, libgdca), Int32,
               (Ptr{Cvoid}, Ptr{Int8}, Int64, Int64, Float64, Float64, Int32, Int64,
                Ptr{Tuple{Int,Int,Float64}}, Int64, Ref{GdcaStats}),
               ctx.handle, Z, N, M, θ === :auto ? -1.0 : Float64(θ), Float64(pseudocount),
               score === :DI ? Int32(1) : Int32(0), min_separation, R, len, stats)
    if st == GDCA_ERR_NOT_SPD
        throw(PosDefException(stats[].posdef_info))           # what cholesky(C) throws at src/GaussDCA.jl:34
    elseif st == GDCA_ERR_Q_TOO_BIG
        error(last_error(ctx))                                 # "parameter q=$q is too big (max 31 is allowed)"
    elseif st == GDCA_ERR_INVALID_ARG
        throw(ArgumentError(last_error(ctx)))
    elseif st != GDCA_OK
        error("libgdca_b200: " * last_error(ctx))
    end

    return R
end

function check_arguments(filename, pseudocount, θ, max_gap_fraction, score, min_separation)
    aerror(s) = throw(ArgumentError(s))
    0 <= pseudocount <= 1 ||
        aerror("invalid pseudocount value: $pseudocount (must be between 0 and 1)")
    θ == :auto || (θ isa Real && 0 <= θ <= 1) ||
        aerror("invalid θ value: $θ (must be either :auto, or a number between 0 and 1)")
    0 <= max_gap_fraction <= 1 ||
        aerror("invalid max_gap_fraction value: $max_gap_fraction (must be between 0 and 1)")
    score in [:DI, :frob] ||
        aerror("invalid score value: $score (must be either :DI or :frob)")
    min_separation >= 1 ||
        aerror("invalid min_separation value: $min_separation (must be >= 1)")
    isfile(filename) ||
        aerror("cannot open file $filename")

    return true
end

function printrank(io::IO, R::Vector{Tuple{Int,Int,Float64}})
    for I in R
        @printf(io, "%i %i %e\n", I[1], I[2], I[3])
    end
end
printrank(R::Vector{Tuple{Int,Int,Float64}}) = printrank(stdout, R)   # the reference says STDOUT (pre-1.0 name)

printrank(outfile::AbstractString, R::Vector{Tuple{Int,Int,Float64}}) = open(f->printrank(f, R), outfile, "w")

# ---- DCAUtils-shaped staged pieces (not exported, like the DCAUtils functions the reference imports at src/GaussDCA.jl:6) ----
# Same names, argument order and return tuples as the calls at src/GaussDCA.jl:28-39, each one ccall (include/gdca_b200.h).
# Julia's column-major n x n arrays are passed as they are: every matrix on this path is symmetric.

function _check(ctx::Context, st::Int32)
    st == GDCA_OK && return
    st == GDCA_ERR_INVALID_ARG && throw(ArgumentError(last_error(ctx)))
    error(last_error(ctx))
end

function compute_weighted_frequencies(Z::Matrix{Int8}, q::Integer, θ)
    N, M = size(Z)
    q == maximum(Z) || throw(ArgumentError("q=$q does not match maximum(Z)"))
    n = (q - 1) * N
    Pi_true = Vector{Float64}(undef, n); Pij_true = Matrix{Float64}(undef, n, n); W = Vector{Float64}(undef, M)
    Meff = Ref(0.0); θused = Ref(0.0); qout = Ref(Int32(0))
    ctx = context()
    _check(ctx, ccall((:gdca_compute_weighted_frequencies, libgdca), Int32,
                      (Ptr{Cvoid}, Ptr{Int8}, Int64, Int64, Float64, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ptr{Float64},
                       Ref{Float64}, Ref{Int32}),
                      ctx.handle, Z, N, M, θ === :auto ? -1.0 : Float64(θ), Pi_true, Pij_true, Meff, W, θused, qout))
    return Pi_true, Pij_true, Meff[], W
end

function add_pseudocount(Pi_true::Vector{Float64}, Pij_true::Matrix{Float64}, pc::Float64, q::Integer)
    n = length(Pi_true)
    Pi = similar(Pi_true); Pij = similar(Pij_true)
    ctx = context()
    _check(ctx, ccall((:gdca_add_pseudocount, libgdca), Int32,
                      (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Float64, Ptr{Float64}, Ptr{Float64}),
                      ctx.handle, Pi_true, Pij_true, n, q, pc, Pi, Pij))
    return Pi, Pij
end

function compute_C(Pi::Vector{Float64}, Pij::Matrix{Float64})        # src/GaussDCA.jl:76
    n = length(Pi)
    C = similar(Pij)
    ctx = context()
    _check(ctx, ccall((:gdca_compute_C, libgdca), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
                      ctx.handle, Pi, Pij, n, C))
    return C
end

function _score(mJ::Matrix{Float64}, C, q::Integer, which::Int32)
    n = size(mJ, 1)
    N = div(n, q - 1)
    S = Matrix{Float64}(undef, N, N)
    ctx = context()
    _check(ctx, ccall((:gdca_score, libgdca), Int32,
                      (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Int32, Ptr{Float64}),
                      ctx.handle, mJ, C === nothing ? C_NULL : C, n, q, which, S))
    return S
end
compute_FN(mJ::Matrix{Float64}, q::Integer) = _score(mJ, nothing, q, Int32(0))                         # src/GaussDCA.jl:39
compute_DI_gauss(mJ::Matrix{Float64}, C::Matrix{Float64}, q::Integer) = _score(mJ, C, q, Int32(1))     # src/GaussDCA.jl:37

end # module
