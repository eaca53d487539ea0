# reo_ccall.jl -- drop-in replacement for `identify_degs` (src/RankCompV3.jl:339-438) that calls
# libreo_cuda.so (include/reo.h) through ccall.  Everything else in RankCompV3.jl stays as it is.
#
# NOTE: Julia is not installed in the build image, so this file is delivered as the binding a maintainer
# would add; the identical ABI is exercised by the Python ctypes harness (rankcompv3.jl_b200/api.py), by the plain-C
# caller examples/reo_driver.c and by tests/test_gpu_parity.py.  The input matrix is ordinary (pageable) Julia memory:
# the library pipelines it through pinned bounce buffers (bench.py reports that path as e2e.pageable).
#
# Usage inside the package:   include("reo_ccall.jl")   after the original definition of identify_degs,
# or replace the body of identify_degs with `return identify_degs_cuda(...)`.

const LIBREO = get(ENV, "LIBREO_CUDA", "libreo_cuda.so")

const REO_OK, REO_ERR_DIM, REO_ERR_ARG, REO_ERR_BOUNDS = 0, -1, -2, -6
const REO_MAX_ITER_LOG = 256

struct ReoStats
    iters_done::Int32
    converged::Int32
    n_deg::NTuple{REO_MAX_ITER_LOG,Int32}
    n_ref::NTuple{REO_MAX_ITER_LOG,Int32}
    rank_bits::Int32
    sample_words::Int32
    compares::Int64
    ms_stage::Float64
    ms_pairs::Float64
    ms_stats::Float64
    ms_total::Float64
    ms_wall::Float64
    pair_launches::Int32
    kernel_launches::Int32
    ordered_triples::Int64
    planes_per_word::Float64
end

const REO_OUT_PINNED = UInt32(2)

# Page-locked output block from reo_host_alloc: the library DMAs the results straight into it (REO_OUT_PINNED); the
# finalizer hands the block back with reo_host_free when the Julia array is collected.
function reo_pinned_array(::Type{T}, dims::Integer...) where {T}
    n = max(prod(dims) * sizeof(T), 1)
    p = ccall((:reo_host_alloc, LIBREO), Ptr{Cvoid}, (Csize_t,), n)
    p == C_NULL && throw(OutOfMemoryError())
    a = unsafe_wrap(Array, Ptr{T}(p), dims; own = false)
    finalizer(x -> ccall((:reo_host_free, LIBREO), Cvoid, (Ptr{Cvoid},), pointer(x)), a)
    return a
end

reo_dtype(::Type{Int64}) = 0; reo_dtype(::Type{Float64}) = 1
reo_dtype(::Type{Int32}) = 2; reo_dtype(::Type{Float32}) = 3

mutable struct ReoHandle
    ptr::Ptr{Cvoid}
    function ReoHandle(devices::Vector{Cint} = Cint[0]; seed::UInt64 = UInt64(0))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:reo_create, LIBREO), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}, UInt64, UInt32),
                   out, length(devices), devices, seed, 0)
        rc == REO_OK || error(unsafe_string(ccall((:reo_last_error, LIBREO), Cstring, (Ptr{Cvoid},), C_NULL)))
        h = new(out[])
        finalizer(x -> ccall((:reo_destroy, LIBREO), Cint, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

const REO_HANDLE = Ref{Union{Nothing,ReoHandle}}(nothing)
reo_handle() = (REO_HANDLE[] === nothing && (REO_HANDLE[] = ReoHandle()); REO_HANDLE[])

function reo_throw(h::ReoHandle, rc::Integer)
    msg = unsafe_string(ccall((:reo_last_error, LIBREO), Cstring, (Ptr{Cvoid},), h.ptr))
    rc == REO_ERR_DIM    && throw(DimensionMismatch(msg))          # src:355-356
    rc == REO_ERR_ARG    && throw(ArgumentError(msg))
    rc == REO_ERR_BOUNDS && throw(BoundsError())                   # src:411 for r <= 10
    error("libreo_cuda ($rc): $msg")
end

# Same signature and return value as identify_degs (src:339-350, 437).
function identify_degs_cuda(data::AbstractMatrix, group::AbstractVector, gene_names::AbstractVector,
                            pval_reo::AbstractFloat, pval_deg::AbstractFloat, padj_deg::AbstractFloat,
                            ref_gene::BitVector, n_iter::Int64, n_conv::Int64; handle::ReoHandle = reo_handle())
    r, c = size(data)
    glev = unique(group)                                           # src:353
    gnum = length(glev)
    c == length(group) || throw(DimensionMismatch("'data' and 'group' do not have compatiable sizes"))
    gnum > 1 || throw(DimensionMismatch("Only 1 level in 'group1, at least 2 levels!"))
    T = eltype(data) <: Integer ? (sizeof(eltype(data)) <= 4 ? Int32 : Int64) : (eltype(data) == Float32 ? Float32 : Float64)
    mat = Matrix{T}(data)                                          # column-major, owned by Julia
    gid = Int32[findfirst(==(g), glev) - 1 for g in group]         # 0-based level ids, order of first appearance
    # thresholds as at src:362 (kept in Julia: integer results cross the ABI)
    gsi1 = [count(==(l), group) for l in glev]; gsi2 = c .- gsi1
    thr = Matrix{Int32}(get_major_reo_lower_count.(Matrix(hcat(gsi1, gsi2)'), pval_reo))   # 2 x gnum
    K = gnum == 2 ? 1 : gnum
    result = reo_pinned_array(Float64, r, 15, K)                   # page-locked: results are DMA-ed straight into them
    updown = reo_pinned_array(Int8, r, K)
    final_ref = reo_pinned_array(UInt8, r, K)
    iters = Vector{Int32}(undef, K)
    stats = Ref{ReoStats}()
    ref = Vector{UInt8}(ref_gene)
    rc = GC.@preserve mat gid thr ref result updown final_ref iters begin
        ccall((:reo_identify_degs, LIBREO), Cint,
              (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Int64, Int64, Ptr{Int32}, Int32, Ptr{Int32}, Float64, Float64, Float64,
               Ptr{UInt8}, Int32, Int32, UInt32, Ptr{Float64}, Ptr{Int8}, Ptr{UInt8}, Ptr{Int32}, Ref{ReoStats}),
              handle.ptr, mat, reo_dtype(T), r, c, r, gid, gnum, thr, pval_reo, pval_deg, padj_deg,
              ref, n_iter, n_conv, REO_OUT_PINNED, result, updown, final_ref, iters, stats)
    end
    rc == REO_OK || reo_throw(handle, rc)
    # the reference's log lines, per level k (src:418-420, 432-435), from the library's per-level iteration log
    n_deg = Vector{Int32}(undef, REO_MAX_ITER_LOG); n_ref = similar(n_deg)
    for k in 1:K
        it = Ref{Int32}(0); cv = Ref{Int32}(0)
        ccall((:reo_iter_log, LIBREO), Cint, (Ptr{Cvoid}, Int32, Ref{Int32}, Ref{Int32}, Ptr{Int32}, Ptr{Int32}, Int32),
              handle.ptr, k - 1, it, cv, n_deg, n_ref, REO_MAX_ITER_LOG)
        for e in 1:min(it[], REO_MAX_ITER_LOG)
            @info "INFO: iteration $(e-1),  # DEGs $(n_deg[e]), # non-DEGs $(r - n_deg[e])"        # src:418
        end
        cv[] == 1 && @info "INFO: Convergence threshold is reached"                                # src:420
        if gnum == 2                                                                               # src:431-435
            @info "INFO: The results of differentially expressed genes in the iteration process of $(glev[1]) vs $(glev[2]) were output."
        else
            @info "INFO: The results of differentially expressed genes in the iteration process of $(glev[k]) vs other were output."
        end
    end
    res = gene_names                                               # src:394
    for k in 1:K
        ud = String.(gene_names); ud .= "no change"
        ud[updown[:, k] .== 1] .= "up"; ud[updown[:, k] .== -1] .= "down"    # src:426-429
        res = hcat(res, result[:, :, k], ud)                        # src:430
    end
    return res                                                     # r x (1 + 16K) Matrix{Any}, src:437
end
