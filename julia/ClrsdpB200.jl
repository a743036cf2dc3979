# ClrsdpB200.jl — the binding a maintainer of the reference adds to drive libclrsdp.so (SURVEY §8 row f1).
#
# `include` this file after MPMP.jl: it replaces the body of `MPMP.solverank1sdp` (MPMP.jl:595-1025) by calls into the
# C ABI of include/clrsdp.h; `solvempmp`, `prepareabc`, `get_block_info` and the helpers stay as they are. Same keyword
# arguments, same log table, same 11-tuple (MPMP.jl:1014-1024), same "higher precision" errors.
#
# NOT EXECUTED IN THIS REPOSITORY: the build image has no Julia toolchain. The Python mirror
# clustered-low-rank-sdp-solver_b200/clrsdp/solver.py performs the same call sequence and is what the tests drive;
# keep the two in step.
module ClrsdpB200

using Arblib, BlockDiagonals, Printf

const LIB = get(ENV, "CLRSDP_LIB", "libclrsdp.so")
const T_COUNT = 17

# ---- wire format (include/clrsdp.h: clrsdp_mp) ----------------------------------------------------------------
# planar: sign[n] (Int8), exp[n] (Int64), limb[nlimb][n] (UInt32, limb 0 least significant); value = sign * 0.limbs * 2^exp
# — MPFR's own representation, so BigFloat <-> wire is a copy.
struct MpArr
    sign::Vector{Int8}
    exp::Vector{Int64}
    limb::Matrix{UInt32}      # n x nlimb, column-major: column k = limb k of all numbers
end
MpArr(n::Integer, nl::Integer) = MpArr(zeros(Int8, n), zeros(Int64, n), zeros(UInt32, n, nl))
struct CMp
    sign::Ptr{Int8}
    exp::Ptr{Int64}
    limb::Ptr{UInt32}
    n::Int64
end
cmp(a::MpArr) = CMp(pointer(a.sign), pointer(a.exp), pointer(a.limb), length(a.sign))
nlimbs() = precision(BigFloat) ÷ 32

function MpArr(v::AbstractVector{BigFloat})
    p = precision(BigFloat)
    (p % 32 == 0 && 128 <= p <= 512) || error("the B200 path needs precision(BigFloat) a multiple of 32 in [128, 512]")
    n, nl = length(v), nlimbs()
    a = MpArr(n, nl)
    nwords = 2 * cld(p, 64)                       # 32-bit words of the mpfr limb array; the mantissa is top-aligned
    for (i, x) in enumerate(v)
        precision(x) == p || (x = BigFloat(x; precision = p))
        iszero(x) && continue
        a.sign[i] = x.sign < 0 ? -1 : 1
        a.exp[i] = x.exp
        d = unsafe_wrap(Array, Ptr{UInt32}(x.d), nwords)
        a.limb[i, :] = d[end-nl+1:end]
    end
    a
end
midpoints(A::ArbMatrix) = BigFloat[BigFloat(Arblib.midref(A[i, j])) for i in 1:size(A, 1) for j in 1:size(A, 2)]  # row-major
MpArr(A::ArbMatrix) = MpArr(midpoints(A))
MpArr(v::AbstractVector) = MpArr(BigFloat[BigFloat(x) for x in v])

function to_bigfloats(a::MpArr)
    p, nl = precision(BigFloat), size(a.limb, 2)
    nwords = 2 * cld(p, 64)
    out = Vector{BigFloat}(undef, length(a.sign))
    for i in eachindex(out)
        x = BigFloat(a.sign[i] == 0 ? 0 : 1; precision = p)          # allocates the limb array; overwritten below
        if a.sign[i] != 0
            d = unsafe_wrap(Array, Ptr{UInt32}(x.d), nwords)
            fill!(d, 0)
            d[end-nl+1:end] = a.limb[i, :]
            x.exp = a.exp[i]
            x.sign = a.sign[i] < 0 ? -1 : 1
        end
        out[i] = x
    end
    out
end
to_arbmatrix(v::Vector{BigFloat}, r, c) = ArbMatrix(permutedims(reshape(v, c, r)); prec = precision(BigFloat))

# ---- clrsdp_iter_info ---------------------------------------------------------------------------------------------
struct IterInfo
    iter::Int32; status::Int32; terminate::Int32; pd_feasible::Int32
    mu::Float64; p_obj::Float64; d_obj::Float64; gap::Float64; P_err::Float64; p_err::Float64; d_err::Float64
    alpha_p::Float64; alpha_d::Float64; beta_c::Float64
    p_obj_new::Float64; d_obj_new::Float64; gap_new::Float64; primal_err_new::Float64; dual_err_new::Float64
    seconds::Float64
    timings::NTuple{T_COUNT,Float64}
end

const HIGHER_PRECISION = (-10, -11, -12, -13, -14, -16)   # NOT_PD_X, NOT_PD_Y, SINGULAR_S, SINGULAR_Q, EIG, DIVERGED
function chk(h, st, what)
    st == 0 && return
    if st in HIGHER_PRECISION                        # the reference's messages (MPMP.jl:793, :1439, :1503, :1882)
        error("$what failed. Try again with higher precision")
    end
    error("$what: " * unsafe_string(ccall((:clrsdp_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))
end

function fetch(h, name::String, j::Integer, l::Integer)
    n = ccall((:clrsdp_fetch, LIB), Int64, (Ptr{Cvoid}, Cstring, Cint, Cint, Ptr{Cvoid}), h, name, j, l, C_NULL)
    n >= 0 || error("clrsdp_fetch($name): $n")
    a = MpArr(n, nlimbs())
    GC.@preserve a begin
        r = ccall((:clrsdp_fetch, LIB), Int64, (Ptr{Cvoid}, Cstring, Cint, Cint, Ref{CMp}), h, name, j, l, cmp(a))
        r >= 0 || error("clrsdp_fetch($name): $r")
    end
    to_bigfloats(a)
end

flatten_blocks(X) = MpArr(vcat([midpoints(blk) for cl in X for blk in cl]...))    # blocks in (j, l) order, row-major

const HEADER = @sprintf("%5s %8s %11s %11s %11s %10s %10s %10s %10s %10s %10s %10s", "iter", "time(s)", "mu", "P-obj",
                        "D-obj", "gap", "P-error", "p-error", "d-error", "alpha_p", "alpha_d", "beta")
print_row(r::IterInfo, t) = @printf("%5d %8.1f %11.3e %11.3e %11.3e %10.2e %10.2e %10.2e %10.2e %10.2e %10.2e %10.2e\n",
                                    r.iter, t, r.mu, r.p_obj, r.d_obj, r.gap, r.P_err, r.p_err, r.d_err, r.alpha_p,
                                    r.alpha_d, r.beta_c)

"""
    solverank1sdp(constraints, b, blockinfo; kwargs...)

Drop-in for `MPMP.solverank1sdp` (MPMP.jl:595-614): `constraints` is the vector of `(A, B, c, H)` tuples of `prepareabc`,
`blockinfo` the result of `get_block_info`. Returns `x, X, y, Y, P, p, d, dual_gap, primal_obj, dual_obj, time_total`.
"""
function solverank1sdp(constraints, b, blockinfo; C = 0, b0 = 0, maxiterations = 500,
        beta_infeasible = BigFloat(3) / 10, beta_feasible = BigFloat(1) / 10, gamma = BigFloat(7) / 10,
        omega_p = BigFloat(10)^10, omega_d = BigFloat(10)^10, duality_gap_threshold = BigFloat(10)^(-15),
        primal_error_threshold = BigFloat(10)^(-30), dual_error_threshold = BigFloat(10)^(-30),
        need_primal_feasible = false, need_dual_feasible = false, testing = true, initial_solutions = [],
        devices = [0])
    # `devices`: CUDA ordinals. One entry = one GPU; several = ONE handle that shards the clusters over those GPUs of the
    # box inside this Julia process (clrsdp_create_multi: weighted partition of the clusters, one library thread per
    # device, NCCL between them) - no extra processes, nothing else changes below.
    bi = blockinfo
    href = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:clrsdp_create_multi, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Cint}), href, precision(BigFloat),
               length(devices), Cint.(devices))
    st == 0 || error("clrsdp_create_multi failed ($st): no sm_100 device or unsupported precision — there is no CPU fallback")
    h = href[]
    try
        delta = Cint[bi.Y_blocksizes[j][l] ÷ bi.m[j] for j in 1:bi.J for l in 1:bi.L[j]]
        ranks = Cint[bi.ranks[j][l][k] for j in 1:bi.J for l in 1:bi.L[j] for k in 1:bi.n_samples[j]]
        chk(h, ccall((:clrsdp_set_structure, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}),
                     h, bi.J, bi.n_y, Cint.(bi.m), Cint.(bi.L), Cint.(bi.n_samples), delta, ranks), "set_structure")
        for (j, (A, B, c, H)) in enumerate(constraints)            # the tuple of prepareabc (MPMP.jl:401-406)
            V = MpArr(vcat([midpoints(A[l, k][r]) for l in 1:bi.L[j] for k in 1:bi.n_samples[j] for r in 1:length(A[l, k])]...))
            Hs = MpArr(BigFloat[BigFloat(Arblib.midref(H[l, k][r])) for l in 1:bi.L[j] for k in 1:bi.n_samples[j] for r in 1:length(H[l, k])])
            Bw, cw = MpArr(B), MpArr(c)
            GC.@preserve V Hs Bw cw begin
                chk(h, ccall((:clrsdp_upload_cluster, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{CMp}, Ref{CMp}, Ref{CMp}, Ref{CMp}),
                             h, j - 1, cmp(V), cmp(Hs), cmp(Bw), cmp(cw)), "upload_cluster")
            end
        end
        bw, b0w = MpArr(b), MpArr([b0])
        GC.@preserve bw b0w chk(h, ccall((:clrsdp_upload_objective, LIB), Cint, (Ptr{Cvoid}, Ref{CMp}, Ref{CMp}),
                                         h, cmp(bw), cmp(b0w)), "upload_objective")
        if C != 0                                                       # objective matrix (MPMP.jl:599, :1108-1118, :1031-1034)
            Cw = flatten_blocks([C.blocks[j].blocks for j in 1:bi.J])
            GC.@preserve Cw chk(h, ccall((:clrsdp_upload_C, LIB), Cint, (Ptr{Cvoid}, Ref{CMp}), h, cmp(Cw)), "upload_C")
        end
        rp = MpArr([beta_infeasible, beta_feasible, gamma, omega_p, omega_d, duality_gap_threshold,
                    primal_error_threshold, dual_error_threshold])   # order: CLRSDP_P_* of clrsdp.h
        ip = Cint[maxiterations, need_primal_feasible, need_dual_feasible, 1]   # 1: fill the 17 timing buckets (:889-898)
        GC.@preserve rp chk(h, ccall((:clrsdp_set_params, LIB), Cint, (Ptr{Cvoid}, Ref{CMp}, Ptr{Cint}), h, cmp(rp), ip), "set_params")
        if length(initial_solutions) == 4                             # warm start (MPMP.jl:689)
            x0, X0, y0, Y0 = initial_solutions
            xw, Xw, yw, Yw = MpArr(x0), flatten_blocks(X0), MpArr(y0), flatten_blocks(Y0)
            GC.@preserve xw Xw yw Yw chk(h, ccall((:clrsdp_upload_point, LIB), Cint,
                (Ptr{Cvoid}, Ref{CMp}, Ref{CMp}, Ref{CMp}, Ref{CMp}), h, cmp(xw), cmp(Xw), cmp(yw), cmp(Yw)), "upload_point")
        else
            chk(h, ccall((:clrsdp_init_point, LIB), Cint, (Ptr{Cvoid},), h), "init_point")
        end
        info = Ref{IterInfo}()
        chk(h, ccall((:clrsdp_prepare, LIB), Cint, (Ptr{Cvoid}, Ref{IterInfo}), h, info), "prepare")
        println(HEADER)                                               # MPMP.jl:700-714
        iter, t0 = 1, time()
        timings = zeros(T_COUNT)
        while info[].terminate == 0 && iter < maxiterations           # MPMP.jl:742-753
            chk(h, ccall((:clrsdp_iterate, LIB), Cint, (Ptr{Cvoid}, Ref{IterInfo}), h, info), "clrsdp_iterate")
            print_row(info[], time() - t0)                            # MPMP.jl:923-937
            iter > 2 && (timings .+= collect(info[].timings))         # the first two iterations are not counted (:889)
            iter += 1
        end
        time_total = time() - t0
        info[].terminate == 3 && println("Optimal")
        println(HEADER)
        @printf("Time spent: total %.5e s; Decomp %.5e predict_dir %.5e correct_dir %.5e alpha %.5e\n", time_total,
                timings[1], timings[2], timings[3], timings[4])
        # ---- results (MPMP.jl:1014-1024) ----
        n_x = sum(bi.dim_S)
        sizes = [bi.Y_blocksizes[j][l] for j in 1:bi.J for l in 1:bi.L[j]]
        n_X = sum(s -> s * s, sizes)
        xw, Xw, yw, Yw = MpArr(n_x, nlimbs()), MpArr(n_X, nlimbs()), MpArr(bi.n_y, nlimbs()), MpArr(n_X, nlimbs())
        GC.@preserve xw Xw yw Yw chk(h, ccall((:clrsdp_download_point, LIB), Cint,
            (Ptr{Cvoid}, Ref{CMp}, Ref{CMp}, Ref{CMp}, Ref{CMp}), h, cmp(xw), cmp(Xw), cmp(yw), cmp(Yw)), "download_point")
        xs, Xs, ys, Ys = to_bigfloats(xw), to_bigfloats(Xw), to_bigfloats(yw), to_bigfloats(Yw)
        function blocks_of(flat)
            off, out = 0, Vector{Any}(undef, bi.J)
            for j in 1:bi.J
                blks = ArbMatrix[]
                for l in 1:bi.L[j]
                    s = bi.Y_blocksizes[j][l]
                    push!(blks, to_arbmatrix(flat[off+1:off+s*s], s, s))
                    off += s * s
                end
                out[j] = BlockDiagonal(blks)
            end
            BlockDiagonal(out)
        end
        X, Y = blocks_of(Xs), blocks_of(Ys)
        P = BlockDiagonal([BlockDiagonal([to_arbmatrix(fetch(h, "P", j - 1, l - 1), bi.Y_blocksizes[j][l], bi.Y_blocksizes[j][l])
                                          for l in 1:bi.L[j]]) for j in 1:bi.J])
        p, d = fetch(h, "p", 0, 0), fetch(h, "d", 0, 0)
        primal_obj, dual_obj = fetch(h, "scalar", 1, 0)[1], fetch(h, "scalar", 2, 0)[1]     # CLRSDP_S_P_OBJ, _D_OBJ (with b0)
        po, dobj = primal_obj - BigFloat(b0), dual_obj - BigFloat(b0)
        dual_gap = abs(po - dobj) / max(one(po), abs(po + dobj))                              # MPMP.jl:1067-1074
        return xs, X, ys, Y, P, p, d, dual_gap, primal_obj, dual_obj, time_total
    finally
        ccall((:clrsdp_destroy, LIB), Cint, (Ptr{Cvoid},), h)
    end
end

# ---- problem files (clrsdp/problem_io.py: CLRSDP1) ---------------------------------------------------------------
# Dumps the reference's own prepareabc output - and, optionally, the result of the reference's solverank1sdp - so that
# the Python test-suite can pin the oracle and the GPU path against the real reference (tests/golden/).
function put_arr(io::IO, a::MpArr)
    write(io, UInt64(length(a.sign)))
    write(io, a.sign)
    write(io, a.exp)
    write(io, a.limb)                 # n x nlimb column-major = limb plane k contiguous = [nlimb][n]
end
jstr(v::AbstractVector) = "[" * join(string.(v), ",") * "]"

"""
    write_problem(path, constraints, b, blockinfo; b0 = 0, solution = nothing)

`solution`: the 11-tuple returned by `solverank1sdp` plus the iteration count, as
`(x, X, y, Y, primal_obj, dual_obj, iterations)`, or `nothing`.
"""
function write_problem(path, constraints, b, bi; b0 = 0, solution = nothing)
    clusters = String[]
    for j in 1:bi.J
        delta = [bi.Y_blocksizes[j][l] ÷ bi.m[j] for l in 1:bi.L[j]]
        ranks = [jstr([bi.ranks[j][l][k] for k in 1:bi.n_samples[j]]) for l in 1:bi.L[j]]
        push!(clusters, "{\"m\":$(bi.m[j]),\"K\":$(bi.n_samples[j]),\"L\":$(bi.L[j]),\"delta\":$(jstr(delta)),\"ranks\":[" *
                        join(ranks, ",") * "]}")
    end
    sol = solution === nothing ? "null" :
          "{\"iterations\":$(solution[7]),\"primal_obj\":\"$(solution[5])\",\"dual_obj\":\"$(solution[6])\",\"has_point\":true}"
    hdr = "{\"prec\":$(precision(BigFloat)),\"n_y\":$(bi.n_y),\"b0\":\"$(b0)\",\"clusters\":[" * join(clusters, ",") *
          "],\"solution\":$sol}"
    open(path, "w") do io
        write(io, "CLRSDP1\n")
        write(io, UInt64(sizeof(hdr)))
        write(io, hdr)
        put_arr(io, MpArr(b))
        for (j, (A, B, c, H)) in enumerate(constraints)
            for l in 1:bi.L[j]
                put_arr(io, MpArr(vcat(BigFloat[], [midpoints(A[l, k][r]) for k in 1:bi.n_samples[j] for r in 1:length(A[l, k])]...)))
            end
            for l in 1:bi.L[j]
                put_arr(io, MpArr(BigFloat[BigFloat(Arblib.midref(H[l, k][r])) for k in 1:bi.n_samples[j] for r in 1:length(H[l, k])]))
            end
            put_arr(io, MpArr(B))
            put_arr(io, MpArr(c))
        end
        if solution !== nothing
            x, X, y, Y = solution[1:4]
            put_arr(io, MpArr(x)); put_arr(io, flatten_blocks(X)); put_arr(io, MpArr(y)); put_arr(io, flatten_blocks(Y))
        end
    end
end

end # module
