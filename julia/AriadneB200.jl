# AriadneB200.jl — Julia host side of libariadne_b200.so (thin `ccall` layer).
#
# NOTE: there is no `julia` binary in the build image or on the GPU box, so this file has never been executed.
# It binds exactly the C entry points of include/ariadne_b200.h that the Python mirror
# (newtonkrylov.jl_b200/host.py) drives in the test-suite; struct layouts are checked against the header by
# tests/test_abi.py (through the ctypes mirror with identical field order).
#
# Usage (drop-in for the hot path of Ariadne):
#     using AriadneB200
#     u   = B200Vector(sin.(π .* x) * sin.(π .* y)')            # uploads, compact slab layout
#     u, stats = newton_krylov!(Bratu2D(), u, (dx, dy, λ); krylov_kwargs = (; restart = true))
module AriadneB200

export B200Vector, newton_krylov!, newton_krylov, JacobianOperator, Fixed, EisenstatWalker, GmresPreconditioner,
       TridiagonalLU, ilu, JacobiPreconditioner, UserPreconditioner, UserResidual,
       Bratu1D, Bratu2D, Heat1D, Diffusion2D, Heat1DDG, GEuler, solve

using LinearAlgebra
import LinearAlgebra: mul!

const lib = get(ENV, "ARIADNE_B200_LIB", joinpath(@__DIR__, "..", "newtonkrylov.jl_b200", "libariadne_b200.so"))

# ---- status handling -------------------------------------------------------------------------------------------
last_error() = unsafe_string(ccall((:ak_last_error, lib), Cstring, ()))
function check(rc::Cint)
    rc < 0 && error("libariadne_b200 error $rc: $(last_error())")
    return rc
end

# ---- PODs (field order == include/ariadne_b200.h) -----------------------------------------------------------------
struct AkProblem
    kind::Int32; bc::Int32; scheme::Int32; jvp_mode::Int32
    nx::Int64; ny::Int64; gny::Int64; gy0::Int64
    dx::Float64; dy::Float64; lambda::Float64; a::Float64; dt::Float64; fd_eps::Float64
    un::Ptr{Float64}; coef::Ptr{Float64}; work::Ptr{Float64}
    user_residual::Ptr{Cvoid}; user_jvp::Ptr{Cvoid}; user_data::Ptr{Cvoid}      # AK_USER
end
# every native residual: no callbacks
AkProblem(kind, bc, scheme, jvp_mode, nx, ny, gny, gy0, dx, dy, lambda, a, dt, fd_eps, un, coef, work) =
    AkProblem(kind, bc, scheme, jvp_mode, nx, ny, gny, gy0, dx, dy, lambda, a, dt, fd_eps, un, coef, work, C_NULL, C_NULL, C_NULL)
struct AkKrylovOpts
    atol::Float64; rtol::Float64; itmax::Int64
    restart::Int32; reorthogonalization::Int32; history::Int32; fuse::Int32
    precond_n::Int32; precond_itmax::Int32; precond_m::Int32; precond_m_itmax::Int32
    n_apply::Ptr{Cvoid}; n_user::Ptr{Cvoid}; m_apply::Ptr{Cvoid}; m_user::Ptr{Cvoid}
end
struct AkKrylovStats
    niter::Int64; solved::Int32; inconsistent::Int32; breakdown::Int32; npass::Int32
    rnorm::Float64; beta::Float64
end
struct AkNewtonOpts
    tol_rel::Float64; tol_abs::Float64; max_niter::Int32; forcing::Int32
    eta::Float64; eta_max::Float64; gamma::Float64
    algo::Int32; memory::Int32; max_basis::Int64
    krylov::AkKrylovOpts
    krylov_rtol_override::Int32; verbose::Int32
end
struct AkNewtonStats
    solved::Int32; outer_iterations::Int32; inner_iterations::Int64
    n_res::Float64; tol::Float64; t_seconds::Float64; flags::Int32
end

const AK_SIMPLE2, AK_BRATU1D, AK_BRATU2D, AK_HEAT1D, AK_HEAT2D, AK_HEAT1D_DG, AK_USER = Int32.(0:6)
const AK_STEADY, AK_EULER = Int32(0), Int32(1)
const AK_JVP_ANALYTIC, AK_JVP_FD = Int32(0), Int32(2)
const AK_ALGO = Dict(:gmres => Int32(0), :cg => Int32(1), :fgmres => Int32(2))
const AK_PRECOND_NONE, AK_PRECOND_INNER_GMRES, AK_PRECOND_USER, AK_PRECOND_JACOBI, AK_PRECOND_TRIDIAG_LU = Int32.(0:4)

# ---- preconditioner objects returned by `M(J)` / `N(J)` (src/Ariadne.jl:324-329) -----------------------------------------
"N = (J) -> GmresPreconditioner(J, itmax) of examples/bratu.jl:141-149; run natively as an inner GMRES"
struct GmresPreconditioner{JOp}
    J::JOp
    itmax::Int
end
"What `ilu(collect(J))` is for the tridiagonal 1-D Bratu Jacobian (examples/bratu.jl:121-139); applied with ldiv = true"
struct TridiagonalLU{JOp}; J::JOp; end
ilu(J) = TridiagonalLU(J)
"y = x ./ diag(J(u))"
struct JacobiPreconditioner{JOp}; J::JOp; end
"Any preconditioner of the caller: `apply!(y::CuPtr, x::CuPtr, n, stream)` enqueues y <- P x on `stream`"
struct UserPreconditioner{F}; apply!::F; ldiv::Bool; end
precond_fields(::Nothing) = (AK_PRECOND_NONE, Int32(0), C_NULL, C_NULL)
precond_fields(P::GmresPreconditioner) = (AK_PRECOND_INNER_GMRES, Int32(P.itmax), C_NULL, C_NULL)
precond_fields(::TridiagonalLU) = (AK_PRECOND_TRIDIAG_LU, Int32(0), C_NULL, C_NULL)
precond_fields(::JacobiPreconditioner) = (AK_PRECOND_JACOBI, Int32(0), C_NULL, C_NULL)
function precond_trampoline(user::Ptr{Cvoid}, stream::UInt64, x::Ptr{Float64}, y::Ptr{Float64})::Cint
    P, n = (unsafe_pointer_to_objref(user)::Base.RefValue{Any})[]
    try
        P.apply!(y, x, n, stream)
        return 0
    catch
        return 1
    end
end

# ---- context and device vectors ------------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ak_ctx_create, lib), Cint, (Cint, Ptr{Ptr{Cvoid}}), device, r))
        c = new(r[])
        finalizer(c -> ccall((:ak_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.h), c)
        return c
    end
end
const CTX = Ref{Union{Nothing, Context}}(nothing)
context() = (CTX[] === nothing && (CTX[] = Context()); CTX[]::Context)

"fp64 vector resident in HBM (compact slab, no ghost cells); `dims` = (n,) or (nx, ny) with x fastest."
mutable struct B200Vector{N} <: AbstractVector{Float64}
    ptr::Ptr{Float64}
    dims::NTuple{N, Int}
    ctx::Context
    function B200Vector{N}(::UndefInitializer, dims::NTuple{N, Int}; ctx = context()) where {N}
        r = Ref{Ptr{Float64}}(C_NULL)
        check(ccall((:ak_malloc, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Float64}}), ctx.h, prod(dims), r))
        v = new{N}(r[], dims, ctx)
        finalizer(v -> ccall((:ak_free, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), v.ctx.h, v.ptr), v)
        return v
    end
end
function B200Vector(a::Array{Float64, N}) where {N}
    v = B200Vector{N}(undef, size(a))
    check(ccall((:ak_upload, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), v.ctx.h, v.ptr, a, length(a)))
    return v
end
Base.size(v::B200Vector) = (prod(v.dims),)          # logical length = interior only (examples/halovector.jl:17-26)
Base.similar(v::B200Vector{N}) where {N} = B200Vector{N}(undef, v.dims; ctx = v.ctx)
Base.zero(v::B200Vector) = (z = similar(v); Krylov_kfill!(z, 0.0); z)
Base.copy(v::B200Vector) = (c = similar(v); Krylov_kcopy!(length(v), c, v); c)
function Base.Array(v::B200Vector{N}) where {N}
    a = Array{Float64, N}(undef, v.dims)
    check(ccall((:ak_download, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), v.ctx.h, a, v.ptr, length(a)))
    return a
end

# ---- Krylov.k* hooks (examples/halovector.jl:51-147).  With Krylov.jl loaded these are attached as
#      `Krylov.kdot(n, x::B200Vector, y::B200Vector) = Krylov_kdot(n, x, y)` etc. so that Krylov.jl itself can drive
#      the vectors; the fused native solver below does not need them. ---------------------------------------------------
function Krylov_kdot(n::Integer, x::B200Vector, y::B200Vector)
    r = Ref{Float64}(0.0)
    check(ccall((:ak_dot, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), x.ctx.h, n, x.ptr, y.ptr, r))
    return r[]
end
function Krylov_knorm(n::Integer, x::B200Vector)
    r = Ref{Float64}(0.0)
    check(ccall((:ak_nrm2, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}), x.ctx.h, n, x.ptr, r))
    return r[]
end
Krylov_kscal!(n, s, x::B200Vector) = (check(ccall((:ak_scal, lib), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Float64}), x.ctx.h, n, s, x.ptr)); x)
Krylov_kaxpy!(n, s, x::B200Vector, y::B200Vector) = (check(ccall((:ak_axpy, lib), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Float64}, Ptr{Float64}), x.ctx.h, n, s, x.ptr, y.ptr)); y)
Krylov_kaxpby!(n, s, x::B200Vector, t, y::B200Vector) = (check(ccall((:ak_axpby, lib), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Float64}, Float64, Ptr{Float64}), x.ctx.h, n, s, x.ptr, t, y.ptr)); y)
Krylov_kcopy!(n, y::B200Vector, x::B200Vector) = (check(ccall((:ak_copy, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}), x.ctx.h, n, y.ptr, x.ptr)); y)
Krylov_kfill!(x::B200Vector, val) = (check(ccall((:ak_fill, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Float64), x.ctx.h, length(x), x.ptr, val)); x)
Krylov_kref!(n, x::B200Vector, y::B200Vector, c, s) = (check(ccall((:ak_ref, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Float64), x.ctx.h, n, x.ptr, y.ptr, c, s)); (x, y))

# ---- native residuals F!(res, u, p) -------------------------------------------------------------------------------------
abstract type NativeResidual end
struct Bratu1D <: NativeResidual end       # examples/bratu.jl:14-24,   p = (Δx, λ)
struct Bratu2D <: NativeResidual end       # 2-D extension,             p = (Δx, Δy, λ)
struct GEuler{R} <: NativeResidual; f::R; end   # G_Euler! ∘ f!,  p = (uₙ, Δt, du, p_f, t)  (examples/implicit.jl:8-13,61)
struct Heat1D end                          # examples/heat_1D.jl:12-25, p_f = (a, Δx, bc) with bc ∈ (:zero, :periodic)
struct Diffusion2D end                     # examples/heat_2D.jl:45-62, p_f = (a, Δx, Δy, bc)
struct Heat1DDG end                        # examples/heat_1D_DG.jl:32-36, p_f = (h,)

bc_code(bc) = bc === :periodic ? Int32(1) : Int32(0)
problem(::Bratu1D, u, p; coef = C_NULL) =
    AkProblem(AK_BRATU1D, 0, AK_STEADY, 0, length(u), 1, 1, 0, p[1], 0.0, p[2], 0.0, 0.0, 0.0, C_NULL, coef, C_NULL)
problem(::Bratu2D, u, p; coef = C_NULL) =
    AkProblem(AK_BRATU2D, 0, AK_STEADY, 0, u.dims[1], u.dims[2], u.dims[2], 0, p[1], p[2], p[3], 0.0, 0.0, 0.0, C_NULL, coef, C_NULL)
function problem(F::GEuler, u, p; coef = C_NULL)
    un, dt, _, pf, _ = p
    if F.f isa Heat1D
        return AkProblem(AK_HEAT1D, bc_code(pf[3]), AK_EULER, 0, length(u), 1, 1, 0, pf[2], 0.0, 0.0, pf[1], dt, 0.0, un.ptr, C_NULL, C_NULL)
    elseif F.f isa Diffusion2D
        return AkProblem(AK_HEAT2D, bc_code(pf[4]), AK_EULER, 0, u.dims[1], u.dims[2], u.dims[2], 0, pf[2], pf[3], 0.0, pf[1], dt, 0.0, un.ptr, C_NULL, C_NULL)
    else
        return AkProblem(AK_HEAT1D_DG, 1, AK_EULER, 0, length(u), 1, 1, 0, pf[1], 0.0, 0.0, 0.0, dt, 0.0, un.ptr, C_NULL, C_NULL)
    end
end
# ---- caller-supplied residuals: the generic seam of newton_krylov!(F!, u, p, res) (src/Ariadne.jl:250-256) ------------------
# `F!(res::Ptr{Float64}, u::Ptr{Float64}, p, n, stream)` and (optionally) `jvp!(out, u, v, p, n, stream)` enqueue device work
# on `stream` (e.g. CUDA.jl kernels launched with `stream = CuStream(stream)`; the tangent is what
# `Enzyme.autodiff(Forward, ...)` of the same kernel computes).  Without `jvp!` the library forms
# (F(u + ε v) - F(u)) / ε itself (AK_JVP_FD).
mutable struct UserResidual{F, T} <: NativeResidual
    F!::F
    jvp!::T
    p::Any
    n::Int
end
UserResidual(F!, jvp! = nothing) = UserResidual(F!, jvp!, nothing, 0)
function user_residual_trampoline(user::Ptr{Cvoid}, stream::UInt64, u::Ptr{Float64}, res::Ptr{Float64})::Cint
    R = unsafe_pointer_to_objref(user)::UserResidual
    try
        R.F!(res, u, R.p, R.n, stream)
        return 0
    catch
        return 1
    end
end
function user_jvp_trampoline(user::Ptr{Cvoid}, stream::UInt64, u::Ptr{Float64}, v::Ptr{Float64}, out::Ptr{Float64})::Cint
    R = unsafe_pointer_to_objref(user)::UserResidual
    try
        R.jvp!(out, u, v, R.p, R.n, stream)
        return 0
    catch
        return 1
    end
end
function problem(R::UserResidual, u, p; coef = C_NULL)
    R.p, R.n = p, length(u)                      # R must stay rooted while the solve runs (it is: the caller holds it)
    fr = @cfunction(user_residual_trampoline, Cint, (Ptr{Cvoid}, UInt64, Ptr{Float64}, Ptr{Float64}))
    fj = R.jvp! === nothing ? C_NULL :
         @cfunction(user_jvp_trampoline, Cint, (Ptr{Cvoid}, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}))
    return AkProblem(AK_USER, 0, AK_STEADY, R.jvp! === nothing ? AK_JVP_FD : AK_JVP_ANALYTIC, length(u), 1, 1, 0,
                     0.0, 0.0, 0.0, 0.0, 0.0, 0.0, C_NULL, coef, C_NULL, fr, fj, pointer_from_objref(R))
end

function (F::NativeResidual)(res::B200Vector, u::B200Vector, p)
    prob = Ref(problem(F, u, p))
    check(ccall((:ak_residual, lib), Cint, (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), u.ctx.h, prob, u.ptr, res.ptr, C_NULL))
    return nothing
end

# ---- JacobianOperator (src/Ariadne.jl:34-57) -------------------------------------------------------------------------------
struct JacobianOperator{F, A, P}
    f::F; res::A; u::A; p::P
    coef::Ptr{Float64}
    JacobianOperator(f::F, res, u, p; coef = C_NULL) where {F} = new{F, typeof(u), typeof(p)}(f, res, u, p, coef)
end
Base.size(J::JacobianOperator) = (length(J.res), length(J.u))
Base.eltype(J::JacobianOperator) = Float64
Base.length(J::JacobianOperator) = prod(size(J))
function mul!(out::B200Vector, J::JacobianOperator, v::B200Vector)
    prob = Ref(problem(J.f, J.u, J.p; coef = J.coef))
    check(ccall((:ak_jvp, lib), Cint, (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), v.ctx.h, prob, J.u.ptr, v.ptr, out.ptr))
    return nothing
end

# ---- forcing (src/Ariadne.jl:180-217) -----------------------------------------------------------------------------------------
abstract type Forcing end
Base.@kwdef struct Fixed <: Forcing; η::Float64 = 0.1; end
(F::Fixed)(args...) = F.η
inital(F::Fixed) = F.η
Base.@kwdef struct EisenstatWalker <: Forcing; η_max::Float64 = 0.999; γ::Float64 = 0.9; end
(F::EisenstatWalker)(η, tol, n_res, n_res_prior) =
    ccall((:ak_forcing_ew, lib), Float64, (Float64, Float64, Float64, Float64, Float64, Float64), F.η_max, F.γ, η, tol, n_res, n_res_prior)
inital(F::EisenstatWalker) = F.η_max

struct Stats
    outer_iterations::Int; inner_iterations::Int; n_res::Float64
end
update(s::Stats, inner, n_res) = Stats(s.outer_iterations + 1, s.inner_iterations + inner, n_res)

# ---- Krylov workspace (krylov_workspace / krylov_solve!: src/Ariadne.jl:317-318,338-340) ------------------------------------------
mutable struct Workspace
    h::Ptr{Cvoid}; proto::B200Vector; niter::Int; solved::Bool
end
function krylov_workspace(algo::Symbol, res::B200Vector; memory = 20, max_basis = 0)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ak_krylov_create, lib), Cint, (Ptr{Cvoid}, Int32, Int64, Int32, Int64, Ptr{Ptr{Cvoid}}), res.ctx.h, AK_ALGO[algo], length(res), memory, max_basis, r))
    ws = Workspace(r[], res, 0, false)
    finalizer(w -> ccall((:ak_krylov_destroy, lib), Cint, (Ptr{Cvoid},), w.h), ws)
    return ws
end
solution(ws::Workspace) = ccall((:ak_krylov_x, lib), Ptr{Float64}, (Ptr{Cvoid},), ws.h)
function krylov_solve!(ws::Workspace, J::JacobianOperator, b::B200Vector; atol = √eps(Float64), rtol = √eps(Float64),
                       itmax = 0, restart = false, reorthogonalization = false, history = false, fuse = 5,
                       M = nothing, N = nothing, ldiv = false)
    (N isa TridiagonalLU || M isa TridiagonalLU) && !ldiv && error("ilu(J) is applied with ldiv = true (examples/bratu.jl:126)")
    # caller-supplied preconditioners travel as (object, n) behind a rooted Ref for the duration of the solve
    nref = N isa UserPreconditioner ? Ref{Any}((N, length(b))) : nothing
    mref = M isa UserPreconditioner ? Ref{Any}((M, length(b))) : nothing
    tramp = @cfunction(precond_trampoline, Cint, (Ptr{Cvoid}, UInt64, Ptr{Float64}, Ptr{Float64}))
    fields(P, r) = r === nothing ? precond_fields(P) : (AK_PRECOND_USER, Int32(0), tramp, pointer_from_objref(r))
    pn, pit, nfn, nus = fields(N, nref)
    pm, pmit, mfn, mus = fields(M, mref)
    o = Ref(AkKrylovOpts(atol, rtol, itmax, restart, reorthogonalization, history, fuse, pn, pit, pm, pmit, nfn, nus, mfn, mus))
    st = Ref(AkKrylovStats(0, 0, 0, 0, 0, 0.0, 0.0))
    prob = Ref(problem(J.f, J.u, J.p; coef = J.coef))
    GC.@preserve nref mref J begin
        check(ccall((:ak_krylov_solve, lib), Cint,
                    (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Ptr{AkKrylovOpts}, Ptr{AkKrylovStats}, Ptr{Float64}, Int64),
                    ws.h, prob, J.u.ptr, b.ptr, o, st, C_NULL, 0))
    end
    ws.niter, ws.solved = st[].niter, st[].solved != 0
    return ws
end

# ---- multi-GPU: one Julia process per GPU; `id` (128 bytes) comes from rank 0 through Distributed / MPI ------------------
comm_unique_id() = (id = zeros(UInt8, 128); check(ccall((:ak_comm_unique_id, lib), Cint, (Ptr{UInt8},), id)); id)
comm_init(ctx::Context, nranks, rank, id::Vector{UInt8}) =
    check(ccall((:ak_comm_init, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx.h, nranks, rank, id))
comm_enable_p2p(ctx::Context, halo_doubles) =
    check(ccall((:ak_comm_enable_p2p, lib), Cint, (Ptr{Cvoid}, Int64), ctx.h, halo_doubles))

# ---- newton_krylov! (src/Ariadne.jl:288-372): the loop is driven from Julia, one ccall per arrowed line ----------------------------
function newton_krylov!(F!::NativeResidual, u::B200Vector, p = nothing, res::B200Vector = zero(u);
                        tol_rel = 1.0e-6, tol_abs = 1.0e-12, max_niter = 50,
                        forcing::Union{Forcing, Nothing} = EisenstatWalker(), verbose = 0, algo = :gmres,
                        M = nothing, N = nothing, krylov_kwargs = (;), callback = (args...) -> nothing)
    t₀ = time_ns()
    # λ·exp(u) cache shared by residual and JVPs (Bratu); F(u) cache for finite-difference JVPs (user F! without tangent)
    coef = (F! isa Union{Bratu1D, Bratu2D} || (F! isa UserResidual && F!.jvp! === nothing)) ? similar(u) : nothing
    cptr = coef === nothing ? Ptr{Float64}(C_NULL) : coef.ptr
    prob = Ref(problem(F!, u, p; coef = cptr))
    nrm = Ref{Float64}(0.0)
    residual_norm() = (check(ccall((:ak_residual, lib), Cint, (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                   u.ctx.h, prob, u.ptr, res.ptr, nrm)); nrm[])
    n_res = residual_norm()
    callback(u, res, n_res)
    tol = tol_rel * n_res + tol_abs
    η = forcing === nothing ? nothing : inital(forcing)
    J = JacobianOperator(F!, res, u, p; coef = cptr)
    workspace = krylov_workspace(algo, res)
    rhs = similar(res)
    stats = Stats(0, 0, n_res)
    while n_res > tol && stats.outer_iterations <= max_niter
        kwargs = krylov_kwargs
        N !== nothing && (kwargs = (; N = N(J), kwargs...))
        M !== nothing && (kwargs = (; M = M(J), kwargs...))
        forcing !== nothing && (kwargs = (; rtol = η, kwargs...))
        Krylov_kcopy!(length(res), rhs, res)                       # copy(res)
        krylov_solve!(workspace, J, rhs; kwargs...)
        check(ccall((:ak_axpy, lib), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Float64}, Ptr{Float64}),
                    u.ctx.h, length(u), -1.0, solution(workspace), u.ptr))   # u .-= s .* d, s = 1
        n_res_prior = n_res
        n_res = residual_norm()
        callback(u, res, n_res)
        if isinf(n_res) || isnan(n_res)
            @error "Inner solver blew up" stats
            break
        end
        forcing !== nothing && (η = forcing(η, tol, n_res, n_res_prior))
        stats = update(stats, workspace.niter, n_res)
    end
    t = (time_ns() - t₀) / 1.0e9
    return u, (; solved = n_res <= tol, stats, t)
end
newton_krylov(F::NativeResidual, u₀::B200Vector, p = nothing; kwargs...) = newton_krylov!(F, copy(u₀), p; kwargs...)

# ---- solve(G!, f!, uₙ, p, Δt, ts) (examples/implicit.jl:54-78) -------------------------------------------------------------------------
function solve(G::GEuler, uₙ::B200Vector, p, Δt, ts; callback = _ -> nothing, verbose = 0, algo = :gmres, krylov_kwargs = (;))
    u = copy(uₙ); du = zero(uₙ); res = zero(uₙ)
    for t in ts
        t == first(ts) && continue
        _, stats = newton_krylov!(G, u, (uₙ, Δt, du, p, t), res; verbose, algo, tol_abs = 6.0e-6, krylov_kwargs)
        stats.solved || @warn "non linear solve failed marching on" t stats
        callback(u)
        Krylov_kcopy!(length(u), uₙ, u)
    end
    return uₙ
end

end # module
