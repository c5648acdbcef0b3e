# AriadneB200.jl — Julia host side of libariadne_b200.so (thin `ccall` layer).
#
# NOTE: there is no `julia` binary in the build image or on the GPU box, so this file has never been executed.
# What IS checked mechanically (tests/test_julia_wrapper.py, CPU): every struct below has the field order and field
# types of include/ariadne_b200.h (through the ctypes mirror), and every `ccall` names an exported symbol with the
# right number and kinds of arguments.  The Python mirror (newtonkrylov.jl_b200/host.py) drives exactly these entry
# points in the GPU test-suite.
#
# Drop-in usage for the hot path of Ariadne (same call shapes as src/Ariadne.jl and the examples):
#     using AriadneB200
#     u = B200Vector(sin.(π .* x) * sin.(π .* y)')                      # uploads; compact slab layout, x fastest
#     u, stats = newton_krylov!(Bratu2D(), u, (Δx, Δy, λ); krylov_kwargs = (; restart = true))
#     u, stats = newton_krylov!(Bratu2D(), Array_u₀, (Δx, Δy, λ))        # host Array in / out (src/Ariadne.jl:259-263)
#     solve(GEuler(Diffusion2D()), uₙ, (a, Δx, Δy, :zero), Δt, ts; krylov_kwargs = (; verbose = 1, reorthogonalization = true))
# With Krylov.jl loaded, ext/AriadneB200KrylovExt.jl attaches Krylov.kdot/knorm/kscal!/kaxpy!/kaxpby!/kcopy!/kfill!/kref!
# to B200Vector (examples/halovector.jl:51-147), so Krylov.jl's own solvers can drive the device vectors as well.
module AriadneB200

export B200Vector, B200Matrix, newton_krylov!, newton_krylov, newton_krylov_native!, JacobianOperator, Fixed, EisenstatWalker,
       GmresPreconditioner, TridiagonalLU, ilu, JacobiPreconditioner, UserPreconditioner, UserResidual,
       Bratu1D, Bratu2D, Heat1D, Diffusion2D, Heat1DDG, GEuler, GMidpoint, GTrapezoid, solve,
       krylov_workspace, krylov_solve!, solution, Context, context

using LinearAlgebra
using SparseArrays
import LinearAlgebra: mul!

const lib = get(ENV, "ARIADNE_B200_LIB", joinpath(@__DIR__, "..", "..", "newtonkrylov.jl_b200", "libariadne_b200.so"))

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
const AK_STEADY, AK_EULER, AK_MIDPOINT, AK_TRAPEZOID = Int32.(0:3)
const AK_JVP_ANALYTIC, AK_JVP_FD_FUSED, AK_JVP_FD = Int32.(0:2)
const AK_ALGO = Dict(:gmres => Int32(0), :cg => Int32(1), :fgmres => Int32(2))
const AK_PRECOND_NONE, AK_PRECOND_INNER_GMRES, AK_PRECOND_USER, AK_PRECOND_JACOBI, AK_PRECOND_TRIDIAG_LU = Int32.(0:4)
const AK_FORCING_NONE, AK_FORCING_FIXED, AK_FORCING_EW = Int32.(0:2)
const AK_FUSE = Dict(:none => Int32(0), :mgs => Int32(1), :full => Int32(2), :pair => Int32(3), :block4 => Int32(4), :block8 => Int32(5), :sweep => Int32(6))
fuse_code(f::Symbol) = AK_FUSE[f]
fuse_code(f::Integer) = Int32(f)

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
"Any preconditioner of the caller: `apply!(y::Ptr{Float64}, x::Ptr{Float64}, n, stream)` enqueues y <- P x on `stream`"
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
    closed::Bool          # once true every finalizer that would touch the handle is a no-op (finalizers run in any order)
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ak_ctx_create, lib), Cint, (Cint, Ptr{Ptr{Cvoid}}), device, r))
        c = new(r[], false)
        finalizer(close, c)
        return c
    end
end
function Base.close(c::Context)
    c.closed && return nothing
    c.closed = true
    ccall((:ak_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.h)
    return nothing
end
const CTX = Ref{Union{Nothing, Context}}(nothing)
context() = (CTX[] === nothing && (CTX[] = Context()); CTX[]::Context)
sync(c::Context = context()) = check(ccall((:ak_ctx_sync, lib), Cint, (Ptr{Cvoid},), c.h))

"fp64 vector resident in HBM (compact slab, no ghost cells); `dims` = (n,) or (nx, ny) with x fastest."
mutable struct B200Vector{N} <: AbstractVector{Float64}
    ptr::Ptr{Float64}
    dims::NTuple{N, Int}
    ctx::Context
    owned::Bool
    function B200Vector{N}(::UndefInitializer, dims::NTuple{N, Int}; ctx = context()) where {N}
        r = Ref{Ptr{Float64}}(C_NULL)
        check(ccall((:ak_malloc, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Float64}}), ctx.h, prod(dims), r))
        v = new{N}(r[], dims, ctx, true)
        finalizer(release, v)
        return v
    end
    # non-owning view of device memory the library (or another vector) owns: workspace.x, callback arguments
    B200Vector{N}(ptr::Ptr{Float64}, dims::NTuple{N, Int}, ctx::Context) where {N} = new{N}(ptr, dims, ctx, false)
end
function release(v::B200Vector)
    (v.owned && !v.ctx.closed && v.ptr != C_NULL) || return nothing
    ccall((:ak_free, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), v.ctx.h, v.ptr)
    v.ptr = C_NULL
    return nothing
end
function B200Vector(a::Array{Float64, N}; ctx = context()) where {N}
    v = B200Vector{N}(undef, size(a); ctx)
    check(ccall((:ak_upload, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), v.ctx.h, v.ptr, a, length(a)))
    return v
end
Base.size(v::B200Vector) = (prod(v.dims),)          # logical length = interior only (examples/halovector.jl:17-26)
Base.similar(v::B200Vector{N}) where {N} = B200Vector{N}(undef, v.dims; ctx = v.ctx)
Base.similar(v::B200Vector, ::Type{Float64}) = similar(v)
Base.zero(v::B200Vector) = (z = similar(v); Krylov_kfill!(z, 0.0); z)
Base.copy(v::B200Vector) = (c = similar(v); Krylov_kcopy!(length(v), c, v); c)
function Base.Array(v::B200Vector{N}) where {N}
    a = Array{Float64, N}(undef, v.dims)
    check(ccall((:ak_download, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), v.ctx.h, a, v.ptr, length(a)))
    return a
end
function Base.copyto!(v::B200Vector, a::Array{Float64})
    length(a) == length(v) || throw(DimensionMismatch("copyto!: $(length(a)) values into a B200Vector of length $(length(v))"))
    check(ccall((:ak_upload, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), v.ctx.h, v.ptr, a, length(a)))
    return v
end
# Scalar indexing (examples/halovector.jl:28-40 defines it for HaloVector): one PCIe round trip per element — meant for
# inspection and printing, not for loops.
function Base.getindex(v::B200Vector, i::Int)
    1 <= i <= length(v) || throw(BoundsError(v, i))
    r = Ref{Float64}(0.0)
    check(ccall((:ak_download, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), v.ctx.h, r, v.ptr + 8 * (i - 1), 1))
    return r[]
end
function Base.setindex!(v::B200Vector, val, i::Int)
    1 <= i <= length(v) || throw(BoundsError(v, i))
    r = Ref{Float64}(Float64(val))
    check(ccall((:ak_upload, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), v.ctx.h, v.ptr + 8 * (i - 1), r, 1))
    return v
end
Base.show(io::IO, v::B200Vector) = print(io, "B200Vector", v.dims, " @ ", v.ptr)
Base.show(io::IO, ::MIME"text/plain", v::B200Vector) = show(io, v)

"Column-major n x ncols matrix of device vectors (the `V` / `Out` of the batched `mul!`, src/Ariadne.jl:69-83)."
struct B200Matrix
    data::B200Vector{1}
    n::Int
    ncols::Int
end
B200Matrix(a::Matrix{Float64}; ctx = context()) = B200Matrix(B200Vector(vec(copy(a)); ctx), size(a, 1), size(a, 2))
B200Matrix(::UndefInitializer, n::Integer, ncols::Integer; ctx = context()) =
    B200Matrix(B200Vector{1}(undef, (n * ncols,); ctx), n, ncols)
Base.size(A::B200Matrix) = (A.n, A.ncols)
Base.Array(A::B200Matrix) = reshape(Array(A.data), A.n, A.ncols)

# ---- Krylov.k* hooks (examples/halovector.jl:51-147); ext/AriadneB200KrylovExt.jl attaches them to Krylov.jl's generic
#      functions when Krylov.jl is loaded.  The fused native solver below does not need them. --------------------------
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
Krylov_kdivcopy!(n, y::B200Vector, x::B200Vector, s) = (check(ccall((:ak_divcopy, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Float64), x.ctx.h, n, y.ptr, x.ptr, s)); y)
LinearAlgebra.norm(x::B200Vector) = Krylov_knorm(length(x), x)              # norm(res): src/Ariadne.jl:303,350
LinearAlgebra.dot(x::B200Vector, y::B200Vector) = Krylov_kdot(length(x), x, y)

# ---- native residuals F!(res, u, p) -------------------------------------------------------------------------------------
abstract type NativeResidual end
struct Bratu1D <: NativeResidual end       # examples/bratu.jl:14-24,   p = (Δx, λ)
struct Bratu2D <: NativeResidual end       # 2-D extension,             p = (Δx, Δy, λ) or, for one slab of a
                                           #                            multi-GPU grid, (Δx, Δy, λ, gny, gy0)
# G!(res, uₙ, Δt, f!, du, u, p, t) ∘ f!,  p = (uₙ, Δt, du, p_f, t)  (examples/implicit.jl:8-37,61)
struct GEuler{R} <: NativeResidual; f::R; end       # G_Euler!      implicit.jl:8-13
struct GMidpoint{R} <: NativeResidual; f::R; end    # G_Midpoint!   implicit.jl:17-25
struct GTrapezoid{R} <: NativeResidual; f::R; end   # G_Trapezoid!  implicit.jl:29-37
const ImplicitResidual = Union{GEuler, GMidpoint, GTrapezoid}
scheme_code(::GEuler) = AK_EULER
scheme_code(::GMidpoint) = AK_MIDPOINT
scheme_code(::GTrapezoid) = AK_TRAPEZOID
struct Heat1D end                          # examples/heat_1D.jl:12-25, p_f = (a, Δx, bc) with bc ∈ (:zero, :periodic)
struct Diffusion2D end                     # examples/heat_2D.jl:45-62, p_f = (a, Δx, Δy, bc) or (a, Δx, Δy, bc, gny, gy0)
struct Heat1DDG end                        # examples/heat_1D_DG.jl:32-36, p_f = (h,)

bc_code(bc) = bc === :periodic ? Int32(1) : Int32(0)
devptr(x::B200Vector) = x.ptr
devptr(::Nothing) = Ptr{Float64}(C_NULL)
devptr(p::Ptr{Float64}) = p
problem(::Bratu1D, u, p; coef = C_NULL) =
    AkProblem(AK_BRATU1D, 0, AK_STEADY, 0, length(u), 1, 1, 0, p[1], 0.0, p[2], 0.0, 0.0, 0.0, C_NULL, coef, C_NULL)
function problem(::Bratu2D, u, p; coef = C_NULL)
    gny, gy0 = length(p) >= 5 ? (p[4], p[5]) : (u.dims[2], 0)
    return AkProblem(AK_BRATU2D, 0, AK_STEADY, 0, u.dims[1], u.dims[2], gny, gy0, p[1], p[2], p[3], 0.0, 0.0, 0.0, C_NULL, coef, C_NULL)
end
function problem(F::ImplicitResidual, u, p; coef = C_NULL)
    un, dt, _, pf, _ = p
    sc = scheme_code(F)
    if F.f isa Heat1D
        return AkProblem(AK_HEAT1D, bc_code(pf[3]), sc, 0, length(u), 1, 1, 0, pf[2], 0.0, 0.0, pf[1], dt, 0.0, devptr(un), C_NULL, C_NULL)
    elseif F.f isa Diffusion2D
        gny, gy0 = length(pf) >= 6 ? (pf[5], pf[6]) : (u.dims[2], 0)
        return AkProblem(AK_HEAT2D, bc_code(pf[4]), sc, 0, u.dims[1], u.dims[2], gny, gy0, pf[2], pf[3], 0.0, pf[1], dt, 0.0, devptr(un), C_NULL, C_NULL)
    else
        return AkProblem(AK_HEAT1D_DG, 1, sc, 0, length(u), 1, 1, 0, pf[1], 0.0, 0.0, 0.0, dt, 0.0, devptr(un), C_NULL, C_NULL)
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
    GC.@preserve res u p begin
        check(ccall((:ak_residual, lib), Cint, (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), u.ctx.h, prob, u.ptr, res.ptr, C_NULL))
    end
    return nothing
end

# ---- JacobianOperator (src/Ariadne.jl:34-57) -------------------------------------------------------------------------------
# `coef` is the B200Vector itself (not its pointer): the operator roots the cache for as long as it is alive.
struct JacobianOperator{F, A, P, C}
    f::F; res::A; u::A; p::P
    coef::C          # λ·exp(u) (Bratu) / F(u) (finite-difference tangents) cached by the last residual evaluation, or nothing
    JacobianOperator(f::F, res, u, p; coef = nothing) where {F} = new{F, typeof(u), typeof(p), typeof(coef)}(f, res, u, p, coef)
end
Base.size(J::JacobianOperator) = (length(J.res), length(J.u))
Base.eltype(J::JacobianOperator) = Float64
Base.length(J::JacobianOperator) = prod(size(J))
function mul!(out::B200Vector, J::JacobianOperator, v::B200Vector)
    prob = Ref(problem(J.f, J.u, J.p; coef = devptr(J.coef)))
    GC.@preserve out J v begin
        check(ccall((:ak_jvp, lib), Cint, (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), v.ctx.h, prob, J.u.ptr, v.ptr, out.ptr))
    end
    return nothing
end
# batched form, src/Ariadne.jl:69-83: Out[:, c] = J V[:, c] (one multi-RHS launch for the Bratu operators)
function mul!(Out::B200Matrix, J::JacobianOperator, V::B200Matrix)
    prob = Ref(problem(J.f, J.u, J.p; coef = devptr(J.coef)))
    GC.@preserve Out J V begin
        check(ccall((:ak_jvp_batched, lib), Cint, (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int32),
                    V.data.ctx.h, prob, J.u.ptr, V.data.ptr, V.n, Out.data.ptr, Out.n, V.ncols))
    end
    return nothing
end
# transpose / adjoint, src/Ariadne.jl:87-107
struct TransposedOperator{JOp}; parent::JOp; end
Base.transpose(J::JacobianOperator) = TransposedOperator(J)
Base.adjoint(J::JacobianOperator) = TransposedOperator(J)
Base.size(Jt::TransposedOperator) = reverse(size(Jt.parent))
Base.eltype(::TransposedOperator) = Float64
Base.length(Jt::TransposedOperator) = prod(size(Jt))
function mul!(out::B200Vector, Jt::TransposedOperator, v::B200Vector)
    J = Jt.parent
    prob = Ref(problem(J.f, J.u, J.p; coef = devptr(J.coef)))
    GC.@preserve out J v begin
        check(ccall((:ak_jvp_transpose, lib), Cint, (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), v.ctx.h, prob, J.u.ptr, v.ptr, out.ptr))
    end
    return nothing
end
# collect(J), src/Ariadne.jl:140-162: one product per column (small n; the Python mirror has the colour-probe version)
function Base.collect(JOp::Union{JacobianOperator, TransposedOperator})
    N, M = size(JOp)
    J = JOp isa TransposedOperator ? JOp.parent : JOp
    e = zeros(M)
    v = similar(JOp isa TransposedOperator ? J.res : J.u)
    out = similar(JOp isa TransposedOperator ? J.u : J.res)
    dense = zeros(N, M)
    for j in 1:M
        e .= 0.0
        e[j] = 1.0
        copyto!(v, e)
        Krylov_kfill!(out, 0.0)
        mul!(out, JOp, v)
        dense[:, j] .= vec(Array(out))
    end
    return sparse(dense)
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
    h::Ptr{Cvoid}; proto::B200Vector; niter::Int; solved::Bool; residuals::Vector{Float64}
end
function krylov_workspace(algo::Symbol, res::B200Vector; memory = 20, max_basis = 0)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ak_krylov_create, lib), Cint, (Ptr{Cvoid}, Int32, Int64, Int32, Int64, Ptr{Ptr{Cvoid}}), res.ctx.h, AK_ALGO[algo], length(res), memory, max_basis, r))
    ws = Workspace(r[], res, 0, false, Float64[])
    finalizer(w -> (w.proto.ctx.closed || ccall((:ak_krylov_destroy, lib), Cint, (Ptr{Cvoid},), w.h); nothing), ws)
    return ws
end
"workspace.x: non-owning view of the solution of the last solve"
solution(ws::Workspace) = B200Vector{1}(ccall((:ak_krylov_x, lib), Ptr{Float64}, (Ptr{Cvoid},), ws.h), (length(ws.proto),), ws.proto.ctx)
# Krylov.jl keyword set of the reference call sites: `verbose`, `timemax`, `callback` are accepted and ignored
# (examples/heat_2D.jl:131 passes `verbose = 1`); `history = true` fills `ws.residuals`.
function krylov_solve!(ws::Workspace, J::JacobianOperator, b::B200Vector; atol = √eps(Float64), rtol = √eps(Float64),
                       itmax = 0, restart = false, reorthogonalization = false, history = false, fuse = :sweep,
                       M = nothing, N = nothing, ldiv = false, verbose = 0, timemax = Inf, callback = nothing)
    (N isa TridiagonalLU || M isa TridiagonalLU) && !ldiv && error("ilu(J) is applied with ldiv = true (examples/bratu.jl:126)")
    # caller-supplied preconditioners travel as (object, n) behind a rooted Ref for the duration of the solve
    nref = N isa UserPreconditioner ? Ref{Any}((N, length(b))) : nothing
    mref = M isa UserPreconditioner ? Ref{Any}((M, length(b))) : nothing
    tramp = @cfunction(precond_trampoline, Cint, (Ptr{Cvoid}, UInt64, Ptr{Float64}, Ptr{Float64}))
    fields(P, r) = r === nothing ? precond_fields(P) : (AK_PRECOND_USER, Int32(0), tramp, pointer_from_objref(r))
    pn, pit, nfn, nus = fields(N, nref)
    pm, pmit, mfn, mus = fields(M, mref)
    o = Ref(AkKrylovOpts(atol, rtol, itmax, restart, reorthogonalization, history, fuse_code(fuse), pn, pit, pm, pmit, nfn, nus, mfn, mus))
    st = Ref(AkKrylovStats(0, 0, 0, 0, 0, 0.0, 0.0))
    prob = Ref(problem(J.f, J.u, J.p; coef = devptr(J.coef)))
    hist = history ? zeros(min((itmax == 0 ? 2 * length(b) : itmax) + 1, 1 << 20)) : Float64[]
    GC.@preserve nref mref J b hist begin
        check(ccall((:ak_krylov_solve, lib), Cint,
                    (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Ptr{AkKrylovOpts}, Ptr{AkKrylovStats}, Ptr{Float64}, Int64),
                    ws.h, prob, J.u.ptr, b.ptr, o, st, history ? pointer(hist) : Ptr{Float64}(C_NULL), length(hist)))
    end
    ws.niter, ws.solved = st[].niter, st[].solved != 0
    ws.residuals = history ? hist[1:min(length(hist), ws.niter + 1)] : Float64[]
    return ws
end

# ---- multi-GPU: one Julia process per GPU; `id` (128 bytes) comes from rank 0 through Distributed / MPI ------------------
comm_unique_id() = (id = zeros(UInt8, 128); check(ccall((:ak_comm_unique_id, lib), Cint, (Ptr{UInt8},), id)); id)
comm_init(ctx::Context, nranks, rank, id::Vector{UInt8}) =
    check(ccall((:ak_comm_init, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx.h, nranks, rank, id))
comm_enable_p2p(ctx::Context, halo_doubles) =
    check(ccall((:ak_comm_enable_p2p, lib), Cint, (Ptr{Cvoid}, Int64), ctx.h, halo_doubles))
comm_use_p2p(ctx::Context, on::Bool) = check(ccall((:ak_comm_use_p2p, lib), Cint, (Ptr{Cvoid}, Cint), ctx.h, on))
comm_barrier(ctx::Context) = check(ccall((:ak_comm_barrier, lib), Cint, (Ptr{Cvoid},), ctx.h))

# ---- newton_krylov! (src/Ariadne.jl:288-372): the loop is driven from Julia, one ccall per arrowed line ----------------------------
wants_coef(F!) = F! isa Union{Bratu1D, Bratu2D} || (F! isa UserResidual && F!.jvp! === nothing)
function newton_krylov!(F!::NativeResidual, u::B200Vector, p = nothing, res::B200Vector = zero(u);
                        tol_rel = 1.0e-6, tol_abs = 1.0e-12, max_niter = 50,
                        forcing::Union{Forcing, Nothing} = EisenstatWalker(), verbose = 0, algo = :gmres,
                        M = nothing, N = nothing, krylov_kwargs = (;), callback = (args...) -> nothing)
    t₀ = time_ns()
    # λ·exp(u) cache shared by residual and JVPs (Bratu); F(u) cache for finite-difference JVPs (user F! without tangent)
    coef = wants_coef(F!) ? similar(u) : nothing
    prob = Ref(problem(F!, u, p; coef = devptr(coef)))
    nrm = Ref{Float64}(0.0)
    residual_norm() = (check(ccall((:ak_residual, lib), Cint, (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                   u.ctx.h, prob, u.ptr, res.ptr, nrm)); nrm[])
    J = JacobianOperator(F!, res, u, p; coef)                         # src/Ariadne.jl:314 — roots `coef`
    workspace = krylov_workspace(algo, res)                           # :317-318
    rhs = similar(res)
    n_res = 0.0
    tol = 0.0
    stats = Stats(0, 0, 0.0)
    # every device buffer whose raw pointer the library holds stays rooted until the loop is over
    GC.@preserve coef rhs workspace J u res p begin
        n_res = residual_norm()                                       # :302-303
        callback(u, res, n_res)                                       # :304
        tol = tol_rel * n_res + tol_abs                               # :306
        η = forcing === nothing ? nothing : inital(forcing)           # :308-310
        verbose > 0 && @info "Jacobian-Free Newton-Krylov" algo res₀ = n_res tol η
        stats = Stats(0, 0, n_res)                                    # :320
        while n_res > tol && stats.outer_iterations <= max_niter     # :321
            kwargs = krylov_kwargs
            N !== nothing && (kwargs = (; N = N(J), kwargs...))       # :324-326
            M !== nothing && (kwargs = (; M = M(J), kwargs...))       # :327-329
            forcing !== nothing && (kwargs = (; rtol = η, kwargs...)) # :330-333 (later keys win)
            Krylov_kcopy!(length(res), rhs, res)                      # copy(res) :338
            krylov_solve!(workspace, J, rhs; kwargs...)
            d = solution(workspace)                                   # :340
            Krylov_kaxpy!(length(u), -1.0, d, u)                      # u .-= s .* d, s = 1  :341-344
            n_res_prior = n_res
            n_res = residual_norm()                                   # :349-350
            callback(u, res, n_res)                                   # :351
            if isinf(n_res) || isnan(n_res)                           # :353-356
                @error "Inner solver blew up" stats
                break
            end
            forcing !== nothing && (η = forcing(η, tol, n_res, n_res_prior))   # :358-360
            stats = update(stats, workspace.niter, n_res)             # :367
            verbose > 0 && @info "Newton" iter = n_res η stats
        end
    end
    t = (time_ns() - t₀) / 1.0e9
    return u, (; solved = n_res <= tol, stats, t)
end
newton_krylov(F::NativeResidual, u₀::B200Vector, p = nothing; kwargs...) = newton_krylov!(F, copy(u₀), p; kwargs...)

# ---- the same solve through ONE C entry point (the loop runs in C++): device vectors / host Arrays ---------------------------------------
function newton_opts(; tol_rel = 1.0e-6, tol_abs = 1.0e-12, max_niter = 50, forcing::Union{Forcing, Nothing} = EisenstatWalker(),
                     verbose = 0, algo = :gmres, memory = 20, max_basis = 0, krylov_kwargs = (;))
    kk = Dict{Symbol, Any}(pairs(krylov_kwargs))
    for k in (:verbose, :timemax, :callback)   # Krylov.jl keywords without a native counterpart (examples/heat_2D.jl:131)
        delete!(kk, k)
    end
    ko = AkKrylovOpts(get(kk, :atol, √eps(Float64)), get(kk, :rtol, √eps(Float64)), get(kk, :itmax, 0),
                      get(kk, :restart, false), get(kk, :reorthogonalization, false), get(kk, :history, false),
                      fuse_code(get(kk, :fuse, :sweep)), AK_PRECOND_NONE, 0, AK_PRECOND_NONE, 0, C_NULL, C_NULL, C_NULL, C_NULL)
    fc, η, ηmax, γ = forcing === nothing ? (AK_FORCING_NONE, 0.1, 0.999, 0.9) :
                     forcing isa Fixed ? (AK_FORCING_FIXED, forcing.η, 0.999, 0.9) : (AK_FORCING_EW, 0.1, forcing.η_max, forcing.γ)
    return AkNewtonOpts(tol_rel, tol_abs, max_niter, fc, η, ηmax, γ, AK_ALGO[algo], memory, max_basis, ko,
                        haskey(kk, :rtol) ? 1 : 0, verbose)
end
function newton_krylov_native!(F!::NativeResidual, u::B200Vector, p = nothing, res::B200Vector = zero(u); kwargs...)
    coef = wants_coef(F!) ? similar(u) : nothing
    prob = Ref(problem(F!, u, p; coef = devptr(coef)))
    o = Ref(newton_opts(; kwargs...))
    st = Ref(AkNewtonStats(0, 0, 0, 0.0, 0.0, 0.0, 0))
    GC.@preserve coef u res p begin
        check(ccall((:ak_newton_solve, lib), Cint,
                    (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Ptr{AkNewtonOpts}, Ptr{AkNewtonStats},
                     Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Int32, Ptr{Cvoid}, Ptr{Cvoid}),
                    u.ctx.h, prob, u.ptr, res.ptr, o, st, C_NULL, C_NULL, C_NULL, 0, C_NULL, C_NULL))
    end
    s = st[]
    return u, (; solved = s.solved != 0, stats = Stats(s.outer_iterations, s.inner_iterations, s.n_res), t = s.t_seconds)
end
# `newton_krylov!(F!, u₀::Array, p)` — src/Ariadne.jl:259-263: the caller holds a host Array; the library uploads it,
# solves and writes the solution back (ak_newton_solve_host).  For the implicit residuals p[1] = uₙ may be a host Array too.
function newton_krylov!(F!::NativeResidual, u::Array{Float64}, p = nothing; ctx::Context = context(), kwargs...)
    F! isa UserResidual && error("UserResidual works on device vectors: pass a B200Vector")
    shape = B200Vector{ndims(u)}(Ptr{Float64}(C_NULL), size(u), ctx)          # carries the dims only
    un_host = (F! isa ImplicitResidual && p[1] isa Array{Float64}) ? p[1] : nothing
    pdev = un_host === nothing ? p : (nothing, p[2:end]...)
    prob = Ref(problem(F!, shape, pdev))
    o = Ref(newton_opts(; kwargs...))
    st = Ref(AkNewtonStats(0, 0, 0, 0.0, 0.0, 0.0, 0))
    GC.@preserve u un_host begin
        check(ccall((:ak_newton_solve_host, lib), Cint,
                    (Ptr{Cvoid}, Ptr{AkProblem}, Ptr{Float64}, Ptr{Float64}, Ptr{AkNewtonOpts}, Ptr{AkNewtonStats},
                     Ptr{Float64}, Ptr{Int64}, Int32),
                    ctx.h, prob, u, un_host === nothing ? Ptr{Float64}(C_NULL) : pointer(un_host), o, st, C_NULL, C_NULL, 0))
    end
    s = st[]
    return u, (; solved = s.solved != 0, stats = Stats(s.outer_iterations, s.inner_iterations, s.n_res), t = s.t_seconds)
end
newton_krylov(F::NativeResidual, u₀::Array{Float64}, p = nothing; kwargs...) = newton_krylov!(F, copy(u₀), p; kwargs...)

# ---- solve(G!, f!, uₙ, p, Δt, ts) (examples/implicit.jl:54-78) -------------------------------------------------------------------------
function solve(G::ImplicitResidual, uₙ::B200Vector, p, Δt, ts; callback = _ -> nothing, verbose = 0, algo = :gmres, krylov_kwargs = (;))
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
