# The call sites of the reference's examples/bratu.jl (lines 40-63, 121-157) through AriadneB200.
using AriadneB200

N = 10_000
λ = 3.5
dx = 1 / (N + 1)
x = collect(LinRange(0.0 + dx, 1.0 - dx, N))
u₀ = sin.(x .* π)

# bratu.jl:59-63 — algo = :cg
_, stats = newton_krylov!(Bratu1D(), B200Vector(copy(u₀)), (dx, λ); algo = :cg)
@show stats
# bratu.jl:121-139 — N = J -> ilu(collect(J)), ldiv = true  (tridiagonal LU on the device)
_, stats = newton_krylov!(Bratu1D(), B200Vector(copy(u₀)), (dx, λ); algo = :gmres, N = J -> ilu(J),
                          krylov_kwargs = (; ldiv = true))
@show stats
# bratu.jl:141-157 — algo = :fgmres, N = J -> GmresPreconditioner(J, 5)
_, stats = newton_krylov!(Bratu1D(), B200Vector(copy(u₀)), (dx, λ); algo = :fgmres, N = J -> GmresPreconditioner(J, 5))
@show stats
