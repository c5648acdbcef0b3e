# The call sites of the reference's examples/heat_2D.jl (lines 64-91, 131, 151-154) through AriadneB200:
# same parameters, same `solve` / `krylov_kwargs`; the HaloVector + OffsetArray state becomes a B200Vector
# (compact slab: the ghost ring is implicit, see include/ariadne_b200.h).
using AriadneB200

a = 0.01
N = M = 40
Δx = 1 / (N + 1)
Δy = 1 / (M + 1)
Δt = Δx^2 * Δy^2 / (2.0 * a * (Δx^2 + Δy^2))
xs = Δx .* (1:N)
ys = Δy .* (1:M)
u₀ = [sin(π * x) * sin(π * y) for x in xs, y in ys]          # interior of heat_2D.jl:83-91

u = B200Vector(copy(u₀))
solve(GEuler(Diffusion2D()), u, (a, Δx, Δy, :zero), Δt, 0.0:Δt:10Δt;
      verbose = 1, krylov_kwargs = (; verbose = 1, reorthogonalization = true))     # heat_2D.jl:131
u = B200Vector(copy(u₀))
solve(GEuler(Diffusion2D()), u, (a, Δx, Δy, :periodic), Δt, 0.0:Δt:2Δt;
      verbose = 1, krylov_kwargs = (; verbose = 1, reorthogonalization = true))     # heat_2D.jl:151-154
