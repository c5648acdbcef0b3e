# baseline/reference_cpu.jl — the REAL reference calls for each BASELINE.json config.
#
# Cannot be run in the build image or on the GPU box (no julia, no network for Pkg); shipped so that anyone with
# Julia ≥ 1.10 and `] add Ariadne Krylov Enzyme OffsetArrays SummationByPartsOperators` can close the
# "parity unpinned" gap of DESIGN.md §6: it prints, per Newton iteration, ‖F‖ and the GMRES iteration count in the
# same format as tests/test_gpu_solvers.py compares.
using Ariadne, Krylov, LinearAlgebra, Printf

function bratu!(res, y, (Δx, λ))                       # examples/bratu.jl:14-24 (verbatim semantics)
    N = length(y)
    for i in 1:N
        y_l = i == 1 ? zero(eltype(y)) : y[i - 1]
        y_r = i == N ? zero(eltype(y)) : y[i + 1]
        res[i] = (y_r - 2y[i] + y_l) / Δx^2 + λ * exp(y[i])
    end
    return nothing
end

# 2-D Bratu as defined by this repository (DESIGN.md §2): u is nx × ny, x = first (fastest) index
function bratu2d!(res, u, (Δx, Δy, λ))
    nx, ny = size(u)
    for j in 1:ny, i in 1:nx
        w = i == 1 ? 0.0 : u[i - 1, j]; e = i == nx ? 0.0 : u[i + 1, j]
        s = j == 1 ? 0.0 : u[i, j - 1]; n = j == ny ? 0.0 : u[i, j + 1]
        res[i, j] = ((e - 2u[i, j] + w) / Δx^2 + (n - 2u[i, j] + s) / Δy^2) + λ * exp(u[i, j])
    end
    return nothing
end

function report(name, F!, u₀, p; kwargs...)
    hist = Tuple{Float64}[]
    cb(u, res, n_res) = push!(hist, (n_res,))
    t = @elapsed (u, stats) = newton_krylov!(F!, copy(u₀), p, similar(u₀); callback = cb, kwargs...)
    @printf("%s solved=%s outer=%d inner=%d t=%.3fs\n", name, stats.solved, stats.stats.outer_iterations, stats.stats.inner_iterations, t)
    foreach(h -> @printf("   %.15e\n", h[1]), hist)
    return u
end

# config 1: 1-D Bratu, λ = 3.5, N = 10_000 (CPU reference path)
let N = 10_000, λ = 3.5, dx = 1 / (N + 1)
    x = LinRange(dx, 1 - dx, N)
    report("bratu1d", bratu!, sin.(x .* π), (dx, λ); algo = :gmres)
end
# config 4 (protocol B): 2-D Bratu to convergence at N ≤ 2048; protocol A at 8192² uses krylov_kwargs = (; restart = true, itmax = 200)
for N in (32, 64, 128)
    dx = 1 / (N + 1); x = dx .* (1:N)
    report("bratu2d_$N", bratu2d!, sin.(π .* x) * sin.(π .* x)', (dx, dx, 3.5); algo = :gmres)
end
# configs 2, 3, 5: run examples/heat_1D.jl, heat_2D.jl, heat_1D_DG.jl of the reference unchanged.
