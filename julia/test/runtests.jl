# Tests of the Julia wrapper, to be run where `julia`, a B200 and the built library exist
# (`julia --project=julia -e 'using Pkg; Pkg.test()'`).  NOT executed in the build image (no `julia` binary);
# the same checks run there through the Python mirror (tests/test_gpu_*.py), and tests/test_julia_wrapper.py
# verifies this package's structs and `ccall`s against include/ariadne_b200.h without Julia.
#
# Layout follows the reference's test/runtests.jl: convergence smoke tests, then JacobianOperator known answers.
using Test
using AriadneB200
using LinearAlgebra

@testset "AriadneB200" begin
    @testset "2-D Bratu, device vector and host Array entry points" begin
        N = 96
        Δ = 1 / (N + 1)
        x = Δ .* (1:N)
        u₀ = sin.(π .* x) .* sin.(π .* x)'
        u, stats = newton_krylov!(Bratu2D(), B200Vector(copy(u₀)), (Δ, Δ, 3.5))
        @test stats.solved
        v, stats_h = newton_krylov!(Bratu2D(), copy(u₀), (Δ, Δ, 3.5))        # src/Ariadne.jl:259-263 shape
        @test stats_h.solved
        @test stats_h.stats.outer_iterations == stats.stats.outer_iterations
        @test norm(vec(Array(u)) .- vec(v)) <= 1.0e-8 * norm(vec(v))
    end

    @testset "JacobianOperator protocol (test/runtests.jl:28-54 shapes)" begin
        N = 50
        Δ = 1 / (N + 1)
        u = B200Vector(sin.(π .* collect(LinRange(Δ, 1 - Δ, N))))
        res = zero(u)
        J = JacobianOperator(Bratu1D(), res, u, (Δ, 3.5))
        @test size(J) == (N, N)
        @test length(J) == N * N
        @test eltype(J) == Float64
        Jd = collect(J)
        @test collect(transpose(J)) == transpose(Jd)
        v = rand(N)
        out = zero(u)
        mul!(out, J, B200Vector(copy(v)))
        @test Array(out) ≈ Jd * v
        V = rand(N, 4)
        Out = B200Matrix(undef, N, 4)
        mul!(Out, J, B200Matrix(V))
        @test Array(Out) ≈ Jd * V
    end

    @testset "implicit heat 2-D with the reference's Krylov kwargs (examples/heat_2D.jl:131)" begin
        N = 40
        a = 0.01
        Δ = 1 / (N + 1)
        Δt = Δ^2 * Δ^2 / (2 * a * (Δ^2 + Δ^2))
        x = Δ .* (1:N)
        uₙ = B200Vector(sin.(π .* x) .* sin.(π .* x)')
        solve(GEuler(Diffusion2D()), uₙ, (a, Δ, Δ, :zero), Δt, 0.0:Δt:(3Δt);
              verbose = 0, krylov_kwargs = (; verbose = 1, reorthogonalization = true))
        @test !any(isnan, Array(uₙ))
        solve(GTrapezoid(Diffusion2D()), uₙ, (a, Δ, Δ, :periodic), Δt, 0.0:Δt:(2Δt))
        @test !any(isnan, Array(uₙ))
    end
end
