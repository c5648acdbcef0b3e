"""Transcribes the reference's own known answers for the hot path into a JSON fixture.

The reference (Julia + Krylov.jl + Enzyme.jl) cannot be executed in this image (no julia
binary, dependencies not vendored), so these are the literal values asserted by the reference's
test-suite and examples, with the file:line each comes from.  Run:  python tests/golden/make_golden.py
"""
import json
import os

GOLDEN = {
    "source": "vchuravy/NewtonKrylov.jl test/runtests.jl, examples/bratu.jl, src/Ariadne.jl",
    "jacobian_2x2": {
        "cite": "test/runtests.jl:28-42",
        "u": [3.0, 5.0],
        "v": [1.0, 0.0],
        "J_times_v": [6.0, 7.38905609893065],       # @test out == [6.0, 7.38905609893065]   :36-38
        "Jt_times_v": [6.0, 10.0],                   # @test out == [6.0, 10.0]               :40-42
        "size": [2, 2], "length": 4, "eltype": "Float64",  # :32-34
    },
    "newton_2x2": {
        "cite": "test/runtests.jl:15-23",
        "cases": [
            {"x0": [2.0, 0.5], "solved": True},     # newton_krylov!(F!, x0)  -> @test stats.solved
            {"x0": [3.0, 5.0], "solved": True},     # newton_krylov(F, x0)    -> @test stats.solved
        ],
    },
    "newton_defaults": {
        "cite": "src/Ariadne.jl:290-299,185-200",
        "tol_rel": 1.0e-6, "tol_abs": 1.0e-12, "max_niter": 50, "eta_max": 0.999, "gamma": 0.9, "fixed_eta": 0.1,
    },
    "bratu_analytic": {
        "cite": "examples/bratu.jl:33-46",
        "lambda": 3.51382, "theta": 4.79173, "N": 10000,
        "formula": "-2*log(cosh(theta*(x-0.5)/2)/cosh(theta/4))",
    },
    "heat_1d_params": {"cite": "examples/heat_1D.jl:46,111-115", "L": 1.0, "M": 100, "a": 0.2, "dt": 0.1, "t_final": 3.0,
                       "ic": "4x(1-x)", "tol_abs": 6.0e-6},
    "heat_2d_params": {"cite": "examples/heat_2D.jl:64-72,83-91,131", "a": 0.01, "N": 40, "M": 40,
                       "ic": "sin(pi x) sin(pi y)", "reorthogonalization": True},
    "dg_params": {"cite": "examples/heat_1D_DG.jl:14-25,40,81-82", "polydeg": 3, "elements": 40, "dt": 0.01,
                  "t_final": 50.0, "ic": "sin(pi x)",
                  "invariants": "docs/src/notebooks/heat_1D_DG.jl:154,157: 2*J_M + I == J_E ; J_M == J_T"},
    "survey_probe_predictions": {
        "cite": "SURVEY.md §6 (NumPy restatement during the survey; NOT reference output)",
        "newton_2x2_[2,0.5]": {"outer": 10, "inner": [1, 1, 2, 1, 1, 2, 1, 1, 1, 2], "n_res_final": 7.15e-07},
        "newton_2x2_[3,5]": {"outer": 7, "inner": [1, 1, 1, 1, 2, 2, 2], "u": [-0.47767, 1.33110]},
        "bratu1d_lambda3.5_N1000": {"outer": 9, "inner": [1, 15, 49, 116, 317, 498, 460, 501, 409]},
        "bratu2d_lambda3.5": {"N": [32, 64, 128], "outer": [7, 8, 8], "inner_total": [123, 245, 477]},
        "heat1d_M100": {"newton_per_step": 7, "inner_first_step": [1, 4, 9, 14, 22, 40, 50]},
    },
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "runtests_known_answers.json")
    with open(out, "w") as f:
        json.dump(GOLDEN, f, indent=1, sort_keys=True)
    print(out)
