"""The C ABI from plain C: examples/c/bratu2d_newton.c compiles against include/ariadne_b200.h with the host C
compiler, links libariadne_b200.so, fails loudly without a CUDA device and solves the 2-D Bratu problem with one."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples", "c")
BIN = os.path.join(EX, "bratu2d_newton")


@pytest.fixture(scope="module")
def binary(nk):
    subprocess.run(["make", "-C", EX, "-B", "bratu2d_newton", "CC=gcc"], check=True, stdout=subprocess.PIPE,
                   stderr=subprocess.STDOUT)
    assert os.path.exists(BIN)
    return BIN


def test_c_consumer_builds_and_refuses_to_run_without_a_gpu(binary):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([binary, "32"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_consumer_solves_bratu2d(binary, oracle):
    import numpy as np

    import problems as P

    r = subprocess.run([binary, "96", "3.5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    last = [line for line in r.stdout.splitlines() if line.startswith("solved")][0]
    fields = dict(zip(last.split()[0::3], last.split()[2::3]))
    d = P.bratu2d(96)
    ur, sr, hr = oracle.newton(P.oracle_problem(oracle, d), d["u0"])
    assert int(fields["solved"]) == 1 and int(fields["outer"]) == sr["outer_iterations"]
    umax = float([line for line in r.stdout.splitlines() if line.startswith("max u")][0].split("=")[1])
    assert abs(umax - float(np.max(ur))) < 1e-7
