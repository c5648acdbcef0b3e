"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU,
exports every symbol include/ariadne_b200.h declares, the ctypes struct mirrors match the C
layouts, and the product refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from newtonkrylov_jl_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ariadne_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(ak_[a-z0-9_]+)\s*\(", src)
    # function-pointer typedefs are not symbols
    return sorted(set(n for n in names if n != "ak_newton_callback"))


def test_library_exports_every_declared_symbol(nk):
    lib = nk._lib.load()
    names = declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/ariadne_b200.h but not exported"
    assert set(nk._lib.SIGNATURES) == set(names), set(nk._lib.SIGNATURES) ^ set(names)
    assert lib.ak_abi_version() == A.ABI_VERSION


def test_struct_layouts_match_the_header():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "ariadne_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(ak_problem), sizeof(ak_krylov_opts), sizeof(ak_krylov_stats),
         sizeof(ak_newton_opts), sizeof(ak_newton_stats));
  printf("%zu %zu %zu %zu\n", offsetof(ak_problem, dx), offsetof(ak_problem, un), offsetof(ak_newton_opts, krylov),
         offsetof(ak_newton_stats, t_seconds));
  printf("%zu %zu %zu %zu %zu %zu %zu\n", offsetof(ak_problem, user_residual), offsetof(ak_problem, user_data),
         offsetof(ak_krylov_opts, precond_m), offsetof(ak_krylov_opts, n_apply), offsetof(ak_krylov_opts, m_user),
         offsetof(ak_newton_opts, krylov_rtol_override), offsetof(ak_newton_opts, verbose));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        cfile = os.path.join(td, "t.c")
        open(cfile, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes[:5] == [C.sizeof(A.ak_problem), C.sizeof(A.ak_krylov_opts), C.sizeof(A.ak_krylov_stats),
                         C.sizeof(A.ak_newton_opts), C.sizeof(A.ak_newton_stats)]
    assert sizes[5:9] == [A.ak_problem.dx.offset, A.ak_problem.un.offset, A.ak_newton_opts.krylov.offset,
                          A.ak_newton_stats.t_seconds.offset]
    # ABI 3: caller-supplied residual / tangent callbacks and the preconditioner hooks
    assert sizes[9:] == [A.ak_problem.user_residual.offset, A.ak_problem.user_data.offset,
                         A.ak_krylov_opts.precond_m.offset, A.ak_krylov_opts.n_apply.offset,
                         A.ak_krylov_opts.m_user.offset, A.ak_newton_opts.krylov_rtol_override.offset,
                         A.ak_newton_opts.verbose.offset]


def test_header_is_plain_c_and_cites_the_reference():
    src = open(HEADER).read()
    assert 'extern "C"' in src
    assert "at::" not in src and "Tensor" not in src  # plain pointers and sizes only, no torch types
    for cite in ("src/Ariadne.jl:288-372", "src/Ariadne.jl:34-57", "examples/halovector.jl:51-147",
                 "examples/implicit.jl:54-78", "examples/bratu.jl:14-24", "examples/heat_2D.jl:45-62"):
        assert cite in src, cite
    # compiles as C
    with tempfile.TemporaryDirectory() as td:
        cfile = os.path.join(td, "t.c")
        open(cfile, "w").write('#include "ariadne_b200.h"\nint main(void){return 0;}\n')
        subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", cfile,
                        "-o", os.path.join(td, "t.o")], check=True)


def test_cuda_binary_targets_sm_100a(nk):
    nk._lib.load()
    r = subprocess.run(["cuobjdump", "--list-elf", nk.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in r.stdout


def test_default_options_match_between_library_and_mirror(nk):
    lib = nk._lib.load()
    o = A.ak_newton_opts()
    lib.ak_newton_default_opts(C.byref(o))
    m = A.default_newton_opts()
    for f, _ in A.ak_newton_opts._fields_:
        if f == "krylov":
            continue
        assert getattr(o, f) == getattr(m, f), f
    for f, _ in A.ak_krylov_opts._fields_:
        assert getattr(o.krylov, f) == getattr(m.krylov, f), f


def test_forcing_matches_oracle_and_reference_formula(nk, oracle):
    """EisenstatWalker / Fixed / inital — src/Ariadne.jl:185-217 (pure host scalar code, no GPU needed)."""
    ew = nk.EisenstatWalker()
    assert nk.inital(ew) == 0.999 and nk.inital(nk.Fixed()) == 0.1 and nk.Fixed(0.3)(1, 2, 3, 4) == 0.3
    import numpy as np
    rng = np.random.default_rng(2)
    for _ in range(100):
        eta, tol = rng.random(), 10 ** rng.uniform(-12, -3)
        nr, nrp = 10 ** rng.uniform(-8, 2), 10 ** rng.uniform(-8, 2)
        assert ew(eta, tol, nr, nrp) == oracle.forcing_ew(0.999, 0.9, eta, tol, nr, nrp)
    s = nk.update(nk.Stats(0, 0, 1.0), 5, 0.5)
    assert s == nk.Stats(1, 5, 0.5)


def test_no_cpu_fallback(nk):
    """Without a CUDA device the product fails loudly instead of computing on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nk.AriadneError) as e:
        nk.Context(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "newtonkrylov.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "nk_oracle" not in txt and "ok_gmres" not in txt, f
