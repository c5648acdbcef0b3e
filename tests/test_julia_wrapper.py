"""CPU checks of the Julia host side (julia/src/AriadneB200.jl) against the C ABI, without a `julia` binary:

* every `struct Ak*` has the field ORDER and field TYPES of the ctypes mirror (`_abi.py`), which tests/test_abi.py in
  turn checks against the compiled header;
* every `ccall((:sym, lib), Ret, (Args...), ...)` names a symbol that include/ariadne_b200.h declares, with the same
  number of arguments and the same kind (integer width / double / pointer) for each, and the same return kind;
* the seams VERDICT r1 listed as missing exist: the Krylov.k* methods are attached in a package extension, the host
  `Array` entry point, Midpoint/Trapezoid, the slab parameters gny/gy0, the Krylov.jl keywords of the reference's
  call sites (`verbose`, `history`, `timemax`), scalar indexing, transpose / collect / batched mul!.
"""
import ctypes as C
import os
import re

from newtonkrylov_jl_b200 import _abi as A
from newtonkrylov_jl_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "julia", "src", "AriadneB200.jl")
EXT = os.path.join(ROOT, "julia", "ext", "AriadneB200KrylovExt.jl")


def source():
    src = open(JL, encoding="utf-8").read()
    # strip comments (a `#` inside a string does not occur in this file)
    return "\n".join(line.split("#", 1)[0] for line in src.splitlines())


def jl_kind(t):
    t = t.strip()
    if t.startswith("Ptr{") or t in ("Cstring",):
        return "ptr"
    return {"Int32": "i32", "Cint": "i32", "Int64": "i64", "UInt64": "u64", "Float64": "f64", "Cvoid": "void"}[t]


def c_kind(t):
    if t is None:
        return "void"
    if t in (C.c_int32, C.c_int):
        return "i32"
    if t in (C.c_int64, C.c_long, C.c_longlong):
        return "i64"
    if t is C.c_uint64:
        return "u64"
    if t is C.c_double:
        return "f64"
    if t in (C.c_void_p, C.c_char_p) or isinstance(t, type(C.POINTER(C.c_int))) or hasattr(t, "_type_") and t.__name__.startswith("LP_"):
        return "ptr"
    if isinstance(t, type) and issubclass(t, C.Structure):
        return "struct:" + t.__name__
    if hasattr(t, "_flags_"):  # CFUNCTYPE prototypes: function pointers
        return "ptr"
    raise AssertionError(f"unmapped ctypes type {t!r}")


def split_top(s):
    """split on commas that are not nested inside braces"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out if x.strip()]


def test_struct_field_order_and_types_match_the_abi():
    src = source()
    pairs = {"AkProblem": A.ak_problem, "AkKrylovOpts": A.ak_krylov_opts, "AkKrylovStats": A.ak_krylov_stats,
             "AkNewtonOpts": A.ak_newton_opts, "AkNewtonStats": A.ak_newton_stats}
    for jname, cstruct in pairs.items():
        m = re.search(r"\bstruct " + jname + r"\b(.*?)\bend\b", src, flags=re.S)
        assert m, jname
        fields = re.findall(r"(\w+)::([\w{}]+)", m.group(1))
        cfields = list(cstruct._fields_)
        assert len(fields) == len(cfields), (jname, len(fields), len(cfields))
        for (fn, ft), (cn, ct) in zip(fields, cfields):
            assert fn == cn.rstrip("_"), (jname, fn, cn)          # `lambda_` in Python (keyword), `lambda` in C / Julia
            if ft.startswith("Ak"):
                assert c_kind(ct) == "struct:" + {"AkKrylovOpts": "ak_krylov_opts"}[ft], (jname, fn)
            else:
                assert jl_kind(ft) == c_kind(ct), (jname, fn, ft, ct)


def ccalls(src):
    flat = re.sub(r"\s+", " ", src)
    out = []
    for m in re.finditer(r"ccall\(\(:(\w+), lib\), ([\w{}]+), \(([^()]*)\)", flat):
        sym, ret, args = m.group(1), m.group(2), split_top(m.group(3))
        out.append((sym, ret, args))
    return out


def test_every_ccall_matches_the_declared_signature():
    calls = ccalls(source())
    assert len(calls) >= 30
    seen = set()
    for sym, ret, args in calls:
        assert sym in L.SIGNATURES, f"ccall of :{sym}, which include/ariadne_b200.h does not declare"
        cres, cargs = L.SIGNATURES[sym]
        assert len(args) == len(cargs), (sym, args, cargs)
        assert jl_kind(ret) == c_kind(cres), (sym, ret, cres)
        for i, (ja, ca) in enumerate(zip(args, cargs)):
            assert jl_kind(ja) == c_kind(ca), (sym, i, ja, ca)
        seen.add(sym)
    # the entry points every seam of INTEGRATION.md §1 needs are actually bound
    for need in ("ak_residual", "ak_jvp", "ak_jvp_transpose", "ak_jvp_batched", "ak_dot", "ak_nrm2", "ak_scal", "ak_axpy",
                 "ak_axpby", "ak_copy", "ak_fill", "ak_ref", "ak_divcopy", "ak_krylov_create", "ak_krylov_solve",
                 "ak_krylov_x", "ak_krylov_destroy", "ak_newton_solve", "ak_newton_solve_host", "ak_forcing_ew",
                 "ak_malloc", "ak_free", "ak_upload", "ak_download", "ak_ctx_create", "ak_ctx_destroy",
                 "ak_comm_unique_id", "ak_comm_init", "ak_comm_enable_p2p"):
        assert need in seen, need


def test_enum_values_match_the_abi():
    src = source()
    assert "const AK_SIMPLE2, AK_BRATU1D, AK_BRATU2D, AK_HEAT1D, AK_HEAT2D, AK_HEAT1D_DG, AK_USER = Int32.(0:6)" in src
    assert "const AK_STEADY, AK_EULER, AK_MIDPOINT, AK_TRAPEZOID = Int32.(0:3)" in src
    assert (A.AK_STEADY, A.AK_EULER, A.AK_MIDPOINT, A.AK_TRAPEZOID) == (0, 1, 2, 3)
    assert "const AK_JVP_ANALYTIC, AK_JVP_FD_FUSED, AK_JVP_FD = Int32.(0:2)" in src
    assert (A.AK_JVP_ANALYTIC, A.AK_JVP_FD_FUSED, A.AK_JVP_FD) == (0, 1, 2)
    m = re.search(r"const AK_FUSE = Dict\((.*?)\)\n", src)
    fuse = dict((k, int(v)) for k, v in re.findall(r":(\w+) => Int32\((\d)\)", m.group(1)))
    assert fuse == {"none": A.AK_FUSE_NONE, "mgs": A.AK_FUSE_MGS, "full": A.AK_FUSE_FULL, "pair": A.AK_FUSE_PAIR,
                    "block4": A.AK_FUSE_BLOCK4, "block8": A.AK_FUSE_BLOCK8, "sweep": A.AK_FUSE_SWEEP}
    m = re.search(r"const AK_ALGO = Dict\((.*?)\)\n", src)
    assert dict((k, int(v)) for k, v in re.findall(r":(\w+) => Int32\((\d)\)", m.group(1))) == \
        {"gmres": A.AK_ALGO_GMRES, "cg": A.AK_ALGO_CG, "fgmres": A.AK_ALGO_FGMRES}
    assert "AK_PRECOND_NONE, AK_PRECOND_INNER_GMRES, AK_PRECOND_USER, AK_PRECOND_JACOBI, AK_PRECOND_TRIDIAG_LU = Int32.(0:4)" in src
    assert "const AK_FORCING_NONE, AK_FORCING_FIXED, AK_FORCING_EW = Int32.(0:2)" in src


def test_krylov_hooks_are_attached_in_a_package_extension():
    """examples/halovector.jl:51-147 overloads Krylov.kdot ... Krylov.kref! for its vector type; so does the extension."""
    ext = open(EXT, encoding="utf-8").read()
    for hook in ("kdot", "knorm", "kscal!", "kaxpy!", "kaxpby!", "kcopy!", "kfill!", "kref!"):
        assert re.search(r"^Krylov\." + re.escape(hook) + r"\(", ext, flags=re.M), hook
    toml = open(os.path.join(ROOT, "julia", "Project.toml")).read()
    assert "[weakdeps]" in toml and "[extensions]" in toml and 'AriadneB200KrylovExt = "Krylov"' in toml
    # argument order of the reference's methods: kcopy!(n, y, x), kaxpby!(n, s, x, t, y), kref!(n, x, y, c, s)
    assert "Krylov.kcopy!(n::Integer, y::B200Vector, x::B200Vector)" in ext
    assert "Krylov.kaxpby!(n::Integer, s::Float64, x::B200Vector, t::Float64, y::B200Vector)" in ext
    assert "Krylov.kref!(n::Integer, x::B200Vector, y::B200Vector, c::Float64, s::Float64)" in ext


def test_seams_of_the_reference_call_sites_exist():
    src = source()
    # src/Ariadne.jl:259-263 shape: host Array in / out through ak_newton_solve_host
    assert re.search(r"function newton_krylov!\(F!::NativeResidual, u::Array\{Float64\}", src)
    # examples/implicit.jl:17-37
    assert "struct GMidpoint" in src and "struct GTrapezoid" in src and "scheme_code(::GTrapezoid) = AK_TRAPEZOID" in src
    # slabs of a multi-GPU grid
    assert "gny, gy0 = length(p) >= 5 ? (p[4], p[5])" in src and "gny, gy0 = length(pf) >= 6 ? (pf[5], pf[6])" in src
    # krylov_kwargs = (; verbose = 1, reorthogonalization = true) of examples/heat_2D.jl:131 must be accepted
    sig = re.search(r"function krylov_solve!\(ws::Workspace.*?\)\n", src, flags=re.S).group(0)
    for kw in ("verbose", "timemax", "history", "reorthogonalization", "restart", "itmax", "callback", "ldiv", "M", "N"):
        assert re.search(r"\b" + kw + r" = ", sig), kw
    # printing / inspection of a device vector
    assert "Base.getindex(v::B200Vector, i::Int)" in src and "Base.setindex!(v::B200Vector, val, i::Int)" in src
    # operator protocol beyond mul!: transpose, collect, batched
    assert "Base.transpose(J::JacobianOperator)" in src and "function Base.collect(JOp::Union{JacobianOperator, TransposedOperator})" in src
    assert "function mul!(Out::B200Matrix, J::JacobianOperator, V::B200Matrix)" in src
    # the Newton loop keeps every buffer whose raw pointer the library holds rooted; finalizers check `closed`
    assert "GC.@preserve coef rhs workspace J u res p begin" in src
    assert "!v.ctx.closed" in src and "c.closed = true" in src
