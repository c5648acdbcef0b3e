# Package extension (loaded automatically when Krylov.jl is present): attaches the vector-kernel protocol of Krylov.jl
# to B200Vector exactly where examples/halovector.jl:51-147 attaches it to HaloVector, so that Krylov.jl itself
# (gmres!, cg!, fgmres!, any other solver) can drive device-resident vectors through libariadne_b200.so:
#
#     using Krylov, AriadneB200
#     ws = krylov_workspace(:gmres, KrylovConstructor(res::B200Vector))      # src/Ariadne.jl:317-318
#     krylov_solve!(ws, J::AriadneB200.JacobianOperator, copy(res); rtol = η) # src/Ariadne.jl:338
#
# (The fused native solver `AriadneB200.krylov_solve!` does not need these methods.)
# NOTE: never executed here (no `julia` binary in the build image); signatures follow examples/halovector.jl.
module AriadneB200KrylovExt

using AriadneB200
using Krylov
import AriadneB200: B200Vector, Krylov_kdot, Krylov_knorm, Krylov_kscal!, Krylov_kaxpy!, Krylov_kaxpby!, Krylov_kcopy!,
                    Krylov_kfill!, Krylov_kref!, Krylov_kdivcopy!

Krylov.kdot(n::Integer, x::B200Vector, y::B200Vector) = Krylov_kdot(n, x, y)                              # halovector.jl:51-62
Krylov.knorm(n::Integer, x::B200Vector) = Krylov_knorm(n, x)                                              # :64-74
Krylov.kscal!(n::Integer, s::Float64, x::B200Vector) = Krylov_kscal!(n, s, x)                             # :76-85
Krylov.kaxpy!(n::Integer, s::Float64, x::B200Vector, y::B200Vector) = Krylov_kaxpy!(n, s, x, y)           # :87-97
Krylov.kaxpby!(n::Integer, s::Float64, x::B200Vector, t::Float64, y::B200Vector) = Krylov_kaxpby!(n, s, x, t, y)  # :99-109
Krylov.kcopy!(n::Integer, y::B200Vector, x::B200Vector) = Krylov_kcopy!(n, y, x)                          # :111-121
Krylov.kfill!(x::B200Vector, val::Float64) = Krylov_kfill!(x, val)                                        # :123-132
Krylov.kref!(n::Integer, x::B200Vector, y::B200Vector, c::Float64, s::Float64) = Krylov_kref!(n, x, y, c, s)  # :134-147
# Krylov 0.10 also routes V[k+1] = q / Hbis through kdivcopy! when the vector type provides it; HaloVector falls back to
# broadcasting over getindex/setindex!, which would be one PCIe round trip per entry here.
if isdefined(Krylov, :kdivcopy!)
    Krylov.kdivcopy!(n::Integer, y::B200Vector, x::B200Vector, s::Float64) = Krylov_kdivcopy!(n, y, x, s)
end
if isdefined(Krylov, :kscalcopy!)
    Krylov.kscalcopy!(n::Integer, y::B200Vector, s::Float64, x::B200Vector) = (Krylov_kcopy!(n, y, x); Krylov_kscal!(n, s, y))
end
if isdefined(Krylov, :kdotr)
    Krylov.kdotr(n::Integer, x::B200Vector, y::B200Vector) = Krylov_kdot(n, x, y)
end

end # module
