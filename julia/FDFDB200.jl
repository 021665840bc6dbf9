# FDFDB200.jl - Julia-side binding of libfdfd_b200.so for MaxwellFDFD.jl users.
#
# UNEXECUTED: there is no Julia in the build image, so this file has never been run.  It shows the
# `ccall` stubs a maintainer would add next to src/model/model.jl so that
#     Ps = create_paramops(mdl); Cs = create_curls(mdl); js = create_srcs(mdl)
#     A, b = create_linsys(EE, ω, Ps, Cs, js);  e = A \ b
# becomes
#     A = create_A_gpu(EE, ω, mdl);  b = create_b_gpu(A, create_srcs(mdl)...);  e = A \ b
# with the operator never assembled.  Every stub cites the C entry point of include/fdfd_b200.h.
module FDFDB200

using MaxwellFDFD
using MaxwellFDFD: create_stretched_∆ls, create_e⁻ⁱᵏᴸ, calc_matparams!
import LinearAlgebra: mul!
import Base: size, *, \

const LIB = get(ENV, "FDFD_B200_LIB", "libfdfd_b200.so")

# mirrors `fdfd_desc` (include/fdfd_b200.h); isbits, passed by reference
struct Desc
    N::NTuple{3,Int64}
    isbloch::NTuple{3,Int32}
    boundft_is_E::NTuple{3,Int32}
    order_cmpfirst::Int32
    field_type::Int32
    device::Int32
    rank::Int32
    nranks::Int32
    weighted_out_avg::Int32
    kernel::Int32
end

mutable struct GpuOperator <: AbstractMatrix{ComplexF64}
    h::Ptr{Cvoid}
    n::Int
    function GpuOperator(h, n)
        A = new(h, n)
        finalizer(a -> ccall((:fdfd_destroy, LIB), Cint, (Ptr{Cvoid},), a.h), A)
        return A
    end
end

function check(code::Cint, h::Ptr{Cvoid}=C_NULL)
    code == 0 && return nothing
    msg = unsafe_string(ccall((:fdfd_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    code == 1 ? throw(ArgumentError(msg)) : error("fdfd_b200 error $code: $msg")   # cf. source.jl:214, model.jl:242
end

"""GPU stand-in for `create_A(ft, ω, create_paramops(mdl), create_curls(mdl))` (model.jl:141-175,225-246)."""
function create_A_gpu(ft::FieldType, ω::Number, mdl::MaxwellFDFD.Model; device::Integer=-1)
    s∆lₑ, s∆lₘ, _, _ = create_stretched_∆ls(mdl)                     # model.jl:122-139
    calc_matparams!(mdl)                                             # model.jl:143 (stays on the host)
    g = mdl.grid
    d = Ref(Desc(Tuple(Int64.(g.N)), Tuple(Int32.(g.isbloch)), Tuple(Int32.(mdl.boundft .== EE)),
                 Int32(mdl.order_cmpfirst), Int32(ft == EE ? 0 : 1), Int32(device), 0, 1, 0, 0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:fdfd_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Ptr{Desc}), h, d))
    A = GpuOperator(h[], 3 * prod(g.N))
    se = [ComplexF64.(v) for v in s∆lₑ]; sm = [ComplexF64.(v) for v in s∆lₘ]
    GC.@preserve se sm begin
        check(ccall((:fdfd_set_coeffs, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{ComplexF64}}, Ptr{Ptr{ComplexF64}}),
                    A.h, pointer.(se), pointer.(sm)), A.h)
    end
    ph = ComplexF64.(create_e⁻ⁱᵏᴸ(mdl))                              # model.jl:91
    check(ccall((:fdfd_set_bloch, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), A.h, ph), A.h)
    check(ccall((:fdfd_set_omega, LIB), Cint, (Ptr{Cvoid}, ComplexF64), A.h, ComplexF64(ω)), A.h)
    ε = Array{ComplexF64,5}(mdl.εarr)                                # (Nx,Ny,Nz,3,3) column-major, model.jl:51
    check(ccall((:fdfd_set_eps, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Cint), A.h, ε, 1), A.h)
    μ = Array{ComplexF64,5}(mdl.μarr)
    check(ccall((:fdfd_set_mu, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), A.h, μ), A.h)
    return A
end

size(A::GpuOperator) = (A.n, A.n)

"""`mul!(y, A, x)`: the per-iteration SparseMatrixCSC product of the reference path -> fdfd_apply."""
function mul!(y::Vector{ComplexF64}, A::GpuOperator, x::Vector{ComplexF64})
    check(ccall((:fdfd_apply, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint), A.h, x, y, 0), A.h)
    return y
end
*(A::GpuOperator, x::Vector{ComplexF64}) = mul!(similar(x), A, x)

"""`A \\ b` -> fdfd_solve (BiCGSTAB by default); returns the field, warns if maxit was hit."""
function solve(A::GpuOperator, b::Vector{ComplexF64}; method::Symbol=:bicgstab, rtol=1e-8, maxit=10_000, x0=zero(b))
    x = copy(x0); iters = Ref{Cint}(0); relres = Ref{Cdouble}(0)
    code = ccall((:fdfd_solve, LIB), Cint,
                 (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint, Cdouble, Cint, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}),
                 A.h, method == :qmr ? 1 : 0, b, x, 0, rtol, maxit, 10, iters, relres, C_NULL)
    code == 5 ? (@warn "not converged" iters[] relres[]) : check(code, A.h)
    return x
end
\(A::GpuOperator, b::Vector{ComplexF64}) = solve(A, b)

"""create_b (model.jl:251-274, EE branch) -> fdfd_create_b."""
function create_b_gpu(A::GpuOperator, jₑ::Vector{ComplexF64}, jₘ::Vector{ComplexF64})
    b = similar(jₑ)
    check(ccall((:fdfd_create_b, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint),
                A.h, jₑ, iszero(jₘ) ? C_NULL : pointer(jₘ), b, 0), A.h)
    return b
end

"""h_from_e (model.jl:276-279) -> fdfd_h_from_e."""
function h_from_e_gpu(A::GpuOperator, e::Vector{ComplexF64}, jₘ::Vector{ComplexF64})
    h = similar(e)
    check(ccall((:fdfd_h_from_e, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint),
                A.h, e, iszero(jₘ) ? C_NULL : pointer(jₘ), h, 0), A.h)
    return h
end

"""e_from_h (model.jl:281-284) -> fdfd_e_from_h (diagonal Pε only, like the reference's `Pε \\`)."""
function e_from_h_gpu(A::GpuOperator, h::Vector{ComplexF64}, jₑ::Vector{ComplexF64})
    e = similar(h)
    check(ccall((:fdfd_e_from_h, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint),
                A.h, h, iszero(jₑ) ? C_NULL : pointer(jₑ), e, 0), A.h)
    return e
end

"""create_Mcs (model.jl:287-306) applied to a field: `Mcₑ * e` (ft = EE) or `Mcₘ * h` (ft = HH) -> fdfd_interp_corners."""
function interp_corners_gpu(A::GpuOperator, ft::FieldType, f::Vector{ComplexF64})
    out = similar(f)
    check(ccall((:fdfd_interp_corners, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint),
                A.h, ft == EE ? 0 : 1, f, out, 0), A.h)
    return out
end

# mirrors `fdfd_shape` / `fdfd_matparams_desc` (include/fdfd_b200.h)
struct CShape
    kind::Int32; axis::Int32; pind::Int32; reserved::Int32
    c::NTuple{3,Float64}
    r::NTuple{3,Float64}
end
struct MatParamsDesc
    N::NTuple{3,Int64}
    isbloch::NTuple{3,Int32}
    boundft_is_E::NTuple{3,Int32}
    field_type::Int32
    field_ortho_shape::Int32
    lprim::NTuple{3,Ptr{Float64}}
    k0::Int64; k1::Int64
    nshape::Int32; nparam::Int32
    shapes::Ptr{CShape}
    params::Ptr{ComplexF64}
    device::Int32
end

"""GPU stand-in for `calc_matparams!(mdl)` (full.jl:16-70): fills `mdl.εarr` (and `mdl.μarr` with `ft = HH`) from
the objects added with `add_obj!`.  `cshapes` / `pinds` / `params` are the model's `oind2shp`, `oind2εind`, `εind2ε`
converted by the caller (Box -> kind 0 with half-widths, Ball -> 1, axis-aligned Cylinder -> 2; row-major 3x3 tensors)."""
function calc_matparams_gpu!(arr::Array{ComplexF64,5}, ft::FieldType, mdl::MaxwellFDFD.Model,
                             cshapes::Vector{CShape}, params::Matrix{ComplexF64}; device::Integer=-1)
    g = mdl.grid
    lprim = ntuple(w -> collect(Float64, g.ghosted.l[1][w]), 3)          # N+1 primal planes incl. the +end ghost point
    GC.@preserve lprim cshapes params begin
        d = MatParamsDesc(Tuple(Int64.(g.N)), Tuple(Int32.(g.isbloch)), Tuple(Int32.(mdl.boundft .== EE)),
                          ft == EE ? 0 : 1, 0, ntuple(w -> pointer(lprim[w]), 3), 0, g.N[3],
                          length(cshapes), size(params, 2), pointer(cshapes), pointer(params), device)
        check(ccall((:fdfd_calc_matparams, LIB), Cint, (Ref{MatParamsDesc}, Ptr{ComplexF64}, Cint), d, arr, 0))
    end
    return arr
end

"""Debug: the assembled A as a SparseMatrixCSC (same colptr/rowval Julia's create_A produces)."""
function sparse_export(A::GpuOperator)
    nnz = Ref{Int64}(0)
    check(ccall((:fdfd_export_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Ptr{Int64}),
                A.h, C_NULL, C_NULL, C_NULL, nnz), A.h)
    colptr = Vector{Int64}(undef, A.n + 1); rowval = Vector{Int64}(undef, nnz[]); nzval = Vector{ComplexF64}(undef, nnz[])
    check(ccall((:fdfd_export_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Ptr{Int64}),
                A.h, colptr, rowval, nzval, nnz), A.h)
    return SparseArrays.SparseMatrixCSC(A.n, A.n, colptr, rowval, nzval)
end

end # module
