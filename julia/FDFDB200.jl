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
using SparseArrays
import LinearAlgebra: mul!
import Base: size, eltype, show, *, \

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

# What create_A returns here.  Deliberately NOT an AbstractMatrix: the operator is matrix-free, there is no getindex, and
# the generic AbstractMatrix fallbacks (show, iteration, dense conversion) would call it element by element.  It supports
# what the reference path does with A: size, eltype, mul!, *, \ (and sparse_export for debugging).
# multi = true: a fdfd_multi handle - ONE value over `ngpu` GPUs of this box (include/fdfd_b200.h, fdfd_multi_*).
mutable struct GpuOperator
    h::Ptr{Cvoid}
    n::Int
    multi::Bool
    function GpuOperator(h, n, multi=false)
        A = new(h, n, multi)
        finalizer(a -> a.multi ? ccall((:fdfd_multi_destroy, LIB), Cint, (Ptr{Cvoid},), a.h) :
                                 ccall((:fdfd_destroy, LIB), Cint, (Ptr{Cvoid},), a.h), A)
        return A
    end
end

function check(code::Cint, h::Ptr{Cvoid}=C_NULL; multi::Bool=false)
    code == 0 && return nothing
    msg = multi ? unsafe_string(ccall((:fdfd_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), h)) :
                  unsafe_string(ccall((:fdfd_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    code == 1 ? throw(ArgumentError(msg)) : error("fdfd_b200 error $code: $msg")   # cf. source.jl:214, model.jl:242
end
check(A::GpuOperator, code::Cint) = check(code, A.h; multi=A.multi)

"""GPU stand-in for `create_A(ft, ω, create_paramops(mdl), create_curls(mdl))` (model.jl:141-175,225-246).
`ngpu > 1` (optionally `devices = [0, 1, ...]`): the same single value, spread over that many GPUs as z-slabs - the
library owns one host thread and one slab per device; `A * x` and `A \\ b` take the full-grid vectors as before."""
function create_A_gpu(ft::FieldType, ω::Number, mdl::MaxwellFDFD.Model; device::Integer=-1, ngpu::Integer=1,
                      devices::Union{Nothing,Vector{<:Integer}}=nothing)
    s∆lₑ, s∆lₘ, _, _ = create_stretched_∆ls(mdl)                     # model.jl:122-139
    calc_matparams!(mdl)                                             # model.jl:143 (stays on the host)
    g = mdl.grid
    d = Ref(Desc(Tuple(Int64.(g.N)), Tuple(Int32.(g.isbloch)), Tuple(Int32.(mdl.boundft .== EE)),
                 Int32(mdl.order_cmpfirst), Int32(ft == EE ? 0 : 1), Int32(device), 0, 1, 0, 0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    multi = ngpu > 1 || devices !== nothing
    if multi
        devs = devices === nothing ? C_NULL : Int32.(devices)
        check(ccall((:fdfd_multi_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Ptr{Desc}, Int32, Ptr{Int32}), h, d, Int32(ngpu), devs);
              multi=true)
    else
        check(ccall((:fdfd_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Ptr{Desc}), h, d))
    end
    A = GpuOperator(h[], 3 * prod(g.N), multi)
    se = [ComplexF64.(v) for v in s∆lₑ]; sm = [ComplexF64.(v) for v in s∆lₘ]
    GC.@preserve se sm begin
        pe, pm = pointer.(se), pointer.(sm)
        check(A, multi ? ccall((:fdfd_multi_set_coeffs, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{ComplexF64}}, Ptr{Ptr{ComplexF64}}), A.h, pe, pm) :
                         ccall((:fdfd_set_coeffs, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{ComplexF64}}, Ptr{Ptr{ComplexF64}}), A.h, pe, pm))
    end
    ph = ComplexF64.(create_e⁻ⁱᵏᴸ(mdl))                              # model.jl:91
    check(A, multi ? ccall((:fdfd_multi_set_bloch, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), A.h, ph) :
                     ccall((:fdfd_set_bloch, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), A.h, ph))
    check(A, multi ? ccall((:fdfd_multi_set_omega, LIB), Cint, (Ptr{Cvoid}, ComplexF64), A.h, ComplexF64(ω)) :
                     ccall((:fdfd_set_omega, LIB), Cint, (Ptr{Cvoid}, ComplexF64), A.h, ComplexF64(ω)))
    ε = Array{ComplexF64,5}(mdl.εarr)                                # (Nx,Ny,Nz,3,3) column-major, model.jl:51
    check(A, multi ? ccall((:fdfd_multi_set_eps, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Cint), A.h, ε, 1) :
                     ccall((:fdfd_set_eps, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Cint), A.h, ε, 1))
    μ = Array{ComplexF64,5}(mdl.μarr)
    check(A, multi ? ccall((:fdfd_multi_set_mu, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), A.h, μ) :
                     ccall((:fdfd_set_mu, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), A.h, μ))
    return A
end

size(A::GpuOperator) = (A.n, A.n)
size(A::GpuOperator, i::Integer) = i <= 2 ? A.n : 1
eltype(::GpuOperator) = ComplexF64
show(io::IO, A::GpuOperator) = print(io, A.n, "×", A.n, " matrix-free FDFD operator on ", A.multi ? "several GPUs" : "one GPU")

"""`mul!(y, A, x)`: the per-iteration SparseMatrixCSC product of the reference path -> fdfd_apply / fdfd_multi_apply."""
function mul!(y::Vector{ComplexF64}, A::GpuOperator, x::Vector{ComplexF64})
    check(A, A.multi ? ccall((:fdfd_multi_apply, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}), A.h, x, y) :
                       ccall((:fdfd_apply, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint), A.h, x, y, 0))
    return y
end
*(A::GpuOperator, x::Vector{ComplexF64}) = mul!(similar(x), A, x)

"""`A \\ b` -> fdfd_solve (BiCGSTAB by default); returns the field, warns if maxit was hit."""
function solve(A::GpuOperator, b::Vector{ComplexF64}; method::Symbol=:bicgstab, rtol=1e-8, maxit=10_000, x0=zero(b))
    x = copy(x0); iters = Ref{Cint}(0); relres = Ref{Cdouble}(0)
    m = method == :qmr ? 1 : 0
    code = A.multi ?
        ccall((:fdfd_multi_solve, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Cdouble, Cint, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}),
              A.h, m, b, x, rtol, maxit, 10, iters, relres, C_NULL) :
        ccall((:fdfd_solve, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint, Cdouble, Cint, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}),
              A.h, m, b, x, 0, rtol, maxit, 10, iters, relres, C_NULL)
    code == 5 ? (@warn "not converged" iters[] relres[]) : check(A, code)
    return x
end
\(A::GpuOperator, b::Vector{ComplexF64}) = solve(A, b)

"""create_b (model.jl:251-274, EE branch) -> fdfd_create_b."""
function create_b_gpu(A::GpuOperator, jₑ::Vector{ComplexF64}, jₘ::Vector{ComplexF64})
    b = similar(jₑ)
    pm = iszero(jₘ) ? Ptr{ComplexF64}(C_NULL) : pointer(jₘ)
    GC.@preserve jₘ check(A, A.multi ?
        ccall((:fdfd_multi_create_b, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}), A.h, jₑ, pm, b) :
        ccall((:fdfd_create_b, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint), A.h, jₑ, pm, b, 0))
    return b
end

"""h_from_e (model.jl:276-279) -> fdfd_h_from_e."""
function h_from_e_gpu(A::GpuOperator, e::Vector{ComplexF64}, jₘ::Vector{ComplexF64})
    h = similar(e)
    pm = iszero(jₘ) ? Ptr{ComplexF64}(C_NULL) : pointer(jₘ)
    GC.@preserve jₘ check(A, A.multi ?
        ccall((:fdfd_multi_h_from_e, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}), A.h, e, pm, h) :
        ccall((:fdfd_h_from_e, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint), A.h, e, pm, h, 0))
    return h
end

"""e_from_h (model.jl:281-284) -> fdfd_e_from_h (diagonal Pε only, like the reference's `Pε \\`)."""
function e_from_h_gpu(A::GpuOperator, h::Vector{ComplexF64}, jₑ::Vector{ComplexF64})
    e = similar(h)
    pe = iszero(jₑ) ? Ptr{ComplexF64}(C_NULL) : pointer(jₑ)
    GC.@preserve jₑ check(A, A.multi ?
        ccall((:fdfd_multi_e_from_h, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}), A.h, h, pe, e) :
        ccall((:fdfd_e_from_h, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint), A.h, h, pe, e, 0))
    return e
end

"""create_Mcs (model.jl:287-306) applied to a field: `Mcₑ * e` (ft = EE) or `Mcₘ * h` (ft = HH) -> fdfd_interp_corners."""
function interp_corners_gpu(A::GpuOperator, ft::FieldType, f::Vector{ComplexF64})
    A.multi && throw(ArgumentError("interp_corners_gpu: single-GPU operators only"))
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
    A.multi && throw(ArgumentError("sparse_export: single-GPU operators only"))
    nnz = Ref{Int64}(0)
    check(ccall((:fdfd_export_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Ptr{Int64}),
                A.h, C_NULL, C_NULL, C_NULL, nnz), A.h)
    colptr = Vector{Int64}(undef, A.n + 1); rowval = Vector{Int64}(undef, nnz[]); nzval = Vector{ComplexF64}(undef, nnz[])
    check(ccall((:fdfd_export_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Ptr{Int64}),
                A.h, colptr, rowval, nzval, nnz), A.h)
    return SparseMatrixCSC(A.n, A.n, colptr, rowval, nzval)
end

end # module
