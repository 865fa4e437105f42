# BlueTangleCUDA.jl -- Julia shim over libbluetangle_cuda.so (include/bluetangle_cuda.h).
#
# NOT EXECUTED in the build container (no Julia there): this file is the reference-side binding a maintainer adds.
# It introduces two device-resident state types and extends the reference's own generic functions for them, branch for
# branch with src/hilbert.jl:469-515 (apply), :639-666 (apply on rho), :322-364 (apply_noise), :669-696 (measurement),
# src/func.jl:91-147 (expect / correlation), src/ops.jl:46-132 (sample / sample_exact).  Every method is a thin ccall.
module BlueTangleCUDA

using BlueTangle
import BlueTangle: apply, apply_noise, get_N, sample, sample_exact, expect, correlation, partial_trace,
                   born_measure_Z, _born_measure, _reset_Z, zero_state, inner, fidelity, measure
using Random

const LIB = get(ENV, "BLUETANGLE_CUDA_LIB", "libbluetangle_cuda")

struct BTError <: Exception
    code::Cint
    msg::String
end
lasterr() = unsafe_string(ccall((:bt_last_error, LIB), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : throw(BTError(rc, lasterr()))

# ---- device-resident states -------------------------------------------------------------------------------------
mutable struct CuState
    h::Ptr{Cvoid}
    N::Int
    n_batch::Int
    function CuState(N::Int, n_batch::Int=1)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bt_sv_create, LIB), Cint, (Cint, Int64, Ref{Ptr{Cvoid}}), N, n_batch, r))
        s = new(r[], N, n_batch)
        finalizer(x -> ccall((:bt_sv_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), s)
        return s
    end
end

mutable struct CuRho
    h::Ptr{Cvoid}
    N::Int
    function CuRho(N::Int)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:bt_dm_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), N, r))
        d = new(r[], N)
        finalizer(x -> ccall((:bt_dm_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), d)
        return d
    end
end

get_N(s::CuState) = s.N
get_N(r::CuRho) = r.N
cu_zero_state(N::Int; n_batch::Int=1) = CuState(N, n_batch)

function CuState(v::AbstractVector{<:Complex})
    N = Int(log2(length(v)))
    s = CuState(N)
    a = Vector{ComplexF64}(v)
    check(ccall((:bt_sv_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, UInt64), s.h, a, length(a)))
    return s
end
function Base.Vector(s::CuState)
    out = Vector{ComplexF64}(undef, s.n_batch << s.N)
    check(ccall((:bt_sv_download, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, UInt64), s.h, out, length(out)))
    return out
end
function Base.Matrix(r::CuRho)
    out = Matrix{ComplexF64}(undef, 2^r.N, 2^r.N)
    check(ccall((:bt_dm_download, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, UInt64), r.h, out, length(out)))
    return out
end

cmat(m) = Matrix{ComplexF64}(m)   # Julia matrices are column-major: exactly what the ABI expects

# ---- gates: op.expand(N)*state (src/hilbert.jl:505) ------------------------------------------------------------------
# `mat`: the op's matrix, or for variational ops (op.type == "f", src/hilbert.jl:500-501) the matrix function applied to `pars`
function _apply_matrix!(s::CuState, op, mat=op.mat)
    if op.q == 1
        check(ccall((:bt_sv_apply_1q, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Cint), s.h, op.qubit, cmat(mat), op.control))
    else
        check(ccall((:bt_sv_apply_2q, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{ComplexF64}, Cint), s.h, op.qubit, op.target_qubit, cmat(mat), op.control))
    end
end
function _apply_matrix!(r::CuRho, op, mat=op.mat)
    if op.q == 1
        check(ccall((:bt_dm_apply_1q, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Cint), r.h, op.qubit, cmat(mat), op.control))
    else
        check(ccall((:bt_dm_apply_2q, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{ComplexF64}, Cint), r.h, op.qubit, op.target_qubit, cmat(mat), op.control))
    end
end

_kraus_table(kraus) = reduce(vcat, [vec(cmat(k)) for k in kraus])

# __QuantumChannel_new_apply (src/struct.jl:31-76): rand() is drawn HERE so the RNG stream equals the CPU path's
function _channel!(s::CuState, kraus, q::Int, qubit::Int, target::Int)
    u = [rand() for _ in 1:s.n_batch]
    chosen = Vector{Int32}(undef, s.n_batch)
    check(ccall((:bt_sv_kraus, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{ComplexF64}, Cint, Ptr{Float64}, Ptr{Int32}),
                s.h, q, qubit, target, _kraus_table(kraus), length(kraus), u, chosen))
    return s
end
function _channel!(r::CuRho, kraus, q::Int, qubit::Int, target::Int)
    check(ccall((:bt_dm_kraus, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{ComplexF64}, Cint), r.h, q, qubit, target, _kraus_table(kraus), length(kraus)))
    return r
end

# born_measure_Z (src/hilbert.jl:682-696) / _reset_Z (:752-759)
function born_measure_Z(N::Int, s::CuState, qubit::Int; reset::Bool=false)
    u = [rand() for _ in 1:s.n_batch]
    out = Vector{Int32}(undef, s.n_batch); p0 = Vector{Float64}(undef, s.n_batch)
    check(ccall((:bt_sv_measure_z, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Cint), s.h, qubit, u, out, p0, reset ? 1 : 0))
    return s, (s.n_batch == 1 ? Int(out[1]) : Int.(out))
end
_reset_Z(s::CuState, qubit::Int) = born_measure_Z(s.N, s, qubit; reset=true)
function born_measure_Z(N::Int, r::CuRho, qubit::Int)      # src/hilbert.jl:784-796
    check(ccall((:bt_dm_dephase, LIB), Cint, (Ptr{Cvoid}, Cint), r.h, qubit)); return r
end

function _born_measure(s::CuState, o::QuantumOps)             # src/hilbert.jl:669-679
    rname = BlueTangle._resolve_measurement_name(o.name)       # "MR": one discrete draw (src/struct.jl:565-571)
    rot = BlueTangle._measurement_mat(rname)
    rot == BlueTangle.gate.I || check(ccall((:bt_sv_apply_1q, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Cint), s.h, o.qubit, cmat(rot), -2))
    s, ind = born_measure_Z(s.N, s, o.qubit)
    rot == BlueTangle.gate.I || check(ccall((:bt_sv_apply_1q, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Cint), s.h, o.qubit, cmat(rot'), -2))
    return s, ind
end
_born_measure(r::CuRho, o::QuantumOps) = throw("fix this:")   # src/hilbert.jl:772: unsupported in the reference

# apply_noise (src/hilbert.jl:322-364)
function apply_noise(x::Union{CuState,CuRho}, op::QuantumOps, noise::NoiseModel)
    (hasproperty(op, :noisy) && op.noisy == true) || return x
    if op.q == 1
        op.control == -2 ? _channel!(x, noise.q1.kraus, 1, op.qubit, -1) : _channel!(x, noise.q2.kraus, 2, op.control, op.qubit)
    elseif op.q == 2
        _channel!(x, noise.q2.kraus, 2, op.qubit, op.target_qubit)
    end
    return x
end

# apply (src/hilbert.jl:469-515 and :639-666): same branch order
function apply(x::Union{CuState,CuRho}, op::QuantumOps; noise::Union{NoiseModel,Bool}=false, track_measurements::Bool=false, kwargs...)
    mid = Int[]
    if isa(op, OpF)
        # src/struct.jl:703-744: op.apply is a closure typed for the CPU states; op.data holds what it was built from
        if op.data isa Function
            r = op.data(x)
            r isa Union{CuState,CuRho} && (x = r)
        elseif op.data isa AbstractVector
            for o in op.data; apply(x, o); end
        else
            throw(ArgumentError("OpF with a full-register matrix cannot run on a device-resident state"))
        end
    elseif isa(op, OpQC)
        if x isa CuState && uppercase(op.name) in ("RES", "RESET")
            _reset_Z(x, op.qubit)
        else
            _channel!(x, op.kraus, op.q, op.qubit, op.target_qubit)
        end
    elseif op.type == "🔬"
        if isa(op, ifOp)
            x isa CuRho && throw("error: fix this!")          # src/struct.jl:616
            _, ind = _born_measure(x, op)
            for ifop in (ind == 0 ? op.if01[1] : op.if01[2])
                apply(x, ifop; noise=noise)
            end
        else
            _, ind = _born_measure(x, op)
        end
        track_measurements && push!(mid, ind)
    elseif op.type == "f"                                      # variational gates: matrix from the parameters (src/hilbert.jl:500-501)
        _apply_matrix!(x, op, op.mat(kwargs[:pars]...))
    else
        _apply_matrix!(x, op)
    end
    isa(noise, NoiseModel) && apply_noise(x, op, noise)
    return track_measurements ? (x, mid) : x
end
Base.:*(op::QuantumOps, x::Union{CuState,CuRho}) = apply(x, op)                     # src/all.jl:41-49

struct BtGate
    nq::Int32; qubit::Int32; target::Int32; control::Int32
    m::NTuple{16,ComplexF64}
end
function _pack(ops::Vector{<:QuantumOps})
    map(ops) do o
        v = zeros(ComplexF64, 16); v[1:length(o.mat)] = vec(cmat(o.mat))
        BtGate(o.q, o.qubit, o.target_qubit, o.control, Tuple(v))
    end
end
# apply(ops, state) (src/hilbert.jl:517-553): a run of plain gates goes down in ONE call (host fusion + tile kernel)
function apply(ops::Vector{<:QuantumOps}, s::CuState; noise::Union{NoiseModel,Bool}=false, track_measurements::Bool=false, kwargs...)
    mids = Int[]
    plain(o) = isa(o, Op) && o.type != "🔬" && o.type != "f" && !isa(noise, NoiseModel)
    i = 1
    while i <= length(ops)
        if plain(ops[i])
            j = i
            while j < length(ops) && plain(ops[j+1]); j += 1; end
            g = _pack(ops[i:j])
            check(ccall((:bt_sv_apply_circuit, LIB), Cint, (Ptr{Cvoid}, Ptr{BtGate}, UInt64, Cint), s.h, g, length(g), 1))
            i = j + 1
        else
            r = apply(s, ops[i]; noise=noise, track_measurements=track_measurements, kwargs...)   # kwargs forwarded like src/hilbert.jl:531-536 (pars of variational ops)
            track_measurements && append!(mids, r[2])
            i += 1
        end
    end
    return track_measurements ? (s, mids) : s
end

# ---- reductions / observables ------------------------------------------------------------------------------------------
function partial_trace(s::CuState, q::Int)                                           # src/linalg.jl:167-192
    out = Matrix{ComplexF64}(undef, 2, 2)
    check(ccall((:bt_sv_rdm1, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}), s.h, q, out)); out
end
function partial_trace(s::CuState, q1::Int, q2::Int)                                 # src/linalg.jl:198-230
    abs(q1 - q2) > 1 && throw("must be local")
    out = Matrix{ComplexF64}(undef, 4, 4)
    check(ccall((:bt_sv_rdm2, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{ComplexF64}), s.h, q1, q2, out)); out
end
function partial_trace(s::CuState, keep::AbstractVector)                             # src/linalg.jl:83-86 (any number of kept qubits up to 12)
    k = length(keep); d = 2^k
    out = Matrix{ComplexF64}(undef, d, d)
    check(ccall((:bt_sv_rdm, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{ComplexF64}), s.h, k, Cint.(keep), out)); out
end
function partial_trace(r::CuRho, dims::Vector, trace_out::Vector)                    # src/linalg.jl:88-140 (qubit registers: all dims == 2)
    all(==(2), dims) || throw(ArgumentError("device density matrices are qubit registers"))
    keep = setdiff(1:length(dims), trace_out); k = length(keep); d = 2^k
    out = Matrix{ComplexF64}(undef, d, d)
    check(ccall((:bt_dm_rdm, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{ComplexF64}), r.h, k, Cint.(keep), out)); out
end
bipartition_trace(r::CuRho) = partial_trace(r, fill(2, r.N), collect(1:Int(r.N / 2)))  # src/linalg.jl:151-161: keeps the last N/2 qubits

# entanglement_entropy src/func.jl:299-312 / :323-328: the spectrum comes from the device (one-sided Jacobi iteration)
function entanglement_entropy(s::CuState; spectrum_bool=false)
    partA = s.N ÷ 2
    spec = Vector{Float64}(undef, 2^partA * s.n_batch)
    check(ccall((:bt_sv_schmidt_spectrum, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Cint}), s.h, partA, spec, C_NULL))
    spec = spec[spec .> 0]
    return spectrum_bool == false ? sum(-spec .* log.(spec)) : (sum(-spec .* log.(spec)), -log.(spec))
end
function entanglement_entropy(r::CuRho)
    nk = Int(r.N / 2)
    spec = Vector{Float64}(undef, 2^nk)
    check(ccall((:bt_dm_bipartition_spectrum, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Cint}), r.h, nk, spec, C_NULL))
    spec = spec[spec .> 0]
    return sum(-spec .* log.(spec)), -log.(spec)
end

function fidelity(r::CuRho, s::CuRho)                                                # src/tensor.jl:222-229
    out = Ref{Float64}(0.0)
    check(ccall((:bt_dm_fidelity, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Float64}), r.h, s.h, out)); out[]
end

# expect(x, op) src/func.jl:91-92 for any Op (1- or 2-qubit, with or without a control)
function expect(s::CuState, op::Op)
    out = Vector{Float64}(undef, s.n_batch)
    check(ccall((:bt_sv_expect_op, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{ComplexF64}, Ptr{Float64}), s.h, op.q, op.qubit, op.target_qubit, op.control, cmat(op.mat), out))
    return s.n_batch == 1 ? out[1] : out
end
function expect(r::CuRho, op::Op)
    out = Ref{Float64}(0.0)
    check(ccall((:bt_dm_expect_op, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{ComplexF64}, Ref{Float64}), r.h, op.q, op.qubit, op.target_qubit, op.control, cmat(op.mat), out))
    return out[]
end

function expect(s::CuState, op_str::String)                                          # src/func.jl:97
    out = Vector{Float64}(undef, s.N * s.n_batch)
    check(ccall((:bt_sv_expect_1q_all, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}), s.h, cmat(gates(op_str)), out)); out
end
function expect(r::CuRho, op_str::String)                                            # src/func.jl:98
    out = Vector{Float64}(undef, r.N)
    check(ccall((:bt_dm_expect_1q_all, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}), r.h, cmat(gates(op_str)), out)); out
end
function correlation(s::CuState, list_of_operators::String, qubits_applied::Vector)  # src/func.jl:139-142
    names = String.(split(list_of_operators, ","))
    mats = reduce(vcat, [vec(cmat(gates(n))) for n in names])
    out = Vector{Float64}(undef, s.n_batch)
    check(ccall((:bt_sv_expect_product, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{ComplexF64}, Ptr{Float64}), s.h, length(names), Cint.(qubits_applied), mats, out))
    return s.n_batch == 1 ? out[1] : out
end
function correlation(r::CuRho, list_of_operators::String, qubits_applied::Vector)    # src/func.jl:144-147
    names = String.(split(list_of_operators, ","))
    mats = reduce(vcat, [vec(cmat(gates(n))) for n in names])
    out = Ref{Float64}(0.0)
    check(ccall((:bt_dm_expect_product, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{ComplexF64}, Ref{Float64}), r.h, length(names), Cint.(qubits_applied), mats, out))
    return out[]
end

# sample (src/ops.jl:46-62): draws come from Julia's RNG; inverse CDF on the device (SURVEY App. A.6)
function sample(s::CuState, shots)
    u = [rand() for _ in 1:shots]
    out = Vector{Int64}(undef, shots)
    check(ccall((:bt_sv_sample, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, UInt64, Ptr{Int64}), s.h, u, shots, out)); out
end
function sample_exact(s::CuState)                                                    # src/ops.jl:98-101
    p = Vector{Float64}(undef, 2^s.N)
    check(ccall((:bt_sv_probs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.h, p))
    idx = findall(!iszero, p)
    return idx .- 1, p[idx]
end
function inner(a::CuState, b::CuState)
    out = Vector{ComplexF64}(undef, a.n_batch)
    check(ccall((:bt_sv_inner, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{ComplexF64}), a.h, b.h, out)); a.n_batch == 1 ? out[1] : out
end
fidelity(a::CuState, b::CuState) = abs2.(inner(a, b))

# Pauli-sum Hamiltonians (hamiltonian src/vqa.jl:36-67 + expect src/func.jl:91, the VQE loss src/vqa.jl:282-283): the term list
# stays a list; one device call evaluates sum_k c_k <P_k> with one read of the state per commuting group of terms.
struct PauliSum
    N::Int
    coefs::Vector{Float64}
    strings::String          # N characters (I/X/Y/Z, qubit 1 first) per term, concatenated
end
function pauli_sum(N::Int, string_of_ops::Vector, boundary::String="open")           # same arguments as hamiltonian(N, ...)
    coefs = Float64[]; io = IOBuffer()
    for id in 2:2:length(string_of_ops)
        names = String.(split(string_of_ops[id], ",")); k = length(names)
        sites = boundary == "open" ? (1:N-(k-1)) : (1:N)
        for site in sites
            s = fill('I', N)
            for (j, n) in enumerate(names); s[mod1(site + j - 1, N)] = uppercase(n)[1]; end
            write(io, String(s)); push!(coefs, Float64(string_of_ops[id-1]))
        end
    end
    PauliSum(N, coefs, String(take!(io)))
end
function expect(s::CuState, H::PauliSum)
    out = Vector{Float64}(undef, s.n_batch)
    check(ccall((:bt_sv_expect_pauli_sum, LIB), Cint, (Ptr{Cvoid}, Cint, Cstring, Ptr{Float64}, Ptr{Float64}), s.h, length(H.coefs), H.strings, H.coefs, out))
    return s.n_batch == 1 ? out[1] : out
end

end # module
