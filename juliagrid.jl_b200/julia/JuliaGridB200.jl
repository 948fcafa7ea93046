# JuliaGridB200.jl — overlay package: the `B200` tag that routes JuliaGrid's Newton-Raphson power flow and
# Gauss-Newton WLS state estimation to libjgb200.so (hand-written sm_100a CUDA kernels) through `ccall`.
#
# Host code only. It reads the reference's own structures (`PowerSystem`, `Measurement`, the tables `acWLS` builds)
# and passes their arrays — Float64 / Int64 (1-based) / Int8 / ComplexF64 — straight to the C ABI of include/jgb200.h.
# NOTE: Julia is not installed in the build image or on the GPU box, so this file has NEVER BEEN EXECUTED: it is a
# reviewed-by-eye sketch of the binding a maintainer adds (INTEGRATION.md), not a tested artefact. The same ABI, call
# for call, is exercised by the Python host mirror (juliagrid.jl_b200/*.py) and by tests/abi_c_test.c in the test-suite.
module JuliaGridB200

using JuliaGrid
using SparseArrays
import JuliaGrid: newtonRaphson, gaussNewton, mismatch!, solve!, increment!, powerFlow!, stateEstimation!,
    setInitialPoint!, power!, current!, AC, Polar, PowerSystem, Measurement, Normal

const libjgb = get(ENV, "JGB200_LIB", joinpath(@__DIR__, "..", "libjgb200.so"))

"Factorisation tag: `newtonRaphson(system, B200)`, `gaussNewton(monitoring, B200)`."
struct B200 <: Normal end

mutable struct Ctx
    handle::Ptr{Cvoid}
    function Ctx(device::Integer = 0, stream::Ptr{Cvoid} = C_NULL)
        rc = Ref{Int32}(0)
        h = ccall((:jgb_create, libjgb), Ptr{Cvoid}, (Int32, Ptr{Cvoid}, Ref{Int32}), device, stream, rc)
        h == C_NULL && throw(ErrorException(unsafe_string(ccall((:jgb_last_error, libjgb), Cstring, (Ptr{Cvoid},), C_NULL))))
        ctx = new(h)
        finalizer(c -> ccall((:jgb_destroy, libjgb), Cvoid, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end

function check(ctx::Ctx, rc::Int32)
    rc < 0 && throw(ErrorException(unsafe_string(ccall((:jgb_last_error, libjgb), Cstring, (Ptr{Cvoid},), ctx.handle))))
    return rc      # rc == 1: iteration cap reached — not an error, as in the reference (print/solver.jl:426-442)
end

##### Newton-Raphson #####
mutable struct NewtonRaphsonB200
    jacobianColptr::Vector{Int64}
    jacobianRowval::Vector{Int64}
    mismatch::Vector{Float64}
    increment::Vector{Float64}
    pq::Vector{Int64}
    pvpq::Vector{Int64}
    pcount::Vector{Int64}
    iteration::Int64
    ctx::Ctx
end

mutable struct AcPowerFlowB200 <: AC
    voltage::Polar                 # the SAME objects as reference.voltage / power / current
    power::JuliaGrid.AcPower
    current::JuliaGrid.AcCurrent
    method::NewtonRaphsonB200
    system::PowerSystem
    reference::Any                 # the reference's own AcPowerFlow{NewtonRaphson{LU}}: start point, power!, current!
    deviceMagnitude::Vector{Float64}   # what the device holds; analysis.voltage is pushed when it differs
    deviceAngle::Vector{Float64}
end

function newtonRaphson(system::PowerSystem, ::Type{B200}; device::Integer = 0)
    ref = newtonRaphson(system, LU)                   # reference constructor: bus-type fix-ups, start point, acModel!
    ac = system.model.ac
    bus = system.bus
    ctx = Ctx(device)
    Y, Yt = ac.nodalMatrix, ac.nodalMatrixTranspose
    GC.@preserve Y Yt begin
        check(ctx, ccall((:jgb_nr_setup, libjgb), Int32,
            (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int8}, Int64),
            ctx.handle, bus.number, Y.colptr, Y.rowval, pointer(reinterpret(Float64, Y.nzval)),
            pointer(reinterpret(Float64, Yt.nzval)), bus.layout.type, bus.layout.slack))
    end
    dim, nnzJ = Ref{Int64}(0), Ref{Int64}(0)
    check(ctx, ccall((:jgb_nr_dims, libjgb), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), ctx.handle, dim, nnzJ))
    m = NewtonRaphsonB200(zeros(Int64, dim[] + 1), zeros(Int64, nnzJ[]), zeros(dim[]), zeros(dim[]),
        zeros(Int64, bus.number), zeros(Int64, bus.number), zeros(Int64, bus.number), 0, ctx)
    check(ctx, ccall((:jgb_nr_pattern, libjgb), Int32,
        (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
        ctx.handle, m.pq, m.pvpq, m.pcount, m.jacobianColptr, m.jacobianRowval))
    check(ctx, ccall((:jgb_nr_set_injection, libjgb), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx.handle, bus.supply.active, bus.supply.reactive, bus.demand.active, bus.demand.reactive))
    analysis = AcPowerFlowB200(ref.voltage, ref.power, ref.current, m, system, ref,
        similar(ref.voltage.magnitude), similar(ref.voltage.angle))
    pushState!(analysis)
    return analysis
end

function pushState!(a::AcPowerFlowB200)
    check(a.method.ctx, ccall((:jgb_nr_set_state, libjgb), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), a.method.ctx.handle, a.voltage.magnitude, a.voltage.angle))
    copyto!(a.deviceMagnitude, a.voltage.magnitude)
    copyto!(a.deviceAngle, a.voltage.angle)
    return nothing
end
function pullState!(a::AcPowerFlowB200)
    check(a.method.ctx, ccall((:jgb_nr_get_state, libjgb), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), a.method.ctx.handle, a.voltage.magnitude, a.voltage.angle))
    copyto!(a.deviceMagnitude, a.voltage.magnitude)
    copyto!(a.deviceAngle, a.voltage.angle)
    return nothing
end
"analysis.voltage is plain Julia data the user may edit or reset between calls: push it whenever it left the device copy"
syncState!(a::AcPowerFlowB200) =
    (a.voltage.magnitude == a.deviceMagnitude && a.voltage.angle == a.deviceAngle) || pushState!(a)

"setInitialPoint!(analysis) (acPowerFlow.jl:1226-1249): the reference resets the shared voltage vectors; push them"
function setInitialPoint!(a::AcPowerFlowB200)
    setInitialPoint!(a.reference)
    a.method.iteration = 0
    pushState!(a)
    return nothing
end

"power! (postprocessing/acAnalysis.jl:30-79) is a method of the reference's AcPowerFlow; the B200 analysis shares its
voltage / power / current containers with that object, so the call is forwarded. current! (:672-700) is generic on `AC`
and works on the B200 analysis as it is."
power!(a::AcPowerFlowB200) = power!(a.reference)

function mismatch!(a::AcPowerFlowB200)
    syncState!(a)
    sp, sq = Ref{Float64}(0), Ref{Float64}(0)
    check(a.method.ctx, ccall((:jgb_nr_mismatch, libjgb), Int32, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}),
        a.method.ctx.handle, sp, sq))
    return sp[], sq[]
end

function solve!(a::AcPowerFlowB200)
    syncState!(a)
    check(a.method.ctx, ccall((:jgb_nr_solve, libjgb), Int32, (Ptr{Cvoid},), a.method.ctx.handle))
    a.method.iteration += 1
    pullState!(a)
    return nothing
end

function powerFlow!(a::AcPowerFlowB200; iteration::Int64 = 20, tolerance::Float64 = 1e-8,
    power::Bool = false, current::Bool = false, verbose::Int64 = 0)
    pushState!(a)
    it, sp, sq = Ref{Int64}(0), Ref{Float64}(0), Ref{Float64}(0)
    check(a.method.ctx, ccall((:jgb_nr_run, libjgb), Int32,
        (Ptr{Cvoid}, Int64, Float64, Ref{Int64}, Ref{Float64}, Ref{Float64}),
        a.method.ctx.handle, iteration, tolerance, it, sp, sq))
    a.method.iteration = it[]
    pullState!(a)
    power && power!(a)
    current && current!(a)
    return nothing
end

##### Gauss-Newton WLS #####
mutable struct GaussNewtonB200
    mean::Vector{Float64}
    residual::Vector{Float64}
    increment::Vector{Float64}
    type::Vector{Int8}
    index::Vector{Int64}
    range::Vector{Int64}
    objective::Float64
    iteration::Int64
    ctx::Ctx
end

mutable struct AcStateEstimationB200 <: AC
    voltage::Polar
    power::JuliaGrid.AcPower
    current::JuliaGrid.AcCurrent
    method::GaussNewtonB200
    system::PowerSystem
    monitoring::Measurement
    deviceMagnitude::Vector{Float64}
    deviceAngle::Vector{Float64}
    reference::Any                 # the reference's AcStateEstimation, built on the first power! call (post-processing only)
end

function pushState!(a::AcStateEstimationB200)
    check(a.method.ctx, ccall((:jgb_wls_set_state, libjgb), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
        a.method.ctx.handle, a.voltage.magnitude, a.voltage.angle))
    copyto!(a.deviceMagnitude, a.voltage.magnitude)
    copyto!(a.deviceAngle, a.voltage.angle)
    return nothing
end
function pullState!(a::AcStateEstimationB200)
    check(a.method.ctx, ccall((:jgb_wls_get_state, libjgb), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
        a.method.ctx.handle, a.voltage.magnitude, a.voltage.angle))
    copyto!(a.deviceMagnitude, a.voltage.magnitude)
    copyto!(a.deviceAngle, a.voltage.angle)
    return nothing
end
syncState!(a::AcStateEstimationB200) =
    (a.voltage.magnitude == a.deviceMagnitude && a.voltage.angle == a.deviceAngle) || pushState!(a)

"setInitialPoint!(analysis) for state estimation (acStateEstimation.jl): back to the bus voltages of the system"
function setInitialPoint!(a::AcStateEstimationB200)
    copyto!(a.voltage.magnitude, a.system.bus.voltage.magnitude)
    copyto!(a.voltage.angle, a.system.bus.voltage.angle)
    a.method.iteration = 0
    pushState!(a)
    return nothing
end

# power! (postprocessing/acAnalysis.jl:221-262) is a method of the reference's AcStateEstimation: the reference object is
# built once, on the first call, and only for this post-processing step; current! (:672-700) is generic on `AC`.
function power!(a::AcStateEstimationB200)
    if a.reference === nothing
        a.reference = gaussNewton(a.monitoring)
        a.power = a.reference.power
    end
    copyto!(a.reference.voltage.magnitude, a.voltage.magnitude)
    copyto!(a.reference.voltage.angle, a.voltage.angle)
    power!(a.reference)
    return nothing
end

function gaussNewton(monitoring::Measurement, ::Type{B200}; device::Integer = 0)
    system = monitoring.system
    jcb, mean, pcs, rsd, type, index, range, power, current, _ = JuliaGrid.acWLS(system, monitoring)
    ac, bus, br = system.model.ac, system.bus, system.branch
    ctx = Ctx(device)
    Y, Yt = ac.nodalMatrix, ac.nodalMatrixTranspose
    GC.@preserve Y Yt jcb pcs begin
        check(ctx, ccall((:jgb_wls_setup, libjgb), Int32,
            (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int8}, Ptr{Int64}, Ptr{Int64},
             Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64},
             Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            ctx.handle, bus.number, length(mean), bus.layout.slack, jcb.colptr, jcb.rowval, type, index, range,
            pcs.colptr, pcs.rowval, pcs.nzval, Y.colptr, Y.rowval, pointer(reinterpret(Float64, Y.nzval)),
            pointer(reinterpret(Float64, Yt.nzval)), br.number, br.layout.from, br.layout.to,
            br.parameter.conductance, br.parameter.susceptance, br.parameter.turnsRatio, br.parameter.shiftAngle,
            pointer(reinterpret(Float64, ac.admittance))))
    end
    check(ctx, ccall((:jgb_wls_set_mean, libjgb), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.handle, mean))
    m = GaussNewtonB200(mean, rsd, zeros(2 * bus.number), type, index, range, 0.0, 0, ctx)
    a = AcStateEstimationB200(Polar(copy(bus.voltage.magnitude), copy(bus.voltage.angle)), power, current, m,
        system, monitoring, zeros(bus.number), zeros(bus.number), nothing)
    pushState!(a)
    return a
end

function increment!(a::AcStateEstimationB200)
    syncState!(a)
    mi, ob = Ref{Float64}(0), Ref{Float64}(0)
    check(a.method.ctx, ccall((:jgb_wls_increment, libjgb), Int32, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}),
        a.method.ctx.handle, mi, ob))
    a.method.objective = ob[]
    return mi[]
end

function solve!(a::AcStateEstimationB200)
    check(a.method.ctx, ccall((:jgb_wls_solve, libjgb), Int32, (Ptr{Cvoid},), a.method.ctx.handle))
    a.method.iteration += 1
    pullState!(a)
    return nothing
end

function stateEstimation!(a::AcStateEstimationB200; iteration::Int64 = 40, tolerance::Float64 = 1e-8,
    power::Bool = false, current::Bool = false, verbose::Int64 = 0)
    syncState!(a)
    it, mi, ob = Ref{Int64}(0), Ref{Float64}(0), Ref{Float64}(0)
    check(a.method.ctx, ccall((:jgb_wls_run, libjgb), Int32,
        (Ptr{Cvoid}, Int64, Float64, Ref{Int64}, Ref{Float64}, Ref{Float64}),
        a.method.ctx.handle, iteration, tolerance, it, mi, ob))
    a.method.iteration, a.method.objective = it[], ob[]
    pullState!(a)
    power && power!(a)
    current && current!(a)
    return nothing
end

"""
    residualTest!(analysis::AcStateEstimationB200; threshold = 3.0)

Largest normalised residual on the device (selected inverse of the gain factor + row projection); the device returns
the 1-based row, the label / status bookkeeping below is JuliaGrid's own (stateEstimation/badData.jl:224-282).
"""
function JuliaGrid.residualTest!(a::AcStateEstimationB200; threshold::Float64 = 3.0)
    ctx, se, mon = a.method.ctx, a.method, a.monitoring
    rn, idx = Ref{Float64}(0), Ref{Int64}(0)
    check(ctx, ccall((:jgb_wls_residual_test, libjgb), Int32, (Ptr{Cvoid}, Float64, Ref{Float64}, Ref{Int64}, Ptr{Float64}),
        ctx.handle, threshold, rn, idx, C_NULL))
    bad = JuliaGrid.ResidualTest(rn[] > threshold, rn[], "", idx[])
    bad.index == 0 && return bad
    nv, na, nw, nq = mon.voltmeter.number, mon.ammeter.number, mon.wattmeter.number, mon.varmeter.number
    rows = [bad.index]
    if bad.index < se.range[2]
        bad.label, k = JuliaGrid.getLabelIdx(mon.voltmeter.label, bad.index)
        bad.detect && (mon.voltmeter.magnitude.status[k] = 0)
    elseif bad.index < se.range[3]
        bad.label, k = JuliaGrid.getLabelIdx(mon.ammeter.label, bad.index - nv)
        bad.detect && (mon.ammeter.magnitude.status[k] = 0)
    elseif bad.index < se.range[4]
        bad.label, k = JuliaGrid.getLabelIdx(mon.wattmeter.label, bad.index - nv - na)
        bad.detect && (mon.wattmeter.active.status[k] = 0)
    elseif bad.index < se.range[5]
        bad.label, k = JuliaGrid.getLabelIdx(mon.varmeter.label, bad.index - nv - na - nw)
        bad.detect && (mon.varmeter.reactive.status[k] = 0)
    else
        loc = bad.index - nv - na - nw - nq
        bad.label, k = JuliaGrid.getLabelIdx(mon.pmu.label, (loc + 1) ÷ 2)
        if bad.detect
            if mon.pmu.layout.polar[k]
                se.type[bad.index] in (2, 3, 4, 5, 12) ? (mon.pmu.magnitude.status[k] = 0) : (mon.pmu.angle.status[k] = 0)
            else
                mon.pmu.magnitude.status[k] = mon.pmu.angle.status[k] = 0
                push!(rows, iseven(loc) ? bad.index - 1 : bad.index + 1)
            end
        end
    end
    if bad.detect
        for r in rows
            check(ctx, ccall((:jgb_wls_remove_row, libjgb), Int32, (Ptr{Cvoid}, Int64), ctx.handle, r))
            se.mean[r] = 0.0
            se.type[r] = 0
        end
        se.iteration = 0
    end
    return bad
end

# ---- fast Newton-Raphson BX / XB: JuliaGrid builds the two constant Jacobians, the device factors them once ----------
mutable struct FastNewtonRaphsonB200
    ctx::Ctx
    reference::Any            # the reference's own analysis (bus-type fix-ups, start point, B' and B'')
    iteration::Int64
end

function fastNewtonRaphsonB200(system::PowerSystem; bx::Bool = true, device::Integer = 0)
    ref = bx ? JuliaGrid.fastNewtonRaphsonBX(system) : JuliaGrid.fastNewtonRaphsonXB(system)
    Y, Yt, bus = system.model.ac.nodalMatrix, system.model.ac.nodalMatrixTranspose, system.bus
    Bp, Bq = ref.method.active.jacobian, ref.method.reactive.jacobian
    ctx = Ctx(device)
    GC.@preserve Y Yt Bp Bq check(ctx, ccall((:jgb_fnr_setup, libjgb), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int8}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64},
         Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
        ctx.handle, bus.number, Y.colptr, Y.rowval, pointer(reinterpret(Float64, Yt.nzval)), bus.layout.type,
        bus.layout.slack, Bp.colptr, Bp.rowval, Bp.nzval, Bq.colptr, Bq.rowval, Bq.nzval))
    check(ctx, ccall((:jgb_fnr_set_injection, libjgb), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx.handle, bus.supply.active, bus.supply.reactive, bus.demand.active, bus.demand.reactive))
    check(ctx, ccall((:jgb_fnr_set_state, libjgb), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
        ctx.handle, ref.voltage.magnitude, ref.voltage.angle))
    return FastNewtonRaphsonB200(ctx, ref, 0)
end

"powerFlow! for the fast method on the device; the voltages land in `a.reference.voltage`."
function powerFlowB200!(a::FastNewtonRaphsonB200; iteration::Int64 = 20, tolerance::Float64 = 1e-8)
    it, sp, sq = Ref{Int64}(0), Ref{Float64}(0), Ref{Float64}(0)
    rc = check(a.ctx, ccall((:jgb_fnr_run, libjgb), Int32, (Ptr{Cvoid}, Int64, Float64, Ref{Int64}, Ref{Float64}, Ref{Float64}),
        a.ctx.handle, iteration, tolerance, it, sp, sq))
    a.iteration = it[]
    v = a.reference.voltage
    check(a.ctx, ccall((:jgb_fnr_get_state, libjgb), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
        a.ctx.handle, v.magnitude, v.angle))
    return rc == 0
end

# ---- linear analyses (DC power flow, DC and PMU state estimation): one device factorisation, many right-hand sides ----
# The reference's own setup functions build every table (`dcPowerFlow`, `dcStateEstimation`, `pmuStateEstimation`
# with the default LU tag); only `factorization / solution!` is replaced (src/backend/utility.jl:470-586).

"Device LDLt of a symmetric SparseMatrixCSC; `skip` = slack index whose row/column becomes the identity (0: none)."
mutable struct LinearB200
    ctx::Ctx
    n::Int64
    m::Int64
end

function LinearB200(A::SparseMatrixCSC{Float64, Int64}; skip::Integer = 0, device::Integer = 0)
    ctx = Ctx(device)
    GC.@preserve A check(ctx, ccall((:jgb_lin_setup, libjgb), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64),
        ctx.handle, size(A, 1), A.colptr, A.rowval, A.nzval, skip))
    return LinearB200(ctx, size(A, 1), 0)
end

"P = precision * coefficient (m x n): right-hand sides are formed on the device as P' z."
function projection!(F::LinearB200, P::SparseMatrixCSC{Float64, Int64})
    GC.@preserve P check(F.ctx, ccall((:jgb_lin_projection, libjgb), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), F.ctx.handle, size(P, 1), P.colptr, P.rowval, P.nzval))
    F.m = size(P, 1)
    return F
end

"`solution!` for every column of B (n x R, one right-hand side per column; Julia's column-major = the ABI's [R][n])."
function solution(F::LinearB200, B::Matrix{Float64})
    X = similar(B)
    check(F.ctx, ccall((:jgb_lin_solve, libjgb), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}),
        F.ctx.handle, size(B, 2), B, X))
    return X
end

"x_r = A^-1 P' z_r for every column of Z (m x R, one measurement vector per Monte-Carlo draw)."
function projectedSolution(F::LinearB200, Z::Matrix{Float64})
    X = Matrix{Float64}(undef, F.n, size(Z, 2))
    check(F.ctx, ccall((:jgb_lin_solve_projected, libjgb), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}),
        F.ctx.handle, size(Z, 2), Z, X))
    return X
end

"DC power flow angles for every column of `rhs` (pf.rhs per injection scenario, dcPowerFlow.jl:99-102)."
function dcPowerFlowB200(system::PowerSystem, rhs::Matrix{Float64}; device::Integer = 0)
    dc, slack = system.model.dc, system.bus.layout.slack
    F = LinearB200(dc.nodalMatrix; skip = slack, device = device)
    θ = solution(F, rhs)
    θ[slack, :] .= 0.0                                  # addSlackAngle! (backend/utility.jl:610-622)
    θ .+= system.bus.voltage.angle[slack]
    return θ
end

"Linear WLS estimates for every column of Z: `se` is the `WLS` method of dcStateEstimation / pmuStateEstimation."
function linearEstimationB200(se, Z::Matrix{Float64}; slack::Integer = 0, device::Integer = 0)
    H = copy(se.coefficient)
    slack > 0 && (H[:, slack] .= 0.0)                   # removeColumn (backend/sparse.jl:155-163)
    P = se.precision * H
    G = transpose(H) * P
    slack > 0 && (G[slack, slack] = 1.0)
    F = projection!(LinearB200(sparse((G + transpose(G)) / 2); skip = slack, device = device), P)
    return projectedSolution(F, Z)
end

export B200, fastNewtonRaphsonB200, powerFlowB200!, LinearB200, projection!, solution, projectedSolution, dcPowerFlowB200, linearEstimationB200

end # module
