# BatchSimulation.jl — Julia host side of the B200 batched C172 engine (thin `ccall` layer over
# include/flightbatch.h). SOURCE ONLY in this repository: Julia is not installed in the build image, so this file has
# not been executed here; the same C ABI is exercised by the Python mirror (flight.jl_b200/__init__.py) and by the C
# driver tests/abi_driver.c, and tests/test_abi_and_host.py checks THIS file against the header mechanically: the layout
# of every struct below (field by field against offsetof in C) and the arity of every `ccall` against its prototype.
#
# It mirrors the user surface of Flight.jl for N members stepped in lock-step:
#   Simulation(world; dt, Δt, t_start, t_end, saveat)  lib/FlightCore/src/sim.jl:183-196  -> BatchSimulation(world, N; ...)
#   init!(sim, init)                                   sim.jl:390-414                      -> init!(bsim, C172.TrimParameters / C172.Init columns)
#   step!(sim, Δt, true) / run!(sim)                   sim.jl:386,611                      -> step!(bsim, Δt) / run!(bsim)
#   TimeSeries(sim)                                    sim.jl:644-704                      -> TimeSeries(bsim, :θ)
#   sim.t / sim.x / sim.u / sim.y                      sim.jl:265-275                      -> bsim.t, bsim.x, getinput/setinput!, getoutput
# There is no CPU fallback: every call goes to libflightbatch.so and throws FlightBatchError on a non-zero return.
module BatchSimulations

export BatchSimulation, state_async!, wait_state, FlightBatchError, init!, step!, run!, setinput!, getinput, getoutput, schedule_inputs!, set_segment!, BatchTimeSeries, BatchLinearizedSS, getstate, getinputs, getoutputs, setinputs!, setstate!,
       status, stats, linearize, TimeSeriesColumn, log_stats, comm_unique_id, comm_init!, log_stats_allreduce!,
       BatchGroup, shard, group_size, release_trim_failed!

const libflightbatch = get(ENV, "FLIGHTBATCH_LIB", joinpath(@__DIR__, "..", "libflightbatch.so"))
const tables_path = get(ENV, "FLIGHTBATCH_TABLES", joinpath(@__DIR__, "..", "data", "c172_tables.bin"))

struct FlightBatchError <: Exception
    code::Int32
    msg::String
end
Base.showerror(io::IO, e::FlightBatchError) = print(io, "flightbatch error ", e.code, ": ", e.msg)

last_error() = unsafe_string(ccall((:fb_last_error, libflightbatch), Cstring, ()))
check(rc::Int32) = rc == 0 ? nothing : throw(FlightBatchError(rc, last_error()))

# enums of include/flightbatch.h (kept in the same order; FB_API_VERSION is checked at load)
@enum Variant::Int32 C172S = 0 C172X0 = 1 C172X1 = 2 C172X2 = 3 C172XA2 = 4   # C172XA2: Cessna172Xv0 with Actuator2 actuation and an optional VariablePitch propeller
@enum KinematicsKind::Int32 WA = 0 ECEF = 1 NED = 2
const FB_API_VERSION = 5

# struct fb_config (must match the C layout field by field)
struct FbConfig
    variant::Int32
    kinematics::Int32
    n_members::Int64
    member_offset::Int64
    dt::Float64
    Dt::Float64
    t_start::Float64
    t_end::Float64
    device::Int32
    log_every::Int32
    log_n_fields::Int32
    log_fields::NTuple{32,Int32}
    steps_per_launch::Int32
    log_flags::Int32          # FB_LOG_MEMBERS (1) | FB_LOG_MOMENTS (2); 0 = both
    prop_pitch_n::Int32       # C172XA2 only: > 1 = Propeller(VariablePitch(lo, hi; length = n)), 0 / 1 = FixedPitch
    reserved0::Int32
    prop_pitch_lo::Float64
    prop_pitch_hi::Float64
    act2_omega_n::Float64     # C172XA2 only: Actuator2 parameters (0 = the reference defaults 5 * 2π, 0.6)
    act2_zeta::Float64
end

mutable struct BatchSimulation
    handle::Ptr{Cvoid}
    n::Int
    nx::Int
    dt::Float64
    device::Int32
    log_fields::Vector{Int32}
    function BatchSimulation(variant::Variant, kin::KinematicsKind, n::Integer;
                             dt = 0.02, Δt = dt, t_start = 0.0, t_end = 10000.0, saveat = nothing,
                             log_fields = Int32[], device = 0, member_offset = 0, steps_per_launch = 0, log_flags = 0,
                             prop_pitch = nothing, act2 = (0.0, 0.0))   # prop_pitch = (lower, upper, length) for a VariablePitch propeller
        @assert ccall((:fb_api_version, libflightbatch), Int32, ()) == FB_API_VERSION
        lf = ntuple(i -> i <= length(log_fields) ? Int32(log_fields[i]) : Int32(0), 32)
        log_every = (saveat === nothing || isempty(log_fields)) ? 0 : max(1, round(Int, saveat / dt))
        cfg = Ref(FbConfig(Int32(variant), Int32(kin), n, member_offset, dt, Δt, t_start, t_end, device, log_every,
                           length(log_fields), lf, steps_per_launch, log_flags,
                           prop_pitch === nothing ? 0 : prop_pitch[3], 0, prop_pitch === nothing ? 0.0 : prop_pitch[1],
                           prop_pitch === nothing ? 0.0 : prop_pitch[2], act2[1], act2[2]))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:fb_create, libflightbatch), Int32, (Ref{FbConfig}, Ref{Ptr{Cvoid}}), cfg, h))
        blob = read(tables_path)
        GC.@preserve blob check(ccall((:fb_load_tables, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t), h[], blob, length(blob)))
        nx = Ref{Int32}(0)
        check(ccall((:fb_nx, libflightbatch), Int32, (Ptr{Cvoid}, Ref{Int32}), h[], nx))
        sim = new(h[], n, nx[], dt, Int32(device), collect(Int32, log_fields))
        finalizer(s -> (s.handle != C_NULL && ccall((:fb_destroy, libflightbatch), Int32, (Ptr{Cvoid},), s.handle); s.handle = C_NULL), sim)
    end
end

# struct fb_trim_params_soa: 12 column pointers + payload[5]; C_NULL = the reference default of that field
const TRIM_FIELDS = (:lat, :lon, :h_e, :psi_nb, :EAS, :gamma_wb_n, :psi_wb_dot, :theta_wb_dot, :beta_a, :fuel_load, :mixture, :flaps,
                     :m_pilot, :m_copilot, :m_lpass, :m_rpass, :m_baggage)

column(sim::BatchSimulation, v::Real) = fill(Float64(v), sim.n)
column(sim::BatchSimulation, v::AbstractVector) = (length(v) == sim.n || throw(ArgumentError("column length != N")); Vector{Float64}(v))

"""
    init!(sim; EAS = ..., h_e = ..., flaps = ..., ...)   # init!(sim, C172.TrimParameters(...)) for every member

Each keyword is a scalar (broadcast) or a length-N vector. Returns `(success::BitVector, trim_state::Matrix (N×7), cost)`.
Members whose trim does not reach the reference's stop value (cost ≤ 1e-16, c172.jl:927) get status 6 and are frozen. This is
a deliberate deviation: the reference only warns, assigns the best trim state found and keeps simulating (c172.jl:934-940);
here a member that failed to trim is excluded from stepping and from every ensemble moment, and reported through `success`.
"""
function init!(sim::BatchSimulation; kw...)
    cols = Dict{Symbol,Vector{Float64}}(k => column(sim, v) for (k, v) in kw)
    all(k -> k in TRIM_FIELDS, keys(cols)) || throw(ArgumentError("unknown TrimParameters field"))
    ptrs = [haskey(cols, f) ? pointer(cols[f]) : Ptr{Float64}(C_NULL) for f in TRIM_FIELDS]
    ts = Matrix{Float64}(undef, sim.n, 7)
    cost = Vector{Float64}(undef, sim.n)
    GC.@preserve cols ptrs begin
        check(ccall((:fb_init_trim, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Float64}, Ptr{Float64}),
                    sim.handle, ptrs, ts, cost))
    end
    (status(sim) .!= 6, ts, cost)
end

step!(sim::BatchSimulation, Δt::Real = sim.dt) = check(ccall((:fb_step, libflightbatch), Int32, (Ptr{Cvoid}, Float64), sim.handle, Δt))
run!(sim::BatchSimulation) = check(ccall((:fb_run, libflightbatch), Int32, (Ptr{Cvoid},), sim.handle))
Base.wait(sim::BatchSimulation) = check(ccall((:fb_sync, libflightbatch), Int32, (Ptr{Cvoid},), sim.handle))

"""
    state_async!(x, sim); ...; wait_state(sim)

Pipelined read of `sim.x` into the caller's `N x nx` matrix `x` (page-locked memory for a truly asynchronous copy):
the state as of the work submitted so far is snapshotted on the device and copied while later `step!` calls run.
The caller keeps `x` alive (`GC.@preserve`) until `wait_state` returns.
"""
state_async!(x::Matrix{Float64}, sim::BatchSimulation) =
    check(ccall((:fb_get_state_async, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Float64}), getfield(sim, :handle), x))
wait_state(sim::BatchSimulation) = check(ccall((:fb_wait_state, libflightbatch), Int32, (Ptr{Cvoid},), getfield(sim, :handle)))

function Base.getproperty(sim::BatchSimulation, s::Symbol)
    if s === :t
        t = Ref(0.0); n = Ref{Int64}(0)
        check(ccall((:fb_time, libflightbatch), Int32, (Ptr{Cvoid}, Ref{Float64}, Ref{Int64}), getfield(sim, :handle), t, n))
        return t[]
    elseif s === :x   # N × nx, column c = component c of `X(md)` (SURVEY.md App. A)
        x = Matrix{Float64}(undef, getfield(sim, :n), getfield(sim, :nx))
        check(ccall((:fb_get_state, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Float64}), getfield(sim, :handle), x))
        return x
    else
        return getfield(sim, s)
    end
end

# mdl.<...>.u.<field> = value : `field` is an fb_input id (see include/flightbatch.h), value a scalar or an N-vector
function setinput!(sim::BatchSimulation, field::Integer, v::Real)
    r = Ref(Float64(v))
    check(ccall((:fb_set_input, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Ref{Float64}, Int64), sim.handle, field, r, 0))
end
function setinput!(sim::BatchSimulation, field::Integer, v::AbstractVector)
    c = column(sim, v)
    check(ccall((:fb_set_input, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64), sim.handle, field, c, 1))
end
function getinput(sim::BatchSimulation, field::Integer)
    out = Vector{Float64}(undef, sim.n)
    check(ccall((:fb_get_input, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), sim.handle, field, out)); out
end
function getoutput(sim::BatchSimulation, field::Integer)
    out = Vector{Float64}(undef, sim.n)
    check(ccall((:fb_get_output, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), sim.handle, field, out)); out
end

# Cessna172Xv2 guidance (c172x_gdc.jl): fb_input ids of GuidanceLaws.u / SegmentGuidance.u
const FB_U_GDC_MODE_REQ, FB_U_GDC_HOR_REQ, FB_U_GDC_VRT_REQ = Int32(70), Int32(71), Int32(72)
const FB_U_GDC_LAT1 = Int32(73)   # then LON1, H1, LAT2, LON2, H2

"""
    set_segment!(sim, p1, p2)

`avionics.gdc.seg.u.target = Segment(p1, p2)` for every member: `p1`, `p2` are `(lat, lon, h_e)` tuples of scalars or
N-vectors, i.e. the fields of the reference's `Geographic{LatLon,Ellipsoidal}` end points. With Flight.jl loaded, pass
`(seg.p1.loc.ϕ, seg.p1.loc.λ, Float64(seg.p1.h))` of a `C172XGuidance.Segment`, whose constructor has already refused
vertical or degenerate segments (c172x_gdc.jl:44-58).
"""
function set_segment!(sim::BatchSimulation, p1, p2)
    for (k, v) in enumerate((p1..., p2...))
        setinput!(sim, FB_U_GDC_LAT1 + Int32(k - 1), v)
    end
end

"declarative replacement of `user_callback!` (sim.jl:331-341): piecewise-constant per-member inputs"
function schedule_inputs!(sim::BatchSimulation, member::Vector{Int64}, step::Vector{Int64}, field::Vector{Int32}, value::Vector{Float64})
    n = length(member)
    @assert length(step) == n && length(field) == n && length(value) == n
    check(ccall((:fb_schedule_inputs, libflightbatch), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}),
                sim.handle, n, member, step, field, value))
end

"release_trim_failed!(sim): the reference's own handling of a failed trim (warn, keep the best state found, go on, c172.jl:934-940); returns the number of members released"
function release_trim_failed!(sim::BatchSimulation)
    n = Ref{Int64}(0)
    check(ccall((:fb_release_trim_failed, libflightbatch), Int32, (Ptr{Cvoid}, Ref{Int64}), sim.handle, n)); Int(n[])
end
function status(sim::BatchSimulation)
    st = Vector{Int32}(undef, sim.n)
    check(ccall((:fb_status, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Int32}), sim.handle, st)); st
end

struct Moments
    count::Int64
    sum::Float64
    sum_sq::Float64
    min::Float64
    max::Float64
end
function stats(sim::BatchSimulation, field::Integer)
    m = Ref(Moments(0, 0.0, 0.0, 0.0, 0.0))
    check(ccall((:fb_stats, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Ref{Moments}), sim.handle, field, m)); m[]
end

"""
    log_stats(sim, field, first, count) -> Vector{Moments}

Ensemble moments of a logged `fb_output` field at log samples `first .. first + count - 1` (0-based), accumulated inside
the integrator as each sample was taken; after `log_stats_allreduce!` they cover every shard of the ensemble.
"""
function log_stats(sim::BatchSimulation, field::Integer, first::Integer, count::Integer)
    m = Vector{Moments}(undef, count)
    check(ccall((:fb_log_stats, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Moments}), sim.handle, field, first, count, m)); m
end

"128-byte NCCL unique id: rank 0 creates it, the host distributes it (MPI.bcast, a file, ...), every rank calls `comm_init!`"
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:fb_comm_unique_id, libflightbatch), Int32, (Ptr{UInt8},), id)); id
end
comm_init!(sim::BatchSimulation, n_ranks::Integer, rank::Integer, id::Vector{UInt8}) =
    check(ccall((:fb_comm_init, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), sim.handle, n_ranks, rank, id))
"all-reduce (NCCL inside the library, side stream, asynchronous) of the moments of the log samples completed since the last call"
log_stats_allreduce!(sim::BatchSimulation) = check(ccall((:fb_log_stats_allreduce, libflightbatch), Int32, (Ptr{Cvoid},), sim.handle))

"TimeSeries(sim) column: `(t, values)` with `values` of size N × n_samples for one logged `fb_output` field"
function TimeSeriesColumn(sim::BatchSimulation, field::Integer)
    ns = Ref{Int64}(0); nf = Ref{Int32}(0)
    check(ccall((:fb_log_info, libflightbatch), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int32}), sim.handle, ns, nf))
    t = Vector{Float64}(undef, ns[]); v = Matrix{Float64}(undef, sim.n, ns[])
    check(ccall((:fb_log_times, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Float64}), sim.handle, t))
    check(ccall((:fb_log_fetch, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), sim.handle, field, v))
    (t, v)
end

# ---- TimeSeries(sim) with the reference's lazy drilling (sim.jl:666-676): `ts.aircraft.vehicle.kinematics.e_nb.θ`
# Path under the root `y` of every fb_output id (same table as Y_PATHS in flight.jl_b200/__init__.py).
const Y_PATHS = Dict{String,Int32}(
    "aircraft.vehicle.kinematics.e_nb.ψ" => 0, "aircraft.vehicle.kinematics.e_nb.θ" => 1, "aircraft.vehicle.kinematics.e_nb.φ" => 2,
    "aircraft.vehicle.kinematics.ϕ_λ.ϕ" => 3, "aircraft.vehicle.kinematics.ϕ_λ.λ" => 4, "aircraft.vehicle.kinematics.h_e" => 5,
    "aircraft.vehicle.kinematics.h_o" => 6, "aircraft.vehicle.kinematics.v_eb_n[1]" => 7, "aircraft.vehicle.kinematics.v_eb_n[2]" => 8,
    "aircraft.vehicle.kinematics.v_eb_n[3]" => 9, "aircraft.vehicle.kinematics.v_gnd" => 10, "aircraft.vehicle.kinematics.χ_gnd" => 11,
    "aircraft.vehicle.kinematics.γ_gnd" => 12, "aircraft.vehicle.kinematics.ω_wb_b[1]" => 13, "aircraft.vehicle.kinematics.ω_wb_b[2]" => 14,
    "aircraft.vehicle.kinematics.ω_wb_b[3]" => 15, "aircraft.vehicle.airflow.TAS" => 16, "aircraft.vehicle.airflow.EAS" => 17,
    "aircraft.vehicle.airflow.CAS" => 18, "aircraft.vehicle.airflow.M" => 19, "aircraft.vehicle.airflow.q" => 20,
    "aircraft.vehicle.systems.aero.α" => 21, "aircraft.vehicle.systems.aero.β" => 22, "aircraft.vehicle.systems.aero.α_filt" => 23,
    "aircraft.vehicle.systems.aero.β_filt" => 24, "aircraft.vehicle.systems.aero.stall" => 25, "aircraft.vehicle.systems.aero.coeffs.C_L" => 26,
    "aircraft.vehicle.systems.aero.coeffs.C_D" => 27, "aircraft.vehicle.systems.ldg.left.strut.wow" => 28,
    "aircraft.vehicle.systems.ldg.right.strut.wow" => 29, "aircraft.vehicle.systems.ldg.nose.strut.wow" => 30,
    "aircraft.vehicle.systems.ldg.left.strut.ξ" => 31, "aircraft.vehicle.systems.ldg.right.strut.ξ" => 32, "aircraft.vehicle.systems.ldg.nose.strut.ξ" => 33,
    "aircraft.vehicle.systems.ldg.left.strut.F_dmp_zs" => 34, "aircraft.vehicle.systems.ldg.right.strut.F_dmp_zs" => 35,
    "aircraft.vehicle.systems.ldg.nose.strut.F_dmp_zs" => 36, "aircraft.vehicle.systems.pwp.engine.ω" => 37,
    "aircraft.vehicle.systems.pwp.engine.state" => 38, "aircraft.vehicle.systems.pwp.engine.MAP" => 39, "aircraft.vehicle.systems.pwp.engine.P_shaft" => 40,
    "aircraft.vehicle.systems.pwp.engine.τ_shaft" => 41, "aircraft.vehicle.systems.pwp.engine.ṁ" => 42, "aircraft.vehicle.systems.fuel.m_total" => 43,
    "aircraft.vehicle.systems.pwp.propeller.wr_p.F[1]" => 44, "aircraft.vehicle.systems.pwp.propeller.J" => 45, "aircraft.vehicle.dynamics.mp_Σ_b.m" => 46,
    "aircraft.vehicle.dynamics.f_c_c[1]" => 47, "aircraft.vehicle.dynamics.f_c_c[2]" => 48, "aircraft.vehicle.dynamics.f_c_c[3]" => 49,
    "aircraft.avionics.ctl.lon.mode" => 50, "aircraft.avionics.ctl.lat.mode" => 51, "aircraft.avionics.ctl.lon.throttle_cmd" => 52,
    "aircraft.avionics.ctl.lon.elevator_cmd" => 53, "aircraft.avionics.ctl.lat.aileron_cmd" => 54, "aircraft.avionics.ctl.lat.rudder_cmd" => 55,
    "aircraft.avionics.ctl.lon.h_state" => 56, "aircraft.avionics.gdc.mode" => 57, "aircraft.avionics.gdc.seg.hor_gdc" => 58,
    "aircraft.avionics.gdc.seg.vrt_gdc" => 59, "aircraft.avionics.gdc.seg.Δχ" => 60, "aircraft.avionics.gdc.seg.χ_ref" => 61,
    "aircraft.avionics.gdc.seg.h_ref" => 62, "aircraft.avionics.gdc.seg.data.χ_12" => 63, "aircraft.avionics.gdc.seg.data.γ_12" => 64,
    "aircraft.avionics.gdc.seg.data.s_12" => 65, "aircraft.avionics.gdc.seg.data.s_1b" => 66, "aircraft.avionics.gdc.seg.data.e_sb" => 67,
    "aircraft.avionics.gdc.seg.data.v_sb" => 68)

"`BatchTimeSeries(sim)`: lazy view over the device log; a leaf is `(t, values)` with `values` of size N × n_samples"
struct BatchTimeSeries
    sim::BatchSimulation
    path::String
end
BatchTimeSeries(sim::BatchSimulation) = BatchTimeSeries(sim, "")
function _resolve(ts::BatchTimeSeries, path::String)
    haskey(Y_PATHS, path) && return TimeSeriesColumn(getfield(ts, :sim), Y_PATHS[path])
    any(k -> startswith(k, path * ".") || startswith(k, path * "["), keys(Y_PATHS)) && return BatchTimeSeries(getfield(ts, :sim), path)
    error("TimeSeries has no path $path")
end
function Base.getproperty(ts::BatchTimeSeries, s::Symbol)
    p = getfield(ts, :path)
    _resolve(ts, isempty(p) ? String(s) : p * "." * String(s))
end
Base.getindex(ts::BatchTimeSeries, i::Integer) = _resolve(ts, getfield(ts, :path) * "[$i]")
Base.propertynames(ts::BatchTimeSeries) = begin
    p = getfield(ts, :path); n = isempty(p) ? 0 : length(p) + 1
    unique(Symbol(first(split(first(split(chop(k, head = n, tail = 0), ".")), "["))) for k in keys(Y_PATHS) if startswith(k, p))
end

"""
    linearize(sim)  ->  (A, B, C, D, x0, u0, y0, ẋ0)  # linearize(aircraft, trim_params), aircraftbase.jl:292-341

Forward-difference Jacobians of every member about its current (trimmed) state; arrays are `N × rows × cols`
(member fastest, exactly the layout `fb_linearize` writes). Needs NED kinematics, like the reference's
`XStateSpace` (c172s.jl:269-412, c172x.jl:332-498).
"""
function linearize(sim::BatchSimulation)
    d = zeros(Int32, 3)
    check(ccall((:fb_linearize_dims, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Int32}), sim.handle, d))
    nxs, nus, nys = d
    A = Array{Float64}(undef, sim.n, nxs, nxs); B = Array{Float64}(undef, sim.n, nxs, nus)
    C = Array{Float64}(undef, sim.n, nys, nxs); D = Array{Float64}(undef, sim.n, nys, nus)
    x0 = Matrix{Float64}(undef, sim.n, nxs); u0 = Matrix{Float64}(undef, sim.n, nus); y0 = Matrix{Float64}(undef, sim.n, nys)
    xdot0 = Matrix{Float64}(undef, sim.n, nxs)
    check(ccall((:fb_linearize, libflightbatch), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                sim.handle, A, B, C, D, x0, u0, y0, xdot0))
    (A, B, C, D, x0, u0, y0, xdot0)
end

# ---- LinearizedSS ensembles (linearization.jl:24-111,157-192; README example 3 per member) ----------------------
"N `LinearizedSS` models stepped in lock-step on the GPU; built device-to-device from a trimmed NED `BatchSimulation`"
mutable struct BatchLinearizedSS
    handle::Ptr{Cvoid}
    n::Int64; nx::Int32; nu::Int32; ny::Int32; dt::Float64
end
function BatchLinearizedSS(sim::BatchSimulation)
    d = zeros(Int32, 3)
    check(ccall((:fb_linearize_dims, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Int32}), sim.handle, d))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:fb_lss_create, libflightbatch), Int32, (Int32, Int64, Int32, Int32, Int32, Float64, Ref{Ptr{Cvoid}}),
                sim.device, sim.n, d[1], d[2], d[3], sim.dt, h))
    check(ccall((:fb_lss_from_linearize, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h[], sim.handle))
    lss = BatchLinearizedSS(h[], sim.n, d[1], d[2], d[3], sim.dt)
    finalizer(l -> ccall((:fb_lss_destroy, libflightbatch), Int32, (Ptr{Cvoid},), l.handle), lss)
end
step!(lss::BatchLinearizedSS, Δt::Real = lss.dt) = check(ccall((:fb_lss_step, libflightbatch), Int32, (Ptr{Cvoid}, Float64), lss.handle, Δt))
# N × rows matrices: the C layout [rows][N]. (`ccall` needs a literal function name, hence one method per entry point.)
function getstate(lss::BatchLinearizedSS)
    out = Matrix{Float64}(undef, lss.n, lss.nx)
    check(ccall((:fb_lss_get_state, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Float64}), lss.handle, out)); out
end
function getinputs(lss::BatchLinearizedSS)
    out = Matrix{Float64}(undef, lss.n, lss.nu)
    check(ccall((:fb_lss_get_input, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Float64}), lss.handle, out)); out
end
function getoutputs(lss::BatchLinearizedSS)
    out = Matrix{Float64}(undef, lss.n, lss.ny)
    check(ccall((:fb_lss_get_output, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Float64}), lss.handle, out)); out
end
setinputs!(lss::BatchLinearizedSS, u::Matrix{Float64}) = check(ccall((:fb_lss_set_input, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Float64}), lss.handle, u))
setstate!(lss::BatchLinearizedSS, x::Matrix{Float64}) = check(ccall((:fb_lss_set_state, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Float64}), lss.handle, x))

################################################################################
# BatchGroup: ONE Julia process driving several GPUs (fb_group_*, include/flightbatch.h). The ensemble is sharded in contiguous
# blocks over `devices`; host arrays are GLOBAL (N_total × columns), the library scatters / gathers the shards, steps them
# concurrently and all-reduces the ensemble moments over its own NCCL communicator: no MPI, no Distributed.jl needed.

mutable struct BatchGroup
    handle::Ptr{Cvoid}
    n::Int
    nx::Int
    dt::Float64
    n_dev::Int
    function BatchGroup(variant::Variant, kin::KinematicsKind, n::Integer, devices::AbstractVector{<:Integer};
                        dt = 0.02, Δt = dt, t_start = 0.0, t_end = 10000.0, saveat = nothing, log_fields = Int32[],
                        steps_per_launch = 0, log_flags = 0, prop_pitch = nothing, act2 = (0.0, 0.0))
        lf = ntuple(i -> i <= length(log_fields) ? Int32(log_fields[i]) : Int32(0), 32)
        log_every = (saveat === nothing || isempty(log_fields)) ? 0 : max(1, round(Int, saveat / dt))
        cfg = Ref(FbConfig(Int32(variant), Int32(kin), n, 0, dt, Δt, t_start, t_end, 0, log_every,
                           length(log_fields), lf, steps_per_launch, log_flags,
                           prop_pitch === nothing ? 0 : prop_pitch[3], 0, prop_pitch === nothing ? 0.0 : prop_pitch[1],
                           prop_pitch === nothing ? 0.0 : prop_pitch[2], act2[1], act2[2]))
        dev = Vector{Int32}(devices)
        g = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:fb_group_create, libflightbatch), Int32, (Ref{FbConfig}, Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), cfg, dev, length(dev), g))
        blob = read(tables_path)
        GC.@preserve blob check(ccall((:fb_group_load_tables, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t), g[], blob, length(blob)))
        h = Ref{Ptr{Cvoid}}(C_NULL); off = Ref{Int64}(0); cnt = Ref{Int64}(0); nx = Ref{Int32}(0)
        check(ccall((:fb_group_shard, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}, Ref{Int64}, Ref{Int64}), g[], 0, h, off, cnt))
        check(ccall((:fb_nx, libflightbatch), Int32, (Ptr{Cvoid}, Ref{Int32}), h[], nx))
        grp = new(g[], n, nx[], dt, length(dev))
        finalizer(x -> (x.handle != C_NULL && ccall((:fb_group_destroy, libflightbatch), Int32, (Ptr{Cvoid},), x.handle); x.handle = C_NULL), grp)
    end
end

"(handle, offset, count) of shard `k` (0-based): everything the per-shard `fb_*` entry points offer, on shard-local arrays"
function shard(grp::BatchGroup, k::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL); off = Ref{Int64}(0); cnt = Ref{Int64}(0)
    check(ccall((:fb_group_shard, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}, Ref{Int64}, Ref{Int64}), grp.handle, k, h, off, cnt))
    (h[], Int(off[]), Int(cnt[]))
end

"init!(grp; EAS = ..., h_e = ...): `init!(sim, C172.TrimParameters(...))` for every member of the sharded ensemble (global columns)"
function init!(grp::BatchGroup; kw...)
    cols = Dict(k => (v isa Real ? fill(Float64(v), grp.n) : Vector{Float64}(v)) for (k, v) in kw)
    ptrs = [haskey(cols, f) ? pointer(cols[f]) : Ptr{Float64}(C_NULL) for f in TRIM_FIELDS]
    ts = Matrix{Float64}(undef, grp.n, 7); cost = Vector{Float64}(undef, grp.n)
    GC.@preserve cols ptrs check(ccall((:fb_group_init_trim, libflightbatch), Int32,
        (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Float64}, Ptr{Float64}), grp.handle, ptrs, ts, cost))
    (status(grp) .!= 6, ts, cost)
end
step!(grp::BatchGroup, Δt::Real = grp.dt) = check(ccall((:fb_group_step, libflightbatch), Int32, (Ptr{Cvoid}, Float64), grp.handle, Δt))
run!(grp::BatchGroup) = check(ccall((:fb_group_run, libflightbatch), Int32, (Ptr{Cvoid},), grp.handle))
Base.wait(grp::BatchGroup) = check(ccall((:fb_group_sync, libflightbatch), Int32, (Ptr{Cvoid},), grp.handle))   # as for BatchSimulation
function setinput!(grp::BatchGroup, field::Integer, v)
    col = v isa Real ? [Float64(v)] : Vector{Float64}(v)
    check(ccall((:fb_group_set_input, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64), grp.handle, field, col, v isa Real ? 0 : 1))
end
function getstate(grp::BatchGroup)
    x = Matrix{Float64}(undef, grp.n, grp.nx)
    check(ccall((:fb_group_get_state, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Float64}), grp.handle, x)); x
end
function status(grp::BatchGroup)
    st = Vector{Int32}(undef, grp.n)
    check(ccall((:fb_group_status, libflightbatch), Int32, (Ptr{Cvoid}, Ptr{Int32}), grp.handle, st)); st
end
"ensemble moments (all shards, all-reduced over NCCL inside the library) of a logged field at samples `first .. first + count - 1`"
function log_stats(grp::BatchGroup, field::Integer, first::Integer, count::Integer)
    m = Vector{Moments}(undef, count)
    check(ccall((:fb_group_log_stats, libflightbatch), Int32, (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Moments}), grp.handle, field, first, count, m)); m
end
function group_size(grp::BatchGroup)
    n = Ref{Int32}(0)
    check(ccall((:fb_group_size, libflightbatch), Int32, (Ptr{Cvoid}, Ref{Int32}), grp.handle, n)); Int(n[])
end

end # module
