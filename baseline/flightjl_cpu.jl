# flightjl_cpu.jl — BASELINE.md §4 Tier A: Flight.jl ITSELF as the CPU arm of bench.py.
#
# NOT RUN IN THIS IMAGE (no Julia toolchain). `bench.py --impl reference` executes it whenever `julia` is on PATH and the
# reference has been instantiated under baseline/_ref (a Julia project whose Manifest resolves `Flight`):
#     julia --project=baseline/_ref --threads=<cores> baseline/flightjl_cpu.jl <cfg2|cfg3> <members.txt> <steps> <warmup> [states_out.txt]
# `members.txt` is written by bench.py: one row per member, whitespace separated, the columns named in its first line
# (the same counter-based draws the GPU arm uses, so the two arms integrate the same aircraft).
#
# What it does (SURVEY.md §8d "CPU reference timing"): one `Simulation(SimpleWorld(aircraft); dt = 0.01, Δt = 0.01, save_on =
# false)` per member — one model instance per member: models built from a shared definition alias the controllers' mutable
# gain arrays (lib/FlightCore/src/modeling.jl:144) — `init!` with the member's C172.TrimParameters, the member's inputs, one
# warm-up step to compile, then K timed bench steps, each `Threads.@threads for sim in sims; step!(sim, 1.0, true); end`
# (lib/FlightCore/src/sim.jl:386), the reference's own fixed-step RK4 (sim.jl:186). Prints ONE JSON line
# {"value": aircraft-RK4-steps/s, "threads": n, "seconds": t, "sample": "..."} and, when asked, writes the final states for the
# parity check against the GPU arm (<= 1e-9 relative, tests/ compare in the reference's X(md) order, SURVEY.md App. A).
using Flight
using DelimitedFiles, Printf

function load_members(path)
    raw = readdlm(path; header = true)
    cols = Dict(Symbol(strip(String(n))) => Float64.(raw[1][:, i]) for (i, n) in enumerate(raw[2]))
    return cols, size(raw[1], 1)
end

trim_params(c, i) = C172.TrimParameters(;
    Ob = Geographic(LatLon(c[:lat][i], c[:lon][i]), HEllip(c[:h_e][i])), ψ_nb = c[:psi_nb][i], EAS = c[:EAS][i],
    fuel_load = c[:fuel_load][i], flaps = c[:flaps][i],
    payload = C172.PayloadY(; m_pilot = c[:m_pilot][i], m_copilot = c[:m_copilot][i], m_lpass = c[:m_lpass][i],
                            m_rpass = c[:m_rpass][i], m_baggage = c[:m_baggage][i]))

function build(workload, c, i)
    if workload == "cfg3"   # BASELINE.json configs[2]: README.md:107-127 per member
        world = SimpleWorld(; aircraft = Cessna172Xv2()) |> Model
        sim = Simulation(world; dt = 0.01, Δt = 0.01, t_end = 1e6, save_on = false)
        world.atmosphere.wind.u.N = c[:wind_n][i]
        world.atmosphere.wind.u.E = c[:wind_e][i]
        init!(sim, trim_params(c, i))
        ctl = world.aircraft.avionics.ctl
        ctl.lon.u.mode_req = C172XControl.ModeControlLon.EAS_clm
        ctl.lat.u.mode_req = C172XControl.ModeControlLat.φ_β
        ctl.lon.u.clm_ref = c[:clm_ref][i]
        ctl.lat.u.φ_ref = c[:phi_ref][i]
    else                    # BASELINE.json configs[1]: Cessna172Sv0 on WA kinematics, doublets applied between bench steps
        world = SimpleWorld(; aircraft = Cessna172Sv0()) |> Model
        sim = Simulation(world; dt = 0.01, Δt = 0.01, t_end = 1e6, save_on = false)
        init!(sim, trim_params(c, i))
    end
    return sim
end

function doublet!(sim, ele0, ail0, t)   # the cfg-2 input schedule (tutorial02.md:164-168 shape), in force during second t
    act = sim.mdl.aircraft.vehicle.systems.act
    act.u.elevator = ele0 + (5 <= t < 7 ? 0.1 : (7 <= t < 9 ? -0.1 : 0.0))
    act.u.aileron = ail0 + (20 <= t < 22 ? 0.1 : (22 <= t < 24 ? -0.1 : 0.0))
end

function main()
    workload, path, K, W = ARGS[1], ARGS[2], parse(Int, ARGS[3]), parse(Int, ARGS[4])
    c, n = load_members(path)
    sims = Vector{Any}(undef, n)
    Threads.@threads for i in 1:n
        sims[i] = build(workload, c, i)
    end
    ele0 = [workload == "cfg2" ? Float64(s.mdl.aircraft.vehicle.systems.act.u.elevator) : 0.0 for s in sims]
    ail0 = [workload == "cfg2" ? Float64(s.mdl.aircraft.vehicle.systems.act.u.aileron) : 0.0 for s in sims]
    advance!(k) = Threads.@threads for i in 1:n
        workload == "cfg2" && doublet!(sims[i], ele0[i], ail0[i], k)
        step!(sims[i], 1.0, true)
    end
    for k in 0:max(W, 1)-1     # warm-up: compiles every method on the path
        advance!(k)
    end
    t = @elapsed for k in 0:K-1
        advance!(max(W, 1) + k)
    end
    value = n * 100 * K / t
    @printf("{\"value\": %.6e, \"threads\": %d, \"seconds\": %.3f, \"sample\": \"%d members x %d RK4 steps of the %s ensemble, Flight.jl Simulation per member under Threads.@threads on %d threads (%s, Julia %s), save_on = false\"}\n",
            value, Threads.nthreads(), t, n, 100 * K, workload, Threads.nthreads(), Sys.cpu_info()[1].model, string(VERSION))
    if length(ARGS) >= 5
        writedlm(ARGS[5], permutedims(reduce(hcat, [collect(Float64, s.mdl.x) for s in sims])))
    end
end

main()
