/* flightbatch.h — C ABI of the B200-native batched C172 flight-dynamics engine.
 *
 * Drop-in boundary for ONE path of e271828e/Flight.jl: N independent
 * `Simulation(SimpleWorld(Cessna172...))` objects stepped in lock-step with the
 * reference's fixed-step RK4 + discrete callbacks.  Each entry point names the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns int32: 0 = ok, <0 = API misuse or CUDA error
 *     (text via fb_last_error(), thread-local, valid until the next call);
 *   - per-member physics terminations are NOT API errors: they set status[i]
 *     (FB_ST_*) and freeze that member (reference: SimulationTermination,
 *     lib/FlightCore/src/sim.jl:32,561-570);
 *   - host arrays are struct-of-arrays, member index fastest: a field with C
 *     components is `double a[C][N]` (a Julia `Matrix{Float64}(undef, N, C)`);
 *   - the caller owns every host pointer; the library owns all device memory;
 *   - one handle = one owner thread at a time; there is NO CPU fallback: every
 *     entry point that computes fails with FB_ERR_CUDA when no device exists.
 */
#ifndef FLIGHTBATCH_H
#define FLIGHTBATCH_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FB_API_VERSION 5

/* error codes */
#define FB_OK 0
#define FB_ERR_ARG (-1)
#define FB_ERR_CUDA (-2)
#define FB_ERR_STATE (-3)   /* call out of order (e.g. step before init) */
#define FB_ERR_ALLOC (-4)

/* aircraft variant: lib/FlightApps/src/c172/c172s/c172s0.jl:14-18, c172x/c172x{0,1,2}.jl */
#define FB_C172S 0   /* Cessna172Sv0: mechanical actuation, NoAvionics          */
#define FB_C172X0 1  /* Cessna172Xv0: fly-by-wire actuators, NoAvionics         */
#define FB_C172X1 2  /* Cessna172Xv1: + C172XControl.ControlLaws                */
#define FB_C172X2 3  /* Cessna172Xv2: + C172XGuidance.GuidanceLaws (direct and segment modes) */
/* Cessna172Xv0 assembled from the non-default components the Model tree admits: a fly-by-wire
 * actuation of seven Actuator2 units (second-order lag, X = (v, p) per unit, c172x.jl:68-106: nx grows by 7) and,
 * when fb_config.prop_pitch_n > 1, a Propeller(VariablePitch(prop_pitch_lo, prop_pitch_hi; length = prop_pitch_n))
 * (propellers.jl:341-400) driven by the input FB_U_PROP_PITCH. NoAvionics; fb_linearize is not offered for it. */
#define FB_C172XA2 4
/* kinematic mechanisation: lib/FlightPhysics/src/kinematics.jl:150,250,329 */
#define FB_WA 0
#define FB_ECEF 1
#define FB_NED 2

/* per-member status */
#define FB_ST_RUNNING 0
#define FB_ST_DONE 1
#define FB_ST_CRASH_ANGLE 2    /* landinggear.jl:336-339  strut/terrain angle > 60 deg   */
#define FB_ST_CRASH_RATE 3     /* landinggear.jl:340-343  damper compression > 10 m/s     */
#define FB_ST_BOUNDS 4         /* geodesy.jl:218-221, atmosphere.jl:133, geodesy.jl:83-86 */
#define FB_ST_CONTACT_ASSERT 5 /* landinggear.jl:321                                      */
#define FB_ST_TRIM_FAILED 6    /* c172.jl:934-940 (warning + success=false)               */

/* engine discrete state, lib/FlightPhysics/src/piston.jl:199-203 */
#define FB_ENG_OFF 0
#define FB_ENG_STARTING 1
#define FB_ENG_RUNNING 2

/* ---- input fields (fb_set_input / fb_get_input / fb_schedule_inputs) --------------------
 * Values are doubles; Bool/enum inputs use 0/1/2.  Ranged inputs saturate on assignment
 * exactly like the reference's `Ranged` (lib/FlightPhysics/src/types.jl:15-20).            */
enum fb_input {
    /* C172S MechanicalActuation.u, c172s.jl:62-72 */
    FB_U_AILERON = 0, FB_U_ELEVATOR, FB_U_RUDDER, FB_U_AILERON_OFFSET, FB_U_ELEVATOR_OFFSET, FB_U_RUDDER_OFFSET,
    FB_U_FLAPS, FB_U_BRAKE_LEFT, FB_U_BRAKE_RIGHT,
    /* PistonEngine.u, piston.jl:259-267 (throttle is an input only for C172S; C172X drives it through its actuator) */
    FB_U_THROTTLE = 9, FB_U_MIXTURE, FB_U_ENG_START, FB_U_ENG_STOP, FB_U_MIXTURE_AUTO,
    FB_U_STEERING_ENGAGED = 14, /* landinggear.jl:45-48 */
    /* Payload.u, c172.jl:521-527 */
    FB_U_M_PILOT = 15, FB_U_M_COPILOT, FB_U_M_LPASS, FB_U_M_RPASS, FB_U_M_BAGGAGE,
    /* SimpleAtmosphere: TunableSeaLevel.u (atmosphere.jl:75-78), TunableWind.u (:165) */
    FB_U_T_SL = 20, FB_U_P_SL, FB_U_WIND_N, FB_U_WIND_E, FB_U_WIND_D,
    /* HorizontalTerrain: elevation (terrain.jl:34-36), surface type u[] (:38) 0 dry 1 wet 2 icy */
    FB_U_TRN_ELEVATION = 25, FB_U_TRN_SURFACE,
    /* C172X FlyByWireActuation commands, c172x.jl:112-120 (written by the control laws for X1/X2) */
    FB_U_ACT_THROTTLE = 27, FB_U_ACT_AILERON, FB_U_ACT_ELEVATOR, FB_U_ACT_RUDDER, FB_U_ACT_FLAPS,
    FB_U_ACT_BRAKE_LEFT, FB_U_ACT_BRAKE_RIGHT,
    /* Propeller{<:VariablePitch}.u: Ranged(0, 0, 1) blade-pitch command, propellers.jl:385,395-398 (FB_C172XA2 only) */
    FB_U_PROP_PITCH = 34,
    /* C172XControl.ControlLawsLon.u, c172x_ctl.jl:226-237 */
    FB_U_LON_MODE_REQ = 40, FB_U_LON_THROTTLE_AXIS, FB_U_LON_THROTTLE_OFFSET, FB_U_LON_ELEVATOR_AXIS, FB_U_LON_ELEVATOR_OFFSET,
    FB_U_LON_Q_REF, FB_U_LON_THETA_REF, FB_U_LON_EAS_REF, FB_U_LON_CLM_REF, FB_U_LON_H_REF,
    /* C172XControl.ControlLawsLat.u, c172x_ctl.jl:826-836 */
    FB_U_LAT_MODE_REQ = 60, FB_U_LAT_AILERON_AXIS, FB_U_LAT_AILERON_OFFSET, FB_U_LAT_RUDDER_AXIS, FB_U_LAT_RUDDER_OFFSET,
    FB_U_LAT_P_REF, FB_U_LAT_BETA_REF, FB_U_LAT_PHI_REF, FB_U_LAT_CHI_REF,
    /* C172XGuidance.GuidanceLaws.u.mode_req, c172x_gdc.jl:281-283 (0 = direct, 1 = segment), and SegmentGuidance.u
     * (c172x_gdc.jl:209-213): the two guidance requests and the target Segment's end points p1, p2 as
     * Geographic{LatLon,Ellipsoidal} (:38-41); default target Segment() = (0, 0, 0) -> (lat 1e-3, 0, 0) (:90) */
    FB_U_GDC_MODE_REQ = 70, FB_U_GDC_HOR_REQ, FB_U_GDC_VRT_REQ,
    FB_U_GDC_LAT1, FB_U_GDC_LON1, FB_U_GDC_H1, FB_U_GDC_LAT2, FB_U_GDC_LON2, FB_U_GDC_H2,
    FB_U__COUNT = 80
};

/* ---- output fields (fb_get_output / fb_log_* / fb_stats) --------------------------------
 * Scalars of the root `y` the reference saves each step (sim.jl:345-347); names follow the
 * reference's dotted path under `y.aircraft.vehicle`.                                     */
enum fb_output {
    FB_Y_PSI = 0, FB_Y_THETA, FB_Y_PHI,            /* kinematics.e_nb                       */
    FB_Y_LAT, FB_Y_LON, FB_Y_H_E, FB_Y_H_O,        /* kinematics.ϕ_λ, h_e, h_o              */
    FB_Y_V_N, FB_Y_V_E, FB_Y_V_D,                  /* kinematics.v_eb_n                     */
    FB_Y_V_GND, FB_Y_CHI_GND, FB_Y_GAMMA_GND,
    FB_Y_P, FB_Y_Q, FB_Y_R,                        /* kinematics.ω_wb_b                     */
    FB_Y_TAS, FB_Y_EAS, FB_Y_CAS, FB_Y_MACH, FB_Y_QBAR, /* airflow                          */
    FB_Y_ALPHA, FB_Y_BETA, FB_Y_ALPHA_FILT, FB_Y_BETA_FILT, FB_Y_STALL, FB_Y_C_L, FB_Y_C_D, /* systems.aero */
    FB_Y_WOW_LEFT, FB_Y_WOW_RIGHT, FB_Y_WOW_NOSE,  /* systems.ldg.*.strut.wow               */
    FB_Y_XI_LEFT, FB_Y_XI_RIGHT, FB_Y_XI_NOSE,     /* systems.ldg.*.strut.ξ                 */
    FB_Y_FDMP_LEFT, FB_Y_FDMP_RIGHT, FB_Y_FDMP_NOSE,
    FB_Y_ENG_OMEGA, FB_Y_ENG_STATE, FB_Y_ENG_MAP, FB_Y_ENG_P_SHAFT, FB_Y_ENG_TAU_SHAFT, FB_Y_ENG_MDOT, /* systems.pwp.engine */
    FB_Y_FUEL_M_TOTAL,                             /* systems.fuel.m_total                  */
    FB_Y_PROP_THRUST, FB_Y_PROP_J,                 /* systems.pwp.propeller.wr_p.F[1], J    */
    FB_Y_MASS,                                     /* dynamics.mp_Σ_b.m                     */
    FB_Y_F_X, FB_Y_F_Y, FB_Y_F_Z,                  /* dynamics.f_c_c (specific force)       */
    FB_Y_LON_MODE, FB_Y_LAT_MODE,                  /* avionics.ctl.lon.mode / lat.mode      */
    FB_Y_THROTTLE_CMD, FB_Y_ELEVATOR_CMD, FB_Y_AILERON_CMD, FB_Y_RUDDER_CMD,
    FB_Y_ALT_STATE,                                /* avionics.ctl.lon.s.h_state: 0 acquire, 1 hold (c172x_ctl.jl:44-47) */
    /* avionics.gdc (Cessna172Xv2): GuidanceLawsY.mode and SegmentGuidanceY / SegmentGuidanceData of the last guidance
     * update (c172x_gdc.jl:101-111,215-223,285-289) */
    FB_Y_GDC_MODE, FB_Y_GDC_HOR, FB_Y_GDC_VRT, FB_Y_GDC_DCHI, FB_Y_GDC_CHI_REF, FB_Y_GDC_H_REF,
    FB_Y_GDC_CHI_12, FB_Y_GDC_GAMMA_12, FB_Y_GDC_S_12, FB_Y_GDC_S_1B, FB_Y_GDC_E_SB, FB_Y_GDC_V_SB,
    FB_Y_PROP_DBETA,                               /* systems.pwp.propeller.Δβ (propellers.jl:372; 0 for FixedPitch) */
    FB_Y__COUNT
};

typedef struct fb_sim* fb_handle;

/* Simulation(mdl; dt, Δt, t_start, t_end, saveat) — lib/FlightCore/src/sim.jl:183-196 */
typedef struct fb_config {
    int32_t variant;        /* FB_C172*                                             */
    int32_t kinematics;     /* FB_WA / FB_ECEF / FB_NED                             */
    int64_t n_members;      /* members held by THIS handle (one handle per GPU)     */
    int64_t member_offset;  /* global index of this handle's member 0 in a sharded ensemble (>= 0; set by fb_group_create, informational elsewhere) */
    double dt;              /* RK4 step (reference default 0.02)                    */
    double Dt;              /* periodic (avionics) interval; must be k*dt, k>=1     */
    double t_start, t_end;
    int32_t device;         /* CUDA device ordinal                                  */
    int32_t log_every;      /* save outputs every `log_every` steps; 0 = no log (save_on=false) */
    int32_t log_n_fields;   /* number of entries of log_fields                      */
    int32_t log_fields[32]; /* fb_output ids to log                                 */
    int32_t steps_per_launch; /* K: RK4 steps fused per kernel launch (0 = default) */
    int32_t log_flags;      /* FB_LOG_* bits; 0 = FB_LOG_MEMBERS | FB_LOG_MOMENTS                       */
    /* FB_C172XA2 only (ignored otherwise; all zero = the reference's defaults) */
    int32_t prop_pitch_n;   /* > 1: VariablePitch with this many blade-pitch knots (reference default 11, propellers.jl:351); 0 / 1: FixedPitch */
    int32_t reserved0;
    double prop_pitch_lo, prop_pitch_hi; /* rad: VariablePitch(lower, upper); must match the "prop.tables3" section of the table blob */
    double act2_omega_n;    /* Actuator2 natural frequency, rad/s (0 = 5 * 2 pi, c172x.jl:71) */
    double act2_zeta;       /* Actuator2 damping ratio (0 = 0.6)                              */
} fb_config;
/* what a log sample keeps: every member's value (TimeSeries(sim), sim.jl:644-704) and / or the ensemble moments of the
 * logged fields, reduced inside the integrator as the sample is taken (no reference counterpart: one Simulation = one aircraft) */
#define FB_LOG_MEMBERS 1
#define FB_LOG_MOMENTS 2

/* C172.TrimParameters — lib/FlightApps/src/c172/c172.jl:806-818, struct-of-arrays; NULL = default */
typedef struct fb_trim_params_soa {
    const double* lat;        /* rad */
    const double* lon;        /* rad */
    const double* h_e;        /* ellipsoidal altitude, m (default 1050) */
    const double* psi_nb;     /* heading */
    const double* EAS;        /* default 50 */
    const double* gamma_wb_n;
    const double* psi_wb_dot;
    const double* theta_wb_dot;
    const double* beta_a;
    const double* fuel_load;  /* default 0.5 */
    const double* mixture;    /* default 0.5 */
    const double* flaps;      /* default 0   */
    const double* payload[5]; /* pilot, copilot, lpass, rpass, baggage: default 75,75,0,0,50 */
} fb_trim_params_soa;

/* C172.Init(KinInit(...), SystemsInitializer(...)) — c172.jl:773-794, kinematics.jl:21-40 */
typedef struct fb_init_soa {
    const double* q_nb[4];    /* attitude quaternion (real, i, j, k); NULL = identity */
    const double* lat; const double* lon;
    const double* h_e;        /* ellipsoidal altitude */
    const double* w_wb_b[3];
    const double* v_eb_n[3];
    const double* n_eng;      /* normalised engine speed */
    const double* engine_state; /* FB_ENG_* as double */
    const double* throttle; const double* mixture; const double* elevator; const double* aileron; const double* rudder;
    const double* flaps; const double* brake_left; const double* brake_right; const double* fuel_load;
    const double* payload[5];
    const double* stall; const double* alpha_filt; const double* beta_filt;
} fb_init_soa;

typedef struct fb_moments {  /* ensemble statistics over members with status RUNNING or DONE */
    int64_t count; double sum, sum_sq, min, max;
} fb_moments;

/* ---- lifecycle ----------------------------------------------------------------------- */
int32_t fb_api_version(void);
const char* fb_last_error(void);
/* Simulation(world; ...) : sim.jl:183-253 */
int32_t fb_create(const fb_config* cfg, fb_handle* out);
/* table blob produced by tools/build_tables.py (aero/engine/propeller/geoid/gain schedules) */
int32_t fb_load_tables(fb_handle h, const void* blob, size_t nbytes);
int32_t fb_destroy(fb_handle h);
int32_t fb_sync(fb_handle h);

/* ---- initialisation -------------------------------------------------------------------- */
/* init!(sim, C172.TrimParameters(...)) : sim.jl:390-414 -> c172.jl:883-942. trim_state_out [7][N] and
 * cost_out [N] may be NULL. Members whose trim does not reach cost <= 1e-16 get FB_ST_TRIM_FAILED and are frozen:
 * a deliberate deviation from the reference, which only warns, assigns the best trim state it found and keeps
 * simulating (c172.jl:934-940). In an ensemble a member that never reached its trim point is excluded from stepping and
 * from every moment instead of flying an untrimmed transient; the caller sees it in fb_status. The solver also differs
 * (damped Newton to 1e-26 here, NLopt BOBYQA stopping at 1e-16 there): trim states agree with the reference's stored
 * trims to ~1e-7 (tests/golden/trim_points.json), which is the optimiser's own stopping distance. */
int32_t fb_init_trim(fb_handle h, const fb_trim_params_soa* p, double* trim_state_out, double* cost_out);
/* The reference's own handling of a failed trim (c172.jl:934-940: warn, keep the best trim state found, go on simulating):
 * every member with FB_ST_TRIM_FAILED goes back to FB_ST_RUNNING from the state the solver left it in. n_released may be NULL. */
int32_t fb_release_trim_failed(fb_handle h, int64_t* n_released);
/* init!(sim, C172.Init(kin, sys)) : aircraftbase.jl:76-86 */
int32_t fb_init(fb_handle h, const fb_init_soa* ic);
/* raw state injection (mdl.x, aero.s.stall, engine.s[]) followed by the reinit! RHS evaluation */
int32_t fb_init_state(fb_handle h, const double* x, const uint8_t* stall, const uint8_t* eng_state);

/* ---- inputs ---------------------------------------------------------------------------- */
/* mdl.<...>.u.<field> = value : stride 0 broadcasts values[0], stride 1 = one value per member */
int32_t fb_set_input(fb_handle h, int32_t field, const double* values, int64_t stride);
int32_t fb_get_input(fb_handle h, int32_t field, double* out);
/* declarative replacement of user_callback! (sim.jl:331-341): event k sets `field[k]` of `member[k]` to
 * `value[k]` in the user-callback slot at the end of the step that reaches step index `step[k]`
 * (i.e. it is in force from t = t_start + step*dt). Replaces any previous schedule. */
int32_t fb_schedule_inputs(fb_handle h, int64_t n_events, const int64_t* member, const int64_t* step,
                           const int32_t* field, const double* value);

/* ---- stepping ---------------------------------------------------------------------------- */
/* step!(sim, Δt, true) : sim.jl:386 — advance every running member by round(t_advance/dt) steps. Asynchronous. Members near
 * the terrain may be advanced by a pass that runs on a second, internal stream beside the next launch; every
 * entry point other than fb_step / fb_run / fb_time / fb_nx first joins the two streams, so no caller ever sees an ensemble
 * at two step indices. */
int32_t fb_step(fb_handle h, double t_advance);
/* run!(sim) headless : sim.jl:611-638 — step until t_end */
int32_t fb_run(fb_handle h);
int32_t fb_time(fb_handle h, double* t, int64_t* n_steps);

/* ---- read-back ------------------------------------------------------------------------- */
int32_t fb_nx(fb_handle h, int32_t* nx);
int32_t fb_get_state(fb_handle h, double* x);                       /* sim.x : [nx][N] */
/* Pipelined read of sim.x (no reference counterpart: the reference reads its own memory). The state as of the work
 * submitted so far is snapshotted on the device and copied to `x` (pinned host memory, or the copy is not
 * asynchronous) on a stream of its own while later fb_step calls already run; `x` is valid after fb_wait_state (or any
 * later fb_get_state_async / fb_sync / fb_destroy on the handle). One read may be in flight per handle: a second
 * call first waits for the previous one. */
int32_t fb_get_state_async(fb_handle h, double* x);
int32_t fb_wait_state(fb_handle h);
int32_t fb_get_discrete(fb_handle h, uint8_t* stall, uint8_t* eng_state);
int32_t fb_get_output(fb_handle h, int32_t field, double* out);     /* sim.y.<field> : [N] */
int32_t fb_status(fb_handle h, int32_t* status);
/* TimeSeries(sim) : sim.jl:644-704 — decimated SoA log */
int32_t fb_log_info(fb_handle h, int64_t* n_samples, int32_t* n_fields);
int32_t fb_log_times(fb_handle h, double* t);                        /* [n_samples] */
int32_t fb_log_fetch(fb_handle h, int32_t field, double* out);       /* [n_samples][N] */
/* moments of output `field` over the members with status RUNNING or DONE, at the current state (one output pass + one
 * warp-shuffle reduction kernel; blocking) */
int32_t fb_stats(fb_handle h, int32_t field, fb_moments* out);
/* Moments of a LOGGED field at log samples [first, first + count), accumulated by the integrator itself when each sample
 * was taken (warp reduction + one atomic set per warp; members frozen before the sample are not counted). After
 * fb_log_stats_allreduce they are the moments of the whole sharded ensemble. fb_log_stats_begin enqueues the device-to-
 * host copy on the read-back stream behind the work submitted so far and returns at once; fb_log_stats_end waits for it
 * and writes out[count] for one field; fb_log_stats = begin + end. */
int32_t fb_log_stats_begin(fb_handle h, int64_t first, int64_t count);
int32_t fb_log_stats_end(fb_handle h, int32_t field, fb_moments* out);
int32_t fb_log_stats(fb_handle h, int32_t field, int64_t first, int64_t count, fb_moments* out);

/* ---- sharded ensembles --------------------------------------------------------------------------
 * The path shards by member with no data-path collective; the one exchange is the all-reduce of the
 * ensemble moments, which the library owns (NCCL, loaded at run time; a side stream, never the launch stream).
 * One process per GPU: rank 0 calls fb_comm_unique_id, the host distributes the 128 bytes by any means (MPI, torch,
 * a file), every rank calls fb_comm_init on its handle. fb_log_stats_allreduce then combines, in place on the device,
 * the moments of every log sample completed since the previous call (SUM of count / sum / sum of squares, MIN, MAX)
 * over all ranks; it is asynchronous, and every rank must call it at the same simulation time. */
int32_t fb_comm_unique_id(void* id128);
int32_t fb_comm_init(fb_handle h, int32_t n_ranks, int32_t rank, const void* id128);
int32_t fb_log_stats_allreduce(fb_handle h);
/* device time spent in the all-reduces so far (CUDA events on the side stream; waits for them) and their number */
int32_t fb_comm_time(fb_handle h, double* ms_total, int64_t* n_calls);

/* One process driving several GPUs: Simulation(...) over `n_members` members sharded in contiguous blocks over the
 * devices devices[0..n_dev) (cfg->device and cfg->member_offset are ignored), with an in-library NCCL communicator for
 * the moments. Host arrays passed to the fb_group_* calls are GLOBAL ([C][n_members]); the group scatters / gathers
 * the shards. fb_group_shard returns the handle of shard k for everything else (per-shard fb_* calls take shard-local
 * arrays of `count` members starting at global member `offset`). */
typedef struct fb_group* fb_group_handle;
int32_t fb_group_create(const fb_config* cfg, const int32_t* devices, int32_t n_dev, fb_group_handle* out);
int32_t fb_group_destroy(fb_group_handle g);
int32_t fb_group_size(fb_group_handle g, int32_t* n_dev);
int32_t fb_group_shard(fb_group_handle g, int32_t k, fb_handle* h, int64_t* offset, int64_t* count);
int32_t fb_group_load_tables(fb_group_handle g, const void* blob, size_t nbytes);
int32_t fb_group_init_trim(fb_group_handle g, const fb_trim_params_soa* p, double* trim_state_out, double* cost_out);
int32_t fb_group_set_input(fb_group_handle g, int32_t field, const double* values, int64_t stride);
int32_t fb_group_step(fb_group_handle g, double t_advance);   /* every shard advances concurrently */
int32_t fb_group_run(fb_group_handle g);
int32_t fb_group_sync(fb_group_handle g);
int32_t fb_group_get_state(fb_group_handle g, double* x);      /* [nx][n_members] */
int32_t fb_group_status(fb_group_handle g, int32_t* status);   /* [n_members] */
/* all-reduce of the samples completed so far + fetch of samples [first, first + count) of one logged field */
int32_t fb_group_log_stats(fb_group_handle g, int32_t field, int64_t first, int64_t count, fb_moments* out);

/* ---- linearisation -------------------------------------------------------------------------
 * linearize(aircraft, trim_params) : lib/FlightPhysics/src/aircraftbase.jl:292-341 -> linearization.jl:55-111, for every
 * member about its current (trimmed) state and inputs with the wind set to zero (the reference linearises under NoWind()).
 * NED kinematics only. dims = {nx, nu, ny} = {16, 4, 33} (C172S, c172s.jl:269-412) or {20, 4, 38} (C172X, c172x.jl:332-498).
 * Forward-difference Jacobians with FiniteDiff's default step max(sqrt(eps)|x_j|, sqrt(eps)) (linearization.jl:4,104-107).
 * Layout, member fastest: A is [nx cols][nx rows][N] (a Julia Array{Float64,3}(undef, N, nx, nx) indexed [i, row, col]),
 * B [nu][nx][N], C [nx][ny][N], D [nu][ny][N]; x0 [nx][N], u0 [nu][N], y0 [ny][N], xdot0 [nx][N]. Any pointer may be NULL. */
int32_t fb_linearize_dims(fb_handle h, int32_t dims[3]);
int32_t fb_linearize(fb_handle h, double* A, double* B, double* C, double* D, double* x0, double* u0, double* y0, double* xdot0);

/* ---- LinearizedSS ensembles ------------------------------------------------------------------
 * N linear state-space models xdot = xdot0 + A (x - x0) + B (u - u0), y = y0 + C (x - x0) + D (u - u0)
 * (lib/FlightPhysics/src/linearization.jl:24-111,157-192) stepped in lock-step by the same fixed-step RK4: the
 * `Simulation(Model(linearize(aircraft, trim)))` of the reference's README example 3 (README.md:142-204), per member.
 * Array layouts are exactly those fb_linearize writes. The state starts at x0 and the input at u0 (X(lss) = copy(x0),
 * U(lss) = copy(u0), linearization.jl:152-153). */
typedef struct fb_lss* fb_lss_handle;
int32_t fb_lss_create(int32_t device, int64_t n_members, int32_t nx, int32_t nu, int32_t ny, double dt, fb_lss_handle* out);
int32_t fb_lss_destroy(fb_lss_handle l);
int32_t fb_lss_set_model(fb_lss_handle l, const double* A, const double* B, const double* C, const double* D,
                         const double* x0, const double* u0, const double* y0, const double* xdot0);
/* LinearizedSS(linearize(aircraft, trim)) of every member of `sim`, device to device (same N, same device) */
int32_t fb_lss_from_linearize(fb_lss_handle l, fb_handle sim);
int32_t fb_lss_set_state(fb_lss_handle l, const double* x);   /* [nx][N] */
int32_t fb_lss_get_state(fb_lss_handle l, double* x);
int32_t fb_lss_set_input(fb_lss_handle l, const double* u);   /* [nu][N]: throttle, aileron, elevator, rudder commands */
int32_t fb_lss_get_input(fb_lss_handle l, double* u);
int32_t fb_lss_get_output(fb_lss_handle l, double* y);        /* [ny][N] */
int32_t fb_lss_step(fb_lss_handle l, double t_advance);       /* step!(sim, t_advance, true); asynchronous */
int32_t fb_lss_sync(fb_lss_handle l);
int32_t fb_lss_stream(fb_lss_handle l, void** stream);

/* ---- test hooks: the Modeling protocol pieces, for parity checks in isolation ---------- */
/* f_ode!(world) at the current state: writes ẋ [nx][N] (modeling.jl:201-254, world.jl:26-32) */
int32_t fb_rhs(fb_handle h, double* xdot);
/* f_step!(world) at the current state (world.jl:34-39): the y of an f_ode! there, then quaternion renormalisation, stall
 * hysteresis, crash checks, contact reset and the engine state machine. changed [N] (may be NULL): 1 where x or s changed. */
int32_t fb_fstep(fb_handle h, int32_t* changed);
/* f_periodic!(avionics, vehicle) + assign! at the current state (c172x2.jl:28-37): guidance, ControlLawsLon, ControlLawsLat
 * from the outputs of an f_ode! there; no-op for variants without avionics. */
int32_t fb_fperiodic(fb_handle h);
/* counters */
int32_t fb_counters(fb_handle h, int64_t* kernel_launches, int64_t* rhs_evals_per_member);
/* utility: measured FP64 FMA throughput of `device` in TFLOP/s (the roofline denominator of this path) */
int32_t fb_measure_fp64_peak(int32_t device, double* tflops);
/* device pointers for zero-copy interop with torch / CUDA.jl (valid until fb_destroy) */
int32_t fb_device_state(fb_handle h, void** d_x, int64_t* pitch_elems);
/* launch stream handle (cudaStream_t) so that callers can time with events on the right stream. To CONSUME the state on
 * this stream after stepping (CUDA.jl / torch interop), call fb_device_state again first: it is the join that orders the
 * internal hand-over stream in front of whatever the caller enqueues next. */
int32_t fb_stream(fb_handle h, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* FLIGHTBATCH_H */
