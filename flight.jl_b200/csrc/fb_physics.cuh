// fb_physics.cuh — one-thread-per-aircraft FP64 physics of the Cessna 172 hot path (sm_100a).
//
// Continuous right-hand side f_ode!(world) and discrete f_step!(world) of
//   Cessna172Sv0 / Cessna172Xv* on SimpleWorld(SimpleAtmosphere, HorizontalTerrain)
// Behavioural spec (reference file:line, paths relative to the reference root):
//   kinematics   lib/FlightPhysics/src/kinematics.jl:181-223 (WA), :363-409 (NED), :226-229 (renormalise)
//   geodesy      lib/FlightPhysics/src/geodesy.jl:62-69,103-106,125-129,204-211,367-428,451-467
//   atmosphere   lib/FlightPhysics/src/atmosphere.jl:116-135,220-242,269-283
//   aero         lib/FlightApps/src/c172/c172.jl:226-245,307-384
//   landing gear lib/FlightPhysics/src/landinggear.jl:228-347,411-483
//   propeller    lib/FlightPhysics/src/propellers.jl:405-452
//   engine       lib/FlightPhysics/src/piston.jl:314-477,575-603
//   fuel/payload lib/FlightApps/src/c172/c172.jl:542-554,607-641
//   dynamics     lib/FlightPhysics/src/dynamics.jl:443-499
//   actuation    lib/FlightApps/src/c172/c172s/c172s.jl:92-120, c172x/c172x.jl:39-52,127-143
//
// This is a B200-first restatement, not a transcription. The dataflow is rearranged so that one RHS needs
// 6 atan2 + 1 log + 1 exp instead of the reference's 21 atan2 + 8 pow + 5 sincos (results agree to rounding):
// frame rotations are built from direction cosines / half-angle identities instead of angle -> sincos round
// trips, the local vertical at the CoM is the n-vector itself, output-only quantities (CAS, total pressure,
// Euler angles, accelerations) are left to the output pass, mass properties are evaluated in their
// affine-in-mass form, divisions by table spacings and constants are multiplications by stored reciprocals,
// the remaining divisions / square roots use MUFU-seeded Newton sequences without IEEE slow paths, and a
// conservative height gate skips the three strut evaluations while the aircraft is well clear of the terrain.
// Derivatives are handed to a `Sink` as they are produced so that the integrator can fold each one into
// its Runge-Kutta accumulators (kept in shared memory) instead of holding a derivative vector in registers.
#pragma once
// FB_LDG_INLINE: 1 in the translation unit of the complete integrator kernels (flightbatch_complete.cu), where the strut
// code is inlined once inside the leg loop (cfg 4: +30 %); 0 elsewhere, where the legs call the out-of-line copy. The
// split exists because the mere presence of the inlined copy in a module changes ptxas' register allocation of the
// airborne kernels compiled beside it (measured: 52.6 -> 56.4 ms on the 1M-member cfg-2 launch).
#ifndef FB_LDG_INLINE
#define FB_LDG_INLINE 0
#endif
#include "fb_math.cuh"
#include "fb_tables.h"
#include "../../include/flightbatch.h"

// ------------------------------------------------------------------------------------------- layout
// AIR = true is the integrator's compact layout for members clear of the terrain: the six friction-regulator states
// of the landing gear are identically zero there (landinggear.jl:479-483 resets them while a strut is unloaded), so
// they get no column; ROW maps a compact component to its row of the ensemble state in HBM.
template <int VAR, int KIN, bool AIR = false>
struct Lay {
    static constexpr int NLDG = AIR ? 0 : 6;
    static constexpr int AERO = 0, LDG = 2, FUEL = 2 + NLDG, ENG = FUEL + 1, ACT = ENG + 3;
    // actuator states: none (C172S), one per first-order unit (C172X), (v, p) per second-order unit (FB_C172XA2, c172x.jl:84)
    static constexpr int NACT = (VAR == FB_C172S) ? 0 : (VAR == FB_C172XA2 ? 14 : 7);
    // ... of which the integrator propagates in closed form (first-order lags only: no stage register, no accumulator column)
    static constexpr int NACT_CF = (VAR == FB_C172X0 || VAR == FB_C172X1) ? 7 : 0;
    static constexpr int KINX = ACT + NACT;
    __host__ __device__ static constexpr int row(int c) { return (AIR && c >= 2) ? c + 6 : c; }
    static constexpr int NKIN = (KIN == FB_WA) ? 9 : (KIN == FB_ECEF ? 8 : 6);
    static constexpr int DYN = KINX + NKIN;
    static constexpr int NX = DYN + 6;
};

namespace wgs {
constexpr double a = 6378137.0;
constexpr double f = 1.0 / 298.257223563;
constexpr double w_ie = 7.292115e-05;
constexpr double GM = 3.986005e+14;
constexpr double b = a * (1 - f);
constexpr double e2 = 2 * f - f * f;
constexpr double m = w_ie * w_ie * a * a * b / GM;
constexpr double g_a = 9.7803253359;
constexpr double g_b = 9.8321849378;
constexpr double k_g = b * g_b / (a * g_a) - 1;
}
namespace isa {
constexpr double R = 287.05287, gamma = 1.40, T_std = 288.15, p_std = 101325.0, g_std = 9.80665;
constexpr double rho_std = p_std / (R * T_std);
}

// flags a single RHS evaluation hands to f_step! and to the status logic
struct RhsOut {
    double alpha;       // unclamped aerodynamic AoA (aero.y.α, c172.jl:324,367)
    int wow;            // bit i = strut i weight-on-wheels
    int crash;          // 2 bits per strut: 0 none, 1 strut/terrain angle > 60 deg, 2 compression rate > 10 m/s (landinggear.jl:331-347)
    int err;            // 0 none, FB_ST_BOUNDS / FB_ST_CONTACT_ASSERT (would have thrown inside f_ode!), FB_NEED_COMPLETE
    int gnd;            // a strut was evaluated (the aircraft is within the height gate of the terrain)
};
#define FB_NEED_COMPLETE 64   // internal: an AIR-layout evaluation came within the height gate (or left the troposphere); the member goes to the complete kernel

// everything the output pass / log / control laws may ask for (filled only by the WY=true instantiation)
struct YFull {
    double psi, theta, phi, lat, lon, h_e, h_o;
    d3 v_eb_n, w_wb_b, w_eb_b;
    double v_gnd, chi_gnd, gamma_gnd;
    double TAS, EAS, CAS, M, q;
    double alpha, beta, alpha_filt, beta_filt, C_L, C_D;
    int wow;
    double xi[3], F_dmp[3];
    double eng_omega, eng_n, eng_MAP, eng_P_shaft, eng_tau_shaft, eng_mdot, fuel_m_total, prop_thrust, prop_J, mass;
    d3 f_c_c;
    double act_pos[7], act_cmd[7];
    double prop_dbeta;
};

static __device__ __noinline__ double pow_slow(double a, double b) { return pow(a, b); }
// pseudo input fields: combinations the RHS uses as one value (so that an input cache can hold them ready-made)
enum { FB_UX_AE = 100,   // -clamp(elevator_offset + elevator)   C172S MechanicalActuation, c172s.jl:92-120
       FB_UX_AA,         //  clamp(aileron_offset + aileron)
       FB_UX_RU,         //  clamp(rudder_offset + rudder)
       FB_UX_TRF,        //  (p_sl / p_std)^(-beta R / g): sea-level factor of T_ISA(p)/T_std (piston.jl:38), 1 for the standard day
       FB_UX_TRC,        //  FB_UX_TRF / T_sl: T_ISA(p)/T_std = T * FB_UX_TRC in the troposphere
       FB_UX_PL0,        //  payload moments (c172.jl:542-554: five point masses, all at z = 0.899): sum m
       FB_UX_PL1, FB_UX_PL2, FB_UX_PL3, FB_UX_PL4, FB_UX_PL5 };   // sum m x, sum m y, sum m x^2, sum m y^2, sum m x y
// The five payload stations share one height and two lateral offsets, so their contribution to the mass properties is
// six moments of the masses; the right-hand side adds them to the airframe's and the fuel's (affine in the masses).
FB_DEV void payload_moments(double m_pilot, double m_copilot, double m_lpass, double m_rpass, double m_baggage, double* S) {
    constexpr double xf = 0.183, xr = -0.681, xb = -1.316, y = 0.356;
    const double front = m_pilot + m_copilot, rear = m_lpass + m_rpass;
    const double dfront = m_copilot - m_pilot, drear = m_rpass - m_lpass;     // right minus left
    S[0] = front + rear + m_baggage;
    S[1] = fma(xf, front, fma(xr, rear, xb * m_baggage));
    S[2] = y * (dfront + drear);
    S[3] = fma(xf * xf, front, fma(xr * xr, rear, (xb * xb) * m_baggage));
    S[4] = (y * y) * (front + rear);
    S[5] = fma(xf * y, dfront, (xr * y) * drear);
}
struct MemberIn {       // per-member input view: field f of this member is u[f * pitch]
    double* u;          // plain (coherent) loads: the owning thread also writes inputs (events, control laws)
    long long pitch;
    FB_DEV double raw(int f) const { return u[(long long)f * pitch]; }
    FB_DEV double operator()(int f) const {
        if (f == FB_UX_AE) return -clampd(raw(FB_U_ELEVATOR_OFFSET) + raw(FB_U_ELEVATOR), -1.0, 1.0);
        if (f == FB_UX_AA) return clampd(raw(FB_U_AILERON_OFFSET) + raw(FB_U_AILERON), -1.0, 1.0);
        if (f == FB_UX_RU) return clampd(raw(FB_U_RUDDER_OFFSET) + raw(FB_U_RUDDER), -1.0, 1.0);
        if (f == FB_UX_TRF || f == FB_UX_TRC) {
            double p_sl = raw(FB_U_P_SL);
            double trf = (p_sl == 101325.0) ? 1.0 : pow_slow(p_sl * (1.0 / 101325.0), 6.5e-3 * 287.05287 / 9.80665);
            return (f == FB_UX_TRF) ? trf : trf * fb_rcp(raw(FB_U_T_SL));
        }
        if (f >= FB_UX_PL0 && f <= FB_UX_PL5) { double S[6]; payload(S); return S[f - FB_UX_PL0]; }
        return raw(f);
    }
    FB_DEV void payload(double* S) const {
        payload_moments(raw(FB_U_M_PILOT), raw(FB_U_M_COPILOT), raw(FB_U_M_LPASS), raw(FB_U_M_RPASS), raw(FB_U_M_BAGGAGE), S);
    }
    FB_DEV void set(int f, double v) const { u[(long long)f * pitch] = v; }
};

// Input view of the integrator: the values every RHS evaluation reads are kept in spare shared-memory columns
// (NSLOT of them, filled in priority order), Bool inputs in a flag word; everything else falls through to HBM/L2.
// The owning thread calls refresh() after anything wrote inputs (events, control laws).
template <int VAR, int T, int NSLOT>
struct CachedIn {
    MemberIn g;
    double* sc;         // shared: column 0 of this thread's slots
    int flags;          // bit 0 mixture_auto, 1 eng_start, 2 eng_stop, 3 any wind component non-zero
    // Slot priority. First what every evaluation reads whatever the model (sea-level temperature and the T_ISA factor,
    // terrain elevation, the surface / throttle / mixture commands), then the six payload moments (all or none: with
    // fewer than 15 slots they are formed from the five masses on every evaluation), the wind, the sea-level pressure
    // (read once per step where the ISA anchor is in use) and the brake commands (ground contact only).
    __host__ __device__ static constexpr int field_of(int s) {
        if (s < 0) return -1;
        if (s < 9) {
            if (VAR == FB_C172S) {
                switch (s) {
                    case 0: return FB_U_T_SL; case 1: return FB_UX_TRC; case 2: return FB_U_TRN_ELEVATION; case 3: return FB_UX_AE;
                    case 4: return FB_UX_AA; case 5: return FB_UX_RU; case 6: return FB_U_FLAPS; case 7: return FB_U_THROTTLE;
                    default: return FB_U_MIXTURE;
                }
            } else {
                switch (s) {
                    case 0: return FB_U_T_SL; case 1: return FB_UX_TRC; case 2: return FB_U_TRN_ELEVATION; case 3: return FB_U_ACT_THROTTLE;
                    case 4: return FB_U_ACT_AILERON; case 5: return FB_U_ACT_ELEVATOR; case 6: return FB_U_ACT_RUDDER; case 7: return FB_U_ACT_FLAPS;
                    default: return FB_U_MIXTURE;
                }
            }
        }
        int r = s - 9;
        if (NSLOT >= 15) { if (r < 6) return FB_UX_PL0 + r; r -= 6; }
        switch (r) {
            case 0: return FB_U_WIND_N; case 1: return FB_U_WIND_E; case 2: return FB_U_WIND_D; case 3: return FB_U_P_SL;
            case 4: return VAR == FB_C172S ? FB_U_BRAKE_LEFT : FB_U_ACT_BRAKE_LEFT; case 5: return VAR == FB_C172S ? FB_U_BRAKE_RIGHT : FB_U_ACT_BRAKE_RIGHT;
            default: return -1;
        }
    }
    static constexpr int NFIELDS = 9 + 6 + 6;     // upper bound of useful slots
    __host__ __device__ static constexpr int slot(int f) {
        for (int s = 0; s < NSLOT; s++) if (field_of(s) == f) return s;
        return -1;
    }
    FB_DEV double operator()(int f) const {
        const int s = slot(f);
        if (s >= 0) return sc[s * T];
        if (f == FB_U_MIXTURE_AUTO) return (flags & 1) ? 1.0 : 0.0;
        if (f == FB_U_ENG_START) return (flags & 2) ? 1.0 : 0.0;
        if (f == FB_U_ENG_STOP) return (flags & 4) ? 1.0 : 0.0;
        if (f >= FB_U_WIND_N && f <= FB_U_WIND_D) return (flags & 8) ? g(f) : 0.0;
        return g(f);
    }
    FB_DEV void payload(double* S) const {
        if (slot(FB_UX_PL5) >= 0) {
#pragma unroll
            for (int j = 0; j < 6; j++) S[j] = sc[slot(FB_UX_PL0 + j) * T];
        } else g.payload(S);
    }
    FB_DEV void set(int f, double v) {       // one plain field (not a flag, not a derived slot): ensemble + cached copy
        g.set(f, v);
        const int s = slot(f);
        if (s >= 0) sc[s * T] = v;
    }
    FB_DEV void refresh() {
        double S[6];
        if (slot(FB_UX_PL5) >= 0) g.payload(S);
#pragma unroll
        for (int s = 0; s < NSLOT; s++) {
            const int f = field_of(s);
            if (f >= FB_UX_PL0 && f <= FB_UX_PL5) sc[s * T] = S[f - FB_UX_PL0];
            else if (f >= 0) sc[s * T] = g(f);
        }
        flags = (g(FB_U_MIXTURE_AUTO) != 0.0 ? 1 : 0) | (g(FB_U_ENG_START) != 0.0 ? 2 : 0) | (g(FB_U_ENG_STOP) != 0.0 ? 4 : 0) |
                ((g(FB_U_WIND_N) != 0.0 || g(FB_U_WIND_E) != 0.0 || g(FB_U_WIND_D) != 0.0) ? 8 : 0);
    }
};

// Output tap of the lean RHS: the handful of intermediate values the control and guidance laws feed back (attitude
// quaternion, NED velocity, wander-frame body rates, TAS, density, alpha, beta, latitude, longitude). The integrator stores them when the evaluation is the
// one at x* (whose outputs the reference's f_periodic! sees), which saves a second, output-producing evaluation.
// In priority order for the shared-memory part (the first NS entries sit in spare columns of the owning thread, the rest in
// the member's column of E.ctl): what the scheduled gains and every law read first.
enum { TAP_TAS = 0, TAP_RHO = 1, TAP_ALPHA = 2, TAP_BETA = 3, TAP_W_WB_B = 4, TAP_V_EB_N = 7, TAP_Q_NB = 10, TAP_LAT = 14, TAP_LON = 15, TAP_COUNT = 16 };
struct NoTap { FB_DEV void operator()(int, double) const {} };
// Tap split between shared-memory columns (entries < NS, compile-time stride T) and the member's column of a
// [TAP_COUNT][pitch] array; written only while `on`. With ANC the latitude / longitude entries are the anchor's (below).
template <int T, int NS, bool ANC>
struct SplitTap {
    double* ps;                      // shared: this thread's tap columns
    double* gbase;                   // row 0 of the [TAP_COUNT][pitch] array (member 0)
    const volatile long long* i0;    // shared: member index of lane 0 of this warp's task; the address of the member's column is
    int lane;                        //   formed where a value is stored (a 64-bit pointer kept across the right-hand side would be spilled)
    long long pitch; bool on;
    FB_DEV void operator()(int k, double v) const {
        if (ANC && k >= TAP_LAT) return;
        if (on) { if (k < NS) ps[k * T] = v; else gbase[(long long)k * pitch + *i0 + lane] = v; }
    }
};

// Anchor of the slowly varying transcendentals. Within one Runge-Kutta step the position moves by |v| dt (metres), so
// the geodetic latitude / longitude and the ISA pressure at the stage states differ from their values at the step-start
// state by ~1e-7 relative. The evaluation at the step-start state (k1, or f(x*) doubling as the next k1) computes them
// exactly (two atan2, exp o log) and leaves an anchor in five shared-memory columns of the owning thread; the other three
// evaluations of the step update from it:
//   lat = lat0 + asin(sin lat cos lat0 - cos lat sin lat0),  lon likewise from the horizontal components of n_e, with
//   asin(s) = s + s^3/6 (|s| < 1e-4 enforced: the next term is 3 s^5 / 40 < 1e-21);
//   p = p0 (T / T0)^k = p0 (1 + e)^k, e = T / T0 - 1, as the binomial series to e^4 (|e| < 2e-4 enforced: the e^5 term is
//   < 6e-19);
// outside those bounds (never at dt = 0.01) the evaluation falls back to the exact formulas. The anchor state itself is
// not stored: it is the step-start column of the integrator (n0 is rebuilt from it), which no stage writes.
enum { ANC_LAT = 0, ANC_LON, ANC_ICLAT, ANC_P, ANC_IT, ANC_COUNT };
struct NoAnchor {
    static constexpr bool ON = false;
    FB_DEV bool exact() const { return true; }
    FB_DEV double get(int) const { return 0.0; }
    FB_DEV void set(int, double) const {}
    FB_DEV d3 n0() const { return mk3(0.0, 0.0, 0.0); }
};
template <int T, int KIN, int KINX>
struct SmemAnchor {
    static constexpr bool ON = true;
    double* p;              // this thread's anchor columns
    const double* sx;       // this thread's step-start state column
    bool ex;                // this evaluation is at the step-start state
    FB_DEV bool exact() const { return ex; }
    FB_DEV double get(int k) const { return p[k * T]; }
    FB_DEV void set(int k, double v) const { p[k * T] = v; }
    FB_DEV d3 n0() const {  // n_e of the step-start state
        if (KIN == FB_WA) {
            const double w = sx[(KINX + 4) * T], x = sx[(KINX + 5) * T], y = sx[(KINX + 6) * T], z = sx[(KINX + 7) * T];
            return mk3(-(2.0 * x * z + 2.0 * w * y), -(2.0 * y * z - 2.0 * w * x), -(1.0 - 2.0 * (x * x + y * y)));
        }
        return mk3(sx[(KINX + 4) * T], sx[(KINX + 5) * T], sx[(KINX + 6) * T]);
    }
};

// derivative sink that simply stores into a local array (test hook, trim, output pass)
struct ArraySink {
    double* xd;
    FB_DEV void put(int c, double v) { xd[c] = v; }
    // first-order actuator i at row c (c172x.jl:39-52): state as stored, derivative (cmd - x) / tau
    FB_DEV double act(int, int c, const double* x, double) const { return x[c]; }
    FB_DEV void put_act(int c, double cmd, double xp) { xd[c] = (1.0 / 0.05) * (cmd - xp); }
};
// state holder over a plain array (f_step! on register / local state)
struct ArrayState {
    double* x;
    FB_DEV double get(int c) const { return x[c]; }
    FB_DEV void set(int c, double v) { x[c] = v; }
};

// ------------------------------------------------------------------------------------ interpolation
// cell index / weight on explicit knots (Gridded(Linear())); x already clamped to the knot span; IK = 1/spacing.
// Branch-free: the span is cut into FB_BK uniform buckets with at most one interior knot each (fb_tables.h: TB_BK),
// so the cell is the one at the bucket's left edge, plus one if x is at or beyond the knot inside the bucket.
template <int AXIS>
FB_DEV void locate(const double* tb, int KN, int IK, double x, int& i, double& w) {
    double t = (x - tb[KN]) * tb[TB_BK_IH + AXIS];
    int b = min(max((int)t, 0), FB_BK - 1);
    const int4 r = *reinterpret_cast<const int4*>(tb + TB_BK + (AXIS * FB_BK + b) * 2);
    double kb = __hiloint2double(r.y, r.x);
    int c = r.z + ((x >= kb) ? 1 : 0);
    i = c;
    w = (x - tb[KN + c]) * tb[IK + c];
}
FB_DEV double lerp(double a, double b, double w) { return fma(w, b - a, a); }
// uniform axis k0 + i/idk, n knots; Flat
FB_DEV void locate_uniform(double k0, double idk, int n, double x, int& i, double& w) {
    double t = clampd((x - k0) * idk, 0.0, (double)(n - 1));
    i = min((int)t, n - 2);
    w = t - (double)i;
}

// geoid undulation, bilinear on the EGM96 15' grid (geodesy.jl:186-211); geoid[lon][lat] float32
FB_DEV double geoid_height(const float* __restrict__ g, double lat, double lon) {
    double l = (lon < 0.0) ? lon + 2.0 * FB_PI : lon;               // mod(lon + 2pi, 2pi) for lon in [-pi, pi]
    double fi = (lat + 0.5 * FB_PI) * (720.0 / FB_PI);
    double fj = l * (1440.0 / (2.0 * FB_PI));
    int i = min(max((int)fi, 0), FB_GEOID_NLAT - 2);
    int j = min(max((int)fj, 0), FB_GEOID_NLON - 2);
    double u = fi - (double)i, w = fj - (double)j;
    const float* p = g + (size_t)j * FB_GEOID_NLAT + i;
    double f00 = (double)__ldg(p), f10 = (double)__ldg(p + 1);
    double f01 = (double)__ldg(p + FB_GEOID_NLAT), f11 = (double)__ldg(p + FB_GEOID_NLAT + 1);
    return lerp(lerp(f00, f10, u), lerp(f01, f11, u), w);
}
FB_DEV double geoid_at_n(const float* __restrict__ g, d3 n) {
    double lat = fb_atan2(n.z, fb_sqrt(fb_max(fma(n.x, n.x, n.y * n.y), 1e-280)));
    double lon = fb_atan2(n.y, n.x);
    return geoid_height(g, lat, lon);
}

// ------------------------------------------------------------------------------------------ geodesy
FB_DEV void radii(double n3, double& M, double& N) {   // geodesy.jl:125-129
    double ifd = fb_rsqrt(fma(-wgs::e2 * n3, n3, 1.0));
    N = wgs::a * ifd;
    M = N * (1 - wgs::e2) * ifd * ifd;
}
FB_DEV d3 cartesian_from_geo(d3 n, double h) {
    double M, N;
    radii(n.z, M, N);
    return mk3((N + h) * n.x, (N + h) * n.y, fma(N, 1 - wgs::e2, h) * n.z);
}
// ECEF -> (n-vector, ellipsoidal altitude): Fukushima's one-step Halley method (geodesy.jl:367-412)
FB_DEV void geo_from_cartesian(d3 r, d3& n, double& h) {
    constexpr double c = wgs::a * wgs::e2;
    constexpr double ec2 = 1 - wgs::e2;
    constexpr double ec = 0.99664718933525254;   // sqrt(1 - e2)
    double p2 = fma(r.x, r.x, r.y * r.y);
    double ip = fb_rsqrt(fb_max(p2, 1e-280));     // on the polar axis p = 0 and the longitude is taken as 0
    double p = p2 * ip;
    double az = fabs(r.z);
    double zc = ec * az, s0 = az, c0 = ec * p;
    double a02 = fma(s0, s0, c0 * c0);
    double a0 = fb_sqrt(a02);
    double a03 = a02 * a0;
    double sc = s0 * c0;
    double b0 = 1.5 * c * sc * (fma(p, s0, -zc * c0) * a0 - c * sc);
    double s1 = fma(zc, a03, c * s0 * s0 * s0) * a03 - b0 * s0;
    double c1 = fma(p, a03, -c * c0 * c0 * c0) * a03 - b0 * c0;
    double cc = ec * c1;
    double s12 = s1 * s1, cc2 = cc * cc;
    double inv = fb_rsqrt(s12 + cc2);
    h = (fma(p, cc, s0 * s1) - wgs::a * fb_sqrt(fma(ec2, s12, cc2))) * inv;
    double cl = cc * inv, sl = s1 * inv;            // cos(lat), |sin(lat)|
    sl = (r.z > 0.0) ? sl : ((r.z < 0.0) ? -sl : 0.0);
    double clon = (p2 > 0.0) ? r.x * ip : 1.0, slon = r.y * ip;
    d3 v = mk3(cl * clon, cl * slon, sl);
    double nn = fb_rsqrt(dot(v, v));                  // NVector constructor normalises (geodesy.jl:49-52)
    n = nn * v;
}
FB_DEV double gravity(double n3, double h) {  // Somigliana + 2nd-order altitude term (geodesy.jl:451-467)
    double s2 = n3 * n3;
    double g0 = wgs::g_a * fma(wgs::k_g, s2, 1.0) * fb_rsqrt(fma(-wgs::e2, s2, 1.0));
    return g0 * (1.0 - 2.0 / wgs::a * (1.0 + wgs::f + wgs::m - 2.0 * wgs::f * s2) * h + 3.0 / (wgs::a * wgs::a) * h * h);
}

// --------------------------------------------------------------------------------------- atmosphere
struct Air { d3 v_wb_b; double T, p, rho, ia, iRT, TAS, iTAS, q, tr; };  // tr = T_ISA(p) / T_std (piston.jl:38)

static __device__ __noinline__ void isa_upper(double hg, double T_sl, double p_sl, double& T, double& p, int& err) {
    const double beta[7] = {-6.5e-3, 0, 1e-3, 2.8e-3, 0, -2.8e-3, -2e-3};
    const double hc[7] = {11000, 20000, 32000, 47000, 51000, 71000, 84852};
    double hb = 0, Tb = T_sl, pb = p_sl;
    for (int i = 0; i < 7; i++) {
        double hh = (hg < hc[i]) ? hg : hc[i];
        double Tn = Tb + beta[i] * (hh - hb);
        double pn = (beta[i] != 0.0) ? pb * pow(1.0 + beta[i] / Tb * (hh - hb), -isa::g_std / (beta[i] * isa::R))
                                     : pb * exp(-isa::g_std / (isa::R * Tb) * (hh - hb));
        if (hg < hc[i]) { T = Tn; p = pn; return; }
        hb = hc[i]; Tb = Tn; pb = pn;
    }
    err = FB_ST_BOUNDS; T = Tb; p = pb;
}
// ISA temperature / pressure at orthometric altitude h_o, plus tr = (p/p_std)^(-beta R/g) for the engine
// trf = (p_sl / p_std)^kT (input-derived, see FB_UX_TRF). TROPO_ONLY (the airborne integrator layout): above 11 km the
// member is handed to the complete kernel instead of carrying the upper-layer code.
template <bool TROPO_ONLY = false, class Anc = NoAnchor, class PSL>
FB_DEV void isa_data(double h_o, double T_sl, const PSL& p_sl_of, double trc, double& T, double& p, double& tr, int& err, const Anc& anc = Anc()) {
    constexpr double beta = -6.5e-3;
    constexpr double k = -isa::g_std / (beta * isa::R);     // 5.2558...
    if (h_o < -1000.0) err = FB_ST_BOUNDS;
    double hg = h_o * wgs::a * fb_rcp(wgs::a + h_o);
    if (TROPO_ONLY || hg < 11000.0) {
        T = fma(beta, hg, T_sl);
        tr = T * trc;                                       // (p/p_std)^kT = (p_sl/p_std)^kT * T / T_sl, since k * kT == 1
        bool ex = !Anc::ON || anc.exact();
        double e = 0.0;
        if (Anc::ON && !ex) { e = fma(T, anc.get(ANC_IT), -1.0); ex = !(fabs(e) < 2e-4); }
        if (ex) {
            double ratio = T * fb_rcp(T_sl);
            p = p_sl_of() * exp(k * log(ratio));
            if (Anc::ON) { anc.set(ANC_P, p); anc.set(ANC_IT, fb_rcp(T)); }
        } else {   // p0 (1 + e)^k
            constexpr double c2 = k * (k - 1) / 2, c3 = c2 * (k - 2) / 3, c4 = c3 * (k - 3) / 4;
            p = anc.get(ANC_P) * fma(e, fma(e, fma(e, fma(e, c4, c3), c2), k), 1.0);
        }
        if (TROPO_ONLY && !(hg < 11000.0)) err = FB_NEED_COMPLETE;
    } else {   // rare: copies keep the callee's by-reference outputs from pinning the caller's values in local memory
        constexpr double kT = -beta * isa::R / isa::g_std;      // 1 / k
        double T2, p2; int e2 = 0;
        isa_upper(hg, T_sl, p_sl_of(), T2, p2, e2);
        T = T2; p = p2; if (e2) err = e2;
        tr = pow_slow(p2 * (1.0 / isa::p_std), kT);
        // the anchor holds a tropospheric state: an evaluation above the tropopause invalidates it, so that the next evaluation
        // below 11 km (a later stage of the same step, or a descent) takes the exact branch and anchors afresh
        if (Anc::ON) anc.set(ANC_IT, 0.0);
    }
}

// ------------------------------------------------------------------------------ PI lane (control.jl:52-81)
FB_DEV double pi_lane(double input, double x_i, double k_p, double k_i, double k_l, double lo, double hi, double& xdot) {
    double out_free = fma(k_p, input, x_i);
    int sat = (out_free >= hi) - (out_free <= lo);
    bool halted = isign(input * (double)sat) > 0;
    xdot = (halted ? 0.0 : k_i * input) - k_l * x_i;
    return clampd(out_free, lo, hi);
}

FB_DEV double get_mu(double mu_s, double mu_d, double v) {  // landinggear.jl:170-174, v_s 0.005, v_d 0.01
    double k = clampd((fabs(v) - 0.005) * (1.0 / (0.01 - 0.005)), 0.0, 1.0);
    return k * mu_d + (1.0 - k) * mu_s;
}

// sin and cos of a small angle (|x| <= pi/6: truncation below 5e-20 and 2e-18), no argument reduction, no slow path
FB_DEV void sincos_small(double x, double& s, double& c) {
    const double x2 = x * x;
    double ps = fma(x2, -1.0 / 1307674368000.0, 1.0 / 6227020800.0);
    ps = fma(ps, x2, -1.0 / 39916800.0); ps = fma(ps, x2, 1.0 / 362880.0); ps = fma(ps, x2, -1.0 / 5040.0);
    ps = fma(ps, x2, 1.0 / 120.0); ps = fma(ps, x2, -1.0 / 6.0);
    s = fma(ps * x2, x, x);
    double pc = fma(x2, -1.0 / 87178291200.0, 1.0 / 479001600.0);
    pc = fma(pc, x2, -1.0 / 3628800.0); pc = fma(pc, x2, 1.0 / 40320.0); pc = fma(pc, x2, -1.0 / 720.0);
    pc = fma(pc, x2, 1.0 / 24.0); pc = fma(pc, x2, -0.5);
    c = fma(pc, x2, 1.0);
}

#ifndef FB_LDG_EXACT
#define FB_LDG_EXACT 1     // 1: strut attachment geodesy by ECEF -> geodetic conversions and ECEF vector differences, as the reference does; 0: to first order in |r_bs| / R (see ldg_unit)
#endif
#ifndef FB_COM_EXACT
#define FB_COM_EXACT 0     // 1: geodetic coordinates of the CoM by a second ECEF -> geodetic conversion, as the reference does (diagnostic)
#endif
// kinematic data shared by the subsystems of one RHS evaluation
struct Kin { q4 q_nb, q_eb; d3 n_e, r_eb_e, w_wb_b, w_eb_b, v_eb_b, v_eb_n; double h_e, h_o, lat, lon;
             double iMh, iNh, clat;      // 1 / (M + h), 1 / (N + h), cos(lat) = |(n_x, n_y)|
             double s_eb, s_ew; };       // norm defects |q_eb|^2 - 1 and |q_ew|^2 - 1 (0 where the quaternion is built from angles)

// one landing-gear unit (strut + contact). Adds its wrench to (F, tau) in body axes. Fast reciprocals / square roots
// where the operand is bounded away from zero (direction cosines of a strut that has not crashed, friction
// coefficients); IEEE operations where a wheel speed can pass through zero (castoring nose wheel). Inlined ONCE, inside the
// non-unrolled loop over the three legs of the RHS: an out-of-line copy would receive the kinematic data through
// local memory (~25 L2 round trips per leg), three inlined copies would triple the code the I-cache has to hold.
FB_DEV void ldg_unit(const Kin& k, const float* __restrict__ geoid, double elev, int surface,
                                      d3 r_bs_b, double k_s, double k_d_ext, double k_d_cmp,
                                      bool has_steering, bool steering_engaged, double steer_input, double brake,
                                      double x0, double x1, double& xd0, double& xd1, d3& F, d3& tau, int leg, RhsOut& out,
                                      double* xi_out, double* fdmp_out) {
    const d3 e3 = mk3(0.0, 0.0, 1.0);
#if FB_LDG_EXACT
    d3 ks_e = qrot(k.q_eb, e3);
    d3 r_bs_e = qrot(k.q_eb, r_bs_b);
    d3 r_es_e = k.r_eb_e + r_bs_e;
    d3 n_t; double he_w;
    geo_from_cartesian(r_es_e, n_t, he_w);
    if (he_w < -1000.0) out.err = FB_ST_BOUNDS;
    // Geoid undulation under the strut: the reference takes LatLon(n_t) (two atan2) for the table lookup; here latitude and
    // longitude of Os come from those of Ob to first order in |r_bs| / R, (dN, dE) = horizontal part of R_nb r_bs — the
    // second-order remainder (1e-14 rad) moves the bilinear undulation by < 1e-11 m, two orders below the rounding of he_w.
    // (he_w and the compression below keep the reference's own formulation: see FB_LDG_EXACT.)
    double he_t;
    {
        const d3 d_s = qrot(k.q_nb, r_bs_b);
        he_t = elev + geoid_height(geoid, fma(d_s.x, k.iMh, k.lat), fma(d_s.y * k.iNh, fb_rcp(fb_max(k.clat, 1e-9)), k.lon));
    }
#else
    // Geodetic coordinates of the strut attachment Os to first order in |r_bs| / R (as for the CoM in the dynamics
    // block): with (dN, dE, dD) = R_nb r_bs, lat_s = lat + dN / (M + h), lon_s = lon + dE / ((N + h) cos lat),
    // h_s = h_e - dD + (curvature term). The reference converts r_es_e = r_eb_e + R_eb r_bs with Fukushima's method
    // (landinggear.jl:245) and later differences two 6.4e6 m ECEF vectors for the compression (:262-270: r_et_e - r_es_e,
    // which is exactly -dh n_t, projected on the terrain normal: l = dh / (u_t . k_s) up to cos(1e-7) - 1); the algebraic
    // form carries less rounding noise than the reference's own (one ulp of 6.4e6 m is 9.3e-10 m), i.e. it differs from
    // it by the reference's rounding spread, which is what tests/test_gpu_ground.py bounds.
    const d3 d_s = qrot(k.q_nb, r_bs_b);
    const double thN = d_s.x * k.iMh, thE = d_s.y * k.iNh;
    const double he_w = k.h_e - d_s.z + 0.5 * fma(thN, d_s.x, thE * d_s.y);
    if (he_w < -1000.0) out.err = FB_ST_BOUNDS;
    double he_t = elev + geoid_height(geoid, k.lat + thN, fma(thE, fb_rcp(fb_max(k.clat, 1e-9)), k.lon));
#endif
    if (he_t < -1000.0 || elev < -1000.0) out.err = FB_ST_BOUNDS;
    double dh = he_w - he_t;
    bool wow = dh <= 0.0;
    double vx = 0.0, vy = 0.0;
    double F_dmp = 0.0;
    q4 q_bc = mkq(1, 0, 0, 0);
    d3 r_bc_b = r_bs_b;
    if (wow) {
        out.wow |= (1 << leg);
#if FB_LDG_EXACT
        d3 r_et_e = cartesian_from_geo(n_t, he_t);
        d3 r_st_e = r_et_e - r_es_e;
        q4 q_en = qmul(k.q_eb, qconj(k.q_nb));          // product first, then rotate: keeps the reference's norm defect
        d3 ut_e = qrot(q_en, e3);                       // terrain normal (NED down) in ECEF: q_en(0,0,1)
        double ut_ks = dot(ut_e, ks_e);
        double l = dot(ut_e, r_st_e) * fb_rcp(ut_ks);
#else
        const double ut_ks = qrot(k.q_nb, e3).z;        // terrain normal (NED down at Ob) . strut axis
        const double l = dh * fb_rcp(ut_ks);
#endif
        // acos(ut_ks) in degrees > 60 (landinggear.jl:331-339) is ut_ks <= 1/2: acos(0.5) rounds to 60.00000000000001
        bool crash_angle = ut_ks <= 0.5;
        if (crash_angle) out.crash |= (1 << (2 * leg));
        double xi = fb_min(0.0, l);
        r_bc_b = mk3(r_bs_b.x, r_bs_b.y, r_bs_b.z + xi);
        d3 v_b = k.v_eb_b + cross(k.w_eb_b, r_bc_b);
        double cs = 1.0, sn = 0.0;  // wheel heading in strut axes
        if (has_steering) {
            if (steering_engaged) sincos_small(steer_input * (FB_PI / 6), sn, cs);   // |angle| <= pi/6: the input is Ranged(-1, 1)
            else { double r = sqrt(fma(v_b.x, v_b.x, v_b.y * v_b.y)); if (r > 0.0) { cs = v_b.x / r; sn = v_b.y / r; } }
        }
        // q_nw = q_nb o Rz(psi_sw) applied to e1 (landinggear.jl:291-296); Rz(psi) built from (cos, sin) by half angles
        double chw, shw;
        half_angle_cs(cs, sn, chw, shw);
        d3 iw_n = qrot(qmul(k.q_nb, mkq(chw, 0.0, 0.0, shw)), mk3(1.0, 0.0, 0.0));
        double ch, sh;
        half_angle(iw_n.y, iw_n.x, ch, sh);              // contact frame = NED rotated about down by azimuth(iw_n)
        q4 q_nc = mkq(ch, 0.0, 0.0, sh);
        q_bc = qmul(qconj(k.q_nb), q_nc);
        d3 v_c_body = qrot_inv(q_bc, v_b);
        d3 ks_c = qrot_inv(q_bc, e3);
        double xi_dot = -v_c_body.z * fb_rcp(ks_c.z);
        if (-xi_dot > 10.0 && !crash_angle) out.crash |= (2 << (2 * leg));
        double k_d = (xi_dot > 0.0) ? k_d_ext : k_d_cmp;
        F_dmp = -(k_s * xi + k_d * xi_dot);
        double vz = fma(ks_c.z, xi_dot, v_c_body.z);
        if (!(fabs(vz) < 1e-8)) out.err = FB_ST_CONTACT_ASSERT;
        vx = fma(ks_c.x, xi_dot, v_c_body.x);
        vy = fma(ks_c.y, xi_dot, v_c_body.y);
    }
    if (xi_out) { *xi_out = wow ? (r_bc_b.z - r_bs_b.z) : 0.0; *fdmp_out = F_dmp; }
    // friction regulator (landinggear.jl:401-409,417-419): k_p 5, k_i 400, k_l 0.2, bounds +-1
    double o0 = pi_lane(-vx, x0, 5.0, 400.0, 0.2, -1.0, 1.0, xd0);
    double o1 = pi_lane(-vy, x1, 5.0, 400.0, 0.2, -1.0, 1.0, xd1);
    if (wow) {
        double norm_v = fb_sqrt(fb_max(fma(vx, vx, vy * vy), 1e-280));   // 1e-140 for a wheel at rest: same mu, same psi_cv branch
        double mu_roll = get_mu(0.03, 0.02, norm_v);
        double mu_skid = (surface == 0) ? get_mu(0.75, 0.25, norm_v) : ((surface == 1) ? get_mu(0.25, 0.15, norm_v) : get_mu(0.075, 0.025, norm_v));
        double mu_x = fma(mu_skid - mu_roll, brake, mu_roll);
        double psi_cv = (norm_v < 1e-3) ? (FB_PI / 2) : fb_atan2(vy, vx);
        const double psi_skid = 10 * FB_PI / 180.0;
        double psi_abs = fabs(psi_cv);
        double mu_y;
        if (psi_abs < psi_skid) mu_y = mu_skid * psi_abs * (1.0 / psi_skid);
        else if (psi_abs > FB_PI - psi_skid) mu_y = mu_skid * (1.0 - (psi_skid + psi_abs - FB_PI) * (1.0 / psi_skid));
        else mu_y = mu_skid;
        double sc = fb_min(1.0, mu_skid * fb_rsqrt(fma(mu_x, mu_x, mu_y * mu_y)));   // mu_x >= mu_roll > 0
        d3 f_c = mk3(o0 * mu_x * sc, o1 * mu_y * sc, -1.0);
        d3 f_s = qrot(q_bc, f_c);
        double Nf = fb_max(0.0, -F_dmp * fb_rcp(f_s.z));
        d3 F_b = Nf * f_s;
        F = F + F_b;
        tau = tau + cross(r_bc_b, F_b);
    }
}

// Out-of-line copy for the output-producing RHS (log samples, test hooks): that function is called from inside the
// integrator kernels, and keeping it small keeps its register demand from leaking into their allocation.
static __device__ __noinline__ void ldg_unit_ool(const Kin& k, const float* __restrict__ geoid, double elev, int surface,
                                          d3 r_bs_b, double k_s, double k_d_ext, double k_d_cmp,
                                          bool has_steering, bool steering_engaged, double steer_input, double brake,
                                          double x0, double x1, double& xd0, double& xd1, d3& F, d3& tau, int leg, RhsOut& out,
                                          double* xi_out, double* fdmp_out) {
    ldg_unit(k, geoid, elev, surface, r_bs_b, k_s, k_d_ext, k_d_cmp, has_steering, steering_engaged, steer_input, brake, x0, x1, xd0, xd1,
             F, tau, leg, out, xi_out, fdmp_out);
}

// 2 x 9 table on uniform ("scaled") axes with Line extrapolation (piston.jl:84-104)
FB_DEV double interp_2x9_line(const double* y, double n0, double n1, double x0, double x1, double n, double x) {
    const double idn = 1.0 / (n1 - n0), idx = 8.0 / (x1 - x0);
    double ns = clampd(n, n0, n1), xs = clampd(x, x0, x1);
    double u = (ns - n0) * idn;
    double t = (xs - x0) * idx;
    int j = min((int)t, 7);
    double w = t - (double)j;
    double f00 = y[j], f01 = y[j + 1], f10 = y[9 + j], f11 = y[9 + j + 1];
    double v = lerp(lerp(f00, f01, w), lerp(f10, f11, w), u);
    double g0 = lerp(f10 - f00, f11 - f01, w) * idn;
    double g1 = lerp(f01 - f00, f11 - f10, u) * idx;
    return fma(n - ns, g0, fma(x - xs, g1, v));
}

// ------------------------------------------------------------------------------------------- the RHS
// x: state (read-only; every component is read before its derivative is put into the sink).
// Geodetic latitude / longitude and cos(lat) of the n-vector k.n_e (geodesy.jl:103-106), exactly at the step-start state
// and by the anchor's increment at the other stage states (see SmemAnchor).
template <class Anc>
FB_DEV void latlon(Kin& k, const Anc& anc) {
    const double p2 = fb_max(fma(k.n_e.x, k.n_e.x, k.n_e.y * k.n_e.y), 1e-280);
    const double ip = fb_rsqrt(p2);
    { const double c = p2 * ip; k.clat = fma(fma(-c, c, p2), 0.5 * ip, c); }   // sqrt(p2) to the last bit: one ulp of the latitude is 7e-10 m, which the
                                                                               // guidance laws differentiate (test_gpu_guidance.py)
    bool ex = !Anc::ON || anc.exact();
    if (Anc::ON && !ex) {
        const d3 n0 = anc.n0();
        const double ip0 = anc.get(ANC_ICLAT);
        const double clat0 = fma(n0.x, n0.x, n0.y * n0.y) * ip0;
        const double s = fma(k.n_e.z, clat0, -(k.clat * n0.z));                        // sin(lat - lat0)
        const double t = fma(k.n_e.y, n0.x, -(k.n_e.x * n0.y)) * (ip * ip0);           // sin(lon - lon0)
        k.lat = anc.get(ANC_LAT) + fma(s * s * (1.0 / 6.0), s, s);
        k.lon = anc.get(ANC_LON) + fma(t * t * (1.0 / 6.0), t, t);
        ex = !(fabs(s) < 1e-4 && fabs(t) < 1e-4 && p2 > 1e-6);
    }
    if (ex) {
        k.lat = fb_atan2(k.n_e.z, k.clat);
        k.lon = fb_atan2(k.n_e.y, k.n_e.x);
        if (Anc::ON && anc.exact()) { anc.set(ANC_LAT, k.lat); anc.set(ANC_LON, k.lon); anc.set(ANC_ICLAT, ip); }
    }
}
template <int VAR, int KIN, bool WY, bool AIR = false, class Sink, class U, class Tap = NoTap, class Anc = NoAnchor>
FB_DEV void rhs(const double* x, Sink& sink, int stall, int eng_state, const U& u,
                const double* __restrict__ tb, const float* __restrict__ geoid, RhsOut& out, YFull* yf, const Tap tap = Tap(), const Anc anc = Anc()) {
    typedef Lay<VAR, KIN, AIR> L;
    out.wow = 0; out.crash = 0; out.err = 0; out.gnd = 0;
    Kin k;
    k.w_eb_b = mk3(x[L::DYN], x[L::DYN + 1], x[L::DYN + 2]);
    k.v_eb_b = mk3(x[L::DYN + 3], x[L::DYN + 4], x[L::DYN + 5]);
    // ------------------------------------------------------------------ kinematics
    if (KIN == FB_WA) {
        q4 q_wb = mkq(x[L::KINX], x[L::KINX + 1], x[L::KINX + 2], x[L::KINX + 3]);
        q4 q_ew = mkq(x[L::KINX + 4], x[L::KINX + 5], x[L::KINX + 6], x[L::KINX + 7]);
        k.h_e = x[L::KINX + 8];
        k.s_ew = fma(q_ew.w, q_ew.w, fma(q_ew.x, q_ew.x, fma(q_ew.y, q_ew.y, q_ew.z * q_ew.z))) - 1.0;
        k.s_eb = fma(q_wb.w, q_wb.w, fma(q_wb.x, q_wb.x, fma(q_wb.y, q_wb.y, q_wb.z * q_wb.z))) - 1.0 + k.s_ew;
        double dq12 = 2.0 * q_ew.w * q_ew.x, dq13 = 2.0 * q_ew.w * q_ew.y;
        double dq24 = 2.0 * q_ew.x * q_ew.z, dq34 = 2.0 * q_ew.y * q_ew.z;
        // psi_nw = atan2(-(dq34+dq12), dq24-dq13); q_nw = Rz(psi_nw), built from (cos, sin) by half-angle identities
        double yy = -(dq34 + dq12), xx = dq24 - dq13;
        double irr = fb_rsqrt(fma(xx, xx, yy * yy));
        double c = xx * irr, s = yy * irr;
        double ch, sh;
        half_angle_cs(c, s, ch, sh);
        q4 q_nw = mkq(ch, 0.0, 0.0, sh);
        k.q_nb = qmul(q_nw, q_wb);
        k.q_eb = qmul(q_ew, q_wb);
        k.n_e = mk3(-(dq24 + dq13), -(dq34 - dq12), -(1.0 - 2.0 * (q_ew.x * q_ew.x + q_ew.y * q_ew.y)));
        latlon(k, anc);
        k.v_eb_n = qrot(k.q_nb, k.v_eb_b);
        double M, N;
        radii(k.n_e.z, M, N);
        k.r_eb_e = mk3((N + k.h_e) * k.n_e.x, (N + k.h_e) * k.n_e.y, fma(N, 1 - wgs::e2, k.h_e) * k.n_e.z);
        k.iNh = fb_rcp(N + k.h_e); k.iMh = fb_rcp(M + k.h_e);
        double wn_x = k.v_eb_n.y * k.iNh, wn_y = -k.v_eb_n.x * k.iMh;
        d3 w_ew_w = mk3(fma(c, wn_x, s * wn_y), fma(-s, wn_x, c * wn_y), 0.0);  // q_nw'(w_ew_n)
        d3 w_ew_b = qrot_inv(q_wb, w_ew_w);
        k.w_wb_b = k.w_eb_b - w_ew_b;
        q4 d1 = qdt(q_wb, k.w_wb_b), d2 = qdt(q_ew, w_ew_w);
        sink.put(L::KINX, d1.w); sink.put(L::KINX + 1, d1.x); sink.put(L::KINX + 2, d1.y); sink.put(L::KINX + 3, d1.z);
        sink.put(L::KINX + 4, d2.w); sink.put(L::KINX + 5, d2.x); sink.put(L::KINX + 6, d2.y); sink.put(L::KINX + 7, d2.z);
        sink.put(L::KINX + 8, -k.v_eb_n.z);
    } else if (KIN == FB_ECEF) {   // kinematics.jl:281-315: x = (q_eb, n_e, h_e), neither block normalised on the fly
        k.q_eb = mkq(x[L::KINX], x[L::KINX + 1], x[L::KINX + 2], x[L::KINX + 3]);
        k.n_e = mk3(x[L::KINX + 4], x[L::KINX + 5], x[L::KINX + 6]);
        k.h_e = x[L::KINX + 7];
        k.s_ew = 0.0;
        k.s_eb = fma(k.q_eb.w, k.q_eb.w, fma(k.q_eb.x, k.q_eb.x, fma(k.q_eb.y, k.q_eb.y, k.q_eb.z * k.q_eb.z))) - 1.0;
        // q_en = ltf(n_e) = Rz(lon) o Ry(-(lat + pi/2)) with (lat, lon) = LatLon(n_e): built from direction cosines by
        // half-angle identities. cos/sin of -(lat + pi/2) are (-sin lat, -cos lat).
        double p2 = fma(k.n_e.x, k.n_e.x, k.n_e.y * k.n_e.y);
        double ip = fb_rsqrt(fb_max(p2, 1e-280));
        double p = p2 * ip;
        double inn = fb_rsqrt(fma(k.n_e.z, k.n_e.z, p2));
        double clo = (p2 > 0.0) ? k.n_e.x * ip : 1.0, slo = k.n_e.y * ip;
        double chl, shl, cha, sha;
        half_angle_cs(clo, slo, chl, shl);
        half_angle_cs(-k.n_e.z * inn, -p * inn, cha, sha);
        q4 q_en = qmul(mkq(chl, 0.0, 0.0, shl), mkq(cha, 0.0, sha, 0.0));
        k.q_nb = qmul(qconj(q_en), k.q_eb);
        latlon(k, anc);
        k.v_eb_n = qrot(k.q_nb, k.v_eb_b);
        double M, N;
        radii(k.n_e.z, M, N);
        k.r_eb_e = mk3((N + k.h_e) * k.n_e.x, (N + k.h_e) * k.n_e.y, fma(N, 1 - wgs::e2, k.h_e) * k.n_e.z);
        k.iNh = fb_rcp(N + k.h_e); k.iMh = fb_rcp(M + k.h_e);
        double wx = k.v_eb_n.y * k.iNh, wy = -k.v_eb_n.x * k.iMh;
        k.w_wb_b = k.w_eb_b - qrot_inv(k.q_nb, mk3(wx, wy, 0.0));
        q4 d1 = qdt(k.q_eb, k.w_eb_b);
        d3 nd = qrot(q_en, mk3(-wy, wx, 0.0));                 // w_ew_n x (0, 0, -1)
        sink.put(L::KINX, d1.w); sink.put(L::KINX + 1, d1.x); sink.put(L::KINX + 2, d1.y); sink.put(L::KINX + 3, d1.z);
        sink.put(L::KINX + 4, nd.x); sink.put(L::KINX + 5, nd.y); sink.put(L::KINX + 6, nd.z);
        sink.put(L::KINX + 7, -k.v_eb_n.z);
    } else if (KIN == FB_NED) {
        double psi = x[L::KINX], th = x[L::KINX + 1], phi = x[L::KINX + 2];
        k.lat = x[L::KINX + 3]; k.lon = x[L::KINX + 4]; k.h_e = x[L::KINX + 5];
        k.s_ew = 0.0; k.s_eb = 0.0;
        if (fabs(k.lat) > 0.5 * FB_PI || fabs(k.lon) > FB_PI) out.err = FB_ST_BOUNDS;
        double sps, cps, sth, cth, sph, cph;
        sincos(0.5 * psi, &sps, &cps); sincos(0.5 * th, &sth, &cth); sincos(0.5 * phi, &sph, &cph);
        k.q_nb = qmul(qmul(mkq(cps, 0, 0, sps), mkq(cth, 0, sth, 0)), mkq(cph, sph, 0, 0));
        double sla, cla, slo, clo;
        sincos(k.lat, &sla, &cla); sincos(k.lon, &slo, &clo);
        k.n_e = mk3(cla * clo, cla * slo, sla);
        // q_en = Rz(lon) o Ry(-(lat + pi/2))
        double sl2, cl2, sa2, ca2;
        sincos(0.5 * k.lon, &sl2, &cl2); sincos(-0.5 * (k.lat + 0.5 * FB_PI), &sa2, &ca2);
        q4 q_en = qmul(mkq(cl2, 0, 0, sl2), mkq(ca2, 0, sa2, 0));
        k.q_eb = qmul(q_en, k.q_nb);
        k.v_eb_n = qrot(k.q_nb, k.v_eb_b);
        double M, N;
        radii(k.n_e.z, M, N);
        k.r_eb_e = mk3((N + k.h_e) * k.n_e.x, (N + k.h_e) * k.n_e.y, fma(N, 1 - wgs::e2, k.h_e) * k.n_e.z);
        double iN = fb_rcp(N + k.h_e);
        k.iNh = iN; k.iMh = fb_rcp(M + k.h_e); k.clat = cla;
        double wx = k.v_eb_n.y * iN, wy = -k.v_eb_n.x * k.iMh;
        double icla = fb_rcp(cla);
        d3 w_en_n = mk3(wx, wy, -k.v_eb_n.y * (sla * icla) * iN);
        d3 w_nb_b = k.w_eb_b - qrot_inv(k.q_nb, w_en_n);
        k.w_wb_b = k.w_eb_b - qrot_inv(k.q_nb, mk3(wx, wy, 0.0));
        double sf = 2.0 * sph * cph, cf = fma(-2.0 * sph, sph, 1.0);   // sin / cos of phi and theta from the half angles
        double st = 2.0 * sth * cth, ct = fma(-2.0 * sth, sth, 1.0);
        double sec = fb_rcp(ct), tn = st * sec;
        sink.put(L::KINX, (sf * w_nb_b.y + cf * w_nb_b.z) * sec);
        sink.put(L::KINX + 1, cf * w_nb_b.y - sf * w_nb_b.z);
        sink.put(L::KINX + 2, w_nb_b.x + (sf * w_nb_b.y + cf * w_nb_b.z) * tn);
        sink.put(L::KINX + 3, -wy);
        sink.put(L::KINX + 4, wx * icla);
        sink.put(L::KINX + 5, -k.v_eb_n.z);
    }
    if (k.h_e < -1000.0) out.err = FB_ST_BOUNDS;
    k.h_o = k.h_e - geoid_height(geoid, k.lat, k.lon);
    tap(TAP_LAT, k.lat); tap(TAP_LON, k.lon);
    tap(TAP_Q_NB, k.q_nb.w); tap(TAP_Q_NB + 1, k.q_nb.x); tap(TAP_Q_NB + 2, k.q_nb.y); tap(TAP_Q_NB + 3, k.q_nb.z);
    tap(TAP_V_EB_N, k.v_eb_n.x); tap(TAP_V_EB_N + 1, k.v_eb_n.y); tap(TAP_V_EB_N + 2, k.v_eb_n.z);
    tap(TAP_W_WB_B, k.w_wb_b.x); tap(TAP_W_WB_B + 1, k.w_wb_b.y); tap(TAP_W_WB_B + 2, k.w_wb_b.z);
    if (WY) {
        q4 q = k.q_nb;
        if (KIN == FB_NED) { yf->psi = x[L::KINX]; yf->theta = x[L::KINX + 1]; yf->phi = x[L::KINX + 2]; }
        else {   // REuler(q_nb), attitude.jl:382-390
            yf->psi = atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
            yf->theta = asin(clampd(2.0 * (q.w * q.y - q.x * q.z), -1.0, 1.0));
            yf->phi = atan2(2.0 * (q.w * q.x + q.y * q.z), 1.0 - 2.0 * (q.x * q.x + q.y * q.y));
        }
        yf->lat = k.lat; yf->lon = k.lon; yf->h_e = k.h_e; yf->h_o = k.h_o;
        yf->v_eb_n = k.v_eb_n; yf->w_wb_b = k.w_wb_b; yf->w_eb_b = k.w_eb_b;
        yf->v_gnd = sqrt(dot(k.v_eb_n, k.v_eb_n));
        bool valid = (KIN == FB_NED) || (yf->v_gnd > 0.1);   // kinematics.jl:212-213 vs :393-394
        yf->chi_gnd = valid ? atan2(k.v_eb_n.y, k.v_eb_n.x) : 0.0;
        yf->gamma_gnd = valid ? atan2(-k.v_eb_n.z, sqrt(fma(k.v_eb_n.x, k.v_eb_n.x, k.v_eb_n.y * k.v_eb_n.y))) : 0.0;
    }

    // ------------------------------------------------------------------ atmosphere + air data
    Air air;
    isa_data<AIR>(k.h_o, u(FB_U_T_SL), [&]() { return u(FB_U_P_SL); }, u(FB_UX_TRC), air.T, air.p, air.tr, out.err, anc);
    {
        air.iRT = fb_rcp(isa::R * air.T);
        air.rho = air.p * air.iRT;
        air.ia = fb_rsqrt(isa::gamma * isa::R * air.T);      // 1 / speed of sound
        d3 wind = mk3(u(FB_U_WIND_N), u(FB_U_WIND_E), u(FB_U_WIND_D));
        air.v_wb_b = k.v_eb_b - qrot_inv(k.q_nb, wind);
        double tas2 = dot(air.v_wb_b, air.v_wb_b);
        air.iTAS = fb_rsqrt(fb_max(tas2, 1e-280));
        air.TAS = tas2 * air.iTAS;
        air.q = 0.5 * air.rho * tas2;
        tap(TAP_TAS, air.TAS); tap(TAP_RHO, air.rho);
    }
    if (WY) {   // atmosphere.jl:220-242
        yf->TAS = air.TAS; yf->q = air.q;
        yf->EAS = air.TAS * sqrt(air.rho / isa::rho_std);
        yf->M = air.TAS * air.ia;
        double Tt = air.T * (1.0 + (isa::gamma - 1) / 2 * yf->M * yf->M);
        double pt = air.p * pow(Tt / air.T, isa::gamma / (isa::gamma - 1));
        yf->CAS = sqrt(2 * isa::gamma / (isa::gamma - 1) * isa::p_std / isa::rho_std * (pow(1.0 + (pt - air.p) / isa::p_std, (isa::gamma - 1) / isa::gamma) - 1.0));
    }

    // ------------------------------------------------------------------ actuation
    double a_e, a_a, a_r, a_f, steer, br_l, br_r, throttle;
    if (VAR == FB_C172S) {
        double ru = u(FB_UX_RU);
        steer = ru; a_r = -ru;
        a_e = u(FB_UX_AE);
        a_a = u(FB_UX_AA);
        a_f = u(FB_U_FLAPS); br_l = u(FB_U_BRAKE_LEFT); br_r = u(FB_U_BRAKE_RIGHT);
        throttle = u(FB_U_THROTTLE);
    } else if (VAR == FB_C172XA2) {   // seven Actuator2 units, c172x.jl:91-106: x = (v, p), v' = w_n^2 (cmd - p) - 2 zeta w_n v, p' = v
        const FbTablesTail* const tl = tables_tail(geoid);
        const double wn2 = tl->act2_wn2, zw2 = tl->act2_2zw;
        double pos[7];
#pragma unroll
        for (int i = 0; i < 7; i++) {
            const double lo = (i == 0 || i >= 4) ? 0.0 : -1.0;
            const double cmd = u(FB_U_ACT_THROTTLE + i);   // stored saturated (Ranged)
            const double xv = x[L::ACT + 2 * i], xp = x[L::ACT + 2 * i + 1];
            pos[i] = clampd(xp, lo, 1.0);
            sink.put(L::ACT + 2 * i, fma(wn2, cmd - xp, -(zw2 * xv)));
            sink.put(L::ACT + 2 * i + 1, xv);
            if (WY) { yf->act_pos[i] = pos[i]; yf->act_cmd[i] = cmd; }
        }
        throttle = pos[0]; a_a = pos[1]; a_e = -pos[2]; a_r = -pos[3]; steer = pos[3]; a_f = pos[4]; br_l = pos[5]; br_r = pos[6];
    } else {
        double pos[7];
#pragma unroll
        for (int i = 0; i < 7; i++) {
            double lo = (i == 0 || i >= 4) ? 0.0 : -1.0;
            double cmd = u(FB_U_ACT_THROTTLE + i);   // stored saturated (Ranged)
            double xp = sink.act(i, L::ACT + i, x, cmd);
            pos[i] = clampd(xp, lo, 1.0);
            sink.put_act(L::ACT + i, cmd, xp);
            if (WY) { yf->act_pos[i] = pos[i]; yf->act_cmd[i] = cmd; }
        }
        throttle = pos[0]; a_a = pos[1]; a_e = -pos[2]; a_r = -pos[3]; steer = pos[3]; a_f = pos[4]; br_l = pos[5]; br_r = pos[6];
    }

    d3 F = mk3(0, 0, 0), tau = mk3(0, 0, 0);
    // ------------------------------------------------------------------ aerodynamics
    {
        constexpr double S = 16.165, b = 10.912, c = 1.494, tau_f = 0.02;
        constexpr double d2r = FB_PI / 180.0;
        double alpha_filt = x[L::AERO], beta_filt = x[L::AERO + 1];
        d3 v = air.v_wb_b;
        double alpha, beta, ca, sa;
        {   // alpha = beta = 0 below TAS 0.1 (covers both thresholds: c172.jl:324, atmosphere.jl:337); evaluated branch-free
            const bool moving = air.TAS > 0.1;
            double vxz2 = fma(v.x, v.x, v.z * v.z);
            const bool planar = vxz2 > 0.0;
            double ivxz = fb_rsqrt(fb_max(vxz2, 1e-280));
            double al = fb_atan2(v.z, v.x);
            double be = fb_atan2(v.y, vxz2 * ivxz);
            alpha = moving ? al : 0.0; beta = moving ? be : 0.0;
            ca = (moving && planar) ? v.x * ivxz : 1.0; sa = (moving && planar) ? v.z * ivxz : 0.0;
        }
        out.alpha = alpha;
        tap(TAP_ALPHA, alpha); tap(TAP_BETA, beta);
        double i2V = (air.TAS > 1.0) ? 0.5 * air.iTAS : 0.5;   // 1 / (2 max(TAS, V_min)), V_min = 1
        double afd = (1.0 / tau_f) * (alpha - alpha_filt), bfd = (1.0 / tau_f) * (beta - beta_filt);
        sink.put(L::AERO, afd); sink.put(L::AERO + 1, bfd);
        double p_nd = k.w_wb_b.x * b * i2V, q_nd = k.w_wb_b.y * c * i2V, r_nd = k.w_wb_b.z * b * i2V;
        double adot_nd = clampd(afd * c * i2V, -0.04, 0.04);
        double de = fma((23 + 28) * d2r * 0.5, a_e + 1.0, -28 * d2r);
        double da = fma(40 * d2r * 0.5, a_a + 1.0, -20 * d2r);
        double dr = fma(32 * d2r * 0.5, a_r + 1.0, -16 * d2r);
        double df = 30 * d2r * a_f;
        double dh_nd = (k.h_o - u(FB_U_TRN_ELEVATION)) * (1.0 / b);
        double a_c = clampd(alpha, -0.1, 0.36), b_c = clampd(beta, -0.2, 0.2);
        // flap axis (knots 0,10,20,30 deg) shared by five tables
        int jf; double wf;
        { double dfc = clampd(df, tb[TB_DF_X], tb[TB_DF_X + 3]); locate<FB_AX_DF>(tb, TB_DF_X, TB_I_DF, dfc, jf, wf); }
        double wf2 = clampd(df * (1.0 / (30 * d2r)), 0.0, 1.0);   // 2-knot flap axis (0, 30 deg)
        int ig; double wg;
        { double g = clampd(dh_nd, 0.0, 1.1); locate<FB_AX_GE>(tb, TB_GE_X, TB_I_GE, g, ig, wg); }
        double CD_ge = lerp(tb[TB_CD_GE + ig], tb[TB_CD_GE + ig + 1], wg);
        double CL_ge = lerp(tb[TB_CL_GE + ig], tb[TB_CL_GE + ig + 1], wg);
        double CD_df = lerp(tb[TB_CD_DF + jf], tb[TB_CD_DF + jf + 1], wf);
        double CL_df = lerp(tb[TB_CL_DF + jf], tb[TB_CL_DF + jf + 1], wf);
        double Cm_df = lerp(tb[TB_CM_DF + jf], tb[TB_CM_DF + jf + 1], wf);
        int ia; double wa;
        { double ac = clampd(a_c, tb[TB_CD_A_X], tb[TB_CD_A_X + 25]); locate<FB_AX_CD_A>(tb, TB_CD_A_X, TB_I_CD_A, ac, ia, wa); }
        const double* cd = tb + TB_CD_A_DF + ia * 4 + jf;
        double CD_a_df = lerp(lerp(cd[0], cd[1], wf), lerp(cd[4], cd[5], wf), wa);
        int il; double wl;
        { double ac = clampd(a_c, tb[TB_CL_A_X], tb[TB_CL_A_X + 16]); locate<FB_AX_CL_A>(tb, TB_CL_A_X, TB_I_CL_A, ac, il, wl); }
        double CL_a = lerp(tb[TB_CL_A + il * 2 + stall], tb[TB_CL_A + il * 2 + 2 + stall], wl);
        double CD_de = 0.06 * fabs(clampd(de, -1.0, 1.0));       // knots (-1,0,1) -> (0.06,0,0.06)
        double CD_b = 0.17 * fabs(b_c);
        // 2x2 tables over alpha in (0, 0.094) and flap (0, 30 deg)
        double wa2 = clampd(a_c * (1.0 / 0.094), 0.0, 1.0);
        auto t22 = [&](int o) { return lerp(lerp(tb[o], tb[o + 1], wf2), lerp(tb[o + 2], tb[o + 3], wf2), wa2); };
        double CY_p = t22(TB_CY_P), CY_r = t22(TB_CY_R), Cl_r = t22(TB_CL_R);
        // C_Y.beta_dflap: beta knots (-0.349, 0, 0.349); beta is clamped to +-0.2 so only the inner cells are used
        double CY_b_df;
        {
            double bb = b_c * (1.0 / 0.3490);
            int ib = (bb >= 0.0) ? 1 : 0;
            double wb = bb + (double)(1 - ib);
            const double* t = tb + TB_CY_B_DF + ib * 2;
            CY_b_df = lerp(lerp(t[0], t[1], wf2), lerp(t[2], t[3], wf2), wb);
        }
        double C_D = 0.027 + CD_ge * (CD_a_df + CD_df) + CD_de + CD_b;
        double C_Y = 0.1870 * dr + 0.0 * da + CY_b_df + CY_p * p_nd + CY_r * r_nd;
        double C_L = CL_ge * (CL_a + CL_df) + 0.43 * de + 3.9 * q_nd + 1.7 * adot_nd;
        double C_l = 0.229 * da + 0.0147 * dr + -0.09226 * b_c + -0.4840 * p_nd + Cl_r * r_nd;
        double C_m = 0.1 + -1.1220 * de + Cm_df + -1.8 * a_c + -12.4 * q_nd + -7.27 * adot_nd;
        double C_n = -0.0430 * dr + -0.0053 * da + 0.05874 * b_c + -0.0278 * p_nd + -0.0937 * r_nd;
        double qS = air.q * S;
        // stability -> body axes: Ry(-alpha) applied to (-C_D, C_Y, -C_L)
        double fx = -C_D, fz = -C_L;
        F = mk3(qS * (ca * fx - sa * fz), qS * C_Y, qS * (sa * fx + ca * fz));
        tau = mk3(qS * C_l * b, qS * C_m * c, qS * C_n * b);
        if (WY) { yf->alpha = alpha; yf->beta = beta; yf->alpha_filt = alpha_filt; yf->beta_filt = beta_filt; yf->C_L = C_L; yf->C_D = C_D; }
    }

    // ------------------------------------------------------------------ landing gear
    {
        double elev = u(FB_U_TRN_ELEVATION);
        if (AIR) {
            // compact layout: regulator states are zero and stay zero while the gate holds; otherwise hand the member over
            if (!(k.h_o - elev > 4.0)) out.err = FB_NEED_COMPLETE;
        } else if (k.h_o - elev > 4.0) {
            // every strut attachment is < 2.4 m from Ob: no wheel can reach the terrain; regulators just leak
#pragma unroll
            for (int i = 0; i < 6; i++) sink.put(L::LDG + i, -0.2 * x[L::LDG + i]);
            if (WY) { for (int i = 0; i < 3; i++) { yf->xi[i] = 0.0; yf->F_dmp[i] = 0.0; } }
        } else {
            out.gnd = 1;
            int surface = (int)u(FB_U_TRN_SURFACE);
            bool eng = u(FB_U_STEERING_ENGAGED) != 0.0;
            d3 Fl = mk3(0, 0, 0), taul = mk3(0, 0, 0); RhsOut ol; ol.wow = 0; ol.crash = 0; ol.err = 0; ol.alpha = 0.0; ol.gnd = 1;
            double xd[6];
            if (WY) {
                // output-producing RHS (log samples, test hooks; an out-of-line function called from inside the integrator
                // kernels): three out-of-line strut evaluations on a copy of the kinematic data. Kept in exactly this shape:
                // the register allocation of the calling kernels depends on it (ptxas allocates across the call).
                Kin kk = k;
                ldg_unit_ool(kk, geoid, elev, surface, mk3(-0.381, -1.092, 1.902), 39404, 9340, 9340, false, false, 0.0, br_l,
                             x[L::LDG], x[L::LDG + 1], xd[0], xd[1], Fl, taul, 0, ol, &yf->xi[0], &yf->F_dmp[0]);
                ldg_unit_ool(kk, geoid, elev, surface, mk3(-0.381, 1.092, 1.902), 39404, 9340, 9340, false, false, 0.0, br_r,
                             x[L::LDG + 2], x[L::LDG + 3], xd[2], xd[3], Fl, taul, 1, ol, &yf->xi[1], &yf->F_dmp[1]);
                ldg_unit_ool(kk, geoid, elev, surface, mk3(1.27, 0.0, 1.9), 26269, 3503, 3503, true, eng, steer, 0.0,
                             x[L::LDG + 4], x[L::LDG + 5], xd[4], xd[5], Fl, taul, 2, ol, &yf->xi[2], &yf->F_dmp[2]);
            } else {
                double xd_a = 0.0, xd_b = 0.0, xd_c = 0.0, xd_d = 0.0, xd_e = 0.0, xd_f = 0.0;
#pragma unroll 1
                for (int leg = 0; leg < 3; leg++) {   // left main, right main, nose (c172.jl:466-500): geometry and rates by select
                    const bool nose = (leg == 2);
                    const d3 r_bs_b = mk3(nose ? 1.27 : -0.381, nose ? 0.0 : (leg == 0 ? -1.092 : 1.092), nose ? 1.9 : 1.902);
                    const double k_s = nose ? 26269.0 : 39404.0, k_d = nose ? 3503.0 : 9340.0;
                    const double brake = nose ? 0.0 : (leg == 0 ? br_l : br_r);
                    const double x0 = (leg == 0) ? x[L::LDG] : ((leg == 1) ? x[L::LDG + 2] : x[L::LDG + 4]);
                    const double x1 = (leg == 0) ? x[L::LDG + 1] : ((leg == 1) ? x[L::LDG + 3] : x[L::LDG + 5]);
                    double a0, a1;
#if !FB_LDG_INLINE
                    ldg_unit_ool(k, geoid, elev, surface, r_bs_b, k_s, k_d, k_d, nose, nose && eng, steer, brake, x0, x1, a0, a1, Fl, taul, leg, ol, nullptr, nullptr);
#else
                    ldg_unit(k, geoid, elev, surface, r_bs_b, k_s, k_d, k_d, nose, nose && eng, steer, brake, x0, x1, a0, a1, Fl, taul, leg, ol, nullptr, nullptr);
#endif
                    if (leg == 0) { xd_a = a0; xd_b = a1; } else if (leg == 1) { xd_c = a0; xd_d = a1; } else { xd_e = a0; xd_f = a1; }
                }
                xd[0] = xd_a; xd[1] = xd_b; xd[2] = xd_c; xd[3] = xd_d; xd[4] = xd_e; xd[5] = xd_f;
            }
            F = F + Fl; tau = tau + taul;
            out.wow = ol.wow; out.crash = ol.crash; if (ol.err) out.err = ol.err;
#pragma unroll
            for (int i = 0; i < 6; i++) sink.put(L::LDG + i, xd[i]);
        }
    }

    // ------------------------------------------------------------------ power plant
    double h_rot_x;
    double mdot;
    {
        constexpr double d = 2.0, J_xx = 0.3, J_eng = 0.05;
        constexpr double w_rated = 2700 * FB_PI / 30, w_idle = 600 * FB_PI / 30, P_rated = 735.49875 * 200, tau_start = 40;
        double w = x[L::ENG];
        const d3 r_bp = mk3(2.055, 0.0, 0.833);
        d3 v = air.v_wb_b + cross(k.w_eb_b, r_bp);
        double vJ2 = dot(v, v);
        double v_J = fb_sqrt(fb_max(vJ2, 1e-280));          // (1e-140 instead of 0 for a propeller at rest in still air)
        double aw = fabs(w);
        double J = (2 * FB_PI / d) * v_J * fb_rcp(fb_max(aw, 1.0));
        double Mt = aw * (d / 2) * air.ia;
        int ij, im; double wj, wm;
        locate_uniform(0.0, 20 / 1.5, 21, J, ij, wj);
        locate_uniform(0.0, 20 / 1.5, 21, Mt, im, wm);
        const double* p00 = tb + TB_PROP + (ij * 21 + im) * 4;
        const double* p10 = p00 + 21 * 4;
        double cf[4];
#pragma unroll
        for (int c = 0; c < 4; c++) cf[c] = lerp(lerp(p00[c], p00[4 + c], wm), lerp(p10[c], p10[4 + c], wm), wj);
        if (VAR == FB_C172XA2) {   // Propeller{<:VariablePitch}: third table axis, the blade-pitch offset (propellers.jl:395-398, 253-273)
            const FbTablesTail* const tl = tables_tail(geoid);
            const double* const t3 = tl->prop3;
            if (WY) yf->prop_dbeta = 0.0;
            if (t3) {   // (uniform over the ensemble)
                const int nb = tl->prop3_n;
                const double db = fma(tl->prop3_hi - tl->prop3_lo, clampd(u(FB_U_PROP_PITCH), 0.0, 1.0), tl->prop3_lo);   // linear_scaling(u, range), types.jl:70-73
                const double tb3 = clampd((db - tl->prop3_lo) * tl->prop3_is, 0.0, (double)(nb - 1));
                const int ib = min((int)tb3, nb - 2);
                const double wb = tb3 - (double)ib;
                const double* q00 = t3 + (((long long)ij * 21 + im) * nb + ib) * 4;   // [iJ][iMt][iB][4]
                const double* q10 = q00 + 21 * nb * 4;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const double f00 = lerp(q00[c], q00[4 + c], wb), f01 = lerp(q00[nb * 4 + c], q00[nb * 4 + 4 + c], wb);
                    const double f10 = lerp(q10[c], q10[4 + c], wb), f11 = lerp(q10[nb * 4 + c], q10[nb * 4 + 4 + c], wb);
                    cf[c] = lerp(lerp(f00, f01, wm), lerp(f10, f11, wm), wj);
                }
                if (WY) yf->prop_dbeta = db;
            }
        } else if (WY) yf->prop_dbeta = 0.0;
        double al_p, be_p;
        {
            double vxz2 = fma(v.x, v.x, v.z * v.z);
            double vxz = fb_sqrt(fb_max(vxz2, 1e-280));
            double al = fb_atan2(v.z, v.x), be = fb_atan2(v.y, vxz);
            al_p = (v_J >= 0.1) ? al : 0.0; be_p = (v_J >= 0.1) ? be : 0.0;
        }
        double f = w * (1.0 / (2 * FB_PI)), f2 = f * f;
        constexpr double d4 = d * d * d * d, d5 = d4 * d;
        double kF = air.rho * f2 * d4, kM = air.rho * f2 * d5;
        d3 Fp = mk3(kF * cf[0], kF * cf[2] * be_p, kF * cf[2] * al_p);
        d3 Mp = mk3(kM * cf[1], kM * cf[3] * be_p, kM * cf[3] * al_p);
        F = F + Fp;
        tau = tau + Mp + cross(r_bp, Fp);
        h_rot_x = J_xx * w;
        double tau_load = Mp.x;
        // engine
        double xd_frc, xd_idle;
        double o_frc = pi_lane(-w, x[L::ENG + 2], 5.0, 200.0, 0.0, -1.0, 1.0, xd_frc);
        double o_idle = pi_lane(1.0 - w * (1.0 / w_idle), x[L::ENG + 1], 4.0, 2.0, 0.0, -0.5, 0.5, xd_idle);
        sink.put(L::ENG + 1, xd_idle); sink.put(L::ENG + 2, xd_frc);
        double tau_shaft;
        mdot = 0.0;
        double MAP_y = air.p, P_y = 0.0;
        if (eng_state == FB_ENG_OFF) {
            tau_shaft = o_frc * (0.01 * P_rated / w_rated);
        } else if (eng_state == FB_ENG_STARTING) {
            tau_shaft = tau_start;
            P_y = tau_shaft * w;
            if (WY) {   // MAP = mu * p_std needs mu also in this state (piston.jl:378-384)
                double delta = (air.p * (1.0 / isa::p_std)) * rsqrt(air.tr);
                double mri = 0.5 + o_idle;
                MAP_y = interp_2x9_line(tb + TB_MU_WOT, 0.667, 1.0, 0.441, 1.0, w * (1.0 / w_rated), delta) * fma(throttle, 1.0 - mri, mri) * isa::p_std;
            }
        } else {
            double n = w * (1.0 / w_rated);
            // delta = (p/p_std) (T_ISA(p)/T_std)^-0.5 (piston.jl:38-39)
            double itr = fb_rsqrt(air.tr);
            double delta = (air.p * (1.0 / isa::p_std)) * itr;
            double mu_ratio_idle = 0.5 + o_idle;
            double mu_wot = interp_2x9_line(tb + TB_MU_WOT, 0.667, 1.0, 0.441, 1.0, n, delta);
            double mu = mu_wot * fma(throttle, 1.0 - mu_ratio_idle, mu_ratio_idle);
            double rr = air.rho * (1.0 / isa::rho_std);
            double isq = fb_rsqrt(rr);          // k_f = 1 / sqrt(rho / rho_std)
            double sq_rr = rr * isq;
            double mixture = u(FB_U_MIXTURE);
            double mixture_pos = (u(FB_U_MIXTURE_AUTO) == 0.0) ? 0.5 * (mixture + 1.0)
                                                               : fma(mixture, 0.0950 - 0.0625, 0.0625) * sq_rr * (1.0 / 0.0950);
            double far = (0.0950 * mixture_pos) * isq;
            double d_wot = interp_2x9_line(tb + TB_D_WOT, 0.667, 1.0, 0.401, 0.936, n, mu);
            // pi_std(n, mu): gridded 13 x 3, Flat
            double pi_std, pi_wot;
            {
                double nc = clampd(n, tb[TB_N_STD], tb[TB_N_STD + 12]);
                int i; double wn; locate<FB_AX_N_STD>(tb, TB_N_STD, TB_I_N_STD, nc, i, wn);
                double mc = clampd(mu, 0.0, 1.0);
                int j = (mc >= 0.568) ? 1 : 0;
                double wmu = (j == 0) ? mc * (1.0 / 0.568) : (mc - 0.568) * (1.0 / (1.0 - 0.568));
                const double* t = tb + TB_PI_STD + i * 3 + j;
                pi_std = lerp(lerp(t[0], t[1], wmu), lerp(t[3], t[4], wmu), wn);
            }
            {   // pi_wot(n, d_wot): gridded 5 x 3 (delta knots 0, 0.441, 1), Flat in n and below, Line above in delta
                double nc = clampd(n, tb[TB_N_WOT], tb[TB_N_WOT + 4]);
                int i; double wn; locate<FB_AX_N_WOT>(tb, TB_N_WOT, TB_I_N_WOT, nc, i, wn);
                double dc = clampd(d_wot, 0.0, 1.0);
                int j = (dc >= 0.441) ? 1 : 0;
                double wd = (j == 0) ? dc * (1.0 / 0.441) : (dc - 0.441) * (1.0 / (1.0 - 0.441));
                const double* t = tb + TB_PI_WOT + i * 3 + j;
                pi_wot = lerp(lerp(t[0], t[1], wd), lerp(t[3], t[4], wd), wn);
                double g = lerp(t[1] - t[0], t[4] - t[3], wn) * (1.0 / (1.0 - 0.441));      // upper cell's slope (j = 1 whenever it is used)
                pi_wot = fma(fb_max(d_wot - 1.0, 0.0), g, pi_wot);                        // Line extrapolation above delta = 1
            }
            const bool at_std = fabs(d_wot - 1.0) < 5e-3;
            double pi_ISA = fma(fb_div(pi_wot - pi_std, at_std ? 1.0 : d_wot - 1.0), delta - 1.0, pi_std);
            pi_ISA = at_std ? pi_std : pi_ISA;
            pi_ISA = fb_max(pi_ISA, 0.0);
            double pi_pow = pi_ISA * fb_sqrt(air.tr * (isa::T_std * isa::R) * air.iRT);
            int jf; double wff;
            { double fc = clampd(far, tb[TB_F_X], tb[TB_F_X + 10]); locate<FB_AX_F>(tb, TB_F_X, TB_I_F, fc, jf, wff); }
            double pi_actual = pi_pow * lerp(tb[TB_PI_RATIO + jf], tb[TB_PI_RATIO + jf + 1], wff);
            double P_shaft = P_rated * pi_actual;
            tau_shaft = fb_div(P_shaft, (w > 0.0) ? w : 1.0);
            tau_shaft = (w > 0.0) ? tau_shaft : 0.0;
            MAP_y = mu * isa::p_std; P_y = P_shaft;
            // sfc_pow(n, pi_actual): gridded 5 x 8, Line
            double sfc;
            {
                double nc = clampd(n, tb[TB_N_SFC], tb[TB_N_SFC + 4]);
                double pc = clampd(pi_actual, tb[TB_PI_SFC], tb[TB_PI_SFC + 7]);
                int i, j; double wn, wp;
                locate<FB_AX_N_SFC>(tb, TB_N_SFC, TB_I_N_SFC, nc, i, wn);
                locate<FB_AX_PI_SFC>(tb, TB_PI_SFC, TB_I_PI_SFC, pc, j, wp);
                const double* t = tb + TB_SFC_POW + i * 8 + j;
                double f00 = t[0], f01 = t[1], f10 = t[8], f11 = t[9];
                sfc = lerp(lerp(f00, f01, wp), lerp(f10, f11, wp), wn);
                double g0 = lerp(f10 - f00, f11 - f01, wp) * tb[TB_I_N_SFC + i];
                double g1 = lerp(f01 - f00, f11 - f10, wn) * tb[TB_I_PI_SFC + j];
                sfc = fma(n - nc, g0, fma(pi_actual - pc, g1, sfc));
            }
            double SFC = sfc * lerp(tb[TB_SFC_RATIO + jf], tb[TB_SFC_RATIO + jf + 1], wff);
            mdot = SFC * P_shaft;
        }
        sink.put(L::ENG, (tau_shaft + tau_load) * (1.0 / (J_eng + J_xx)));
        if (WY) {
            yf->eng_omega = w; yf->eng_n = w * (1.0 / w_rated); yf->eng_MAP = MAP_y; yf->eng_P_shaft = P_y; yf->eng_tau_shaft = tau_shaft;
            yf->eng_mdot = mdot; yf->prop_thrust = Fp.x; yf->prop_J = J;
        }
    }
    // ------------------------------------------------------------------ fuel
    constexpr double m_full = 114.4, m_res = 1.0;
    const double x_fuel = x[L::FUEL];
    sink.put(L::FUEL, -mdot * (1.0 / (m_full - m_res)));
    double m_fuel = fb_max(0.0, fma(x_fuel, m_full - m_res, m_res));

    // ------------------------------------------------------------------ mass properties (affine in the masses)
    // airframe at Ob: m 767, J_G diag(820,1164,1702), r (0.056,0,0.582): J_O = J_G + m (|r|^2 I - r r^T)
    // fuel: two tanks of m_fuel / 2 at (0.325, -+2.845, 0); payload: six moments of the five station masses (payload_moments)
    double m, mrx, mry, mrz, Jxx, Jyy, Jzz, Jxy, Jxz, Jyz;
    {
        constexpr double ma = 767.0, ax = 0.056, az = 0.582, fx = 0.325, fy = 2.845, pz = 0.899;
        double S[6];
        u.payload(S);
        m = (ma + m_fuel) + S[0];
        mrx = fma(fx, m_fuel, ma * ax) + S[1]; mry = S[2]; mrz = fma(pz, S[0], ma * az);
        Jxx = fma(fy * fy, m_fuel, 820.0 + ma * (az * az)) + fma(pz * pz, S[0], S[4]);
        Jyy = fma(fx * fx, m_fuel, 1164.0 + ma * (ax * ax + az * az)) + fma(pz * pz, S[0], S[3]);
        Jzz = fma(fx * fx + fy * fy, m_fuel, 1702.0 + ma * (ax * ax)) + (S[3] + S[4]);
        Jxy = -S[5]; Jxz = fma(-pz, S[1], -ma * ax * az); Jyz = -pz * S[2];
    }
    // ------------------------------------------------------------------ rigid-body dynamics (dynamics.jl:443-499)
    {
        double im = fb_rcp(m);
        d3 rc = mk3(mrx * im, mry * im, mrz * im);
        // inertia about the CoM: J_c = J_O - m (|rc|^2 I - rc rc^T)
        double a11 = Jxx - m * (rc.y * rc.y + rc.z * rc.z), a22 = Jyy - m * (rc.x * rc.x + rc.z * rc.z), a33 = Jzz - m * (rc.x * rc.x + rc.y * rc.y);
        double a12 = Jxy + m * rc.x * rc.y, a13 = Jxz + m * rc.x * rc.z, a23 = Jyz + m * rc.y * rc.z;
        d3 w = k.w_eb_b;
        d3 tau_c = tau - cross(rc, F);
        d3 v_c = k.v_eb_b + cross(w, rc);
#if FB_COM_EXACT
        d3 w_ie_b = qrot_inv(k.q_eb, mk3(0.0, 0.0, wgs::w_ie));
        d3 w_ic = w_ie_b + w;
        d3 r_ec_e = k.r_eb_e + qrot(k.q_eb, rc);
        d3 n_c; double h_c;
        geo_from_cartesian(r_ec_e, n_c, h_c);
        if (h_c < -1000.0) out.err = FB_ST_BOUNDS;
        double g = gravity(n_c.z, h_c);
        // Gravity in body axes. The reference rotates (0,0,g) by q_cl = q_eb' o ltf(Oc) (dynamics.jl:481-488); its
        // quaternion rotation v + 2 q_im x (q_re v + q_im x v) equals v + |q|^2 (R v - v), and the state quaternions
        // are only renormalised when | |q| - 1 | > 1e-8 (kinematics.jl:114-118), so the norm defect is part of the
        // reference's arithmetic. The local vertical at the CoM is -n_c, hence R (0,0,g) = R_be (-g n_c); the term
        // (|q_eb|^2 - 1)(u - v) restores the reference's defect exactly without building ltf(Oc).
        d3 u_dn = (-g) * n_c;
        d3 g_c = qrot_inv(k.q_eb, u_dn);
        {
            double s1 = fma(k.q_eb.w, k.q_eb.w, fma(k.q_eb.x, k.q_eb.x, fma(k.q_eb.y, k.q_eb.y, k.q_eb.z * k.q_eb.z))) - 1.0;
            g_c = mk3(fma(s1, u_dn.x, g_c.x), fma(s1, u_dn.y, g_c.y), fma(s1, u_dn.z - g, g_c.z));
        }
#else
        // Geodetic coordinates of the CoM without a second ECEF -> geodetic conversion. The reference converts
        // r_ec_e = r_eb_e + R_eb rc back with Fukushima's method (dynamics.jl:481-487, geodesy.jl:367-412) to get the
        // ellipsoid normal n_c and the altitude h_c at Oc for gravity(Oc) and ltf(Oc). Oc is |rc| < 1.5 m from Ob, so
        // with (dN, dE, dD) = R_nb rc the normal tilts by thN = dN / (M + h) towards north and thE = dE / (N + h)
        // towards east, sin(lat_c) = sin(lat) + thN cos(lat), and h_c = h_e - dD + (thN dN + thE dE) / 2 (the
        // ellipsoid falling away under the tangent plane); what is dropped is of second order in |rc| / R,
        // < 1e-13 relative on the direction and 1e-14 m on the altitude. Gravity at Oc in the local-level axes of Ob
        // is then g (-thN, -thE, 1), and q_nb takes it to body axes; the Earth rate comes the same way from its
        // local-level components w_ie (cos lat, 0, -sin lat).
        // Norm defect (the state quaternions are only renormalised when | |q| - 1 | > 1e-8, kinematics.jl:114-118, and a
        // steady turn holds |q_wb|^2 - 1 at a few 1e-10, so it is part of the reference's arithmetic — dropping it moves a
        // 600 s cfg-3 trajectory by 2e-9): the reference rotates a vector v given in the axes it starts from by a
        // quaternion q with the formula v + 2 q_im x (q_re v + q_im x v) = R v + (|q|^2 - 1)(R v - v). It does so with
        // q_eb' on the Earth rate (0, 0, w_ie) and with q_eb' o ltf(Oc) on (0, 0, g); here both are rotated by q_nb' from
        // their local-level components t, which carries (|q_nb|^2 - 1)(R t - t) instead. To first order in the defects
        // the difference is (|q_eb|^2 - 1)(t - v) on the Earth rate (t and v are far apart) and (|q_ew|^2 - 1)(R v - v) on
        // gravity (t = v to 2e-7); both are restored below.
        const d3 t_ie = mk3(wgs::w_ie * k.clat, 0.0, -wgs::w_ie * k.n_e.z);
        d3 w_ie_b = qrot_inv(k.q_nb, t_ie);
        w_ie_b = mk3(fma(k.s_eb, t_ie.x, w_ie_b.x), w_ie_b.y, fma(k.s_eb, t_ie.z - wgs::w_ie, w_ie_b.z));
        d3 w_ic = w_ie_b + w;
        d3 d_n = qrot(k.q_nb, rc);
        const double thN = d_n.x * k.iMh, thE = d_n.y * k.iNh;
        const double h_c = k.h_e - d_n.z + 0.5 * fma(thN, d_n.x, thE * d_n.y);
        if (h_c < -1000.0) out.err = FB_ST_BOUNDS;
        const double g = gravity(fma(thN, k.clat, k.n_e.z), h_c);
        d3 g_c = qrot_inv(k.q_nb, mk3(-g * thN, -g * thE, g));
        if (KIN == FB_WA) g_c = mk3(fma(k.s_ew, g_c.x, g_c.x), fma(k.s_ew, g_c.y, g_c.y), fma(k.s_ew, g_c.z - g, g_c.z));
#endif
        auto Jmul = [&](d3 v) { return mk3(a11 * v.x + a12 * v.y + a13 * v.z, a12 * v.x + a22 * v.y + a23 * v.z, a13 * v.x + a23 * v.y + a33 * v.z); };
        d3 hc = Jmul(w_ic);
        hc.x += h_rot_x;
        d3 rhs3 = tau_c - Jmul(cross(w_ie_b, w)) - cross(w_ic, hc);
        // symmetric 3x3 solve by cofactors
        double c11 = a22 * a33 - a23 * a23, c12 = a13 * a23 - a12 * a33, c13 = a12 * a23 - a13 * a22;
        double c22 = a11 * a33 - a13 * a13, c23 = a12 * a13 - a11 * a23, c33 = a11 * a22 - a12 * a12;
        double idet = fb_rcp(a11 * c11 + a12 * c12 + a13 * c13);
        d3 wd = mk3((c11 * rhs3.x + c12 * rhs3.y + c13 * rhs3.z) * idet,
                    (c12 * rhs3.x + c22 * rhs3.y + c23 * rhs3.z) * idet,
                    (c13 * rhs3.x + c23 * rhs3.y + c33 * rhs3.z) * idet);
        d3 vd_c = im * F + g_c - cross(w + 2.0 * w_ie_b, v_c);
        d3 vd_b = vd_c - cross(wd, rc);
        sink.put(L::DYN, wd.x); sink.put(L::DYN + 1, wd.y); sink.put(L::DYN + 2, wd.z);
        sink.put(L::DYN + 3, vd_b.x); sink.put(L::DYN + 4, vd_b.y); sink.put(L::DYN + 5, vd_b.z);
        if (WY) {   // specific force at the CoM, dynamics.jl:503-508: f = a_ic - gamma = vd_c + (w + 2 w_ie) x v_c - g_c
            yf->f_c_c = vd_c + cross(w + 2.0 * w_ie_b, v_c) - g_c;
            yf->mass = m; yf->fuel_m_total = fma(x_fuel, m_full - m_res, m_res); yf->wow = out.wow;
        }
    }
}

// f_step!(world) (world.jl:34-39 -> aircraftbase.jl:172-181 -> c172.jl:715-724). `y` is the RhsOut of the
// evaluation at the post-step state. X is any indexable state holder (registers or a shared-memory column).
// Returns true when something the RHS reads changed (so that the first-same-as-last derivative must be
// re-evaluated).
template <int VAR, int KIN, bool AIR = false, class X, class U>
FB_DEV bool f_step(X& x, int& stall, int& eng_state, const U& u, const RhsOut& y, int& status) {
    typedef Lay<VAR, KIN, AIR> L;
    bool changed = false;
    if (KIN == FB_WA) {   // kinematics.jl:114-118,226-229
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int o = L::KINX + 4 * b;
            double q0 = x.get(o), q1 = x.get(o + 1), q2 = x.get(o + 2), q3 = x.get(o + 3);
            double nn = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
            if (fabs(nn - 1.0) > 1e-8) { x.set(o, q0 / nn); x.set(o + 1, q1 / nn); x.set(o + 2, q2 / nn); x.set(o + 3, q3 / nn); changed = true; }
        }
    } else if (KIN == FB_ECEF) {   // kinematics.jl:317-320: q_eb and n_e blocks
        {
            const int o = L::KINX;
            double q0 = x.get(o), q1 = x.get(o + 1), q2 = x.get(o + 2), q3 = x.get(o + 3);
            double nn = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
            if (fabs(nn - 1.0) > 1e-8) { x.set(o, q0 / nn); x.set(o + 1, q1 / nn); x.set(o + 2, q2 / nn); x.set(o + 3, q3 / nn); changed = true; }
        }
        {
            const int o = L::KINX + 4;
            double n0 = x.get(o), n1 = x.get(o + 1), n2 = x.get(o + 2);
            double nn = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
            if (fabs(nn - 1.0) > 1e-8) { x.set(o, n0 / nn); x.set(o + 1, n1 / nn); x.set(o + 2, n2 / nn); changed = true; }
        }
    }
    int st0 = stall;
    if (y.alpha > 0.36) stall = 1; else if (y.alpha < 0.09) stall = 0;   // c172.jl:375-384
    changed |= (stall != st0);
#pragma unroll
    for (int leg = 0; leg < (AIR ? 0 : 3); leg++) {
        int cr = (y.crash >> (2 * leg)) & 3;                             // landinggear.jl:331-347
        if (cr) { status = (cr == 1) ? FB_ST_CRASH_ANGLE : FB_ST_CRASH_RATE; return changed; }
        if (!(y.wow & (1 << leg))) {                                     // landinggear.jl:479-483
            double a = x.get(L::LDG + 2 * leg), b = x.get(L::LDG + 2 * leg + 1);
            if (a != 0.0 || b != 0.0) { changed = true; x.set(L::LDG + 2 * leg, 0.0); x.set(L::LDG + 2 * leg + 1, 0.0); }
        }
    }
    constexpr double w_stall = 300 * FB_PI / 30, w_idle = 600 * FB_PI / 30;
    int e0 = eng_state;                                                  // piston.jl:428-453
    double w = x.get(L::ENG);
    bool fuel_available = (x.get(L::FUEL) * (114.4 - 1.0)) > 0.0;
    bool start = u(FB_U_ENG_START) != 0.0, stop = u(FB_U_ENG_STOP) != 0.0;
    if (eng_state == FB_ENG_OFF) { if (start) eng_state = FB_ENG_STARTING; }
    else if (eng_state == FB_ENG_STARTING) {
        if (!start) eng_state = FB_ENG_OFF;
        if (w > w_idle && fuel_available) eng_state = FB_ENG_RUNNING;
    } else {
        if (stop || w < w_stall || !fuel_available) eng_state = FB_ENG_OFF;
    }
    changed |= (eng_state != e0);
    return changed;
}
