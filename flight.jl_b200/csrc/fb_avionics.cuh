// fb_avionics.cuh — discrete-time side of the path: output extraction (the root `y` the reference saves),
// and the C172X gain-scheduled control laws run once per Δt by the thread that owns the aircraft.
// Behavioural spec (paths relative to the reference root):
//   discrete compensators  lib/FlightPhysics/src/control.jl:123-185 (Integrator), :370-471 (PID), :620-743 (LQR)
//   gain lookups           control.jl:950-994: scale(interpolate(A, BSpline(Linear())), 25:5:55, 50:1000:3050), Flat
//   ControlLawsLon         lib/FlightApps/src/c172/c172x/control/c172x_ctl.jl:201-284 (model), :286-446 (f_periodic!),
//                          :449-519 (assign!, f_init!)
//   ControlLawsLat         c172x_ctl.jl:813-878 (model), :880-983 (f_periodic!), :986-1032 (assign!, f_init!)
//   Cessna172Xv2           c172x/c172x2.jl:28-37 + guidance/c172x_gdc.jl:297-329 (ModeGuidance.direct forwards)
// Layout: the per-member controller memory (integrator states, saturation flags, previous mode, last commands) is
// one column of E.ctl ([FB_CTL_STATE_COUNT][pitch], member fastest); a thread loads its column once per launch,
// keeps it in thread-local storage across the fused steps and writes it back at the end. Gain schedules stay in
// global memory (45 KB, read-only path): only the compensators of the active mode are looked up.
#pragma once
#include "fb_physics.cuh"

// The control and guidance laws are inlined into the integrator (one call site each): out of line they would take the
// controller memory and the vehicle outputs through local memory (measured: +8 % on cfg 3). f_init!, which runs each
// law several times, goes through the out-of-line wrappers below.
#define FB_GDC_FN FB_DEV
#define FB_CTL_FN FB_DEV
#define FB_CTL_STATE_COUNT 80
#define FB_CTL_TAP_ROW 48      // rows 48..63 of the column: output tap of the integrator (TAP_COUNT values, fb_physics.cuh)
#define FB_CTL_GDC_ROW 64      // rows 64..75: outputs of the last guidance update (GDC_* below), written in place
enum { GDC_MODE = 0, GDC_HOR, GDC_VRT, GDC_DCHI, GDC_CHI_REF, GDC_H_REF, GDC_CHI_12, GDC_GAMMA_12, GDC_S_12, GDC_S_1B, GDC_E_SB, GDC_V_SB, GDC_COUNT };

enum { LON_DIRECT = 0, LON_SAS, LON_THR_Q, LON_THR_THETA, LON_THR_EAS, LON_EAS_Q, LON_EAS_THETA, LON_EAS_CLM, LON_EAS_ALT };  // c172x_ctl.jl:29-41
enum { LAT_DIRECT = 0, LAT_SAS, LAT_P_BETA, LAT_PHI_BETA, LAT_CHI_BETA };                                                  // :733-741

struct LqrS { double io[2]; int sat[2]; };          // LQR.s: int_out_0, out_sat_0 (control.jl:678-681)
struct PidS { double xi, xd; int sat; };            // PID.s: x_i0, x_d0, sat_out_0 (control.jl:386-390)
struct IntS { double x0; int sat; };                // Integrator.s (control.jl:133-136)

struct CtlState {
    int lon_mode, lat_mode, h_state;                // y.lon.mode, y.lat.mode (previous mode for bumpless transfer), s.h_state
    double thr_cmd, ele_cmd, ail_cmd, rud_cmd;      // y.*_cmd of the last update
    LqrS te2te, tv2te, vh2te, ar2ar, pb2ar;
    double te2te_zref1, pb2ar_zref0;                // te2te_lqr.u.z_ref.elevator_cmd, φβ2ar_lqr.u.z_ref.φ (read on resets)
    IntS q2e_int, p2phi_int;
    PidS q2e_pid, c2t_pid, v2t_pid, p2phi_pid, chi2phi_pid;
    double* gdc = nullptr; long long gdc_pitch = 0;   // this member's guidance outputs (rows FB_CTL_GDC_ROW.. of its column)

    template <bool STORE>
    FB_DEV void xfer(double* p, long long pitch) {
        int k = 0;
        auto D = [&](double& v) { if (STORE) p[k * pitch] = v; else v = p[k * pitch]; k++; };
        auto I = [&](int& v) { if (STORE) p[k * pitch] = (double)v; else v = (int)p[k * pitch]; k++; };
        auto L = [&](LqrS& s) { D(s.io[0]); D(s.io[1]); I(s.sat[0]); I(s.sat[1]); };
        auto P = [&](PidS& s) { D(s.xi); D(s.xd); I(s.sat); };
        auto N = [&](IntS& s) { D(s.x0); I(s.sat); };
        I(lon_mode); I(lat_mode); I(h_state); D(thr_cmd); D(ele_cmd); D(ail_cmd); D(rud_cmd);
        L(te2te); L(tv2te); L(vh2te); L(ar2ar); L(pb2ar); D(te2te_zref1); D(pb2ar_zref0);
        N(q2e_int); N(p2phi_int); P(q2e_pid); P(c2t_pid); P(v2t_pid); P(p2phi_pid); P(chi2phi_pid);
    }
    FB_DEV void load(double* p, long long pitch) { xfer<false>(p, pitch); gdc = p + (long long)FB_CTL_GDC_ROW * pitch; gdc_pitch = pitch; }
    FB_DEV void store(double* p, long long pitch) { xfer<true>(p, pitch); }
};
#define FB_CTL_ROW_H_STATE 2   // row of h_state in the column (constructed-model default: hold = 1)

__host__ __device__ inline double avionics_saturate_input(int f, double v) {
    switch (f) {
        case FB_U_LON_THROTTLE_AXIS: case FB_U_LON_THROTTLE_OFFSET: return clampd(v, 0.0, 1.0);
        case FB_U_LON_ELEVATOR_AXIS: case FB_U_LON_ELEVATOR_OFFSET:
        case FB_U_LAT_AILERON_AXIS: case FB_U_LAT_AILERON_OFFSET: case FB_U_LAT_RUDDER_AXIS: case FB_U_LAT_RUDDER_OFFSET:
            return clampd(v, -1.0, 1.0);
        default: return v;
    }
}

// the WY=true RHS as an out-of-line function: called once per log sample / control period, never in the
// RK stages, so the hot loop keeps a single inlined copy of the lean RHS
template <int VAR, int KIN>
__device__ __noinline__ void rhs_with_outputs(const double* x, double* xd, int stall, int eng, MemberIn u,
                                              const double* tb, const float* geoid, RhsOut* out, YFull* yf) {
    ArraySink sink; sink.xd = xd;
    rhs<VAR, KIN, true>(x, sink, stall, eng, u, tb, geoid, *out, yf);
}

FB_DEV double pick_output(const YFull& y, int stall, int eng, const CtlState& cs, int field) {
    switch (field) {
        case FB_Y_PSI: return y.psi; case FB_Y_THETA: return y.theta; case FB_Y_PHI: return y.phi;
        case FB_Y_LAT: return y.lat; case FB_Y_LON: return y.lon; case FB_Y_H_E: return y.h_e; case FB_Y_H_O: return y.h_o;
        case FB_Y_V_N: return y.v_eb_n.x; case FB_Y_V_E: return y.v_eb_n.y; case FB_Y_V_D: return y.v_eb_n.z;
        case FB_Y_V_GND: return y.v_gnd; case FB_Y_CHI_GND: return y.chi_gnd; case FB_Y_GAMMA_GND: return y.gamma_gnd;
        case FB_Y_P: return y.w_wb_b.x; case FB_Y_Q: return y.w_wb_b.y; case FB_Y_R: return y.w_wb_b.z;
        case FB_Y_TAS: return y.TAS; case FB_Y_EAS: return y.EAS; case FB_Y_CAS: return y.CAS; case FB_Y_MACH: return y.M; case FB_Y_QBAR: return y.q;
        case FB_Y_ALPHA: return y.alpha; case FB_Y_BETA: return y.beta; case FB_Y_ALPHA_FILT: return y.alpha_filt; case FB_Y_BETA_FILT: return y.beta_filt;
        case FB_Y_STALL: return (double)stall; case FB_Y_C_L: return y.C_L; case FB_Y_C_D: return y.C_D;
        case FB_Y_WOW_LEFT: return (double)(y.wow & 1); case FB_Y_WOW_RIGHT: return (double)((y.wow >> 1) & 1); case FB_Y_WOW_NOSE: return (double)((y.wow >> 2) & 1);
        case FB_Y_XI_LEFT: return y.xi[0]; case FB_Y_XI_RIGHT: return y.xi[1]; case FB_Y_XI_NOSE: return y.xi[2];
        case FB_Y_FDMP_LEFT: return y.F_dmp[0]; case FB_Y_FDMP_RIGHT: return y.F_dmp[1]; case FB_Y_FDMP_NOSE: return y.F_dmp[2];
        case FB_Y_ENG_OMEGA: return y.eng_omega; case FB_Y_ENG_STATE: return (double)eng; case FB_Y_ENG_MAP: return y.eng_MAP;
        case FB_Y_ENG_P_SHAFT: return y.eng_P_shaft; case FB_Y_ENG_TAU_SHAFT: return y.eng_tau_shaft; case FB_Y_ENG_MDOT: return y.eng_mdot;
        case FB_Y_FUEL_M_TOTAL: return y.fuel_m_total; case FB_Y_PROP_THRUST: return y.prop_thrust; case FB_Y_PROP_J: return y.prop_J;
        case FB_Y_PROP_DBETA: return y.prop_dbeta;
        case FB_Y_MASS: return y.mass; case FB_Y_F_X: return y.f_c_c.x; case FB_Y_F_Y: return y.f_c_c.y; case FB_Y_F_Z: return y.f_c_c.z;
        case FB_Y_LON_MODE: return (double)cs.lon_mode; case FB_Y_LAT_MODE: return (double)cs.lat_mode;
        case FB_Y_THROTTLE_CMD: return cs.thr_cmd; case FB_Y_ELEVATOR_CMD: return cs.ele_cmd;
        case FB_Y_AILERON_CMD: return cs.ail_cmd; case FB_Y_RUDDER_CMD: return cs.rud_cmd;
        case FB_Y_ALT_STATE: return (double)cs.h_state;
        default:
            if (field >= FB_Y_GDC_MODE && field <= FB_Y_GDC_V_SB && cs.gdc) return cs.gdc[(long long)(field - FB_Y_GDC_MODE) * cs.gdc_pitch];
            return nan("");
    }
}

// ------------------------------------------------------------------------------------------ gain schedules
// One bilinear cell of the (EAS, h_e) design grid, Flat outside (control.jl:965); value k of a schedule with P
// values per design point is read on demand.
struct GainCell {
    int o;          // (i * 4 + j): design-point index of the cell's low corner
    double u, w;    // weights along EAS and h
    FB_DEV void locate(double EAS, double h) {
        double fe = (clampd(EAS, 25.0, 55.0) - 25.0) * (1.0 / 5.0), fh = (clampd(h, 50.0, 3050.0) - 50.0) * (1.0 / 1000.0);
        int i = min((int)fe, 5), j = min((int)fh, 2);
        u = fe - (double)i; w = fh - (double)j; o = i * 4 + j;
    }
};
// Values of one schedule at the member's cell. SH: the schedules sit in the CTA's shared memory (the integrator stages
// them with the table pack: every member has its own cell, so from L2 each value was four uncoalesced round trips —
// the dominant stall of the fly-by-wire kernels); otherwise they are read from the device buffer.
template <bool SH>
struct GainT {
    const double* p; unsigned a; double w00, w01, w10, w11;   // low corner's record (generic pointer, or shared-window address when SH), corner weights
    int P;
    FB_DEV double ld(int idx) const {
        if (SH) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a + 8u * (unsigned)idx)); return v; }
        return __ldg(p + idx);
    }
    FB_DEV double operator()(int k) const {
        double f00 = ld(k), f01 = ld(P + k), f10 = ld(4 * P + k), f11 = ld(5 * P + k);
// Four corner weights (formed once per schedule, gain()) instead of the nested form (1-u)((1-w) f00 + w f01) + u (...):
        // one multiplication + three FMAs per value, same value to rounding
        return fma(w11, f11, fma(w10, f10, fma(w01, f01, w00 * f00)));
    }
    template <int N>
    FB_DEV void batch(int k0, double* out) const {
#pragma unroll
        for (int n = 0; n < N; n++) out[n] = (*this)(k0 + n);
    }
};
template <bool SH>
struct GainTables {       // the ten schedules, [7 EAS][4 h][P] each, concatenated (fb_tables.h)
    const double* base; unsigned sbase;
    FB_DEV GainT<SH> gain(int s, const GainCell& c) const {
        GainT<SH> g; g.P = gs_P(s);
        { const double um = 1.0 - c.u, wm = 1.0 - c.w; g.w00 = um * wm; g.w01 = um * c.w; g.w10 = c.u * wm; g.w11 = c.u * c.w; }
        g.p = base + gs_off(s) + c.o * gs_P(s);
        g.a = sbase + 8u * (unsigned)(gs_off(s) + c.o * gs_P(s));
        return g;
    }
};
typedef GainTables<false> GainsGlobal;
typedef GainTables<true> GainsShared;
FB_DEV GainsGlobal gains_global(const FbDeviceTables& T) { GainsGlobal g; g.base = T.gains; g.sbase = 0; return g; }
struct PidGains {      // (k_p, k_i, k_d, tau_f) of a scheduled PID
    double k_p, k_i, k_d, tau_f;
    template <class GT> FB_DEV PidGains(const GT& T, int s, const GainCell& c) { double g[4]; T.gain(s, c).template batch<4>(0, g); k_p = g[0]; k_i = g[1]; k_d = g[2]; tau_f = g[3]; }
};

// ------------------------------------------------------------------------------------------ compensators
// PID f_periodic! (control.jl:431-471), β_p = β_d = 1; gains g = (k_p, k_i, k_d, τ_f)
FB_DEV double pid_update(PidS& s, double k_p, double k_i, double k_d, double tau_f, double lo, double hi,
                         double input, int sat_ext, double Dt) {
    double alpha = fb_rcp(tau_f + Dt);
    bool halted = (isign(input * (double)s.sat) > 0) || (isign(input * (double)sat_ext) > 0);
    double x_i = s.xi + Dt * k_i * input * (halted ? 0.0 : 1.0);
    double x_d = alpha * tau_f * s.xd + Dt * alpha * k_d * input;
    double out_free = k_p * input + x_i + alpha * (-s.xd + k_d * input);
    s.sat = (out_free >= hi) - (out_free <= lo);
    s.xi = x_i; s.xd = x_d;
    return clampd(out_free, lo, hi);
}
FB_DEV void pid_reset(PidS& s) { s.xi = 0.0; s.xd = 0.0; s.sat = 0; }   // f_init!(pid): zero input leaves a zero state
// Integrator f_periodic! (control.jl:164-185), unbounded
FB_DEV double int_update(IntS& s, double input, int sat_ext, double Dt) {
    bool halted = (isign(input * (double)s.sat) > 0) || (isign(input * (double)sat_ext) > 0);
    double x1 = s.x0 + Dt * input * (halted ? 0.0 : 1.0);
    s.x0 = x1; s.sat = 0;   // bounds are +-Inf: never saturated
    return x1;
}
FB_DEV void int_reset(IntS& s) { s.x0 = 0.0; s.sat = 0; }

// LQR f_periodic! (control.jl:708-743), NU = NZ = 2. Schedule record: K_fbk (2 x NX row-major), K_fwd 2x2, K_int 2x2,
// x_trim NX, u_trim 2, z_trim 2 (packing of tools/build_tables.py).
template <int NX, class G>
FB_DEV void lqr_update(LqrS& s, const G& g, const double* x, double z0, double z1, double zr0, double zr1,
                       double lo0, double lo1, double Dt, double& out0, double& out1) {
    double dx[NX], t[NX], c[8], uz[4];
    g.template batch<NX>(2 * NX + 8, t);            // x_trim
#pragma unroll
    for (int j = 0; j < NX; j++) dx[j] = x[j] - t[j];
    g.template batch<8>(2 * NX, c);                 // K_fwd (2x2), K_int (2x2)
    g.template batch<4>(3 * NX + 8, uz);            // u_trim, z_trim
    const double zt0 = uz[2], zt1 = uz[3];
    double out[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        double int_in = c[4 + 2 * i] * (zr0 - z0) + c[5 + 2 * i] * (zr1 - z1);
        bool halted = isign(int_in * (double)s.sat[i]) > 0;
        double int_out = s.io[i] + Dt * int_in * (halted ? 0.0 : 1.0);
        double fwd = c[2 * i] * (zr0 - zt0) + c[1 + 2 * i] * (zr1 - zt1);
        g.template batch<NX>(i * NX, t);            // row i of K_fbk
        double fbk = 0.0;
#pragma unroll
        for (int j = 0; j < NX; j++) fbk = fbk + t[j] * dx[j];
        double out_free = uz[i] + int_out + fwd - fbk;
        const double lo = i ? lo1 : lo0;
        s.sat[i] = (out_free >= 1.0) - (out_free <= lo);
        out[i] = clampd(out_free, lo, 1.0);
        s.io[i] = int_out;
    }
    out0 = out[0]; out1 = out[1];
}
// f_init!(lqr) with the gains currently assigned: zero inputs, one update (leaves int_out_0 = 0 and the saturation
// flags of the zero-input output)
template <int NX, class G>
FB_DEV void lqr_reset(LqrS& s, const G& g, double lo0, double lo1, double Dt) {
    s.io[0] = 0.0; s.io[1] = 0.0; s.sat[0] = 0; s.sat[1] = 0;
    double x[NX];
#pragma unroll
    for (int j = 0; j < NX; j++) x[j] = 0.0;
    double a, b;
    lqr_update<NX>(s, g, x, 0.0, 0.0, 0.0, 0.0, lo0, lo1, Dt, a, b);
}

// ------------------------------------------------------------------------------------------ control laws
// f_periodic!(ControlLawsLon) — c172x_ctl.jl:286-446
template <class U, class GT>
FB_CTL_FN void lon_periodic(CtlState& cs, const YFull& v, int wow, const U& u,
                                          const GT& T, double Dt, int mode_req, const GainCell& c) {
    const double EAS = v.EAS, h_e = v.h_e, q = v.w_wb_b.y, r = v.w_wb_b.z, theta = v.theta, phi = v.phi;
    const double clm = -v.v_eb_n.z;
    const double h_ref = u(FB_U_LON_H_REF), EAS_ref = u(FB_U_LON_EAS_REF);
    const double h_err = h_ref - h_e;
    const int mode_prev = cs.lon_mode;
    double throttle_ref = clampd(u(FB_U_LON_THROTTLE_AXIS) + u(FB_U_LON_THROTTLE_OFFSET), 0.0, 1.0);
    double elevator_ref = clampd(u(FB_U_LON_ELEVATOR_AXIS) + u(FB_U_LON_ELEVATOR_OFFSET), -1.0, 1.0);
    double throttle_cmd = throttle_ref, elevator_cmd = elevator_ref;
    double q_ref = u(FB_U_LON_Q_REF), theta_ref = u(FB_U_LON_THETA_REF);
    int mode;
    if (wow) mode = LON_DIRECT;
    else if (mode_req == LON_EAS_ALT) {
        if (cs.h_state == 0) {   // acquire
            mode = LON_THR_EAS;
            throttle_ref = h_err > 0.0 ? 1.0 : 0.0;
            if (fabs(h_err) < 10.0 - 1.0) cs.h_state = 1;
        } else {
            mode = LON_EAS_ALT;
            if (fabs(h_err) > 10.0 + 1.0) cs.h_state = 0;
        }
    } else mode = mode_req;
    const bool te2te_en = (mode >= LON_SAS && mode <= LON_THR_THETA) || (mode >= LON_EAS_Q && mode <= LON_EAS_CLM);
    const bool q2e_en = te2te_en && mode != LON_SAS;
    const bool t2q_en = mode == LON_THR_THETA || mode == LON_EAS_THETA || mode == LON_EAS_CLM;
    const bool v2t_en = mode >= LON_EAS_Q && mode <= LON_EAS_CLM;
    const double xred[8] = {v.w_eb_b.y, theta, EAS, v.alpha, v.alpha_filt, v.eng_n, v.act_pos[0], v.act_pos[2]};   // XLonRed :132-147
    if (te2te_en) {
        const int sat_thr = cs.te2te.sat[0], sat_ele = cs.te2te.sat[1];
        if (v2t_en) {
            const PidGains g(T, GS_V2T, c);
            if (mode != mode_prev) { pid_reset(cs.v2t_pid); if (g.k_i != 0.0) cs.v2t_pid.xi = cs.thr_cmd; }
            throttle_ref = pid_update(cs.v2t_pid, g.k_p, g.k_i, g.k_d, g.tau_f, -INFINITY, INFINITY, EAS_ref - EAS, sat_thr, Dt);
        }
        if (q2e_en) {
            const PidGains g(T, GS_Q2E, c);
            if (mode != mode_prev) { int_reset(cs.q2e_int); pid_reset(cs.q2e_pid); if (g.k_i != 0.0) cs.q2e_pid.xi = cs.te2te_zref1; }
            if (t2q_en) {
                if (mode == LON_EAS_CLM) {
                    const PidGains gc(T, GS_C2THETA, c);
                    if (mode != mode_prev) { pid_reset(cs.c2t_pid); if (gc.k_i != 0.0) cs.c2t_pid.xi = theta; }
                    theta_ref = pid_update(cs.c2t_pid, gc.k_p, gc.k_i, gc.k_d, gc.tau_f, -INFINITY, INFINITY, u(FB_U_LON_CLM_REF) - clm, sat_ele, Dt);
                }
                double theta_dot_ref = 1.0 * (theta_ref - theta);
                double phi_bnd = clampd(phi, -FB_PI / 3, FB_PI / 3);
                // 1 / cos(phi) * theta_dot_ref + r * tan(phi) (c172x_ctl.jl:375-377): sin and cos of |phi| <= pi/3 by the
                // double-angle identities from the small-angle series at phi / 2
                double sh, ch; sincos_small(0.5 * phi_bnd, sh, ch);
                const double sp = 2.0 * sh * ch, cp = fma(-2.0 * sh, sh, 1.0), icp = fb_rcp(cp);
                q_ref = fma(icp, theta_dot_ref, r * (sp * icp));
            }
            double io = int_update(cs.q2e_int, q_ref - q, sat_ele, Dt);
            elevator_ref = pid_update(cs.q2e_pid, g.k_p, g.k_i, g.k_d, g.tau_f, -INFINITY, INFINITY, io, sat_ele, Dt);
        }
        const auto g = T.gain(GS_TE2TE, c);
        cs.te2te_zref1 = elevator_ref;
        lqr_update<8>(cs.te2te, g, xred, v.act_cmd[0], v.act_cmd[2], throttle_ref, elevator_ref, 0.0, -1.0, Dt, throttle_cmd, elevator_cmd);   // Zte :167-171
    }
    if (mode == LON_THR_EAS) {
        const auto g = T.gain(GS_TV2TE, c);
        if (mode != mode_prev) lqr_reset<8>(cs.tv2te, g, 0.0, -1.0, Dt);
        lqr_update<8>(cs.tv2te, g, xred, v.act_cmd[0], EAS, throttle_ref, EAS_ref, 0.0, -1.0, Dt, throttle_cmd, elevator_cmd);   // Ztv :190-194
    }
    if (mode == LON_EAS_ALT) {
        const auto g = T.gain(GS_VH2TE, c);
        if (mode != mode_prev) lqr_reset<9>(cs.vh2te, g, 0.0, -1.0, Dt);
        const double xfull[9] = {v.w_eb_b.y, theta, EAS, v.alpha, h_e, v.alpha_filt, v.eng_n, v.act_pos[0], v.act_pos[2]};   // XLonFull :97-118
        lqr_update<9>(cs.vh2te, g, xfull, EAS, h_e, EAS_ref, h_ref, 0.0, -1.0, Dt, throttle_cmd, elevator_cmd);   // Zvh :178-182
    }
    cs.lon_mode = mode;
    cs.thr_cmd = clampd(throttle_cmd, 0.0, 1.0);
    cs.ele_cmd = clampd(elevator_cmd, -1.0, 1.0);
}

FB_DEV double wrap_to_pi(double x) { return x + 2.0 * FB_PI * floor((FB_PI - x) / (2.0 * FB_PI)); }   // attitude.jl wrap_to_π

// f_periodic!(ControlLawsLat) — c172x_ctl.jl:880-983
template <class U, class GT>
FB_CTL_FN void lat_periodic(CtlState& cs, const YFull& v, int wow, const U& u,
                                          const GT& T, double Dt, int mode_req, const GainCell& c) {
    const double EAS = v.EAS, p = v.w_wb_b.x;
    const int mode_prev = cs.lat_mode;
    const int mode = wow ? LAT_DIRECT : mode_req;
    double aileron_ref = clampd(u(FB_U_LAT_AILERON_AXIS) + u(FB_U_LAT_AILERON_OFFSET), -1.0, 1.0);
    double rudder_ref = clampd(u(FB_U_LAT_RUDDER_AXIS) + u(FB_U_LAT_RUDDER_OFFSET), -1.0, 1.0);
    double aileron_cmd = aileron_ref, rudder_cmd = rudder_ref;
    double phi_ref = u(FB_U_LAT_PHI_REF);
    const double xlat[8] = {v.w_eb_b.x, v.w_eb_b.z, v.phi, EAS, v.beta, v.beta_filt, v.act_pos[1], v.act_pos[3]};   // XLatRed :767-779
    if (mode == LAT_SAS) {
        const auto g = T.gain(GS_AR2AR, c);
        lqr_update<8>(cs.ar2ar, g, xlat, v.act_cmd[1], v.act_cmd[3], aileron_ref, rudder_ref, -1.0, -1.0, Dt, aileron_cmd, rudder_cmd);   // Zar :806-810
    }
    if (mode == LAT_P_BETA || mode == LAT_PHI_BETA || mode == LAT_CHI_BETA) {
        const int sat_ail = cs.pb2ar.sat[0];
        if (mode == LAT_P_BETA) {
            const PidGains g(T, GS_P2PHI, c);
            if (mode != mode_prev) { int_reset(cs.p2phi_int); pid_reset(cs.p2phi_pid); if (g.k_i != 0.0) cs.p2phi_pid.xi = cs.pb2ar_zref0; }
            double io = int_update(cs.p2phi_int, u(FB_U_LAT_P_REF) - p, sat_ail, Dt);
            phi_ref = pid_update(cs.p2phi_pid, g.k_p, g.k_i, g.k_d, g.tau_f, -INFINITY, INFINITY, io, sat_ail, Dt);
        } else if (mode == LAT_CHI_BETA) {
            const PidGains g(T, GS_CHI2PHI, c);
            if (mode != mode_prev) { pid_reset(cs.chi2phi_pid); if (g.k_i != 0.0) cs.chi2phi_pid.xi = cs.pb2ar_zref0; }
            phi_ref = pid_update(cs.chi2phi_pid, g.k_p, g.k_i, g.k_d, g.tau_f, -FB_PI / 4, FB_PI / 4, wrap_to_pi(u(FB_U_LAT_CHI_REF) - v.chi_gnd), sat_ail, Dt);
        }
        const auto g = T.gain(GS_PHIBETA2AR, c);
        if (mode != mode_prev) lqr_reset<8>(cs.pb2ar, g, -1.0, -1.0, Dt);
        cs.pb2ar_zref0 = phi_ref;
        lqr_update<8>(cs.pb2ar, g, xlat, v.phi, v.beta, phi_ref, u(FB_U_LAT_BETA_REF), -1.0, -1.0, Dt, aileron_cmd, rudder_cmd);   // Zφβ :789-793
    }
    cs.lat_mode = mode;
    cs.ail_cmd = clampd(aileron_cmd, -1.0, 1.0);
    cs.rud_cmd = clampd(rudder_cmd, -1.0, 1.0);
}

// ------------------------------------------------------------------------------------------ guidance (Cessna172Xv2)
// f_periodic!(GuidanceLaws, ctl, vehicle): c172x_gdc.jl:297-329 -> SegmentGuidance :228-252 -> SegmentGuidanceData :113-148
// (dchi_inf = pi/2, e_sf = 250 m, e_thr = 1000 m, :203-207). On a Cessna172Xv1 the mode request stays `direct` (the host
// refuses guidance inputs there), which makes this a no-op. Cold code: runs once per control period, IEEE arithmetic.
template <class U>
FB_GDC_FN void gdc_periodic(const CtlState& cs, double lat, double lon, double h_e, int wow, U& u) {
    const int mode = wow ? 0 : (int)u(FB_U_GDC_MODE_REQ);
    auto out = [&](int k, double v) { cs.gdc[(long long)k * cs.gdc_pitch] = v; };
    if (mode == 1) {
        auto cart = [](double la, double lo, double h) {   // Cartesian(Geographic{LatLon}), geodesy.jl:418-428
            double sl, cl, so, co;
            sincos(la, &sl, &cl); sincos(lo, &so, &co);
            double N = wgs::a / sqrt(1.0 - wgs::e2 * sl * sl);
            return mk3((N + h) * cl * co, (N + h) * cl * so, (N * (1 - wgs::e2) + h) * sl);
        };
        const double h1 = u(FB_U_GDC_H1), h2 = u(FB_U_GDC_H2);
        d3 r1 = cart(u(FB_U_GDC_LAT1), u(FB_U_GDC_LON1), h1), r2 = cart(u(FB_U_GDC_LAT2), u(FB_U_GDC_LON2), h2), rb = cart(lat, lon, h_e);
        double sl2, cl2, sa2, ca2;   // ltf(Ob) = Rz(lon) o Ry(-(lat + pi/2)), geodesy.jl:132-135
        sincos(0.5 * lon, &sl2, &cl2); sincos(-0.5 * (lat + 0.5 * FB_PI), &sa2, &ca2);
        q4 q_en = qmul(mkq(cl2, 0, 0, sl2), mkq(ca2, 0, sa2, 0));
        d3 r_1b = qrot_inv(q_en, rb - r1), r_12 = qrot_inv(q_en, r2 - r1);
        double s_12 = sqrt(r_12.x * r_12.x + r_12.y * r_12.y);
        double ux = r_12.x / s_12, uy = r_12.y / s_12;
        double s_1b = ux * r_1b.x + uy * r_1b.y;
        double e_sb = ux * r_1b.y - uy * r_1b.x;
        double h_s = h1 + (h2 - h1) * s_1b / s_12;
        double chi_12 = atan2(uy, ux);
        double dchi = -(0.5 * FB_PI) / (0.5 * FB_PI) * atan(e_sb / 250.0);
        double xw = chi_12 + dchi;
        double chi_ref = xw + 2.0 * FB_PI * floor((FB_PI - xw) / (2.0 * FB_PI));   // wrap_to_pi, attitude.jl:478
        const bool hor = u(FB_U_GDC_HOR_REQ) != 0.0;
        const bool vrt = (fabs(e_sb) < 1000.0) ? (u(FB_U_GDC_VRT_REQ) != 0.0) : false;
        out(GDC_HOR, hor ? 1.0 : 0.0); out(GDC_VRT, vrt ? 1.0 : 0.0); out(GDC_DCHI, dchi); out(GDC_CHI_REF, chi_ref); out(GDC_H_REF, h_s);
        out(GDC_CHI_12, chi_12); out(GDC_GAMMA_12, atan2(h2 - h1, s_12)); out(GDC_S_12, s_12); out(GDC_S_1B, s_1b); out(GDC_E_SB, e_sb);
        out(GDC_V_SB, h_e - h_s);
        if (hor) { u.set(FB_U_LAT_CHI_REF, chi_ref); u.set(FB_U_LAT_MODE_REQ, (double)LAT_CHI_BETA); }
        if (vrt) { u.set(FB_U_LON_H_REF, h_s); u.set(FB_U_LON_MODE_REQ, (double)LON_EAS_ALT); }
    }
    out(GDC_MODE, (double)mode);
}

static __device__ __noinline__ void lon_periodic_ool(CtlState& cs, const YFull& v, int wow, const MemberIn& u, const FbDeviceTables& T, double Dt, int mode_req) {
    GainCell c; c.locate(v.EAS, v.h_e);
    lon_periodic(cs, v, wow, u, gains_global(T), Dt, mode_req, c);
}
static __device__ __noinline__ void lat_periodic_ool(CtlState& cs, const YFull& v, int wow, const MemberIn& u, const FbDeviceTables& T, double Dt, int mode_req) {
    GainCell c; c.locate(v.EAS, v.h_e);
    lat_periodic(cs, v, wow, u, gains_global(T), Dt, mode_req, c);
}

// The control- and guidance-law inputs of one member (fb_input 40..49, 60..68, 70) read in ONE batch of independent
// loads; the laws then index the copy with compile-time field ids. set() keeps the copy and the ensemble in step (the
// segment guidance hands references to the control laws that way, c172x_gdc.jl:316-326).
struct CtlIn {
    MemberIn g;
    double lon[10], lat[9], gdc_mode;
    FB_DEV void load() {   // rows 40..49, 60..68, 70 of the member's input column: one pointer walked by the pitch
        const double* p = g.u + (long long)FB_U_LON_MODE_REQ * g.pitch;
#pragma unroll
        for (int k = 0; k < 10; k++) { lon[k] = *p; p += g.pitch; }
        p += 10 * g.pitch;
#pragma unroll
        for (int k = 0; k < 9; k++) { lat[k] = *p; p += g.pitch; }
        gdc_mode = p[g.pitch];
    }
    FB_DEV double operator()(int f) const {
        if (f >= FB_U_LON_MODE_REQ && f < FB_U_LON_MODE_REQ + 10) return lon[f - FB_U_LON_MODE_REQ];
        if (f >= FB_U_LAT_MODE_REQ && f < FB_U_LAT_MODE_REQ + 9) return lat[f - FB_U_LAT_MODE_REQ];
        if (f == FB_U_GDC_MODE_REQ) return gdc_mode;
        return g(f);
    }
    FB_DEV void set(int f, double v) {
        g.set(f, v);
        if (f >= FB_U_LON_MODE_REQ && f < FB_U_LON_MODE_REQ + 10) lon[f - FB_U_LON_MODE_REQ] = v;
        if (f >= FB_U_LAT_MODE_REQ && f < FB_U_LAT_MODE_REQ + 9) lat[f - FB_U_LAT_MODE_REQ] = v;
    }
};

// assign!(vehicle, avionics): c172x_ctl.jl:449-458, 986-995 — the four fly-by-wire commands the control laws own.
// U: MemberIn (ensemble) or the integrator's CachedIn (compares against, and updates, its shared-memory copy).
template <class U>
FB_DEV bool avionics_assign(const CtlState& cs, U& u) {
    bool ch = false;
    double v;
    // (the four commands are stored already saturated: lon_periodic / lat_periodic clamp them, f_init! starts them at 0)
    v = cs.thr_cmd; if (u(FB_U_ACT_THROTTLE) != v) { u.set(FB_U_ACT_THROTTLE, v); ch = true; }
    v = cs.ail_cmd; if (u(FB_U_ACT_AILERON) != v) { u.set(FB_U_ACT_AILERON, v); ch = true; }
    v = cs.ele_cmd; if (u(FB_U_ACT_ELEVATOR) != v) { u.set(FB_U_ACT_ELEVATOR, v); ch = true; }
    v = cs.rud_cmd; if (u(FB_U_ACT_RUDDER) != v) { u.set(FB_U_ACT_RUDDER, v); ch = true; }
    return ch;
}

// f_periodic!(avionics, vehicle) (guidance in direct mode forwards, then lon, then lat) followed by assign!.
// Returns true when an actuator command changed (only ẋ.act of the pending derivative depends on it).
template <int VAR, int KIN, class UA, class GT>
FB_DEV bool avionics_periodic(const YFull& yf, const RhsOut& y, CtlIn& u, UA& ua,
                              CtlState& cs, const GT& T, double Dt) {
    if (VAR < FB_C172X1) return false;
    gdc_periodic(cs, yf.lat, yf.lon, yf.h_e, y.wow, u);
    GainCell c; c.locate(yf.EAS, yf.h_e);   // both laws schedule on the same (EAS, h_e) (c172x_ctl.jl:296, :889)
    lon_periodic(cs, yf, y.wow, u, T, Dt, (int)u(FB_U_LON_MODE_REQ), c);
    lat_periodic(cs, yf, y.wow, u, T, Dt, (int)u(FB_U_LAT_MODE_REQ), c);
    return avionics_assign(cs, ua);
}

// f_init!(avionics, vehicle): c172x_ctl.jl:1246-1250 -> :463-519 (lon), :998-1032 (lat). `yf` is the vehicle output
// left by f_init!(vehicle, ...). A freshly constructed model is assumed (all compensator memory zero).
FB_DEV void avionics_init(const YFull& v, int wow, const MemberIn& u, CtlState& cs, const FbDeviceTables& T, double Dt) {
    LqrS z; z.io[0] = z.io[1] = 0.0; z.sat[0] = z.sat[1] = 0;
    cs.te2te = z; cs.tv2te = z; cs.vh2te = z; cs.ar2ar = z; cs.pb2ar = z;
    cs.te2te_zref1 = 0.0; cs.pb2ar_zref0 = 0.0;
    int_reset(cs.q2e_int); int_reset(cs.p2phi_int);
    pid_reset(cs.q2e_pid); pid_reset(cs.c2t_pid); pid_reset(cs.v2t_pid); pid_reset(cs.p2phi_pid); pid_reset(cs.chi2phi_pid);
    u.set(FB_U_LON_THROTTLE_AXIS, clampd(v.act_pos[0], 0.0, 1.0)); u.set(FB_U_LON_ELEVATOR_AXIS, clampd(v.act_pos[2], -1.0, 1.0));
    u.set(FB_U_LON_THROTTLE_OFFSET, 0.0); u.set(FB_U_LON_ELEVATOR_OFFSET, 0.0);
    u.set(FB_U_LON_Q_REF, v.w_wb_b.y); u.set(FB_U_LON_THETA_REF, v.theta); u.set(FB_U_LON_EAS_REF, v.EAS);
    u.set(FB_U_LON_CLM_REF, -v.v_eb_n.z); u.set(FB_U_LON_H_REF, v.h_e);
    lon_periodic_ool(cs, v, wow, u, T, Dt, LON_SAS);
    lon_periodic_ool(cs, v, wow, u, T, Dt, LON_THR_EAS);
    lon_periodic_ool(cs, v, wow, u, T, Dt, LON_EAS_ALT);
    lon_periodic_ool(cs, v, wow, u, T, Dt, LON_DIRECT);
    u.set(FB_U_LON_MODE_REQ, (double)LON_DIRECT);
    u.set(FB_U_LAT_AILERON_AXIS, clampd(v.act_pos[1], -1.0, 1.0)); u.set(FB_U_LAT_RUDDER_AXIS, clampd(v.act_pos[3], -1.0, 1.0));
    u.set(FB_U_LAT_AILERON_OFFSET, 0.0); u.set(FB_U_LAT_RUDDER_OFFSET, 0.0);
    u.set(FB_U_LAT_P_REF, v.w_wb_b.x); u.set(FB_U_LAT_PHI_REF, v.phi); u.set(FB_U_LAT_BETA_REF, v.beta); u.set(FB_U_LAT_CHI_REF, v.chi_gnd);
    lat_periodic_ool(cs, v, wow, u, T, Dt, LAT_SAS);
    lat_periodic_ool(cs, v, wow, u, T, Dt, LAT_PHI_BETA);
    lat_periodic_ool(cs, v, wow, u, T, Dt, LAT_DIRECT);
    u.set(FB_U_LAT_MODE_REQ, (double)LAT_DIRECT);
}
