// TEST INFRASTRUCTURE ONLY (see orc_real.hpp).
// Avionics of the C172X variants: discrete PID / Integrator / LQR compensators and the gain-scheduled
// longitudinal / lateral control laws, restating
//   FP/control.jl:123-185 (Integrator), :370-471 (PID), :620-743 (LQR), :950-994 (gain lookups)
//   FA/c172/c172x/control/c172x_ctl.jl:18-24,29-77,97-171 (modes, state / command vectors),
//       :275-446 (ControlLawsLon), :449-519 (assign!, f_init!), :733-751,756-812 (lateral vectors),
//       :858-1032 (ControlLawsLat), :1190-1250 (ControlLaws)
//   FA/c172/c172x/c172x2.jl:18-37 + guidance/c172x_gdc.jl:297-329 (ModeGuidance.direct pass-through)
// in the reference's operation order.
#pragma once
#include <limits>
#include "orc_vehicle.hpp"

// gain schedules: [7 EAS][4 h][P] row-major, from the table blob (data of control/data/*.h5)
struct GainTables {
    enum { TE2TE = 0, TV2TE, VH2TE, AR2AR, PHIBETA2AR, Q2E, C2THETA, V2T, P2PHI, CHI2PHI, COUNT };
    std::vector<double> t[COUNT];
    int P[COUNT] = {0};
    bool loaded = false;
    // extrapolate(scale(interpolate(A, BSpline(Linear())), range(25,55,7), range(50,3050,4)), Flat()) — control.jl:950-973
    void lookup(int which, real EAS, real h, real* out) const {
        real e = rclamp(EAS, 25.0, 55.0), hh = rclamp(h, 50.0, 3050.0);
        real fe = (e - 25.0) / 5.0, fh = (hh - 50.0) / 1000.0;
        int i = (int)std::floor(val(fe)), j = (int)std::floor(val(fh));
        if (i > 5) i = 5; if (j > 2) j = 2; if (i < 0) i = 0; if (j < 0) j = 0;
        real u = fe - (double)i, w = fh - (double)j;
        const int p = P[which];
        const double* a = t[which].data();
        for (int k = 0; k < p; k++) {
            double f00 = a[((i * 4) + j) * p + k], f01 = a[((i * 4) + j + 1) * p + k];
            double f10 = a[(((i + 1) * 4) + j) * p + k], f11 = a[(((i + 1) * 4) + j + 1) * p + k];
            out[k] = (1.0 - u) * ((1.0 - w) * f00 + w * f01) + u * ((1.0 - w) * f10 + w * f11);
        }
    }
};
extern GainTables g_gains;

static const double ORC_INF = std::numeric_limits<double>::infinity();

struct PIDc {   // control.jl:370-471
    real k_p = 1.0, k_i = 0.0, k_d = 0.0, tau_f = 0.01, beta_p = 1.0, beta_d = 1.0;
    double lo = -ORC_INF, hi = ORC_INF;
    real input = 0.0; int sat_ext = 0;
    real x_i0 = 0.0, x_d0 = 0.0; int sat_out_0 = 0;
    real output = 0.0; int sat_out = 0;
    real y_p = 0.0, y_i = 0.0, y_d = 0.0, out_free = 0.0, u_i_y = 0.0; bool int_halted = false;   // the rest of PIDDiscreteY (control.jl:392-407)
    void assign(const real* g) { k_p = g[0]; k_i = g[1]; k_d = g[2]; tau_f = g[3]; }
    void periodic(double Dt) {
        real alpha = 1.0 / (tau_f + Dt);
        real u_p = beta_p * input, u_d = beta_d * input, u_i = input;
        bool halted = (isign(u_i * (double)sat_out_0) > 0) || (isign(u_i * (double)sat_ext) > 0);
        real x_i = x_i0 + Dt * k_i * u_i * (halted ? 0.0 : 1.0);
        real x_d = alpha * tau_f * x_d0 + Dt * alpha * k_d * u_d;
        y_p = k_p * u_p; y_i = x_i; y_d = alpha * (-x_d0 + k_d * u_d);
        out_free = y_p + y_i + y_d;
        u_i_y = u_i; int_halted = halted;
        sat_out = (out_free >= hi) - (out_free <= lo);
        output = rclamp(out_free, lo, hi);
        x_i0 = x_i; x_d0 = x_d; sat_out_0 = sat_out;
    }
    void init(double Dt) { input = 0.0; sat_ext = 0; x_i0 = 0.0; x_d0 = 0.0; sat_out_0 = 0; periodic(Dt); }
};

struct Integ {   // control.jl:123-185
    double lo = -ORC_INF, hi = ORC_INF;
    real input = 0.0; int sat_ext = 0;
    real x0 = 0.0; int sat_out_0 = 0;
    real output = 0.0;
    real x1_y = 0.0; int sat_out_y = 0; bool halted_y = false;   // the rest of IntegratorDiscreteY (control.jl:140-147)
    void periodic(double Dt) {
        bool halted = (isign(input * (double)sat_out_0) > 0) || (isign(input * (double)sat_ext) > 0);
        real x1 = x0 + Dt * input * (halted ? 0.0 : 1.0);
        output = rclamp(x1, lo, hi);
        int sat_out = (x1 >= hi) - (x1 <= lo);
        x1_y = x1; sat_out_y = sat_out; halted_y = halted;
        x0 = x1; sat_out_0 = sat_out;
    }
    void init(double Dt, real x_init = 0.0) { input = 0.0; sat_ext = 0; x0 = x_init; sat_out_0 = 0; periodic(Dt); }
};

template <int NX>
struct LQRc {   // control.jl:620-743, NU = NZ = 2
    real K_fbk[2][NX], K_fwd[2][2], K_int[2][2], x_trim[NX], u_trim[2], z_trim[2];
    double lo[2] = {-ORC_INF, -ORC_INF}, hi[2] = {ORC_INF, ORC_INF};
    int sat_ext[2] = {0, 0}; real z_ref[2], z[2], x[NX];
    real int_out_0[2]; int out_sat_0[2] = {0, 0};
    real output[2]; int out_sat[2] = {0, 0};
    LQRc() {
        for (int i = 0; i < 2; i++) { for (int j = 0; j < NX; j++) K_fbk[i][j] = 0.0; for (int j = 0; j < 2; j++) { K_fwd[i][j] = 0.0; K_int[i][j] = 0.0; }
            u_trim[i] = 0.0; z_trim[i] = 0.0; z_ref[i] = 0.0; z[i] = 0.0; int_out_0[i] = 0.0; output[i] = 0.0; }
        for (int j = 0; j < NX; j++) { x_trim[j] = 0.0; x[j] = 0.0; }
    }
    void assign(const real* g) {   // packed: K_fbk (2 x NX row-major), K_fwd 2x2, K_int 2x2, x_trim NX, u_trim 2, z_trim 2
        int p = 0;
        for (int i = 0; i < 2; i++) for (int j = 0; j < NX; j++) K_fbk[i][j] = g[p++];
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) K_fwd[i][j] = g[p++];
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) K_int[i][j] = g[p++];
        for (int j = 0; j < NX; j++) x_trim[j] = g[p++];
        for (int i = 0; i < 2; i++) u_trim[i] = g[p++];
        for (int i = 0; i < 2; i++) z_trim[i] = g[p++];
    }
    void periodic(double Dt) {
        for (int i = 0; i < 2; i++) {
            real int_in = K_int[i][0] * (z_ref[0] - z[0]) + K_int[i][1] * (z_ref[1] - z[1]);
            bool halted = (isign(int_in * (double)out_sat_0[i]) > 0) || (isign(int_in * (double)sat_ext[i]) > 0);
            real int_out = int_out_0[i] + Dt * int_in * (halted ? 0.0 : 1.0);
            real fwd = K_fwd[i][0] * (z_ref[0] - z_trim[0]) + K_fwd[i][1] * (z_ref[1] - z_trim[1]);
            real fbk = 0.0;
            for (int j = 0; j < NX; j++) fbk = fbk + K_fbk[i][j] * (x[j] - x_trim[j]);
            real out_free = u_trim[i] + int_out + fwd - fbk;
            out_sat[i] = (out_free >= hi[i]) - (out_free <= lo[i]);
            output[i] = rclamp(out_free, lo[i], hi[i]);
            int_out_0[i] = int_out;
        }
        out_sat_0[0] = out_sat[0]; out_sat_0[1] = out_sat[1];
    }
    void init(double Dt) {
        for (int i = 0; i < 2; i++) { z_ref[i] = 0.0; z[i] = 0.0; sat_ext[i] = 0; int_out_0[i] = 0.0; out_sat_0[i] = 0; }
        for (int j = 0; j < NX; j++) x[j] = 0.0;
        periodic(Dt);
    }
};

// ------------------------------------------------------------------------------------------------ guidance
// Segment / SegmentGuidanceData: lib/FlightApps/src/c172/c172x/guidance/c172x_gdc.jl:38-148
struct GeoLLH { LatLon ll; real h; };
struct SegmentData { real chi_12 = 0.0, gamma_12 = 0.0, s_12 = 0.0, s_1b = 0.0, s_2b = 0.0, e_sb = 0.0, v_sb = 0.0, h_s = 0.0; };
inline SegmentData segment_data(GeoLLH p1, GeoLLH p2, GeoLLH ob) {   // c172x_gdc.jl:113-148
    V3 r_e1_e = cartesian_from_geo(n_from_latlon(p1.ll), p1.h);
    V3 r_e2_e = cartesian_from_geo(n_from_latlon(p2.ll), p2.h);
    V3 r_eb_e = cartesian_from_geo(n_from_latlon(ob.ll), ob.h);
    Q4 q_en = ltf(n_from_latlon(ob.ll));
    V3 r_1b_n = qrot_inv(q_en, r_eb_e - r_e1_e);
    V3 r_1b_h(r_1b_n.x, r_1b_n.y, 0.0);
    V3 r_12_n = qrot_inv(q_en, r_e2_e - r_e1_e);
    V3 r_12_h(r_12_n.x, r_12_n.y, 0.0);
    SegmentData d;
    d.s_12 = norm(r_12_h);
    V3 u_12 = r_12_h / d.s_12;
    d.s_1b = dot(u_12, r_1b_h);
    d.s_2b = d.s_1b - d.s_12;
    d.e_sb = cross(u_12, r_1b_h).z;
    d.h_s = p1.h + (p2.h - p1.h) * d.s_1b / d.s_12;
    d.v_sb = ob.h - d.h_s;
    d.chi_12 = azimuth(u_12);
    d.gamma_12 = atan2(p2.h - p1.h, d.s_12);
    return d;
}
// Segment(p1; s, chi, dh): c172x_gdc.jl:64-88 (gamma is turned into dh = s tan(gamma) by the caller)
inline GeoLLH segment_end(GeoLLH p1, real s, real chi, real dh) {
    Q4 q_en1 = ltf(n_from_latlon(p1.ll));
    V3 r_12_e = qrot(q_en1, V3(s * cos(chi), s * sin(chi), 0.0));
    V3 r_e2_e = cartesian_from_geo(n_from_latlon(p1.ll), p1.h) + r_12_e;
    V3 n2; real h2;
    geo_from_cartesian(r_e2_e, n2, h2);
    GeoLLH p2; p2.ll = latlon_from_n(n2); p2.h = p1.h + dh;
    return p2;
}
// the Segment constructor's validity check (:44-58): horizontal length, seen from p1, of at least 1e-6 m
inline bool segment_valid(GeoLLH p1, GeoLLH p2) {
    V3 r_12_e = cartesian_from_geo(n_from_latlon(p2.ll), p2.h) - cartesian_from_geo(n_from_latlon(p1.ll), p1.h);
    V3 r_12_n = qrot_inv(ltf(n_from_latlon(p1.ll)), r_12_e);
    return val(sqrt(r_12_n.x * r_12_n.x + r_12_n.y * r_12_n.y)) >= 1e-6;
}

enum LonMode { LON_DIRECT = 0, LON_SAS, LON_THR_Q, LON_THR_THETA, LON_THR_EAS, LON_EAS_Q, LON_EAS_THETA, LON_EAS_CLM, LON_EAS_ALT };
enum LatMode { LAT_DIRECT = 0, LAT_SAS, LAT_P_BETA, LAT_PHI_BETA, LAT_CHI_BETA };

struct Avionics {
    int variant = VAR_C172S;
    // ---- ControlLawsLon (c172x_ctl.jl:201-274)
    struct { int mode_req = LON_DIRECT; real throttle_axis = 0.0, throttle_offset = 0.0, elevator_axis = 0.0, elevator_offset = 0.0,
             q_ref = 0.0, theta_ref = 0.0, EAS_ref = 50.0, clm_ref = 0.0, h_ref = 0.0; } ulon;
    struct { int mode = LON_DIRECT; real throttle_cmd = 0.0, elevator_cmd = 0.0; } ylon;
    int h_state = 1;   // AltTrackingState.hold
    LQRc<8> te2te, tv2te; LQRc<9> vh2te;
    Integ q2e_int; PIDc q2e_pid, c2t_pid, v2t_pid;
    // ---- ControlLawsLat (c172x_ctl.jl:813-857)
    struct { int mode_req = LAT_DIRECT; real aileron_axis = 0.0, aileron_offset = 0.0, rudder_axis = 0.0, rudder_offset = 0.0,
             p_ref = 0.0, beta_ref = 0.0, phi_ref = 0.0, chi_ref = 0.0; } ulat;
    struct { int mode = LAT_DIRECT; real aileron_cmd = 0.0, rudder_cmd = 0.0; } ylat;
    LQRc<8> ar2ar, pb2ar;
    Integ p2phi_int; PIDc p2phi_pid, chi2phi_pid;
    // ---- GuidanceLaws (Cessna172Xv2 only): c172x_gdc.jl:201-223,264-289
    struct { int mode_req = 0; bool hor_req = false, vrt_req = false;
             GeoLLH p1{{0.0, 0.0}, 0.0}, p2{{1e-3, 0.0}, 0.0}; } ugdc;
    struct { int mode = 0; SegmentData data; real dchi = 0.0, chi_ref = 0.0, h_ref = 0.0; bool hor = false, vrt = false; } ygdc;

    bool active() const { return variant == VAR_C172X1 || variant == VAR_C172X2; }

    // assign!(vehicle, avionics): c172x_ctl.jl:449-458, 986-995 — commands are Ranged on assignment
    void assign(Member& m) {
        if (!active()) return;
        m.u.act_cmd[0] = rclamp(ylon.throttle_cmd, 0.0, 1.0);
        m.u.act_cmd[2] = rclamp(ylon.elevator_cmd, -1.0, 1.0);
        m.u.act_cmd[1] = rclamp(ylat.aileron_cmd, -1.0, 1.0);
        m.u.act_cmd[3] = rclamp(ylat.rudder_cmd, -1.0, 1.0);
    }
    void f_step(Member&) {}   // @no_step

    static bool on_gnd(const Member& m) { return m.y.strut[0].wow || m.y.strut[1].wow || m.y.strut[2].wow; }   // c172.jl:998-1001

    void lon_periodic(const Member& m, double Dt) {   // c172x_ctl.jl:286-446
        const VehicleY& v = m.y;
        real EAS = v.air.EAS, h_e = v.kin.h_e;
        real q = v.kin.w_wb_b.y, r = v.kin.w_wb_b.z;
        real theta = v.kin.e_nb.theta, phi = v.kin.e_nb.phi;
        real clm = -v.kin.v_eb_n.z;
        real h_err = ulon.h_ref - h_e;
        int mode_prev = ylon.mode;
        real throttle_ref = rclamp(ulon.throttle_axis + ulon.throttle_offset, 0.0, 1.0);
        real elevator_ref = rclamp(ulon.elevator_axis + ulon.elevator_offset, -1.0, 1.0);
        real throttle_cmd = throttle_ref, elevator_cmd = elevator_ref;
        real q_ref = ulon.q_ref, theta_ref = ulon.theta_ref;
        int mode;
        const double h_thr = 10.0, h_hys = 1.0, k_p_theta = 1.0;
        if (on_gnd(m)) mode = LON_DIRECT;
        else if (ulon.mode_req == LON_EAS_ALT) {
            if (h_state == 0) {   // acquire
                mode = LON_THR_EAS;
                throttle_ref = h_err > 0.0 ? 1.0 : 0.0;
                if (rabs(h_err) < h_thr - h_hys) h_state = 1;
            } else {
                mode = LON_EAS_ALT;
                if (rabs(h_err) > h_thr + h_hys) h_state = 0;
            }
        } else mode = ulon.mode_req;
        auto te2te_en = [](int md) { return md == LON_SAS || md == LON_THR_Q || md == LON_THR_THETA || md == LON_EAS_Q || md == LON_EAS_THETA || md == LON_EAS_CLM; };
        auto q2e_en = [](int md) { return md == LON_THR_Q || md == LON_THR_THETA || md == LON_EAS_Q || md == LON_EAS_THETA || md == LON_EAS_CLM; };
        auto t2q_en = [](int md) { return md == LON_THR_THETA || md == LON_EAS_THETA || md == LON_EAS_CLM; };
        auto v2t_en = [](int md) { return md == LON_EAS_Q || md == LON_EAS_THETA || md == LON_EAS_CLM; };
        real xred[8] = {v.kin.w_eb_b.y, theta, EAS, v.aero.alpha, v.aero.alpha_filt, v.eng.n, v.act.pos[0], v.act.pos[2]};   // XLonRed :132-147
        real g[64];
        if (te2te_en(mode)) {
            int sat_thr = te2te.out_sat[0], sat_ele = te2te.out_sat[1];
            if (v2t_en(mode)) {
                g_gains.lookup(GainTables::V2T, EAS, h_e, g); v2t_pid.assign(g);
                if (mode != mode_prev) { v2t_pid.init(Dt); if (v2t_pid.k_i != 0.0) v2t_pid.x_i0 = rclamp(ylon.throttle_cmd, 0.0, 1.0); }
                v2t_pid.input = ulon.EAS_ref - EAS; v2t_pid.sat_ext = sat_thr;
                v2t_pid.periodic(Dt);
                throttle_ref = v2t_pid.output;
            }
            if (q2e_en(mode)) {
                g_gains.lookup(GainTables::Q2E, EAS, h_e, g); q2e_pid.assign(g);
                if (mode != mode_prev) { q2e_int.init(Dt); q2e_pid.init(Dt); if (q2e_pid.k_i != 0.0) q2e_pid.x_i0 = te2te.z_ref[1]; }
                if (t2q_en(mode)) {
                    if (mode == LON_EAS_CLM) {
                        g_gains.lookup(GainTables::C2THETA, EAS, h_e, g); c2t_pid.assign(g);
                        if (mode != mode_prev) { c2t_pid.init(Dt); if (c2t_pid.k_i != 0.0) c2t_pid.x_i0 = theta; }
                        c2t_pid.input = ulon.clm_ref - clm; c2t_pid.sat_ext = sat_ele;
                        c2t_pid.periodic(Dt);
                        theta_ref = c2t_pid.output;
                    }
                    real theta_dot_ref = k_p_theta * (theta_ref - theta);
                    real phi_bnd = rclamp(phi, -ORC_PI / 3, ORC_PI / 3);
                    q_ref = 1.0 / cos(phi_bnd) * theta_dot_ref + r * tan(phi_bnd);
                }
                q2e_int.input = q_ref - q; q2e_int.sat_ext = sat_ele;
                q2e_int.periodic(Dt);
                q2e_pid.input = q2e_int.output; q2e_pid.sat_ext = sat_ele;
                q2e_pid.periodic(Dt);
                elevator_ref = q2e_pid.output;
            }
            g_gains.lookup(GainTables::TE2TE, EAS, h_e, g); te2te.assign(g);
            for (int i = 0; i < 8; i++) te2te.x[i] = xred[i];
            te2te.z[0] = v.act.cmd[0]; te2te.z[1] = v.act.cmd[2];          // Zte(vehicle) :167-171
            te2te.z_ref[0] = throttle_ref; te2te.z_ref[1] = elevator_ref;
            te2te.periodic(Dt);
            throttle_cmd = te2te.output[0]; elevator_cmd = te2te.output[1];
        }
        if (mode == LON_THR_EAS) {
            g_gains.lookup(GainTables::TV2TE, EAS, h_e, g); tv2te.assign(g);
            if (mode != mode_prev) tv2te.init(Dt);
            for (int i = 0; i < 8; i++) tv2te.x[i] = xred[i];
            tv2te.z[0] = v.act.cmd[0]; tv2te.z[1] = EAS;                    // Ztv :190-194
            tv2te.z_ref[0] = throttle_ref; tv2te.z_ref[1] = ulon.EAS_ref;
            tv2te.periodic(Dt);
            throttle_cmd = tv2te.output[0]; elevator_cmd = tv2te.output[1];
        }
        if (mode == LON_EAS_ALT) {
            g_gains.lookup(GainTables::VH2TE, EAS, h_e, g); vh2te.assign(g);
            if (mode != mode_prev) vh2te.init(Dt);
            real xfull[9] = {v.kin.w_eb_b.y, theta, EAS, v.aero.alpha, h_e, v.aero.alpha_filt, v.eng.n, v.act.pos[0], v.act.pos[2]};   // XLonFull :97-118
            for (int i = 0; i < 9; i++) vh2te.x[i] = xfull[i];
            vh2te.z[0] = EAS; vh2te.z[1] = h_e;                              // Zvh :178-182
            vh2te.z_ref[0] = ulon.EAS_ref; vh2te.z_ref[1] = ulon.h_ref;
            vh2te.periodic(Dt);
            throttle_cmd = vh2te.output[0]; elevator_cmd = vh2te.output[1];
        }
        ylon.mode = mode;
        ylon.throttle_cmd = rclamp(throttle_cmd, 0.0, 1.0);
        ylon.elevator_cmd = rclamp(elevator_cmd, -1.0, 1.0);
    }

    void lat_periodic(const Member& m, double Dt) {   // c172x_ctl.jl:880-983
        const VehicleY& v = m.y;
        real EAS = v.air.EAS, h_e = v.kin.h_e;
        real p = v.kin.w_wb_b.x;
        int mode_prev = ylat.mode;
        int mode = on_gnd(m) ? LAT_DIRECT : ulat.mode_req;
        real aileron_ref = rclamp(ulat.aileron_axis + ulat.aileron_offset, -1.0, 1.0);
        real rudder_ref = rclamp(ulat.rudder_axis + ulat.rudder_offset, -1.0, 1.0);
        real aileron_cmd = aileron_ref, rudder_cmd = rudder_ref;
        real phi_ref = ulat.phi_ref;
        real xlat[8] = {v.kin.w_eb_b.x, v.kin.w_eb_b.z, v.kin.e_nb.phi, EAS, v.aero.beta, v.aero.beta_filt, v.act.pos[1], v.act.pos[3]};   // XLatRed :767-779
        real g[64];
        if (mode == LAT_SAS) {
            g_gains.lookup(GainTables::AR2AR, EAS, h_e, g); ar2ar.assign(g);
            for (int i = 0; i < 8; i++) ar2ar.x[i] = xlat[i];
            ar2ar.z[0] = v.act.cmd[1]; ar2ar.z[1] = v.act.cmd[3];           // Zar :806-810
            ar2ar.z_ref[0] = aileron_ref; ar2ar.z_ref[1] = rudder_ref;
            ar2ar.periodic(Dt);
            aileron_cmd = ar2ar.output[0]; rudder_cmd = ar2ar.output[1];
        }
        if (mode == LAT_P_BETA || mode == LAT_PHI_BETA || mode == LAT_CHI_BETA) {
            int sat_ail = pb2ar.out_sat[0];
            if (mode == LAT_P_BETA) {
                g_gains.lookup(GainTables::P2PHI, EAS, h_e, g); p2phi_pid.assign(g);
                if (mode != mode_prev) { p2phi_int.init(Dt); p2phi_pid.init(Dt); if (p2phi_pid.k_i != 0.0) p2phi_pid.x_i0 = pb2ar.z_ref[0]; }
                p2phi_int.input = ulat.p_ref - p; p2phi_int.sat_ext = sat_ail;
                p2phi_int.periodic(Dt);
                p2phi_pid.input = p2phi_int.output; p2phi_pid.sat_ext = sat_ail;
                p2phi_pid.periodic(Dt);
                phi_ref = p2phi_pid.output;
            } else if (mode == LAT_CHI_BETA) {
                g_gains.lookup(GainTables::CHI2PHI, EAS, h_e, g); chi2phi_pid.assign(g);
                if (mode != mode_prev) { chi2phi_pid.init(Dt); if (chi2phi_pid.k_i != 0.0) chi2phi_pid.x_i0 = pb2ar.z_ref[0]; }
                chi2phi_pid.input = wrap_to_pi(ulat.chi_ref - v.kin.chi_gnd); chi2phi_pid.sat_ext = sat_ail;
                chi2phi_pid.periodic(Dt);
                phi_ref = chi2phi_pid.output;
            }
            g_gains.lookup(GainTables::PHIBETA2AR, EAS, h_e, g); pb2ar.assign(g);
            if (mode != mode_prev) pb2ar.init(Dt);
            for (int i = 0; i < 8; i++) pb2ar.x[i] = xlat[i];
            pb2ar.z[0] = v.kin.e_nb.phi; pb2ar.z[1] = v.aero.beta;          // Zphibeta :789-793
            pb2ar.z_ref[0] = phi_ref; pb2ar.z_ref[1] = ulat.beta_ref;
            pb2ar.periodic(Dt);
            aileron_cmd = pb2ar.output[0]; rudder_cmd = pb2ar.output[1];
        }
        ylat.mode = mode;
        ylat.aileron_cmd = rclamp(aileron_cmd, -1.0, 1.0);
        ylat.rudder_cmd = rclamp(rudder_cmd, -1.0, 1.0);
    }

    // f_periodic!(GuidanceLaws, ctl, vehicle): c172x_gdc.jl:297-329 -> SegmentGuidance :228-252 (dchi_inf = pi/2,
    // e_sf = 250 m, e_thr = 1000 m, :203-207). Writes the control laws' inputs when a guidance channel is active.
    void gdc_periodic(const Member& m) {
        const int mode = on_gnd(m) ? 0 : ugdc.mode_req;
        if (mode == 1) {
            GeoLLH ob; ob.ll = m.y.kin.ll; ob.h = m.y.kin.h_e;
            SegmentData d = segment_data(ugdc.p1, ugdc.p2, ob);
            const real dchi_inf = 0.5 * ORC_PI;
            real dchi = -dchi_inf / (0.5 * ORC_PI) * atan(d.e_sb / 250.0);
            ygdc.data = d; ygdc.dchi = dchi;
            ygdc.chi_ref = wrap_to_pi(d.chi_12 + dchi);
            ygdc.h_ref = d.h_s;
            ygdc.hor = ugdc.hor_req;
            ygdc.vrt = (rabs(d.e_sb) < 1000.0) ? ugdc.vrt_req : false;
            if (ygdc.hor) { ulat.chi_ref = ygdc.chi_ref; ulat.mode_req = LAT_CHI_BETA; }
            if (ygdc.vrt) { ulon.h_ref = ygdc.h_ref; ulon.mode_req = LON_EAS_ALT; }
        }
        ygdc.mode = mode;   // circular guidance is a stub in the reference (:257-262)
    }

    // f_periodic!(avionics, vehicle): c172x2.jl:28-37 (guidance, then the control laws), lon then lat (@sm_periodic)
    void f_periodic(Member& m, const World&, double Dt) {
        if (!active()) return;
        if (variant == VAR_C172X2) gdc_periodic(m);
        lon_periodic(m, Dt);
        lat_periodic(m, Dt);
    }

    // f_init!(avionics, vehicle): c172x_ctl.jl:1246-1250 -> :463-519 (lon), :998-1032 (lat)
    void f_init(Member& m, const World& W, double Dt = 0.0) {
        if (!active()) return;
        const VehicleY& v = m.y;
        // f_init!(lon): every submodel, then the LQR bounds (:275-284)
        te2te.init(Dt); tv2te.init(Dt); vh2te.init(Dt); q2e_int.init(Dt); q2e_pid.init(Dt); c2t_pid.init(Dt); v2t_pid.init(Dt);
        LQRc<8>* l8[2] = {&te2te, &tv2te};
        for (auto* l : l8) { l->lo[0] = 0; l->lo[1] = -1; l->hi[0] = 1; l->hi[1] = 1; }
        vh2te.lo[0] = 0; vh2te.lo[1] = -1; vh2te.hi[0] = 1; vh2te.hi[1] = 1;
        ulon.throttle_axis = rclamp(v.act.pos[0], 0.0, 1.0); ulon.elevator_axis = rclamp(v.act.pos[2], -1.0, 1.0);
        ulon.throttle_offset = 0.0; ulon.elevator_offset = 0.0;
        ulon.q_ref = v.kin.w_wb_b.y; ulon.theta_ref = v.kin.e_nb.theta; ulon.EAS_ref = v.air.EAS; ulon.clm_ref = -v.kin.v_eb_n.z; ulon.h_ref = v.kin.h_e;
        const int lon_seq[4] = {LON_SAS, LON_THR_EAS, LON_EAS_ALT, LON_DIRECT};
        for (int md : lon_seq) { ulon.mode_req = md; lon_periodic(m, Dt); }
        // f_init!(lat): :868-878
        ar2ar.init(Dt); pb2ar.init(Dt); p2phi_int.init(Dt); p2phi_pid.init(Dt); chi2phi_pid.init(Dt);
        LQRc<8>* la[2] = {&pb2ar, &ar2ar};
        for (auto* l : la) { l->lo[0] = -1; l->lo[1] = -1; l->hi[0] = 1; l->hi[1] = 1; }
        chi2phi_pid.lo = -ORC_PI / 4; chi2phi_pid.hi = ORC_PI / 4;
        ulat.aileron_axis = rclamp(v.act.pos[1], -1.0, 1.0); ulat.rudder_axis = rclamp(v.act.pos[3], -1.0, 1.0);
        ulat.aileron_offset = 0.0; ulat.rudder_offset = 0.0;
        ulat.p_ref = v.kin.w_wb_b.x; ulat.phi_ref = v.kin.e_nb.phi; ulat.beta_ref = v.aero.beta; ulat.chi_ref = v.kin.chi_gnd;
        const int lat_seq[3] = {LAT_SAS, LAT_PHI_BETA, LAT_DIRECT};
        for (int md : lat_seq) { ulat.mode_req = md; lat_periodic(m, Dt); }
        (void)W;
    }

    bool set_input(int f, double v) {
        auto c01 = [](double x) { return x < 0 ? 0.0 : (x > 1 ? 1.0 : x); };
        auto c11 = [](double x) { return x < -1 ? -1.0 : (x > 1 ? 1.0 : x); };
        switch (f) {
            case FB_U_LON_MODE_REQ: ulon.mode_req = (int)v; return true;
            case FB_U_LON_THROTTLE_AXIS: ulon.throttle_axis = c01(v); return true;
            case FB_U_LON_THROTTLE_OFFSET: ulon.throttle_offset = c01(v); return true;
            case FB_U_LON_ELEVATOR_AXIS: ulon.elevator_axis = c11(v); return true;
            case FB_U_LON_ELEVATOR_OFFSET: ulon.elevator_offset = c11(v); return true;
            case FB_U_LON_Q_REF: ulon.q_ref = v; return true;
            case FB_U_LON_THETA_REF: ulon.theta_ref = v; return true;
            case FB_U_LON_EAS_REF: ulon.EAS_ref = v; return true;
            case FB_U_LON_CLM_REF: ulon.clm_ref = v; return true;
            case FB_U_LON_H_REF: ulon.h_ref = v; return true;
            case FB_U_LAT_MODE_REQ: ulat.mode_req = (int)v; return true;
            case FB_U_LAT_AILERON_AXIS: ulat.aileron_axis = c11(v); return true;
            case FB_U_LAT_AILERON_OFFSET: ulat.aileron_offset = c11(v); return true;
            case FB_U_LAT_RUDDER_AXIS: ulat.rudder_axis = c11(v); return true;
            case FB_U_LAT_RUDDER_OFFSET: ulat.rudder_offset = c11(v); return true;
            case FB_U_LAT_P_REF: ulat.p_ref = v; return true;
            case FB_U_LAT_BETA_REF: ulat.beta_ref = v; return true;
            case FB_U_LAT_PHI_REF: ulat.phi_ref = v; return true;
            case FB_U_LAT_CHI_REF: ulat.chi_ref = v; return true;
            case FB_U_GDC_MODE_REQ: ugdc.mode_req = (int)v; return true;
            case FB_U_GDC_HOR_REQ: ugdc.hor_req = v != 0.0; return true;
            case FB_U_GDC_VRT_REQ: ugdc.vrt_req = v != 0.0; return true;
            case FB_U_GDC_LAT1: ugdc.p1.ll.lat = v; return true; case FB_U_GDC_LON1: ugdc.p1.ll.lon = v; return true; case FB_U_GDC_H1: ugdc.p1.h = v; return true;
            case FB_U_GDC_LAT2: ugdc.p2.ll.lat = v; return true; case FB_U_GDC_LON2: ugdc.p2.ll.lon = v; return true; case FB_U_GDC_H2: ugdc.p2.h = v; return true;
            default: return false;
        }
    }
    double get_input(int f) const {
        switch (f) {
            case FB_U_LON_MODE_REQ: return ulon.mode_req; case FB_U_LON_THROTTLE_AXIS: return val(ulon.throttle_axis);
            case FB_U_LON_THROTTLE_OFFSET: return val(ulon.throttle_offset); case FB_U_LON_ELEVATOR_AXIS: return val(ulon.elevator_axis);
            case FB_U_LON_ELEVATOR_OFFSET: return val(ulon.elevator_offset); case FB_U_LON_Q_REF: return val(ulon.q_ref);
            case FB_U_LON_THETA_REF: return val(ulon.theta_ref); case FB_U_LON_EAS_REF: return val(ulon.EAS_ref);
            case FB_U_LON_CLM_REF: return val(ulon.clm_ref); case FB_U_LON_H_REF: return val(ulon.h_ref);
            case FB_U_LAT_MODE_REQ: return ulat.mode_req; case FB_U_LAT_AILERON_AXIS: return val(ulat.aileron_axis);
            case FB_U_LAT_AILERON_OFFSET: return val(ulat.aileron_offset); case FB_U_LAT_RUDDER_AXIS: return val(ulat.rudder_axis);
            case FB_U_LAT_RUDDER_OFFSET: return val(ulat.rudder_offset); case FB_U_LAT_P_REF: return val(ulat.p_ref);
            case FB_U_LAT_BETA_REF: return val(ulat.beta_ref); case FB_U_LAT_PHI_REF: return val(ulat.phi_ref);
            case FB_U_LAT_CHI_REF: return val(ulat.chi_ref); case FB_U_GDC_MODE_REQ: return ugdc.mode_req;
            case FB_U_GDC_HOR_REQ: return ugdc.hor_req; case FB_U_GDC_VRT_REQ: return ugdc.vrt_req;
            case FB_U_GDC_LAT1: return val(ugdc.p1.ll.lat); case FB_U_GDC_LON1: return val(ugdc.p1.ll.lon); case FB_U_GDC_H1: return val(ugdc.p1.h);
            case FB_U_GDC_LAT2: return val(ugdc.p2.ll.lat); case FB_U_GDC_LON2: return val(ugdc.p2.ll.lon); case FB_U_GDC_H2: return val(ugdc.p2.h);
            default: return 0.0 / 0.0;
        }
    }
    double output(const Member&, int f) const {
        switch (f) {
            case FB_Y_LON_MODE: return ylon.mode; case FB_Y_LAT_MODE: return ylat.mode;
            case FB_Y_THROTTLE_CMD: return val(ylon.throttle_cmd); case FB_Y_ELEVATOR_CMD: return val(ylon.elevator_cmd);
            case FB_Y_AILERON_CMD: return val(ylat.aileron_cmd); case FB_Y_RUDDER_CMD: return val(ylat.rudder_cmd);
            case FB_Y_ALT_STATE: return h_state;
            case FB_Y_GDC_MODE: return ygdc.mode; case FB_Y_GDC_HOR: return ygdc.hor; case FB_Y_GDC_VRT: return ygdc.vrt;
            case FB_Y_GDC_DCHI: return val(ygdc.dchi); case FB_Y_GDC_CHI_REF: return val(ygdc.chi_ref); case FB_Y_GDC_H_REF: return val(ygdc.h_ref);
            case FB_Y_GDC_CHI_12: return val(ygdc.data.chi_12); case FB_Y_GDC_GAMMA_12: return val(ygdc.data.gamma_12); case FB_Y_GDC_S_12: return val(ygdc.data.s_12);
            case FB_Y_GDC_S_1B: return val(ygdc.data.s_1b); case FB_Y_GDC_E_SB: return val(ygdc.data.e_sb); case FB_Y_GDC_V_SB: return val(ygdc.data.v_sb);
            default: return 0.0 / 0.0;
        }
    }
};
