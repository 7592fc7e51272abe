// TEST INFRASTRUCTURE ONLY (see orc_real.hpp).
// One Cessna 172 `SimpleWorld`: continuous RHS f_ode!, discrete f_step!, restating
//   FP/kinematics.jl:150-425, FP/dynamics.jl:190-317,443-525, FP/landinggear.jl:26-552,
//   FP/propellers.jl:405-452, FP/piston.jl:299-453,575-603, FP/control.jl:52-88,
//   FA/c172/c172.jl:26-44,226-245,307-384,442-476,513-641,697-724,
//   FA/c172/c172s/c172s.jl:92-120, FA/c172/c172x/c172x.jl:39-52,127-143,
//   FP/aircraftbase.jl:142-181, FP/world.jl:26-39
// in the reference's evaluation order.
#pragma once
#include "orc_geo_atm.hpp"
#include "orc_tables.hpp"

enum KinType { KIN_WA = 0, KIN_ECEF = 1, KIN_NED = 2 };
enum Variant { VAR_C172S = 0, VAR_C172X0 = 1, VAR_C172X1 = 2, VAR_C172X2 = 3,
               // Cessna172Xv0 assembled from the non-default components the Model tree admits: a
               // FlyByWireActuation of seven Actuator2 units (c172x.jl:68-106) and, optionally, a VariablePitch propeller
               // (propellers.jl:341-400)
               VAR_C172XA2 = 4 };
enum EngState { ENG_OFF = 0, ENG_STARTING = 1, ENG_RUNNING = 2 };
enum Surface { SURF_DRY = 0, SURF_WET = 1, SURF_ICY = 2 };

struct Wrench { V3 F, tau; };
inline Wrench operator+(Wrench a, Wrench b) { Wrench w; w.F = a.F + b.F; w.tau = a.tau + b.tau; return w; }
struct Frame { V3 r; Q4 q; };  // FrameTransform
inline Wrench translate(const Frame& t, const Wrench& w) {  // dynamics.jl:117-128
    Wrench o;
    V3 F = qrot(t.q, w.F);
    V3 tau = qrot(t.q, w.tau);
    o.F = F;
    o.tau = tau + cross(t.r, F);
    return o;
}
struct MassProps { real m; M3 J; V3 r; MassProps() : m(0.0) {} };
inline MassProps mp_point(real m, V3 r) {  // dynamics.jl:190-193
    MassProps p;
    M3 s = v2skew(r);
    p.m = m; p.J = (-m) * (s * s); p.r = r;
    return p;
}
inline MassProps mp_add(const MassProps& a, const MassProps& b) {  // dynamics.jl:231-241
    MassProps p;
    real m = a.m + b.m;
    if (m > 0.0) {
        p.m = m; p.J = a.J + b.J;
        p.r = (1.0 / m) * (a.m * a.r + b.m * b.r);
    }
    return p;
}
inline MassProps mp_translate(V3 r_bc_b, const MassProps& c) {  // dynamics.jl:254-272, q_bc = identity
    MassProps p;
    M3 s = v2skew(c.r);
    M3 J_G_c = c.J + c.m * (s * s);
    V3 r_cG_b = qrot(Q4(), c.r);
    V3 r_bG_b = r_bc_b + r_cG_b;
    M3 s2 = v2skew(r_bG_b);
    p.m = c.m; p.J = J_G_c - c.m * (s2 * s2); p.r = r_bG_b;
    return p;
}

struct KinData {
    Euler e_nb; Q4 q_nb, q_eb, q_en; LatLon ll; V3 n_e; real h_e, h_o; V3 r_eb_e, w_wb_b, w_eb_b, v_eb_b, v_eb_n;
    real v_gnd, chi_gnd, gamma_gnd;
};
struct PIY { real out[2]; real out_free[2]; int sat_out[2]; bool halted[2]; };
struct StrutY { real dh = 0.0; bool wow = false; real xi = 0.0, xi_dot = 0.0, F_dmp = 0.0, psi_sw = 0.0, alpha_ts = 0.0; Frame t_sc, t_bc; real v_xy[2] = {0.0, 0.0}; int surface = SURF_DRY; };
struct ContactY { real mu_roll = 0.0, mu_skid = 0.0, kappa_br = 0.0, psi_cv = 0.0; real mu_max[2] = {0.0, 0.0}, mu_eff[2] = {0.0, 0.0}; V3 f_c, F_c; Wrench wr_b; PIY frc; };
struct AeroY { real e, a, r, f, de, da, dr, df, alpha, beta, alpha_filt, beta_filt, alpha_filt_dot, beta_filt_dot; bool stall; real C_D, C_Y, C_L, C_l, C_m, C_n; Wrench wr_b; };
struct PropY { V3 v_wOp_p; real omega, J, Mt, dbeta = 0.0; Wrench wr_p, wr_b; V3 hr_p, hr_b; real P, eta_p; };
struct EngY { bool start, stop; int state; real throttle, MAP, mixture, mixture_pos, f, mdot, omega, n, tau_shaft, P_shaft, SFC; PIY idle, frc; };
struct FuelY { real x_avail, m_total, m_avail; };
struct ActY { real cmd[7], pos[7], vel[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}; int sat[7]; };  // throttle, aileron, elevator, rudder, flaps, brake_left, brake_right
struct DynY { Wrench wr_c, wr_b; MassProps mp_c, mp_b; V3 ho_b, wd_ec_c, vd_ec_c, a_ec_c, a_ic_c, g_c_c, gam_c_c, f_c_c, wd_eb_b, vd_eb_b, al_ib_b, a_eb_b, a_ib_b; };
struct VehicleY { KinData kin; AirData air; AeroY aero; StrutY strut[3]; ContactY contact[3]; PropY prop; EngY eng; FuelY fuel; ActY act; DynY dyn; };

// state vector layout: systems{aero 2, ldg 6, fuel 1, engine 3, [act 7]}, kinematics, dynamics
struct Layout {
    int variant, kin;
    int aero = 0, ldg = 2, fuel = 8, eng = 9, act = 12, kinx, dyn, nx, nkin;
    Layout(int v = VAR_C172S, int k = KIN_WA) : variant(v), kin(k) {
        int nact = (v == VAR_C172S) ? 0 : (v == VAR_C172XA2 ? 14 : 7);   // Actuator2: X = (v, p) per unit (c172x.jl:84)
        kinx = 12 + nact;
        nkin = (k == KIN_WA) ? 9 : (k == KIN_ECEF ? 8 : 6);
        dyn = kinx + nkin;
        nx = dyn + 6;
    }
};
static const int ORC_NX_MAX = 41;

struct Inputs {
    // C172S MechanicalActuation.u (c172s.jl:62-72); every assignment saturates (types.jl:15-20)
    real aileron = 0.0, elevator = 0.0, rudder = 0.0, aileron_offset = 0.0, elevator_offset = 0.0, rudder_offset = 0.0;
    real flaps = 0.0, brake_left = 0.0, brake_right = 0.0;
    // engine.u (piston.jl:259-267)
    bool eng_start = false, eng_stop = false;
    real throttle = 0.0; int mixture_auto = 1; real mixture = 0.5;
    bool steering_engaged = true;
    // payload.u (c172.jl:521-527)
    real m_pilot = 75.0, m_copilot = 75.0, m_lpass = 0.0, m_rpass = 0.0, m_baggage = 50.0;
    // atmosphere / terrain
    real T_sl = isa::T_std, p_sl = isa::p_std; V3 wind_n;
    real trn_elev = 0.0; int surface = SURF_DRY;
    // C172X fly-by-wire actuator commands (c172x.jl:33); index order as ActY
    real act_cmd[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    // Propeller{<:VariablePitch}.u: Ranged(0, 0, 1) (propellers.jl:385)
    real prop_pitch = 0.0;
};
struct Discrete { bool stall = false; int eng_state = ENG_OFF; };

static const double ACT_LO[7] = {0.0, -1.0, -1.0, -1.0, 0.0, 0.0, 0.0};
static const double ACT_HI[7] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0};

struct World {
    const Geoid* geoid;
    AeroTables aero;
    EngineTables eng;
    PropTables prop;
    const PropTables3* prop3 = nullptr;   // set: the propeller is Propeller(VariablePitch(prop3->lo, prop3->hi)) instead of FixedPitch
    // Actuator2 parameters (c172x.jl:71: defaults 5 Hz natural frequency, damping ratio 0.6)
    double act2_wn = 5 * 2 * ORC_PI, act2_zeta = 0.6;
    // engine parameters (c172s.jl:16-27)
    double P_rated = 735.49875 * 200, w_rated = 2700 * ORC_PI / 30, w_stall = 300 * ORC_PI / 30, w_max = 3100 * ORC_PI / 30,
           w_idle = 600 * ORC_PI / 30, tau_start = 40, J_eng = 0.05;
    MassProps afm;  // const mp_b_afm (c172.jl:26-35), evaluated once like the reference's module constant
    explicit World(const Geoid* g) : geoid(g), eng(300.0 / 2700.0, 3100.0 / 2700.0) {
        V3 r(0.056, 0.0, 0.582);
        M3 Jc; Jc.m[0][0] = 820.0; Jc.m[1][1] = 1164.0; Jc.m[2][2] = 1702.0;
        M3 sk = v2skew(r);
        afm.m = 767.0; afm.J = Jc - 767.0 * (sk * sk); afm.r = r;  // dynamics.jl:209-225, q_bc = identity
    }
};

// continuous PI compensator with anti-windup, FP/control.jl:52-81 (scalar lane)
inline void pi_lane(real input, int sat_ext, real x_i, double k_p, double k_i, double k_l, double lo, double hi,
                    real& xdot, real& output, real& out_free, int& sat_out, bool& halted) {
    real u_p = 1.0 * input;
    real u_i = input;
    real y_p = k_p * u_p;
    real y_i = x_i;
    out_free = y_p + y_i;
    output = rclamp(out_free, lo, hi);
    int sat_hi = out_free >= hi, sat_lo = out_free <= lo;
    sat_out = sat_hi - sat_lo;
    halted = (isign(u_i * (double)sat_out) > 0) || (isign(u_i * (double)sat_ext) > 0);
    xdot = k_i * u_i * (halted ? 0.0 : 1.0) - k_l * x_i;
}

inline real get_mu(double mu_s, double mu_d, double v_s, double v_d, real v) {  // landinggear.jl:170-174
    real k = rclamp((rabs(v) - v_s) / (v_d - v_s), 0.0, 1.0);
    return k * mu_d + (1.0 - k) * mu_s;
}

struct Member {
    Layout L;
    real x[ORC_NX_MAX];
    real xdot[ORC_NX_MAX];
    Inputs u;
    Discrete s;
    VehicleY y;
    int status = ST_RUNNING;
    bool want_y_extra = false;  // compute dynamics.jl:502-513 (outputs only; excluded from the FLOP count)

    Member() : L() { for (int i = 0; i < ORC_NX_MAX; i++) { x[i] = 0.0; xdot[i] = 0.0; } }

    // ------------------------------------------------------------------ kinematics
    void kin_ode(const World& W) {
        KinData& k = y.kin;
        const real* xk = x + L.kinx; real* xd = xdot + L.kinx;
        V3 w_eb_b(x[L.dyn], x[L.dyn + 1], x[L.dyn + 2]);      // kinematics.u .= dynamics.x
        V3 v_eb_b(x[L.dyn + 3], x[L.dyn + 4], x[L.dyn + 5]);
        if (L.kin == KIN_WA) {  // kinematics.jl:181-223
            Q4 q_wb(xk[0], xk[1], xk[2], xk[3]), q_ew(xk[4], xk[5], xk[6], xk[7]);
            real h_e = xk[8]; check_alt(h_e);
            real psi_nw = get_psi_nw(q_ew);
            Q4 q_nw = Rz(psi_nw);
            Q4 q_nb = qmul(q_nw, q_wb);
            Q4 q_eb = qmul(q_ew, q_wb);
            Q4 q_en = qmul(q_eb, qconj(q_nb));
            Euler e_nb = euler_from_q(q_nb);
            V3 n_e = n_from_q(q_ew);
            LatLon ll = latlon_from_n(n_e);
            real h_o = h_e - W.geoid->height(n_e); check_alt(h_o);
            V3 v_eb_n = qrot(q_nb, v_eb_b);
            V3 r_eb_e = cartesian_from_geo(n_e, h_e);
            real M, N; radii(n_e, M, N);
            V3 w_ew_n(v_eb_n.y / (N + h_e), -v_eb_n.x / (M + h_e), 0.0);
            V3 w_ew_w = qrot_inv(q_nw, w_ew_n);
            V3 w_ew_b = qrot_inv(q_wb, w_ew_w);
            V3 w_wb_b = w_eb_b - w_ew_b;
            real v_gnd = norm(v_eb_n);
            real chi = 0.0, gam = 0.0;
            if (v_gnd > 0.1) { chi = azimuth(v_eb_n); gam = inclination(v_eb_n); }
            Q4 d1 = qdt(q_wb, w_wb_b), d2 = qdt(q_ew, w_ew_w);
            xd[0] = d1.w; xd[1] = d1.x; xd[2] = d1.y; xd[3] = d1.z;
            xd[4] = d2.w; xd[5] = d2.x; xd[6] = d2.y; xd[7] = d2.z;
            xd[8] = -v_eb_n.z;
            k.e_nb = e_nb; k.q_nb = q_nb; k.q_eb = q_eb; k.q_en = q_en; k.ll = ll; k.n_e = n_e; k.h_e = h_e; k.h_o = h_o;
            k.r_eb_e = r_eb_e; k.w_wb_b = w_wb_b; k.w_eb_b = w_eb_b; k.v_eb_b = v_eb_b; k.v_eb_n = v_eb_n;
            k.v_gnd = v_gnd; k.chi_gnd = chi; k.gamma_gnd = gam;
        } else if (L.kin == KIN_ECEF) {  // kinematics.jl:281-315
            Q4 q_eb(xk[0], xk[1], xk[2], xk[3]);
            V3 n_e(xk[4], xk[5], xk[6]);
            real h_e = xk[7]; check_alt(h_e);
            real h_o = h_e - W.geoid->height(n_e); check_alt(h_o);
            LatLon ll = latlon_from_n(n_e);
            Q4 q_en = ltf(n_e);
            Q4 q_nb = qmul(qconj(q_en), q_eb);
            Euler e_nb = euler_from_q(q_nb);
            V3 r_eb_e = cartesian_from_geo(n_e, h_e);
            V3 v_eb_n = qrot(q_nb, v_eb_b);
            real M, N; radii(n_e, M, N);
            V3 w_ew_n(v_eb_n.y / (N + h_e), -v_eb_n.x / (M + h_e), 0.0);
            V3 w_ew_b = qrot_inv(q_nb, w_ew_n);
            V3 w_wb_b = w_eb_b - w_ew_b;
            real v_gnd = norm(v_eb_n);
            real chi = 0.0, gam = 0.0;
            if (v_gnd > 0.1) { chi = azimuth(v_eb_n); gam = inclination(v_eb_n); }
            Q4 d1 = qdt(q_eb, w_eb_b);
            V3 nd = qrot(q_en, cross(w_ew_n, V3(0.0, 0.0, -1.0)));
            xd[0] = d1.w; xd[1] = d1.x; xd[2] = d1.y; xd[3] = d1.z;
            xd[4] = nd.x; xd[5] = nd.y; xd[6] = nd.z;
            xd[7] = -v_eb_n.z;
            k.e_nb = e_nb; k.q_nb = q_nb; k.q_eb = q_eb; k.q_en = q_en; k.ll = ll; k.n_e = n_e; k.h_e = h_e; k.h_o = h_o;
            k.r_eb_e = r_eb_e; k.w_wb_b = w_wb_b; k.w_eb_b = w_eb_b; k.v_eb_b = v_eb_b; k.v_eb_n = v_eb_n;
            k.v_gnd = v_gnd; k.chi_gnd = chi; k.gamma_gnd = gam;
        } else {  // NED, kinematics.jl:363-409
            Euler e_nb(xk[0], xk[1], xk[2]);
            LatLon ll; ll.lat = xk[3]; ll.lon = xk[4];
            if (rabs(ll.lat) > 0.5 * ORC_PI || rabs(ll.lon) > ORC_PI) throw OrcTermination(ST_ALT_BOUNDS);  // geodesy.jl:83-86
            real h_e = xk[5]; check_alt(h_e);
            V3 n_e = n_from_latlon(ll);
            real h_o = h_e - W.geoid->height(n_e); check_alt(h_o);
            Q4 q_nb = q_from_euler(e_nb);
            Q4 q_en = ltf(n_e);
            Q4 q_eb = qmul(q_en, q_nb);
            V3 v_eb_n = qrot(q_nb, v_eb_b);
            V3 r_eb_e = cartesian_from_geo(n_e, h_e);
            real M, N; radii(n_e, M, N);
            real lat2 = latlon_from_n(n_e).lat;  // LatLon(Ob).lat in get_w_en_n (kinematics.jl:417)
            V3 w_en_n(v_eb_n.y / (N + h_e), -v_eb_n.x / (M + h_e), -v_eb_n.y * tan(lat2) / (N + h_e));
            V3 w_en_b = qrot_inv(q_nb, w_en_n);
            V3 w_nb_b = w_eb_b - w_en_b;
            real M2, N2; radii(n_e, M2, N2);
            V3 w_ew_n(v_eb_n.y / (N2 + h_e), -v_eb_n.x / (M2 + h_e), 0.0);
            V3 w_ew_b = qrot_inv(q_nb, w_ew_n);
            V3 w_wb_b = w_eb_b - w_ew_b;
            real v_gnd = norm(v_eb_n);
            real chi = azimuth(v_eb_n), gam = inclination(v_eb_n);
            V3 ed = euler_dt(e_nb, w_nb_b);
            xd[0] = ed.x; xd[1] = ed.y; xd[2] = ed.z;
            xd[3] = -w_en_n.y;                  // geodesy.jl:110-114
            xd[4] = w_en_n.x / cos(ll.lat);
            xd[5] = -v_eb_n.z;
            k.e_nb = e_nb; k.q_nb = q_nb; k.q_eb = q_eb; k.q_en = q_en; k.ll = ll; k.n_e = n_e; k.h_e = h_e; k.h_o = h_o;
            k.r_eb_e = r_eb_e; k.w_wb_b = w_wb_b; k.w_eb_b = w_eb_b; k.v_eb_b = v_eb_b; k.v_eb_n = v_eb_n;
            k.v_gnd = v_gnd; k.chi_gnd = chi; k.gamma_gnd = gam;
        }
    }

    // ------------------------------------------------------------------ aerodynamics, c172.jl:307-373
    void aero_ode(const World& W, real u_e, real u_a, real u_r, real u_f) {
        const AeroTables& T = W.aero;
        const AirData& air = y.air; const KinData& kin = y.kin;
        const double S = 16.165, b = 10.912, c = 1.494, V_min = 1.0, tau = 0.02;
        real alpha_filt = x[L.aero], beta_filt = x[L.aero + 1];
        bool stall = s.stall;
        V3 v_wb_a = qrot_inv(Q4(), air.v_wb_b);  // f_ba.q'(v_wb_b), f_ba = identity (c172.jl:204)
        real alpha = 0.0, beta = 0.0;
        if (air.TAS > 0.1) airflow_angles(v_wb_a, alpha, beta);
        real V = rmax(air.TAS, V_min);
        real alpha_filt_dot = 1.0 / tau * (alpha - alpha_filt);
        real beta_filt_dot = 1.0 / tau * (beta - beta_filt);
        real p_nd = kin.w_wb_b.x * b / (2.0 * V);
        real q_nd = kin.w_wb_b.y * c / (2.0 * V);
        real r_nd = kin.w_wb_b.z * b / (2.0 * V);
        real adot_nd = alpha_filt_dot * c / (2.0 * V);
        real bdot_nd = beta_filt_dot * b / (2.0 * V);
        const double d2r_ = ORC_PI / 180.0;
        auto lin = [](real u, double umin, double umax, double lo, double hi) { return lo + (hi - lo) / (umax - umin) * (u - umin); };
        real de = lin(u_e, -1, 1, -28 * d2r_, 23 * d2r_);
        real da = lin(u_a, -1, 1, -20 * d2r_, 20 * d2r_);
        real dr = lin(u_r, -1, 1, -16 * d2r_, 16 * d2r_);
        real df = lin(u_f, 0, 1, 0 * d2r_, 30 * d2r_);
        real dh_nd = (kin.h_o - u.trn_elev) / b;
        // get_aero_coeffs, c172.jl:226-245
        real a_c = rclamp(alpha, -0.1, 0.36), b_c = rclamp(beta, -0.2, 0.2);
        real ad_c = rclamp(adot_nd, -0.04, 0.04), bd_c = rclamp(bdot_nd, -0.2, 0.2);
        (void)bd_c;
        real st = stall ? 1.0 : 0.0;
        real C_D = T.CD_zero + T.CD_ge(dh_nd) * (T.CD_a_df(a_c, df) + T.CD_df(df)) + T.CD_de(de) + T.CD_b(b_c);
        real C_Y = T.CY_dr * dr + T.CY_da * da + T.CY_b_df(b_c, df) + T.CY_p(a_c, df) * p_nd + T.CY_r(a_c, df) * r_nd;
        real C_L = T.CL_ge(dh_nd) * (T.CL_a(a_c, st) + T.CL_df(df)) + T.CL_de * de + T.CL_q * q_nd + T.CL_adot * ad_c;
        real C_l = T.Cl_da * da + T.Cl_dr * dr + T.Cl_b * b_c + T.Cl_p * p_nd + T.Cl_r(a_c, df) * r_nd;
        real C_m = T.Cm_zero + T.Cm_de * de + T.Cm_df(df) + T.Cm_a * a_c + T.Cm_q * q_nd + T.Cm_adot * ad_c;
        real C_n = T.Cn_dr * dr + T.Cn_da * da + T.Cn_b * b_c + T.Cn_p * p_nd + T.Cn_r * r_nd;
        Q4 q_as = Ry(-alpha);
        V3 F_s = (air.q * S) * V3(-C_D, C_Y, -C_L);
        V3 F_a = qrot(q_as, F_s);
        V3 tau_a = (air.q * S) * V3(C_l * b, C_m * c, C_n * b);
        xdot[L.aero] = alpha_filt_dot; xdot[L.aero + 1] = beta_filt_dot;
        AeroY& a = y.aero;
        a.e = u_e; a.a = u_a; a.r = u_r; a.f = u_f; a.de = de; a.da = da; a.dr = dr; a.df = df;
        a.alpha = alpha; a.beta = beta; a.alpha_filt = alpha_filt; a.beta_filt = beta_filt;
        a.alpha_filt_dot = alpha_filt_dot; a.beta_filt_dot = beta_filt_dot; a.stall = stall;
        a.C_D = C_D; a.C_Y = C_Y; a.C_L = C_L; a.C_l = C_l; a.C_m = C_m; a.C_n = C_n;
        a.wr_b.F = F_a; a.wr_b.tau = tau_a;
    }

    // ------------------------------------------------------------------ landing gear unit, landinggear.jl:228-328, 411-476
    void ldg_unit(const World& W, int leg, V3 r_bs_b, double k_s, double k_d_ext, double k_d_cmp,
                  bool has_steering, real steer_input, real brake_input, double l_0 = 0.0, double psi_max = ORC_PI / 6) {
        const KinData& kin = y.kin;
        StrutY st; ContactY ct;
        Q4 q_bs;  // identity
        Q4 q_es = qmul(kin.q_eb, q_bs);
        V3 e3(0.0, 0.0, 1.0), e1(1.0, 0.0, 0.0);
        V3 ks_e = qrot(q_es, e3);
        V3 r_bs_e = qrot(kin.q_eb, r_bs_b);
        V3 r_sw0_e = l_0 * ks_e;
        V3 r_ew0_e = kin.r_eb_e + r_bs_e + r_sw0_e;
        V3 loc_Ot; real he_Ow0;
        geo_from_cartesian(r_ew0_e, loc_Ot, he_Ow0);
        real he_Ot = u.trn_elev + W.geoid->height(loc_Ot);
        check_alt(he_Ot);
        st.dh = he_Ow0 - he_Ot;
        st.wow = st.dh <= 0.0;
        st.surface = u.surface;
        real v_xy[2] = {0.0, 0.0};
        if (st.wow) {
            V3 r_et_e = cartesian_from_geo(loc_Ot, he_Ot);
            V3 r_es_e = kin.r_eb_e + r_bs_e;
            V3 r_st_e = r_et_e - r_es_e;
            V3 ut_n(0.0, 0.0, 1.0);
            V3 ut_e = qrot(kin.q_en, ut_n);
            real ut_ks = dot(ut_e, ks_e);
            real l = dot(ut_e, r_st_e) / ut_ks;
            st.alpha_ts = acos(rmax(rmin(ut_ks, 1.0), -1.0));
            st.xi = rmin(0.0, l - l_0);
            V3 r_sc_s = e3 * (l_0 + st.xi);
            V3 r_sc_b = qrot(q_bs, r_sc_s);
            V3 r_bc_b = r_sc_b + r_bs_b;
            V3 v_b = kin.v_eb_b + cross(kin.w_eb_b, r_bc_b);
            V3 v_s = qrot_inv(q_bs, v_b);
            real psi_v = atan2(v_s.y, v_s.x);
            real psi_sw = 0.0;
            if (has_steering) psi_sw = u.steering_engaged ? steer_input * psi_max : psi_v;  // landinggear.jl:68-72
            st.psi_sw = psi_sw;
            Q4 q_sw = Rz(psi_sw);
            Q4 q_ns = qmul(kin.q_nb, q_bs);
            Q4 q_nw = qmul(q_ns, q_sw);
            V3 kc_n = ut_n;
            V3 iw_n = qrot(q_nw, e1);
            V3 iw_n_trn = iw_n - dot(iw_n, kc_n) * kc_n;
            V3 ic_n = normalize(iw_n_trn);
            V3 jc_n = cross(kc_n, ic_n);
            M3 R_nc;
            R_nc.m[0][0] = ic_n.x; R_nc.m[1][0] = ic_n.y; R_nc.m[2][0] = ic_n.z;
            R_nc.m[0][1] = jc_n.x; R_nc.m[1][1] = jc_n.y; R_nc.m[2][1] = jc_n.z;
            R_nc.m[0][2] = kc_n.x; R_nc.m[1][2] = kc_n.y; R_nc.m[2][2] = kc_n.z;
            Q4 q_sc = qmul(qconj(q_ns), q_from_rmat(R_nc));
            Q4 q_bc = qmul(q_bs, q_sc);
            st.t_sc.r = r_sc_s; st.t_sc.q = q_sc;
            st.t_bc.r = r_bc_b; st.t_bc.q = q_bc;
            V3 v_c_body = qrot_inv(q_bc, v_b);
            Q4 q_sc2 = qmul(qconj(q_bs), q_bc);
            V3 ks_c = qrot_inv(q_sc2, e3);
            st.xi_dot = -v_c_body.z / ks_c.z;
            real k_d = (st.xi_dot > 0.0 ? k_d_ext : k_d_cmp);
            st.F_dmp = -(k_s * st.xi + k_d * st.xi_dot);
            V3 v_dmp_c = ks_c * st.xi_dot;
            V3 v_c = v_c_body + v_dmp_c;
            if (!(rabs(v_c.z) < 1e-8)) throw OrcTermination(ST_CONTACT_ASSERT);  // landinggear.jl:321
            v_xy[0] = v_c.x; v_xy[1] = v_c.y;
        }
        st.v_xy[0] = v_xy[0]; st.v_xy[1] = v_xy[1];
        // contact: friction regulator first (landinggear.jl:417-419), gains :401-409
        int xo = L.ldg + 2 * leg;
        for (int i = 0; i < 2; i++)
            pi_lane(-v_xy[i], 0, x[xo + i], 5.0, 400.0, 0.2, -1.0, 1.0, xdot[xo + i], ct.frc.out[i], ct.frc.out_free[i], ct.frc.sat_out[i], ct.frc.halted[i]);
        if (st.wow) {
            real norm_v = sqrt(v_xy[0] * v_xy[0] + v_xy[1] * v_xy[1]);
            ct.mu_roll = get_mu(0.03, 0.02, 0.005, 0.01, norm_v);
            if (u.surface == SURF_DRY) ct.mu_skid = get_mu(0.75, 0.25, 0.005, 0.01, norm_v);
            else if (u.surface == SURF_WET) ct.mu_skid = get_mu(0.25, 0.15, 0.005, 0.01, norm_v);
            else ct.mu_skid = get_mu(0.075, 0.025, 0.005, 0.01, norm_v);
            ct.kappa_br = brake_input * 1.0;  // DirectBraking eta_br = 1 (landinggear.jl:97-112); NoBraking -> 0
            real mu_x = ct.mu_roll + (ct.mu_skid - ct.mu_roll) * ct.kappa_br;
            if (norm_v < 1e-3) ct.psi_cv = ORC_PI / 2; else ct.psi_cv = atan2(v_xy[1], v_xy[0]);
            double psi_skid = 10 * ORC_PI / 180.0;
            real psi_abs = rabs(ct.psi_cv);
            real mu_y;
            if (psi_abs < psi_skid) mu_y = ct.mu_skid * psi_abs / psi_skid;
            else if (psi_abs > ORC_PI - psi_skid) mu_y = ct.mu_skid * (1.0 - (psi_skid + psi_abs - ORC_PI) / psi_skid);
            else mu_y = ct.mu_skid;
            real sc = rmin(1.0, ct.mu_skid / sqrt(mu_x * mu_x + mu_y * mu_y));
            ct.mu_max[0] = mu_x * sc; ct.mu_max[1] = mu_y * sc;
            ct.mu_eff[0] = ct.frc.out[0] * ct.mu_max[0]; ct.mu_eff[1] = ct.frc.out[1] * ct.mu_max[1];
            ct.f_c = V3(ct.mu_eff[0], ct.mu_eff[1], -1.0);
            V3 f_s = qrot(st.t_sc.q, ct.f_c);
            real Nf = -st.F_dmp / f_s.z;
            Nf = rmax(0.0, Nf);
            ct.F_c = ct.f_c * Nf;
            Wrench wr_c; wr_c.F = ct.F_c;
            ct.wr_b = translate(st.t_bc, wr_c);
        }
        y.strut[leg] = st; y.contact[leg] = ct;
    }

    // ------------------------------------------------------------------ power plant, piston.jl:575-595
    void pwp_ode(const World& W, real throttle_in) {
        const AirData& air = y.air; const KinData& kin = y.kin;
        const double gear_ratio = 1.0, d = 2.0, J_xx = 0.3;
        real w_eng = x[L.eng];
        real w = gear_ratio * w_eng;
        // propeller, propellers.jl:405-452
        PropY& p = y.prop;
        V3 r_bp(2.055, 0.0, 0.833);
        V3 v_wOp_b = air.v_wb_b + cross(kin.w_eb_b, r_bp);
        V3 v_wOp_p = qrot_inv(Q4(), v_wOp_b);
        real v_J = norm(v_wOp_p);
        real w_J = rmax(rabs(w), 1.0);
        real J = 2 * ORC_PI * v_J / (w_J * d);
        real Mt = rabs(w) * (d / 2) / air.a;
        const PropTables& PT = W.prop;
        real C_Fx, C_Mx, C_Fz_a, C_Mz_a, C_P, eta_p;
        if (W.prop3) {   // get_dbeta(::Model{<:Propeller{<:VariablePitch}}) = linear_scaling(u[], range) (propellers.jl:395-398, types.jl:70-73)
            const PropTables3& P3 = *W.prop3;
            real dbeta = P3.lo + (P3.hi - P3.lo) / (1.0 - 0.0) * (rclamp(u.prop_pitch, 0.0, 1.0) - 0.0);
            C_Fx = P3.C_Fx(J, Mt, dbeta); C_Mx = P3.C_Mx(J, Mt, dbeta); C_Fz_a = P3.C_Fz_a(J, Mt, dbeta);
            C_Mz_a = P3.C_Mz_a(J, Mt, dbeta); C_P = P3.C_P(J, Mt, dbeta); eta_p = P3.eta_p(J, Mt, dbeta);
            p.dbeta = dbeta;
        } else {
            C_Fx = PT.C_Fx(J, Mt); C_Mx = PT.C_Mx(J, Mt); C_Fz_a = PT.C_Fz_a(J, Mt); C_Mz_a = PT.C_Mz_a(J, Mt); C_P = PT.C_P(J, Mt); eta_p = PT.eta_p(J, Mt);
            p.dbeta = 0.0;
        }
        real al_p, be_p;
        airflow_angles(v_wOp_p, al_p, be_p);
        V3 C_F(C_Fx, C_Fz_a * be_p, C_Fz_a * al_p);
        V3 C_M = 1.0 * V3(C_Mx, C_Mz_a * be_p, C_Mz_a * al_p);
        real f = w / (2 * ORC_PI), f2 = f * f, f3 = f * f2;
        double d4 = d * d * d * d, d5 = d * d4;
        p.wr_p.F = (air.rho * f2 * d4) * C_F;
        p.wr_p.tau = (air.rho * f2 * d5) * C_M;
        p.P = air.rho * rabs(f3) * d5 * C_P;
        Frame t_bp; t_bp.r = r_bp;
        p.wr_b = translate(t_bp, p.wr_p);
        p.hr_p = V3(J_xx * w, 0.0, 0.0);
        p.hr_b = qrot(t_bp.q, p.hr_p);
        p.v_wOp_p = v_wOp_p; p.omega = w; p.J = J; p.Mt = Mt; p.eta_p = eta_p;
        real tau_load = gear_ratio * p.wr_p.tau.x;
        double J_load = gear_ratio * gear_ratio * J_xx;
        // engine, piston.jl:314-426
        EngY& e = y.eng;
        const EngineTables& ET = W.eng;
        real throttle = throttle_in, mixture = u.mixture;
        int state = s.eng_state;
        pi_lane(-w_eng, 0, x[L.eng + 2], 5.0, 200.0, 0.0, -1.0, 1.0, xdot[L.eng + 2], e.frc.out[0], e.frc.out_free[0], e.frc.sat_out[0], e.frc.halted[0]);
        pi_lane(1.0 - w_eng / W.w_idle, 0, x[L.eng + 1], 4.0, 2.0, 0.0, -0.5, 0.5, xdot[L.eng + 1], e.idle.out[0], e.idle.out_free[0], e.idle.sat_out[0], e.idle.halted[0]);
        real mu_ratio_idle = 0.5 + e.idle.out[0];
        real n = w_eng / W.w_rated;
        auto T_ISA = [](real p_) { return isa::T_std * pow(p_ / isa::p_std, -isa::beta[0] * isa::R / isa::g_std); };
        real delta = (air.p / isa::p_std) * pow(T_ISA(air.p) / isa::T_std, -0.5);
        real mu_wot = ET.mu_wot(n, delta);
        real mu = mu_wot * (mu_ratio_idle + throttle * (1.0 - mu_ratio_idle));
        real k_f = 1.0 / sqrt(air.rho / isa::rho_std);
        real mixture_pos;
        if (!u.mixture_auto) mixture_pos = 0.5 * (mixture + 1.0);
        else {
            real f_target = ET.f_lean + mixture * (ET.f_rich - ET.f_lean);
            mixture_pos = f_target / (k_f * ET.f_rich);
        }
        real MAP, far, tau_shaft, P_shaft, SFC, mdot;
        if (state == ENG_OFF) {
            double tau_fr_max = 0.01 * W.P_rated / W.w_rated;
            real tau_fr = e.frc.out[0] * tau_fr_max;
            MAP = air.p; far = 0.0; tau_shaft = tau_fr; P_shaft = 0.0; SFC = 0.0; mdot = 0.0;
        } else if (state == ENG_STARTING) {
            MAP = mu * isa::p_std; far = 0.0; tau_shaft = W.tau_start; P_shaft = tau_shaft * w_eng; SFC = 0.0; mdot = 0.0;
        } else {
            real f_sl = ET.f_rich * mixture_pos;
            far = k_f * f_sl;
            // compute_pi_ISA_pow, piston.jl:457-477
            real d_wot = ET.d_wot(n, mu);
            real pi_std = ET.pi_std(n, mu);
            real pi_wot = ET.pi_wot(n, d_wot);
            real pi_ISA;
            if (rabs(d_wot - 1.0) < 5e-3) pi_ISA = pi_std;
            else pi_ISA = pi_std + (pi_wot - pi_std) / (d_wot - 1.0) * (delta - 1.0);
            pi_ISA = rmax(pi_ISA, 0.0);
            real pi_pow = pi_ISA * sqrt(T_ISA(air.p) / air.T);
            real pi_actual = pi_pow * ET.pi_ratio(far);
            MAP = mu * isa::p_std;
            P_shaft = W.P_rated * pi_actual;
            tau_shaft = (w_eng > 0.0 ? P_shaft / w_eng : real(0.0));
            SFC = ET.sfc_pow(n, pi_actual) * ET.sfc_ratio(far);
            mdot = SFC * P_shaft;
        }
        real sum_tau = tau_shaft + tau_load;
        real sum_J = W.J_eng + J_load;
        xdot[L.eng] = sum_tau / sum_J;
        e.start = u.eng_start; e.stop = u.eng_stop; e.state = state; e.throttle = throttle; e.MAP = MAP; e.mixture = mixture;
        e.mixture_pos = mixture_pos; e.f = far; e.mdot = mdot; e.omega = w_eng; e.n = n; e.tau_shaft = tau_shaft; e.P_shaft = P_shaft; e.SFC = SFC;
    }

    // ------------------------------------------------------------------ rigid body dynamics, dynamics.jl:443-525
    void dyn_ode(const MassProps& mp_b, const Wrench& wr_b, V3 ho_b) {
        const KinData& kin = y.kin;
        Q4 q_eb = kin.q_eb; V3 r_eb_e = kin.r_eb_e;
        V3 w_eb_b(x[L.dyn], x[L.dyn + 1], x[L.dyn + 2]);
        V3 v_eb_b(x[L.dyn + 3], x[L.dyn + 4], x[L.dyn + 5]);
        V3 w_ie_e(0.0, 0.0, wgs::w_ie);
        V3 w_ie_b = qrot_inv(q_eb, w_ie_e);
        V3 r_bc_b = mp_b.r;
        MassProps mp_c = mp_translate(-r_bc_b, mp_b);
        Frame t_cb; t_cb.r = -r_bc_b;
        Wrench wr_c = translate(t_cb, wr_b);
        V3 ho_c = ho_b;
        real m = mp_c.m; const M3& J = mp_c.J;
        V3 w_ec_c = w_eb_b;
        V3 v_ec_c = v_eb_b + cross(w_ec_c, r_bc_b);
        V3 w_ie_c = w_ie_b;
        V3 w_ic_c = w_ie_c + w_ec_c;
        V3 r_bc_e = qrot(q_eb, r_bc_b);
        V3 r_ec_e = r_eb_e + r_bc_e;
        V3 n_c; real h_c;
        geo_from_cartesian(r_ec_e, n_c, h_c);   // ltf(Oc)
        Q4 q_el = ltf(n_c);
        Q4 q_be = qconj(q_eb);
        Q4 q_ce = q_be;
        Q4 q_cl = qmul(q_ce, q_el);
        V3 n_c2; real h_c2;
        geo_from_cartesian(r_ec_e, n_c2, h_c2);  // gravity(Oc) converts again (geodesy.jl:452)
        V3 g_c_l(0.0, 0.0, gravity(n_c2, h_c2));
        V3 g_c_c = qrot(q_cl, g_c_l);
        V3 hc = J * w_ic_c + ho_c;
        V3 wd_ec_c = solve3(J, wr_c.tau - J * cross(w_ie_c, w_ec_c) - cross(w_ic_c, hc));
        V3 vd_ec_c = (1.0 / m) * wr_c.F + g_c_c - cross(w_ec_c + 2.0 * w_ie_c, v_ec_c);
        V3 wd_eb_b = wd_ec_c;
        V3 vd_eb_b = vd_ec_c - cross(wd_ec_c, r_bc_b);
        xdot[L.dyn] = wd_eb_b.x; xdot[L.dyn + 1] = wd_eb_b.y; xdot[L.dyn + 2] = wd_eb_b.z;
        xdot[L.dyn + 3] = vd_eb_b.x; xdot[L.dyn + 4] = vd_eb_b.y; xdot[L.dyn + 5] = vd_eb_b.z;
        DynY& d = y.dyn;
        d.wr_c = wr_c; d.wr_b = wr_b; d.mp_c = mp_c; d.mp_b = mp_b; d.ho_b = ho_b;
        d.wd_ec_c = wd_ec_c; d.vd_ec_c = vd_ec_c; d.g_c_c = g_c_c; d.wd_eb_b = wd_eb_b; d.vd_eb_b = vd_eb_b;
        if (want_y_extra) {  // dynamics.jl:502-513, outputs only
            V3 r_ec_c = qrot(q_ce, r_ec_e);
            V3 r_eb_b = qrot(q_be, r_eb_e);
            d.a_ec_c = vd_ec_c + cross(w_ec_c, v_ec_c);
            d.a_ic_c = vd_ec_c + cross(w_ec_c + 2.0 * w_ie_c, v_ec_c) + cross(w_ie_c, cross(w_ie_c, r_ec_c));
            d.gam_c_c = g_c_c + cross(w_ie_c, cross(w_ie_c, r_ec_c));
            d.f_c_c = d.a_ic_c - d.gam_c_c;
            d.al_ib_b = wd_eb_b - cross(w_eb_b, w_ie_b);
            d.a_eb_b = vd_eb_b + cross(w_eb_b, v_eb_b);
            d.a_ib_b = vd_eb_b + cross(w_eb_b + 2.0 * w_ie_b, v_eb_b) + cross(w_ie_b, cross(w_ie_b, r_eb_b));
        }
    }

    // ------------------------------------------------------------------ f_ode!(vehicle), aircraftbase.jl:142-170
    void f_ode(const World& W) {
        kin_ode(W);
        const KinData& kin = y.kin;
        AtmData at = atm_data(kin.h_o, u.T_sl, u.p_sl, u.wind_n);
        y.air = air_data(at, kin.v_eb_b, kin.q_nb);
        // systems, c172.jl:697-713
        real a_e, a_a, a_r, a_f, steer, br_l, br_r, throttle;
        if (L.variant == VAR_C172S) {  // c172s.jl:92-120
            steer = rclamp(u.rudder_offset + u.rudder, -1.0, 1.0);
            br_l = u.brake_left; br_r = u.brake_right;
            a_e = -rclamp(u.elevator_offset + u.elevator, -1.0, 1.0);
            a_a = rclamp(u.aileron_offset + u.aileron, -1.0, 1.0);
            a_r = -rclamp(u.rudder_offset + u.rudder, -1.0, 1.0);
            a_f = u.flaps;
            throttle = u.throttle;
        } else if (L.variant == VAR_C172XA2) {  // Actuator2, c172x.jl:91-106: second-order lag, clamped position, x = (v, p)
            ActY& a = y.act;
            const double wn = W.act2_wn, zt = W.act2_zeta;
            for (int i = 0; i < 7; i++) {
                real cmd = rclamp(u.act_cmd[i], ACT_LO[i], ACT_HI[i]);
                real xv = x[L.act + 2 * i], xp = x[L.act + 2 * i + 1];
                a.cmd[i] = cmd;
                a.pos[i] = rclamp(xp, ACT_LO[i], ACT_HI[i]);
                a.vel[i] = xv;
                a.sat[i] = (cmd == ACT_HI[i]) - (cmd == ACT_LO[i]);
                xdot[L.act + 2 * i] = (wn * wn) * (cmd - xp) - 2 * zt * wn * xv;
                xdot[L.act + 2 * i + 1] = xv;
            }
            a_e = -a.pos[2]; a_a = a.pos[1]; a_r = -a.pos[3]; a_f = a.pos[4];
            throttle = a.pos[0]; steer = a.pos[3]; br_l = a.pos[5]; br_r = a.pos[6];
        } else {  // c172x.jl:39-52, 127-143
            ActY& a = y.act;
            for (int i = 0; i < 7; i++) {
                real cmd = rclamp(u.act_cmd[i], ACT_LO[i], ACT_HI[i]);
                real xp = x[L.act + i];
                a.cmd[i] = cmd;
                a.pos[i] = rclamp(xp, ACT_LO[i], ACT_HI[i]);
                a.sat[i] = (cmd == ACT_HI[i]) - (cmd == ACT_LO[i]);
                xdot[L.act + i] = 1.0 / 0.05 * (cmd - xp);
            }
            a_e = -a.pos[2]; a_a = a.pos[1]; a_r = -a.pos[3]; a_f = a.pos[4];
            throttle = a.pos[0]; steer = a.pos[3]; br_l = a.pos[5]; br_r = a.pos[6];
        }
        aero_ode(W, a_e, a_a, a_r, a_f);
        ldg_unit(W, 0, V3(-0.381, -1.092, 1.902), 39404, 9340, 9340, false, 0.0, br_l);
        ldg_unit(W, 1, V3(-0.381, 1.092, 1.902), 39404, 9340, 9340, false, 0.0, br_r);
        ldg_unit(W, 2, V3(1.27, 0.0, 1.9), 26269, 3503, 3503, true, steer, 0.0);
        pwp_ode(W, throttle);
        {  // fuel, c172.jl:607-616
            const double m_full = 114.4, m_res = 1.0;
            FuelY& f = y.fuel;
            f.x_avail = x[L.fuel];
            f.m_total = m_res + f.x_avail * (m_full - m_res);
            f.m_avail = f.m_total - m_res;
            xdot[L.fuel] = -y.eng.mdot / (m_full - m_res);
        }
        // get_mp_b(systems): afm, aero, ldg, fuel, pld, pwp, act (dynamics.jl:328-345)
        MassProps zero;
        MassProps mp = zero;
        mp = mp_add(mp, W.afm);  // airframe: `const mp_b_afm`, c172.jl:26-35
        mp = mp_add(mp, zero);  // aero
        mp = mp_add(mp, zero);  // ldg
        {   // fuel, c172.jl:618-636
            real m_fuel = rmax(0.0, y.fuel.m_total);
            MassProps f = zero;
            f = mp_add(f, mp_point(0.5 * m_fuel, V3(0.325, -2.845, 0.0)));
            f = mp_add(f, mp_point(0.5 * m_fuel, V3(0.325, 2.845, 0.0)));
            mp = mp_add(mp, f);
        }
        {   // payload, c172.jl:542-554
            MassProps p = zero;
            p = mp_add(p, mp_point(u.m_pilot, V3(0.183, -0.356, 0.899)));
            p = mp_add(p, mp_point(u.m_copilot, V3(0.183, 0.356, 0.899)));
            p = mp_add(p, mp_point(u.m_lpass, V3(-0.681, -0.356, 0.899)));
            p = mp_add(p, mp_point(u.m_rpass, V3(-0.681, 0.356, 0.899)));
            p = mp_add(p, mp_point(u.m_baggage, V3(-1.316, 0.0, 0.899)));
            mp = mp_add(mp, p);
        }
        mp = mp_add(mp, zero);  // pwp
        mp = mp_add(mp, zero);  // act
        Wrench wz, wr = wz;
        wr = wr + wz;                                              // afm
        wr = wr + y.aero.wr_b;
        { Wrench l = wz; l = l + y.contact[0].wr_b; l = l + y.contact[1].wr_b; l = l + y.contact[2].wr_b; wr = wr + l; }
        wr = wr + wz; wr = wr + wz;                                // fuel, pld
        wr = wr + y.prop.wr_b;
        wr = wr + wz;                                              // act
        V3 hz, hr = hz;
        hr = hr + hz; hr = hr + hz; hr = hr + hz; hr = hr + hz; hr = hr + hz;
        hr = hr + y.prop.hr_b;
        hr = hr + hz;
        dyn_ode(mp, wr, hr);
    }

    // ------------------------------------------------------------------ f_step!(vehicle), aircraftbase.jl:172-181
    // returns true if anything the RHS reads was changed
    bool f_step(const World& W) {
        bool changed = false;
        if (L.kin == KIN_WA || L.kin == KIN_ECEF) {  // kinematics.jl:114-118, 226-229, 317-320
            int blocks[2][2] = {{L.kinx, 4}, {L.kinx + 4, L.kin == KIN_WA ? 4 : 3}};
            for (auto& b : blocks) {
                real n2 = 0.0;
                for (int i = 0; i < b[1]; i++) n2 = n2 + x[b[0] + i] * x[b[0] + i];
                real nn = sqrt(n2);
                if (rabs(nn - 1.0) > 1e-8) { for (int i = 0; i < b[1]; i++) x[b[0] + i] = x[b[0] + i] / nn; changed = true; }
            }
        }
        // aero stall hysteresis, c172.jl:375-384
        bool st0 = s.stall;
        if (y.aero.alpha > 0.36) s.stall = true;
        else if (y.aero.alpha < 0.09) s.stall = false;
        changed |= (s.stall != st0);
        // landing gear, landinggear.jl:331-347, 479-483
        for (int leg = 0; leg < 3; leg++) {
            const StrutY& st = y.strut[leg];
            if (st.wow && st.alpha_ts * (180.0 / ORC_PI) > 60.0) throw OrcTermination(ST_CRASH_ANGLE);
            if (-st.xi_dot > 10.0) throw OrcTermination(ST_CRASH_RATE);
            if (!st.wow) {
                int xo = L.ldg + 2 * leg;
                if (val(x[xo]) != 0.0 || val(x[xo + 1]) != 0.0) changed = true;
                x[xo] = 0.0; x[xo + 1] = 0.0;
            }
        }
        // engine state machine, piston.jl:428-453
        int e0 = s.eng_state;
        real w = x[L.eng];
        bool fuel_available = y.fuel.m_avail > 0.0;
        if (s.eng_state == ENG_OFF) { if (u.eng_start) s.eng_state = ENG_STARTING; }
        else if (s.eng_state == ENG_STARTING) {
            if (!u.eng_start) s.eng_state = ENG_OFF;
            if (w > W.w_idle && fuel_available) s.eng_state = ENG_RUNNING;
        } else {
            if (u.eng_stop || w < W.w_stall || !fuel_available) s.eng_state = ENG_OFF;
        }
        changed |= (s.eng_state != e0);
        return changed;
    }
};
