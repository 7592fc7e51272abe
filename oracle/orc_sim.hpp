// TEST INFRASTRUCTURE ONLY (see orc_real.hpp).
// Step loop, initialisation and trim, restating
//   FC/sim.jl:183-253 (Simulation), :301-381 (callbacks), :386 (step!), :390-414 (init!)
//   + the fixed-step RK4 of OrdinaryDiffEqLowOrderRK (compat "2", source not under
//     /root/reference; published tableau restated)
//   FA/c172/c172.jl:796-942 (TrimState/TrimParameters, trim kinematics, cost, f_init!)
//   FA/c172/c172s/c172s.jl:168-263, FA/c172/c172x/c172x.jl:222-326 (systems init, trim assign)
//   FP/kinematics.jl:155-178,255-279,335-361 (kinematics f_init!), FP/aircraftbase.jl:76-86,110-118
#pragma once
#include <vector>
#include <algorithm>
#include "orc_vehicle.hpp"
#include "orc_ctl.hpp"

struct KinInit { Q4 q_nb; V3 n_e; real h_e; V3 w_wb_b, v_eb_n; KinInit() : n_e(1.0, 0.0, 0.0), h_e(0.0) {} };

struct SysInit {  // c172.jl:773-790
    int engine_state = ENG_OFF; int mixture_auto = 1;
    real n_eng = 0.0, throttle = 0.0, mixture = 0.5, elevator = 0.0, aileron = 0.0, rudder = 0.0, flaps = 0.0,
         brake_left = 0.0, brake_right = 0.0, fuel_load = 0.5;
    real payload[5] = {75.0, 75.0, 0.0, 0.0, 50.0};
    bool stall = false; real alpha_filt = 0.0, beta_filt = 0.0;
};

struct TrimParams {  // c172.jl:806-818
    V3 n_e = V3(1.0, 0.0, 0.0); real h_e = 1050.0;
    real psi_nb = 0.0, EAS = 50.0, gamma_wb_n = 0.0, psi_wb_dot = 0.0, theta_wb_dot = 0.0, beta_a = 0.0;
    real fuel_load = 0.5, mixture = 0.5, flaps = 0.0;
    real payload[5] = {75.0, 75.0, 0.0, 0.0, 50.0};
};
struct TrimState { real v[7]; TrimState() { double d[7] = {0.1, 0.0, 0.75, 0.47, 0.014, -0.0015, 0.02}; for (int i = 0; i < 7; i++) v[i] = d[i]; } };
// order: alpha_a, phi_nb, n_eng, throttle, aileron, elevator, rudder (c172.jl:796-804)

inline real clampR(real v, double lo, double hi) { return rclamp(v, lo, hi); }

// f_init!(kinematics, ic) — kinematics.jl:155-178 (WA), :255-279 (ECEF), :335-361 (NED); dynamics.x .= kinematics.u
inline void kin_init(Member& m, const KinInit& ic) {
    const Layout& L = m.L;
    real M, N; radii(ic.n_e, M, N);
    V3 w_ew_n(ic.v_eb_n.y / (N + ic.h_e), -ic.v_eb_n.x / (M + ic.h_e), 0.0);
    V3 w_ew_b = qrot_inv(ic.q_nb, w_ew_n);
    V3 w_eb_b = w_ew_b + ic.w_wb_b;
    V3 v_eb_b = qrot_inv(ic.q_nb, ic.v_eb_n);
    real* xk = m.x + L.kinx;
    if (L.kin == KIN_WA) {
        Q4 q_ew = ltf(ic.n_e);
        xk[0] = ic.q_nb.w; xk[1] = ic.q_nb.x; xk[2] = ic.q_nb.y; xk[3] = ic.q_nb.z;
        xk[4] = q_ew.w; xk[5] = q_ew.x; xk[6] = q_ew.y; xk[7] = q_ew.z;
        xk[8] = ic.h_e;
    } else if (L.kin == KIN_ECEF) {
        Q4 q_eb = qmul(ltf(ic.n_e), ic.q_nb);
        xk[0] = q_eb.w; xk[1] = q_eb.x; xk[2] = q_eb.y; xk[3] = q_eb.z;
        xk[4] = ic.n_e.x; xk[5] = ic.n_e.y; xk[6] = ic.n_e.z;
        xk[7] = ic.h_e;
    } else {
        Euler e = euler_from_q(ic.q_nb);
        LatLon ll = latlon_from_n(ic.n_e);
        xk[0] = e.psi; xk[1] = e.theta; xk[2] = e.phi; xk[3] = ll.lat; xk[4] = ll.lon; xk[5] = ic.h_e;
    }
    m.x[L.dyn] = w_eb_b.x; m.x[L.dyn + 1] = w_eb_b.y; m.x[L.dyn + 2] = w_eb_b.z;
    m.x[L.dyn + 3] = v_eb_b.x; m.x[L.dyn + 4] = v_eb_b.y; m.x[L.dyn + 5] = v_eb_b.z;
}

// f_init!(systems, init) — c172s.jl:168-220 / c172x.jl:222-281
inline void sys_init(Member& m, const World& W, const SysInit& si) {
    const Layout& L = m.L;
    Inputs& u = m.u;
    u.m_pilot = clampR(si.payload[0], 0, 100); u.m_copilot = clampR(si.payload[1], 0, 100); u.m_lpass = clampR(si.payload[2], 0, 100);
    u.m_rpass = clampR(si.payload[3], 0, 100); u.m_baggage = clampR(si.payload[4], 0, 100);
    u.mixture = clampR(si.mixture, 0, 1); u.mixture_auto = si.mixture_auto;
    if (L.variant == VAR_C172S) {
        u.throttle = clampR(si.throttle, 0, 1);
        u.elevator = clampR(si.elevator, -1, 1); u.aileron = clampR(si.aileron, -1, 1); u.rudder = clampR(si.rudder, -1, 1);
        u.flaps = clampR(si.flaps, 0, 1);
        u.aileron_offset = 0.0; u.elevator_offset = 0.0; u.rudder_offset = 0.0;
        u.brake_left = clampR(si.brake_left, 0, 1); u.brake_right = clampR(si.brake_right, 0, 1);
    } else {
        real c[7] = {si.throttle, si.aileron, si.elevator, si.rudder, si.flaps, si.brake_left, si.brake_right};
        // act.<unit>.u[] = value; act.x.<unit>.p = value (c172x.jl:258-274); an Actuator2's velocity state keeps its zero
        const bool a2 = (L.variant == VAR_C172XA2);
        for (int i = 0; i < 7; i++) {
            u.act_cmd[i] = clampR(c[i], ACT_LO[i], ACT_HI[i]);
            if (a2) { m.x[L.act + 2 * i] = 0.0; m.x[L.act + 2 * i + 1] = u.act_cmd[i]; } else m.x[L.act + i] = u.act_cmd[i];
        }
    }
    m.s.eng_state = si.engine_state;
    m.x[L.eng] = si.n_eng * W.w_rated; m.x[L.eng + 1] = 0.0; m.x[L.eng + 2] = 0.0;
    m.s.stall = si.stall;
    m.x[L.aero] = si.alpha_filt; m.x[L.aero + 1] = si.beta_filt;
    m.x[L.fuel] = clampR(si.fuel_load, 0, 1);
    // contact friction regulators keep their (zero) state: Model construction zero-initialises x
}

// f_init!(vehicle, VehicleInitializer) — aircraftbase.jl:76-86
inline void vehicle_init(Member& m, const World& W, const KinInit& ki, const SysInit& si) {
    for (int i = 0; i < 6; i++) m.x[m.L.ldg + i] = 0.0;
    kin_init(m, ki);
    sys_init(m, W, si);
    m.f_ode(W);
}

// aircraftbase.jl:110-118
inline real theta_constraint(V3 v_wb_b, real gamma, real phi) {
    real TAS = norm(v_wb_b);
    real a = v_wb_b.x / TAS;
    real b = (v_wb_b.y * sin(phi) + v_wb_b.z * cos(phi)) / TAS;
    real sg = sin(gamma);
    return atan((a * b + sg * sqrt(a * a + b * b - sg * sg)) / (a * a - sg * sg));
}

// Kinematics.Initializer(trim_state, trim_params, atmosphere) — c172.jl:825-854
inline KinInit trim_kin_init(const Member& m, const World& W, const TrimParams& tp, const TrimState& ts) {
    real alpha_a = ts.v[0], phi_nb = ts.v[1];
    real h_o = tp.h_e - W.geoid->height(tp.n_e);
    AtmData at = atm_data(h_o, m.u.T_sl, m.u.p_sl, m.u.wind_n);
    real TAS = tp.EAS * sqrt(isa::rho_std / at.rho);
    real cb = cos(tp.beta_a);
    V3 v_wb_a = TAS * V3(cos(alpha_a) * cb, sin(tp.beta_a), sin(alpha_a) * cb);
    V3 v_wb_b = qrot(Q4(), v_wb_a);
    real theta_nb = theta_constraint(v_wb_b, tp.gamma_wb_n, phi_nb);
    Euler e_nb(tp.psi_nb, theta_nb, phi_nb);
    KinInit ki;
    ki.q_nb = q_from_euler(e_nb);
    ki.w_wb_b = omega_from_euler_dt(e_nb, V3(tp.psi_wb_dot, tp.theta_wb_dot, 0.0));
    ki.n_e = tp.n_e; ki.h_e = tp.h_e;
    V3 v_wb_n = qrot(ki.q_nb, v_wb_b);
    ki.v_eb_n = at.v + v_wb_n;
    return ki;
}

// AircraftBase.assign!(vehicle, trim_params, trim_state, ...) — c172s.jl:227-263 / c172x.jl:288-326
inline void trim_assign(Member& m, const World& W, const TrimParams& tp, const TrimState& ts) {
    KinInit ki = trim_kin_init(m, W, tp, ts);
    SysInit si;
    si.engine_state = ENG_RUNNING; si.mixture_auto = 1;
    si.n_eng = ts.v[2]; si.throttle = ts.v[3]; si.aileron = ts.v[4]; si.elevator = ts.v[5]; si.rudder = ts.v[6];
    si.mixture = tp.mixture; si.flaps = tp.flaps; si.brake_left = 0.0; si.brake_right = 0.0; si.fuel_load = tp.fuel_load;
    for (int i = 0; i < 5; i++) si.payload[i] = tp.payload[i];
    si.stall = false; si.alpha_filt = ts.v[0]; si.beta_filt = tp.beta_a;
    vehicle_init(m, W, ki, si);
}

// residuals whose squared sum is cost(vehicle) — c172.jl:857-867
inline void trim_residual(const Member& m, const World& W, real r[7]) {
    const Layout& L = m.L;
    real vn = norm(m.y.kin.v_eb_b);
    for (int i = 0; i < 3; i++) r[i] = m.xdot[L.dyn + 3 + i] / vn;
    for (int i = 0; i < 3; i++) r[3 + i] = m.xdot[L.dyn + i];
    r[6] = m.xdot[L.eng] / W.w_rated;
}
inline real trim_cost(const Member& m, const World& W) {
    real r[7]; trim_residual(m, W, r);
    real c = 0.0;
    for (int i = 0; i < 7; i++) c = c + r[i] * r[i];
    return c;
}

static const double TRIM_LO[7] = {-ORC_PI / 12, -ORC_PI / 3, 0.4, 0, -1, -1, -1};
static const double TRIM_HI[7] = {0.36, ORC_PI / 3, 1.1, 1, 1, 1, 1};

// f_init!(vehicle, TrimParameters) — c172.jl:883-942. The reference minimises `cost` with NLopt's
// LN_BOBYQA (third-party C library, compat "1.2.1", not under /root/reference) down to stopval 1e-16.
// The cost is a sum of 7 squared residuals in 7 unknowns, so the oracle finds the same root with a
// damped Newton iteration on the residual vector (forward-difference Jacobian) inside the same bounds;
// parity is on the converged point, pinned by the stored trim solutions.
inline bool trim_solve(Member& m, const World& W, const TrimParams& tp, TrimState& ts_out, double* cost_out = nullptr, int* iters_out = nullptr) {
    TrimState ts;
    auto eval = [&](const TrimState& s, real r[7]) -> real {
        try { trim_assign(m, W, tp, s); } catch (OrcTermination&) { for (int i = 0; i < 7; i++) r[i] = 1e3; return 7e6; }
        trim_residual(m, W, r);
        real c = 0.0; for (int i = 0; i < 7; i++) c = c + r[i] * r[i];
        return c;
    };
    real r0[7]; real c0 = eval(ts, r0);
    int it = 0;
    for (; it < 60 && val(c0) > 1e-26; it++) {
        double Jm[7][7];
        for (int j = 0; j < 7; j++) {
            TrimState tj = ts;
            double h = 1e-7;
            if (val(tj.v[j]) + h > TRIM_HI[j]) h = -h;
            tj.v[j] = tj.v[j] + h;
            real rj[7]; eval(tj, rj);
            for (int i = 0; i < 7; i++) Jm[i][j] = (val(rj[i]) - val(r0[i])) / h;
        }
        // solve J dx = -r (Gaussian elimination, partial pivoting)
        double A[7][8];
        for (int i = 0; i < 7; i++) { for (int j = 0; j < 7; j++) A[i][j] = Jm[i][j]; A[i][7] = -val(r0[i]); }
        bool sing = false;
        for (int c = 0; c < 7; c++) {
            int p = c; for (int i = c + 1; i < 7; i++) if (std::fabs(A[i][c]) > std::fabs(A[p][c])) p = i;
            if (std::fabs(A[p][c]) < 1e-300) { sing = true; break; }
            for (int j = 0; j < 8; j++) std::swap(A[c][j], A[p][j]);
            for (int i = c + 1; i < 7; i++) { double f = A[i][c] / A[c][c]; for (int j = c; j < 8; j++) A[i][j] -= f * A[c][j]; }
        }
        if (sing) break;
        double dx[7];
        for (int i = 6; i >= 0; i--) { double sum = A[i][7]; for (int j = i + 1; j < 7; j++) sum -= A[i][j] * dx[j]; dx[i] = sum / A[i][i]; }
        double lam = 1.0; bool improved = false;
        for (int ls = 0; ls < 30; ls++) {
            TrimState tn = ts;
            for (int j = 0; j < 7; j++) tn.v[j] = rclamp(ts.v[j] + lam * dx[j], TRIM_LO[j], TRIM_HI[j]);
            real rn[7]; real cn = eval(tn, rn);
            if (val(cn) < val(c0)) { ts = tn; c0 = cn; for (int i = 0; i < 7; i++) r0[i] = rn[i]; improved = true; break; }
            lam *= 0.5;
        }
        if (!improved) break;
    }
    trim_assign(m, W, tp, ts);  // leave the vehicle at the optimum (c172.jl:941)
    ts_out = ts;
    if (cost_out) *cost_out = val(c0);
    if (iters_out) *iters_out = it;
    return val(c0) <= 1e-16;  // success = STOPVAL_REACHED (c172.jl:927,933)
}

// ---------------------------------------------------------------------------------- input schedule
struct InputEvent { int64_t step; int field; double value; };  // applied in the user-callback slot after step `step`

bool set_input_field(Member& m, int field, double v);   // defined in orc_capi.cpp
double get_input_field(const Member& m, int field);

struct Sim {
    const World* W;
    World* W_own = nullptr;    // a private copy when the member's components differ from the shared World's (orc_sim_set_components)
    Member m;
    Avionics av;
    double dt = 0.02, Dt = 0.02, t_start = 0.0;
    int64_t n_steps = 0;       // steps taken since init
    int64_t n_periodic = 0;    // mdl._n (sim.jl:358-375)
    int periodic_every = 1;    // Dt / dt (integral; tstops land exactly on the step grid)
    std::vector<InputEvent> schedule;
    size_t sched_pos = 0;
    uint64_t rhs_evals = 0;

    void rhs() { m.f_ode(*W); rhs_evals++; }

    void apply_events() {
        while (sched_pos < schedule.size() && schedule[sched_pos].step <= n_steps) {
            if (!set_input_field(m, schedule[sched_pos].field, schedule[sched_pos].value))
                av.set_input(schedule[sched_pos].field, schedule[sched_pos].value);   // control-law inputs (c172x_ctl.jl:226-237,826-836)
            sched_pos++;
        }
    }

    // root f_ode!: avionics no-op, assign!(vehicle, avionics) then vehicle (aircraftbase.jl:221-230)
    void world_ode() { av.assign(m); rhs(); }

    // one integrator step n -> n+1: re-evaluated k1, k2..k4, update, FSAL-last RHS,
    // then callbacks in CallbackSet order (step, periodic, user).
    void step() {
        if (m.status != ST_RUNNING) return;
        apply_events();   // inputs scheduled at or before the current step index are already in force
        const int nx = m.L.nx;
        real x0[ORC_NX_MAX], k1[ORC_NX_MAX], k2[ORC_NX_MAX], k3[ORC_NX_MAX], k4[ORC_NX_MAX];
        for (int i = 0; i < nx; i++) x0[i] = m.x[i];
        Discrete s0 = m.s;
        try {
            world_ode(); for (int i = 0; i < nx; i++) k1[i] = m.xdot[i];
            real half = dt / 2;
            for (int i = 0; i < nx; i++) m.x[i] = x0[i] + half * k1[i];
            world_ode(); for (int i = 0; i < nx; i++) k2[i] = m.xdot[i];
            for (int i = 0; i < nx; i++) m.x[i] = x0[i] + half * k2[i];
            world_ode(); for (int i = 0; i < nx; i++) k3[i] = m.xdot[i];
            for (int i = 0; i < nx; i++) m.x[i] = x0[i] + dt * k3[i];
            world_ode(); for (int i = 0; i < nx; i++) k4[i] = m.xdot[i];
        } catch (OrcTermination& e) {   // thrown inside a stage evaluation: the member freezes at the step start
            for (int i = 0; i < nx; i++) m.x[i] = x0[i];
            m.s = s0; m.status = e.code;
            return;
        }
        for (int i = 0; i < nx; i++) m.x[i] = x0[i] + (dt / 6) * (2.0 * (k2[i] + k3[i]) + (k1[i] + k4[i]));
        n_steps++;
        try {
            world_ode();  // fsallast: leaves y*
        } catch (OrcTermination& e) { m.status = e.code; return; }   // freezes at x*
        try {
            av.f_step(m);
            av.assign(m);
            m.f_step(*W);
        } catch (OrcTermination& e) { m.status = e.code; return; }
        if (n_steps % periodic_every == 0) {   // t == t_start + (n+1)*Dt (sim.jl:364-368)
            av.f_periodic(m, *W, Dt);          // control laws read vehicle.y = y*
            av.assign(m);
            n_periodic++;
        }
        apply_events();                        // user callback slot
    }
};

// ---------------------------------------------------------------------------------- linearization
// linearize(vehicle, trim_params): FP/aircraftbase.jl:292-341 -> FP/linearization.jl:55-111, with the state / input /
// output vectors of FA/c172/c172s/c172s.jl:269-412 (16 / 4 / 33) and FA/c172/c172x/c172x.jl:332-498 (20 / 4 / 38); NED
// kinematics only. Jacobians by FiniteDiff.finite_difference_jacobian! defaults (third-party, compat "2.28.1", source
// not under /root/reference; published rule restated): forward differences, column j uses
// eps_j = max(sqrt(eps(Float64)) * |x_j|, sqrt(eps(Float64))) and J[:, j] = (f(x + eps_j e_j) - f(x)) / eps_j.
struct SSDims { int nx, nu, ny; };
inline SSDims ss_dims(int variant) { SSDims d; if (variant == VAR_C172S) { d.nx = 16; d.nu = 4; d.ny = 33; } else { d.nx = 20; d.nu = 4; d.ny = 38; } return d; }

inline void ss_get_x(const Layout& L, const real* x, real* o) {   // XStateSpace(x_vehicle)
    for (int i = 0; i < 3; i++) o[i] = x[L.dyn + i];                 // p q r
    for (int i = 0; i < 3; i++) o[3 + i] = x[L.kinx + i];            // psi theta phi
    for (int i = 0; i < 3; i++) o[6 + i] = x[L.dyn + 3 + i];         // v_x v_y v_z
    for (int i = 0; i < 3; i++) o[9 + i] = x[L.kinx + 3 + i];        // lat lon h
    o[12] = x[L.aero]; o[13] = x[L.aero + 1]; o[14] = x[L.eng]; o[15] = x[L.fuel];
    if (L.variant != VAR_C172S) for (int i = 0; i < 4; i++) o[16 + i] = x[L.act + i];   // thr ail ele rud
}
inline void ss_assign_x(const Layout& L, real* x, const real* o) {   // assign_x_ss!
    for (int i = 0; i < 3; i++) x[L.dyn + i] = o[i];
    for (int i = 0; i < 3; i++) x[L.kinx + i] = o[3 + i];
    for (int i = 0; i < 3; i++) x[L.dyn + 3 + i] = o[6 + i];
    for (int i = 0; i < 3; i++) x[L.kinx + 3 + i] = o[9 + i];
    x[L.aero] = o[12]; x[L.aero + 1] = o[13]; x[L.eng] = o[14]; x[L.fuel] = o[15];
    if (L.variant != VAR_C172S) for (int i = 0; i < 4; i++) x[L.act + i] = o[16 + i];
}
inline void ss_get_u(const Member& m, real* o) {   // UStateSpace(vehicle): throttle, aileron, elevator, rudder
    if (m.L.variant == VAR_C172S) { o[0] = m.u.throttle; o[1] = m.u.aileron; o[2] = m.u.elevator; o[3] = m.u.rudder; }
    else for (int i = 0; i < 4; i++) o[i] = m.u.act_cmd[i];
}
inline void ss_assign_u(Member& m, const real* o) {   // assign_u_ss!: Ranged fields saturate on assignment
    if (m.L.variant == VAR_C172S) {
        m.u.throttle = rclamp(o[0], 0.0, 1.0); m.u.aileron = rclamp(o[1], -1.0, 1.0);
        m.u.elevator = rclamp(o[2], -1.0, 1.0); m.u.rudder = rclamp(o[3], -1.0, 1.0);
    } else for (int i = 0; i < 4; i++) m.u.act_cmd[i] = rclamp(o[i], ACT_LO[i], ACT_HI[i]);
}
inline void ss_get_y(const Member& m, real* o) {   // YStateSpace(vehicle)
    const VehicleY& y = m.y;
    int k = 0;
    o[k++] = y.kin.w_eb_b.x; o[k++] = y.kin.w_eb_b.y; o[k++] = y.kin.w_eb_b.z;
    o[k++] = y.kin.e_nb.psi; o[k++] = y.kin.e_nb.theta; o[k++] = y.kin.e_nb.phi;
    o[k++] = y.kin.v_eb_b.x; o[k++] = y.kin.v_eb_b.y; o[k++] = y.kin.v_eb_b.z;
    o[k++] = y.kin.ll.lat; o[k++] = y.kin.ll.lon; o[k++] = y.kin.h_e;
    o[k++] = y.aero.alpha_filt; o[k++] = y.aero.beta_filt; o[k++] = y.eng.omega;
    if (m.L.variant != VAR_C172S) o[k++] = y.eng.n;
    o[k++] = y.fuel.x_avail;
    if (m.L.variant != VAR_C172S) for (int i = 0; i < 4; i++) o[k++] = y.act.pos[i];
    o[k++] = y.dyn.f_c_c.x; o[k++] = y.dyn.f_c_c.y; o[k++] = y.dyn.f_c_c.z;
    o[k++] = y.aero.alpha; o[k++] = y.aero.beta; o[k++] = y.air.EAS; o[k++] = y.air.TAS;
    o[k++] = y.kin.v_eb_n.x; o[k++] = y.kin.v_eb_n.y; o[k++] = y.kin.v_eb_n.z;
    o[k++] = y.kin.chi_gnd; o[k++] = y.kin.gamma_gnd; o[k++] = -y.kin.v_eb_n.z;
    real u[4]; ss_get_u(m, u);
    for (int i = 0; i < 4; i++) o[k++] = u[i];
}

// A [nx][nx], B [nx][nu], C [ny][nx], D [ny][nu] row-major; x0, u0, y0, xdot0 as named. Evaluated about the member's
// current state and inputs with the wind set to zero (the reference linearizes under NoWind()); the member is restored.
inline bool linearize_member(Member& m, const World& W, double* A, double* B, double* C, double* D,
                             double* x0o, double* u0o, double* y0o, double* xd0o) {
    if (m.L.kin != KIN_NED) return false;
    const SSDims d = ss_dims(m.L.variant);
    const int nx = m.L.nx;
    real xsave[ORC_NX_MAX]; for (int i = 0; i < nx; i++) xsave[i] = m.x[i];
    const Inputs usave = m.u; const bool wy = m.want_y_extra;
    m.want_y_extra = true;
    m.u.wind_n = V3(0.0, 0.0, 0.0);
    real x0[20], u0[4], f0[20], y0[38];
    ss_get_x(m.L, m.x, x0); ss_get_u(m, u0);
    auto eval = [&](const real* xs, const real* us, real* f, real* y) {
        ss_assign_x(m.L, m.x, xs); ss_assign_u(m, us);
        m.f_ode(W);
        ss_get_x(m.L, m.xdot, f); ss_get_y(m, y);
    };
    eval(x0, u0, f0, y0);
    const double rs = std::sqrt(std::numeric_limits<double>::epsilon());
    for (int j = 0; j < d.nx + d.nu; j++) {
        real xs[20], us[4], f[20], y[38];
        for (int i = 0; i < d.nx; i++) xs[i] = x0[i];
        for (int i = 0; i < d.nu; i++) us[i] = u0[i];
        const bool is_x = j < d.nx;
        const double v0 = val(is_x ? x0[j] : u0[j - d.nx]);
        const double eps = std::max(rs * std::fabs(v0), rs);
        if (is_x) xs[j] = xs[j] + eps; else us[j - d.nx] = us[j - d.nx] + eps;
        eval(xs, us, f, y);
        for (int r = 0; r < d.nx; r++) { double v = (val(f[r]) - val(f0[r])) / eps; if (is_x) A[r * d.nx + j] = v; else B[r * d.nu + (j - d.nx)] = v; }
        for (int r = 0; r < d.ny; r++) { double v = (val(y[r]) - val(y0[r])) / eps; if (is_x) C[r * d.nx + j] = v; else D[r * d.nu + (j - d.nx)] = v; }
    }
    for (int i = 0; i < d.nx; i++) { if (x0o) x0o[i] = val(x0[i]); if (xd0o) xd0o[i] = val(f0[i]); }
    for (int i = 0; i < d.nu; i++) if (u0o) u0o[i] = val(u0[i]);
    for (int i = 0; i < d.ny; i++) if (y0o) y0o[i] = val(y0[i]);
    for (int i = 0; i < nx; i++) m.x[i] = xsave[i];
    m.u = usave; m.want_y_extra = wy;
    m.f_ode(W);
    return true;
}
