// fb_init.cuh — ensemble initialisation kernels.
//   init_kernel : init!(sim, C172.Init(KinInit(...), SystemsInitializer(...)))
//                 (lib/FlightPhysics/src/aircraftbase.jl:76-86, kinematics.jl:155-178,335-361,
//                  lib/FlightApps/src/c172/c172s/c172s.jl:168-220, c172x/c172x.jl:222-281)
//   trim_kernel : init!(sim, C172.TrimParameters(...)) (lib/FlightApps/src/c172/c172.jl:825-942,
//                 c172s.jl:227-263, c172x.jl:288-326, aircraftbase.jl:110-118)
// The reference minimises a sum of 7 squared residuals in 7 unknowns with NLopt's BOBYQA until the cost
// drops below 1e-16; each GPU thread finds the same root with a damped Newton iteration on the residual
// vector (forward-difference Jacobian, 8 RHS evaluations per iteration) inside the same bounds.
#pragma once
#include "fb_kernels.cuh"

struct InitSoA {   // device pointers, each [pitch] or nullptr (= default)
    const double* q_nb[4]; const double* lat; const double* lon; const double* h_e;
    const double* w_wb_b[3]; const double* v_eb_n[3];
    const double* n_eng; const double* engine_state; const double* throttle; const double* mixture;
    const double* elevator; const double* aileron; const double* rudder; const double* flaps;
    const double* brake_left; const double* brake_right; const double* fuel_load; const double* payload[5];
    const double* stall; const double* alpha_filt; const double* beta_filt;
};
struct TrimSoA {
    const double* lat; const double* lon; const double* h_e; const double* psi_nb; const double* EAS;
    const double* gamma_wb_n; const double* psi_wb_dot; const double* theta_wb_dot; const double* beta_a;
    const double* fuel_load; const double* mixture; const double* flaps; const double* payload[5];
};

FB_DEV double opt(const double* p, long long i, double dflt) { return p ? p[i] : dflt; }

struct KinIC { q4 q_nb; double lat, lon, h_e; d3 w_wb_b, v_eb_n; };

// f_init!(kinematics, ic); dynamics.x .= kinematics.u
template <int VAR, int KIN>
FB_DEV void kin_init(double* x, const KinIC& ic) {
    typedef Lay<VAR, KIN> L;
    double sla, cla, slo, clo;
    sincos(ic.lat, &sla, &cla); sincos(ic.lon, &slo, &clo);
    double M, N;
    radii(sla, M, N);
    d3 w_ew_n = mk3(ic.v_eb_n.y / (N + ic.h_e), -ic.v_eb_n.x / (M + ic.h_e), 0.0);
    d3 w_eb_b = qrot_inv(ic.q_nb, w_ew_n) + ic.w_wb_b;
    d3 v_eb_b = qrot_inv(ic.q_nb, ic.v_eb_n);
    if (KIN == FB_WA) {
        double sl2, cl2, sa2, ca2;
        sincos(0.5 * ic.lon, &sl2, &cl2); sincos(-0.5 * (ic.lat + 0.5 * FB_PI), &sa2, &ca2);
        q4 q_ew = qmul(mkq(cl2, 0, 0, sl2), mkq(ca2, 0, sa2, 0));   // ltf(n_e), geodesy.jl:132-135
        x[L::KINX] = ic.q_nb.w; x[L::KINX + 1] = ic.q_nb.x; x[L::KINX + 2] = ic.q_nb.y; x[L::KINX + 3] = ic.q_nb.z;
        x[L::KINX + 4] = q_ew.w; x[L::KINX + 5] = q_ew.x; x[L::KINX + 6] = q_ew.y; x[L::KINX + 7] = q_ew.z;
        x[L::KINX + 8] = ic.h_e;
    } else if (KIN == FB_ECEF) {   // kinematics.jl:256-279: q_eb = ltf(n_e) o q_nb, n_e, h_e
        double sl2, cl2, sa2, ca2;
        sincos(0.5 * ic.lon, &sl2, &cl2); sincos(-0.5 * (ic.lat + 0.5 * FB_PI), &sa2, &ca2);
        q4 q_en = qmul(mkq(cl2, 0, 0, sl2), mkq(ca2, 0, sa2, 0));
        q4 q_eb = qmul(q_en, ic.q_nb);
        x[L::KINX] = q_eb.w; x[L::KINX + 1] = q_eb.x; x[L::KINX + 2] = q_eb.y; x[L::KINX + 3] = q_eb.z;
        x[L::KINX + 4] = cla * clo; x[L::KINX + 5] = cla * slo; x[L::KINX + 6] = sla;
        x[L::KINX + 7] = ic.h_e;
    } else if (KIN == FB_NED) {
        q4 q = ic.q_nb;
        x[L::KINX] = atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
        x[L::KINX + 1] = asin(clampd(2.0 * (q.w * q.y - q.x * q.z), -1.0, 1.0));
        x[L::KINX + 2] = atan2(2.0 * (q.w * q.x + q.y * q.z), 1.0 - 2.0 * (q.x * q.x + q.y * q.y));
        // LatLon(NVector(lat, lon)): the round trip through the n-vector (kinematics.jl:347, geodesy.jl:103-106)
        d3 n = mk3(cla * clo, cla * slo, sla);
        x[L::KINX + 3] = atan2(n.z, sqrt(fma(n.x, n.x, n.y * n.y)));
        x[L::KINX + 4] = atan2(n.y, n.x);
        x[L::KINX + 5] = ic.h_e;
    }
    x[L::DYN] = w_eb_b.x; x[L::DYN + 1] = w_eb_b.y; x[L::DYN + 2] = w_eb_b.z;
    x[L::DYN + 3] = v_eb_b.x; x[L::DYN + 4] = v_eb_b.y; x[L::DYN + 5] = v_eb_b.z;
}

struct SysIC {
    double n_eng, throttle, mixture, elevator, aileron, rudder, flaps, brake_left, brake_right, fuel_load;
    double payload[5]; double alpha_filt, beta_filt; int engine_state, stall, mixture_auto;
};

// f_init!(systems, init): writes x, inputs and discrete state
template <int VAR, int KIN>
FB_DEV void sys_init(double* x, const MemberIn& u, int& stall, int& eng, const SysIC& s) {
    typedef Lay<VAR, KIN> L;
    constexpr double w_rated = 2700 * FB_PI / 30;
    for (int j = 0; j < 5; j++) u.set(FB_U_M_PILOT + j, clampd(s.payload[j], 0.0, 100.0));
    u.set(FB_U_MIXTURE, clampd(s.mixture, 0.0, 1.0));
    u.set(FB_U_MIXTURE_AUTO, (double)s.mixture_auto);
    if (VAR == FB_C172S) {
        u.set(FB_U_THROTTLE, clampd(s.throttle, 0.0, 1.0));
        u.set(FB_U_ELEVATOR, clampd(s.elevator, -1.0, 1.0));
        u.set(FB_U_AILERON, clampd(s.aileron, -1.0, 1.0));
        u.set(FB_U_RUDDER, clampd(s.rudder, -1.0, 1.0));
        u.set(FB_U_FLAPS, clampd(s.flaps, 0.0, 1.0));
        u.set(FB_U_AILERON_OFFSET, 0.0); u.set(FB_U_ELEVATOR_OFFSET, 0.0); u.set(FB_U_RUDDER_OFFSET, 0.0);
        u.set(FB_U_BRAKE_LEFT, clampd(s.brake_left, 0.0, 1.0));
        u.set(FB_U_BRAKE_RIGHT, clampd(s.brake_right, 0.0, 1.0));
    } else {
        double c[7] = {s.throttle, s.aileron, s.elevator, s.rudder, s.flaps, s.brake_left, s.brake_right};
#pragma unroll
        for (int j = 0; j < 7; j++) {
            double v = clampd(c[j], (j == 0 || j >= 4) ? 0.0 : -1.0, 1.0);
            u.set(FB_U_ACT_THROTTLE + j, v);
            if (VAR == FB_C172XA2) { x[L::ACT + 2 * j] = 0.0; x[L::ACT + 2 * j + 1] = v; }   // act.x.<unit>.p = value; an Actuator2's v keeps its zero
            else x[L::ACT + j] = v;
        }
    }
    eng = s.engine_state; stall = s.stall;
    x[L::ENG] = s.n_eng * w_rated; x[L::ENG + 1] = 0.0; x[L::ENG + 2] = 0.0;
    x[L::AERO] = s.alpha_filt; x[L::AERO + 1] = s.beta_filt;
    x[L::FUEL] = clampd(s.fuel_load, 0.0, 1.0);
#pragma unroll
    for (int j = 0; j < 6; j++) x[L::LDG + j] = 0.0;
}

template <int VAR, int KIN>
__global__ void __launch_bounds__(FB_BLOCK, 2)
init_kernel(EnsembleDev E, InitSoA I) {
    typedef Lay<VAR, KIN> L;
    constexpr int NX = L::NX;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E.n) return;
    MemberIn u; u.u = E.u + i; u.pitch = E.pitch;
    double x[NX];
    KinIC k;
    q4 q = mkq(opt(I.q_nb[0], i, 1.0), opt(I.q_nb[1], i, 0.0), opt(I.q_nb[2], i, 0.0), opt(I.q_nb[3], i, 0.0));
    double qn = rsqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);   // RQuat(v) normalises (attitude.jl:69-71)
    k.q_nb = mkq(q.w * qn, q.x * qn, q.y * qn, q.z * qn);
    k.lat = opt(I.lat, i, 0.0); k.lon = opt(I.lon, i, 0.0); k.h_e = opt(I.h_e, i, 0.0);
    k.w_wb_b = mk3(opt(I.w_wb_b[0], i, 0.0), opt(I.w_wb_b[1], i, 0.0), opt(I.w_wb_b[2], i, 0.0));
    k.v_eb_n = mk3(opt(I.v_eb_n[0], i, 0.0), opt(I.v_eb_n[1], i, 0.0), opt(I.v_eb_n[2], i, 0.0));
    kin_init<VAR, KIN>(x, k);
    SysIC s;
    s.n_eng = opt(I.n_eng, i, 0.0); s.engine_state = (int)opt(I.engine_state, i, 0.0);
    s.throttle = opt(I.throttle, i, 0.0); s.mixture = opt(I.mixture, i, 0.5); s.elevator = opt(I.elevator, i, 0.0);
    s.aileron = opt(I.aileron, i, 0.0); s.rudder = opt(I.rudder, i, 0.0); s.flaps = opt(I.flaps, i, 0.0);
    s.brake_left = opt(I.brake_left, i, 0.0); s.brake_right = opt(I.brake_right, i, 0.0); s.fuel_load = opt(I.fuel_load, i, 0.5);
    const double pd[5] = {75.0, 75.0, 0.0, 0.0, 50.0};
    for (int j = 0; j < 5; j++) s.payload[j] = opt(I.payload[j], i, pd[j]);
    s.stall = opt(I.stall, i, 0.0) != 0.0; s.alpha_filt = opt(I.alpha_filt, i, 0.0); s.beta_filt = opt(I.beta_filt, i, 0.0);
    s.mixture_auto = 1;
    int stall, eng;
    sys_init<VAR, KIN>(x, u, stall, eng, s);
#pragma unroll
    for (int c = 0; c < NX; c++) E.x[(long long)c * E.pitch + i] = x[c];
    E.stall[i] = (unsigned char)stall; E.eng[i] = (unsigned char)eng; E.status[i] = FB_ST_RUNNING;
}

// trim bounds, c172.jl:897-917 (alpha_a, phi_nb, n_eng, throttle, aileron, elevator, rudder)
__constant__ double c_trim_lo[7] = {-FB_PI / 12, -FB_PI / 3, 0.4, 0, -1, -1, -1};
__constant__ double c_trim_hi[7] = {0.36, FB_PI / 3, 1.1, 1, 1, 1, 1};

struct TrimP { double lat, lon, h_e, psi_nb, EAS, gamma, psi_dot, theta_dot, beta_a, fuel_load, mixture, flaps, payload[5]; };

// assign!(vehicle, trim_params, trim_state) up to (not including) the final f_ode!
template <int VAR, int KIN>
FB_DEV void trim_assign(double* x, const MemberIn& u, int& stall, int& eng, const TrimP& p, const double* ts,
                        const float* geoid, int& err) {
    double alpha_a = ts[0], phi_nb = ts[1];
    // AtmosphericData(atmosphere, Ob): Ob has an ellipsoidal altitude -> orthometric through the geoid
    double h_o = p.h_e - geoid_height(geoid, p.lat, p.lon);
    double T, pr, lnp;
    isa_data(h_o, u(FB_U_T_SL), [&]() { return u(FB_U_P_SL); }, u(FB_UX_TRC), T, pr, lnp, err);
    double rho = pr / (isa::R * T);
    double TAS = p.EAS * sqrt(isa::rho_std / rho);
    double sa, ca, sb, cb;
    sincos(alpha_a, &sa, &ca); sincos(p.beta_a, &sb, &cb);
    d3 v_wb_b = mk3(TAS * ca * cb, TAS * sb, TAS * sa * cb);
    // theta_constraint, aircraftbase.jl:110-118
    double sphi, cphi;
    sincos(phi_nb, &sphi, &cphi);
    double a = v_wb_b.x / TAS, b = (v_wb_b.y * sphi + v_wb_b.z * cphi) / TAS, sg = sin(p.gamma);
    double theta_nb = atan((a * b + sg * sqrt(a * a + b * b - sg * sg)) / (a * a - sg * sg));
    KinIC k;
    double sps, cps, sth, cth, sph, cph;
    sincos(0.5 * p.psi_nb, &sps, &cps); sincos(0.5 * theta_nb, &sth, &cth); sincos(0.5 * phi_nb, &sph, &cph);
    k.q_nb = qmul(qmul(mkq(cps, 0, 0, sps), mkq(cth, 0, sth, 0)), mkq(cph, sph, 0, 0));
    double st, ct;
    sincos(theta_nb, &st, &ct);
    // omega(e_wb, (psi_dot, theta_dot, 0)), attitude.jl:460-474
    k.w_wb_b = mk3(-st * p.psi_dot, ct * sphi * p.psi_dot + cphi * p.theta_dot, ct * cphi * p.psi_dot - sphi * p.theta_dot);
    k.lat = p.lat; k.lon = p.lon; k.h_e = p.h_e;
    d3 wind = mk3(u(FB_U_WIND_N), u(FB_U_WIND_E), u(FB_U_WIND_D));
    k.v_eb_n = wind + qrot(k.q_nb, v_wb_b);
    kin_init<VAR, KIN>(x, k);
    SysIC s;
    s.engine_state = FB_ENG_RUNNING; s.mixture_auto = 1; s.stall = 0;
    s.n_eng = ts[2]; s.throttle = ts[3]; s.aileron = ts[4]; s.elevator = ts[5]; s.rudder = ts[6];
    s.mixture = p.mixture; s.flaps = p.flaps; s.brake_left = 0.0; s.brake_right = 0.0; s.fuel_load = p.fuel_load;
    for (int j = 0; j < 5; j++) s.payload[j] = p.payload[j];
    s.alpha_filt = alpha_a; s.beta_filt = p.beta_a;
    sys_init<VAR, KIN>(x, u, stall, eng, s);
}

template <int VAR, int KIN>
__global__ void __launch_bounds__(FB_BLOCK, 2)
trim_kernel(EnsembleDev E, TrimSoA P, double* trim_state_out, double* cost_out) {
    typedef Lay<VAR, KIN> L;
    constexpr int NX = L::NX;
    constexpr double w_rated = 2700 * FB_PI / 30;
    extern __shared__ __align__(16) double smem_tb[];
    stage_tables(smem_tb, E.tables.pack);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E.n) return;
    MemberIn u; u.u = E.u + i; u.pitch = E.pitch;
    TrimP p;
    p.lat = opt(P.lat, i, 0.0); p.lon = opt(P.lon, i, 0.0); p.h_e = opt(P.h_e, i, 1050.0); p.psi_nb = opt(P.psi_nb, i, 0.0);
    p.EAS = opt(P.EAS, i, 50.0); p.gamma = opt(P.gamma_wb_n, i, 0.0); p.psi_dot = opt(P.psi_wb_dot, i, 0.0);
    p.theta_dot = opt(P.theta_wb_dot, i, 0.0); p.beta_a = opt(P.beta_a, i, 0.0);
    p.fuel_load = clampd(opt(P.fuel_load, i, 0.5), 0.0, 1.0); p.mixture = clampd(opt(P.mixture, i, 0.5), 0.0, 1.0);
    p.flaps = clampd(opt(P.flaps, i, 0.0), 0.0, 1.0);
    const double pd[5] = {75.0, 75.0, 0.0, 0.0, 50.0};
    for (int j = 0; j < 5; j++) p.payload[j] = clampd(opt(P.payload[j], i, pd[j]), 0.0, 100.0);

    const double* lo = c_trim_lo;
    const double* hi = c_trim_hi;
    double ts[7] = {0.1, 0.0, 0.75, 0.47, 0.014, -0.0015, 0.02};   // TrimState() defaults, c172.jl:796-804
    double x[NX], xd[NX];
    int stall = 0, eng = FB_ENG_RUNNING;

    auto eval = [&](const double* t, double* r) -> double {
        int err = 0;
        trim_assign<VAR, KIN>(x, u, stall, eng, p, t, E.tables.geoid, err);
        RhsOut y;
        ArraySink sink; sink.xd = xd;
        rhs<VAR, KIN, false>(x, sink, stall, eng, u, smem_tb, E.tables.geoid, y, nullptr);
        if (err || y.err) { for (int j = 0; j < 7; j++) r[j] = 1e3; return 7e6; }
        double vn = rsqrt(x[L::DYN + 3] * x[L::DYN + 3] + x[L::DYN + 4] * x[L::DYN + 4] + x[L::DYN + 5] * x[L::DYN + 5]);
        for (int j = 0; j < 3; j++) r[j] = xd[L::DYN + 3 + j] * vn;       // cost(vehicle), c172.jl:857-867
        for (int j = 0; j < 3; j++) r[3 + j] = xd[L::DYN + j];
        r[6] = xd[L::ENG] * (1.0 / w_rated);
        double c = 0.0;
        for (int j = 0; j < 7; j++) c = fma(r[j], r[j], c);
        return c;
    };
    double r0[7];
    double c0 = eval(ts, r0);
#pragma unroll 1
    for (int it = 0; it < 60 && c0 > 1e-26; it++) {
        double A[7][8];
#pragma unroll 1
        for (int j = 0; j < 7; j++) {
            double tj[7];
            for (int q = 0; q < 7; q++) tj[q] = ts[q];
            double h = 1e-7;
            if (tj[j] + h > hi[j]) h = -h;
            tj[j] += h;
            double rj[7];
            eval(tj, rj);
            for (int q = 0; q < 7; q++) A[q][j] = (rj[q] - r0[q]) / h;
        }
        for (int q = 0; q < 7; q++) A[q][7] = -r0[q];
        bool sing = false;
#pragma unroll 1
        for (int c = 0; c < 7; c++) {   // Gaussian elimination with partial pivoting
            int pv = c;
            for (int q = c + 1; q < 7; q++) if (fabs(A[q][c]) > fabs(A[pv][c])) pv = q;
            if (fabs(A[pv][c]) < 1e-300) { sing = true; break; }
            for (int q = 0; q < 8; q++) { double t = A[c][q]; A[c][q] = A[pv][q]; A[pv][q] = t; }
            for (int q = c + 1; q < 7; q++) { double f = A[q][c] / A[c][c]; for (int w = c; w < 8; w++) A[q][w] -= f * A[c][w]; }
        }
        if (sing) break;
        double dx[7];
        for (int q = 6; q >= 0; q--) { double s = A[q][7]; for (int w = q + 1; w < 7; w++) s -= A[q][w] * dx[w]; dx[q] = s / A[q][q]; }
        double lam = 1.0; bool improved = false;
#pragma unroll 1
        for (int ls = 0; ls < 30; ls++) {
            double tn[7], rn[7];
            for (int q = 0; q < 7; q++) tn[q] = clampd(fma(lam, dx[q], ts[q]), lo[q], hi[q]);
            double cn = eval(tn, rn);
#ifdef FB_TRIM_DEBUG
            if (i == 0 && it >= 5 && ls == 0) printf("  lo %.6e %.6e %.6e hi %.6e %.6e %.6e dx0 %.6e fma %.10e clamp %.10e\n", lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], dx[0], fma(lam, dx[0], ts[0]), clampd(fma(lam, dx[0], ts[0]), lo[0], hi[0]));
            if (i == 0 && it >= 5) printf("  ls %d lam %.3e ts0 %.10e tn0 %.10e ts1 %.10e tn1 %.10e cn %.6e c0 %.6e\n", ls, lam, ts[0], tn[0], ts[1], tn[1], cn, c0);
#endif
            if (cn < c0) { for (int q = 0; q < 7; q++) { ts[q] = tn[q]; r0[q] = rn[q]; } c0 = cn; improved = true; break; }
            lam *= 0.5;
        }
#ifdef FB_TRIM_DEBUG
        if (i == 0) printf("it %d c0 %.3e lam %.3e improved %d dx %.3e %.3e %.3e %.3e %.3e %.3e %.3e\n", it, c0, lam, (int)improved, dx[0], dx[1], dx[2], dx[3], dx[4], dx[5], dx[6]);
#endif
        if (!improved) break;
    }
    int err = 0;
    trim_assign<VAR, KIN>(x, u, stall, eng, p, ts, E.tables.geoid, err);   // leave the vehicle at the optimum
    // the trimmed condition must be airborne with the engine above idle (asserts at c172s.jl:256-261)
    RhsOut y;
    ArraySink sink; sink.xd = xd;
    rhs<VAR, KIN, false>(x, sink, stall, eng, u, smem_tb, E.tables.geoid, y, nullptr);
    bool ok = (c0 <= 1e-16) && !err && !y.err && y.wow == 0;
#pragma unroll
    for (int c = 0; c < NX; c++) E.x[(long long)c * E.pitch + i] = x[c];
    E.stall[i] = (unsigned char)stall; E.eng[i] = (unsigned char)eng;
    E.status[i] = ok ? FB_ST_RUNNING : FB_ST_TRIM_FAILED;
    if (trim_state_out) for (int j = 0; j < 7; j++) trim_state_out[(long long)j * E.pitch + i] = ts[j];
    if (cost_out) cost_out[i] = c0;
}

// ------------------------------------------------------------------------------------------ linearization
// linearize(aircraft, trim_params) (lib/FlightPhysics/src/aircraftbase.jl:292-341 -> linearization.jl:55-111) for every
// member about its current state and inputs, wind set to zero as the reference does (NoWind()). State / input /
// output vectors: lib/FlightApps/src/c172/c172s/c172s.jl:269-412 (16 / 4 / 33), c172x/c172x.jl:332-498 (20 / 4 / 38);
// NED kinematics only (the reference defines them for Vehicle{NED}). Jacobians follow FiniteDiff's default rule
// (third-party, restated): forward differences, eps_j = max(sqrt(eps) |x_j|, sqrt(eps)).
// Output layout (member fastest): A[(c * NXS + r) * pitch + i] = dxdot_r / dx_c, likewise B, C, D.
template <int VAR> struct SS {
    static constexpr int NX = (VAR == FB_C172S) ? 16 : 20;
    static constexpr int NU = 4;
    static constexpr int NY = (VAR == FB_C172S) ? 33 : 38;
};
template <int VAR, int KIN>
FB_DEV void ss_get_x(const double* x, double* o) {
    typedef Lay<VAR, KIN> L;
    for (int i = 0; i < 3; i++) { o[i] = x[L::DYN + i]; o[3 + i] = x[L::KINX + i]; o[6 + i] = x[L::DYN + 3 + i]; o[9 + i] = x[L::KINX + 3 + i]; }
    o[12] = x[L::AERO]; o[13] = x[L::AERO + 1]; o[14] = x[L::ENG]; o[15] = x[L::FUEL];
    if (VAR != FB_C172S) for (int i = 0; i < 4; i++) o[16 + i] = x[L::ACT + i];
}
template <int VAR, int KIN>
FB_DEV void ss_assign_x(double* x, const double* o) {
    typedef Lay<VAR, KIN> L;
    for (int i = 0; i < 3; i++) { x[L::DYN + i] = o[i]; x[L::KINX + i] = o[3 + i]; x[L::DYN + 3 + i] = o[6 + i]; x[L::KINX + 3 + i] = o[9 + i]; }
    x[L::AERO] = o[12]; x[L::AERO + 1] = o[13]; x[L::ENG] = o[14]; x[L::FUEL] = o[15];
    if (VAR != FB_C172S) for (int i = 0; i < 4; i++) x[L::ACT + i] = o[16 + i];
}
template <int VAR> FB_DEV int ss_u_field(int k) {
    if (VAR == FB_C172S) return k == 0 ? FB_U_THROTTLE : (k == 1 ? FB_U_AILERON : (k == 2 ? FB_U_ELEVATOR : FB_U_RUDDER));
    return FB_U_ACT_THROTTLE + k;
}
template <int VAR, int KIN>
FB_DEV void ss_get_y(const double* x, const YFull& y, const MemberIn& u, double* o) {
    typedef Lay<VAR, KIN> L;
    int k = 0;
    o[k++] = y.w_eb_b.x; o[k++] = y.w_eb_b.y; o[k++] = y.w_eb_b.z;
    o[k++] = y.psi; o[k++] = y.theta; o[k++] = y.phi;
    o[k++] = x[L::DYN + 3]; o[k++] = x[L::DYN + 4]; o[k++] = x[L::DYN + 5];
    o[k++] = y.lat; o[k++] = y.lon; o[k++] = y.h_e;
    o[k++] = y.alpha_filt; o[k++] = y.beta_filt; o[k++] = y.eng_omega;
    if (VAR != FB_C172S) o[k++] = y.eng_n;
    o[k++] = x[L::FUEL];
    if (VAR != FB_C172S) for (int i = 0; i < 4; i++) o[k++] = y.act_pos[i];
    o[k++] = y.f_c_c.x; o[k++] = y.f_c_c.y; o[k++] = y.f_c_c.z;
    o[k++] = y.alpha; o[k++] = y.beta; o[k++] = y.EAS; o[k++] = y.TAS;
    o[k++] = y.v_eb_n.x; o[k++] = y.v_eb_n.y; o[k++] = y.v_eb_n.z;
    o[k++] = y.chi_gnd; o[k++] = y.gamma_gnd; o[k++] = -y.v_eb_n.z;
    for (int i = 0; i < 4; i++) o[k++] = u(ss_u_field<VAR>(i));
}

template <int VAR, int KIN>
__global__ void __launch_bounds__(FB_BLOCK, 2)
linearize_kernel(EnsembleDev E, double* A, double* B, double* C, double* D, double* x0o, double* u0o, double* y0o, double* xd0o) {
    typedef Lay<VAR, KIN> L;
    constexpr int NX = L::NX, NXS = SS<VAR>::NX, NUS = SS<VAR>::NU, NYS = SS<VAR>::NY;
    extern __shared__ __align__(16) double smem_tb[];
    stage_tables(smem_tb, E.tables.pack);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E.n) return;
    if (KIN != FB_NED) return;
    const long long P = E.pitch;
    double ul[FB_U__COUNT];                                   // private copy of the inputs: perturbations never reach HBM
    for (int f = 0; f < FB_U__COUNT; f++) ul[f] = E.u[(long long)f * P + i];
    ul[FB_U_WIND_N] = 0.0; ul[FB_U_WIND_E] = 0.0; ul[FB_U_WIND_D] = 0.0;
    MemberIn u; u.u = ul; u.pitch = 1;
    double x[NX], xd[NX];
    for (int c = 0; c < NX; c++) x[c] = E.x[(long long)c * P + i];
    const int stall = E.stall[i], eng = E.eng[i];
    double x0[NXS], u0[NUS], f0[NXS], y0[NYS];
    ss_get_x<VAR, KIN>(x, x0);
    for (int k = 0; k < NUS; k++) u0[k] = ul[ss_u_field<VAR>(k)];
    YFull yf; RhsOut ro;
    rhs_with_outputs<VAR, KIN>(x, xd, stall, eng, u, smem_tb, E.tables.geoid, &ro, &yf);
    ss_get_x<VAR, KIN>(xd, f0);
    ss_get_y<VAR, KIN>(x, yf, u, y0);
    for (int r = 0; r < NXS; r++) { if (x0o) x0o[r * P + i] = x0[r]; if (xd0o) xd0o[r * P + i] = f0[r]; }
    for (int r = 0; r < NUS; r++) if (u0o) u0o[r * P + i] = u0[r];
    for (int r = 0; r < NYS; r++) if (y0o) y0o[r * P + i] = y0[r];
    const double rs = 1.4901161193847656e-08;   // sqrt(eps(Float64))
#pragma unroll 1
    for (int j = 0; j < NXS + NUS; j++) {
        double xs[NXS], f[NXS], y[NYS];
        for (int r = 0; r < NXS; r++) xs[r] = x0[r];
        for (int k = 0; k < NUS; k++) ul[ss_u_field<VAR>(k)] = u0[k];
        const bool is_x = j < NXS;
        const double v0 = is_x ? x0[j] : u0[j - NXS];
        const double eps = fmax(rs * fabs(v0), rs);
        if (is_x) xs[j] += eps;
        else { const int fld = ss_u_field<VAR>(j - NXS); ul[fld] = saturate_input(fld, u0[j - NXS] + eps); }   // Ranged assignment
        ss_assign_x<VAR, KIN>(x, xs);
        rhs_with_outputs<VAR, KIN>(x, xd, stall, eng, u, smem_tb, E.tables.geoid, &ro, &yf);
        ss_get_x<VAR, KIN>(xd, f);
        ss_get_y<VAR, KIN>(x, yf, u, y);
        if (is_x) {
            for (int r = 0; r < NXS; r++) A[((long long)j * NXS + r) * P + i] = (f[r] - f0[r]) / eps;
            for (int r = 0; r < NYS; r++) C[((long long)j * NYS + r) * P + i] = (y[r] - y0[r]) / eps;
        } else {
            for (int r = 0; r < NXS; r++) B[((long long)(j - NXS) * NXS + r) * P + i] = (f[r] - f0[r]) / eps;
            for (int r = 0; r < NYS; r++) D[((long long)(j - NXS) * NYS + r) * P + i] = (y[r] - y0[r]) / eps;
        }
    }
}
