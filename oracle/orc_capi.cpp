// TEST INFRASTRUCTURE ONLY (see orc_real.hpp). extern "C" surface used by tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs through ctypes.
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <memory>
#include <thread>
#include <atomic>
#include "../include/flightbatch.h"
#include "orc_sim.hpp"

#if ORC_COUNT
thread_local OpCounts g_ops;
#endif

static std::vector<float> g_geoid_data;
static Geoid g_geoid;
GainTables g_gains;

static real clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

bool set_input_field(Member& m, int f, double v) {
    Inputs& u = m.u;
    switch (f) {
        case FB_U_AILERON: u.aileron = clampd(v, -1, 1); break;
        case FB_U_ELEVATOR: u.elevator = clampd(v, -1, 1); break;
        case FB_U_RUDDER: u.rudder = clampd(v, -1, 1); break;
        case FB_U_AILERON_OFFSET: u.aileron_offset = clampd(v, -1, 1); break;
        case FB_U_ELEVATOR_OFFSET: u.elevator_offset = clampd(v, -1, 1); break;
        case FB_U_RUDDER_OFFSET: u.rudder_offset = clampd(v, -1, 1); break;
        case FB_U_FLAPS: u.flaps = clampd(v, 0, 1); break;
        case FB_U_BRAKE_LEFT: u.brake_left = clampd(v, 0, 1); break;
        case FB_U_BRAKE_RIGHT: u.brake_right = clampd(v, 0, 1); break;
        case FB_U_THROTTLE: u.throttle = clampd(v, 0, 1); break;
        case FB_U_MIXTURE: u.mixture = clampd(v, 0, 1); break;
        case FB_U_ENG_START: u.eng_start = v != 0; break;
        case FB_U_ENG_STOP: u.eng_stop = v != 0; break;
        case FB_U_MIXTURE_AUTO: u.mixture_auto = v != 0; break;
        case FB_U_STEERING_ENGAGED: u.steering_engaged = v != 0; break;
        case FB_U_M_PILOT: u.m_pilot = clampd(v, 0, 100); break;
        case FB_U_M_COPILOT: u.m_copilot = clampd(v, 0, 100); break;
        case FB_U_M_LPASS: u.m_lpass = clampd(v, 0, 100); break;
        case FB_U_M_RPASS: u.m_rpass = clampd(v, 0, 100); break;
        case FB_U_M_BAGGAGE: u.m_baggage = clampd(v, 0, 100); break;
        case FB_U_T_SL: u.T_sl = clampd(v, isa::T_std - 50, isa::T_std + 50); break;
        case FB_U_P_SL: u.p_sl = clampd(v, isa::p_std - 10000, isa::p_std + 10000); break;
        case FB_U_WIND_N: u.wind_n.x = v; break;
        case FB_U_WIND_E: u.wind_n.y = v; break;
        case FB_U_WIND_D: u.wind_n.z = v; break;
        case FB_U_TRN_ELEVATION: u.trn_elev = v; break;
        case FB_U_TRN_SURFACE: u.surface = (int)v; break;
        case FB_U_PROP_PITCH: u.prop_pitch = clampd(v, 0, 1); break;
        default:
            if (f >= FB_U_ACT_THROTTLE && f <= FB_U_ACT_BRAKE_RIGHT) { int i = f - FB_U_ACT_THROTTLE; u.act_cmd[i] = clampd(v, ACT_LO[i], ACT_HI[i]); break; }
            return false;
    }
    return true;
}
double get_input_field(const Member& m, int f) {
    const Inputs& u = m.u;
    switch (f) {
        case FB_U_AILERON: return val(u.aileron); case FB_U_ELEVATOR: return val(u.elevator); case FB_U_RUDDER: return val(u.rudder);
        case FB_U_AILERON_OFFSET: return val(u.aileron_offset); case FB_U_ELEVATOR_OFFSET: return val(u.elevator_offset);
        case FB_U_RUDDER_OFFSET: return val(u.rudder_offset); case FB_U_FLAPS: return val(u.flaps);
        case FB_U_BRAKE_LEFT: return val(u.brake_left); case FB_U_BRAKE_RIGHT: return val(u.brake_right);
        case FB_U_THROTTLE: return val(u.throttle); case FB_U_MIXTURE: return val(u.mixture);
        case FB_U_ENG_START: return u.eng_start; case FB_U_ENG_STOP: return u.eng_stop; case FB_U_MIXTURE_AUTO: return u.mixture_auto;
        case FB_U_STEERING_ENGAGED: return u.steering_engaged;
        case FB_U_M_PILOT: return val(u.m_pilot); case FB_U_M_COPILOT: return val(u.m_copilot); case FB_U_M_LPASS: return val(u.m_lpass);
        case FB_U_M_RPASS: return val(u.m_rpass); case FB_U_M_BAGGAGE: return val(u.m_baggage);
        case FB_U_T_SL: return val(u.T_sl); case FB_U_P_SL: return val(u.p_sl);
        case FB_U_WIND_N: return val(u.wind_n.x); case FB_U_WIND_E: return val(u.wind_n.y); case FB_U_WIND_D: return val(u.wind_n.z);
        case FB_U_TRN_ELEVATION: return val(u.trn_elev); case FB_U_TRN_SURFACE: return u.surface;
        case FB_U_PROP_PITCH: return val(u.prop_pitch);
        default:
            if (f >= FB_U_ACT_THROTTLE && f <= FB_U_ACT_BRAKE_RIGHT) return val(u.act_cmd[f - FB_U_ACT_THROTTLE]);
    }
    return 0.0 / 0.0;
}

static double get_output_field(const Sim& s, int f) {
    const VehicleY& y = s.m.y;
    switch (f) {
        case FB_Y_PSI: return val(y.kin.e_nb.psi); case FB_Y_THETA: return val(y.kin.e_nb.theta); case FB_Y_PHI: return val(y.kin.e_nb.phi);
        case FB_Y_LAT: return val(y.kin.ll.lat); case FB_Y_LON: return val(y.kin.ll.lon); case FB_Y_H_E: return val(y.kin.h_e); case FB_Y_H_O: return val(y.kin.h_o);
        case FB_Y_V_N: return val(y.kin.v_eb_n.x); case FB_Y_V_E: return val(y.kin.v_eb_n.y); case FB_Y_V_D: return val(y.kin.v_eb_n.z);
        case FB_Y_V_GND: return val(y.kin.v_gnd); case FB_Y_CHI_GND: return val(y.kin.chi_gnd); case FB_Y_GAMMA_GND: return val(y.kin.gamma_gnd);
        case FB_Y_P: return val(y.kin.w_wb_b.x); case FB_Y_Q: return val(y.kin.w_wb_b.y); case FB_Y_R: return val(y.kin.w_wb_b.z);
        case FB_Y_TAS: return val(y.air.TAS); case FB_Y_EAS: return val(y.air.EAS); case FB_Y_CAS: return val(y.air.CAS); case FB_Y_MACH: return val(y.air.M); case FB_Y_QBAR: return val(y.air.q);
        case FB_Y_ALPHA: return val(y.aero.alpha); case FB_Y_BETA: return val(y.aero.beta); case FB_Y_ALPHA_FILT: return val(y.aero.alpha_filt); case FB_Y_BETA_FILT: return val(y.aero.beta_filt);
        case FB_Y_STALL: return y.aero.stall; case FB_Y_C_L: return val(y.aero.C_L); case FB_Y_C_D: return val(y.aero.C_D);
        case FB_Y_WOW_LEFT: return y.strut[0].wow; case FB_Y_WOW_RIGHT: return y.strut[1].wow; case FB_Y_WOW_NOSE: return y.strut[2].wow;
        case FB_Y_XI_LEFT: return val(y.strut[0].xi); case FB_Y_XI_RIGHT: return val(y.strut[1].xi); case FB_Y_XI_NOSE: return val(y.strut[2].xi);
        case FB_Y_FDMP_LEFT: return val(y.strut[0].F_dmp); case FB_Y_FDMP_RIGHT: return val(y.strut[1].F_dmp); case FB_Y_FDMP_NOSE: return val(y.strut[2].F_dmp);
        case FB_Y_ENG_OMEGA: return val(y.eng.omega); case FB_Y_ENG_STATE: return y.eng.state; case FB_Y_ENG_MAP: return val(y.eng.MAP);
        case FB_Y_ENG_P_SHAFT: return val(y.eng.P_shaft); case FB_Y_ENG_TAU_SHAFT: return val(y.eng.tau_shaft); case FB_Y_ENG_MDOT: return val(y.eng.mdot);
        case FB_Y_FUEL_M_TOTAL: return val(y.fuel.m_total);
        case FB_Y_PROP_THRUST: return val(y.prop.wr_p.F.x); case FB_Y_PROP_J: return val(y.prop.J);
        case FB_Y_PROP_DBETA: return val(y.prop.dbeta);
        case FB_Y_MASS: return val(y.dyn.mp_b.m);
        case FB_Y_F_X: return val(y.dyn.f_c_c.x); case FB_Y_F_Y: return val(y.dyn.f_c_c.y); case FB_Y_F_Z: return val(y.dyn.f_c_c.z);
        default: return s.av.output(s.m, f);
    }
}

extern "C" {

// reads the "geoid" section (721x1441 float32, latitude fastest) of the table blob written by
// tools/build_tables.py — data only; the blob format is documented there.
int orc_load_geoid(const char* path) {
    FILE* fp = fopen(path, "rb");
    if (!fp) return -1;
    char magic[8]; uint32_t n = 0, pad = 0;
    if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, "FBTABLE1", 8) != 0) { fclose(fp); return -2; }
    if (fread(&n, 4, 1, fp) != 1 || fread(&pad, 4, 1, fp) != 1) { fclose(fp); return -2; }
    static const char* gnames[GainTables::COUNT] = {"gain.te2te", "gain.tv2te", "gain.vh2te", "gain.ar2ar", "gain.phibeta2ar",
                                                    "gain.q2e", "gain.c2theta", "gain.v2t", "gain.p2phi", "gain.chi2phi"};
    for (uint32_t i = 0; i < n; i++) {   // gain schedules first (data of control/data/*.h5, packed [7 EAS][4 h][P])
        struct { char name[24]; uint32_t dtype, ndim, dims[4]; uint64_t off, nbytes; } rec;
        fseek(fp, 16 + 64L * i, SEEK_SET);
        if (fread(&rec, 64, 1, fp) != 1) break;
        for (int k = 0; k < GainTables::COUNT; k++) if (strncmp(rec.name, gnames[k], 24) == 0) {
            g_gains.t[k].resize(rec.nbytes / 8);
            g_gains.P[k] = (int)rec.dims[2];
            fseek(fp, (long)rec.off, SEEK_SET);
            if (fread(g_gains.t[k].data(), 8, g_gains.t[k].size(), fp) != g_gains.t[k].size()) { fclose(fp); return -6; }
            g_gains.loaded = true;
        }
    }
    for (uint32_t i = 0; i < n; i++) {
        struct { char name[24]; uint32_t dtype, ndim, dims[4]; uint64_t off, nbytes; } rec;
        fseek(fp, 16 + 64L * i, SEEK_SET);
        if (fread(&rec, 64, 1, fp) != 1) break;
        if (strncmp(rec.name, "geoid", 24) == 0) {
            g_geoid_data.resize((size_t)Geoid::NLAT * Geoid::NLON);
            if (rec.nbytes != g_geoid_data.size() * 4) { fclose(fp); return -3; }
            fseek(fp, (long)rec.off, SEEK_SET);
            size_t got = fread(g_geoid_data.data(), sizeof(float), g_geoid_data.size(), fp);
            fclose(fp);
            if (got != g_geoid_data.size()) return -4;
            g_geoid.data = g_geoid_data.data();
            return 0;
        }
    }
    fclose(fp);
    return -5;
}

void* orc_world_create() { if (!g_geoid.data) return nullptr; return new World(&g_geoid); }
void orc_world_destroy(void* w) { delete (World*)w; }
// propeller table as generated by the oracle: out[6][21][21] in order C_Fx, C_Mx, C_Fz_a, C_Mz_a, C_P, eta_p
void orc_world_prop_table(void* w, double* out) {
    World* W = (World*)w;
    const Interp2* all[6] = {&W->prop.C_Fx, &W->prop.C_Mx, &W->prop.C_Fz_a, &W->prop.C_Mz_a, &W->prop.C_P, &W->prop.eta_p};
    for (int k = 0; k < 6; k++) for (int i = 0; i < 441; i++) out[k * 441 + i] = all[k]->y[i];
}
// engine tables: mu_wot[2][9], pi_std[13][3], pi_wot[5][3]
void orc_world_engine_tables(void* w, double* mu_wot, double* pi_std, double* pi_wot) {
    World* W = (World*)w;
    for (int i = 0; i < 18; i++) mu_wot[i] = W->eng.mu_wot.y[i];
    for (int i = 0; i < 39; i++) pi_std[i] = W->eng.pi_std.y[i];
    for (int i = 0; i < 15; i++) pi_wot[i] = W->eng.pi_wot.y[i];
}

void* orc_sim_create(void* w, int variant, int kin, double dt, double Dt) {
    Sim* s = new Sim();
    s->W = (World*)w;
    s->m.L = Layout(variant, kin);
    s->av.variant = variant;
    s->dt = dt; s->Dt = Dt;
    s->periodic_every = (int)std::llround(Dt / dt);
    if (s->periodic_every < 1) s->periodic_every = 1;
    return s;
}
void orc_sim_destroy(void* s) { Sim* S = (Sim*)s; delete S->W_own; delete S; }
// Non-default components of the variant VAR_C172XA2: the Sim gets its own copy of the World with the Actuator2 parameters and,
// for n_pitch > 1, the Lookup of Propeller(VariablePitch(lo, hi; length = n_pitch)) (built once per (lo, hi, n) and cached).
void orc_sim_set_components(void* s, double act2_wn, double act2_zeta, int n_pitch, double pitch_lo, double pitch_hi) {
    Sim* S = (Sim*)s;
    if (!S->W_own) { S->W_own = new World(*S->W); S->W = S->W_own; }
    if (act2_wn > 0) S->W_own->act2_wn = act2_wn;
    if (act2_zeta > 0) S->W_own->act2_zeta = act2_zeta;
    if (n_pitch > 1) {
        static std::mutex mu; static std::vector<PropTables3*> cache; static std::vector<int> cache_n;
        std::lock_guard<std::mutex> g(mu);
        const PropTables3* hit = nullptr;
        for (size_t i = 0; i < cache.size(); i++) if (cache[i]->lo == pitch_lo && cache[i]->hi == pitch_hi && cache_n[i] == n_pitch) hit = cache[i];
        if (!hit) { cache.push_back(new PropTables3(pitch_lo, pitch_hi, n_pitch)); cache_n.push_back(n_pitch); hit = cache.back(); }
        S->W_own->prop3 = hit;
    } else S->W_own->prop3 = nullptr;
}
// prop3 table of a Sim as [21][21][n][6] (C_Fx, C_Mx, C_Fz_a, C_Mz_a, C_P, eta_p): golden check of the blob section
int orc_sim_prop3(void* s, double* out) {
    const PropTables3* P = ((Sim*)s)->W->prop3;
    if (!P) return 0;
    const int n = P->C_Fx.a2.n();
    if (out) {
        const Interp3* all[6] = {&P->C_Fx, &P->C_Mx, &P->C_Fz_a, &P->C_Mz_a, &P->C_P, &P->eta_p};
        for (size_t o = 0; o < (size_t)21 * 21 * n; o++) for (int c = 0; c < 6; c++) out[o * 6 + c] = all[c]->y[o];
    }
    return n;
}
int orc_sim_nx(void* s) { return ((Sim*)s)->m.L.nx; }
int orc_sim_status(void* s) { return ((Sim*)s)->m.status; }
void orc_sim_want_y_extra(void* s, int on) { ((Sim*)s)->m.want_y_extra = on != 0; }

static void reinit(Sim* S) {
    S->n_steps = 0; S->n_periodic = 0; S->sched_pos = 0; S->m.status = ST_RUNNING;
}

// p[17]: lat, lon, h_e, psi_nb, EAS, gamma_wb_n, psi_wb_dot, theta_wb_dot, beta_a, fuel_load, mixture, flaps, payload[5]
int orc_sim_init_trim(void* s, const double* p, double* trim_state_out, double* cost_out) {
    Sim* S = (Sim*)s;
    TrimParams tp;
    LatLon ll; ll.lat = p[0]; ll.lon = p[1];
    tp.n_e = n_from_latlon(ll); tp.h_e = p[2]; tp.psi_nb = p[3]; tp.EAS = p[4]; tp.gamma_wb_n = p[5];
    tp.psi_wb_dot = p[6]; tp.theta_wb_dot = p[7]; tp.beta_a = p[8];
    tp.fuel_load = clampd(p[9], 0, 1); tp.mixture = clampd(p[10], 0, 1); tp.flaps = clampd(p[11], 0, 1);
    for (int i = 0; i < 5; i++) tp.payload[i] = clampd(p[12 + i], 0, 100);
    reinit(S);
    TrimState ts; double cost = 0; int iters = 0;
    bool ok = false;
    try { ok = trim_solve(S->m, *S->W, tp, ts, &cost, &iters); } catch (OrcTermination& e) { S->m.status = e.code; return -1; }
    if (trim_state_out) for (int i = 0; i < 7; i++) trim_state_out[i] = val(ts.v[i]);
    if (cost_out) *cost_out = cost;
    if (!ok) S->m.status = ST_TRIM_FAILED;
    S->av.f_init(S->m, *S->W, S->Dt);
    S->apply_events();
    return ok ? 0 : 1;
}
// evaluate the trim cost at a given trim state (no optimisation): pins the stored reference trims
double orc_sim_trim_cost_at(void* s, const double* p, const double* trim_state) {
    Sim* S = (Sim*)s;
    TrimParams tp;
    LatLon ll; ll.lat = p[0]; ll.lon = p[1];
    tp.n_e = n_from_latlon(ll); tp.h_e = p[2]; tp.psi_nb = p[3]; tp.EAS = p[4]; tp.gamma_wb_n = p[5];
    tp.psi_wb_dot = p[6]; tp.theta_wb_dot = p[7]; tp.beta_a = p[8];
    tp.fuel_load = clampd(p[9], 0, 1); tp.mixture = clampd(p[10], 0, 1); tp.flaps = clampd(p[11], 0, 1);
    for (int i = 0; i < 5; i++) tp.payload[i] = clampd(p[12 + i], 0, 100);
    TrimState ts; for (int i = 0; i < 7; i++) ts.v[i] = trim_state[i];
    reinit(S);
    try { trim_assign(S->m, *S->W, tp, ts); } catch (OrcTermination& e) { return 1e300; }
    return val(trim_cost(S->m, *S->W));
}
// ic[40]: q_nb[4], lat, lon, h_e, w_wb_b[3], v_eb_n[3], n_eng, engine_state, throttle, mixture, elevator, aileron,
//         rudder, flaps, brake_left, brake_right, fuel_load, payload[5], stall, alpha_filt, beta_filt
int orc_sim_init(void* s, const double* ic) {
    Sim* S = (Sim*)s;
    KinInit ki; SysInit si;
    ki.q_nb = qnormalize(Q4(ic[0], ic[1], ic[2], ic[3]));   // RQuat(v) normalises (attitude.jl:69-71)
    LatLon ll; ll.lat = ic[4]; ll.lon = ic[5];
    ki.n_e = n_from_latlon(ll); ki.h_e = ic[6];
    ki.w_wb_b = V3(ic[7], ic[8], ic[9]); ki.v_eb_n = V3(ic[10], ic[11], ic[12]);
    si.n_eng = ic[13]; si.engine_state = (int)ic[14]; si.throttle = ic[15]; si.mixture = ic[16]; si.elevator = ic[17];
    si.aileron = ic[18]; si.rudder = ic[19]; si.flaps = ic[20]; si.brake_left = ic[21]; si.brake_right = ic[22]; si.fuel_load = ic[23];
    for (int i = 0; i < 5; i++) si.payload[i] = ic[24 + i];
    si.stall = ic[29] != 0; si.alpha_filt = ic[30]; si.beta_filt = ic[31];
    reinit(S);
    try { vehicle_init(S->m, *S->W, ki, si); } catch (OrcTermination& e) { S->m.status = e.code; return -1; }
    S->av.f_init(S->m, *S->W, S->Dt);
    S->apply_events();
    return 0;
}
int orc_sim_init_state(void* s, const double* x, int stall, int eng_state) {
    Sim* S = (Sim*)s;
    for (int i = 0; i < S->m.L.nx; i++) S->m.x[i] = x[i];
    S->m.s.stall = stall != 0; S->m.s.eng_state = eng_state;
    reinit(S);
    try { S->world_ode(); } catch (OrcTermination& e) { S->m.status = e.code; return -1; }
    S->apply_events();
    return 0;
}
void orc_sim_get_state(void* s, double* x) { Sim* S = (Sim*)s; for (int i = 0; i < S->m.L.nx; i++) x[i] = val(S->m.x[i]); }
void orc_sim_get_xdot(void* s, double* xd) { Sim* S = (Sim*)s; for (int i = 0; i < S->m.L.nx; i++) xd[i] = val(S->m.xdot[i]); }
void orc_sim_get_discrete(void* s, int* stall, int* eng_state) { Sim* S = (Sim*)s; *stall = S->m.s.stall; *eng_state = S->m.s.eng_state; }
int orc_sim_set_input(void* s, int field, double v) { Sim* S = (Sim*)s; return (set_input_field(S->m, field, v) || S->av.set_input(field, v)) ? 0 : -1; }
double orc_sim_get_input(void* s, int field) { Sim* S = (Sim*)s; double v = get_input_field(S->m, field); return v == v ? v : S->av.get_input(field); }
double orc_sim_get_output(void* s, int field) { return get_output_field(*(Sim*)s, field); }
void orc_sim_add_event(void* s, long long step, int field, double value) {
    Sim* S = (Sim*)s;
    S->schedule.push_back({step, field, value});
    std::stable_sort(S->schedule.begin(), S->schedule.end(), [](const InputEvent& a, const InputEvent& b) { return a.step < b.step; });
}
void orc_sim_clear_events(void* s) { Sim* S = (Sim*)s; S->schedule.clear(); S->sched_pos = 0; }
// f_ode!(world) at the current state
int orc_sim_rhs(void* s) {
    Sim* S = (Sim*)s;
    try { S->world_ode(); } catch (OrcTermination& e) { return e.code; }
    return 0;
}
int orc_sim_step(void* s, long long n) {
    Sim* S = (Sim*)s;
    for (long long i = 0; i < n; i++) S->step();
    return S->m.status;
}
// Rounding-sensitivity probe (tests/ground_util.py): n steps with every state component moved by one ulp in a random
// direction after each step — the per-step difference any other correct implementation of the same algorithm may show.
int orc_sim_step_noisy(void* s, long long n, unsigned long long seed) {
    Sim* S = (Sim*)s;
    unsigned long long z = seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    for (long long k = 0; k < n; k++) {
        S->step();
        if (S->m.status != ST_RUNNING) break;
        for (int i = 0; i < S->m.L.nx; i++) {
            z += 0x9E3779B97F4A7C15ull;
            unsigned long long r = z; r = (r ^ (r >> 30)) * 0xBF58476D1CE4E5B9ull; r = (r ^ (r >> 27)) * 0x94D049BB133111EBull; r ^= r >> 31;
            const double v = val(S->m.x[i]);
            if (v != 0.0) S->m.x[i] = std::nextafter(v, (r & 1) ? INFINITY : -INFINITY);
        }
    }
    return S->m.status;
}
// Modeling-protocol pieces in isolation (parity hooks): f_ode! at the current state for y, then f_step!(world) /
// f_periodic!(avionics) + assign!, exactly the callback bodies of Sim::step()
int orc_sim_fstep(void* s) {
    Sim* S = (Sim*)s;
    if (S->m.status != ST_RUNNING) return S->m.status;
    try { S->world_ode(); S->av.f_step(S->m); S->av.assign(S->m); S->m.f_step(*S->W); } catch (OrcTermination& e) { S->m.status = e.code; }
    return S->m.status;
}
int orc_sim_fperiodic(void* s) {
    Sim* S = (Sim*)s;
    if (S->m.status != ST_RUNNING) return S->m.status;
    try { S->world_ode(); S->av.f_periodic(S->m, *S->W, S->Dt); S->av.assign(S->m); } catch (OrcTermination& e) { S->m.status = e.code; }
    return S->m.status;
}
long long orc_sim_steps_taken(void* s) { return ((Sim*)s)->n_steps; }
// linearize(aircraft, trim) about the current state: dims[3] = nx, nu, ny; matrices row-major
void orc_sim_linearize_dims(void* s, int* dims) { SSDims d = ss_dims(((Sim*)s)->m.L.variant); dims[0] = d.nx; dims[1] = d.nu; dims[2] = d.ny; }
int orc_sim_linearize(void* s, double* A, double* B, double* C, double* D, double* x0, double* u0, double* y0, double* xd0) {
    Sim* S = (Sim*)s;
    try { S->av.assign(S->m); return linearize_member(S->m, *S->W, A, B, C, D, x0, u0, y0, xd0) ? 0 : -1; } catch (OrcTermination& e) { return e.code; }
}

// ---------------------------------------------------------------------------------------------------------
// Component-level hooks used by tests/ to replay the reference's own unit tests against the oracle.
static KinInit kin_init_from(const World* W, const double* ic) {   // ic: q_nb[4], lat, lon, h_orth, w_wb_b[3], v_eb_n[3]
    KinInit ki;
    ki.q_nb = qnormalize(Q4(ic[0], ic[1], ic[2], ic[3]));
    LatLon ll; ll.lat = ic[4]; ll.lon = ic[5];
    ki.n_e = n_from_latlon(ll);
    ki.h_e = ic[6] + W->geoid->height(ki.n_e);    // KinInit(h = HOrth(..)) -> HEllip(h, n_e) (kinematics.jl:36-38)
    ki.w_wb_b = V3(ic[7], ic[8], ic[9]); ki.v_eb_n = V3(ic[10], ic[11], ic[12]);
    return ki;
}
static void put_kin(const KinData& k, double* out) {   // q_nb4, n_e3, h_e, v_eb_b3, w_eb_b3, v_eb_n3, q_en4, r_eb_e3, h_o  (25)
    const real v[] = {k.q_nb.w, k.q_nb.x, k.q_nb.y, k.q_nb.z, k.n_e.x, k.n_e.y, k.n_e.z, k.h_e, k.v_eb_b.x, k.v_eb_b.y, k.v_eb_b.z,
                      k.w_eb_b.x, k.w_eb_b.y, k.w_eb_b.z, k.v_eb_n.x, k.v_eb_n.y, k.v_eb_n.z, k.q_en.w, k.q_en.x, k.q_en.y, k.q_en.z,
                      k.r_eb_e.x, k.r_eb_e.y, k.r_eb_e.z, k.h_o};
    for (int i = 0; i < 25; i++) out[i] = val(v[i]);
}
// kinematics model alone, constant (w_eb_b, v_eb_b) input, fixed-step RK4 + f_step! renormalisation
// (lib/FlightPhysics/test/test_kinematics.jl:37-41,75-93)
int orc_kin_sim(void* w, int kin, const double* ic, double dt, long long nsteps, double* out25) {
    World* W = (World*)w;
    Member m; m.L = Layout(VAR_C172S, kin);
    KinInit ki = kin_init_from(W, ic);
    kin_init(m, ki);
    const Layout& L = m.L;
    try {
        for (long long s = 0; s < nsteps; s++) {
            real x0[12], k1[12], k2[12], k3[12], k4[12];
            int n = L.nkin;
            auto f = [&](real* kd) { m.kin_ode(*W); for (int i = 0; i < n; i++) kd[i] = m.xdot[L.kinx + i]; };
            for (int i = 0; i < n; i++) x0[i] = m.x[L.kinx + i];
            f(k1); for (int i = 0; i < n; i++) m.x[L.kinx + i] = x0[i] + dt / 2 * k1[i];
            f(k2); for (int i = 0; i < n; i++) m.x[L.kinx + i] = x0[i] + dt / 2 * k2[i];
            f(k3); for (int i = 0; i < n; i++) m.x[L.kinx + i] = x0[i] + dt * k3[i];
            f(k4); for (int i = 0; i < n; i++) m.x[L.kinx + i] = x0[i] + (dt / 6) * (2.0 * (k2[i] + k3[i]) + (k1[i] + k4[i]));
            if (kin != KIN_NED) {
                int blocks[2][2] = {{L.kinx, 4}, {L.kinx + 4, kin == KIN_WA ? 4 : 3}};
                for (auto& b : blocks) {
                    real n2 = 0.0; for (int i = 0; i < b[1]; i++) n2 = n2 + m.x[b[0] + i] * m.x[b[0] + i];
                    real nn = sqrt(n2);
                    if (rabs(nn - 1.0) > 1e-8) for (int i = 0; i < b[1]; i++) m.x[b[0] + i] = m.x[b[0] + i] / nn;
                }
            }
        }
        m.kin_ode(*W);
    } catch (OrcTermination& e) { return e.code; }
    put_kin(m.y.kin, out25);
    return 0;
}
// VehicleDynamics.f_ode! with a rigid body (mass m, inertia diag(J) at its CoM, CoM at r_bc_b) and a body wrench
// (lib/FlightPhysics/test/test_dynamics.jl:21-66). out: wdot3, vdot3, a_eb_b3, a_ib_b3, gravity, kin(25)
int orc_dyn_test(void* w, const double* ic, double mass, const double* Jd, const double* r_bc, const double* F, const double* tau, double* out) {
    World* W = (World*)w;
    Member m; m.L = Layout(VAR_C172S, KIN_WA); m.want_y_extra = true;
    KinInit ki = kin_init_from(W, ic);
    kin_init(m, ki);
    try {
        m.kin_ode(*W);
        MassProps c; c.m = mass; c.J.m[0][0] = Jd[0]; c.J.m[1][1] = Jd[1]; c.J.m[2][2] = Jd[2];
        MassProps b = mp_translate(V3(r_bc[0], r_bc[1], r_bc[2]), c);   // t_bc(mp_c), dynamics.jl:254-272
        Wrench wr; wr.F = V3(F[0], F[1], F[2]); wr.tau = V3(tau[0], tau[1], tau[2]);
        for (int i = 0; i < 6; i++) m.x[m.L.dyn + i] = 0.0;   // the reference test leaves dyn.x at zero (test_dynamics.jl:21)
        m.dyn_ode(b, wr, V3());
    } catch (OrcTermination& e) { return e.code; }
    const Layout& L = m.L;
    for (int i = 0; i < 6; i++) out[i] = val(m.xdot[L.dyn + i]);
    const DynY& d = m.y.dyn;
    out[6] = val(d.a_eb_b.x); out[7] = val(d.a_eb_b.y); out[8] = val(d.a_eb_b.z);
    out[9] = val(d.a_ib_b.x); out[10] = val(d.a_ib_b.y); out[11] = val(d.a_ib_b.z);
    out[12] = val(gravity(m.y.kin.n_e, m.y.kin.h_e));
    put_kin(m.y.kin, out + 13);
    return 0;
}
// LandingGearUnit{DirectSteering, DirectBraking, Strut(l_0 = 1, SimpleDamper(25000, 1000, 1000))}.f_ode! at the body
// origin on HorizontalTerrain (lib/FlightPhysics/test/test_landing_gear.jl:104-209).
// out: wow, xi, xi_dot, F_dmp, v_xy[2], mu_roll, mu_skid, mu_max[2], mu_eff[2], f_c[3], wr_b.F[3], xdot[2], sat_out[2]  (22)
int orc_ldg_test(void* w, const double* ic, double steer_input, double brake, int engaged, double psi_max,
                 double frc_x0, double frc_x1, int surface, double* out) {
    World* W = (World*)w;
    Member m; m.L = Layout(VAR_C172S, KIN_WA);
    KinInit ki = kin_init_from(W, ic);
    kin_init(m, ki);
    m.u.steering_engaged = engaged != 0; m.u.surface = surface;
    m.x[m.L.ldg] = frc_x0; m.x[m.L.ldg + 1] = frc_x1;
    try {
        m.kin_ode(*W);
        m.ldg_unit(*W, 0, V3(0.0, 0.0, 0.0), 25000, 1000, 1000, true, clampd(steer_input, -1, 1), clampd(brake, 0, 1), 1.0, psi_max);
    } catch (OrcTermination& e) { return e.code; }
    const StrutY& s = m.y.strut[0]; const ContactY& c = m.y.contact[0];
    const real v[] = {s.wow ? 1.0 : 0.0, s.xi, s.xi_dot, s.F_dmp, s.v_xy[0], s.v_xy[1], c.mu_roll, c.mu_skid, c.mu_max[0], c.mu_max[1],
                      c.mu_eff[0], c.mu_eff[1], c.f_c.x, c.f_c.y, c.f_c.z, c.wr_b.F.x, c.wr_b.F.y, c.wr_b.F.z,
                      m.xdot[m.L.ldg], m.xdot[m.L.ldg + 1], (double)c.frc.sat_out[0], (double)c.frc.sat_out[1]};
    for (int i = 0; i < 22; i++) out[i] = val(v[i]);
    return 0;
}
double orc_get_mu(int skidding, int surface, double v) {   // landinggear.jl:170-193
    if (!skidding) return val(get_mu(0.03, 0.02, 0.005, 0.01, v));
    if (surface == SURF_DRY) return val(get_mu(0.75, 0.25, 0.005, 0.01, v));
    if (surface == SURF_WET) return val(get_mu(0.25, 0.15, 0.005, 0.01, v));
    return val(get_mu(0.075, 0.025, 0.005, 0.01, v));
}
// PistonEngineLookup(n_stall, n_max) queries (lib/FlightPhysics/test/test_piston.jl:57-126).
// which: 0 d_wot(n, mu), 1 pi_std(n, mu), 2 pi_wot(n, delta), 3 mu_wot(n, delta), 4 pi_ISA_pow(n, mu=a, delta=b)
double orc_engine_lookup(double n_stall, double n_max, int which, double n, double a, double b) {
    EngineTables E(n_stall, n_max);
    switch (which) {
        case 0: return val(E.d_wot(n, a));
        case 1: return val(E.pi_std(n, a));
        case 2: return val(E.pi_wot(n, a));
        case 3: return val(E.mu_wot(n, a));
        default: {
            real d_wot = E.d_wot(n, a), pi_std = E.pi_std(n, a), pi_wot = E.pi_wot(n, d_wot), r;
            if (rabs(d_wot - 1.0) < 5e-3) r = pi_std; else r = pi_std + (pi_wot - pi_std) / (d_wot - 1.0) * (b - 1.0);
            return val(rmax(r, 0.0));
        }
    }
}
// gain-schedule lookup (control.jl:975-994): which = GainTables index; returns the packed vector length
int orc_gain_lookup(int which, double EAS, double h, double* out) {
    real g[64];
    g_gains.lookup(which, EAS, h, g);
    for (int k = 0; k < g_gains.P[which]; k++) out[k] = val(g[k]);
    return g_gains.P[which];
}
// h2delta(h) = p/p_std / sqrt(T/T_std) of the standard atmosphere (piston.jl:41-44)
double orc_h2delta(double h_geop) {
    real T, p; isa_data(h_geop, isa::T_std, isa::p_std, T, p);
    return val(p / isa::p_std / sqrt(T / isa::T_std));
}
// geodesy round trip (lib/FlightPhysics/test/test_geodesy.jl:181-192): (lat, lon, h_e) -> ECEF -> (n, h_e)
void orc_geo_roundtrip(double lat, double lon, double h_e, double* out7) {
    LatLon ll; ll.lat = lat; ll.lon = lon;
    V3 n = n_from_latlon(ll);
    V3 r = cartesian_from_geo(n, h_e);
    V3 n2; real h2;
    geo_from_cartesian(r, n2, h2);
    out7[0] = val(r.x); out7[1] = val(r.y); out7[2] = val(r.z); out7[3] = val(n2.x); out7[4] = val(n2.y); out7[5] = val(n2.z); out7[6] = val(h2);
}

// op counters (only meaningful in the ORC_COUNT build)
// Segment(p1; s, chi, dh) -> p2 = out3 {lat, lon, h}; SegmentGuidanceData(seg, Ob) -> out8 {chi_12, gamma_12, s_12, s_1b,
// s_2b, e_sb, v_sb, h_s}; validity of Segment(p1, p2) (c172x_gdc.jl:38-148)
void orc_segment_end(double lat1, double lon1, double h1, double s, double chi, double dh, double* out3) {
    GeoLLH p1; p1.ll.lat = lat1; p1.ll.lon = lon1; p1.h = h1;
    GeoLLH p2 = segment_end(p1, s, chi, dh);
    out3[0] = val(p2.ll.lat); out3[1] = val(p2.ll.lon); out3[2] = val(p2.h);
}
void orc_segment_data(const double* p1, const double* p2, const double* ob, double* out8) {
    GeoLLH a, b, o;
    a.ll.lat = p1[0]; a.ll.lon = p1[1]; a.h = p1[2]; b.ll.lat = p2[0]; b.ll.lon = p2[1]; b.h = p2[2]; o.ll.lat = ob[0]; o.ll.lon = ob[1]; o.h = ob[2];
    SegmentData d = segment_data(a, b, o);
    out8[0] = val(d.chi_12); out8[1] = val(d.gamma_12); out8[2] = val(d.s_12); out8[3] = val(d.s_1b); out8[4] = val(d.s_2b);
    out8[5] = val(d.e_sb); out8[6] = val(d.v_sb); out8[7] = val(d.h_s);
}
int orc_segment_valid(const double* p1, const double* p2) {
    GeoLLH a, b;
    a.ll.lat = p1[0]; a.ll.lon = p1[1]; a.h = p1[2]; b.ll.lat = p2[0]; b.ll.lon = p2[1]; b.h = p2[2];
    return segment_valid(a, b) ? 1 : 0;
}

// LinearizedSS f_ode! (linearization.jl:157-192) under the fixed-step RK4 of Sim::step(): row-major A [nx][nx], B [nx][nu];
// x is advanced in place by n_steps steps with the input held at u
void orc_lss_step(int nx, int nu, const double* A, const double* B, const double* x0, const double* u0, const double* xd0,
                  double* x, const double* u, double dt, long long n_steps) {
    std::vector<double> k1(nx), k2(nx), k3(nx), k4(nx), xs(nx), dx(nx), du(nu), c(nx);
    auto f = [&](const double* xv, double* out) {
        for (int i = 0; i < nx; i++) dx[i] = xv[i] - x0[i];
        for (int i = 0; i < nu; i++) du[i] = u[i] - u0[i];
        for (int i = 0; i < nx; i++) out[i] = xd0[i];
        for (int i = 0; i < nx; i++) { double a = 0.0; for (int j = 0; j < nx; j++) a += A[i * nx + j] * dx[j]; c[i] = a; }
        for (int i = 0; i < nx; i++) out[i] += c[i];
        for (int i = 0; i < nx; i++) { double a = 0.0; for (int j = 0; j < nu; j++) a += B[i * nu + j] * du[j]; c[i] = a; }
        for (int i = 0; i < nx; i++) out[i] += c[i];
    };
    for (long long s = 0; s < n_steps; s++) {
        f(x, k1.data());
        for (int i = 0; i < nx; i++) xs[i] = x[i] + (dt / 2) * k1[i];
        f(xs.data(), k2.data());
        for (int i = 0; i < nx; i++) xs[i] = x[i] + (dt / 2) * k2[i];
        f(xs.data(), k3.data());
        for (int i = 0; i < nx; i++) xs[i] = x[i] + dt * k3[i];
        f(xs.data(), k4.data());
        for (int i = 0; i < nx; i++) x[i] = x[i] + (dt / 6) * (2.0 * (k2[i] + k3[i]) + (k1[i] + k4[i]));
    }
}

// Discrete compensators in isolation (control.jl:123-185 Integrator, :370-471 PID), for the known answers of the reference's
// test_control.jl:88-146, 254-372: n periodic updates with the given input / external saturation.
// par = {k_p, k_i, k_d, tau_f, lo, hi}; st (in/out) = {x_i0, x_d0, sat_out_0}; out = {y_p, y_i, y_d, out_free, output, sat_out, int_halted, u_i}
void orc_pid_run(const double* par, double* st, double input, int sat_ext, double Dt, long long n, double* out) {
    PIDc c; c.k_p = par[0]; c.k_i = par[1]; c.k_d = par[2]; c.tau_f = par[3]; c.lo = par[4]; c.hi = par[5];
    c.x_i0 = st[0]; c.x_d0 = st[1]; c.sat_out_0 = (int)st[2]; c.input = input; c.sat_ext = sat_ext;
    for (long long k = 0; k < n; k++) c.periodic(Dt);
    st[0] = val(c.x_i0); st[1] = val(c.x_d0); st[2] = c.sat_out_0;
    out[0] = val(c.y_p); out[1] = val(c.y_i); out[2] = val(c.y_d); out[3] = val(c.out_free); out[4] = val(c.output); out[5] = c.sat_out;
    out[6] = c.int_halted; out[7] = val(c.u_i_y);
}
// lo, hi; st (in/out) = {x0, sat_out_0}; out = {x1, output, sat_out, halted}
void orc_integ_run(double lo, double hi, double* st, double input, int sat_ext, double Dt, long long n, double* out) {
    Integ c; c.lo = lo; c.hi = hi; c.x0 = st[0]; c.sat_out_0 = (int)st[1]; c.input = input; c.sat_ext = sat_ext;
    for (long long k = 0; k < n; k++) c.periodic(Dt);
    st[0] = val(c.x0); st[1] = c.sat_out_0;
    out[0] = val(c.x1_y); out[1] = val(c.output); out[2] = c.sat_out_y; out[3] = c.halted_y;
}

// One lane of the continuous PI compensator with anti-windup (control.jl:52-81: the landing gear's friction regulators, the
// engine's idle and friction compensators) under the fixed-step RK4, for the reference's test_control.jl:36-84.
// par = {k_p, k_i, k_l, lo, hi}; x (in/out) = the integrator state; out = {output, out_free, sat_out, int_halted, y_i, u_i}
void orc_pi_run(const double* par, double* x, double input, int sat_ext, double dt, long long n, double* out) {
    real xi = *x, xd, o, of; int so; bool h;
    auto f = [&](real xv) { pi_lane(input, sat_ext, xv, par[0], par[1], par[2], par[3], par[4], xd, o, of, so, h); return xd; };
    for (long long k = 0; k < n; k++) {
        real k1 = f(xi), k2 = f(xi + (dt / 2) * k1), k3 = f(xi + (dt / 2) * k2), k4 = f(xi + dt * k3);
        xi = xi + (dt / 6) * (2.0 * (k2 + k3) + (k1 + k4));
    }
    f(xi);   // y of the f_ode! at the final state
    *x = val(xi);
    out[0] = val(o); out[1] = val(of); out[2] = so; out[3] = h; out[4] = val(xi); out[5] = input;
}

// Wander-azimuth geodesy known answers (lib/FlightPhysics/test/test_geodesy.jl:20-62 NVector, :64-94 LatLon, :96-112 Altitude),
// with the functions the WA / ECEF kinematics and the terrain lookup use (orc_geo_atm.hpp):
// out[0]     get_psi_nw(ltf(n_e))                                   (:45, = 0),  n_e = normalize(1, 2, 3)
// out[1]     get_psi_nw(ltf(n_e) o Rz(pi/3)) - pi/3                 (:50)
// out[2..4]  NVector(ltf(n_e) o Rz(pi/3)) - n_e                     (:55)
// out[5]     worst |LatLon - LatLon(NVector(ltf(NVector(LatLon)) o Rz(psi_nw)))| over the 10 x 10 x 10 lattice of :84-92
//            (longitude compared modulo 2 pi; at the poles every longitude is the same point)
// out[6]     Orthometric(Ellipsoidal(h_orth, loc), loc) - h_orth, h_orth = 1500, loc = NVector(3, 1, -5)   (:108, == 0)
// out[7]     norm(n_e) - 1                                           (:26)
void orc_geodesy_test(double* out) {
    V3 n_e = normalize(V3(1.0, 2.0, 3.0));
    Q4 r_en = ltf(n_e, 0.0);
    out[0] = val(get_psi_nw(r_en));
    Q4 r_ew = qmul(r_en, Rz(ORC_PI / 3));
    out[1] = val(get_psi_nw(r_ew)) - ORC_PI / 3;
    V3 d = n_from_q(r_ew) - n_e;
    out[2] = val(d.x); out[3] = val(d.y); out[4] = val(d.z);
    double worst = 0.0;
    for (int i = 0; i < 10; i++) for (int j = 0; j < 10; j++) for (int k = 0; k < 10; k++) {
        LatLon ll; ll.lat = -ORC_PI / 2 + ORC_PI * i / 9.0; ll.lon = -ORC_PI + 2 * ORC_PI * j / 9.0;
        const double psi = -ORC_PI + 2 * ORC_PI * k / 9.0;
        LatLon t = latlon_from_n(n_from_q(qmul(ltf(n_from_latlon(ll), 0.0), Rz(psi))));
        double dlat = std::fabs(val(t.lat) - val(ll.lat));
        double dlon = std::fabs(val(t.lon) - val(ll.lon)); dlon = std::fmin(dlon, std::fabs(dlon - 2 * ORC_PI));
        if (i == 0 || i == 9) dlon = 0.0;
        worst = std::fmax(worst, std::fmax(dlat, dlon));
    }
    out[5] = worst;
    V3 loc = normalize(V3(3.0, 1.0, -5.0));
    Geoid g; g.data = g_geoid_data.data();
    real h_orth = 1500.0;
    real h_ell = h_orth + g.height(loc);          // Altitude{Ellipsoidal}(h_orth, loc), geodesy.jl:238-240
    out[6] = val((h_ell - g.height(loc)) - h_orth);
    out[7] = val(norm(n_e)) - 1.0;
}

// theta constraint of the trim (aircraftbase.jl:110-118) on the reference's test values (test_aircraft_base.jl:16-40): alpha_a 0.15,
// beta_a -0.11, TAS 100, gamma_wb_n -0.07, psi_nb 0.3, phi_nb 0.7 -> theta_nb; returns inclination(REuler(psi, theta, phi)(v_wb_b)) - gamma
double orc_theta_constraint_test() {
    const double al = 0.15, be = -0.11, TAS = 100.0, gam = -0.07, psi = 0.3, phi = 0.7;
    V3 v_wb_b = TAS * V3(std::cos(al) * std::cos(be), std::sin(be), std::sin(al) * std::cos(be));    // get_velocity_vector, atmosphere.jl:323-326
    real th = theta_constraint(v_wb_b, gam, phi);
    V3 v_wb_n = qrot(q_from_euler(Euler(psi, th, phi)), v_wb_b);
    return val(inclination(v_wb_n)) - gam;
}

// Hamilton-product identities of lib/FlightPhysics/test/test_quaternions.jl:60-105 with the oracle's product (quaternions.jl:109-115):
// out[0] |(q1 q2) q3 - q1 (q2 q3)| (:94), out[1] |(q1 + q2) q3 - (q1 q3 + q2 q3)| (:95), out[2] |q1 inv(q1) - 1| (:96),
// out[3] |q1 1 - q1| (:93), out[4] |norm(normalize(q1)) - 1| (:39), out[5] |conj(q1).imag + q1.imag| (:79), out[6] |q1 \ q2 - q2 / q1| (:99, must NOT vanish)
void orc_quat_test(double* out) {
    Q4 q1(1.0, 2.0, 3.0, 4.0), q2(-0.5, 0.7, 2.0, -1.5), q3(0.3, -4.0, 1.0, 2.5);
    auto add = [](Q4 a, Q4 b) { return Q4(a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z); };
    auto sub = [](Q4 a, Q4 b) { return Q4(a.w - b.w, a.x - b.x, a.y - b.y, a.z - b.z); };
    auto inv = [](Q4 a) { real n2 = a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z; Q4 c = qconj(a); return Q4(c.w / n2, c.x / n2, c.y / n2, c.z / n2); };
    out[0] = val(qnorm(sub(qmul(qmul(q1, q2), q3), qmul(q1, qmul(q2, q3)))));
    out[1] = val(qnorm(sub(qmul(add(q1, q2), q3), add(qmul(q1, q3), qmul(q2, q3)))));
    out[2] = val(qnorm(sub(qmul(q1, inv(q1)), Q4(1.0, 0.0, 0.0, 0.0))));
    out[3] = val(qnorm(sub(qmul(q1, Q4(1.0, 0.0, 0.0, 0.0)), q1)));
    out[4] = std::fabs(val(qnorm(qnormalize(q1))) - 1.0);
    Q4 c = qconj(q1);
    out[5] = std::fabs(val(c.x + q1.x)) + std::fabs(val(c.y + q1.y)) + std::fabs(val(c.z + q1.z)) + std::fabs(val(c.w - q1.w));
    out[6] = val(qnorm(sub(qmul(inv(q1), q2), qmul(q2, inv(q1)))));
}

// Rotation algebra known answers (lib/FlightPhysics/test/test_attitude.jl:29-82 RQuat, :182-240 REuler), evaluated with the
// primitives every right-hand side uses (orc_math.hpp = attitude.jl:98-103,118,304-308,382-395; quaternions.jl:109-115).
// out[0..2]   q_ab * (q_bc * x_c) - (q_ab o q_bc) * x_c          (:66)
// out[3..5]   x_c - q_bc' * (q_bc * x_c)                          (:67)
// out[6..9]   dt(q_ab, w) - 0.5 q_ab (x) (0, w)                   (:76)
// out[10..12] w - omega(q_ab, qdot)                               (:78)
// out[13..15] dt(REuler(phi = pi/2), (0,1,0))                     (:203, = (1,0,0))
// out[16..18] dt(REuler(phi = pi), (0,1,0))                       (:204, = (0,-1,0))
// out[19..21] dt(REuler(theta = pi/3, phi = pi/4), (1,0,0))       (:208, = (0,0,1))
// out[22..24] dt(REuler(psi = pi/3, theta = pi/4), (-1,2,1)) - dt(REuler(psi = pi/6, theta = pi/4), (-1,2,1))   (:209, = 0)
// out[25..27] w - omega(r_ab, dt(r_ab, w)), r_ab = REuler(-1.234, sqrt 2, sqrt 3), w = (2,-3.2,4)               (:213-215)
// out[28]     worst |REuler - REuler(RQuat(REuler))| over the 10 x 10 x 10 angle lattice of :225-233 (gimbal lock avoided by 1e-3)
// out[29..32] RQuat(REuler()) - (1,0,0,0)                          (:221)
// out[33..35] r_ab * (r_bc * x_c) - (r_ab o r_bc) * x_c, r_bc = REuler(1.34, -sqrt 2, sqrt 2), x_c = (-1,2,3)    (:199)
void orc_attitude_test(double* out) {
    auto put3 = [&](int o, V3 v) { out[o] = val(v.x); out[o + 1] = val(v.y); out[o + 2] = val(v.z); };
    Q4 q_ab = qnormalize(Q4(0.8, 0.1, 0.1, 0.3)), q_bc = qnormalize(Q4(0.2, -0.5, 0.6, 0.1));   // RQuat(v) normalises (attitude.jl:69-71)
    V3 x_c(1.0, 2.0, 3.0);
    put3(0, qrot(q_ab, qrot(q_bc, x_c)) - qrot(qmul(q_ab, q_bc), x_c));
    put3(3, x_c - qrot(qconj(q_bc), qrot(q_bc, x_c)));
    V3 w(10.0, -4.0, 2.0);
    Q4 qd = qdt(q_ab, w);
    Q4 h = qmul(q_ab, Q4(0.0, w.x, w.y, w.z));
    out[6] = val(qd.w - 0.5 * h.w); out[7] = val(qd.x - 0.5 * h.x); out[8] = val(qd.y - 0.5 * h.y); out[9] = val(qd.z - 0.5 * h.z);
    Q4 wq = qmul(qconj(q_ab), qd);                                            // omega = 2 Im(q' (x) qdot), attitude.jl:120-122
    put3(10, w - V3(2.0 * wq.x, 2.0 * wq.y, 2.0 * wq.z));
    put3(13, euler_dt(Euler(0.0, 0.0, ORC_PI / 2), V3(0.0, 1.0, 0.0)));
    put3(16, euler_dt(Euler(0.0, 0.0, ORC_PI), V3(0.0, 1.0, 0.0)));
    put3(19, euler_dt(Euler(0.0, ORC_PI / 3, ORC_PI / 4), V3(1.0, 0.0, 0.0)));
    put3(22, euler_dt(Euler(ORC_PI / 3, ORC_PI / 4, 0.0), V3(-1.0, 2.0, 1.0)) - euler_dt(Euler(ORC_PI / 6, ORC_PI / 4, 0.0), V3(-1.0, 2.0, 1.0)));
    Euler r_ab(-1.234, std::sqrt(2.0), std::sqrt(3.0)), r_bc(1.34, -std::sqrt(2.0), std::sqrt(2.0));
    V3 w2(2.0, -3.2, 4.0);
    put3(25, w2 - omega_from_euler_dt(r_ab, euler_dt(r_ab, w2)));
    double worst = 0.0;
    for (int i = 0; i < 10; i++) for (int j = 0; j < 10; j++) for (int k = 0; k < 10; k++) {
        const double psi = -ORC_PI + 2 * ORC_PI * i / 9.0, th = -(ORC_PI / 2 - 0.001) + 2 * (ORC_PI / 2 - 0.001) * j / 9.0, phi = -ORC_PI + 2 * ORC_PI * k / 9.0;
        Euler e = euler_from_q(q_from_euler(Euler(psi, th, phi)));
        auto dang = [](double a, double b) { double d = std::fabs(a - b); return std::fmin(d, std::fabs(d - 2 * ORC_PI)); };   // psi, phi = +-pi are one rotation (:194-195)
        worst = std::fmax(worst, std::fmax(dang(val(e.psi), psi), std::fmax(std::fabs(val(e.theta) - th), dang(val(e.phi), phi))));
    }
    out[28] = worst;
    Q4 q0 = q_from_euler(Euler());
    out[29] = val(q0.w) - 1.0; out[30] = val(q0.x); out[31] = val(q0.y); out[32] = val(q0.z);
    V3 xc2(-1.0, 2.0, 3.0);
    Q4 qa = q_from_euler(r_ab), qb = q_from_euler(r_bc);
    put3(33, qrot(qa, qrot(qb, xc2)) - qrot(qmul(qa, qb), xc2));
}

// LinearizedSS f_ode! alone (linearization.jl:164-190), in the reference's order: dx = x - x0, du = u - u0; xdot = xdot0, += A dx,
// += B du; y = y0, += C dx, += D du. Row-major matrices. (Known answers: test_linearization.jl:27-45.)
void orc_lss_rhs(int nx, int nu, int ny, const double* A, const double* B, const double* C, const double* D, const double* xd0,
                 const double* x0, const double* u0, const double* y0, const double* x, const double* u, double* xdot, double* y) {
    std::vector<double> dx(nx), du(nu);
    for (int i = 0; i < nx; i++) dx[i] = x[i] - x0[i];
    for (int i = 0; i < nu; i++) du[i] = u[i] - u0[i];
    auto mul_add = [](int n, int m, const double* M, const double* v, double* out) {
        for (int i = 0; i < n; i++) { double a = 0.0; for (int j = 0; j < m; j++) a += M[i * m + j] * v[j]; out[i] += a; }
    };
    for (int i = 0; i < nx; i++) xdot[i] = xd0[i];
    mul_add(nx, nx, A, dx.data(), xdot); mul_add(nx, nu, B, du.data(), xdot);
    for (int i = 0; i < ny; i++) y[i] = y0[i];
    mul_add(ny, nx, C, dx.data(), y); mul_add(ny, nu, D, du.data(), y);
}

void orc_ops_reset() {
#if ORC_COUNT
    g_ops = OpCounts();
#endif
}
void orc_ops_get(unsigned long long* out4) {
#if ORC_COUNT
    out4[0] = g_ops.plain; out4[1] = g_ops.div; out4[2] = g_ops.sqrt; out4[3] = g_ops.trans;
#else
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
#endif
}
int orc_is_counting() { return ORC_COUNT; }

// ---- batch helper for the CPU baseline: step an array of sims on `threads` OpenMP threads
int orc_batch_step(void** sims, long long n_sims, long long n_steps, int threads) {
    if (threads < 1) threads = 1;
    std::atomic<long long> next(0);
    auto work = [&]() {
        for (;;) {
            long long i0 = next.fetch_add(8);
            if (i0 >= n_sims) break;
            long long i1 = std::min(n_sims, i0 + 8);
            for (long long i = i0; i < i1; i++) {
                Sim* S = (Sim*)sims[i];
                for (long long k = 0; k < n_steps; k++) S->step();
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return 0;
}

}  // extern "C"
