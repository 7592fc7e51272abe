// fb_avionics.cuh — discrete-time side of the path: output extraction (the root `y` the reference saves),
// the decimated log writer, and the C172X control laws (ControlLawsLon / ControlLawsLat,
// lib/FlightApps/src/c172/c172x/control/c172x_ctl.jl:286-446, 880-983) run once per Δt.
#pragma once
#include "fb_physics.cuh"

#define FB_CTL_STATE_COUNT 64

struct CtlState {
    // placeholder until the control laws land: modes and commands only
    double lon_mode, lat_mode, thr_cmd, ele_cmd, ail_cmd, rud_cmd;
    FB_DEV void load(const double* p, long long pitch) {
        lon_mode = p[0]; lat_mode = p[pitch]; thr_cmd = p[2 * pitch]; ele_cmd = p[3 * pitch]; ail_cmd = p[4 * pitch]; rud_cmd = p[5 * pitch];
    }
    FB_DEV void store(double* p, long long pitch) const {
        p[0] = lon_mode; p[pitch] = lat_mode; p[2 * pitch] = thr_cmd; p[3 * pitch] = ele_cmd; p[4 * pitch] = ail_cmd; p[5 * pitch] = rud_cmd;
    }
};

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
        case FB_Y_MASS: return y.mass; case FB_Y_F_X: return y.f_c_c.x; case FB_Y_F_Y: return y.f_c_c.y; case FB_Y_F_Z: return y.f_c_c.z;
        case FB_Y_LON_MODE: return cs.lon_mode; case FB_Y_LAT_MODE: return cs.lat_mode;
        case FB_Y_THROTTLE_CMD: return cs.thr_cmd; case FB_Y_ELEVATOR_CMD: return cs.ele_cmd;
        case FB_Y_AILERON_CMD: return cs.ail_cmd; case FB_Y_RUDDER_CMD: return cs.rud_cmd;
        default: return nan("");
    }
}

// f_periodic!(avionics, vehicle) + assign!(vehicle, avionics): placeholder (NoAvionics behaviour)
template <int VAR, int KIN>
FB_DEV bool avionics_periodic(const YFull& yf, const RhsOut& y, const MemberIn& u,
                              CtlState& cs, const FbDeviceTables& T, double Dt) {
    return false;
}
