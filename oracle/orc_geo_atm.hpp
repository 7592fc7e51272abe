// TEST INFRASTRUCTURE ONLY (see orc_real.hpp).
// WGS84 geodesy, EGM96 geoid, ISA atmosphere and air data, restating
//   lib/FlightPhysics/src/geodesy.jl   (constants :15-35, NVector/LatLon :47-106, radii :125-129,
//       ltf :132-135, get_psi_nw :140-147, geoid :186-211, altitude datums :213-246,
//       ECEF->geodetic :367-412, geodetic->ECEF :418-428, gravity :451-467)
//   lib/FlightPhysics/src/atmosphere.jl (ISA :99-135, AtmosphericData :269-278, AirData :220-242)
// in the reference's operation order.
#pragma once
#include <stdexcept>
#include "orc_math.hpp"

struct OrcTermination {  // stands in for the reference's exceptions (SURVEY.md §8b status codes)
    int code;
    explicit OrcTermination(int c) : code(c) {}
};
enum OrcStatus { ST_RUNNING = 0, ST_DONE = 1, ST_CRASH_ANGLE = 2, ST_CRASH_RATE = 3, ST_ALT_BOUNDS = 4, ST_CONTACT_ASSERT = 5, ST_TRIM_FAILED = 6 };

namespace wgs {
static const double GM = 3.986005e+14;
static const double a = 6378137.0;
static const double f = 1.0 / 298.257223563;
static const double w_ie = 7.292115e-05;
static const double b = a * (1 - f);
static const double e2 = 2 * f - f * f;
static const double a2 = a * a;
static const double m = w_ie * w_ie * a * a * b / GM;
static const double g_a = 9.7803253359;
static const double g_b = 9.8321849378;
static const double k_g = b * g_b / (a * g_a) - 1;
static const double h_min = -1000.0;
}

struct LatLon { real lat, lon; };

inline void check_alt(real h) { if (h < wgs::h_min) throw OrcTermination(ST_ALT_BOUNDS); }  // geodesy.jl:218-221

inline LatLon latlon_from_n(V3 n) {  // geodesy.jl:103-106
    LatLon ll;
    ll.lat = atan2(n.z, sqrt(n.x * n.x + n.y * n.y));
    ll.lon = atan2(n.y, n.x);
    return ll;
}
inline V3 n_from_latlon(LatLon ll) {  // geodesy.jl:97-101
    real c = cos(ll.lat);
    return V3(c * cos(ll.lon), c * sin(ll.lon), sin(ll.lat));
}
inline V3 n_from_q(Q4 q) {  // geodesy.jl:62-69
    real dq12 = 2.0 * q.w * q.x, dq13 = 2.0 * q.w * q.y;
    real dq24 = 2.0 * q.x * q.z, dq34 = 2.0 * q.y * q.z;
    return -V3(dq24 + dq13, dq34 - dq12, 1.0 - 2.0 * (q.x * q.x + q.y * q.y));
}
inline void radii(V3 n, real& M, real& N) {  // geodesy.jl:125-129
    real f_den = sqrt(1.0 - wgs::e2 * sq(n.z));
    M = wgs::a * (1 - wgs::e2) / cube(f_den);
    N = wgs::a / f_den;
}
inline Q4 ltf(V3 n, real psi_nw = 0.0) {  // geodesy.jl:132-135
    LatLon ll = latlon_from_n(n);
    return qmul(qmul(Rz(ll.lon), Ry(-(ll.lat + 0.5 * ORC_PI))), Rz(psi_nw));
}
inline real get_psi_nw(Q4 q) {  // geodesy.jl:142-147
    real dq12 = 2.0 * q.w * q.x, dq13 = 2.0 * q.w * q.y;
    real dq24 = 2.0 * q.x * q.z, dq34 = 2.0 * q.y * q.z;
    return atan2(-(dq34 + dq12), dq24 - dq13);
}

struct Geoid {  // geodesy.jl:186-211 ; 721x1441 Float32, latitude fastest
    const float* data = nullptr;
    static const int NLAT = 721, NLON = 1441;
    real height(V3 n) const {
        LatLon ll = latlon_from_n(n);
        real lat = ll.lat;
        real lon = ll.lon + 2.0 * ORC_PI;             // mod(lon + 2pi, 2pi)
        lon = lon - 2.0 * ORC_PI * floor(lon / (2.0 * ORC_PI));
        real fi = (lat + 0.5 * ORC_PI) / (ORC_PI / 720.0);
        real fj = lon / (2.0 * ORC_PI / 1440.0);
        int i = (int)std::floor(val(fi)), j = (int)std::floor(val(fj));
        if (i < 0) i = 0; if (i > NLAT - 2) i = NLAT - 2;
        if (j < 0) j = 0; if (j > NLON - 2) j = NLON - 2;
        real u = fi - (double)i, w = fj - (double)j;
        double f00 = data[i + NLAT * j], f10 = data[i + 1 + NLAT * j];
        double f01 = data[i + NLAT * (j + 1)], f11 = data[i + 1 + NLAT * (j + 1)];
        return (1.0 - w) * ((1.0 - u) * f00 + u * f10) + w * ((1.0 - u) * f01 + u * f11);
    }
};

inline real h_geop_from_orth(real h_o) { return h_o * wgs::a / (wgs::a + h_o); }  // geodesy.jl:235

inline V3 cartesian_from_geo(V3 n, real h_e) {  // geodesy.jl:418-428
    real M, N;
    radii(n, M, N);
    return V3((N + h_e) * n.x, (N + h_e) * n.y, (N * (1 - wgs::e2) + h_e) * n.z);
}
// geodesy.jl:367-412 (Fukushima, one Halley iteration); returns normalised n-vector + ellipsoidal h
inline void geo_from_cartesian(V3 r, V3& n_e, real& h_e) {
    real x = r.x, y = r.y, z = r.z;
    real p = sqrt(x * x + y * y);
    double c = wgs::a * wgs::e2;
    double ec2 = 1 - wgs::e2;
    double ec = std::sqrt(ec2);
    real zc = ec * rabs(z);
    real s0 = rabs(z);
    real c0 = ec * p;
    real a0 = sqrt(s0 * s0 + c0 * c0);
    real a03 = cube(a0);
    real b0 = 1.5 * c * s0 * c0 * ((p * s0 - zc * c0) * a0 - c * s0 * c0);
    real s1 = (zc * a03 + c * cube(s0)) * a03 - b0 * s0;
    real c1 = (p * a03 - c * cube(c0)) * a03 - b0 * c0;
    real cc = ec * c1;
    real s12 = s1 * s1, cc2 = cc * cc;
    real h = (p * cc + s0 * s1 - wgs::a * sqrt(ec2 * s12 + cc2)) / sqrt(s12 + cc2);
    real sgn = (double)isign(z);
    real cos_lat, sin_lat;
    if (s1 < cc) {
        real t = s1 / cc;
        cos_lat = 1.0 / sqrt(1.0 + t * t);
        sin_lat = t * cos_lat * sgn;
    } else {
        real t = cc / s1;
        real as = 1.0 / sqrt(1.0 + t * t);
        cos_lat = t * as;
        sin_lat = as * sgn;
    }
    real cos_lon = 1.0, sin_lon = 0.0;
    if (p > 0.0) { cos_lon = x / p; sin_lon = y / p; }
    n_e = normalize(V3(cos_lat * cos_lon, cos_lat * sin_lon, sin_lat));  // NVector ctor normalises (:49-52)
    check_alt(h);
    h_e = h;
}
inline real gravity(V3 n, real h) {  // geodesy.jl:451-467
    real s2 = sq(n.z);
    real g0 = wgs::g_a * (1.0 + wgs::k_g * s2) / sqrt(1.0 - wgs::e2 * s2);
    return g0 * (1.0 - 2.0 / wgs::a * (1.0 + wgs::f + wgs::m - 2.0 * wgs::f * s2) * h + 3.0 / wgs::a2 * sq(h));
}

// ---------------------------------------------------------------- atmosphere
namespace isa {
static const double R = 287.05287, gamma = 1.40, betas = 1.458e-6, S = 110.4;
static const double T_std = 288.15, p_std = 101325.0, rho_std = p_std / (R * T_std), g_std = 9.80665;
static const double beta[7] = {-6.5e-3, 0, 1e-3, 2.8e-3, 0, -2.8e-3, -2e-3};
static const double h_ceil[7] = {11000, 20000, 32000, 47000, 51000, 71000, 84852};
}
inline real isa_T(real h, real Tb, real hb, double beta) { return Tb + beta * (h - hb); }
inline real isa_p(real h, double g0, real pb, real Tb, real hb, double beta) {  // atmosphere.jl:108-114
    if (beta != 0.0) return pb * pow(1.0 + beta / Tb * (h - hb), -g0 / (beta * isa::R));
    return pb * exp(-g0 / (isa::R * Tb) * (h - hb));
}
inline void isa_data(real h_geop, real T_sl, real p_sl, real& T, real& p) {  // atmosphere.jl:116-135
    real hb = 0.0, Tb = T_sl, pb = p_sl;
    for (int i = 0; i < 7; i++) {
        if (h_geop < isa::h_ceil[i]) {
            T = isa_T(h_geop, Tb, hb, isa::beta[i]);
            p = isa_p(h_geop, isa::g_std, pb, Tb, hb, isa::beta[i]);
            return;
        }
        real Tc = isa_T(isa::h_ceil[i], Tb, hb, isa::beta[i]);
        real pc = isa_p(isa::h_ceil[i], isa::g_std, pb, Tb, hb, isa::beta[i]);
        hb = isa::h_ceil[i]; Tb = Tc; pb = pc;
    }
    throw OrcTermination(ST_ALT_BOUNDS);
}

struct AtmData { real T, p, rho, a, mu; V3 v; };
struct AirData { V3 v_ew_n, v_ew_b, v_wb_b; real T, p, rho, a, mu, M, Tt, pt, dp, q, TAS, EAS, CAS; };

// AtmosphericData(SimpleAtmosphere, Geographic(n_e, h_o)) — atmosphere.jl:269-278
inline AtmData atm_data(real h_o, real T_sl, real p_sl, V3 wind_n) {
    AtmData d;
    check_alt(h_o);
    real h_g = h_geop_from_orth(h_o);
    check_alt(h_g);
    isa_data(h_g, T_sl, p_sl, d.T, d.p);
    d.rho = d.p / (isa::R * d.T);
    d.a = sqrt(isa::gamma * isa::R * d.T);
    d.mu = (isa::betas * pow(d.T, 1.5)) / (d.T + isa::S);
    d.v = wind_n;
    return d;
}
// AirData(atm_data, kin_data) — atmosphere.jl:220-242
inline AirData air_data(const AtmData& at, V3 v_eb_b, Q4 q_nb) {
    AirData d;
    d.T = at.T; d.p = at.p; d.rho = at.rho; d.a = at.a; d.mu = at.mu;
    d.v_ew_n = at.v;
    d.v_ew_b = qrot_inv(q_nb, d.v_ew_n);
    d.v_wb_b = v_eb_b - d.v_ew_b;
    d.TAS = norm(d.v_wb_b);
    d.M = d.TAS / d.a;
    d.Tt = d.T * (1.0 + (isa::gamma - 1) / 2 * sq(d.M));
    d.pt = d.p * pow(d.Tt / d.T, isa::gamma / (isa::gamma - 1));
    d.dp = d.pt - d.p;
    d.q = 1.0 / 2 * d.rho * sq(d.TAS);
    d.EAS = d.TAS * sqrt(d.rho / isa::rho_std);
    d.CAS = sqrt(2 * isa::gamma / (isa::gamma - 1) * isa::p_std / isa::rho_std *
                 (pow(1.0 + d.dp / isa::p_std, (isa::gamma - 1) / isa::gamma) - 1.0));
    return d;
}
// atmosphere.jl:336-344
inline void airflow_angles(V3 v, real& alpha, real& beta) {
    if (norm(v) < 0.1) { alpha = 0.0; beta = 0.0; return; }
    alpha = atan2(v.z, v.x);
    beta = atan2(v.y, sqrt(v.x * v.x + v.z * v.z));
}
