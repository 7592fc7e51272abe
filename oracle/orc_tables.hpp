// TEST INFRASTRUCTURE ONLY (see orc_real.hpp).
// Lookup tables of the C172 path, built the way the reference builds them:
//   aero      : lib/FlightApps/src/c172/c172.jl:51-199 (JSBSim C172R data, gridded, Flat)
//   engine    : lib/FlightPhysics/src/piston.jl:70-195 (IO-360 charts)
//   propeller : lib/FlightPhysics/src/propellers.jl:46-208 (blade-element table generation)
// The numbers below are the reference's published data tables, restated.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <map>
#include "orc_interp.hpp"
#include "orc_math.hpp"

static inline std::vector<double> d2r(std::vector<double> v) { for (auto& x : v) x = x * ORC_PI / 180.0; return v; }

struct AeroTables {
    // scalars
    double CD_zero = 0.027, CY_dr = 0.1870, CY_da = 0.0, CL_de = 0.43, CL_q = 3.9, CL_adot = 1.7;
    double Cl_da = 0.229, Cl_dr = 0.0147, Cl_b = -0.09226, Cl_p = -0.4840;
    double Cm_zero = 0.1, Cm_de = -1.1220, Cm_a = -1.8, Cm_q = -12.4, Cm_adot = -7.27;
    double Cn_dr = -0.0430, Cn_da = -0.0053, Cn_b = 0.05874, Cn_p = -0.0278, Cn_r = -0.0937;
    Interp1 CD_b, CD_de, CD_df, CD_ge, CL_df, CL_ge, Cm_df;
    Interp2 CD_a_df, CY_b_df, CY_p, CY_r, CL_a, Cl_r;

    AeroTables() {
        std::vector<double> ge_x = {0.0, 0.1, 0.15, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1};
        std::vector<double> df_x = d2r({0, 10, 20, 30});
        std::vector<double> df2_x = d2r({0, 30});
        CD_de = {Axis({-1.0, 0.0, 1.0}, FLAT, FLAT), {0.06, 0.0, 0.06}};
        CD_b = {Axis({-1.0, 0.0, 1.0}, FLAT, FLAT), {0.17, 0.0, 0.17}};
        CD_ge = {Axis(ge_x, FLAT, FLAT), {0.48, 0.515, 0.629, 0.709, 0.815, 0.882, 0.928, 0.962, 0.988, 1.0, 1.0, 1.0, 1.0}};
        CD_df = {Axis(df_x, FLAT, FLAT), {0.0, 0.007, 0.012, 0.018}};
        std::vector<double> a_x = {-0.0873, -0.0698, -0.0524, -0.0349, -0.0175, 0.0, 0.0175, 0.0349, 0.0524, 0.0698, 0.0873, 0.1047, 0.1222,
                                   0.1396, 0.1571, 0.1745, 0.1920, 0.2094, 0.2269, 0.2443, 0.2618, 0.2793, 0.2967, 0.3142, 0.3316, 0.3491};
        static const double cd_a_df[4][26] = {
            {0.0041, 0.0013, 0.0001, 0.0003, 0.0020, 0.0052, 0.0099, 0.0162, 0.0240, 0.0334, 0.0442, 0.0566, 0.0706, 0.0860, 0.0962, 0.1069, 0.1180, 0.1298, 0.1424, 0.1565, 0.1727, 0.1782, 0.1716, 0.1618, 0.1475, 0.1097},
            {0.0000, 0.0004, 0.0023, 0.0057, 0.0105, 0.0168, 0.0248, 0.0342, 0.0452, 0.0577, 0.0718, 0.0874, 0.1045, 0.1232, 0.1353, 0.1479, 0.1610, 0.1746, 0.1892, 0.2054, 0.2240, 0.2302, 0.2227, 0.2115, 0.1951, 0.1512},
            {0.0005, 0.0025, 0.0059, 0.0108, 0.0172, 0.0251, 0.0346, 0.0457, 0.0583, 0.0724, 0.0881, 0.1053, 0.1240, 0.1442, 0.1573, 0.1708, 0.1849, 0.1995, 0.2151, 0.2323, 0.2521, 0.2587, 0.2507, 0.2388, 0.2214, 0.1744},
            {0.0014, 0.0041, 0.0084, 0.0141, 0.0212, 0.0299, 0.0402, 0.0521, 0.0655, 0.0804, 0.0968, 0.1148, 0.1343, 0.1554, 0.1690, 0.1830, 0.1975, 0.2126, 0.2286, 0.2464, 0.2667, 0.2735, 0.2653, 0.2531, 0.2351, 0.1866}};
        CD_a_df.a0 = Axis(a_x, FLAT, FLAT); CD_a_df.a1 = Axis(df_x, FLAT, FLAT);
        CD_a_df.y.resize(26 * 4);
        for (int i = 0; i < 26; i++) for (int j = 0; j < 4; j++) CD_a_df.y[i * 4 + j] = cd_a_df[j][i];
        CY_b_df = {Axis({-0.3490, 0.0, 0.3490}, FLAT, FLAT), Axis(df2_x, FLAT, FLAT), {0.1370, 0.1060, 0.0, 0.0, -0.1370, -0.1060}};
        CY_p = {Axis({0.0, 0.094}, FLAT, FLAT), Axis(df2_x, FLAT, FLAT), {-0.0750, -0.1610, -0.1450, -0.2310}};
        CY_r = {Axis({0.0, 0.094}, FLAT, FLAT), Axis(df2_x, FLAT, FLAT), {0.2140, 0.1620, 0.2670, 0.2150}};
        CL_ge = {Axis(ge_x, FLAT, FLAT), {1.2030, 1.1270, 1.0900, 1.0730, 1.0460, 1.0550, 1.0190, 1.0130, 1.0080, 1.0060, 1.0030, 1.0020, 1.0000}};
        std::vector<double> cl_a_x = {-0.09, 0.0, 0.09, 0.1, 0.12, 0.14, 0.16, 0.17, 0.19, 0.21, 0.24, 0.26, 0.28, 0.3, 0.32, 0.34, 0.36};
        static const double cl_a[2][17] = {
            {-0.22, 0.25, 0.73, 0.83, 0.92, 1.02, 1.08, 1.13, 1.19, 1.25, 1.35, 1.44, 1.47, 1.43, 1.38, 1.30, 1.15},
            {-0.22, 0.25, 0.73, 0.78, 0.79, 0.81, 0.82, 0.83, 0.85, 0.86, 0.88, 0.90, 0.92, 0.95, 0.99, 1.05, 1.15}};
        CL_a.a0 = Axis(cl_a_x, FLAT, FLAT); CL_a.a1 = Axis({0.0, 1.0}, FLAT, FLAT);
        CL_a.y.resize(17 * 2);
        for (int i = 0; i < 17; i++) for (int j = 0; j < 2; j++) CL_a.y[i * 2 + j] = cl_a[j][i];
        CL_df = {Axis(df_x, FLAT, FLAT), {0.0, 0.2, 0.3, 0.35}};
        Cl_r = {Axis({0.0, 0.094}, FLAT, FLAT), Axis(df2_x, FLAT, FLAT), {0.0798, 0.1246, 0.1869, 0.2317}};
        Cm_df = {Axis(df_x, FLAT, FLAT), {0.0, -0.0654, -0.0981, -0.1140}};
    }
};

// --------------------------------------------------------------- piston engine lookups
struct EngineTables {
    Interp2 d_wot, mu_wot, pi_std, pi_wot, sfc_pow;
    Interp1 pi_ratio, sfc_ratio;
    static constexpr double f_cutoff = 0.0580, f_lean = 0.0625, f_rich = 0.0950;

    EngineTables(double n_stall, double n_max) {  // piston.jl:70-195
        assert(n_stall < 0.667 && n_max > 1.074);
        Axis n_rng = Axis::range(0.667, 1.0, 2, LINE, LINE);
        Axis mu_rng = Axis::range(0.401, 0.936, 9, LINE, LINE);
        static const double dd[2][9] = {{0.455, 0.523, 0.587, 0.652, 0.718, 0.781, 0.844, 0.906, 0.965},
                                        {0.464, 0.530, 0.596, 0.662, 0.727, 0.792, 0.855, 0.921, 0.981}};
        d_wot.a0 = n_rng; d_wot.a1 = mu_rng; d_wot.y.resize(18);
        for (int i = 0; i < 2; i++) for (int j = 0; j < 9; j++) d_wot.y[i * 9 + j] = dd[i][j];
        // mu_wot by inverse interpolation (piston.jl:92-104)
        Axis d_rng = Axis::range(0.441, 1.0, 9, LINE, LINE);
        mu_wot.a0 = n_rng; mu_wot.a1 = d_rng; mu_wot.y.resize(18);
        for (int i = 0; i < 2; i++) {
            std::vector<double> dk(9);
            for (int j = 0; j < 9; j++) dk[j] = val(d_wot(n_rng.knots[i], mu_rng.knots[j]));
            Interp1 inv{Axis(dk, LINE, LINE), mu_rng.knots};
            for (int j = 0; j < 9; j++) mu_wot.y[i * 9 + j] = val(inv(d_rng.knots[j]));
        }
        // pi_std (piston.jl:108-134)
        std::vector<double> n_data = {n_stall, 0.667, 0.704, 0.741, 0.778, 0.815, 0.852, 0.889, 0.926, 0.963, 1.000, 1.074, n_max};
        std::vector<double> mu_data = {0, 0.568, 1.0};
        static const double mk2[13] = {1.000, 0.836, 0.854, 0.874, 0.898, 0.912, 0.939, 0.961, 0.959, 0.958, 0.956, 0.953, 1.000};
        static const double pk1[13] = {0, 0.270, 0.305, 0.335, 0.360, 0.380, 0.405, 0.428, 0.450, 0.476, 0.498, 0.498, 0};
        static const double pk2[13] = {0, 0.489, 0.548, 0.609, 0.680, 0.729, 0.810, 0.880, 0.920, 0.965, 1.000, 0.950, 0};
        pi_std.a0 = Axis(n_data, FLAT, FLAT); pi_std.a1 = Axis(mu_data, FLAT, FLAT); pi_std.y.resize(13 * 3);
        for (int i = 0; i < 13; i++) {
            Interp1 t{Axis({0.0, 0.568, mk2[i]}, LINE, LINE), {0.0, pk1[i], pk2[i]}};
            for (int j = 0; j < 3; j++) pi_std.y[i * 3 + j] = val(t(mu_data[j]));
        }
        // pi_wot (piston.jl:138-151)
        std::vector<double> n5 = {n_stall, 0.667, 1.000, 1.074, n_max};
        static const double col2[5] = {0, 0.23, 0.409, 0.409, 0};
        pi_wot.a0 = Axis(n5, FLAT, FLAT); pi_wot.a1 = Axis({0.0, 0.441, 1.0}, FLAT, LINE); pi_wot.y.resize(15);
        for (int i = 0; i < 5; i++) {
            pi_wot.y[i * 3 + 0] = 0.0;
            pi_wot.y[i * 3 + 1] = col2[i];
            pi_wot.y[i * 3 + 2] = val(pi_std(n5[i], mu_wot(n5[i], 1.0)));
        }
        std::vector<double> f_data(11);
        f_data[0] = f_cutoff;
        for (int i = 0; i < 10; i++) f_data[1 + i] = f_lean + (f_rich - f_lean) * i / 9.0;
        f_data[10] = f_rich;
        pi_ratio = {Axis(f_data, FLAT, FLAT), {0.000, 0.8600, 0.9492, 0.9776, 0.9933, 1.000, 0.9983, 0.9910, 0.9798, 0.9657, 0.9500}};
        sfc_ratio = {Axis(f_data, FLAT, FLAT), {5, 0.8700, 0.8524, 0.8818, 0.9261, 0.9839, 1.0510, 1.1279, 1.2135, 1.3163, 1.4280}};
        std::vector<double> n_sfc = {2000.0 / 2700, 2200.0 / 2700, 2400.0 / 2700, 2600.0 / 2700, 2700.0 / 2700};
        std::vector<double> pi_sfc(8);
        for (int i = 0; i < 8; i++) pi_sfc[i] = std::pow(10.0, -1.0 + (double)i / 7.0);
        pi_sfc[7] = 1.0;
        static const double sfc[5][8] = {
            {1.7671, 1.43728, 1.19992, 1.02909, 0.906153, 0.817674, 0.753997, 0.708169},
            {1.83791, 1.49664, 1.25103, 1.07427, 0.947056, 0.855503, 0.789613, 0.742193},
            {1.98614, 1.60588, 1.3322, 1.13524, 0.993496, 0.891482, 0.818064, 0.765226},
            {2.11663, 1.70062, 1.40123, 1.18576, 1.03069, 0.919083, 0.838765, 0.780961},
            {2.33484, 1.85418, 1.50825, 1.2593, 1.08012, 0.951177, 0.858376, 0.791588}};
        sfc_pow.a0 = Axis(n_sfc, LINE, LINE); sfc_pow.a1 = Axis(pi_sfc, LINE, LINE); sfc_pow.y.resize(40);
        for (int i = 0; i < 5; i++) for (int j = 0; j < 8; j++) sfc_pow.y[i * 8 + j] = 1e-7 * sfc[i][j];
    }
};

// --------------------------------------------------------------- propeller table generation
// propellers.jl:46-94 (DefaultAirfoil), :96-101 (Blade), :131-208 (Coefficients + induced-angle eq.)
namespace prop {
inline double cL(double al, double M);
inline double cL(double al, double M) {
    if (M <= 0.8) return (al < 0.25 ? 2 * ORC_PI * al : ORC_PI / 2 * std::cos(al) / std::cos(0.25)) / std::sqrt(1 - M * M);
    if (M >= 1.2) return (al < 0.25 ? 4 * al : std::cos(al) / std::cos(0.25)) / std::sqrt(M * M - 1);
    return cL(al, 0.8) + (cL(al, 1.2) - cL(al, 0.8)) / 0.4 * (M - 0.8);
}
inline double cL_a(double al, double M) {
    if (M <= 0.8) return (al < 0.25 ? 2 * ORC_PI : -ORC_PI / 2 * std::sin(al) / std::cos(0.25)) / std::sqrt(1 - M * M);
    if (M >= 1.2) return (al < 0.25 ? 4.0 : -std::sin(al) / std::cos(0.25)) / std::sqrt(M * M - 1);
    return cL_a(al, 0.8) + (cL_a(al, 1.2) - cL_a(al, 0.8)) / 0.4 * (M - 0.8);
}
inline double cD(double al, double M) {
    double inc;
    if (al < 0.25) inc = 0.006 + 0.224 * al * al;
    else if (al < 0.3) inc = -1.0234 + 16.6944 * al * al;
    else inc = ORC_PI / 2 * std::sin(al) / std::cos(0.25);
    double k;
    if (M <= 0.8) k = 1.0;
    else if (M <= 0.95) k = 1.0 + 160000 * std::pow(M - 0.8, 4) / 27;
    else if (M <= 1.0) k = 6.0 - 800 * (1 - M) * (1 - M);
    else k = 6 - 5 * (M - 1);
    return k * inc;
}
struct BladeP { double zeta_h = 0.2, chord_a = 0.075, pitch = 0.8, alpha0 = -2.1 * ORC_PI / 180.0; };
inline double M_section(double J, double Mt, double z, double ei) {
    return Mt * std::sqrt((ORC_PI * ORC_PI * z * z + J * J) / (ORC_PI * ORC_PI + J * J)) * std::cos(ei);
}
inline double induced_eq(int nb, double c, double ba_t, double J, double Mt, double ba, double e_inf, double z, double ei) {
    double al = ba - e_inf - ei;
    double M = M_section(J, Mt, z, ei);
    return nb * c / (8 * z) * cL(al, M) - std::acos(std::exp(-nb * (1 - z) / (2 * std::sin(ba_t)))) * std::tan(ei) * std::sin(e_inf + ei);
}
// Root of f near x0. Roots.find_zero(f, x0) (Roots "3", Order0: secant iterations with a bracketing
// fallback) converges to the root nearest the start to machine precision; here a secant iteration
// polished by bisection on the final bracket reaches the same root to <= 1 ulp-level tolerance.
template <class F>
inline double find_zero(F f, double x0) {
    double x_a = x0, f_a = f(x_a);
    if (f_a == 0.0) return x_a;
    double h = std::fabs(x0) * 1e-4 + 1e-6;
    double x_b = x0 + h, f_b = f(x_b);
    for (int it = 0; it < 200; it++) {
        if (f_b == 0.0) return x_b;
        double denom = (f_b - f_a);
        if (denom == 0.0) break;
        double x_c = x_b - f_b * (x_b - x_a) / denom;
        x_a = x_b; f_a = f_b;
        x_b = x_c; f_b = f(x_b);
        if (std::fabs(x_b - x_a) <= 4e-16 * std::fabs(x_b)) break;
    }
    // polish: if the last two iterates bracket the root, bisect to the last bit
    if (f_a * f_b < 0) {
        double lo = x_a, hi = x_b, flo = f_a;
        for (int it = 0; it < 200; it++) {
            double mid = 0.5 * (lo + hi);
            if (mid == lo || mid == hi) break;
            double fm = f(mid);
            if (fm == 0.0) return mid;
            if ((fm < 0) == (flo < 0)) { lo = mid; flo = fm; } else hi = mid;
        }
        return std::fabs(f(lo)) < std::fabs(f(hi)) ? lo : hi;
    }
    return x_b;
}
struct Coefs { double C_Fx, C_Mx, C_Fz_a, C_Mz_a, C_P, eta_p; };
inline Coefs coefficients(int nb, const BladeP& bl, double J, double Mt, double dbeta, int n_z = 101) {
    std::vector<double> z(n_z), dFx(n_z), dMx(n_z), dFz(n_z), dMz(n_z);
    for (int i = 0; i < n_z; i++) z[i] = bl.zeta_h + (1.0 - bl.zeta_h) * i / (n_z - 1);
    z[n_z - 1] = 1.0;
    auto beta_a = [&](double zz) { return std::atan(bl.pitch / (ORC_PI * zz)) + dbeta - bl.alpha0; };
    double ba_t = beta_a(1.0);
    double ei = 1.0;
    const double pi2 = ORC_PI * ORC_PI;
    for (int i = 0; i < n_z; i++) {
        double zz = z[i];
        double e_inf = std::atan(J / (ORC_PI * zz));
        double ba = beta_a(zz);
        double c = bl.chord_a * std::sqrt(1 - zz * zz);
        ei = find_zero([&](double e) { return induced_eq(nb, c, ba_t, J, Mt, ba, e_inf, zz, e); }, ei);
        double e = e_inf + ei;
        double al = ba - e;
        double M = M_section(J, Mt, zz, ei);
        double kc = nb * c;
        double z2 = zz * zz, z3 = z2 * zz;
        double ce = std::cos(e), se = std::sin(e);
        double c2i = std::cos(ei) * std::cos(ei), c2inf = std::cos(e_inf) * std::cos(e_inf);
        double t_inf = std::tan(e_inf), t2_inf = t_inf * t_inf;
        double l = cL(al, M), d = cD(al, M), la = cL_a(al, M);
        dFx[i] = pi2 / 4 * z2 * kc * c2i / c2inf * (l * ce - d * se);
        dMx[i] = -pi2 / 8 * z3 * kc * c2i / c2inf * (d * ce + l * se);
        dFz[i] = -pi2 / 8 * z2 * kc * c2i * (2 * t_inf * (d * ce + l * se) - t2_inf * (l * ce - (la + d) * se));
        dMz[i] = -pi2 / 16 * z3 * kc * c2i * (2 * t_inf * (l * ce - d * se) + t2_inf * ((la + d) * ce + l * se));
    }
    auto trapz = [&](const std::vector<double>& y) {
        double s = 0;
        for (int i = 0; i + 1 < n_z; i++) s += (z[i + 1] - z[i]) * (y[i] + y[i + 1]) / 2;
        return s;
    };
    Coefs k;
    k.C_Fx = trapz(dFx); k.C_Mx = trapz(dMx); k.C_Fz_a = trapz(dFz); k.C_Mz_a = trapz(dMz);
    k.C_P = 2 * ORC_PI * k.C_Mx;
    k.eta_p = (k.C_Fx > 0 ? -J * k.C_Fx / k.C_P : 0.0);
    return k;
}
}  // namespace prop

struct PropTables {  // 21x21 over J in [0,1.5], Mt in [0,1.5]; scaled, Flat (propellers.jl:235-273)
    Interp2 C_Fx, C_Mx, C_Fz_a, C_Mz_a, C_P, eta_p;
    PropTables() {
        const int n = 21;
        Axis aj = Axis::range(0, 1.5, n, FLAT, FLAT), am = Axis::range(0, 1.5, n, FLAT, FLAT);
        Interp2* all[6] = {&C_Fx, &C_Mx, &C_Fz_a, &C_Mz_a, &C_P, &eta_p};
        for (auto* t : all) { t->a0 = aj; t->a1 = am; t->y.resize(n * n); }
        prop::BladeP bl;
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
            prop::Coefs k = prop::coefficients(2, bl, aj.knots[i], am.knots[j], 0.0);
            C_Fx.y[i * n + j] = k.C_Fx; C_Mx.y[i * n + j] = k.C_Mx; C_Fz_a.y[i * n + j] = k.C_Fz_a;
            C_Mz_a.y[i * n + j] = k.C_Mz_a; C_P.y[i * n + j] = k.C_P; eta_p.y[i * n + j] = k.eta_p;
        }
    }
};

// Lookup of a Propeller(VariablePitch(lo, hi; length = n)) (propellers.jl:235-251, 341-356): the same blade-element
// coefficients over a third, blade-pitch-offset axis. Built on demand (the C172 itself flies a FixedPitch propeller).
struct PropTables3 {
    Interp3 C_Fx, C_Mx, C_Fz_a, C_Mz_a, C_P, eta_p;
    double lo, hi;
    PropTables3(double lo_, double hi_, int nb = 11, int n_blades = 2) : lo(lo_), hi(hi_) {
        const int n = 21;
        Axis aj = Axis::range(0, 1.5, n, FLAT, FLAT), am = Axis::range(0, 1.5, n, FLAT, FLAT), ab = Axis::range(lo, hi, nb, FLAT, FLAT);
        Interp3* all[6] = {&C_Fx, &C_Mx, &C_Fz_a, &C_Mz_a, &C_P, &eta_p};
        for (auto* t : all) { t->a0 = aj; t->a1 = am; t->a2 = ab; t->y.resize((size_t)n * n * nb); }
        prop::BladeP bl;
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < nb; k++) {
            prop::Coefs c = prop::coefficients(n_blades, bl, aj.knots[i], am.knots[j], ab.knots[k]);
            const size_t o = ((size_t)i * n + j) * nb + k;
            C_Fx.y[o] = c.C_Fx; C_Mx.y[o] = c.C_Mx; C_Fz_a.y[o] = c.C_Fz_a; C_Mz_a.y[o] = c.C_Mz_a; C_P.y[o] = c.C_P; eta_p.y[o] = c.eta_p;
        }
    }
};
