#!/usr/bin/env python
"""Build-time data tool: writes flight.jl_b200/data/c172_tables.bin, the table blob that
fb_load_tables() takes.

It gathers every lookup table the C172 hot path reads:
  * aerodynamic coefficient tables   — data of lib/FlightApps/src/c172/c172.jl:51-159 (JSBSim C172R)
  * IO-360 engine charts             — data of lib/FlightPhysics/src/piston.jl:76-191, with the derived
                                       tables (mu_wot inverse, pi_std, pi_wot) built the same way
  * propeller coefficient tables     — regenerated with a blade-element solver restating
                                       lib/FlightPhysics/src/propellers.jl:46-208 (numpy/scipy; an
                                       implementation independent of the oracle's C++ one — the test
                                       suite checks that the two agree)
  * EGM96 geoid (721x1441 float32)   — lib/FlightPhysics/src/data/ww15mgh_le.bin (SHA-256 checked)
  * gain schedules                   — the ten HDF5 files under lib/FlightApps/src/c172/c172x/control/data

Run here (needs /root/reference); the blob is committed so nothing on the GPU box reads the reference.

Blob layout: magic "FBTABLE1", uint32 n_sections, uint32 0, then n x 64-byte section records
{char name[24]; uint32 dtype (0=f64, 1=f32); uint32 ndim; uint32 dims[4]; uint64 offset; uint64 nbytes},
then 16-byte aligned little-endian payloads in C order.
"""
import hashlib
import os
import struct
import sys

import numpy as np
from scipy import optimize

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import h5mini  # noqa: E402

REF = os.environ.get("FLIGHT_REF", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "flight.jl_b200", "data", "c172_tables.bin")


# ----------------------------------------------------------------------------- interpolation helpers
def interp1_line(xk, yk, x):
    """linear_interpolation(xk, yk, extrapolation_bc=Line()) for scalar x."""
    xk = np.asarray(xk, float); yk = np.asarray(yk, float)
    i = int(np.clip(np.searchsorted(xk, x, side="right") - 1, 0, len(xk) - 2))
    return yk[i] + (yk[i + 1] - yk[i]) / (xk[i + 1] - xk[i]) * (x - xk[i])


def interp2_scaled_line(r0, r1, A, x0, x1):
    """extrapolate(scale(interpolate(A, BSpline(Linear())), r0, r1), Line()) at a point inside the grid
    or on its boundary (all the construction-time calls are in range)."""
    i = int(np.clip(np.searchsorted(r0, x0, side="right") - 1, 0, len(r0) - 2))
    j = int(np.clip(np.searchsorted(r1, x1, side="right") - 1, 0, len(r1) - 2))
    u = (x0 - r0[i]) / (r0[i + 1] - r0[i]); w = (x1 - r1[j]) / (r1[j + 1] - r1[j])
    return (1 - u) * ((1 - w) * A[i, j] + w * A[i, j + 1]) + u * ((1 - w) * A[i + 1, j] + w * A[i + 1, j + 1])


def interp2_gridded_flat(k0, k1, A, x0, x1):
    x0 = min(max(x0, k0[0]), k0[-1]); x1 = min(max(x1, k1[0]), k1[-1])
    return interp2_scaled_line(np.asarray(k0), np.asarray(k1), A, x0, x1)


# ----------------------------------------------------------------------------- aero (c172.jl:51-159)
def aero_tables():
    d2r = np.pi / 180
    t = {}
    t["aero.ge_x"] = [0.0, 0.1, 0.15, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1]
    t["aero.CD_ge"] = [0.48, 0.515, 0.629, 0.709, 0.815, 0.882, 0.928, 0.962, 0.988, 1.0, 1.0, 1.0, 1.0]
    t["aero.CL_ge"] = [1.203, 1.127, 1.09, 1.073, 1.046, 1.055, 1.019, 1.013, 1.008, 1.006, 1.003, 1.002, 1.0]
    t["aero.df_x"] = [0 * d2r, 10 * d2r, 20 * d2r, 30 * d2r]
    t["aero.CD_df"] = [0.0, 0.007, 0.012, 0.018]
    t["aero.CL_df"] = [0.0, 0.2, 0.3, 0.35]
    t["aero.Cm_df"] = [0.0, -0.0654, -0.0981, -0.1140]
    t["aero.CD_de"] = [0.06, 0.0, 0.06]      # knots -1, 0, 1
    t["aero.CD_b"] = [0.17, 0.0, 0.17]       # knots -1, 0, 1
    t["aero.CD_a_x"] = [-0.0873, -0.0698, -0.0524, -0.0349, -0.0175, 0.0, 0.0175, 0.0349, 0.0524, 0.0698, 0.0873, 0.1047,
                        0.1222, 0.1396, 0.1571, 0.1745, 0.1920, 0.2094, 0.2269, 0.2443, 0.2618, 0.2793, 0.2967, 0.3142,
                        0.3316, 0.3491]
    cd = np.array([
        [0.0041, 0.0013, 0.0001, 0.0003, 0.0020, 0.0052, 0.0099, 0.0162, 0.0240, 0.0334, 0.0442, 0.0566, 0.0706, 0.0860, 0.0962, 0.1069, 0.1180, 0.1298, 0.1424, 0.1565, 0.1727, 0.1782, 0.1716, 0.1618, 0.1475, 0.1097],
        [0.0000, 0.0004, 0.0023, 0.0057, 0.0105, 0.0168, 0.0248, 0.0342, 0.0452, 0.0577, 0.0718, 0.0874, 0.1045, 0.1232, 0.1353, 0.1479, 0.1610, 0.1746, 0.1892, 0.2054, 0.2240, 0.2302, 0.2227, 0.2115, 0.1951, 0.1512],
        [0.0005, 0.0025, 0.0059, 0.0108, 0.0172, 0.0251, 0.0346, 0.0457, 0.0583, 0.0724, 0.0881, 0.1053, 0.1240, 0.1442, 0.1573, 0.1708, 0.1849, 0.1995, 0.2151, 0.2323, 0.2521, 0.2587, 0.2507, 0.2388, 0.2214, 0.1744],
        [0.0014, 0.0041, 0.0084, 0.0141, 0.0212, 0.0299, 0.0402, 0.0521, 0.0655, 0.0804, 0.0968, 0.1148, 0.1343, 0.1554, 0.1690, 0.1830, 0.1975, 0.2126, 0.2286, 0.2464, 0.2667, 0.2735, 0.2653, 0.2531, 0.2351, 0.1866]])
    t["aero.CD_a_df"] = cd.T.copy()          # [26 alpha][4 flap]
    t["aero.CY_b_df"] = [[0.1370, 0.1060], [0.0, 0.0], [-0.1370, -0.1060]]  # beta knots -0.349,0,0.349 ; df 0,30deg
    t["aero.CY_p"] = [[-0.0750, -0.1610], [-0.1450, -0.2310]]               # alpha knots 0,0.094 ; df 0,30deg
    t["aero.CY_r"] = [[0.2140, 0.1620], [0.2670, 0.2150]]
    t["aero.Cl_r"] = [[0.0798, 0.1246], [0.1869, 0.2317]]
    t["aero.CL_a_x"] = [-0.09, 0.0, 0.09, 0.1, 0.12, 0.14, 0.16, 0.17, 0.19, 0.21, 0.24, 0.26, 0.28, 0.3, 0.32, 0.34, 0.36]
    cl = np.array([
        [-0.22, 0.25, 0.73, 0.83, 0.92, 1.02, 1.08, 1.13, 1.19, 1.25, 1.35, 1.44, 1.47, 1.43, 1.38, 1.30, 1.15],
        [-0.22, 0.25, 0.73, 0.78, 0.79, 0.81, 0.82, 0.83, 0.85, 0.86, 0.88, 0.90, 0.92, 0.95, 0.99, 1.05, 1.15]])
    t["aero.CL_a"] = cl.T.copy()             # [17 alpha][2 stall]
    # scalar derivatives, c172.jl:52-159
    t["aero.scalars"] = [0.027, 0.1870, 0.0, 0.43, 3.9, 1.7,        # CD_zero, CY_dr, CY_da, CL_de, CL_q, CL_adot
                         0.229, 0.0147, -0.09226, -0.4840,          # Cl_da, Cl_dr, Cl_b, Cl_p
                         0.1, -1.1220, -1.8, -12.4, -7.27,          # Cm_zero, Cm_de, Cm_a, Cm_q, Cm_adot
                         -0.0430, -0.0053, 0.05874, -0.0278, -0.0937]  # Cn_dr, Cn_da, Cn_b, Cn_p, Cn_r
    return {k: np.asarray(v, float) for k, v in t.items()}


# ----------------------------------------------------------------------------- engine (piston.jl:70-195)
def engine_tables(n_stall=300.0 / 2700.0, n_max=3100.0 / 2700.0):
    t = {}
    n_rng = np.linspace(0.667, 1.0, 2); mu_rng = np.linspace(0.401, 0.936, 9); d_rng = np.linspace(0.441, 1.0, 9)
    d_wot = np.array([[0.455, 0.523, 0.587, 0.652, 0.718, 0.781, 0.844, 0.906, 0.965],
                      [0.464, 0.530, 0.596, 0.662, 0.727, 0.792, 0.855, 0.921, 0.981]])
    mu_wot = np.zeros((2, 9))
    for i in range(2):
        dk = [interp2_scaled_line(n_rng, mu_rng, d_wot, n_rng[i], m) for m in mu_rng]
        mu_wot[i] = [interp1_line(dk, mu_rng, d) for d in d_rng]
    n_data = np.array([n_stall, 0.667, 0.704, 0.741, 0.778, 0.815, 0.852, 0.889, 0.926, 0.963, 1.000, 1.074, n_max])
    mu_data = np.array([0, 0.568, 1.0])
    mk2 = [1.000, 0.836, 0.854, 0.874, 0.898, 0.912, 0.939, 0.961, 0.959, 0.958, 0.956, 0.953, 1.000]
    pk1 = [0, 0.270, 0.305, 0.335, 0.360, 0.380, 0.405, 0.428, 0.450, 0.476, 0.498, 0.498, 0]
    pk2 = [0, 0.489, 0.548, 0.609, 0.680, 0.729, 0.810, 0.880, 0.920, 0.965, 1.000, 0.950, 0]
    pi_std = np.zeros((13, 3))
    for i in range(13):
        pi_std[i] = [interp1_line([0.0, 0.568, mk2[i]], [0.0, pk1[i], pk2[i]], m) for m in mu_data]
    n5 = np.array([n_stall, 0.667, 1.000, 1.074, n_max])
    pi_wot = np.zeros((5, 3))
    pi_wot[:, 1] = [0, 0.23, 0.409, 0.409, 0]
    for i, n in enumerate(n5):
        # mu_wot(n, 1) with Line extrapolation in n (scaled BSpline): in-range in delta (=1, upper knot)
        nn = n
        u = (nn - n_rng[0]) / (n_rng[1] - n_rng[0])          # may be outside [0,1]: Line extrapolation in n
        m = (1 - u) * mu_wot[0, -1] + u * mu_wot[1, -1]
        pi_wot[i, 2] = interp2_gridded_flat(n_data, mu_data, pi_std, n, m)
    f_data = np.concatenate([[0.0580], np.linspace(0.0625, 0.0950, 10)])
    t["eng.d_wot"] = d_wot; t["eng.mu_wot"] = mu_wot
    t["eng.n_std"] = n_data; t["eng.mu_std"] = mu_data; t["eng.pi_std"] = pi_std
    t["eng.n_wot"] = n5; t["eng.d_wot_x"] = np.array([0, 0.441, 1.0]); t["eng.pi_wot"] = pi_wot
    t["eng.f_x"] = f_data
    t["eng.pi_ratio"] = np.array([0.000, 0.8600, 0.9492, 0.9776, 0.9933, 1.000, 0.9983, 0.9910, 0.9798, 0.9657, 0.9500])
    t["eng.sfc_ratio"] = np.array([5, 0.8700, 0.8524, 0.8818, 0.9261, 0.9839, 1.0510, 1.1279, 1.2135, 1.3163, 1.4280])
    t["eng.n_sfc"] = np.array([2000, 2200, 2400, 2600, 2700]) / 2700.0
    t["eng.pi_sfc"] = 10.0 ** np.linspace(-1, 0, 8)
    t["eng.sfc_pow"] = 1e-7 * np.array([
        [1.7671, 1.43728, 1.19992, 1.02909, 0.906153, 0.817674, 0.753997, 0.708169],
        [1.83791, 1.49664, 1.25103, 1.07427, 0.947056, 0.855503, 0.789613, 0.742193],
        [1.98614, 1.60588, 1.3322, 1.13524, 0.993496, 0.891482, 0.818064, 0.765226],
        [2.11663, 1.70062, 1.40123, 1.18576, 1.03069, 0.919083, 0.838765, 0.780961],
        [2.33484, 1.85418, 1.50825, 1.2593, 1.08012, 0.951177, 0.858376, 0.791588]])
    return t


# ----------------------------------------------------------------------------- propeller (propellers.jl:46-208)
def _cL(al, M):
    if M <= 0.8:
        return (2 * np.pi * al if al < 0.25 else np.pi / 2 * np.cos(al) / np.cos(0.25)) / np.sqrt(1 - M * M)
    if M >= 1.2:
        return (4 * al if al < 0.25 else np.cos(al) / np.cos(0.25)) / np.sqrt(M * M - 1)
    return _cL(al, 0.8) + (_cL(al, 1.2) - _cL(al, 0.8)) / 0.4 * (M - 0.8)


def _cL_a(al, M):
    if M <= 0.8:
        return (2 * np.pi if al < 0.25 else -np.pi / 2 * np.sin(al) / np.cos(0.25)) / np.sqrt(1 - M * M)
    if M >= 1.2:
        return (4.0 if al < 0.25 else -np.sin(al) / np.cos(0.25)) / np.sqrt(M * M - 1)
    return _cL_a(al, 0.8) + (_cL_a(al, 1.2) - _cL_a(al, 0.8)) / 0.4 * (M - 0.8)


def _cD(al, M):
    if al < 0.25:
        inc = 0.006 + 0.224 * al ** 2
    elif al < 0.3:
        inc = -1.0234 + 16.6944 * al ** 2
    else:
        inc = np.pi / 2 * np.sin(al) / np.cos(0.25)
    if M <= 0.8:
        k = 1.0
    elif M <= 0.95:
        k = 1.0 + 160000 * (M - 0.8) ** 4 / 27
    elif M <= 1.0:
        k = 6.0 - 800 * (1 - M) ** 2
    else:
        k = 6 - 5 * (M - 1)
    return k * inc


def prop_coefficients(J, Mt, n_blades=2, zeta_h=0.2, chord=0.075, pitch=0.8, alpha0=np.deg2rad(-2.1), dbeta=0.0, n_z=101):
    z = np.linspace(zeta_h, 1.0, n_z)
    beta_a = lambda zz: np.arctan(pitch / (np.pi * zz)) + dbeta - alpha0
    ba_t = beta_a(1.0)
    dFx = np.zeros(n_z); dMx = np.zeros(n_z); dFz = np.zeros(n_z); dMz = np.zeros(n_z)
    ei = 1.0
    for i, zz in enumerate(z):
        e_inf = np.arctan(J / (np.pi * zz))
        ba = beta_a(zz)
        c = chord * np.sqrt(max(0.0, 1 - zz * zz))

        def Msec(e):
            return Mt * np.sqrt((np.pi ** 2 * zz ** 2 + J ** 2) / (np.pi ** 2 + J ** 2)) * np.cos(e)

        def f(e):
            return (n_blades * c / (8 * zz) * _cL(ba - e_inf - e, Msec(e))
                    - np.arccos(np.exp(-n_blades * (1 - zz) / (2 * np.sin(ba_t)))) * np.tan(e) * np.sin(e_inf + e))
        if f(ei) != 0.0:
            ei = optimize.newton(f, ei, tol=1e-15, maxiter=200)   # secant from the previous station's root
        e = e_inf + ei
        al = ba - e
        M = Msec(ei)
        kc = n_blades * c
        ce, se = np.cos(e), np.sin(e)
        c2i, c2inf = np.cos(ei) ** 2, np.cos(e_inf) ** 2
        ti = np.tan(e_inf)
        l, d, la = _cL(al, M), _cD(al, M), _cL_a(al, M)
        dFx[i] = np.pi ** 2 / 4 * zz ** 2 * kc * c2i / c2inf * (l * ce - d * se)
        dMx[i] = -np.pi ** 2 / 8 * zz ** 3 * kc * c2i / c2inf * (d * ce + l * se)
        dFz[i] = -np.pi ** 2 / 8 * zz ** 2 * kc * c2i * (2 * ti * (d * ce + l * se) - ti ** 2 * (l * ce - (la + d) * se))
        dMz[i] = -np.pi ** 2 / 16 * zz ** 3 * kc * c2i * (2 * ti * (l * ce - d * se) + ti ** 2 * ((la + d) * ce + l * se))
    tz = lambda y: float(np.sum(np.diff(z) * (y[:-1] + y[1:]) / 2))
    C_Fx, C_Mx, C_Fz, C_Mz = tz(dFx), tz(dMx), tz(dFz), tz(dMz)
    C_P = 2 * np.pi * C_Mx
    eta = -J * C_Fx / C_P if C_Fx > 0 else 0.0
    return C_Fx, C_Mx, C_Fz, C_Mz, C_P, eta


def prop_tables(n=21):
    Jr = np.linspace(0, 1.5, n); Mr = np.linspace(0, 1.5, n)
    out = np.zeros((6, n, n))
    for i, J in enumerate(Jr):
        for j, Mt in enumerate(Mr):
            out[:, i, j] = prop_coefficients(J, Mt)
    return {"prop.tables": out}


def prop_tables3(lo=-0.1, hi=0.2, nb=11, n=21):
    """Lookup of Propeller(VariablePitch(lo, hi; length = nb)) (propellers.jl:235-251, 341-356): the same blade-element
    coefficients over a third axis, the blade-pitch offset. Default range: the reference's own VariablePitch test case
    (lib/FlightPhysics/test/test_propellers.jl:152). Layout [iJ][iMt][ib][6]; only FB_C172XA2 reads it."""
    Jr = np.linspace(0, 1.5, n); Mr = np.linspace(0, 1.5, n); Br = np.linspace(lo, hi, nb)
    out = np.zeros((n, n, nb, 6))
    for i, J in enumerate(Jr):
        for j, Mt in enumerate(Mr):
            for k, db in enumerate(Br):
                out[i, j, k, :] = prop_coefficients(J, Mt, dbeta=db)
    return {"prop.tables3": out, "prop.pitch3": np.array([lo, hi], dtype=np.float64)}


# ----------------------------------------------------------------------------- geoid + gains
def geoid_table():
    path = os.path.join(REF, "lib/FlightPhysics/src/data/ww15mgh_le.bin")
    raw = open(path, "rb").read()
    assert hashlib.sha256(raw).hexdigest() == "9d190e021672769b508547021bcaebcc7d13558d66d215d019675a5f595f5cae"
    a = np.frombuffer(raw, dtype="<f4").reshape(1441, 721)   # [lon][lat]: file is column-major 721x1441
    return {"geoid": a.copy()}


LQR_FILES = ["te2te", "tv2te", "vh2te", "ar2ar", "φβ2ar"]
PID_FILES = ["q2e", "c2θ", "v2t", "p2φ", "χ2φ"]
ASCII = {"φβ2ar": "phibeta2ar", "c2θ": "c2theta", "p2φ": "p2phi", "χ2φ": "chi2phi"}


def gain_tables():
    d = os.path.join(REF, "lib/FlightApps/src/c172/c172x/control/data")
    t = {}
    for name in LQR_FILES:
        f = h5mini.read(os.path.join(d, name + ".h5"))
        dat = f["data"]
        key = "gain." + ASCII.get(name, name)
        nx = dat["x_trim"].shape[-1]
        # on-disk [h][EAS][x][u] (= Julia K[u,x,EAS,h]); pack per grid point as
        # [K_fbk (2 x nx row-major), K_fwd (2x2), K_int (2x2), x_trim nx, u_trim 2, z_trim 2] -> [EAS][h][P]
        P = 2 * nx + 4 + 4 + nx + 2 + 2
        pack = np.zeros((7, 4, P))
        for ih in range(4):
            for ie in range(7):
                parts = [dat["K_fbk"][ih, ie].T.ravel(), dat["K_fwd"][ih, ie].T.ravel(), dat["K_int"][ih, ie].T.ravel(),
                         dat["x_trim"][ih, ie], dat["u_trim"][ih, ie], dat["z_trim"][ih, ie]]
                pack[ie, ih] = np.concatenate(parts)
        t[key] = pack
        assert np.allclose(f["bounds"], [[25, 55], [50, 3050]])
    for name in PID_FILES:
        f = h5mini.read(os.path.join(d, name + ".h5"))
        dat = f["data"]
        pack = np.zeros((7, 4, 4))
        for k, fld in enumerate(["k_p", "k_i", "k_d", "τ_f"]):
            pack[:, :, k] = dat[fld].T      # on-disk [h][EAS]
        t["gain." + ASCII.get(name, name)] = pack
        assert np.allclose(f["bounds"], [[25, 55], [50, 3050]])
    return t


def write_blob(tables, path):
    names = list(tables.keys())
    hdr = 16 + 64 * len(names)
    off = (hdr + 15) // 16 * 16
    recs, payload = [], []
    for k in names:
        a = tables[k]
        if a.dtype == np.float32:
            dt, a = 1, np.ascontiguousarray(a, "<f4")
        else:
            dt, a = 0, np.ascontiguousarray(a, "<f8")
        dims = list(a.shape) + [1] * (4 - a.ndim)
        b = a.tobytes()
        nm = k.encode("ascii")
        assert len(nm) < 24
        recs.append(struct.pack("<24sII4IQQ", nm, dt, a.ndim, *dims, off, len(b)))
        pad = (-len(b)) % 16
        payload.append(b + b"\0" * pad)
        off += len(b) + pad
    with open(path, "wb") as f:
        f.write(b"FBTABLE1" + struct.pack("<II", len(names), 0))
        f.write(b"".join(recs))
        f.write(b"\0" * ((hdr + 15) // 16 * 16 - hdr))
        f.write(b"".join(payload))


def main():
    import argparse
    ap = argparse.ArgumentParser(description="write flight.jl_b200/data/c172_tables.bin")
    ap.add_argument("--pitch", nargs=3, metavar=("LOWER", "UPPER", "LENGTH"), default=("-0.1", "0.2", "11"),
                    help="blade-pitch offset range of the VariablePitch propeller whose Lookup goes into prop.tables3 (rad, rad, knots); "
                         "default: the reference's test case VariablePitch(-0.1, 0.2), length 11. fb_config.prop_pitch_* must match it.")
    ap.add_argument("--out", default=OUT)
    args = ap.parse_args()
    tables = {}
    tables.update(aero_tables())
    tables.update(engine_tables())
    tables.update(prop_tables())
    tables.update(prop_tables3(float(args.pitch[0]), float(args.pitch[1]), int(args.pitch[2])))
    tables.update(gain_tables())
    tables.update(geoid_table())
    write_blob(tables, args.out)
    print("wrote", args.out, os.path.getsize(args.out), "bytes;", len(tables), "sections")


if __name__ == "__main__":
    main()
