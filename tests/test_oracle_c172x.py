"""CPU tests: closed-loop behaviour of the oracle's Cessna172Xv1 (fly-by-wire + gain-scheduled control laws),
transcribed from the reference's lib/FlightApps/test/c172/test_c172x1.jl (dt = Δt = 0.01, :42). Line numbers in the
comments point at the reference assertions being replayed, including the Bool / enum expectations."""
import numpy as np
import pytest

DT = 0.01
LON = dict(direct=0, sas=1, thr_q=2, thr_theta=3, thr_EAS=4, EAS_q=5, EAS_theta=6, EAS_clm=7, EAS_alt=8)
LAT = dict(direct=0, sas=1, p_beta=2, phi_beta=3, chi_beta=4)


class X1:
    """Thin driver: Cessna172Xv1 on SimpleWorld, init!(sim, C172.TrimParameters()) (the EAS 50 / h 1050 design point)."""

    def __init__(self, orc, fb):
        self.fb = fb
        self.o = orc.Sim(variant=2, kin=0, dt=DT, Dt=DT)
        ok, ts, cost = self.o.init_trim()
        assert ok
        x = self.o.x
        self.L_dyn = self.o.nx - 6
        self.trim = dict(w=np.array([self.y("p"), self.y("q"), self.y("r")]), v=x[self.L_dyn + 3:].copy(), h_e=self.y("h_e"))

    def u(self, name, v):
        self.o.set_input(self.fb.INPUT_NAMES[name], v)

    def y(self, name):
        return self.o.y(self.fb.OUTPUT_NAMES[name])

    def step(self, t):
        st = self.o.step(int(round(t / DT)))
        assert st == 0

    def q_close(self, atol=1e-5):
        return abs(self.y("q") - self.trim["w"][1]) < atol

    def vx_close(self, atol=1e-2):
        return abs(self.o.x[self.L_dyn + 3] - self.trim["v"][0]) < atol


def test_on_ground_modes_are_overridden(orc, fb):                                   # test_c172x1.jl:55-86
    import bench
    o = orc.Sim(variant=2, kin=0, dt=DT, Dt=DT)
    o.init(h_e=1.9 + float(bench.geoid_height(0.0, 0.0)))
    for k, v in (("lon_mode_req", LON["EAS_clm"]), ("lat_mode_req", LAT["p_beta"]), ("lon_throttle_axis", 0.1),
                 ("lon_elevator_axis", 0.3), ("lat_aileron_axis", 0.2), ("lat_rudder_axis", 0.4)):
        o.set_input(fb.INPUT_NAMES[k], v)
    assert o.step(1) == 0
    assert o.y(fb.Y["FB_Y_WOW_LEFT"]) == 1 or o.y(fb.Y["FB_Y_WOW_NOSE"]) == 1        # is_on_gnd :70
    assert o.y(fb.Y["FB_Y_LON_MODE"]) == LON["direct"] and o.y(fb.Y["FB_Y_LAT_MODE"]) == LAT["direct"]   # :72-73
    assert o.get_input(fb.U["FB_U_ACT_THROTTLE"]) == 0.1 and o.get_input(fb.U["FB_U_ACT_ELEVATOR"]) == 0.3   # :77-80
    assert o.get_input(fb.U["FB_U_ACT_RUDDER"]) == 0.4 and o.get_input(fb.U["FB_U_ACT_AILERON"]) == 0.2


def test_direct_mode_preserves_trim(orc, fb):                                      # :98-113
    s = X1(orc, fb)
    s.step(DT)
    assert s.y("lon_mode") == LON["direct"] and s.y("lat_mode") == LAT["direct"]     # :105-106
    s.step(10)
    w = np.array([s.y("p"), s.y("q"), s.y("r")])
    assert np.all(np.abs(w - s.trim["w"]) < 1e-5)                                    # :110
    assert np.all(np.abs(s.o.x[s.L_dyn + 3:] - s.trim["v"]) < 1e-2)                  # :111


def test_gain_lookup_equals_tables(orc):                                            # :130-132,156-158,181-183 ... (atol 1e-6)
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "trim_points.json")))
    pt = [p for p in gold["points"] if p["EAS"] == 50 and p["h_e"] == 1050][0]
    g = orc.gain_lookup(0, 50.0, 1050.0)          # te2te at the design point: packed K_fbk(16) K_fwd(4) K_int(4) x_trim(8) u_trim(2) z_trim(2)
    assert np.allclose(g[24:32], pt["x_lon"], atol=1e-12) and np.allclose(g[32:34], pt["u_lon"], atol=1e-12)
    a, b, m = orc.gain_lookup(3, 45.0, 1050.0), orc.gain_lookup(3, 50.0, 2050.0), orc.gain_lookup(3, 47.5, 1550.0)
    c, d = orc.gain_lookup(3, 50.0, 1050.0), orc.gain_lookup(3, 45.0, 2050.0)
    assert np.allclose(m, 0.25 * (a + b + c + d), atol=1e-12)                        # bilinear on the scaled grid
    assert np.allclose(orc.gain_lookup(5, 10.0, -500.0), orc.gain_lookup(5, 25.0, 50.0))   # Flat extrapolation (control.jl:965)


def test_lon_sas_then_lat_sas_hold_trim(orc, fb):                                   # :119-168
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["sas"]); s.step(DT)
    assert s.y("lon_mode") == LON["sas"]                                             # :127
    s.step(30)
    assert s.q_close() and s.vx_close()                                              # :137-138
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["sas"]); s.u("lat_mode_req", LAT["sas"]); s.step(DT)
    assert s.y("lat_mode") == LAT["sas"]                                             # :153
    s.step(10)
    assert abs(s.y("p") - s.trim["w"][0]) < 1e-5 and s.vx_close()                    # :162-163


def test_phi_beta_tracking(orc, fb):                                               # :172-202
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["sas"]); s.u("lat_mode_req", LAT["phi_beta"]); s.step(DT)
    assert s.y("lat_mode") == LAT["phi_beta"]                                        # :178
    s.step(10)
    assert s.vx_close()                                                              # :188
    s.u("lat_phi_ref", np.pi / 12); s.u("lat_beta_ref", np.deg2rad(3)); s.step(10)
    assert abs(s.y("phi") - np.pi / 12) < 1e-3 and abs(s.y("beta") - np.deg2rad(3)) < 1e-3   # :193-194


def test_p_beta_tracking(orc, fb):                                                 # :206-243
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["sas"]); s.u("lat_mode_req", LAT["sas"]); s.step(1)
    s.u("lat_mode_req", LAT["p_beta"]); s.step(DT)
    assert s.y("lat_mode") == LAT["p_beta"]                                          # :216
    s.step(1); assert s.q_close() and s.vx_close()                                   # :225-226
    s.step(10); assert s.q_close() and s.vx_close()                                  # :230-231
    s.u("lat_p_ref", 0.02); s.u("lat_beta_ref", np.deg2rad(3)); s.step(10)
    assert abs(s.y("p") - 0.02) < 1e-3 and abs(s.y("beta") - np.deg2rad(3)) < 1e-3    # :236-237


def test_chi_beta_tracking_with_crosswind(orc, fb):                                # :247-287
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["sas"]); s.u("lat_mode_req", LAT["sas"]); s.step(1)
    s.u("lat_mode_req", LAT["chi_beta"]); s.step(DT)
    assert s.y("lat_mode") == LAT["chi_beta"]                                        # :257
    s.step(1); assert s.q_close() and s.vx_close()                                   # :267-268
    s.u("lat_chi_ref", np.pi / 2); s.step(29)
    assert abs(s.y("chi_gnd") - np.pi / 2) < 1e-2                                    # :274
    s.u("wind_n", 10.0); s.step(10)
    assert abs(s.y("chi_gnd") - np.pi / 2) < 1e-2                                    # :280


def test_thr_q_and_thr_theta(orc, fb):                                             # :295-353
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["thr_q"]); s.u("lat_mode_req", LAT["phi_beta"]); s.step(DT)
    assert s.y("lon_mode") == LON["thr_q"]                                           # :301
    s.step(1); assert s.q_close() and s.vx_close()                                   # :311-312
    s.u("lat_phi_ref", np.pi / 12); s.u("lon_q_ref", 0.01); s.step(10)
    assert abs(s.y("q") - 0.01) < 1e-3                                               # :320
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["thr_theta"]); s.u("lat_mode_req", LAT["phi_beta"]); s.step(DT)
    assert s.y("lon_mode") == LON["thr_theta"]                                       # :338
    s.step(1); assert s.q_close() and s.vx_close()
    s.u("lat_phi_ref", np.pi / 6); s.u("lon_theta_ref", np.deg2rad(5)); s.step(10)
    assert abs(s.y("theta") - np.deg2rad(5)) < 1e-4                                  # :350


def test_thr_EAS_and_EAS_q(orc, fb):                                               # :359-428
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["thr_EAS"]); s.u("lat_mode_req", LAT["phi_beta"]); s.step(DT)
    assert s.y("lon_mode") == LON["thr_EAS"]                                         # :366
    s.step(1); assert s.q_close() and s.vx_close()
    s.u("lat_phi_ref", np.pi / 6); s.u("lon_eas_ref", 45.0); s.step(30)
    assert abs(s.y("eas") - 45.0) < 1e-1                                             # :384
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["EAS_q"]); s.u("lat_mode_req", LAT["phi_beta"]); s.step(DT)
    assert s.y("lon_mode") == LON["EAS_q"]                                           # :399
    s.step(1); assert s.q_close() and s.vx_close()
    for qr in (-0.005, 0.005, 0.0):
        s.u("lon_q_ref", qr); s.step(20)
        assert abs(s.y("q") - qr) < 1e-3 and abs(s.y("eas") - 50.0) < 1              # :416-427


def test_EAS_theta_and_EAS_clm(orc, fb):                                           # :436-493
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["EAS_theta"]); s.u("lat_mode_req", LAT["phi_beta"]); s.step(DT)
    assert s.y("lon_mode") == LON["EAS_theta"]                                       # :443
    s.step(0.1); assert s.q_close() and s.vx_close()
    s.u("lat_phi_ref", np.pi / 6); s.u("lon_theta_ref", np.deg2rad(3)); s.step(10)
    s.u("lon_theta_ref", -np.deg2rad(3)); s.step(60)
    assert abs(s.y("theta") + np.deg2rad(3)) < 1e-3 and abs(s.y("eas") - 50.0) < 1e-1   # :458-459
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["EAS_clm"]); s.u("lat_mode_req", LAT["phi_beta"]); s.step(DT)
    assert s.y("lon_mode") == LON["EAS_clm"]                                         # :475
    s.step(1); assert s.q_close() and s.vx_close()                                   # :485-486
    s.u("lat_phi_ref", np.pi / 6); s.u("lon_eas_ref", 45.0); s.u("lon_clm_ref", 2.0); s.step(30)
    assert abs(s.y("v_d") + 2.0) < 1e-1 and abs(s.y("eas") - 45.0) < 2e-1            # :492-493


def test_altitude_acquire_and_hold_transitions(orc, fb):                           # :501-556
    s = X1(orc, fb)
    s.u("lon_mode_req", LON["EAS_alt"]); s.u("lat_mode_req", LAT["phi_beta"]); s.step(DT)
    assert s.y("alt_state") == 1 and s.y("lon_mode") == LON["EAS_alt"]               # :518-519
    s.step(1); assert s.q_close() and s.vx_close()                                   # :524-525
    s.u("lat_phi_ref", np.pi / 12)
    s.u("lon_h_ref", s.trim["h_e"] + 100); s.step(1)
    assert s.y("alt_state") == 0 and s.y("lon_mode") == LON["thr_EAS"]               # :532-533
    s.step(60)
    assert s.y("alt_state") == 1 and abs(s.y("h_e") - (s.trim["h_e"] + 100)) < 1e-1   # :536-537
    h_ref = s.y("h_e") - 5.0                                                         # within h_thr / 2 (:540)
    s.u("lon_h_ref", h_ref); s.step(1)
    assert s.y("alt_state") == 1                                                     # :542
    s.step(30); assert abs(s.y("h_e") - h_ref) < 1e-1                                # :544
    s.u("lon_h_ref", s.trim["h_e"] - 100); s.step(1)
    assert s.y("alt_state") == 0                                                     # :548
    s.step(80)
    assert s.y("alt_state") == 1 and abs(s.y("h_e") - (s.trim["h_e"] - 100)) < 1e-1   # :550-551
    assert s.y("lon_mode") == LON["EAS_alt"]                                         # :554
