"""CPU tests: Segment / SegmentGuidanceData / GuidanceLaws of Cessna172Xv2 in the oracle, transcribed from the reference's
lib/FlightApps/test/c172/test_c172x2.jl (line numbers in the comments), plus the host-side Segment of the Python mirror
against the oracle's restatement of the same constructor (c172x_gdc.jl:38-98)."""
import numpy as np
import pytest

DT = 0.01
LON = dict(direct=0, sas=1, thr_EAS=4, EAS_alt=8)
LAT = dict(direct=0, sas=1, chi_beta=4)
E_THR = 1000.0   # SegmentGuidance.e_thr, c172x_gdc.jl:206


def test_segment_geometry_known_answers(orc):                                       # test_c172x2.jl:25-56
    p0 = (0.0, 0.0, 0.0)
    assert orc.segment_valid(p0, (1e-3, 0.0, 0.0))                                   # Segment() default :29
    assert not orc.segment_valid(p0, p0)                                             # :31
    assert not orc.segment_valid(p0, (0.0, 0.0, 100.0))                              # :32
    chi, dchi, s = np.pi / 3, np.pi / 4, 1e3                                         # :34-36
    p2 = orc.segment_end(p0, 1e4, chi, gamma=np.deg2rad(5))                          # :37
    p = orc.segment_end(p0, s, chi + dchi, gamma=0.0)                                # :38
    d = orc.segment_data(p0, p2, p)
    assert abs(d["s_1b"] - s * np.cos(dchi)) < 1e-2                                  # :41
    assert abs(d["e_sb"] - s * np.sin(dchi)) < 1e-2                                  # :42
    assert abs(d["h_s"] - d["s_1b"] * np.tan(np.deg2rad(5))) < 1e-2                  # :43
    assert abs(d["chi_12"] - chi) < 1e-3 and abs(d["gamma_12"] - np.deg2rad(5)) < 1e-3
    assert abs(d["s_2b"] - (d["s_1b"] - d["s_12"])) < 1e-9 and abs(d["v_sb"] + d["h_s"]) < 1e-9


def test_host_segment_matches_oracle(orc, fb):
    rng = np.random.default_rng(20261018)
    n = 200
    lat, lon, h = np.deg2rad(rng.uniform(-80, 80, n)), rng.uniform(-3.1, 3.1, n), rng.uniform(0, 3000, n)
    s, chi, g = rng.uniform(10, 5e4, n), rng.uniform(-3.1, 3.1, n), np.deg2rad(rng.uniform(-5, 5, n))
    seg = fb.Segment.from_origin((lat, lon, h), s, chi, gamma=g)
    for i in range(n):
        o = orc.segment_end((lat[i], lon[i], h[i]), s[i], chi[i], gamma=g[i])
        assert abs(o[0] - seg.p2[0][i]) < 1e-12 and abs(o[1] - seg.p2[1][i]) < 1e-12 and abs(o[2] - seg.p2[2][i]) < 1e-9
    inv = -seg                                                                       # :45-47
    assert np.array_equal(inv.p1[0], seg.p2[0]) and np.array_equal(inv.p2[2], seg.p1[2])
    with pytest.raises(ValueError):
        fb.Segment((0.0, 0.0, 0.0), (0.0, 0.0, 100.0))
    with pytest.raises(TypeError):
        fb.Segment.from_origin((0.0, 0.0, 0.0), 1e3, 0.0)


class X2:
    def __init__(self, orc, fb):
        self.fb, self.orc = fb, orc
        self.o = orc.Sim(variant=3, kin=0, dt=DT, Dt=DT)

    def u(self, name, v):
        self.o.set_input(self.fb.INPUT_NAMES[name], float(v))

    def y(self, name):
        return self.o.y(self.fb.OUTPUT_NAMES[name])

    def target(self, p1, p2):
        for k, v in zip(("lat1", "lon1", "h1", "lat2", "lon2", "h2"), tuple(p1) + tuple(p2)):
            self.u("gdc_" + k, v)

    def step(self):
        assert self.o.step(1) == 0


def test_segment_guidance_sequence(orc, fb):                                        # test_c172x2.jl:60-172
    import bench
    s = X2(orc, fb)
    # ---- on the ground the guidance mode request is ignored (:77-91)
    s.o.init(h_e=1.81 + float(bench.geoid_height(0.0, 0.0)))
    s.u("gdc_mode_req", 1); s.u("gdc_hor_req", 1); s.u("gdc_vrt_req", 1)
    s.step()
    assert s.y("wow_left") == 1 or s.y("wow_nose") == 1
    assert s.y("gdc_mode") == 0                                                      # :91
    # ---- in the air (:95-)
    ok, _, _ = s.o.init_trim(psi_nb=0.0)
    assert ok
    s.u("gdc_mode_req", 1); s.u("gdc_hor_req", 1); s.u("gdc_vrt_req", 1)
    ob = (s.y("lat"), s.y("lon"), s.y("h_e"))
    chi_ac, dh = s.y("chi_gnd"), 100.0
    aux = orc.segment_end(ob, E_THR / 2, chi_ac + np.pi / 2, dh=dh)                  # :106
    s.target(aux, orc.segment_end(aux, 1e4, 0.0, gamma=np.deg2rad(5)))               # :107
    s.step()
    assert s.y("gdc_mode") == 1                                                      # :113
    assert s.y("gdc_hor") == 1 and s.y("lat_mode") == LAT["chi_beta"]                # :116-117
    assert s.o.get_input(fb.U["FB_U_LAT_CHI_REF"]) == s.y("gdc_chi_ref")             # :118
    assert s.y("gdc_vrt") == 1                                                       # :122
    assert s.y("lon_mode") in (LON["EAS_alt"], LON["thr_EAS"])                       # :123-124
    assert s.o.get_input(fb.U["FB_U_LON_H_REF"]) == s.y("gdc_h_ref")                 # :125
    assert abs(s.y("gdc_h_ref") - (ob[2] + dh)) < 1.0                                # :126
    assert s.y("gdc_dchi") > 0                                                       # :129
    # segment to the left (:139-145)
    aux = orc.segment_end(ob, E_THR / 2, chi_ac - np.pi / 2, gamma=0.0)
    s.target(aux, orc.segment_end(aux, 1e4, 0.0, gamma=np.deg2rad(5)))
    s.step()
    assert s.y("gdc_dchi") < 0                                                       # :145
    # outside the vertical guidance threshold (:147-154)
    aux = orc.segment_end(ob, 2 * E_THR, chi_ac + np.pi / 2, gamma=0.0)
    s.target(aux, orc.segment_end(aux, 1e4, 0.0, gamma=np.deg2rad(5)))
    s.step()
    assert s.y("gdc_vrt") == 0                                                       # :154
    # vertical guidance off: lon mode unchanged, and settable again (:158-169)
    lon_prev = s.y("lon_mode")
    s.u("gdc_vrt_req", 0)
    s.step()
    assert s.y("gdc_vrt") == 0 and s.y("lon_mode") == lon_prev                       # :161-162
    s.u("lon_mode_req", LON["sas"])
    s.step()
    assert s.y("lon_mode") == LON["sas"]                                             # :168
    # horizontal guidance off: lat mode unchanged, and settable again (:172-)
    lat_prev = s.y("lat_mode")
    s.u("gdc_hor_req", 0)
    s.step()
    assert s.y("gdc_hor") == 0 and s.y("lat_mode") == lat_prev
    s.u("lat_mode_req", LAT["sas"])
    s.step()
    assert s.y("lat_mode") == LAT["sas"]


def test_segment_capture_closed_loop(orc, fb):
    """Not in the reference's tests: fly the guidance for 120 s and require that the aircraft captures the segment
    (cross-track and vertical errors decay, course aligned with the segment) — a behavioural check of sign conventions."""
    s = X2(orc, fb)
    ok, _, _ = s.o.init_trim(psi_nb=0.3, EAS=45.0, h_e=1200.0)
    assert ok
    ob = (s.y("lat"), s.y("lon"), s.y("h_e"))
    p1 = orc.segment_end(ob, 400.0, 0.3 + np.pi / 2, dh=30.0)
    s.target(p1, orc.segment_end(p1, 3e4, 0.5, gamma=0.0))
    s.u("gdc_mode_req", 1); s.u("gdc_hor_req", 1); s.u("gdc_vrt_req", 1)
    e0 = None
    for k in range(120):
        assert s.o.step(100) == 0
        if k == 0:
            e0 = abs(s.y("gdc_e_sb"))
    assert e0 > 300.0
    assert abs(s.y("gdc_e_sb")) < 5.0 and abs(s.y("gdc_v_sb")) < 3.0
    assert abs(s.y("chi_gnd") - 0.5) < 0.02
    assert s.y("lat_mode") == LAT["chi_beta"] and s.y("lon_mode") == LON["EAS_alt"]
