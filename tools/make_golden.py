#!/usr/bin/env python
"""Writes tests/golden/trim_points.json: the reference's OWN trim solutions, stored in its gain-schedule files
(lib/FlightApps/src/c172/c172x/control/data/{te2te,ar2ar}.h5, datasets data/x_trim, data/u_trim), at the 28
design points (7 EAS x 4 altitudes). They were produced by design/c172/c172x_design.jl:106-118 with
C172X.Vehicle{NED}, Geographic(LatLon(), HEllip(h)), flaps = flaps_schedule(EAS), everything else default.
Run here (needs /root/reference); the JSON is committed so nothing at test time reads the reference."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import h5mini
REF = os.environ.get("FLIGHT_REF", "/root/reference")
D = os.path.join(REF, "lib/FlightApps/src/c172/c172x/control/data")
lon = h5mini.read(os.path.join(D, "te2te.h5"))["data"]
lat = h5mini.read(os.path.join(D, "ar2ar.h5"))["data"]
EAS = [25, 30, 35, 40, 45, 50, 55]; H = [50, 1050, 2050, 3050]
pts = []
for ih, h in enumerate(H):
    for ie, e in enumerate(EAS):
        xl, xa = lon["x_trim"][ih, ie], lat["x_trim"][ih, ie]
        # lon: q, theta, EAS, alpha, alpha_filt, n_eng, thr_p, ele_p (c172x_ctl.jl:120-129); lat: p, r, phi, EAS, beta, beta_filt, ail_p, rud_p (:756-765)
        pts.append(dict(EAS=e, h_e=h, flaps=(1.0 if e <= 30 else 0.0),
                        x_lon=[float(v) for v in xl], x_lat=[float(v) for v in xa],
                        u_lon=[float(v) for v in lon["u_trim"][ih, ie]], u_lat=[float(v) for v in lat["u_trim"][ih, ie]],
                        trim_state=[float(xl[3]), float(xa[2]), float(xl[5]), float(xl[6]), float(xa[6]), float(xl[7]), float(xa[7])]))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "trim_points.json")
json.dump(dict(source="te2te.h5 / ar2ar.h5 data/x_trim (reference gain-schedule files)", trim_state_order=["alpha_a", "phi_nb", "n_eng", "throttle", "aileron", "elevator", "rudder"], points=pts), open(out, "w"), indent=1)
print("wrote", out, len(pts), "points")
