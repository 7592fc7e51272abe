"""TEST INFRASTRUCTURE ONLY — ctypes binding of the CPU oracle (oracle/liborc*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (flight.jl_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_GEOID = os.path.join(_HERE, "..", "flight.jl_b200", "data", "c172_tables.bin")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


FAST_FLAGS = "-O3 -march=native"


def _fast_name():
    """liborc_fast_<hash of this host's CPU flags>.so: the -march=native timing build is only valid where it was built."""
    import hashlib
    flags = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                flags = line
                break
    except OSError:
        pass
    return "liborc_fast_%s.so" % hashlib.sha1(flags.encode()).hexdigest()[:10]


def build_fast():
    name = _fast_name()
    subprocess.check_call(["make", "-s", "-C", _HERE, name])
    return name


def _load(name):
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    vp, dp, i32, i64, dbl = C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_longlong, C.c_double
    sig = {
        "orc_load_geoid": (i32, [C.c_char_p]),
        "orc_world_create": (vp, []),
        "orc_world_destroy": (None, [vp]),
        "orc_world_prop_table": (None, [vp, dp]),
        "orc_world_engine_tables": (None, [vp, dp, dp, dp]),
        "orc_sim_create": (vp, [vp, i32, i32, dbl, dbl]),
        "orc_sim_set_components": (None, [vp, dbl, dbl, i32, dbl, dbl]),
        "orc_sim_prop3": (i32, [vp, dp]),
        "orc_sim_destroy": (None, [vp]),
        "orc_sim_nx": (i32, [vp]),
        "orc_sim_status": (i32, [vp]),
        "orc_sim_want_y_extra": (None, [vp, i32]),
        "orc_sim_init_trim": (i32, [vp, dp, dp, dp]),
        "orc_sim_trim_cost_at": (dbl, [vp, dp, dp]),
        "orc_sim_init": (i32, [vp, dp]),
        "orc_sim_init_state": (i32, [vp, dp, i32, i32]),
        "orc_sim_get_state": (None, [vp, dp]),
        "orc_sim_get_xdot": (None, [vp, dp]),
        "orc_sim_get_discrete": (None, [vp, C.POINTER(i32), C.POINTER(i32)]),
        "orc_sim_set_input": (i32, [vp, i32, dbl]),
        "orc_sim_get_input": (dbl, [vp, i32]),
        "orc_sim_get_output": (dbl, [vp, i32]),
        "orc_sim_add_event": (None, [vp, i64, i32, dbl]),
        "orc_sim_clear_events": (None, [vp]),
        "orc_sim_rhs": (i32, [vp]),
        "orc_sim_step": (i32, [vp, i64]),
        "orc_sim_step_noisy": (i32, [vp, i64, C.c_ulonglong]),
        "orc_sim_steps_taken": (i64, [vp]),
        "orc_sim_fstep": (i32, [vp]), "orc_sim_fperiodic": (i32, [vp]),
        "orc_sim_linearize_dims": (None, [vp, C.POINTER(i32)]),
        "orc_sim_linearize": (i32, [vp, dp, dp, dp, dp, dp, dp, dp, dp]),
        "orc_ops_reset": (None, []),
        "orc_ops_get": (None, [C.POINTER(C.c_ulonglong)]),
        "orc_is_counting": (i32, []),
        "orc_batch_step": (i32, [C.POINTER(vp), i64, i64, i32]),
        "orc_kin_sim": (i32, [vp, i32, dp, dbl, i64, dp]),
        "orc_dyn_test": (i32, [vp, dp, dbl, dp, dp, dp, dp, dp]),
        "orc_ldg_test": (i32, [vp, dp, dbl, dbl, i32, dbl, dbl, dbl, i32, dp]),
        "orc_get_mu": (dbl, [i32, i32, dbl]),
        "orc_engine_lookup": (dbl, [dbl, dbl, i32, dbl, dbl, dbl]),
        "orc_h2delta": (dbl, [dbl]),
        "orc_gain_lookup": (i32, [i32, dbl, dbl, dp]),
        "orc_geo_roundtrip": (None, [dbl, dbl, dbl, dp]),
        "orc_lss_step": (None, [i32, i32, dp, dp, dp, dp, dp, dp, dp, dbl, i64]),
        "orc_attitude_test": (None, [dp]),
        "orc_quat_test": (None, [dp]),
        "orc_theta_constraint_test": (dbl, []),
        "orc_geodesy_test": (None, [dp]),
        "orc_pi_run": (None, [dp, dp, dbl, i32, dbl, i64, dp]),
        "orc_pid_run": (None, [dp, dp, dbl, i32, dbl, i64, dp]),
        "orc_integ_run": (None, [dbl, dbl, dp, dbl, i32, dbl, i64, dp]),
        "orc_lss_rhs": (None, [i32, i32, i32, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]),
        "orc_segment_end": (None, [dbl, dbl, dbl, dbl, dbl, dbl, dp]),
        "orc_segment_data": (None, [dp, dp, dp, dp]),
        "orc_segment_valid": (i32, [dp, dp]),
    }
    for k, (res, args) in sig.items():
        f = getattr(lib, k)
        f.restype = res
        f.argtypes = args
    rc = lib.orc_load_geoid(os.path.abspath(_GEOID).encode())
    if rc != 0:
        raise RuntimeError(f"oracle: cannot load geoid ({rc})")
    return lib


_libs = {}


def build_fma():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liborc_fma.so"])
    return "liborc_fma.so"


def lib(counting=False, fast=False):
    # fast: False = the checker, True = the -march=native timing build of THIS host, "fma" = the portable FMA-contracted build
    # (same code on every host: the rounding-spread tests must not depend on which CPU the GPU box has)
    key = build_fma() if fast == "fma" else (build_fast() if fast else ("liborc_count.so" if counting else "liborc.so"))
    if key not in _libs:
        _libs[key] = _load(key)
    return _libs[key]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


TRIM_DEFAULT = dict(lat=0.0, lon=0.0, h_e=1050.0, psi_nb=0.0, EAS=50.0, gamma_wb_n=0.0, psi_wb_dot=0.0,
                    theta_wb_dot=0.0, beta_a=0.0, fuel_load=0.5, mixture=0.5, flaps=0.0,
                    m_pilot=75.0, m_copilot=75.0, m_lpass=0.0, m_rpass=0.0, m_baggage=50.0)
TRIM_KEYS = list(TRIM_DEFAULT.keys())

INIT_DEFAULT = dict(q0=1.0, q1=0.0, q2=0.0, q3=0.0, lat=0.0, lon=0.0, h_e=0.0, w0=0.0, w1=0.0, w2=0.0,
                    vN=0.0, vE=0.0, vD=0.0, n_eng=0.0, engine_state=0.0, throttle=0.0, mixture=0.5, elevator=0.0,
                    aileron=0.0, rudder=0.0, flaps=0.0, brake_left=0.0, brake_right=0.0, fuel_load=0.5,
                    m_pilot=75.0, m_copilot=75.0, m_lpass=0.0, m_rpass=0.0, m_baggage=50.0, stall=0.0,
                    alpha_filt=0.0, beta_filt=0.0)
INIT_KEYS = list(INIT_DEFAULT.keys())


class World:
    def __init__(self, counting=False, fast=False):
        self.lib = lib(counting, fast)
        self.h = self.lib.orc_world_create()
        assert self.h

    def prop_table(self):
        out = np.zeros((6, 21, 21))
        self.lib.orc_world_prop_table(self.h, _dp(out))
        return out

    def engine_tables(self):
        a, b, c = np.zeros((2, 9)), np.zeros((13, 3)), np.zeros((5, 3))
        self.lib.orc_world_engine_tables(self.h, _dp(a), _dp(b), _dp(c))
        return a, b, c


_world = {}


def world(counting=False, fast=False):
    key = (counting, fast)
    if key not in _world:
        _world[key] = World(counting, fast)
    return _world[key]


class Sim:
    """One reference `Simulation(SimpleWorld(Cessna172...); dt, Δt)` on the CPU oracle."""

    def __init__(self, variant=0, kin=0, dt=0.02, Dt=None, counting=False, fast=False, prop_pitch=None, act2=None):
        """fast=True: the -O3 -march=native timing build (bench.py's CPU arm); fast="fma": the portable FMA-contracted build
        (rounding-spread tests); the checker is the default build. variant 4 (Actuator2 actuation): prop_pitch = (lo, hi[, n])
        makes the propeller a Propeller(VariablePitch(lo, hi; length = n)), act2 = (omega_n, zeta) overrides the actuator
        defaults."""
        self.w = world(counting, fast)
        self.lib = self.w.lib
        self.h = self.lib.orc_sim_create(self.w.h, variant, kin, dt, dt if Dt is None else Dt)
        self.nx = self.lib.orc_sim_nx(self.h)
        self.dt = dt
        if prop_pitch is not None or act2 is not None:
            lo, hi, n = (tuple(prop_pitch) + (11,))[:3] if prop_pitch is not None else (0.0, 0.0, 0)
            wn, zt = act2 if act2 is not None else (0.0, 0.0)
            self.lib.orc_sim_set_components(self.h, float(wn), float(zt), int(n), float(lo), float(hi))

    def prop3(self):
        """[21][21][n][6] coefficient table of the variable-pitch propeller (None for FixedPitch)"""
        n = self.lib.orc_sim_prop3(self.h, None)
        if n <= 0:
            return None
        out = np.zeros((21, 21, n, 6))
        self.lib.orc_sim_prop3(self.h, _dp(out))
        return out

    def __del__(self):
        try:
            self.lib.orc_sim_destroy(self.h)
        except Exception:
            pass

    def init_trim(self, **kw):
        p = dict(TRIM_DEFAULT); p.update(kw)
        pv = np.array([p[k] for k in TRIM_KEYS], dtype=np.float64)
        ts = np.zeros(7); cost = C.c_double(0)
        rc = self.lib.orc_sim_init_trim(self.h, _dp(pv), _dp(ts), C.byref(cost))
        return rc == 0, ts, cost.value

    def trim_cost_at(self, trim_state, **kw):
        p = dict(TRIM_DEFAULT); p.update(kw)
        pv = np.array([p[k] for k in TRIM_KEYS], dtype=np.float64)
        ts = np.ascontiguousarray(trim_state, dtype=np.float64)
        return self.lib.orc_sim_trim_cost_at(self.h, _dp(pv), _dp(ts))

    def init(self, **kw):
        p = dict(INIT_DEFAULT); p.update(kw)
        pv = np.array([p[k] for k in INIT_KEYS], dtype=np.float64)
        return self.lib.orc_sim_init(self.h, _dp(pv))

    def init_state(self, x, stall=0, eng_state=2):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return self.lib.orc_sim_init_state(self.h, _dp(x), int(stall), int(eng_state))

    @property
    def x(self):
        out = np.zeros(self.nx)
        self.lib.orc_sim_get_state(self.h, _dp(out))
        return out

    @property
    def xdot(self):
        out = np.zeros(self.nx)
        self.lib.orc_sim_get_xdot(self.h, _dp(out))
        return out

    @property
    def discrete(self):
        a, b = C.c_int(0), C.c_int(0)
        self.lib.orc_sim_get_discrete(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    @property
    def status(self):
        return self.lib.orc_sim_status(self.h)

    def set_input(self, field, v):
        rc = self.lib.orc_sim_set_input(self.h, int(field), float(v))
        assert rc == 0, f"unknown input field {field}"

    def get_input(self, field):
        return self.lib.orc_sim_get_input(self.h, int(field))

    def y(self, field):
        return self.lib.orc_sim_get_output(self.h, int(field))

    def add_event(self, step, field, value):
        self.lib.orc_sim_add_event(self.h, int(step), int(field), float(value))

    def rhs(self):
        return self.lib.orc_sim_rhs(self.h)

    def step(self, n=1):
        return self.lib.orc_sim_step(self.h, int(n))

    def step_noisy(self, n, seed):
        """n steps with a one-ulp random nudge of every state component after each (rounding-sensitivity probe)."""
        return self.lib.orc_sim_step_noisy(self.h, int(n), int(seed))

    def f_step(self):
        return self.lib.orc_sim_fstep(self.h)

    def f_periodic(self):
        return self.lib.orc_sim_fperiodic(self.h)

    def linearize(self):
        """linearize(aircraft, trim) about the current state (NED kinematics): dict of A, B, C, D (row-major), x0, u0, y0, xdot0."""
        d = (C.c_int * 3)()
        self.lib.orc_sim_linearize_dims(self.h, d)
        nx, nu, ny = d[0], d[1], d[2]
        r = dict(A=np.zeros((nx, nx)), B=np.zeros((nx, nu)), C=np.zeros((ny, nx)), D=np.zeros((ny, nu)),
                 x0=np.zeros(nx), u0=np.zeros(nu), y0=np.zeros(ny), xdot0=np.zeros(nx))
        rc = self.lib.orc_sim_linearize(self.h, *[_dp(r[k]) for k in ("A", "B", "C", "D", "x0", "u0", "y0", "xdot0")])
        assert rc == 0, rc
        return r

    def want_y_extra(self, on=True):
        self.lib.orc_sim_want_y_extra(self.h, 1 if on else 0)


KIN_KEYS = ["q_nb", "n_e", "h_e", "v_eb_b", "w_eb_b", "v_eb_n", "q_en", "r_eb_e", "h_o"]


def _kin_dict(a):
    return dict(q_nb=a[0:4], n_e=a[4:7], h_e=a[7], v_eb_b=a[8:11], w_eb_b=a[11:14], v_eb_n=a[14:17], q_en=a[17:21], r_eb_e=a[21:24], h_o=a[24])


def _ic(q_nb=(1, 0, 0, 0), lat=0.0, lon=0.0, h_orth=0.0, w_wb_b=(0, 0, 0), v_eb_n=(0, 0, 0)):
    return np.array(list(q_nb) + [lat, lon, h_orth] + list(w_wb_b) + list(v_eb_n), dtype=np.float64)


def kin_sim(kin, dt, nsteps, **ic):
    """Kinematics model alone (reference test_kinematics.jl): returns the KinData fields after nsteps."""
    w = world()
    out = np.zeros(25)
    rc = w.lib.orc_kin_sim(w.h, int(kin), _dp(_ic(**ic)), float(dt), int(nsteps), _dp(out))
    assert rc == 0
    return _kin_dict(out)


def dyn_test(mass, Jdiag, r_bc, F, tau, **ic):
    w = world()
    out = np.zeros(38)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (Jdiag, r_bc, F, tau)]
    rc = w.lib.orc_dyn_test(w.h, _dp(_ic(**ic)), float(mass), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(out))
    assert rc == 0
    d = dict(wdot=out[0:3], vdot=out[3:6], a_eb_b=out[6:9], a_ib_b=out[9:12], gravity=out[12])
    d.update(_kin_dict(out[13:]))
    return d


def ldg_test(steer_input=0.0, brake=0.0, engaged=True, psi_max=np.pi / 6, frc_x=(0.0, 0.0), surface=0, **ic):
    w = world()
    out = np.zeros(22)
    rc = w.lib.orc_ldg_test(w.h, _dp(_ic(**ic)), float(steer_input), float(brake), int(engaged), float(psi_max),
                            float(frc_x[0]), float(frc_x[1]), int(surface), _dp(out))
    assert rc == 0, rc
    return dict(wow=bool(out[0]), xi=out[1], xi_dot=out[2], F_dmp=out[3], v_xy=out[4:6], mu_roll=out[6], mu_skid=out[7], mu_max=out[8:10],
                mu_eff=out[10:12], f_c=out[12:15], F_b=out[15:18], xdot=out[18:20], sat_out=out[20:22])


def gain_lookup(which, EAS, h):
    out = np.zeros(64)
    n = lib().orc_gain_lookup(int(which), float(EAS), float(h), _dp(out))
    return out[:n]


def quat_test():
    """Hamilton-product identities of test_quaternions.jl (layout: orc_capi.cpp:orc_quat_test)"""
    out = np.zeros(7)
    lib().orc_quat_test(_dp(out))
    return out


def theta_constraint_test():
    """test_aircraft_base.jl:16-40: flight-path angle recovered from the theta the trim constraint returns, minus the one asked for"""
    return float(lib().orc_theta_constraint_test())


def geodesy_test():
    """Known answers of test_geodesy.jl (layout: orc_capi.cpp:orc_geodesy_test)"""
    out = np.zeros(8)
    lib().orc_geodesy_test(_dp(out))
    return out


def pi_run(k_p, k_i, k_l, lo, hi, x, input, sat_ext, dt, n):
    """One lane of PIContinuous (control.jl:52-81) under RK4: returns (x_new, PIContinuousY of the lane as a dict)"""
    par = np.array([k_p, k_i, k_l, lo, hi], dtype=np.float64); xs = np.array([x], dtype=np.float64); out = np.zeros(6)
    lib().orc_pi_run(_dp(par), _dp(xs), float(input), int(sat_ext), float(dt), int(n), _dp(out))
    return float(xs[0]), dict(output=out[0], out_free=out[1], sat_out=int(out[2]), int_halted=bool(out[3]), y_i=out[4], u_i=out[5])


class DiscretePID:
    """PIDDiscrete (control.jl:370-471) in isolation: parameters as attributes, `run(input, sat_ext, Dt, n)` = n periodic updates,
    returns PIDDiscreteY as a dict."""

    def __init__(self, k_p=1.0, k_i=0.0, k_d=0.0, tau_f=0.01, lo=-np.inf, hi=np.inf):
        self.k_p, self.k_i, self.k_d, self.tau_f, self.lo, self.hi = k_p, k_i, k_d, tau_f, lo, hi
        self.st = np.zeros(3)

    def init(self):
        self.st[:] = 0.0

    def run(self, input, sat_ext, Dt, n):
        par = np.array([self.k_p, self.k_i, self.k_d, self.tau_f, self.lo, self.hi], dtype=np.float64)
        out = np.zeros(8)
        lib().orc_pid_run(_dp(par), _dp(self.st), float(input), int(sat_ext), float(Dt), int(n), _dp(out))
        return dict(y_p=out[0], y_i=out[1], y_d=out[2], out_free=out[3], output=out[4], sat_out=int(out[5]), int_halted=bool(out[6]), u_i=out[7])


class DiscreteIntegrator:
    """IntegratorDiscrete (control.jl:123-185) in isolation."""

    def __init__(self, lo=-np.inf, hi=np.inf):
        self.lo, self.hi = lo, hi
        self.st = np.zeros(2)

    def init(self):
        self.st[:] = 0.0

    def run(self, input, sat_ext, Dt, n):
        out = np.zeros(4)
        lib().orc_integ_run(float(self.lo), float(self.hi), _dp(self.st), float(input), int(sat_ext), float(Dt), int(n), _dp(out))
        return dict(x1=out[0], output=out[1], sat_out=int(out[2]), halted=bool(out[3]), x0=self.st[0])


def attitude_test():
    """Residuals / values of the rotation-algebra known answers of test_attitude.jl (layout: orc_capi.cpp:orc_attitude_test)"""
    out = np.zeros(36)
    lib().orc_attitude_test(_dp(out))
    return out


def lss_rhs(A, B, Cm, D, xdot0, x0, u0, y0, x, u):
    """LinearizedSS f_ode! (linearization.jl:164-190): returns (xdot, y)"""
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    A, B, Cm, D, xdot0, x0, u0, y0, x, u = map(f, (A, B, Cm, D, xdot0, x0, u0, y0, x, u))
    xd, y = np.zeros(A.shape[0]), np.zeros(Cm.shape[0])
    lib().orc_lss_rhs(A.shape[0], B.shape[1], Cm.shape[0], _dp(A), _dp(B), _dp(Cm), _dp(D), _dp(xdot0), _dp(x0), _dp(u0), _dp(y0), _dp(x), _dp(u), _dp(xd), _dp(y))
    return xd, y


def lss_step(A, B, x0, u0, xdot0, x, u, dt, n_steps):
    """LinearizedSS (linearization.jl:157-192) under the path's fixed-step RK4; A [nx, nx], B [nx, nu] row-major. Returns x."""
    A, B, x0, u0, xdot0, u = (np.ascontiguousarray(v, dtype=np.float64) for v in (A, B, x0, u0, xdot0, u))
    x = np.array(x, dtype=np.float64)
    lib().orc_lss_step(A.shape[0], B.shape[1], _dp(A), _dp(B), _dp(x0), _dp(u0), _dp(xdot0), _dp(x), _dp(u), float(dt), int(n_steps))
    return x


def segment_end(p1, s, chi, dh=None, gamma=None):
    """Segment(p1; s, chi, gamma | dh).p2 with p1 = (lat, lon, h_e) (c172x_gdc.jl:64-88)."""
    assert (dh is None) != (gamma is None)
    if dh is None:
        dh = s * np.tan(gamma)
    out = np.zeros(3)
    lib().orc_segment_end(float(p1[0]), float(p1[1]), float(p1[2]), float(s), float(chi), float(dh), _dp(out))
    return tuple(out)


SEG_FIELDS = ("chi_12", "gamma_12", "s_12", "s_1b", "s_2b", "e_sb", "v_sb", "h_s")


def segment_data(p1, p2, ob):
    """SegmentGuidanceData(Segment(p1, p2), Ob) as a dict (c172x_gdc.jl:113-148)."""
    a, b, o, out = (np.array(v, dtype=np.float64) for v in (p1, p2, ob, np.zeros(8)))
    lib().orc_segment_data(_dp(a), _dp(b), _dp(o), _dp(out))
    return dict(zip(SEG_FIELDS, out))


def segment_valid(p1, p2):
    a, b = np.array(p1, dtype=np.float64), np.array(p2, dtype=np.float64)
    return bool(lib().orc_segment_valid(_dp(a), _dp(b)))


def batch_step(sims, n_steps, threads):
    arr = (C.c_void_p * len(sims))(*[s.h for s in sims])
    sims[0].lib.orc_batch_step(arr, len(sims), int(n_steps), int(threads))


def ops_reset(counting_lib):
    counting_lib.orc_ops_reset()


def ops_get(counting_lib):
    out = (C.c_ulonglong * 4)()
    counting_lib.orc_ops_get(out)
    return dict(plain=out[0], div=out[1], sqrt=out[2], trans=out[3], total=sum(out))
