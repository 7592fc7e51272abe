"""flightbatch — host-side mirror of the Flight.jl user API for the batched C172 RK4 path.

Reference surface kept (paths relative to the reference root):
  Simulation(mdl; dt, Δt, t_start, t_end, saveat)   lib/FlightCore/src/sim.jl:183-196  -> BatchSimulation
  init!(sim, init)                                   sim.jl:390-414                      -> init_b / sim.init
  step!(sim, Δt, true)                               sim.jl:386                          -> step_b / sim.step
  run!(sim)                                          sim.jl:611-638                      -> run_b / sim.run
  TimeSeries(sim)                                    sim.jl:644-704                      -> TimeSeries
  C172.TrimParameters / C172.Init(KinInit, ...)      lib/FlightApps/src/c172/c172.jl:773-818
  Cessna172Sv0 / Cessna172Xv0/v1/v2, WA / NED, SimpleWorld
Everything numerical happens in libflightbatch.so (hand-written sm_100a CUDA behind the C ABI of
include/flightbatch.h), loaded with ctypes. There is no CPU fallback: importing works without a GPU,
but creating a BatchSimulation raises when the library or a CUDA device is missing.
"""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLIGHTBATCH_LIB") or os.path.join(_HERE, "libflightbatch.so")   # override: development builds only
TABLES_PATH = os.path.join(_HERE, "data", "c172_tables.bin")
HEADER_PATH = os.path.join(_HERE, "..", "include", "flightbatch.h")


class FlightBatchError(RuntimeError):
    pass


# ----------------------------------------------------------------------------- enums from the C header
def _parse_header():
    txt = open(HEADER_PATH).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(FB_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", txt)}
    enums = {}
    for name in ("fb_input", "fb_output"):
        body = re.search(r"enum\s+%s\s*\{(.*?)\};" % name, txt, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        val = -1
        out = {}
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = [t.strip() for t in item.split("=")]
                val = int(v)
            else:
                k = item
                val += 1
            out[k] = val
        enums[name] = out
    return defs, enums


DEFS, _ENUMS = _parse_header()
U = _ENUMS["fb_input"]     # e.g. U["FB_U_ELEVATOR"]
Y = _ENUMS["fb_output"]
globals().update(DEFS)
globals().update(U)
globals().update(Y)

INPUT_NAMES = {k[len("FB_U_"):].lower(): v for k, v in U.items() if not k.endswith("_COUNT")}
OUTPUT_NAMES = {k[len("FB_Y_"):].lower(): v for k, v in Y.items() if not k.endswith("_COUNT")}


# ----------------------------------------------------------------------------- ctypes structs
class fb_config(C.Structure):
    _fields_ = [("variant", C.c_int32), ("kinematics", C.c_int32), ("n_members", C.c_int64), ("member_offset", C.c_int64),
                ("dt", C.c_double), ("Dt", C.c_double), ("t_start", C.c_double), ("t_end", C.c_double),
                ("device", C.c_int32), ("log_every", C.c_int32), ("log_n_fields", C.c_int32), ("log_fields", C.c_int32 * 32),
                ("steps_per_launch", C.c_int32), ("log_flags", C.c_int32),
                ("prop_pitch_n", C.c_int32), ("reserved0", C.c_int32), ("prop_pitch_lo", C.c_double), ("prop_pitch_hi", C.c_double),
                ("act2_omega_n", C.c_double), ("act2_zeta", C.c_double)]


_DP = C.POINTER(C.c_double)


class fb_trim_params_soa(C.Structure):
    _fields_ = [(k, _DP) for k in ("lat", "lon", "h_e", "psi_nb", "EAS", "gamma_wb_n", "psi_wb_dot", "theta_wb_dot",
                                   "beta_a", "fuel_load", "mixture", "flaps")] + [("payload", _DP * 5)]


class fb_init_soa(C.Structure):
    _fields_ = ([("q_nb", _DP * 4), ("lat", _DP), ("lon", _DP), ("h_e", _DP), ("w_wb_b", _DP * 3), ("v_eb_n", _DP * 3)] +
                [(k, _DP) for k in ("n_eng", "engine_state", "throttle", "mixture", "elevator", "aileron", "rudder", "flaps",
                                    "brake_left", "brake_right", "fuel_load")] +
                [("payload", _DP * 5), ("stall", _DP), ("alpha_filt", _DP), ("beta_filt", _DP)])


class fb_moments(C.Structure):
    _fields_ = [("count", C.c_int64), ("sum", C.c_double), ("sum_sq", C.c_double), ("min", C.c_double), ("max", C.c_double)]


EXPORTS = ["fb_api_version", "fb_last_error", "fb_create", "fb_load_tables", "fb_destroy", "fb_sync", "fb_init_trim", "fb_init",
           "fb_init_state", "fb_set_input", "fb_get_input", "fb_schedule_inputs", "fb_step", "fb_run", "fb_time", "fb_nx",
           "fb_get_state", "fb_get_state_async", "fb_wait_state", "fb_get_discrete", "fb_get_output", "fb_status", "fb_release_trim_failed", "fb_log_info", "fb_log_times", "fb_log_fetch",
           "fb_stats", "fb_log_stats_begin", "fb_log_stats_end", "fb_log_stats", "fb_comm_unique_id", "fb_comm_init", "fb_log_stats_allreduce", "fb_comm_time",
           "fb_group_create", "fb_group_destroy", "fb_group_size", "fb_group_shard", "fb_group_load_tables", "fb_group_init_trim", "fb_group_set_input",
           "fb_group_step", "fb_group_run", "fb_group_sync", "fb_group_get_state", "fb_group_status", "fb_group_log_stats", "fb_rhs", "fb_fstep", "fb_fperiodic", "fb_counters", "fb_linearize", "fb_linearize_dims", "fb_device_state", "fb_stream", "fb_measure_fp64_peak",
           "fb_lss_create", "fb_lss_destroy", "fb_lss_set_model", "fb_lss_from_linearize", "fb_lss_set_state", "fb_lss_get_state",
           "fb_lss_set_input", "fb_lss_get_input", "fb_lss_get_output", "fb_lss_step", "fb_lss_sync", "fb_lss_stream"]

_lib = None


def load_library():
    """Load libflightbatch.so (the CUDA engine). Raises when it has not been built: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlightBatchError(f"{LIB_PATH} not built; run __graft_entry__.build() (nvcc, sm_100a)")
    # The engine opens NCCL by soname at run time, and a process holds ONE libnccl.so.2: whichever is loaded first. If this
    # interpreter also has the pip-packaged NCCL (the one torch links against), point the engine at that file, so that a later
    # `import torch` finds the NCCL it was built for instead of an older system library the engine happened to load first.
    if "FLIGHTBATCH_NCCL" not in os.environ:
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            for d in (spec.submodule_search_locations if spec else ()):
                cand = os.path.join(d, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["FLIGHTBATCH_NCCL"] = cand
                    break
        except Exception:
            pass
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    u8p = C.POINTER(C.c_uint8)
    sig = {
        "fb_api_version": (i32, []), "fb_last_error": (C.c_char_p, []),
        "fb_create": (i32, [C.POINTER(fb_config), C.POINTER(vp)]), "fb_load_tables": (i32, [vp, C.c_char_p, C.c_size_t]),
        "fb_destroy": (i32, [vp]), "fb_sync": (i32, [vp]),
        "fb_init_trim": (i32, [vp, C.POINTER(fb_trim_params_soa), _DP, _DP]), "fb_init": (i32, [vp, C.POINTER(fb_init_soa)]),
        "fb_init_state": (i32, [vp, _DP, u8p, u8p]),
        "fb_set_input": (i32, [vp, i32, _DP, i64]), "fb_get_input": (i32, [vp, i32, _DP]),
        "fb_schedule_inputs": (i32, [vp, i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32), _DP]),
        "fb_step": (i32, [vp, dbl]), "fb_run": (i32, [vp]), "fb_time": (i32, [vp, _DP, C.POINTER(i64)]),
        "fb_nx": (i32, [vp, C.POINTER(i32)]), "fb_get_state": (i32, [vp, _DP]), "fb_get_state_async": (i32, [vp, _DP]), "fb_wait_state": (i32, [vp]), "fb_get_discrete": (i32, [vp, u8p, u8p]),
        "fb_get_output": (i32, [vp, i32, _DP]), "fb_status": (i32, [vp, C.POINTER(i32)]),
        "fb_release_trim_failed": (i32, [vp, C.POINTER(i64)]),
        "fb_log_info": (i32, [vp, C.POINTER(i64), C.POINTER(i32)]), "fb_log_times": (i32, [vp, _DP]),
        "fb_log_fetch": (i32, [vp, i32, _DP]), "fb_stats": (i32, [vp, i32, C.POINTER(fb_moments)]),
        "fb_rhs": (i32, [vp, _DP]), "fb_fstep": (i32, [vp, C.POINTER(i32)]), "fb_fperiodic": (i32, [vp]), "fb_counters": (i32, [vp, C.POINTER(i64), C.POINTER(i64)]),
        "fb_device_state": (i32, [vp, C.POINTER(vp), C.POINTER(i64)]), "fb_stream": (i32, [vp, C.POINTER(vp)]),
        "fb_measure_fp64_peak": (i32, [i32, _DP]),
        "fb_linearize_dims": (i32, [vp, C.POINTER(i32)]), "fb_linearize": (i32, [vp] + [_DP] * 8),
        "fb_lss_create": (i32, [i32, i64, i32, i32, i32, dbl, C.POINTER(vp)]), "fb_lss_destroy": (i32, [vp]),
        "fb_lss_set_model": (i32, [vp] + [_DP] * 8), "fb_lss_from_linearize": (i32, [vp, vp]),
        "fb_lss_set_state": (i32, [vp, _DP]), "fb_lss_get_state": (i32, [vp, _DP]), "fb_lss_set_input": (i32, [vp, _DP]),
        "fb_lss_get_input": (i32, [vp, _DP]), "fb_lss_get_output": (i32, [vp, _DP]), "fb_lss_step": (i32, [vp, dbl]),
        "fb_lss_sync": (i32, [vp]), "fb_lss_stream": (i32, [vp, C.POINTER(vp)]),
        "fb_log_stats_begin": (i32, [vp, i64, i64]), "fb_log_stats_end": (i32, [vp, i32, C.POINTER(fb_moments)]),
        "fb_log_stats": (i32, [vp, i32, i64, i64, C.POINTER(fb_moments)]),
        "fb_comm_unique_id": (i32, [C.c_char_p]), "fb_comm_init": (i32, [vp, i32, i32, C.c_char_p]),
        "fb_log_stats_allreduce": (i32, [vp]), "fb_comm_time": (i32, [vp, _DP, C.POINTER(i64)]),
        "fb_group_create": (i32, [C.POINTER(fb_config), C.POINTER(i32), i32, C.POINTER(vp)]), "fb_group_destroy": (i32, [vp]),
        "fb_group_size": (i32, [vp, C.POINTER(i32)]), "fb_group_shard": (i32, [vp, i32, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64)]),
        "fb_group_load_tables": (i32, [vp, C.c_char_p, C.c_size_t]), "fb_group_init_trim": (i32, [vp, C.POINTER(fb_trim_params_soa), _DP, _DP]),
        "fb_group_set_input": (i32, [vp, i32, _DP, i64]), "fb_group_step": (i32, [vp, dbl]), "fb_group_run": (i32, [vp]), "fb_group_sync": (i32, [vp]),
        "fb_group_get_state": (i32, [vp, _DP]), "fb_group_status": (i32, [vp, C.POINTER(i32)]),
        "fb_group_log_stats": (i32, [vp, i32, i64, i64, C.POINTER(fb_moments)]),
    }
    for name in EXPORTS:
        f = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        f.restype, f.argtypes = sig[name]
    _lib = lib
    return lib


def _ck(rc, lib):
    if rc != 0:
        raise FlightBatchError(f"flightbatch error {rc}: {lib.fb_last_error().decode()}")


def _dp(a):
    return a.ctypes.data_as(_DP)


# ----------------------------------------------------------------------------- model definitions
class WA:
    kin = DEFS["FB_WA"]


class NED:
    kin = DEFS["FB_NED"]


class ECEF:
    kin = DEFS["FB_ECEF"]


class _Aircraft:
    variant = None

    def __init__(self, kinematics=None):
        self.kinematics = kinematics if kinematics is not None else WA()


class Cessna172Sv0(_Aircraft):
    variant = DEFS["FB_C172S"]


class Cessna172Xv0(_Aircraft):
    variant = DEFS["FB_C172X0"]


class Cessna172Xv1(_Aircraft):
    variant = DEFS["FB_C172X1"]


class Cessna172Xv2(_Aircraft):
    variant = DEFS["FB_C172X2"]


class FixedPitch:
    """Propellers.FixedPitch — lib/FlightPhysics/src/propellers.jl:337-339 (what every C172 of the reference flies)"""
    n = 0
    lower = upper = 0.0


class VariablePitch:
    """Propellers.VariablePitch(lower, upper; length = 11) — propellers.jl:341-349: the blade-pitch offset follows the
    propeller input `prop_pitch` in [0, 1] linearly over [lower, upper] (rad)"""

    def __init__(self, lower, upper, length=11):
        self.lower, self.upper, self.n = float(lower), float(upper), int(length)


class Actuator2:
    """C172X.Actuator2(; ω_n = 5 * 2π, ζ = 0.6) — lib/FlightApps/src/c172/c172x/c172x.jl:68-106"""

    def __init__(self, omega_n=5 * 2 * np.pi, zeta=0.6):
        self.omega_n, self.zeta = float(omega_n), float(zeta)


class Cessna172Xv0A2(_Aircraft):
    """Cessna172Xv0 assembled from the non-default components the reference's Model tree admits: a
    FlyByWireActuation of seven Actuator2 units (x.act = (v, p) per unit: nx grows by 7) and a FixedPitch or VariablePitch
    propeller. NoAvionics."""
    variant = DEFS["FB_C172XA2"]

    def __init__(self, kinematics=None, propeller=None, actuator=None):
        super().__init__(kinematics)
        self.propeller = propeller if propeller is not None else FixedPitch()
        self.actuator = actuator if actuator is not None else Actuator2()


class SimpleWorld:
    """SimpleWorld(aircraft, SimpleAtmosphere(), HorizontalTerrain()) — lib/FlightPhysics/src/world.jl:20-24"""

    def __init__(self, aircraft=None):
        self.aircraft = aircraft if aircraft is not None else Cessna172Sv0()


class TrimParameters:
    """C172.TrimParameters (c172.jl:806-818), one value or one array of N values per field."""
    FIELDS = dict(lat=0.0, lon=0.0, h_e=1050.0, psi_nb=0.0, EAS=50.0, gamma_wb_n=0.0, psi_wb_dot=0.0, theta_wb_dot=0.0,
                  beta_a=0.0, fuel_load=0.5, mixture=0.5, flaps=0.0,
                  m_pilot=75.0, m_copilot=75.0, m_lpass=0.0, m_rpass=0.0, m_baggage=50.0)

    def __init__(self, **kw):
        for k in kw:
            if k not in self.FIELDS:
                raise TypeError(f"TrimParameters has no field {k}")
        self.values = dict(kw)


# WGS84 (lib/FlightPhysics/src/geodesy.jl:23-38)
_WGS_A, _WGS_F = 6378137.0, 1.0 / 298.257223563
_WGS_E2 = 2 * _WGS_F - _WGS_F * _WGS_F


def _cartesian(lat, lon, h):
    """Cartesian(Geographic{LatLon,Ellipsoidal}) (geodesy.jl:418-428): ECEF position [3, ...]."""
    sl, cl = np.sin(lat), np.cos(lat)
    N = _WGS_A / np.sqrt(1.0 - _WGS_E2 * sl * sl)
    return np.stack([(N + h) * cl * np.cos(lon), (N + h) * cl * np.sin(lon), (N * (1 - _WGS_E2) + h) * sl])


def _ned_axes(lat, lon):
    """Unit vectors of the local tangent frame NED(lat, lon) in ECEF axes (the rotation ltf() encodes, geodesy.jl:132-135)."""
    sl, cl, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
    north = np.stack([-sl * co, -sl * so, cl])
    east = np.stack([-so, co, np.zeros_like(sl)])
    return north, east


def _latlon_from_cartesian(r):
    """LatLon(Cartesian) (geodesy.jl:367-412: Fukushima's method, one Halley iteration)."""
    x, y, z = r
    p = np.sqrt(x * x + y * y)
    c, ec2 = _WGS_A * _WGS_E2, 1.0 - _WGS_E2
    ec = np.sqrt(ec2)
    zc, s0, c0 = ec * np.abs(z), np.abs(z), ec * p
    a0 = np.sqrt(s0 * s0 + c0 * c0)
    a03 = a0 ** 3
    b0 = 1.5 * c * s0 * c0 * ((p * s0 - zc * c0) * a0 - c * s0 * c0)
    s1 = (zc * a03 + c * s0 ** 3) * a03 - b0 * s0
    cc = ec * ((p * a03 - c * c0 ** 3) * a03 - b0 * c0)
    lat = np.sign(z) * np.arctan2(s1, cc)
    return lat, np.arctan2(y, x)


class Segment:
    """C172XGuidance.Segment (c172x_gdc.jl:38-98): a straight target segment p1 -> p2, each point (lat, lon, h_e) as
    Geographic{LatLon,Ellipsoidal}; scalars or arrays of N values. Vertical or degenerate segments are refused as the
    reference does (horizontal length seen from p1 below 1e-6 m -> ValueError, the reference's ArgumentError)."""

    def __init__(self, p1, p2):
        self.p1 = tuple(np.asarray(v, dtype=np.float64) for v in p1)
        self.p2 = tuple(np.asarray(v, dtype=np.float64) for v in p2)
        r12 = _cartesian(*self.p2) - _cartesian(*self.p1)
        north, east = _ned_axes(self.p1[0], self.p1[1])
        s12 = np.hypot((north * r12).sum(axis=0), (east * r12).sum(axis=0))
        if np.any(s12 < 1e-6):
            raise ValueError("Segment is vertical or degenerate")

    @classmethod
    def from_origin(cls, p1, s, chi, gamma=None, dh=None):
        """Segment(p1; s, χ, γ | Δh) (c172x_gdc.jl:64-88): horizontal length s along azimuth χ from p1."""
        if (gamma is None) == (dh is None):
            raise TypeError("provide either gamma (flight path angle) or dh (altitude increment)")
        lat1, lon1, h1 = (np.asarray(v, dtype=np.float64) for v in p1)
        s, chi = np.asarray(s, dtype=np.float64), np.asarray(chi, dtype=np.float64)
        if dh is None:
            dh = s * np.tan(gamma)
        north, east = _ned_axes(lat1, lon1)
        r2 = _cartesian(lat1, lon1, h1) + north * (s * np.cos(chi)) + east * (s * np.sin(chi))
        lat2, lon2 = _latlon_from_cartesian(r2)
        return cls((lat1, lon1, h1), (lat2, lon2, h1 + dh))

    def __neg__(self):
        return Segment(self.p2, self.p1)


class Init:
    """C172.Init(KinInit(q_nb, location, h, ω_wb_b, v_eb_n), SystemsInitializer(...)) — c172.jl:773-794"""
    FIELDS = dict(q0=1.0, q1=0.0, q2=0.0, q3=0.0, lat=0.0, lon=0.0, h_e=0.0, w0=0.0, w1=0.0, w2=0.0, vN=0.0, vE=0.0, vD=0.0,
                  n_eng=0.0, engine_state=0.0, throttle=0.0, mixture=0.5, elevator=0.0, aileron=0.0, rudder=0.0, flaps=0.0,
                  brake_left=0.0, brake_right=0.0, fuel_load=0.5, m_pilot=75.0, m_copilot=75.0, m_lpass=0.0, m_rpass=0.0,
                  m_baggage=50.0, stall=0.0, alpha_filt=0.0, beta_filt=0.0)

    def __init__(self, **kw):
        for k in kw:
            if k not in self.FIELDS:
                raise TypeError(f"Init has no field {k}")
        self.values = dict(kw)


class C172:
    TrimParameters = TrimParameters
    Init = Init


def _make_config(world, n, dt, Dt, t_start, t_end, saveat, log_fields, device, member_offset, steps_per_launch, log_members, log_moments):
    world = world if world is not None else SimpleWorld()
    ac = world.aircraft
    cfg = fb_config()
    cfg.variant = ac.variant
    cfg.kinematics = ac.kinematics.kin
    cfg.n_members = int(n)
    cfg.member_offset = int(member_offset)
    cfg.dt = float(dt)
    cfg.Dt = float(dt if Dt is None else Dt)
    cfg.t_start = float(t_start)
    cfg.t_end = float(t_end)
    cfg.device = int(device)
    fields = [OUTPUT_NAMES[f] if isinstance(f, str) else int(f) for f in log_fields]
    cfg.log_every = 0 if (saveat is None or not fields) else max(1, int(round(saveat / dt)))
    cfg.log_n_fields = len(fields)
    for i, f in enumerate(fields):
        cfg.log_fields[i] = f
    cfg.steps_per_launch = int(steps_per_launch)
    cfg.log_flags = (DEFS["FB_LOG_MEMBERS"] if log_members else 0) | (DEFS["FB_LOG_MOMENTS"] if log_moments else 0)
    if isinstance(ac, Cessna172Xv0A2):
        cfg.prop_pitch_n = ac.propeller.n
        cfg.prop_pitch_lo, cfg.prop_pitch_hi = ac.propeller.lower, ac.propeller.upper
        cfg.act2_omega_n, cfg.act2_zeta = ac.actuator.omega_n, ac.actuator.zeta
    return cfg, fields


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, the host distributes it, every rank passes it to comm_init)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    _ck(lib.fb_comm_unique_id(buf), lib)
    return buf.raw


class _InputProxy:
    """sim.u.<field>: read returns an [N] array, assignment broadcasts a scalar or uploads an [N] array.
    Mirrors direct mutation of nested inputs in the reference (README.md:166)."""

    def __init__(self, sim):
        object.__setattr__(self, "_sim", sim)

    def __getattr__(self, name):
        if name not in INPUT_NAMES:
            raise AttributeError(name)
        return self._sim.get_input(INPUT_NAMES[name])

    def __setattr__(self, name, value):
        if name not in INPUT_NAMES:
            raise AttributeError(name)
        self._sim.set_input(INPUT_NAMES[name], value)

    __getitem__ = __getattr__
    __setitem__ = __setattr__


class _OutputProxy:
    def __init__(self, sim):
        self._sim = sim

    def __getattr__(self, name):
        if name not in OUTPUT_NAMES:
            raise AttributeError(name)
        return self._sim.get_output(OUTPUT_NAMES[name])

    __getitem__ = __getattr__


class BatchSimulation:
    """N independent `Simulation(world; dt, Δt, t_start, t_end, saveat)` stepped in lock-step on one B200."""

    def __init__(self, world=None, n=1, dt=0.02, Dt=None, t_start=0.0, t_end=10000.0, saveat=None, log_fields=(),
                 device=0, member_offset=0, steps_per_launch=0, tables_path=None, log_members=True, log_moments=True, _adopt=None):
        """saveat / log_fields: the outputs saved every `saveat` seconds (TimeSeries(sim)); log_members=False keeps only the
        ensemble moments of each sample (log_stats), log_moments=False only the per-member samples."""
        self._h = None
        self._lib = load_library()
        cfg, fields = _make_config(world, n, dt, Dt, t_start, t_end, saveat, log_fields, device, member_offset, steps_per_launch,
                                   log_members, log_moments)
        self.cfg = cfg
        self.n = int(n)
        self.dt = float(dt)
        self.log_fields = fields
        self._owned = _adopt is None
        if _adopt is None:
            h = C.c_void_p()
            _ck(self._lib.fb_create(C.byref(cfg), C.byref(h)), self._lib)
            self._h = h
            blob = open(tables_path or TABLES_PATH, "rb").read()
            _ck(self._lib.fb_load_tables(self._h, blob, len(blob)), self._lib)
        else:
            self._h = _adopt      # a shard of a BatchGroup: the group owns the handle
        nx = C.c_int32()
        _ck(self._lib.fb_nx(self._h, C.byref(nx)), self._lib)
        self.nx = nx.value
        self.u = _InputProxy(self)
        self.y = _OutputProxy(self)

    def close(self):
        if self._h is not None:
            if self._owned:
                self._lib.fb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers
    def _col(self, v, keep):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.n,)))
        keep.append(a)
        return _dp(a)

    # -- init!(sim, ...)
    def init(self, init=None, **kw):
        """init!(sim, C172.TrimParameters(...)) or init!(sim, C172.Init(...)); returns (success[N], trim_state[7,N], cost[N]) for trims."""
        if init is None:
            init = TrimParameters(**kw)
        keep = []
        if isinstance(init, TrimParameters):
            p = fb_trim_params_soa()
            v = init.values
            for k in ("lat", "lon", "h_e", "psi_nb", "EAS", "gamma_wb_n", "psi_wb_dot", "theta_wb_dot", "beta_a", "fuel_load", "mixture", "flaps"):
                if k in v:
                    setattr(p, k, self._col(v[k], keep))
            for j, k in enumerate(("m_pilot", "m_copilot", "m_lpass", "m_rpass", "m_baggage")):
                if k in v:
                    p.payload[j] = self._col(v[k], keep)
            ts = np.zeros((7, self.n))
            cost = np.zeros(self.n)
            _ck(self._lib.fb_init_trim(self._h, C.byref(p), _dp(ts), _dp(cost)), self._lib)
            return self.status() != DEFS["FB_ST_TRIM_FAILED"], ts, cost
        if isinstance(init, Init):
            ic = fb_init_soa()
            v = init.values
            for j in range(4):
                if f"q{j}" in v:
                    ic.q_nb[j] = self._col(v[f"q{j}"], keep)
            for j in range(3):
                if f"w{j}" in v:
                    ic.w_wb_b[j] = self._col(v[f"w{j}"], keep)
            for j, k in enumerate(("vN", "vE", "vD")):
                if k in v:
                    ic.v_eb_n[j] = self._col(v[k], keep)
            for k in ("lat", "lon", "h_e", "n_eng", "engine_state", "throttle", "mixture", "elevator", "aileron", "rudder", "flaps",
                      "brake_left", "brake_right", "fuel_load", "stall", "alpha_filt", "beta_filt"):
                if k in v:
                    setattr(ic, k, self._col(v[k], keep))
            for j, k in enumerate(("m_pilot", "m_copilot", "m_lpass", "m_rpass", "m_baggage")):
                if k in v:
                    ic.payload[j] = self._col(v[k], keep)
            _ck(self._lib.fb_init(self._h, C.byref(ic)), self._lib)
            return None
        raise TypeError("init! expects C172.TrimParameters or C172.Init")

    def init_state(self, x, stall=None, eng_state=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape == (self.nx, self.n)
        u8 = C.POINTER(C.c_uint8)
        st = None if stall is None else np.ascontiguousarray(stall, dtype=np.uint8)
        en = None if eng_state is None else np.ascontiguousarray(eng_state, dtype=np.uint8)
        _ck(self._lib.fb_init_state(self._h, _dp(x), st.ctypes.data_as(u8) if st is not None else None,
                                    en.ctypes.data_as(u8) if en is not None else None), self._lib)

    # -- inputs
    def set_input(self, field, value):
        field = INPUT_NAMES[field] if isinstance(field, str) else int(field)
        a = np.asarray(value, dtype=np.float64)
        if a.ndim == 0:
            a1 = np.array([float(a)])
            _ck(self._lib.fb_set_input(self._h, field, _dp(a1), 0), self._lib)
        else:
            a = np.ascontiguousarray(a)
            assert a.shape == (self.n,)
            _ck(self._lib.fb_set_input(self._h, field, _dp(a), 1), self._lib)

    def set_segment(self, seg):
        """avionics.gdc.seg.u.target = seg (Cessna172Xv2): one Segment for all members or one per member."""
        for k, v in zip(("lat1", "lon1", "h1", "lat2", "lon2", "h2"), seg.p1 + seg.p2):
            self.set_input("gdc_" + k, v if np.ndim(v) else float(v))

    def get_input(self, field):
        field = INPUT_NAMES[field] if isinstance(field, str) else int(field)
        out = np.zeros(self.n)
        _ck(self._lib.fb_get_input(self._h, field, _dp(out)), self._lib)
        return out

    def schedule_inputs(self, member, step, field, value):
        """Declarative user_callback! replacement: piecewise-constant inputs (see fb_schedule_inputs)."""
        member = np.ascontiguousarray(member, dtype=np.int64)
        step = np.ascontiguousarray(step, dtype=np.int64)
        field = np.ascontiguousarray(field, dtype=np.int32)
        value = np.ascontiguousarray(value, dtype=np.float64)
        n = len(member)
        assert len(step) == n and len(field) == n and len(value) == n
        _ck(self._lib.fb_schedule_inputs(self._h, n, member.ctypes.data_as(C.POINTER(C.c_int64)), step.ctypes.data_as(C.POINTER(C.c_int64)),
                                         field.ctypes.data_as(C.POINTER(C.c_int32)), _dp(value)), self._lib)

    # -- step! / run!
    def step(self, t_advance=None):
        _ck(self._lib.fb_step(self._h, float(self.dt if t_advance is None else t_advance)), self._lib)

    def run(self):
        _ck(self._lib.fb_run(self._h), self._lib)

    def sync(self):
        _ck(self._lib.fb_sync(self._h), self._lib)

    # -- read-back
    @property
    def t(self):
        t = C.c_double()
        ns = C.c_int64()
        _ck(self._lib.fb_time(self._h, C.byref(t), C.byref(ns)), self._lib)
        return t.value

    @property
    def n_steps(self):
        t = C.c_double()
        ns = C.c_int64()
        _ck(self._lib.fb_time(self._h, C.byref(t), C.byref(ns)), self._lib)
        return ns.value

    @property
    def x(self):
        return self.get_state()

    def get_state(self, out=None):
        """sim.x as an [nx][N] array. `out` may be a preallocated C-contiguous float64 array (e.g. the numpy view of a
        pinned host buffer, which makes the device-to-host copy run at PCIe speed); it is filled and returned."""
        if out is None:
            out = np.empty((self.nx, self.n))
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.shape != (self.nx, self.n):
            raise FlightBatchError(f"get_state: out must be a C-contiguous float64 array of shape {(self.nx, self.n)}")
        _ck(self._lib.fb_get_state(self._h, _dp(out)), self._lib)
        return out

    def get_state_async(self, out):
        """Pipelined sim.x: snapshot the state as of the work submitted so far and copy it into `out` (an [nx][N] float64
        array over PINNED host memory) while later step() calls already run. `out` is valid after wait_state()."""
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.shape != (self.nx, self.n):
            raise FlightBatchError(f"get_state_async: out must be a C-contiguous float64 array of shape {(self.nx, self.n)}")
        self._snap_out = out   # keep the buffer alive while the copy is in flight
        _ck(self._lib.fb_get_state_async(self._h, _dp(out)), self._lib)
        return out

    def wait_state(self):
        _ck(self._lib.fb_wait_state(self._h), self._lib)
        self._snap_out = None

    @property
    def xdot(self):
        out = np.zeros((self.nx, self.n))
        _ck(self._lib.fb_rhs(self._h, _dp(out)), self._lib)
        return out

    def discrete(self):
        st = np.zeros(self.n, dtype=np.uint8)
        en = np.zeros(self.n, dtype=np.uint8)
        u8 = C.POINTER(C.c_uint8)
        _ck(self._lib.fb_get_discrete(self._h, st.ctypes.data_as(u8), en.ctypes.data_as(u8)), self._lib)
        return st, en

    def status(self):
        out = np.zeros(self.n, dtype=np.int32)
        _ck(self._lib.fb_status(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))), self._lib)
        return out

    def release_trim_failed(self):
        """The reference's handling of a failed trim (c172.jl:934-940): warn, keep the best trim state found, go on simulating.
        Members frozen with FB_ST_TRIM_FAILED become RUNNING again; returns how many."""
        n = C.c_int64(0)
        _ck(self._lib.fb_release_trim_failed(self._h, C.byref(n)), self._lib)
        return int(n.value)

    def get_output(self, field):
        field = OUTPUT_NAMES[field] if isinstance(field, str) else int(field)
        out = np.zeros(self.n)
        _ck(self._lib.fb_get_output(self._h, field, _dp(out)), self._lib)
        return out

    def stats(self, field):
        field = OUTPUT_NAMES[field] if isinstance(field, str) else int(field)
        m = fb_moments()
        _ck(self._lib.fb_stats(self._h, field, C.byref(m)), self._lib)
        return dict(count=m.count, sum=m.sum, sum_sq=m.sum_sq, min=m.min, max=m.max)

    def log_stats_begin(self, first=0, count=None):
        """Start the asynchronous fetch of the ensemble moments of log samples [first, first + count)."""
        ns, nf = C.c_int64(), C.c_int32()
        _ck(self._lib.fb_log_info(self._h, C.byref(ns), C.byref(nf)), self._lib)
        count = ns.value - first if count is None else count
        _ck(self._lib.fb_log_stats_begin(self._h, int(first), int(count)), self._lib)
        self._stats_range = (int(first), int(count))

    def log_stats_end(self):
        """Wait for log_stats_begin: {field name: {count, sum, sum_sq, min, max, mean, std: arrays over the samples}}."""
        first, count = self._stats_range
        names = {v: k for k, v in OUTPUT_NAMES.items()}
        out = {}
        for f in self.log_fields:
            m = (fb_moments * count)()
            _ck(self._lib.fb_log_stats_end(self._h, int(f), m), self._lib)
            d = dict(count=np.array([r.count for r in m], dtype=np.int64), sum=np.array([r.sum for r in m]), sum_sq=np.array([r.sum_sq for r in m]),
                     min=np.array([r.min for r in m]), max=np.array([r.max for r in m]))
            with np.errstate(invalid="ignore", divide="ignore"):
                d["mean"] = d["sum"] / d["count"]
                d["std"] = np.sqrt(np.maximum(d["sum_sq"] / d["count"] - d["mean"] ** 2, 0.0))
            out[names[f]] = d
        return out

    def log_stats(self, first=0, count=None):
        """Ensemble moments of the logged fields at log samples [first, first + count), accumulated inside the integrator as
        each sample was taken (after log_stats_allreduce: over every shard of the ensemble)."""
        self.log_stats_begin(first, count)
        return self.log_stats_end()

    def comm_init(self, n_ranks, rank, unique_id):
        """Join the NCCL communicator of the ensemble-moment all-reduce (unique_id: the 128 bytes of comm_unique_id() of rank 0)."""
        _ck(self._lib.fb_comm_init(self._h, int(n_ranks), int(rank), bytes(unique_id)), self._lib)

    def log_stats_allreduce(self):
        """All-reduce (NCCL, side stream, asynchronous) the moments of the log samples completed since the previous call."""
        _ck(self._lib.fb_log_stats_allreduce(self._h), self._lib)

    def comm_time_ms(self):
        ms, n = C.c_double(), C.c_int64()
        _ck(self._lib.fb_comm_time(self._h, C.byref(ms), C.byref(n)), self._lib)
        return ms.value

    def linearize(self):
        """linearize(aircraft, trim) for every member (NED kinematics, trimmed state): dict of A [N, nx, nx], B [N, nx, nu],
        C [N, ny, nx], D [N, ny, nu] (row = output, column = perturbed variable) and x0, u0, y0, xdot0 [N, n]."""
        d = (C.c_int32 * 3)()
        _ck(self._lib.fb_linearize_dims(self._h, d), self._lib)
        nx, nu, ny = d[0], d[1], d[2]
        raw = dict(A=np.zeros((nx, nx, self.n)), B=np.zeros((nu, nx, self.n)), C=np.zeros((nx, ny, self.n)), D=np.zeros((nu, ny, self.n)),
                   x0=np.zeros((nx, self.n)), u0=np.zeros((nu, self.n)), y0=np.zeros((ny, self.n)), xdot0=np.zeros((nx, self.n)))
        _ck(self._lib.fb_linearize(self._h, *[_dp(raw[k]) for k in ("A", "B", "C", "D", "x0", "u0", "y0", "xdot0")]), self._lib)
        out = {k: np.transpose(raw[k], (2, 1, 0)) for k in ("A", "B", "C", "D")}      # [col][row][N] -> [N][row][col]
        out.update({k: raw[k].T for k in ("x0", "u0", "y0", "xdot0")})
        return out

    def f_step(self):
        """f_step!(world) at the current state (test hook); returns the per-member `changed` flags."""
        out = np.zeros(self.n, dtype=np.int32)
        _ck(self._lib.fb_fstep(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))), self._lib)
        return out

    def f_periodic(self):
        """f_periodic!(avionics, vehicle) + assign! at the current state (test hook)."""
        _ck(self._lib.fb_fperiodic(self._h), self._lib)

    def counters(self):
        a, b = C.c_int64(), C.c_int64()
        _ck(self._lib.fb_counters(self._h, C.byref(a), C.byref(b)), self._lib)
        return dict(kernel_launches=a.value, rhs_evals=b.value)

    def stream_handle(self):
        p = C.c_void_p()
        _ck(self._lib.fb_stream(self._h, C.byref(p)), self._lib)
        return p.value

    def device_state(self):
        p = C.c_void_p()
        pitch = C.c_int64()
        _ck(self._lib.fb_device_state(self._h, C.byref(p), C.byref(pitch)), self._lib)
        return p.value, pitch.value


class BatchGroup:
    """One process driving several GPUs: the ensemble sharded in contiguous blocks over `devices`, host arrays global.
    `shards[k]` is a BatchSimulation view of shard k (members offsets[k] .. offsets[k + 1]) for everything else."""

    def __init__(self, world=None, n=1, devices=(0,), dt=0.02, Dt=None, t_start=0.0, t_end=10000.0, saveat=None, log_fields=(),
                 steps_per_launch=0, tables_path=None, log_members=True, log_moments=True):
        self._g = None
        self._lib = load_library()
        cfg, fields = _make_config(world, n, dt, Dt, t_start, t_end, saveat, log_fields, 0, 0, steps_per_launch, log_members, log_moments)
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        g = C.c_void_p()
        _ck(self._lib.fb_group_create(C.byref(cfg), devs, len(devices), C.byref(g)), self._lib)
        self._g = g
        blob = open(tables_path or TABLES_PATH, "rb").read()
        _ck(self._lib.fb_group_load_tables(self._g, blob, len(blob)), self._lib)
        self.n, self.dt, self.log_fields = int(n), float(dt), fields
        self.shards, self.offsets = [], [0]
        for k in range(len(devices)):
            h, off, cnt = C.c_void_p(), C.c_int64(), C.c_int64()
            _ck(self._lib.fb_group_shard(self._g, k, C.byref(h), C.byref(off), C.byref(cnt)), self._lib)
            self.shards.append(BatchSimulation(world, n=cnt.value, dt=dt, Dt=Dt, t_start=t_start, t_end=t_end, saveat=saveat, log_fields=log_fields,
                                               device=devices[k], member_offset=off.value, steps_per_launch=steps_per_launch,
                                               log_members=log_members, log_moments=log_moments, _adopt=h))
            self.offsets.append(off.value + cnt.value)
        self.nx = self.shards[0].nx

    def close(self):
        if self._g is not None:
            for s in self.shards:
                s._h = None
            self._lib.fb_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self, init=None, **kw):
        if init is None:
            init = TrimParameters(**kw)
        if not isinstance(init, TrimParameters):
            raise TypeError("BatchGroup.init takes C172.TrimParameters (use shards[k].init for C172.Init)")
        keep, p, v = [], fb_trim_params_soa(), init.values

        def col(a):
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (self.n,)))
            keep.append(a)
            return _dp(a)
        for k in ("lat", "lon", "h_e", "psi_nb", "EAS", "gamma_wb_n", "psi_wb_dot", "theta_wb_dot", "beta_a", "fuel_load", "mixture", "flaps"):
            if k in v:
                setattr(p, k, col(v[k]))
        for j, k in enumerate(("m_pilot", "m_copilot", "m_lpass", "m_rpass", "m_baggage")):
            if k in v:
                p.payload[j] = col(v[k])
        ts, cost = np.zeros((7, self.n)), np.zeros(self.n)
        _ck(self._lib.fb_group_init_trim(self._g, C.byref(p), _dp(ts), _dp(cost)), self._lib)
        return self.status() != DEFS["FB_ST_TRIM_FAILED"], ts, cost

    def set_input(self, field, value):
        field = INPUT_NAMES[field] if isinstance(field, str) else int(field)
        a = np.asarray(value, dtype=np.float64)
        if a.ndim == 0:
            _ck(self._lib.fb_group_set_input(self._g, field, _dp(np.array([float(a)])), 0), self._lib)
        else:
            a = np.ascontiguousarray(a)
            assert a.shape == (self.n,)
            _ck(self._lib.fb_group_set_input(self._g, field, _dp(a), 1), self._lib)

    def step(self, t_advance=None):
        _ck(self._lib.fb_group_step(self._g, float(self.dt if t_advance is None else t_advance)), self._lib)

    def run(self):
        _ck(self._lib.fb_group_run(self._g), self._lib)

    def sync(self):
        _ck(self._lib.fb_group_sync(self._g), self._lib)

    @property
    def x(self):
        out = np.empty((self.nx, self.n))
        _ck(self._lib.fb_group_get_state(self._g, _dp(out)), self._lib)
        return out

    def status(self):
        out = np.zeros(self.n, dtype=np.int32)
        _ck(self._lib.fb_group_status(self._g, out.ctypes.data_as(C.POINTER(C.c_int32))), self._lib)
        return out

    def log_stats(self, field, first=0, count=1):
        """Moments of one logged field over the WHOLE ensemble (in-library NCCL all-reduce over the shards) at samples [first, first + count)."""
        field = OUTPUT_NAMES[field] if isinstance(field, str) else int(field)
        m = (fb_moments * count)()
        _ck(self._lib.fb_group_log_stats(self._g, field, int(first), int(count), m), self._lib)
        return [dict(count=r.count, sum=r.sum, sum_sq=r.sum_sq, min=r.min, max=r.max) for r in m]


# The reference's dotted path under the root `y` (what `TimeSeries(sim)` drills into, sim.jl:666-676) of every fb_output
# field; vector components carry the reference's 1-based index. Same table as BatchSimulation.jl.
Y_PATHS = {
    "aircraft.vehicle.kinematics.e_nb.ψ": "psi", "aircraft.vehicle.kinematics.e_nb.θ": "theta", "aircraft.vehicle.kinematics.e_nb.φ": "phi",
    "aircraft.vehicle.kinematics.ϕ_λ.ϕ": "lat", "aircraft.vehicle.kinematics.ϕ_λ.λ": "lon",
    "aircraft.vehicle.kinematics.h_e": "h_e", "aircraft.vehicle.kinematics.h_o": "h_o",
    "aircraft.vehicle.kinematics.v_eb_n[1]": "v_n", "aircraft.vehicle.kinematics.v_eb_n[2]": "v_e", "aircraft.vehicle.kinematics.v_eb_n[3]": "v_d",
    "aircraft.vehicle.kinematics.v_gnd": "v_gnd", "aircraft.vehicle.kinematics.χ_gnd": "chi_gnd", "aircraft.vehicle.kinematics.γ_gnd": "gamma_gnd",
    "aircraft.vehicle.kinematics.ω_wb_b[1]": "p", "aircraft.vehicle.kinematics.ω_wb_b[2]": "q", "aircraft.vehicle.kinematics.ω_wb_b[3]": "r",
    "aircraft.vehicle.airflow.TAS": "tas", "aircraft.vehicle.airflow.EAS": "eas", "aircraft.vehicle.airflow.CAS": "cas",
    "aircraft.vehicle.airflow.M": "mach", "aircraft.vehicle.airflow.q": "qbar",
    "aircraft.vehicle.systems.aero.α": "alpha", "aircraft.vehicle.systems.aero.β": "beta", "aircraft.vehicle.systems.aero.α_filt": "alpha_filt",
    "aircraft.vehicle.systems.aero.β_filt": "beta_filt", "aircraft.vehicle.systems.aero.stall": "stall",
    "aircraft.vehicle.systems.aero.coeffs.C_L": "c_l", "aircraft.vehicle.systems.aero.coeffs.C_D": "c_d",
    "aircraft.vehicle.systems.ldg.left.strut.wow": "wow_left", "aircraft.vehicle.systems.ldg.right.strut.wow": "wow_right",
    "aircraft.vehicle.systems.ldg.nose.strut.wow": "wow_nose", "aircraft.vehicle.systems.ldg.left.strut.ξ": "xi_left",
    "aircraft.vehicle.systems.ldg.right.strut.ξ": "xi_right", "aircraft.vehicle.systems.ldg.nose.strut.ξ": "xi_nose",
    "aircraft.vehicle.systems.ldg.left.strut.F_dmp_zs": "fdmp_left", "aircraft.vehicle.systems.ldg.right.strut.F_dmp_zs": "fdmp_right",
    "aircraft.vehicle.systems.ldg.nose.strut.F_dmp_zs": "fdmp_nose",
    "aircraft.vehicle.systems.pwp.engine.ω": "eng_omega", "aircraft.vehicle.systems.pwp.engine.state": "eng_state",
    "aircraft.vehicle.systems.pwp.engine.MAP": "eng_map", "aircraft.vehicle.systems.pwp.engine.P_shaft": "eng_p_shaft",
    "aircraft.vehicle.systems.pwp.engine.τ_shaft": "eng_tau_shaft", "aircraft.vehicle.systems.pwp.engine.ṁ": "eng_mdot",
    "aircraft.vehicle.systems.fuel.m_total": "fuel_m_total",
    "aircraft.vehicle.systems.pwp.propeller.wr_p.F[1]": "prop_thrust", "aircraft.vehicle.systems.pwp.propeller.J": "prop_j",
    "aircraft.vehicle.dynamics.mp_Σ_b.m": "mass",
    "aircraft.vehicle.dynamics.f_c_c[1]": "f_x", "aircraft.vehicle.dynamics.f_c_c[2]": "f_y", "aircraft.vehicle.dynamics.f_c_c[3]": "f_z",
    "aircraft.avionics.ctl.lon.mode": "lon_mode", "aircraft.avionics.ctl.lat.mode": "lat_mode",
    "aircraft.avionics.ctl.lon.throttle_cmd": "throttle_cmd", "aircraft.avionics.ctl.lon.elevator_cmd": "elevator_cmd",
    "aircraft.avionics.ctl.lat.aileron_cmd": "aileron_cmd", "aircraft.avionics.ctl.lat.rudder_cmd": "rudder_cmd",
    "aircraft.avionics.ctl.lon.h_state": "alt_state",
    "aircraft.avionics.gdc.mode": "gdc_mode", "aircraft.avionics.gdc.seg.hor_gdc": "gdc_hor", "aircraft.avionics.gdc.seg.vrt_gdc": "gdc_vrt",
    "aircraft.avionics.gdc.seg.Δχ": "gdc_dchi", "aircraft.avionics.gdc.seg.χ_ref": "gdc_chi_ref", "aircraft.avionics.gdc.seg.h_ref": "gdc_h_ref",
    "aircraft.avionics.gdc.seg.data.χ_12": "gdc_chi_12", "aircraft.avionics.gdc.seg.data.γ_12": "gdc_gamma_12",
    "aircraft.avionics.gdc.seg.data.s_12": "gdc_s_12", "aircraft.avionics.gdc.seg.data.s_1b": "gdc_s_1b",
    "aircraft.avionics.gdc.seg.data.e_sb": "gdc_e_sb", "aircraft.avionics.gdc.seg.data.v_sb": "gdc_v_sb",
}


class _YNode:
    """Lazy drill-down node of a TimeSeries: `ts.aircraft.vehicle.kinematics.e_nb.θ`, `...v_eb_n[0]` (0-based here)."""

    def __init__(self, ts, prefix):
        object.__setattr__(self, "_ts", ts)
        object.__setattr__(self, "_prefix", prefix)

    def _resolve(self, path):
        if path in Y_PATHS:
            return self._ts._column(Y_PATHS[path])
        if any(k.startswith(path + ".") or k.startswith(path + "[") for k in Y_PATHS):
            return _YNode(self._ts, path)
        raise AttributeError(f"TimeSeries has no logged path {path}")

    def __getattr__(self, name):
        return self._resolve(f"{self._prefix}.{name}" if self._prefix else name)

    def __getitem__(self, i):
        return self._resolve(f"{self._prefix}[{int(i) + 1}]")

    def __dir__(self):
        n = len(self._prefix) + 1 if self._prefix else 0
        return sorted({k[n:].split(".")[0].split("[")[0] for k in Y_PATHS if k.startswith(self._prefix)})


class TimeSeries:
    """TimeSeries(sim) (sim.jl:644-704): columnar view over the decimated device log. A column is [n_samples, N]; it is
    fetched from the device on first use, by flat name (`ts.theta`, `ts["theta"]`) or by the reference's path under the
    root `y` (`ts.aircraft.vehicle.kinematics.e_nb.θ`, lazily drilled like the reference's getproperty, sim.jl:666-676).
    `ts[k]` with an integer or slice restricts the time axis of every column (sim.jl:683-699)."""

    def __init__(self, sim, _sel=None):
        self._sim = sim
        ns, nf = C.c_int64(), C.c_int32()
        _ck(sim._lib.fb_log_info(sim._h, C.byref(ns), C.byref(nf)), sim._lib)
        self._sel = slice(None) if _sel is None else _sel
        t = np.zeros(ns.value)
        if ns.value:
            _ck(sim._lib.fb_log_times(sim._h, _dp(t)), sim._lib)
        self.t = t[self._sel]
        self.n_samples = int(np.size(self.t))
        self._n_all = ns.value
        self._cache = {}

    def _column(self, name):
        if name not in self._cache:
            out = np.zeros((self._n_all, self._sim.n))
            _ck(self._sim._lib.fb_log_fetch(self._sim._h, OUTPUT_NAMES[name], _dp(out)), self._sim._lib)
            self._cache[name] = out[self._sel]
        return self._cache[name]

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name in OUTPUT_NAMES:
            return self._column(name)
        return _YNode(self, "")._resolve(name)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.__getattr__(key)
        return TimeSeries(self._sim, _sel=key)

    def __len__(self):
        return self.n_samples


class LinearizedSS:
    """N LinearizedSS models (linearization.jl:24-111) stepped in lock-step: `Simulation(Model(linearize(aircraft, trim)))`
    of the reference's README example 3 for every member. Build it from a trimmed NED BatchSimulation (`LinearizedSS(sim)`,
    Jacobians stay on the device) or from arrays in the layout `BatchSimulation.linearize()` returns. State, input and
    output vectors are the reference's XStateSpace / UStateSpace / YStateSpace (c172s.jl:269-412, c172x.jl:332-498)."""

    def __init__(self, sim=None, model=None, n=None, dt=0.01, device=0):
        self._lib = load_library()
        self._h = C.c_void_p()
        if sim is not None:
            d = (C.c_int32 * 3)()
            _ck(self._lib.fb_linearize_dims(sim._h, d), self._lib)
            self.nx, self.nu, self.ny, self.n, dt, device = d[0], d[1], d[2], sim.n, sim.dt, sim.cfg.device
        else:
            self.nx, self.nu, self.ny = model["A"].shape[1], model["B"].shape[2], model["C"].shape[1]
            self.n = model["A"].shape[0] if n is None else n
        self.dt = dt
        _ck(self._lib.fb_lss_create(int(device), self.n, self.nx, self.nu, self.ny, float(dt), C.byref(self._h)), self._lib)
        if sim is not None:
            _ck(self._lib.fb_lss_from_linearize(self._h, sim._h), self._lib)
        else:
            raw = [np.ascontiguousarray(np.transpose(model[k], (2, 1, 0))) for k in ("A", "B", "C", "D")]   # [N][row][col] -> [col][row][N]
            raw += [np.ascontiguousarray(np.asarray(model[k]).T) for k in ("x0", "u0", "y0", "xdot0")]
            _ck(self._lib.fb_lss_set_model(self._h, *[_dp(a) for a in raw]), self._lib)

    def _get(self, fn, rows):
        out = np.empty((rows, self.n))
        _ck(fn(self._h, _dp(out)), self._lib)
        return out

    @property
    def x(self):
        return self._get(self._lib.fb_lss_get_state, self.nx)

    @x.setter
    def x(self, v):
        _ck(self._lib.fb_lss_set_state(self._h, _dp(np.ascontiguousarray(v, dtype=np.float64))), self._lib)

    @property
    def u(self):
        return self._get(self._lib.fb_lss_get_input, self.nu)

    @u.setter
    def u(self, v):
        _ck(self._lib.fb_lss_set_input(self._h, _dp(np.ascontiguousarray(v, dtype=np.float64))), self._lib)

    @property
    def y(self):
        return self._get(self._lib.fb_lss_get_output, self.ny)

    def step(self, t_advance=None):
        _ck(self._lib.fb_lss_step(self._h, float(self.dt if t_advance is None else t_advance)), self._lib)

    def sync(self):
        _ck(self._lib.fb_lss_sync(self._h), self._lib)

    def stream_handle(self):
        p = C.c_void_p()
        _ck(self._lib.fb_lss_stream(self._h, C.byref(p)), self._lib)
        return p.value

    def close(self):
        if self._h:
            self._lib.fb_lss_destroy(self._h)
            self._h = C.c_void_p()


def measure_fp64_peak(device=0):
    """Measured FP64 FMA throughput (TFLOP/s) of `device`: the roofline denominator of this path."""
    lib = load_library()
    v = C.c_double()
    _ck(lib.fb_measure_fp64_peak(int(device), C.byref(v)), lib)
    return v.value


# reference-style free functions (Julia's init!, step!, run!)
def init_b(sim, *a, **kw):
    return sim.init(*a, **kw)


def step_b(sim, *a):
    return sim.step(*a)


def run_b(sim):
    return sim.run()
