"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports every symbol the header declares,
computing calls fail loudly without a GPU (there is no CPU fallback), workload generation is sharding-invariant, and the
N > 1 path (member sharding + statistics all-reduce) works on a world_size-2 gloo group."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "flightbatch.h")).read()
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(fb):
    lib = fb.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(lib, name), f"libflightbatch.so does not export {name}"
    assert sorted(fb.EXPORTS) == declared
    assert lib.fb_api_version() == fb.DEFS["FB_API_VERSION"]


def test_header_enums_are_consistent(fb):
    assert fb.U["FB_U_AILERON"] == 0 and fb.U["FB_U_THROTTLE"] == 9 and fb.U["FB_U_ACT_THROTTLE"] == 27
    assert len(set(fb.U.values())) == len(fb.U) and len(set(fb.Y.values())) == len(fb.Y)
    assert max(v for k, v in fb.U.items() if not k.endswith("_COUNT")) < fb.U["FB_U__COUNT"]


def test_no_gpu_means_loud_failure_not_a_fallback(fb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fb.FlightBatchError, match="no CUDA device"):
        fb.BatchSimulation(n=4)
    with pytest.raises(fb.FlightBatchError):
        fb.measure_fp64_peak(0)


def test_api_misuse_is_reported(fb):
    lib = fb.load_library()
    cfg = fb.fb_config()
    cfg.variant = 99
    h = C.c_void_p()
    assert lib.fb_create(C.byref(cfg), C.byref(h)) == fb.DEFS["FB_ERR_ARG"]
    assert b"variant" in lib.fb_last_error()
    assert lib.fb_step(None, 1.0) == fb.DEFS["FB_ERR_ARG"]


def test_table_blob_has_every_section():
    import struct
    blob = open(os.path.join(ROOT, "flight.jl_b200", "data", "c172_tables.bin"), "rb").read()
    assert blob[:8] == b"FBTABLE1"
    n = struct.unpack_from("<I", blob, 8)[0]
    names = {struct.unpack_from("<24s", blob, 16 + 64 * i)[0].rstrip(b"\0").decode() for i in range(n)}
    for need in ("aero.CD_a_df", "aero.CL_a", "eng.pi_std", "eng.sfc_pow", "prop.tables", "geoid", "gain.te2te", "gain.phibeta2ar", "gain.q2e", "gain.chi2phi"):
        assert need in names


def test_workload_is_sharding_invariant():
    import bench
    a = bench.member_params(0, 12288)
    b = bench.member_params(4096, 8192)
    c = bench.member_params(5000, 100)
    for k in a:
        assert np.array_equal(a[k][4096:], b[k])
        assert np.array_equal(a[k][5000:5100], c[k])
    assert a["EAS"].min() >= 30 and a["EAS"].max() <= 55
    assert (a["flaps"][a["EAS"] < 35] == 1.0).all() and (a["flaps"][a["EAS"] > 42] == 0.0).all()


def test_geoid_height_of_the_workload_generator_matches_the_oracle(orc, fb):
    import bench
    rng = np.random.default_rng(3)
    for _ in range(20):
        lat, lon = np.deg2rad(rng.uniform(-60, 60)), rng.uniform(-np.pi, np.pi)
        o = orc.Sim(0, 0, 0.01)
        o.init_trim(lat=lat, lon=lon, h_e=1000.0)
        n_oracle = o.y(fb.Y["FB_Y_H_E"]) - o.y(fb.Y["FB_Y_H_O"])
        assert abs(bench.geoid_height(lat, lon) - n_oracle) < 1e-9


def test_doublet_schedule_shape():
    import bench
    n = 16
    m, s, f, v = bench.doublet_events(n, np.zeros(n), np.ones(n), 1, 0)
    assert len(m) == 6 * n and set(s) == {500, 700, 900, 2000, 2200, 2400}
    assert np.allclose(sorted(set(v[f == 1])), [-0.1, 0.0, 0.1]) and np.allclose(sorted(set(v[f == 0])), [0.9, 1.0, 1.1])


def test_shard_range_partitions_exactly():
    import importlib
    import flightbatch  # noqa: F401
    dist = importlib.import_module("flightbatch.dist")
    for n in (1, 7, 65536, 1_000_003):
        for w in (1, 2, 4, 8):
            pieces = [dist.shard_range(n, r, w) for r in range(w)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(pieces[i][1] == pieces[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in pieces) - min(hi - lo for lo, hi in pieces) <= 1


def _gloo_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import flightbatch  # noqa: F401
    from flightbatch import dist as fbd
    import bench
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n_total = 10000
    lo, hi = fbd.shard_range(n_total, rank, world)
    p = bench.member_params(lo, hi - lo)                  # each rank generates only its shard
    status = np.zeros(hi - lo, dtype=np.int32)
    status[::97] = 6                                      # a few failed members are excluded, as fb_stats does
    g = fbd.allreduce_moments(fbd.host_moments(p["EAS"], status), device="cpu")
    q.put((rank, g))
    dist.destroy_process_group()


def test_two_rank_gloo_statistics_allreduce():
    import torch.multiprocessing as mp
    import bench
    from flightbatch import dist as fbd
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = bench.member_params(0, 10000)["EAS"]
    keep = np.ones(10000, dtype=bool)
    for r in range(2):
        lo, hi = fbd.shard_range(10000, r, 2)
        keep[lo:hi][::97] = False
    ref = fbd.host_moments(full[keep])
    for r in range(2):
        g = res[r]
        assert g["count"] == ref["count"] and g["min"] == ref["min"] and g["max"] == ref["max"]
        assert abs(g["sum"] - ref["sum"]) < 1e-9 * abs(ref["sum"]) and abs(g["mean"] - full[keep].mean()) < 1e-12 * 50


# ------------------------------------------------------------------------------------------------ the boundary from outside Python
def build_abi_driver():
    """tests/abi_driver.c compiled with gcc against include/flightbatch.h and linked to libflightbatch.so (prototypes checked by
    the compiler, -Werror)."""
    import subprocess
    import tempfile
    libdir = os.path.join(ROOT, "flight.jl_b200")
    exe = os.path.join(tempfile.gettempdir(), "fb_abi_driver_%d" % os.getpid())
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "abi_driver.c"), "-L", libdir, "-lflightbatch", "-lm", f"-Wl,-rpath,{libdir}"])
    return exe


JL_TYPES = {"Int32": (4, 4), "Int64": (8, 8), "Float64": (8, 8), "UInt8": (1, 1), "NTuple{32,Int32}": (128, 4)}


def _julia_struct_layout(name):
    """Field offsets of an isbits `struct` of BatchSimulation.jl under Julia's C-compatible layout rules (natural alignment)."""
    txt = open(os.path.join(ROOT, "flight.jl_b200", "julia", "BatchSimulation.jl")).read()
    body = re.search(r"^struct %s\n(.*?)^end" % name, txt, re.S | re.M).group(1)
    off, out, amax = 0, {}, 1
    for line in body.splitlines():
        line = line.split("#")[0].strip()
        if not line:
            continue
        field, typ = [t.strip() for t in line.split("::")]
        size, align = JL_TYPES[typ]
        off = (off + align - 1) // align * align
        out[field] = off
        off += size
        amax = max(amax, align)
    out["size"] = (off + amax - 1) // amax * amax
    return out


def test_struct_layouts_agree_between_c_ctypes_and_julia(fb):
    import json
    import subprocess
    c = json.loads(subprocess.check_output([build_abi_driver(), "layout"]))
    assert c["api_version"] == fb.DEFS["FB_API_VERSION"] and c["FB_U__COUNT"] == fb.U["FB_U__COUNT"] and c["FB_Y__COUNT"] == fb.Y["FB_Y__COUNT"]
    # ctypes mirror (what every Python test goes through)
    for sname, cls in (("fb_config", fb.fb_config), ("fb_moments", fb.fb_moments), ("fb_trim_params_soa", fb.fb_trim_params_soa), ("fb_init_soa", fb.fb_init_soa)):
        assert C.sizeof(cls) == c[sname]["size"], sname
        for f, o in c[sname].items():
            if f != "size":
                assert getattr(cls, f).offset == o, (sname, f)
    # Julia mirror (never executed in this image: checked mechanically instead)
    jl = _julia_struct_layout("FbConfig")
    assert jl == {k: v for k, v in c["fb_config"].items()}, (jl, c["fb_config"])
    assert _julia_struct_layout("Moments") == c["fb_moments"]
    txt = open(os.path.join(ROOT, "flight.jl_b200", "julia", "BatchSimulation.jl")).read()
    assert int(re.search(r"const FB_API_VERSION = (\d+)", txt).group(1)) == c["api_version"]
    assert len(re.search(r"const TRIM_FIELDS = \((.*?)\)", txt, re.S).group(1).split(",")) * 8 == c["fb_trim_params_soa"]["size"]


def test_every_entry_point_rejects_null_handles_from_c():
    import subprocess
    r = subprocess.run([build_abi_driver(), "null"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def _split_top(s):
    """Split at commas that are not nested in (), {} or []."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [t.strip() for t in out if t.strip()]


def test_julia_ccalls_match_the_header_prototypes():
    """Every `ccall((:fb_x, libflightbatch), ret, (argtypes...), args...)` of BatchSimulation.jl: the function exists in the header,
    the number of argument types equals the prototype's parameter count and the number of values passed, and the return type is
    the prototype's. (Round 1 shipped an 8-argument call of the 9-parameter fb_linearize.)"""
    hdr = open(os.path.join(ROOT, "include", "flightbatch.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"^(int32_t|const char\*)\s+(fb_[a-z0-9_]+)\s*\((.*?)\);", hdr, re.S | re.M):
        a = [t for t in _split_top(" ".join(args.split())) if t != "void"]
        protos[name] = (ret, len(a))
    txt = open(os.path.join(ROOT, "flight.jl_b200", "julia", "BatchSimulation.jl")).read()
    calls = 0
    for m in re.finditer(r"ccall\(\((:?\w+), libflightbatch\)", txt):
        assert m.group(1).startswith(":"), f"ccall needs a literal function name, got `{m.group(1)}`"
        name = m.group(1)[1:]
        # the whole argument list of this ccall( ... ) by bracket matching
        i = m.start() + len("ccall(")
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(txt[j], 0)
            j += 1
        parts = _split_top(txt[i:j - 1])
        ret, argtypes, values = parts[1], parts[2], parts[3:]
        assert name in protos, f"{name} is not declared in flightbatch.h"
        ntypes = len(_split_top(argtypes[1:-1]))
        assert ntypes == protos[name][1], f"{name}: {ntypes} argument types, the prototype has {protos[name][1]}"
        assert len(values) == ntypes, f"{name}: {len(values)} values for {ntypes} argument types"
        assert (ret == "Cstring") == (protos[name][0] == "const char*") and (ret == "Int32") == (protos[name][0] == "int32_t"), (name, ret)
        calls += 1
    assert calls >= 40
    # fields the binding reads must exist on the struct it reads them from
    sim_fields = set(re.findall(r"^\s+(\w+)::", re.search(r"mutable struct BatchSimulation\n(.*?)function BatchSimulation", txt, re.S).group(1), re.M))
    code = "\n".join(line.split("#")[0] for line in txt.splitlines())      # comments and docstring references do not count
    code = re.sub(r'"""(.*?)"""', "", code, flags=re.S)
    code = re.sub(r'^"[^"\n]*"$', "", code, flags=re.M)                     # one-line docstrings
    for f in set(re.findall(r"\bsim\.(\w+)", code)):
        assert f in sim_fields | {"t", "x"}, f"BatchSimulation has no field {f}"


def test_julia_enums_follow_the_header(fb):
    """The Julia mirror's `Variant` / `KinematicsKind` enums and the input / output ids it names must be the header's values
    (the mirror is never executed in this image)."""
    txt = open(os.path.join(ROOT, "flight.jl_b200", "julia", "BatchSimulation.jl")).read()
    var = dict(re.findall(r"(C172\w+) = (\d+)", re.search(r"@enum Variant::Int32 ([^\n#]*)", txt).group(1)))
    assert {k: int(v) for k, v in var.items()} == {"C172S": fb.DEFS["FB_C172S"], "C172X0": fb.DEFS["FB_C172X0"], "C172X1": fb.DEFS["FB_C172X1"],
                                                   "C172X2": fb.DEFS["FB_C172X2"], "C172XA2": fb.DEFS["FB_C172XA2"]}
    kin = dict(re.findall(r"(\w+) = (\d+)", re.search(r"@enum KinematicsKind::Int32 ([^\n#]*)", txt).group(1)))
    assert {k: int(v) for k, v in kin.items()} == {"WA": fb.DEFS["FB_WA"], "ECEF": fb.DEFS["FB_ECEF"], "NED": fb.DEFS["FB_NED"]}
    for name, val in re.findall(r"const (FB_[UY]_\w+) = Int32\((\d+)\)", txt):
        table = fb.U if name.startswith("FB_U_") else fb.Y
        assert table[name] == int(val), name


def test_table_blob_holds_the_variable_pitch_sections():
    """The FB_C172XA2 variant's VariablePitch Lookup travels in the same blob (prop.tables3 [21][21][n][6] + prop.pitch3)."""
    import struct
    b = open(os.path.join(ROOT, "flight.jl_b200", "data", "c172_tables.bin"), "rb").read()
    n = struct.unpack_from("<I", b, 8)[0]
    secs = {}
    for i in range(n):
        nm, dt, nd, d0, d1, d2, d3, off, nb = struct.unpack_from("<24sII4IQQ", b, 16 + 64 * i)
        secs[nm.rstrip(b"\0").decode()] = ([d0, d1, d2, d3][:nd], nb)
    assert secs["prop.tables3"][0][:2] == [21, 21] and secs["prop.tables3"][0][3] == 6
    assert secs["prop.tables3"][1] == 21 * 21 * secs["prop.tables3"][0][2] * 6 * 8 and secs["prop.pitch3"][1] == 16
