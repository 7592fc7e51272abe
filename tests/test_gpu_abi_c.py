"""The hot path driven from plain C through include/flightbatch.h (tests/abi_driver.c): create -> load tables -> trim ->
schedule -> step -> read back -> statistics -> linearize -> LinearizedSS -> destroy on cuda:0 — the calls a Julia `ccall` host
makes, with no Python between the caller and the library."""
import os
import subprocess

import pytest

from test_abi_and_host import ROOT, build_abi_driver

pytestmark = pytest.mark.gpu


def test_c_host_drives_the_whole_path():
    r = subprocess.run([build_abi_driver(), "run", os.path.join(ROOT, "flight.jl_b200", "data", "c172_tables.bin")], capture_output=True, text=True)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
