"""The C ABI driven from plain C (tests/c_abi/drive.c, gcc, `#include "tnuts.h"`): proof that the
boundary works without Python's struct mirrors.  CPU: it compiles and links.  GPU: it runs
create -> set_model -> init -> run -> fetch -> get_state -> set_state -> destroy and its output is
compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
import turing_jl_b200 as T

from helpers import oracle_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "drive")
    libdir = os.path.join(ROOT, "turing.jl_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "drive.c"), "-o", exe,
                           "-L", libdir, "-ltnuts", "-lm", "-Wl,-rpath," + libdir])
    return exe


def test_c_driver_compiles_and_links(built, tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_c_driver_matches_the_oracle(built, tmp_path):
    exe = _build(tmp_path)
    out = str(tmp_path / "out.bin")
    r = subprocess.run([exe, out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "resume_identical 1" in r.stdout
    with open(out, "rb") as f:
        K, ncol, C, D = np.fromfile(f, dtype=np.int64, count=4)
        draws = np.fromfile(f, dtype=np.float64, count=K * ncol * C).reshape(K, ncol, C)
        theta = np.fromfile(f, dtype=np.float64, count=K * D * C).reshape(K, D, C)
        eps = np.fromfile(f, dtype=np.float64, count=C)
        lp0 = np.fromfile(f, dtype=np.float64, count=1)[0]
    model = T.EightSchools()
    om = oracle_of(model)
    assert lp0 == pytest.approx(om.lpgrad(np.zeros(10))[0], rel=1e-13)
    cfg = O.make_config(delta=0.8, n_adapts=20, seed=11, sampler_mode=1)
    ref = O.sample_chains(om, cfg, int(C), int(K))
    st = draws[:, D + 3:, :].transpose(2, 0, 1)
    same = (st[..., 0] == ref["stats"][..., 0]) & (st[..., 6] == ref["stats"][..., 6])
    assert np.all(same[:, :6]) and same.mean() > 0.97
    np.testing.assert_allclose(theta.transpose(2, 0, 1)[:, :6], ref["theta"][:, :6], atol=1e-8)
    np.testing.assert_allclose(draws[:6, :D, :].transpose(2, 0, 1), ref["params"][:, :6], atol=1e-7)
