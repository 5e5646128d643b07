"""The C++ `MPC` adapter (mpc_ros_b200/include/mpc_planner.h): the class a maintainer drops in place of the
reference's mpc_planner.{h,cpp}.  CPU tier: it is built and exports the reference's class surface.  GPU
tier: BASELINE config 1 -- one solve through MPC::LoadParams / MPC::Solve (ROS-free harness
bench/mpc_bench.cpp), checked against the oracle, plus the single-solve latency target."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import ROOT

LIBDIR = os.path.join(ROOT, "mpc_ros_b200", "lib")


def test_adapter_exports_reference_class_surface():
    out = subprocess.run(["nm", "-DC", "--defined-only", os.path.join(LIBDIR, "libmpc_adapter.so")],
                         capture_output=True, text=True).stdout
    # mpc_ros/include/mpc_planner.h:26-47
    assert "MPC::MPC()" in out
    assert "MPC::Solve(Eigen::VectorXd, Eigen::VectorXd)" in out
    assert "MPC::LoadParams(std::map<" in out
    assert os.access(os.path.join(LIBDIR, "mpc_bench"), os.X_OK)


@pytest.mark.gpu
def test_config1_single_solve_through_adapter(oracle):
    from bench import gen_py
    from oracle.oracle_py import YAML_DEFAULT
    r = subprocess.run([os.path.join(LIBDIR, "mpc_bench"), "latency", "2000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "init mpc" in r.stdout                      # the reference prints this on construction (mpc_planner.cpp:225)
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    g = gen_py.problems(20261019, 1)                   # config 1 = problem #0 of seed 20261019
    c, cte, eth = oracle.prestep(g["wx"][:, 0], g["wy"][:, 0], *g["pose"][:, 0])
    o = oracle.solve(YAML_DEFAULT, [0.0, 0.0, 0.0, g["vel"][0, 0], cte, eth], c)
    assert o["status"] == 1 and res["status"] == 1
    assert abs(res["w0"] - o["u0"][0]) <= 1e-5 and abs(res["a0"] - o["u0"][1]) <= 1e-5
    assert res["kkt"] <= 1e-8
    assert res["p99_us"] < 1000.0                      # north-star: p99 single-solve latency < 1 ms


@pytest.mark.gpu
def test_higher_order_polynomial_through_adapter(oracle):
    """MPC::Solve with coeffs.size() == 6 (a quintic): the adapter switches the handle to the higher-order path for
    that call and back; lower orders are zero-padded."""
    from oracle.oracle_py import YAML_DEFAULT
    co = [0.05, -0.1, 0.02, 0.003, 0.04, -0.03]
    r = subprocess.run([os.path.join(LIBDIR, "mpc_bench"), "poly"] + ["%.17g" % c for c in co], capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stderr
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    state = [0.0, 0.0, 0.0, 0.3, co[0], -0.1]
    o = oracle.solve(YAML_DEFAULT, state, co)
    o3 = oracle.solve(YAML_DEFAULT, state, co[:4])
    assert res["status"] == 1 and o["status"] == 1 and res["status_cubic"] == 1
    assert abs(res["w0"] - o["u0"][0]) <= 1e-5 and abs(res["a0"] - o["u0"][1]) <= 1e-5
    assert abs(res["w0_cubic"] - o3["u0"][0]) <= 1e-5 and abs(res["a0_cubic"] - o3["u0"][1]) <= 1e-5
    assert abs(o["u0"][0] - o3["u0"][0]) > 1e-4          # (the two problems differ)
    # a quadratic: zero-padded
    r = subprocess.run([os.path.join(LIBDIR, "mpc_bench"), "poly", "0.05", "-0.1", "0.02"], capture_output=True, text=True, timeout=120)
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    o2 = oracle.solve(YAML_DEFAULT, [0.0, 0.0, 0.0, 0.3, 0.05, -0.1], [0.05, -0.1, 0.02, 0.0])
    assert res["status"] == 1 and abs(res["w0"] - o2["u0"][0]) <= 1e-5
