"""CPU tier: the oracle's restatement of the reference's ROS-side functions (pre-step, state assembly, post-step,
REF_V schedule, plan cut-off and down-sampling) against the reference's OWN, unmodified sources --
mpc_ros/src/driving_state.cpp and mpc_ros/src/mpc_planner_ros.cpp compiled behind the stand-in ROS headers of
oracle/shim/ros_stubs into oracle/_ref/libros_ref.so (oracle/ros_ref_driver.cpp, recipe oracle/Makefile).
SURVEY 8a rows "waypoint -> robot frame" .. "result post-step", 8f-1, 8f-2."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from bench import gen_py
from oracle.oracle_py import CFG_DEFAULT, YAML_DEFAULT, Reference, RosReference, ros_ref_available
from tests.conftest import ROOT

pytestmark = pytest.mark.skipif(not ros_ref_available(), reason="oracle/_ref/libros_ref.so not built (needs /root/reference)")
REF = "/root/reference/mpc_ros"


@pytest.fixture(scope="module")
def ros():
    return RosReference()


def test_polyfit_and_polyeval_match_reference(oracle, ros):
    rng = np.random.default_rng(5)
    for order in (1, 3, 5):
        for _ in range(20):
            m = int(rng.integers(order + 1, 30))
            xs = np.sort(rng.uniform(-1.0, 5.0, m)); ys = rng.normal(size=m)
            c_ref = ros.polyfit(xs, ys, order)            # driving_state.cpp:283-300 through the stand-in householderQr
            c = oracle.polyfit(xs, ys, order)
            assert np.abs(c - c_ref).max() <= 1e-9 * max(1.0, np.abs(c_ref).max())
            x = float(rng.uniform(-2, 2))
            assert abs(ros.polyeval(c_ref, x) - np.polyval(c_ref[::-1], x)) <= 1e-9 * max(1.0, abs(np.polyval(c_ref[::-1], x)))
            assert ros.polyeval(c_ref, 0.0) == c_ref[0]   # cte = polyeval(coeffs, 0) = c[0] (:211)


@pytest.mark.parametrize("delay_mode", [False, True])
def test_tracking_tick_matches_oracle_chain(oracle, ros, delay_mode):
    """Tracking::mpcComputeVelocityCommands (driving_state.cpp:105-119: deceleration, findBestPath = transform, polyfit,
    etheta rule, state assembly, MPC::Solve, speed clamp) == the oracle's prestep -> state -> solve -> poststep."""
    pm = dict(YAML_DEFAULT)
    g = gen_py.problems(20261018 + 5, 48)
    h = ros.tracker(pm, delay_mode, 0.5)
    R = Reference(pm)
    rng = np.random.default_rng(9)
    for i in range(48):
        wx, wy = g["wx"][:, i], g["wy"][:, i]
        pose = g["pose"][:, i]; v = g["vel"][0, i]
        w_prev, thr_prev = rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5)
        ref_v = 0.5
        far_goal = (pose[0] + 100.0, pose[1])                        # no braking
        t = ros.tick(h, pose, far_goal, v, wx, wy, [w_prev, thr_prev, ref_v])
        assert t["ok"] == 1
        c, cte, eth = oracle.prestep(wx, wy, *pose)
        s6 = oracle.state(delay_mode, v, w_prev, thr_prev, pm["DT"], cte, eth)
        r = R.solve(s6, c)                                           # the same reference MPC the tick used
        assert abs(t["w"] - r["u0"][0]) <= 1e-9 and abs(t["throttle"] - r["u0"][1]) <= 1e-9
        speed = oracle.poststep_speed(v, r["u0"][1], pm["DT"], ref_v)
        assert abs(t["cmd"][0] - speed) <= 1e-12 and t["cmd"][1] == t["w"]
        assert np.abs(t["pred"] - r["pred"]).max() <= 1e-9
        assert t["ref_v"] == ref_v


def test_deceleration_schedule_matches_reference(oracle, ros):
    pm = dict(YAML_DEFAULT)
    h = ros.tracker(pm, False, 0.7)
    g = gen_py.problems(20261018 + 6, 24)
    rng = np.random.default_rng(3)
    for i in range(24):
        pose = g["pose"][:, i]; v = float(rng.uniform(0.05, 0.8))
        dist = float(rng.uniform(0.0, 1.0)); ang = float(rng.uniform(-3, 3))
        goal = (pose[0] + dist * np.cos(ang), pose[1] + dist * np.sin(ang))
        ref_v0 = float(rng.choice([0.5, 0.05, 0.3]))
        t = ros.tick(h, pose, goal, v, g["wx"][:, i], g["wy"][:, i], [0.0, 0.0, ref_v0])
        want = oracle.decel(pose[0], pose[1], goal[0], goal[1], v, pm["MAXTHR"], 0.7, 0.05, ref_v0)
        assert t["ref_v"] == want, (i, t["ref_v"], want)


def test_cutoff_and_downsample_match_reference(oracle, ros):
    """MPCPlannerROS::getCutOffPlan (:266-291) and downSamplePlan (:365-391) on open plans cut from the three tracks."""
    rng = np.random.default_rng(11)
    for kind in range(3):
        tx, ty = gen_py.path(kind)
        n = len(tx)
        for _ in range(40):
            start = int(rng.integers(0, n)); length = int(rng.integers(110, 260))
            idx = (start + np.arange(length)) % n
            px, py = tx[idx], ty[idx]
            a = int(rng.integers(0, 40))                             # the robot is near plan point a
            rx = px[a] + rng.uniform(-0.3, 0.3); ry = py[a] + rng.uniform(-0.3, 0.3)
            r = ros.window(px, py, rx, ry, 5.0)
            e = oracle.cutoff(px, py, 0, rx, ry, ring=False)
            assert e == r["erased"], (kind, e, r["erased"])
            wd = float(np.hypot(px[e + 1] - px[e], py[e + 1] - py[e]))
            step = oracle.downsample_step(5.0, wd)                   # int(_pathLength / 10 / _waypointsDist), :374
            assert step == r["step"]
            wx, wy, m = oracle.downsample(px, py, e, length - e, step, ring=False, cap=256)
            assert m == r["m"]
            assert np.array_equal(wx, r["wx"]) and np.array_equal(wy, r["wy"])


def test_oracle_loop_equals_the_reference_code_loop():
    """BASELINE config 5's checker: the oracle loop (restatements + oracle solve) and the loop through the reference's own
    getCutOffPlan / downSamplePlan / Tracking tick / MPC::Solve command the same robots identically."""
    from tests.closed_loop_ref import run_oracle, run_reference_ros
    o = run_oracle(6, 30); r = run_reference_ros(6, 30)
    assert np.abs(o["w"] - r["w"]).max() <= 1e-9 and np.abs(o["thr"] - r["thr"]).max() <= 1e-9
    assert np.abs(o["dist"] - r["dist"]).max() <= 1e-9


def test_reference_ros_sources_compile_against_the_adapter_header():
    """SURVEY 8f-3 / INTEGRATION section 1: the reference's two ROS sources compile UNMODIFIED against this repository's
    mpc_planner.h (the MPC adapter) -- stand-in ROS / Eigen headers, real reference headers; the maintainer's swap is
    "replace mpc_planner.h / mpc_planner.cpp", so the reference's own headers are staged next to OUR mpc_planner.h
    (a quoted #include looks in the including file's directory first)."""
    if not os.path.isdir(REF):
        pytest.skip("reference checkout not present")
    with tempfile.TemporaryDirectory() as d:
        os.symlink(os.path.join(REF, "include", "driving_state.h"), os.path.join(d, "driving_state.h"))
        os.symlink(os.path.join(REF, "include", "mpc_planner_ros.h"), os.path.join(d, "mpc_planner_ros.h"))
        os.symlink(os.path.join(ROOT, "mpc_ros_b200", "include", "mpc_planner.h"), os.path.join(d, "mpc_planner.h"))
        inc = ["-I" + d, "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "oracle", "shim", "ros_stubs"),
               "-I" + os.path.join(ROOT, "include")]
        objs = []
        for src in ("driving_state.cpp", "mpc_planner_ros.cpp"):
            o = os.path.join(d, src + ".o")
            r = subprocess.run(["g++", "-std=c++14", "-O1", "-fPIC", "-w", "-c", os.path.join(REF, "src", src), "-o", o] + inc,
                               capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-3000:]
            objs.append(o)
        # ... and link with the adapter into the plugin library the catkin build would produce
        libdir = os.path.join(ROOT, "mpc_ros_b200", "lib")
        r = subprocess.run(["g++", "-shared", "-o", os.path.join(d, "libmpc_ros.so")] + objs +
                           ["-L" + libdir, "-lmpc_adapter", "-lmpc_b200", "-Wl,--no-undefined", "-Wl,-rpath," + libdir],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        syms = subprocess.run(["nm", "-DC", "--defined-only", os.path.join(d, "libmpc_ros.so")], capture_output=True, text=True).stdout
        assert "mpc_ros::MPCPlannerROS::computeVelocityCommands" in syms
        assert "Tracking::findBestPath" in syms
