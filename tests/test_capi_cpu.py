"""CPU tier: the C-ABI library loads, exports every symbol include/mpc_b200.h declares, the host-side
parameter logic mirrors the reference, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

from mpc_ros_b200 import capi
from tests.conftest import ROOT, has_gpu


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "mpc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mpc_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), "libmpc_b200.so does not export " + name
    assert sorted(capi.EXPORTS) == declared
    assert L.mpc_b200_version() == 100


def test_param_defaults_mirror_reference():
    p = capi.default_params()
    # MPC::MPC (mpc_planner.cpp:227-230) and FG_eval::FG_eval (:47-57)
    assert (p.mpc_steps, p.max_angvel, p.max_throttle, p.bound_value) == (20, 3.0, 1.0, 1000.0)
    assert (p.dt, p.ref_vel, p.w_cte, p.w_etheta, p.w_vel, p.w_angvel, p.w_accel) == (0.1, 0.5, 100.0, 100.0, 1.0, 100.0, 50.0)
    y = capi.yaml_default_params()
    # mpc_ros/params/mpc_params.yaml:12-25
    assert (y.mpc_steps, y.ref_vel, y.w_cte, y.w_etheta, y.w_vel, y.w_angvel, y.w_accel) == (20, 0.5, 100.0, 0.0, 1000.0, 100.0, 50.0)
    assert (y.max_angvel, y.max_throttle, y.bound_value, y.dt) == (1.5, 1.0, 1000.0, 0.1)


def test_loadparams_keys():
    p = capi.default_params()
    pm = dict(DT=0.05, STEPS=30, REF_CTE=0.1, REF_ETHETA=0.2, REF_V=0.9, W_CTE=1, W_EPSI=2, W_V=3, W_ANGVEL=4, W_A=5,
              W_DANGVEL=6, W_DA=7, ANGVEL=0.7, MAXTHR=0.8, BOUND=99)
    capi.params_from_map(pm, p)
    assert (p.dt, p.mpc_steps, p.ref_cte, p.ref_etheta, p.ref_vel) == (0.05, 30, 0.1, 0.2, 0.9)
    assert (p.w_cte, p.w_etheta, p.w_vel, p.w_angvel, p.w_accel, p.w_angvel_d, p.w_accel_d) == (1, 2, 3, 4, 5, 6, 7)
    assert (p.max_angvel, p.max_throttle, p.bound_value) == (0.7, 0.8, 99)
    # missing key keeps the previous value (mpc_planner.cpp:73: find() != end() ? at() : old)
    capi.params_from_map(dict(REF_V=0.3), p)
    assert p.ref_vel == 0.3 and p.w_cte == 1
    assert capi.lib().mpc_b200_params_set(C.byref(p), b"NOPE", 1.0) == -1


def test_yaml_loader(tmp_path):
    y = tmp_path / "mpc_params.yaml"
    y.write_text("# Parameters for control loop\npub_twist_cmd: true\ndebug_info: false\ndelay_mode: true\n"
                 "max_speed: 0.5 # unit: m/s \nwaypoints_dist: -1.0\npath_length: 5.0 # unit: m\ngoal_radius: 0.5\n"
                 "controller_freq: 10\n\n# Parameter for MPC solver\nmpc_steps: 20.0\nmpc_ref_cte: 0.0\n"
                 "mpc_ref_vel: 0.5\nmpc_ref_etheta: 0.0\nmpc_w_cte: 100.0\nmpc_w_etheta: 0000.0\nmpc_w_vel: 1000.0\n"
                 "mpc_w_angvel: 100.0\nmpc_w_angvel_d: 0.0\nmpc_w_accel: 50.0\nmpc_w_accel_d: 0.0\n"
                 "mpc_max_angvel: 1.5 \nmpc_max_throttle: 1.0 # Maximal throttle accel\nmpc_bound_value: 1.0e3 # Bound\n")
    p = capi.params_from_yaml(str(y))
    q = capi.yaml_default_params()
    for f, _ in capi.Params._fields_:
        assert getattr(p, f) == getattr(q, f), f
    y.write_text("mpc_max_throttle: 0.01\ncontroller_freq: 20\n")
    p = capi.params_from_yaml(str(y))
    assert p.max_throttle == 0.1          # floor of driving_state.cpp:63
    assert p.dt == 0.05
    with pytest.raises(capi.MpcError):
        capi.params_from_yaml(str(tmp_path / "missing.yaml"))


def test_create_validates_and_fails_loudly_without_gpu():
    L = capi.lib()
    p = capi.yaml_default_params()
    h = C.c_void_p()
    assert L.mpc_b200_create(C.byref(h), C.byref(p), 0, 0) == -1            # max_batch < 1
    bad = capi.yaml_default_params(); bad.mpc_steps = 1
    assert L.mpc_b200_create(C.byref(h), C.byref(bad), 8, 0) == -1
    neg = capi.yaml_default_params(); neg.w_accel_d = -1.0
    assert L.mpc_b200_create(C.byref(h), C.byref(neg), 8, 0) == -1          # negative rate weight
    if not has_gpu():
        assert L.mpc_b200_device_count() == 0
        assert L.mpc_b200_create(C.byref(h), C.byref(p), 8, 0) == -2        # no CPU fallback
        assert not h.value
        with pytest.raises(capi.MpcError):
            capi.Solver(p, 8)
    assert L.mpc_b200_warm_size(20) == 158 + 120 + 76
    assert b"no CPU path" in L.mpc_b200_strerror(-2)


def test_null_handle_is_rejected_everywhere():
    """Every batched entry point validates its handle before touching CUDA (returns MPC_B200_ERR_INVALID)."""
    L = capi.lib()
    z = None
    assert L.mpc_b200_solve_batch(z, 1, z, z, z, z, z, z, z, z, z, z, z, z) == -1
    assert L.mpc_b200_track_batch(z, 1, 11, z, z, z, z, z, z, z, z, z, z, z, z) == -1
    assert L.mpc_b200_track_submit(z, 1, 11, z, z, z, z, z, z, z, z, z, z, z, z) == -1
    assert L.mpc_b200_track_wait(z) == -1
    assert L.mpc_b200_decel_batch(z, 1, z, z, z, 0.05, z, z) == -1
    assert L.mpc_b200_plant_step_batch(z, 1, z, z, z, z) == -1
    assert L.mpc_b200_poststep_batch(z, 1, z, z, z, z, z) == -1
    assert L.mpc_b200_warm_shift(z, 1, z, z, z) == -1
    assert L.mpc_b200_stream_create(0, None) == -1
    assert L.mpc_b200_stream_destroy(None) == -1
    if not has_gpu():
        s = C.c_void_p()
        assert L.mpc_b200_stream_create(0, C.byref(s)) == -2               # no device, no stream
    p = capi.yaml_default_params()
    assert L.mpc_b200_num_waypoints(C.byref(p)) == 11                      # 5 m window, every 10th of 100 points + the last


def test_product_library_has_no_oracle_or_emulator_symbols():
    """The product .so must not contain the oracle / emulator (no CPU fallback inside)."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    for bad in ("ipm_solve", "mpc_oracle", "nmpc_emu", "ldl_factor"):
        assert bad not in out


def test_bench_issue_loop_library_loads():
    """bench.py's device-resident leg is issued by bench/libmpc_issue.so (C++ over the C ABI): it loads without a GPU and
    exports its four entry points; it links the product library, not a copy of it."""
    path = os.path.join(ROOT, "bench", "libmpc_issue.so")
    if not os.path.exists(path):
        pytest.skip("bench/libmpc_issue.so not built (run __graft_entry__.build())")
    L = C.CDLL(path)
    for name in ("mpcb_issue_create", "mpcb_issue_run", "mpcb_issue_samples", "mpcb_issue_destroy"):
        assert hasattr(L, name)
    import subprocess
    deps = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert "libmpc_b200.so" in deps
