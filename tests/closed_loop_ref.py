"""tests/closed_loop_ref.py -- TEST INFRASTRUCTURE: the reference control loop on the CPU, the checker for BASELINE
config 5 (bench/closed_loop.py drives the GPU loops).

run_oracle         cold-started oracle solve every tick (the reference cold-starts every call), pre-step / state assembly
                   / post-step through the oracle's restatements, plan windowing through mpc_oracle_cutoff / _downsample.
run_reference_ros  the same loop through the reference's OWN code: MPCPlannerROS::getCutOffPlan + downSamplePlan and
                   Tracking::mpcComputeVelocityCommands (deceleration, findBestPath, MPC::Solve) from oracle/_ref/libros_ref.so.
Same Fleet (robots, tracks, plant) as the GPU loops."""
import numpy as np

from bench.closed_loop import Fleet


def _window(orc, fleet, i, max_erase=64):
    px, py = fleet.paths[fleet.kind[i]]
    e = orc.cutoff(px, py, int(fleet.idx[i]), fleet.pose[0, i], fleet.pose[1, i], ring=True, max_erase=max_erase)
    fleet.idx[i] = (int(fleet.idx[i]) + e) % len(px)
    wx, wy, m = orc.downsample(px, py, int(fleet.idx[i]), fleet.win, fleet.step, ring=True)
    return wx, wy


def run_oracle(R, T, seed=20261018 + 5, pm=None, delay_mode=True):
    from oracle.oracle_py import Oracle, YAML_DEFAULT
    pm = dict(pm or YAML_DEFAULT)
    orc = Oracle()
    dt = pm["DT"]
    fleet = Fleet(R, seed)
    trace = dict(cte=[], eth=[], w=[], thr=[], iters=[], dist=[], status=[])
    for t in range(T):
        w = np.zeros(R); thr = np.zeros(R); cte = np.zeros(R); eth = np.zeros(R); its = np.zeros(R); stt = np.zeros(R)
        for i in range(R):
            wx, wy = _window(orc, fleet, i)
            c, ct, e = orc.prestep(wx, wy, *fleet.pose[:, i])
            st = orc.state(delay_mode, fleet.v[i], fleet.w[i], fleet.thr[i], dt, ct, e)
            r = orc.solve(pm, st, c)
            w[i], thr[i] = r["u0"]; cte[i] = ct; eth[i] = e; its[i] = r["iters"]; stt[i] = r["status"]
        trace["cte"].append(cte); trace["eth"].append(eth); trace["w"].append(w.copy()); trace["thr"].append(thr.copy())
        trace["iters"].append(its); trace["status"].append(stt); trace["dist"].append(fleet.track_distance())
        fleet.actuate(w, thr, dt, pm["REF_V"])
    out = {k: np.array(v) for k, v in trace.items()}
    out["kind"] = fleet.kind.copy()
    return out


def run_reference_ros(R, T, seed=20261018 + 5, pm=None, delay_mode=True):
    """Every robot gets its own DrivingStateContext (it keeps _w, _throttle and REF_V between ticks, as the node does)."""
    from oracle.oracle_py import RosReference, YAML_DEFAULT
    pm = dict(pm or YAML_DEFAULT)
    ros = RosReference()
    dt = pm["DT"]
    fleet = Fleet(R, seed)
    ctx = [ros.tracker(pm, delay_mode, 0.5) for _ in range(R)]
    # (the node starts with _w = 0, _throttle = 1.0, driving_state.cpp:24-27; here every loop starts from the fleet's state)
    st3 = [np.array([fleet.w[i], fleet.thr[i], pm["REF_V"]]) for i in range(R)]
    trace = dict(w=[], thr=[], dist=[], cmd_v=[])
    for t in range(T):
        w = np.zeros(R); thr = np.zeros(R); sp = np.zeros(R)
        for i in range(R):
            px, py = fleet.paths[fleet.kind[i]]
            n = len(px)
            # the open plan the node would hold: from the current plan index, one lap at most
            idx = (int(fleet.idx[i]) + np.arange(0, 64 + fleet.win)) % n
            r = ros.window(px[idx][:64 + fleet.win], py[idx][:64 + fleet.win], fleet.pose[0, i], fleet.pose[1, i], 5.0)
            fleet.idx[i] = (int(fleet.idx[i]) + r["erased"]) % n
            # the reference down-samples the WHOLE remaining plan; the window is the first path_length of it
            sel = (int(fleet.idx[i]) + np.array(list(range(0, fleet.win, fleet.step)) + [fleet.win - 1])) % n
            goal = (fleet.pose[0, i] + 1e3, fleet.pose[1, i])           # closed track: never near a goal
            k = ros.tick(ctx[i], fleet.pose[:, i], goal, fleet.v[i], px[sel], py[sel], st3[i])
            st3[i] = k["state3"]
            w[i] = k["w"]; thr[i] = k["throttle"]; sp[i] = k["cmd"][0]
        trace["w"].append(w.copy()); trace["thr"].append(thr.copy()); trace["cmd_v"].append(sp.copy())
        trace["dist"].append(fleet.track_distance())
        fleet.actuate(w, thr, dt, pm["REF_V"])
    out = {k: np.array(v) for k, v in trace.items()}
    out["kind"] = fleet.kind.copy()
    return out
