"""mpc_ros_b200/capi.py -- ctypes view of the C ABI in include/mpc_b200.h.

This is harness plumbing for tests/ and bench.py (the product's host side is C++:
mpc_ros_b200/include/mpc_planner.h).  It never computes anything itself and has no fallback:
every solve goes through libmpc_b200.so's CUDA kernels, and loading fails loudly when the
library has not been built (python -c "import __graft_entry__ as g; g.build()").
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPC_B200_LIB") or os.path.join(_HERE, "lib", "libmpc_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class Params(C.Structure):
    """mirror of struct mpc_b200_params"""
    _fields_ = [
        ("mpc_steps", C.c_int32),
        ("dt", C.c_double), ("ref_cte", C.c_double), ("ref_etheta", C.c_double), ("ref_vel", C.c_double),
        ("w_cte", C.c_double), ("w_etheta", C.c_double), ("w_vel", C.c_double), ("w_angvel", C.c_double),
        ("w_accel", C.c_double), ("w_angvel_d", C.c_double), ("w_accel_d", C.c_double),
        ("max_angvel", C.c_double), ("max_throttle", C.c_double), ("bound_value", C.c_double),
        ("tol", C.c_double), ("max_iter", C.c_int32),
        ("delay_mode", C.c_int32), ("max_speed", C.c_double), ("path_length", C.c_double),
        ("waypoints_dist", C.c_double), ("goal_radius", C.c_double), ("controller_freq", C.c_double),
        ("warm_mu_init", C.c_double),
    ]


# every symbol include/mpc_b200.h declares
EXPORTS = [
    "mpc_b200_params_default", "mpc_b200_params_yaml_default", "mpc_b200_params_from_yaml",
    "mpc_b200_params_set", "mpc_b200_create", "mpc_b200_destroy", "mpc_b200_set_params",
    "mpc_b200_get_params", "mpc_b200_set_option", "mpc_b200_warm_size", "mpc_b200_solve_batch", "mpc_b200_polyfit_batch",
    "mpc_b200_prestep_batch", "mpc_b200_warm_shift", "mpc_b200_window_batch", "mpc_b200_poststep_batch",
    "mpc_b200_num_waypoints", "mpc_b200_decel_batch", "mpc_b200_plant_step_batch", "mpc_b200_stream_create", "mpc_b200_stream_destroy", "mpc_b200_stream_synchronize", "mpc_b200_track_batch", "mpc_b200_track_submit", "mpc_b200_track_wait",
    "mpc_b200_last_kernel_seconds", "mpc_b200_launch_count", "mpc_b200_strerror",
    "mpc_b200_last_cuda_error", "mpc_b200_version", "mpc_b200_device_count", "mpc_b200_measure_fp64_peak",
    "mpc_b200_track_packed_layout", "mpc_b200_track_packed_submit", "mpc_b200_host_alloc", "mpc_b200_host_free",
    "mpc_b200_track_slice_submit",
    "mpc_b200_debug_profile", "mpc_b200_debug_fp64_probe",
]

_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libmpc_b200.so is not built (%s); run __graft_entry__.build(). "
                           "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.mpc_b200_params_default.argtypes = [C.POINTER(Params)]
    L.mpc_b200_params_yaml_default.argtypes = [C.POINTER(Params)]
    L.mpc_b200_params_from_yaml.argtypes = [C.c_char_p, C.POINTER(Params)]
    L.mpc_b200_params_from_yaml.restype = C.c_int
    L.mpc_b200_params_set.argtypes = [C.POINTER(Params), C.c_char_p, C.c_double]
    L.mpc_b200_params_set.restype = C.c_int
    L.mpc_b200_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Params), C.c_int32, C.c_int32]
    L.mpc_b200_create.restype = C.c_int
    L.mpc_b200_destroy.argtypes = [C.c_void_p]
    L.mpc_b200_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
    L.mpc_b200_set_params.restype = C.c_int
    L.mpc_b200_get_params.argtypes = [C.c_void_p, C.POINTER(Params)]
    L.mpc_b200_get_params.restype = C.c_int
    L.mpc_b200_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.mpc_b200_set_option.restype = C.c_int
    L.mpc_b200_warm_size.argtypes = [C.c_int32]
    L.mpc_b200_warm_size.restype = C.c_int32
    L.mpc_b200_solve_batch.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 12
    L.mpc_b200_solve_batch.restype = C.c_int
    L.mpc_b200_polyfit_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 6
    L.mpc_b200_polyfit_batch.restype = C.c_int
    L.mpc_b200_prestep_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 7
    L.mpc_b200_prestep_batch.restype = C.c_int
    L.mpc_b200_window_batch.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 10
    L.mpc_b200_window_batch.restype = C.c_int
    L.mpc_b200_poststep_batch.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 5
    L.mpc_b200_poststep_batch.restype = C.c_int
    L.mpc_b200_num_waypoints.argtypes = [C.POINTER(Params)]
    L.mpc_b200_num_waypoints.restype = C.c_int
    L.mpc_b200_decel_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    L.mpc_b200_decel_batch.restype = C.c_int
    L.mpc_b200_plant_step_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mpc_b200_plant_step_batch.restype = C.c_int
    L.mpc_b200_stream_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
    L.mpc_b200_stream_create.restype = C.c_int
    L.mpc_b200_stream_destroy.argtypes = [C.c_void_p]
    L.mpc_b200_stream_destroy.restype = C.c_int
    L.mpc_b200_stream_synchronize.argtypes = [C.c_void_p]
    L.mpc_b200_stream_synchronize.restype = C.c_int
    L.mpc_b200_track_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 12
    L.mpc_b200_track_batch.restype = C.c_int
    L.mpc_b200_track_submit.argtypes = L.mpc_b200_track_batch.argtypes
    L.mpc_b200_track_submit.restype = C.c_int
    L.mpc_b200_track_wait.argtypes = [C.c_void_p]
    L.mpc_b200_track_wait.restype = C.c_int
    L.mpc_b200_warm_shift.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mpc_b200_warm_shift.restype = C.c_int
    L.mpc_b200_last_kernel_seconds.argtypes = [C.c_void_p]
    L.mpc_b200_last_kernel_seconds.restype = C.c_double
    L.mpc_b200_launch_count.argtypes = [C.c_void_p]
    L.mpc_b200_launch_count.restype = C.c_int64
    L.mpc_b200_strerror.argtypes = [C.c_int]
    L.mpc_b200_strerror.restype = C.c_char_p
    L.mpc_b200_last_cuda_error.argtypes = [C.c_void_p]
    L.mpc_b200_last_cuda_error.restype = C.c_char_p
    L.mpc_b200_version.restype = C.c_int
    L.mpc_b200_device_count.restype = C.c_int
    L.mpc_b200_measure_fp64_peak.argtypes = [C.c_int32, C.c_int32]
    L.mpc_b200_measure_fp64_peak.restype = C.c_double
    L.mpc_b200_track_packed_layout.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
    L.mpc_b200_track_packed_layout.restype = C.c_int64
    L.mpc_b200_track_packed_submit.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.mpc_b200_track_packed_submit.restype = C.c_int
    L.mpc_b200_track_slice_submit.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 12
    L.mpc_b200_track_slice_submit.restype = C.c_int
    L.mpc_b200_host_alloc.argtypes = [C.c_size_t]
    L.mpc_b200_host_alloc.restype = C.c_void_p
    L.mpc_b200_host_free.argtypes = [C.c_void_p]
    if hasattr(L, "mpc_b200_debug_profile"):
        L.mpc_b200_debug_profile.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.mpc_b200_debug_profile.restype = C.c_int
    if hasattr(L, "mpc_b200_debug_fp64_probe"):
        L.mpc_b200_debug_fp64_probe.argtypes = [C.c_int32] * 4
        L.mpc_b200_debug_fp64_probe.restype = C.c_double
    _LIB = L
    return L


class MpcError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = lib().mpc_b200_strerror(code).decode()
        RuntimeError.__init__(self, "mpc_b200 error %d: %s %s" % (code, msg, detail))


def yaml_default_params():
    p = Params()
    lib().mpc_b200_params_yaml_default(C.byref(p))
    return p


def default_params():
    p = Params()
    lib().mpc_b200_params_default(C.byref(p))
    return p


def params_from_yaml(path, base=None):
    p = base if base is not None else default_params()
    rc = lib().mpc_b200_params_from_yaml(path.encode(), C.byref(p))
    if rc != 0:
        raise MpcError(rc, path)
    return p


def stream_create(device=0):
    """A non-blocking CUDA stream owned by the library (an integer cudaStream_t)."""
    s = C.c_void_p()
    rc = lib().mpc_b200_stream_create(device, C.byref(s))
    if rc != 0:
        raise MpcError(rc)
    return s.value


def stream_destroy(stream):
    lib().mpc_b200_stream_destroy(stream)


def params_from_map(pm, base=None):
    """pm: the reference's LoadParams map (DT, STEPS, REF_V, ...)."""
    p = base if base is not None else default_params()
    for k, v in pm.items():
        rc = lib().mpc_b200_params_set(C.byref(p), k.encode(), float(v))
        if rc != 0:
            raise MpcError(rc, k)
    return p


def _addr(x):
    """host numpy array or anything with data_ptr() (a device tensor) -> integer address"""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if isinstance(x, int):          # a raw address (mpc_b200_host_alloc)
        return x
    return x.data_ptr()


class Solver:
    """Owns one mpc_b200_handle."""

    def __init__(self, params, max_batch, device=0):
        self._h = C.c_void_p()
        self.max_batch = max_batch
        rc = lib().mpc_b200_create(C.byref(self._h), C.byref(params), max_batch, device)
        if rc != 0:
            self._h = None
            raise MpcError(rc)
        self.N = params.mpc_steps
        self.nc = 4          # rows of the coeffs arrays (option poly_coeffs)

    def close(self):
        if self._h:
            lib().mpc_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params):
        rc = lib().mpc_b200_set_params(self._h, C.byref(params))
        if rc != 0:
            raise MpcError(rc)
        self.N = params.mpc_steps

    def set_option(self, name, value):
        rc = lib().mpc_b200_set_option(self._h, name.encode(), float(value))
        if rc != 0:
            raise MpcError(rc, name)
        if name == "poly_coeffs":
            self.nc = int(value)

    def solve_raw(self, batch, state, coeffs, u0, pred, ref_vel=None, warm_in=None, obj=None, status=None,
                  iters=None, kkt=None, warm_out=None, stream=None):
        rc = lib().mpc_b200_solve_batch(self._h, batch, _addr(state), _addr(coeffs), _addr(ref_vel), _addr(warm_in),
                                        _addr(u0), _addr(pred), _addr(obj), _addr(status), _addr(iters), _addr(kkt),
                                        _addr(warm_out), stream)
        if rc != 0:
            raise MpcError(rc, lib().mpc_b200_last_cuda_error(self._h).decode())

    def solve(self, state, coeffs, ref_vel=None):
        """Host numpy in/out.  state 6 x B, coeffs 4 x B."""
        state = np.ascontiguousarray(state, dtype=np.float64)
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        B = state.shape[1]
        N = self.N
        out = dict(u0=np.zeros((2, B)), pred=np.zeros((3 * N, B)), obj=np.zeros(B),
                   status=np.zeros(B, dtype=np.int32), iters=np.zeros(B, dtype=np.int32), kkt=np.zeros(B))
        if ref_vel is not None:
            ref_vel = np.ascontiguousarray(ref_vel, dtype=np.float64)
        self.solve_raw(B, state, coeffs, out["u0"], out["pred"], ref_vel=ref_vel, obj=out["obj"],
                       status=out["status"], iters=out["iters"], kkt=out["kkt"])
        return out

    def polyfit(self, wx, wy, pose):
        wx = np.ascontiguousarray(wx, dtype=np.float64)
        wy = np.ascontiguousarray(wy, dtype=np.float64)
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        M, B = wx.shape
        coeffs = np.zeros((self.nc, B)); ce = np.zeros((2, B))
        rc = lib().mpc_b200_polyfit_batch(self._h, B, M, _addr(wx), _addr(wy), _addr(pose), _addr(coeffs), _addr(ce), None)
        if rc != 0:
            raise MpcError(rc, lib().mpc_b200_last_cuda_error(self._h).decode())
        return coeffs, ce[0], ce[1]

    def polyfit_raw(self, batch, M, wx, wy, pose, coeffs, cte_etheta=None, stream=None):
        rc = lib().mpc_b200_polyfit_batch(self._h, batch, M, _addr(wx), _addr(wy), _addr(pose), _addr(coeffs),
                                          _addr(cte_etheta), stream)
        if rc != 0:
            raise MpcError(rc, lib().mpc_b200_last_cuda_error(self._h).decode())

    def window_raw(self, batch, path_x, path_y, track_off, track_len, track_id, idx, pose, wx, wy, stream=None):
        rc = lib().mpc_b200_window_batch(self._h, batch, _addr(path_x), _addr(path_y), _addr(track_off), _addr(track_len),
                                         _addr(track_id), _addr(idx), _addr(pose), _addr(wx), _addr(wy), stream)
        if rc != 0:
            raise MpcError(rc)

    def poststep_raw(self, batch, u0, vel, ref_vel, cmd, stream=None):
        rc = lib().mpc_b200_poststep_batch(self._h, batch, _addr(u0), _addr(vel), _addr(ref_vel), _addr(cmd), stream)
        if rc != 0:
            raise MpcError(rc)

    def decel_raw(self, batch, pose, goal, vel, min_speed, ref_vel, stream=None):
        rc = lib().mpc_b200_decel_batch(self._h, batch, _addr(pose), _addr(goal), _addr(vel), min_speed, _addr(ref_vel), stream)
        if rc != 0:
            raise MpcError(rc)

    def plant_step_raw(self, batch, cmd, pose, vel, stream=None):
        rc = lib().mpc_b200_plant_step_batch(self._h, batch, _addr(cmd), _addr(pose), _addr(vel), stream)
        if rc != 0:
            raise MpcError(rc)

    def track_raw(self, batch, M, wx, wy, pose, vel, u0, pred, ref_vel=None, cmd=None, obj=None, status=None, iters=None,
                  kkt=None):
        rc = lib().mpc_b200_track_batch(self._h, batch, M, _addr(wx), _addr(wy), _addr(pose), _addr(vel), _addr(ref_vel),
                                        _addr(u0), _addr(pred), _addr(cmd), _addr(obj), _addr(status), _addr(iters), _addr(kkt))
        if rc != 0:
            raise MpcError(rc, lib().mpc_b200_last_cuda_error(self._h).decode())

    def track_submit_raw(self, batch, M, wx, wy, pose, vel, u0, pred, ref_vel=None, cmd=None, obj=None, status=None,
                         iters=None, kkt=None):
        rc = lib().mpc_b200_track_submit(self._h, batch, M, _addr(wx), _addr(wy), _addr(pose), _addr(vel), _addr(ref_vel),
                                         _addr(u0), _addr(pred), _addr(cmd), _addr(obj), _addr(status), _addr(iters), _addr(kkt))
        if rc != 0:
            raise MpcError(rc, lib().mpc_b200_last_cuda_error(self._h).decode())

    def track_wait(self):
        rc = lib().mpc_b200_track_wait(self._h)
        if rc != 0:
            raise MpcError(rc, lib().mpc_b200_last_cuda_error(self._h).decode())

    def track(self, wx, wy, pose, vel, ref_vel=None):
        """One control tick for B robots, host numpy in/out (Tracking::findBestPath, driving_state.cpp:175-271).
        vel (3 x B: v, previous w, previous throttle) is updated in place for the next tick."""
        wx = np.ascontiguousarray(wx, dtype=np.float64); wy = np.ascontiguousarray(wy, dtype=np.float64)
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        if not (isinstance(vel, np.ndarray) and vel.dtype == np.float64 and vel.flags.c_contiguous):
            raise TypeError("vel must be a C-contiguous float64 array (it is updated in place)")
        M, B = wx.shape
        N = self.N
        out = dict(u0=np.zeros((2, B)), pred=np.zeros((3 * N, B)), cmd=np.zeros((2, B)), obj=np.zeros(B),
                   status=np.zeros(B, dtype=np.int32), iters=np.zeros(B, dtype=np.int32), kkt=np.zeros(B))
        if ref_vel is not None:
            ref_vel = np.ascontiguousarray(ref_vel, dtype=np.float64)
        self.track_raw(B, M, wx, wy, pose, vel, out["u0"], out["pred"], ref_vel=ref_vel, cmd=out["cmd"], obj=out["obj"],
                       status=out["status"], iters=out["iters"], kkt=out["kkt"])
        return out

    def packed_layout(self, batch, M, with_ref_vel=False):
        """(total bytes, dict name -> byte offset) of the one-buffer tick (mpc_b200_track_packed_submit)."""
        off = (C.c_int64 * 12)()
        total = lib().mpc_b200_track_packed_layout(self._h, batch, M, int(with_ref_vel), off)
        if total < 0:
            raise MpcError(int(total))
        names = ["wx", "wy", "pose", "ref_vel", "vel", "u0", "pred", "cmd", "obj", "kkt", "status", "iters"]
        return int(total), {n: int(o) for n, o in zip(names, off)}

    def packed_views(self, buf, batch, M, with_ref_vel=False):
        """numpy views of the blocks of a packed tick buffer (`buf`: uint8 array of packed_layout()[0] bytes)."""
        total, off = self.packed_layout(batch, M, with_ref_vel)
        N = self.N
        def f64(name, rows):
            return np.frombuffer(buf, dtype=np.float64, count=rows * batch, offset=off[name]).reshape(rows, batch)
        v = dict(wx=f64("wx", M), wy=f64("wy", M), pose=f64("pose", 3), vel=f64("vel", 3), u0=f64("u0", 2),
                 pred=f64("pred", 3 * N), cmd=f64("cmd", 2), obj=f64("obj", 1)[0], kkt=f64("kkt", 1)[0],
                 status=np.frombuffer(buf, dtype=np.int32, count=batch, offset=off["status"]),
                 iters=np.frombuffer(buf, dtype=np.int32, count=batch, offset=off["iters"]))
        if with_ref_vel:
            v["ref_vel"] = f64("ref_vel", 1)[0]
        return v

    def track_packed_submit(self, batch, M, io, with_ref_vel=False):
        rc = lib().mpc_b200_track_packed_submit(self._h, batch, M, int(with_ref_vel), _addr(io))
        if rc != 0:
            raise MpcError(rc, lib().mpc_b200_last_cuda_error(self._h).decode())

    def warm_shift(self, batch, prev, nxt, stream=None):
        rc = lib().mpc_b200_warm_shift(self._h, batch, _addr(prev), _addr(nxt), stream)
        if rc != 0:
            raise MpcError(rc)

    def prestep_raw(self, batch, M, wx, wy, pose, vel, coeffs, state, stream=None):
        rc = lib().mpc_b200_prestep_batch(self._h, batch, M, _addr(wx), _addr(wy), _addr(pose), _addr(vel),
                                          _addr(coeffs), _addr(state), stream)
        if rc != 0:
            raise MpcError(rc, lib().mpc_b200_last_cuda_error(self._h).decode())

    def prestep(self, wx, wy, pose, vel):
        wx = np.ascontiguousarray(wx, dtype=np.float64); wy = np.ascontiguousarray(wy, dtype=np.float64)
        pose = np.ascontiguousarray(pose, dtype=np.float64); vel = np.ascontiguousarray(vel, dtype=np.float64)
        M, B = wx.shape
        coeffs = np.zeros((self.nc, B)); state = np.zeros((6, B))
        self.prestep_raw(B, M, wx, wy, pose, vel, coeffs, state)
        return coeffs, state

    @property
    def last_kernel_seconds(self):
        return lib().mpc_b200_last_kernel_seconds(self._h)

    @property
    def launch_count(self):
        return lib().mpc_b200_launch_count(self._h)
