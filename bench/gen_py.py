"""bench/gen_py.py -- ctypes binding of the synthetic problem generator (bench/problem_gen.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmpc_gen.so")
        if not os.path.exists(path):
            subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++14", "-o", path,
                                   os.path.join(_HERE, "problem_gen.cpp")])
        _LIB = C.CDLL(path)
        _LIB.mpcgen_num_waypoints.restype = C.c_int
        _LIB.mpcgen_num_waypoints.argtypes = [C.c_double]
        _LIB.mpcgen_path_size.restype = C.c_int
        _LIB.mpcgen_path_size.argtypes = [C.c_int]
        dp = C.POINTER(C.c_double)
        _LIB.mpcgen_path_copy.argtypes = [C.c_int, dp, dp]
        _LIB.mpcgen_problems.argtypes = [C.c_uint64, C.c_int, C.c_double, dp, dp, dp, dp, C.POINTER(C.c_int)]
    return _LIB


def path(kind):
    L = _lib()
    n = L.mpcgen_path_size(kind)
    x = np.zeros(n); y = np.zeros(n)
    dp = C.POINTER(C.c_double)
    L.mpcgen_path_copy(kind, x.ctypes.data_as(dp), y.ctypes.data_as(dp))
    return x, y


def problems(seed, batch, path_length=5.0):
    """Returns dict(wx, wy [M x batch], pose [3 x batch], vel [3 x batch] = v, prev w, prev throttle, kind)."""
    L = _lib()
    M = L.mpcgen_num_waypoints(path_length)
    wx = np.zeros((M, batch)); wy = np.zeros((M, batch)); pose = np.zeros((3, batch)); vel = np.zeros((3, batch))
    kind = np.zeros(batch, dtype=np.int32)
    dp = C.POINTER(C.c_double)
    L.mpcgen_problems(C.c_uint64(seed), batch, path_length, wx.ctypes.data_as(dp), wy.ctypes.data_as(dp),
                      pose.ctypes.data_as(dp), vel.ctypes.data_as(dp), kind.ctypes.data_as(C.POINTER(C.c_int)))
    return dict(wx=wx, wy=wy, pose=pose, vel=vel, kind=kind, M=M)
