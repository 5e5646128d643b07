"""mpc_ros_b200 -- B200-native batched NMPC solver behind mpc_ros's MPC::Solve.

The product is the C-ABI shared library (include/mpc_b200.h, built from csrc/ into
lib/libmpc_b200.so) and the C++ MPC adapter class (include/mpc_planner.h).  The Python in this
package is only the ctypes plumbing the tests and bench.py use to call that library.
"""
from . import capi  # noqa: F401
