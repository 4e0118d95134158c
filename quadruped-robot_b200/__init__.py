"""quadruped-robot_b200 -- B200-native batched convex-MPC / WBC engine (host-side Python mirror).

The product is the C-ABI library built from csrc/ (libqr_gpu.so, declared in include/qr_gpu.h);
this package only carries the workload generator, the ctypes binding used by tests and bench.py,
and the build recipe.  It never imports anything from oracle/.
"""
from . import robots, synth  # noqa: F401
