"""Shared helpers of the parity tests: golden fixtures and the tolerance of BASELINE.json.

Tolerance (north_star): contact forces within 1e-4 relative / 1e-5 absolute.  It is applied
element-wise against x*, the exact optimum of the reference's QP (oracle.polish_from_working_set on
the converged qpOASES working set).  Converged qpOASES itself is only within ~3e-3 N of x*
(SURVEY.md section 0 fact 3), so against qpOASES the check is norm-wise:
    ||f - f_qpOASES||_inf <= max(1e-4 ||f||_inf, ||f_qpOASES - x*||_inf) + 1e-5.
"""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL, ATOL = 1e-4, 1e-5
KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait", "mu", "f_max")


def golden_files(reference_runnable=False):
    """All MPC fixtures; reference_runnable=True leaves out the horizon-30 ones (BASELINE configs[4]): the reference's
    own arrays stop at K_MAX_GAIT_SEGMENTS = 16, so its source build cannot be run on them."""
    files = sorted(glob.glob(os.path.join(HERE, "golden", "mpc_*.npz")))
    return [f for f in files if "_h30_" not in f] if reference_runnable else files


def load_golden(path, pkg):
    z = np.load(path)
    robot, h, dt, gait, seed, mu_sweep = z["meta"]
    b = {k: np.ascontiguousarray(z[k]) for k in KEYS}
    b["robot"] = pkg.robots.ROBOTS[str(robot)]
    b["horizon"] = int(h)
    b["dt"] = float(dt)
    return z, b, int(h), float(dt), bool(int(mu_sweep))


def assert_elementwise(u, x_star, what=""):
    u = np.asarray(u, float)
    err = np.abs(u - x_star)
    tol = RTOL * np.abs(x_star) + ATOL
    worst = (err / tol).max()
    assert worst <= 1.0, f"{what}: element-wise parity violated, worst err/tol = {worst:.3g}"
    return worst


def assert_vs_qpoases(u, x_conv, x_star, what=""):
    u = np.asarray(u, float)
    d = np.abs(u - x_conv).max()
    bound = max(RTOL * np.abs(x_conv).max(), np.abs(x_conv - x_star).max()) + ATOL
    assert d <= bound, f"{what}: ||f - f_qpOASES||_inf = {d:.3g} > {bound:.3g}"


def swing_mask(gait_row, h):
    """Boolean [12h]: True where the variable belongs to a swing foot-step (table entry 0)."""
    return np.repeat(np.asarray(gait_row).reshape(4 * h) == 0, 3)
