"""TEST INFRASTRUCTURE -- CPU restatement of the trajectory initialisation either side of the CHOMP loop
(SURVEY 8f-4).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

  interpolate_waypoints   omg/util.py:238-258 (scipy CubicSpline(bc_type="clamped") / interp1d "linear"), restated
                          without scipy: the tridiagonal system of scipy/interpolate/_cubic.py solved by Thomas
                          elimination, CubicHermiteSpline's piece coefficients, PPoly's power-sum evaluation.
  dynamic_timesteps       omg/core.py:64-72
  trajectory_init         omg/core.py:59-78 (Trajectory.interpolate_waypoints: knots are always start and end)

Pinned: tests/golden/assets_traj.npz holds the outputs of the reference's own omg.util.interpolate_waypoints /
omg.core.Trajectory (tools/make_golden_assets.py); tests/test_oracle_assets.py replays them (<= 1e-13)."""
import numpy as np


def _clamped_slopes(x, y):
    """Knot derivatives of the clamped cubic spline through (x, y[K, m])."""
    K = x.shape[0]
    s = np.zeros_like(y)
    if K == 2:
        return s
    dx = np.diff(x)
    slope = np.diff(y, axis=0) / dx[:, None]
    cp = np.zeros(K)
    dp = np.zeros_like(y)
    for k in range(1, K - 1):
        lower, diag, upper = dx[k], 2.0 * (dx[k - 1] + dx[k]), dx[k - 1]
        rhs = 3.0 * (dx[k] * slope[k - 1] + dx[k - 1] * slope[k])
        den = diag - lower * cp[k - 1]
        cp[k] = upper / den
        dp[k] = (rhs - lower * dp[k - 1]) / den
    for k in range(K - 2, 0, -1):
        s[k] = dp[k] - cp[k] * s[k + 1]
    return s


def interpolate_waypoints(waypoints, n, m=None, mode="cubic"):
    """waypoints [K, m] -> [n, m] at the interior points of linspace(0, 1, n + 2)."""
    y = np.asarray(waypoints, dtype=np.float64)
    K = y.shape[0]
    x = np.linspace(0, 1, K)
    t = np.linspace(0, 1, n + 2)[1:-1]
    lo = np.clip(np.searchsorted(x, t, side="right") - 1, 0, K - 2)
    x_lo, x_hi = x[lo], x[lo + 1]
    if mode == "linear":
        return ((t - x_lo) / (x_hi - x_lo))[:, None] * y[lo + 1] + ((x_hi - t) / (x_hi - x_lo))[:, None] * y[lo]
    s = _clamped_slopes(x, y)
    dx = (x_hi - x_lo)[:, None]
    slope = (y[lo + 1] - y[lo]) / dx
    tt = (s[lo] + s[lo + 1] - 2 * slope) / dx
    c0 = tt / dx
    c1 = (slope - s[lo]) / dx - tt
    c2, c3 = s[lo], y[lo]
    h = (t - x_lo)[:, None]
    res = c3 + c2 * h
    z = h * h
    res = res + c1 * z
    z = z * h
    return res + c0 * z


def dynamic_timesteps(start, end, traj_delta=0.05, traj_min_step=2, traj_max_step=50):
    d = np.linalg.norm(np.asarray(start, dtype=np.float64) - np.asarray(end, dtype=np.float64), axis=-1)
    return np.minimum(np.maximum((d / traj_delta).astype(int), traj_min_step), traj_max_step)


def trajectory_init(start, end, n, mode="cubic"):
    """[B,9],[B,9] -> [B,n,9]."""
    start, end = np.atleast_2d(start), np.atleast_2d(end)
    return np.stack([interpolate_waypoints(np.stack([start[b], end[b]]), n, mode=mode) for b in range(start.shape[0])])
