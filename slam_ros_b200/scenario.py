"""Deterministic synthetic scan sequences for the EKF-SLAM hot path.

Everything here produces *inputs* only -- (u, lines[]) per step in the reference's (alfa, r) line
convention (slam_ros/simplifyPath.h:62-79: angle first, then distance; C_AR = diag(var_alfa, var_r),
lineFitting.cpp:446-448) -- so that the CPU oracle and libekfcuda consume identical bytes.

Two families:

* ``room_scenario``  -- BASELINE.json configs[0]: a 2-D room swept by a 361-beam scanner; beams are
  ray-cast against wall segments, grouped by the wall they hit and fitted by total least squares
  (a fixed, deterministic extractor -- the reference's own extractor reads uninitialised memory,
  SURVEY.md 8c, so its output cannot be a fixture).
* ``map_scenario``   -- configs[1..4]: N random world lines seeded through the augmentation path
  (one scan on the empty map at the origin), then m observed landmarks per step.

The declared line covariance is deliberately conservative relative to the actual noise (Q9: the
reference gate is sqrt(|v' S^-1 v|) <= 0.4, Robot.h:15 / Robot.cpp:489).
"""
import math

import numpy as np

SIGMA_ALFA = 2e-3     # declared std of a line angle  [rad]
SIGMA_R = 5e-3        # declared std of a line distance [m]
NOISE_FRAC = 0.05     # actual noise = NOISE_FRAC * declared


def wrap_pi(a):
    """Wrap to (-pi, pi]."""
    a = np.asarray(a, dtype=np.float64)
    w = np.mod(a + math.pi, 2.0 * math.pi) - math.pi
    return np.where(w <= -math.pi, w + 2.0 * math.pi, w)


def declared_R(m):
    R = np.zeros((m, 4))
    R[:, 0] = SIGMA_ALFA ** 2
    R[:, 3] = SIGMA_R ** 2
    return R


def circle_odometry(steps, d=0.02, dtheta=0.02):
    """u = (d, 0, dtheta) per step: a circle of radius d/dtheta that starts at the origin."""
    u = np.zeros((steps, 3))
    u[:, 0] = d
    u[:, 2] = dtheta
    return u


def true_trajectory(u, pose0=(0.0, 0.0, 0.0)):
    """Integrate the reference motion model (Robot.cpp:148) without noise."""
    poses = np.zeros((u.shape[0] + 1, 3))
    poses[0] = pose0
    for s in range(u.shape[0]):
        x, y, th = poses[s]
        a = th + u[s, 2] / 2.0
        poses[s + 1] = (x + u[s, 0] * math.cos(a), y + u[s, 0] * math.sin(a), th + u[s, 2])
    return poses


def observe(world_lines, pose):
    """Reference observation model h (Robot.cpp:423-425) for world lines (alfa_w, r_w)."""
    x, y, th = pose
    al = world_lines[:, 0]
    z = np.empty_like(world_lines)
    z[:, 0] = wrap_pi(al - th)
    z[:, 1] = world_lines[:, 1] - (x * np.cos(al) + y * np.sin(al))
    return z


def map_scenario(n_landmarks, steps, m=8, seed=1234, stride=7, d=0.02, dtheta=0.02):
    """configs[1..4] generator.

    Returns a dict: world (N,2); seed_z (N,2), seed_R (N,4) -- the single scan that seeds the map on the
    empty filter at the origin; u (S,3); z (S,m,2); R (S,m,4); idx (S,m) true landmark of each line.
    Step s observes landmarks (stride*s + q) mod N, q = 0..m-1.
    """
    rng = np.random.default_rng(seed)
    N = int(n_landmarks)
    world = np.empty((N, 2))
    world[:, 0] = wrap_pi(rng.uniform(-math.pi, math.pi, N))
    world[:, 1] = rng.uniform(2.0, 8.0, N)
    noise = np.array([SIGMA_ALFA, SIGMA_R]) * NOISE_FRAC
    seed_z = observe(world, (0.0, 0.0, 0.0)) + rng.standard_normal((N, 2)) * noise
    seed_z[:, 0] = wrap_pi(seed_z[:, 0])
    u = circle_odometry(steps, d, dtheta)
    poses = true_trajectory(u)
    z = np.empty((steps, m, 2))
    idx = np.empty((steps, m), dtype=np.int64)
    for s in range(steps):
        ids = (stride * s + np.arange(m)) % N
        idx[s] = ids
        zz = observe(world[ids], poses[s + 1]) + rng.standard_normal((m, 2)) * noise
        zz[:, 0] = wrap_pi(zz[:, 0])
        z[s] = zz
    R = np.broadcast_to(declared_R(m), (steps, m, 4)).copy()
    return {"world": world, "seed_z": seed_z, "seed_R": declared_R(N), "u": u, "z": z, "R": R, "idx": idx,
            "poses": poses, "seed": seed, "m": m, "N": N}


# --------------------------------------------------------------------------------------------
# configs[0]: 2-D room, 361 beams
# --------------------------------------------------------------------------------------------
def room_walls():
    """10 m x 8 m room centred on the origin plus interior partitions: ~20 distinct wall lines."""
    W = []

    def seg(x0, y0, x1, y1):
        W.append((x0, y0, x1, y1))

    seg(-5, -4, 5, -4); seg(5, -4, 5, 4); seg(5, 4, -5, 4); seg(-5, 4, -5, -4)      # outer walls
    seg(-5, 1.5, -3.2, 1.5); seg(-3.2, 1.5, -3.2, 2.6)                                # alcove NW
    seg(3.0, -4, 3.0, -2.4); seg(3.0, -2.4, 4.1, -2.4)                                # closet SE
    seg(1.5, 4, 1.5, 2.7); seg(-1.0, -4, -1.0, -3.0)                                  # stubs
    seg(2.2, 1.2, 3.6, 2.0); seg(-3.5, -2.6, -2.3, -1.7)                              # slanted furniture
    seg(3.6, 2.0, 3.0, 3.0); seg(-2.3, -1.7, -3.0, -0.9)
    seg(-0.6, 2.4, 0.6, 2.4); seg(0.6, 2.4, 0.6, 3.1)                                 # desk
    seg(4.0, -0.5, 5.0, -0.5); seg(-5, -1.0, -4.2, -1.0)                              # shelves
    seg(0.8, -2.2, 1.9, -2.9); seg(-1.9, 0.4, -1.2, 1.3)
    return np.array(W, dtype=np.float64)


def _raycast(walls, pose, angles, max_range):
    x, y, th = pose
    dx = np.cos(angles + th)[:, None]
    dy = np.sin(angles + th)[:, None]
    x0, y0, x1, y1 = walls[:, 0][None], walls[:, 1][None], walls[:, 2][None], walls[:, 3][None]
    ex, ey = x1 - x0, y1 - y0
    den = dx * ey - dy * ex
    with np.errstate(divide="ignore", invalid="ignore"):
        t = ((x0 - x) * ey - (y0 - y) * ex) / den
        s = ((x0 - x) * dy - (y0 - y) * dx) / den
    ok = (np.abs(den) > 1e-12) & (t > 1e-6) & (s >= 0.0) & (s <= 1.0)
    t = np.where(ok, t, np.inf)
    wall = np.argmin(t, axis=1)
    rng_ = t[np.arange(t.shape[0]), wall]
    hit = np.isfinite(rng_) & (rng_ <= max_range)
    return rng_, wall, hit


def _fit_line(px, py):
    """Total least squares (alfa, r) with r >= 0, alfa in (-pi, pi]: x cos(alfa) + y sin(alfa) = r."""
    mx, my = px.mean(), py.mean()
    sxx = ((px - mx) ** 2).sum(); syy = ((py - my) ** 2).sum(); sxy = ((px - mx) * (py - my)).sum()
    alfa = 0.5 * math.atan2(-2.0 * sxy, syy - sxx)
    r = mx * math.cos(alfa) + my * math.sin(alfa)
    if r < 0:
        r = -r
        alfa += math.pi
    return float(wrap_pi(alfa)), float(r)


def room_scenario(steps=1000, seed=7, beams=361, range_sigma=1e-3, max_range=10.0, min_points=10,
                  max_lines=9, d=0.02):
    """configs[0]: per step up to ``max_lines`` fitted lines (robot frame).  Returns ragged lists packed as
    z (S, max_lines, 2), R (S, max_lines, 4), count (S,), u (S,3)."""
    rng = np.random.default_rng(seed)
    walls = room_walls()
    u = np.zeros((steps, 3))
    u[:, 0] = d
    u[:, 2] = 0.01 * np.sin(0.01 * np.arange(steps)) + 0.012
    poses = true_trajectory(u)
    angles = np.deg2rad(np.arange(beams, dtype=np.float64)) - math.pi
    z = np.zeros((steps, max_lines, 2)); R = np.zeros((steps, max_lines, 4)); cnt = np.zeros(steps, dtype=np.int64)
    for s in range(steps):
        rng_, wall, hit = _raycast(walls, poses[s + 1], angles, max_range)
        rr = rng_ + rng.standard_normal(beams) * range_sigma
        lines = []
        start = 0
        for b in range(1, beams + 1):
            if b == beams or wall[b] != wall[start] or not hit[b] or not hit[start]:
                if hit[start] and b - start >= min_points:
                    sl = slice(start + 1, b - 1)      # drop the corner beams
                    px = rr[sl] * np.cos(angles[sl]); py = rr[sl] * np.sin(angles[sl])
                    if px.size >= min_points - 2:
                        lines.append(_fit_line(px, py))
                start = b
        lines = lines[:max_lines]
        cnt[s] = len(lines)
        for i, (a, r) in enumerate(lines):
            z[s, i] = (a, r)
        R[s, :len(lines)] = declared_R(len(lines))
    return {"u": u, "z": z, "R": R, "count": cnt, "poses": poses, "seed": seed, "walls": walls}


def room_scan(pose, beams=361, range_sigma=1e-3, max_range=10.0, rng=None, step_deg=1.0):
    """One `mappingPoints` payload as the node receives it (slam_ros/main.cpp:37-56): float32 pairs (r, angle),
    angle in [0, 2 pi] at `step_deg` steps (1 degree in the reference's simulator; the reference subtracts pi),
    r = 0 for beams without a return."""
    walls = room_walls()
    angles = np.deg2rad(np.arange(beams, dtype=np.float64) * step_deg)
    rng_, wall, hit = _raycast(walls, pose, angles - math.pi, max_range)
    rr = np.where(hit, rng_, 0.0)
    if rng is not None and range_sigma > 0:
        rr = np.where(hit, rr + rng.standard_normal(beams) * range_sigma, 0.0)
    out = np.empty((beams, 2), dtype=np.float32)
    out[:, 0] = rr
    out[:, 1] = angles
    return out


def room_scans(steps=20, seed=7, beams=361, range_sigma=1e-3, d=0.05):
    """A short trajectory through the room: (S, beams, 2) float32 payloads plus the true poses."""
    rng = np.random.default_rng(seed)
    u = np.zeros((steps, 3))
    u[:, 0] = d
    u[:, 2] = 0.03 * np.sin(0.2 * np.arange(steps)) + 0.02
    poses = true_trajectory(u)
    scans = np.stack([room_scan(poses[s + 1], beams, range_sigma, rng=rng) for s in range(steps)])
    return {"scans": scans, "poses": poses, "u": u}


def encoder_for(pose_est, u):
    """The `encoder` argument that makes Robot.cpp:140-145 recover the intended odometry (Q5):
    u[2] = theta - enc[2], u[0] = |xy - enc_xy|."""
    return np.array([pose_est[0] - u[0], pose_est[1], pose_est[2] - u[2]], dtype=np.float64)
