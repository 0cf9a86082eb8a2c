"""Seeded synthetic inputs for the scan-registration hot path.

Pure numpy, no oracle and no reference dependency: an analytic ray-caster over a "city block" scene
(ground plane, 6x6 box buildings on a 24 m pitch, 150 vertical poles) producing ring scans in the wire
format the hot path consumes (PointXYZIRT, 32 B AoS: x@0 y@4 z@8 intensity@16 ring@20 time@24 — the
layout of liosam_ws/src/LIO-SAM/src/imageProjection.cpp:4-15), plus the gyro table of
imageProjection.cpp:305-362 for the deskew config.
"""
import numpy as np

MASTER_SEED = 20261018

XYZIRT = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                   "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                   "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})


class CityBlock:
    """160 x 160 m ground, 6x6 buildings 12x12x10 m on a 24 m pitch, 150 poles r=0.15 m h=6 m."""

    def __init__(self, seed=MASTER_SEED, tiles=1):
        rng = np.random.default_rng(seed)
        self.half = 80.0 * tiles
        c = -60.0 + 24.0 * np.arange(6)
        offs = (np.arange(tiles) - (tiles - 1) / 2.0) * 160.0
        bx, by = [], []
        for ox in offs:
            for oy in offs:
                gx, gy = np.meshgrid(c + ox, c + oy, indexing="ij")
                bx.append(gx.ravel()); by.append(gy.ravel())
        bx, by = np.concatenate(bx), np.concatenate(by)
        # per-building jitter in footprint/height so walls are not all coplanar across blocks
        n = bx.size
        hx = 6.0 + rng.uniform(-0.5, 0.5, n)
        hy = 6.0 + rng.uniform(-0.5, 0.5, n)
        hz = 10.0 + rng.uniform(-2.0, 2.0, n)
        self.box_min = np.stack([bx - hx, by - hy, np.zeros(n)], 1)
        self.box_max = np.stack([bx + hx, by + hy, hz], 1)
        # poles along the street centre lines +-4.5 m (kerb side), never inside a building footprint
        npole = 150 * tiles * tiles
        streets = -48.0 + 24.0 * np.arange(5)
        px, py = np.empty(npole), np.empty(npole)
        for i in range(npole):
            along = rng.uniform(-self.half + 5, self.half - 5)
            s = streets[rng.integers(0, 5)] + offs[rng.integers(0, tiles)] + rng.choice([-4.5, 4.5])
            if rng.random() < 0.5:
                px[i], py[i] = along, s
            else:
                px[i], py[i] = s, along
        self.pole_xy = np.stack([px, py], 1)
        self.pole_r = 0.15
        self.pole_h = 6.0

    def raycast(self, origin, dirs, max_range=120.0):
        """origin (3,), dirs (R,3) unit. Returns range (R,) with inf where nothing is hit."""
        o = np.asarray(origin, np.float64)
        d = np.asarray(dirs, np.float64)
        R = d.shape[0]
        best = np.full(R, np.inf)
        # ground z = 0
        with np.errstate(divide="ignore", invalid="ignore"):
            t = -o[2] / d[:, 2]
        hit = (d[:, 2] < 0) & (t > 0)
        gx = o[0] + t * d[:, 0]; gy = o[1] + t * d[:, 1]
        hit &= (np.abs(gx) <= self.half) & (np.abs(gy) <= self.half)
        best = np.where(hit, t, best)
        # boxes: slab method, chunked over rays to bound memory
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
        B = self.box_min.shape[0]
        chunk = max(1, int(4e6 // max(B, 1)))
        for s in range(0, R, chunk):
            e = min(R, s + chunk)
            t0 = (self.box_min[None, :, :] - o[None, None, :]) * inv[s:e, None, :]
            t1 = (self.box_max[None, :, :] - o[None, None, :]) * inv[s:e, None, :]
            tmin = np.minimum(t0, t1).max(axis=2)
            tmax = np.maximum(t0, t1).min(axis=2)
            ok = (tmax >= tmin) & (tmax > 0)
            tt = np.where(ok, np.where(tmin > 0, tmin, tmax), np.inf).min(axis=1)
            best[s:e] = np.minimum(best[s:e], tt)
        # vertical cylinders
        P = self.pole_xy.shape[0]
        chunk = max(1, int(4e6 // max(P, 1)))
        a = d[:, 0] ** 2 + d[:, 1] ** 2
        for s in range(0, R, chunk):
            e = min(R, s + chunk)
            fx = o[0] - self.pole_xy[None, :, 0]
            fy = o[1] - self.pole_xy[None, :, 1]
            b = fx * d[s:e, None, 0] + fy * d[s:e, None, 1]
            c = fx * fx + fy * fy - self.pole_r ** 2
            disc = b * b - a[s:e, None] * c
            with np.errstate(divide="ignore", invalid="ignore"):
                tc = (-b - np.sqrt(np.maximum(disc, 0))) / a[s:e, None]
            z = o[2] + tc * d[s:e, None, 2]
            ok = (disc > 0) & (tc > 0) & (z >= 0) & (z <= self.pole_h)
            tt = np.where(ok, tc, np.inf).min(axis=1)
            best[s:e] = np.minimum(best[s:e], tt)
        best[best > max_range] = np.inf
        return best


def rot_zyx(roll, pitch, yaw):
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, sy * sr + cy * sp * cr],
                     [sy * cp, cy * cr + sy * sp * sr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def ring_scan(scene, pose, n_rings=16, n_cols=1800, elev_deg=(-15.0, 15.0), seed=0, noise=0.01, dropout=0.02,
              scan_time=0.1, omega=None, max_range=120.0):
    """One ring scan in the XYZIRT wire format, firing order = column-major (all rings of a column together).

    pose = (roll, pitch, yaw, x, y, z) of the sensor at scan start. omega: optional callable t -> (3,) rad/s body
    rates; when given, the sensor orientation is integrated over the scan (the motion deskew undoes).
    Azimuths sit at bin centres with +-0.35 bin jitter so atan2 rounding rarely decides a column.
    """
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(elev_deg[0], elev_deg[1], n_rings))
    res = 2 * np.pi / n_cols
    cols = np.arange(n_cols)
    az = -np.pi + (cols + 0.5) * res + rng.uniform(-0.35, 0.35, n_cols) * res
    t_col = cols / n_cols * scan_time
    A, E = np.meshgrid(az, elev, indexing="ij")         # (cols, rings)
    d_s = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
    ring = np.tile(np.arange(n_rings), n_cols).astype(np.uint16)
    tpt = np.repeat(t_col, n_rings)
    R0 = rot_zyx(pose[0], pose[1], pose[2])
    if omega is None:
        d_w = d_s @ R0.T
    else:
        # integrate body rates column by column (first-order, matches what a gyro table can undo)
        d_w = np.empty_like(d_s)
        Rt = R0.copy()
        dt = scan_time / n_cols
        for j in range(n_cols):
            sl = slice(j * n_rings, (j + 1) * n_rings)
            d_w[sl] = d_s[sl] @ Rt.T
            w = np.asarray(omega(t_col[j])) * dt
            th = np.linalg.norm(w)
            if th > 0:
                k = w / th
                K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
                Rt = Rt @ (np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K)
    rngs = scene.raycast(np.asarray(pose[3:6], np.float64), d_w, max_range)
    keep = np.isfinite(rngs) & (rng.random(rngs.size) >= dropout)
    r = rngs[keep] + rng.normal(0.0, noise, int(keep.sum()))
    pts = d_s[keep] * r[:, None]
    out = np.zeros(pts.shape[0], XYZIRT)
    out["x"], out["y"], out["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    out["intensity"] = rng.uniform(1.0, 100.0, pts.shape[0])
    out["ring"] = ring[keep]
    out["time"] = tpt[keep]
    return out


def imu_table(t0, scan_time, omega, rate=500.0):
    """Gyro table as imuDeskewInfo builds it (imageProjection.cpp:305-362): samples over
    [t0-0.01, t0+scan_time+0.01], rotation integrated as rot[i] = rot[i-1] + w(t_i) * dt, rot[0] = 0."""
    t = np.arange(t0 - 0.01, t0 + scan_time + 0.01 + 1e-12, 1.0 / rate)
    rot = np.zeros((t.size, 3))
    for i in range(1, t.size):
        rot[i] = rot[i - 1] + np.asarray(omega(t[i] - t0)) * (t[i] - t[i - 1])
    return t, rot[:, 0].copy(), rot[:, 1].copy(), rot[:, 2].copy()


def street_loop(n_blocks_x=2, n_blocks_y=1, spacing=1.0, z=1.8):
    """Key-frame poses every `spacing` m along a rectangular street loop around n_blocks_x x n_blocks_y blocks."""
    x0, y0 = -24.0, -24.0
    x1, y1 = x0 + 24.0 * n_blocks_x, y0 + 24.0 * n_blocks_y
    corners = [(x0, y0), (x1, y0), (x1, y1), (x0, y1), (x0, y0)]
    poses = []
    for (ax, ay), (bx, by) in zip(corners[:-1], corners[1:]):
        L = np.hypot(bx - ax, by - ay)
        yaw = np.arctan2(by - ay, bx - ax)
        for s in np.arange(0.0, L, spacing):
            poses.append((0.0, 0.0, yaw, ax + (bx - ax) * s / L, ay + (by - ay) * s / L, z))
    return np.array(poses)


def sample_surfaces(scene, n, seed=0, noise=0.02):
    """n points uniform on the scene surfaces (ground, walls, roofs, poles) + N(0, noise). float64 (n,3)."""
    rng = np.random.default_rng(seed)
    ext = scene.box_max - scene.box_min
    wall_area = 2 * (ext[:, 0] + ext[:, 1]) * ext[:, 2]
    roof_area = ext[:, 0] * ext[:, 1]
    ground_area = (2 * scene.half) ** 2
    pole_area = 2 * np.pi * scene.pole_r * scene.pole_h * scene.pole_xy.shape[0]
    areas = np.array([ground_area, wall_area.sum(), roof_area.sum(), pole_area])
    counts = rng.multinomial(n, areas / areas.sum())
    out = []
    g = rng.uniform(-scene.half, scene.half, (counts[0], 2))
    out.append(np.column_stack([g, np.zeros(counts[0])]))
    b = rng.choice(ext.shape[0], counts[1], p=wall_area / wall_area.sum())
    per = rng.uniform(0, 1, counts[1]) * 2 * (ext[b, 0] + ext[b, 1])
    zz = rng.uniform(0, 1, counts[1]) * ext[b, 2]
    ex, ey = ext[b, 0], ext[b, 1]
    x = np.where(per < ex, per, np.where(per < ex + ey, ex, np.where(per < 2 * ex + ey, 2 * ex + ey - per, 0.0)))
    y = np.where(per < ex, 0.0, np.where(per < ex + ey, per - ex, np.where(per < 2 * ex + ey, ey, 2 * (ex + ey) - per)))
    out.append(np.column_stack([scene.box_min[b, 0] + x, scene.box_min[b, 1] + y, zz]))
    b = rng.choice(ext.shape[0], counts[2], p=roof_area / roof_area.sum())
    out.append(np.column_stack([scene.box_min[b, 0] + rng.uniform(0, 1, counts[2]) * ext[b, 0],
                                scene.box_min[b, 1] + rng.uniform(0, 1, counts[2]) * ext[b, 1], scene.box_max[b, 2]]))
    p = rng.integers(0, scene.pole_xy.shape[0], counts[3])
    th = rng.uniform(0, 2 * np.pi, counts[3])
    out.append(np.column_stack([scene.pole_xy[p, 0] + scene.pole_r * np.cos(th), scene.pole_xy[p, 1] + scene.pole_r * np.sin(th),
                                rng.uniform(0, scene.pole_h, counts[3])]))
    pts = np.concatenate(out)
    pts += rng.normal(0, noise, pts.shape)
    return pts[rng.permutation(pts.shape[0])]


def keyframes_from_map(map_corner, map_surf, n_keys=12, seed=5):
    """Synthetic key frames for the local-map assembly (mapOptmization.cpp:899-938): chunk k of a map, moved into the sensor
    frame of a random key pose, so that transforming each chunk by its pose and concatenating restores the map.
    Returns a list of (corner (n, 4) f32, surf (n, 4) f32, pose6 f32 = roll, pitch, yaw, x, y, z)."""
    rng = np.random.default_rng(seed)
    keys = []
    cc = np.array_split(map_corner, n_keys)
    ss = np.array_split(map_surf, n_keys)
    for k in range(n_keys):
        pose = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(-3, 3),
                         rng.uniform(-30, 30), rng.uniform(-30, 30), rng.uniform(-0.5, 0.5)], np.float32)
        R = rot_zyx(*pose[:3].astype(np.float64))
        t = pose[3:].astype(np.float64)
        loc = []
        for cloud in (cc[k], ss[k]):
            q = cloud.copy()
            q[:, :3] = ((cloud[:, :3].astype(np.float64) - t) @ R).astype(np.float32)
            loc.append(np.ascontiguousarray(q))
        keys.append((loc[0], loc[1], pose))
    return keys


def keyframes_overlapping(map_corner, map_surf, n_keys=50, frac=0.2, noise=0.02, seed=11):
    """Key frames as LIO-SAM really accumulates them for the local map (mapOptmization.cpp:899-938): consecutive scans see mostly
    the SAME surfaces, so the concatenation the two map VoxelGrids receive is several times the size of their output. Key frame k
    is a random `frac` of the map points (+ N(0, noise) range noise), moved into the sensor frame of a key pose.
    Returns a list of (corner (n, 4) f32, surf (n, 4) f32, pose6 f32)."""
    rng = np.random.default_rng(seed)
    keys = []
    for k in range(n_keys):
        pose = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(-3, 3),
                         rng.uniform(-30, 30), rng.uniform(-30, 30), rng.uniform(-0.5, 0.5)], np.float32)
        R = rot_zyx(*pose[:3].astype(np.float64))
        t = pose[3:].astype(np.float64)
        loc = []
        for cloud in (map_corner, map_surf):
            pick = rng.random(len(cloud)) < frac
            q = cloud[pick].copy()
            w = q[:, :3].astype(np.float64) + rng.normal(0.0, noise, (len(q), 3))
            q[:, :3] = ((w - t) @ R).astype(np.float32)
            loc.append(np.ascontiguousarray(q))
        keys.append((loc[0], loc[1], pose))
    return keys
