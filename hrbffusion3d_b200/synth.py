"""Seeded synthetic RGB-D data: analytic scenes ray-cast with a pinhole
camera, Kinect-style axial noise / drop-outs, procedural texture.  numpy only (host side data
generation; not part of the hot path).

Conventions follow the reference: depth u16 at 5000 units / metre (TUM), K = (528, 528, 320, 240)
at 640x480, camera looks along +z, pose = camera-to-world 4x4 (row-major)."""
import numpy as np

SEED_BASE = 0x48524246  # "HRBF"


def default_camera(width=640, height=480):
    s = width / 640.0
    return (528.0 * s, 528.0 * s, 320.0 * s, 240.0 * s)  # fx, fy, cx, cy


def rot_xyz(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def make_pose(rx=0.0, ry=0.0, rz=0.0, t=(0.0, 0.0, 0.0)):
    T = np.eye(4)
    T[:3, :3] = rot_xyz(rx, ry, rz)
    T[:3, 3] = t
    return T.astype(np.float32)


class Scene:
    """Union of planes (n.X = c) and spheres; 'plane' = single tilted plane, 'room' = box + sphere."""

    def __init__(self, kind="room"):
        self.kind = kind
        if kind == "plane":
            n = rot_xyz(np.deg2rad(15.0), 0, 0) @ np.array([0, 0, 1.0])
            self.planes = [(n, float(n @ np.array([0, 0, 1.5])))]
            self.spheres = []
        else:
            # 4 x 3 x 2.5 m box seen from the inside, camera near the middle looking at +z
            # (the wall behind the start pose, z = -1.8, is only ever seen by the room loop of SURVEY 8d config 3)
            self.planes = [(np.array([0, 0, 1.0]), 2.2), (np.array([1.0, 0, 0]), 2.0), (np.array([-1.0, 0, 0]), 2.0),
                           (np.array([0, 1.0, 0]), 1.2), (np.array([0, -1.0, 0]), 1.3), (np.array([0, 0, -1.0]), 1.8)]
            self.spheres = [(np.array([0.3, 0.6, 1.7]), 0.45), (np.array([-0.8, 0.2, 1.9]), 0.35)]

    def raycast(self, o, D):
        """o (3,), D (...,3) world rays -> (s, normal, albedo) with s = ray parameter (inf if none)."""
        best = np.full(D.shape[:-1], np.inf)
        nrm = np.zeros(D.shape)
        for n, c in self.planes:
            den = D @ n
            with np.errstate(divide="ignore", invalid="ignore"):
                s = (c - o @ n) / den
            ok = (s > 1e-6) & (s < best) & np.isfinite(s)
            best = np.where(ok, s, best)
            nrm[ok] = n
        for c0, r in self.spheres:
            oc = o - c0
            b = D @ oc
            a = np.sum(D * D, -1)
            disc = b * b - a * (oc @ oc - r * r)
            with np.errstate(invalid="ignore"):
                s = (-b - np.sqrt(disc)) / a
            ok = (disc > 0) & (s > 1e-6) & (s < best)
            best = np.where(ok, s, best)
            P = o + D * best[..., None]
            nn = (P - c0) / r
            nrm[ok] = nn[ok]
        return best, nrm

    def texture(self, P):
        """procedural grey-ish RGB in [0,255] from world position"""
        f = 25.0
        a = 0.5 + 0.5 * np.sin(f * P[..., 0] + 1.3) * np.cos(f * 0.8 * P[..., 1] - 0.4) * np.cos(f * 0.7 * P[..., 2])
        b = 0.5 + 0.5 * np.sin(2.3 * f * P[..., 0] * 0.5 + 2.0 * P[..., 2] + 4.0 * P[..., 1])
        r = 30 + 200 * (0.6 * a + 0.4 * b)
        g = 30 + 200 * (0.5 * a + 0.5 * (1 - b))
        bl = 30 + 200 * (0.3 + 0.7 * a * b)
        return np.clip(np.stack([r, g, bl], -1), 1, 254).astype(np.uint8)


def render_depth(scene, pose, width=640, height=480, cam=None, noise=True, seed=0, dropout=0.05, pixel_center=0.0):
    """-> depth u16 [h,w] (5000/m, 0 = invalid), rgb u8 [h,w,3]"""
    fx, fy, cx, cy = cam or default_camera(width, height)
    rng = np.random.default_rng(SEED_BASE + seed)
    u, v = np.meshgrid(np.arange(width) + pixel_center, np.arange(height) + pixel_center)
    if noise:
        u = u + rng.uniform(-0.5, 0.5, u.shape)
        v = v + rng.uniform(-0.5, 0.5, v.shape)
    d = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1)
    R, t = pose[:3, :3].astype(np.float64), pose[:3, 3].astype(np.float64)
    D = d @ R.T
    s, nrm = scene.raycast(t, D)
    z = s.copy()  # d has z == 1 -> camera-frame depth equals the ray parameter
    P = t + D * np.where(np.isfinite(s), s, 0)[..., None]
    rgb = scene.texture(P)
    valid = np.isfinite(z)
    if noise:
        sigma = 0.0012 + 0.0019 * (np.where(valid, z, 1.0) - 0.4) ** 2
        z = z + rng.normal(0, 1, z.shape) * sigma
        ncam = nrm @ R  # world -> camera
        cosang = np.abs(np.sum(ncam * d, -1)) / np.linalg.norm(d, axis=-1)
        valid &= cosang > np.cos(np.deg2rad(70.0))
        valid &= rng.uniform(0, 1, z.shape) > dropout
    valid &= (z > 0.3) & (z < 10.0)
    depth = np.where(valid, np.round(z * 5000.0), 0).astype(np.uint16)
    rgb[~np.isfinite(s)] = 0
    return depth, rgb


def ideal_maps(scene, pose, width=640, height=480, cam=None, seed=0, invalid_frac=0.03, half_pixel=False):
    """Noise-free per-pixel maps in the CAMERA frame, shaped like the reference's RGBA32F textures:
    vertex (xyz, conf), normal (xyz, radius), k1/k2 (dir xyz, value), icp weight, rgba u8.
    A seeded fraction of pixels is invalidated (z = 0) to exercise the NaN paths."""
    fx, fy, cx, cy = cam or default_camera(width, height)
    rng = np.random.default_rng(SEED_BASE + 1000 + seed)
    off = 0.5 if half_pixel else 0.0
    u, v = np.meshgrid(np.arange(width) + off, np.arange(height) + off)
    d = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1)
    R, t = pose[:3, :3].astype(np.float64), pose[:3, 3].astype(np.float64)
    D = d @ R.T
    s, nrm = scene.raycast(t, D)
    valid = np.isfinite(s) & (s > 0.3) & (s < 10)
    valid &= rng.uniform(0, 1, s.shape) > invalid_frac
    sv = np.where(valid, s, 0.0)
    V = d * sv[..., None]
    ncam = nrm @ R
    flip = ncam[..., 2] < 0     # reference convention: normals have n.z >= 0 in the camera frame
    ncam[flip] *= -1
    P = t + D * sv[..., None]
    vertex = np.zeros((height, width, 4), np.float32)
    normal = np.zeros((height, width, 4), np.float32)
    vertex[..., :3] = V
    vertex[..., 3] = np.where(valid, 4.0 + 2.0 * rng.uniform(0, 1, s.shape), 0.0)
    normal[..., :3] = np.where(valid[..., None], ncam, 0.0)
    normal[..., 3] = np.where(valid, 1.41421356 * sv / (0.5 * (fx + fy)) * 4.0, 0.0)
    # smooth pseudo curvature fields (principal directions tangent-ish, values mostly within +-300)
    k1 = np.zeros((height, width, 4), np.float32)
    k2 = np.zeros((height, width, 4), np.float32)
    tx = np.cross(ncam, np.array([0.0, 1.0, 0.0]))
    tx /= np.maximum(np.linalg.norm(tx, axis=-1, keepdims=True), 1e-9)
    ty = np.cross(ncam, tx)
    k1[..., :3], k2[..., :3] = tx, ty
    k1[..., 3] = 3.0 * np.sin(4.0 * P[..., 0]) + 5.0
    k2[..., 3] = 2.0 * np.cos(3.0 * P[..., 1]) - 4.0
    bad = rng.uniform(0, 1, s.shape) < 0.01
    k1[..., 3] = np.where(bad, 1000.0, k1[..., 3])
    k1[~valid] = (0, 0, 0, 1000.0)
    k2[~valid] = (0, 0, 0, 1000.0)
    w = np.where(valid, (1.0 / np.maximum(sv, 1e-3) ** 2) * (vertex[..., 3] / 256.0 + 0.6), 0.0).astype(np.float32)
    rgb = scene.texture(P)
    rgb[~valid] = 0
    rgba = np.concatenate([rgb, np.full((height, width, 1), 255, np.uint8)], -1)
    return {"vertex": vertex, "normal": normal, "k1": k1, "k2": k2, "icpw": w, "rgba": np.ascontiguousarray(rgba)}


def circle_trajectory(n_frames, radius=0.05, yaw_deg=2.0, frames_per_rev=200):
    """5 cm-radius circle with a 2 degree yaw wobble."""
    poses = []
    for i in range(n_frames):
        a = 2 * np.pi * i / frames_per_rev
        poses.append(make_pose(0.0, np.deg2rad(yaw_deg) * np.sin(a), 0.0,
                               (radius * np.cos(a) - radius, radius * np.sin(a), 0.0)))
    return poses


def room_loop_trajectory(n_frames, radius=0.05):
    """a closed loop inside the room -- the camera turns once about the vertical axis (360 / n degrees per
    frame: 0.36 degrees at n = 1000) while it moves on a small circle (<= 1 cm per frame); the last frame closes onto the first,
    so the map first grows around the whole room and then the start of the loop is seen again."""
    poses = []
    for i in range(n_frames):
        a = 2 * np.pi * i / n_frames
        poses.append(make_pose(0.0, a, 0.0, (radius * np.sin(3 * a), 0.02 * np.sin(2 * a), radius * (1 - np.cos(3 * a)))))
    return poses


def _render_one(args):
    kind, pose, W, H, cam, seed = args
    return render_depth(Scene(kind), pose, W, H, cam, noise=True, seed=seed)


def render_sequence(kind, poses, W, H, cam, seed0=0, workers=None):
    """[(depth u16, rgb u8)] for every pose, rendered on `workers` processes (host-side data generation only; call before CUDA is
    initialised in this process -- the pool forks)."""
    import multiprocessing as mp
    import os
    jobs = [(kind, p, W, H, cam, seed0 + i) for i, p in enumerate(poses)]
    workers = workers or min(len(jobs), max(1, (os.cpu_count() or 2) - 1))
    if workers <= 1 or len(jobs) < 4:
        return [_render_one(j) for j in jobs]
    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(_render_one, jobs, chunksize=max(1, len(jobs) // (4 * workers)))


def surfels_from_maps(maps, pose, time=1, submap=0, stride=1):
    """Global-frame surfel records (float32 [n, 20], the reference's 80-B layout, Shaders/Vertex.cpp:20-44) from
    camera-frame ideal maps: {pos.xyz, conf | colour24, submap, initTime, lastTime | n.xyz, radius | k1dir, k1 | k2dir, k2}.
    Pixel order is x outer / y inner like the reference's uv VBO (GlobalModel.cpp:89-96)."""
    V, N, K1, K2, rgba = maps["vertex"], maps["normal"], maps["k1"], maps["k2"], maps["rgba"]
    H, W = V.shape[:2]
    R, t = pose[:3, :3].astype(np.float32), pose[:3, 3].astype(np.float32)
    xs, ys = np.meshgrid(np.arange(0, W, stride), np.arange(0, H, stride), indexing="ij")
    xs, ys = xs.reshape(-1), ys.reshape(-1)
    v, n, k1, k2, c = V[ys, xs], N[ys, xs], K1[ys, xs], K2[ys, xs], rgba[ys, xs].astype(np.int64)
    ok = v[:, 2] > 0
    v, n, k1, k2, c = v[ok], n[ok], k1[ok], k2[ok], c[ok]
    s = np.zeros((v.shape[0], 20), np.float32)
    s[:, 0:3] = v[:, :3] @ R.T + t
    s[:, 3] = v[:, 3]
    s[:, 4] = ((c[:, 0] << 16) + (c[:, 1] << 8) + c[:, 2]).astype(np.float32)
    s[:, 5], s[:, 6], s[:, 7] = submap, time, time
    s[:, 8:11] = n[:, :3] @ R.T
    s[:, 11] = n[:, 3]
    s[:, 12:15] = k1[:, :3] @ R.T
    s[:, 15] = k1[:, 3]
    s[:, 16:19] = k2[:, :3] @ R.T
    s[:, 19] = k2[:, 3]
    return s
