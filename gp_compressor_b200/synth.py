"""Synthetic point clouds of the shapes BASELINE.json / SURVEY.md section 8(d) name (C1..C5).

The reference ships no data (its test mains read private PCD files,
/root/reference/src/test_gp_compress.cpp:14), so every workload is generated here from a
seed.  Clouds use PCL's PointXYZRGB memory layout, 32 bytes per point:
float x, y, z, 1.0f | uint8 b, g, r, a | 12 bytes of padding   ([RECALLED] PCL point_types).
They are returned as a C-contiguous uint8 array of shape (n, 32).
"""
import numpy as np

POINT_BYTES = 32


def pack_cloud(xyz, rgb=None):
    """xyz: (n,3) float; rgb: (n,3) uint8 (r,g,b) or None -> (n,32) uint8."""
    xyz = np.asarray(xyz, dtype=np.float32)
    n = xyz.shape[0]
    out = np.zeros((n, POINT_BYTES), dtype=np.uint8)
    f = out.view(np.float32).reshape(n, 8)
    f[:, 0:3] = xyz
    f[:, 3] = 1.0
    if rgb is not None:
        rgb = np.asarray(rgb, dtype=np.uint8)
        out[:, 16] = rgb[:, 2]
        out[:, 17] = rgb[:, 1]
        out[:, 18] = rgb[:, 0]
    out[:, 19] = 255
    return out


def cloud_xyz(cloud32):
    return cloud32.view(np.float32).reshape(-1, 8)[:, 0:3]


def _colors(x, y):
    r = 128 + 100 * np.sin(x)
    g = 128 + 100 * np.cos(y)
    b = 128 + 100 * np.sin(x + y)
    return np.stack([r, g, b], axis=1).astype(np.uint8)


def c1_planar_bumps(n=100_000, seed=1):
    """C1: planar-with-bumps cloud of the test_gp_compress configuration (res 0.15f, sz 20)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 10, n)
    y = rng.uniform(0, 10, n)
    c = rng.uniform(0, 10, (20, 2))
    z = 0.05 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y / 1.3)
    for j in range(20):
        z = z + 0.1 * np.exp(-((x - c[j, 0]) ** 2 + (y - c[j, 1]) ** 2) / (2 * 0.3 ** 2))
    z = z + rng.normal(0, 0.002, n)
    return pack_cloud(np.stack([x, y, z], axis=1), _colors(x, y))


def _sample_rect(rng, n, origin, u, v, noise):
    """n points uniform on the parallelogram origin + a*u + b*v, Gaussian noise along the normal."""
    a = rng.random(n, dtype=np.float32)
    b = rng.random(n, dtype=np.float32)
    nrm = np.cross(u, v)
    nrm = nrm / np.linalg.norm(nrm)
    d = rng.standard_normal(n, dtype=np.float32) * np.float32(noise)
    p = (np.asarray(origin, np.float32)[None, :] + a[:, None] * np.asarray(u, np.float32)[None, :]
         + b[:, None] * np.asarray(v, np.float32)[None, :] + d[:, None] * nrm.astype(np.float32)[None, :])
    return p


def _box_faces(lo, hi):
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    return [(lo, ex, ey), (lo + ez, ex, ey), (lo, ex, ez), (lo + ey, ex, ez), (lo, ey, ez), (lo + ex, ey, ez)]


def c2_indoor(n=5_000_000, seed=2, noise=0.003):
    """C2: room 10 x 8 x 3 m (floor, ceiling, 4 walls) plus 12 axis-aligned boxes; area-uniform."""
    rng = np.random.default_rng(seed)
    faces = _box_faces([0, 0, 0], [10, 8, 3])
    for _ in range(12):
        s = rng.uniform(0.3, 1.5, 3)
        lo = np.array([rng.uniform(0.2, 10 - 0.2 - s[0]), rng.uniform(0.2, 8 - 0.2 - s[1]), 0.0])
        faces += _box_faces(lo, lo + s)[1:]  # no bottom face
    areas = np.array([np.linalg.norm(np.cross(u, v)) for _, u, v in faces])
    counts = rng.multinomial(n, areas / areas.sum())
    parts = [_sample_rect(rng, int(c), o, u, v, noise) for (o, u, v), c in zip(faces, counts) if c > 0]
    xyz = np.concatenate(parts, axis=0)
    xyz = xyz[rng.permutation(xyz.shape[0])]
    return pack_cloud(xyz, _colors(xyz[:, 0].astype(np.float64), xyz[:, 1].astype(np.float64)))


def c3_dense_floor(n=10_000_000, seed=3, noise=0.003, side=10.0):
    """C3: 100 m^2 of gently undulating surface (about n/10^4 points per 0.1 m patch)."""
    rng = np.random.default_rng(seed)
    x = rng.random(n, dtype=np.float32) * np.float32(side)
    y = rng.random(n, dtype=np.float32) * np.float32(side)
    z = (0.03 * np.sin(3.0 * x) * np.cos(2.0 * y) + 0.01 * np.sin(17.0 * x + 5.0 * y)).astype(np.float32)
    z = z + rng.standard_normal(n, dtype=np.float32) * np.float32(noise)
    xyz = np.stack([x, y, z], axis=1)
    return pack_cloud(xyz, _colors(x.astype(np.float64), y.astype(np.float64)))


def c4_patch_params(n_patches=1_000_000, nbv=30, seed=4, res=float(np.float32(0.1))):
    """C4: decode-only workload. Random fitted parameters for n_patches patches."""
    rng = np.random.default_rng(seed)
    T = n_patches * nbv
    bv1 = rng.uniform(-res / 2, res / 2, T)
    bv2 = rng.uniform(-res / 2, res / 2, T)
    alpha = rng.standard_normal(T)
    q = rng.standard_normal((n_patches, 4))
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    mean = rng.uniform(0, 100, (n_patches, 3))
    rgbmean = rng.uniform(0, 255, (n_patches, 3))
    return dict(nbv=np.full(n_patches, nbv, dtype=np.int32), bv1=bv1, bv2=bv2, alpha=alpha,
                quat=q.reshape(-1), mean=mean.reshape(-1), rgbmean=rgbmean.reshape(-1))


def c5_outdoor(n=50_000_000, seed=5, noise=0.01):
    """C5: LiDAR-like outdoor cloud: ground over r <= 100 m with density ~ 1/r, 40 facades, 200 trunks."""
    rng = np.random.default_rng(seed)
    n_ground = int(n * 0.7)
    n_fac = int(n * 0.2)
    n_tree = n - n_ground - n_fac
    parts = []
    # ground: density ~ 1/r  <=>  r uniform in [1, 100]
    r = rng.uniform(1.0, 100.0, n_ground).astype(np.float32)
    th = rng.uniform(0, 2 * np.pi, n_ground).astype(np.float32)
    x = r * np.cos(th)
    y = r * np.sin(th)
    z = (0.3 * np.sin(x / 15.0) * np.cos(y / 20.0)).astype(np.float32) + rng.standard_normal(n_ground, dtype=np.float32) * np.float32(noise)
    parts.append(np.stack([x, y, z], axis=1))
    # facades: vertical planes 10-30 m wide, 5-15 m tall
    w = rng.uniform(10, 30, 40)
    h = rng.uniform(5, 15, 40)
    cnt = rng.multinomial(n_fac, (w * h) / (w * h).sum())
    for j in range(40):
        ang = rng.uniform(0, 2 * np.pi)
        c = rng.uniform(-80, 80, 2)
        u = np.array([np.cos(ang) * w[j], np.sin(ang) * w[j], 0.0])
        v = np.array([0.0, 0.0, h[j]])
        o = np.array([c[0], c[1], 0.0]) - u / 2
        if cnt[j] > 0:
            parts.append(_sample_rect(rng, int(cnt[j]), o, u, v, noise))
    # tree-like vertical cylinders
    cnt = rng.multinomial(n_tree, np.full(200, 1.0 / 200))
    for j in range(200):
        c = rng.uniform(-90, 90, 2)
        rad = rng.uniform(0.15, 0.5)
        hh = rng.uniform(3, 10)
        m = int(cnt[j])
        if m == 0:
            continue
        a = rng.uniform(0, 2 * np.pi, m).astype(np.float32)
        zz = rng.uniform(0, hh, m).astype(np.float32)
        rr = np.float32(rad) + rng.standard_normal(m, dtype=np.float32) * np.float32(noise)
        parts.append(np.stack([np.float32(c[0]) + rr * np.cos(a), np.float32(c[1]) + rr * np.sin(a), zz], axis=1))
    xyz = np.concatenate(parts, axis=0)
    xyz = xyz[rng.permutation(xyz.shape[0])]
    return pack_cloud(xyz, _colors(xyz[:, 0].astype(np.float64), xyz[:, 1].astype(np.float64)))


# hyper-parameter sets of SURVEY.md section 8(d)
def hyper_ref():
    """The reference defaults: rbf_kernel(100, 1) (rbf_kernel.h:24), s0 = 1e-1f (sparse_gp.h:48)."""
    return dict(sigmaf_sq=100.0, l_sq=1.0, s0=float(np.float32(1e-1)))


def hyper_bind(res):
    """A hyper-set under which the BV capacity binds (length scale of a patch cell)."""
    return dict(sigmaf_sq=1.0, l_sq=float((res / 12.0) ** 2), s0=1e-4)
