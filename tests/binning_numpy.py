"""Second witness for the binning rows (a-1 .. a-5): a brute-force, SEQUENTIAL numpy restatement of
gp_compressor::project_cloud written from the reference text (gp_compressor.cpp:29-118, 177-249) and the PCL semantics
list of SURVEY.md section 8a -- independent of oracle/gpc_oracle.cpp, which restates the same path as an order-free,
data-parallel rule.  It builds the voxel lattice by replaying the bounding-box growth point by point, visits the leaves in
(reverse) Morton order, does the radius search by brute force with float32 arithmetic, takes the plane normal from numpy's
SVD of the m x 4 homogeneous matrix and claims points greedily with an `occupied` array, exactly as the reference loops do.
Small clouds only (O(leaves x points)).  Test infrastructure."""
import numpy as np

F32 = np.float32
EPS = float(np.finfo(np.float32).eps)


def lattice(xyz, res):
    """OctreePointCloud::addPointsFromInputCloud -> adoptBoundingBoxToPoint, finite points in index order."""
    mn = mx = None
    depth = 0
    for p in xyz:
        if not np.all(np.isfinite(p)):
            continue
        p = p.astype(np.float64)
        while True:
            if mn is None:
                mn, mx = p - res / 2.0, p + res / 2.0
                depth = 1                                   # getKeyBitSize(): max(keys, 2) = 2 voxels -> 1 level
                side = (1 << depth) * res - EPS
                over = (side - (mx - mn)) / 2.0
                mn, mx = mn - over, mx + over
                break
            lo, up = p < mn, p >= mx
            if not (lo.any() or up.any()):
                break
            side = (1 << depth) * res
            mn = np.where(up, mn, mn - side)                # every axis that does not violate the upper bound
            depth += 1
            mx = mn + ((1 << depth) * res - EPS)
    return mn, depth


def morton(k):
    code = 0
    for b in range(21):
        code |= ((int(k[0]) >> b) & 1) << (3 * b + 2) | ((int(k[1]) >> b) & 1) << (3 * b + 1) | ((int(k[2]) >> b) & 1) << (3 * b)
    return code


def rotation(pts4):
    """gp_compressor::compute_rotation, gp_compressor.cpp:29-64 (pts4: m x 4 homogeneous rows)."""
    if pts4.shape[0] < 4:
        return np.eye(3)
    _, _, vt = np.linalg.svd(pts4, full_matrices=False)
    n = vt[3, :3] / np.linalg.norm(vt[3, :3])
    a = np.abs(n)
    if a[0] > a[1] and a[0] > a[2]:
        n = -n if n[0] < 0 else n
        c1 = np.cross([0.0, 0.0, 1.0], n)
    elif a[1] > a[0] and a[1] > a[2]:
        n = -n if n[1] < 0 else n
        c1 = np.cross([1.0, 0.0, 0.0], n)
    else:
        n = -n if n[2] < 0 else n
        c1 = np.cross([0.0, 1.0, 0.0], n)
    c1 = c1 / np.linalg.norm(c1)
    return np.stack([n, c1, np.cross(n, c1)], axis=1)


def project_cloud(xyz, res, leaf_order=0, ref=None, rgb=None, sz=10):
    """xyz: n x 3 float32.  Returns lattice, leaves in visiting order and the greedy assignment.
    ref: oracle.ref_source -- when given, the frame of every leaf and its claim come from THE REFERENCE'S OWN
    compute_rotation and project_points (gp_compressor.cpp:29-118 compiled by oracle/ref_build.py) instead of the numpy
    restatement below; rgb: n x 3 colours (r, g, b) handed to project_points."""
    res = float(res)
    mn, depth = lattice(xyz, res)
    finite = np.isfinite(xyz).all(axis=1)
    idx = np.nonzero(finite)[0]
    keys = np.floor((xyz[idx].astype(np.float64) - mn) / res).astype(np.int64)      # (unsigned)((double(p) - min) / res)
    codes = np.array([morton(k) for k in keys], dtype=np.uint64)
    leaf_codes = np.unique(codes)                                                    # ascending Morton
    if leaf_order == 0:
        leaf_codes = leaf_codes[::-1]                                                # PCL 1.7/1.8 depth-first iterator
    first_key = {}
    for k, c in zip(keys, codes):
        first_key.setdefault(int(c), k)
    # search order inside radiusSearch: ascending Morton of the point's own voxel, then ascending index
    order = idx[np.lexsort((idx, codes))]
    radius = float(np.sqrt(F32(3.0)) / F32(2.0)) * res
    r2 = radius * radius
    occupied = np.zeros(xyz.shape[0], dtype=bool)
    occ32 = np.zeros(xyz.shape[0], dtype=np.int32)
    if rgb is None:
        rgb = np.zeros((xyz.shape[0], 3))
    owner = np.full(xyz.shape[0], -1, dtype=np.int32)
    out = dict(lattice_min=mn, depth=depth, leaf_code=leaf_codes, ncand=[], R=[], stream=[], x1=[], x2=[], y=[], center=[])
    P = xyz[order]
    for i, c in enumerate(leaf_codes):
        k = first_key[int(c)]
        centre = ((k.astype(np.float64) + 0.5) * res + mn).astype(np.float32)       # genLeafNodeCenterFromOctreeKey -> float
        d = P - centre                                                               # float32, getVector3fMap
        # float32 squaredNorm of a fixed-size 3-vector: Eigen's unrolled reduction splits in halves, x^2 + (y^2 + z^2)
        # [RECALLED Eigen 3 redux_novec_unroller]; one candidate in ~10^5 sits close enough to the radius to notice
        d2 = d[:, 0] * d[:, 0] + (d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        cand = order[d2.astype(np.float64) <= r2]
        out["ncand"].append(cand.size)
        pts = xyz[cand].astype(np.float64)
        mid = centre.astype(np.float64)
        if ref is not None:
            R = ref.compute_rotation(pts)                                            # gp_compressor.cpp:237
            pr = ref.project_points(res, sz, pts, rgb[cand], cand.astype(np.int32), occ32, R, mid)   # :239
            cl = cand[pr["claimed"]]
            owner[cl] = i
            out["R"].append(R); out["stream"].append(cl.astype(np.int64))
            out["x1"].append(pr["local"][:, 1]); out["x2"].append(pr["local"][:, 2]); out["y"].append(pr["local"][:, 0])
            out["center"].append(pr["center"] if cl.size else mid)
            out.setdefault("colour", []).append(pr["colour"]); out.setdefault("rgb_mean", []).append(pr["rgb_mean"])
            continue
        R = rotation(np.concatenate([pts, np.ones((cand.size, 1))], axis=1))
        claimed, h, a1, a2 = [], [], [], []
        for m, j in enumerate(cand):                                                 # project_points, :78-100
            if occupied[j]:
                continue
            pt = R.T @ (pts[m] - mid)
            half = res / 2.0
            if pt[1] > half or pt[1] < -half or pt[2] > half or pt[2] < -half:
                continue
            occupied[j] = True
            owner[j] = i
            claimed.append(j); h.append(pt[0]); a1.append(pt[1]); a2.append(pt[2])
        mean = (np.sum(h) / len(h)) if h else np.nan
        out["R"].append(R); out["stream"].append(np.array(claimed, dtype=np.int64))
        out["x1"].append(np.array(a1)); out["x2"].append(np.array(a2)); out["y"].append(np.array(h) - mean)
        out["center"].append(mid + mean * R[:, 0] if h else mid)
    out["owner"] = owner
    return out
