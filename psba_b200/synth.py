"""Seeded synthetic bundle-adjustment problems in the array layouts the SBA reader produces
(SURVEY.md 8(d)): harness code for bench.py and the tests, not part of the engine.

ring_problem: m cameras on a circle of radius 40 in the XZ-plane looking at the origin (fu=1500,
u0=v0=0, ar=1, s=0), points uniform in a ball of radius 10, every point observed by exactly d distinct
cameras inside a window of w consecutive cameras (mod m) whose start is uniform.  Observation =
reference projection (CL_files/compute_exQT.cl:68-69) + N(0, 0.5^2) px; initial points = truth +
N(0, (0.01*depth)^2); cameras start at the truth (+ optional translation noise).  The reference itself
cannot hold the 2000-camera instance (dense comm3DIdx would be 14.6 TB).
"""
import numpy as np


def ring_problem(m=2000, n=1_000_000, d=5, w=64, seed=20262000, cam_noise=0.0, chunk=200_000):
    rng = np.random.Generator(np.random.PCG64(seed))
    w = min(w, m)
    theta = 2.0 * np.pi * np.arange(m) / m
    phi = theta + 0.5 * np.pi                       # world->camera rotation = R_y(phi)
    q = np.stack([np.cos(0.5 * phi), np.zeros(m), np.sin(0.5 * phi), np.zeros(m)], axis=1)
    q[q[:, 0] < 0] *= -1.0                          # quat2vec: scalar part >= 0 (PSBA/misc.cpp:21-49)
    v = q[:, 1:4]
    initrot = np.concatenate([np.sqrt(1.0 - (v * v).sum(1, keepdims=True)), v], axis=1)  # readparams.cpp:222-226
    K = np.tile(np.array([1500.0, 0.0, 0.0, 1.0, 0.0]), (m, 1))
    cams = np.zeros((m, 6))
    cams[:, 5] = 40.0
    if cam_noise > 0:
        cams[:, 3:6] += rng.normal(0.0, cam_noise, (m, 3))
    # rotation matrices M(q) of the true cameras
    s0, x, y, z = initrot[:, 0], initrot[:, 1], initrot[:, 2], initrot[:, 3]
    R = np.empty((m, 3, 3))
    R[:, 0, 0] = s0 * s0 + x * x - y * y - z * z; R[:, 0, 1] = 2 * (x * y - s0 * z); R[:, 0, 2] = 2 * (x * z + s0 * y)
    R[:, 1, 0] = 2 * (x * y + s0 * z); R[:, 1, 1] = s0 * s0 - x * x + y * y - z * z; R[:, 1, 2] = 2 * (y * z - s0 * x)
    R[:, 2, 0] = 2 * (x * z - s0 * y); R[:, 2, 1] = 2 * (y * z + s0 * x); R[:, 2, 2] = s0 * s0 - x * x - y * y + z * z
    t_true = np.array([0.0, 0.0, 40.0])

    pts_true = np.empty((n, 3))
    pts0 = np.empty((n, 3))
    jidx = np.empty(n * d, dtype=np.int32)
    impts = np.empty((n * d, 2))
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        c = b - a
        dirs = rng.normal(size=(c, 3))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        X = dirs * (10.0 * rng.random(c) ** (1.0 / 3.0))[:, None]
        pts_true[a:b] = X
        start = rng.integers(0, m, c)
        off = np.argsort(rng.random((c, w), dtype=np.float32), axis=1)[:, :d]     # d distinct offsets in the window
        cj = np.sort((start[:, None] + off) % m, axis=1).astype(np.int32)          # cameras ascending (misc.cpp:191)
        jidx[a * d:b * d] = cj.ravel()
        Xc = np.einsum("oij,oj->oi", R[cj.ravel()], np.repeat(X, d, axis=0)) + t_true
        fu = 1500.0
        px = (fu * Xc[:, 0] + 0.0 * Xc[:, 1] + 0.0 * Xc[:, 2]) / Xc[:, 2]
        py = (fu * 1.0 * Xc[:, 1] + 0.0 * Xc[:, 2]) / Xc[:, 2]
        impts[a * d:b * d, 0] = px + rng.normal(0.0, 0.5, c * d)
        impts[a * d:b * d, 1] = py + rng.normal(0.0, 0.5, c * d)
        depth = Xc[:, 2].reshape(c, d).mean(axis=1)
        pts0[a:b] = X + rng.normal(size=(c, 3)) * (0.01 * depth)[:, None]
    iidx = np.repeat(np.arange(n, dtype=np.int32), d)
    import os
    if os.environ.get("PSBA_SYNTH_SORT"):
        # EXPERIMENT ONLY (memory-locality probe, never the bench default): points ordered by their first camera
        order = np.argsort(jidx.reshape(n, d)[:, 0], kind="stable")
        pts_true, pts0 = pts_true[order], pts0[order]
        jidx = jidx.reshape(n, d)[order].ravel()
        impts = impts.reshape(n, d, 2)[order].reshape(n * d, 2)
    return dict(m=m, n=n, o=n * d, K=K, initrot=initrot, cams=cams, pts=pts0, impts=impts, iidx=iidx, jidx=jidx,
                pts_true=pts_true, name="ring-m%d-n%d-o%d-w%d" % (m, n, n * d, w), window=w, seed=seed)


def write_sba_text(prob, cams_path, pts_path, max_points=None):
    """Dump a problem in the SBA text format (12-column cameras: fu u0 v0 ar s q0 qx qy qz tx ty tz) so that
    the text loaders can be exercised on synthetic data."""
    m, n = prob["m"], prob["n"] if max_points is None else min(prob["n"], max_points)
    with open(cams_path, "w") as f:
        for j in range(m):
            f.write(" ".join("%.17g" % x for x in list(prob["K"][j]) + list(prob["initrot"][j]) + list(prob["cams"][j, 3:6])) + "\n")
    ptr = np.zeros(prob["n"] + 1, dtype=np.int64)
    np.add.at(ptr, prob["iidx"] + 1, 1)
    ptr = np.cumsum(ptr)
    with open(pts_path, "w") as f:
        for i in range(n):
            a, b = ptr[i], ptr[i + 1]
            parts = ["%.17g %.17g %.17g" % tuple(prob["pts"][i]), str(b - a)]
            for k in range(a, b):
                parts.append("%d %.17g %.17g" % (prob["jidx"][k], prob["impts"][k, 0], prob["impts"][k, 1]))
            f.write(" ".join(parts) + "\n")


BAL_OBS = {"Ladybug-138-19878": 85217, "Venice-52-64053": 347173, "Dubrovnik-88-64298": 383937}


def _quat2vec_rows(raw):
    """PSBA/misc.cpp:21-49 + readparams.cpp:222-226 on a (m, 12) camera table -> K, initrot, t"""
    q = raw[:, 5:9]
    mag = np.sqrt((q * q).sum(1))
    sg = np.where(q[:, 0] >= 0.0, 1.0, -1.0) / mag
    v = q[:, 1:4] * sg[:, None]
    initrot = np.concatenate([np.sqrt(1.0 - (v * v).sum(1, keepdims=True)), v], axis=1)
    return raw[:, 0:5].copy(), initrot, raw[:, 9:12].copy()


def _rotmats(q):
    s0, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((len(q), 3, 3))
    R[:, 0, 0] = s0 * s0 + x * x - y * y - z * z; R[:, 0, 1] = 2 * (x * y - s0 * z); R[:, 0, 2] = 2 * (x * z + s0 * y)
    R[:, 1, 0] = 2 * (x * y + s0 * z); R[:, 1, 1] = s0 * s0 - x * x + y * y - z * z; R[:, 1, 2] = 2 * (y * z - s0 * x)
    R[:, 2, 0] = 2 * (x * z - s0 * y); R[:, 2, 1] = 2 * (y * z + s0 * x); R[:, 2, 2] = s0 * s0 - x * x - y * y + z * z
    return R


def bal_structure_problem(cams_txt, n, o, seed=None, name="bal-synth"):
    """Synthetic structure on REAL shipped BAL cameras: the six BAL `-pts.txt` files are
    missing from the reference checkout, so points and observations are generated on the shipped
    `*-cams.txt`:  track lengths d_i = 2 + Poisson(o/n - 2) clipped to [2, m] with sum exactly o; first
    camera uniform; X_i = back-projection of a pixel (u,v) ~ U(-0.4 fu, 0.4 fu)^2 at depth z ~ logU(2, 20);
    the other d_i - 1 cameras uniform among those that see X_i (depth > 0.1, |u|,|v| <= fu), cameras of a
    track ascending; observation = reference projection + N(0, 0.5^2) px; initial point = truth +
    N(0, (0.01 z)^2); initial cameras = shipped.  Label results "synthetic structure on real BAL cameras"."""
    raw = np.loadtxt(cams_txt)
    m = raw.shape[0]
    rng = np.random.Generator(np.random.PCG64(20260000 + m if seed is None else seed))
    K, initrot, t = _quat2vec_rows(raw)
    R = _rotmats(initrot)
    fu, ar = K[:, 0], K[:, 3]
    # track lengths
    d = 2 + rng.poisson(max(o / n - 2.0, 0.0), n)
    d = np.clip(d, 2, m)
    diff = o - int(d.sum())
    i = n - 1
    while diff != 0:
        step = 1 if diff > 0 else -1
        if 2 <= d[i] + step <= m:
            d[i] += step; diff -= step
        i = i - 1 if i > 0 else n - 1
    X = np.empty((n, 3)); zdepth = np.empty(n); first = np.empty(n, dtype=np.int64)
    cand = np.zeros((n, m), dtype=bool)
    todo = np.arange(n)
    for _ in range(200):
        c = len(todo)
        if c == 0:
            break
        a = rng.integers(0, m, c)
        u = rng.uniform(-0.4, 0.4, c) * fu[a]; v = rng.uniform(-0.4, 0.4, c) * fu[a]
        z = np.exp(rng.uniform(np.log(2.0), np.log(20.0), c))
        Xc = np.stack([u / fu[a] * z, v / (fu[a] * ar[a]) * z, z], axis=1)
        Xw = np.einsum("cji,cj->ci", R[a], Xc - t[a])                      # R^T (Xc - t)
        P = np.einsum("mij,cj->cmi", R, Xw) + t[None, :, :]               # all cameras
        dep = P[:, :, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            uu = fu[None, :] * P[:, :, 0] / dep; vv = fu[None, :] * ar[None, :] * P[:, :, 1] / dep
        ok = (dep > 0.1) & (np.abs(uu) <= fu[None, :]) & (np.abs(vv) <= fu[None, :])
        ok[np.arange(c), a] = False
        good = ok.sum(1) >= d[todo] - 1
        g = todo[good]
        X[g] = Xw[good]; zdepth[g] = z[good]; first[g] = a[good]; cand[g] = ok[good]
        todo = todo[~good]
    if len(todo):
        raise RuntimeError("bal_structure_problem: %d points could not be placed" % len(todo))
    # choose the other d_i - 1 cameras uniformly among the candidates
    key = rng.random((n, m))
    key[~cand] = 2.0
    order = np.argsort(key, axis=1)
    take = np.arange(m)[None, :] < (d - 1)[:, None]
    sel = np.zeros((n, m), dtype=bool)
    rows = np.repeat(np.arange(n), m).reshape(n, m)
    sel[rows[take], order[take]] = True
    sel[np.arange(n), first] = True
    iidx, jidx = np.nonzero(sel)                                            # point-major, cameras ascending
    iidx = iidx.astype(np.int32); jidx = jidx.astype(np.int32)
    assert len(iidx) == o
    Xc = np.einsum("oij,oj->oi", R[jidx], X[iidx]) + t[jidx]
    px = (K[jidx, 0] * Xc[:, 0] + K[jidx, 4] * Xc[:, 1] + K[jidx, 1] * Xc[:, 2]) / Xc[:, 2]
    py = (K[jidx, 0] * K[jidx, 3] * Xc[:, 1] + K[jidx, 2] * Xc[:, 2]) / Xc[:, 2]
    impts = np.stack([px, py], axis=1) + rng.normal(0.0, 0.5, (o, 2))
    pts0 = X + rng.normal(size=(n, 3)) * (0.01 * zdepth)[:, None]
    cams = np.zeros((m, 6)); cams[:, 3:6] = t
    return dict(m=m, n=n, o=o, K=K, initrot=initrot, cams=cams, pts=pts0, impts=impts, iidx=iidx, jidx=jidx,
                pts_true=X, name=name, label="synthetic structure on real BAL cameras")
