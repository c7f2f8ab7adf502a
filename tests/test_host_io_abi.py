"""CPU tests (no GPU): host-side code of the product (readers, partitioning) and the C-ABI surface.

The product library is LOADED here (its host functions run on the CPU) but no compute entry point is
called -- those need a CUDA device and have no fallback.
"""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle
import psba_b200
from util import HERE, data_file, dataset_paths

ROOT = os.path.dirname(HERE)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "psba_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(psba_[a-zA-Z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 40
    L = psba_b200.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.psba_version()


def test_library_has_sm100a_code_only():
    """the shipped kernels are sm_100a SASS (no PTX-JIT fallback, no other architectures)"""
    out = subprocess.run(["cuobjdump", "-lelf", psba_b200.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


@pytest.mark.parametrize("key", ["7", "9", "54", "54KD", "T21"])
def test_product_reader_matches_oracle_reader(key):
    c, p, cnp = dataset_paths(key)
    a = psba_b200.read_sba(c, p, cnp)
    b = oracle.read_sba(c, p, cnp, kind="restatement")
    assert (a["m"], a["n"], a["o"]) == (b["m"], b["n"], b["o"])
    for k in ("K", "initrot", "cams", "pts", "impts", "iidx", "jidx"):
        assert np.array_equal(a[k], b[k]), k        # bit-exact: same text -> same doubles, same index lists


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("key", ["7", "54KD", "T21"])
def test_product_reader_matches_reference_reader(key):
    """against readInitialSBAEstimate + quat2vec + generate_idxs compiled from the reference sources"""
    c, p, cnp = dataset_paths(key)
    a = psba_b200.read_sba(c, p, cnp)
    b = oracle.read_sba(c, p, cnp, kind="reference")
    for k in ("K", "initrot", "cams", "pts", "impts", "iidx", "jidx"):
        assert np.array_equal(a[k], b[k]), k
    # the reference's dense tables say the same thing as our lists
    blk = b["blk_idx"]
    assert np.array_equal(blk[a["iidx"], a["jidx"]], np.arange(a["o"]))
    assert int((blk >= 0).sum()) == a["o"]


def test_varKD_equals_varK():
    """the 17-column file carries five zero distortion coefficients; results must equal varK"""
    a = psba_b200.read_sba(*dataset_paths("54"))
    b = psba_b200.read_sba(*dataset_paths("54KD"))
    for k in ("K", "initrot", "cams", "pts", "impts", "iidx", "jidx"):
        assert np.array_equal(a[k], b[k]), k


def test_seven_column_cameras_need_default_K():
    c, p = data_file("7cams.txt"), data_file("7pts.txt")
    with pytest.raises(RuntimeError):
        psba_b200.read_sba(c, p, 6)
    K = [851.57945, 330.24755, 262.19500, 1.00169, 0.0]          # SURVEY F6 (data/7camsvarK.txt)
    a = psba_b200.read_sba(c, p, 6, Kdefault=K)
    b = psba_b200.read_sba(*dataset_paths("7"))
    assert np.allclose(a["K"], b["K"]) and np.array_equal(a["impts"], b["impts"])
    assert np.allclose(a["initrot"], b["initrot"], atol=1e-6) and np.allclose(a["cams"], b["cams"], atol=1e-6)


def test_reader_edge_cases(tmp_path):
    cams = tmp_path / "c.txt"
    pts = tmp_path / "p.txt"
    # (the reference only recognises a camera-file comment at the start of the file or after another
    #  comment: after a data record its fgetc sees the newline, readparams.cpp:203-219; the points loader
    #  consumes the newline, :418, so comments may appear between point lines)
    cams.write_text("# comment\n# another\n1000 0 0 1 0  1 0 0 0  0 0 5\n900 1 2 1 0  0.9 0.1 0 0  1 2 3\n")
    # frames out of order in the second point; comment line in between
    pts.write_text("0 0 1 2 0 10 20 1 30 40\n# c\n1 1 2 2 1 5 6 0 7 8\n")
    a = psba_b200.read_sba(str(cams), str(pts), 11)
    assert (a["m"], a["n"], a["o"]) == (2, 2, 4)
    assert a["jidx"].tolist() == [0, 1, 0, 1]                 # generate_idxs: cameras ascending
    # documented deviation: the reference leaves the image points in FILE order while its indices ascend
    # (readparams.cpp:332-423 vs misc.cpp:189-197), which attaches (5,6) to camera 0 although the file gives it to
    # camera 1; the product sorts (frame, x, y) together.  Every shipped file lists its frames ascending.
    assert a["impts"].tolist() == [[10, 20], [30, 40], [7, 8], [5, 6]]
    b = oracle.read_sba(str(cams), str(pts), 11)
    assert b["impts"].tolist() == [[10, 20], [30, 40], [5, 6], [7, 8]]   # the reference's reader: file order
    for k in ("K", "initrot", "cams", "pts", "iidx", "jidx"):
        assert np.array_equal(a[k], b[k]), k
    q = a["initrot"][1]
    assert abs(np.dot(q, q) - 1.0) < 1e-15 and q[0] > 0
    # the product reader is line based and also accepts a comment between camera lines
    cams.write_text("1000 0 0 1 0  1 0 0 0  0 0 5\n# mid\n900 1 2 1 0  0.9 0.1 0 0  1 2 3\n")
    a2 = psba_b200.read_sba(str(cams), str(pts), 11)
    assert np.array_equal(a2["K"], a["K"]) and np.array_equal(a2["initrot"], a["initrot"])
    # wrong column count on the first line -> error code, no crash (readparams.cpp:189-192)
    cams.write_text("1000 0 0 1 0  1 0 0 0  0 0\n")
    with pytest.raises(RuntimeError):
        psba_b200.read_sba(str(cams), str(pts), 11)
    # frame index beyond the camera count (readparams.cpp:366-370)
    cams.write_text("1000 0 0 1 0  1 0 0 0  0 0 5\n")
    with pytest.raises(RuntimeError):
        psba_b200.read_sba(str(cams), str(pts), 11)
    # covariances are detected from the first line and skipped
    cams.write_text("1000 0 0 1 0  1 0 0 0  0 0 5\n")
    pts.write_text("0 0 1 1 0 10 20 1 0 0 1\n")
    a = psba_b200.read_sba(str(cams), str(pts), 11)
    assert a["impts"].tolist() == [[10, 20]] and a["o"] == 1


def test_quat2vec_matches():
    L = psba_b200.lib()
    rng = np.random.default_rng(0)
    for nin in (7, 12, 17):
        x = rng.normal(size=nin)
        a = np.zeros(nin - 1); b = np.zeros(nin - 1)
        L.psba_quat2vec(psba_b200._d(x), nin, psba_b200._d(a), nin - 1)
        oracle.lib().orc_quat2vec(oracle._d(x), nin, oracle._d(b), nin - 1)
        assert np.array_equal(a, b)
        q = x[nin - 7:nin - 3]
        v = a[nin - 7:nin - 4]
        assert abs(np.linalg.norm(v) ** 2 + (q[0] / np.linalg.norm(q)) ** 2 - 1) < 1e-14


def test_synthetic_problem_through_the_text_reader(tmp_path):
    """a text dump of the synthetic generator goes through the reader as a format check"""
    from psba_b200 import synth
    prob = synth.ring_problem(m=12, n=300, d=4, w=8, seed=7)
    c, p = str(tmp_path / "c.txt"), str(tmp_path / "p.txt")
    synth.write_sba_text(prob, c, p)
    a = psba_b200.read_sba(c, p, 11)
    assert np.array_equal(a["iidx"], prob["iidx"]) and np.array_equal(a["jidx"], prob["jidx"])
    assert np.allclose(a["impts"], prob["impts"], rtol=0, atol=0)
    assert np.allclose(a["pts"], prob["pts"], rtol=0, atol=0)
    assert np.allclose(a["initrot"], prob["initrot"], atol=1e-15)
    assert np.all(np.diff(a["jidx"].reshape(-1, 4), axis=1) > 0)


def test_local_range_partitions_points_by_observation_count():
    rng = np.random.default_rng(1)
    n = 1000
    d = rng.integers(1, 9, n)
    iidx = np.repeat(np.arange(n, dtype=np.int32), d)
    o = int(d.sum())
    for R in (1, 2, 3, 8):
        prev_p, prev_o, sizes = 0, 0, []
        for r in range(R):
            p0, p1, o0, o1 = psba_b200.local_range(n, o, iidx, r, R)
            assert p0 == prev_p and o0 == prev_o and p1 >= p0
            assert o1 - o0 == int(d[p0:p1].sum())
            prev_p, prev_o = p1, o1
            sizes.append(o1 - o0)
        assert prev_p == n and prev_o == o
        assert max(sizes) - min(sizes) <= 2 * 8          # balanced to within one point's track


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(psba_b200, "_lib", None)
    monkeypatch.setattr(psba_b200, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        psba_b200.lib()


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under psba_b200/ may reference it"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "psba_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"import oracle|from oracle|liboracle|psba_oracle\.h|orc_[a-z]", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_result_writer_round_trip(tmp_path):
    """refined parameters written as SBA text (rotations recomposed by vec2quat) and read
    back by the reader describe the same cameras: the oracle's reprojection cost of the re-read problem equals
    the cost of the parameters that were written."""
    c, p, cnp = dataset_paths("7")
    prob = psba_b200.read_sba(c, p, cnp)
    rng = np.random.default_rng(3)
    cams = prob["cams"].copy()
    cams[:, :3] = rng.normal(0, 0.02, (prob["m"], 3))          # a refined local rotation
    cams[:, 3:] += rng.normal(0, 0.01, (prob["m"], 3))
    pts = prob["pts"] + rng.normal(0, 0.01, prob["pts"].shape)
    moved = dict(prob, cams=cams, pts=pts)
    O = oracle.Problem(moved)
    cost = O.call("exQT")
    O.close()
    co, po, ply = str(tmp_path / "c_out.txt"), str(tmp_path / "p_out.txt"), str(tmp_path / "r.ply")
    psba_b200.write_result(prob, cams, pts, co, po, ply)
    back = psba_b200.read_sba(co, po, 11)
    assert back["m"] == prob["m"] and back["n"] == prob["n"] and back["o"] == prob["o"]
    assert np.array_equal(back["iidx"], prob["iidx"]) and np.array_equal(back["jidx"], prob["jidx"])
    assert np.array_equal(back["impts"], prob["impts"]) and np.array_equal(back["pts"], pts)
    assert np.all(back["cams"][:, :3] == 0) and np.array_equal(back["cams"][:, 3:], cams[:, 3:])
    O2 = oracle.Problem(back)
    cost2 = O2.call("exQT")
    O2.close()
    assert abs(cost2 - cost) / cost < 1e-12
    # vec2quat is the operation of compute_exQT.cl:46-49 and gives unit quaternions
    q = np.zeros(4)
    psba_b200.lib().psba_vec2quat(psba_b200._d(np.ascontiguousarray(prob["initrot"][0])), psba_b200._d(np.ascontiguousarray(cams[0, :3])), psba_b200._d(q))
    assert abs(np.linalg.norm(q) - 1) < 1e-14
    head = open(ply).read().split("end_header")[0]
    assert "element vertex %d" % (prob["n"] + prob["m"]) in head
    assert len(open(ply).read().strip().splitlines()) == 10 + prob["n"] + prob["m"]


def _rodrigues(r):
    th = np.linalg.norm(r)
    if th < 1e-300:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def test_native_bal_loader(tmp_path):
    """a BAL problem-*.txt (Rodrigues, p = -P/P.z, observations in arbitrary order) is converted
    to PSBA's quaternion / pinhole form.  The oracle's reprojection cost of the loaded problem must equal the BAL
    cost computed here from the BAL parameters with the BAL camera model (k1 = k2 = 0)."""
    rng = np.random.default_rng(11)
    m, n = 6, 40
    rv = rng.normal(0, 0.4, (m, 3)); rv[0] = 0.0                 # one identity rotation
    rv[1] = np.array([np.pi - 1e-3, 0.0, 0.0])                   # close to 180 degrees: exercises every quaternion branch
    t = rng.normal(0, 0.3, (m, 3)); f = rng.uniform(500, 900, m); kc = np.zeros((m, 2))
    X = rng.normal(0, 1.0, (n, 3)) + np.array([0, 0, -8.0])      # in front of BAL cameras (negative z)
    obs = []
    for i in range(n):
        for j in rng.choice(m, size=rng.integers(2, m + 1), replace=False):
            P = _rodrigues(rv[j]) @ X[i] + t[j]
            p = -P[:2] / P[2]
            obs.append((int(j), i, *(f[j] * p + rng.normal(0, 0.5, 2))))
    order = rng.permutation(len(obs))                            # BAL files need not be sorted
    path = str(tmp_path / "problem-6-40-pre.txt")
    with open(path, "w") as fh:
        fh.write("%d %d %d\n" % (m, n, len(obs)))
        for q in order:
            fh.write("%d %d %.17e %.17e\n" % obs[q])
        for j in range(m):
            for v in list(rv[j]) + list(t[j]) + [f[j], kc[j, 0], kc[j, 1]]:
                fh.write("%.17e\n" % v)
        for i in range(n):
            for v in X[i]:
                fh.write("%.17e\n" % v)
    prob = psba_b200.read_bal(path)
    assert (prob["m"], prob["n"], prob["o"]) == (m, n, len(obs))
    key = prob["iidx"].astype(np.int64) * m + prob["jidx"]
    assert np.all(np.diff(key) > 0)                              # point-major, cameras ascending
    assert np.array_equal(prob["K"], np.stack([f, 0 * f, 0 * f, 0 * f + 1, 0 * f], axis=1))
    assert np.all(prob["initrot"][:, 0] >= 0) and np.allclose(np.linalg.norm(prob["initrot"], axis=1), 1, atol=1e-14)
    bal_cost = 0.0
    for (j, i, x, y) in obs:
        P = _rodrigues(rv[j]) @ X[i] + t[j]
        e = np.array([x, y]) - f[j] * (-P[:2] / P[2])
        bal_cost += e @ e
    O = oracle.Problem(prob)
    cost = O.call("exQT")
    O.close()
    assert abs(cost - bal_cost) / bal_cost < 1e-10
    with open(path, "w") as fh:
        fh.write("2 2 1\n0 5 1.0 1.0\n")
    with pytest.raises(RuntimeError):
        psba_b200.read_bal(path)


def test_cpp_host_program_is_built_and_links_against_the_header():
    """psba_b200/bin/psba_main (PSBA/main.cpp:70-231 as a g++ consumer of include/psba_b200.h) exists after build() and
    prints its usage without touching a GPU"""
    import subprocess
    exe = os.path.join(ROOT, "psba_b200", "bin", "psba_main")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
