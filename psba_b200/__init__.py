"""psba_b200 -- B200-native (sm_100a CUDA, FP64) engine behind the host API of eglrp/PSBA.

This package is a thin ctypes binding over ``libpsba_b200.so`` (C ABI: include/psba_b200.h).  The
product is the shared library (hand-written CUDA kernels + C++ host drivers); Python is only the
harness that tests and bench.py drive it with.  There is no CPU fallback: importing works anywhere,
but creating a problem without the built library or without a CUDA device raises / aborts.

Method names follow the reference's operator API (PSBA/sba_func.h:10-138, cl_spdinv.h, cl_cholmod.h,
cl_linearalg.h) and drivers (PSBA/levmar.h, trust_region.h).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpsba_b200.so")
_lib = None

ITER_NAMES = {1: "TURN_TO_LM", 2: "TURN_TO_TR", 3: "CONTINUE", 4: "ERR", 5: "DP_NO_CHANGE",
              6: "ERR_SMALL_ENOUGH", 7: "PASS"}
PARAMS_CUR, PARAMS_NEW = 0, 1
VEC_G, VEC_DP = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class TraceRec(C.Structure):
    _fields_ = [("phase", C.c_int), ("itno", C.c_int), ("err", C.c_double), ("rho", C.c_double),
                ("mu", C.c_double), ("delta", C.c_double), ("pnorm", C.c_double), ("accepted", C.c_int)]


class TryResult(C.Structure):
    _fields_ = [("cost_new", C.c_double), ("dp_L2", C.c_double), ("dp_dot", C.c_double), ("solve_status", C.c_double), ("p_new_L2", C.c_double)]


def build():
    """Compile libpsba_b200.so in-tree with nvcc for sm_100a (psba_b200/csrc/Makefile)."""
    import subprocess
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], check=True, capture_output=True)


def lib():
    """Load the CUDA library; fail loudly if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("psba_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    i, d, vp = C.c_int, C.c_double, C.c_void_p
    dims = [vp, i, i, i, i, i, i]
    L.psba_setup_cl.restype = vp
    L.psba_setup_cl.argtypes = [i, i, i, i, i, i]
    L.psba_fill_initBuffer2.argtypes = dims + [_dp, _dp, _dp, _dp, _dp]
    L.psba_fill_idxBuffer.argtypes = [vp, i, i, i, _ip, _ip]
    L.psba_release_buffer.argtypes = [vp]
    L.psba_host_alloc.restype = vp
    L.psba_host_alloc.argtypes = [C.c_size_t]
    L.psba_host_free.argtypes = [vp]
    L.psba_compute_exQT.restype = d
    L.psba_compute_exQT.argtypes = dims + [i, _dp]
    L.psba_compute_jacobiQT.argtypes = dims + [_dp, _dp]
    for f in ("psba_compute_U", "psba_compute_V", "psba_compute_g"):
        getattr(L, f).argtypes = dims + [d, _dp]
    L.psba_compute_Wblks.argtypes = dims + [_ip, _ip, d, _dp]
    L.psba_maxElmOfUV.restype = d
    L.psba_maxElmOfUV.argtypes = [vp, i, _dp]
    L.psba_update_UV.argtypes = [vp, i, i, i, i, d, _dp, _dp]
    L.psba_restore_UVdiag.argtypes = [vp, i, i, i, i]
    L.psba_compute_Vinv.restype = d
    L.psba_compute_Vinv.argtypes = dims + [_dp]
    L.psba_compute_Yblks.argtypes = dims + [_ip, _ip, _dp]
    L.psba_compute_S.argtypes = dims + [_dp]
    L.psba_compute_ea.argtypes = dims + [_dp]
    L.psba_SPDinv.restype = d
    L.psba_SPDinv.argtypes = [vp, i, _dp]
    L.psba_matVec_mul.argtypes = [vp, i, i, _dp]
    L.psba_cholesky.restype = d
    L.psba_cholesky.argtypes = [vp, i, _dp]
    L.psba_trigMat_inv.restype = d
    L.psba_trigMat_inv.argtypes = [vp, i, _dp]
    L.psba_trigMat_mul.argtypes = [vp, i, _dp]
    L.psba_get_delta_beta.argtypes = [vp, i, _dp, _dp]
    L.psba_compute_cholmod_E.argtypes = [vp, i, _dp]
    L.psba_compute_eb.argtypes = dims + [_dp]
    L.psba_compute_dpb.argtypes = [vp, i, i, i, i, _dp]
    L.psba_compute_newp.argtypes = [vp, i, i, _dp]
    L.psba_update_p.argtypes = [vp, i, i, _dp]
    L.psba_compute_Jmultiply.restype = d
    L.psba_compute_Jmultiply.argtypes = [vp, i, i, i, i, i, _dp]
    L.psba_upload_vec.argtypes = [vp, i, _dp, i]
    L.psba_cholmod_blk.restype = d
    L.psba_cholmod_blk.argtypes = [vp, i, _dp, _dp, _dp, _ip]
    L.psba_cholmod_blk_mat.restype = d
    L.psba_cholmod_blk_mat.argtypes = [vp, i, _dp, _dp, _dp, _dp, _ip]
    L.psba_levmar.argtypes = dims + [_dp]
    L.psba_trust_region.argtypes = dims + [_ip, _dp]
    L.psba_solve.argtypes = [vp, _dp, _dp, _ip]
    L.psba_trace_count.argtypes = [vp]
    L.psba_trace_get.argtypes = [vp, i, C.POINTER(TraceRec)]
    L.psba_set_option.argtypes = [vp, C.c_char_p, d]
    L.psba_get_stat.restype = d
    L.psba_get_stat.argtypes = [vp, C.c_char_p]
    L.psba_get_index.restype = C.c_longlong
    L.psba_get_index.argtypes = [vp, C.c_char_p, vp, C.c_longlong]
    L.psba_force_lambda.argtypes = [vp, _dp, i]
    L.psba_get_params.argtypes = [vp, i, _dp, _dp]
    L.psba_set_params.argtypes = [vp, _dp, _dp]
    L.psba_linearize.argtypes = [vp, d, d]
    L.psba_try_step.argtypes = [vp, d, C.POINTER(TryResult)]
    L.psba_readInitialSBAEstimate.argtypes = [C.c_char_p, C.c_char_p, i, _dp, _ip, _ip, _ip,
                                              C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp),
                                              C.POINTER(_dp), C.POINTER(_ip), C.POINTER(_ip)]
    L.psba_readInitialSBAEstimate_ext.argtypes = L.psba_readInitialSBAEstimate.argtypes + [C.POINTER(_dp), C.POINTER(_dp), _ip]
    L.psba_set_distortion.argtypes = [vp, _dp]
    L.psba_set_covariances.argtypes = [vp, _dp, i]
    L.psba_set_covariances.restype = i
    L.psba_quat2vec.argtypes = [_dp, i, _dp, i]
    L.psba_free.argtypes = [vp]
    L.psba_read_bal.argtypes = [C.c_char_p, _ip, _ip, _ip] + [C.POINTER(_dp)] * 5 + [C.POINTER(_ip)] * 2 + [C.POINTER(_dp)]
    L.psba_vec2quat.argtypes = [_dp, _dp, _dp]
    L.psba_write_sba_result.argtypes = [C.c_char_p, C.c_char_p, i, i, i, _dp, _dp, _dp, _dp, _dp, _ip, _ip]
    L.psba_write_ply.argtypes = [C.c_char_p, i, i, _dp, _dp, _dp]
    L.psba_comm_unique_id.argtypes = [C.c_char_p]
    L.psba_comm_init.argtypes = [i, i, C.c_char_p]
    L.psba_local_range.argtypes = [i, i, _ip, i, i, _ip, _ip, _ip, _ip]
    L.psba_plan_open.restype = vp
    L.psba_plan_open.argtypes = [i, C.c_longlong, _ip, _ip]
    L.psba_plan_get.restype = C.c_longlong
    L.psba_plan_get.argtypes = [vp, C.c_char_p, _ip, C.c_longlong]
    L.psba_plan_close.argtypes = [vp]
    L.psba_version.restype = C.c_char_p
    _lib = L
    return L


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def read_sba(cams_path, pts_path, origin_cnp=11, Kdefault=None, ext=False):
    """readInitialSBAEstimate + quat2vec + the split of main.cpp:131-149 (host code of the library).
    ext=True also returns what the reference parses and drops: kc[m,5] (varKD files) and cov[o,covsz] (or None)."""
    L = lib()
    m, n, o = C.c_int(), C.c_int(), C.c_int()
    K, rot, ex, pts, im = _dp(), _dp(), _dp(), _dp(), _dp()
    ii, jj = _ip(), _ip()
    kc, cov, covsz = _dp(), _dp(), C.c_int()
    kd = None if Kdefault is None else np.ascontiguousarray(Kdefault, dtype=np.float64)
    args = [cams_path.encode(), pts_path.encode(), origin_cnp, _d(kd), C.byref(m), C.byref(n), C.byref(o), C.byref(K), C.byref(rot),
            C.byref(ex), C.byref(pts), C.byref(im), C.byref(ii), C.byref(jj)]
    rc = (L.psba_readInitialSBAEstimate_ext(*args, C.byref(kc), C.byref(cov), C.byref(covsz)) if ext
          else L.psba_readInitialSBAEstimate(*args))
    if rc:
        raise RuntimeError("psba_readInitialSBAEstimate failed with code %d" % rc)
    m, n, o = m.value, n.value, o.value
    take = lambda p, k, shp: np.ctypeslib.as_array(p, shape=(k,)).copy().reshape(shp)
    out = dict(m=m, n=n, o=o, K=take(K, m * 5, (m, 5)), initrot=take(rot, m * 4, (m, 4)), cams=take(ex, m * 6, (m, 6)),
               pts=take(pts, n * 3, (n, 3)), impts=take(im, o * 2, (o, 2)), iidx=take(ii, o, (o,)), jidx=take(jj, o, (o,)))
    if ext:
        out["kc"] = take(kc, m * 5, (m, 5)) if kc else None
        out["cov"] = take(cov, o * covsz.value, (o, covsz.value)) if cov else None
        for p in (kc, cov):
            if p:
                L.psba_free(p)
    for p in (K, rot, ex, pts, im, ii, jj):
        L.psba_free(p)
    return out


def read_bal(path):
    """Native BAL problem-*.txt loader (psba_read_bal): the same dict as read_sba plus kc[m,2] = (k1, k2)."""
    L = lib()
    m, n, o = C.c_int(), C.c_int(), C.c_int()
    K, rot, ex, pts, im, kc = _dp(), _dp(), _dp(), _dp(), _dp(), _dp()
    ii, jj = _ip(), _ip()
    rc = L.psba_read_bal(path.encode(), C.byref(m), C.byref(n), C.byref(o), C.byref(K), C.byref(rot), C.byref(ex),
                         C.byref(pts), C.byref(im), C.byref(ii), C.byref(jj), C.byref(kc))
    if rc:
        raise RuntimeError("psba_read_bal failed with code %d" % rc)
    m, n, o = m.value, n.value, o.value
    take = lambda p, k, shp: np.ctypeslib.as_array(p, shape=(k,)).copy().reshape(shp)
    out = dict(m=m, n=n, o=o, K=take(K, m * 5, (m, 5)), initrot=take(rot, m * 4, (m, 4)), cams=take(ex, m * 6, (m, 6)),
               pts=take(pts, n * 3, (n, 3)), impts=take(im, o * 2, (o, 2)), iidx=take(ii, o, (o,)), jidx=take(jj, o, (o,)),
               kc=take(kc, m * 2, (m, 2)))
    for p in (K, rot, ex, pts, im, ii, jj, kc):
        L.psba_free(p)
    return out


def write_result(prob, cams, pts, cams_path, pts_path, ply_path=None):
    """Save refined parameters (cams[m,6] local rotation | t, pts[n,3]) as SBA text files that read_sba loads
    back (12-column cameras with the rotations recomposed by vec2quat) and optionally as a PLY cloud."""
    L = lib()
    a = lambda x, t=np.float64: np.ascontiguousarray(x, dtype=t)
    K, rot, ce, p3, im = a(prob["K"]), a(prob["initrot"]), a(cams), a(pts), a(prob["impts"])
    ii, jj = a(prob["iidx"], np.int32), a(prob["jidx"], np.int32)
    rc = L.psba_write_sba_result(cams_path.encode(), pts_path.encode(), int(prob["m"]), int(prob["n"]), int(prob["o"]),
                                 _d(K), _d(rot), _d(ce), _d(p3), _d(im), _i(ii), _i(jj))
    if rc == 0 and ply_path:
        rc = L.psba_write_ply(ply_path.encode(), int(prob["m"]), int(prob["n"]), _d(rot), _d(ce), _d(p3))
    if rc:
        raise RuntimeError("writing the result failed")


def local_range(n, o, iidx, rank, nranks):
    L = lib()
    v = [C.c_int() for _ in range(4)]
    a = np.ascontiguousarray(iidx, dtype=np.int32)
    L.psba_local_range(n, o, _i(a), rank, nranks, *[C.byref(x) for x in v])
    return tuple(x.value for x in v)


PLAN_TABLES = ("stats", "cam2pos", "tile_index", "step_panels", "step_panel_ptr", "crit_I", "crit_K", "step_crit_ptr", "psrc_ptr", "psrc",
               "def_I", "def_J", "def_sptr", "def_src", "step_def_ptr", "b_J", "b_sptr", "b_slot", "step_b_ptr",
               "crit_desc", "crit_src", "def_desc", "def_srcs", "bw_order", "coltile_ptr", "coltile_row", "coltile_slot")


def plan_tiles(m, pair_k, pair_l):
    """Host stage of the camera-system set-up (ordering, symbolic factor, step schedule, task lists) as numpy tables; no GPU needed
    (psba_plan_open / psba_plan_get in include/psba_b200.h)."""
    L = lib()
    k = np.ascontiguousarray(pair_k, dtype=np.int32)
    l = np.ascontiguousarray(pair_l, dtype=np.int32)
    h = L.psba_plan_open(int(m), len(k), _i(k), _i(l))
    if not h:
        raise ValueError("psba_plan_open refused the pair list")
    try:
        out = {}
        for name in PLAN_TABLES:
            n = L.psba_plan_get(h, name.encode(), None, 0)
            a = np.zeros(max(int(n), 1), dtype=np.int32)
            L.psba_plan_get(h, name.encode(), _i(a), n)
            out[name] = a[:n]
        return out
    finally:
        L.psba_plan_close(h)


class _Pinned(np.ndarray):
    """numpy view of page-locked memory from psba_host_alloc; the block is freed with the last view"""


_pinned_blocks = {}


def pinned_array(src):
    """Copy of `src` (C-contiguous) in page-locked host memory (psba_host_alloc): uploads and downloads through the
    ABI then run at the PCIe rate instead of the pageable-memory rate."""
    src = np.ascontiguousarray(src)
    L = lib()
    ptr = L.psba_host_alloc(max(src.nbytes, 16))
    buf = (C.c_char * max(src.nbytes, 16)).from_address(ptr)
    out = np.frombuffer(buf, dtype=src.dtype, count=src.size).reshape(src.shape)
    out[...] = src
    _pinned_blocks[ptr] = buf
    return out


def pinned_problem(prob):
    """the problem dictionary with every array the engine uploads moved to page-locked memory"""
    q = dict(prob)
    for k, t in (("K", np.float64), ("impts", np.float64), ("initrot", np.float64), ("cams", np.float64), ("pts", np.float64),
                 ("iidx", np.int32), ("jidx", np.int32)):
        q[k] = pinned_array(np.ascontiguousarray(prob[k], dtype=t))
    return q


def free_pinned():
    L = lib()
    for ptr in list(_pinned_blocks):
        del _pinned_blocks[ptr]
        L.psba_host_free(ptr)


class PSBA:
    """One problem on one GPU (struct psba_ctx); mirrors PSBA_struct + the operator wrappers."""

    def __init__(self, prob):
        self.L = lib()
        self.m, self.n, self.o = int(prob["m"]), int(prob["n"]), int(prob["o"])
        self.N = 6 * self.m
        a = lambda x, t=np.float64: np.ascontiguousarray(x, dtype=t)
        self.h = C.c_void_p(self.L.psba_setup_cl(6, 3, 2, self.m, self.n, self.o))
        K, imp, rot, cams, pts = a(prob["K"]), a(prob["impts"]), a(prob["initrot"]), a(prob["cams"]), a(prob["pts"])
        self.L.psba_fill_initBuffer2(self.h, 6, 3, 2, self.m, self.n, self.o, _d(K), _d(imp), _d(rot), _d(cams), _d(pts))
        ii, jj = a(prob["iidx"], np.int32), a(prob["jidx"], np.int32)
        self.L.psba_fill_idxBuffer(self.h, self.m, self.n, self.o, _i(ii), _i(jj))
        self.n_loc = int(self.stat("n_local"))
        self.o_loc = int(self.stat("o_local"))
        self.T_loc = self.N + 3 * self.n_loc
        self._dims = (6, 3, 2, self.n, self.m, self.o)

    def set_distortion(self, kc):
        a = None if kc is None else np.ascontiguousarray(kc, dtype=np.float64)
        assert a is None or a.shape == (self.m, 5)
        self.L.psba_set_distortion(self.h, _d(a))

    def set_covariances(self, cov):
        a = None if cov is None else np.ascontiguousarray(cov, dtype=np.float64)
        assert a is None or (a.shape[0] == self.o and a.shape[1] in (3, 4))
        if self.L.psba_set_covariances(self.h, _d(a), 0 if a is None else a.shape[1]):
            raise ValueError("a covariance is not positive definite")

    # ---- operators (names of PSBA/sba_func.h) ------------------------------------------------
    def compute_exQT(self, params=PARAMS_CUR, want=False):
        ex = np.zeros((self.o_loc, 2)) if want else None
        cost = self.L.psba_compute_exQT(self.h, *self._dims, params, _d(ex))
        return (cost, ex) if want else cost

    def compute_jacobiQT(self, want=False):
        JA = np.zeros((self.o_loc, 2, 6)) if want else None
        JB = np.zeros((self.o_loc, 2, 3)) if want else None
        self.L.psba_compute_jacobiQT(self.h, *self._dims, _d(JA), _d(JB))
        return (JA, JB) if want else None

    def compute_U(self, coeff=1.0):
        out = np.zeros((self.m, 6, 6))
        self.L.psba_compute_U(self.h, *self._dims, coeff, _d(out))
        return out

    def compute_V(self, coeff=1.0):
        out = np.zeros((self.n_loc, 3, 3))
        self.L.psba_compute_V(self.h, *self._dims, coeff, _d(out))
        return out

    def compute_Wblks(self, coeff=1.0):
        out = np.zeros((self.o_loc, 6, 3))
        self.L.psba_compute_Wblks(self.h, *self._dims, None, None, coeff, _d(out))
        return out

    def compute_g(self, coeff=1.0):
        out = np.zeros(self.T_loc)
        self.L.psba_compute_g(self.h, *self._dims, coeff, _d(out))
        return out

    def maxElmOfUV(self):
        uv = np.zeros(self.T_loc)
        mx = self.L.psba_maxElmOfUV(self.h, self.T_loc, _d(uv))
        return mx, uv

    def update_UV(self, mu):
        self.L.psba_update_UV(self.h, 6, 3, self.n, self.m, mu, None, None)

    def restore_UVdiag(self):
        self.L.psba_restore_UVdiag(self.h, 6, 3, self.n, self.m)

    def compute_Vinv(self):
        out = np.zeros((self.n_loc, 3, 3))
        ret = self.L.psba_compute_Vinv(self.h, *self._dims, _d(out))
        return ret, out

    def compute_Yblks(self):
        out = np.zeros((self.o_loc, 6, 3))
        self.L.psba_compute_Yblks(self.h, *self._dims, None, None, _d(out))
        return out

    def compute_S(self, want=True):
        out = np.zeros((self.N, self.N)) if want else None
        self.L.psba_compute_S(self.h, *self._dims, _d(out))
        return out

    def compute_ea(self):
        out = np.zeros(self.N)
        self.L.psba_compute_ea(self.h, *self._dims, _d(out))
        return out

    def SPDinv(self, want=False):
        out = np.zeros((self.N, self.N)) if want else None
        ret = self.L.psba_SPDinv(self.h, self.N, _d(out))
        return (ret, out) if want else ret

    def cholesky(self):
        out = np.zeros((self.N, self.N))
        return self.L.psba_cholesky(self.h, self.N, _d(out)), out

    def trigMat_inv(self):
        out = np.zeros((self.N, self.N))
        self.L.psba_trigMat_inv(self.h, self.N, _d(out))
        return out

    def trigMat_mul(self):
        out = np.zeros((self.N, self.N))
        self.L.psba_trigMat_mul(self.h, self.N, _d(out))
        return out

    def get_delta_beta(self):
        dl, bt = C.c_double(), C.c_double()
        self.L.psba_get_delta_beta(self.h, self.N, C.byref(dl), C.byref(bt))
        return dl.value, bt.value

    def compute_cholmod_E(self):
        out = np.zeros(self.N)
        self.L.psba_compute_cholmod_E(self.h, self.N, _d(out))
        return out

    def matVec_mul(self):
        out = np.zeros(self.N)
        self.L.psba_matVec_mul(self.h, self.N, self.N, _d(out))
        return out

    def compute_eb(self):
        out = np.zeros(self.T_loc)
        self.L.psba_compute_eb(self.h, *self._dims, _d(out))
        return out

    def compute_dpb(self):
        out = np.zeros(self.T_loc)
        self.L.psba_compute_dpb(self.h, 6, 3, self.m, self.n, _d(out))
        return out

    def compute_newp(self):
        out = np.zeros(self.T_loc)
        self.L.psba_compute_newp(self.h, self.N, 3 * self.n, _d(out))
        return out

    def update_p(self):
        out = np.zeros(self.T_loc)
        self.L.psba_update_p(self.h, self.N, 3 * self.n, _d(out))
        return out

    def compute_Jmultiply(self, vec=VEC_G, want=False):
        out = np.zeros((self.o_loc, 2)) if want else None
        s = self.L.psba_compute_Jmultiply(self.h, 2, self.n, self.m, self.o, vec, _d(out))
        return (s, out) if want else s

    def upload_vec(self, vec, host):
        h = np.ascontiguousarray(host, dtype=np.float64)
        self.L.psba_upload_vec(self.h, vec, _d(h), h.size)

    def cholmod_blk(self):
        E = np.zeros(self.N)
        delta, beta, ns = C.c_double(), C.c_double(), C.c_int()
        s = self.L.psba_cholmod_blk(self.h, self.N, _d(E), C.byref(delta), C.byref(beta), C.byref(ns))
        return dict(sumE=s, E=E, delta=delta.value, beta=beta.value, n_scalar_blocks=ns.value)

    def cholmod_blk_mat(self, mat):
        """modified Cholesky of a host matrix (psba_cholmod_blk_mat): returns dict(sumE, E, L, delta, beta, n_scalar_blocks)"""
        A = np.ascontiguousarray(mat, dtype=np.float64).copy()
        n = A.shape[0]
        E = np.zeros(n)
        delta, beta, ns = C.c_double(), C.c_double(), C.c_int()
        s = self.L.psba_cholmod_blk_mat(self.h, n, _d(A), _d(E), C.byref(delta), C.byref(beta), C.byref(ns))
        return dict(sumE=s, E=E, L=A, delta=delta.value, beta=beta.value, n_scalar_blocks=ns.value)

    # ---- fused steps and drivers -----------------------------------------------------------
    def linearize(self, coeff_uvw=1.0, coeff_g=1.0):
        self.L.psba_linearize(self.h, coeff_uvw, coeff_g)

    def try_step(self, mu):
        r = TryResult()
        self.L.psba_try_step(self.h, mu, C.byref(r))
        return dict(cost_new=r.cost_new, dp_L2=r.dp_L2, dp_dot=r.dp_dot, solve_status=r.solve_status)

    def levmar(self):
        fe = C.c_double()
        flag = self.L.psba_levmar(self.h, *self._dims, C.byref(fe))
        return flag, fe.value

    def trust_region(self):
        fe = C.c_double()
        flag = self.L.psba_trust_region(self.h, *self._dims, None, C.byref(fe))
        return flag, fe.value

    def solve(self):
        ie, fe, it = C.c_double(), C.c_double(), C.c_int()
        flag = self.L.psba_solve(self.h, C.byref(ie), C.byref(fe), C.byref(it))
        return dict(flag=flag, initErr=ie.value, finalErr=fe.value, itno=it.value)

    def trace(self):
        out = []
        r = TraceRec()
        for k in range(self.L.psba_trace_count(self.h)):
            self.L.psba_trace_get(self.h, k, C.byref(r))
            out.append(dict(phase=r.phase, itno=r.itno, err=r.err, rho=r.rho, mu=r.mu, delta=r.delta,
                            pnorm=r.pnorm, accepted=r.accepted))
        return out

    def set_option(self, name, v):
        self.L.psba_set_option(self.h, name.encode(), float(v))

    def index(self, name):
        """device-built index table by name (include/psba_b200.h: psba_get_index)"""
        cnt = self.L.psba_get_index(self.h, name.encode(), None, 0)
        out = np.zeros(cnt, dtype=np.int64 if name in ("pchunk_beg", "pchunk_end") else np.int32)
        if cnt:
            self.L.psba_get_index(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), cnt)
        return out

    def stat(self, name):
        return self.L.psba_get_stat(self.h, name.encode())

    def force_lambda(self, lams):
        a = np.ascontiguousarray(lams, dtype=np.float64)
        self.L.psba_force_lambda(self.h, _d(a), a.size)

    def get_params(self, params=PARAMS_CUR, out=None):
        """refined cameras (m x 6) and this rank's points; `out` = (cams, pts) arrays to fill (e.g. page-locked ones)"""
        cams, pts = out if out is not None else (np.zeros((self.m, 6)), np.zeros((self.n_loc, 3)))
        assert cams.shape == (self.m, 6) and pts.shape == (self.n_loc, 3) and cams.flags.c_contiguous and pts.flags.c_contiguous
        self.L.psba_get_params(self.h, params, _d(cams), _d(pts))
        return cams, pts

    def set_params(self, cams=None, pts=None):
        a = None if cams is None else np.ascontiguousarray(cams, dtype=np.float64)
        b = None if pts is None else np.ascontiguousarray(pts, dtype=np.float64)
        self.L.psba_set_params(self.h, _d(a), _d(b))

    def close(self):
        if self.h:
            self.L.psba_release_buffer(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
