"""CPU ORACLE for the PSBA hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this package.  The product package psba_b200 never does.

Two back ends behind one interface (oracle/psba_oracle.h):
  kind="restatement": oracle/liboracle.so, the plain-C restatement in this directory;
  kind="reference"  : oracle/_ref/libpsba_ref.so, the reference's OWN kernel bodies and loaders
                      compiled in place from /root/reference (see oracle/ref_shim/), driven by
                      the same restated LM / trust-region drivers.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

ITER_NAMES = {1: "TURN_TO_LM", 2: "TURN_TO_TR", 3: "CONTINUE", 4: "ERR", 5: "DP_NO_CHANGE",
              6: "ERR_SMALL_ENOUGH", 7: "PASS"}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(ref=True):
    """Compile the restatement (always) and oracle/_ref (only when the reference tree exists)."""
    subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    if ref and os.path.isdir("/root/reference/CL_files"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libpsba_ref.so"))


def lib(kind="restatement"):
    if kind in _LIBS:
        return _LIBS[kind]
    path = os.path.join(_HERE, "liboracle.so") if kind == "restatement" else os.path.join(_HERE, "_ref", "libpsba_ref.so")
    if not os.path.exists(path):
        if kind == "restatement":
            build(ref=False)
        else:
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
    L = C.CDLL(path)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _ip, _ip, C.c_int]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_ptr_d.restype = _dp
    L.orc_ptr_d.argtypes = [C.c_void_p, C.c_char_p]
    L.orc_ptr_i.restype = _ip
    L.orc_ptr_i.argtypes = [C.c_void_p, C.c_char_p]
    L.orc_pair_ptr.restype = C.POINTER(C.c_longlong)
    L.orc_pair_ptr.argtypes = [C.c_void_p]
    L.orc_get.restype = C.c_double
    L.orc_get.argtypes = [C.c_void_p, C.c_char_p]
    L.orc_set.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.orc_call.restype = C.c_double
    L.orc_call.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.orc_trace_get.argtypes = [C.c_void_p, C.c_int, _dp]
    L.orc_force_lambda.argtypes = [C.c_void_p, _dp, C.c_int]
    L.orc_set_ext.argtypes = [C.c_void_p, _dp, _dp]
    L.orc_solve.argtypes = [C.c_void_p]
    L.orc_levmar.argtypes = [C.c_void_p]
    L.orc_trust_region.argtypes = [C.c_void_p]
    L.orc_use_ops.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_native_ops.restype = C.c_void_p
    L.orc_read_sba.argtypes = [C.c_char_p, C.c_char_p, C.c_int, _ip, _ip, _ip,
                               C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_ip), C.POINTER(_ip)]
    L.orc_generate_idxs.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _ip]
    L.orc_split_motion.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp]
    L.orc_quat2vec.argtypes = [_dp, C.c_int, _dp, C.c_int]
    for f in ("orc_cholesky", "orc_SPDinv"):
        getattr(L, f).restype = C.c_double
    L.orc_cholesky.argtypes = [_dp, _dp, C.c_int]
    L.orc_SPDinv.argtypes = [_dp, _dp, C.c_int]
    L.orc_trigMat_inv.argtypes = [_dp, _dp, C.c_int]
    L.orc_trigMat_mul.argtypes = [_dp, _dp, C.c_int]
    L.orc_get_delta_beta.argtypes = [_dp, C.c_int, _dp, _dp]
    L.orc_cholmod_blk.argtypes = [_dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double, _ip]
    L.orc_cholmod_E.argtypes = [_dp, _dp, C.c_int]
    L.orc_potrf_solve.restype = C.c_double
    L.orc_potrf_solve.argtypes = [_dp, _dp, _dp, C.c_int]
    if kind == "reference":
        L.ref_ops.restype = C.c_void_p
        L.ref_read_sba.argtypes = [C.c_char_p, C.c_char_p, C.c_int, _ip, _ip, _ip,
                                   C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp), C.POINTER(C.c_void_p)]
        L.ref_generate_idxs.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_void_p, _ip, _ip, _ip, _ip, _ip]
        L.ref_quat2vec.argtypes = [_dp, C.c_int, _dp, C.c_int]
        L.ref_free.argtypes = [C.c_void_p]
    _LIBS[kind] = L
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def read_sba(cams_path, pts_path, origin_cnp=11, kind="restatement"):
    """Load an SBA-format problem exactly as PSBA/main.cpp:102-149 does.

    Returns a dict with m, n, o, K[m,5], initrot[m,4], cams[m,6], pts[n,3], impts[o,2],
    iidx[o], jidx[o] (+ motstruct, raw per-camera parameters after quat2vec).
    """
    L = lib(kind)
    m, n, o = C.c_int(), C.c_int(), C.c_int()
    mot, rot, imp = _dp(), _dp(), _dp()
    if kind == "reference":
        vmp = C.c_void_p()
        L.ref_read_sba(cams_path.encode(), pts_path.encode(), origin_cnp, C.byref(m), C.byref(n), C.byref(o),
                       C.byref(mot), C.byref(rot), C.byref(imp), C.byref(vmp))
        m, n, o = m.value, n.value, o.value
        vmask = np.ctypeslib.as_array(C.cast(vmp, C.POINTER(C.c_ubyte)), shape=(n * m,)).copy().reshape(n, m)
        iidx = np.zeros(o, np.int32); jidx = np.zeros(o, np.int32)
        blk = np.zeros(n * m, np.int32); comm = np.zeros(n * m * m, np.int32); cnt = np.zeros(m * m, np.int32)
        impts = np.ctypeslib.as_array(imp, shape=(o * 2,)).copy()
        L.ref_generate_idxs(m, n, o, _d(impts), vmp, _i(comm), _i(cnt), _i(iidx), _i(jidx), _i(blk))
        extra = dict(vmask=vmask, blk_idx=blk.reshape(n, m), comm3DIdx=comm.reshape(m, m, n), comm3DIdxCnt=cnt.reshape(m, m))
        L.ref_free(vmp)
    else:
        nf, fr = _ip(), _ip()
        rc = L.orc_read_sba(cams_path.encode(), pts_path.encode(), origin_cnp, C.byref(m), C.byref(n), C.byref(o),
                            C.byref(mot), C.byref(rot), C.byref(imp), C.byref(nf), C.byref(fr))
        if rc:
            raise RuntimeError("orc_read_sba failed with code %d" % rc)
        m, n, o = m.value, n.value, o.value
        impts = np.ctypeslib.as_array(imp, shape=(o * 2,)).copy()
        nfa = np.ctypeslib.as_array(nf, shape=(n,)).copy()
        fra = np.ctypeslib.as_array(fr, shape=(o,)).copy()
        iidx = np.zeros(o, np.int32); jidx = np.zeros(o, np.int32)
        L.orc_generate_idxs(m, n, o, _i(nfa), _i(fra), _i(iidx), _i(jidx))
        extra = dict(pt_nframes=nfa, frames=fra)
    motstruct = np.ctypeslib.as_array(mot, shape=(m * origin_cnp + n * 3,)).copy()
    initrot = np.ctypeslib.as_array(rot, shape=(m * 4,)).copy().reshape(m, 4)
    K = np.zeros((m, 5)); cams = np.zeros((m, 6))
    lib("restatement").orc_split_motion(_d(motstruct), origin_cnp, m, _d(K), _d(cams))
    pts = motstruct[m * origin_cnp:].copy().reshape(n, 3)
    out = dict(m=m, n=n, o=o, K=K, initrot=initrot, cams=cams, pts=pts, impts=impts.reshape(o, 2),
               iidx=iidx, jidx=jidx, motstruct=motstruct, origin_cnp=origin_cnp)
    out.update(extra)
    return out


class Problem:
    """One bundle-adjustment problem held by the oracle (struct orc_state)."""

    def __init__(self, prob, kind="restatement", dense=None):
        self.kind = kind
        self.L = lib(kind)
        self.m, self.n, self.o = int(prob["m"]), int(prob["n"]), int(prob["o"])
        if dense is None:
            dense = kind == "reference"
        a = lambda x, t=np.float64: np.ascontiguousarray(x, dtype=t)
        K, imp, rot, cams, pts = a(prob["K"]), a(prob["impts"]), a(prob["initrot"]), a(prob["cams"]), a(prob["pts"])
        ii, jj = a(prob["iidx"], np.int32), a(prob["jidx"], np.int32)
        self.h = C.c_void_p(self.L.orc_create(self.m, self.n, self.o, _d(K), _d(imp), _d(rot), _d(cams), _d(pts),
                                              _i(ii), _i(jj), 1 if dense else 0))
        if kind == "reference":
            self.L.orc_use_ops(self.h, self.L.ref_ops())
        self.N, self.T = 6 * self.m, 6 * self.m + 3 * self.n
        self._shapes = dict(K=(self.m, 5), impts=(self.o, 2), initcams=(self.m, 4), cams=(self.m, 6), newcams=(self.m, 6),
                            pts=(self.n, 3), newpts=(self.n, 3), ex=(self.o, 2), JA=(self.o, 2, 6), JB=(self.o, 2, 3),
                            U=(self.m, 6, 6), V=(self.n, 3, 3), UVdiag=(self.T,), W=(self.o, 6, 3), Y=(self.o, 6, 3),
                            S=(self.N, self.N), Saux=(self.N, self.N), diagAux=(3 * self.N,), blkBackup=(3 * self.N,),
                            E=(self.N,), g=(self.T,), dp=(self.T,), eab=(self.T,), Jx1=(self.o, 2), Jx2=(self.o, 2))

    def buf(self, name):
        """numpy VIEW of a double buffer of the oracle state."""
        shp = self._shapes[name]
        p = self.L.orc_ptr_d(self.h, name.encode())
        return np.ctypeslib.as_array(p, shape=(int(np.prod(shp)),)).reshape(shp)

    def ibuf(self, name, size):
        p = self.L.orc_ptr_i(self.h, name.encode())
        return np.ctypeslib.as_array(p, shape=(size,))

    def pair_ptr(self):
        return np.ctypeslib.as_array(self.L.orc_pair_ptr(self.h), shape=(self.m * self.m + 1,))

    def get(self, name):
        return self.L.orc_get(self.h, name.encode())

    def set(self, name, v):
        self.L.orc_set(self.h, name.encode(), float(v))

    def call(self, op, arg=0.0):
        return self.L.orc_call(self.h, op.encode(), float(arg))

    def set_ext(self, kc=None, wgt=None):
        """extended camera model: kc[m,5] distortion (sba varKD), wgt[o,3] lower-triangular residual weights"""
        a = None if kc is None else np.ascontiguousarray(kc, dtype=np.float64)
        b = None if wgt is None else np.ascontiguousarray(wgt, dtype=np.float64)
        self.L.orc_set_ext(self.h, None if a is None else _d(a), None if b is None else _d(b))

    def force_lambda(self, lams):
        a = np.ascontiguousarray(lams, dtype=np.float64)
        self.L.orc_force_lambda(self.h, _d(a), len(a))

    def solve(self):
        return self.L.orc_solve(self.h)

    def levmar(self):
        return self.L.orc_levmar(self.h)

    def trust_region(self):
        return self.L.orc_trust_region(self.h)

    def trace(self):
        out = []
        rec = np.zeros(8)
        for k in range(int(self.get("ntrace"))):
            self.L.orc_trace_get(self.h, k, _d(rec))
            out.append(dict(phase=int(rec[0]), itno=int(rec[1]), err=rec[2], rho=rec[3], mu=rec[4],
                            delta=rec[5], pnorm=rec[6], accepted=int(rec[7])))
        return out

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
