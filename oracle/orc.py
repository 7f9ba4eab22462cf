"""ctypes binding of the CPU oracle (oracle/liborc*.so).  TEST INFRASTRUCTURE ONLY: importable
from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import ekfb200  # noqa: E402  (import shim at the repository root)

abi = ekfb200.load_package()._abi


def build(force=False):
    """Compile liborc.so / liborc_omp.so with oracle/Makefile (gcc only, no GPU needed)."""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


_libs = {}


def lib(omp=False):
    key = "omp" if omp else "st"
    if key in _libs:
        return _libs[key]
    path = os.path.join(_HERE, "liborc_omp.so" if omp else "liborc.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    vp, i32, f64, f32 = C.c_void_p, C.c_int, C.c_double, C.c_float
    P = C.POINTER
    sig = {
        "orc_create": (vp, [P(abi.EkfConfig), i32]),
        "orc_destroy": (None, [vp]),
        "orc_capture": (None, [vp, vp, i32, i32, i32, f64]),
        "orc_add_feature": (i32, [vp, f32, f32]),
        "orc_add_features_structured": (i32, [vp, vp, i32]),
        "orc_remove_feature": (None, [vp, i32]),
        "orc_predict": (None, [vp, vp, vp, i32]),
        "orc_match": (i32, [vp]),
        "orc_update_after_match": (None, [vp, vp, i32]),
        "orc_update": (None, [vp, vp, i32]),
        "orc_inject_match": (None, [vp, i32, f64, f64, i32]),
        "orc_convert2xyz": (None, [vp, i32]),
        "orc_state_dim": (i32, [vp]),
        "orc_num_features": (i32, [vp]),
        "orc_get_full": (None, [vp, vp, vp, i32]),
        "orc_set_full": (None, [vp, vp, vp, i32]),
        "orc_get_feature": (None, [vp, i32, P(abi.EkfFeatureInfo)]),
        "orc_get_template": (None, [vp, i32, i32, vp]),
        "orc_get_S_blocks": (None, [vp, vp]),
        "orc_get_St": (i32, [vp, vp, i32]),
        "orc_get_step_stats": (None, [vp, P(abi.EkfStepStats)]),
        "orc_covariance_parameter": (f64, [vp]),
        "orc_get_dt": (f64, [vp]),
        "orc_min_margin": (f64, [vp]),
        "orc_match_batch": (None, [vp, i32, i32, i32, i32, vp, i32, i32, vp, vp, f32, f32, f32, vp, vp, i32]),
        "orc_import_state": (None, [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32]),
        "orc_num_deleted": (i32, [vp]),
        "orc_get_deleted": (None, [vp, i32, vp, vp, vp]),
        "orc_points_features": (i32, [vp, vp, i32]),
        "orc_rts_epoch": (None, [vp, vp, vp, vp, vp, vp, vp, f64]),
        "orc_num_threads": (i32, []),
        "orc_set_num_threads": (None, [i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _libs[key] = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleFilter:
    """Same method names as the product's VSlamFilter mirror, backed by the CPU oracle.
    kind: 0 = fp64 state / float matcher (parity target), 1 = all float, 2 = all double."""

    def __init__(self, cfg, kind=0, omp=False):
        self.L = lib(omp)
        self.cfg = cfg
        self.h = self.L.orc_create(C.byref(cfg), kind)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def captureNewFrame(self, img, stamp=-1.0):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        self.L.orc_capture(self.h, _ptr(img), img.shape[1], img.shape[0], img.strides[0], float(stamp))

    def addFeature(self, u, v):
        return self.L.orc_add_feature(self.h, float(u), float(v))

    def addFeatures(self, pixels):
        """addFeature (vslamRansac.cpp:309-371) for every pixel of an (m, 2) array in order, through the sparsity-exploiting
        form (bit-identical to the dense per-feature calls, O(n) instead of O(n^3) each).  Returns the number added."""
        uv = np.ascontiguousarray(pixels, dtype=np.float32).reshape(-1, 2)
        return self.L.orc_add_features_structured(self.h, _ptr(uv), int(uv.shape[0]))

    def removeFeature(self, i):
        self.L.orc_remove_feature(self.h, int(i))

    def predict(self, dv=(0, 0, 0), dw=(0, 0, 0), vcontrol=False):
        a = np.asarray(dv, dtype=np.float64); b = np.asarray(dw, dtype=np.float64)
        self.L.orc_predict(self.h, _ptr(a), _ptr(b), int(bool(vcontrol)))

    def match(self):
        return self.L.orc_match(self.h)

    def update_after_match(self, picks=None):
        p = np.ascontiguousarray(picks if picks is not None else np.zeros(0), dtype=np.uint32)
        self.L.orc_update_after_match(self.h, _ptr(p), int(p.size))

    def update(self, picks=None):
        self.match()
        self.update_after_match(picks)

    def inject_match(self, i, zu, zv, accepted=True):
        self.L.orc_inject_match(self.h, int(i), float(zu), float(zv), int(bool(accepted)))

    def convert2XYZ_ifLinear(self, i):
        self.L.orc_convert2xyz(self.h, int(i))

    def convert2XYZ_ifLinearAll(self):
        self.L.orc_convert2xyz(self.h, -1)

    def numOfFeatures(self):
        return self.L.orc_num_features(self.h)

    def state_dim(self):
        return self.L.orc_state_dim(self.h)

    def get_full(self):
        n = self.state_dim()
        mu = np.zeros(n); S = np.zeros((n, n))
        self.L.orc_get_full(self.h, _ptr(mu), _ptr(S), n)
        return mu, S

    def set_full(self, mu, S):
        mu = np.ascontiguousarray(mu, dtype=np.float64); S = np.ascontiguousarray(S, dtype=np.float64)
        assert mu.size == self.state_dim() and S.shape == (mu.size, mu.size)
        self.L.orc_set_full(self.h, _ptr(mu), _ptr(S), mu.size)

    def getState(self):
        return self.get_full()[0][:14].copy()

    def getSigma(self):
        return self.get_full()[1][:14, :14].copy()

    def Covariance_Parameter(self):
        return self.L.orc_covariance_parameter(self.h)

    def getDt(self):
        return self.L.orc_get_dt(self.h)

    def feature(self, i):
        o = abi.EkfFeatureInfo()
        self.L.orc_get_feature(self.h, int(i), C.byref(o))
        return o

    def template(self, i, which=0):
        w = self.cfg.window_size
        out = np.zeros((w, w), dtype=np.uint8)
        self.L.orc_get_template(self.h, int(i), int(which), _ptr(out))
        return out

    def S_blocks(self):
        out = np.zeros((self.numOfFeatures(), 2, 2))
        self.L.orc_get_S_blocks(self.h, _ptr(out))
        return out

    def St(self):
        k = self.L.orc_get_St(self.h, None, 0)
        out = np.zeros((k, k))
        self.L.orc_get_St(self.h, _ptr(out), k * k)
        return out

    def stats(self):
        s = abi.EkfStepStats()
        self.L.orc_get_step_stats(self.h, C.byref(s))
        return s

    def min_margin(self):
        return self.L.orc_min_margin(self.h)

    def deleted(self):
        """Archive of removed XYZ features (vslamRansac.cpp:394-404): list of (real_index, XYZ_pos[3], cov_4_delete[9])."""
        out = []
        for i in range(self.L.orc_num_deleted(self.h)):
            ri = np.zeros(1, dtype=np.int32); xyz = np.zeros(3); cov = np.zeros(9)
            self.L.orc_get_deleted(self.h, i, _ptr(ri), _ptr(xyz), _ptr(cov))
            out.append((int(ri[0]), xyz, cov))
        return out

    def getPointsFeatures(self):
        rows = self.L.orc_points_features(self.h, None, 0)
        out = np.zeros((rows, 12))
        self.L.orc_points_features(self.h, _ptr(out), rows)
        return out

    def rts_epoch(self, MU, SIGMA, MU_S, SIGMA_S, dTspeed, dRspeed, deltaT):
        """VSlamFilter::rts_epoch on 13-dimensional camera states; returns the smoothed (MU, SIGMA)."""
        mu = np.array(MU, dtype=np.float64).copy(); sg = np.array(SIGMA, dtype=np.float64).reshape(13, 13).copy()
        mus = np.ascontiguousarray(MU_S, dtype=np.float64); sgs = np.ascontiguousarray(SIGMA_S, dtype=np.float64)
        a = np.ascontiguousarray(dTspeed, dtype=np.float64); b = np.ascontiguousarray(dRspeed, dtype=np.float64)
        self.L.orc_rts_epoch(self.h, _ptr(mu), _ptr(sg), _ptr(mus), _ptr(sgs), _ptr(a), _ptr(b), float(deltaT))
        return mu, sg

    def import_from(self, other):
        """Copy mu, Sigma and the feature table of another filter object (product or oracle) that
        offers get_full / feature / template.  Used to seed large maps without the dense O(n^3)
        addFeature of the reference."""
        mu, S = other.get_full()
        N = other.numOfFeatures()
        w = self.cfg.window_size
        feats = [other.feature(i) for i in range(N)]
        i32 = lambda f: np.ascontiguousarray([getattr(x, f) for x in feats], dtype=np.int32)  # noqa: E731
        pos, cod, ntot, nfind, real = (i32(f) for f in ("position_in_state", "coding", "n_tot", "n_find", "real_index"))
        cen = np.ascontiguousarray([[x.center[0], x.center[1]] for x in feats], dtype=np.float32).reshape(-1)
        tm = np.ascontiguousarray(np.stack([other.template(i) for i in range(N)]) if N else np.zeros((0, w, w)), dtype=np.uint8)
        mu = np.ascontiguousarray(mu); S = np.ascontiguousarray(S)
        nxt = int(real.max()) + 1 if N else 1
        self.L.orc_import_state(self.h, mu.size, N, _ptr(mu), _ptr(S), _ptr(pos), _ptr(cod), _ptr(ntot), _ptr(nfind),
                                _ptr(real), _ptr(cen), _ptr(tm), nxt)


def match_batch(frames, templates, h, S, sigma_size=3.0, ncc_threshold=0.8, search_clamp=20.0, kind_mf=0, omp=True):
    """Patch::findMatch over a batch on the CPU oracle.  Same argument meaning as ekf_match_batch."""
    L = lib(omp)
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    F, H, W = frames.shape
    templates = np.ascontiguousarray(templates, dtype=np.uint8)
    w = templates.shape[-1]
    M = templates.shape[0] // F
    h = np.ascontiguousarray(h, dtype=np.float64); S = np.ascontiguousarray(S, dtype=np.float64)
    uv = np.zeros((F * M, 2), dtype=np.int32); score = np.zeros(F * M, dtype=np.float32)
    L.orc_match_batch(_ptr(frames), F, W, H, frames.strides[1], _ptr(templates), M, w, _ptr(h), _ptr(S),
                      float(sigma_size), float(ncc_threshold), float(search_clamp), _ptr(uv), _ptr(score), int(kind_mf))
    return uv, score
