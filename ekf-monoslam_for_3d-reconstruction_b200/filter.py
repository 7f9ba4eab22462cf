"""Python mirror of the reference's `class VSlamFilter` (mono-slam/src/vslamRansac.hpp:27-141) over
the C ABI.  Same method names, argument meaning and call order as the reference class; Eigen / OpenCV
types become numpy arrays.  Used by the tests and bench.py; the C++ twin is host/vslam_filter.hpp.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._lib import lib


class EkfError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class VSlamFilter:
    """VSlamFilter(cfg) — cfg is an EkfConfig (see default_config) instead of a libconfig file name
    (vslamRansac.cpp:142).  feature_capacity bounds the map size of this handle."""

    def __init__(self, cfg=None, feature_capacity=128, device=0):
        self.L = lib()
        self.cfg = cfg if cfg is not None else _abi.default_config()
        h = C.c_void_p()
        rc = self.L.ekf_create(C.byref(self.cfg), int(feature_capacity), int(device), C.byref(h))
        if rc != 0:
            raise EkfError(f"ekf_create failed with {rc} "
                           f"({'unsupported configuration' if rc == _abi.EKF_ERR_UNSUPPORTED else 'see stderr'})")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.ekf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise EkfError(f"ekf error {rc}: {self.L.ekf_last_error(self.h).decode()}")
        return rc

    # ---- the per-frame path ------------------------------------------------------------------
    def captureNewFrame(self, img, stamp=-1.0):
        """vslamRansac.cpp:226-245; img is an HxW (gray) or HxWx3 (BGR) uint8 numpy array (host) — copied,
        resized by 1 / cfg.scale and converted to gray inside the call."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        if img.ndim == 3:
            self._ck(self.L.ekf_capture_frame_bgr(self.h, _ptr(img), img.shape[1], img.shape[0], img.strides[0], float(stamp)))
        else:
            self._ck(self.L.ekf_capture_frame(self.h, _ptr(img), img.shape[1], img.shape[0], img.strides[0], float(stamp)))

    def returnGrayImg(self):
        """vslamRansac.cpp:1364."""
        w, h = C.c_int(0), C.c_int(0)
        self._ck(self.L.ekf_get_frame(self.h, None, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), dtype=np.uint8)
        self._ck(self.L.ekf_get_frame(self.h, _ptr(out), C.byref(w), C.byref(h)))
        return out

    def captureNewFrame_device(self, dev_ptr, width, height, stride, stamp=-1.0):
        """Same for a frame already in device memory (raw pointer, e.g. tensor.data_ptr())."""
        self._ck(self.L.ekf_capture_frame_device(self.h, C.c_void_p(int(dev_ptr)), width, height, stride, float(stamp)))

    def predict(self, dv=(0.0, 0.0, 0.0), dw=(0.0, 0.0, 0.0), vcontrol=False):
        a = np.asarray(dv, dtype=np.float64); b = np.asarray(dw, dtype=np.float64)
        self._ck(self.L.ekf_predict(self.h, _ptr(a), _ptr(b), int(bool(vcontrol))))

    def match(self, want_count=True):
        n = C.c_int(0)
        self._ck(self.L.ekf_match(self.h, C.byref(n) if want_count else None))
        return n.value

    def update_after_match(self, picks=None):
        p = np.ascontiguousarray(picks if picks is not None else np.zeros(0), dtype=np.uint32)
        self._ck(self.L.ekf_update_after_match(self.h, _ptr(p), int(p.size)))

    def update(self, picks=None):
        """vslamRansac.cpp:868; `picks` replaces rand() (vslamRansac.cpp:989)."""
        p = np.ascontiguousarray(picks if picks is not None else np.zeros(0), dtype=np.uint32)
        self._ck(self.L.ekf_update(self.h, _ptr(p), int(p.size)))

    def inject_match(self, i, zu, zv, accepted=True):
        self._ck(self.L.ekf_inject_match(self.h, int(i), float(zu), float(zv), int(bool(accepted))))

    # ---- map management ----------------------------------------------------------------------
    def addFeature(self, u, v):
        return self._ck(self.L.ekf_add_feature(self.h, float(u), float(v)))

    def removeFeature(self, i):
        self._ck(self.L.ekf_remove_feature(self.h, int(i)))

    def findNewFeatures(self, num=-1):
        """vslamRansac.cpp:783-839; returns the number of features added."""
        return self._ck(self.L.ekf_find_new_features(self.h, int(num)))

    def detect_corners(self, num):
        out = np.zeros((1024, 2), dtype=np.float32); n = C.c_int(0)
        self._ck(self.L.ekf_detect_corners(self.h, int(num), _ptr(out), C.byref(n)))
        return out[:n.value].copy()

    def convert2XYZ_ifLinear(self, i):
        """vslamRansac.cpp:741-772."""
        self._ck(self.L.ekf_convert2xyz_if_linear(self.h, int(i)))

    def convert2XYZ_ifLinearAll(self):
        """vslamRansac.cpp:775-780."""
        self._ck(self.L.ekf_convert2xyz_if_linear_all(self.h))

    # ---- accessors ---------------------------------------------------------------------------
    def numOfFeatures(self):
        return self.L.ekf_num_features(self.h)

    def state_dim(self):
        return self.L.ekf_state_dim(self.h)

    def getState(self):
        out = np.zeros(14)
        self._ck(self.L.ekf_get_state(self.h, _ptr(out)))
        return out

    def getSigma(self):
        out = np.zeros((14, 14))
        self._ck(self.L.ekf_get_sigma(self.h, _ptr(out)))
        return out

    def Covariance_Parameter(self):
        v = C.c_double(0)
        self._ck(self.L.ekf_covariance_parameter(self.h, C.byref(v)))
        return v.value

    def getDt(self):
        return self.L.ekf_get_dt(self.h)

    def returnCentrPatchIndx(self, i):
        out = np.zeros(2, dtype=np.float32)
        self._ck(self.L.ekf_get_center(self.h, int(i), _ptr(out)))
        return out

    def feature(self, i):
        o = _abi.EkfFeatureInfo()
        self._ck(self.L.ekf_get_feature(self.h, int(i), C.byref(o)))
        return o

    def deleted(self):
        """VSlamFilter::deleted_patches (vslamRansac.cpp:394-404): list of (real_index, XYZ_pos[3], cov_4_delete[9])."""
        out = []
        for i in range(self.L.ekf_num_deleted(self.h)):
            o = _abi.EkfDeletedInfo()
            self._ck(self.L.ekf_get_deleted(self.h, i, C.byref(o)))
            out.append((o.real_index, np.array(o.xyz_pos), np.array(o.cov_4_delete)))
        return out

    def getPointsFeatures(self):
        """RosVSLAM::getPointsFeatures (RosVSLAMRansac.cpp:340-418): (last real_index + 1) x 12."""
        rows = C.c_int(0)
        self._ck(self.L.ekf_get_points_features(self.h, None, 0, C.byref(rows)))
        out = np.zeros((rows.value, 12))
        self._ck(self.L.ekf_get_points_features(self.h, _ptr(out), rows.value, C.byref(rows)))
        return out

    def rts_epoch(self, MU, SIGMA, MU_S, SIGMA_S, dTspeed, dRspeed, deltaT):
        """VSlamFilter::rts_epoch (vslamRansac.cpp:423-449) on 13-dimensional camera states; returns (MU, SIGMA)."""
        mu = np.array(MU, dtype=np.float64).copy(); sg = np.array(SIGMA, dtype=np.float64).reshape(13, 13).copy()
        mus = np.ascontiguousarray(MU_S, dtype=np.float64); sgs = np.ascontiguousarray(SIGMA_S, dtype=np.float64)
        a = np.ascontiguousarray(dTspeed, dtype=np.float64); b = np.ascontiguousarray(dRspeed, dtype=np.float64)
        assert mu.size == 13 and mus.size == 13 and sgs.size == 169
        self._ck(self.L.ekf_rts_epoch(self.h, _ptr(mu), _ptr(sg), _ptr(mus), _ptr(sgs), _ptr(a), _ptr(b), float(deltaT)))
        return mu, sg

    def template(self, i, which=0):
        w = self.cfg.window_size
        out = np.zeros((w, w), dtype=np.uint8)
        self._ck(self.L.ekf_get_template(self.h, int(i), int(which), _ptr(out)))
        return out

    def stats(self):
        s = _abi.EkfStepStats()
        self._ck(self.L.ekf_get_step_stats(self.h, C.byref(s)))
        return s

    def get_full(self):
        n = self.state_dim()
        mu = np.zeros(n); S = np.zeros((n, n))
        self._ck(self.L.ekf_get_full(self.h, _ptr(mu), _ptr(S), n))
        return mu, S

    def set_full(self, mu, S):
        mu = np.ascontiguousarray(mu, dtype=np.float64); S = np.ascontiguousarray(S, dtype=np.float64)
        n = self.state_dim()
        if mu.size != n or S.shape != (n, n):
            raise ValueError("set_full: shape mismatch")
        self._ck(self.L.ekf_set_full(self.h, _ptr(mu), _ptr(S), n))

    def S_blocks(self):
        out = np.zeros((self.numOfFeatures(), 2, 2))
        self._ck(self.L.ekf_get_S_blocks(self.h, _ptr(out)))
        return out

    def set_profiling(self, on=True):
        self._ck(self.L.ekf_set_profiling(self.h, int(bool(on))))

    def profile(self, reset=True):
        """dict class -> (ms, launches) accumulated since the last reset."""
        p = _abi.EkfProfile()
        self._ck(self.L.ekf_get_profile(self.h, C.byref(p), int(bool(reset))))
        return {name: (p.ms[i], p.launches[i]) for i, name in enumerate(_abi.PROF_CLASSES)}

    def set_symmetric_downdate(self, on=True):
        self._ck(self.L.ekf_set_symmetric_downdate(self.h, int(bool(on))))

    def debug_time_downdate(self, reps=20):
        """Mean duration (ms) of one covariance-downdate launch timed ALONE (diagnostic; perturbs Sigma by rounding errors)."""
        ms = C.c_float(0)
        self._ck(self.L.ekf_debug_time_downdate(self.h, int(reps), C.byref(ms)))
        return float(ms.value)

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.ekf_set_stream(self.h, C.c_void_p(int(cuda_stream_ptr) if cuda_stream_ptr else 0)))

    def sync(self):
        self._ck(self.L.ekf_sync(self.h))

    # ---- large map split across GPUs (row-block partitioned update, ekf_dist_*) -------------------
    def dist_attach(self, nccl_id: bytes, rank: int, world: int):
        buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
        self._ck(self.L.ekf_dist_attach(self.h, C.cast(buf, C.c_void_p), int(rank), int(world)))

    def dist_detach(self):
        self._ck(self.L.ekf_dist_detach(self.h))

    def dist_info(self):
        r, w, b = C.c_int(0), C.c_int(0), C.c_int64(0)
        self._ck(self.L.ekf_dist_info(self.h, C.byref(r), C.byref(w), C.byref(b)))
        return dict(rank=r.value, world=w.value, allgather_bytes=b.value, peer_memory=bool(self.L.ekf_dist_peer_memory(self.h)))


def nccl_unique_id(libnccl_path=None) -> bytes:
    """Loads libnccl (the copy torch uses unless a path is given) and returns a fresh 128-byte id."""
    load_nccl(libnccl_path)
    buf = (C.c_char * 128)()
    rc = lib().ekf_dist_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise EkfError(f"ekf_dist_unique_id failed: {rc}")
    return bytes(buf)


def load_nccl(libnccl_path=None):
    if libnccl_path is None:
        try:
            import os
            import nvidia.nccl as _n
            cand = os.path.join(os.path.dirname(_n.__file__), "lib", "libnccl.so.2")
            libnccl_path = cand if os.path.exists(cand) else None
        except Exception:
            libnccl_path = None
    rc = lib().ekf_dist_load_nccl(libnccl_path.encode() if libnccl_path else None)
    if rc != 0:
        raise EkfError(f"ekf_dist_load_nccl failed: {rc}")


def match_batch(frames_dev, n_frames, width, height, stride, templates_dev, features_per_frame, window,
                h_dev, S_dev, out_uv_dev, out_score_dev, sigma_size=3.0, ncc_threshold=0.8, search_clamp=20.0,
                stream=0):
    """ekf_match_batch on raw device pointers (ints)."""
    rc = lib().ekf_match_batch(C.c_void_p(frames_dev), n_frames, width, height, stride, C.c_void_p(templates_dev),
                               features_per_frame, window, C.c_void_p(h_dev), C.c_void_p(S_dev), float(sigma_size),
                               float(ncc_threshold), float(search_clamp), C.c_void_p(out_uv_dev),
                               C.c_void_p(out_score_dev), C.c_void_p(stream))
    if rc != 0:
        raise EkfError(f"ekf_match_batch failed: {rc}")


class FilterBatch:
    """B independent VSlamFilter instances stepped together on one device (BASELINE config 3).
    Seed it from a single VSlamFilter, optionally perturb the camera states, then call
    captureNewFrame / step once per frame."""

    def __init__(self, cfg, n_filters, feature_capacity=32, device=0):
        self.L = lib()
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.L.ekf_batch_create(C.byref(cfg), int(n_filters), int(feature_capacity), int(device), C.byref(h))
        if rc != 0:
            raise EkfError(f"ekf_batch_create failed with {rc}")
        self.h = h
        self.B = int(n_filters)

    def close(self):
        if getattr(self, "h", None):
            self.L.ekf_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise EkfError(f"ekf batch error {rc}: {self.L.ekf_batch_last_error(self.h).decode()}")
        return rc

    def describe(self):
        d = _abi.EkfBatchDesc()
        self._ck(self.L.ekf_batch_describe(self.h, C.byref(d)))
        return d

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.ekf_batch_set_stream(self.h, C.c_void_p(int(cuda_stream_ptr) if cuda_stream_ptr else 0)))

    def sync(self):
        self._ck(self.L.ekf_batch_sync(self.h))

    def seed_from(self, filt):
        self._ck(self.L.ekf_batch_seed_from(self.h, filt.h))

    def set_camera_states(self, mu14):
        a = np.ascontiguousarray(mu14, dtype=np.float64)
        if a.shape != (self.B, 14):
            raise ValueError("set_camera_states expects (n_filters, 14)")
        self._ck(self.L.ekf_batch_set_camera_states(self.h, _ptr(a)))

    def captureNewFrame(self, img, stamp=-1.0):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        self._ck(self.L.ekf_batch_capture_frame(self.h, _ptr(img), img.shape[1], img.shape[0], img.strides[0], float(stamp)))

    def captureNewFrame_device(self, dev_ptr, width, height, stride, stamp=-1.0):
        self._ck(self.L.ekf_batch_capture_frame_device(self.h, C.c_void_p(int(dev_ptr)), width, height, stride, float(stamp)))

    def step(self, picks=None, dv=(0.0, 0.0, 0.0), dw=(0.0, 0.0, 0.0), vcontrol=False):
        """predict + update of every filter."""
        p = np.ascontiguousarray(picks if picks is not None else np.zeros(0), dtype=np.uint32)
        a = np.asarray(dv, dtype=np.float64); b = np.asarray(dw, dtype=np.float64)
        self._ck(self.L.ekf_batch_step(self.h, _ptr(a), _ptr(b), int(bool(vcontrol)), _ptr(p), int(p.size)))

    def camera_states(self, want_sigma=False):
        mu = np.zeros((self.B, 14)); st = np.zeros((self.B, _abi.BATCH_STAT_FIELDS), dtype=np.int32)
        S = np.zeros((self.B, 14, 14)) if want_sigma else None
        self._ck(self.L.ekf_batch_get_camera_states(self.h, _ptr(mu), _ptr(S) if want_sigma else None, _ptr(st)))
        return (mu, S, st) if want_sigma else (mu, st)

    def numOfFeatures(self, f):
        return self._ck(self.L.ekf_batch_num_features(self.h, int(f)))

    def state_dim(self, f):
        return self._ck(self.L.ekf_batch_state_dim(self.h, int(f)))

    def get_full(self, f):
        n = self.state_dim(f)
        mu = np.zeros(n); S = np.zeros((n, n))
        self._ck(self.L.ekf_batch_get_full(self.h, int(f), _ptr(mu), _ptr(S), n))
        return mu, S

    def set_full(self, f, mu, S):
        mu = np.ascontiguousarray(mu, dtype=np.float64); S = np.ascontiguousarray(S, dtype=np.float64)
        n = self.state_dim(f)
        if mu.size != n or S.shape != (n, n):
            raise ValueError("set_full: shape mismatch")
        self._ck(self.L.ekf_batch_set_full(self.h, int(f), _ptr(mu), _ptr(S), n))

    def feature(self, f, i):
        o = _abi.EkfFeatureInfo()
        self._ck(self.L.ekf_batch_get_feature(self.h, int(f), int(i), C.byref(o)))
        return o

    def kernel_launches(self):
        return int(self.L.ekf_batch_kernel_launches(self.h))

    def last_match_deferred(self):
        """(filter, feature) pairs of the last step that the warp-per-feature matcher left to the CTA matcher."""
        return self._ck(self.L.ekf_batch_last_match_deferred(self.h))

    def last_step_ms(self):
        out = np.zeros(3, dtype=np.float32)
        self._ck(self.L.ekf_batch_last_step_ms(self.h, _ptr(out)))
        return dict(predict=float(out[0]), match=float(out[1]), update=float(out[2]))
