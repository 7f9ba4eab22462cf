#!/usr/bin/env python
"""oracle/gen_golden_export.py — TEST INFRASTRUCTURE ONLY.

Golden vectors for the caller-side rows of SURVEY.md section 8(f) item 4, produced by the REFERENCE'S OWN sources
(oracle/_ref/libref_f64.so: vslamRansac.cpp and RosVSLAMRansac.cpp compiled unmodified against oracle/shim):
  * VSlamFilter::rts_epoch (vslamRansac.cpp:423-449) on three seeded 13-dimensional states (one with |w| = 0);
  * RosVSLAM::getPointsFeatures (RosVSLAMRansac.cpp:340-418) and the archive of removed features
    (vslamRansac.cpp:394-404) on a scripted sequence with XYZ conversions and removals.
Run where /root/reference exists:  python oracle/gen_golden_export.py  ->  tests/golden/export_rts_points.npz
tests/test_golden_export.py (CPU oracle) and tests/test_gpu_export.py (CUDA path) replay it anywhere.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, HERE, os.path.join(ROOT, "tests")]

RTS_SEEDS = ((1, False), (2, False), (3, True))
SCENE = dict(n_features=10, n_frames=9, seed=77)
XYZ = (1, 4, 9)


def rts_inputs(seed, zero_w):
    rng = np.random.default_rng(seed)

    def state():
        mu = np.zeros(13)
        mu[:3] = rng.normal(0, 0.5, 3)
        q = rng.normal(0, 1, 4); mu[3:7] = q / np.linalg.norm(q)
        mu[7:10] = rng.normal(0, 0.2, 3); mu[10:13] = rng.normal(0, 0.1, 3)
        A = rng.normal(0, 1, (13, 13))
        return mu, A @ A.T * 1e-3 + np.eye(13) * 1e-4

    (mu, sg), (mus, sgs) = state(), state()
    dts, drs = rng.normal(0, 0.01, 3), rng.normal(0, 0.01, 3)
    if zero_w:
        mu[10:13] = 0; drs[:] = 0
    return mu, sg, mus, sgs, dts, drs, 1 / 30


def run_export_case(pkg, make_filter):
    """Replays the scripted case on any filter object with the VSlamFilter method names (reference, oracle, CUDA path)."""
    rec = {}
    f = make_filter(dict())
    for k, (seed, zero_w) in enumerate(RTS_SEEDS):
        m, S = f.rts_epoch(*rts_inputs(seed, zero_w))
        rec[f"rts{k}_mu"] = np.asarray(m); rec[f"rts{k}_sigma"] = np.asarray(S)
    sc = pkg.synth.Scene(**SCENE)
    over = sc.config_overrides(); over["xyz_conversion"] = 1
    f = make_filter(over)
    f.captureNewFrame(sc.frame(0), sc.stamps[0])
    for p in sc.feature_pixels:
        f.addFeature(*p)
    for t in range(1, sc.n_frames):
        f.captureNewFrame(sc.frame(t), sc.stamps[t]); f.predict(); f.update(sc.picks(t, sc.n_features))
    mu, S = f.get_full()
    for i in XYZ:
        pos = 14 + 6 * i
        S[pos + 5, :] *= 1e-4; S[:, pos + 5] *= 1e-4
    f.set_full(mu, S)
    f.convert2XYZ_ifLinearAll()
    rec["coding"] = np.array([f.feature(i).coding for i in range(f.numOfFeatures())], dtype=np.int32)
    rec["points_before"] = np.asarray(f.getPointsFeatures())
    for i in (4, 1, 0):   # two archived XYZ features and one inverse-depth feature; the last feature stays
        f.removeFeature(i)
    d = f.deleted()
    rec["deleted_index"] = np.array([x[0] for x in d], dtype=np.int32)
    rec["deleted_xyz"] = np.array([x[1] for x in d]).reshape(len(d), 3)
    rec["deleted_cov"] = np.array([x[2] for x in d]).reshape(len(d), 9)
    rec["points_after"] = np.asarray(f.getPointsFeatures())
    return rec


def main():
    import ekfb200
    import refbind
    pkg = ekfb200.load_package()
    refbind.build()
    rec = run_export_case(pkg, lambda over: refbind.ReferenceFilter(pkg.default_config(**over), fp64=True))
    path = os.path.join(ROOT, "tests", "golden", "export_rts_points.npz")
    np.savez_compressed(path, **rec)
    print(f"{path}: {os.path.getsize(path) / 1024:.1f} KiB, coding {rec['coding'].tolist()}, archived {rec['deleted_index'].tolist()}, "
          f"points {rec['points_before'].shape} -> {rec['points_after'].shape}")


if __name__ == "__main__":
    main()
