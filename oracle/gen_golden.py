#!/usr/bin/env python
"""oracle/gen_golden.py — TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the REFERENCE'S OWN sources (oracle/_ref/libref_f64.so: the
reference compiled unmodified against oracle/shim with `float` re-typed to double; see
oracle/ref_capi.cpp) on seeded synthetic sequences.  Run in the container that has /root/reference:

    python oracle/gen_golden.py

The fixtures travel to the GPU box, where tests/test_golden.py (CPU oracle) and
tests/test_gpu_golden.py (CUDA path) replay the same inputs and compare.
Per step the fixture keeps the full state vector, the 14x14 camera block and the Frobenius norm of
Sigma, every feature's integer table and match; the full Sigma is kept for the last step.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, HERE, os.path.join(ROOT, "tests")]
import ekfb200  # noqa: E402
import refbind  # noqa: E402

INT_FIELDS = ("position_in_state", "position_in_z", "coding", "n_tot", "n_find", "real_index",
              "is_in_innovation", "is_in_li", "is_in_hi", "remove_flag")

# name -> Scene kwargs, extra config, scripted actions
CASES = {
    "seq_n12_easy": dict(scene=dict(n_features=12, n_frames=6, seed=3), controls=False),
    "seq_n30_easy": dict(scene=dict(n_features=30, n_frames=5, seed=41), controls=False),
    "seq_n40_hard": dict(scene=dict(n_features=40, n_frames=7, seed=9, hard=True), controls=False),
    "seq_n16_controls": dict(scene=dict(n_features=16, n_frames=5, seed=17), controls=True),
    "seq_n10_xyz": dict(scene=dict(n_features=10, n_frames=4, seed=23), controls=False, xyz=(1, 4, 9)),
    # motion-blur templates (kernel_size = 3, T_camera = 0.5 as conf_sim.cfg) with a fast camera
    "seq_n16_blur": dict(scene=dict(n_features=16, n_frames=5, seed=29, speed=0.9, omega=0.5, template_smooth=2.5), controls=False,
                         cfg=dict(kernel_size=3, T_camera=0.5)),
}


def controls_for(t):
    rng = np.random.default_rng([515, t])
    return rng.normal(scale=0.01, size=3), rng.normal(scale=0.003, size=3), bool(t % 2)


def run_case(pkg, name, spec, make_filter):
    """Replays the scripted case on any filter object with the VSlamFilter method names and returns
    the record dict (shared by the generator and the replaying tests)."""
    sc = pkg.synth.Scene(**spec["scene"])
    over = sc.config_overrides()
    over["xyz_conversion"] = 1  # the reference always converts at the end of update (vslamRansac.cpp:1317)
    over.update(spec.get("cfg", {}))
    f = make_filter(over)
    N = sc.n_features
    f.captureNewFrame(sc.frame(0), sc.stamps[0])
    added = [f.addFeature(*p) for p in sc.feature_pixels]
    if spec.get("xyz"):
        mu, S = f.get_full()
        for i in spec["xyz"]:
            pos = 14 + 6 * i
            S[pos + 5, :] *= 1e-3; S[:, pos + 5] *= 1e-3
        f.set_full(mu, S)
        f.convert2XYZ_ifLinearAll()
    rec = {"added": np.array(added, dtype=np.int32)}
    mu, S = f.get_full()
    rec["mu_init"] = mu; rec["Sigma_init"] = S
    h = hashlib.sha256()
    for t in range(1, sc.n_frames):
        img = sc.frame(t)
        h.update(img.tobytes())
        f.captureNewFrame(img, sc.stamps[t])
        if spec["controls"]:
            dv, dw, vc = controls_for(t)
            f.predict(dv=dv, dw=dw, vcontrol=vc)
        else:
            f.predict()
        mu, S = f.get_full()
        rec[f"t{t}_pred_mu"] = mu; rec[f"t{t}_pred_S14"] = S[:14, :14].copy(); rec[f"t{t}_pred_Sfro"] = np.linalg.norm(S)
        rec[f"t{t}_pred_Sblocks"] = f.S_blocks()
        rec[f"t{t}_pred_h"] = np.array([list(f.feature(i).h) for i in range(f.numOfFeatures())])
        rec[f"t{t}_pred_innov"] = np.array([f.feature(i).is_in_innovation for i in range(f.numOfFeatures())], dtype=np.int32)
        f.update(sc.picks(t, N))
        mu, S = f.get_full()
        nf = f.numOfFeatures()
        rec[f"t{t}_mu"] = mu; rec[f"t{t}_S14"] = S[:14, :14].copy(); rec[f"t{t}_Sfro"] = np.linalg.norm(S)
        rec[f"t{t}_Sdiag"] = np.diag(S).copy()
        rec[f"t{t}_tab"] = np.array([[getattr(f.feature(i), fld) for fld in INT_FIELDS] for i in range(nf)], dtype=np.int32).reshape(nf, len(INT_FIELDS))
        rec[f"t{t}_center"] = np.array([list(f.feature(i).center) for i in range(nf)], dtype=np.float32).reshape(nf, 2)
        rec[f"t{t}_covpar"] = f.Covariance_Parameter()
    rec["Sigma_last"] = S
    rec["frames_sha256"] = np.frombuffer(h.digest(), dtype=np.uint8)
    return rec


def main():
    pkg = ekfb200.load_package()
    refbind.build()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, spec in CASES.items():
        rec = run_case(pkg, name, spec, lambda over: refbind.ReferenceFilter(pkg.default_config(**over), fp64=True))
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **rec)
        tabs = [k for k in rec if k.endswith("_tab")]
        ncode = int(rec[tabs[-1]][:, 2].sum()) if tabs else 0
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, final n = {rec['Sigma_last'].shape[0]}, "
              f"features {rec[tabs[-1]].shape[0]}, XYZ-coded {ncode}")


if __name__ == "__main__":
    main()
