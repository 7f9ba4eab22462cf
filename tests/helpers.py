"""Shared helpers for the parity tests (oracle vs CUDA path on the same seeded inputs)."""
import numpy as np

# BASELINE.json tolerance: state and covariance within 1e-9 relative in fp64 (norm-wise, per step from
# identical inputs — SURVEY.md §7 "Hard parts").
TOL = 1e-9

INT_FIELDS = ("position_in_state", "position_in_z", "coding", "n_tot", "n_find", "real_index",
              "is_in_innovation", "is_in_li", "is_in_hi", "remove_flag")


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(a - b)
    s = np.linalg.norm(b)
    return d / s if s > 0 else d


def make_pair(pkg, orc, scene, capacity=None, **cfg_over):
    over = scene.config_overrides()
    over.update(cfg_over)
    cfg = pkg.default_config(**over)
    g = pkg.VSlamFilter(cfg, feature_capacity=capacity or max(8, scene.n_features + 8))
    o = orc.OracleFilter(cfg, kind=0, omp=True)
    return g, o


def make_oracle(pkg, orc, scene, kind=0, omp=False, **cfg_over):
    """A CPU oracle filter configured for `scene` (no GPU needed)."""
    over = scene.config_overrides()
    over.update(cfg_over)
    return orc.OracleFilter(pkg.default_config(**over), kind=kind, omp=omp)


def seed_features(filt, scene):
    filt.captureNewFrame(scene.frame(0), scene.stamps[0])
    return [filt.addFeature(*p) for p in scene.feature_pixels]


def assert_tables_equal(g, o, fields=INT_FIELDS, ctx=""):
    """Integer / flag state of every feature must be bit-identical."""
    assert g.numOfFeatures() == o.numOfFeatures(), f"{ctx}: feature count {g.numOfFeatures()} vs {o.numOfFeatures()}"
    for i in range(g.numOfFeatures()):
        a, b = g.feature(i), o.feature(i)
        for f in fields:
            assert getattr(a, f) == getattr(b, f), f"{ctx}: feature {i} field {f}: gpu {getattr(a, f)} oracle {getattr(b, f)}"


def assert_state_close(g, o, tol=TOL, ctx=""):
    mg, Sg = g.get_full()
    mo, So = o.get_full()
    assert mg.shape == mo.shape, ctx
    em, es = relerr(mg, mo), relerr(Sg, So)
    assert em <= tol, f"{ctx}: mu rel err {em:.3e}"
    assert es <= tol, f"{ctx}: Sigma rel err {es:.3e}"
    return em, es
