"""CPU: the oracle against the REFERENCE'S OWN sources (oracle/_ref: vslamRansac.cpp, Patch.cpp,
camModel.cpp, utils.cpp compiled unmodified against the API stand-ins of oracle/shim).  Runs only
where /root/reference exists (this container); tests/test_golden.py carries the same evidence to the
GPU box as committed fixtures.

fp64 variant (the reference with `float` re-typed to double) vs the oracle's all-double kind, and
the fp32 variant (as written) vs the all-float kind: integer tables must be identical and the state
must agree to rounding (observed: bit-identical — same operation order)."""
import numpy as np
import pytest

from helpers import INT_FIELDS, relerr

refbind = pytest.importorskip("refbind")
if not refbind.available():
    pytest.skip("reference sources not present (GPU box): see tests/test_golden.py", allow_module_level=True)

VARIANTS = [pytest.param(True, 2, 1e-12, id="fp64"), pytest.param(False, 1, 1e-5, id="fp32")]


@pytest.fixture(scope="module")
def ref():
    refbind.build()
    return refbind


def _pair(pkg, orc, ref, sc, fp64, kind, **over):
    o = sc.config_overrides()
    o["xyz_conversion"] = 1  # VSlamFilter::update always ends with convert2XYZ_ifLinearAll (vslamRansac.cpp:1317)
    o.update(over)
    cfg = pkg.default_config(**o)
    return ref.ReferenceFilter(cfg, fp64=fp64), orc.OracleFilter(cfg, kind=kind)


def _same_tables(r, o, ctx, skip=()):
    assert r.numOfFeatures() == o.numOfFeatures(), ctx
    for i in range(r.numOfFeatures()):
        a, b = r.feature(i), o.feature(i)
        for f in INT_FIELDS:
            if f in skip:
                continue
            assert getattr(a, f) == getattr(b, f), f"{ctx}: feature {i} {f}: reference {getattr(a, f)} oracle {getattr(b, f)}"
        assert tuple(a.center) == tuple(b.center), f"{ctx}: feature {i} match"
        assert np.array_equal(r.template(i, 0), o.template(i, 0))


def _same_state(r, o, tol, ctx):
    (mr, Sr), (mo, So) = r.get_full(), o.get_full()
    assert mr.shape == mo.shape, ctx
    assert relerr(mo, mr) <= tol and relerr(So, Sr) <= tol, f"{ctx}: mu {relerr(mo, mr):.2e} Sigma {relerr(So, Sr):.2e}"
    return relerr(mo, mr), relerr(So, Sr)


@pytest.mark.parametrize("fp64,kind,tol", VARIANTS)
@pytest.mark.parametrize("n_features,hard,seed", [(12, False, 3), (30, False, 41), (40, True, 9)])
def test_sequence(pkg, orc, ref, fp64, kind, tol, n_features, hard, seed):
    sc = pkg.synth.Scene(n_features=n_features, n_frames=7, seed=seed, hard=hard)
    r, o = _pair(pkg, orc, ref, sc, fp64, kind)
    for f in (r, o):
        f.captureNewFrame(sc.frame(0), sc.stamps[0])
        assert sum(f.addFeature(*p) for p in sc.feature_pixels) == n_features
    _same_state(r, o, tol, "after addFeature")
    worst = 0.0
    for t in range(1, sc.n_frames):
        for f in (r, o):
            f.captureNewFrame(sc.frame(t), sc.stamps[t]); f.predict()
        assert r.getDt() == o.getDt()
        _same_state(r, o, tol, f"frame {t} predict")
        assert relerr(o.S_blocks(), r.S_blocks()) <= tol
        for i in range(r.numOfFeatures()):
            a, b = r.feature(i), o.feature(i)
            assert a.is_in_innovation == b.is_in_innovation
            if a.is_in_innovation:
                assert relerr(list(b.h), list(a.h)) <= tol and relerr(list(b.H), list(a.H)) <= tol
        picks = sc.picks(t, n_features)
        r.update(picks); o.update(picks)
        assert r.last_hypotheses() == o.stats().ransac_hypotheses
        _same_tables(r, o, f"frame {t} update")
        worst = max(worst, *_same_state(r, o, tol, f"frame {t} update"))
    if hard:
        assert any(not o.feature(i).is_in_li for i in range(o.numOfFeatures())) or o.numOfFeatures() < n_features
    print(f"worst rel err vs reference: {worst:.2e}")


@pytest.mark.parametrize("fp64,kind,tol", VARIANTS)
def test_controls_remove_and_accessors(pkg, orc, ref, fp64, kind, tol):
    sc = pkg.synth.Scene(n_features=14, n_frames=4, seed=17)
    r, o = _pair(pkg, orc, ref, sc, fp64, kind)
    for f in (r, o):
        f.captureNewFrame(sc.frame(0), sc.stamps[0])
        for p in sc.feature_pixels:
            f.addFeature(*p)
        assert f.addFeature(2.0, 3.0) == 0           # outside the gate (vslamRansac.cpp:314)
        f.removeFeature(3); f.removeFeature(0); f.removeFeature(f.numOfFeatures() - 1)
    # Patch::position_in_z is not initialised by the reference's constructor (Patch.cpp:78-103): it is
    # garbage until the first predict assigns it (vslamRansac.cpp:589)
    _same_tables(r, o, "after remove", skip=("position_in_z",)); _same_state(r, o, tol, "after remove")
    for f in (r, o):
        f.captureNewFrame(sc.frame(1), sc.stamps[1])
        f.predict(dv=(0.01, -0.02, 0.005), dw=(0.002, 0.001, -0.003), vcontrol=True)
    _same_state(r, o, tol, "predict with controls")
    for f in (r, o):
        f.update(sc.picks(1, 14))
    _same_tables(r, o, "update"); _same_state(r, o, tol, "update")
    assert relerr(o.getState(), r.getState()) <= tol and relerr(o.getSigma(), r.getSigma()) <= tol
    assert abs(o.Covariance_Parameter() - r.Covariance_Parameter()) <= max(tol, 1e-6 if not fp64 else 0) * abs(r.Covariance_Parameter())


@pytest.mark.parametrize("fp64,kind,tol", VARIANTS)
def test_xyz_conversion(pkg, orc, ref, fp64, kind, tol):
    """convert2XYZ_ifLinear (vslamRansac.cpp:741-780): shrink rho's variance until the linearity index
    passes, convert, then run a step with mixed inverse-depth / XYZ features."""
    sc = pkg.synth.Scene(n_features=10, n_frames=4, seed=23)
    r, o = _pair(pkg, orc, ref, sc, fp64, kind)
    for f in (r, o):
        f.captureNewFrame(sc.frame(0), sc.stamps[0])
        for p in sc.feature_pixels:
            f.addFeature(*p)
        mu, S = f.get_full()
        for i in (1, 4, 9):
            pos = 14 + 6 * i
            S[pos + 5, :] *= 1e-3; S[:, pos + 5] *= 1e-3
        f.set_full(mu, S)
        f.convert2XYZ_ifLinearAll()
    assert o.state_dim() == 14 + 6 * 10 - 3 * 3
    assert [o.feature(i).coding for i in range(10)] == [0, 1, 0, 0, 1, 0, 0, 0, 0, 1]
    _same_tables(r, o, "after conversion", skip=("position_in_z",)); _same_state(r, o, tol, "after conversion")
    for t in (1, 2):
        for f in (r, o):
            f.captureNewFrame(sc.frame(t), sc.stamps[t]); f.predict(); f.update(sc.picks(t, 10))
        _same_tables(r, o, f"frame {t}"); _same_state(r, o, tol, f"frame {t}")


def test_find_match_standalone(pkg, orc, ref):
    """Patch::findMatch of the reference (fp32, as written) vs the oracle's batched matcher."""
    d = pkg.synth.match_batch_inputs(n_frames=1, features_per_frame=24, width=320, height=240, window=11, seed=5, s_diag=20.0)
    h = d["h"].copy(); S = d["S"].copy(); tm = d["templates"].copy()
    h[0] = (3.2, 4.9); h[1] = (317.5, 238.5); tm[2] = 128; S[3] = (400.0, 390.0, 390.0, 400.0)
    uv, _ = orc.match_batch(d["frames"], tm, h, S, sigma_size=3.0, kind_mf=0)
    for k in range(24):
        got = ref.find_match(d["frames"][0], tm[k], h[k], S[k], sigma_size=3.0, fp64=False)
        assert got == (uv[k, 0], uv[k, 1]), k


@pytest.mark.parametrize("fp64,kind,tol", VARIANTS)
def test_motion_blur_templates(pkg, orc, ref, fp64, kind, tol):
    """Patch::blur / blurPatch / evaluateKernel (Patch.cpp:50-57, libblur.cpp:17-81) with the reference's own
    libblur.cpp compiled in: kernel_size = 3 as conf_sim.cfg, T_camera = 0.5, a fast camera so that the
    predicted displacement over the exposure exceeds the threshold for most features."""
    sc = pkg.synth.Scene(n_features=16, n_frames=6, seed=29, speed=0.9, omega=0.5, template_smooth=2.5)
    r, o = _pair(pkg, orc, ref, sc, fp64, kind, kernel_size=3, T_camera=0.5)
    for f in (r, o):
        f.captureNewFrame(sc.frame(0), sc.stamps[0])
        for p in sc.feature_pixels:
            f.addFeature(*p)
    blurred = 0
    for t in range(1, sc.n_frames):
        for f in (r, o):
            f.captureNewFrame(sc.frame(t), sc.stamps[t]); f.predict()
        for i in range(r.numOfFeatures()):
            if o.feature(i).is_in_innovation:
                a, b = r.template(i, 1), o.template(i, 1)       # matching_patch after Patch::blur
                assert np.array_equal(a, b), f"frame {t} feature {i}: blurred template differs in {(a != b).sum()} pixels"
                blurred += int(not np.array_equal(b, o.template(i, 0)))
        picks = sc.picks(t, 16)
        r.update(picks); o.update(picks)
        _same_tables(r, o, f"frame {t} update"); _same_state(r, o, tol, f"frame {t} update")
    assert blurred >= 10, f"only {blurred} templates were blurred: the scene does not exercise the path"
    assert o.stats().blur_requests > 0


@pytest.mark.parametrize("fp64,kind,tol", VARIANTS)
def test_forse_plane(pkg, orc, ref, fp64, kind, tol):
    """forsePlane pseudo-measurement (vslamRansac.cpp:1245-1263, 1272)."""
    sc = pkg.synth.Scene(n_features=14, n_frames=5, seed=61, hard=True)
    r, o = _pair(pkg, orc, ref, sc, fp64, kind, forsePlane=1)
    for f in (r, o):
        f.captureNewFrame(sc.frame(0), sc.stamps[0])
        for p in sc.feature_pixels:
            f.addFeature(*p)
    for t in range(1, sc.n_frames):
        for f in (r, o):
            f.captureNewFrame(sc.frame(t), sc.stamps[t]); f.predict(); f.update(sc.picks(t, 14))
        _same_tables(r, o, f"frame {t}"); _same_state(r, o, tol, f"frame {t}")


def _rts_inputs(seed):
    rng = np.random.default_rng(seed)
    def state():
        mu = np.zeros(13)
        mu[:3] = rng.normal(0, 0.5, 3)
        q = rng.normal(0, 1, 4); mu[3:7] = q / np.linalg.norm(q)
        mu[7:10] = rng.normal(0, 0.2, 3); mu[10:13] = rng.normal(0, 0.1, 3)
        A = rng.normal(0, 1, (13, 13))
        return mu, A @ A.T * 1e-3 + np.eye(13) * 1e-4
    return state(), state(), rng.normal(0, 0.01, 3), rng.normal(0, 0.01, 3)


@pytest.mark.parametrize("fp64,kind,tol", VARIANTS)
@pytest.mark.parametrize("seed,zero_w", [(1, False), (2, False), (3, True)])
def test_rts_epoch(pkg, orc, ref, fp64, kind, tol, seed, zero_w):
    """VSlamFilter::rts_epoch (vslamRansac.cpp:423-449) on 13-dimensional camera states."""
    cfg = pkg.default_config()
    r, o = ref.ReferenceFilter(cfg, fp64=fp64), orc.OracleFilter(cfg, kind=kind)
    (mu, sg), (mus, sgs), dts, drs = _rts_inputs(seed)
    if zero_w:
        mu[10:13] = 0; drs[:] = 0   # |w| = 0 branch of Jacobian_qt_w (Sinc = 1, n_w = 0)
    mr, Sr = r.rts_epoch(mu, sg, mus, sgs, dts, drs, 1 / 30)
    mo, So = o.rts_epoch(mu, sg, mus, sgs, dts, drs, 1 / 30)
    assert relerr(mo, mr) <= tol and relerr(So, Sr) <= tol, f"mu {relerr(mo, mr):.2e} Sigma {relerr(So, Sr):.2e}"
    assert relerr(mo, mu) > 1e-6   # the epoch did something


@pytest.mark.parametrize("fp64,kind,tol", VARIANTS)
def test_deleted_archive(pkg, orc, ref, fp64, kind, tol):
    """removeFeature archives XYZ features seen more than five times (vslamRansac.cpp:394-404)."""
    sc = pkg.synth.Scene(n_features=10, n_frames=9, seed=77)
    r, o = _pair(pkg, orc, ref, sc, fp64, kind)
    for f in (r, o):
        f.captureNewFrame(sc.frame(0), sc.stamps[0])
        for p in sc.feature_pixels:
            f.addFeature(*p)
    for t in range(1, 9):
        for f in (r, o):
            f.captureNewFrame(sc.frame(t), sc.stamps[t]); f.predict(); f.update(sc.picks(t, 10))
    for f in (r, o):   # shrink rho's variance until the linearity index passes (as test_xyz_conversion)
        mu, S = f.get_full()
        for i in (1, 4, 9):
            pos = 14 + 6 * i
            S[pos + 5, :] *= 1e-4; S[:, pos + 5] *= 1e-4
        f.set_full(mu, S)
        f.convert2XYZ_ifLinearAll()
    xyz = [i for i in range(r.numOfFeatures()) if r.feature(i).coding]
    assert len(xyz) >= 2 and xyz == [i for i in range(o.numOfFeatures()) if o.feature(i).coding]
    # RosVSLAM::getPointsFeatures, from the reference's own RosVSLAMRansac.cpp (oracle/_ref), before and after removals
    Pr, Po = r.getPointsFeatures(), o.getPointsFeatures()
    assert Pr.shape == Po.shape and np.count_nonzero(np.abs(Pr).sum(axis=1)) == len(xyz)
    assert relerr(Po, Pr) <= tol, f"points matrix {relerr(Po, Pr):.2e}"
    # feature 0 is inverse-depth: removed, not archived.  The LAST feature stays: once it is gone the reference's
    # getPointsFeatures writes the archived rows with a larger real_index past its matrix (RosVSLAMRansac.cpp:349 sizes
    # it by the last live patch) — undefined behaviour that the oracle and the CUDA path replace by dropping those rows.
    victims = [i for i in xyz if i != r.numOfFeatures() - 1][:3]
    for i in sorted(victims + [0], reverse=True):
        r.removeFeature(i); o.removeFeature(i)
    dr, do = r.deleted(), o.deleted()
    assert len(dr) == len(do) == len(victims) >= 2
    Pr, Po = r.getPointsFeatures(), o.getPointsFeatures()
    assert Pr.shape == Po.shape and relerr(Po, Pr) <= tol, "points matrix with archived rows"
    assert np.count_nonzero(np.abs(Pr).sum(axis=1)) == len(xyz)
    # oracle only: removing the last (XYZ, archived) feature drops its row instead of overflowing
    last = o.numOfFeatures() - 1
    if o.feature(last).coding:
        o.removeFeature(last)
        P2 = o.getPointsFeatures()
        assert P2.shape[0] == o.feature(o.numOfFeatures() - 1).real_index + 1 and np.isfinite(P2).all()
    for (ia, xa, ca), (ib, xb, cb) in zip(dr, do):
        assert ia == ib
        assert relerr(xb, xa) <= tol and relerr(cb, ca) <= tol
