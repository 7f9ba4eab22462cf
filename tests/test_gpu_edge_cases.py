"""-m gpu: edge cases of the path against the oracle — empty map, frames in which nothing matches
(every candidate window flat -> NaN scores, vslamRansac.cpp:1062 "No Matching li"), features dropped by
the quality rule over several frames (ragged maps), a feature at infinity (rho <= 0, :517-522),
capacity and argument errors."""
import numpy as np
import pytest

from helpers import assert_state_close, assert_tables_equal, make_pair, seed_features

pytestmark = pytest.mark.gpu


def test_empty_map_steps(gpu_pkg, orc):
    sc = gpu_pkg.synth.Scene(n_features=4, n_frames=4, seed=1)
    g, o = make_pair(gpu_pkg, orc, sc)
    for t in range(0, 4):
        for f in (g, o):
            f.captureNewFrame(sc.frame(t), sc.stamps[t])
            if t:
                f.predict(dv=(0.01, 0.0, -0.01), dw=(0.0, 0.02, 0.0), vcontrol=bool(t & 1)); f.update(sc.picks(t, 4))
        assert g.numOfFeatures() == o.numOfFeatures() == 0 and g.state_dim() == 14
        assert_state_close(g, o, ctx=f"empty map frame {t}")
    assert g.stats().n_matched == 0 and g.stats().n_li == 0


def test_nothing_matches_then_features_are_dropped(gpu_pkg, orc):
    """A blank frame: every window is flat, every NCC is 0/0, every match is rejected; after enough
    misses update_quality_index (Patch.cpp:143-150) removes the features — in both implementations
    on the same frame, leaving identical ragged maps."""
    sc = gpu_pkg.synth.Scene(n_features=14, n_frames=3, seed=4)
    g, o = make_pair(gpu_pkg, orc, sc)
    seed_features(g, sc); seed_features(o, sc)
    for f in (g, o):                                  # one good frame first
        f.captureNewFrame(sc.frame(1), sc.stamps[1]); f.predict(); f.update(sc.picks(1, 14))
    assert_tables_equal(g, o, ctx="good frame")
    blank = np.full((sc.height, sc.width), 90, np.uint8)
    blank[::2, :] = 91                                # not flat globally, flat inside no window? rows alternate -> windows are NOT flat
    flat = np.full((sc.height, sc.width), 90, np.uint8)
    removed_any = False
    for k, img in enumerate([flat, flat, blank, flat]):
        stamp = sc.stamps[1] + (k + 1) / 30.0
        for f in (g, o):
            f.captureNewFrame(img, stamp); f.predict(); f.update(sc.picks(2 + k, 14))
        sg, so = g.stats(), o.stats()
        assert (sg.n_matched, sg.n_li, sg.n_hi, sg.n_removed) == (so.n_matched, so.n_li, so.n_hi, so.n_removed)
        if img is flat:
            assert sg.n_matched == 0
        removed_any |= sg.n_removed > 0
        assert g.numOfFeatures() == o.numOfFeatures()
        assert_tables_equal(g, o, ctx=f"bad frame {k}")
        assert_state_close(g, o, ctx=f"bad frame {k}")
    assert removed_any and g.numOfFeatures() < 14


def test_feature_at_infinity_is_removed(gpu_pkg, orc):
    sc = gpu_pkg.synth.Scene(n_features=10, n_frames=3, seed=12)
    g, o = make_pair(gpu_pkg, orc, sc)
    seed_features(g, sc); seed_features(o, sc)
    mu, S = o.get_full()
    mu[14 + 6 * 3 + 5] = -0.05                        # rho <= 0 (vslamRansac.cpp:517)
    mu[14 + 6 * 7 + 5] = 0.0
    for f in (g, o):
        f.set_full(mu, S)
        f.captureNewFrame(sc.frame(1), sc.stamps[1]); f.predict()
    assert [g.feature(i).remove_flag for i in range(10)] == [o.feature(i).remove_flag for i in range(10)]
    assert g.feature(3).remove_flag == 1 and g.feature(7).remove_flag == 1 and not g.feature(3).is_in_innovation
    for f in (g, o):
        f.update(sc.picks(1, 10))
    assert g.numOfFeatures() == o.numOfFeatures() == 8
    assert_tables_equal(g, o, ctx="after removing features at infinity")
    assert_state_close(g, o, ctx="after removing features at infinity")


def test_capacity_and_argument_errors(gpu_pkg):
    import ctypes as C
    sc = gpu_pkg.synth.Scene(n_features=6, n_frames=2, seed=3)
    cfg = gpu_pkg.default_config(**sc.config_overrides())
    f = gpu_pkg.VSlamFilter(cfg, feature_capacity=4)
    with pytest.raises(gpu_pkg.EkfError):
        f.addFeature(100.0, 100.0)                    # before captureNewFrame
    with pytest.raises(gpu_pkg.EkfError):
        f.predict()
    f.captureNewFrame(sc.frame(0), sc.stamps[0])
    for p in sc.feature_pixels[:4]:
        assert f.addFeature(*p) == 1
    with pytest.raises(gpu_pkg.EkfError, match="capacity"):
        f.addFeature(*sc.feature_pixels[4])
    with pytest.raises(gpu_pkg.EkfError):
        f.removeFeature(9)
    with pytest.raises(gpu_pkg.EkfError):
        f.update(sc.picks(1, 4))                      # update before predict
    L = gpu_pkg.lib()
    h = C.c_void_p()
    assert L.ekf_create(C.byref(cfg), 0, 0, C.byref(h)) == -1 and L.ekf_create(C.byref(cfg), 9000, 0, C.byref(h)) == -1
    assert L.ekf_create(C.byref(cfg), 8, 99, C.byref(h)) == -2       # no such device: EKF_ERR_CUDA, no fallback
    small = np.zeros((4, 4), np.uint8)
    with pytest.raises(gpu_pkg.EkfError):
        f.captureNewFrame(small, 2.0)


def test_camera_accessors_answer_from_the_step_record(gpu_pkg, orc):
    """getState / getSigma after update() are served from the packed result record (no device round trip): they must equal the
    device state bit for bit after the update, after removeFeature / addFeature (which leave the camera block alone) and — read from
    the device again — after predict and after set_full."""
    sc = gpu_pkg.synth.Scene(n_features=24, n_frames=4, seed=77)
    g, _ = make_pair(gpu_pkg, orc, sc)
    seed_features(g, sc)

    def check(ctx):
        mu, S = g.get_full()
        assert np.array_equal(g.getState(), mu[:14]), ctx
        assert np.array_equal(g.getSigma(), S[:14, :14]), ctx

    for t in (1, 2):
        g.captureNewFrame(sc.frame(t), sc.stamps[t]); g.predict()
        check(f"after predict {t}")
        g.update(sc.picks(t, 24))
        check(f"after update {t}")
    g.removeFeature(3)
    check("after removeFeature")
    g.addFeature(*sc.feature_pixels[3])
    check("after addFeature")
    mu, S = g.get_full()
    mu = mu.copy(); mu[0] += 0.25
    g.set_full(mu, S)
    check("after set_full")
