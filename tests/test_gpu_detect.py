"""-m gpu: findNewFeatures (vslamRansac.cpp:783-839) — the reference's mask + a Shi-Tomasi detector
standing in for OpenCV goodFeaturesToTrack(frame, features, num, 0.01, 12, mask).  The arithmetic of
that call lives in an un-vendored, unpinned dependency of the reference (OpenCV): PARITY UNPINNED;
the checker here is the cv2 wheel of this image with the reference's arguments."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _reference_mask(shape, centers, w):
    """vslamRansac.cpp:788-818."""
    H, W = shape
    mask = np.zeros((H, W), np.uint8)
    mask[w:H - w, w:W - w] = 255
    win = 2 * w + 1
    for cx, cy in centers:
        if cx > w and cy > w and cx < W - w and cy < H - w:
            x = int(cx - win // 2) if cx - win // 2 > 0 else 0
            y = int(cy - win // 2) if cy - win // 2 > 0 else 0
            mask[y:y + win, x:x + win] = 0
    return mask


def _textured(seed, W=640, H=480):
    """Smooth random texture with a few hundred distinct corners (blurred noise + rectangles)."""
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.integers(0, 256, (H, W)).astype(np.float32), (0, 0), 2.5)
    img = (img - img.min()) / (img.max() - img.min()) * 160 + 40
    for _ in range(60):
        x, y = int(rng.integers(20, W - 60)), int(rng.integers(20, H - 60))
        img[y:y + int(rng.integers(8, 40)), x:x + int(rng.integers(8, 40))] += rng.uniform(-40, 40)
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("seed,num", [(1, 30), (2, 100), (3, 12), (4, 200), (5, 60), (6, 150), (8, 300)])
def test_detector_agrees_with_cv2(gpu_pkg, seed, num):
    """EXACT: every one of the `num` corners, in order.  Order rule (both sides): descending min-eigenvalue score as float bits;
    equal scores: the pixel with the higher address (row-major index) first — OpenCV's greaterThanPtr, here the low 32 bits of the
    (score bits << 32 | pixel index) sort key (csrc/ekf_detect.cu); then the greedy 12 px minimum-distance selection in that order."""
    img = _textured(seed)
    cfg = gpu_pkg.default_config(window_size=11, xyz_conversion=0, min_features=0)
    f = gpu_pkg.VSlamFilter(cfg, feature_capacity=512)
    f.captureNewFrame(img, 1.0)
    seeds = [(100.0, 100.0), (320.0, 240.0), (500.0, 400.0)]
    for p in seeds:
        assert f.addFeature(*p) == 1
    got = f.detect_corners(num)
    mask = _reference_mask(img.shape, seeds, 11)
    want = cv2.goodFeaturesToTrack(img, num, 0.01, 12, mask=mask).reshape(-1, 2)
    assert len(got) == len(want) == num
    assert np.array_equal(got, want), f"corner lists differ: first difference at rank {int(np.argmax((got != want).any(axis=1)))}"
    # invariants of the reference's call: minimum distance 12 px, inside the mask
    d = np.linalg.norm(got[:, None, :] - got[None, :, :], axis=2) + 1e9 * np.eye(len(got))
    assert d.min() >= 12.0
    assert all(mask[int(y), int(x)] for x, y in got)


def test_find_new_features_adds_them(gpu_pkg):
    img = _textured(7)
    cfg = gpu_pkg.default_config(window_size=11, xyz_conversion=0, min_features=0, nInitFeatures=5)
    f = gpu_pkg.VSlamFilter(cfg, feature_capacity=64)
    f.captureNewFrame(img, 1.0)
    want = cv2.goodFeaturesToTrack(img, 10, 0.01, 12, mask=_reference_mask(img.shape, [], 11)).reshape(-1, 2)
    assert f.findNewFeatures(10) == 10              # first frame of the ROS node: findNewFeatures(10), monoslam_ransac.cpp:406
    assert f.numOfFeatures() == 10 and f.state_dim() == 14 + 60
    centers = np.array([f.returnCentrPatchIndx(i) for i in range(10)])
    assert np.array_equal(centers, want)
    assert f.findNewFeatures(-1) == 5               # num <= 0 -> nInitFeatures (vslamRansac.cpp:786)
    c2 = np.array([f.returnCentrPatchIndx(i) for i in range(10, 15)])
    d = np.linalg.norm(c2[:, None, :] - centers[None, :, :], axis=2)
    assert d.min() > 11                             # new ones stay out of the boxes around the existing patches


def test_update_tops_up_when_too_few_features_are_visible(gpu_pkg):
    """vslamRansac.cpp:1312-1315: fewer than min_features visible after the update ->
    findNewFeatures(min_features - visible)."""
    img = _textured(11)
    cfg = gpu_pkg.default_config(window_size=11, xyz_conversion=0, min_features=15, max_features=100, sigma_size=3,
                                 fx=525.0, fy=525.0, u0=320.0, v0=240.0, k1=0.0, k2=0.0, p1=0.0, p2=0.0, T_camera=0.0)
    f = gpu_pkg.VSlamFilter(cfg, feature_capacity=64)
    f.captureNewFrame(img, 1.0)
    assert f.findNewFeatures(6) == 6
    f.captureNewFrame(img, 1.0 + 1 / 30)          # the camera does not move: every template matches where it was
    f.predict(); f.update(np.arange(8, dtype=np.uint32))
    st = f.stats()
    assert st.n_matched == 6 and st.topup_request == 9
    assert f.numOfFeatures() == 15 and f.state_dim() == 14 + 6 * 15
    mu, S = f.get_full()
    assert np.isfinite(mu).all() and np.isfinite(S).all() and np.all(np.diag(S) > 0)


@pytest.mark.parametrize("scale,shape,color", [(2, (480, 640), False), (2, (480, 640), True), (10, (1080, 1920), True),
                                               (3, (481, 643), False), (1, (240, 320), True)])
def test_capture_resize_and_gray_match_cv2(gpu_pkg, scale, shape, color):
    """captureNewFrame (vslamRansac.cpp:234-245): cv::resize by 1 / scale, then BGR2GRAY — bit-exact against cv2."""
    rng = np.random.default_rng(scale * 7 + int(color))
    H, W = shape
    img = rng.integers(0, 256, (H, W, 3) if color else (H, W), dtype=np.uint8)
    cfg = gpu_pkg.default_config(window_size=11, xyz_conversion=0, min_features=0, scale=scale)
    f = gpu_pkg.VSlamFilter(cfg, feature_capacity=8)
    f.captureNewFrame(img, 1.0)
    want = cv2.resize(img, (W // scale, H // scale))
    if color:
        want = cv2.cvtColor(want, cv2.COLOR_BGR2GRAY)
    got = f.returnGrayImg()
    assert got.shape == want.shape
    assert np.array_equal(got, want), f"{(got != want).sum()} of {want.size} pixels differ (max {np.abs(got.astype(int) - want).max()})"
    # the filter works on the resized frame: the in-image gate of addFeature uses its size
    assert f.addFeature(W // scale - 20.0, H // scale - 20.0) == 1
    assert f.addFeature(W // scale + 5.0, 30.0) == 0
