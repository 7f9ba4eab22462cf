"""CPU: caller-side formats — the key-frame selector and the writers of
nodes_and_prjcts.txt / cams_cov.txt / points.txt (monoslam_ransac.cpp:585-687, 232-275), Python twin and C++
twin (host/keyframe_recorder.hpp) on the same trajectory, read back through a restatement of the consumer's
parser (sba_add.cpp:76-180).  The comparison with the reference's OWN selector (compiled from monoslam_ransac.cpp) is in
tests/test_keyframe_pinned.py."""
import os
import struct
import subprocess

import numpy as np
import pytest


class Stub:
    """Plays back a trajectory through the accessors the recorder reads."""

    def __init__(self, states, sigmas, covs, points):
        self.states, self.sigmas, self.covs, self.points = states, sigmas, covs, points
        self.t = 0
        self.Point4sba = np.zeros((1, 3), dtype=np.int32)

    def getState(self): return self.states[self.t]
    def getSigma(self): return self.sigmas[self.t]
    def Covariance_Parameter(self): return self.covs[self.t]
    def getPointsFeatures(self): return self.points


def _trajectory(n=60, seed=5):
    rng = np.random.default_rng(seed)
    states, sigmas, covs = [], [], []
    for t in range(n):
        s = np.zeros(14)
        s[:3] = [0.55 * t, 0.05 * np.sin(0.3 * t), 0.02 * t]
        ang = np.pi + 0.01 * t                      # starts at the reference's initial attitude (180 deg about (0,-1,1)/sqrt2)
        ax = np.array([0.0, -1.0, 1.0]) / np.sqrt(2)
        s[3] = np.cos(ang / 2); s[4:7] = np.sin(ang / 2) * ax
        s[7:13] = rng.normal(0, 0.1, 6); s[13] = 1.0
        A = rng.normal(0, 1e-2, (14, 14))
        states.append(s); sigmas.append(A @ A.T + np.eye(14) * 1e-6)
        covs.append(1e-3 * (1.5 + np.sin(0.9 * t)))  # wiggles so that both key-frame sub-branches occur
    pts = np.zeros((9, 12))
    pts[2] = rng.normal(0, 1, 12); pts[5] = rng.normal(0, 100, 12); pts[8] = rng.normal(0, 1e-4, 12)
    return states, sigmas, covs, pts


def test_eigen_format_matches_eigen_default_ioformat(pkg):
    kf = pkg.keyframes
    assert kf.eigen_format(np.array([[1.0, 2.5], [-3.0, 4.0]])) == "  1 2.5\n -3   4"
    assert kf.eigen_format(np.array([0.123456789, 100.0, 1e-7])) == "0.123457\n     100\n   1e-07"
    assert kf.eigen_format(np.zeros((1, 3), dtype=np.int32)) == "0 0 0"


def test_poses_diff_and_quat2vec(pkg):
    kf = pkg.keyframes
    assert np.allclose(kf.quat2vec([1.0, 0, 0, 0]), 0)
    v = kf.quat2vec([np.cos(0.25), np.sin(0.25), 0, 0])            # rotation of 0.5 rad about x
    assert np.allclose(v, [0.5, 0, 0])
    d = kf.poses_diff(np.zeros(7), np.array([3.0, 4.0, 0, np.cos(0.25), np.sin(0.25), 0, 0]), np.zeros(3))
    assert abs(d - (5 * 3.33 + 0.5 * 57.29577951308232)) < 1e-9


def _run_python(pkg, tmp, traj):
    states, sigmas, covs, pts = traj
    stub = Stub(states, sigmas, covs, pts)
    saved = []
    rec = pkg.keyframes.KeyframeRecorder(str(tmp), image_writer=lambda path, img: saved.append(path))
    for t in range(len(states)):
        stub.t = t
        rec.on_frame(stub, t + 1, np.zeros((2, 2), dtype=np.uint8))
    rec.finish(stub)
    return rec, saved


def test_selector_and_files(pkg, tmp_path):
    traj = _trajectory()
    rec, saved = _run_python(pkg, tmp_path, traj)
    pts, cov, nodes = pkg.keyframes.read_sba_inputs(str(tmp_path))
    assert len(nodes) >= 4 and len(nodes) == cov.shape[0] == len(rec.key_frames) == len(saved)
    assert nodes[0][0] == 1                                        # "Taking current Pose": the first frame (:668-682)
    assert [os.path.basename(p) for p in saved] == ["%d.png" % n[0] for n in nodes]
    states, sigmas, covs, P = traj
    kinds = set()
    for (cam, pose, projs), C in zip(nodes, cov):
        assert np.allclose(pose, states[cam - 1][:7], rtol=1e-5, atol=1e-12)      # 6 significant digits
        assert np.allclose(C, sigmas[cam - 1][:7, :7], rtol=1e-5, atol=1e-12)
        assert projs == [(0, 0, 0)]
        kinds.add(cam)
    assert pts.shape == P.shape and np.allclose(pts, P, rtol=1e-5, atol=1e-12)
    # both sub-branches of the DistWalked >= MoveThresh case were taken: some key frames are the frame that crossed
    # the threshold, others an earlier (minimum-covariance) candidate
    text = open(tmp_path / "nodes_and_prjcts.txt").read()
    assert "0  0  0" in text and "\n0 0 0\n" in text


def test_cpp_twin_writes_identical_files(pkg, tmp_path):
    traj = _trajectory(seed=11)
    pydir = tmp_path / "py"; cdir = tmp_path / "cpp"
    pydir.mkdir(); cdir.mkdir()
    _, saved = _run_python(pkg, pydir, traj)
    states, sigmas, covs, pts = traj
    blob = struct.pack("i", len(states))
    for s, S, c in zip(states, sigmas, covs):
        blob += np.asarray(s, dtype=np.float64).tobytes() + np.asarray(S, dtype=np.float64).tobytes() + struct.pack("d", c)
    blob += struct.pack("i", pts.shape[0]) + np.asarray(pts, dtype=np.float64).tobytes()
    (tmp_path / "traj.bin").write_bytes(blob)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "keyframe_stub"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-o", str(exe), os.path.join(root, "tests", "keyframe_stub.cpp")])
    out = subprocess.check_output([str(exe), str(tmp_path / "traj.bin"), str(cdir)], text=True)
    assert [os.path.basename(p) for p in out.split()] == [os.path.basename(p) for p in saved]
    for name in ("nodes_and_prjcts.txt", "cams_cov.txt", "cams_cov2.txt", "points.txt"):
        assert (pydir / name).read_bytes() == (cdir / name).read_bytes(), name


def test_recorder_driven_by_the_oracle_filter(pkg, orc, tmp_path):
    """End to end on the CPU: the oracle filter runs a short sequence, the recorder selects key frames from its accessors and
    the files read back as the states / covariances the filter had at those frames; points.txt is getPointsFeatures()."""
    from helpers import make_oracle, seed_features
    kf = pkg.keyframes
    sc = pkg.synth.Scene(n_features=20, n_frames=16, seed=13)
    o = make_oracle(pkg, orc, sc, xyz_conversion=1)
    seed_features(o, sc)
    rec = kf.KeyframeRecorder(str(tmp_path))
    rec.MoveThresh = 0.6           # the synthetic camera moves centimetres per frame
    states, sigmas = {}, {}
    for t in range(1, 16):
        o.captureNewFrame(sc.frame(t), sc.stamps[t]); o.predict(); o.update(sc.picks(t, 20))
        states[t], sigmas[t] = o.getState(), o.getSigma()
        rec.on_frame(o, t, sc.frame(t))
    rec.finish(o)
    pts, cov, nodes = kf.read_sba_inputs(str(tmp_path))
    assert len(nodes) >= 2 and len(nodes) == cov.shape[0] and nodes[0][0] == 1
    for (cam, pose, projs), C in zip(nodes, cov):
        assert np.allclose(pose, states[cam][:7], rtol=1e-5, atol=1e-9)
        assert np.allclose(C, sigmas[cam][:7, :7], rtol=1e-5, atol=1e-15)
    assert pts.shape == o.getPointsFeatures().shape
