"""Caller-side data formats of the reference's ROS node, Python twin of
host/keyframe_recorder.hpp: the key-frame selector of ImageConverter::imageCb (monoslam_ransac.cpp:585-687;
:689-752 is commented out in the reference and is not reproduced) and the writers of nodes_and_prjcts.txt / cams_cov.txt / cams_cov2.txt / points.txt
(monoslam_ransac.cpp:232-236, 262-275) consumed by sparse_bundle_adjustment/src/nodes/sba_add.cpp:76-180.

Works with any object offering getState / getSigma / Covariance_Parameter / getPointsFeatures (the product's
VSlamFilter mirror, the oracle, or a stub in CPU tests).  Pure host logic: no arithmetic of the filter lives here."""
import math
import os

import numpy as np


def quat2vec(q):
    """monoslam_ransac.cpp:40-50."""
    n = math.acos(q[0]) * 2
    if n > 0.0001:
        n1 = n / math.sin(n / 2)
        return np.array([q[1] * n1, q[2] * n1, q[3] * n1])
    return np.zeros(3)


def poses_diff(state_old, state_new, last_rot):
    """monoslam_ransac.cpp:52-60."""
    d = np.asarray(state_old[:3], dtype=np.float64) - np.asarray(state_new[:3], dtype=np.float64)
    a = math.sqrt(float(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])) * 3.33
    b = (np.asarray(last_rot) - quat2vec(state_new[3:7])) * 57.29577951308232
    return a + abs(b[0]) + abs(b[1]) + abs(b[2])


def _cell(v, precision):
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    s = "%.*g" % (precision, float(v))
    if "e" in s:                       # ostream prints at least two exponent digits, as %g does; nothing to fix
        return s
    return s


def eigen_format(m, precision=6):
    """Eigen's default IOFormat: cells padded to the widest, ' ' between columns, one row per line."""
    m = np.asarray(m)
    if m.ndim == 1:
        m = m.reshape(-1, 1)
    cells = [[_cell(v, precision) for v in row] for row in m]
    width = max((len(c) for row in cells for c in row), default=0)
    return "\n".join(" ".join(c.rjust(width) for c in row) for row in cells)


class KeyframeRecorder:
    def __init__(self, directory=".", image_writer=None):
        self.dir = directory
        self.image_writer = image_writer      # callable(path, image) — cv2.imwrite in the ROS node
        os.makedirs(directory, exist_ok=True)
        self.node_proj = open(os.path.join(directory, "nodes_and_prjcts.txt"), "w")
        self.cov_cams = open(os.path.join(directory, "cams_cov.txt"), "w")
        self.cov_cams2 = open(os.path.join(directory, "cams_cov2.txt"), "w")
        self.MoveThresh = 18.0                 # :195
        self.Num_of_points_thershold = 10      # :186 (only read by the reference's commented-out rule)
        self.min_cov_for_pose = 10000000.0     # :187
        self.last_vrot = np.zeros(3)
        self.last_image_pose = np.zeros(7)
        self.Pose_id = 0
        self.min_stat = np.zeros(7)
        self.min_camscov = np.zeros((14, 14))
        self.min_projs = np.zeros((1, 3), dtype=np.int32)
        self.Selected_Pose = None
        self.key_frames = []

    def _write_node(self, fid, pose7, projs):
        self.node_proj.write("P%d\n" % fid)
        self.node_proj.write(eigen_format(np.asarray(pose7, dtype=np.float64)) + "\n")
        self.node_proj.write(("0  0  0" if projs is None else eigen_format(np.asarray(projs, dtype=np.int64))) + "\n")
        self.key_frames.append(int(fid))

    def _write_cov(self, f, S14):
        f.write(eigen_format(np.asarray(S14, dtype=np.float64)[:7, :7]) + "\n")

    def _save(self, fid, img):
        if self.image_writer is not None and img is not None:
            self.image_writer(os.path.join(self.dir, "%d.png" % fid), img)

    def _candidate(self, slam, fid, stat14, cov, img, with_cov):
        self.min_cov_for_pose = cov
        self.Pose_id = fid
        self.min_stat = np.array(stat14[:7], dtype=np.float64)
        if with_cov:
            self.min_camscov = np.array(slam.getSigma(), dtype=np.float64).reshape(14, 14)
        self.min_projs = np.array(getattr(slam, "Point4sba", np.zeros((1, 3), dtype=np.int32))).reshape(-1, 3)
        self.Selected_Pose = None if img is None else np.array(img).copy()

    def _set_last(self, stat14):
        self.last_vrot = quat2vec(stat14[3:7])
        self.last_image_pose = np.array(stat14[:7], dtype=np.float64)

    def on_frame(self, slam, frame_id, image=None):
        """One camera frame after slam.update() (monoslam_ransac.cpp:560, 585, 609-687)."""
        stat14 = np.asarray(slam.getState(), dtype=np.float64)
        dist = poses_diff(self.last_image_pose, stat14, self.last_vrot)
        if self.MoveThresh / 2 < dist < self.MoveThresh:
            some_var = slam.Covariance_Parameter()
            if some_var < self.min_cov_for_pose:
                self._candidate(slam, frame_id, stat14, some_var, image, True)
        elif dist >= self.MoveThresh:
            if self.min_cov_for_pose < 1000000:
                if (slam.Covariance_Parameter() - self.min_cov_for_pose) < 0.000085:
                    self._write_node(frame_id, stat14[:7], None)
                    S = np.array(slam.getSigma(), dtype=np.float64).reshape(14, 14)
                    self._write_cov(self.cov_cams2, S); self._write_cov(self.cov_cams, S)
                    self._save(frame_id, image)
                else:
                    self._write_node(self.Pose_id, self.min_stat, self.min_projs)
                    self._write_cov(self.cov_cams, self.min_camscov)
                    self._save(self.Pose_id, self.Selected_Pose)
                self._set_last(stat14)
            elif frame_id < 5:
                self._write_node(frame_id, stat14[:7], None)
                self.min_camscov = np.array(slam.getSigma(), dtype=np.float64).reshape(14, 14)
                self._write_cov(self.cov_cams, self.min_camscov)
                self._save(frame_id, image)
                self._set_last(stat14)
            self.min_cov_for_pose = 10000000.0
        # monoslam_ransac.cpp:689-752 (the Point4sba.rows() >= Num_of_points_thershold candidate rule and the
        # take_image_every_x_frame rule) sits inside a /* ... */ block in the reference: dead code, not reproduced.

    def finish(self, slam):
        """~ImageConverter (monoslam_ransac.cpp:262-275): close the files, write points.txt."""
        self.node_proj.close(); self.cov_cams.close(); self.cov_cams2.close()
        with open(os.path.join(self.dir, "points.txt"), "w") as f:
            f.write(eigen_format(np.asarray(slam.getPointsFeatures(), dtype=np.float64)))


def read_sba_inputs(directory):
    """The consumer's view (sba_add.cpp:76-180 restated): returns (points[rows,12], cams_cov[k,7,7], nodes) with
    nodes = list of (cam_index, pose7, [(point_index, u, v), ...]).  Token-based like the consumer's `>>`."""
    pts = np.array(open(os.path.join(directory, "points.txt")).read().split(), dtype=np.float64).reshape(-1, 12)
    cov = np.array(open(os.path.join(directory, "cams_cov.txt")).read().split(), dtype=np.float64).reshape(-1, 7, 7)
    tok = open(os.path.join(directory, "nodes_and_prjcts.txt")).read().split()
    nodes, i = [], 0
    while i < len(tok):
        assert tok[i][0] == "P", tok[i]
        cam = int(tok[i][1:]); i += 1
        pose = [float(t) for t in tok[i:i + 7]]; i += 7
        projs = []
        while i < len(tok) and tok[i][0] != "P":
            projs.append(tuple(int(t) for t in tok[i:i + 3])); i += 3
        nodes.append((cam, pose, projs))
    return pts, cov, nodes
