#!/usr/bin/env python
"""oracle/gen_golden_keyframes.py — TEST INFRASTRUCTURE ONLY.  Writes tests/golden/keyframes/<trajectory>/ with the files the
REFERENCE'S OWN key-frame selector (oracle/_ref/libref_selector_f64.so, compiled from monoslam_ransac.cpp:585, 609-687 by
oracle/build_ref_selector.py) produces on the trajectories of tests/test_keyframe_pinned.py, plus images.txt (the names
passed to cv::imwrite, in order).  Needs /root/reference; the outputs are committed so that the comparison also runs
where the reference is absent."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import refselbind  # noqa: E402
from test_keyframe_pinned import FILES, trajectories  # noqa: E402

if __name__ == "__main__":
    base = os.path.join(ROOT, "tests", "golden", "keyframes")
    for name, traj in trajectories().items():
        d = os.path.join(base, name)
        os.makedirs(d, exist_ok=True)
        imgs = refselbind.run(d, traj[0], traj[1], traj[2])
        with open(os.path.join(d, "images.txt"), "w") as f:
            f.write("\n".join(imgs) + "\n")
        print(name, len(imgs), "key frames", [os.path.getsize(os.path.join(d, x)) for x in FILES])
