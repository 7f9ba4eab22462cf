#!/usr/bin/env python
"""oracle/build_ref_selector.py — TEST INFRASTRUCTURE ONLY.

Builds oracle/_ref/libref_selector_f64.so: the reference's own key-frame selector (the live part of
ImageConverter::imageCb, mono-slam/src/monoslam_ransac.cpp:585, 609-687, with quat2vec / poses_diff, :40-60) pasted at
BUILD time from the unmodified source under /root/reference into oracle/ref_selector_harness.cpp.in and compiled against
the API stand-ins of oracle/shim/.  The generated translation unit lives in oracle/_ref/ (git-ignored): no reference
source is stored in this repository.  The line ranges are checked against marker text so that a different revision of the
reference fails loudly instead of pinning the wrong lines.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("EKF_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "mono-slam", "src", "monoslam_ransac.cpp")
OUT = os.path.join(HERE, "_ref")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
# (first line, last line, text that must appear on the first line, text that must appear on the last line)
HELPERS = (40, 60, "Vector3f quat2vec(Vector4f quat)", "}")
DIST = (585, 585, "float DistWalked = poses_diff(last_image_pose,stat14.segment<7>(0),last_vrot);", "DistWalked")
BODY = (609, 687, "if ( DistWalked>(MoveThresh/2)   &&   DistWalked<MoveThresh ) {", "}")


def _cut(lines, spec):
    a, b, first, last = spec
    chunk = lines[a - 1:b]
    if first not in chunk[0] or last not in chunk[-1]:
        raise RuntimeError(f"monoslam_ransac.cpp:{a}-{b} is not the expected text (reference revision changed?)")
    return "".join(chunk)


def build(force=False):
    if not os.path.exists(SRC):
        raise RuntimeError(f"reference source not found: {SRC}")
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, "libref_selector_f64.so")
    tmpl = os.path.join(HERE, "ref_selector_harness.cpp.in")
    deps = [SRC, tmpl, os.path.abspath(__file__)]
    for root, _, names in os.walk(os.path.join(HERE, "shim")):
        deps += [os.path.join(root, n) for n in names]
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(d) for d in deps):
        return out
    lines = open(SRC).read().splitlines(keepends=True)
    text = open(tmpl).read()
    text = text.replace("@@HELPERS@@", _cut(lines, HELPERS))
    text = text.replace("@@SELECTOR@@", _cut(lines, DIST) + _cut(lines, BODY))
    gen = os.path.join(OUT, "ref_selector_gen.cpp")
    with open(gen, "w") as f:
        f.write(text)
    cmd = [CXX, "-std=c++17", "-O2", "-msse4", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-DEKF_SHIM_DOUBLE",
           "-I", os.path.join(HERE, "shim"), "-I", OUT, "-o", out, gen]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"building libref_selector_f64.so failed:\n{r.stderr[-6000:]}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
