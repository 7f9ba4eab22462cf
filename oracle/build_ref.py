#!/usr/bin/env python
"""oracle/build_ref.py — TEST INFRASTRUCTURE ONLY.

Compiles the reference's own hot-path sources, unmodified and from where they lie under
/root/reference/mono-slam/src, against the API stand-ins in oracle/shim/ plus the C harness
oracle/ref_capi.cpp, into oracle/_ref/libref_f32.so (as written, fp32) and oracle/_ref/libref_f64.so
(-DEKF_SHIM_DOUBLE: `float` re-typed to double).  Nothing is copied out of /root/reference; the
outputs are git-ignored.  Does not use the reference's build system (catkin/cmake + ROS).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("EKF_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "mono-slam", "src")
OUT = os.path.join(HERE, "_ref")
FILES = ["vslamRansac.cpp", "Patch.cpp", "camModel.cpp", "utils.cpp", "libblur.cpp", "RosVSLAMRansac.cpp"]
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
# -O2 -msse4 as mono-slam/CMakeLists.txt:3; asserts of the shim stay on (no -DNDEBUG); no FMA contraction
COMMON = ["-std=c++17", "-O2", "-msse4", "-fPIC", "-shared", "-ffp-contract=off", "-w",
          "-I", os.path.join(HERE, "shim"), "-I", SRC]


def build(force=False):
    if not os.path.isdir(SRC):
        raise RuntimeError(f"reference sources not found under {SRC}")
    os.makedirs(OUT, exist_ok=True)
    deps = [os.path.join(SRC, f) for f in FILES] + [os.path.join(HERE, "ref_capi.cpp"), os.path.abspath(__file__)]
    for root, _, names in os.walk(os.path.join(HERE, "shim")):
        deps += [os.path.join(root, n) for n in names]
    for root, _, names in os.walk(SRC):
        deps += [os.path.join(root, n) for n in names if n.endswith((".hpp", ".h"))]
    newest = max(os.path.getmtime(d) for d in deps)
    built = []
    for name, defs in (("libref_f32.so", []), ("libref_f64.so", ["-DEKF_SHIM_DOUBLE"])):
        out = os.path.join(OUT, name)
        if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
            built.append(out)
            continue
        cmd = [CXX] + COMMON + defs + ["-o", out, os.path.join(HERE, "ref_capi.cpp")] + [os.path.join(SRC, f) for f in FILES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"building {name} failed:\n{r.stderr[-6000:]}")
        built.append(out)
    return built


if __name__ == "__main__":
    for p in build(force="--force" in sys.argv):
        print(p)
