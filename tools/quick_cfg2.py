"""tools/quick_cfg2.py — a short cfg2 timing (resident pass + per-class events) without the baselines, for A/B runs on the GPU box.
Usage: python tools/quick_cfg2.py [steps]   (environment switches such as EKF_SCHED are read by the library)"""
import json, subprocess, sys
steps = sys.argv[1] if len(sys.argv) > 1 else "20"
out = subprocess.run([sys.executable, "bench.py", "--steps", steps, "--warmup", "3", "--no-cpu-baseline", "--no-sharded", "--no-parity"],
                     capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not line:
    print(out.stdout[-2000:], out.stderr[-2000:]); sys.exit(1)
d = json.loads(line[-1])
print("cfg2", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("kernel_ms_per_step"))
