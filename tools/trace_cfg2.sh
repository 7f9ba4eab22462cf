#!/bin/bash
# tools/trace_cfg2.sh — per-launch timeline of one cfg2 update (EKF_TRACE), last traced step of a short run
EKF_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sharded --no-parity 2>&1 | grep "^trace" > /tmp/trace_all.txt
# the 4th traced update (warm)
awk '/^trace end/{n++; next} n==4' /tmp/trace_all.txt
