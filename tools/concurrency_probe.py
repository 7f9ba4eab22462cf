import sys, time, threading
sys.path[:0]=['/root/repo','/root/repo/tests']
import numpy as np, torch
import ekfb200
pkg=ekfb200.load_package()
import bench
sc=bench.make_scene(pkg,"cfg2_n500",14)
cfg=pkg.default_config(**sc.config_overrides())
def mk():
    f=pkg.VSlamFilter(cfg,feature_capacity=504); f.set_symmetric_downdate(True); bench.seed_filter(f,sc); return f
frames=[sc.frame(t) for t in range(14)]
picks=[sc.picks(t,500) for t in range(14)]
def run(f, ts, out):
    t0=time.perf_counter()
    for t in ts:
        f.captureNewFrame(frames[t], sc.stamps[t]); f.predict(); f.update(picks[t])
    f.sync(); out.append(time.perf_counter()-t0)
f1=mk(); f2=mk()
for f in (f1,f2): run(f, range(1,4), [])
o=[]; run(f1, range(4,14), o); print("one filter, 10 steps: %.2f ms/step"%(o[0]*100))
f1b=mk(); f2b=mk()
for f in (f1b,f2b): run(f, range(1,4), [])
o1,o2=[],[]
th=[threading.Thread(target=run,args=(f1b,range(4,14),o1)),threading.Thread(target=run,args=(f2b,range(4,14),o2))]
t0=time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; wall=time.perf_counter()-t0
print("two filters concurrently: wall %.2f ms per step-pair (%.2f, %.2f)"%(wall*100,o1[0]*100,o2[0]*100))
