"""tools/match_ab.py — A/B of the warp tile matcher builds (EKF_MATCH_W2VAR 0 / 1 / 2) on the stateless batch matcher inside one
process: 64 frames x 200 features of 1920 x 1080, template side 11, windows at the 41 x 41 clamp.  Prints M matches/s per build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ekfb200
pkg = ekfb200.load_package(); pkg.build(); pkg.lib()
F, M, W, H, w = 64, 200, 1920, 1080, 11
d = pkg.synth.match_batch_inputs(n_frames=F, features_per_frame=M, width=W, height=H, window=w, seed=1239, s_diag=60.0)
dev = torch.device("cuda:0")
frames = torch.from_numpy(d["frames"]).to(dev); tm = torch.from_numpy(d["templates"]).to(dev)
h = torch.from_numpy(d["h"]).to(dev); S = torch.from_numpy(d["S"]).to(dev)
uv = torch.zeros((F * M, 2), dtype=torch.int32, device=dev); sc = torch.zeros(F * M, dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream
def launch():
    pkg.match_batch(frames.data_ptr(), F, W, H, W, tm.data_ptr(), M, w, h.data_ptr(), S.data_ptr(), uv.data_ptr(), sc.data_ptr(),
                    sigma_size=3.0, ncc_threshold=0.8, search_clamp=20.0, stream=st)
ref = None
for var in sys.argv[1:] or ["0", "1", "2"]:
    os.environ["EKF_MATCH_W2VAR"] = var
    for _ in range(3): launch()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): launch()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    res = (uv.cpu().numpy().copy(), sc.cpu().numpy().copy())
    same = True if ref is None else (np.array_equal(ref[0], res[0]) and np.array_equal(ref[1].view(np.uint32), res[1].view(np.uint32)))
    ref = ref or res
    print(f"W2VAR={var}: {ms:.4f} ms per launch, {F * M / ms / 1e3:.1f} M matches/s, found {(res[0][:, 0] >= 0).mean():.3f}, same bits as first: {same}")
