"""-m gpu: stateless batched NCC matcher (BASELINE config 5) vs the oracle's Patch::findMatch.
Match coordinates, accept/reject and the float NCC score must be bit-exact."""
import numpy as np
import pytest

from helpers import near_tie_scene

pytestmark = pytest.mark.gpu


def _run_gpu(pkg, d, F, M, W, H, w, sigma_size, clamp=20.0):
    import torch
    dev = torch.device("cuda:0")
    frames = torch.from_numpy(d["frames"]).to(dev)
    tm = torch.from_numpy(d["templates"]).to(dev)
    h = torch.from_numpy(d["h"]).to(dev)
    S = torch.from_numpy(d["S"]).to(dev)
    uv = torch.zeros((F * M, 2), dtype=torch.int32, device=dev)
    sc = torch.zeros(F * M, dtype=torch.float32, device=dev)
    pkg.match_batch(frames.data_ptr(), F, W, H, W, tm.data_ptr(), M, w, h.data_ptr(), S.data_ptr(), uv.data_ptr(),
                    sc.data_ptr(), sigma_size=sigma_size, ncc_threshold=0.8, search_clamp=clamp,
                    stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return uv.cpu().numpy(), sc.cpu().numpy()


@pytest.mark.parametrize("F,M,W,H,w,s_diag,sigma_size", [
    (2, 40, 640, 480, 11, 16.0, 3.0),     # delta = 12 px
    (2, 30, 640, 480, 11, 60.0, 3.0),     # clamp at +-20 px
    (1, 24, 320, 240, 21, 9.0, 2.0),      # reference default window (ConfigVSLAM.cpp:31)
    (1, 16, 320, 240, 30, 25.0, 3.0),     # even window (conf_sim.cfg)
])
def test_match_batch_bit_exact(gpu_pkg, orc, F, M, W, H, w, s_diag, sigma_size):
    d = gpu_pkg.synth.match_batch_inputs(n_frames=F, features_per_frame=M, width=W, height=H, window=w,
                                         seed=31 + w, s_diag=s_diag)
    uv_o, sc_o = orc.match_batch(d["frames"], d["templates"], d["h"], d["S"], sigma_size=sigma_size)
    uv_g, sc_g = _run_gpu(gpu_pkg, d, F, M, W, H, w, sigma_size)
    assert np.array_equal(uv_g, uv_o)
    assert np.array_equal(sc_g.view(np.uint32), sc_o.view(np.uint32)), "float NCC scores differ bitwise"
    found = (uv_o[:, 0] >= 0)
    assert found.mean() > 0.5
    assert np.array_equal(uv_o[found], d["truth"][found])


def test_match_batch_edges(gpu_pkg, orc):
    """Windows clipped by the border, predictions outside the frame, flat templates (NaN score),
    near-singular and huge covariances."""
    rng = np.random.default_rng(3)
    W, H, w, M = 160, 120, 11, 12
    d = gpu_pkg.synth.match_batch_inputs(n_frames=1, features_per_frame=M, width=W, height=H, window=w, seed=77)
    h = d["h"].copy(); S = d["S"].copy(); tm = d["templates"].copy()
    h[0] = (3.2, 4.9); h[1] = (W - 2.5, H - 1.5); h[2] = (-30.0, 50.0); h[3] = (W + 40.0, 10.0)
    tm[4] = 128                      # flat template -> 0/0
    S[5] = (1e-8, 0, 0, 1e-8)        # tiny ellipse: only the centre pixel
    S[6] = (400.0, 390.0, 390.0, 400.0)  # elongated, clamped
    S[7] = (16.0, -15.9, -15.9, 16.0)
    d2 = dict(frames=d["frames"], templates=tm, h=h, S=S)
    uv_o, sc_o = orc.match_batch(d2["frames"], tm, h, S, sigma_size=3.0)
    uv_g, sc_g = _run_gpu(gpu_pkg, d2, 1, M, W, H, w, 3.0)
    assert np.array_equal(uv_g, uv_o)
    assert np.array_equal(sc_g.view(np.uint32), sc_o.view(np.uint32))


def test_match_batch_near_ties_bit_exact(gpu_pkg, orc):
    """Smooth, periodic, coarse and saturated frames: the full-window tile matcher (template side 11) decides what it can and
    marks exact ties / crowded bands for the CTA matcher; both together must give the reference's answer on every feature."""
    frames, templates, h, S, F, M = near_tie_scene()
    uv_o, sc_o = orc.match_batch(frames, templates, h, S, sigma_size=3.0)
    d = dict(frames=frames, templates=templates, h=h, S=S)
    uv_g, sc_g = _run_gpu(gpu_pkg, d, F, M, frames.shape[2], frames.shape[1], 11, 3.0)
    assert np.array_equal(uv_g, uv_o)
    assert np.array_equal(sc_g.view(np.uint32), sc_o.view(np.uint32)), "float NCC scores differ bitwise"
    assert (uv_o[:M, 0] >= 0).mean() > 0.9


def test_match_batch_unaligned_frames(gpu_pkg, orc):
    """Frame width 323 (row stride not a multiple of 4): the tile matcher stages the window with byte loads."""
    d = gpu_pkg.synth.match_batch_inputs(n_frames=2, features_per_frame=20, width=323, height=241, window=11, seed=91, s_diag=40.0)
    uv_o, sc_o = orc.match_batch(d["frames"], d["templates"], d["h"], d["S"], sigma_size=3.0)
    uv_g, sc_g = _run_gpu(gpu_pkg, d, 2, 20, 323, 241, 11, 3.0)
    assert np.array_equal(uv_g, uv_o)
    assert np.array_equal(sc_g.view(np.uint32), sc_o.view(np.uint32))
