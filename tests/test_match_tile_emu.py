"""CPU: the scoring core of the full-window warp matcher (csrc/ekf_match_tile.cuh — the same __host__ __device__ code the
kernels k_match_filter_batch_warp2 / k_match_batch_warp2 run) executed lane by lane on the host (tests/match_tile_emu.cu)
against the oracle's Patch::findMatch: match coordinates, accept / reject and the float NCC score bit for bit.

What this pins without a GPU: the tile index arithmetic (funnel shifts, last-word mask, running-sum differences), the exact
int32 numerators, and above all the pre-selection argument — a float score good to 6e-7 only has to keep every candidate of the
double-precision guard band inside a 4e-6 band; the decision itself is the reference's exact operation sequence."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import near_tie_scene

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "ekf-monoslam_for_3d-reconstruction_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = os.path.join(HERE, "_build", "libmatch_tile_emu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(HERE, "match_tile_emu.cu")
    deps = [src, os.path.join(CSRC, "ekf_match_tile.cuh"), os.path.join(CSRC, "ekf_math.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "--fmad=false", "-std=c++17", "-Xcompiler", "-fPIC",
                        "-diag-suppress", "550", "-shared", "-I", CSRC, "-I", os.path.join(ROOT, "include"), src, "-o", out], check=True)
    lib = ctypes.CDLL(out)
    lib.emu_match_batch.restype = ctypes.c_int
    return lib


def _emu_run(lib, frames, templates, h, S, sigma_size, thr=0.8, clamp=20.0, tile_rows=4):
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    F, H, W = frames.shape
    templates = np.ascontiguousarray(templates, dtype=np.uint8)
    total = templates.shape[0]
    M = total // F
    h = np.ascontiguousarray(h, dtype=np.float64); S = np.ascontiguousarray(S, dtype=np.float64)
    uv = np.full((total, 2), -7, dtype=np.int32); sc = np.zeros(total, dtype=np.float32)
    dec = np.zeros(total, dtype=np.int32); nl = np.zeros(total, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.emu_match_batch(p(frames), F, W, H, W, p(templates), M, p(h), p(S), ctypes.c_float(sigma_size), ctypes.c_float(thr),
                             ctypes.c_float(clamp), p(uv), p(sc), p(dec), p(nl), int(tile_rows))
    assert rc == 0
    return uv, sc, dec.astype(bool), nl


def _check(uv_e, sc_e, dec, uv_o, sc_o):
    assert np.array_equal(uv_e[dec], uv_o[dec])
    assert np.array_equal(sc_e[dec].view(np.uint32), sc_o[dec].view(np.uint32)), "float NCC scores differ bitwise"


@pytest.mark.parametrize("tile_rows", [4, 2, 40, 20])
@pytest.mark.parametrize("s_diag,sigma_size,seed", [(16.0, 3.0, 5), (60.0, 3.0, 6), (200.0, 3.0, 7), (2.0, 2.0, 8)])
def test_tile_core_bit_exact_on_planted_scenes(pkg, orc, emu, s_diag, sigma_size, seed, tile_rows):
    d = pkg.synth.match_batch_inputs(n_frames=2, features_per_frame=40, width=640, height=480, window=11, seed=seed, s_diag=s_diag)
    uv_o, sc_o = orc.match_batch(d["frames"], d["templates"], d["h"], d["S"], sigma_size=sigma_size)
    uv_e, sc_e, dec, nl = _emu_run(emu, d["frames"], d["templates"], d["h"], d["S"], sigma_size, tile_rows=tile_rows)
    assert dec.all(), "noise frames have no near-ties: nothing should be handed to the CTA matcher"
    assert nl.max() <= 2
    _check(uv_e, sc_e, dec, uv_o, sc_o)
    if s_diag >= 16.0:   # the 2-sigma window of the last case often misses the planted position
        assert (uv_o[:, 0] >= 0).mean() > 0.5


@pytest.mark.parametrize("tile_rows", [4, 2, 40, 20])
def test_tile_core_edges(pkg, orc, emu, tile_rows):
    """Windows clipped by the border, predictions outside the frame, flat templates, tiny / elongated ellipses."""
    W, H, w, M = 160, 120, 11, 12
    d = pkg.synth.match_batch_inputs(n_frames=1, features_per_frame=M, width=W, height=H, window=w, seed=77)
    h = d["h"].copy(); S = d["S"].copy(); tm = d["templates"].copy()
    h[0] = (3.2, 4.9); h[1] = (W - 2.5, H - 1.5); h[2] = (-30.0, 50.0); h[3] = (W + 40.0, 10.0)
    tm[4] = 128
    S[5] = (1e-8, 0, 0, 1e-8)
    S[6] = (400.0, 390.0, 390.0, 400.0)
    S[7] = (16.0, -15.9, -15.9, 16.0)
    uv_o, sc_o = orc.match_batch(d["frames"], tm, h, S, sigma_size=3.0)
    uv_e, sc_e, dec, nl = _emu_run(emu, d["frames"], tm, h, S, 3.0, tile_rows=tile_rows)
    assert dec.all()
    _check(uv_e, sc_e, dec, uv_o, sc_o)


@pytest.mark.parametrize("tile_rows", [4, 2, 40, 20])
def test_tile_core_smooth_and_repetitive_images(orc, emu, tile_rows):
    """Images made to produce near-ties: a smooth gradient (neighbouring candidates score almost equally), an exactly periodic
    texture (many candidates with IDENTICAL sums: the band list overflows or a lane sees three of them and the feature goes to
    the CTA matcher), saturated patches (flat windows), and coarse quantisation.  Whatever the warp matcher decides itself must
    be the reference's answer; what it hands over is counted."""
    frames, templates, h, S, F, M = near_tie_scene()
    uv_o, sc_o = orc.match_batch(frames, templates, h, S, sigma_size=3.0)
    uv_e, sc_e, dec, nl = _emu_run(emu, frames, templates, h, S, 3.0, tile_rows=tile_rows)
    _check(uv_e, sc_e, dec, uv_o, sc_o)
    per_frame = dec.reshape(F, M).mean(axis=1)
    assert per_frame[0] == 1.0 and per_frame[3] > 0.9, per_frame     # smooth / saturated: decided by the warp
    assert per_frame[1] < 1.0, per_frame                               # periodic: exact ties go to the CTA matcher
    assert (uv_o[:M, 0] >= 0).mean() > 0.9


def test_float_score_error_bound():
    """The pre-selection argument in numbers: over random integer sums the float score stays within 6e-7 of ncc* in double."""
    rng = np.random.default_rng(5)
    n = 121
    t = rng.integers(0, 256, size=(20000, n)); p = rng.integers(0, 256, size=(20000, n))
    # correlated pairs as well, so that scores near 1 are covered
    p[::2] = np.clip(t[::2] + rng.integers(-3, 4, size=(10000, n)), 0, 255)
    T = t.sum(1); TT = (t * t).sum(1); P = p.sum(1); PP = (p * p).sum(1); Stp = (t * p).sum(1)
    d1 = n * TT - T * T; d2 = n * PP - P * P; num = n * Stp - T * P
    ok = (d1 > 0) & (d2 > 0)
    v = num[ok] / np.sqrt(d1[ok].astype(np.float64)) / np.sqrt(d2[ok].astype(np.float64))
    rd1f = (1.0 / np.sqrt(d1[ok].astype(np.float64))).astype(np.float32)
    f = num[ok].astype(np.float32) * (rd1f * (np.float32(1.0) / np.sqrt(d2[ok].astype(np.float32))))
    assert np.abs(num).max() < 2 ** 31 and d2.max() < 2 ** 31
    assert np.abs(f.astype(np.float64) - v).max() < 6e-7


@pytest.mark.parametrize("tile_rows", [4, 2, 40])
def test_tile_core_fuzz(orc, emu, tile_rows):
    """Random frame sizes and kinds (noise, smooth products of sines, three grey levels, 3 x 3 blocks), predictions up to 10 px
    outside the frame, anisotropic covariances with |rho| up to 0.98 and variances from 0.01 to 150 px^2, exact and perturbed
    templates, sigma_size 2 or 3: whatever the tile matcher decides is the oracle's answer bit for bit."""
    rng = np.random.default_rng(2024)
    w, F, M = 11, 2, 60
    for trial in range(8):
        Wd, Ht = int(rng.integers(64, 400)), int(rng.integers(64, 300))
        kind = trial % 4
        if kind == 0:
            frames = rng.integers(0, 256, size=(F, Ht, Wd), dtype=np.uint8)
        elif kind == 1:
            yy, xx = np.mgrid[0:Ht, 0:Wd]
            frames = np.stack([(127 + 80 * np.sin(xx / rng.uniform(5, 40)) * np.cos(yy / rng.uniform(5, 40))
                                + rng.normal(scale=rng.uniform(0, 3), size=(Ht, Wd))).clip(0, 255) for _ in range(F)]).astype(np.uint8)
        elif kind == 2:
            frames = (rng.integers(0, 3, size=(F, Ht, Wd)) * 100).astype(np.uint8)
        else:
            frames = np.repeat(np.repeat(rng.integers(0, 256, size=(F, Ht // 3 + 1, Wd // 3 + 1), dtype=np.uint8), 3, 1), 3, 2)[:, :Ht, :Wd].copy()
        u = rng.integers(-10, Wd + 10, size=(F, M)); v = rng.integers(-10, Ht + 10, size=(F, M))
        tm = np.zeros((F, M, w, w), dtype=np.uint8)
        for f in range(F):
            for i in range(M):
                uu = int(np.clip(u[f, i], 6, Wd - 7)); vv = int(np.clip(v[f, i], 6, Ht - 7))
                tm[f, i] = frames[f, vv - 5:vv + 6, uu - 5:uu + 6]
                if rng.random() < 0.3:
                    tm[f, i] = np.clip(tm[f, i].astype(int) + rng.integers(-20, 21, size=(w, w)), 0, 255)
        h = np.stack([u, v], -1).astype(np.float64) + rng.normal(scale=3.0, size=(F, M, 2))
        a = rng.uniform(0.01, 150, size=(F, M)); b = rng.uniform(0.01, 150, size=(F, M)); rho = rng.uniform(-0.98, 0.98, size=(F, M))
        S = np.zeros((F, M, 2, 2)); S[..., 0, 0] = a; S[..., 1, 1] = b; S[..., 0, 1] = S[..., 1, 0] = rho * np.sqrt(a * b)
        tm = tm.reshape(F * M, w, w); h = h.reshape(F * M, 2); S = S.reshape(F * M, 4)
        ss = float(rng.choice([2.0, 3.0]))
        uv_o, sc_o = orc.match_batch(frames, tm, h, S, sigma_size=ss)
        uv_e, sc_e, dec, nl = _emu_run(emu, frames, tm, h, S, ss, tile_rows=tile_rows)
        _check(uv_e, sc_e, dec, uv_o, sc_o)
        assert dec.mean() > 0.9
