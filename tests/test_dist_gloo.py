"""CPU, world_size 2 over gloo: the host logic of the one sharded path (batch of filters split by
filter index, no data-path collective) — slices, deterministic per-filter perturbations, the
max-over-ranks timing rule and the reporting all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import ekfb200
    pkg = ekfb200.load_package()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 6
    sl = pkg.dist.ensemble_slice(rank, world, B)
    mu0 = np.array([0, 0, 0, 0, 0, -0.707106781, 0.707106781, 0, 0, 0, 0, 0, 0, 1.0])
    cams = pkg.dist.ensemble_camera_states(mu0, sl)
    t = pkg.dist.max_over_ranks(10.0 + 5.0 * rank, dist)
    allc = pkg.dist.gather_camera_states(cams, dist)
    dist.barrier()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), sl=np.array(list(sl)), cams=cams, t=t, allc=allc)
    dist.destroy_process_group()


def test_ensemble_sharding_world2(tmp_path, pkg):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"r{k}.npz") for k in range(world)]
    # slices are disjoint and cover the ensemble in order
    assert list(r[0]["sl"]) == list(range(0, 6)) and list(r[1]["sl"]) == list(range(6, 12))
    # timing rule: max over ranks, identical on every rank
    assert r[0]["t"] == r[1]["t"] == 15.0
    # reporting all-gather: global filter order, same on every rank
    assert np.array_equal(r[0]["allc"], r[1]["allc"])
    assert np.array_equal(r[0]["allc"], np.concatenate([r[0]["cams"], r[1]["cams"]]))
    # a filter's perturbation depends on its GLOBAL index only (re-sharding does not change the ensemble)
    mu0 = r[0]["cams"][0] * 0 + np.array([0, 0, 0, 0, 0, -0.707106781, 0.707106781, 0, 0, 0, 0, 0, 0, 1.0])
    single = pkg.dist.ensemble_camera_states(mu0, range(12))
    assert np.array_equal(single, r[0]["allc"])
    assert np.allclose(np.linalg.norm(single[:, 3:7], axis=1), 1.0)
    assert len({tuple(row) for row in single}) == 12


def test_single_rank_paths(pkg):
    assert pkg.dist.max_over_ranks(3.5) == 3.5
    assert pkg.dist.gather_camera_states(np.ones((2, 14))).shape == (2, 14)
    with pytest.raises(ValueError):
        pkg.dist.ensemble_slice(2, 2, 4)
