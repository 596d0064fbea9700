"""Host-side logic of the N>1 paths on CPU: world_size 2, gloo, 127.0.0.1 rendezvous.
The all-gather must equal torch.stack of the per-rank payloads (bit-exact), the timing reduction
must be the max over ranks, and the stream->rank / frame-seed sharding must be disjoint."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from autoware_vision_pilot_b200 import multicam
from oracle import post


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_rank_data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    feat = torch.randn(multicam.FEAT_SHAPE, generator=g).to(torch.float16)
    rng = np.random.default_rng(rank)
    lc = [1e-3 * rng.normal(), 0.02 * rng.normal(), -1.8]
    rc = [1e-3 * rng.normal(), 0.02 * rng.normal(), 1.9] if rank != 1 else [float("nan")] * 3
    meas = torch.from_numpy(post.pathfinder_measurement(lc, rc, 0.01 * rank, 4.0))
    return feat, meas


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        feat, meas = _make_rank_data(rank)
        feats, measurements = multicam.all_gather_cameras(feat, meas)
        mx = multicam.max_over_ranks(10.0 + rank, torch.device("cpu"))
        torch.save({"feats": feats, "meas": measurements, "max": mx}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_all_gather_and_max_reduce_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    exp_f = torch.stack([_make_rank_data(r)[0] for r in range(world)])
    exp_m = torch.stack([_make_rank_data(r)[1] for r in range(world)])
    for r in range(world):
        d = torch.load(tmp_path / f"r{r}.pt")
        assert torch.equal(d["feats"], exp_f)                              # bit-exact
        assert torch.equal(torch.nan_to_num(d["meas"], nan=-7.0), torch.nan_to_num(exp_m, nan=-7.0))
        assert d["max"] == 11.0
    # fusing the gathered measurements with the oracle rule gives every rank the same state
    st = post.initial_state()
    for m in exp_m.numpy():
        st = post.estimator_update(st, m)
    assert np.isfinite(st[3]).all() and st[3, 1] < 1e3


def test_payload_roundtrip_and_sharding():
    feat, meas = _make_rank_data(0)
    f2, m2 = multicam.unpack_payload(multicam.pack_payload(feat, meas), torch.float16)
    assert torch.equal(f2, feat) and torch.equal(torch.nan_to_num(m2), torch.nan_to_num(meas))
    seeds = {multicam.frame_seed(r, f) for r in range(8) for f in range(100)}
    assert len(seeds) == 800                                               # disjoint streams
    assert multicam.max_over_ranks(3.5, torch.device("cpu")) == 3.5        # single process: identity


def test_fusion_refuses_cpu_tensors():
    with pytest.raises(RuntimeError):
        multicam.fuse_measurements(torch.zeros(14, 2, dtype=torch.float64), torch.zeros(2, 14, 2, dtype=torch.float64))
