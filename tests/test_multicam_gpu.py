"""Config 5 on real GPUs: every rank runs EgoLanes on its own camera, one NCCL
all-gather exchanges the fused feature maps + PathFinder measurements, every rank fuses the
measurements with the reference's Estimator::update rule on the GPU.  Needs >= 2 GPUs (skipped on the
single-GPU test box; run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, vpw, out_dir):
    import torch.distributed as dist
    from autoware_vision_pilot_b200 import engine as E
    from autoware_vision_pilot_b200 import multicam
    from oracle import post, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        eng = E.Engine([E.EGO_LANES], [vpw], gpu_id=rank, resize_mode=E.RESIZE_PIL_BICUBIC)
        frame = synth.synth_frame(multicam.frame_seed(rank, 0))
        eng.infer(frame)
        fused = torch.from_numpy(eng.read_tap("0/fused")).permute(1, 2, 0).contiguous().half().cuda()   # [10,20,1456]
        rng = np.random.default_rng(rank)
        lc = [1e-3 * rng.normal(), 0.02 * rng.normal(), -1.8]
        rc = [1e-3 * rng.normal(), 0.02 * rng.normal(), 1.9]
        meas = torch.from_numpy(post.pathfinder_measurement(lc, rc, 0.01 * rank, 4.0)).cuda()
        feats, allm = multicam.all_gather_cameras(fused, meas)
        state = torch.from_numpy(post.initial_state()).cuda()
        multicam.fuse_measurements(state, allm)
        torch.cuda.synchronize()
        torch.save({"feat_local": fused.cpu(), "meas_local": meas.cpu(), "feats": feats.cpu(), "meas": allm.cpu(),
                    "state": state.cpu()}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_multicamera_allgather_and_fusion(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from autoware_vision_pilot_b200 import weights as W
    from oracle import post, synth
    world = min(torch.cuda.device_count(), 8)
    vpw = W.write_vpw(synth.synth_state_dict("ego_lanes"), str(tmp_path / "ego.vpw"))
    mp.spawn(_worker, args=(world, _free_port(), vpw, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    exp_f = torch.stack([o["feat_local"] for o in outs])
    exp_m = torch.stack([o["meas_local"] for o in outs])
    ref = post.initial_state()
    for m in exp_m.numpy():
        ref = post.estimator_update(ref, m)
    for o in outs:
        assert torch.equal(o["feats"], exp_f)                       # all-gather == stack of per-rank payloads
        assert torch.equal(torch.nan_to_num(o["meas"], nan=-7.0), torch.nan_to_num(exp_m, nan=-7.0))
        assert np.allclose(o["state"].numpy(), ref, rtol=1e-13, atol=0)
