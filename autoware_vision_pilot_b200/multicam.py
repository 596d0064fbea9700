"""Multi-GPU plumbing of the hot path.

* Configs 3/4: camera frames are independent units — stream k runs on rank k, one engine per GPU,
  no data-path collective.  `max_over_ranks` is the only cross-rank operation of the benchmark
  (device time, max over ranks).
* Config 5 (extension; the reference's PathFinder is single-camera, path_finder.cpp:48): every rank
  contributes its EgoLanes fused feature map [1456,10,20] 16-bit (BackboneFeatureFusion output,
  backbone_feature_fusion.py:37) and its 14-slot PathFinder measurement (mean, variance) fp64
  (path_finder.cpp:97-157); ONE all-gather over NVLink, then every rank applies the reference's own
  Gaussian-product / inverse-variance rule (estimator.cpp:24-74) to the gathered measurements.

torch.distributed is plumbing only: NCCL on the GPUs, gloo in the CPU tests of this logic.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch
import torch.distributed as dist

from . import _lib as L

STATE_DIM = 14
FEAT_SHAPE = (10, 20, 1456)          # NHWC fused feature map (582 400 B at 16 bit)


def stream_for_rank(rank: int) -> int:
    """Camera stream handled by `rank` (one stream per GPU)."""
    return rank


def frame_seed(rank: int, frame: int) -> int:
    """Synthetic frame seed of stream `rank`, frame `frame` (1000*k + f)."""
    return 1000 * stream_for_rank(rank) + frame


def max_over_ranks(value_ms: float, device: torch.device, group=None) -> float:
    """Multi-GPU numbers are the max over ranks of the device-timed duration."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value_ms)
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def pack_payload(features: torch.Tensor, measurement: torch.Tensor) -> torch.Tensor:
    """One contiguous byte buffer per rank: [features (16-bit) | measurement (14x2 fp64)]."""
    assert measurement.dtype == torch.float64 and measurement.numel() == STATE_DIM * 2
    assert features.element_size() == 2
    f = features.contiguous().view(torch.uint8).reshape(-1)
    m = measurement.contiguous().view(torch.uint8).reshape(-1)
    return torch.cat([f, m])


def unpack_payload(buf: torch.Tensor, feat_dtype: torch.dtype, feat_shape=FEAT_SHAPE) -> Tuple[torch.Tensor, torch.Tensor]:
    nfeat = 2
    for d in feat_shape:
        nfeat *= d
    f = buf[:nfeat].view(feat_dtype).reshape(feat_shape)
    m = buf[nfeat:nfeat + STATE_DIM * 2 * 8].view(torch.float64).reshape(STATE_DIM, 2)
    return f, m


def all_gather_cameras(features: torch.Tensor, measurement: torch.Tensor, group=None):
    """The single collective of config 5.  Returns (features [world, ...], measurements [world,14,2])."""
    world = dist.get_world_size(group)
    payload = pack_payload(features, measurement)
    gathered = torch.empty(world * payload.numel(), dtype=torch.uint8, device=payload.device)
    dist.all_gather_into_tensor(gathered, payload, group=group)   # ncclAllGather on the GPUs
    gathered = gathered.view(world, payload.numel())
    feats, meas = [], []
    for r in range(world):
        f, m = unpack_payload(gathered[r], features.dtype, tuple(features.shape))
        feats.append(f)
        meas.append(m)
    return torch.stack(feats), torch.stack(meas)


def fuse_measurements(state: torch.Tensor, measurements: torch.Tensor) -> torch.Tensor:
    """Estimator::update over the gathered per-camera measurements, on the GPU (vpb_bayes_fuse).
    state [14,2] fp64 CUDA (updated in place and returned), measurements [n,14,2] fp64 CUDA."""
    if not (state.is_cuda and measurements.is_cuda):
        raise RuntimeError("fuse_measurements runs on the GPU only (no CPU fallback)")
    lib = L.lib()
    lib.vpb_bayes_fuse.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    m = measurements.contiguous()
    L.check(lib.vpb_bayes_fuse(state.data_ptr(), m.data_ptr(), m.shape[0],
                               torch.cuda.current_stream().cuda_stream), "vpb_bayes_fuse")
    return state
