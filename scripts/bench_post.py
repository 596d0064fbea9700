"""Device timings of the widened output-side ops: the fused mask overlay (HBM-bound) and the
lateral post-process kernel (latency-bound).  Run on the GPU box: python scripts/bench_post.py
Inputs are rotated over 24 frame buffers (149 MB > L2) so every overlay launch streams from HBM."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from autoware_vision_pilot_b200 import _lib as L  # noqa: E402
from autoware_vision_pilot_b200.lateral import LateralPostProcess  # noqa: E402
from oracle import lateral as LT  # noqa: E402


def main():
    lib = L.lib()
    lib.vpb_visualize_mask.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_int, C.c_void_p]
    h, w, nbuf, iters = 1080, 1920, 24, 240
    frames = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
    outs = [torch.empty_like(f) for f in frames]
    mask = torch.randint(0, 3, (320, 640), dtype=torch.uint8, device="cuda")
    for i in range(nbuf):
        lib.vpb_visualize_mask(mask.data_ptr(), 320, 640, 2, frames[i].data_ptr(), h, w, 3 * w, outs[i].data_ptr(), 3 * w, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        k = i % nbuf
        lib.vpb_visualize_mask(mask.data_ptr(), 320, 640, 2, frames[k].data_ptr(), h, w, 3 * w, outs[k].data_ptr(), 3 * w, None)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    nbytes = 2 * 3 * h * w + 320 * 640
    peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {}
    res = {"visualize_mask": {"us": us, "bytes": nbytes, "GBps": nbytes / us / 1e3}}

    post = LateralPostProcess()
    masks = [torch.from_numpy(LT.synth_lane_masks(50 + i)).cuda() for i in range(8)]
    for m in masks:
        post.update_device(m.data_ptr())
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        post.update_device(masks[i % 8].data_ptr())
    e1.record()
    torch.cuda.synchronize()
    res["lateral_update"] = {"us": e0.elapsed_time(e1) / iters * 1e3, "bytes": 3 * 80 * 160 * 4 + 1024}
    res["peaks_file"] = peaks
    print(json.dumps(res))


if __name__ == "__main__":
    main()
