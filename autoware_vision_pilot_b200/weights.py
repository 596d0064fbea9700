"""Checkpoint conversion: the reference's `.pth` state_dict -> `.vpw` (the flat file the C++ engine
reads; csrc/engine.cu load_vpw).

The reference loads `torch.load(path, weights_only=True)` into the nn.Module
(Models/inference/scene_seg_infer.py:30-31); its C++ side never reads a .pth either — it consumes a
converted artefact (ONNX, Models/exports/convert_pytorch_to_onnx.py).  The `.vpw` plays that role
here: same tensor names and shapes as the state_dict, raw little-endian
fp32, no pickle.  BN folding and K-major 16-bit repacking happen inside the engine at load.
"""
from __future__ import annotations

import os
import struct
from typing import Dict

import numpy as np


def write_vpw(state_dict: Dict[str, "np.ndarray"], path: str) -> str:
    """state_dict values may be torch tensors or numpy arrays."""
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(b"VPW1")
        f.write(struct.pack("<I", len(state_dict)))
        for name, t in state_dict.items():
            a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
            if a.dtype == np.int64:
                dt = 1
            else:
                dt, a = 0, a.astype(np.float32, copy=False)
            a = np.ascontiguousarray(a)
            nb = name.encode("utf-8")
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<II", dt, a.ndim))
            for d in a.shape:
                f.write(struct.pack("<I", d))
            f.write(struct.pack("<Q", a.nbytes))
            f.write(a.tobytes())
    os.replace(tmp, path)
    return path


def convert_checkpoint(pth_path: str, vpw_path: str | None = None) -> str:
    """.pth (torch.save(model.state_dict())) -> .vpw next to it (cached by mtime)."""
    import torch
    if vpw_path is None:
        vpw_path = os.path.splitext(pth_path)[0] + ".vpw"
    if os.path.exists(vpw_path) and os.path.getmtime(vpw_path) >= os.path.getmtime(pth_path):
        return vpw_path
    sd = torch.load(pth_path, weights_only=True, map_location="cpu")
    return write_vpw(sd, vpw_path)
