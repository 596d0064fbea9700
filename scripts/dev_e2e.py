"""GPU-box development check: engine vs oracle per tap, for one model.  python scripts/dev_e2e.py scene_seg"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from autoware_vision_pilot_b200 import engine as E  # noqa: E402
from autoware_vision_pilot_b200 import weights as W  # noqa: E402
from oracle import net, resize, synth  # noqa: E402


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "scene_seg"
    dtype = sys.argv[2] if len(sys.argv) > 2 else "fp16"
    sd = synth.synth_state_dict(model)
    path = f"/tmp/{model}.vpw"
    W.write_vpw(sd, path)
    frame = synth.synth_frame(0)
    small = resize.pil_bicubic_resize(frame, 640, 320)
    t = time.time()
    eng = E.Engine([E.KIND_BY_NAME[model]], [path], dtype=dtype, resize_mode=E.RESIZE_PIL_BICUBIC)
    print(f"engine create {time.time() - t:.2f}s", eng.stats())
    eng.infer(frame)
    rs = eng.read_resized()
    print("resized bit-exact:", np.array_equal(rs, small), int((rs != small).sum()))
    taps = {}
    x = net.to_tensor_normalize(small)
    out = net.forward(model, sd, x, taps=taps)
    pre = eng.read_tap("pre")
    d = np.abs(pre - x[0].numpy())
    print(f"pre: max abs err {d.max():.3e}  (fp16 ulp at 2.6 = {2.0 ** -9:.3e})")
    names = ["f0", "f1", "f2", "f3", "f4", "context", "neck"] + (["fused"] if model == "ego_lanes" else [])
    for k in names:
        got = eng.read_tap("0/" + k)
        ref = taps[k][0].numpy()
        err = np.abs(got - ref)
        print(f"{k:8s} shape {got.shape}  ref std {ref.std():.3f}  max|d|/std {err.max() / ref.std():.4f}  mean|d|/std {err.mean() / ref.std():.5f}")
    raw = eng.raw(0).copy()
    ref = out[0].numpy()
    err = np.abs(raw - ref)
    sig = ref.std()
    print(f"logits: max|d| {err.max():.4f} ({err.max() / sig:.4f} sigma)  mean|d| {err.mean():.5f} ({err.mean() / sig:.5f} sigma)")
    if model == "scene_seg":
        cls = eng.cls(0)
        exp = ref.argmax(0)
        srt = np.sort(ref, axis=0)
        margin = srt[-1] - srt[-2]
        bad = cls != exp
        print(f"argmax mismatch {bad.mean() * 100:.3f}%  largest margin flipped {margin[bad].max() if bad.any() else 0:.4f}  tau=2*max|d|={2 * err.max():.4f}")
    # timing: device-resident frame, graph replay
    dframe = torch.from_numpy(frame).cuda()
    for _ in range(5):
        eng.infer_device(dframe.data_ptr(), 1080, 1920, 1920 * 3)
    eng.sync()
    n = 50
    t = time.time()
    for _ in range(n):
        eng.infer_device(dframe.data_ptr(), 1080, 1920, 1920 * 3)
    eng.sync()
    dt = (time.time() - t) / n
    st = eng.stats()
    print(f"device-resident: {dt * 1e3:.3f} ms/frame  {1 / dt:.1f} FPS  {st['total_flops'] / dt / 1e12:.1f} TFLOP/s")
    t = time.time()
    for _ in range(20):
        eng.infer(frame)
    dt2 = (time.time() - t) / 20
    print(f"host e2e (pageable in, sync): {dt2 * 1e3:.3f} ms/frame")
    prof = eng.profile()
    tot = sum(p["ms"] for p in prof)
    print(f"eager per-op total {tot:.3f} ms over {len(prof)} launches")
    for p in sorted(prof, key=lambda p: -p["ms"])[:25]:
        tf = p["flops"] / p["ms"] / 1e9 if p["ms"] > 0 else 0
        print(f"  {p['name']:22s} {p['ms'] * 1e3:8.1f} us  {tf:7.1f} TF/s")
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/profile_{model}.txt", "w") as f:
        for p in prof:
            tf = p["flops"] / p["ms"] / 1e9 if p["ms"] > 0 else 0
            f.write(f"{p['name']:24s} {p['ms'] * 1e3:9.1f} us {tf:8.1f} TF/s gemm={int(p['gemm'])}\n")
    enc = sum(p["ms"] for p in prof if "mb" in p["name"] or "stem" in p["name"] or "enc8" in p["name"])
    print(f"encoder total {enc * 1e3:.1f} us; preprocess {prof[0]['ms'] * 1e3:.1f} us")


if __name__ == "__main__":
    main()
