"""Per-tap error of the engine vs the fp32 oracle on a list of frames (GPU box): where does a frame's error enter?"""
import os
import sys
import tempfile

import numpy as np
from PIL import Image

sys.path.insert(0, ".")
from autoware_vision_pilot_b200 import engine as E  # noqa: E402
from autoware_vision_pilot_b200 import weights as W  # noqa: E402
from oracle import net, resize, synth  # noqa: E402


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "scene_seg"
    sd = synth.synth_state_dict(model)
    vpw = W.write_vpw(sd, os.path.join(tempfile.mkdtemp(), "m.vpw"))
    frames = {"synth0": synth.synth_frame(0), "s1_11": synth.synth_frame(synth.stream_seed(1, 11)),
              "s2_12": synth.synth_frame(synth.stream_seed(2, 12))}
    real = np.asarray(Image.open("tests/golden/real/frame_00.png").convert("RGB"))
    for dtype in ("fp16", "fp32"):
        eng = E.Engine([E.KIND_BY_NAME[model]], [vpw], dtype=dtype, resize_mode=E.RESIZE_PIL_BICUBIC)
        eng_n = E.Engine([E.KIND_BY_NAME[model]], [vpw], dtype=dtype, resize_mode=E.RESIZE_NONE)
        for name, f in list(frames.items()) + [("real0", real)]:
            if name == "real0":
                small = f
                eng_n.infer(small)
                g = eng_n
            else:
                small = resize.pil_bicubic_resize(f, 640, 320)
                eng.infer(f)
                g = eng
            taps = {}
            ref = net.forward(model, sd, net.to_tensor_normalize(small), taps=taps)[0].numpy()
            row = []
            for k in ("f0", "f1", "f2", "f3", "f4", "context", "neck"):
                t = taps[k][0].numpy()
                e = np.abs(g.read_tap("0/" + k) - t)
                row.append(f"{k} {e.max() / t.std():.4f}/{e.mean() / t.std():.5f}")
            e = np.abs(g.raw(0) - ref)
            print(f"{dtype} {name:7s} out {e.max() / ref.std():.4f}/{e.mean() / ref.std():.5f} | " + " | ".join(row), flush=True)


if __name__ == "__main__":
    main()
