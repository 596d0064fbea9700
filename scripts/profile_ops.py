"""Per-op device times of the 4-task engine (eager, one CUDA-event pair per launch; vp_engine_profile), sorted by lane
and by time — run on the GPU box: python scripts/profile_ops.py [n_runs]"""
import os
import sys
import tempfile

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

from autoware_vision_pilot_b200 import engine as E  # noqa: E402
from autoware_vision_pilot_b200 import weights as W  # noqa: E402
from oracle import synth  # noqa: E402

MODELS = ("scene_seg", "scene_3d", "domain_seg", "ego_lanes")


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    tmp = tempfile.mkdtemp()
    paths = [W.write_vpw(synth.synth_state_dict(m), os.path.join(tmp, m + ".vpw")) for m in MODELS]
    eng = E.Engine([E.KIND_BY_NAME[m] for m in MODELS], paths, resize_mode=E.RESIZE_PIL_BICUBIC, fetch_raw=False)
    eng.infer(synth.synth_frame(0))
    acc = None
    for _ in range(runs):
        prof = eng.profile()
        if acc is None:
            acc = [dict(p) for p in prof]
        else:
            for a, p in zip(acc, prof):
                a["ms"] = min(a["ms"], p["ms"])
    tot = sum(a["ms"] for a in acc)
    print(f"{len(acc)} ops, serial total {tot * 1e3:.0f} us")
    for a in acc:
        tf = a["flops"] / a["ms"] / 1e9 if a["ms"] > 0 and a["flops"] else 0
        print(f"{a['name']:28s} {a['ms'] * 1e3:8.1f} us  {a['flops'] / 1e9:8.2f} GF {tf:8.1f} TF/s  {a['kernel'] or ''}")


if __name__ == "__main__":
    main()
