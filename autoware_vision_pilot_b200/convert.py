"""CLI: convert a reference checkpoint (.pth state_dict) to the engine's .vpw weight file.

    python -m autoware_vision_pilot_b200.convert scene_seg.pth [out.vpw]

Plays the role Models/exports/convert_pytorch_to_onnx.py plays for the reference's C++ backends.
"""
import sys

from .weights import convert_checkpoint


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 2
    out = convert_checkpoint(argv[0], argv[1] if len(argv) > 1 else None)
    print(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
