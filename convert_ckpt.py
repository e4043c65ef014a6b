"""Add the 5 inpainting input channels to the first conv of a GLIGEN checkpoint (the reference's convert_ckpt.py:21-43 CLI).

    python convert_ckpt.py --ckpt_path gligen.pth --new_ckpt_path gligen_inpaint.pth
"""
import argparse

import torch

from gligen_b200.checkpoint import add_additional_channels  # noqa: F401  (same name as the reference's helper)

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--ckpt_path", type=str, default=None)
    parser.add_argument("--new_ckpt_path", type=str, default=None)
    args = parser.parse_args()
    ckpt = torch.load(args.ckpt_path, map_location="cpu")
    add_additional_channels(ckpt["model"], 4 + 1)
    torch.save({"model": ckpt["model"]}, args.new_ckpt_path)
