"""Drop-in for ``deepOF_fc.py`` (deepOF_fc.py:5-7): ``deepOF(data_path)`` trains the guided VGG16 flow model at 320x448."""
from __future__ import annotations

import argparse

from .flyingChairsTrain_vgg import train

IMAGE_SIZE = [320, 448]              # deepOF_fc.py:6


def deepOF(data_path, **kw):
    """deepOF(data_path): ``train(data_path, image_size)`` of flyingChairsTrain_vgg.py with image_size = [320, 448].
    Extra keyword arguments (batch_size, math_mode, max_iters, max_epochs, sample_fn, ...) go to the trainer."""
    display = kw.pop("display", 0)
    t = train(kw.pop("sample_fn", None), image_size=tuple(IMAGE_SIZE), data_path=data_path, **kw)
    t.trainNet(display=display)
    return t


def main(argv=None):
    parser = argparse.ArgumentParser(description="Unsupervised motion estimation from videos (B200-native training step)")
    parser.add_argument("data_path", type=str, help="Path to FlyingChairs_release/")
    parser.add_argument("--math", default="bf16", choices=["fp32", "tf32", "bf16"])
    parser.add_argument("--batch-size", type=int, default=8)
    parser.add_argument("--max-iters", type=int, default=100)
    args = parser.parse_args(argv)
    deepOF(args.data_path, math_mode=args.math, batch_size=args.batch_size, max_iters=args.max_iters, tc_wgrad=args.math != "fp32")


if __name__ == "__main__":
    main()
