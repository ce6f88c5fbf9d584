"""Drop-in for ``flyingChairsLoader.py`` with the decode on the device.

The reference reads every image with a single-threaded ``cv2.imread`` + ``cv2.resize`` and every flow with ``np.fromfile``
(flyingChairsLoader.py:64-82,88-104).  Here the host only moves BYTES: the .ppm / .flo files of a batch are read by a small thread pool
straight into one pinned buffer, copied to the device in one transfer, and decoded there (dofb_decode_ppm: RGB -> BGR float + OpenCV's
fixed-point bilinear resize; dofb_decode_flo: header check + payload).  Same class / method names and return order as the reference;
the arrays returned are CUDA tensors ([B,H,W,3] float32 BGR 0..255, [B,h,w,2] float32), which TrainStep.run takes without another copy."""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import ops
from ._lib import DeepOFError

SPLIT_FILE = "FlyingChairs_train_val.txt"


def _ppm_header(buf: memoryview):
    """-> (width, height, maxval, offset of the first pixel byte) of a binary P6 header (comments allowed)."""
    pos, fields = 0, []
    n = len(buf)
    while len(fields) < 4:
        while pos < n and buf[pos] in b" \t\r\n":
            pos += 1
        if pos < n and buf[pos] == ord("#"):
            while pos < n and buf[pos] not in b"\r\n":
                pos += 1
            continue
        start = pos
        while pos < n and buf[pos] not in b" \t\r\n":
            pos += 1
        if start == pos:
            raise DeepOFError("truncated .ppm header")
        fields.append(bytes(buf[start:pos]))
    if fields[0] != b"P6":
        raise DeepOFError(f"not a binary P6 .ppm (magic {fields[0]!r})")
    w, h, maxval = int(fields[1]), int(fields[2]), int(fields[3])
    if maxval != 255:
        raise DeepOFError(f".ppm with maxval {maxval}: only 8-bit images are supported (FlyingChairs is 8-bit)")
    return w, h, maxval, pos + 1                      # exactly one whitespace byte separates the header from the pixels


class flyingChairsLoader:
    """Pipeline for preparing the Flying Chairs data (flyingChairsLoader.py:7-113): 512 x 384 images, 22872 pairs, standard split."""

    def __init__(self, data_path, image_size, device="cuda", split_file=None, io_threads: int = 16):
        self.data_path = data_path
        self.image_size = image_size
        self.img_path = os.path.join(self.data_path, "data")
        self.device = torch.device(device)
        self._pool = ThreadPoolExecutor(max_workers=io_threads)
        self._pin = None
        self.trainValGT = self.trainValSplit(split_file)
        self.trainList, self.valList = self.getData(self.img_path)
        print("We have %d training samples and %d validation samples." % (len(self.trainList), len(self.valList)))

    def trainValSplit(self, split_file=None):
        """The standard split file (:30-39); looked up in the working directory (as the reference does) or next to the data.
        There is no network here: a missing file is an error instead of a wget."""
        for cand in (split_file, SPLIT_FILE, os.path.join(self.data_path, SPLIT_FILE)):
            if cand and os.path.exists(cand):
                with open(cand, "r") as f:
                    return f.readlines()
        raise DeepOFError(f"{SPLIT_FILE} not found (looked in the working directory and in {self.data_path})")

    def getData(self, img_path):
        assert os.path.exists(img_path)
        train, val = [], []
        for imgIdx in range(len(self.trainValGT)):
            frameID = "%05d" % (imgIdx + 1)             # the image index starts at 00001 (:49)
            tag = self.trainValGT[imgIdx][0]
            if tag == "1":
                train.append(frameID)
            elif tag == "2":
                val.append(frameID)
            else:
                print("Something wrong with the split file.")
        return train, val

    # ------------------------------------------------------------------ device decode
    def _read_all(self, paths):
        """Read the files into ONE pinned uint8 buffer (thread pool) -> (device byte tensor, per-file offsets, per-file sizes)."""
        sizes = [os.path.getsize(p) for p in paths]
        offs = np.concatenate([[0], np.cumsum([(s + 15) // 16 * 16 for s in sizes])]).astype(np.int64)
        total = int(offs[-1])
        if self._pin is None or self._pin.numel() < total:
            self._pin = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
        host = self._pin.numpy()

        def load(i):
            with open(paths[i], "rb", buffering=0) as f:
                got = f.readinto(memoryview(host)[int(offs[i]):int(offs[i]) + sizes[i]])
            if got != sizes[i]:
                raise DeepOFError(f"short read on {paths[i]}")
        list(self._pool.map(load, range(len(paths))))
        dev = self._pin[:total].to(self.device, non_blocking=True)
        return dev, offs[:-1], sizes, host

    def _hook(self, frame_ids):
        assert len(frame_ids) > 0, "we need a non-empty batch list"
        paths = []
        for fid in frame_ids:
            paths += [os.path.join(self.img_path, fid + "_img1.ppm"), os.path.join(self.img_path, fid + "_img2.ppm"),
                      os.path.join(self.img_path, fid + "_flow.flo")]
        dev, offs, sizes, host = self._read_all(paths)
        n = len(frame_ids)
        data_off = np.empty(2 * n, dtype=np.int64)
        wh = None
        for j in range(n):
            for k in range(2):
                i = 3 * j + k
                w, h, _mv, hdr = _ppm_header(memoryview(host)[int(offs[i]):int(offs[i]) + min(sizes[i], 256)])
                if sizes[i] < hdr + 3 * w * h:
                    raise DeepOFError(f"{paths[i]}: truncated pixel data")
                if wh is None:
                    wh = (w, h)
                elif wh != (w, h):
                    raise DeepOFError(f"{paths[i]}: image size {w}x{h} differs from {wh[0]}x{wh[1]} inside one batch")
                data_off[k * n + j] = offs[i] + hdr           # all sources first, then all targets
        imgs = ops.decode_ppm(dev, torch.from_numpy(data_off).to(self.device), (wh[1], wh[0]), self.image_size)
        flo_off = torch.from_numpy(np.ascontiguousarray(offs[2::3])).to(self.device)
        fh = int(np.frombuffer(host[int(offs[2]) + 8:int(offs[2]) + 12].tobytes(), np.int32)[0])
        fw = int(np.frombuffer(host[int(offs[2]) + 4:int(offs[2]) + 8].tobytes(), np.int32)[0])
        flow, status = ops.decode_flo(dev, flo_off, (fh, fw))
        if int(status.item()) != 0:
            raise DeepOFError("Magic number incorrect. Invalid .flo file")       # utils.py:13
        return imgs[:n], imgs[n:], flow

    # ------------------------------------------------------------------ the reference's sampling interface
    def sampleTrain(self, batch_size, batch_id):
        assert batch_size > 0, "we need a batch size larger than 0"
        idxs = range((batch_id - 1) * batch_size, batch_id * batch_size)           # :61
        return self.hookTrainData(idxs)

    def hookTrainData(self, sampleIdxs):
        return self._hook([self.trainList[i] for i in sampleIdxs])

    def sampleVal(self, batch_size, batch_id):
        assert batch_size > 0, "we need a batch size larger than 0"
        idxs = range((batch_id - 1) * batch_size, batch_id * batch_size)
        return (self.hookValData(idxs), idxs)

    def hookValData(self, sampleIdxs):
        return self._hook([self.valList[i] for i in sampleIdxs])


def evaluate_aee(flows_scale1, flow_gt, mult: float = 2.0, clip=(-300.0, 250.0)) -> float:
    """The AEE of evaluateNet (flyingChairsTrain.py:264-266,294-296): flows_all[0] * 2, clip to the data set's range, bilinear resize to the
    ground-truth size, mean end-point error -- one device pass (dofb_eval_flow_aee_sum)."""
    return float(ops.eval_flow_aee(flows_scale1.contiguous(), flow_gt.contiguous(), mult, clip).item())
