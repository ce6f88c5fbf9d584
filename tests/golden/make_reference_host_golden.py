"""Reference-derived fixtures for the host-side pieces of the path that the reference holds as plain NumPy / cv2 code.

Nothing is copied: the script READS the cited lines from the reference checkout at run time, makes them Python-3 runnable (``xrange`` ->
``range``, py2 ``print x`` statements dropped), executes them on synthetic inputs and stores inputs + outputs in
``tests/golden/reference_host.npz``:

  * ``utils.readFlow``              /root/reference/utils.py:4-21        (.flo parser)
  * ``utils.flow_ee``               /root/reference/utils.py:64-68       (average end-point error)
  * ``train.load_deconv_weights``   /root/reference/flyingChairsTrain.py:78-92   (bilinear initialisation of every 'up*' filter; the TF
                                    session is replaced by a stand-in that hands the shape in and takes the assigned array out)
  * evaluation recipe               /root/reference/flyingChairsTrain.py:263-267 + :294-296  (x2, clip, cv2.resize, AEE)
  * ``flyingChairsLoader.hookTrainData``  /root/reference/flyingChairsLoader.py:64-82  (cv2.imread + cv2.resize + readFlow on a tiny
                                    synthetic data-set directory written by this script)

    python tests/golden/make_reference_host_golden.py        # writes the fixture (needs /root/reference, cv2)

``tests/test_oracle_reference_pin.py`` checks the host modules / the oracle metric against the committed fixture on CPU, the GPU tests check
the device decode + evaluation kernels against it, and -- when /root/reference is present -- the extraction is re-run and compared.
"""
from __future__ import annotations

import math
import os
import re
import struct
import sys
import tempfile
import textwrap
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "reference_host.npz"


def _lines(rel: str, first: int, last: int, expect_first: str) -> str:
    """Lines first..last (1-based, inclusive) of a reference file, dedented, with the py2-isms of these ranges made runnable."""
    src = (REF / rel).read_text().splitlines()[first - 1:last]
    assert expect_first in src[0], f"{rel}:{first} changed: {src[0]!r}"
    out = []
    for ln in src:
        if re.match(r"^\s*print\s+[^(]", ln):             # py2 print statement -> no-op of the same indentation
            ln = re.sub(r"print\s+.*$", "pass", ln)
        out.append(ln.replace("xrange(", "range(").replace("\t", "    "))
    return textwrap.dedent("\n".join(out)) + "\n"


class _NumpyOfItsTime:
    """The NumPy the reference was written against took one-element arrays where an integer is expected (``count=2*w*h`` and the shape
    ``(h, w, 2)`` in utils.readFlow, with w and h read as one-element arrays); today's NumPy refuses.  Same functions, arguments unwrapped."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def _int(v):
        return int(np.asarray(v).reshape(-1)[0]) if isinstance(v, np.ndarray) else v

    def fromfile(self, f, dtype=float, count=-1, **kw):
        return np.fromfile(f, dtype, count=self._int(count), **kw)

    def resize(self, a, new_shape):
        return np.resize(a, tuple(self._int(v) for v in new_shape))


def reference_utils():
    """Namespace holding the reference's readFlow and flow_ee."""
    ns = dict(np=_NumpyOfItsTime(), os=os, sys=sys)
    exec(compile(_lines("utils.py", 4, 21, "def readFlow"), str(REF / "utils.py"), "exec"), ns)            # noqa: S102
    exec(compile(_lines("utils.py", 64, 68, "def flow_ee"), str(REF / "utils.py"), "exec"), ns)            # noqa: S102
    return ns


def reference_bilinear(shape):
    """load_deconv_weights (flyingChairsTrain.py:78-92) with a stand-in session: returns the array the reference assigns."""
    src = _lines("flyingChairsTrain.py", 78, 92, "def load_deconv_weights")

    class _Var:
        def __init__(self, shape):
            self.shape = tuple(shape)

        def assign(self, w):
            return ("assign", w)

    class _Sess:
        assigned = None

        def run(self, x):
            if isinstance(x, _Var):
                return np.zeros(x.shape)
            if isinstance(x, tuple) and x[0] == "assign":
                self.assigned = np.array(x[1])
                return None
            raise TypeError(x)

    ns = dict(np=np, math=math)
    exec(compile(src, str(REF / "flyingChairsTrain.py"), "exec"), ns)                                      # noqa: S102
    sess = _Sess()
    ns["load_deconv_weights"](None, _Var(shape), sess)
    return sess.assigned


def reference_eval_aee(flows_all0, prev_all, flow_gt, origin_size, utils_ns):
    """flyingChairsTrain.py:263-267 (per-sample x2 / clip / cv2.resize) and :294-296 (concatenate + utils.flow_ee)."""
    import cv2
    body = _lines("flyingChairsTrain.py", 263, 267, "for batch_idx in xrange(testBatchSize)")

    class _Self:
        pass
    me = _Self()
    me.origin_size = origin_size
    ns = dict(np=np, cv2=cv2, testBatchSize=flows_all0.shape[0], flows_all=[flows_all0], prev_all=prev_all, self=me,
              flow1_list=[], previous_img_list=[])
    exec(compile(body, str(REF / "flyingChairsTrain.py"), "exec"), ns)                                     # noqa: S102
    flow_1 = [np.concatenate(ns["flow1_list"], axis=0)]
    tail = _lines("flyingChairsTrain.py", 294, 296, "f1 = np.concatenate(flow_1")

    class _U:
        flow_ee = staticmethod(utils_ns["flow_ee"])
    ns2 = dict(np=np, flow_1=flow_1, flow_gt=[flow_gt], utils=_U)
    exec(compile(tail, str(REF / "flyingChairsTrain.py"), "exec"), ns2)                                    # noqa: S102
    return flow_1[0], float(ns2["AEE"])


def reference_hook_train_data(data_dir: Path, frame_ids, image_size, utils_ns):
    """flyingChairsLoader.hookTrainData (flyingChairsLoader.py:64-82) on a data-set directory; returns (source, target, flow)."""
    import cv2
    src = _lines("flyingChairsLoader.py", 64, 82, "def hookTrainData")

    class _U:
        readFlow = staticmethod(utils_ns["readFlow"])
    ns = dict(np=np, os=os, cv2=cv2, utils=_U)
    exec(compile(src, str(REF / "flyingChairsLoader.py"), "exec"), ns)                                     # noqa: S102

    class _Self:
        pass
    me = _Self()
    me.trainList, me.img_path, me.image_size = list(frame_ids), str(data_dir), list(image_size)
    return ns["hookTrainData"](me, range(len(frame_ids)))


def write_flo(path: Path, flow: np.ndarray):
    """Middlebury .flo (the reference's own writeFlow references an undefined TAG_CHAR, utils.py:44, so the files are written here)."""
    h, w, _ = flow.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<f", 202021.25))
        f.write(struct.pack("<ii", w, h))
        f.write(flow.astype("<f4").tobytes())


def build():
    import cv2
    rng = np.random.RandomState(20260921)
    U = reference_utils()
    out = {}
    # ---- tiny data set: 3 pairs of 48 x 64 images, flows of the same size, loader output at 32 x 48 (shrinking, like 384x512 -> 320x448) ----
    ids = ["00001", "00002", "00003"]
    with tempfile.TemporaryDirectory() as td:
        d = Path(td)
        for i, fid in enumerate(ids):
            for k in (1, 2):
                img = rng.randint(0, 256, size=(48, 64, 3)).astype(np.uint8)
                img = cv2.GaussianBlur(img, (5, 5), 1.2)                      # (smooth: resize differences would show)
                assert cv2.imwrite(str(d / f"{fid}_img{k}.ppm"), img)
                out[f"ppm_{fid}_{k}"] = np.frombuffer((d / f"{fid}_img{k}.ppm").read_bytes(), dtype=np.uint8)
            flow = (rng.rand(48, 64, 2).astype(np.float32) * 2 - 1) * (3.0 + i)
            write_flo(d / f"{fid}_flow.flo", flow)
            out[f"flo_{fid}"] = np.frombuffer((d / f"{fid}_flow.flo").read_bytes(), dtype=np.uint8)
            out[f"readflow_{fid}"] = U["readFlow"](str(d / f"{fid}_flow.flo"))
        src, tgt, flo = reference_hook_train_data(d, ids, (32, 48), U)
        out["loader_source"], out["loader_target"], out["loader_flow"] = src, tgt, flo
        assert src.dtype == np.uint8 and src.shape == (3, 32, 48, 3) and flo.shape == (3, 48, 64, 2)
    out["ids"] = np.array(ids)
    # ---- average end-point error ----
    f1 = rng.randn(4, 24, 32, 2).astype(np.float32) * 3
    f2 = rng.randn(4, 24, 32, 2).astype(np.float32) * 3
    out["ee_f1"], out["ee_f2"], out["ee_aee"] = f1, f2, np.float64(U["flow_ee"](f1, f2))
    # ---- evaluation recipe: network output at half resolution -> ground-truth size ----
    pr1 = (rng.randn(3, 24, 32, 2) * 40).astype(np.float32)          # (x2 = +-240 with a few values past the clip limits)
    pr1[0, 0, 0] = (200.0, -180.0)
    prev = rng.rand(3, 24, 32, 3).astype(np.float32)
    gt = (rng.randn(3, 48, 64, 2) * 5).astype(np.float32)
    up, aee = reference_eval_aee(pr1, prev, gt, (48, 64), U)
    out["eval_pr1"], out["eval_gt"], out["eval_up"], out["eval_aee"] = pr1, gt, up, np.float64(aee)
    # ---- bilinear deconvolution filters ----
    for shape in ((4, 4, 2, 2), (4, 4, 32, 194), (4, 4, 64, 386)):
        out["bilinear_" + "x".join(map(str, shape))] = reference_bilinear(shape)
    return out


if __name__ == "__main__":
    data = build()
    np.savez_compressed(OUT, **data)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes, {len(data)} arrays)")
