"""Reference-derived fixture for the bilinear flow-warp: runs the reference's OWN NumPy restatement of the warp.

The reference holds exactly one runnable (NumPy-only) piece of its hot path: the author's check of the warp,
``/root/reference/check_loss.py:61-135``.  This script does not copy it: it READS those lines from the reference checkout at run time,
executes them (the only py2-isms in that range are ``xrange`` and commented-out ``print`` statements) on synthetic inputs and
stores inputs + outputs in ``tests/golden/check_loss_warp.npz``.  The inputs are chosen so that every bilinear corner stays inside
the image, where the script's flat-index clamp (:106-109) and the TF graph's per-axis clamp (flyingChairsWrapFlow.py:815-818) agree.

    python tests/golden/make_check_loss_golden.py            # writes the fixture (needs /root/reference)

``tests/test_oracle_reference_pin.py`` checks the oracle (and, on the GPU box, the CUDA kernel) against the committed fixture, and --
when /root/reference is present -- re-runs this extraction and checks that the fixture still is what the reference code produces.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/check_loss.py")
FIRST, LAST = 61, 135             # 1-based, inclusive: "flows = pr6" ... "recons = np.transpose(np.asarray(reconstructs))"
OUT = Path(__file__).resolve().parent / "check_loss_warp.npz"


def reference_warp_source() -> str:
    lines = REF.read_text().splitlines()[FIRST - 1:LAST]
    assert lines[0].strip() == "flows = pr6" and lines[-1].startswith("recons = np.transpose"), "reference file changed"
    return "\n".join(lines) + "\n"


def run_reference_warp(flow_scaled: np.ndarray, target: np.ndarray) -> np.ndarray:
    """flow_scaled [h,w,2] (already multiplied by flow_scale, as check_loss.py:42 does), target [h,w,C] -> recon [h*w, C]."""
    height, width, channels = target.shape
    ns = dict(np=np, xrange=range, pr6=flow_scaled, outputs_flat=target.reshape(-1, channels), height=height, width=width,
              channels=channels)
    exec(compile(reference_warp_source(), str(REF), "exec"), ns)        # noqa: S102 -- the reference's own code, read-only checkout
    return ns["recons"]


def cases():
    """(name, flow_scaled [h,w,2], target [h,w,3]); corners stay inside the image (interior property asserted)."""
    rng = np.random.RandomState(20260921)
    out = []
    for name, (h, w), amp in (("s6_6x8", (6, 8), 1.4), ("s4_24x32", (24, 32), 4.0), ("s3_48x64", (48, 64), 7.5)):
        flow = (rng.rand(h, w, 2) * 2 - 1) * amp
        rows, cols = np.mgrid[0:h, 0:w]
        # pull every sample position into [0, size-1): floor >= 0 and floor + 1 <= size - 1
        flow[..., 0] = np.clip(cols + flow[..., 0], 0, w - 1 - 1e-3) - cols
        flow[..., 1] = np.clip(rows + flow[..., 1], 0, h - 1 - 1e-3) - rows
        sub = flow[::5, ::3]                                # exact-integer flows (zero fractional weight) as well
        sub[..., 0] = np.clip(np.round(cols[::5, ::3] + sub[..., 0]), 0, w - 2) - cols[::5, ::3]
        sub[..., 1] = np.clip(np.round(rows[::5, ::3] + sub[..., 1]), 0, h - 2) - rows[::5, ::3]
        target = rng.rand(h, w, 3) - 0.4
        fx, fy = np.floor(cols + flow[..., 0]), np.floor(rows + flow[..., 1])
        assert fx.min() >= 0 and fy.min() >= 0 and fx.max() <= w - 2 and fy.max() <= h - 2      # all four corners inside the image
        out.append((name, flow, target))
    return out


def main():
    if not REF.exists():
        sys.exit(f"{REF} not found: the fixture can only be generated where the reference checkout is present")
    blob = {}
    for name, flow, target in cases():
        recon = run_reference_warp(flow, target)
        blob[name + "_flow"] = flow.astype(np.float64)
        blob[name + "_target"] = target.astype(np.float64)
        blob[name + "_recon"] = recon.astype(np.float64)
    np.savez_compressed(OUT, **blob)
    print(OUT, OUT.stat().st_size, "bytes;", len(blob) // 3, "cases")


if __name__ == "__main__":
    main()
