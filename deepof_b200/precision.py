"""Accuracy accounting of the tensor-core math modes: per-scale end-point distances between two sets of flow maps and
the evaluation-recipe EPE.  Pure tensor arithmetic -- the REFERENCE side (the CPU oracle's flows) is supplied by the caller (bench.py /
tests), this module never imports it."""
from __future__ import annotations

import torch

# stated tolerances: [mean, max] end-point distance in px of flows_all[s] against the fp32 CPU oracle on the held-out batch,
# and the north_star bound on the evaluation-recipe EPE
TOLERANCE = {
    "fp32": {"epd_mean_px": 2e-5, "epd_max_px": 5e-4, "epe_abs": 1e-3},
    # measured (scripts/precision_report.py, 192x256; bench.py re-measures at 384x512): tf32 mean 1.2e-3..3.3e-3 / max 4e-3..1.7e-2 px,
    # bf16 mean 2.6e-3..4.1e-3 / max 1.2e-2..1.4e-2 px at scale 1 (init .. trained weights); kind::tf32 TRUNCATES its operands to 10
    # mantissa bits while the bf16 operands are rounded to nearest, which is why bf16 is no worse than tf32 here
    "tf32": {"epd_mean_px": 8e-3, "epd_max_px": 5e-2, "epe_abs": 1e-3},
    "bf16": {"epd_mean_px": 8e-3, "epd_max_px": 5e-2, "epe_abs": 1e-3},
}


def end_point_distance(flows_a, flows_b):
    """[(mean, max)] per scale of ||a - b||_2 over all pixels (a, b: lists of [B,h,w,2])."""
    out = []
    for a, b in zip(flows_a, flows_b):
        d = (a.double() - b.double().to(a.device)).pow(2).sum(dim=3).sqrt()
        out.append((float(d.mean()), float(d.max())))
    return out


def flow_magnitude(flows):
    return [float(f.double().pow(2).sum(dim=3).sqrt().mean()) for f in flows]


def within(stats, mode: str) -> bool:
    tol = TOLERANCE[mode]
    return all(m <= tol["epd_mean_px"] and x <= tol["epd_max_px"] for m, x in stats)
