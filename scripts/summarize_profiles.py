"""Turn the files scripts/profile_gpu.sh leaves in gpurun_out/ into the tables of profiles/rNN_ncu_summary.md (run in the dev container).

    python scripts/summarize_profiles.py bf16 > /tmp/tables.md
"""
import collections
import csv
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
math = sys.argv[1] if len(sys.argv) > 1 else "bf16"

COLS = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def launch_table():
    path = OUT / f"launches_{math}.csv"
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("dofb::", "")
        v = float(r[iv].replace(",", ""))
        v = v / 1000.0 if r[iu] in ("ns", "nsecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total µs | share |\n|---|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:80]}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% |")
    tc = sum(a[1] for k, a in agg.items() if k.startswith("tc_"))
    print(f"\ntcgen05 kernels (`tc_*`): {100 * tc / tot:.1f}% of the device time of this capture; total {tot / 1000:.2f} ms over {sum(a[0] for a in agg.values())} launches.\n")


def rep_table(rep):
    txt = subprocess.run(["ncu", "-i", str(OUT / rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    print("| " + " | ".join(hdr[i].replace("__", "\\_\\_") for i in idx) + " |\n|" + "---|" * len(idx))
    print("| " + " | ".join(units[i] for i in idx) + " |")
    for r in rows[2:]:
        print("| " + " | ".join(re.sub(r"\(.*", "", r[i]).replace("void ", "")[:60] if i == idx[0] else r[i] for i in idx) + " |")
    print()


print(f"## Launch list (`ncu --metrics gpu__time_duration.sum`; full CSV: launches_{math}.csv)\n")
launch_table()
for rep, title in (("prof_tc_gemm.ncu-rep", "tcgen05 gather-GEMM (conv fwd / dgrad / transposed conv)"), ("prof_tc_wgrad.ncu-rep", "tcgen05 weight gradient"),
                   ("prof_hbm.ncu-rep", "streaming kernels (flow heads, warp+loss, Adam)")):
    if (OUT / rep).exists():
        print(f"## {title}, `ncu --set full`  (`gpurun_out/{rep}`)\n")
        rep_table(rep)


def traffic_json():
    """gpurun_out/tc_traffic_<math>.csv (ncu dram bytes per tc_* launch of one step) -> profiles/r02_tc_traffic_<math>.json (bench.py: roofline.traffic)."""
    import json
    path = OUT / f"tc_traffic_{math}.csv"
    if not path.exists():
        return
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[iid], {"kernel": re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("dofb::", ""), "bytes": 0.0, "us": 0.0})
        v = float(r[iv].replace(",", ""))
        if r[im].startswith("dram__bytes"):
            d["bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[iu], 1.0)
        elif r[im].startswith("gpu__time"):
            d["us"] = v / 1000.0 if r[iu] in ("ns", "nsecond") else v
    launches = list(per.values())
    build = subprocess.run(["git", "log", "--format=%h", "-1"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:tc_ python scripts/prof_heads.py {math} "
                     "(one training step, B=32, 384x512)", "build": build, "launches": len(launches),
           "dram_bytes_total": sum(d["bytes"] for d in launches), "dram_bytes_per_launch": sum(d["bytes"] for d in launches) / max(len(launches), 1),
           "time_us_total": sum(d["us"] for d in launches),
           "per_launch": [{"kernel": d["kernel"], "us": round(d["us"], 1), "dram_MB": round(d["bytes"] / 1e6, 1)} for d in sorted(launches, key=lambda d: -d["us"])]}
    (ROOT / "profiles" / f"r02_tc_traffic_{math}.json").write_text(json.dumps(out, indent=1))
    print(f"tc_* DRAM traffic: {out['dram_bytes_total'] / 1e9:.2f} GB over {len(launches)} launches -> profiles/r02_tc_traffic_{math}.json\n")


traffic_json()


def layers_table(path=OUT / "prof_layers_raw.csv"):
    """Per-layer table from the NVTX-filtered capture of scripts/profile_layers.sh (raw page exported on the GPU box)."""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "µs"), ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
            ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2→SM GB"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
    cols = [(c, t) for c, t in cols if c in ix]
    print("| NVTX range / kernel | " + " | ".join(t for _, t in cols) + " | DRAM TB/s | L2→SM TB/s |\n|---|" + "---:|" * (len(cols) + 2))
    seen = set()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("dofb::", "")
        name = re.sub(r"<.*", "", name)
        if name in seen or "vectorized_elementwise" in name:
            continue
        seen.add(name)
        us = float(r[ix["gpu__time_duration.sum"]])
        rd, wr = float(r[ix["dram__bytes_read.sum"]]), float(r[ix["dram__bytes_write.sum"]])
        xb = float(r[ix["l1tex__m_xbar2l1tex_read_bytes.sum"]])
        vals = [r[ix[c]] for c, _ in cols]
        vals = [f"{float(v):.1f}" if re.fullmatch(r"[0-9.]+", v) and "." in v else v for v in vals]
        print(f"| `{name[:70]}` | " + " | ".join(vals) + f" | {(rd + wr) / us:.2f} | {xb * 1e3 / us:.2f} |")
    print()


if (OUT / "prof_layers_raw.csv").exists():
    print("## Named layers of one bf16 step (B = 32, 384x512), `ncu --set full --nvtx-include <tag>/` (scripts/profile_layers.sh)\n")
    layers_table()
