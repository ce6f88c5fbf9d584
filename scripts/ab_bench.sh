#!/bin/bash
# A/B of two builds of the library on ONE box: ab/libA.so vs ab/libB.so, alternated; prints ms per step and a few layers.
for rep in 1 2; do for v in ${VARIANTS:-A B}; do
  cp ab/lib$v.so deepof_b200/libdeepof_b200.so
  python bench.py --steps 20 --warmup 5 --no-cpu --no-accuracy --no-extras > gpurun_out/ab_$v$rep.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  cp gpurun_out/bench_layers_flownets_bf16_n1.json gpurun_out/ab_layers_$v$rep.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$v$rep.json")); pt=json.load(open("gpurun_out/bench_layers_flownets_bf16_n1.json"))["per_tag_ms"]
print("$v$rep", round(d["ms_per_step"],4), {k: round(pt[k],4) for k in [k for k in pt if any(t in k for t in "${TAGS:-conv_fwd:conv1 conv_dgrad:conv2}".split())]})
PY
done; done
