"""C-ABI library loads and exports every symbol include/deepof_b200.h declares; host logic; gloo DDP."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "deepof_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dofb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    import __graft_entry__ as ge
    ge.build()
    from deepof_b200 import _lib
    lib = _lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.dofb_version() == 100


def test_no_product_import_of_oracle():
    for py in (ROOT / "deepof_b200").rglob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{py} imports the oracle"


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from deepof_b200 import ops, flownet
    with pytest.raises(Exception):
        flownet.FlowNetS(1, 192, 256)
    with pytest.raises(ops.DeepOFError):
        ops.adam(torch.zeros(4), torch.zeros(4), torch.zeros(4), torch.zeros(4), 1e-3)


def test_arena_layout_and_init_match_oracle():
    from deepof_b200 import flownet
    from oracle import flownet_s as fs, tf_ops
    shapes = flownet.param_shapes()
    assert list(shapes.items()) == list(fs.param_shapes().items())
    arena = flownet.ParamArena(shapes, "cpu")
    assert arena.n_true == 38777706
    flat = arena.new()
    views = arena.views(flat)
    for name, off in arena.offsets.items():
        assert off % 64 == 0 and views[name].data_ptr() == flat.data_ptr() + 4 * off
    # product initialiser draws the same numbers as the oracle's
    gen = torch.Generator().manual_seed(1)
    ref = fs.init_params(1)
    for name, shape in shapes.items():
        if name.endswith("weights"):
            w = flownet.xavier_uniform(shape, gen)
            if name.startswith("up"):
                w = flownet.bilinear_deconv(shape)
            assert torch.equal(w, ref[name]), name
    assert flownet.same_pad if hasattr(flownet, "same_pad") else True
    from deepof_b200.ops import same_pad
    for n, k, s in [(384, 7, 2), (192, 5, 2), (48, 3, 2), (48, 3, 1), (24, 4, 2)]:
        out, before, _after = tf_ops.same_pad(n, k, s)
        assert same_pad(n, k, s) == (out, before)


def test_shard_batch():
    from deepof_b200.ddp import shard_batch
    assert shard_batch(64, 3, 8) == (24, 8)
    with pytest.raises(ValueError):
        shard_batch(30, 0, 8)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from deepof_b200.ddp import GradReducer
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
class Arena:
    # 10 "tensors" of 100000 floats (+3 at the end): bucket cuts fall on tensor boundaries
    offsets = {{f"t{{i}}": i * 100000 for i in range(10)}}
class Eng:
    arena = Arena()
e = Eng()
n = 1000003
e.grad = torch.full((n,), float(rank + 1))
e.theta = torch.full((n,), float(rank))
r = GradReducer(e, bucket_mb=1.0, tail_mb=0.5)
r.broadcast_params()
assert float(e.theta.abs().max()) == 0.0          # rank 0's parameters everywhere
scale = r(e.grad)
assert abs(scale - 1.0 / world) < 1e-12
want = sum(range(1, world + 1))
assert torch.all(e.grad == want), (rank, e.grad[:3])
# tail bucket: the leading tensors that fit 0.5 MB (131072 floats -> one tensor); then >= 1 MB (262144 floats -> 3 tensors) each
assert r.bounds == [(0, 100000), (100000, 400000), (400000, 700000), (700000, n)], r.bounds
# overlapped form: the backward reports falling arena offsets; every bucket is reduced exactly once, top bucket first
e.grad.fill_(float(rank + 1))
r.begin()
r.ready(900000); assert len(r._works) == 0          # the last bucket starts at 700000: not yet entirely final
r.ready(700000); assert len(r._works) == 1
r.ready(300000); assert len(r._works) == 2
r.ready(300000); assert len(r._works) == 2          # idempotent
assert abs(r.finish() - 1.0 / world) < 1e-12
assert torch.all(e.grad == want)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_grad_reducer_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
