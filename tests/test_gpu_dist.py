"""GPU, >= 2 devices: the real NCCL path of the sharded driver against the single-GPU fused run."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import __graft_entry__ as g
pkg = g.load_package()
from gaussdca_jl_b200 import dist as gd
from gaussdca_jl_b200._lib import ptr
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = pkg.Context(local)
out = {}
# second shape is larger: the library's buffers move, the peer mappings must be re-exchanged transparently
for (L, M, theta, score) in ((64, 30000, "auto", "frob"), (64, 30000, 0.25, "DI"), (96, 41000, "auto", "frob")):
    Z = np.empty((M, L), dtype=np.int8)
    ctx.check(ctx.lib.gdca_synth_alignment(ctx.h, ptr(Z), L, M, 77))
    R, info = gd.gdca_sharded(Z, 0.8, theta, score, 5, ctx=ctx)
    if rank == 0:
        # single-GPU fused run on a SEPARATE context (rank 0's sharded context keeps its peer mappings)
        ctx1 = globals().setdefault("_ctx1", pkg.Context(local))
        R1, st = pkg.gdca_from_alignment(Z, 0.8, theta, score, 5, ctx=ctx1, return_stats=True, as_array=True)
        same_keys = bool(np.array_equal(R["i"], R1["i"]) and np.array_equal(R["j"], R1["j"]))
        out[f"{L}x{M}-{theta}-{score}"] = dict(same_keys=same_keys, max_abs=float(np.max(np.abs(R["score"] - R1["score"]))),
                                               thresh=(info["thresh"], st["thresh"]), meff=(info["meff"], st["meff"]),
                                               passes=info["passes"])
    dist.barrier()
if rank == 0:
    print("RESULT " + json.dumps(out))
dist.destroy_process_group()
'''


def test_nccl_sharded_equals_single_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    import json
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[7:])
    for k, v in out.items():
        assert v["same_keys"], (k, v)            # identical ranking order
        assert v["max_abs"] == 0.0, (k, v)       # bit-identical scores: integer exchange + sums with zeros
        assert v["thresh"][0] == v["thresh"][1] and v["meff"][0] == v["meff"][1]
    assert len(out) == 3 and out["64x30000-auto-frob"]["passes"] == 1
