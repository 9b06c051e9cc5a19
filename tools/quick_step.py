"""Stage times of a few resident steps of config C (no argument) or config D (argument 1: DI score, pseudocount 0.2) on one GPU:
   python tools/quick_step.py [0|1]      (developer loop: one gpurun call of ~25 s)"""
import sys, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
pkg = g.load_package()
import torch
from gaussdca_jl_b200 import _lib
ctx = pkg.Context(0)
L, M = 500, 200000
SCORE = int(sys.argv[1]) if len(sys.argv) > 1 else 0   # 0 frob (config C), 1 DI with pseudocount 0.2 (config D)
PC = 0.2 if SCORE else 0.8
Z = torch.empty((M, L), dtype=torch.int8, device='cuda')
ctx.check(ctx.lib.gdca_synth_alignment_dev(ctx.h, ctypes.c_void_p(Z.data_ptr()), L, M, 20140321))
n_out = int(ctx.lib.gdca_ranking_length(L, 5))
R = np.empty(n_out, dtype=_lib.RANK_DTYPE)
acc = {}
for it in range(6):
    st = _lib.Stats()
    ctx.check(ctx.lib.gdca_run_resident(ctx.h, ctypes.c_void_p(Z.data_ptr()), L, M, -1.0, PC, SCORE, 5, _lib.ptr(R), n_out, ctypes.byref(st)))
    if it >= 2:
        for k, v in st.asdict().items():
            if k.startswith('ms_'): acc[k] = acc.get(k, 0) + v / 4
print({k: round(v, 3) for k, v in acc.items()}, R[0], flush=True)
