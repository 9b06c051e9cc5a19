"""Config E (L=1500, M=1e6) on one GPU: stage times of a few steps (covariance experiments)."""
import sys, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
pkg = g.load_package()
import torch
from gaussdca_jl_b200 import _lib
ctx = pkg.Context(0)
L, M = 1500, 1000000
Z = torch.empty((M, L), dtype=torch.int8, device='cuda')
ctx.check(ctx.lib.gdca_synth_alignment_dev(ctx.h, ctypes.c_void_p(Z.data_ptr()), L, M, 20140321))
n_out = int(ctx.lib.gdca_ranking_length(L, 5))
for it in range(2):
    st = _lib.Stats()
    ctx.check(ctx.lib.gdca_run_resident(ctx.h, ctypes.c_void_p(Z.data_ptr()), L, M, -1.0, 0.8, 0, 5, None, n_out, ctypes.byref(st)))
    ms = ctypes.c_float()
    ctx.lib.gdca_dev_cov_kernel_ms(ctx.h, ctypes.byref(ms))
    print({k: round(v, 1) for k, v in st.asdict().items() if k.startswith('ms_')}, 'cov kernel', round(ms.value, 1), ctx.cov_info()['clusters'], flush=True)
