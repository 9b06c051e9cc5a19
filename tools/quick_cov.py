"""Covariance stage alone at config C, both engines (developer timing loop)."""
import sys, time, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
pkg = g.load_package()
import torch
ctx = pkg.Context(0)
L, M = int(sys.argv[1]) if len(sys.argv) > 1 else 500, int(sys.argv[2]) if len(sys.argv) > 2 else 200000
Z = torch.empty((M, L), dtype=torch.int8, device='cuda')
ctx.check(ctx.lib.gdca_synth_alignment_dev(ctx.h, ctypes.c_void_p(Z.data_ptr()), L, M, 20140321))
n_out = int(ctx.lib.gdca_ranking_length(L, 5))
R = np.empty(n_out, dtype=pkg.RANK_DTYPE) if hasattr(pkg, 'RANK_DTYPE') else None
from gaussdca_jl_b200 import _lib
R = np.empty(n_out, dtype=_lib.RANK_DTYPE)
for eng in (2, 1, 0):
    ctx.set_cov_engine(eng)
    for it in range(3):
        st = _lib.Stats()
        ctx.check(ctx.lib.gdca_run_resident(ctx.h, ctypes.c_void_p(Z.data_ptr()), L, M, -1.0, 0.8, 0, 5, _lib.ptr(R), n_out, ctypes.byref(st)))
    ms = ctypes.c_float()
    ctx.lib.gdca_dev_cov_kernel_ms(ctx.h, ctypes.byref(ms))
    d = st.asdict()
    print(eng, ctx.cov_info(), 'cov kernel ms', ms.value, {k: round(v, 3) for k, v in d.items() if k.startswith('ms_')}, R[0], flush=True)
