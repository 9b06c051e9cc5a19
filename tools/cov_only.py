"""One covariance stage at a given shape (profiling helper): load, weights, then gdca_dev_covariance a few times."""
import sys, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
pkg = g.load_package()
import torch
from gaussdca_jl_b200 import _lib
ctx = pkg.Context(0)
L, M = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
Z = torch.empty((M, L), dtype=torch.int8, device='cuda')
ctx.check(ctx.lib.gdca_synth_alignment_dev(ctx.h, ctypes.c_void_p(Z.data_ptr()), L, M, 20140321))
ctx.check(ctx.lib.gdca_dev_load_resident(ctx.h, ctypes.c_void_p(Z.data_ptr()), L, M))
ident = ctypes.c_uint64()
ctx.check(ctx.lib.gdca_dev_ident_sum(ctx.h, ctypes.byref(ident)))
th, thr = ctypes.c_double(), ctypes.c_int64()
ctx.lib.gdca_theta_from_ident_sum(L, M, ident.value, ctypes.byref(th), ctypes.byref(thr))
ctx.check(ctx.lib.gdca_dev_pair_pass(ctx.h, 1, thr.value))
meff = ctypes.c_double()
ctx.check(ctx.lib.gdca_dev_finish_weights(ctx.h, 0, ctypes.byref(meff)))
for it in range(reps):
    ctx.check(ctx.lib.gdca_dev_covariance(ctx.h, 0.8))
    ctx.check(ctx.lib.gdca_dev_sync(ctx.h))
    ms = ctypes.c_float()
    ctx.lib.gdca_dev_cov_kernel_ms(ctx.h, ctypes.byref(ms))
    print(it, ctx.cov_info(), ms.value, meff.value, flush=True)
