"""Config C with the sequences in random order (Cs): stage times."""
import sys, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
pkg = g.load_package()
import torch
from gaussdca_jl_b200 import _lib
ctx = pkg.Context(0)
L, M = 500, 200000
Z = torch.empty((M, L), dtype=torch.int8, device='cuda')
ctx.check(ctx.lib.gdca_synth_alignment_dev(ctx.h, ctypes.c_void_p(Z.data_ptr()), L, M, 20140321))
perm = torch.from_numpy(np.random.default_rng(7).permutation(M)).cuda()
Zs = Z[perm].contiguous()
n_out = int(ctx.lib.gdca_ranking_length(L, 5))
R = np.empty(n_out, dtype=_lib.RANK_DTYPE)
for name, A in (("C", Z), ("Cs", Zs)):
    for it in range(3):
        st = _lib.Stats()
        ctx.check(ctx.lib.gdca_run_resident(ctx.h, ctypes.c_void_p(A.data_ptr()), L, M, -1.0, 0.8, 0, 5, _lib.ptr(R), n_out, ctypes.byref(st)))
    c, cap = ctypes.c_int64(), ctypes.c_int64()
    ctx.lib.gdca_dev_pair_list_info(ctx.h, ctypes.byref(c), ctypes.byref(cap))
    mf, mx = ctypes.c_float(), ctypes.c_float()
    ctx.lib.gdca_dev_sweep_info(ctx.h, None, None, None, None, ctypes.byref(mf), ctypes.byref(mx), None)
    print('filter ms', round(mf.value, 3), 'exact ms', round(mx.value, 3))
    d = st.asdict()
    print(name, 'candidates', c.value, 'cap', cap.value, 'meff', d['meff'], {k: round(v, 3) for k, v in d.items() if k.startswith('ms_')}, R[0], flush=True)
