"""GPU scratch: time the neighbour-count sweep (tensor-core prefilter + exact sweep) on synthetic config C, both operand types.
Usage: python tools/filter_probe.py [LxM] [reps]   (run under ncu with -k regex:tc_filter_kernel to capture the kernels)"""
import ctypes
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as g

pkg = g.load_package()
from gaussdca_jl_b200._lib import ptr  # noqa: E402

L, M = (500, 200000) if len(sys.argv) < 2 else tuple(map(int, sys.argv[1].split("x")))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = pkg.Context(0)
lib = ctx.lib
Z = np.empty((M, L), dtype=np.int8)
ctx.check(lib.gdca_synth_alignment(ctx.h, ptr(Z), L, M, 20140321))
ctx.check(lib.gdca_dev_load(ctx.h, ptr(Z), L, M))
thresh = L // 2
for bits, mc in ((8, 1), (80, 1), (4, 1)):
    ctx.check(lib.gdca_set_tc_filter_bits(ctx.h, bits))
    ctx.check(lib.gdca_set_tc_filter_multicast(ctx.h, mc))
    for rep in range(reps):
        ctx.check(lib.gdca_dev_pair_pass(ctx.h, 1, thresh))
        filt, tiles, blocks = ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64()
        tf, l2 = ctypes.c_double(), ctypes.c_double()
        msf, msx = ctypes.c_float(), ctypes.c_float()
        ctx.check(lib.gdca_dev_sweep_info(ctx.h, ctypes.byref(filt), ctypes.byref(tiles), ctypes.byref(tf), ctypes.byref(blocks),
                                          ctypes.byref(msf), ctypes.byref(msx), ctypes.byref(l2)))
    print(json.dumps(dict(bits=filt.value, multicast=mc, tiles=tiles.value, tflop=tf.value, ms_filter=msf.value, ms_exact=msx.value,
                          blocks=blocks.value, tflops=tf.value / (msf.value / 1e3) if msf.value else None,
                          l2_tb_s=l2.value / (msf.value / 1e3) / 1e12 if msf.value else None)))
