"""ctypes binding of libgdca_b200.so (include/gdca_b200.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgdca_b200.so")

GDCA_OK, GDCA_ERR_INVALID_ARG, GDCA_ERR_Q_TOO_BIG, GDCA_ERR_NOT_SPD = 0, 1, 2, 3
GDCA_ERR_CUDA, GDCA_ERR_OOM, GDCA_ERR_NO_DEVICE, GDCA_ERR_STATE = 4, 5, 6, 7
SCORE_CODES = {"frob": 0, "DI": 1}

# Tuple{Int,Int,Float64}  (src/GaussDCA.jl:90): 24 bytes, offsets 0/8/16
RANK_DTYPE = np.dtype([("i", np.int64), ("j", np.int64), ("score", np.float64)])
assert RANK_DTYPE.itemsize == 24


class Stats(ctypes.Structure):
    _fields_ = [
        ("L", ctypes.c_int64), ("M", ctypes.c_int64), ("n", ctypes.c_int64),
        ("q", ctypes.c_int32), ("posdef_info", ctypes.c_int32),
        ("theta", ctypes.c_double), ("thresh", ctypes.c_int64), ("meff", ctypes.c_double),
        ("ident_sum", ctypes.c_uint64), ("theta_passes", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("ms_h2d", ctypes.c_float), ("ms_pack", ctypes.c_float), ("ms_theta", ctypes.c_float),
        ("ms_weights", ctypes.c_float), ("ms_cov", ctypes.c_float), ("ms_chol", ctypes.c_float),
        ("ms_inv", ctypes.c_float), ("ms_score", ctypes.c_float), ("ms_apc", ctypes.c_float),
        ("ms_rank", ctypes.c_float), ("ms_d2h", ctypes.c_float), ("ms_total", ctypes.c_float),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class PosDefException(ArithmeticError):
    """Mirror of Julia's LinearAlgebra.PosDefException thrown by cholesky(C) (src/GaussDCA.jl:34)."""

    def __init__(self, info):
        super().__init__(f"matrix is not positive definite; Cholesky factorization failed (info={info})")
        self.info = info


class GdcaError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


_p, _i32, _i64, _u64, _dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
_pi32, _pi64, _pu64, _pdbl = (ctypes.POINTER(t) for t in (_i32, _i64, _u64, _dbl))

# name -> (restype, argtypes): every symbol include/gdca_b200.h declares
SIGNATURES = {
    "gdca_abi_version": (_i32, []),
    "gdca_create": (_i32, [ctypes.POINTER(_p), _i32]),
    "gdca_create_multi": (_i32, [ctypes.POINTER(_p), _pi32, _i32]),
    "gdca_group_size": (_i32, [_p]),
    "gdca_destroy": (None, [_p]),
    "gdca_last_error": (ctypes.c_char_p, [_p]),
    "gdca_status_string": (ctypes.c_char_p, [_i32]),
    "gdca_set_shard": (_i32, [_p, _i32, _i32]),
    "gdca_run": (_i32, [_p, _p, _i64, _i64, _dbl, _dbl, _i32, _i64, _p, _i64, ctypes.POINTER(Stats)]),
    "gdca_ranking_length": (_i64, [_i64, _i64]),
    "gdca_run_resident": (_i32, [_p, _p, _i64, _i64, _dbl, _dbl, _i32, _i64, _p, _i64, ctypes.POINTER(Stats)]),
    "gdca_dev_R_ptr": (_p, [_p]),
    "gdca_compute_weights": (_i32, [_p, _p, _i64, _i64, _dbl, _p, _p, _pdbl, _pdbl, _pi64, _pu64]),
    "gdca_compute_covariance": (_i32, [_p, _p, _i64, _i64, _p, _dbl, _dbl, _p, _p, _pi32]),
    "gdca_inverse": (_i32, [_p, _p, _i64, _p, _pi32]),
    "gdca_score": (_i32, [_p, _p, _p, _i64, _i32, _i32, _p]),
    "gdca_apc": (_i32, [_p, _p, _i64, _p]),
    "gdca_ranking": (_i32, [_p, _p, _i64, _i64, _p, _i64]),
    "gdca_dev_load": (_i32, [_p, _p, _i64, _i64]),
    "gdca_dev_load_resident": (_i32, [_p, _p, _i64, _i64]),
    "gdca_compute_weighted_frequencies": (_i32, [_p, _p, _i64, _i64, _dbl, _p, _p, _pdbl, _p, _pdbl, _pi32]),
    "gdca_add_pseudocount": (_i32, [_p, _p, _p, _i64, _i32, _dbl, _p, _p]),
    "gdca_compute_C": (_i32, [_p, _p, _p, _i64, _p]),
    "gdca_dev_pair_pass": (_i32, [_p, _i32, _i64]),
    "gdca_dev_pair_sample": (_i32, [_p, _i32]),
    "gdca_set_tc_filter": (_i32, [_p, _i32]),
    "gdca_dev_tc_filter": (_i32, [_p, _i64, _p, _p, _i64]),
    "gdca_dev_cov_kernel_ms": (_i32, [_p, ctypes.POINTER(ctypes.c_float)]),
    "gdca_dev_tc_filter_launch_mode": (_i32, [_p]),
    "gdca_set_pair_list": (_i32, [_p, _i32]),
    "gdca_dev_pair_list_info": (_i32, [_p, _pi64, _pi64]),
    "gdca_set_cov_engine": (_i32, [_p, _i32]),
    "gdca_dev_cov_info": (_i32, [_p, _pi32, _pi32, _pi32, _pi64, _pi32, _pdbl, _pdbl]),
    "gdca_set_tc_filter_bits": (_i32, [_p, _i32]),
    "gdca_set_tc_filter_multicast": (_i32, [_p, _i32]),
    "gdca_tc_filter_tile_order": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _p, _i64, _pi64]),
    "gdca_dev_sweep_info": (_i32, [_p, _pi32, _pi64, _pdbl, _pi64, ctypes.POINTER(ctypes.c_float),
                                   ctypes.POINTER(ctypes.c_float), _pdbl]),
    "gdca_dev_ham_sum_ptr": (_p, [_p]),
    "gdca_dev_counts_ptr": (_p, [_p]),
    "gdca_dev_counts_stride": (_i64, [_p]),
    "gdca_theta_from_ham_sum": (_i32, [_i64, _i64, _u64, _pdbl, _pi64, _pu64]),
    "gdca_theta_from_ident_sum": (_i32, [_i64, _i64, _u64, _pdbl, _pi64]),
    "gdca_dev_ident_sum": (_i32, [_p, _pu64]),
    "gdca_dev_finish_weights": (_i32, [_p, _i32, _pdbl]),
    "gdca_dev_set_weights": (_i32, [_p, _p, _dbl]),
    "gdca_dev_covariance": (_i32, [_p, _dbl]),
    "gdca_dev_C_ptr": (_p, [_p]),
    "gdca_dev_npad": (_i64, [_p]),
    "gdca_dev_W_ptr": (_p, [_p]),
    "gdca_dev_inverse": (_i32, [_p, _pi32]),
    "gdca_dev_mJ_ptr": (_p, [_p]),
    "gdca_set_ozaki": (_i32, [_p, _i32]),
    "gdca_set_di_engine": (_i32, [_p, _i32]),
    "gdca_dev_inverse_info": (_i32, [_p, _pi32, _pdbl, _pdbl]),
    "gdca_dev_inverse_shared": (_i32, [_p]),
    "gdca_test_fp64_gemm": (_i32, [_p, _i32, _p, _i32, _p, _i32, _p, _i64, _i64, _i64, _i32, _dbl, _dbl]),
    "gdca_dev_score_rank": (_i32, [_p, _i32, _i64, _p, _i64]),
    "gdca_dev_S_ptr": (_p, [_p]),
    "gdca_dev_peer_export": (_i32, [_p, _p]),
    "gdca_dev_peer_import": (_i32, [_p, _i32, _p]),
    "gdca_dev_peer_valid": (_i32, [_p]),
    "gdca_dev_peer_close": (_i32, [_p]),
    "gdca_dev_zero_counts": (_i32, [_p]),
    "gdca_dev_zero_C": (_i32, [_p]),
    "gdca_dev_sync": (_i32, [_p]),
    "gdca_dev_copy_to_host": (_i32, [_p, _p, _p, _i64]),
    "gdca_dev_get_stats": (_i32, [_p, ctypes.POINTER(Stats)]),
    "gdca_dev_stream": (_p, [_p]),
    "gdca_dev_kernel_launches": (_i64, [_p]),
    "gdca_synth_alignment_dev": (_i32, [_p, _p, _i64, _i64, _u64]),
    "gdca_synth_alignment": (_i32, [_p, _p, _i64, _i64, _u64]),
    "gdca_probe_peaks": (_i32, [_p, _pdbl, _pdbl, _pdbl, _pdbl]),
    "gdca_read_fasta_alignment": (_i32, [ctypes.c_char_p, _dbl, ctypes.POINTER(_p), _pi64, _pi64]),
    "gdca_remove_duplicate_sequences": (_i32, [_p, _i64, _i64, _p, _pi64, _p]),
    "gdca_write_rank": (_i32, [ctypes.c_char_p, _p, _i64]),
    "gdca_format_rank": (_i32, [_p, _i64, _p, _i64, _pi64]),
    "gdca_free_host": (None, [_p]),
    "gdca_host_last_error": (ctypes.c_char_p, []),
}

_lib = None


def load():
    """dlopen the in-tree CUDA library and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  gaussdca.jl_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


def devices_from_env(default=(0,)):
    """GDCA_B200_DEVICES="0,1,2,3" (or a count: "4" = devices 0..3) selects the GPUs of the default context -- the knob a
    Julia or Python user of gDCA(filename) turns to run on more than one GPU (SURVEY 5 "Config/flags", 8b-1)."""
    v = os.environ.get("GDCA_B200_DEVICES", "").strip()
    if not v:
        return tuple(default)
    if "," not in v and v.isdigit() and int(v) > 0 and len(v) <= 2 and not v.startswith("0"):
        return tuple(range(int(v)))
    return tuple(int(x) for x in v.split(",") if x.strip() != "")


class Context:
    """Owns one gdca_ctx: one GPU, or -- devices=[...] -- a group of GPUs of one node driven from this process
    (gdca_create_multi: the alignment is copied once and broadcast over NVLink, the sweep, the covariance and the inversion are
    sharded with their exchanges fused into the kernels).  Reusable across calls."""

    def __init__(self, device: int = 0, devices=None):
        self.lib = load()
        h = _p()
        if devices is not None and len(devices) > 1:
            arr = (ctypes.c_int32 * len(devices))(*[int(d) for d in devices])
            st = self.lib.gdca_create_multi(ctypes.byref(h), arr, len(devices))
            device = int(devices[0])
        else:
            if devices is not None and len(devices) == 1:
                device = int(devices[0])
            st = self.lib.gdca_create(ctypes.byref(h), int(device))
        if st != GDCA_OK:
            msg = self.lib.gdca_last_error(None).decode()
            raise GdcaError(st, f"gdca_create(device={device}, devices={devices}) failed: {msg}")
        self.h = h
        self.device = int(device)
        self.devices = tuple(int(d) for d in devices) if devices is not None else (int(device),)

    def close(self):
        if getattr(self, "h", None):
            self.lib.gdca_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st):
        """Map a gdca_status_t to the exception class the reference would raise."""
        if st == GDCA_OK:
            return
        msg = self.lib.gdca_last_error(self.h).decode()
        if st == GDCA_ERR_INVALID_ARG:
            raise ValueError(msg)  # ArgumentError (src/GaussDCA.jl:50)
        if st == GDCA_ERR_NOT_SPD:
            s = Stats()
            self.lib.gdca_dev_get_stats(self.h, ctypes.byref(s))
            raise PosDefException(s.posdef_info)
        raise GdcaError(st, f"{self.lib.gdca_status_string(st).decode()}: {msg}")

    def set_cov_engine(self, mode: int):
        """0 auto, 1 scatter-add engine, 2 co-occurrence counts per weight class on the FP4 tensor cores (gdca_set_cov_engine)."""
        self.check(self.lib.gdca_set_cov_engine(self.h, int(mode)))

    def set_di_engine(self, mode: int):
        """1 (default) tridiagonalisation + implicit QL, one lane per site pair; 0 one-sided Jacobi (gdca_set_di_engine)."""
        self.check(self.lib.gdca_set_di_engine(self.h, int(mode)))

    def cov_info(self) -> dict:
        """What the last covariance stage ran on (gdca_dev_cov_info)."""
        eng, cls, seg, clu = _i32(), _i32(), _i32(), _i32()
        kb = _i64()
        tf, l2 = _dbl(), _dbl()
        self.check(self.lib.gdca_dev_cov_info(self.h, ctypes.byref(eng), ctypes.byref(cls), ctypes.byref(seg), ctypes.byref(kb),
                                              ctypes.byref(clu), ctypes.byref(tf), ctypes.byref(l2)))
        return dict(engine=eng.value, classes=cls.value, segments=seg.value, kblocks=kb.value, clusters=clu.value,
                    tflop=tf.value, l2_bytes=l2.value)

    def stats(self) -> dict:
        s = Stats()
        self.lib.gdca_dev_get_stats(self.h, ctypes.byref(s))
        return s.asdict()


_default_ctx = {}


def default_context(device=None) -> Context:
    """The process-wide context: device `device`, or the GPUs named by GDCA_B200_DEVICES (default: GPU 0)."""
    key = devices_from_env() if device is None else (int(device),)
    if key not in _default_ctx:
        _default_ctx[key] = Context(devices=list(key))
    return _default_ctx[key]
