"""Several GPUs behind ONE context (gdca_create_multi): the call a Julia or Python user makes, gDCA() / gdca_run(), on a device
group.  The alignment goes to the first device once and is broadcast over NVLink; sweep, covariance and inversion are sharded with
their exchanges fused into the kernels.  The contract: the ranking is BIT-identical to the single-GPU run (integer exchanges,
disjoint writes, exact digit arithmetic in the shared GEMMs).

Two flavours: members that share one physical GPU (env GDCA_GROUP_ALLOW_SAME_DEVICE=1) -- this runs on a one-GPU box and covers
every line of the group path -- and real groups on >= 2 GPUs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # L, M, theta, score, pc           n = 20 L
    (128, 6000, "auto", "frob", 0.8),  # n = 2560: sliced engine, shared trtri levels (h = 4, 8, 16), sharded lauum
    (130, 3000, 0.3, "DI", 0.2),       # n = 2600 -> padded to 2688 (21 blocks): partial last group
    (40, 20000, "auto", "frob", 0.8),  # n = 800: DMMA only, prefilter on (M >= 16384): group = sweep + covariance sharding
    (64, 500, 0.0, "frob", 0.5),       # theta = 0: no sweep at all
]


def run_cases(pkg, orc, ctx1, ctxg):
    for (L, M, theta, score, pc) in CASES:
        Z = orc.synth_alignment(L, M, seed=L + M)
        R1, s1 = pkg.gdca_from_alignment(Z, pc, theta, score, 5, ctx=ctx1, return_stats=True, as_array=True)
        Rg, sg = pkg.gdca_from_alignment(Z, pc, theta, score, 5, ctx=ctxg, return_stats=True, as_array=True)
        assert sg["thresh"] == s1["thresh"] and sg["theta"] == s1["theta"] and sg["meff"] == s1["meff"], (L, M)
        assert np.array_equal(Rg["i"], R1["i"]) and np.array_equal(Rg["j"], R1["j"]), (L, M, theta, score)
        assert np.array_equal(Rg["score"], R1["score"]), (L, M, float(np.max(np.abs(Rg["score"] - R1["score"]))))
    # and the group context is reusable, also after an error
    Zs = np.ones((50, 30), dtype=np.int8)
    Zs[:, 0] = np.arange(50) % 21 + 1
    with pytest.raises(pkg.PosDefException):
        pkg.gdca_from_alignment(Zs, 0.0, 0.2, "frob", 5, ctx=ctxg)
    Z = orc.synth_alignment(128, 3000, seed=1)
    assert np.array_equal(pkg.gdca_from_alignment(Z, ctx=ctxg, as_array=True), pkg.gdca_from_alignment(Z, ctx=ctx1, as_array=True))


@pytest.mark.parametrize("members", [2, 3, 4])
def test_group_of_members_on_one_gpu_is_bit_identical_to_single(pkg, orc, ctx, monkeypatch, members):
    monkeypatch.setenv("GDCA_GROUP_ALLOW_SAME_DEVICE", "1")
    ctxg = pkg.Context(devices=[0] * members)
    assert ctxg.lib.gdca_group_size(ctxg.h) == members
    try:
        run_cases(pkg, orc, ctx, ctxg)
    finally:
        ctxg.close()


@pytest.mark.parametrize("members", [2, 3, 4])
def test_shared_factorisation_is_bit_identical_to_single(pkg, orc, ctx, monkeypatch, members):
    """The trailing update of the Cholesky shared by block columns (owners apply the broadcast panels and ship the next panel's
    columns back one step ahead).  It switches on by itself at n >= 16 384; GDCA_SHARE_MIN_NB=16 forces it at test sizes."""
    import ctypes
    monkeypatch.setenv("GDCA_GROUP_ALLOW_SAME_DEVICE", "1")
    monkeypatch.setenv("GDCA_SHARE_MIN_NB", "16")
    ctxg = pkg.Context(devices=[0] * members)
    try:
        # n = 2560 (20 blocks), 4000 (32 blocks, padded), 4100 (33 blocks: partial last outer block)
        for (L, M, theta, score, pc) in [(128, 6000, "auto", "frob", 0.8), (200, 3000, 0.3, "DI", 0.2), (205, 2500, "auto", "frob", 0.5)]:
            Z = orc.synth_alignment(L, M, seed=3 * L + M)
            R1 = pkg.gdca_from_alignment(Z, pc, theta, score, 5, ctx=ctx, as_array=True)
            Rg = pkg.gdca_from_alignment(Z, pc, theta, score, 5, ctx=ctxg, as_array=True)
            assert np.array_equal(Rg["i"], R1["i"]) and np.array_equal(Rg["j"], R1["j"]), (L, M)
            assert np.array_equal(Rg["score"], R1["score"]), (L, M, float(np.max(np.abs(Rg["score"] - R1["score"]))))
            assert ctxg.lib.gdca_dev_inverse_shared(ctxg.h) == 1 and ctx.lib.gdca_dev_inverse_shared(ctx.h) == 0
    finally:
        ctxg.close()


def test_repeated_device_is_rejected_outside_test_mode(pkg, monkeypatch):
    monkeypatch.delenv("GDCA_GROUP_ALLOW_SAME_DEVICE", raising=False)
    with pytest.raises(pkg.GdcaError, match="listed twice"):
        pkg.Context(devices=[0, 0])


def test_real_device_group_is_bit_identical_to_single(pkg, orc, ctx):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    for world in [w for w in (2, 4, 8) if w <= n]:
        ctxg = pkg.Context(devices=list(range(world)))
        try:
            run_cases(pkg, orc, ctx, ctxg)
        finally:
            ctxg.close()


def test_devices_env_knob(pkg, monkeypatch):
    from gaussdca_jl_b200 import _lib
    monkeypatch.setenv("GDCA_B200_DEVICES", "0,2,3")
    assert _lib.devices_from_env() == (0, 2, 3)
    monkeypatch.setenv("GDCA_B200_DEVICES", "4")
    assert _lib.devices_from_env() == (0, 1, 2, 3)
    monkeypatch.delenv("GDCA_B200_DEVICES")
    assert _lib.devices_from_env() == (0,)
