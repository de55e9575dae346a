"""CPU: host-side planners of the kernels added in the second half of round 2 (no kernel is launched):
  * which RGB layers qualify for the thin16 kernels (bf16 fat side), pass by pass;
  * how the one-pass instance norm cuts an image into a thread-block cluster - a function of the plane and the storage
    type only, never of the batch (N-GPU = 1-GPU relies on that);
  * argument validation of the new entry points fails loudly without a GPU."""
import ctypes

import _srgan_lib


def _desc(N, H, W, C, K, R, stride, pad):
    d = _srgan_lib.ConvDesc()
    d.N, d.H, d.W, d.C, d.K, d.R, d.S = N, H, W, C, K, R, R
    d.stride, d.pad = stride, pad
    d.P = (H + 2 * pad - R) // stride + 1
    d.Q = (W + 2 * pad - R) // stride + 1
    d.xs_n = d.xs_h = d.xs_w = d.xs_c = 0
    return d


def test_thin16_eligibility():
    lib = _srgan_lib.load()
    sup = lambda d, p: lib.srgan_conv2d_thin16_supported(d, p)
    stem = _desc(64, 128, 128, 3, 64, 7, 1, 3)            # generator stem: every pass
    assert [sup(stem, p) for p in (0, 1, 2)] == [1, 1, 1]
    head = _desc(64, 128, 128, 64, 3, 7, 1, 3)            # generator head: every pass
    assert [sup(head, p) for p in (0, 1, 2)] == [1, 1, 1]
    d1 = _desc(128, 128, 128, 3, 64, 4, 2, 1)             # wide discriminator stem (stride 2): no thin16 dgrad
    assert [sup(d1, p) for p in (0, 1, 2)] == [1, 0, 1]
    d2 = _desc(128, 64, 64, 3, 32, 4, 2, 1)               # narrow tower: 32 channels cannot feed the bf16 wgrad
    assert sup(d2, 0) == 1 and sup(d2, 2) == 0
    efirst = _desc(64, 128, 128, 3, 64, 7, 2, 1)          # encoder stem
    assert [sup(efirst, p) for p in (0, 1, 2)] == [1, 0, 1]
    trunk = _desc(64, 32, 32, 256, 256, 3, 1, 1)          # not an RGB layer
    assert [sup(trunk, p) for p in (0, 1, 2)] == [0, 0, 0]
    wide = _desc(2, 256, 256, 64, 3, 7, 1, 3)             # head on 256-wide images: the row kernel holds <= 128 pixels
    assert sup(wide, 0) == 0
    for d, p in ((stem, 0), (stem, 1), (stem, 2), (head, 0), (head, 1), (head, 2)):
        assert lib.srgan_conv2d_thin16_workspace(d, p) > 0
    assert lib.srgan_conv2d_thin16_workspace(trunk, 0) == 0


def test_onepass_norm_plan_depends_on_the_plane_only():
    lib = _srgan_lib.load()
    F32, BF16 = 0, 1
    before = lib.srgan_inorm_onepass_enable(1)
    try:
        def plan(hw, c, dt):
            cl, sl = ctypes.c_int(0), ctypes.c_int(0)
            ok = lib.srgan_inorm_onepass_plan(hw, c, dt, ctypes.byref(cl), ctypes.byref(sl))
            return ok, cl.value, sl.value
        assert plan(32 * 32, 256, BF16) == (1, 8, 128)         # residual trunk: 8 CTAs x 128 pixels x 512 B = 64 KB each
        assert plan(32 * 32, 256, F32) == (0, 0, 0)            # 1 MB per image: more than 8 x 64 KB
        assert plan(15 * 15, 256, BF16) == (1, 2, 128)
        assert plan(31 * 31, 128, F32) == (1, 8, 128)
        assert plan(31 * 31, 128, BF16) == (1, 4, 256)
        assert plan(24 * 24, 128, BF16)[:2] == (1, 3)          # odd cluster sizes are fine
        assert plan(16 * 16, 64, BF16)[:2] == (1, 1)
        assert plan(128 * 128, 64, BF16)[0] == 0 and plan(49, 512, BF16)[0] == 0 and plan(64, 12, BF16)[0] == 0
        for hw, c, dt in ((1024, 256, BF16), (961, 128, F32), (3844, 64, BF16)):
            ok, cl, sl = plan(hw, c, dt)
            assert ok and cl <= 8 and cl * sl >= hw and (cl - 1) * sl < hw      # every CTA of the cluster has rows
            assert sl * c * (2 if dt == BF16 else 4) <= 65536
        lib.srgan_inorm_onepass_enable(0)
        assert plan(32 * 32, 256, BF16)[0] == 0                # switched off: the two-kernel path
        assert lib.srgan_inorm_onepass_enable(-1) == 0         # query only
    finally:
        lib.srgan_inorm_onepass_enable(before)


def test_new_entry_points_validate_arguments():
    lib = _srgan_lib.load()
    assert lib.srgan_cross_entropy_fwd(None, None, None, None, 4, 4, None) == -1
    assert lib.srgan_cross_entropy_fwd(16, 16, 16, 16, 0, 4, None) == -1
    assert lib.srgan_prdc_kth_radius(16, 16, 8, 8, None) == -1 and b"k < N" in lib.srgan_last_error()
    assert lib.srgan_prdc_pairdist2(16, 16, 16, 0, 0, 8, None) == 0            # empty: nothing to do
    assert lib.srgan_grad_fold(16, 16, None, 6, None) == -1                    # n % 4
    assert lib.srgan_grad_fold(16, 16, None, 0, None) == 0
    trunk = _desc(1, 8, 8, 64, 64, 3, 1, 1)
    assert lib.srgan_conv2d_fprop_thin16(trunk, 16, 16, None, 16, 0, 0.0, None, 0, None) < 0
    assert b"thin16" in lib.srgan_last_error()
