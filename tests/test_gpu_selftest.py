"""Exhaustive device self-tests of the hand-scheduled fp32 sequences in csrc/rt_math.cuh (rt_selftest, include/rtb200.h):
every input of their domain must give the bits of the compiler's IEEE-correct code — the property the whole parity claim
(`-fmad=false -prec-div=true -prec-sqrt=true`) rests on once a sequence is written by hand."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    import rtb200
    c = rtb200.Context([0])
    yield c
    c.close()


def test_inverse_length_is_two_correctly_rounded_operations(ctx):
    import rtb200
    n, bad = ctx.selftest(rtb200.RT_SELFTEST_INV_LEN)
    assert n == 0x41000000 and bad == 0          # every float in [2^-65, 2^65)


def test_packed_inverse_length_pair_equals_the_scalar_one(ctx):
    """rt_inv_len2 (two normalisations per pass of packed FMUL2 / FFMA2, the two-light Phong pass of shade_light_pair)."""
    import rtb200
    n, bad = ctx.selftest(rtb200.RT_SELFTEST_INV_LEN_PAIR)
    assert n == 2 * 0x41000000 and bad == 0      # every float in [2^-65, 2^65), in either half of the pair


def test_pixel_division_equals_ieee_division(ctx):
    import rtb200
    n, bad = ctx.selftest(rtb200.RT_SELFTEST_PIXEL_DIV)
    assert n == 16384 * 16385 // 2 and bad == 0  # every 0 <= x < w <= 16384


def test_rejected_rsqrt_seed_variant_really_differs(ctx):
    """Documents why rt_inv_len spends a second MUFU: seeding the reciprocal with the rsqrt estimate is NOT correctly rounded."""
    import rtb200
    n, bad = ctx.selftest(rtb200.RT_SELFTEST_INV_LEN_RSQ_SEED)
    print("rsqrt-seeded variant: %d mismatches of %d" % (bad, n))
    assert bad > 0


def test_unknown_selftest_is_rejected(ctx):
    with pytest.raises(RuntimeError):
        ctx.selftest(99)


def test_big_frame_takes_the_ieee_division_kernels(ctx):
    """A frame side above RT_FASTDIV_MAX is outside the verified range of the fast division: the launch must fall back to the
    IEEE-division kernels and still equal the oracle."""
    import numpy as np
    import oracle_lib as O
    import scenes
    sc = scenes.default_scene()
    w, h = 16400, 8
    cam = scenes.make_camera(width=w, height=h)
    ctx.set_scene(sc)
    px, _ = ctx.render(cam, w, h, 4, 1, 0)
    ref = O.render(sc, cam, w, h, 4)
    assert np.array_equal(px.reshape(h, w), ref["pixels"])
