"""GPU parity of the speculative segment-parallel dither (csrc/nq_dither_spec.cuh, the default path for the images
that qualify): the result must be bit-identical to the oracle's sequential GilbertCurve, and the path must actually
have taken the images it is meant for (nq_get_spec_stats). The stage bodies are checked on the CPU in
tests/test_spec_dither_host.py; these tests add the kernels and the orchestration around them (rolling admission into
a pool of work-array slots, the serial kernels running next to it). Set NQ_SPEC_DITHER_TEST=0 to skip these tests."""
import os

import numpy as np
import pytest

from nquant_android_b200.synth import make_image

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("NQ_SPEC_DITHER_TEST") == "0", reason="NQ_SPEC_DITHER_TEST=0")]


@pytest.fixture()
def spec_ctx():
    from nquant_android_b200 import _build
    from nquant_android_b200.quantizer import Context
    _build.build()
    ctx = Context(0)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("w,h,nmax,cls,seg,warm,seed", [
    (256, 256, 256, "noisy", 4096, 1024, 0xC0FFEE),
    (256, 192, 256, "rand", 2048, 512, 7),
    (320, 180, 128, "noisy", 2048, 512, 99),
    (173, 211, 200, "noisy", 1000, 300, 5),        # warm-up too short for some segments: exact re-runs
    (512, 512, 256, "noisy", 8192, 1024, 0xC0FFEE),
    (256, 256, 128, "smooth", 4096, 1024, 3),         # DITHER_MAX 9
    (256, 192, 256, "rand", 2048, 512, 0xC0FFEE),    # a memo entry created by an error-dependent lookup is patched in (1 patch on the CPU harness)
])
def test_spec_dither_is_bit_identical(spec_ctx, oracle, w, h, nmax, cls, seg, warm, seed):
    img = make_image(w, h, cls, "opaque")
    ref = oracle.convert(1, img, w, h, nmax, True, seed=seed, trace=False)
    spec_ctx.set_spec_dither(True, seg, warm)
    out, pal, plen, _ = spec_ctx.convert_batch(1, img[None, :], w, h, nmax, True, seeds=[seed])
    assert np.array_equal(pal[0, :plen[0]], ref.palette)
    assert np.array_equal(out[0], ref.out)
    assert spec_ctx.image_info(0)["rng_draws"] == ref.scalars["rng_draws"]
    assert spec_ctx.spec_stats()["images"] == 1, "the speculative path did not take (or finish) the image"


def test_spec_dither_corrects_draw_mispredictions(spec_ctx, oracle):
    """Image 13 of the 1080p probe batch: five error-dependent lookups draw against their prediction (CPU harness: 5 re-resolves,
    6 rounds); the path must finish it itself, bit-identical."""
    w, h, i = 1920, 1080, 13
    img = make_image(w, h, "noisy", "opaque", seed=0x5EED0000 + i)
    ref = oracle.convert(1, img, w, h, 256, True, seed=0xC0FFEE + i, trace=False)
    spec_ctx.set_spec_dither(True, 8192, 1024)
    out, pal, plen, _ = spec_ctx.convert_batch(1, img[None, :], w, h, 256, True, seeds=[0xC0FFEE + i])
    assert np.array_equal(out[0], ref.out)
    assert spec_ctx.image_info(0)["rng_draws"] == ref.scalars["rng_draws"]
    st = spec_ctx.spec_stats()
    assert st["images"] == 1 and st["fallbacks"] == 0 and st["rounds"] >= 2, st


def test_spec_dither_batch_and_mixed_eligibility(spec_ctx, oracle):
    w, h = 256, 256
    imgs = np.stack([make_image(w, h, "noisy", "opaque", seed=0x5EED0000 + i) for i in range(3)] +
                    [make_image(w, h, "smooth", "opaque")])
    seeds = [11, 12, 13, 14]
    spec_ctx.set_spec_dither(True, 4096, 1024)
    out, pal, plen, _ = spec_ctx.convert_batch(1, imgs, w, h, 256, True, seeds=seeds)
    for i in range(len(imgs)):
        ref = oracle.convert(1, imgs[i], w, h, 256, True, seed=seeds[i], trace=False)
        assert np.array_equal(pal[i, :plen[i]], ref.palette), i
        assert np.array_equal(out[i], ref.out), i
    assert spec_ctx.spec_stats()["images"] == 3       # the PriorityQueue-mode image belongs to k_dither_sorted
    # an image that k_dither_fifo has to do (a transparent pixel) runs NEXT to the speculative rounds of the others
    imgs2 = np.stack([imgs[0], make_image(w, h, "rand", "transparent")])
    out2, pal2, plen2, _ = spec_ctx.convert_batch(1, imgs2, w, h, 256, True, seeds=[11, 15])
    for i, sd in enumerate([11, 15]):
        ref = oracle.convert(1, imgs2[i], w, h, 256, True, seed=sd, trace=False)
        assert np.array_equal(out2[i], ref.out), i
    assert spec_ctx.spec_stats()["images"] == 4


def test_spec_dither_slot_reuse_and_chunks(oracle, monkeypatch):
    """More images than work-array slots (NQ_SPEC_SLOTS=3: images are admitted as slots free up, every launch mixes images
    of different ages) and more than one pipeline chunk (nq_set_chunk_images): bit-identical to the oracle, all completed
    by the speculative path."""
    from nquant_android_b200.quantizer import Context
    monkeypatch.setenv("NQ_SPEC_SLOTS", "3")
    ctx = Context(0)
    try:
        w, h, n = 256, 192, 11
        imgs = np.stack([make_image(w, h, "noisy" if i % 3 else "rand", "opaque", seed=0x5EED0000 + i) for i in range(n)])
        seeds = [100 + i for i in range(n)]
        ctx.set_spec_dither(True, 2048, 512)
        ctx.set_chunk_images(6)
        out, pal, plen, _ = ctx.convert_batch(1, imgs, w, h, 256, True, seeds=seeds)
        for i in range(n):
            ref = oracle.convert(1, imgs[i], w, h, 256, True, seed=seeds[i], trace=False)
            assert np.array_equal(pal[i, :plen[i]], ref.palette), i
            assert np.array_equal(out[i], ref.out), i
            assert ctx.image_info(i)["rng_draws"] == ref.scalars["rng_draws"], i
        st = ctx.spec_stats()
        assert st["images"] == n and st["fallbacks"] == 0, st
    finally:
        ctx.close()


def test_spec_dither_leaves_other_quantizers_alone(spec_ctx, oracle):
    w, h = 192, 160
    img = make_image(w, h, "noisy", "opaque")
    spec_ctx.set_spec_dither(True, 2048, 512)
    for kind, nmax, dither in ((0, 256, True), (1, 16, True), (1, 256, False)):
        ref = oracle.convert(kind, img, w, h, nmax, dither, seed=3, trace=False)
        out, pal, plen, _ = spec_ctx.convert_batch(kind, img[None, :], w, h, nmax, dither, seeds=[3])
        assert np.array_equal(out[0], ref.out), (kind, nmax, dither)
    assert spec_ctx.spec_stats()["images"] == 0


def test_spec_dither_hard_image_sequential_chains(spec_ctx, oracle):
    """A saturated bright corner: thousands of pixels that are exactly a palette colour, whose error-dependent lookups draw
    or not depending on the diffused error, so the draw prediction fails hundreds of times. The mispredicted stretch is
    re-run as the sequential algorithm from its predecessor's exact state (a chain of segments on one thread) and everything
    behind it re-resolved: the path finishes the image itself, bit-identical, in a handful of rounds."""
    from test_spec_dither_host import _saturated_corner
    w = h = 512
    img = _saturated_corner(w, h, 0x5EED0001, 30)
    ref = oracle.convert(1, img, w, h, 256, True, seed=0xC0FFEE, trace=False)
    spec_ctx.set_spec_dither(True, 2048, 1024)
    out, pal, plen, _ = spec_ctx.convert_batch(1, img[None, :], w, h, 256, True, seeds=[0xC0FFEE])
    assert np.array_equal(pal[0, :plen[0]], ref.palette)
    assert np.array_equal(out[0], ref.out)
    assert spec_ctx.image_info(0)["rng_draws"] == ref.scalars["rng_draws"]
    st = spec_ctx.spec_stats()
    assert st["images"] == 1 and st["fallbacks"] == 0 and st["rounds"] <= 12, st


def test_spec_dither_follows_a_rejected_nextint(spec_ctx, oracle):
    """Random.nextInt(32767) draws again when next(31) is one of its two top values (2 in 2^31 per draw: a few 4K images of a
    1024-image batch meet one). The path maps draw indices to generator steps around such steps (k_spec_rejects,
    lcg_step_of) instead of handing the image to the serial kernel. Seed built so that the 100 000th step is rejected."""
    from test_spec_dither_host import seed_with_rejected_step
    seed = seed_with_rejected_step(100000)
    w = h = 512
    img = make_image(w, h, "noisy", "opaque")
    ref = oracle.convert(1, img, w, h, 256, True, seed=seed, trace=False)
    for spec in (True, False):                 # the speculative path and the serial kernels must both follow it
        spec_ctx.set_spec_dither(spec, 2048, 1024)
        out, pal, plen, _ = spec_ctx.convert_batch(1, img[None, :], w, h, 256, True, seeds=[seed])
        assert np.array_equal(out[0], ref.out), spec
        assert spec_ctx.image_info(0)["rng_draws"] == ref.scalars["rng_draws"]
    st = spec_ctx.spec_stats()
    assert st["images"] == 1 and st["fallbacks"] == 0, st
