"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libnquant_b200.so), against the
oracle on the same seeded inputs -- stage by stage (bins, initial find_nn, merge sequence, palette,
saliency) and end to end (ARGB output). Everything is integer/bit exact: no tolerances anywhere."""
import hashlib
import json
import os

import numpy as np
import pytest

from nquant_android_b200.synth import make_image

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype("<u4").tobytes()).hexdigest()


def test_native_library_is_the_cuda_build(gpu_ctx):
    from nquant_android_b200 import _lib
    assert os.path.basename(_lib.SO) == "libnquant_b200.so" and os.path.exists(_lib.SO)
    assert gpu_ctx.kernel_launches() >= 1


def test_device_math_is_bit_identical_to_host(gpu_ctx, oracle):
    rng = np.random.default_rng(11)
    n = 4000
    sets = {
        "pow": (rng.uniform(0.003, 1.3, n), np.repeat([2.4, 1 / 3.0, 1 / 2.4, 7.0], n // 4)),
        "exp": (rng.uniform(-40, 5, n), None), "tanh": (rng.uniform(-25, 25, n), None),
        "cbrt": (rng.integers(1, 1 << 24, n).astype(np.float64), None),
        "atan2": (rng.uniform(-130, 130, n), rng.uniform(-130, 130, n)),
        "sin": (rng.uniform(-55, 55, n), None), "cos": (rng.uniform(-55, 55, n), None),
    }
    for name, (x, y) in sets.items():
        got = gpu_ctx.math(name, x, y)
        ref = np.array([oracle.math_fn(name, float(a), 0.0 if y is None else float(b), 0)
                        for a, b in zip(x, y if y is not None else np.zeros(n))])
        assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), name



def _lab_pairs(rng, n):
    """Lab pairs shaped like histogram bins: random, near-duplicates, opposite/equal hues, axis cases."""
    L = rng.uniform(0, 100, (n, 2)).astype(np.float32)
    A = rng.uniform(-110, 110, (n, 2)).astype(np.float32)
    B = rng.uniform(-110, 110, (n, 2)).astype(np.float32)
    k = n // 8
    # neighbours (what find_nn mostly sees): second colour = first + small step
    A[:k, 1] = A[:k, 0] + rng.normal(0, 1.5, k).astype(np.float32)
    B[:k, 1] = B[:k, 0] + rng.normal(0, 1.5, k).astype(np.float32)
    # same hue, different chroma (dh ~ 0) and exactly opposite hues (|dh| ~ pi, hsum ~ 2 pi)
    s = rng.uniform(0.1, 3.0, k).astype(np.float32)
    A[k:2 * k, 1] = A[k:2 * k, 0] * s
    B[k:2 * k, 1] = B[k:2 * k, 0] * s
    A[2 * k:3 * k, 1] = -A[2 * k:3 * k, 0] * s
    B[2 * k:3 * k, 1] = -B[2 * k:3 * k, 0] * s
    A[3 * k:4 * k, 1] = A[3 * k:4 * k, 0]
    B[3 * k:4 * k, 1] = -B[3 * k:4 * k, 0]
    # axis cases and greys
    A[4 * k:4 * k + k // 4, 0] = 0
    B[4 * k + k // 4:4 * k + k // 2, 1] = 0
    A[4 * k + k // 2:5 * k, :] = 0
    B[4 * k + k // 2:4 * k + 3 * k // 4, :] = 0
    # tiny chroma
    A[5 * k:6 * k] *= 1e-3
    B[5 * k:6 * k] *= 1e-3
    lab1 = np.stack([L[:, 0], A[:, 0], B[:, 0]], axis=1)
    lab2 = np.stack([L[:, 1], A[:, 1], B[:, 1]], axis=1)
    return lab1, lab2


def test_ciede_filter_never_changes_a_float(gpu_ctx, oracle):
    """The plain-double filter in front of the correctly rounded CIEDE2000 kernels (nq_fastmath.cuh) must
    return exactly the floats of the exact path (oracle: CIELABConvertor.java:91-194 restated)."""
    rng = np.random.default_rng(2024)
    n = 400000
    lab1, lab2 = _lab_pairs(rng, n)
    got, n_exact = gpu_ctx.ciede(lab1, lab2)
    ref = oracle.ciede_parts_batch(lab1, lab2)
    bad = np.nonzero((got.view(np.uint32) != ref.view(np.uint32)).any(axis=1))[0]
    assert bad.size == 0, (bad[:5], lab1[bad[:5]], lab2[bad[:5]], got[bad[:5]], ref[bad[:5]])
    # 3/8 of that set is built to defeat the filter; on bin-like pairs it must decide nearly everything
    m = 100000
    L = rng.uniform(0, 100, (m, 2)).astype(np.float32)
    A = rng.uniform(-110, 110, (m, 1)).astype(np.float32) + rng.normal(0, 3, (m, 2)).astype(np.float32)
    B = rng.uniform(-110, 110, (m, 1)).astype(np.float32) + rng.normal(0, 3, (m, 2)).astype(np.float32)
    l1, l2 = np.stack([L[:, 0], A[:, 0], B[:, 0]], axis=1), np.stack([L[:, 1], A[:, 1], B[:, 1]], axis=1)
    got, n_exact = gpu_ctx.ciede(l1, l2)
    assert np.array_equal(got.view(np.uint32), oracle.ciede_parts_batch(l1, l2).view(np.uint32))
    assert n_exact < m // 100, f"exact path used for {n_exact} of {m} neighbouring pairs"


CASES = [
    # kind, class, alpha, W, H, K, dither
    (0, "noisy", "opaque", 160, 120, 256, True), (1, "noisy", "opaque", 160, 120, 256, True),
    (0, "smooth", "opaque", 160, 120, 256, True), (1, "smooth", "opaque", 160, 120, 256, True),
    (0, "rand", "opaque", 128, 128, 64, True), (1, "rand", "opaque", 128, 128, 64, True),
    (0, "noisy", "semi", 128, 96, 16, True), (1, "noisy", "semi", 128, 96, 16, True),
    (0, "noisy", "transparent", 128, 96, 64, True), (1, "noisy", "transparent", 128, 96, 64, True),
    (0, "noisy", "opaque", 128, 96, 256, False), (0, "noisy", "opaque", 128, 96, 16, False),
    (1, "noisy", "opaque", 128, 96, 16, False), (1, "smooth", "opaque", 128, 96, 32, True),
    (0, "smooth", "opaque", 64, 48, 2, True), (1, "smooth", "transparent", 64, 48, 2, True),
    (1, "smooth", "opaque", 128, 96, 4, True), (0, "noisy", "opaque", 128, 96, 3, True),
    (0, "noisy", "semi", 96, 64, 256, True), (1, "smooth", "semi", 96, 64, 256, True),
    (0, "noisy", "opaque", 1, 97, 8, True), (1, "noisy", "opaque", 97, 1, 8, True),
    (0, "noisy", "opaque", 37, 211, 128, True), (1, "noisy", "opaque", 211, 37, 128, True),
    # dither == false with more than 32 colours: Gilbert pass + BlueNoise.dither second pass; for PnnLABQuantizer its
    # weight depends on pixelMap.size() (PnnLABQuantizer.java:511-513)
    (1, "noisy", "opaque", 160, 120, 64, False), (1, "noisy", "opaque", 160, 120, 256, False),
    (1, "rand", "opaque", 400, 300, 64, False), (1, "noisy", "transparent", 128, 96, 64, False),
    (0, "noisy", "transparent", 128, 96, 64, False), (1, "smooth", "opaque", 160, 120, 256, False),
    (1, "smooth", "opaque", 640, 480, 64, False), (1, "smooth", "opaque", 800, 600, 40, False),   # weight != 1
    # at most 12 (13 with the transparent one) distinct colours: PnnLABQuantizer returns pixelMap.keySet() in HashMap
    # order (PnnLABQuantizer.java:193-206); PnnQuantizer keeps its bins
    (1, "few", "opaque", 96, 64, 32, True), (1, "few", "transparent", 96, 64, 32, True), (1, "few", "opaque", 96, 64, 16, False),
    (1, "few", "opaque", 96, 64, 8, True), (0, "few", "transparent", 96, 64, 32, True), (1, "few", "semi", 64, 48, 256, True),
]


@pytest.mark.parametrize("kind,cls,alpha,W,H,K,dither", CASES)
def test_stage_and_output_parity(gpu_ctx, oracle, kind, cls, alpha, W, H, K, dither):
    img = make_image(W, H, cls, alpha)
    seed = 0xC0FFEE + K
    ref = oracle.convert(kind, img, W, H, K, dither, seed=seed)
    gpu_ctx.set_debug(True)
    try:
        out, pal, plen, ha = gpu_ctx.convert_batch(kind, img[None, :], W, H, K, dither, seeds=[seed])
        info = gpu_ctx.image_info(0)
        s = ref.scalars
        assert info["has_semi_transparency"] == s["hasSemiTransparency"]
        assert info["transparent_pixel_index"] == s["transparentPixelIndex"]
        assert bool(ha[0]) == (s["transparentPixelIndex"] >= 0)
        if K > 2 and len(ref.bins):
            assert info["maxbins"] == s["maxbins"] and info["quan_rt"] == s["quan_rt"]
            assert info["weight"] == s["weight"] and info["ratio_merge"] == s["ratio_merge"]
            bins, ierr, inn = gpu_ctx.debug_bins(0)
            assert np.array_equal(bins, ref.bins), "histogram bins (means, counts) differ"
            assert np.array_equal(ierr.view(np.uint32), ref.init_err.view(np.uint32)) and np.array_equal(inn, ref.init_nn)
            assert np.array_equal(gpu_ctx.debug_merges(0), ref.merges), "merge sequence differs"
        for g, r in [("g_margin", "margin"), ("g_thresold", "thresold"), ("g_dither_max_q", "DITHER_MAX"),
                     ("g_dither_max", "ditherMax"), ("g_sorted", "sortedByYDiff"), ("g_has_alpha", "hasAlpha")]:
            assert info[g] == s[r], g
        assert np.float32(info["g_beta"]) == np.float32(s["beta"])
        assert np.array_equal(pal[0, :plen[0]], ref.palette), "palette differs"
        if len(ref.saliencies):
            assert np.array_equal(gpu_ctx.debug_saliencies(W * H, 0).view(np.uint32), ref.saliencies.view(np.uint32))
        assert info["rng_draws"] == s["rng_draws"]
        if kind == 1 and not dither and plen[0] > 32:
            assert np.float32(info["bn_weight"]) == np.float32(s["bn_weight"]), (info["bn_weight"], s["bn_weight"], s["pixelMapSize"])
        assert np.array_equal(out[0], ref.out), f"{int((out[0] != ref.out).sum())} of {W * H} output pixels differ"
    finally:
        gpu_ctx.set_debug(False)


def test_golden_fixtures(gpu_ctx):
    cases = json.load(open(os.path.join(HERE, "golden", "oracle_cases.json")))
    for c in cases:
        img = make_image(c["w"], c["h"], c["cls"], c["alpha"])
        out, pal, plen, _ = gpu_ctx.convert_batch(c["kind"], img[None, :], c["w"], c["h"], c["k"], bool(c["dither"]), seeds=[c["seed"]])
        assert plen[0] == c["palette_len"] and sha(pal[0, :plen[0]]) == c["palette_sha"], c
        assert sha(out[0]) == c["output_sha"], c


def test_reference_config0_512_rgb(gpu_ctx, oracle):
    """BASELINE.json configs[0]: PnnQuantizer 256 colours, dither on, 512x512 gradient+noise."""
    W = H = 512
    img = make_image(W, H, "noisy", "opaque")
    ref = oracle.convert(0, img, W, H, 256, True, trace=False)
    out, pal, plen, _ = gpu_ctx.convert_batch(0, img[None, :], W, H, 256, True)
    assert np.array_equal(pal[0, :plen[0]], ref.palette) and np.array_equal(out[0], ref.out)


def test_headline_class_512_lab(gpu_ctx, oracle):
    """The bench workload's class and quantizer (noisy, PnnLABQuantizer, 256 colours, dither on) at the largest
    size the oracle finishes in seconds: 26 778 bins, 58 016 rescans, every pruning layer of the merge loop and the
    pre-lookup dither path at work. Palette and output bit-exact, RNG draws equal."""
    W = H = 512
    img = make_image(W, H, "noisy", "opaque")
    ref = oracle.convert(1, img, W, H, 256, True, seed=0xC0FFEE, trace=False)
    out, pal, plen, _ = gpu_ctx.convert_batch(1, img[None, :], W, H, 256, True, seeds=[0xC0FFEE])
    assert np.array_equal(pal[0, :plen[0]], ref.palette) and np.array_equal(out[0], ref.out)
    assert gpu_ctx.image_info(0)["rng_draws"] == ref.scalars["rng_draws"]


def test_reference_config1_1080p_lab(gpu_ctx, oracle):
    """BASELINE.json configs[1]: PnnLABQuantizer 256 colours, dither on, 1920x1080 (smooth class so
    the oracle finishes in seconds)."""
    W, H = 1920, 1080
    img = make_image(W, H, "smooth", "opaque")
    ref = oracle.convert(1, img, W, H, 256, True, seed=99, trace=False)
    out, pal, plen, _ = gpu_ctx.convert_batch(1, img[None, :], W, H, 256, True, seeds=[99])
    assert np.array_equal(pal[0, :plen[0]], ref.palette) and np.array_equal(out[0], ref.out)


def test_batch_matches_single_and_device_matches_host(gpu_ctx):
    import torch
    W, H, K = 96, 80, 64
    imgs = np.stack([make_image(W, H, c, "opaque", seed=0x5EED0000 + i) for i, c in enumerate(["noisy", "smooth", "rand", "noisy"])])
    seeds = np.arange(4, dtype=np.uint64) + 7
    for kind in (0, 1):
        out, pal, plen, _ = gpu_ctx.convert_batch(kind, imgs, W, H, K, True, seeds=seeds)
        for i in range(4):
            o1, p1, l1, _ = gpu_ctx.convert_batch(kind, imgs[i:i + 1], W, H, K, True, seeds=seeds[i:i + 1])
            assert np.array_equal(o1[0], out[i]) and np.array_equal(p1[0], pal[i])
        din = torch.from_numpy(imgs.view(np.int32)).cuda()
        dout = torch.empty_like(din)
        gpu_ctx.convert_batch_ptr(kind, din.data_ptr(), dout.data_ptr(), 4, W, H, K, True, seeds=seeds, device=True)
        assert np.array_equal(dout.cpu().numpy().view(np.uint32), out)


def test_large_mixed_batch_matches_single(gpu_ctx):
    """40 different images in one call (more than the 32 sets of CIELAB sort scratch, several merge CTAs per SM,
    every synthetic class and alpha mode side by side) must give each image the result it gets alone."""
    W, H, K = 72, 56, 64
    classes, alphas = ["noisy", "smooth", "rand", "few"], ["opaque", "transparent", "semi", "opaque"]
    imgs = np.stack([make_image(W, H, classes[i % 4], alphas[(i // 4) % 4], seed=0x5EED0000 + i) for i in range(40)])
    seeds = np.arange(40, dtype=np.uint64) * 977 + 5
    for kind in (0, 1):
        out, pal, plen, ha = gpu_ctx.convert_batch(kind, imgs, W, H, K, True, seeds=seeds)
        for i in range(0, 40, 3):
            o1, p1, l1, h1 = gpu_ctx.convert_batch(kind, imgs[i:i + 1], W, H, K, True, seeds=seeds[i:i + 1])
            assert l1[0] == plen[i] and np.array_equal(p1[0], pal[i]) and np.array_equal(o1[0], out[i]) and h1[0] == ha[i], (kind, i)


def test_device_synth_matches_numpy(gpu_ctx):
    import torch
    W, H = 70, 50
    for cls_i, cls in enumerate(["smooth", "noisy", "rand"]):
        for am_i, am in enumerate(["opaque", "transparent", "semi"]):
            d = torch.empty(2 * W * H, dtype=torch.int32, device="cuda")
            gpu_ctx.synth_device(d.data_ptr(), 2, W, H, cls_i, am_i, 0x5EED0000)
            got = d.cpu().numpy().view(np.uint32).reshape(2, -1)
            for i in range(2):
                assert np.array_equal(got[i], make_image(W, H, cls, am, seed=0x5EED0000 + i)), (cls, am, i)


@pytest.mark.parametrize("kind,K,alpha", [(1, 256, "opaque"), (0, 16, "semi")])
def test_full_size_4k_properties(gpu_ctx, kind, K, alpha):
    """BASELINE.json configs[2]/[3] at full 3840x2160: size-independent properties (the oracle needs
    ~1 minute per 4K image, so it is not run here)."""
    W, H = 3840, 2160
    img = make_image(W, H, "noisy", alpha)
    imgs = np.stack([img, img])
    out, pal, plen, ha = gpu_ctx.convert_batch(kind, imgs, W, H, K, True, seeds=[5, 5])
    assert plen[0] == K and np.array_equal(pal[0], pal[1])
    assert np.array_equal(out[0], out[1]), "two copies of one image in a batch must quantize identically"
    palette = pal[0, :K]
    assert np.isin(out[0], palette).all(), "every output pixel must be a palette colour (GilbertCurve.java:279)"
    info = gpu_ctx.image_info(0)
    assert info["merges"] == info["maxbins"] - K
    again = gpu_ctx.dither_with_palette(kind, img, W, H, K, True, palette, seed=5)
    assert np.array_equal(again, out[0]), "stage hook with the same palette must reproduce convert()"
    err = np.abs(((out[0][:, None] >> np.array([16, 8, 0])) & 255).astype(np.int32) - ((img[:, None] >> np.array([16, 8, 0])) & 255).astype(np.int32))
    assert err.mean() < 40


def test_full_size_8192_properties(gpu_ctx):
    """BASELINE.json configs[4] size (8192x8192, 268 MB per image): size-independent properties of the headline
    quantizer (PnnLABQuantizer, 256 colours, dither on) -- the oracle would need minutes for one such image."""
    W = H = 8192
    img = make_image(W, H, "noisy", "opaque")
    out, pal, plen, ha = gpu_ctx.convert_batch(1, img[None, :], W, H, 256, True, seeds=[11])
    assert plen[0] == 256 and not ha[0]
    assert np.isin(out[0], pal[0, :256]).all(), "every output pixel must be a palette colour (GilbertCurve.java:279)"
    info = gpu_ctx.image_info(0)
    assert info["merges"] == info["maxbins"] - 256 and info["rng_draws"] > W * H * 0.9
    sub = slice(0, W * 64)   # the first 64 rows are enough for the error statistic
    err = np.abs(((out[0][sub, None] >> np.array([16, 8, 0])) & 255).astype(np.int32) - ((img[sub, None] >> np.array([16, 8, 0])) & 255).astype(np.int32))
    assert err.mean() < 40


def test_error_codes(gpu_ctx):
    from nquant_android_b200.quantizer import NQuantError
    with pytest.raises(NQuantError) as e:   # pixelMap.size() is not tracked when getLab's calls depend on scan order
        gpu_ctx.convert_batch(1, make_image(96, 64, "noisy", "semi")[None, :], 96, 64, 64, False)
    assert e.value.code == -4
    img = make_image(8, 8)
    with pytest.raises(NQuantError) as e:
        gpu_ctx.convert_batch(0, img[None, :], 8, 8, 1, True)
    assert e.value.code == -2
    with pytest.raises(NQuantError) as e:
        gpu_ctx.convert_batch(0, img[None, :], 8, 8, 257, True)
    assert e.value.code == -4
    with pytest.raises(NQuantError):
        gpu_ctx.convert_batch(7, img[None, :], 8, 8, 16, True)


def test_mirror_classes(oracle):
    from nquant_android_b200.quantizer import PnnQuantizer, PnnLABQuantizer
    W, H = 64, 64
    img = make_image(W, H, "noisy", "transparent")
    for cls, kind in ((PnnQuantizer, 0), (PnnLABQuantizer, 1)):
        q = cls(img, W, H, rng_seed=3)
        out = q.convert(32, True)
        ref = oracle.convert(kind, img, W, H, 32, True, seed=3, trace=False)
        assert q.hasAlpha() and np.array_equal(out, ref.out) and np.array_equal(q.palette, ref.palette)


def test_reference_demo_picture(gpu_ctx):
    """The one real image the reference ships (its demo's sample.jpg, here as the lossless tests/golden/sample_495x438.png,
    tools/make_sample_fixture.py) through imageio.load_argb, as the demo converts it (MainActivity.java:190-194:
    PnnQuantizer, 256 colours, dither) and four other settings, against the oracle's frozen palette and output."""
    from nquant_android_b200.imageio import load_argb
    argb, w, h = load_argb(os.path.join(HERE, "golden", "sample_495x438.png"))
    for c in json.load(open(os.path.join(HERE, "golden", "sample_cases.json"))):
        assert (w, h) == (c["w"], c["h"]) and sha(argb) == c["input_sha"]
        out, pal, plen, _ = gpu_ctx.convert_batch(c["kind"], argb[None, :], w, h, c["k"], bool(c["dither"]), seeds=[c["seed"]])
        assert [int(v) for v in pal[0, :plen[0]]] == c["palette"], c
        assert sha(out[0]) == c["output_sha"], (c["kind"], c["k"], c["dither"])
        assert gpu_ctx.image_info(0)["rng_draws"] == c["rng_draws"]


def test_caller_stream_orders_input_and_output(gpu_ctx):
    """nq_set_stream: the library's work starts after what the caller's stream already holds and the stream continues only
    after the call. Handle 0 is CUDA's legacy default stream (torch's default stream), not the context's own: a torch kernel
    that PRODUCES the input immediately before convert_batch_ptr(device=True) must be seen, on the default stream and on a
    side stream."""
    import torch
    W, H, K = 256, 128, 64
    img = make_image(W, H, "noisy", "opaque")
    ref_out, ref_pal, ref_len, _ = gpu_ctx.convert_batch(1, img[None, :], W, H, K, True, seeds=[3])
    src = torch.from_numpy(img.view(np.int32)).cuda()
    try:
        for stream in (torch.cuda.default_stream(), torch.cuda.Stream()):
            gpu_ctx.set_stream(stream.cuda_stream)
            with torch.cuda.stream(stream):
                din = torch.zeros_like(src)
                dout = torch.empty_like(src)
                big = torch.empty(1 << 28, dtype=torch.int32, device="cuda")
                for _ in range(4):
                    big.fill_(1)                 # keeps the stream busy so an unordered library would read zeros
                din.copy_(src + big[:src.numel()] - 1)
                gpu_ctx.convert_batch_ptr(1, din.data_ptr(), dout.data_ptr(), 1, W, H, K, True, seeds=np.array([3], dtype=np.uint64), device=True)
                res = dout.clone()               # enqueued on the caller's stream right behind the call
            stream.synchronize()
            assert np.array_equal(res.cpu().numpy().view(np.uint32), ref_out[0]), stream
            del big
    finally:
        gpu_ctx.reset_stream()


def test_debug_records_survive_small_workspaces(oracle):
    """Debug mode with more images than workspace slots would index the records of a later group with the group-local
    index; here: several images in debug mode, every image's merge sequence must be its own."""
    from nquant_android_b200.quantizer import Context
    ctx = Context(0)
    try:
        W, H, K = 64, 48, 16
        imgs = np.stack([make_image(W, H, "noisy", "opaque", seed=0x5EED0000 + i) for i in range(5)])
        ctx.set_debug(True)
        ctx.convert_batch(0, imgs, W, H, K, True)
        for i in range(5):
            ref = oracle.convert(0, imgs[i], W, H, K, True)
            assert np.array_equal(ctx.debug_merges(i), ref.merges), i
    finally:
        ctx.close()


def test_convert_batch_multi_dynamic_queue():
    """nq_convert_batch_multi: several contexts pull pieces of one batch from a shared queue (one context per GPU of the
    node; on a single-GPU box two contexts on device 0 exercise the same threads and queue). Every image must land at its
    own index with the result a single-context call gives."""
    import torch
    from nquant_android_b200.quantizer import Context, convert_batch_multi, NQuantError
    W, H, K, n = 96, 64, 64, 13
    imgs = np.stack([make_image(W, H, ["noisy", "smooth", "rand"][i % 3], "opaque", seed=0x5EED0000 + i) for i in range(n)])
    seeds = np.arange(n, dtype=np.uint64) + 21
    ndev = torch.cuda.device_count()
    ctxs = [Context(g % ndev) for g in range(max(2, min(ndev, 4)))]
    try:
        ref = ctxs[0].convert_batch(1, imgs, W, H, K, True, seeds=seeds)
        for q in (0, 3, 1):
            got = convert_batch_multi(ctxs, 1, imgs, W, H, K, True, seeds=seeds, queue_images=q)
            for a, b in zip(ref, got):
                assert np.array_equal(a, b), q
        with pytest.raises(NQuantError):
            convert_batch_multi(ctxs, 1, imgs, W, H, 1, True)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_stage_hook_histogram(gpu_ctx, oracle, kind):
    """nq_histogram: the front of pnnquan alone (alpha scan, histogram, initial find_nn sweep), against the oracle's bins and
    first nearest neighbours; and the context still converts afterwards."""
    W, H, K = 200, 150, 64
    img = make_image(W, H, "noisy", "transparent")
    ref = oracle.convert(kind, img, W, H, K, True, seed=9)
    bins, err, nn = gpu_ctx.histogram(kind, img, W, H, K)
    info = gpu_ctx.image_info(0)
    assert info["maxbins"] == ref.scalars["maxbins"] == len(bins) and info["weight"] == ref.scalars["weight"]
    assert np.array_equal(bins, ref.bins)
    assert np.array_equal(err.view(np.uint32), ref.init_err.view(np.uint32)) and np.array_equal(nn, ref.init_nn)
    out, pal, plen, _ = gpu_ctx.convert_batch(kind, img[None, :], W, H, K, True, seeds=[9])
    assert np.array_equal(out[0], ref.out)
