"""The stage bodies of csrc/nq_dither_spec.cuh (speculative segment-parallel Gilbert dither) are scalar
host/device functions: compiled here with g++ and run as a whole pipeline on the CPU, they must reproduce
the oracle's sequential GilbertCurve bit for bit (tests/spec_host_harness.cpp). No GPU involved; the CUDA
kernels wrap the same functions and get their own parity tests under -m gpu."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from nquant_android_b200.synth import make_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "build", "libnq_spec_host.so")
SRC = os.path.join(ROOT, "tests", "spec_host_harness.cpp")
DEPS = [SRC, os.path.join(ROOT, "oracle", "nq_oracle.cpp"), os.path.join(ROOT, "nquant_android_b200", "csrc", "nq_dither_spec.cuh"),
        os.path.join(ROOT, "nquant_android_b200", "csrc", "nq_color.h"), os.path.join(ROOT, "nquant_android_b200", "csrc", "nq_math.h")]
KEYS = ["eligible", "exact", "rounds", "anomaly", "nseg", "segRuns", "slowPixels", "notes", "mismatches", "rejected", "patches", "redos"]


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in DEPS):
        fma = ["-mfma"] if " fma " in open("/proc/cpuinfo").read() else []
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w"] + fma +
                              ["-o", SO, SRC])
    L = ctypes.CDLL(SO)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    L.nqs_spec_host.argtypes = [vp, ci, ci, ci, ci, ctypes.c_uint64, ci, ci, ci, vp]
    L.nqs_spec_batch_add.argtypes = [vp, ci, ci, ci, ctypes.c_uint64, ci, ci]
    L.nqs_spec_batch_run.argtypes = [ci, ci, ci, vp]
    return L


def run(L, w, h, nmax, cls, alpha, seg, warm, cells=1, seed=0xC0FFEE, img_seed=0x5EED0000):
    img = np.ascontiguousarray(make_image(w, h, cls, alpha, seed=img_seed))
    out = np.zeros(12, np.int64)
    assert L.nqs_spec_host(img.ctypes.data, w, h, nmax, 1, seed, seg, warm, cells, out.ctypes.data) == 0
    return dict(zip(KEYS, [int(v) for v in out]))


@pytest.mark.parametrize("w,h,nmax,cls,alpha,seg,warm,cells", [
    (256, 256, 256, "noisy", "opaque", 4096, 1024, 1),      # headline class; error-dependent lookups at the bright corner
    (256, 256, 256, "noisy", "opaque", 4096, 1024, 0),      # same without the candidate lists
    (256, 192, 256, "rand", "opaque", 2048, 512, 1),
    (320, 180, 128, "noisy", "opaque", 2048, 512, 1),
    (173, 211, 200, "noisy", "opaque", 1000, 300, 1),        # ragged sizes, warm-up too short for some segments: re-runs
    (256, 256, 128, "smooth", "opaque", 4096, 1024, 1),      # DITHER_MAX 9 (weight >= .015), ArrayDeque because of <= 128 colours
    (97, 1, 256, "noisy", "opaque", 64, 16, 1),              # a single row: the curve degenerates to a line, two short segments
    (640, 360, 256, "noisy", "opaque", 8192, 1024, 1),       # the production segment length and warm-up
    (333, 127, 65, "rand", "opaque", 1024, 256, 1),          # just above the 64-colour gate; 9 memo patches, 11 rounds
    (256, 256, 256, "noisy", "opaque", 512, 64, 1),          # warm-up far too short: more than two runs per segment
])
def test_spec_pipeline_matches_sequential_oracle(lib, w, h, nmax, cls, alpha, seg, warm, cells):
    r = run(lib, w, h, nmax, cls, alpha, seg, warm, cells)
    assert r["eligible"] == 1
    assert r["anomaly"] == 0 and r["rejected"] == 0
    assert r["mismatches"] == 0 and r["exact"] == 1, r
    assert r["rounds"] >= 1 and r["segRuns"] >= r["nseg"]


def test_spec_declines_what_it_does_not_cover(lib):
    # smooth class -> PriorityQueue mode (GC:87-94); 16 colours -> lookups read the diffused colour
    assert run(lib, 128, 128, 256, "smooth", "opaque", 2048, 512)["eligible"] == 0
    assert run(lib, 128, 128, 16, "noisy", "opaque", 2048, 512)["eligible"] == 0
    # a transparent pixel leaves a constant alpha error in the queue for ever (alpha is never shaped, GC:248): no re-synchronisation
    assert run(lib, 256, 256, 256, "rand", "transparent", 4096, 1024)["eligible"] == 0


def test_kernels_and_wave_loop_emulated_on_a_batch(lib):
    """The CUDA kernels themselves (k_spec_*), run thread by thread with emulated blockIdx/threadIdx, driven by the same
    spec_drive() wave / round loop nq_api.cu uses: five images through wave slots of two (slot reuse across waves), one of
    them needing memo patches. Every image must be completed here and match the sequential oracle."""
    w, h, seg, warm = 256, 192, 2048, 512
    items = [("rand", 0x5EED0000, 0xC0FFEE), ("noisy", 0x5EED0001, 1), ("noisy", 0x5EED0002, 2), ("rand", 0x5EED0003, 3), ("noisy", 0x5EED0004, 4)]
    lib.nqs_spec_batch_begin()
    for cls, iseed, rseed in items:
        img = np.ascontiguousarray(make_image(w, h, cls, "opaque", seed=iseed))
        assert lib.nqs_spec_batch_add(img.ctypes.data, w, h, 256, rseed, seg, warm) == 0
    out = np.zeros(9, np.int64)
    assert lib.nqs_spec_batch_run(seg, warm, 2, out.ctypes.data) == 0
    r = dict(zip(["images", "eligible", "completed", "handed_back", "rounds", "patches", "redos", "wrong", "launches"], [int(v) for v in out]))
    assert r["images"] == 5 and r["eligible"] == 5 and r["completed"] == 5 and r["handed_back"] == 0, r
    assert r["wrong"] == 0, r
    assert r["patches"] >= 1 and r["rounds"] >= 4, r       # three waves, at least one extra round for the patches


def _saturated_corner(w, h, seed, boost):
    """The noisy class with its bright corner pushed into saturation: thousands of pixels that are exactly a palette colour
    (white), whose error-dependent lookups draw or not depending on the diffused error -- the case the draw PREDICTION gets
    wrong hundreds of times."""
    img = make_image(w, h, "noisy", "opaque", seed=seed).astype(np.int64)
    x, y = np.arange(w * h) % w, np.arange(w * h) // w
    m = (x > 0.55 * w) & (y > 0.55 * h)
    ch = [np.where(m, np.minimum(255, c + boost), c) for c in ((img >> 16) & 255, (img >> 8) & 255, img & 255)]
    return np.ascontiguousarray(((255 << 24) | (ch[0] << 16) | (ch[1] << 8) | ch[2]).astype(np.uint32))


def test_hard_image_is_finished_by_sequential_chains(lib):
    """An image whose draws are mispredicted hundreds of times (1 025 'risk' pixels, 2 % error-dependent lookups) used to need
    one validation round per misprediction (257 rounds, then handed back). The first mispredicted segment is now run again
    from its predecessor's exact state as the sequential algorithm itself, its thread going on through the following
    segments that hold error-dependent lookups (a chain), after which everything behind is re-resolved once: a handful of
    rounds, bit-identical to the oracle."""
    w = h = 512
    img = _saturated_corner(w, h, 0x5EED0001, 30)
    out = np.zeros(12, np.int64)
    assert lib.nqs_spec_host(img.ctypes.data, w, h, 256, 1, 0xC0FFEE, 2048, 1024, 1, out.ctypes.data) == 0
    r = dict(zip(KEYS, [int(v) for v in out]))
    assert r["eligible"] == 1 and r["anomaly"] == 0 and r["mismatches"] == 0 and r["exact"] == 1, r
    assert r["rounds"] <= 12 and r["redos"] >= 1, r


def seed_with_rejected_step(t, low17=0x1234):
    """A java.util.Random seed whose t-th generator step delivers next(31) == 2^31 - 1, one of the two values
    Random.nextInt(32767) rejects and draws again for (2 in 2^31 per draw: a few 4K images of every large batch meet one).
    Built by running the 48-bit LCG backwards from such a state."""
    m, a, c = 1 << 48, 0x5DEECE66D, 0xB
    a_inv = pow(a, -1, m)
    s = (0x7FFFFFFF << 17) | low17
    for _ in range(t):
        s = ((s - c) * a_inv) % m
    return s ^ a            # setSeed scrambles with the multiplier


def test_a_draw_that_nextint_rejects_is_followed_exactly(lib, oracle=None):
    """The draw index -> generator step mapping (lcg_step_of): with a seed whose 100 000th step is rejected by nextInt the
    speculative path must still reproduce the sequential oracle (every later draw uses the step after the one its index
    suggests), instead of handing the image to the serial kernel."""
    from oracle import pyoracle
    seed = seed_with_rejected_step(100000)
    vals = pyoracle.java_random_next_int(seed, 32767, 100002)
    vals0 = pyoracle.java_random_next_int(seed_with_rejected_step(100000, 0x1235), 32767, 100002)
    assert len(vals) == 100002            # (the oracle's Random really loops there: the 100 000th call consumes two steps)
    w, h = 512, 512
    img = np.ascontiguousarray(make_image(w, h, "noisy", "opaque"))
    out = np.zeros(12, np.int64)
    assert lib.nqs_spec_host(img.ctypes.data, w, h, 256, 1, seed, 2048, 1024, 1, out.ctypes.data) == 0
    r = dict(zip(KEYS, [int(v) for v in out]))
    assert r["eligible"] == 1 and r["anomaly"] == 0 and r["rejected"] == 0, r
    assert r["mismatches"] == 0 and r["exact"] == 1, r
