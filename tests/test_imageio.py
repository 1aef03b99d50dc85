"""Indexed-PNG writer / ARGB loader either side of the path. CPU only: the quantized image comes from
the oracle."""
import numpy as np
import pytest

from nquant_android_b200 import imageio
from nquant_android_b200.synth import make_image

PIL = pytest.importorskip("PIL.Image")


@pytest.mark.parametrize("kind,nmax,alpha", [(0, 16, "opaque"), (1, 256, "opaque"), (0, 32, "semi"), (1, 64, "transparent")])
def test_indexed_png_round_trip(tmp_path, oracle, kind, nmax, alpha):
    w, h = 96, 64
    img = make_image(w, h, "noisy", alpha)
    ref = oracle.convert(kind, img, w, h, nmax, True, seed=5, trace=False)
    path = tmp_path / "q.png"
    nbytes = imageio.write_indexed_png(str(path), ref.out, ref.palette, w, h)
    assert nbytes == path.stat().st_size
    im = PIL.open(str(path))
    assert im.mode == "P" and im.size == (w, h)
    back, w2, h2 = imageio.load_argb(str(path))
    assert (w2, h2) == (w, h)
    assert np.array_equal(back, ref.out)          # decoding the indexed file gives convert()'s ARGB back, alpha included
    idx = imageio.to_indices(ref.out, ref.palette)
    assert np.array_equal(ref.palette[idx], ref.out)


def test_to_indices_rejects_foreign_colours():
    with pytest.raises(ValueError):
        imageio.to_indices(np.array([0xFF000001], np.uint32), np.array([0xFF000000, 0xFFFFFFFF], np.uint32))
