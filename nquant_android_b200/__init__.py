"""B200-native implementation of nQuant's quantizer hot path (PnnQuantizer / PnnLABQuantizer
convert). CUDA kernels + C ABI in csrc/, host-side mirror of the reference interface in quantizer.py."""
from ._build import build  # noqa: F401


def __getattr__(name):
    if name in ("PnnQuantizer", "PnnLABQuantizer", "Context", "NQuantError", "default_context", "gilbert_order"):
        from . import quantizer
        return getattr(quantizer, name)
    raise AttributeError(name)
