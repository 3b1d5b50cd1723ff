"""mdgen_b200 — B200-native (sm_100a) MDGen denoiser / Euler-sampling hot path.

Host side: Python mirror of the reference's `mdgen.wrapper.NewMDGenWrapper` surface.
Device side: libmdgen_b200.so (hand-written CUDA behind the C ABI of include/mdgen_b200.h).
"""
__all__ = ["NewMDGenWrapper", "LatentMDGenModel"]


def __getattr__(name):
    if name == "NewMDGenWrapper":
        from .wrapper import NewMDGenWrapper
        return NewMDGenWrapper
    if name == "LatentMDGenModel":
        from .model import LatentMDGenModel
        return LatentMDGenModel
    raise AttributeError(name)
