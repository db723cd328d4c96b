"""efgh_b200 - B200-native permutohedral-lattice / bilateral-convolution hot path of EFGHNet.

Public surface mirrors the reference modules it replaces:
    efgh_b200.GenerateData       <- reference nets/generate_data.py  GenerateData
    efgh_b200.BilateralConvFlex  <- reference nets/bilateralNN.py    BilateralConvFlex
Both call hand-written sm_100a CUDA through the C ABI in include/efgh_b200.h
(efgh_b200/lib/libefgh_b200.so); there is no CPU or PyTorch fallback.
"""
from . import synth  # noqa: F401


def __getattr__(name):
    # torch-dependent modules load lazily so that `import efgh_b200.synth` stays light
    if name == "GenerateData":
        from .generate_data import GenerateData
        return GenerateData
    if name == "BilateralConvFlex":
        from .bilateralNN import BilateralConvFlex
        return BilateralConvFlex
    if name == "Enet":
        from .enet import Enet
        return Enet
    if name == "ScanPipeline":
        from .pipeline import ScanPipeline
        return ScanPipeline
    raise AttributeError(name)
