"""movenet_b200: the WaveNet hot path of cosmicBboy/movenet, hand-written for B200 (sm_100a).

    from movenet_b200 import WaveNet          # drop-in for movenet.wavenet.WaveNet
    from movenet_b200 import mu_law_encoding, mu_law_decoding

The package is a thin Python/PyTorch host layer over a C-ABI CUDA library
(include/movenet_b200.h, built by ``python -m movenet_b200.build``).
"""
from .wavenet import (MAX_AUDIO_FRAMES, MAX_VIDEO_FRAMES, UPSAMPLE_STRIDE, VIDEO_KERNEL_SIZE, WaveNet,
                      upsample_kernel_size_solver)
from .modules import CausalConv1d, DenseConv, DilatedCausalConv1d, GatedResidualConv1d, ResidualConvStack
from .mulaw import mu_law_decoding, mu_law_encoding, one_hot
from .loss import softmax_cross_entropy
from . import optim

__all__ = ["WaveNet", "MAX_AUDIO_FRAMES", "MAX_VIDEO_FRAMES", "VIDEO_KERNEL_SIZE", "UPSAMPLE_STRIDE",
           "upsample_kernel_size_solver", "CausalConv1d", "DilatedCausalConv1d", "GatedResidualConv1d",
           "ResidualConvStack", "DenseConv", "mu_law_encoding", "mu_law_decoding", "one_hot", "softmax_cross_entropy", "optim"]
