"""Layout contract of the hot path (mirrors movenet/types.py:4-5 of the reference).

The reference spells these with ``torchtyping.TensorType``; they are only
annotations, so plain aliases keep the package free of that dependency.
"""
import torch

#: [batch, channels, frames] -- channels-first audio / probabilities / logits
AudioTensor = torch.Tensor
#: [batch, frames, height, width, channels]
VideoTensor = torch.Tensor
