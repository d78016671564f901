"""``movenet.wavenet`` -> ``movenet_b200.wavenet`` (same names as movenet/wavenet.py:27-50)."""
from movenet_b200.wavenet import (MAX_AUDIO_FRAMES, MAX_VIDEO_FRAMES, UPSAMPLE_STRIDE, VIDEO_KERNEL_SIZE,  # noqa: F401
                                  WaveNet, upsample_kernel_size_solver)
