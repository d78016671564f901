"""``movenet.types`` -> ``movenet_b200.types`` (movenet/types.py:4-5)."""
from movenet_b200.types import AudioTensor, VideoTensor  # noqa: F401
