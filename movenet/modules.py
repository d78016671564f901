"""``movenet.modules`` -> ``movenet_b200.modules`` (same names as movenet/modules.py:15-142)."""
from movenet_b200.modules import (CausalConv1d, DenseConv, DilatedCausalConv1d, GatedResidualConv1d,  # noqa: F401
                                  ResidualConvStack)
