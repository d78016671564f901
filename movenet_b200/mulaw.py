"""mu-law companding on the GPU, bit-exact with torchaudio's CPU functions.

``mu_law_encoding`` / ``mu_law_decoding`` keep torchaudio's signatures (the
reference calls them at movenet/dataset.py:284 and movenet/callbacks.py:66-76).

Encoding is monotone in x, so the code is the number of decision thresholds
<= x.  The thresholds are found once per (channels, dtype) on the host by
bisection over the exact operation sequence of the published formula,
evaluated with the same torch CPU elementwise kernels torchaudio runs; the
CUDA kernel then only compares, so the GPU's own log1p rounding never enters.
This is table construction (A-1 numbers), not a data path: samples never
leave the device.
"""
import ctypes as C
from functools import lru_cache

import torch

from . import _lib


def _encode_formula(x: torch.Tensor, channels: int) -> torch.Tensor:
    mu = torch.tensor(channels - 1.0, dtype=x.dtype)
    y = torch.sign(x) * torch.log1p(mu * torch.abs(x)) / torch.log1p(mu)
    return ((y + 1) / 2 * mu + 0.5).to(torch.int64)


def _to_ordered(x: torch.Tensor) -> torch.Tensor:
    """floats -> integers with the same ordering (so bisection lands on ADJACENT floats)."""
    it = torch.int64 if x.dtype == torch.float64 else torch.int32
    i = x.view(it).to(torch.int64)
    mask = (1 << (63 if x.dtype == torch.float64 else 31)) - 1
    return torch.where(i >= 0, i, -(i & mask))


def _from_ordered(i: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if dtype == torch.float64:
        raw = torch.where(i >= 0, i, (-i) | torch.iinfo(torch.int64).min)
        return raw.view(torch.float64)
    raw = torch.where(i >= 0, i, (-i) | (1 << 31))
    raw = torch.where(raw >= (1 << 31), raw - (1 << 32), raw).to(torch.int32)
    return raw.view(torch.float32)


@lru_cache(maxsize=None)
def _thresholds(channels: int, dtype: torch.dtype) -> torch.Tensor:
    """thr[k] = smallest representable x in [-1, 1] whose code is >= k+1 (k = 0..A-2)."""
    k = torch.arange(1, channels, dtype=torch.int64)
    lo = _to_ordered(torch.full((channels - 1,), -1.0, dtype=dtype))   # code(-1) == 0  <  k
    hi = _to_ordered(torch.full((channels - 1,), 1.0, dtype=dtype))    # code(+1) == A-1 >= k
    for _ in range(66):
        mid = lo + (hi - lo) // 2
        ge = _encode_formula(_from_ordered(mid, dtype), channels) >= k
        hi = torch.where(ge, mid, hi)
        lo = torch.where(ge, lo, mid)
    assert bool((hi - lo == 1).all())
    thr = _from_ordered(hi, dtype)
    assert bool((_encode_formula(thr, channels) >= k).all())
    assert bool((_encode_formula(_from_ordered(lo, dtype), channels) < k).all())
    return thr


@lru_cache(maxsize=None)
def _decode_lut(channels: int) -> torch.Tensor:
    q = torch.arange(channels).to(torch.float)
    mu = torch.tensor(channels - 1.0, dtype=torch.float)
    x = (q / mu) * 2 - 1.0
    return torch.sign(x) * (torch.exp(torch.abs(x) * torch.log1p(mu)) - 1.0) / mu


_DEVICE_TABLES = {}


def _on_device(kind, channels, dtype, device):
    key = (kind, channels, dtype, str(device))
    if key not in _DEVICE_TABLES:
        t = _thresholds(channels, dtype) if kind == "thr" else _decode_lut(channels)
        _DEVICE_TABLES[key] = t.to(device)
    return _DEVICE_TABLES[key]


def _require_cuda(t, who):
    if not t.is_cuda:
        raise RuntimeError(f"{who}: expected a CUDA tensor (movenet_b200 has no CPU path)")


def mu_law_encoding(x: torch.Tensor, quantization_channels: int) -> torch.Tensor:
    """torchaudio.functional.mu_law_encoding on the GPU: float (B...,) -> int64 codes."""
    _require_cuda(x, "mu_law_encoding")
    if not x.is_floating_point():
        x = x.to(torch.float)
    if x.dtype not in (torch.float32, torch.float64):
        raise TypeError("mu_law_encoding: float32 or float64 input expected")
    x = x.contiguous()
    thr = _on_device("thr", quantization_channels, x.dtype, x.device)
    out = torch.empty(x.shape, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("mvn_mulaw_encode", x.data_ptr(), int(x.dtype == torch.float64), thr.data_ptr(),
                  quantization_channels, out.data_ptr(), x.numel(), torch.cuda.current_stream().cuda_stream)
    return out


def mu_law_decoding(x_mu: torch.Tensor, quantization_channels: int) -> torch.Tensor:
    """torchaudio.functional.mu_law_decoding on the GPU: integer codes -> float32 in [-1, 1]."""
    _require_cuda(x_mu, "mu_law_decoding")
    if x_mu.is_floating_point():
        raise TypeError("mu_law_decoding: integer codes expected on the GPU path")
    codes = x_mu.to(torch.int64).contiguous()
    lut = _on_device("lut", quantization_channels, torch.float32, codes.device)
    out = torch.empty(codes.shape, dtype=torch.float32, device=codes.device)
    with torch.cuda.device(codes.device):
        _lib.call("mvn_mulaw_decode", codes.data_ptr(), lut.data_ptr(), quantization_channels, out.data_ptr(),
                  codes.numel(), torch.cuda.current_stream().cuda_stream)
    return out


def one_hot(codes: torch.Tensor, quantization_channels: int) -> torch.Tensor:
    """(B, T) integer codes -> (B, A, T) fp32 one-hot, the dataset's encoding (movenet/dataset.py:285-288)."""
    _require_cuda(codes, "one_hot")
    codes = codes.to(torch.int64).contiguous()
    B, T = codes.shape
    out = torch.empty(B, quantization_channels, T, dtype=torch.float32, device=codes.device)
    with torch.cuda.device(codes.device):
        _lib.call("mvn_one_hot", codes.data_ptr(), out.data_ptr(), B, quantization_channels, T,
                  torch.cuda.current_stream().cuda_stream)
    return out
