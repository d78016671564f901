"""mu-law companding oracle (test infrastructure, see oracle/__init__.py).

Restates ``torchaudio.functional.mu_law_encoding`` / ``mu_law_decoding``
(torchaudio is a third-party dependency of the reference:
/root/reference/requirements.txt:12; call sites movenet/dataset.py:284,
movenet/callbacks.py:66,73, tests/test_model.py:22,71).  The published formula:

    encode: x_mu = sign(x) * log1p(mu*|x|) / log1p(mu)
            code = int64((x_mu + 1) / 2 * mu + 0.5)         (trunc toward 0)
    decode: x    = code / mu * 2 - 1
            out  = sign(x) * (exp(|x| * log1p(mu)) - 1) / mu

with mu = quantization_channels - 1 held in the dtype of the input.  Every
intermediate is evaluated in the input dtype with torch CPU elementwise
kernels, in the same operation order, which is what makes the codes
bit-identical to torchaudio's on a CPU.
"""
import torch


def mu_law_encode(x: torch.Tensor, quantization_channels: int) -> torch.Tensor:
    if not x.is_floating_point():
        x = x.to(torch.float)
    x = x.detach().cpu()
    mu = torch.tensor(quantization_channels - 1.0, dtype=x.dtype)
    companded = torch.sign(x) * torch.log1p(mu * torch.abs(x)) / torch.log1p(mu)
    return ((companded + 1) / 2 * mu + 0.5).to(torch.int64)


def mu_law_decode(codes: torch.Tensor, quantization_channels: int) -> torch.Tensor:
    codes = codes.detach().cpu()
    if not codes.is_floating_point():
        codes = codes.to(torch.float)
    mu = torch.tensor(quantization_channels - 1.0, dtype=codes.dtype)
    x = (codes / mu) * 2 - 1.0
    return torch.sign(x) * (torch.exp(torch.abs(x) * torch.log1p(mu)) - 1.0) / mu


def one_hot(codes: torch.Tensor, quantization_channels: int) -> torch.Tensor:
    """(B, T) int64 codes -> (B, A, T) fp32 one-hot, as movenet/dataset.py:285-288."""
    b, t = codes.shape
    out = torch.zeros(b, quantization_channels, t, dtype=torch.float32)
    return out.scatter_(1, codes.unsqueeze(1), 1.0)
