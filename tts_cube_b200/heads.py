"""Output heads of the reference's ``cube/networks/loss.py`` on the GPU (same class / method names).

``encode`` / ``decode`` keep the reference's signatures.  ``sample`` takes the random draws as explicit
arguments (the reference draws them internally from torch's global RNG, which cannot be replayed);
when omitted they are drawn on the device with the reference's distributions.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import check, lib

LOG_SCALE_MIN = float(math.log(1e-14))


def _need_cuda(t: torch.Tensor, name: str):
    if t.device.type != "cuda":
        raise _lib.CubeVocError(f"{name} must be a CUDA tensor (no CPU path)")


def _sp(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


class MULAWOutput:
    """cube/networks/loss.py:218-277."""
    sample_size = 256
    stats = (-0.019, 0.51)

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        _need_cuda(x, "x")
        x = x.to(torch.float32).contiguous()
        q = torch.empty(x.shape, dtype=torch.int64, device=x.device)
        with torch.cuda.device(x.device):
            check(lib().cube_mulaw_encode(_p(x), _p(q), x.numel(), _sp(x)))
        return q

    def decode(self, x_mu: torch.Tensor) -> torch.Tensor:
        _need_cuda(x_mu, "x_mu")
        q = x_mu.to(torch.int64).contiguous()
        x = torch.empty(q.shape, dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            check(lib().cube_mulaw_decode(_p(q), _p(x), q.numel(), _sp(q)))
        return x

    def sample(self, y: torch.Tensor, u: torch.Tensor = None) -> torch.Tensor:
        return self.decode(_categorical(y, u))


class RAWOutput:
    """cube/networks/loss.py:280-307."""
    sample_size = 256
    stats = (-0.019, 0.15)

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        _need_cuda(x, "x")
        x = x.to(torch.float32).contiguous()
        q = torch.empty(x.shape, dtype=torch.int64, device=x.device)
        with torch.cuda.device(x.device):
            check(lib().cube_raw_encode(_p(x), _p(q), x.numel(), _sp(x)))
        return q

    def decode(self, x: torch.Tensor) -> torch.Tensor:
        _need_cuda(x, "x")
        q = x.to(torch.int64).contiguous()
        o = torch.empty(q.shape, dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            check(lib().cube_raw_decode(_p(q), _p(o), q.numel(), _sp(q)))
        return o

    def sample(self, y: torch.Tensor, u: torch.Tensor = None) -> torch.Tensor:
        return self.decode(_categorical(y, u))


def _categorical(y: torch.Tensor, u: torch.Tensor = None) -> torch.Tensor:
    _need_cuda(y, "y")
    y = y.to(torch.float32).contiguous()
    Cn = y.shape[-1]
    if u is None:
        # full-range uniforms: truncating them to (1e-5, 1 - 1e-5) caps the Gumbel noise at 11.5 and measurably starves the
        # rare classes (chi-square test in tests/test_gpu_parity.py); u = 0 maps to -inf (never picked), as it should
        u = torch.rand_like(y)
    u = u.to(y.device, torch.float32).contiguous()
    idx = torch.empty(y.shape[:-1], dtype=torch.int64, device=y.device)
    with torch.cuda.device(y.device):
        check(lib().cube_categorical_sample(_p(y), _p(u), _p(idx), idx.numel(), Cn, _sp(y)))
    return idx


class MOLOutput:
    """cube/networks/loss.py:109-215 (sample only; the loss is training-side)."""
    sample_size = 30
    stats = (6e-6, 0.15)

    def sample(self, y: torch.Tensor, log_scale_min: float = None, temperature: float = 1.0,
               u_mix: torch.Tensor = None, u_x: torch.Tensor = None) -> torch.Tensor:
        _need_cuda(y, "y")
        if log_scale_min is None:
            log_scale_min = LOG_SCALE_MIN
        assert y.shape[2] % 3 == 0
        nr_mix = y.shape[2] // 3
        y = y.to(torch.float32).contiguous()
        if u_mix is None:
            u_mix = torch.empty(y.shape[0], y.shape[1], nr_mix, device=y.device).uniform_(1e-5, 1 - 1e-5)
        if u_x is None:
            u_x = torch.empty(y.shape[0], y.shape[1], device=y.device).uniform_(1e-5, 1.0 - 1e-5)
        u_mix = u_mix.to(y.device, torch.float32).contiguous()
        u_x = u_x.to(y.device, torch.float32).contiguous()
        x = torch.empty(y.shape[0], y.shape[1], dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            check(lib().cube_mol_sample(_p(y), _p(u_mix), _p(u_x), _p(x), x.numel(), nr_mix,
                                        float(log_scale_min), float(temperature), _sp(y)))
        return x

    def encode(self, x):
        return x

    def decode(self, x):
        return x


class GaussianOutput:
    """cube/networks/loss.py:35-66 (sample only)."""
    sample_size = 2
    stats = (6e-6, 0.15)

    def sample(self, y_hat: torch.Tensor, temperature: float = 1.0, eps: torch.Tensor = None) -> torch.Tensor:
        _need_cuda(y_hat, "y_hat")
        y = y_hat.to(torch.float32).contiguous()
        if eps is None:
            eps = torch.randn(y.shape[0], y.shape[1], device=y.device)
        eps = eps.to(y.device, torch.float32).contiguous()
        x = torch.empty(y.shape[0], y.shape[1], dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            check(lib().cube_gaussian_sample(_p(y), _p(eps), _p(x), x.numel(), _sp(y)))
        return x

    def encode(self, x):
        return x

    def decode(self, x):
        return x


def wav_to_int16(wav: torch.Tensor) -> torch.Tensor:
    """cube/api.py:65 on the device."""
    _need_cuda(wav, "wav")
    w = wav.to(torch.float32).contiguous()
    o = torch.empty(w.shape, dtype=torch.int16, device=w.device)
    with torch.cuda.device(w.device):
        check(lib().cube_wav_to_int16(_p(w), _p(o), w.numel(), _sp(w)))
    return o
