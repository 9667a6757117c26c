"""Log-mel spectrogram on the GPU - the feature front-end either side of the vocoders.

Mirrors the reference's two entry points (same argument names and order, same output layout):
  ``mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False)``
        hifigan/meldataset.py:50-74 (called by hifigan/inference.py:26 and cube/networks/cubegan.py:137)
  ``MelVocoder().melspectrogram(y, sample_rate, num_mels, hop_size, use_preemphasis=False)``
        cube/io_utils/vocoder.py:54-62 (called by cube/io_utils/io_vocoder.py:56)
The mel filter bank is host logic in the reference too (``librosa.filters.mel``); ``slaney_mel_basis`` builds the
same matrix with numpy.  STFT, magnitude, filter bank and log run in libcube_vocoder.so (``cube_mel_*``): no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f * 3.0 / 200.0
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * (27.0 / np.log(6.4)), lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((m - 15.0) * (np.log(6.4) / 27.0)), m * 200.0 / 3.0)


def slaney_mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax: float | None = None) -> np.ndarray:
    """``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` with its defaults (Slaney mel scale, area-normalised
    triangles) -> float32 [n_mels, 1 + n_fft // 2]."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    fft_f = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    hz = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    ramps = hz[:, None] - fft_f[None, :]
    width = np.diff(hz)
    lower = -ramps[:-2] / width[:-1, None]
    upper = ramps[2:] / width[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper)) * (2.0 / (hz[2:] - hz[:-2]))[:, None]
    return w.astype(np.float32)


class _MelHandle:
    def __init__(self, cfg: _lib.MelConfig, basis: np.ndarray, device: int, window: np.ndarray | None = None):
        self.ptr = C.c_void_p()
        basis = np.ascontiguousarray(basis, dtype=np.float32)
        wptr = None
        if window is not None:
            window = np.ascontiguousarray(window, dtype=np.float32)
            wptr = window.ctypes.data_as(C.c_void_p)
        check(lib().cube_mel_create(C.byref(self.ptr), C.byref(cfg), wptr, basis.ctypes.data_as(C.c_void_p), device))

    def __del__(self):
        try:
            if self.ptr:
                lib().cube_mel_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class MelSpectrogram:
    """One configured front-end on one device.  ``flavor='hifigan'``: reflect pad (n_fft-hop)/2, eps 1e-9, natural log,
    output [B, num_mels, F].  ``flavor='cube'``: librosa ``center=True`` framing (reflect pad n_fft/2), log10, output
    [B, F, num_mels] (time-major, what WaveRNN / the ClariNet-era tools consume)."""

    def __init__(self, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0.0, fmax=None,
                 flavor="hifigan", preemphasis=0.0, pad_value=None, device=None, mel_basis=None, pad_mode="reflect"):
        if flavor not in ("hifigan", "cube"):
            raise ValueError("flavor must be 'hifigan' or 'cube'")
        if pad_mode not in ("reflect", "constant"):
            # the reference calls librosa.stft with library defaults and does not pin librosa: < 0.10 pads with
            # 'reflect', >= 0.10 with 'constant' zeros (first / last ~2 frames differ); hifigan/meldataset.py:62 is
            # always 'reflect'
            raise ValueError("pad_mode must be 'reflect' or 'constant'")
        self.n_fft, self.num_mels, self.hop_size, self.flavor = int(n_fft), int(num_mels), int(hop_size), flavor
        hif = flavor == "hifigan"
        pad = (self.n_fft - self.hop_size) // 2 if hif else self.n_fft // 2
        floor = 1e-5
        if pad_value is None:   # what a silent frame maps to: the clamp floor
            pad_value = float(np.log(floor)) if hif else -5.0
        self.cfg = _lib.MelConfig(n_fft=self.n_fft, win_size=int(win_size), hop_size=self.hop_size, n_mels=self.num_mels,
                                  pad_left=pad, pad_right=pad, log10_out=0 if hif else 1, layout=0 if hif else 1,
                                  pad_mode=0 if pad_mode == "reflect" else 1,
                                  mag_eps=1e-9 if hif else 0.0, floor_val=floor, pad_value=float(pad_value),
                                  preemph=float(preemphasis))
        self.basis = slaney_mel_basis(sampling_rate, n_fft, num_mels, fmin, fmax) if mel_basis is None else np.asarray(mel_basis)
        if self.basis.shape != (self.num_mels, self.n_fft // 2 + 1):
            raise ValueError(f"mel_basis must be [{self.num_mels}, {self.n_fft // 2 + 1}]")
        self._dev = None if device is None else torch.device(device)
        self._h = None

    def n_frames(self, n_samples: int) -> int:
        c = self.cfg
        if (c.pad_mode == 0 and n_samples <= c.pad_left) or n_samples < 1 or n_samples + c.pad_left + c.pad_right < c.n_fft:
            return 0
        return 1 + (n_samples + c.pad_left + c.pad_right - c.n_fft) // c.hop_size

    def __call__(self, y: torch.Tensor, n_samples=None) -> torch.Tensor:
        """y: CUDA float32 [B, T] (or [T]); n_samples: per-utterance valid lengths of a padded batch."""
        squeeze = y.dim() == 1
        if squeeze:
            y = y[None]
        if y.device.type != "cuda":
            raise _lib.CubeVocError("y must be a CUDA tensor: the mel front-end has no CPU path")
        if self._h is None or self._dev != y.device:
            self._dev = y.device
            self._h = _MelHandle(self.cfg, self.basis, y.device.index if y.device.index is not None else torch.cuda.current_device())
        y = y.to(torch.float32).contiguous()
        B, T = y.shape
        lens = None
        longest = T
        if n_samples is not None:
            lens = (C.c_int32 * B)(*[int(v) for v in n_samples])
            longest = max(int(v) for v in n_samples)
        F = max(1, self.n_frames(longest))
        shape = (B, self.num_mels, F) if self.cfg.layout == 0 else (B, F, self.num_mels)
        out = torch.empty(shape, dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            check(lib().cube_mel_forward(self._h.ptr, C.c_void_p(y.data_ptr()), lens, C.c_void_p(out.data_ptr()), B, T, F,
                                         C.c_void_p(torch.cuda.current_stream(y.device).cuda_stream)))
        return out[0] if squeeze else out


_CACHE: dict = {}


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False):
    """Drop-in for hifigan/meldataset.py:50 on CUDA tensors: y [B, T] in [-1, 1] -> log-mel [B, num_mels, F]."""
    if center:
        raise NotImplementedError("the reference only ever calls mel_spectrogram with center=False")
    key = ("hifigan", n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, str(y.device))
    if key not in _CACHE:
        _CACHE[key] = MelSpectrogram(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, flavor="hifigan")
    return _CACHE[key](y)


class MelVocoder:
    """``cube/io_utils/vocoder.py:MelVocoder`` (the melspectrogram part) with the STFT on the GPU."""

    def melspectrogram(self, y, sample_rate, num_mels, hop_size, use_preemphasis=False, device="cuda", pad_mode="reflect"):
        """y: 1-D numpy array or tensor -> numpy float32 [F, num_mels] (log10 mel, floor 1e-5), like the reference."""
        key = ("cube", sample_rate, num_mels, hop_size, bool(use_preemphasis), str(device), pad_mode)
        if key not in _CACHE:
            _CACHE[key] = MelSpectrogram(1024, num_mels, sample_rate, hop_size, 1024, 0.0, None, flavor="cube",
                                         preemphasis=0.97 if use_preemphasis else 0.0, pad_mode=pad_mode)
        t = torch.as_tensor(np.asarray(y, dtype=np.float32) if not torch.is_tensor(y) else y, dtype=torch.float32).to(device)
        return _CACHE[key](t).cpu().numpy()
