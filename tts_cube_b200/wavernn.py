"""``WaveRNNVocoder`` - the reference's autoregressive ``WaveRNN`` (cube/networks/modules.py:392-503) with its whole
sample loop inside one persistent cooperative CUDA kernel, and ``CubenetVocoder`` (cube/networks/vocoder.py:33-131):
the low-rate + high-rate pair with the time-into-batch fold (W5).

Same constructor arguments and ``state_dict`` keys as the reference classes; ``forward({'mel': [B,F,80],
'x_low': [B,Tl]})`` returns ``np.float32 [B,T,1]`` like the reference.  The sampling head's random numbers may be
passed in (``draws``) so that a run can be replayed; otherwise they are drawn on the device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional

import numpy as np
import torch

from . import _lib
from ._lib import VocConfig, check, lib
from .generator import _Handle

_SAMPLE_SIZE = {"mol": 30, "gm": 2, "mulaw": 256, "raw": 256}


class WaveRNNVocoder(torch.nn.Module):
    def __init__(self, num_layers: int = 2, layer_size: int = 512, upsample=100, upsample_low=10, use_lowres=True,
                 learning_rate=1e-4, output="mol"):
        super().__init__()
        lib()
        if output not in _lib.HEADS:
            raise _lib.CubeVocError(
                f"output '{output}' is not supported on the GPU path (mol, gm, mulaw, raw).  'beta' (reference cube/networks/loss.py:69-106) "
                "samples through torch.distributions.Beta, whose internal gamma rejection sampler cannot be replayed from injected draws; "
                "no shipped model uses it")
        self._output = output
        self._upsample, self._upsample_low, self._use_lowres = int(upsample), int(upsample_low), bool(use_lowres)
        cfg = VocConfig()
        cfg.arch = _lib.ARCH_WAVERNN
        cfg.num_mels = 80
        cfg.wrnn_layers, cfg.wrnn_size = int(num_layers), int(layer_size)
        cfg.wrnn_upsample, cfg.wrnn_upsample_low = int(upsample), int(upsample_low)
        cfg.wrnn_use_lowres = int(bool(use_lowres))
        cfg.wrnn_head = _lib.HEADS[output]
        self._cfg = cfg
        self._sd: Dict[str, torch.Tensor] = {}
        self._handle: Optional[_Handle] = None
        self.register_buffer("_device_tracker", torch.zeros(1), persistent=False)

    def load_state_dict(self, state_dict: Mapping[str, torch.Tensor], strict: bool = True, assign: bool = False):
        self._sd = {k: v.detach().to("cpu", torch.float32).clone() for k, v in state_dict.items()}
        if self._handle is not None:
            self._handle.close()
            self._handle = None
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def state_dict(self, *a, **k):
        return dict(self._sd)

    @property
    def device(self) -> torch.device:
        return self._device_tracker.device

    def _apply(self, fn, *a, **k):
        before = self._device_tracker.device
        r = super()._apply(fn, *a, **k)
        if self._device_tracker.device != before and self._handle is not None:
            self._handle.close()
            self._handle = None
        return r

    def _ensure(self) -> _Handle:
        if self._handle is None:
            dev = self._device_tracker.device
            if dev.type != "cuda":
                raise _lib.CubeVocError("WaveRNNVocoder must be moved to a CUDA device; there is no CPU path")
            if not self._sd:
                raise _lib.CubeVocError("load_state_dict() must be called before forward()")
            dev = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
            hd = _Handle(self._cfg, dev)
            hd.load(self._sd)
            hd.finalize()
            self._handle = hd
        return self._handle

    @property
    def sample_size(self) -> int:
        return _SAMPLE_SIZE[self._output]

    def draws_shape(self, B: int, T: int):
        k = {"mol": 11, "gm": 1}.get(self._output, 256)
        return (T, B, k)

    def make_draws(self, B: int, T: int, device) -> torch.Tensor:
        if self._output == "gm":
            return torch.randn(T, B, 1, device=device)
        return torch.empty(*self.draws_shape(B, T), device=device).uniform_(1e-5, 1.0 - 1e-5)

    @torch.no_grad()
    def inference(self, mel: torch.Tensor, x_low: Optional[torch.Tensor] = None, draws: Optional[torch.Tensor] = None) -> torch.Tensor:
        """mel [B,F,80] cuda, x_low [B,Tl] cuda (when use_lowres) -> x [B,T] cuda."""
        hd = self._ensure()
        if mel.device.type != "cuda":
            raise _lib.CubeVocError("mel must live on a CUDA device (no CPU path)")
        mel = mel.to(torch.float32).contiguous()
        B, F, _ = mel.shape
        Tl = 0
        if self._use_lowres:
            if x_low is None:
                raise _lib.CubeVocError("x_low is required when use_lowres=True")
            x_low = x_low.to(mel.device, torch.float32).contiguous()
            Tl = x_low.shape[1]
        T = int(lib().cube_wavernn_out_len(hd.ptr, F, Tl))
        if draws is None:
            draws = self.make_draws(B, T, mel.device)
        draws = draws.to(mel.device, torch.float32).contiguous()
        if tuple(draws.shape) != self.draws_shape(B, T):
            raise _lib.CubeVocError(f"draws must have shape {self.draws_shape(B, T)}, got {tuple(draws.shape)}")
        out = torch.empty(B, T, device=mel.device, dtype=torch.float32)
        bmax = int(lib().cube_wavernn_max_batch(hd.ptr))
        stream = C.c_void_p(torch.cuda.current_stream(mel.device).cuda_stream)
        with torch.cuda.device(mel.device):
            for b0 in range(0, B, bmax):          # shared memory bounds the lock-step batch (20 at layer_size 512)
                b1 = min(B, b0 + bmax)
                o = torch.empty(b1 - b0, T, device=mel.device, dtype=torch.float32)
                d = draws[:, b0:b1].contiguous()
                xl = x_low[b0:b1].contiguous() if self._use_lowres else None
                check(lib().cube_wavernn_forward(hd.ptr, C.c_void_p(mel[b0:b1].contiguous().data_ptr()),
                                                 C.c_void_p(xl.data_ptr()) if xl is not None else None,
                                                 C.c_void_p(d.data_ptr()), C.c_void_p(o.data_ptr()), b1 - b0, F, Tl, stream))
                out[b0:b1] = o
        return out

    def forward(self, X, draws: Optional[torch.Tensor] = None):
        """Reference call shape: ``wavernn({'mel': ..., 'x_low': ...}) -> np.float32 [B, T, 1]`` (modules.py:447-503)."""
        if "x" in X:
            raise _lib.CubeVocError("WaveRNNVocoder is inference-only (training goes through the reference WaveRNN)")
        dev = self._device_tracker.device
        x = self.inference(X["mel"].to(dev), X["x_low"].to(dev) if "x_low" in X and self._use_lowres else None, draws)
        return x.unsqueeze(2).cpu().numpy()


def fold_batch(mel: torch.Tensor, x_low: torch.Tensor, upsample_low: int, num_batches: int = 20):
    """CubenetVocoder._inference_batch (cube/networks/vocoder.py:113-131) on tensors (any device)."""
    if mel.shape[1] < num_batches:
        num_batches = mel.shape[1]
    mel = mel[:, : mel.shape[1] // num_batches * num_batches]
    x_low = x_low[:, : x_low.shape[1] // num_batches * num_batches]
    ms = mel.reshape(num_batches, -1, mel.shape[2])
    xs = x_low.reshape(num_batches, -1)
    m = torch.full((ms.shape[0], ms.shape[1] + 1, ms.shape[2]), -5.0, dtype=torch.float32, device=mel.device)
    m[:, 1:, :] = ms
    m[1:, 0, :] = ms[:-1, -1, :]
    x = torch.zeros(xs.shape[0], xs.shape[1] + upsample_low, dtype=torch.float32, device=mel.device)
    x[:, upsample_low:] = xs
    x[1:, :upsample_low] = xs[:-1, -upsample_low:]
    return m, x


def unfold_batch(batched_x: torch.Tensor, upsample: int) -> torch.Tensor:
    """CubenetVocoder._compose_batched_inference (vocoder.py:109-111)."""
    return batched_x[:, upsample:].reshape(1, -1)


class CubenetVocoder(torch.nn.Module):
    """cube/networks/vocoder.py:33-131 (inference): low-rate WaveRNN, then the high-rate WaveRNN over 20 folded chunks."""

    def __init__(self, num_layers_lr=2, layer_size_lr=512, num_layers_hr=2, layer_size_hr=512, upsample=100, upsample_low=10,
                 learning_rate=1e-4, output="mol"):
        super().__init__()
        self._wavernn_hr = WaveRNNVocoder(num_layers_hr, layer_size_hr, upsample, upsample_low, True, learning_rate, output)
        self._wavernn_lr = WaveRNNVocoder(num_layers_lr, layer_size_lr, upsample // upsample_low, upsample_low, False, learning_rate, output)
        self._upsample, self._upsample_low = upsample, upsample_low

    def load_state_dict(self, state_dict, strict=True, assign=False):
        self._wavernn_hr.load_state_dict({k[len("_wavernn_hr."):]: v for k, v in state_dict.items() if k.startswith("_wavernn_hr.")})
        self._wavernn_lr.load_state_dict({k[len("_wavernn_lr."):]: v for k, v in state_dict.items() if k.startswith("_wavernn_lr.")})
        return torch.nn.modules.module._IncompatibleKeys([], [])

    @torch.no_grad()
    def forward(self, X, draws_lr=None, draws_hr=None):
        dev = self._wavernn_hr._device_tracker.device
        mel = X["mel"].to(dev)
        x_lr = self._wavernn_lr.inference(mel, None, draws_lr)                       # [1, T/upsample_low]
        m, xl = fold_batch(mel, x_lr, self._upsample_low, num_batches=20)
        x_hr = unfold_batch(self._wavernn_hr.inference(m, xl, draws_hr), self._upsample)
        return x_lr.unsqueeze(2).cpu().numpy(), x_hr.cpu()


class UpsampleNet(torch.nn.Module):
    """``cube/networks/modules.py:317-343``: 3 x (Conv1d(k) + tanh), then per scale s a weight-normed
    ConvTranspose1d(2s, stride s, padding s // 2) + tanh - the learned mel upsampler of the WaveRNN family (the reference
    keeps it as an alternative to ``UpsampleNetR`` / ``UpsampleNetI``, modules.py:411).  Same constructor arguments, same
    state-dict keys (``_conv.{0,2,4}.*``, ``_upsample_conv.{0,2,..}.weight_g/weight_v/bias``), ``forward(c [B, C, T']) ->
    [B, out_channels, T' * prod(scales)]``; ``n_frames`` masks a right-padded batch.  Even scales only (an odd scale
    changes the reference's own length law to T' * s + 1)."""

    def __init__(self, upsample_scales=(2, 2, 4), in_channels=80, out_channels=80, kernel_size=3):
        super().__init__()
        lib()
        cfg = VocConfig()
        cfg.arch = _lib.ARCH_UPSAMPLENET
        cfg.num_mels, cfg.res_channels, cfg.kernel_size = int(in_channels), int(out_channels), int(kernel_size)
        scales = [int(s) for s in upsample_scales]
        if not 1 <= len(scales) <= 4:
            raise _lib.CubeVocError("UpsampleNet takes 1..4 upsample scales")
        cfg.n_upsample = len(scales)
        for i, s in enumerate(scales):
            cfg.upsample_scales[i] = s
        self._cfg, self._scales, self._out = cfg, scales, int(out_channels)
        self._sd: Dict[str, torch.Tensor] = {}
        self._handle: Optional[_Handle] = None
        self.register_buffer("_device_tracker", torch.zeros(1), persistent=False)

    @property
    def device(self) -> torch.device:
        return self._device_tracker.device

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):  # noqa: D102
        self._sd = {k: v.detach().to("cpu", torch.float32).clone() for k, v in state_dict.items()}
        if self._handle is not None:
            self._handle.close()
            self._handle = None
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def state_dict(self, *a, **k):  # noqa: D102
        return dict(self._sd)

    def _apply(self, fn, *a, **k):
        before = self._device_tracker.device
        r = super()._apply(fn, *a, **k)
        if self._device_tracker.device != before and self._handle is not None:
            self._handle.close()
            self._handle = None
        return r

    def _ensure(self) -> _Handle:
        if self._handle is None:
            dev = self.device
            if dev.type != "cuda":
                raise _lib.CubeVocError("UpsampleNet must be moved to a CUDA device; there is no CPU path")
            if not self._sd:
                raise _lib.CubeVocError("load_state_dict() must be called before forward()")
            dev = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
            hd = _Handle(self._cfg, dev)
            hd.load(self._sd)
            hd.finalize()
            self._handle = hd
        return self._handle

    def out_len(self, n_frames: int) -> int:
        t = int(n_frames)
        for s in self._scales:
            t *= s
        return t

    def forward(self, c: torch.Tensor, n_frames=None) -> torch.Tensor:
        import ctypes as C
        hd = self._ensure()
        if c.device.type != "cuda" or c.dtype != torch.float32 or c.dim() != 3:
            raise _lib.CubeVocError(f"c must be a CUDA float32 [B, C, T'] tensor, got {c.dtype} {tuple(c.shape)} on {c.device}")
        hd._check_shapes(c, n_frames)
        if c.device != hd.device:
            raise _lib.CubeVocError(f"c lives on {c.device} but this module's weights live on {hd.device}")
        c = c.contiguous()
        B, _, F = c.shape
        out = torch.empty(B, self._out, self.out_len(F), device=c.device, dtype=torch.float32)
        nf = (C.c_int32 * B)(*[int(v) for v in n_frames]) if n_frames is not None else None
        with torch.cuda.device(c.device):
            _lib.check(lib().cube_voc_forward(hd.ptr, C.c_void_p(c.data_ptr()), nf, None, C.c_void_p(out.data_ptr()), None, B, F,
                                              C.c_void_p(torch.cuda.current_stream(c.device).cuda_stream)))
        return out
