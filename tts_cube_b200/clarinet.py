"""``ParallelWaveNetVocoder`` - ClariNet IAF student (4 flows, [6,6,6,24] gated residual blocks) with
the teacher's ``UpsampleNet2`` mel upsampler (reference cube/networks/modules.py:357-375), fed from
the two checkpoints the reference ships (data/models/pnn_vocoder.network, nn_vocoder.network).

``forward(mel[B,80,F], z[B,1,256F]) -> wav[B,1,256F]``; z ~ N(0,1) is an input so runs are
reproducible sample for sample.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Sequence

import torch

from . import _lib
from ._lib import VocConfig, lib
from .generator import _Handle


def student_config(student_sd: Mapping[str, torch.Tensor], math=None,
                   dilation_base: int = 3, dilation_cycle: int = 6,
                   upsample_scales: Sequence[int] = (16, 16)) -> VocConfig:
    nb: Dict[int, int] = {}
    for k in student_sd:
        if k.startswith("iafs.") and ".res_blocks." in k:
            p = k.split(".")
            nb[int(p[1])] = max(nb.get(int(p[1]), 0), int(p[3]) + 1)
    if not nb:
        raise _lib.CubeVocError("state_dict has no 'iafs.N.res_blocks.M.*' keys")
    w = student_sd["iafs.0.res_blocks.0.filter_conv.conv.weight_v"] if "iafs.0.res_blocks.0.filter_conv.conv.weight_v" in student_sd \
        else student_sd["iafs.0.res_blocks.0.filter_conv.conv.weight"]
    wc = student_sd.get("iafs.0.res_blocks.0.filter_conv_c.weight_v", student_sd.get("iafs.0.res_blocks.0.filter_conv_c.weight"))
    wf = student_sd.get("iafs.0.front_conv.0.conv.weight_v", student_sd.get("iafs.0.front_conv.0.conv.weight"))
    if math is None:  # auto: tensor cores for the shipped 128/256/128 geometry, else fp32 kernels
        math = _lib.MATH_TC_SPLIT16 if (int(w.shape[1]) == 128 and int(w.shape[0]) % 128 == 0) else _lib.MATH_FP32_SIMT
    cfg = VocConfig()
    cfg.arch = _lib.ARCH_PWN_STUDENT
    cfg.math = math
    cfg.num_mels = int(wc.shape[1])
    cfg.n_flows = len(nb)
    for f in range(len(nb)):
        cfg.flow_blocks[f] = nb[f]
    cfg.gate_channels, cfg.res_channels, cfg.kernel_size = int(w.shape[0]), int(w.shape[1]), int(w.shape[2])
    cfg.skip_channels = cfg.res_channels
    cfg.front_kernel = int(wf.shape[2])
    cfg.dilation_base, cfg.dilation_cycle = dilation_base, dilation_cycle
    cfg.n_upsample = len(upsample_scales)
    for i, s in enumerate(upsample_scales):
        cfg.upsample_scales[i] = int(s)
    return cfg


class ParallelWaveNetVocoder(torch.nn.Module):
    def __init__(self, student_sd: Mapping[str, torch.Tensor], teacher_sd: Mapping[str, torch.Tensor],
                 math=None, dilation_base: int = 3, dilation_cycle: int = 6):
        super().__init__()
        lib()
        self._cfg = student_config(student_sd, math, dilation_base, dilation_cycle)
        self._sd = {k: v.detach().to("cpu", torch.float32).clone() for k, v in student_sd.items()}
        # only the upsampler of the teacher is used at synthesis time
        self._sd.update({k: v.detach().to("cpu", torch.float32).clone() for k, v in teacher_sd.items()
                         if k.startswith("upsample_conv.")})
        self._handle: Optional[_Handle] = None
        self.register_buffer("_device_tracker", torch.zeros(1), persistent=False)

    @property
    def device(self) -> torch.device:
        return self._device_tracker.device

    @property
    def hop(self) -> int:
        h = 1
        for i in range(self._cfg.n_upsample):
            h *= self._cfg.upsample_scales[i]
        return h

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
                raise _lib.CubeVocError("ParallelWaveNetVocoder must be moved to a CUDA device; there is no CPU path")
            dev = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
            hd = _Handle(self._cfg, dev)
            hd.load(self._sd)
            hd.finalize()
            self._handle = hd
        return self._handle

    def out_len(self, n_frames: int) -> int:
        return int(n_frames) * self.hop

    def forward(self, mel: torch.Tensor, z: torch.Tensor, n_frames: Optional[Sequence[int]] = None) -> torch.Tensor:
        if torch.is_grad_enabled() and (mel.requires_grad or z.requires_grad):
            raise _lib.CubeVocError("ParallelWaveNetVocoder is inference-only")
        return self._ensure().forward(mel, n_frames, z)

    def forward_host(self, mel: torch.Tensor, z: torch.Tensor, n_frames=None, out: Optional[torch.Tensor] = None,
                     int16: bool = False) -> torch.Tensor:
        hd = self._ensure()
        B, _, F = mel.shape
        if out is None:
            out = torch.empty(B, F * self.hop, dtype=torch.int16 if int16 else torch.float32).pin_memory()
        return hd.forward_host(mel.contiguous(), n_frames, z.contiguous(), out)

    def conditioning(self, B: int, T: int) -> torch.Tensor:
        """c_up [B,80,T] of the last forward (validation tap)."""
        import ctypes as C
        hd = self._ensure()
        out = torch.empty(B, self._cfg.num_mels, T, device=hd.device, dtype=torch.float32)
        _lib.check(lib().cube_voc_get_cond(hd.ptr, C.c_void_p(out.data_ptr()), B, T,
                                           C.c_void_p(torch.cuda.current_stream(hd.device).cuda_stream)))
        return out
