"""CPU oracle for Path W - the autoregressive WaveRNN vocoder (TEST INFRASTRUCTURE, not product).

Functional fp32 restatement of ``WaveRNN._inference`` (reference ``cube/networks/modules.py:453-503``),
the conditioning it builds from ``UpsampleNetR`` (``:378-389``), ``UpsampleNetI`` (``:346-354``) and the
``_lowres_conv`` stack (``:419-423``), and of ``CubenetVocoder._inference_batch`` /
``_compose_batched_inference`` (``cube/networks/vocoder.py:109-131``).  The per-step random draws of the
output head are INJECTED, in the order the reference consumes them from torch's global RNG
(MOL: ``uniform_`` on [B,1,nr_mix] then on [B,1]; Gaussian: ``randn`` on [B,1,1]).

Pinned: tests/test_oracle.py replays tests/golden/wavernn_{mol,gm}.npz, produced by oracle/make_goldens.py
from the unmodified reference class with seeded weights and ``torch.manual_seed``.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import heads_ref as W


def build_cond(sd: Dict[str, torch.Tensor], mel: torch.Tensor, x_low: Optional[torch.Tensor], upsample: int,
               upsample_low: int) -> torch.Tensor:
    """mel [B,F,80] (time-major, cube/networks/modules.py:455), x_low [B,Tl] -> cond [B,T,80(+21)]."""
    up_mel = mel.permute(0, 2, 1).repeat_interleave(upsample, dim=2).permute(0, 2, 1)            # UpsampleNetR
    if x_low is None:
        return up_mel
    interp = F.interpolate(x_low.unsqueeze(1), upsample_low * x_low.shape[1], mode="linear").permute(0, 2, 1)  # UpsampleNetI
    hid = x_low.unsqueeze(1)
    for i in range(3):
        hid = torch.tanh(F.conv1d(hid, sd[f"_lowres_conv.{i}.conv.weight"], sd[f"_lowres_conv.{i}.conv.bias"], padding=3))
    up_x = hid.repeat_interleave(upsample_low, dim=2).permute(0, 2, 1)
    m = min(up_mel.shape[1], up_x.shape[1], interp.shape[1])
    return torch.cat([up_mel[:, :m], up_x[:, :m], interp[:, :m]], dim=-1)


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRU single step (gate order r, z, n)."""
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    H = h.shape[-1]
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1 - z) * n + z * h


@torch.no_grad()
def wavernn_inference(sd: Dict[str, torch.Tensor], mel: torch.Tensor, x_low: Optional[torch.Tensor], upsample: int,
                      upsample_low: int, output: str, draws: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Returns x [B, T].  draws: 'mol' -> u_mix [T,B,nr_mix], u_x [T,B];  'gm' -> eps [T,B];
    'mulaw'/'raw' -> u [T,B,256] (Gumbel-max form, see heads_ref.categorical_sample_gumbel)."""
    sd = {k: v.float() for k, v in sd.items()}
    cond = build_cond(sd, mel.float(), None if x_low is None else x_low.float(), upsample, upsample_low)
    B, T, _ = cond.shape
    n_layers = len([k for k in sd if k.startswith("_rnns.") and k.endswith("weight_ih_l0")])
    H = sd["_rnns.0.weight_hh_l0"].shape[1]
    hs = [torch.zeros(B, H) for _ in range(n_layers)]
    last = torch.zeros(B, 1)
    out = []
    for t in range(T):
        x = torch.cat([cond[:, t], last], dim=-1)
        for l in range(n_layers):
            p = f"_rnns.{l}."
            hs[l] = gru_cell(x, hs[l], sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"], sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"])
            x = hs[l]
        pre = torch.tanh(x @ sd["_preoutput.linear_layer.weight"].t() + sd["_preoutput.linear_layer.bias"])
        y = (pre @ sd["_output.linear_layer.weight"].t() + sd["_output.linear_layer.bias"]).unsqueeze(1)  # [B,1,S]
        if output == "mol":
            s = W.mol_sample(y, draws["u_mix"][t].unsqueeze(1), draws["u_x"][t].unsqueeze(1))      # [B,1]
        elif output == "gm":
            s = W.gaussian_sample(y, draws["eps"][t].unsqueeze(1))
        elif output == "mulaw":
            s = W.mulaw_decode(W.categorical_sample_gumbel(y, draws["u"][t].unsqueeze(1)))
        elif output == "raw":
            s = W.raw_decode(W.categorical_sample_gumbel(y, draws["u"][t].unsqueeze(1)))
        else:
            raise ValueError(output)
        last = s.reshape(B, 1)
        out.append(last)
    return torch.cat(out, dim=1)


def fold_batch(mel: torch.Tensor, x_low: torch.Tensor, upsample_low: int, num_batches: int = 20):
    """CubenetVocoder._inference_batch (cube/networks/vocoder.py:113-131): one utterance [1,F,80] / [1,Tl] ->
    num_batches chunks with one frame / upsample_low samples of left context (first chunk: mel pad -5, x pad 0)."""
    if mel.shape[1] < num_batches:
        num_batches = mel.shape[1]
    mel = mel[:, : mel.shape[1] // num_batches * num_batches]
    x_low = x_low[:, : x_low.shape[1] // num_batches * num_batches]
    ms = mel.reshape(num_batches, -1, mel.shape[2]).numpy()
    xs = x_low.reshape(num_batches, -1).numpy()
    m = np.ones((ms.shape[0], ms.shape[1] + 1, ms.shape[2])) * -5
    m[:, 1:, :] = ms
    m[1:, 0, :] = ms[:-1, -1, :]
    x = np.zeros((xs.shape[0], xs.shape[1] + upsample_low))
    x[:, upsample_low:] = xs
    x[1:, 0:upsample_low] = xs[:-1, -upsample_low:]
    return torch.tensor(m, dtype=torch.float), torch.tensor(x, dtype=torch.float)


def unfold_batch(batched_x: torch.Tensor, upsample: int) -> torch.Tensor:
    """CubenetVocoder._compose_batched_inference (vocoder.py:109-111)."""
    return batched_x[:, upsample:].reshape(1, -1)


def random_state_dict(H: int = 64, n_layers: int = 2, use_lowres: bool = True, S: int = 30, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    ic = 81 + (21 if use_lowres else 0)
    sd = {}

    def lin(name, o, i, gain=1.0):
        sd[name + ".weight"] = torch.randn(o, i, generator=g) * gain / (i ** 0.5)
        sd[name + ".bias"] = torch.randn(o, generator=g) * 0.05

    if use_lowres:
        c = 1
        for i in range(3):
            sd[f"_lowres_conv.{i}.conv.weight"] = torch.randn(20, c, 7, generator=g) / ((c * 7) ** 0.5)
            sd[f"_lowres_conv.{i}.conv.bias"] = torch.randn(20, generator=g) * 0.05
            c = 20
    lin("_skip.linear_layer", H, ic)
    i = ic
    for l in range(n_layers):
        sd[f"_rnns.{l}.weight_ih_l0"] = torch.randn(3 * H, i, generator=g) / (i ** 0.5)
        sd[f"_rnns.{l}.weight_hh_l0"] = torch.randn(3 * H, H, generator=g) / (H ** 0.5)
        sd[f"_rnns.{l}.bias_ih_l0"] = torch.randn(3 * H, generator=g) * 0.05
        sd[f"_rnns.{l}.bias_hh_l0"] = torch.randn(3 * H, generator=g) * 0.05
        i = H
    lin("_preoutput.linear_layer", 256, H)
    lin("_output.linear_layer", S, 256)
    return sd


@torch.no_grad()
def upsamplenet_forward(sd: Dict[str, torch.Tensor], c: torch.Tensor, scales, kernel_size: int = 3) -> torch.Tensor:
    """Reference ``UpsampleNet.forward`` (cube/networks/modules.py:317-343) restated functionally: three
    ``tanh(Conv1d(k, padding=k // 2))`` (ModuleList indices 0, 2, 4), then per scale s
    ``tanh(ConvTranspose1d(2 s, stride=s, padding=s // 2))`` with weight-norm (g over dim 0 of the [C_in, C_out, K] weight).
    ``sd`` holds the reference module's own state_dict.  Pinned by tests/golden/upsamplenet.npz (reference-run)."""
    x = c.to(torch.float32)
    for i in range(3):
        x = torch.tanh(F.conv1d(x, sd[f"_conv.{2 * i}.weight"], sd[f"_conv.{2 * i}.bias"], padding=kernel_size // 2))
    for n, s in enumerate(scales):
        p = f"_upsample_conv.{2 * n}."
        w = torch._weight_norm(sd[p + "weight_v"], sd[p + "weight_g"], 0) if p + "weight_v" in sd else sd[p + "weight"]
        x = torch.tanh(F.conv_transpose1d(x, w, sd[p + "bias"], stride=s, padding=s // 2))
    return x
