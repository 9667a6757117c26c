"""CPU oracle for Path H - the HiFi-GAN generator (TEST INFRASTRUCTURE, not product).

Functional fp32 restatement of ``hifigan/models.py:Generator.forward`` (reference
``hifigan/models.py:100-116``), ``ResBlock1.forward`` (``:35-42``), ``ResBlock2.forward``
(``:63-68``) and ``hifigan/utils.py:get_padding`` (``:34-35``).  It consumes the reference's
own ``state_dict`` (weight-norm ``weight_g``/``weight_v`` pairs or folded ``weight``) so the same
tensors feed the oracle and the CUDA path.

Pinned: tests/test_oracle.py::test_hifigan_oracle_matches_reference_* compares this against outputs of the unmodified reference
module executed in the build container (tests/golden/hifigan_*.npz, made by
oracle/make_goldens.py); the two agree bit-for-bit on CPU because both end in the same ATen
conv calls.
"""
from __future__ import annotations

import json
from typing import Dict, Optional

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # hifigan/models.py:8

CONFIG_V1 = {  # hifigan/config_v1.json - the config Cubegan hard-codes (cube/networks/cubegan.py:41)
    "resblock": "1",
    "upsample_rates": [5, 3, 4, 4],
    "upsample_kernel_sizes": [16, 16, 4, 4],
    "upsample_initial_channel": 512,
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "num_mels": 80,
}
# data/models/vocoder/neb-noft/config.json - the config of the shipped generator g_00600000
CONFIG_NEB = dict(CONFIG_V1, upsample_rates=[3, 5, 4, 4])


def load_config(path: str) -> dict:
    with open(path) as f:
        return json.load(f)


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """hifigan/utils.py:34-35."""
    return int((kernel_size * dilation - dilation) / 2)


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """w = g * v / ||v|| per dim-0 slice (torch.nn.utils.weight_norm, dim=0), what
    ``Generator.remove_weight_norm`` (hifigan/models.py:118-125) leaves behind.  Uses the very
    ATen primitive the reference's weight_norm hook calls, so folded == unfolded bit-for-bit."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.endswith("weight_g"):
            base = k[: -len("weight_g")]
            out[base + "weight"] = torch._weight_norm(sd[base + "weight_v"], v, 0)
        elif k.endswith("weight_v"):
            continue
        else:
            out[k] = v
    return out


def out_len(cfg: dict, n_frames: int) -> int:
    """ConvTranspose1d length law L_out = (L-1)*u - 2*((k-u)//2) + k (hifigan/models.py:84-88)."""
    L = n_frames
    for u, k in zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"]):
        L = (L - 1) * u - 2 * ((k - u) // 2) + k
    return L


def random_state_dict(cfg: dict, seed: int = 0, std: float = 0.01,
                      g_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded random-init generator weights in the reference's weight-norm (g, v) form with the
    reference's key names (what ``Generator(h).state_dict()`` yields).  ``g_scale`` > 1 makes a
    random net loud enough for the parity check to mean something."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def add(name, shape, norm_dim0=True):
        v = torch.randn(shape, generator=gen) * std
        n = v.reshape(shape[0], -1).norm(dim=1).reshape([shape[0]] + [1] * (len(shape) - 1))
        sd[name + ".weight_g"] = n * g_scale * (0.75 + 0.5 * torch.rand(n.shape, generator=gen))
        sd[name + ".weight_v"] = v

    c0 = cfg["upsample_initial_channel"]
    nm = cfg.get("num_mels", 80)
    sd["conv_pre.bias"] = torch.randn(c0, generator=gen) * 0.1
    add("conv_pre", (c0, nm, 7))
    ch = c0
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        sd[f"ups.{i}.bias"] = torch.randn(ch // 2, generator=gen) * 0.1
        add(f"ups.{i}", (ch, ch // 2, k))
        ch //= 2
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            idx = i * len(cfg["resblock_kernel_sizes"]) + j
            if cfg["resblock"] == "1":
                for m in range(len(rd)):
                    for nm_ in ("convs1", "convs2"):
                        sd[f"resblocks.{idx}.{nm_}.{m}.bias"] = torch.randn(ch, generator=gen) * 0.1
                        add(f"resblocks.{idx}.{nm_}.{m}", (ch, ch, rk))
            else:
                for m in range(len(rd)):
                    sd[f"resblocks.{idx}.convs.{m}.bias"] = torch.randn(ch, generator=gen) * 0.1
                    add(f"resblocks.{idx}.convs.{m}", (ch, ch, rk))
    sd["conv_post.bias"] = torch.randn(1, generator=gen) * 0.1
    add("conv_post", (1, ch, 7))
    return sd


def _resblock1(x, w, prefix, k, dil):
    # hifigan/models.py:35-42
    for m, d in enumerate(dil):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, w[f"{prefix}.convs1.{m}.weight"], w[f"{prefix}.convs1.{m}.bias"],
                      padding=get_padding(k, d), dilation=d)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, w[f"{prefix}.convs2.{m}.weight"], w[f"{prefix}.convs2.{m}.bias"],
                      padding=get_padding(k, 1), dilation=1)
        x = xt + x
    return x


def _resblock2(x, w, prefix, k, dil):
    # hifigan/models.py:63-68
    for m, d in enumerate(dil):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, w[f"{prefix}.convs.{m}.weight"], w[f"{prefix}.convs.{m}.bias"],
                      padding=get_padding(k, d), dilation=d)
        x = xt + x
    return x


@torch.no_grad()
def generator_forward(sd: Dict[str, torch.Tensor], cfg: dict, mel: torch.Tensor,
                      dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """mel [B, num_mels, F] -> wav [B, 1, T].  hifigan/models.py:100-116."""
    w = {k: v.to(dtype) for k, v in fold_weight_norm(sd).items()}
    x = mel.to(dtype)
    nk = len(cfg["resblock_kernel_sizes"])
    rb = _resblock1 if cfg["resblock"] == "1" else _resblock2
    x = F.conv1d(x, w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, w[f"ups.{i}.weight"], w[f"ups.{i}.bias"], stride=u,
                               padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            r = rb(x, w, f"resblocks.{i * nk + j}", cfg["resblock_kernel_sizes"][j],
                   cfg["resblock_dilation_sizes"][j])
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01 - hifigan/models.py:112
    x = F.conv1d(x, w["conv_post.weight"], w["conv_post.bias"], padding=3)
    return torch.tanh(x)


@torch.no_grad()
def generator_forward_ragged(sd, cfg, mel: torch.Tensor, n_frames, dtype=torch.float32):
    """Per-utterance unpadded semantics for a right-padded batch (SURVEY hard-part 6): every
    utterance is run alone at its own length; the result is zero-padded to the batch maximum."""
    B = mel.shape[0]
    T = out_len(cfg, int(mel.shape[2]))
    out = torch.zeros(B, 1, T, dtype=dtype)
    for b in range(B):
        f = int(n_frames[b])
        if f <= 0:
            continue
        y = generator_forward(sd, cfg, mel[b:b + 1, :, :f], dtype)
        out[b, :, : y.shape[2]] = y[0]
    return out


def wav_to_int16(wav: torch.Tensor) -> torch.Tensor:
    """cube/api.py:65 ``np.asarray(audio * 32767, dtype=np.int16)``: fp32 multiply, truncation
    toward zero, no clipping (|tanh| <= 1 so the product is always in range)."""
    return (wav.to(torch.float32) * 32767).to(torch.int16)


def synthetic_mel(B: int, F_: int, seed: int, level: float = 0.0, num_mels: int = 80) -> torch.Tensor:
    """SURVEY 8(d) cfg3 input: natural-log-mel-like smooth field, loud enough (level=0) that the
    shipped generator peaks near 0.9 - a flat randn mel drives it to silence."""
    g = torch.Generator().manual_seed(seed)
    r = 6.0 * torch.randn(B, num_mels, F_ + 8, generator=g)
    sm = F.avg_pool1d(r, 9, stride=1)[:, :, :F_]
    tilt = torch.linspace(0, 4, num_mels)[None, :, None]
    return torch.clamp(0.5 * sm - tilt + level, -11.5, 2.5).contiguous()
