"""CPU oracle for the Path-W output heads (TEST INFRASTRUCTURE, not product).

Restates, in the reference's float32 torch op order, the ``encode`` / ``decode`` / ``sample`` members of
``MULAWOutput`` (reference ``cube/networks/loss.py:218-277``), ``RAWOutput`` (``:280-307``),
``MOLOutput.sample`` (``:163-201``) and ``GaussianOutput.sample`` (``:50-52``).  Random draws are
INJECTED (the reference draws them inside ``sample``) so that a GPU kernel can be compared sample
for sample: the draw order is the reference's (``uniform_`` on [B,T,nr_mix] then on [B,T]).

Pinned: tests/test_oracle_heads.py checks every function against the unmodified reference classes
executed in the build container (tests/golden/heads.npz made by oracle/make_goldens.py), including
the full 256-entry decode table and the 255 float32 bin edges of ``encode``.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

QC = 256
LOG_SCALE_MIN = float(np.log(1e-14))  # cube/networks/loss.py:174-175


# ---- MULAWOutput ------------------------------------------------------------------------------
def mulaw_encode(x: torch.Tensor) -> torch.Tensor:
    """float32 [*] in [-1,1] -> int64 codes 0..255.  cube/networks/loss.py:236-254 (torch branch)."""
    x = x.to(torch.float32)
    mu = torch.FloatTensor([QC - 1])
    x_mu = torch.sign(x) * torch.log1p(mu * torch.abs(x)) / torch.log1p(mu)
    x_mu = ((x_mu + 1) / 2 * mu + 0.5).long()
    return torch.clip(x_mu, 0, QC - 1)


def mulaw_decode(q: torch.Tensor) -> torch.Tensor:
    """int64 codes -> float32.  cube/networks/loss.py:256-269 (torch branch)."""
    x_mu = q.float()
    mu = torch.FloatTensor([QC - 1.0])
    x = (x_mu / mu) * 2 - 1.0
    return torch.sign(x) * (torch.exp(torch.abs(x) * torch.log1p(mu)) - 1.0) / mu


def mulaw_decode_table() -> torch.Tensor:
    return mulaw_decode(torch.arange(QC))


def mulaw_encode_edges(encode=mulaw_encode) -> np.ndarray:
    """The 255 float32 bin edges e_k = min{x : encode(x) >= k}, k=1..255, found by bisection over
    the float32 ordering.  encode is monotone, so ``code(x) = #{k : e_k <= x}`` reproduces it bit
    for bit - this table is what the CUDA kernel searches (DESIGN.md 'heads')."""
    def f2o(f):  # float32 -> order-preserving uint32 key
        u = np.asarray(f, dtype=np.float32).view(np.uint32).astype(np.uint64)
        return np.where(u & 0x80000000, 0xFFFFFFFF - u, u + 0x80000000)

    def o2f(o):
        o = np.asarray(o, dtype=np.uint64)
        u = np.where(o >= 0x80000000, o - 0x80000000, 0xFFFFFFFF - o).astype(np.uint32)
        return u.view(np.float32)

    ks = np.arange(1, QC)
    lo = np.full(ks.shape, f2o(np.float32(-1.0)), dtype=np.uint64)   # encode(lo) < k
    hi = np.full(ks.shape, f2o(np.float32(1.0)), dtype=np.uint64)    # encode(hi) >= k
    assert int(encode(torch.tensor([-1.0]))[0]) == 0 and int(encode(torch.tensor([1.0]))[0]) == QC - 1
    while np.any(hi - lo > 1):
        mid = (lo + hi) // 2
        c = encode(torch.from_numpy(o2f(mid).copy())).numpy()
        ge = c >= ks
        hi = np.where(ge, mid, hi)
        lo = np.where(ge, lo, mid)
    return o2f(hi).astype(np.float32)


def mulaw_encode_by_edges(x: torch.Tensor, edges: np.ndarray) -> torch.Tensor:
    xs = x.to(torch.float32).numpy()
    q = np.searchsorted(edges, xs, side="right").astype(np.int64)
    # x > 1 keeps climbing past the last edge: clip as the reference does; NaN -> reference yields
    # an implementation-defined cast, the contract excludes it
    return torch.from_numpy(np.clip(q, 0, QC - 1))


# ---- RAWOutput --------------------------------------------------------------------------------
def raw_encode(x: torch.Tensor) -> torch.Tensor:
    """cube/networks/loss.py:293-295."""
    return torch.clip(((x.to(torch.float32) + 1.0) / 2) * 255, 0, 255).long()


def raw_decode(q: torch.Tensor) -> torch.Tensor:
    """cube/networks/loss.py:297-299 (int64 / 255 is a true division -> float32)."""
    return ((q / 255) - 0.5) * 2


# ---- MOLOutput.sample -------------------------------------------------------------------------
def mol_sample(y: torch.Tensor, u_mix: torch.Tensor, u_x: torch.Tensor,
               log_scale_min: float = LOG_SCALE_MIN, temperature: float = 1.0) -> torch.Tensor:
    """y [B,T,3*nr_mix]; u_mix [B,T,nr_mix], u_x [B,T] ~ U(1e-5, 1-1e-5) (injected)
    -> x [B,T] in [-1,1].  cube/networks/loss.py:176-199."""
    nr_mix = y.shape[2] // 3
    logit_probs = y[:, :, :nr_mix]
    temp = u_mix * temperature
    temp = logit_probs - torch.log(-torch.log(temp))
    _, argmax = temp.max(dim=-1)
    one_hot = F.one_hot(argmax, nr_mix).float()
    means = torch.sum(y[:, :, nr_mix:2 * nr_mix] * one_hot, dim=-1)
    log_scales = torch.clamp(torch.sum(y[:, :, 2 * nr_mix:3 * nr_mix] * one_hot, dim=-1), min=log_scale_min)
    x = means + torch.exp(log_scales) * (torch.log(u_x) - torch.log(1.0 - u_x))
    return torch.clamp(torch.clamp(x, min=-1.0), max=1.0)


def mol_argmax_margin(y: torch.Tensor, u_mix: torch.Tensor) -> torch.Tensor:
    """float64 gap between the best and second-best Gumbel score: positions where it is < 1e-5 are
    legitimate float32 near-ties (the test lists them instead of demanding equality there)."""
    nr_mix = y.shape[2] // 3
    s = y[:, :, :nr_mix].double() - torch.log(-torch.log(u_mix.double()))
    top = torch.topk(s, 2, dim=-1).values
    return top[..., 0] - top[..., 1]


# ---- GaussianOutput.sample --------------------------------------------------------------------
def gaussian_sample(y_hat: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """y_hat [B,T,2] (mean, log_std); eps [B,T] ~ N(0,1) (injected) -> [B,T].
    cube/networks/loss.py:50-52: mean + (eps*0.8) * exp(log_std)."""
    z = eps * 0.8
    return y_hat[:, :, 0] + z * torch.exp(y_hat[:, :, 1])


# ---- Categorical sampling (MULAWOutput.sample / RAWOutput.sample) ------------------------------
def categorical_sample_gumbel(logits: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """The reference draws ``Categorical(logits=y).sample()`` (cube/networks/loss.py:227-229,
    288-290) from torch's global RNG, which no other implementation can replay.  With injected
    uniforms u [*, 256] ~ U(1e-5, 1-1e-5) we use the Gumbel-max form the reference itself uses for
    the mixture pick in MOLOutput.sample (loss.py:182-184): argmax(logits - log(-log u)) is an
    exact sample of the same categorical distribution."""
    s = logits - torch.log(-torch.log(u))
    return s.argmax(dim=-1)


def categorical_margin(logits: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    s = logits.double() - torch.log(-torch.log(u.double()))
    top = torch.topk(s, 2, dim=-1).values
    return top[..., 0] - top[..., 1]


# ---- WaveRNN-path upsamplers (cube/networks/modules.py:346-354, 378-389) -----------------------
def upsample_repeat(c: torch.Tensor, r: int) -> torch.Tensor:
    """UpsampleNetR: [B,C,F] -> [B,C,F*r], each frame repeated r times."""
    return c.repeat_interleave(r, dim=2)


def upsample_linear(c: torch.Tensor, r: int) -> torch.Tensor:
    """UpsampleNetI: F.interpolate(mode='linear', align_corners=False)."""
    return F.interpolate(c, r * c.shape[2], mode="linear")
