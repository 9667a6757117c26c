"""CPU oracle for Path C - ClariNet teacher / ParallelWaveNet (IAF) student.
TEST INFRASTRUCTURE, not product.

*** PARITY UNPINNED BY THE REFERENCE *** : the reference tree ships only the trained weights
(``data/models/nn_vocoder.network`` teacher, ``data/models/pnn_vocoder.network`` student; both plain
``torch.save`` state-dicts) - the forward code lived in upstream ``ksw0306/ClariNet`` (credited at
reference ``README.md:65``; not vendored, no pinned commit, absent from ``.gitmodules``).  This file
restates that published algorithm and anchors it on what the reference does hold:

  * key names / shapes of both checkpoints - ``check_state_dict`` demands an
    exact (strict) match;
  * the mel upsampler ``UpsampleNet2`` (reference ``cube/networks/modules.py:357-375``), whose
    weights are the teacher's ``upsample_conv.{0,2}`` - pinned bit-for-bit by
    tests/golden/upsample2.npz;
  * the Gaussian head (reference ``cube/networks/loss.py:35-66``);
  * a self-consistency probe (tests/test_oracle.py::test_clarinet_self_consistency_probe): student samples must score a far
    better teacher NLL under dilation 3^(i mod 6) + sqrt(0.5) residual scaling than under the
    alternatives.

Algorithm (upstream modules.py / wavenet.py / wavenet_iaf.py, restated):
  Conv(causal):  weight-normed Conv1d(padding=d*(k-1)) then drop the last d*(k-1) outputs.
  ResBlock:      f = filter_conv(h) + filter_conv_c(c);  g = gate_conv(h) + gate_conv_c(c)
                 o = tanh(f) * sigmoid(g);  h' = (h + res_conv(o)) * sqrt(0.5);  skip = skip_conv(o)
  Wavenet:       h = relu(front_conv(x));  skip_sum = sum_i skip_i;
                 out = conv1x1(relu(conv1x1(relu(skip_sum))))              -> [B, 2, T]
  Student IAF:   per flow: (mu, logs) = Wavenet_f(z, c);  z[t] <- z[t]*exp(logs[t-1]) + mu[t-1], z[0] <- 0
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

TEACHER_BLOCKS = 24          # res_blocks.{0..23} in nn_vocoder.network
STUDENT_BLOCKS = (6, 6, 6, 24)  # iafs.{0..3}.res_blocks.* in pnn_vocoder.network
RES_CH, GATE_CH, SKIP_CH, CIN_CH = 128, 256, 128, 80
KERNEL, FRONT_K = 3, 32
CYCLE = 6                    # dilation restarts every 6 layers
UPSAMPLE_SCALES = (16, 16)   # x256, hop 256


def dilation_of(i: int, base: int = 3, cycle: int = CYCLE) -> int:
    """dilation = kernel_size ** (i mod num_layers)  ->  1,3,9,27,81,243 (probe-selected)."""
    return base ** (i % cycle)


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        if k.endswith("weight_g"):
            base = k[: -len("weight_g")]
            out[base + "weight"] = torch._weight_norm(sd[base + "weight_v"], v, 0)
        elif not k.endswith("weight_v"):
            out[k] = v
    return out


def expected_keys(prefix: str, n_blocks: int) -> Dict[str, tuple]:
    """Key -> shape table of one Wavenet in the shipped checkpoints."""
    e = {}

    def wn(name, shape):
        e[f"{prefix}{name}.bias"] = (shape[0],)
        e[f"{prefix}{name}.weight_g"] = (shape[0],) + (1,) * (len(shape) - 1)
        e[f"{prefix}{name}.weight_v"] = shape

    wn("front_conv.0.conv", (RES_CH, 1, FRONT_K))
    for i in range(n_blocks):
        p = f"res_blocks.{i}."
        wn(p + "filter_conv.conv", (GATE_CH, RES_CH, KERNEL))
        wn(p + "gate_conv.conv", (GATE_CH, RES_CH, KERNEL))
        wn(p + "res_conv", (RES_CH, GATE_CH, 1))
        wn(p + "skip_conv", (SKIP_CH, GATE_CH, 1))
        wn(p + "filter_conv_c", (GATE_CH, CIN_CH, 1))
        wn(p + "gate_conv_c", (GATE_CH, CIN_CH, 1))
    wn("final_conv.1.conv", (SKIP_CH, SKIP_CH, 1))
    wn("final_conv.3.conv", (2, SKIP_CH, 1))
    return e


def check_state_dict(sd: Dict[str, torch.Tensor], kind: str) -> None:
    """Strict key/shape match against the module tree of the shipped checkpoints."""
    exp = {}
    if kind == "teacher":
        exp.update(expected_keys("", TEACHER_BLOCKS))
        for i in (0, 2):
            exp[f"upsample_conv.{i}.bias"] = (1,)
            exp[f"upsample_conv.{i}.weight_g"] = (1, 1, 1, 1)
            exp[f"upsample_conv.{i}.weight_v"] = (1, 1, 3, 32)
    elif kind == "student":
        for f, nb in enumerate(STUDENT_BLOCKS):
            exp.update(expected_keys(f"iafs.{f}.", nb))
    else:
        raise ValueError(kind)
    got = {k: tuple(v.shape) for k, v in sd.items()}
    missing = sorted(set(exp) - set(got))
    extra = sorted(set(got) - set(exp))
    bad = sorted(k for k in exp if k in got and exp[k] != got[k])
    if missing or extra or bad:
        raise KeyError(f"{kind} state_dict mismatch: missing={missing[:4]} extra={extra[:4]} shape={bad[:4]}")


def random_state_dict(kind: str, seed: int = 0, blocks: Sequence[int] | None = None) -> Dict[str, torch.Tensor]:
    """Seeded random weights with the shipped key layout (used when the checkpoints are not staged,
    e.g. a GPU box without oracle/_ref/weights).  Scales are chosen so activations stay O(1)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def fill(table, gain_last=None):
        for k, shape in table.items():
            if k.endswith("weight_v"):
                fan_in = 1
                for s in shape[1:]:
                    fan_in *= s
                sd[k] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
            elif k.endswith("weight_g"):
                v = table[k[:-1] + "v"]
                fan_in = 1
                for s in v[1:]:
                    fan_in *= s
                sd[k] = torch.ones(shape) * (0.8 + 0.4 * torch.rand(shape, generator=g))
            else:
                sd[k] = 0.05 * torch.randn(shape, generator=g)

    if kind == "teacher":
        fill(expected_keys("", blocks[0] if blocks else TEACHER_BLOCKS))
        for i in (0, 2):
            sd[f"upsample_conv.{i}.bias"] = torch.tensor([0.01 * (i + 1)])
            sd[f"upsample_conv.{i}.weight_g"] = torch.full((1, 1, 1, 1), 1.3)
            sd[f"upsample_conv.{i}.weight_v"] = torch.randn(1, 1, 3, 32, generator=g) * 0.2 + 0.05
    else:
        for f, nb in enumerate(blocks if blocks else STUDENT_BLOCKS):
            fill(expected_keys(f"iafs.{f}.", nb))
            # keep log-scales small so exp(logs) stays tame through the flows
            sd[f"iafs.{f}.final_conv.3.conv.weight_g"] = torch.tensor([[[0.5]], [[0.1]]])
            sd[f"iafs.{f}.final_conv.3.conv.bias"] = torch.tensor([0.0, -1.0])
    return sd


def student_blocks_of(sd: Dict[str, torch.Tensor]) -> List[int]:
    nb: Dict[int, int] = {}
    for k in sd:
        if k.startswith("iafs.") and ".res_blocks." in k:
            p = k.split(".")
            nb[int(p[1])] = max(nb.get(int(p[1]), 0), int(p[3]) + 1)
    return [nb[i] for i in sorted(nb)]


def teacher_blocks_of(sd: Dict[str, torch.Tensor]) -> int:
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("res_blocks."))


# ----------------------------------------------------------------------------------------------
# upsampler: cube/networks/modules.py:357-375 (UpsampleNet2), weights = teacher upsample_conv.{0,2}
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def upsample_mel(teacher_sd: Dict[str, torch.Tensor], mel: torch.Tensor,
                 scales: Sequence[int] = UPSAMPLE_SCALES) -> torch.Tensor:
    """mel [B, 80, F] -> c_up [B, 80, F*prod(scales)].  ConvTranspose2d(1,1,(3,2s),stride=(1,s),
    padding=(1,s//2)) + LeakyReLU(0.4), twice."""
    w = fold_weight_norm({k: v for k, v in teacher_sd.items() if k.startswith("upsample_conv.")})
    c = mel.to(torch.float32).unsqueeze(1)
    for n, s in enumerate(scales):
        c = F.conv_transpose2d(c, w[f"upsample_conv.{2 * n}.weight"], w[f"upsample_conv.{2 * n}.bias"],
                               stride=(1, s), padding=(1, s // 2))
        c = F.leaky_relu(c, 0.4)
    return c.squeeze(1)


# ----------------------------------------------------------------------------------------------
# Wavenet (one flow of the student == the teacher network)
# ----------------------------------------------------------------------------------------------
def _causal_conv(x, weight, bias, dilation):
    k = weight.shape[2]
    pad = dilation * (k - 1)
    y = F.conv1d(x, weight, bias, padding=pad, dilation=dilation)
    return y[:, :, :-pad] if pad else y


@torch.no_grad()
def wavenet_forward(w: Dict[str, torch.Tensor], prefix: str, n_blocks: int, x: torch.Tensor,
                    c_up: torch.Tensor, dil_base: int = 3, res_scale: float = math.sqrt(0.5),
                    cycle: int = CYCLE) -> torch.Tensor:
    """x [B,1,T], c_up [B,80,T] -> [B,2,T] (mu, log-scale).  ``w`` holds folded weights."""
    p = prefix
    h = F.relu(_causal_conv(x, w[p + "front_conv.0.conv.weight"], w[p + "front_conv.0.conv.bias"], 1))
    skip = None
    for i in range(n_blocks):
        b = f"{p}res_blocks.{i}."
        d = dil_base ** (i % cycle)
        hf = _causal_conv(h, w[b + "filter_conv.conv.weight"], w[b + "filter_conv.conv.bias"], d)
        hg = _causal_conv(h, w[b + "gate_conv.conv.weight"], w[b + "gate_conv.conv.bias"], d)
        hf = hf + F.conv1d(c_up, w[b + "filter_conv_c.weight"], w[b + "filter_conv_c.bias"])
        hg = hg + F.conv1d(c_up, w[b + "gate_conv_c.weight"], w[b + "gate_conv_c.bias"])
        o = torch.tanh(hf) * torch.sigmoid(hg)
        res = F.conv1d(o, w[b + "res_conv.weight"], w[b + "res_conv.bias"])
        s = F.conv1d(o, w[b + "skip_conv.weight"], w[b + "skip_conv.bias"])
        h = (h + res) * res_scale
        skip = s if skip is None else skip + s
    y = F.relu(skip)
    y = F.conv1d(y, w[p + "final_conv.1.conv.weight"], w[p + "final_conv.1.conv.bias"])
    y = F.relu(y)
    return F.conv1d(y, w[p + "final_conv.3.conv.weight"], w[p + "final_conv.3.conv.bias"])


@torch.no_grad()
def student_forward(student_sd: Dict[str, torch.Tensor], z: torch.Tensor, c_up: torch.Tensor,
                    dil_base: int = 3, res_scale: float = math.sqrt(0.5)) -> torch.Tensor:
    """ParallelWaveNet sampling: z [B,1,T] ~ N(0,1) (injected), c_up [B,80,T] -> x [B,1,T]."""
    w = {k: v.float() for k, v in fold_weight_norm(student_sd).items()}
    z = z.to(torch.float32)
    for f, nb in enumerate(student_blocks_of(student_sd)):
        ml = wavenet_forward(w, f"iafs.{f}.", nb, z, c_up, dil_base, res_scale)
        mu, logs = ml[:, 0:1, :-1], ml[:, 1:2, :-1]
        z = F.pad(z[:, :, 1:] * torch.exp(logs) + mu, (1, 0))
    return z


@torch.no_grad()
def vocode_student(student_sd, teacher_sd, mel: torch.Tensor, z: torch.Tensor, **kw) -> torch.Tensor:
    """mel [B,80,F], z [B,1,256F] -> waveform [B,1,256F] (config 2 of BASELINE.json)."""
    return student_forward(student_sd, z, upsample_mel(teacher_sd, mel), **kw)


@torch.no_grad()
def teacher_nll(teacher_sd, x: torch.Tensor, c_up: torch.Tensor, dil_base: int = 3,
                res_scale: float = math.sqrt(0.5)) -> float:
    """Teacher-forced Gaussian NLL (nats/sample) of a waveform under the teacher: out[:, :, t]
    predicts x[t+1] (cube/networks/loss.py:36-48 with log_std_min=-14)."""
    w = {k: v.float() for k, v in fold_weight_norm(teacher_sd).items()}
    ml = wavenet_forward(w, "", teacher_blocks_of(teacher_sd), x, c_up, dil_base, res_scale)
    mean, log_std = ml[:, 0, :-1], torch.clamp(ml[:, 1, :-1], min=-14.0)
    y = x[:, 0, 1:]
    nll = 0.5 * math.log(2 * math.pi) + log_std + 0.5 * (y - mean) ** 2 * torch.exp(-2 * log_std)
    return float(nll.mean())


@torch.no_grad()
def teacher_generate(teacher_sd, c_up: torch.Tensor, eps: torch.Tensor, dil_base: int = 3,
                     res_scale: float = math.sqrt(0.5)) -> torch.Tensor:
    """Autoregressive teacher sampling with injected noise eps [B,T] (config 1; CPU-only by
    definition): x[t+1] = mean_t + 0.8 * eps_t * exp(log_std_t) (cube/networks/loss.py:50-52).
    O(T * receptive field) - naive recompute over the receptive field; use for short clips."""
    w = {k: v.float() for k, v in fold_weight_norm(teacher_sd).items()}
    nb = teacher_blocks_of(teacher_sd)
    B, _, T = c_up.shape
    rf = (FRONT_K - 1) + sum((KERNEL - 1) * dil_base ** (i % CYCLE) for i in range(nb)) + 1
    x = torch.zeros(B, 1, T)
    for t in range(T - 1):
        lo = max(0, t + 1 - rf)
        ml = wavenet_forward(w, "", nb, x[:, :, lo:t + 1], c_up[:, :, lo:t + 1], dil_base, res_scale)
        x[:, 0, t + 1] = ml[:, 0, -1] + 0.8 * eps[:, t] * torch.exp(ml[:, 1, -1])
    return x


def synthetic_mel01(B: int, F_: int, seed: int, num_mels: int = 80) -> torch.Tensor:
    """SURVEY 8(d) cfg1/cfg2 input: smooth field in [0,1] (ClariNet-era min/max normalisation,
    reference cube/io_utils/vocoder.py:86-88)."""
    g = torch.Generator().manual_seed(seed)
    r = torch.randn(B, num_mels, F_ + 8, generator=g)
    sm = F.avg_pool1d(r, 9, stride=1)[:, :, :F_] * 3.0
    return torch.sigmoid(sm).contiguous()
