"""CPU emulation of the tensor-core operand roundings for the ClariNet student.

Every conv of oracle/clarinet_ref.py is replaced by  sum of products of ROUNDED operands accumulated in float64
(the tensor core accumulates exact products in fp32: float64 here only removes the emulation's own noise), and the
waveform is compared with the plain fp32 oracle.  Schemes:
  fp16x3      a_hi*w_hi + a_hi*w_lo + a_lo*w_hi           (what tc_conv.cuh / tc_block.cuh do today)
  fp16x2_a    a_hi*w_hi + a_lo*w_hi                        (weights rounded to 11 bits)
  fp16x2_w    a_hi*w_hi + a_hi*w_lo                        (activations rounded to 11 bits)
  fp16x1      a_hi*w_hi
  e4m3_same   corrections with e4m3 operands, scales cancelling inside one accumulator (a_lo*2^4, w_hi*2^-4)
  e4m3_sep    corrections in a second accumulator carrying 2^8 (a_lo*2^12, w_hi*2^-4; a_hi, w_lo*2^8)
  e5m2_alo    like e4m3_same but a_lo in e5m2
Usage: python tools/split_precision_study.py [frames]   (needs the staged checkpoints, oracle/stage_weights.py)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import clarinet_ref as C  # noqa: E402

W = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "weights")
real_conv1d = torch.nn.functional.conv1d


def h16(x):
    return x.to(torch.float16).to(torch.float64)


def q8(x, dt):
    return x.to(torch.float32).to(dt).to(torch.float64)


def row_scale(w):
    m = w.abs().flatten(1).max(1).values.clamp_min(1e-30)
    e = torch.floor(torch.log2(m))
    return torch.pow(2.0, 11 - e).to(torch.float64).view(-1, 1, 1)       # row max -> [2^11, 2^12)


real_convT1d = torch.nn.functional.conv_transpose1d


def make_conv(scheme, transposed=False):
    E4, E5 = torch.float8_e4m3fn, torch.float8_e5m2

    def conv(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
        x64, w64 = x.to(torch.float64), w.to(torch.float64)
        if transposed:      # weight [C_in][C_out][K]: the scale belongs to the OUTPUT channel (dim 1)
            s = row_scale(w64.transpose(0, 1)).view(1, -1, 1)
        else:
            s = row_scale(w64)
        ws = w64 * s
        a_hi = h16(x64); a_lo = h16(x64 - a_hi)
        w_hi = h16(ws); w_lo = h16(ws - w_hi)
        if transposed:
            cv = lambda a, ww: real_convT1d(a, ww, None, stride=stride, padding=padding)
        else:
            cv = lambda a, ww: real_conv1d(a, ww, None, stride, padding, dilation, groups)
        y = cv(a_hi, w_hi)
        if scheme == "fp16x3":
            y = y + cv(a_hi, w_lo) + cv(a_lo, w_hi)
        elif scheme == "fp16x2_a":
            y = y + cv(a_lo, w_hi)
        elif scheme == "fp16x2_w":
            y = y + cv(a_hi, w_lo)
        elif scheme == "fp16x1":
            pass
        elif scheme == "e4m3_same":
            y = y + cv(q8(a_hi, E4), q8(w_lo, E4)) + cv(q8(a_lo * 16.0, E4), q8(w_hi / 16.0, E4))
        elif scheme == "e4m3_sep":
            y = y + (cv(q8(a_hi, E4), q8(w_lo * 256.0, E4)) + cv(q8(a_lo * 4096.0, E4), q8(w_hi / 16.0, E4))) / 256.0
        elif scheme == "e5m2_alo":
            y = y + cv(q8(a_hi, E4), q8(w_lo, E4)) + cv(q8(a_lo * 16.0, E5), q8(w_hi / 16.0, E4))
        else:
            raise ValueError(scheme)
        y = y / s.view(1, -1, 1)
        if b is not None:
            y = y + b.to(torch.float64).view(1, -1, 1)
        return y.to(torch.float32)

    return conv


SCHEMES = ("fp16x3", "fp16x2_a", "fp16x2_w", "fp16x1", "e4m3_same", "e4m3_sep", "e5m2_alo")


def hifigan(frames):
    """same study for the shipped HiFi-GAN generator (the tighter case: the trained net amplifies rounding by ~1e3)"""
    from oracle import hifigan_ref as H
    sd = torch.load(os.path.join(W, "g_00600000"), map_location="cpu", weights_only=False)["generator"]
    cfg = dict(H.CONFIG_NEB)
    for level in (0.0, 1.0):
        g = torch.Generator().manual_seed(1236)
        mel = torch.clamp(0.5 * torch.nn.functional.avg_pool1d(6 * torch.randn(2, 80, frames + 8, generator=g), 9, 1)
                          - torch.linspace(0, 4, 80)[None, :, None] + level, -11.5, 2.5)
        ref = H.generator_forward(sd, cfg, mel)
        print(f"hifigan level {level:+.0f}: fp32 oracle peak {float(ref.abs().max()):.3f}, T={ref.shape[-1]}")
        for scheme in SCHEMES:
            torch.nn.functional.conv1d = make_conv(scheme)
            torch.nn.functional.conv_transpose1d = lambda x, w, b=None, stride=1, padding=0, _s=scheme: make_conv(_s, True)(x, w, b, stride, padding)
            try:
                x = H.generator_forward(sd, cfg, mel)
            finally:
                torch.nn.functional.conv1d = real_conv1d
                torch.nn.functional.conv_transpose1d = real_convT1d
            err = (x - ref).abs()
            print(f"  {scheme:10s} max-abs {float(err.max()):.3e}   rms {float(err.pow(2).mean().sqrt()):.3e}")


def main():
    if "--hifigan" in sys.argv:
        return hifigan(int(sys.argv[-1]) if sys.argv[-1].isdigit() else 24)
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    ssd = torch.load(os.path.join(W, "pnn_vocoder.network"), map_location="cpu", weights_only=False)
    tsd = torch.load(os.path.join(W, "nn_vocoder.network"), map_location="cpu", weights_only=False)
    mel = C.synthetic_mel01(2, frames, seed=8)
    z = torch.randn(2, 1, frames * 256, generator=torch.Generator().manual_seed(9))
    c_up = C.upsample_mel(tsd, mel)
    ref = C.student_forward(ssd, z, c_up)
    print(f"fp32 oracle: peak {float(ref.abs().max()):.3f}, T={ref.shape[-1]}")
    for scheme in SCHEMES:
        torch.nn.functional.conv1d = make_conv(scheme)
        try:
            x = C.student_forward(ssd, z, c_up)
        finally:
            torch.nn.functional.conv1d = real_conv1d
        err = (x - ref).abs()
        print(f"{scheme:10s} max-abs {float(err.max()):.3e}   rms {float(err.pow(2).mean().sqrt()):.3e}")


if __name__ == "__main__":
    main()
