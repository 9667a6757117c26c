#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE in the build container.

Run here (``python oracle/make_goldens.py``); needs /root/reference, which does not exist on the
GPU box - that is why the outputs are committed.  Nothing is copied from the reference: modules are
imported from where they lie and only their numerical outputs are stored.

  hifigan_mini.npz     reference hifigan/models.py:Generator, mini config (C0=32), seeded weights
                       stored in the file, ResBlock1; loud input
  hifigan_mini_rb2.npz same with resblock "2" (config_v3-style, hifigan/models.py:63-68)
  hifigan_neb.npz      reference Generator + shipped data/models/vocoder/neb-noft/g_00600000
                       (weights NOT stored - staged by oracle/stage_weights.py), F=24
  heads.npz            cube/networks/loss.py MULAW/RAW/MOL/Gaussian encode/decode/sample
  upsample2.npz        cube/networks/modules.py UpsampleNet2/R/I (+ teacher upsample weights)
  upsamplenet.npz      cube/networks/modules.py:317-343 UpsampleNet (3 conv+tanh, weight-normed transposed convs), seeded weights stored
  mel_hifigan.npz      hifigan/meldataset.py:mel_spectrogram (librosa stubbed with torchaudio's Slaney filter bank)
  clarinet_regress.npz NOT reference-derived (no forward code in the reference): regression
                       snapshot of oracle/clarinet_ref.py with the shipped checkpoints
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("CUBE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))


def _stub_matplotlib():
    # hifigan/utils.py:3,6-7 imports matplotlib (not installed); the generator never uses it
    m = types.ModuleType("matplotlib")
    m.use = lambda *a, **k: None
    p = types.ModuleType("matplotlib.pylab")
    m.pylab = p
    sys.modules.setdefault("matplotlib", m)
    sys.modules.setdefault("matplotlib.pylab", p)


def import_reference_hifigan():
    _stub_matplotlib()
    sys.path.insert(0, os.path.join(REF, "hifigan"))
    import models as ref_models  # noqa  (reference hifigan/models.py)
    from env import AttrDict      # noqa  (reference hifigan/env.py)
    return ref_models, AttrDict


def main():
    from oracle import hifigan_ref as H, clarinet_ref as C, heads_ref as W
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref_models, AttrDict = import_reference_hifigan()

    # ---------------- Path H, mini configs with stored weights ----------------
    for tag, resblock, dil, gs in (("hifigan_mini", "1", [[1, 3, 5]] * 3, 0.42),
                                   ("hifigan_mini_rb2", "2", [[1, 3]] * 3, 0.5)):
        cfg = dict(H.CONFIG_V1, upsample_initial_channel=32, resblock=resblock, resblock_dilation_sizes=dil)
        sd = H.random_state_dict(cfg, seed=11, std=0.3, g_scale=gs)
        gen = ref_models.Generator(AttrDict(cfg)).eval()
        gen.load_state_dict(sd, strict=True)
        mel = H.synthetic_mel(2, 13, seed=5, level=0.0)
        with torch.no_grad():
            y = gen(mel)
        print(tag, tuple(y.shape), "peak", float(y.abs().max()))
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), mel=mel.numpy(), wav=y.numpy(),
                            cfg=np.array(repr(cfg)), **{"w:" + k: v.numpy() for k, v in sd.items()})

    # ---------------- Path H, shipped generator ----------------
    ck = torch.load(os.path.join(REF, "data/models/vocoder/neb-noft/g_00600000"), map_location="cpu",
                    weights_only=False)["generator"]
    cfg = H.load_config(os.path.join(REF, "data/models/vocoder/neb-noft/config.json"))
    gen = ref_models.Generator(AttrDict(cfg)).eval()
    gen.load_state_dict(ck, strict=True)
    mel = H.synthetic_mel(2, 24, seed=1237, level=0.0)
    with torch.no_grad():
        y_wn = gen(mel)                 # weight-norm live (the cube/api.py path)
        gen.remove_weight_norm()        # the cube/io_utils/runtime.py:53 path
        y = gen(mel)
    print("hifigan_neb", tuple(y.shape), "peak", float(y.abs().max()), "wn-vs-folded", float((y - y_wn).abs().max()))
    np.savez_compressed(os.path.join(OUT, "hifigan_neb.npz"), mel=mel.numpy(), wav=y.numpy(),
                        wav_int16=(y.numpy().squeeze(1) * 32767).astype(np.int16))

    # ---------------- heads ----------------
    sys.path.insert(0, REF)
    from cube.networks import loss as ref_loss  # reference cube/networks/loss.py, unmodified
    mu = ref_loss.MULAWOutput()
    raw = ref_loss.RAWOutput()
    mol = ref_loss.MOLOutput()
    gau = ref_loss.GaussianOutput()
    g = torch.Generator().manual_seed(77)
    x = torch.cat([torch.linspace(-1, 1, 20001), torch.rand(20000, generator=g) * 2 - 1,
                   torch.tensor([1.0, 0.9, 0.0, -0.9, -1.0, 1.5, -1.5, 1e-8, -1e-8])])
    enc = mu.encode(x)
    table = mu.decode(torch.arange(256))
    edges = W.mulaw_encode_edges(encode=mu.encode)
    # neighbours of every edge (the hard cases)
    nb = np.concatenate([np.nextafter(edges, -np.inf, dtype=np.float32), edges,
                         np.nextafter(edges, np.inf, dtype=np.float32)])
    enc_nb = mu.encode(torch.from_numpy(nb))
    y = torch.randn(3, 50, 30, generator=g)
    y[:, :, 20:] = y[:, :, 20:] * 0.5 - 3.0
    torch.manual_seed(4242)
    x_mol = mol.sample(y)
    torch.manual_seed(4242)
    u_mix = torch.empty(3, 50, 10).uniform_(1e-5, 1 - 1e-5)
    u_x = torch.empty(3, 50).uniform_(1e-5, 1.0 - 1e-5)
    yh = torch.randn(3, 50, 2, generator=g)
    yh[:, :, 1] = yh[:, :, 1] * 0.3 - 2.0
    torch.manual_seed(99)
    x_g = gau.sample(yh)
    torch.manual_seed(99)
    eps = torch.randn(3, 50, 1)
    xr = torch.rand(5000, generator=g) * 2.4 - 1.2
    np.savez_compressed(
        os.path.join(OUT, "heads.npz"), mulaw_x=x.numpy(), mulaw_q=enc.numpy(), mulaw_table=table.numpy(),
        mulaw_edges=edges, mulaw_nb_x=nb, mulaw_nb_q=enc_nb.numpy(),
        raw_x=xr.numpy(), raw_q=raw.encode(xr).numpy(), raw_table=raw.decode(torch.arange(256)).numpy(),
        mol_y=y.numpy(), mol_u_mix=u_mix.numpy(), mol_u_x=u_x.numpy(), mol_x=x_mol.numpy(),
        gau_y=yh.numpy(), gau_eps=eps.squeeze(2).numpy(), gau_x=x_g.reshape(3, 50).numpy())
    print("heads: KAT", mu.encode(np.array([1, 0.9, 0, -0.9, -1])), "edges", edges[:3], edges[-3:])

    # ---------------- upsamplers ----------------
    from cube.networks import modules as ref_mod  # reference cube/networks/modules.py, unmodified
    tsd = torch.load(os.path.join(REF, "data/models/nn_vocoder.network"), map_location="cpu", weights_only=False)
    ssd = torch.load(os.path.join(REF, "data/models/pnn_vocoder.network"), map_location="cpu", weights_only=False)
    C.check_state_dict(tsd, "teacher")
    C.check_state_dict(ssd, "student")
    up = ref_mod.UpsampleNet2([16, 16])
    up.load_state_dict({k.replace("upsample_conv.", "_upsample_conv."): v for k, v in tsd.items()
                        if k.startswith("upsample_conv.")}, strict=True)
    mel01 = C.synthetic_mel01(1, 5, seed=3)
    with torch.no_grad():
        c_up = up(mel01)
        r3 = ref_mod.UpsampleNetR(3)(mel01[:, :4])
        i3 = ref_mod.UpsampleNetI(3)(mel01[:, :4])
    np.savez_compressed(os.path.join(OUT, "upsample2.npz"), mel=mel01.numpy(), c_up=c_up.numpy(),
                        rep3=r3.numpy(), lin3=i3.numpy(),
                        **{"w:" + k: v.numpy() for k, v in tsd.items() if k.startswith("upsample_conv.")})
    print("upsample2", tuple(c_up.shape))

    # ---------------- ClariNet regression snapshot (oracle-derived, NOT reference-derived) -------
    melc = C.synthetic_mel01(1, 6, seed=8)
    z = torch.randn(1, 1, 6 * 256, generator=torch.Generator().manual_seed(9))
    xs = C.vocode_student(ssd, tsd, melc, z)
    np.savez_compressed(os.path.join(OUT, "clarinet_regress.npz"), mel=melc.numpy(), z=z.numpy(), wav=xs.numpy())
    print("clarinet_regress", tuple(xs.shape), "std", float(xs.std()), "peak", float(xs.abs().max()))


_ONLY = {"--inc-only", "--wavernn", "--mel", "--upsamplenet"} & set(sys.argv)      # no flag: regenerate everything

if __name__ == "__main__" and not _ONLY:
    main()


def write_mulaw_inc():
    """Emit tts_cube_b200/csrc/mulaw_tables.inc from tests/golden/heads.npz (the reference's own
    float32 bin edges / decode values, as exact hex-float literals)."""
    d = np.load(os.path.join(OUT, "heads.npz"))
    dst = os.path.join(os.path.dirname(HERE), "tts_cube_b200", "csrc", "mulaw_tables.inc")
    with open(dst, "w") as f:
        f.write("// GENERATED by oracle/make_goldens.py from the reference's float32 torch path\n"
                "// (cube/networks/loss.py:236-269 executed in the build container). Do not edit.\n")
        f.write("static const float MULAW_EDGES_H[255] = {\n")
        f.write(",\n".join("  " + float(v).hex() + "f" for v in d["mulaw_edges"]))
        f.write("\n};\nstatic const float MULAW_DECODE_H[256] = {\n")
        f.write(",\n".join("  " + float(v).hex() + "f" for v in d["mulaw_table"]))
        f.write("\n};\n")


if __name__ == "__main__" and (not _ONLY or "--inc-only" in _ONLY):
    write_mulaw_inc()


def wavernn_goldens():
    """tests/golden/wavernn_{mol,gm}.npz: the reference WaveRNN (cube/networks/modules.py:392-503) with seeded
    weights (stored) and torch.manual_seed; the head's per-step draws are replayed from the same seed."""
    import contextlib, io
    sys.path.insert(0, REF)
    from cube.networks import modules as ref_mod
    from oracle import wavernn_ref as R
    for out, S in (("mol", 30), ("gm", 2)):
        H_, up, upl, B, Fr = 32, 6, 3, 2, 4
        m = ref_mod.WaveRNN(num_layers=2, layer_size=H_, upsample=up, upsample_low=upl, use_lowres=True, output=out).eval()
        sd = R.random_state_dict(H_, 2, True, S, seed=31)
        m.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(17)
        mel = torch.rand(B, Fr, 80, generator=g)
        x_low = torch.rand(B, Fr * up // upl, generator=g) * 1.6 - 0.8
        T = Fr * up
        torch.manual_seed(123)
        with contextlib.redirect_stderr(io.StringIO()):
            y = m({"mel": mel, "x_low": x_low})
        torch.manual_seed(123)
        draws = {}
        if out == "mol":
            um, ux = [], []
            for _ in range(T):
                um.append(torch.empty(B, 1, 10).uniform_(1e-5, 1 - 1e-5)[:, 0])
                ux.append(torch.empty(B, 1).uniform_(1e-5, 1.0 - 1e-5)[:, 0])
            draws = {"u_mix": torch.stack(um), "u_x": torch.stack(ux)}
        else:
            draws = {"eps": torch.stack([torch.randn(B, 1, 1)[:, 0, 0] for _ in range(T)])}
        np.savez_compressed(os.path.join(OUT, f"wavernn_{out}.npz"), mel=mel.numpy(), x_low=x_low.numpy(),
                            x=np.asarray(y).reshape(B, T), upsample=up, upsample_low=upl,
                            **{"d:" + k: v.numpy() for k, v in draws.items()}, **{"w:" + k: v.numpy() for k, v in sd.items()})
        print("wavernn", out, np.asarray(y).shape, float(np.abs(y).max()))


if __name__ == "__main__" and (not _ONLY or "--wavernn" in _ONLY):
    wavernn_goldens()


def mel_goldens():
    """tests/golden/mel_hifigan.npz: the UNMODIFIED hifigan/meldataset.py:mel_spectrogram executed here.
    Harness shims (the reference files are untouched): `librosa` is not installed, so a stub module provides
    `filters.mel` from torchaudio's Slaney filter bank (an implementation independent of ours) and inert
    `load` / `util.normalize`; torch.stft gets `return_complex=False` (mandatory since torch 2.0, absent in the
    torch==1.4 reference call)."""
    import torchaudio
    from oracle import mel_ref as M

    def mel_fn(sr, n_fft, n_mels, fmin, fmax):
        return torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, float(fmin), float(fmax), n_mels, sr, norm="slaney",
                                                     mel_scale="slaney").T.contiguous().numpy()

    lib = types.ModuleType("librosa")
    lib.load = lambda *a, **k: None
    lib.util = types.ModuleType("librosa.util")
    lib.util.normalize = lambda x, *a, **k: x
    lib.filters = types.ModuleType("librosa.filters")
    lib.filters.mel = mel_fn
    for k, v in (("librosa", lib), ("librosa.util", lib.util), ("librosa.filters", lib.filters)):
        sys.modules[k] = v
    sys.path.insert(0, os.path.join(REF, "hifigan"))
    import meldataset as ref_md  # noqa (reference hifigan/meldataset.py)

    real_stft = torch.stft
    torch.stft = lambda *a, **k: real_stft(*a, **{"return_complex": False, **k})
    try:
        out = {}
        # (a) the shipped vocoder's front-end (neb-noft config: 22 050 Hz, hop 256... values from its config.json),
        # (b) the call Cubegan makes (cube/networks/cubegan.py:137: 1024, 80, 24000, 240, 1024, 0, 12000)
        for tag, args in (("a", (1024, 80, 22050, 256, 1024, 0, 8000)), ("b", (1024, 80, 24000, 240, 1024, 0, 12000))):
            ref_md.mel_basis.clear(); ref_md.hann_window.clear()
            y = M.test_signal(2, 47 * args[3], seed=5 if tag == "a" else 6, sr=args[2])
            mel = ref_md.mel_spectrogram(y, *args)
            out[f"y_{tag}"] = y.numpy(); out[f"mel_{tag}"] = mel.numpy(); out[f"args_{tag}"] = np.asarray(args)
            out[f"basis_{tag}"] = mel_fn(args[2], args[0], args[1], args[5], args[6])
            print("mel", tag, tuple(mel.shape), float(mel.min()), float(mel.max()))
        np.savez_compressed(os.path.join(OUT, "mel_hifigan.npz"), **out)
    finally:
        torch.stft = real_stft


if __name__ == "__main__" and (not _ONLY or "--mel" in _ONLY):
    mel_goldens()


def upsamplenet_golden():
    """tests/golden/upsamplenet.npz: the UNMODIFIED reference UpsampleNet (cube/networks/modules.py:317-343), default-initialised
    under torch.manual_seed(77), scales [2, 2, 4], 80 -> 24 channels; weights, input and output stored."""
    sys.path.insert(0, REF)
    from cube.networks import modules as ref_mod     # reference
    from oracle import wavernn_ref as R
    torch.manual_seed(77)
    scales = [2, 2, 4]
    m = ref_mod.UpsampleNet(upsample_scales=scales, in_channels=80, out_channels=24, kernel_size=3).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    c = torch.rand(2, 80, 9, generator=torch.Generator().manual_seed(5)) * 2 - 1
    with torch.no_grad():
        y = m(c)
    assert float((R.upsamplenet_forward(sd, c, scales) - y).abs().max()) == 0.0
    np.savez_compressed(os.path.join(OUT, "upsamplenet.npz"), c=c.numpy(), y=y.numpy(), scales=np.array(scales), in_channels=80,
                        out_channels=24, kernel_size=3, **{"w:" + k: v.numpy() for k, v in sd.items()})
    print("upsamplenet", tuple(y.shape), float(y.abs().max()))


if __name__ == "__main__" and (not _ONLY or "--upsamplenet" in _ONLY):
    upsamplenet_golden()
