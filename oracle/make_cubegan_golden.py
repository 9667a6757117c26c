#!/usr/bin/env python
"""tests/golden/cubegan_e2e.npz: the UNMODIFIED reference ``Cubegan.inference`` (cube/networks/cubegan.py:74-83)
executed in the build container, end to end from a phoneme list to the waveform - BASELINE configs[4]'s call chain
``TTSCube.__call__`` (cube/api.py:45-66) minus the phonemizer (its model files are download-only,
cube/io_utils/repository.py:27-28: the phoneme list is fed straight into ``rez['meta']``, as SURVEY 8(c) describes).

TEST INFRASTRUCTURE: needs /root/reference (absent on the GPU box), so the outputs are committed.  Nothing is copied
from the reference; its modules are imported from where they lie, with harness-side stubs for the packages this image
lacks (pytorch_lightning, matplotlib, librosa, fasttext).  CWD must be the reference root because Cubegan opens
'hifigan/config_v1.json' by relative path (cubegan.py:41).

The published Cubegan models are download-only, so the frontend (Languasito2, 13.5 M parameters) is seeded random-init;
the generator gets the seeded LOUD weights of ``oracle.hifigan_ref.random_state_dict`` (a default-initialised generator
is near-silent and would make a 1e-3 comparison vacuous).  Stored: the phoneme-level inputs, the conditioning the
frontend handed to the generator ([1, F, 80], captured by a forward pre-hook), the reference waveform and the int16
audio ``TTSCube.__call__`` derives from it; the generator weights are regenerated from their seed by the test.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("CUBE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden", "cubegan_e2e.npz")
GEN_SEED, GEN_STD, GEN_GSCALE = 21, 0.3, 0.125
PHONES = ["h", "e", "l", "o", "w"]            # a short utterance keeps the fixture small
ALPHABET = [chr(ord("a") + i) for i in range(26)] + ["_", ".", ","]      # 29 phones


def _stubs():
    import transformers  # noqa: F401  (first: it probes find_spec('librosa') and chokes on a spec-less stub)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    pl = mod("pytorch_lightning", LightningModule=torch.nn.Module)
    cb = mod("pytorch_lightning.callbacks", Callback=object)
    pl.callbacks = cb
    mp = mod("matplotlib", use=lambda *a, **k: None)
    mp.pylab = mod("matplotlib.pylab")
    unused = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("librosa is stubbed: not on the inference path"))
    lb = mod("librosa", load=unused)
    lb.util = mod("librosa.util", normalize=unused)
    lb.filters = mod("librosa.filters", mel=unused)
    ft = mod("fasttext")
    ft.util = mod("fasttext.util")


def main():
    _stubs()
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "hifigan"))
    os.chdir(REF)
    from oracle import hifigan_ref as H
    from cube.io_utils.io_cubegan import CubeganCollate, CubeganEncodings     # reference
    from cube.networks.cubegan import Cubegan                                 # reference

    enc = CubeganEncodings()
    enc.phon2int = {p: i for i, p in enumerate(ALPHABET)}
    enc.speaker2int = {"neb": 0}
    enc.max_pitch, enc.max_duration = 400, 100
    torch.manual_seed(1234)
    model = Cubegan(enc, conditioning=None, train=False)
    model.eval()
    sd = H.random_state_dict(H.CONFIG_V1, seed=GEN_SEED, std=GEN_STD, g_scale=GEN_GSCALE)
    missing = model._generator.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    captured = {}
    model._generator.register_forward_pre_hook(lambda m, a: captured.__setitem__("x", a[0].detach().clone()))

    # the dict cube/api.py:47-57 fabricates around the phonemizer's output
    rez = {"meta": {"phones": list(PHONES), "phon2word": [0] * len(PHONES), "words": ["hello"], "speaker": "neb"}}
    rez["pitch"] = np.zeros((100))
    rez["mgc"] = np.zeros((100, 80))
    rez["meta"]["words_left"] = []
    rez["meta"]["words_right"] = []
    rez["meta"]["frame2phon"] = [0] * 100
    collate = CubeganCollate(enc, conditioning_type=None)
    with torch.no_grad():
        X = collate.collate_fn([rez])
        x_char, x_speaker = X["x_char"].clone(), X["x_speaker"].clone()
        audio = model.inference(X)                                            # cube/networks/cubegan.py:74-83
    cond = captured["x"].permute(0, 2, 1).contiguous()                        # what languasito.inference returned: [1, F, 80]
    wav = audio.detach().cpu().numpy()
    a16 = np.asarray(wav.squeeze() * 32767, dtype=np.int16)                   # cube/api.py:64-65
    F_ = cond.shape[1]
    assert wav.shape == (1, 1, H.out_len(H.CONFIG_V1, F_)), wav.shape
    # the restatement agrees with the reference run (same ATen calls)
    ref2 = H.generator_forward(sd, H.CONFIG_V1, captured["x"])
    print(f"F={F_} T={wav.shape[-1]} peak={np.abs(wav).max():.3f} restatement-vs-reference max-abs {float((ref2 - audio).abs().max()):.2e}")
    np.savez_compressed(OUT, phones=np.array(PHONES), alphabet=np.array(ALPHABET), x_char=x_char.numpy(), x_speaker=x_speaker.numpy(),
                        conditioning=cond.numpy().astype(np.float32), wav=wav.astype(np.float32), wav_int16=a16,
                        gen_seed=GEN_SEED, gen_std=GEN_STD, gen_gscale=GEN_GSCALE, torch_seed=1234)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
