#!/usr/bin/env python
"""CPU timing of the reference frontend at batch 1 (``Languasito2.inference`` per utterance, as ``Cubegan.inference`` runs it) against
the batched driver of the same modules (tts_cube_b200/frontend.py).  BUILD CONTAINER ONLY: imports the unmodified reference class from
/root/reference (seeded random weights; the published models are download-only).  Prints one JSON line; CPU numbers - on a GPU the
per-utterance path pays ~10 cuDNN LSTM launches per utterance on top."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "oracle", "make_cubegan_golden.py"))
mk = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mk)
mk._stubs()
sys.path[:0] = [mk.REF, os.path.join(mk.REF, "hifigan")]
os.chdir(mk.REF)
from cube.io_utils.io_cubegan import CubeganCollate, CubeganEncodings      # noqa: E402  (reference)
from cube.networks.cubegan import Cubegan                                  # noqa: E402  (reference)
from tts_cube_b200 import frontend as FE                                   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.set_num_threads(int(os.environ.get("THREADS", "16")))
enc = CubeganEncodings()
enc.phon2int = {p: i for i, p in enumerate(mk.ALPHABET)}
enc.speaker2int = {"neb": 0}
enc.max_pitch, enc.max_duration = 400, 100
torch.manual_seed(1234)
model = Cubegan(enc, conditioning=None, train=False).eval()
lang = model._languasito
collate = CubeganCollate(enc, conditioning_type=None)
rng = np.random.default_rng(5)
Xs = []
for _ in range(N):
    phones = [mk.ALPHABET[i] for i in rng.integers(0, len(mk.ALPHABET), int(rng.integers(20, 61)))]
    rez = {"meta": {"phones": phones, "phon2word": [0] * len(phones), "words": ["w"], "speaker": "neb", "words_left": [],
                    "words_right": [], "frame2phon": [0] * 100}, "pitch": np.zeros(100), "mgc": np.zeros((100, 80))}
    Xs.append(collate.collate_fn([rez]))
xs = [X["x_char"][0].clone() for X in Xs]
sp = [X["x_speaker"][0].clone() for X in Xs]
with torch.no_grad():
    t0 = time.perf_counter()
    want = [lang.inference(dict(X))[0] for X in Xs]
    t_ref = time.perf_counter() - t0
    FE.languasito_inference_batch(lang, xs[:4], sp[:4])
    t0 = time.perf_counter()
    got = FE.languasito_inference_batch(lang, xs, sp)
    t_b = time.perf_counter() - t0
err = max(float((a - b).abs().max()) for a, b in zip(want, got))
frames = sum(int(w.shape[0]) for w in want)
print(json.dumps({"what": "Languasito2.inference, %d utterances of 20-60 phones, CPU fp32, %d threads" % (N, torch.get_num_threads()),
                  "reference_per_utterance_s": round(t_ref, 3), "batched_s": round(t_b, 3), "speedup": round(t_ref / t_b, 2),
                  "frames_total": frames, "max_abs_difference": err}))
