"""Batch > 1 glue for the reference's text frontend (tts_cube_b200/frontend.py): CPU tests.
Index building is integer work -> bit-exact against the reference's loops; the batched frontend is compared with the batch-1
algorithm (oracle/frontend_ref.py) at 1e-5, and - when /root/reference is present, i.e. in the build container - both are compared
with the UNMODIFIED reference class on its own seeded-random 13.5 M-parameter instance."""
import os
import sys

import numpy as np
import pytest
import torch
from torch import nn

from oracle import frontend_ref as FR
from tts_cube_b200 import frontend as FE

REF = os.environ.get("CUBE_REFERENCE", "/root/reference")


# ------------------------------------------------ index building ------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_frame_index_matches_reference_loops(seed):
    g = torch.Generator().manual_seed(seed)
    B, P = 7, 13
    durs = torch.randint(0, 6, (B, P), generator=g)
    durs[2] = 0                                            # an utterance without a single frame
    durs[3, 5:] = 0                                        # trailing zero-length phones
    n_phones = torch.tensor([13, 1, 13, 13, 4, 9, 13])
    idx, nf = FE.durations_to_frame_index(durs, n_phones)
    lists = [FR.frame2phone(durs[b, : int(n_phones[b])].tolist()) for b in range(B)]
    assert nf.tolist() == [len(a) for a in lists]
    for b, a in enumerate(lists):
        assert idx[b, : len(a)].tolist() == a
        if a:                                              # padding repeats the last frame's phone (modules.py:1051-1053)
            assert (idx[b, len(a):] == a[-1]).all()
    # the gather equals the reference's element-wise index build, for the utterances that have frames
    x = torch.randn(B, P, 5, generator=g)
    keep = [b for b in range(B) if lists[b]]
    want = FR.expand_i(x[keep], [lists[b] for b in keep])
    got = FE.expand_rows(x, idx)[keep]
    assert torch.equal(got[:, : want.shape[1]], want)


def test_frame_index_edge_cases():
    idx, nf = FE.durations_to_frame_index(torch.zeros((3, 4), dtype=torch.int64))
    assert idx.shape == (3, 0) and nf.tolist() == [0, 0, 0]
    idx, nf = FE.durations_to_frame_index(torch.tensor([[2, 0, 3]]))
    assert idx.tolist() == [[0, 0, 2, 2, 2]] and nf.tolist() == [5]
    with pytest.raises(ValueError):
        FE.durations_to_frame_index(torch.zeros(4, dtype=torch.int64))


# ------------------------------------------------ a small module with the reference's layout ------------------------------------------------
class _LinearNorm(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.linear_layer = nn.Linear(i, o)

    def forward(self, x):
        return self.linear_layer(x)


class TinyLanguasito(nn.Module):
    """Same attribute names, layer types and wiring as cube/networks/modules.py:825-914, at toy sizes (and an output bias on the
    duration head that makes durations 0..4 all occur)."""

    def __init__(self, n_phones=12, n_speakers=3, E=8, S=4, C=10, R=6, D=5, max_dur=4, max_pitch=300):
        super().__init__()
        self._pframes, self._use_cond, self._max_pitch = 1, False, max_pitch
        for tag in ("t", "g"):
            setattr(self, f"_phon_emb_{tag}", nn.Embedding(n_phones + 1, E, padding_idx=0))
            setattr(self, f"_speaker_emb_{tag}", nn.Embedding(n_speakers + 1, S, padding_idx=0))
            cnn, inp = [], E
            for _ in range(2):
                cnn += [nn.Conv1d(inp, C, 5, padding=2), nn.Tanh()]
                inp = C
            setattr(self, f"_char_cnn_{tag}", nn.ModuleList(cnn))
            setattr(self, f"_char_rnn_{tag}", nn.LSTM(C, R, num_layers=2, bidirectional=True, batch_first=True))
        self._dur_rnn = nn.LSTM(2 * R + S, D, num_layers=2, bidirectional=True, batch_first=True)
        self._dur_output = _LinearNorm(2 * D, max_dur + 1)
        self._pitch_rnn = nn.LSTM(2 * R + S, D, num_layers=2, bidirectional=True, batch_first=True)
        self._pitch_output = _LinearNorm(2 * D, 2)
        self._cond_rnn = nn.LSTM(2 * R + S + 1, D, num_layers=2, bidirectional=True, batch_first=True)
        self._cond_output = _LinearNorm(2 * D, 80)


def _utterances(seed, n, n_phones, lo=1, hi=15):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randint(1, n_phones + 1, (int(torch.randint(lo, hi + 1, (1,), generator=g)),), generator=g) for _ in range(n)]
    sp = [torch.randint(1, 3, (1,), generator=g) for _ in range(n)]
    return xs, sp


def test_batched_frontend_equals_batch1_algorithm():
    torch.manual_seed(5)
    lang = TinyLanguasito().eval()
    with torch.no_grad():
        lang._dur_output.linear_layer.weight.mul_(8.0)      # spread the duration logits: every class 0..4 gets picked
    assert FE.supports(lang)
    xs, sp = _utterances(3, 9, 12)
    xs[4] = xs[4][:1]                                       # a one-phone utterance
    got = FE.languasito_inference_batch(lang, xs, sp)
    frames = []
    for i in range(len(xs)):
        want = FR.languasito_inference(lang, xs[i][None], sp[i][None])[0]
        assert got[i].shape == want.shape, (i, got[i].shape, want.shape)
        frames.append(want.shape[0])
        if want.numel():
            assert float((got[i] - want).abs().max()) <= 1e-5
    assert len(set(frames)) > 3 and max(frames) > 0         # the batch really was ragged
    # batch composition does not matter: every utterance is computed as if alone
    again = FE.languasito_inference_batch(lang, xs[::-1], sp[::-1])[::-1]
    for a, b in zip(got, again):
        assert a.shape == b.shape and (a.numel() == 0 or float((a - b).abs().max()) <= 1e-5)


def test_unsupported_layouts_are_refused():
    lang = TinyLanguasito()
    lang._use_cond = True
    assert not FE.supports(lang)
    with pytest.raises(ValueError):
        FE.languasito_inference_batch(lang, [torch.tensor([1, 2])], [torch.tensor([1])])
    assert not FE.supports(nn.Linear(2, 2))


# ------------------------------------------------ against the unmodified reference class ------------------------------------------------
@pytest.fixture(scope="module")
def reference_cubegan():
    if not os.path.isdir(os.path.join(REF, "cube")):
        pytest.skip("reference checkout not present (GPU box): the live comparison runs in the build container")
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_cubegan_golden", os.path.join(os.path.dirname(here), "oracle", "make_cubegan_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    mk._stubs()
    cwd = os.getcwd()
    sys.path[:0] = [REF, os.path.join(REF, "hifigan")]
    os.chdir(REF)                                           # Cubegan opens 'hifigan/config_v1.json' by relative path
    try:
        from cube.io_utils.io_cubegan import CubeganCollate, CubeganEncodings
        from cube.networks.cubegan import Cubegan
        enc = CubeganEncodings()
        enc.phon2int = {p: i for i, p in enumerate(mk.ALPHABET)}
        enc.speaker2int = {"neb": 0, "anca": 1}
        enc.max_pitch, enc.max_duration = 400, 100
        torch.manual_seed(1234)
        model = Cubegan(enc, conditioning=None, train=False).eval()
    finally:
        os.chdir(cwd)
    collate = CubeganCollate(enc, conditioning_type=None)

    def make_x(phones, speaker):
        rez = {"meta": {"phones": list(phones), "phon2word": [0] * len(phones), "words": ["w"], "speaker": speaker,
                        "words_left": [], "words_right": [], "frame2phon": [0] * 100},
               "pitch": np.zeros(100), "mgc": np.zeros((100, 80))}
        return collate.collate_fn([rez])
    return model, make_x, mk.ALPHABET


# (the reference itself cannot take a ONE-phone utterance: its `.squeeze()` leaves a 0-d duration array, modules.py:945-949)
TEXTS = ["hello", "ab", "the_quick_brown_fox,jumps.", "zz", "over_the_lazy_dog", "abcabcabcabcabcabcabcabcabcabcabc."]


def test_restatement_matches_reference_class(reference_cubegan):
    model, make_x, _ = reference_cubegan
    lang = model._languasito
    assert FE.supports(lang)
    for i, text in enumerate(TEXTS[:3]):
        X = make_x(text, "neb" if i % 2 == 0 else "anca")
        x_char, x_speaker = X["x_char"].clone(), X["x_speaker"].clone()
        with torch.no_grad():
            want = lang.inference(X)                        # the reference, unmodified (cube/networks/modules.py:1000-1008)
        got = FR.languasito_inference(lang, x_char, x_speaker)
        assert got.shape == want.shape and want.shape[1] > 0
        assert float((got - want).abs().max()) <= 1e-6


def test_batched_frontend_matches_reference_class(reference_cubegan):
    model, make_x, _ = reference_cubegan
    lang = model._languasito
    Xs = [make_x(t, "neb" if i % 2 == 0 else "anca") for i, t in enumerate(TEXTS)]
    xs = [X["x_char"][0].clone() for X in Xs]
    sp = [X["x_speaker"][0].clone() for X in Xs]
    with torch.no_grad():
        want = [lang.inference(X)[0] for X in Xs]
    got = FE.languasito_inference_batch(lang, xs, sp)
    assert len({w.shape[0] for w in want}) > 2
    for w, g_ in zip(want, got):
        assert g_.shape == w.shape
        assert float((g_ - w).abs().max()) <= 1e-5


# ------------------------------------------------ the configs[4] glue uses it ------------------------------------------------
def test_cubegan_inference_batch_uses_the_batched_frontend():
    """``cubegan_inference_batch`` with a Languasito2-shaped frontend: the batched path and the per-utterance path (the reference's
    own ``inference``, here the batch-1 restatement) hand the same conditionings to the vocoder."""
    from tts_cube_b200.api import cubegan_inference_batch

    class Vocoder:                                          # stands in for CubeGenerator: [B, 80, F] -> [B, 1, 4 F]
        device = torch.device("cpu")
        seen = []

        def out_len(self, f):
            return 4 * f

        def __call__(self, mel, frames):
            self.seen.append(tuple(mel.shape))
            return mel.mean(1, keepdim=True).repeat_interleave(4, dim=2)

    class Lang(TinyLanguasito):
        calls = 0

        def inference(self, X, hf_cond=None):               # what Cubegan.inference calls at batch 1
            Lang.calls += 1
            return FR.languasito_inference(self, X["x_char"], X["x_speaker"])

    torch.manual_seed(11)
    model = nn.Module()
    model._languasito = Lang().eval()
    with torch.no_grad():
        model._languasito._dur_output.linear_layer.weight.mul_(8.0)
    model._generator = Vocoder()
    xs, sp = _utterances(7, 6, 12, lo=2)
    Xs = [{"x_char": x[None], "x_speaker": s[None]} for x, s in zip(xs, sp)]
    batched = cubegan_inference_batch(model, Xs, max_batch=4, frontend_batch=4)
    assert Lang.calls == 0                                  # the per-utterance entry point was not needed
    single = cubegan_inference_batch(model, Xs, max_batch=4, frontend_batch=0)
    assert Lang.calls == len(Xs)
    assert len(batched) == len(single) == len(Xs)
    for a, b in zip(batched, single):
        assert a.shape == b.shape and a.numel() > 0
        assert float((a - b).abs().max()) <= 1e-5


def test_utterances_without_frames():
    """Predicted durations all zero: the reference returns a [1, 0, 80] conditioning (``_expand_i`` on an empty alignment) and
    ``Cubegan.inference`` replaces it by one zero frame (cube/networks/cubegan.py:78-80); the batched path returns [0, 80]."""
    torch.manual_seed(2)
    lang = TinyLanguasito().eval()
    with torch.no_grad():                                   # the duration head always answers "0 frames"
        lang._dur_output.linear_layer.weight.zero_()
        lang._dur_output.linear_layer.bias.copy_(torch.tensor([5.0, 0, 0, 0, 0]))
    xs, sp = _utterances(9, 3, 12)
    got = FE.languasito_inference_batch(lang, xs, sp)
    assert [tuple(g.shape) for g in got] == [(0, 80)] * 3
    assert FR.languasito_inference(lang, xs[0][None], sp[0][None]).shape == (1, 0, 80)
    # one silent utterance inside a batch of speaking ones: rows of the others are unaffected
    with torch.no_grad():
        lang._dur_output.linear_layer.bias.copy_(torch.tensor([0.0, 0, 5.0, 0, 0]))      # 2 frames per phone
        lang._phon_emb_t.weight[3].zero_()
    base = FE.languasito_inference_batch(lang, xs, sp)
    assert [g.shape[0] for g in base] == [2 * x.numel() for x in xs]
    import tts_cube_b200 as cube
    assert cube.languasito_inference_batch is FE.languasito_inference_batch
