"""TEST INFRASTRUCTURE (CPU oracle): the batch-1 algorithm of the reference's ``Languasito2.inference`` restated step by step,
for a module with the reference's attribute layout and NO external conditioning (``_use_cond`` False).  Only tests/ may import this;
the product (tts_cube_b200/frontend.py) drives the same modules over a padded batch and is checked against it.

Pinned by tests/test_frontend_glue.py::test_restatement_matches_reference_class, which runs the UNMODIFIED reference class
(cube/networks/modules.py:825-1008, imported from /root/reference with the harness stubs of oracle/make_cubegan_golden.py) on the
same inputs - in the build container only: the class is 13.5 M seeded-random parameters, too large for a committed fixture.
"""
from typing import List

import torch


def frame2phone(durs) -> List[int]:
    """cube/networks/modules.py:946-952: phone index of every frame, by the reference's two nested loops."""
    out, phon_index = [], 0
    for dur in durs:
        for _ in range(int(dur)):
            out.append(phon_index)
        phon_index += 1
    return out


def expand_i(x: torch.Tensor, alignments: List[List[int]]) -> torch.Tensor:
    """cube/networks/modules.py:1043-1054 (``_pframes`` = 1): rows of x picked by the alignment, short utterances padded with their
    last alignment entry."""
    m_size = max(len(a) for a in alignments)
    idx = torch.zeros((x.shape[0], m_size), dtype=torch.int64)
    for ii, a in enumerate(alignments):
        for jj, v in enumerate(a):
            idx[ii, jj] = v
        for jj in range(len(a), m_size):
            idx[ii, jj] = a[-1]
    bi = torch.arange(x.shape[0])[:, None].expand(-1, m_size)
    return x[bi, idx.to(x.device)]


def _trunk(emb, cnn, rnn, spk_emb, x_char, x_speaker):
    hidden = emb(x_char).permute(0, 2, 1)              # :920-925 / :970-975
    for layer in cnn:
        hidden = layer(hidden)
    hidden, _ = rnn(hidden.permute(0, 2, 1))
    spk = spk_emb(x_speaker).repeat(1, hidden.shape[1], 1)
    return torch.cat([hidden, spk], dim=-1)


@torch.no_grad()
def languasito_inference(lang, x_char: torch.Tensor, x_speaker: torch.Tensor) -> torch.Tensor:
    """x_char int64 [1, P], x_speaker int64 [1, 1] -> conditioning [1, F, 80]  (modules.py:1000-1008)."""
    t = _trunk(lang._phon_emb_t, lang._char_cnn_t, lang._char_rnn_t, lang._speaker_emb_t, x_char, x_speaker)
    hidden_dur, _ = lang._dur_rnn(t)
    durs = torch.argmax(lang._dur_output(hidden_dur), dim=-1).reshape(-1).tolist()      # :945 (squeeze: batch 1)
    f2p = [frame2phone(durs)]
    if len(f2p[0]) == 0:
        return torch.zeros((1, 0, lang._cond_output.linear_layer.out_features))
    hidden_pitch, _ = lang._pitch_rnn(expand_i(t, f2p))
    out = lang._pitch_output(hidden_pitch)
    vuv = torch.round(torch.sigmoid(out[:, :, 1]))                                       # :1003
    pitch = torch.sigmoid(out[:, :, 0]) * lang._max_pitch * vuv                          # :1004
    g = _trunk(lang._phon_emb_g, lang._char_cnn_g, lang._char_rnn_g, lang._speaker_emb_g, x_char, x_speaker)
    h = expand_i(g, f2p)
    m = min(h.shape[1], pitch.shape[1])                                                  # :988
    h = torch.cat([h[:, :m], (pitch.unsqueeze(2) / lang._max_pitch)[:, :m]], dim=-1)
    h, _ = lang._cond_rnn(h)
    return lang._cond_output(h)
