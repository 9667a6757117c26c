"""Batch > 1 for the reference's text frontend: ``Languasito2.inference`` is batch-1 by construction -
it squeezes the duration matrix and builds ``frame2phone`` with Python loops (cube/networks/modules.py:945-953), then
``_expand_i`` fills a numpy index array element by element and gathers with it (:1043-1054).  BASELINE configs[4] pushes 128
strings through that per call, and at one utterance at a time the frontend is 96 % of the step.

Here the SAME modules of a ``Languasito2`` instance (its embeddings, char CNNs, BiLSTMs, output layers - nothing is copied or
re-implemented, the weights stay where they are) are driven over a padded batch so that every utterance is computed as if alone:

* phone positions past an utterance's end are zero at the input of every conv (what its "same" zero padding sees at batch 1),
* every LSTM runs on a packed sequence (``pack_padded_sequence``), so the backward direction starts at the utterance's own end,
* durations -> frame index on the device: ``cumsum`` + ``searchsorted`` instead of the Python loops, one ``gather`` instead of the
  element-wise numpy index build.

Host-side glue in PyTorch (device memory, library kernels): the frontend is the reference's side of the vocoder boundary, this
module only removes its batch-1 restriction.  ``oracle/frontend_ref.py`` restates the batch-1 algorithm; tests/test_frontend_glue.py
checks both against the unmodified reference class when /root/reference is present.
"""
from typing import List, Optional, Sequence, Tuple

import torch
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

_ATTRS = ("_phon_emb_t", "_phon_emb_g", "_speaker_emb_t", "_speaker_emb_g", "_char_cnn_t", "_char_cnn_g", "_char_rnn_t",
          "_char_rnn_g", "_dur_rnn", "_dur_output", "_pitch_rnn", "_pitch_output", "_cond_rnn", "_cond_output", "_max_pitch")


def supports(lang) -> bool:
    """True when ``lang`` has the layout of the reference ``Languasito2`` (cube/networks/modules.py:825-914) WITHOUT external
    (word-level) conditioning - the case the batched path covers; anything else runs per utterance through ``lang.inference``."""
    return all(hasattr(lang, a) for a in _ATTRS) and not getattr(lang, "_use_cond", False) and getattr(lang, "_pframes", 1) == 1


def durations_to_frame_index(durs: torch.Tensor, n_phones: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``durs`` [B, P] (frames per phone, integer) -> (``idx`` [B, Fmax], ``n_frames`` [B]): ``idx[b, t]`` = the phone that frame t of
    utterance b belongs to - the ``frame2phone`` list the reference builds with two nested Python loops
    (cube/networks/modules.py:945-953), for all utterances at once and on the device of ``durs``.  Frames past an utterance's
    end repeat its last frame's phone, as ``_expand_i`` pads a batch (:1051-1053); phones past ``n_phones[b]`` count as 0 frames."""
    if durs.dim() != 2:
        raise ValueError("durs must be [B, P]")
    durs = durs.to(torch.int64).clamp_min(0)
    B, P = durs.shape
    if n_phones is not None:
        durs = durs * (torch.arange(P, device=durs.device)[None, :] < n_phones.to(durs.device)[:, None])
    cum = torch.cumsum(durs, dim=1)                                   # [B, P] frames up to and including phone p
    n_frames = cum[:, -1] if P > 0 else torch.zeros(B, dtype=torch.int64, device=durs.device)
    fmax = int(n_frames.max()) if B > 0 and P > 0 else 0
    t = torch.arange(fmax, device=durs.device)[None, :].expand(B, fmax).contiguous()
    idx = torch.searchsorted(cum, t, right=True)                      # first phone whose cumulative count exceeds t
    last = torch.searchsorted(cum, (n_frames - 1).clamp_min(0)[:, None], right=True)      # phone of the last real frame
    idx = torch.minimum(idx, last.expand(B, fmax)).clamp_max(max(P - 1, 0))
    return idx, n_frames


def expand_rows(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """``x`` [B, P, C], ``idx`` [B, F] -> [B, F, C] with out[b, t] = x[b, idx[b, t]]: the gather of ``Languasito2._expand_i``
    (cube/networks/modules.py:1043-1054) as one ``torch.gather``."""
    return torch.gather(x, 1, idx[:, :, None].expand(idx.shape[0], idx.shape[1], x.shape[2]))


def _lstm(rnn, x: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """A batch-first (Bi)LSTM over sequences of different lengths, each as if alone; rows past a sequence's end come back as zero."""
    T = x.shape[1]
    lens = lengths.clamp_min(1).cpu()                                 # pack_padded_sequence refuses empty sequences
    out, _ = rnn(pack_padded_sequence(x, lens, batch_first=True, enforce_sorted=False))
    out, _ = pad_packed_sequence(out, batch_first=True, total_length=T)
    return out * (torch.arange(T, device=x.device)[None, :] < lengths[:, None])[:, :, None]


def _char_path(emb, cnn, rnn, spk_emb, x_char, x_speaker, n_phones, pmask):
    """The phone-level trunk both halves of the frontend share (cube/networks/modules.py:917-931 and :967-981)."""
    h = emb(x_char) * pmask[:, :, None]
    h = h.permute(0, 2, 1)
    for layer in cnn:
        h = layer(h) * pmask[:, None, :]          # what lies past the end must be the conv's zero padding, not tanh(bias)
    h = _lstm(rnn, h.permute(0, 2, 1), n_phones)
    spk = spk_emb(x_speaker).repeat(1, h.shape[1], 1)
    return torch.cat([h, spk], dim=-1)


@torch.no_grad()
def languasito_inference_batch(lang, x_chars: Sequence[torch.Tensor], x_speakers: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """``Languasito2.inference`` (cube/networks/modules.py:1000-1008 -> ``_text_forward`` :916-962, ``_cond_forward`` :964-993) for
    many utterances in one pass.  ``x_chars[i]``: int64 [P_i] phone ids (0 = padding, as the reference's collate encodes them),
    ``x_speakers[i]``: int64 scalar / [1] speaker id.  Returns the conditionings [F_i, 80] (``F_i`` may be 0: an utterance whose
    predicted durations are all zero; ``Cubegan.inference`` replaces that by one zero frame, cube/networks/cubegan.py:78-80)."""
    if not supports(lang):
        raise ValueError("languasito_inference_batch: not a Languasito2 without external conditioning")
    dev = lang._dur_output.linear_layer.weight.device if hasattr(lang._dur_output, "linear_layer") else next(lang.parameters()).device
    B = len(x_chars)
    if B == 0:
        return []
    n_phones = torch.tensor([int(x.numel()) for x in x_chars], dtype=torch.int64, device=dev)
    P = int(n_phones.max())
    if int(n_phones.min()) < 1:
        raise ValueError("every utterance needs at least one phone")
    x_char = torch.zeros((B, P), dtype=torch.int64, device=dev)
    for i, x in enumerate(x_chars):
        x_char[i, : x.numel()] = x.reshape(-1).to(dev)
    x_speaker = torch.stack([torch.as_tensor(s).reshape(-1)[:1].to(dev) for s in x_speakers]).to(torch.int64)     # [B, 1]
    pmask = torch.arange(P, device=dev)[None, :] < n_phones[:, None]

    # ---- text half: durations, then pitch at frame rate (modules.py:916-962) ----
    trunk_t = _char_path(lang._phon_emb_t, lang._char_cnn_t, lang._char_rnn_t, lang._speaker_emb_t, x_char, x_speaker, n_phones, pmask)
    out_dur = lang._dur_output(_lstm(lang._dur_rnn, trunk_t, n_phones))
    durs = torch.argmax(out_dur, dim=-1) * pmask                       # :945
    idx, n_frames = durations_to_frame_index(durs)                     # :946-953, on the device
    F = idx.shape[1]
    if F == 0:
        return [torch.zeros((0, lang._cond_output.linear_layer.out_features), device=dev) for _ in range(B)]
    out_pitch = lang._pitch_output(_lstm(lang._pitch_rnn, expand_rows(trunk_t, idx), n_frames))
    vuv = torch.round(torch.sigmoid(out_pitch[:, :, 1]))               # :1003
    pitch = torch.sigmoid(out_pitch[:, :, 0]) * lang._max_pitch * vuv  # :1004

    # ---- conditioning half (modules.py:964-993) ----
    trunk_g = _char_path(lang._phon_emb_g, lang._char_cnn_g, lang._char_rnn_g, lang._speaker_emb_g, x_char, x_speaker, n_phones, pmask)
    h = torch.cat([expand_rows(trunk_g, idx), (pitch / lang._max_pitch)[:, :, None]], dim=-1)
    cond = lang._cond_output(_lstm(lang._cond_rnn, h, n_frames))       # [B, F, 80]
    nf = n_frames.tolist()
    return [cond[i, : nf[i]] for i in range(B)]
