"""Batch entry point: a list of variable-length mels -> a list of waveforms, sharded over the ranks
of one node when ``torch.distributed`` is initialised (one process per GPU, NCCL over NVLink).

The vocoder has no cross-utterance dependency, so the data path has NO collective: rank 0 scatters
mel shards (point-to-point sends) and gathers audio; every rank runs its shard independently
(SURVEY 8(e)).  Sharding is longest-processing-time-first over frame counts, batches inside a rank are
length-sorted so padding stays small, and each utterance is computed exactly as if run alone
(``n_frames`` masks, see include/cube_vocoder.h).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch


def lpt_shard(n_frames: Sequence[int], world: int) -> List[List[int]]:
    """Greedy LPT: utterances by frame count descending, each to the least-loaded rank.
    Returns, per rank, the utterance indices (longest first).  Deterministic."""
    order = sorted(range(len(n_frames)), key=lambda i: (-int(n_frames[i]), i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(n_frames[i])
    return out


def make_batches(idx: Sequence[int], n_frames: Sequence[int], max_batch: int,
                 max_frames: Optional[int] = None, batch_overhead_frames: int = 1600) -> List[List[int]]:
    """Cut a length-sorted (longest first) index list into consecutive batches of <= max_batch utterances whose padded
    size (count * longest) stays <= max_frames, minimising  sum(count * longest) + batch_overhead_frames * n_batches.
    Every kernel works on the padded rows (masked), so a batch that mixes 15-s and 2-s utterances - what an LPT shard of a
    ragged set looks like - spends half its time on padding (measured at N=8: 0.73 strong-scaling efficiency with one
    padded batch per rank), while every extra batch costs a launch sequence (61 launches: ~1.9 ms at the batch-1 latency
    floor = the time of ~1 600 frames of HiFi-GAN work on a B200).  Exact by dynamic programming over the cut points (the list is sorted, so the longest of a
    batch is its first item); O(n * max_batch)."""
    idx = list(idx)
    n = len(idx)
    if n == 0:
        return []
    f = [int(n_frames[i]) for i in idx]
    INF = float("inf")
    best = [0.0] + [INF] * n           # best[j] = minimal cost of batching the first j items
    cut = [0] * (n + 1)
    for j in range(1, n + 1):
        for i in range(max(0, j - max_batch), j):       # batch = items i .. j-1, longest = f[i]
            padded = (j - i) * f[i]
            if max_frames and padded > max_frames and j - i > 1:
                continue
            c = best[i] + padded + batch_overhead_frames
            if c < best[j]:
                best[j], cut[j] = c, i
    out: List[List[int]] = []
    j = n
    while j > 0:
        out.append(idx[cut[j]:j])
        j = cut[j]
    return out[::-1]


def pad_mels(mels: Sequence[torch.Tensor], pad_value: float = 0.0) -> torch.Tensor:
    F = max(int(m.shape[-1]) for m in mels)
    out = torch.full((len(mels), mels[0].shape[0], F), pad_value, dtype=torch.float32, device=mels[0].device)
    for b, m in enumerate(mels):
        out[b, :, : m.shape[-1]] = m
    return out


def _batch_out(wav: torch.Tensor, batch: Sequence[int], nf: Sequence[int], out_len: Callable[[int], int], res: list) -> None:
    for b, i in enumerate(batch):
        res[i] = wav[b].reshape(-1)[: out_len(nf[i])]


def _run_local(vocode: Callable, mels: Sequence[torch.Tensor], zs: Optional[Sequence[torch.Tensor]],
               out_len: Callable[[int], int], device, max_batch: int, max_frames: Optional[int]) -> List[torch.Tensor]:
    nf = [int(m.shape[-1]) for m in mels]
    order = sorted(range(len(mels)), key=lambda i: (-nf[i], i))
    res: List[Optional[torch.Tensor]] = [None] * len(mels)
    for batch in make_batches(order, nf, max_batch, max_frames):
        mb = pad_mels([mels[i].to(device, non_blocking=True) for i in batch])
        frames = [nf[i] for i in batch]
        if zs is not None:
            wav = vocode(mb, _pad_z([zs[i] for i in batch], out_len(int(mb.shape[-1])), device), frames)
        else:
            wav = vocode(mb, frames)
        _batch_out(wav, batch, nf, out_len, res)
    return res  # type: ignore


def _pad_z(zs: Sequence[torch.Tensor], T: int, device) -> torch.Tensor:
    zb = torch.zeros(len(zs), 1, T, dtype=torch.float32, device=device)
    for b, z in enumerate(zs):
        zb[b, 0, : z.numel()] = z.to(device).reshape(-1)
    return zb


def synthesize(vocoder, mels: Optional[Sequence[torch.Tensor]], zs: Optional[Sequence[torch.Tensor]] = None,
               device=None, max_batch: int = 64, max_frames: Optional[int] = None, group=None,
               vocode: Optional[Callable] = None, out_len: Optional[Callable[[int], int]] = None,
               stats: Optional[dict] = None) -> Optional[List[torch.Tensor]]:
    """mels: list of [80, F_i] float32 tensors (on rank 0; other ranks pass None when distributed).
    Returns the list of waveforms [T_i] on rank 0 (None elsewhere), in input order.

    ``vocoder`` is a CubeGenerator / ParallelWaveNetVocoder; ``vocode`` / ``out_len`` may be given
    instead (any callable with the same contract - the CPU tests of the sharding logic do that).

    Distributed data path (no collective inside the network, point-to-point only):
      * every rank derives the same plan from the broadcast frame counts: LPT shards, then length-sorted batches;
      * rank 0 pads ONE block per (rank, batch) and posts all sends in one group - the blocks are what the receiving
        rank feeds to its vocoder as they are (no re-padding on the receiver);
      * a remote rank sends the output tensor of each batch as the vocoder produced it (no staging copy, one group);
        rank 0 receives every block into its own buffer and returns views of them (no per-utterance copy);
      * on NCCL nothing blocks the host: every wait() is a stream dependency.
    ``stats`` (a dict) receives host-side wall times per phase and the bytes moved, for the bench."""
    import time
    import torch.distributed as dist
    if out_len is None:
        out_len = vocoder.out_len
    if zs is None and vocode is None and _needs_noise(vocoder) and mels is not None:
        # the IAF student needs z ~ N(0,1) per utterance; callers that want a replayable run pass `zs`
        zs = [torch.randn(out_len(int(m.shape[-1]))) for m in mels]
    if vocode is None:
        vocode = (lambda m, z, f: vocoder(m, z, f)) if zs is not None or _needs_noise(vocoder) else (lambda m, f: vocoder(m, f))
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if device is None:
        device = getattr(vocoder, "device", None) or (mels[0].device if mels else torch.device("cpu"))
    if not distributed:
        return _run_local(vocode, mels, zs, out_len, device, max_batch, max_frames)

    t0 = time.perf_counter()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    comm_dev = device if dist.get_backend(group) == "nccl" else torch.device("cpu")
    # ---- frame counts to every rank: two small tensor broadcasts (header, then the counts) ----
    hdr = torch.zeros(3, dtype=torch.int64, device=comm_dev)
    if rank == 0:
        hdr = torch.tensor([len(mels), int(mels[0].shape[0]) if mels else 0, int(zs is not None)], dtype=torch.int64, device=comm_dev)
    dist.broadcast(hdr, src=0, group=group)
    n_utt, C, has_z = (int(v) for v in hdr.tolist())
    nft = torch.zeros(max(n_utt, 1), dtype=torch.int64, device=comm_dev)
    if rank == 0 and n_utt:
        nft = torch.tensor([int(m.shape[-1]) for m in mels], dtype=torch.int64, device=comm_dev)
    dist.broadcast(nft, src=0, group=group)
    nf = nft.tolist()[:n_utt]
    plan = lpt_shard(nf, world)
    batches = [make_batches(plan[r], nf, max_batch, max_frames) for r in range(world)]     # plan[r] is length-sorted
    t1 = time.perf_counter()

    # ---- scatter: one padded block per (rank, batch); all sends / receives of a rank in ONE group ----
    ops, blocks, zblocks = [], {}, {}
    sent = 0
    if rank == 0:
        for r in range(1, world):
            for k, batch in enumerate(batches[r]):
                blk = pad_mels([mels[i].to(comm_dev, non_blocking=True) for i in batch])
                blocks[(r, k)] = blk                          # keep alive until the group completes
                ops.append(dist.P2POp(dist.isend, blk, r, group))
                sent += blk.numel() * 4
                if has_z:
                    zb = _pad_z([zs[i] for i in batch], out_len(int(blk.shape[-1])), comm_dev)
                    zblocks[(r, k)] = zb
                    ops.append(dist.P2POp(dist.isend, zb, r, group))
                    sent += zb.numel() * 4
    else:
        for k, batch in enumerate(batches[rank]):
            Fm = max(nf[i] for i in batch)
            blocks[k] = torch.empty(len(batch), C, Fm, dtype=torch.float32, device=comm_dev)
            ops.append(dist.P2POp(dist.irecv, blocks[k], 0, group))
            if has_z:
                zblocks[k] = torch.empty(len(batch), 1, out_len(Fm), dtype=torch.float32, device=comm_dev)
                ops.append(dist.P2POp(dist.irecv, zblocks[k], 0, group))
    scatter_work = dist.batch_isend_irecv(ops) if ops else []
    t2 = time.perf_counter()

    for w in scatter_work:            # NCCL: makes the current stream wait (not the host); gloo: completes the transfers
        w.wait()
    t3 = time.perf_counter()

    # ---- local compute ----
    out: List[Optional[torch.Tensor]] = [None] * n_utt
    if rank == 0:
        my = plan[0]
        if my:
            local = _run_local(vocode, [mels[i] for i in my], [zs[i] for i in my] if has_z else None, out_len, device, max_batch, max_frames)
            for j, i in enumerate(my):
                out[i] = local[j]
    else:
        sends, keep = [], []
        for k, batch in enumerate(batches[rank]):
            frames = [nf[i] for i in batch]
            mb = blocks[k].to(device, non_blocking=True)
            wav = vocode(mb, zblocks[k].to(device, non_blocking=True), frames) if has_z else vocode(mb, frames)
            wav = wav.reshape(len(batch), 1, -1).to(comm_dev)
            keep.append(wav)
            sends.append(dist.P2POp(dist.isend, wav, 0, group))
        for w in (dist.batch_isend_irecv(sends) if sends else []):
            w.wait()
    t4 = time.perf_counter()

    # ---- gather: posted AFTER rank 0's own compute is enqueued.  The receives are stream-ordered behind it, so the
    # NCCL kernels never sit on SMs while the persistent compute kernels (one CTA per SM, ~220 KB of shared memory each)
    # run - a co-resident NCCL CTA would push one compute CTA into a second wave and double every kernel's tail.  The
    # remote ranks finish at about the same time (LPT), and 4 B/sample over NVLink is < 1 % of the compute time.
    recvd = 0
    if rank == 0:
        gops, rbuf = [], {}
        for r in range(1, world):
            for k, batch in enumerate(batches[r]):
                Tm = out_len(max(nf[i] for i in batch))
                rbuf[(r, k)] = torch.empty(len(batch), 1, Tm, dtype=torch.float32, device=comm_dev)
                gops.append(dist.P2POp(dist.irecv, rbuf[(r, k)], r, group))
                recvd += rbuf[(r, k)].numel() * 4
        for w in (dist.batch_isend_irecv(gops) if gops else []):
            w.wait()
        for (r, k), blk in rbuf.items():
            _batch_out(blk, batches[r][k], nf, out_len, out)
    if stats is not None:
        stats.update({"plan_s": t1 - t0, "scatter_issue_s": t2 - t1, "scatter_wait_s": t3 - t2, "compute_issue_s": t4 - t3,
                      "gather_wait_s": time.perf_counter() - t4, "scatter_bytes": sent, "gather_bytes": recvd,
                      "batches_per_rank": [len(b) for b in batches]})
    return out if rank == 0 else None  # type: ignore


def cubegan_inference_batch(model, Xs: Sequence[dict], max_batch: int = 64, max_frames: Optional[int] = None,
                            int16: bool = False, group=None, frontend_batch: int = 32) -> Optional[List[torch.Tensor]]:
    """BASELINE configs[4] glue: many utterances through the reference ``Cubegan`` whose ``_generator`` is a CubeGenerator
    (``install_into_cubegan``).  The reference frontend is batch-1 by construction (``Languasito2.inference`` squeezes the
    duration matrix, cube/networks/modules.py:945-953).  When ``model._languasito`` has the reference layout without external
    conditioning, its own modules are driven over padded batches of ``frontend_batch`` utterances, each computed as if alone
    (tts_cube_b200/frontend.py; ``frontend_batch=0`` disables this); otherwise it runs per utterance exactly as
    ``Cubegan.inference`` runs it (cube/networks/cubegan.py:74-81: optional HF conditioning, ``_languasito.inference``, the
    empty-utterance guard).  The vocoder then takes ALL conditionings in length-sorted batches (each utterance computed as if
    alone) - sharded over the ranks when torch.distributed is up (rank 0 holds ``Xs``).
    Returns the waveforms [T_i] (float32, or int16 with the ``*32767`` epilogue of cube/api.py:64-65) on rank 0."""
    from . import frontend as FE
    from .heads import wav_to_int16
    conds = None
    if Xs is not None:
        conds = []
        lang = model._languasito
        hf = getattr(model, "_hf", None)
        with torch.no_grad():
            if frontend_batch > 0 and hf is None and FE.supports(lang):
                for i0 in range(0, len(Xs), frontend_batch):
                    part = Xs[i0:i0 + frontend_batch]
                    cs = FE.languasito_inference_batch(lang, [X["x_char"][0] for X in part], [X["x_speaker"][0] for X in part])
                    for c in cs:                                              # [F, 80]
                        if c.shape[0] == 0:                                   # the empty-utterance guard of cubegan.py:78-80
                            c = torch.zeros((1, c.shape[1]), device=c.device)
                        conds.append(c.t())
            else:
                for X in Xs:
                    hf_cond = hf(X["x_tok_ids"])["last_hidden_state"] if hf is not None else None
                    c = lang.inference(X, hf_cond=hf_cond)                    # [1, F, 80]
                    if c.shape[1] == 0:
                        c = torch.zeros((c.shape[0], 1, c.shape[2]), device=c.device)
                    conds.append(c[0].t())                                   # [80, F] view; padded / copied per batch
    wavs = synthesize(model._generator, conds, max_batch=max_batch, max_frames=max_frames, group=group)
    if wavs is not None and int16:
        wavs = [wav_to_int16(w) for w in wavs]
    return wavs


def _needs_noise(vocoder) -> bool:
    return type(vocoder).__name__ == "ParallelWaveNetVocoder"
