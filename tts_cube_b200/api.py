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
                 max_frames: Optional[int] = None) -> List[List[int]]:
    """Cut a length-sorted index list into batches of <= max_batch utterances whose padded size
    (count * longest) stays <= max_frames."""
    batches: List[List[int]] = []
    cur: List[int] = []
    for i in idx:
        longest = int(n_frames[cur[0]]) if cur else int(n_frames[i])
        if cur and (len(cur) >= max_batch or (max_frames and (len(cur) + 1) * longest > max_frames)):
            batches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def pad_mels(mels: Sequence[torch.Tensor], pad_value: float = 0.0) -> torch.Tensor:
    F = max(int(m.shape[-1]) for m in mels)
    out = torch.full((len(mels), mels[0].shape[0], F), pad_value, dtype=torch.float32, device=mels[0].device)
    for b, m in enumerate(mels):
        out[b, :, : m.shape[-1]] = m
    return out


def _run_local(vocode: Callable, mels: Sequence[torch.Tensor], zs: Optional[Sequence[torch.Tensor]],
               out_len: Callable[[int], int], device, max_batch: int, max_frames: Optional[int]) -> List[torch.Tensor]:
    nf = [int(m.shape[-1]) for m in mels]
    order = sorted(range(len(mels)), key=lambda i: (-nf[i], i))
    res: List[Optional[torch.Tensor]] = [None] * len(mels)
    for batch in make_batches(order, nf, max_batch, max_frames):
        mb = pad_mels([mels[i].to(device, non_blocking=True) for i in batch])
        frames = [nf[i] for i in batch]
        if zs is not None:
            T = out_len(int(mb.shape[-1]))
            zb = torch.zeros(len(batch), 1, T, dtype=torch.float32, device=device)
            for b, i in enumerate(batch):
                zb[b, 0, : zs[i].shape[-1]] = zs[i].to(device).reshape(-1)
            wav = vocode(mb, zb, frames)
        else:
            wav = vocode(mb, frames)
        for b, i in enumerate(batch):
            res[i] = wav[b, 0, : out_len(nf[i])]
    return res  # type: ignore


def synthesize(vocoder, mels: Optional[Sequence[torch.Tensor]], zs: Optional[Sequence[torch.Tensor]] = None,
               device=None, max_batch: int = 64, max_frames: Optional[int] = None, group=None,
               vocode: Optional[Callable] = None, out_len: Optional[Callable[[int], int]] = None) -> Optional[List[torch.Tensor]]:
    """mels: list of [80, F_i] float32 tensors (on rank 0; other ranks pass None when distributed).
    Returns the list of waveforms [T_i] on rank 0 (None elsewhere), in input order.

    ``vocoder`` is a CubeGenerator / ParallelWaveNetVocoder; ``vocode`` / ``out_len`` may be given
    instead (any callable with the same contract - the CPU tests of the sharding logic do that)."""
    import torch.distributed as dist
    if vocode is None:
        vocode = (lambda m, z, f: vocoder(m, z, f)) if zs is not None or _needs_noise(vocoder) else (lambda m, f: vocoder(m, f))
    if out_len is None:
        out_len = vocoder.out_len
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if device is None:
        device = getattr(vocoder, "device", None) or (mels[0].device if mels else torch.device("cpu"))
    if not distributed:
        return _run_local(vocode, mels, zs, out_len, device, max_batch, max_frames)

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    meta = [None]
    if rank == 0:
        meta = [dict(nf=[int(m.shape[-1]) for m in mels], C=int(mels[0].shape[0]), z=zs is not None)]
    dist.broadcast_object_list(meta, src=0, group=group)
    nf, C, has_z = meta[0]["nf"], meta[0]["C"], meta[0]["z"]
    plan = lpt_shard(nf, world)
    comm_dev = device if dist.get_backend(group) == "nccl" else torch.device("cpu")

    # ---- scatter: rank 0 -> r, one padded block per rank ----
    mine = plan[rank]
    my_mels: List[torch.Tensor] = []
    my_zs: Optional[List[torch.Tensor]] = [] if has_z else None
    ops, bufs = [], {}
    if rank == 0:
        for r in range(1, world):
            if not plan[r]:
                continue
            blk = pad_mels([mels[i].to(comm_dev) for i in plan[r]])
            ops.append(dist.P2POp(dist.isend, blk, r, group))
            if has_z:
                T = out_len(int(blk.shape[-1]))
                zb = torch.zeros(len(plan[r]), T, dtype=torch.float32, device=comm_dev)
                for b, i in enumerate(plan[r]):
                    zb[b, : zs[i].numel()] = zs[i].to(comm_dev).reshape(-1)
                ops.append(dist.P2POp(dist.isend, zb, r, group))
        my_mels = [mels[i] for i in mine]
        if has_z:
            my_zs = [zs[i] for i in mine]
    elif mine:
        Fm = max(nf[i] for i in mine)
        bufs["mel"] = torch.empty(len(mine), C, Fm, dtype=torch.float32, device=comm_dev)
        ops.append(dist.P2POp(dist.irecv, bufs["mel"], 0, group))
        if has_z:
            bufs["z"] = torch.empty(len(mine), out_len(Fm), dtype=torch.float32, device=comm_dev)
            ops.append(dist.P2POp(dist.irecv, bufs["z"], 0, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if rank != 0 and mine:
        my_mels = [bufs["mel"][b, :, : nf[i]] for b, i in enumerate(mine)]
        if has_z:
            my_zs = [bufs["z"][b, : out_len(nf[i])] for b, i in enumerate(mine)]

    # ---- local compute ----
    wavs = _run_local(vocode, my_mels, my_zs, out_len, device, max_batch, max_frames) if mine else []

    # ---- gather: r -> rank 0, one padded block per rank ----
    ops = []
    if rank == 0:
        rbuf = {}
        for r in range(1, world):
            if plan[r]:
                Tm = out_len(max(nf[i] for i in plan[r]))
                rbuf[r] = torch.empty(len(plan[r]), Tm, dtype=torch.float32, device=comm_dev)
                ops.append(dist.P2POp(dist.irecv, rbuf[r], r, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        out: List[Optional[torch.Tensor]] = [None] * len(nf)
        for b, i in enumerate(mine):
            out[i] = wavs[b]
        for r, blk in rbuf.items():
            for b, i in enumerate(plan[r]):
                out[i] = blk[b, : out_len(nf[i])]
        return out  # type: ignore
    if mine:
        Tm = out_len(max(nf[i] for i in mine))
        blk = torch.zeros(len(mine), Tm, dtype=torch.float32, device=comm_dev)
        for b in range(len(mine)):
            blk[b, : wavs[b].numel()] = wavs[b].to(comm_dev)
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, blk, 0, group)]):
            w.wait()
    return None


def _needs_noise(vocoder) -> bool:
    return type(vocoder).__name__ == "ParallelWaveNetVocoder"
