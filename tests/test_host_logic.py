"""Host-side logic and the C-ABI surface - CPU only (no compute calls without a GPU)."""
import ctypes
import os
import re
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_library_loads_and_exports_every_declared_symbol():
    from tts_cube_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cube_vocoder.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cube_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    l = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(l, name), f"{name} declared in include/cube_vocoder.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert "sm_100a" in _lib.build_info()


def test_config_struct_matches_header_size():
    from tts_cube_b200 import _lib
    # 3 + 2 + 8 + 8 + 2 + 8 + 8 + 64 + 1 + 8 + 3 + 2 + 2 + 1 + 4 int32 fields
    assert ctypes.sizeof(_lib.VocConfig) == 4 * (3 + 2 + 16 + 2 + 16 + 64 + 1 + 8 + 3 + 2 + 2 + 1 + 4 + 6)
    assert ctypes.sizeof(_lib.MelConfig) == 4 * 12      # cube_mel_config: 8 int32 + 4 float


def test_mel_front_end_host_logic():
    """filter bank = the torchaudio/librosa Slaney bank stored with the reference-made golden; frame-count law;
    without a GPU the front-end refuses to run (no CPU path)."""
    import tts_cube_b200 as cube
    from conftest import load_golden
    d = load_golden("mel_hifigan.npz")
    for tag in "ab":
        n_fft, n_mels, sr, hop, win, fmin, fmax = [int(v) for v in d[f"args_{tag}"]]
        b = cube.slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax)
        assert b.shape == (n_mels, n_fft // 2 + 1) and b.dtype == np.float32
        assert float(np.abs(b - d[f"basis_{tag}"]).max()) <= 1e-6 * float(b.max()) * 10
        m = cube.MelSpectrogram(n_fft, n_mels, sr, hop, win, fmin, fmax)
        assert m.n_frames(47 * hop) == 47 and m.n_frames(47 * hop + hop - 1) == 47 and m.n_frames((n_fft - hop) // 2) == 0
    c = cube.MelSpectrogram(flavor="cube", hop_size=256)
    assert c.n_frames(1000) == 1 + 1000 // 256 and c.cfg.layout == 1 and c.cfg.log10_out == 1
    if not torch.cuda.is_available():
        with pytest.raises(cube.CubeVocError):
            cube.mel_spectrogram(torch.zeros(1, 4096), 1024, 80, 22050, 256, 1024, 0, 8000)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_fails_loudly():
    import tts_cube_b200 as cube
    from oracle import hifigan_ref as H
    g = cube.CubeGenerator(dict(H.CONFIG_V1, upsample_initial_channel=32))
    g.load_state_dict(H.random_state_dict(dict(H.CONFIG_V1, upsample_initial_channel=32)))
    with pytest.raises(cube.CubeVocError):
        g(torch.zeros(1, 80, 4))
    with pytest.raises(cube.CubeVocError):
        cube.MULAWOutput().encode(torch.zeros(4))
    # the C ABI itself refuses to create a handle without a device
    from tts_cube_b200 import _lib
    from tts_cube_b200.generator import hifigan_config
    hp = ctypes.c_void_p()
    cfg = hifigan_config(dict(H.CONFIG_V1))
    assert _lib.lib().cube_voc_create(ctypes.byref(hp), ctypes.byref(cfg), 0) != 0
    assert b"no CUDA device" in _lib.lib().cube_voc_last_error()


def test_missing_library_is_an_error(monkeypatch):
    from tts_cube_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcube_vocoder.so")
    with pytest.raises(_lib.CubeVocError):
        _lib.lib()


def test_lpt_shard_properties():
    from tts_cube_b200 import lpt_shard
    g = torch.Generator().manual_seed(4)
    nf = torch.randint(172, 1292, (256,), generator=g).tolist()  # 2-15 s at hop 256 / 22.05 kHz
    plan = lpt_shard(nf, 8)
    flat = sorted(i for p in plan for i in p)
    assert flat == list(range(256))
    loads = [sum(nf[i] for i in p) for p in plan]
    assert max(loads) - min(loads) <= max(nf)          # LPT bound
    assert max(loads) <= 1.02 * sum(nf) / 8
    for p in plan:
        assert [nf[i] for i in p] == sorted((nf[i] for i in p), reverse=True)
    assert lpt_shard([], 4) == [[], [], [], []]
    assert lpt_shard([5], 2) == [[0], []]


def test_make_batches_and_padding():
    from tts_cube_b200.api import make_batches, pad_mels
    nf = [10, 9, 9, 4, 3, 1]
    assert make_batches(range(6), nf, 4) == [[0, 1, 2, 3], [4, 5]]
    assert make_batches(range(6), nf, 64, max_frames=25) == [[0, 1], [2, 3], [4, 5]]
    m = pad_mels([torch.ones(80, 3), torch.ones(80, 5)], pad_value=-5.0)
    assert m.shape == (2, 80, 5) and float(m[0, 0, 4]) == -5.0 and float(m[1, 0, 4]) == 1.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, {root!r})
    from tts_cube_b200.api import synthesize
    from oracle import hifigan_ref as H
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    cfg = dict(H.CONFIG_V1, upsample_initial_channel=32)
    sd = H.random_state_dict(cfg, seed=11, std=0.3, g_scale=0.42)
    # the oracle stands in for the CUDA vocoder here: this test is about sharding/scatter/gather
    vocode = lambda mel, frames: H.generator_forward_ragged(sd, cfg, mel, frames)
    out_len = lambda f: H.out_len(cfg, f)
    rank = dist.get_rank()
    g = torch.Generator().manual_seed(3)
    lens = [9, 4, 7, 2, 5]
    mels = [H.synthetic_mel(1, f, seed=50 + i)[0] for i, f in enumerate(lens)] if rank == 0 else None
    res = synthesize(None, mels, device=torch.device("cpu"), max_batch=2, vocode=vocode, out_len=out_len)
    if rank == 0:
        for i, f in enumerate(lens):
            alone = H.generator_forward(sd, cfg, mels[i][None])[0, 0]
            assert res[i].shape == alone.shape, (res[i].shape, alone.shape)
            assert float((res[i] - alone).abs().max()) <= 1e-6, i
        print("OK")
    else:
        assert res is None
    dist.barrier()
""")


def test_sharded_synthesize_world2_gloo(tmp_path):
    port = _free_port()
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0]
