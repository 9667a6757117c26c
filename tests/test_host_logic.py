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


def _header_struct_words(name):
    """count the 4-byte fields of `typedef struct <name> { ... }` in include/cube_vocoder.h (arrays by their extents)"""
    hdr = open(os.path.join(ROOT, "include", "cube_vocoder.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    consts = {k: int(v) for k, v in re.findall(r"#define\s+(CUBE_MAX_\w+)\s+(\d+)", hdr)}
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, flags=re.S).group(1)
    n = 0
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(uint32_t|int32_t|float)\s+(.*)", decl, flags=re.S)
        assert m, decl
        for var in m.group(2).split(","):
            k = 1
            for ext in re.findall(r"\[(\w+)\]", var):
                k *= consts.get(ext, None) or int(ext)
            n += k
    return n


def test_config_struct_matches_header_size():
    from tts_cube_b200 import _lib
    assert ctypes.sizeof(_lib.VocConfig) == 4 * _header_struct_words("cube_voc_config")
    assert ctypes.sizeof(_lib.MelConfig) == 4 * _header_struct_words("cube_mel_config")
    assert _lib.VocConfig().struct_size == ctypes.sizeof(_lib.VocConfig)
    assert _lib.MelConfig(n_fft=1024).struct_size == ctypes.sizeof(_lib.MelConfig)


def integration_stub_namespace():
    """exec the raw ctypes stub of INTEGRATION.md verbatim (only the library path is made absolute)"""
    from tts_cube_b200 import _lib
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```python\n(# raw-ctypes-stub:.*?)```", md, flags=re.S)
    assert m, "INTEGRATION.md lost its raw ctypes stub"
    code = m.group(1)
    assert 'C.CDLL("libcube_vocoder.so")' in code
    ns = {}
    exec(compile(code.replace('"libcube_vocoder.so"', repr(_lib.LIB_PATH)), "INTEGRATION.md:stub", "exec"), ns)
    return ns


def test_integration_stub_matches_the_abi():
    """the documented stub's struct is the library's struct, and a stale struct (the round-1 stub without the
    wrnn_* fields) is rejected by struct_size instead of being read out of bounds"""
    from tts_cube_b200 import _lib
    ns = integration_stub_namespace()
    Stub = ns["VocConfig"]
    assert ctypes.sizeof(Stub) == ctypes.sizeof(_lib.VocConfig)
    assert [f[0] for f in Stub._fields_] == [f[0] for f in _lib.VocConfig._fields_]
    lib = ns["lib"]
    hp = ctypes.c_void_p()

    class Stale(ctypes.Structure):
        _fields_ = Stub._fields_[:-6]
    st = Stale(struct_size=ctypes.sizeof(Stale))
    assert lib.cube_voc_create(ctypes.byref(hp), ctypes.byref(st), 0) != 0
    assert "struct_size" in ns["last_error"]()
    zero = Stub()          # struct_size left 0
    assert lib.cube_voc_create(ctypes.byref(hp), ctypes.byref(zero), 0) != 0
    assert "struct_size" in ns["last_error"]()
    if not torch.cuda.is_available():
        ok = Stub(struct_size=ctypes.sizeof(Stub), n_ups=1, n_resblock_kernels=1, resblock_type=1)
        assert lib.cube_voc_create(ctypes.byref(hp), ctypes.byref(ok), 0) != 0
        assert "no CUDA device" in ns["last_error"]()


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
    assert make_batches(range(6), nf, 4, batch_overhead_frames=0) == [[0], [1, 2], [3], [4], [5]]    # padding only: equal lengths pair up
    one = make_batches(range(6), nf, 64, batch_overhead_frames=10 ** 6)
    assert one == [[0, 1, 2, 3, 4, 5]]                                          # launches dominate: one batch
    capped = make_batches(range(6), nf, 64, max_frames=25, batch_overhead_frames=10 ** 6)
    assert all(len(b) * nf[b[0]] <= 25 for b in capped) and sorted(i for b in capped for i in b) == list(range(6))
    assert make_batches([], nf, 4) == []
    # padding-aware: an LPT shard mixes long and short utterances; a batch must not be mostly padding
    g = torch.Generator().manual_seed(9)
    nf2 = sorted(torch.randint(188, 1407, (32,), generator=g).tolist(), reverse=True)
    assert 32 * nf2[0] > 1.4 * sum(nf2)                                          # one padded batch would be > 40 % waste
    bs = make_batches(range(32), nf2, 64)
    assert sorted(i for b in bs for i in b) == list(range(32)) and 2 <= len(bs) <= 6
    assert all(b == list(range(b[0], b[-1] + 1)) for b in bs)                    # consecutive runs of the sorted list
    padded = sum(len(b) * nf2[b[0]] for b in bs)
    assert padded <= 1.25 * sum(nf2)
    # optimal for its cost model: no single cut or merge improves it
    cost = lambda bb: sum(len(b) * nf2[b[0]] + 1600 for b in bb)
    for k in range(len(bs) - 1):
        assert cost(bs) <= cost(bs[:k] + [bs[k] + bs[k + 1]] + bs[k + 2:])
    m = pad_mels([torch.ones(80, 3), torch.ones(80, 5)], pad_value=-5.0)
    assert m.shape == (2, 80, 5) and float(m[0, 0, 4]) == -5.0 and float(m[1, 0, 4]) == 1.0


def test_synthesize_draws_noise_for_the_student_when_zs_is_none():
    """a vocoder that needs z (the IAF student) used to get a 2-argument call when zs=None."""
    from tts_cube_b200.api import synthesize

    class ParallelWaveNetVocoder:                       # stands in for the CUDA class: same name, same call contract
        device = torch.device("cpu")
        calls = []

        def out_len(self, f):
            return 4 * f

        def __call__(self, mel, z, frames):
            assert z.shape == (mel.shape[0], 1, 4 * mel.shape[2]) and len(frames) == mel.shape[0]
            self.calls.append((tuple(mel.shape), float(z.abs().sum())))
            return mel.mean(1, keepdim=True).repeat_interleave(4, dim=2) + 0 * z

    v = ParallelWaveNetVocoder()
    mels = [torch.full((80, f), float(f)) for f in (5, 3, 7)]
    out = synthesize(v, mels, max_batch=2)
    assert [o.shape[0] for o in out] == [20, 12, 28]
    assert all(float(o[0]) == f for o, f in zip(out, (5, 3, 7)))
    assert all(zsum > 0 for _, zsum in v.calls)          # noise was drawn, not zeros
    zs = [torch.zeros(4 * f) for f in (5, 3, 7)]
    v.calls.clear()
    synthesize(v, mels, zs=zs, max_batch=2)
    assert all(zsum == 0 for _, zsum in v.calls)         # injected z is what reaches the vocoder


def generator_shell(sd, h):
    """A parameter tree with the reference Generator's state_dict keys and `.h` - what install_into_cubegan reads of
    `model._generator` (cube/networks/cubegan.py:41-43) - without importing the reference."""
    root = torch.nn.Module()
    root.h = h
    for key, val in sd.items():
        mod = root
        parts = key.split(".")
        for part in parts[:-1]:
            if not hasattr(mod, part):
                mod.add_module(part, torch.nn.Module())
            mod = getattr(mod, part)
        mod.register_parameter(parts[-1], torch.nn.Parameter(val.clone(), requires_grad=False))
    return root


class CubeganStandIn(torch.nn.Module):
    """The attributes and the 8 lines of `Cubegan.inference` (cube/networks/cubegan.py:74-83) that the drop-in touches;
    the frontend returns a stored conditioning (made by the real Languasito2 in oracle/make_cubegan_golden.py)."""

    class _Frontend(torch.nn.Module):
        def __init__(self, conds):
            super().__init__()
            self.conds = conds

        def inference(self, X, hf_cond=None):
            return self.conds[int(X["utt"])]

    def __init__(self, generator, conds):
        super().__init__()
        self._generator = generator
        self._languasito = CubeganStandIn._Frontend(conds)
        self._hf = None

    def get_device(self):
        return next(self._generator.parameters()).device if list(self._generator.parameters()) else self._generator.device

    def inference(self, X):
        with torch.no_grad():
            hf_cond = None
            conditioning = self._languasito.inference(X, hf_cond=hf_cond)
            if conditioning.shape[1] == 0:
                conditioning = torch.zeros((conditioning.shape[0], 1, conditioning.shape[2]), device=self.get_device())
            return self._generator(conditioning.permute(0, 2, 1))


def test_install_into_cubegan_swaps_the_generator():
    import types
    import tts_cube_b200 as cube
    from oracle import hifigan_ref as H
    cfg = dict(H.CONFIG_V1, upsample_initial_channel=32)
    sd = H.random_state_dict(cfg, seed=11)
    model = CubeganStandIn(generator_shell(sd, types.SimpleNamespace(**cfg)), [])
    cube.install_into_cubegan(model)
    g = model._generator
    assert isinstance(g, cube.CubeGenerator) and g.device.type == "cpu"
    assert set(g.state_dict()) == set(sd) and all(torch.equal(g.state_dict()[k], sd[k]) for k in sd)
    assert g._cfg.upsample_initial_channel == 32 and g._cfg.math == 0      # auto: stage widths 16..2 -> the fp32 kernels
    from tts_cube_b200.generator import hifigan_config
    assert hifigan_config(H.CONFIG_V1).math == 1                           # config_v1 (what Cubegan hard-codes): tensor cores
    if not torch.cuda.is_available():
        with pytest.raises(cube.CubeVocError):                             # still no CPU path behind the reference call
            g(torch.zeros(1, 80, 4))


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
    else:
        assert res is None
    # the noise path (IAF student): z blocks travel with the mel blocks, utterance by utterance
    hop = 4
    voc3 = lambda mel, z, frames: mel.mean(1, keepdim=True).repeat_interleave(hop, dim=2) + z
    lens3 = [6, 11, 3, 8, 8, 2, 9]
    mels3 = [torch.full((80, f), float(i + 1)) for i, f in enumerate(lens3)] if rank == 0 else None
    zs3 = [torch.arange(hop * f, dtype=torch.float32) * 1e-3 + i for i, f in enumerate(lens3)] if rank == 0 else None
    stats = {{}}
    res3 = synthesize(None, mels3, zs=zs3, device=torch.device("cpu"), max_batch=3, vocode=voc3, out_len=lambda f: hop * f, stats=stats)
    if rank == 0:
        for i, f in enumerate(lens3):
            want = float(i + 1) + zs3[i]
            assert res3[i].shape == want.shape and float((res3[i] - want).abs().max()) <= 1e-6, i
        assert stats["scatter_bytes"] > 0 and stats["gather_bytes"] > 0 and len(stats["batches_per_rank"]) == 2
        print("OK")
    else:
        assert res3 is None
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


def test_every_runtime_switch_is_documented():
    """The engine reads its kernel-selection switches from the environment; each one must be listed in README.md ("Runtime
    switches") and be exercised by a variant run of the GPU suite or named there as a measurement aid."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for f in glob.glob(os.path.join(root, "tts_cube_b200", "csrc", "*.cu*")):
        names |= set(re.findall(r'getenv\("(CUBE_[A-Z0-9_]+)"\)', open(f).read()))
    assert len(names) >= 10
    readme = open(os.path.join(root, "README.md")).read()
    missing = sorted(n for n in names if n not in readme)
    assert not missing, f"switches read by the engine but absent from README.md: {missing}"
    gpu_tests = open(os.path.join(root, "tests", "test_gpu_parity.py")).read()
    aids = {"CUBE_BLOCK_STATS", "CUBE_GRAPH", "CUBE_TC_PREFETCH"}       # instrumentation / covered by their own tests or opt-in only
    untested = sorted(n for n in names - aids if f'"{n}"' not in gpu_tests)
    assert not untested, f"switches without a variant run in tests/test_gpu_parity.py: {untested}"
