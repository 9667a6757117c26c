"""bench.py's host-side logic (CPU only): the CPU legs must stay bounded whatever the host looks like - a thread per
visible core under a cgroup quota, or a sample sized from an unrepresentative probe, lost round 1 its GPU box."""
import builtins
import importlib
import io
import os
import sys

import pytest

from conftest import ROOT


@pytest.fixture()
def bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module("bench")


def test_host_cpus_respects_the_cgroup_quota(bench, monkeypatch):
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(128)), raising=False)
    real_open = builtins.open

    def fake_open(path, *a, **k):
        if path == "/sys/fs/cgroup/cpu.max":
            return io.StringIO("1600000 100000\n")          # what the GPU boxes report: 16 CPUs of 128
        if str(path).startswith("/sys/fs/cgroup/cpu/"):
            raise FileNotFoundError(path)
        return real_open(path, *a, **k)

    monkeypatch.setattr(builtins, "open", fake_open)
    assert bench.host_cpus() == 16
    monkeypatch.setattr(builtins, "open", lambda path, *a, **k: io.StringIO("max 100000\n") if path == "/sys/fs/cgroup/cpu.max"
                        else (_ for _ in ()).throw(FileNotFoundError(path)) if str(path).startswith("/sys/fs/cgroup/cpu/") else real_open(path, *a, **k))
    assert bench.host_cpus() == 128


def test_cpu_sample_only_shrinks(bench, monkeypatch):
    calls = []

    def fake_rate(arch, weights, frames, threads, batch=1):
        calls.append((batch, frames))
        n = batch * frames * 256
        dt = n / 5000.0                                      # a slow host: 5 k samples/s
        return n / dt, n, dt

    monkeypatch.setattr(bench, "cpu_oracle_rate", fake_rate)
    cb, frames = bench.cpu_fit_sample("student", None, 16, 862, n_calls=25, wall_budget_s=100.0)
    b0, f0 = bench.CPU_SAMPLE["step"]["student"]
    assert calls == [(b0, f0)]                               # one measured call, at the fixed shape
    assert cb * frames <= b0 * f0 and cb * frames * 256 / 5000.0 * 25 <= 100.0 * 1.05
    calls.clear()
    monkeypatch.setattr(bench, "cpu_oracle_rate", lambda a, w, fr, t, batch=1: (1e6, batch * fr * 256, batch * fr * 256 / 1e6))
    assert bench.cpu_fit_sample("student", None, 16, 862, n_calls=25, wall_budget_s=100.0) == (b0, f0)     # a fast host: never grown
    assert bench.cpu_fit_sample("hifigan", None, 16, 40, n_calls=25, wall_budget_s=100.0)[1] == 40          # clipped to the workload's frames


def test_input_generators_match_the_oracles(bench):
    import torch
    from oracle import clarinet_ref as C, hifigan_ref as H
    assert torch.equal(bench.synthetic_mel01(2, 30, 5), C.synthetic_mel01(2, 30, 5))
    assert torch.equal(bench.synthetic_logmel(2, 30, 5), H.synthetic_mel(2, 30, 5))
    assert bench.HIFIGAN_NEB_CONFIG == H.CONFIG_NEB
    assert set(bench.WORKLOADS) >= {"pwn", "hifigan", "ragged", "e2e", "api1", "teacher_cpu"}
