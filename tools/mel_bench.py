"""Device log-mel front-end: throughput on one B200 next to the CPU restatement (oracle/mel_ref.py, torch CPU).
Workload: B=64 x 10 s of 22 050 Hz audio (the input side of BASELINE configs[2]); CUDA events, 3 warm-ups, 10 runs;
algorithmic bytes per audio sample = 4 (read) + 80*4/256 (write) = 5.25; 8.3 KFLOP per sample (direct DFT)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tts_cube_b200 as cube  # noqa: E402
from oracle import mel_ref as M  # noqa: E402

dev = torch.device("cuda:0")
B, T = 64, 220672
args = (1024, 80, 22050, 256, 1024, 0, 8000)
y = (torch.rand(B, T, device=dev) * 1.6 - 0.8)
fe = cube.MelSpectrogram(*args)
for _ in range(3):
    m = fe(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    m = fe(y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
torch.set_num_threads(16)
yc = y[:4].cpu()
M.hifigan_mel_spectrogram(yc, *args)
t0 = time.perf_counter()
M.hifigan_mel_spectrogram(yc, *args)
cpu = 4 * T / (time.perf_counter() - t0)
flops = 2.0 * 1024 * 1026 / 256 + 2.0 * 80 * 513 / 256
print(json.dumps({"metric": "audio samples/sec (log-mel front-end)", "value": B * T / (ms / 1e3), "ms": ms, "frames": int(m.shape[2]),
                  "hbm_gbs_algorithmic": 5.25 * B * T / (ms / 1e3) / 1e9, "tflops_fp32": flops * B * T / (ms / 1e3) / 1e12,
                  "cpu_oracle_samples_per_s_16_threads": cpu}))
