"""CPU restatement of the reference's two log-mel front-ends.  TEST INFRASTRUCTURE ONLY - the product
(`tts_cube_b200`) never imports this file.

Pinning: `hifigan_mel_spectrogram` is checked against the UNMODIFIED reference function
(hifigan/meldataset.py:50-74) executed in the build container (tests/golden/mel_hifigan.npz, made by
oracle/make_goldens.py --mel; the harness passes `return_complex=False` to torch.stft, which torch >= 2.0
requires and the torch==1.4 reference omits).  The mel filter bank is third-party arithmetic
(`librosa.filters.mel`, hifigan/requirements.txt:3 pins librosa==0.7.2, not installed here and not vendored):
`slaney_mel_basis` restates its published algorithm (Slaney auditory-toolbox mel scale, htk=False, area
normalisation) and is checked against torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney'),
an independent implementation of the same definition.  `cube_melspectrogram` (cube/io_utils/vocoder.py:54-98)
calls librosa.stft, which cannot run here: PARITY UNPINNED for that flavour beyond the STFT identity
librosa.stft(center=True, pad_mode='reflect', window='hann') == torch.stft(center=True, pad_mode='reflect',
periodic hann), which is how it is restated.
"""
import numpy as np
import torch


def _hz_to_mel(f):
    """librosa.core.hz_to_mel(htk=False): linear below 1 kHz (200/3 Hz per mel), log above (step ln(6.4)/27)."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def slaney_mel_basis(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, Slaney norm), float32
    [n_mels, 1 + n_fft//2]: triangles between neighbouring mel centres, each scaled by 2 / (its band width in Hz).
    Plain loops on purpose (the package has its own vectorised copy; the two are compared in the tests)."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = 1 + n_fft // 2
    fft_f = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = np.linspace(float(_hz_to_mel(fmin)), float(_hz_to_mel(fmax)), n_mels + 2)
    hz_pts = _mel_to_hz(mel_pts)
    w = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lo, ce, hi = hz_pts[i], hz_pts[i + 1], hz_pts[i + 2]
        for k in range(n_bins):
            up = (fft_f[k] - lo) / (ce - lo)
            dn = (hi - fft_f[k]) / (hi - ce)
            w[i, k] = max(0.0, min(up, dn)) * 2.0 / (hi - lo)
    return w.astype(np.float32)


def hifigan_mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, basis=None):
    """hifigan/meldataset.py:50-74 for y [B, T] float32 -> [B, num_mels, F] (natural log)."""
    y = torch.as_tensor(y, dtype=torch.float32)
    if basis is None:
        basis = slaney_mel_basis(sampling_rate, n_fft, num_mels, fmin, fmax)
    basis = torch.as_tensor(basis, dtype=torch.float32)
    pad = int((n_fft - hop_size) / 2)                                                   # meldataset.py:62
    yp = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(yp, n_fft, hop_length=hop_size, win_length=win_size, window=torch.hann_window(win_size),
                      center=False, pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + 1e-9)                        # meldataset.py:70
    return torch.log(torch.clamp(torch.matmul(basis, mag), min=1e-5))                   # :72-73, :27-28


def cube_melspectrogram(y, sample_rate, num_mels, hop_size, use_preemphasis=False, basis=None, pad_mode="reflect"):
    """cube/io_utils/vocoder.py:54-62 for ONE utterance y [T] -> [F, num_mels] (log10, time-major):
    optional lfilter([1, -0.97]) (:64-65), librosa.stft(n_fft=1024, hop, win 1024, 'hann') = centred frames of the
    reflect-padded signal (:71-73), |.|, librosa mel basis with fmin=0, fmax=sr/2 (:80-82), log10(max(1e-5, .)) (:96-98)."""
    n_fft = 1024
    y = torch.as_tensor(y, dtype=torch.float32)
    if use_preemphasis:
        y = torch.cat([y[:1], y[1:] - 0.97 * y[:-1]])
    if basis is None:
        basis = slaney_mel_basis(sample_rate, n_fft, num_mels)
    spec = torch.stft(y[None], n_fft, hop_length=hop_size, win_length=n_fft, window=torch.hann_window(n_fft),
                      center=True, pad_mode=pad_mode, return_complex=True)[0]   # librosa < 0.10: reflect; >= 0.10: constant
    mel = torch.matmul(torch.as_tensor(basis), spec.abs())
    return torch.log10(torch.clamp(mel, min=1e-5)).T.contiguous()


def test_signal(B, T, seed=0, sr=22050):
    """speech-like synthetic audio: a few harmonics with slow amplitude/pitch drift + coloured noise, peak ~0.8"""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(T, dtype=torch.float64) / sr
    out = []
    for _ in range(B):
        f0 = 90.0 + 160.0 * float(torch.rand(1, generator=g))
        ph = 2 * np.pi * torch.cumsum(f0 * (1 + 0.1 * torch.sin(2 * np.pi * 3.1 * t)) / sr * torch.ones(T, dtype=torch.float64), 0)
        s = sum((0.6 ** h) * torch.sin((h + 1) * ph + float(torch.rand(1, generator=g)) * 6.28) for h in range(8))
        env = 0.55 + 0.45 * torch.sin(2 * np.pi * 1.7 * t + float(torch.rand(1, generator=g)) * 6.28)
        n = torch.randn(T, generator=g, dtype=torch.float64)
        n = torch.nn.functional.avg_pool1d(n[None, None], 5, 1, 2)[0, 0] * 0.05
        x = s * env + n
        out.append(0.8 * x / x.abs().max())
    return torch.stack(out).float()
