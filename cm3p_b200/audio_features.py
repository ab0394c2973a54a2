"""Log-mel spectrogram on the GPU (the "f3" row of SURVEY.md §8f).

The reference computes its `input_features` with `transformers.WhisperFeatureExtractor` on the CPU
(cm3p/processing_cm3p.py:284-304; numpy STFT), which becomes the bottleneck once the model embeds
thousands of windows per second.  `LogMelSpectrogram` produces the same features from waveforms that are
already on the GPU: framing + Hann window, the DFT as three tcgen05 GEMMs on hi/lo-split bf16 operands
(fp32 accumulation, ~1e-5 relative accuracy), |.|^2 -> slaney mel filter bank -> log10 -> dynamic-range
clamp -> (x+4)/4.  Window, DFT basis and mel filters are built once on the host with numpy in float64.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib, ops


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mel = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) * logstep, mel)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), f)


def slaney_mel_filters(n_bins: int, n_mels: int, sampling_rate: int, fmin: float = 0.0, fmax: float | None = None):
    """[n_bins, n_mels] triangular filters, slaney scale + slaney area normalisation (the Whisper filter bank)."""
    fmax = sampling_rate / 2.0 if fmax is None else fmax
    fft_freqs = np.linspace(0, sampling_rate // 2, n_bins)
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    diff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    filt = np.maximum(0.0, np.minimum(down, up))
    filt *= (2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels]))[None, :]
    return filt


def _split(x: torch.Tensor):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


class LogMelSpectrogram:
    def __init__(self, device, feature_size: int = 80, sampling_rate: int = 16000, hop_length: int = 160,
                 n_fft: int = 400):
        self.device = torch.device(device)
        self.n_mels, self.hop, self.n_fft = feature_size, hop_length, n_fft
        self.bins = n_fft // 2 + 1
        n = np.arange(n_fft, dtype=np.float64)
        window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)  # periodic Hann
        k = np.arange(self.bins, dtype=np.float64)[:, None]
        ang = 2.0 * np.pi * k * n[None, :] / n_fft
        basis = np.concatenate([np.cos(ang), -np.sin(ang)], axis=0)  # [2*bins, n_fft]
        self.ld = (n_fft + 7) // 8 * 8
        self.ld_spec = (2 * self.bins + 7) // 8 * 8
        basis_t = torch.zeros((2 * self.bins, self.ld), dtype=torch.float32)
        basis_t[:, :n_fft] = torch.from_numpy(basis).float()
        self.w_hi, self.w_lo = (t.to(self.device) for t in _split(basis_t))
        self.window = torch.from_numpy(window).float().to(self.device)
        self.filters = torch.from_numpy(slaney_mel_filters(self.bins, feature_size, sampling_rate)).float().contiguous().to(
            self.device)

    @torch.no_grad()
    def __call__(self, waveforms: torch.Tensor) -> torch.Tensor:
        """waveforms [B, samples] fp32 on the GPU -> [B, n_mels, samples // hop] fp32 (Whisper log-mel)."""
        if not waveforms.is_cuda:
            raise RuntimeError("LogMelSpectrogram: waveforms must be on the GPU (there is no CPU fallback)")
        wave = waveforms.float().contiguous()
        B, N = wave.shape
        frames = N // self.hop  # the extractor computes N/hop + 1 frames and drops the last one
        lib, st = _lib.load(), torch.cuda.current_stream().cuda_stream
        rows = B * frames
        f_hi = torch.empty((rows, self.ld), device=self.device, dtype=torch.bfloat16)
        f_lo = torch.empty_like(f_hi)
        _lib.check(lib.cm3p_logmel_frames(wave.data_ptr(), self.window.data_ptr(), f_hi.data_ptr(), f_lo.data_ptr(), B, N,
                                          frames, self.n_fft, self.hop, self.ld, st), "cm3p_logmel_frames")
        buf = torch.zeros((rows, self.ld_spec), device=self.device, dtype=torch.float32)
        spec = buf[:, :2 * self.bins]
        for a, w in ((f_hi, self.w_hi), (f_hi, self.w_lo), (f_lo, self.w_hi)):
            ops.gemm(a, w, epilogue=ops.EPI_SCALE_F32, accumulate=True, out=spec)
        out = torch.empty((B, self.n_mels, frames), device=self.device, dtype=torch.float32)
        clip_max = torch.full((B,), -math.inf, device=self.device, dtype=torch.float32)
        _lib.check(lib.cm3p_logmel_power_mel(buf.data_ptr(), self.ld_spec, self.filters.data_ptr(), out.data_ptr(),
                                             clip_max.data_ptr(), B, frames, self.bins, self.n_mels, st),
                   "cm3p_logmel_power_mel")
        _lib.check(lib.cm3p_logmel_finalize(out.data_ptr(), clip_max.data_ptr(), B, self.n_mels * frames, st),
                   "cm3p_logmel_finalize")
        return out
