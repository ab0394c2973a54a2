"""GPU log-mel front-end vs the CPU `WhisperFeatureExtractor` the reference's processor uses
(cm3p/processing_cm3p.py:284-304; configs/train/default.yaml processor.audio_feature_extractor)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _waves(batch, samples, seed=0):
    rs = np.random.RandomState(seed)
    t = np.arange(samples) / 16000.0
    out = []
    for b in range(batch):
        w = sum(a * np.sin(2 * np.pi * f * t + p) for a, f, p in
                zip(rs.uniform(0.02, 0.4, 6), rs.uniform(60, 7000, 6), rs.uniform(0, 6.28, 6)))
        w = w * (0.3 + 0.7 * (np.sin(2 * np.pi * 0.7 * t + b) > 0)) + 0.01 * rs.standard_normal(samples)
        out.append(w.astype(np.float32))
    return np.stack(out)


@pytest.mark.parametrize("samples", [256000, 48000])
def test_logmel_matches_whisper_feature_extractor(samples):
    from transformers import WhisperFeatureExtractor
    from cm3p_b200.audio_features import LogMelSpectrogram
    fe = WhisperFeatureExtractor(feature_size=80, sampling_rate=16000, hop_length=160, chunk_length=30, n_fft=400,
                                 padding_value=0, dither=0, return_attention_mask=False)
    waves = _waves(3, samples)
    want = np.stack([fe._np_extract_fbank_features(w[None], "cpu")[0] for w in waves])  # (B, 80, samples/160)
    got = LogMelSpectrogram("cuda")(torch.from_numpy(waves).cuda())
    torch.cuda.synchronize()
    assert tuple(got.shape) == want.shape == (3, 80, samples // 160)
    err = np.abs(got.cpu().numpy() - want)
    assert float(err.max()) <= 5e-3, float(err.max())  # isolated near-floor bins; bf16 consumers resolve 4e-3
    assert float(err.mean()) <= 1e-4, float(err.mean())


def test_logmel_feeds_the_model():
    """(B, 80, 1600) features from 16 s windows go straight into the audio encoder."""
    import copy
    from cm3p_b200.audio_features import LogMelSpectrogram
    from cm3p_b200.configuration_cm3p import CM3PConfig, small_config_dict
    from cm3p_b200.modeling_cm3p import CM3PModel
    from cm3p_b200.synthetic import synthetic_batch, synthetic_state_dict
    cfg = CM3PConfig(**copy.deepcopy(small_config_dict()))
    model = CM3PModel(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=0), strict=True)
    model = model.cuda().eval()
    batch = synthetic_batch(cfg, batch=2, seq_len=320, seed=3)
    feats = LogMelSpectrogram("cuda")(torch.from_numpy(_waves(2, 256000)).cuda())
    assert feats.shape == (2, 80, 1600)
    with torch.no_grad():
        out = model(input_ids=batch["input_ids"].cuda(), attention_mask=batch["attention_mask"].cuda(),
                    input_features=feats, return_loss=False)
    assert torch.isfinite(out.beatmap_embeds).all()
