"""train.py entry: Hydra-style config composition (CPU) and a short end-to-end run with resume (GPU)."""
import json
import os
import textwrap

import pytest

from cm3p_b200 import hydra_lite as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(textwrap.dedent(text))


def test_compose_reference_style_tree(tmp_path):
    """Same constructs as the reference's configs/train/v7.yaml -> default.yaml -> ../model@model: v1 -> default."""
    _write(tmp_path / "train" / "default.yaml", """
        defaults:
          - /train/base@_here_
        model_cls: "CM3PModel"
        training:
          learning_rate: 1e-4
          max_steps: 30000
          bf16: true
        dataset:
          train_metadata_variations: 1
        """)
    _write(tmp_path / "train" / "v7.yaml", """
        defaults:
          - default
          - ../model@model: v1
          - _self_
        training:
          learning_rate: 4e-4
          optim: "muon"
        model:
          has_decoder_head: true
          beatmap_config:
            cls_embed: true
        """)
    _write(tmp_path / "model" / "default.yaml", """
        projection_dim: 512
        has_decoder_head: false
        beatmap_config:
          cls_embed: false
          hidden_size: 768
        """)
    _write(tmp_path / "model" / "v1.yaml", """
        defaults:
          - default
          - _self_
        """)
    cfg = H.compose(str(tmp_path / "train"), "v7", ["training.max_steps=5", "+training.global_negatives=true",
                                                     "dataset.train_metadata_variations=256"])
    assert cfg.training.learning_rate == pytest.approx(4e-4) and isinstance(cfg.training.learning_rate, float)
    assert cfg.training.optim == "muon" and cfg.training.max_steps == 5 and cfg.training.bf16 is True
    assert cfg.training.global_negatives is True
    assert cfg.model.has_decoder_head is True and cfg.model.projection_dim == 512
    assert cfg.model.beatmap_config.cls_embed is True and cfg.model.beatmap_config.hidden_size == 768
    assert cfg.dataset.train_metadata_variations == 256 and cfg.model_cls == "CM3PModel"
    with pytest.raises(KeyError):
        H.compose(str(tmp_path / "train"), "v7", ["training.not_a_key=1"])


def test_cli_parsing_and_shipped_configs():
    d, n, ov = H.parse_cli(["-cn", "synthetic_small", "training.max_steps=3", "--config-dir", "x"], "configs/train", "v1")
    assert (d, n, ov) == ("x", "synthetic_small", ["training.max_steps=3"])
    cfg = H.compose(os.path.join(ROOT, "configs", "train"), "synthetic_small", ["training.max_steps=3"])
    assert cfg.model.beatmap_config.hidden_size == 128 and cfg.dataset.seq_len == 384
    assert cfg.training.optim == "muon" and cfg.training.max_steps == 3
    base = H.compose(os.path.join(ROOT, "configs", "train"), "synthetic", [])
    assert base.model.beatmap_config.vocab_size == 3968 and base.training.per_device_train_batch_size == 8


@pytest.mark.gpu
@pytest.mark.parametrize("model_cls,extra", [
    ("CM3PModel", []),
    ("CM3PModel", ["model.has_decoder_head=true", "+model.loss_type=ForMaskedLM", "dataset.labels=masked_lm",
                   "training.optim=adamw_torch"]),
    ("CM3PForMaskedLM", ["dataset.labels=masked_lm"]),
])
def test_train_entry_runs_saves_and_resumes(tmp_path, model_cls, extra):
    import train
    out = str(tmp_path / "run")
    common = ["-cn", "synthetic_small", f"model_cls={model_cls}", f"training.output_dir={out}",
              "training.logging_steps=2", "training.save_steps=4", "training.learning_rate=2e-3",
              "dataset.fixed_batch=true"] + extra
    res = train.main(common + ["training.max_steps=8"])
    hist = res["log_history"]
    assert hist and all(h["loss"] == h["loss"] for h in hist)  # finite
    assert hist[-1]["loss"] < hist[0]["loss"], hist            # it fits the fixed synthetic batch
    assert os.path.isfile(os.path.join(out, "checkpoint-8", "model.safetensors"))
    assert os.path.isfile(os.path.join(out, "checkpoint-8", "optimizer.pt"))
    with open(os.path.join(out, "checkpoint-8", "trainer_state.json")) as f:
        assert json.load(f)["global_step"] == 8
    # auto-resume from the last checkpoint (train.py:204-223)
    res2 = train.main(common + ["training.max_steps=10"])
    assert [h["step"] for h in res2["log_history"]] == [10]
    assert os.path.isdir(os.path.join(out, "checkpoint-10"))


@pytest.mark.gpu
def test_train_entry_evaluates_with_compute_metrics(tmp_path):
    """eval_strategy=steps: eval loss + the reference's compute_metrics (zero-shot variation accuracies per class,
    masked-LM accuracy) on held-out synthetic batches with more metadata variations than in training."""
    import train
    out = str(tmp_path / "run")
    res = train.main(["-cn", "synthetic_small", f"training.output_dir={out}", "training.max_steps=4",
                      "training.logging_steps=2", "training.save_steps=0", "training.eval_strategy=steps",
                      "training.eval_steps=2", "training.per_device_eval_batch_size=6",
                      "dataset.test_metadata_variations=12", "dataset.labels=masked_lm",
                      "model.has_decoder_head=true", "+model.loss_type=ForMaskedLM", "training.optim=adamw_torch"])
    evals = [h for h in res["log_history"] if "eval_loss" in h]
    assert [h["step"] for h in evals] == [2, 4]
    for h in evals:
        assert h["eval_loss"] == h["eval_loss"] and h["eval_loss"] > 0
        for name in ("year", "status", "tags", "mapper", "masked_lm"):
            assert f"eval_accuracy_{name}" in h
            v = h[f"eval_accuracy_{name}"]
            assert v is None or 0.0 <= v <= 1.0
        assert "eval_top5_accuracy_tags" in h and "eval_top5_accuracy_masked_lm" in h
