"""Build the UNMODIFIED reference model.  TEST INFRASTRUCTURE ONLY.

In the build container the reference package is imported from where it lies (/root/reference): that is what
`oracle/make_golden.py` uses to pin `oracle/cm3p_oracle.py`.  On the GPU box, where /root/reference does not
exist, the untouched copies that `oracle/build_ref.py` placed under the git-ignored `oracle/_ref/` are imported
instead (CPU baseline of bench.py, tools/bench_reference_gpu.py).

The installed transformers (5.5.0) ModernBERT reads `layer_types`, `rope_parameters` and
`sliding_window` from its config, which the reference's configs (written for 4.55.0) do not have,
so those attributes are set on the three sub-configs before the model is constructed
(SURVEY.md §8c).  Nothing in the reference's code is edited.
"""
from __future__ import annotations

import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("CM3P_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "cm3p", "modeling_cm3p.py")):
    REFERENCE_ROOT = os.path.join(_HERE, "_ref")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "cm3p", "modeling_cm3p.py"))


def _shim_encoder_config(c):
    c.layer_types = ["sliding_attention" if i % c.global_attn_every_n_layers else "full_attention"
                     for i in range(c.num_hidden_layers)]
    c.rope_parameters = {
        "full_attention": {"rope_type": "default", "rope_theta": c.global_rope_theta},
        "sliding_attention": {"rope_type": "default", "rope_theta": c.local_rope_theta},
    }
    c.sliding_window = c.local_attention // 2
    for name in ("pad_token_id", "bos_token_id", "eos_token_id"):
        if not hasattr(c, name):
            setattr(c, name, None)


def build_reference_model(config_dict: dict, attn_implementation: str = "sdpa"):
    """-> (reference CM3PModel in eval mode, its CM3PConfig)."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from cm3p import CM3PConfig, CM3PModel  # the reference package, not cm3p_b200

    import copy
    cfg = CM3PConfig(attn_implementation=attn_implementation, **copy.deepcopy(config_dict))
    for c in (cfg.metadata_config, cfg.beatmap_config, cfg.beatmap_config.audio_config):
        _shim_encoder_config(c)
    model = CM3PModel(cfg)
    model.eval()
    return model, cfg
