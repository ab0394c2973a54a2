"""CM3P configuration classes for the B200-native implementation.

Field names and defaults are the drop-in *contract* of the reference
(`/root/reference/cm3p/configuration_cm3p.py:10-343`): a YAML tree written for the reference
(`configs/model/default.yaml`) or a `config.json` saved by it must construct these classes
unchanged.  The three encoder configs share one table of ModernBERT hyper-parameters here
instead of three hand-expanded constructors.

There is exactly one attention backend in this framework (the sm_100a varlen kernel), so
`attn_implementation` is accepted (any of "flash_attention_2" / "sdpa" / "eager") and ignored.
"""
from __future__ import annotations

from transformers import AutoConfig
from transformers.configuration_utils import PretrainedConfig

# ModernBERT hyper-parameters common to the metadata / audio / beatmap encoders, with the
# per-tower defaults of the reference (configuration_cm3p.py:14-44, 93-128, 186-226).
_ENCODER_FIELDS = (
    # name,                    metadata,  audio,     beatmap
    ("hidden_size",             256,       512,       768),
    ("intermediate_size",       512,       1024,      1152),
    ("num_hidden_layers",       6,         6,         22),
    ("num_attention_heads",     4,         8,         12),
    ("hidden_activation",       "gelu",    "gelu",    "gelu"),
    ("max_position_embeddings", 128,       4096,      8192),
    ("initializer_range",       0.02,      0.02,      0.02),
    ("initializer_cutoff_factor", 2.0,     2.0,       2.0),
    ("norm_eps",                1e-5,      1e-5,      1e-5),
    ("norm_bias",               False,     False,     False),
    ("global_rope_theta",       10000.0,   160000.0,  160000.0),
    ("attention_bias",          False,     False,     False),
    ("attention_dropout",       0.0,       0.0,       0.0),
    ("global_attn_every_n_layers", 1,      3,         3),
    ("local_attention",         128,       128,       128),
    ("local_rope_theta",        10000.0,   10000.0,   10000.0),
    ("embedding_dropout",       0.0,       0.0,       0.0),
    ("mlp_bias",                False,     False,     False),
    ("mlp_dropout",             0.0,       0.0,       0.0),
    ("decoder_bias",            True,      True,      True),
    ("deterministic_flash_attn", False,    False,     False),
    ("reference_compile",       None,      None,      None),
)
_TOWER_COLUMN = {"metadata": 1, "audio": 2, "beatmap": 3}


class _EncoderConfig(PretrainedConfig):
    """Shared machinery: pops the ModernBERT fields out of **kwargs with per-tower defaults."""

    _tower = "metadata"
    _extra_defaults = {}

    def _take_fields(self, kwargs: dict) -> dict:
        col = _TOWER_COLUMN[self._tower]
        taken = {}
        for row in _ENCODER_FIELDS:
            taken[row[0]] = kwargs.pop(row[0], row[col])
        for name, default in self._extra_defaults.items():
            taken[name] = kwargs.pop(name, default)
        return taken

    def _set_fields(self, taken: dict) -> None:
        for name, value in taken.items():
            setattr(self, name, value)

    # Derived quantities used by the kernels (not serialised).
    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def layer_is_global(self, layer_idx: int) -> bool:
        """ModernBERT alternation: layer i is global iff i % global_attn_every_n_layers == 0."""
        return layer_idx % self.global_attn_every_n_layers == 0

    @property
    def window_half(self) -> int:
        """Sliding layers attend iff |i - j| <= local_attention // 2."""
        return self.local_attention // 2

    def to_dict(self):
        out = super().to_dict()
        out.pop("reference_compile", None)  # same as the reference: never serialised
        return out


class CM3PMetadataConfig(_EncoderConfig):
    model_type = "CM3PMetadata"
    base_config_key = "metadata_config"
    _tower = "metadata"
    _extra_defaults = dict(cls_embed=True, projection_dim=512, initializer_factor=1.0, vocab_size=1000)

    def __init__(self, pad_token_id=0, bos_token_id=1, eos_token_id=2, **kwargs):
        taken = self._take_fields(kwargs)
        super().__init__(pad_token_id=pad_token_id, bos_token_id=bos_token_id, eos_token_id=eos_token_id, **kwargs)
        self._set_fields(taken)


class CM3PAudioConfig(_EncoderConfig):
    model_type = "CM3PAudio"
    base_config_key = "audio_config"
    _tower = "audio"
    _extra_defaults = dict(
        projector_intermediate_size=2048,  # 4 consecutive 512-d frames -> one audio token
        projector_dim=768,
        projector_hidden_act="gelu",
        sample_rate=16000,
        n_ftt=2048,
        n_mels=80,
        hop_length=128,
        f_min=0,
        f_max=8000,
        pad_mode="constant",
    )

    def __init__(self, **kwargs):
        taken = self._take_fields(kwargs)
        kwargs.pop("vocab_size", None)
        super().__init__(**kwargs)
        self._set_fields(taken)
        self.vocab_size = 1  # the audio tower is fed embeddings, never ids (reference :129)


class CM3PBeatmapConfig(_EncoderConfig):
    model_type = "CM3PBeatmap"
    is_composition = True
    base_config_key = "beatmap_config"
    sub_configs = {"audio_config": CM3PAudioConfig}
    _tower = "beatmap"
    _extra_defaults = dict(
        audio_sos_token_id=3164,
        audio_eos_token_id=3165,
        audio_token_id=3166,
        cls_embed=True,
        projection_dim=512,
        initializer_factor=1.0,
        vocab_size=3167,
        classifier_bias=False,
        classifier_activation="gelu",
        sparse_prediction=False,
        sparse_pred_ignore_index=-100,
        repad_logits_with_grad=False,
    )

    def __init__(self, audio_config=None, pad_token_id=0, bos_token_id=1, eos_token_id=2,
                 attn_implementation=None, **kwargs):
        taken = self._take_fields(kwargs)
        super().__init__(pad_token_id=pad_token_id, bos_token_id=bos_token_id, eos_token_id=eos_token_id,
                         attn_implementation=attn_implementation, **kwargs)
        if isinstance(audio_config, CM3PAudioConfig):
            self.audio_config = audio_config
        else:
            self.audio_config = CM3PAudioConfig(attn_implementation=attn_implementation, **(audio_config or {}))
        self._set_fields(taken)


class CM3PConfig(PretrainedConfig):
    model_type = "CM3P"
    is_composition = True
    sub_configs = {"metadata_config": CM3PMetadataConfig, "beatmap_config": CM3PBeatmapConfig}

    def __init__(self, metadata_config=None, beatmap_config=None, projection_dim=512,
                 logit_scale_init_value=2.6592, initializer_factor=1.0, initializer_range=0.02,
                 loss_type=None, has_decoder_head=False, attn_implementation=None, **kwargs):
        super().__init__(attn_implementation=attn_implementation, **kwargs)
        if isinstance(metadata_config, CM3PMetadataConfig):
            self.metadata_config = metadata_config
        else:
            self.metadata_config = CM3PMetadataConfig(attn_implementation=attn_implementation,
                                                      **(metadata_config or {}))
        if isinstance(beatmap_config, CM3PBeatmapConfig):
            self.beatmap_config = beatmap_config
        else:
            self.beatmap_config = CM3PBeatmapConfig(attn_implementation=attn_implementation,
                                                    **(beatmap_config or {}))
        self.projection_dim = projection_dim
        self.logit_scale_init_value = logit_scale_init_value
        self.initializer_factor = initializer_factor
        self.initializer_range = initializer_range
        self.loss_type = loss_type
        self.has_decoder_head = has_decoder_head


def _register():
    for cls in (CM3PMetadataConfig, CM3PAudioConfig, CM3PBeatmapConfig, CM3PConfig):
        try:
            AutoConfig.register(cls.model_type, cls)
        except ValueError:
            pass  # already registered (e.g. the reference package was imported in the same process)


_register()


# --------------------------------------------------------------------------------------------
# Named configurations used by tests, the oracle goldens and bench.py.

def small_config_dict() -> dict:
    """CPU-smoke config: head_dim stays 64 and 4 layers so both attention layer types occur."""
    return dict(
        projection_dim=64,
        metadata_config=dict(cls_embed=True, hidden_size=128, intermediate_size=96, num_hidden_layers=2,
                             num_attention_heads=2, vocab_size=200, max_position_embeddings=128),
        beatmap_config=dict(
            cls_embed=True, hidden_size=128, intermediate_size=192, num_hidden_layers=4,
            num_attention_heads=2, vocab_size=500, max_position_embeddings=1024,
            audio_token_id=499, audio_sos_token_id=497, audio_eos_token_id=498,
            audio_config=dict(hidden_size=64, intermediate_size=96, num_hidden_layers=4,
                              num_attention_heads=1, projector_intermediate_size=256, projector_dim=128),
        ),
    )


def base_config_dict(has_decoder_head: bool = False) -> dict:
    """`configs/model/default.yaml` + the v7 overrides (`configs/train/v7.yaml:28-36`), v7 vocabulary."""
    return dict(
        projection_dim=512,
        has_decoder_head=has_decoder_head,
        loss_type="ForMaskedLM" if has_decoder_head else None,
        metadata_config=dict(cls_embed=True, vocab_size=1000),
        beatmap_config=dict(cls_embed=True, vocab_size=3968, audio_token_id=3967, audio_sos_token_id=3965,
                            audio_eos_token_id=3966, pad_token_id=3962, bos_token_id=3958, eos_token_id=3959),
    )


__all__ = ["CM3PConfig", "CM3PMetadataConfig", "CM3PAudioConfig", "CM3PBeatmapConfig",
           "small_config_dict", "base_config_dict"]
