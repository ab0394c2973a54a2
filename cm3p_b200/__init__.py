"""cm3p_b200 — B200-native implementation of the CM3P contrastive / embedding hot path."""
from .configuration_cm3p import CM3PAudioConfig, CM3PBeatmapConfig, CM3PConfig, CM3PMetadataConfig

__all__ = ["CM3PConfig", "CM3PMetadataConfig", "CM3PAudioConfig", "CM3PBeatmapConfig"]
