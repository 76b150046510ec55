"""mclstexp_b200: B200-native contrastive-alignment + retrieval hot path of mclSTExp."""
__version__ = "0.1.0"
