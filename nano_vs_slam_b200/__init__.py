"""nano_vs_slam_b200 -- B200-native (sm_100a) implementation of the Nano-VS-SLAM perception hot path.

Public surface mirrors the reference:
    KP2DTinyV2, KP2DTinyV3, KP2DTINY_CONFIGS, KP2DTINYV3_CONFIGS, get_config, tiny_factory
    KP2DtinyFrontend (visual_odometry/frontend.py), BfFeatureMatcher (feature_matcher.py),
    IndexFlatL2 / ShardedIndexFlatL2 (faiss call site in evaluation/global_descriptor.py)
"""
from .kp2dtiny import (KP2DTINY_CONFIGS, KP2DTINYV3_CONFIGS, KP2DTinyV2, KP2DTinyV3, get_config,  # noqa: F401
                       tiny_factory)

__all__ = ["KP2DTinyV2", "KP2DTinyV3", "KP2DTINY_CONFIGS", "KP2DTINYV3_CONFIGS", "get_config", "tiny_factory"]
