"""Drop-in for src/kp2dtiny/models/kp2dtiny.py: re-exports the B200 implementation under the reference's
import path (put nano_vs_slam_b200/compat first on PYTHONPATH)."""
from nano_vs_slam_b200.kp2dtiny import (KP2DTINY_CONFIGS, KP2DTINYV3_CONFIGS, KP2DTinyV2, KP2DTinyV3,  # noqa: F401
                                        get_config, tiny_factory)
