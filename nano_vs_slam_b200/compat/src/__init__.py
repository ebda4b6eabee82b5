"""Import-path shim: lets the reference scripts `from src.kp2dtiny.models.kp2dtiny import ...` resolve to nano_vs_slam_b200."""
