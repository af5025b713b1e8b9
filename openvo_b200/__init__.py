"""openvo_b200 — B200-native drop-in for the per-frame stereo-VO hot path of KevinSpevak/openVO.

Import surface mirrors ref: src/openVO/__init__.py:2-5.
"""
from .stereo_camera import StereoCamera
from .stereo_odometer import StereoOdometer
from .utils.rot2RPY import rot2RPY
from .utils.drawPoseOnImage import drawPoseOnImage

__all__ = ["StereoCamera", "StereoOdometer", "rot2RPY", "drawPoseOnImage"]
