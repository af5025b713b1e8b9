from .rot2RPY import rot2RPY
from .drawPoseOnImage import drawPoseOnImage
