"""Visualisation helper kept only for import-surface compatibility (ref: src/openVO/utils/drawPoseOnImage.py:5-38;
SURVEY.md §2 C6: out of the hot path)."""
import numpy as np

from .rot2RPY import rot2RPY


def drawPoseOnImage(T, img):
    import cv2
    roll, pitch, yaw = rot2RPY(T)
    i = 1 if np.linalg.norm([roll[0], pitch[0], yaw[0]]) > np.linalg.norm([roll[1], pitch[1], yaw[1]]) else 0
    h = img.shape[0]
    # aircraft-style labelling: camera z -> roll, -camera y -> pitch, camera x -> yaw
    lines = [("Roll = " + str(np.round(yaw[i], 3)), 180, 2.0), ("Pitch = " + str(np.round(-pitch[i], 3)), 120, 2.0),
             ("Yaw = " + str(np.round(roll[i], 3)), 60, 2.0),
             ("x,y,z = " + ", ".join(str(np.round(float(T[k, 3]), 1)) for k in range(3)), 10, 1.6)]
    for text, up, scale in lines:
        cv2.putText(img, text=text, org=(0, h - up), fontFace=cv2.FONT_HERSHEY_SIMPLEX, fontScale=scale, color=(0, 0, 255),
                    thickness=3)
