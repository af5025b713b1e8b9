"""Recover OpenCV's rBRIEF sampling pattern (bit_pattern_31_) from the installed cv2 binary.

The reference (openVO) delegates ORB to cv2.ORB_create (ref: src/openVO/stereo_odometer.py:22); the
pattern is data inside that un-vendored dependency.  SURVEY.md A.2.2 gives the signature + sha1.
Writes openvo_b200/csrc/orb_pattern.inc (shared by the CUDA kernels and the oracle).
"""
import hashlib, os, sys
import numpy as np
import cv2

def main():
    so = os.path.join(os.path.dirname(cv2.__file__), "cv2.abi3.so")
    data = open(so, "rb").read()
    sig = np.array([8, -3, 9, 5, 4, 2, 7, -12, -11, 9, -8, 2, 7, -12, 12, -13], dtype="<i4").tobytes()
    off = data.find(sig)
    assert off >= 0 and data.find(sig, off + 1) < 0, "pattern signature not unique"
    blob = data[off:off + 4096]
    assert hashlib.sha1(blob).hexdigest() == "c9ecd83d8d9c918348516d67381c3c8fe65fc014"
    pat = np.frombuffer(blob, dtype="<i4").reshape(256, 4)
    out = os.path.join(os.path.dirname(__file__), "..", "openvo_b200", "csrc", "orb_pattern.inc")
    with open(out, "w") as f:
        f.write("// rBRIEF learned sampling pattern (256 pairs x0,y0,x1,y1), OpenCV's bit_pattern_31_.\n")
        f.write("// Recovered from the installed cv2 4.13.0 binary by tools/extract_orb_pattern.py (sha1 of the 4096\n")
        f.write("// little-endian bytes = c9ecd83d8d9c918348516d67381c3c8fe65fc014, SURVEY.md A.2.2). Data only.\n")
        for r in pat:
            f.write("%d,%d,%d,%d,\n" % tuple(r))
    print("wrote", out)

if __name__ == "__main__":
    sys.exit(main())
