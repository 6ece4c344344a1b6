"""The synthetic-stack generator lives in /synthetic.py (it makes inputs; it is not part of the checker).  This shim
keeps `from oracle import synth` working for the tests and the oracle's own modules."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from synthetic import *  # noqa: F401,F403,E402
from synthetic import (MARGIN, MOTION_AFFINE, MOTION_EUCLIDEAN, MOTION_HOMOGRAPHY, MOTION_TRANSLATION, Stack,  # noqa: F401,E402
                       config_stack, corner_displacement, make_scene, random_warp, render_frame)
