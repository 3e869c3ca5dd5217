"""Constants of the delay path (reference: tools/RAiDER/constants.py:11-23)."""
import numpy as np

_ZMIN = np.float64(-100)  # minimum required height
_ZREF = np.float64(26000)  # maximum integration height when not specified by user
_STEP = np.float64(15.0)  # integration step size in meters

R_EARTH_MAX_WGS84 = 6378137
R_EARTH_MIN_WGS84 = 6356752

_CUBE_SPACING_IN_M = float(2000)  # Horizontal spacing of cube
