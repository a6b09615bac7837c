/* TEST INFRASTRUCTURE ONLY: RGBDOdometry.h includes PlaneExtraction.h; the functions that use it (correspondPlaneSearch*, never called:
 * SURVEY section 2 row 14) are not part of the trimmed translation unit oracle/build_ref_odometry.py generates. */
#pragma once
class PlaneExtraction {};
