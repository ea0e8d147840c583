// compat: <opencv2/opencv.hpp> -> the cv:: subset of fm3d_cv.h (see compat/README.md)
#ifndef FM3D_COMPAT_OPENCV_HPP_
#define FM3D_COMPAT_OPENCV_HPP_
#include "../../fm3d_cv.h"
#endif
