// compat: <pcl/visualization/pcl_visualizer.h>: the viewers of the reference (tools.cpp:243-765) are not reproduced
#ifndef FM3D_COMPAT_PCL_VISUALIZER_H_
#define FM3D_COMPAT_PCL_VISUALIZER_H_
#include "../common/common_headers.h"
#endif
