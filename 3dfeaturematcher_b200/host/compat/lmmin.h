// compat: <lmmin.h> of lmfit (main.cpp:8, normaloptimizer.h:37 in the reference).  The Levenberg-Marquardt solver runs
// inside the GPU kernel (csrc/fm3d_lm2.h); clients include this header but call nothing from it.
#ifndef FM3D_COMPAT_LMMIN_H_
#define FM3D_COMPAT_LMMIN_H_
#endif
