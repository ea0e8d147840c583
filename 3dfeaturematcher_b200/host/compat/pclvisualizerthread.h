// compat: the interface of the reference's pclVisualizerThread (pclvisualizerthread.h:38-73) with methods that do
// nothing.  The reference's version rebuilds a PCL cloud under a mutex inside EVERY Levenberg-Marquardt evaluation
// (normaloptimizer.cpp:121-123) and needs a display; SURVEY D12 excludes it from product and baseline alike.
#ifndef FM3D_COMPAT_PCLVISUALIZERTHREAD_H_
#define FM3D_COMPAT_PCLVISUALIZERTHREAD_H_
#include <vector>
#include <opencv2/opencv.hpp>
class pclVisualizerThread {
public:
    pclVisualizerThread() {}
    void operator()() {}
    void updateClouds(const std::vector<cv::Vec3d>&, const cv::Vec3d&, const cv::Scalar&) {}
    void keepLastCloud() {}
};
#endif
