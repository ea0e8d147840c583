// NormalOptimizer -- same public interface as the reference class
// (Triangulator/normaloptimizer.h:43-59).  The PCL visualiser thread of the reference
// (normaloptimizer.cpp:121-123,304-318) is not reproduced: start/stopVisualizerThread are no-ops.
#ifndef FM3D_HOST_NORMALOPTIMIZER_H_
#define FM3D_HOST_NORMALOPTIMIZER_H_
#include <vector>
#include "../../fm3d_cv.h"
#include "singlecameratriangulator.h"

class NormalOptimizer {
public:
    NormalOptimizer(const cv::FileStorage settings, SingleCameraTriangulator* sct);

    void setImages(const cv::Mat& img1, const cv::Mat& img2);
    void computeOptimizedNormals(std::vector<cv::Vec3d>& points3D, std::vector<cv::Vec3d>& normalsVector);
    void computeOptimizedNormals(std::vector<cv::Vec3d>& points3D, std::vector<cv::Vec3d>& normalsVector, std::vector<cv::Scalar>& colors);
    void computeFeaturesFrames(std::vector<cv::Vec3d>& points3D, std::vector<cv::Vec3d>& normalsVector, std::vector<cv::Matx44d>& featuresFrames);
    void startVisualizerThread() {}
    void stopVisualizerThread() {}
    cv::Vec3d getGravity();

    // fm3d extensions: abs() semantics of the penalty wall (0 fabs = as compiled today, 1 int abs, 2 off)
    void setPenaltyMode(int mode) { penalty_mode_ = mode; }
    // what is minimised: 0 the reference's SSD (default), 1 the zero-mean normalised cost (fm3d_cost_mode)
    void setCostMode(int mode) { cost_mode_ = mode; }
    // per-feature outcome of the last computeOptimizedNormals, indexed like the INPUT points3D
    const std::vector<int>& lastStatus() const { return status_; }
    const std::vector<int>& lastEvaluations() const { return nfev_; }

private:
    NormalOptimizer();
    SingleCameraTriangulator* sct_;
    int pyr_levels_, penalty_mode_, cost_mode_;
    double epsilon_lmmin_;
    cv::Vec3d gravity_;
    std::vector<int> status_, nfev_;
};
#endif
