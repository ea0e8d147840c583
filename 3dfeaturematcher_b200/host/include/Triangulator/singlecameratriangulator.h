// SingleCameraTriangulator -- same public interface as the reference class
// (Triangulator/singlecameratriangulator.h:42-93); arithmetic on the GPU through libfm3d.
#ifndef FM3D_HOST_SINGLECAMERATRIANGULATOR_H_
#define FM3D_HOST_SINGLECAMERATRIANGULATOR_H_
#include <vector>
#include "../../fm3d_cv.h"
#include "../tools.h"

typedef struct {
    double x_;  // x coordinate of the pixel
    double y_;  // y coordinate of the pixel
    float i_;   // intensity of the pixel
} Pixel;

enum IMAGE_ID { image1, image2 };

struct fm3d_ctx;

class SingleCameraTriangulator {
public:
    SingleCameraTriangulator(cv::FileStorage& settings);

    void setImages(const cv::Mat& img1, const cv::Mat& img2);
    void setg12(const cv::Vec3d& T1, const cv::Vec3d& T2, const cv::Vec3d& rodrigues1, const cv::Vec3d& rodrigues2, cv::Matx44d& g12);
    void setKeypoints(const std::vector<cv::KeyPoint>& kpts1, const std::vector<cv::KeyPoint>& kpts2, const std::vector<cv::DMatch>& matches);
    void triangulate(std::vector<cv::Vec3d>& triangulatedPoints, std::vector<bool>& outliersMask);

    // Older residual formulation (singlecameratriangulator.cpp:279-339).  Point groups are 3 x N CV_64FC1
    // (N = cols, as the reference counts them); image points come back as N x 1 CV_64FC2.
    void projectPointsAndComputeResidual(const std::vector<cv::Mat>& pointsGroupVector, std::vector<cv::Mat>& imagePointsVector1,
                                         std::vector<cv::Mat>& imagePointsVector2, std::vector<std::vector<double> >& residualsVectors);
    void projectPointsAndComputeResidual(const cv::Mat& pointsGroup, cv::Mat& imagePoints1, cv::Mat& imagePoints2,
                                         std::vector<double>& residualsVector);

    // The four steps of NormalOptimizer::evaluateNormal (normaloptimizer.cpp:65-149), one by one.  They act
    // on the images given to setImages (the reference passes one pyramid level at a time) with the
    // caller's scale.  Return values as in the reference: 0, or -1 at the first point outside the
    // bounding box / pixel that is not good (the output then holds the elements before it).
    void extractPixelsContourAndGet3DPoints(const cv::Vec3d& point, const cv::Vec3d& normal, std::vector<Pixel>& pixels,
                                            std::vector<cv::Vec3d>& pointsGroup);
    void extractPixelsContour(const cv::Vec3d& point, std::vector<Pixel>& pixels);
    int get3dPointsFromImage1Pixels(const cv::Vec3d& point, const cv::Vec3d& normal, const cv::Mat& pixelMat,
                                    std::vector<cv::Vec3d>& pointsGroup);
    inline void projectPointsToImage2(const std::vector<cv::Vec3d>& pointsGroup, std::vector<Pixel>& pixels) {
        projectPointsToImage2(pointsGroup, 1.0, pixels);
    }
    int projectPointsToImage2(const std::vector<cv::Vec3d>& pointsGroup, const double scale, std::vector<Pixel>& pixels);
    int updateImage1PixelsIntensity(const double scale, std::vector<Pixel>& pixels);

    void projectPointsToImage(const IMAGE_ID id, const std::vector<std::vector<cv::Vec3d> >& pointsGroupVector,
                              std::vector<cv::Mat>& patchesVector, std::vector<cv::Mat>& imagePointsVector);
    void projectReferencePointsToImageWithFrames(const std::vector<cv::Vec3d>& referenceNeighborhooh,
                                                 const std::vector<cv::Matx44d>& featureFrames,
                                                 std::vector<cv::Mat>& patchesVector, std::vector<cv::Mat>& imagePointsVector);

    // fm3d extensions used by NormalOptimizer and by callers that do not want patch_N.pgm files
    fm3d_ctx* context() const { return ctx_; }
    int pixelsRay() const { return pixels_ray_; }
    void setWritePatchFiles(bool on) { write_patch_files_ = on; }   // reference behaviour: true (:799-802)
    void setPatchGeometry(double epsilon_m, double cm_per_pixel) { patch_eps_ = epsilon_m; patch_cmpp_ = cm_per_pixel; }

private:
    SingleCameraTriangulator();
    fm3d_ctx* ctx_;
    cv::Matx44d g_IC_, g_12_;
    cv::Vec3d rodrigues_IC_, translation_IC_;
    std::vector<float> kp1_, kp2_;
    std::vector<int> qidx_, tidx_;
    std::vector<bool> outliers_mask_;
    cv::Mat img_1_, img_2_;
    double z_threshold_min_, z_threshold_max_, patch_eps_, patch_cmpp_;
    int pixels_ray_, pyramids_;
    bool write_patch_files_;
    friend class NormalOptimizer;
};
#endif
