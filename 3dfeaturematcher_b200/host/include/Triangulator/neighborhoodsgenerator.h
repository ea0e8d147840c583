// NeighborhoodsGenerator -- same public interface as the reference class
// (Triangulator/neighborhoodsgenerator.h:81-97): `square` (the method main.cpp / mosaic.cpp use) and
// `circular` (neighborhoodsgenerator.cpp:160-277).  An unknown method throws (the reference exit(-10)s).
#ifndef FM3D_HOST_NEIGHBORHOODSGENERATOR_H_
#define FM3D_HOST_NEIGHBORHOODSGENERATOR_H_
#include <vector>
#include "../../fm3d_cv.h"

class NeighborhoodsGenerator {
public:
    NeighborhoodsGenerator(cv::FileStorage settings);
    // points / normals: 3 x N CV_64FC1 (one column per feature); an empty normals Mat is filled with the
    // initial guess P/|P|; every neighbourhood is a 1 x (thetas*rays) CV_64FC3 Mat
    void computeCircularNeighborhoodsByNormals(const cv::Mat& points, cv::Mat& normals, std::vector<cv::Mat>& neighborhoodsVector);
    void computeCircularNeighborhoodByNormal(const cv::Vec3d& point, cv::Vec3d& normal, cv::Mat& neighborhood);
    void computeSquareNeighborhoodsByNormals(const std::vector<cv::Matx44d>& featuresFrames,
                                             std::vector<std::vector<cv::Vec3d> >& neighborhoodsVector);
    void computeSquareNeighborhoodByNormal(const cv::Matx44d& featureFrame, std::vector<cv::Vec3d>& neighborhood);
    void getReferenceSquaredNeighborhood(std::vector<cv::Vec3d>& neighborhood);
    double epsilon() const { return epsilon_; }
    double cmPerPixel() const { return cm_per_pixel_; }

private:
    NeighborhoodsGenerator();
    double epsilon_, cm_per_pixel_;
    int number_of_angles_, number_of_rays_;
    bool circular_;
};
#endif
