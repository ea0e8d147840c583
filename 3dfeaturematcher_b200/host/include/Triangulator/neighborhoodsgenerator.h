// NeighborhoodsGenerator -- public interface of the reference class for the `square` method
// (Triangulator/neighborhoodsgenerator.h:81-97).  The circular variant is never called by
// main.cpp / mosaic.cpp (settings.yml: method: square) and is not provided: constructing with
// method: circular throws (the reference exit(-10)s only for unknown methods).
#ifndef FM3D_HOST_NEIGHBORHOODSGENERATOR_H_
#define FM3D_HOST_NEIGHBORHOODSGENERATOR_H_
#include <vector>
#include "../../fm3d_cv.h"

class NeighborhoodsGenerator {
public:
    NeighborhoodsGenerator(cv::FileStorage settings);
    void computeSquareNeighborhoodsByNormals(const std::vector<cv::Matx44d>& featuresFrames,
                                             std::vector<std::vector<cv::Vec3d> >& neighborhoodsVector);
    void computeSquareNeighborhoodByNormal(const cv::Matx44d& featureFrame, std::vector<cv::Vec3d>& neighborhood);
    void getReferenceSquaredNeighborhood(std::vector<cv::Vec3d>& neighborhood);
    double epsilon() const { return epsilon_; }
    double cmPerPixel() const { return cm_per_pixel_; }

private:
    NeighborhoodsGenerator();
    double epsilon_, cm_per_pixel_;
};
#endif
