// tools.h -- the helpers of the reference's tools.h that its clients and the four classes call
// (tools.h:43-94 in the reference), host-side, on top of fm3d_cv.h: the 4x4 / Rodrigues / spherical
// maths that crosses the class boundary, the two drawing helpers main.cpp uses for its output
// artefacts (drawMatches -> matches.pgm, drawBackProjectedPoints -> projectedPatches.pgm;
// main.cpp:139-141,190-194, tools.cpp:146-240) and the PCL viewer main.cpp ends with, as a
// function that shows nothing (the viewers are out of scope).
#ifndef FM3D_HOST_TOOLS_H_
#define FM3D_HOST_TOOLS_H_
#include <sstream>
#include <string>
#include <vector>
#include "../fm3d_cv.h"
// the reference's tools.h pulls PCL in for its clients (tools.h:38-41); do the same when some PCL is on the include path
// (the real one, or compat/)
#if defined(__has_include)
#if __has_include(<pcl/common/common_headers.h>)
#include <pcl/common/common_headers.h>
#endif
#endif

template <typename T> std::string NumberToString(T v) { std::ostringstream o; o << v; return o.str(); }

// [R|T] <-> 4x4 (tools.cpp:87-114); rotation vectors follow cv::Rodrigues
void composeTransformation(const cv::Matx33d& R, const cv::Vec3d& T, cv::Matx44d& G);
void decomposeTransformation(const cv::Matx44d& G, cv::Vec3d& r, cv::Vec3d& t);
void rodriguesToMatrix(const cv::Vec3d& r, cv::Matx33d& R);
void matrixToRodrigues(const cv::Matx33d& R, cv::Vec3d& r);
void getSkewMatrix(const cv::Vec3d& vec, cv::Matx33d& skew);
// (phi, theta) <-> unit vector (tools.cpp:767-777): theta = elevation, phi = azimuth
void car2sph(const cv::Vec3d& v, double& phi, double& theta);
void sph2car(const double phi, const double theta, cv::Vec3d& v);

// tools.cpp:116-120
cv::Scalar random_color(cv::RNG& rng);
// tools.cpp:146-186: both frames side by side (CV_8UC3), one random colour per INLIER match (appended to `colors`), two
// circles of radius 4 and the connecting line per inlier
void drawMatches(const cv::Mat& img1, const cv::Mat& img2, cv::Mat& window, const std::vector<cv::KeyPoint>& kpts1,
                 const std::vector<cv::KeyPoint>& kpts2, const std::vector<cv::DMatch>& matches, std::vector<cv::Scalar>& colors,
                 const std::vector<bool> outliersMask);
// tools.cpp:188-239: the image points of every patch painted into a BGR copy of the frame in the patch's colour
void drawBackProjectedPoints(const cv::Mat& input, cv::Mat& output, const std::vector<cv::Mat>& points, const std::vector<cv::Scalar>& colors);
void drawBackProjectedPoints(const cv::Mat& input, cv::Mat& output, const cv::Mat& points, const cv::Scalar& colors);
// tools.cpp:640-765: PCL viewer, blocks until the window is closed.  Here: returns at once.
void viewPointCloudNormalsFramesNeighborhoodAndGravity(const std::vector<std::vector<cv::Vec3d> >& neighborhoodsVector,
                                                       std::vector<cv::Vec3d>& normals, const std::vector<cv::Scalar>& colors,
                                                       std::vector<cv::Matx44d>& featuresFrames, cv::Vec3d& gravity);
#endif
