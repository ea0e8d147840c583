// tools.h -- the math helpers of the reference's tools.h that cross the class boundary
// (tools.h:48-56,64,68,93-94 in the reference), host-side, on top of fm3d_cv.h.
// Drawing and PCL viewers are out of scope.
#ifndef FM3D_HOST_TOOLS_H_
#define FM3D_HOST_TOOLS_H_
#include <sstream>
#include <string>
#include "../fm3d_cv.h"

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
#endif
