// fm3d_draw.cpp -- the two drawing helpers behind main.cpp's output artefacts (matches.pgm, projectedPatches.pgm) and the
// viewer it ends with, for builds that do not link the reference's tools.cpp (which needs PCL / VTK / Eigen).
// Behaviour follows tools.cpp:116-120,146-239; the rasterisers are cv::circle / cv::line (fm3d_cv.h, or the real OpenCV
// with -DFM3D_USE_OPENCV).  Compile with -DFM3D_EXTERNAL_TOOLS to leave all of these to the reference's own tools.cpp.
#ifndef FM3D_EXTERNAL_TOOLS
#include <cmath>

#include "include/tools.h"

cv::Scalar random_color(cv::RNG& rng) {
    const int color = (int)rng.next();
    return CV_RGB(color & 255, (color >> 8) & 255, (color >> 16) & 255);
}

void drawMatches(const cv::Mat& img1, const cv::Mat& img2, cv::Mat& window, const std::vector<cv::KeyPoint>& kpts1,
                 const std::vector<cv::KeyPoint>& kpts2, const std::vector<cv::DMatch>& matches, std::vector<cv::Scalar>& colors,
                 const std::vector<bool> outliersMask) {
    window = cv::Mat(cv::Size(img1.cols * 2, img1.rows), CV_8UC3, cv::Scalar(0));
    cv::Mat bgr[2];
    cv::cvtColor(img1, bgr[0], CV_GRAY2BGR);
    cv::cvtColor(img2, bgr[1], CV_GRAY2BGR);
    // img?BGR.copyTo(window(Rect(...))) (tools.cpp:156-157): the second frame starts at column img1.cols
    for (int k = 0; k < 2; k++) {
        const int rows = bgr[k].rows < window.rows ? bgr[k].rows : window.rows;
        int cols = bgr[k].cols;
        if (k * img1.cols + cols > window.cols) cols = window.cols - k * img1.cols;
        for (int r = 0; r < rows && cols > 0; r++)
            memcpy(window.ptr<unsigned char>(r) + (size_t)3 * k * img1.cols, bgr[k].ptr<unsigned char>(r), (size_t)3 * cols);
    }
    cv::RNG rng(0xFFF0FF0F);
    for (size_t i = 0; i < matches.size(); i++) {
        if (!outliersMask[i]) continue;                  // colours exist for inliers only (:162-169)
        const cv::Scalar color = random_color(rng);
        colors.push_back(color);
        cv::Point2f pt1 = kpts1.at(matches[i].queryIdx).pt, pt2 = kpts2.at(matches[i].trainIdx).pt;
        pt2.x = pt2.x + img1.cols;
        cv::circle(window, pt1, 4, color);               // Point2f -> Point: rounded, as cv::Point_'s conversion
        cv::circle(window, pt2, 4, color);
        cv::line(window, pt1, pt2, color);
    }
}

void drawBackProjectedPoints(const cv::Mat& input, cv::Mat& output, const std::vector<cv::Mat>& points, const std::vector<cv::Scalar>& colors) {
    cv::cvtColor(input, output, CV_GRAY2BGR);
    for (size_t i = 0; i < points.size(); i++) {
        const cv::Scalar& col = colors.at(i);
        for (int k = 0; k < points[i].rows; k++) {
            const double* p = points[i].ptr<double>(k);                          // S*S x 1 CV_64FC2
            const int x = (int)round(p[0]), y = (int)round(p[1]);
            // the reference tests `!(x > cols)` / `!(y > rows)` (:205-206) and so writes one column / one row past the image for
            // x == cols / y == rows; those two cases are skipped here (same deviation as D2 for the sampler)
            if (x < 0 || y < 0 || x >= output.cols || y >= output.rows) continue;
            unsigned char* o = output.ptr<unsigned char>(y) + 3 * x;
            o[0] = (unsigned char)col[0]; o[1] = (unsigned char)col[1]; o[2] = (unsigned char)col[2];
        }
    }
}

void drawBackProjectedPoints(const cv::Mat& input, cv::Mat& output, const cv::Mat& points, const cv::Scalar& colors) {
    output = input;                                          // shallow: the reference paints into the caller's image (:223)
    if (output.type() != CV_8UC3) return;                    // at<Vec3b> on anything else is undefined in the reference
    for (int k = 0; k < points.rows; k++) {
        const double* p = points.ptr<double>(k);
        const int x = (int)round(p[0]), y = (int)round(p[1]);
        if (x < 0 || y < 0 || x >= output.cols || y >= output.rows) continue;   // unchecked in the reference (:226-236)
        unsigned char* o = output.ptr<unsigned char>(y) + 3 * x;
        o[0] = (unsigned char)colors[0]; o[1] = (unsigned char)colors[1]; o[2] = (unsigned char)colors[2];
    }
}

void viewPointCloudNormalsFramesNeighborhoodAndGravity(const std::vector<std::vector<cv::Vec3d> >&, std::vector<cv::Vec3d>&,
                                                       const std::vector<cv::Scalar>&, std::vector<cv::Matx44d>&, cv::Vec3d&) {}
#endif  // FM3D_EXTERNAL_TOOLS
