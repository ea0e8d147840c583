// Test harness: the rasterisers and the PxM writer of fm3d_cv.h behind a C interface, so that tests/test_draw_pins.py can
// pin them to cv2.circle / cv2.line / cv2.imwrite and drawMatches / drawBackProjectedPoints to their cv2 restatement.
#include <cstring>
#include <vector>

#include "tools.h"

extern "C" {

// prims: n x 6 ints (kind 0 circle / 1 line, x1, y1, x2-or-radius, y2, colour index); colours: 3 doubles each (B, G, R)
void fm3d_test_draw(unsigned char* bgr, int w, int h, const int* prims, int n, const double* colours) {
    cv::Mat img(h, w, CV_8UC3);
    memcpy(img.data, bgr, (size_t)w * h * 3);
    for (int i = 0; i < n; i++) {
        const int* p = prims + 6 * i;
        const double* c = colours + 3 * p[5];
        const cv::Scalar col(c[0], c[1], c[2]);
        if (p[0] == 0) cv::circle(img, cv::Point(p[1], p[2]), p[3], col);
        else cv::line(img, cv::Point(p[1], p[2]), cv::Point(p[3], p[4]), col);
    }
    memcpy(bgr, img.data, (size_t)w * h * 3);
}

int fm3d_test_imwrite(const char* path, const unsigned char* data, int w, int h, int channels) {
    cv::Mat img(h, w, channels == 3 ? CV_8UC3 : CV_8UC1);
    memcpy(img.data, data, (size_t)w * h * channels);
    return cv::imwrite(path, img) ? 0 : -1;
}

int fm3d_test_imread_gray(const char* path, unsigned char* out, int cap, int* w, int* h) {
    cv::Mat img = cv::imread(path, CV_LOAD_IMAGE_GRAYSCALE);
    if (img.empty()) return -1;
    *w = img.cols; *h = img.rows;
    if ((size_t)img.cols * img.rows > (size_t)cap) return -2;
    memcpy(out, img.data, (size_t)img.cols * img.rows);
    return 0;
}

// drawMatches: kp1 / kp2 as float pairs, matches as (q, t) int pairs, mask bytes; returns the number of colours
int fm3d_test_draw_matches(const unsigned char* img1, const unsigned char* img2, int w, int h, const float* kp1, int n1, const float* kp2,
                           int n2, const int* matches, const unsigned char* mask, int nm, unsigned char* window_bgr, double* colours_out) {
    cv::Mat a(h, w, CV_8UC1), b(h, w, CV_8UC1), window;
    memcpy(a.data, img1, (size_t)w * h);
    memcpy(b.data, img2, (size_t)w * h);
    std::vector<cv::KeyPoint> k1(n1), k2(n2);
    for (int i = 0; i < n1; i++) k1[i] = cv::KeyPoint(kp1[2 * i], kp1[2 * i + 1], 1.f);
    for (int i = 0; i < n2; i++) k2[i] = cv::KeyPoint(kp2[2 * i], kp2[2 * i + 1], 1.f);
    std::vector<cv::DMatch> m(nm);
    std::vector<bool> msk(nm);
    for (int i = 0; i < nm; i++) { m[i] = cv::DMatch(matches[2 * i], matches[2 * i + 1], 0.f); msk[i] = mask[i] != 0; }
    std::vector<cv::Scalar> colours;
    drawMatches(a, b, window, k1, k2, m, colours, msk);
    memcpy(window_bgr, window.data, (size_t)2 * w * h * 3);
    for (size_t i = 0; i < colours.size(); i++) for (int c = 0; c < 3; c++) colours_out[3 * i + c] = colours[i][c];
    return (int)colours.size();
}

// drawBackProjectedPoints (vector variant): npatch patches of npts image points each
void fm3d_test_draw_points(const unsigned char* img, int w, int h, const double* pts, int npatch, int npts, const double* colours,
                           unsigned char* out_bgr) {
    cv::Mat in(h, w, CV_8UC1), out;
    memcpy(in.data, img, (size_t)w * h);
    std::vector<cv::Mat> pv;
    std::vector<cv::Scalar> cv_;
    for (int i = 0; i < npatch; i++) {
        cv::Mat p(npts, 1, CV_64FC2);
        memcpy(p.data, pts + (size_t)i * npts * 2, sizeof(double) * 2 * npts);
        pv.push_back(p);
        cv_.push_back(cv::Scalar(colours[3 * i], colours[3 * i + 1], colours[3 * i + 2]));
    }
    drawBackProjectedPoints(in, out, pv, cv_);
    memcpy(out_bgr, out.data, (size_t)w * h * 3);
}

}  // extern "C"
