// fm3d_cv.h -- the OpenCV types that appear in the public signatures of the reference's four
// classes (DescriptorsMatcher, SingleCameraTriangulator, NormalOptimizer,
// NeighborhoodsGenerator).
//
// With -DFM3D_USE_OPENCV the real <opencv2/opencv.hpp> is used and the adapters compile against
// the caller's OpenCV.  This image has no OpenCV C++ headers, so by default a minimal stand-in
// of exactly the used subset is provided in namespace cv: Mat (8U / 32F / 64F, 1-2 channels,
// reference-counted shallow copies like cv::Mat), Vec / Matx / Point / Size / Scalar, KeyPoint,
// DMatch, FileStorage / FileNode for the `%YAML:1.0` subset of build/settings.yml,
// imread / imwrite for binary PGM / PPM, and what the reference's two drawing helpers use
// (tools.cpp:146-240: RNG, cvtColor GRAY2BGR, circle, line -- integer rasterisers restated from
// OpenCV's published algorithms and pinned to cv2 in tests/).  Interface plumbing only: nothing
// of the hot path lives here.
#ifndef FM3D_CV_H_
#define FM3D_CV_H_

#ifdef FM3D_USE_OPENCV
#include <opencv2/opencv.hpp>
#else

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_64FC2 CV_MAKETYPE(CV_64F, 2)
#define CV_64FC3 CV_MAKETYPE(CV_64F, 3)
#define CV_LOAD_IMAGE_GRAYSCALE 0
#define CV_GRAY2BGR 8
#define CV_RGB(r, g, b) cv::Scalar((b), (g), (r), 0)

namespace cv {

template <typename T, int N>
struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
    Vec(T a, T b) { static_assert(N == 2, ""); val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { static_assert(N == 3, ""); val[0] = a; val[1] = b; val[2] = c; }
    Vec(T a, T b, T c, T d) { static_assert(N == 4, ""); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
};
typedef Vec<double, 2> Vec2d;
typedef Vec<double, 3> Vec3d;
typedef Vec<double, 4> Vec4d;
typedef Vec<uchar, 3> Vec3b;
template <typename T, int N>
std::ostream& operator<<(std::ostream& o, const Vec<T, N>& v) {      // cv's formatter: [a, b, c]
    o << "[";
    for (int i = 0; i < N; i++) o << (i ? ", " : "") << v.val[i];
    return o << "]";
}

template <typename T, int M, int N>
struct Matx {
    T val[M * N];
    Matx() { for (int i = 0; i < M * N; i++) val[i] = T(0); }
    T& operator()(int i, int j) { return val[i * N + j]; }
    const T& operator()(int i, int j) const { return val[i * N + j]; }
};
typedef Matx<double, 3, 3> Matx33d;
typedef Matx<double, 3, 4> Matx34d;
typedef Matx<double, 4, 4> Matx44d;

inline int cvRound(double v) { return (int)std::nearbyint(v); }      // round half to even, as cv::saturate_cast<int>
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    // cv::Point_'s converting constructor: saturate_cast, i.e. rounding when the target is integral
    template <typename U> Point_(const Point_<U>& p) : x(conv(p.x)), y(conv(p.y)) {}
private:
    template <typename U> static T conv(U v) { return std::is_integral<T>::value && !std::is_integral<U>::value ? (T)cvRound((double)v) : (T)v; }
};
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point_<int> Point2i;
typedef Point2i Point;
template <typename T> std::ostream& operator<<(std::ostream& o, const Point_<T>& p) { return o << "[" << p.x << ", " << p.y << "]"; }

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s) : pt(x, y), size(s), angle(-1), response(0), octave(0), class_id(-1) {}
};

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4e38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(0), distance(d) {}
};

// Dense 2-D matrix with shared storage (copy = shallow, clone = deep), continuous rows.
class Mat {
public:
    int rows, cols;
    uchar* data;
    Mat() : rows(0), cols(0), data(nullptr), type_(CV_8UC1) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(Size s, int type, const Scalar& v) { create(s.height, s.width, type); setTo(v); }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        buf_ = std::shared_ptr<std::vector<uchar>>(new std::vector<uchar>((size_t)r * c * elemSize(), 0));
        data = buf_->empty() ? nullptr : buf_->data();
    }
    static Mat zeros(Size s, int type) { return Mat(s, type); }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == CV_8U ? 1 : depth() == CV_32F ? 4 : 8); }
    size_t step() const { return (size_t)cols * elemSize(); }
    size_t step1(int = 0) const { return (size_t)cols * channels(); }   // cv::Mat::step1: row pitch in elements of one channel
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }
    Size size() const { return Size(cols, rows); }
    Mat clone() const { Mat m(rows, cols, type_); if (data) memcpy(m.data, data, buf_->size()); return m; }
    Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.type_ = type_; m.buf_ = buf_; m.data = data + (size_t)r * step(); return m; }
    void copyTo(Mat m) const { memcpy(m.data, data, (size_t)rows * step()); }
    void setTo(const Scalar& v) {
        for (size_t i = 0; i < (size_t)rows * cols; i++)
            for (int c = 0; c < channels(); c++) {
                if (depth() == CV_8U) data[i * channels() + c] = (uchar)v[c];
                else if (depth() == CV_32F) ((float*)data)[i * channels() + c] = (float)v[c];
                else ((double*)data)[i * channels() + c] = v[c];
            }
    }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step() + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step() + (size_t)c * sizeof(T)); }
    template <typename T> T& at(int i) { return ((T*)data)[i]; }
    template <typename T> const T& at(int i) const { return ((const T*)data)[i]; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step()); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step()); }
private:
    int type_;
    std::shared_ptr<std::vector<uchar>> buf_;
};

// cv::Ptr.  The assignment from an object (a shallow copy, like `new cv::Mat(img)` in descriptorsmatcher.cpp:40-41) is NOT
// part of OpenCV's cv::Ptr: the reference's work-in-progress mosaic.cpp:39-40 writes `imgA_ = imgA;` with a cv::Ptr<cv::Mat>
// on the left, and it is provided here so that file builds against the adapters.
template <typename T>
class Ptr : public std::shared_ptr<T> {
public:
    Ptr() {}
    Ptr(T* p) : std::shared_ptr<T>(p) {}
    Ptr(const std::shared_ptr<T>& p) : std::shared_ptr<T>(p) {}
    Ptr& operator=(const T& obj) { this->reset(new T(obj)); return *this; }
    bool empty() const { return !this->get(); }
};

// base class of the reference's MOSAIC (mosaic.h:47); nothing of it is used
class DescriptorExtractor {
public:
    virtual ~DescriptorExtractor() {}
};

// ---- `%YAML:1.0` subset of cv::FileStorage: nested maps by indentation, scalars, [a, b, c] lists
class FileNode {
public:
    FileNode() {}
    bool empty() const { return !n_; }
    FileNode operator[](const char* key) const {
        if (!n_) return FileNode();
        auto it = n_->children.find(key);
        return it == n_->children.end() ? FileNode() : FileNode(it->second);
    }
    FileNode operator[](const std::string& key) const { return (*this)[key.c_str()]; }
    operator double() const { return n_ && !n_->scalar.empty() ? atof(n_->scalar.c_str()) : 0.0; }
    operator int() const { return n_ && !n_->scalar.empty() ? (int)atof(n_->scalar.c_str()) : 0; }
    operator std::string() const { return n_ ? n_->scalar : std::string(); }
    const std::vector<std::string>& list() const { static std::vector<std::string> e; return n_ ? n_->seq : e; }
    struct Node { std::string scalar; std::vector<std::string> seq; std::map<std::string, std::shared_ptr<Node>> children; };
    explicit FileNode(std::shared_ptr<Node> n) : n_(n) {}
private:
    std::shared_ptr<Node> n_;
};
inline void operator>>(const FileNode& n, double& v) { v = (double)n; }
inline void operator>>(const FileNode& n, int& v) { v = (int)n; }
inline void operator>>(const FileNode& n, std::string& v) { v = (std::string)n; }
inline void operator>>(const FileNode& n, std::vector<double>& v) {
    v.clear();
    for (const std::string& s : n.list()) v.push_back(atof(s.c_str()));
}

class FileStorage {
public:
    enum { READ = 0 };
    FileStorage() {}
    FileStorage(const std::string& path, int) { open(path, READ); }
    bool open(const std::string& path, int) {
        std::ifstream f(path.c_str());
        if (!f) return false;
        root_.reset(new FileNode::Node());
        std::vector<std::pair<int, std::shared_ptr<FileNode::Node>>> stack;
        stack.push_back(std::make_pair(-1, root_));
        std::string line;
        while (std::getline(f, line)) {
            size_t hash = line.find('#');
            if (hash != std::string::npos) line = line.substr(0, hash);
            if (line.find("%YAML") == 0) continue;
            size_t first = line.find_first_not_of(" \t");
            if (first == std::string::npos) continue;
            size_t colon = line.find(':', first);
            if (colon == std::string::npos) continue;
            std::string key = trim(line.substr(first, colon - first)), val = trim(line.substr(colon + 1));
            while (stack.size() > 1 && stack.back().first >= (int)first) stack.pop_back();
            std::shared_ptr<FileNode::Node> node(new FileNode::Node());
            stack.back().second->children[key] = node;
            if (val.empty()) stack.push_back(std::make_pair((int)first, node));
            else if (val[0] == '[') {
                std::string body = val.substr(1, val.find(']') == std::string::npos ? std::string::npos : val.find(']') - 1);
                std::stringstream ss(body);
                std::string item;
                while (std::getline(ss, item, ',')) node->seq.push_back(trim(item));
            } else {
                if (val.size() >= 2 && (val[0] == '"' || val[0] == '\'')) val = val.substr(1, val.size() - 2);
                node->scalar = val;
            }
        }
        return true;
    }
    bool isOpened() const { return (bool)root_; }
    void release() { root_.reset(); }
    FileNode operator[](const char* key) const { return FileNode(root_)[key]; }
    FileNode operator[](const std::string& key) const { return FileNode(root_)[key]; }
private:
    static std::string trim(const std::string& s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
    std::shared_ptr<FileNode::Node> root_;
};

// binary PGM (P5) / PPM (P6) in / out -- the only image formats the reference's pipeline touches.  Like cv::imread with
// CV_LOAD_IMAGE_GRAYSCALE a P6 file is reduced to one channel (cv's fixed-point BT.601 weights).
inline Mat imread(const std::string& path, int = 0) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return Mat();
    char magic[3] = {0, 0, 0};
    int w = 0, h = 0, maxv = 0;
    auto skip = [&]() { int c; while ((c = fgetc(f)) != EOF) { if (c == '#') { while ((c = fgetc(f)) != EOF && c != '\n') {} } else if (c > ' ') { ungetc(c, f); break; } } };
    if (fscanf(f, "%2s", magic) != 1 || (strcmp(magic, "P5") != 0 && strcmp(magic, "P6") != 0)) { fclose(f); return Mat(); }
    const int cn = magic[1] == '6' ? 3 : 1;
    skip(); if (fscanf(f, "%d", &w) != 1) { fclose(f); return Mat(); }
    skip(); if (fscanf(f, "%d", &h) != 1) { fclose(f); return Mat(); }
    skip(); if (fscanf(f, "%d", &maxv) != 1) { fclose(f); return Mat(); }
    fgetc(f);
    if (w <= 0 || h <= 0 || maxv != 255) { fclose(f); return Mat(); }
    std::vector<uchar> raw((size_t)w * h * cn);
    size_t got = fread(raw.data(), 1, raw.size(), f);
    fclose(f);
    if (got != raw.size()) return Mat();
    Mat m(h, w, CV_8UC1);
    if (cn == 1) memcpy(m.data, raw.data(), raw.size());
    else for (size_t i = 0; i < (size_t)w * h; i++)      // RGB in the file; cv: (R*4899 + G*9617 + B*1868 + 8192) >> 14
        m.data[i] = (uchar)((raw[3 * i] * 4899 + raw[3 * i + 1] * 9617 + raw[3 * i + 2] * 1868 + 8192) >> 14);
    return m;
}
// cv::imwrite's PxM encoder picks the format by the channel count, not by the extension: main.cpp:141,194 write CV_8UC3
// images to "*.pgm" and get binary PPM (P6) with the channels swapped from cv's BGR to the file's RGB.
inline bool imwrite(const std::string& path, const Mat& m) {
    if (m.empty() || (m.type() != CV_8UC1 && m.type() != CV_8UC3)) return false;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const int cn = m.channels();
    fprintf(f, "P%c\n%d %d\n255\n", cn == 3 ? '6' : '5', m.cols, m.rows);
    if (cn == 1) fwrite(m.data, 1, (size_t)m.rows * m.cols, f);
    else {
        std::vector<uchar> row((size_t)m.cols * 3);
        for (int r = 0; r < m.rows; r++) {
            const uchar* p = m.ptr<uchar>(r);
            for (int c = 0; c < m.cols; c++) { row[3 * c] = p[3 * c + 2]; row[3 * c + 1] = p[3 * c + 1]; row[3 * c + 2] = p[3 * c]; }
            fwrite(row.data(), 1, row.size(), f);
        }
    }
    fclose(f);
    return true;
}

// cv::RNG: multiply-with-carry, as used by the reference's random_color (tools.cpp:116-120)
class RNG {
public:
    uint64_t state;
    RNG() : state(0xffffffff) {}
    RNG(uint64_t s) : state(s ? s : 0xffffffff) {}
    unsigned next() { state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state; }
};

// cv::cvtColor, CV_GRAY2BGR only (tools.cpp:153-154,190)
inline void cvtColor(const Mat& src, Mat& dst, int code) {
    if (code != CV_GRAY2BGR || src.type() != CV_8UC1) { dst = Mat(); return; }
    Mat out(src.rows, src.cols, CV_8UC3);
    for (int r = 0; r < src.rows; r++) {
        const uchar* s = src.ptr<uchar>(r);
        uchar* d = out.ptr<uchar>(r);
        for (int c = 0; c < src.cols; c++) d[3 * c] = d[3 * c + 1] = d[3 * c + 2] = s[c];
    }
    dst = out;
}

namespace detail {
inline void put_pixel(Mat& img, int x, int y, const uchar* color) { memcpy(img.data + (size_t)y * img.step() + (size_t)x * img.elemSize(), color, img.elemSize()); }
inline void scalar_to_raw(const Scalar& s, const Mat& img, uchar* buf) {       // saturate_cast<uchar>(cvRound) per channel
    for (int c = 0; c < img.channels() && c < 4; c++) { int v = cvRound(s[c]); buf[c] = (uchar)(v < 0 ? 0 : v > 255 ? 255 : v); }
}
// cv::clipLine on the rectangle [0, w) x [0, h) (64-bit intermediates as in OpenCV)
inline bool clip_line(int w, int h, long long& x1, long long& y1, long long& x2, long long& y2) {
    if (w <= 0 || h <= 0) return false;
    const long long right = w - 1, bottom = h - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) { a = c1 < 8 ? 0 : bottom; x1 += (long long)((double)(a - y1) * (x2 - x1) / (y2 - y1)); y1 = a; c1 = (x1 < 0) + (x1 > right) * 2; }
        if (c2 & 12) { a = c2 < 8 ? 0 : bottom; x2 += (long long)((double)(a - y2) * (x2 - x1) / (y2 - y1)); y2 = a; c2 = (x2 < 0) + (x2 > right) * 2; }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) { a = c1 == 1 ? 0 : right; y1 += (long long)((double)(a - x1) * (y2 - y1) / (x2 - x1)); x1 = a; c1 = 0; }
            if (c2) { a = c2 == 1 ? 0 : right; y2 += (long long)((double)(a - x2) * (y2 - y1) / (x2 - x1)); x2 = a; c2 = 0; }
        }
    }
    return (c1 | c2) == 0;
}
}  // namespace detail

// cv::line with thickness 1, 8-connected, no shift (tools.cpp:181): cv::LineIterator's Bresenham walk, left to right
inline void line(Mat& img, Point p1, Point p2, const Scalar& color, int = 1, int = 8, int = 0) {
    uchar raw[4];
    detail::scalar_to_raw(color, img, raw);
    long long x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    if ((unsigned long long)x1 >= (unsigned long long)img.cols || (unsigned long long)x2 >= (unsigned long long)img.cols ||
        (unsigned long long)y1 >= (unsigned long long)img.rows || (unsigned long long)y2 >= (unsigned long long)img.rows)
        if (!detail::clip_line(img.cols, img.rows, x1, y1, x2, y2)) return;
    long long dx = x2 - x1, dy = y2 - y1;
    if (dx < 0) { dx = -dx; dy = -dy; x1 = x2; y1 = y2; }       // left to right: start from the left end
    long long sx = 1, sy = dy < 0 ? -1 : 1;
    if (dy < 0) dy = -dy;
    const bool steep = dy > dx;                                  // walk along the longer axis
    const long long dmaj = steep ? dy : dx, dmin = steep ? dx : dy;
    long long err = dmaj - 2 * dmin;
    long long x = x1, y = y1;
    for (long long i = 0; i <= dmaj; i++) {
        detail::put_pixel(img, (int)x, (int)y, raw);
        const bool diag = err < 0;
        err += -2 * dmin + (diag ? 2 * dmaj : 0);
        if (steep) { y += sy; if (diag) x += sx; }
        else { x += sx; if (diag) y += sy; }
    }
}

// cv::circle with thickness 1, 8-connected, no shift (tools.cpp:178-179): OpenCV's midpoint circle with its clipping
inline void circle(Mat& img, Point center, int radius, const Scalar& color, int = 1, int = 8, int = 0) {
    uchar raw[4];
    detail::scalar_to_raw(color, img, raw);
    const int W = img.cols, H = img.rows;
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    auto put = [&](int x, int y) { if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) detail::put_pixel(img, x, y, raw); };
    while (dx >= dy) {
        const int y11 = center.y - dy, y12 = center.y + dy, y21 = center.y - dx, y22 = center.y + dx;
        const int x11 = center.x - dx, x12 = center.x + dx, x21 = center.x - dy, x22 = center.x + dy;
        put(x11, y11); put(x11, y12); put(x12, y11); put(x12, y12);
        put(x21, y21); put(x21, y22); put(x22, y21); put(x22, y22);
        dy++;
        err += plus;
        plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

}  // namespace cv
#endif  // FM3D_USE_OPENCV
#endif  // FM3D_CV_H_
