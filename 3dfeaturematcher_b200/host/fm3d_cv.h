// fm3d_cv.h -- the OpenCV types that appear in the public signatures of the reference's four
// classes (DescriptorsMatcher, SingleCameraTriangulator, NormalOptimizer,
// NeighborhoodsGenerator).
//
// With -DFM3D_USE_OPENCV the real <opencv2/opencv.hpp> is used and the adapters compile against
// the caller's OpenCV.  This image has no OpenCV C++ headers, so by default a minimal stand-in
// of exactly the used subset is provided in namespace cv: Mat (8U / 32F / 64F, 1-2 channels,
// reference-counted shallow copies like cv::Mat), Vec / Matx / Point / Size / Scalar, KeyPoint,
// DMatch, FileStorage / FileNode for the `%YAML:1.0` subset of build/settings.yml, and
// imread / imwrite for binary PGM.  It is interface plumbing only: no image processing.
#ifndef FM3D_CV_H_
#define FM3D_CV_H_

#ifdef FM3D_USE_OPENCV
#include <opencv2/opencv.hpp>
#else

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_64FC2 CV_MAKETYPE(CV_64F, 2)
#define CV_64FC3 CV_MAKETYPE(CV_64F, 3)
#define CV_LOAD_IMAGE_GRAYSCALE 0

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

template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point_<int> Point2i;

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

template <typename T> using Ptr = std::shared_ptr<T>;

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

// binary PGM (P5) in / out -- the only image format the reference's pipeline touches
inline Mat imread(const std::string& path, int = 0) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return Mat();
    char magic[3] = {0, 0, 0};
    int w = 0, h = 0, maxv = 0;
    auto skip = [&]() { int c; while ((c = fgetc(f)) != EOF) { if (c == '#') { while ((c = fgetc(f)) != EOF && c != '\n') {} } else if (c > ' ') { ungetc(c, f); break; } } };
    if (fscanf(f, "%2s", magic) != 1 || strcmp(magic, "P5") != 0) { fclose(f); return Mat(); }
    skip(); if (fscanf(f, "%d", &w) != 1) { fclose(f); return Mat(); }
    skip(); if (fscanf(f, "%d", &h) != 1) { fclose(f); return Mat(); }
    skip(); if (fscanf(f, "%d", &maxv) != 1) { fclose(f); return Mat(); }
    fgetc(f);
    Mat m(h, w, CV_8UC1);
    size_t got = fread(m.data, 1, (size_t)w * h, f);
    fclose(f);
    return got == (size_t)w * h ? m : Mat();
}
inline bool imwrite(const std::string& path, const Mat& m) {
    if (m.empty() || m.type() != CV_8UC1) return false;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "P5\n%d %d\n255\n", m.cols, m.rows);
    fwrite(m.data, 1, (size_t)m.rows * m.cols, f);
    fclose(f);
    return true;
}

}  // namespace cv
#endif  // FM3D_USE_OPENCV
#endif  // FM3D_CV_H_
