// fm3d_host.cpp -- the reference's four classes as thin adapters over the C-ABI of libfm3d.
//
// Same class names, method names, argument meaning and output conventions as the reference
// (DescriptorsMatcher/descriptorsmatcher.cpp, Triangulator/{singlecameratriangulator,
// normaloptimizer,neighborhoodsgenerator}.cpp): outputs that the reference appends to are
// appended to, outputs it clears are cleared, failed features are erased from points3D in
// place.  What differs is documented in the headers (injected features, no visualiser thread,
// errors as std::runtime_error instead of exit()).  No arithmetic of the hot path lives here.
#include <cmath>
#include <stdexcept>
#include <string>
#include <thread>
#include <memory>
#include <vector>

#include "../../include/fm3d.h"
#include "include/DescriptorsMatcher/descriptorsmatcher.h"
#include "include/Triangulator/neighborhoodsgenerator.h"
#include "include/Triangulator/normaloptimizer.h"
#include "include/Triangulator/singlecameratriangulator.h"
#include "include/tools.h"

namespace {

// One fm3d_ctx per GPU of the process.  FM3D_DEVICE=<i> picks a single device (default 0);
// FM3D_DEVICES=<i,j,...> lists several: matching is sharded by query keypoint and the normal search by feature
// over all of them (contiguous shards, so concatenating the shards reproduces the single-GPU order); triangulation
// and patch extraction stay on the first.  What every GPU needs in full -- the train descriptors, both frames --
// is uploaded ONCE to the first GPU and replicated with one ncclBroadcast over NVLink (fm3d_comm_init_all +
// fm3d_broadcast_all_dev); when NCCL cannot be loaded every GPU gets its own host upload instead (FM3D_NO_NCCL=1
// forces that).
bool& host_comm_ok() { static bool ok = false; return ok; }

std::vector<fm3d_ctx*>& host_ctxs() {
    static std::vector<fm3d_ctx*> ctxs;
    if (ctxs.empty()) {
        std::vector<int> devs;
        if (const char* e = getenv("FM3D_DEVICES")) {
            for (const char* p = e; *p;) {
                char* end = nullptr;
                const long d = strtol(p, &end, 10);
                if (end == p) break;
                devs.push_back((int)d);
                p = *end == ',' ? end + 1 : end;
            }
        }
        if (devs.empty()) devs.push_back(getenv("FM3D_DEVICE") ? atoi(getenv("FM3D_DEVICE")) : 0);
        for (size_t k = 0; k < devs.size(); k++) {
            fm3d_ctx* c = nullptr;
            const int rc = fm3d_ctx_create(devs[k], &c);
            if (rc != FM3D_OK) {
                for (size_t j = 0; j < ctxs.size(); j++) fm3d_ctx_destroy(ctxs[j]);
                ctxs.clear();
                throw std::runtime_error("fm3d: no usable sm_100 GPU at device " + std::to_string(devs[k]) +
                                         " (libfm3d has no CPU path), code " + std::to_string(rc));
            }
            ctxs.push_back(c);
        }
        if (ctxs.size() > 1 && !getenv("FM3D_NO_NCCL"))
            host_comm_ok() = fm3d_comm_init_all(ctxs.data(), (int)ctxs.size()) == FM3D_OK;
    }
    return ctxs;
}

fm3d_ctx* host_ctx() { return host_ctxs()[0]; }

// fn(context index) on every context, concurrently when there are several; returns the first failure
template <typename F>
void for_each_ctx(const char* what, F fn) {
    std::vector<fm3d_ctx*>& cs = host_ctxs();
    std::vector<int> rc(cs.size(), FM3D_OK);
    if (cs.size() == 1) {
        rc[0] = fn(0);
    } else {
        std::vector<std::thread> th;
        for (size_t k = 0; k < cs.size(); k++) th.emplace_back([&rc, &fn, k]() { rc[k] = fn((int)k); });
        for (size_t k = 0; k < th.size(); k++) th[k].join();
    }
    for (size_t k = 0; k < cs.size(); k++)
        if (rc[k] != FM3D_OK) throw std::runtime_error(std::string("fm3d: ") + what + " (context " + std::to_string(k) + "): " + fm3d_last_error(cs[k]));
}

void check(fm3d_ctx* ctx, int rc, const char* what) {
    if (rc != FM3D_OK) throw std::runtime_error(std::string("fm3d: ") + what + ": " + fm3d_last_error(ctx));
}

// ---- several GPUs: grow-only device buffers per (context, slot), and replication of a host buffer to all contexts
enum { SLOT_TRAIN = 0, SLOT_QUERY, SLOT_IDX, SLOT_DIST, SLOT_IMG1, SLOT_IMG2, N_SLOTS };
struct DevBuf { void* p = nullptr; size_t cap = 0; };

void* dev_buf(int k, int slot, size_t bytes) {
    static std::vector<std::vector<DevBuf> > bufs;
    if (bufs.size() < host_ctxs().size()) bufs.resize(host_ctxs().size(), std::vector<DevBuf>(N_SLOTS));
    DevBuf& b = bufs[k][slot];
    if (b.cap < bytes) {
        fm3d_ctx* c = host_ctxs()[k];
        if (b.p) check(c, fm3d_dev_free(c, b.p), "dev_free");
        b.p = nullptr; b.cap = 0;
        check(c, fm3d_dev_malloc(c, bytes, &b.p), "dev_malloc");
        b.cap = bytes;
    }
    return b.p;
}

// host bytes -> slot `slot` of every context.  One upload + ONE ncclBroadcast over NVLink, or one upload per GPU.
std::vector<void*> replicate(int slot, const void* host, size_t bytes) {
    std::vector<fm3d_ctx*>& cs = host_ctxs();
    std::vector<void*> dev(cs.size());
    for (size_t k = 0; k < cs.size(); k++) dev[k] = dev_buf((int)k, slot, bytes);     // same thread: allocation is not re-entrant
    if (host_comm_ok()) {
        check(cs[0], fm3d_copy_h2d(cs[0], dev[0], host, bytes), "copy_h2d");
        check(cs[0], fm3d_broadcast_all_dev(cs.data(), (int)cs.size(), dev.data(), bytes, 0), "broadcast");
        for (size_t k = 0; k < cs.size(); k++) check(cs[k], fm3d_sync(cs[k]), "sync");
    } else {
        for_each_ctx("copy_h2d", [&](int k) { return fm3d_copy_h2d(cs[k], dev[k], host, bytes); });
    }
    return dev;
}

// continuous copy of a matrix's rows (descriptor matrices usually are; a ROI is not)
const unsigned char* packed_rows(const cv::Mat& m, std::vector<unsigned char>& tmp) {
    const size_t row = (size_t)m.cols * m.elemSize();
    if (m.isContinuous()) return m.ptr<unsigned char>(0);
    tmp.resize(row * m.rows);
    for (int r = 0; r < m.rows; r++) memcpy(tmp.data() + row * r, m.ptr<unsigned char>(r), row);
    return tmp.data();
}

// Exact 2-NN of every row of q in t, query rows sharded over the GPUs of the process (descriptorsmatcher.cpp:84-85,101,117).
void knn_sharded(bool binary, const cv::Mat& q, const cv::Mat& t, std::vector<int32_t>& idx, std::vector<float>& dist) {
    std::vector<fm3d_ctx*>& cs = host_ctxs();
    const int G = (int)cs.size(), nq = q.rows, nt = t.rows, cols = q.cols;
    const size_t row = (size_t)cols * q.elemSize();
    std::vector<unsigned char> tq, tt;
    const unsigned char* qh = packed_rows(q, tq);
    const unsigned char* th = packed_rows(t, tt);
    std::vector<void*> train = replicate(SLOT_TRAIN, th, row * nt);
    std::vector<void*> dq(G), di(G), dd(G);
    for (int k = 0; k < G; k++) {
        const int lo = (int)((long long)nq * k / G), hi = (int)((long long)nq * (k + 1) / G), n = hi - lo;
        dq[k] = dev_buf(k, SLOT_QUERY, row * (n > 0 ? n : 1));
        di[k] = dev_buf(k, SLOT_IDX, sizeof(int32_t) * 2 * (n > 0 ? n : 1));
        dd[k] = dev_buf(k, SLOT_DIST, sizeof(float) * 2 * (n > 0 ? n : 1));
    }
    for_each_ctx("knn (sharded)", [&](int k) {
        const int lo = (int)((long long)nq * k / G), hi = (int)((long long)nq * (k + 1) / G), n = hi - lo;
        if (n <= 0) return (int)FM3D_OK;
        fm3d_ctx* c = cs[k];
        if (int rc = fm3d_copy_h2d(c, dq[k], qh + row * lo, row * n)) return rc;
        int rc = binary ? fm3d_match_knn2_hamming_dev(c, (const uint8_t*)dq[k], n, (const uint8_t*)train[k], nt, cols, (int32_t*)di[k], (float*)dd[k])
                        : fm3d_match_knn2_f32_dev(c, (const float*)dq[k], n, (const float*)train[k], nt, cols, (int32_t*)di[k], (float*)dd[k]);
        if (rc) return rc;
        if (int r2 = fm3d_copy_d2h(c, idx.data() + 2 * (size_t)lo, di[k], sizeof(int32_t) * 2 * n)) return r2;
        return fm3d_copy_d2h(c, dist.data() + 2 * (size_t)lo, dd[k], sizeof(float) * 2 * n);
    });
}

}  // namespace

// ------------------------------------------------------------------------------------ tools
void rodriguesToMatrix(const cv::Vec3d& r, cv::Matx33d& R) {
    const double th = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R(i, j) = i == j ? 1.0 : 0.0;
    if (th < 2.220446049250313e-16) return;
    const double c = std::cos(th), s = std::sin(th), c1 = 1 - c;
    const double k[3] = {r[0] / th, r[1] / th, r[2] / th};
    const double kx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R(i, j) = c * (i == j ? 1.0 : 0.0) + c1 * k[i] * k[j] + s * kx[i * 3 + j];
}

void matrixToRodrigues(const cv::Matx33d& R, cv::Vec3d& r) {
    const double rx = R(2, 1) - R(1, 2), ry = R(0, 2) - R(2, 0), rz = R(1, 0) - R(0, 1);
    const double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R(0, 0) + R(1, 1) + R(2, 2) - 1) * 0.5;
    c = c > 1 ? 1 : (c < -1 ? -1 : c);
    const double th = std::acos(c);
    if (s < 1e-5) {
        if (c > 0) { r = cv::Vec3d(0, 0, 0); return; }
        double t = (R(0, 0) + 1) * 0.5; double x = std::sqrt(t > 0 ? t : 0);
        t = (R(1, 1) + 1) * 0.5; double y = std::sqrt(t > 0 ? t : 0) * (R(0, 1) < 0 ? -1.0 : 1.0);
        t = (R(2, 2) + 1) * 0.5; double z = std::sqrt(t > 0 ? t : 0) * (R(0, 2) < 0 ? -1.0 : 1.0);
        if (std::fabs(x) < std::fabs(y) && std::fabs(x) < std::fabs(z) && (R(1, 2) > 0) != (y * z > 0)) z = -z;
        const double n = th / std::sqrt(x * x + y * y + z * z);
        r = cv::Vec3d(x * n, y * n, z * n);
        return;
    }
    const double vth = 1 / (2 * s) * th;
    r = cv::Vec3d(rx * vth, ry * vth, rz * vth);
}

// The five helpers below are also defined by the reference's tools.cpp (tools.cpp:87-127, 767-777).  When the
// adapters are linked next to that file (main.cpp keeps tools.cpp for its drawing and viewer
// functions), compile this file with -DFM3D_EXTERNAL_TOOLS so that there is one definition of each.
#ifndef FM3D_EXTERNAL_TOOLS
void composeTransformation(const cv::Matx33d& R, const cv::Vec3d& T, cv::Matx44d& G) {
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) G(i, j) = R(i, j); G(i, 3) = T(i); }
    G(3, 0) = 0; G(3, 1) = 0; G(3, 2) = 0; G(3, 3) = 1;
}

void decomposeTransformation(const cv::Matx44d& G, cv::Vec3d& r, cv::Vec3d& t) {
    cv::Matx33d R;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R(i, j) = G(i, j); t(i) = G(i, 3); }
    matrixToRodrigues(R, r);
}

void getSkewMatrix(const cv::Vec3d& v, cv::Matx33d& k) {
    k(0, 0) = 0; k(0, 1) = -v[2]; k(0, 2) = v[1];
    k(1, 0) = v[2]; k(1, 1) = 0; k(1, 2) = -v[0];
    k(2, 0) = -v[1]; k(2, 1) = v[0]; k(2, 2) = 0;
}

void car2sph(const cv::Vec3d& v, double& phi, double& theta) {
    theta = std::atan2(v[2], std::sqrt(v[0] * v[0] + v[1] * v[1]));
    phi = std::atan2(v[1], v[0]);
}

void sph2car(const double phi, const double theta, cv::Vec3d& v) {
    v[0] = std::cos(theta) * std::cos(phi);
    v[1] = std::cos(theta) * std::sin(phi);
    v[2] = std::sin(theta);
}
#endif  // FM3D_EXTERNAL_TOOLS

// ------------------------------------------------------------------------- DescriptorsMatcher
DescriptorsMatcher::DescriptorsMatcher(cv::FileStorage& fs, cv::Mat& frame_a, cv::Mat& frame_b)
    : image_a_(frame_a), image_b_(frame_b), binary_(false), have_features_(false) {
    std::string extractorType = (std::string)fs["FeatureOptions"]["ExtractorType"];
    extractor_type_ = extractorType;
    // the reference picks the LSH (binary) matcher for these (descriptorsmatcher.cpp:64-67)
    binary_ = extractorType == "ORB" || extractorType == "BRISK" || extractorType == "FREAK";
    // generateDetector (descriptorsmatcher.cpp:176-289)
    detector_type_ = (std::string)fs["FeatureOptions"]["DetectorType"];
    detector_mode_ = (std::string)fs["FeatureOptions"]["DetectorMode"];
    fast_threshold_ = (int)fs["FeatureOptions"]["FastDetector"]["Threshold"];                 // :218
    fast_nonmax_ = (int)fs["FeatureOptions"]["FastDetector"]["NonMaxSuppression"] > 0;        // :219-220
    adaptive_min_ = (int)fs["FeatureOptions"]["Adaptive"]["MinFeatures"];                     // :194-197
    adaptive_max_ = (int)fs["FeatureOptions"]["Adaptive"]["MaxFeatures"];
    adaptive_iters_ = (int)fs["FeatureOptions"]["Adaptive"]["MaxIters"];
    // generateDetector / generateExtractor for SIFT (descriptorsmatcher.cpp:243-256, :304-314): cv::SIFT(NumFeatures,
    // NumOctaveLayers, ContrastThreshold, EdgeThreshold, Sigma).  build/settings.yml carries no SiftDetector block (the
    // reference would then hand cv::SIFT zeros): a key that is absent keeps cv::SIFT's own default.  cv::ORB's ScaleFactor /
    // NumLevels act only on keypoints of octave > 0 and cv::BRISK::compute uses neither Threshold nor Octaves: with FAST
    // keypoints (octave 0) they change nothing.
    if (!fs["FeatureOptions"]["SiftDetector"].empty()) {
        const cv::FileNode sd = fs["FeatureOptions"]["SiftDetector"];
        if (!sd["NumFeatures"].empty()) sift_nfeatures_ = (int)sd["NumFeatures"];
        if (!sd["NumOctaveLayers"].empty()) sift_layers_ = (int)sd["NumOctaveLayers"];
        if (!sd["ContrastThreshold"].empty()) sift_contrast_ = (double)sd["ContrastThreshold"];
        if (!sd["EdgeThreshold"].empty()) sift_edge_ = (double)sd["EdgeThreshold"];
        if (!sd["Sigma"].empty()) sift_sigma_ = (double)sd["Sigma"];
        if ((detector_type_ == "SIFT" || extractorType == "SIFT") && (sift_layers_ < 1 || sift_layers_ > 5 || !(sift_sigma_ > 0)))
            throw std::runtime_error("fm3d: FeatureOptions.SiftDetector needs 1 <= NumOctaveLayers <= 5 and Sigma > 0");
    }
    if (!fs["FeatureOptions"]["OrbDetector"].empty()) {
        const cv::FileNode od = fs["FeatureOptions"]["OrbDetector"];
        if (!od["NumFeatures"].empty()) orb_nfeatures_ = (int)od["NumFeatures"];
        if (!od["ScaleFactor"].empty()) orb_scale_ = (double)(float)(double)od["ScaleFactor"];       // cv::ORB takes a float
        if (!od["NumLevels"].empty()) orb_levels_ = (int)od["NumLevels"];
        if (detector_type_ == "ORB" && (orb_nfeatures_ < 0 || !(orb_scale_ > 1.0) || orb_levels_ < 1 || orb_levels_ > 16))
            throw std::runtime_error("fm3d: FeatureOptions.OrbDetector needs NumFeatures >= 0, ScaleFactor > 1 and 1 <= NumLevels <= 16");
    }
    host_ctx();
}

DescriptorsMatcher::~DescriptorsMatcher() {}

void DescriptorsMatcher::setFeatures(const std::vector<cv::KeyPoint>& ka, const cv::Mat& da,
                                     const std::vector<cv::KeyPoint>& kb, const cv::Mat& db) {
    kpts_a_ = ka; kpts_b_ = kb; desc_a_ = da; desc_b_ = db;
    have_features_ = true;
}

void DescriptorsMatcher::features(std::vector<cv::KeyPoint>& ka, std::vector<cv::KeyPoint>& kb, cv::Mat& da, cv::Mat& db) {
    // Injected features (setFeatures) are the only way around detection: like the reference (descriptorsmatcher.cpp:110-115)
    // every call otherwise re-detects and OVERWRITES the four output arguments, whatever a previous call left in them.
    if (have_features_) {
        if ((int)kpts_a_.size() != desc_a_.rows || (int)kpts_b_.size() != desc_b_.rows)
            throw std::runtime_error("fm3d: setFeatures needs one descriptor row per keypoint (" + std::to_string(kpts_a_.size()) + " / " +
                                     std::to_string(desc_a_.rows) + ", " + std::to_string(kpts_b_.size()) + " / " + std::to_string(desc_b_.rows) + ")");
        ka = kpts_a_; kb = kpts_b_; da = desc_a_; db = desc_b_;
        return;
    }
    // descriptorsmatcher.cpp:110-115: detect on both frames, then compute on both frames.  DetectorType FAST
    // (STATIC) with ExtractorType SIFT, BRISK or ORB runs on the GPU (K10 + K11 / K12 / K13); the other detectors /
    // extractors of the reference (SURF, STAR, MSER, FREAK, ORB's own detector) are upstream code this library does not carry.
    // DetectorType SIFT (STATIC) with ExtractorType SIFT: cv::SIFT's scale-space detector and its descriptors on the pyramid
    // layers (K14 + K11).
    if ((detector_type_ == "FAST" && (detector_mode_ == "STATIC" || detector_mode_ == "ADAPTIVE") && (extractor_type_ == "SIFT" || extractor_type_ == "BRISK" || extractor_type_ == "ORB")) ||
        (detector_type_ == "SIFT" && detector_mode_ == "STATIC" && extractor_type_ == "SIFT") ||
        (detector_type_ == "ORB" && detector_mode_ == "STATIC" && extractor_type_ == "ORB")) {
        detectAndDescribe(image_a_, ka, da);
        detectAndDescribe(image_b_, kb, db);
        return;
    }
    throw std::runtime_error("fm3d: DescriptorsMatcher detects and describes on the GPU for DetectorType FAST (STATIC or ADAPTIVE) + "
                             "ExtractorType SIFT, BRISK or ORB, for DetectorType SIFT (STATIC) + ExtractorType SIFT and for DetectorType ORB (STATIC) + ExtractorType ORB only (settings: " + detector_type_ + " / " + detector_mode_ + " / " + extractor_type_ +
                             "); inject the features of other detectors with setFeatures");
}

// feature_detector_->detect + descriptor_extractor_->compute for one frame.  cv::FastFeatureDetector appends
// KeyPoint(x, y, 7.f, -1, score) in row-major order; DescriptorExtractor::compute would drop keypoints outside the
// image or of size 0 before cv::SIFT sees them -- FAST produces neither.
void DescriptorsMatcher::detectAndDescribe(const cv::Mat& image, std::vector<cv::KeyPoint>& kpts, cv::Mat& desc) {
    if (image.empty() || image.type() != CV_8UC1) throw std::runtime_error("fm3d: DescriptorsMatcher needs CV_8UC1 frames");
    fm3d_ctx* ctx = host_ctx();
    const uint8_t* px = image.ptr<uint8_t>(0);
    const int w = image.cols, h = image.rows, stride = (int)image.step1();
    int n = 0;
    if (detector_type_ == "SIFT") {
        // cv::SIFT::detect, then cv::SIFT::compute on its keypoints (every descriptor from the keypoint's own pyramid layer)
        // one detection in the usual case (room for a keypoint per 32 pixels); again with the real count if the frame has more
        int cap = std::max(4096, (w * h) / 32);
        // detector and extractor are both built from FeatureOptions.SiftDetector: one pyramid serves both (cv::SIFT::detectAndCompute)
        std::vector<float> xy, size, angle, resp;
        std::unique_ptr<float[]> rows;          // cap x 128, deliberately not value-initialised: only the rows found are written and read
        std::vector<int32_t> oct;
        kpts.clear();
        desc = cv::Mat();
        for (int attempt = 0; attempt < 2; attempt++) {
            xy.resize((size_t)2 * cap); size.resize(cap); angle.resize(cap); resp.resize(cap); oct.resize(cap);
            rows.reset(new float[(size_t)128 * cap]);
            check(ctx, fm3d_detect_and_describe_sift(ctx, px, w, h, stride, sift_nfeatures_, sift_layers_, sift_contrast_, sift_edge_, sift_sigma_, cap,
                                                     xy.data(), size.data(), angle.data(), resp.data(), oct.data(), &n, rows.get()),
                  "detect + compute (SIFT on one pyramid)");
            if (n <= cap) break;
            cap = n;
        }
        if (n == 0) return;
        kpts.reserve(n);
        for (int i = 0; i < n; i++) {
            cv::KeyPoint kp(xy[2 * i], xy[2 * i + 1], size[i]);
            kp.angle = angle[i]; kp.response = resp[i]; kp.octave = oct[i];
            kpts.push_back(kp);
        }
        desc = cv::Mat::zeros(cv::Size(128, n), CV_32F);
        memcpy(desc.ptr<float>(), rows.get(), sizeof(float) * 128 * (size_t)n);
        return;
    }
    if (detector_type_ == "ORB") {
        // cv::ORB::detect, then cv::ORB::compute on its keypoints: one pyramid, FAST + Harris ranking + intensity-centroid angles per
        // level, rBRIEF rows at the level positions (K15); the keypoint set and the rows of cv::ORB::detectAndCompute
        int cap = std::max(64, 2 * orb_nfeatures_ + 64);
        std::vector<float> xy, size, angle, resp;
        std::vector<int32_t> oct;
        std::vector<uint8_t> rows;
        kpts.clear();
        desc = cv::Mat();
        for (int attempt = 0; attempt < 2; attempt++) {
            xy.resize((size_t)2 * cap); size.resize(cap); angle.resize(cap); resp.resize(cap); oct.resize(cap); rows.resize((size_t)32 * cap);
            check(ctx, fm3d_detect_orb(ctx, px, w, h, stride, orb_nfeatures_, orb_scale_, orb_levels_, 20, cap, xy.data(), size.data(), angle.data(),
                                       resp.data(), oct.data(), rows.data(), &n), "detect + compute (ORB)");
            if (n <= cap) break;
            cap = n;
        }
        if (n == 0) return;
        kpts.reserve(n);
        desc = cv::Mat::zeros(cv::Size(32, n), CV_8U);
        for (int i = 0; i < n; i++) {
            cv::KeyPoint kp(xy[2 * i], xy[2 * i + 1], size[i]);
            kp.angle = angle[i]; kp.response = resp[i]; kp.octave = oct[i];
            kpts.push_back(kp);
            memcpy(desc.ptr<uint8_t>(i), rows.data() + (size_t)32 * i, 32);
        }
        return;
    }
    int threshold = fast_threshold_, nonmax = fast_nonmax_;
    if (detector_mode_ == "ADAPTIVE") {
        // descriptorsmatcher.cpp:186-199: cv::DynamicAdaptedFeatureDetector(AdjusterAdapter::create("FAST"), MinFeatures,
        // MaxFeatures, MaxIters) of OpenCV 2.4 (removed in 3.0, so not in the cv2 of this image: restated from the published
        // 2.4 algorithm, parity unpinned).  cv::FastAdjuster starts at threshold 20 with non-maximum suppression and moves the
        // threshold by one per iteration: down when there are too few keypoints, up when there are too many; the loop ends
        // when the count is inside [MinFeatures, MaxFeatures], after MaxIters detections, when it has gone both ways
        // (oscillation), or when the threshold leaves (1, 200).  The keypoints of the LAST detection are the result.
        threshold = 20; nonmax = 1;
        bool down = false, up = false, good = false;
        int iters = adaptive_iters_, last = threshold;
        while (iters > 0 && !(down && up) && !good && threshold > 1 && threshold < 200) {
            last = threshold;
            check(ctx, fm3d_detect_fast(ctx, px, w, h, stride, threshold, nonmax, 0, nullptr, nullptr, &n), "detect (adaptive)");
            if (n < adaptive_min_) { down = true; threshold--; }
            else if (n > adaptive_max_) { up = true; threshold++; }
            else good = true;
            iters--;
        }
        if (iters == adaptive_iters_) {          // the loop never ran (MaxIters <= 0): cv leaves the keypoint list empty
            kpts.clear();
            desc = cv::Mat();
            adaptive_threshold_used_ = -1;
            return;
        }
        threshold = last;
        adaptive_threshold_used_ = last;
    }
    check(ctx, fm3d_detect_fast(ctx, px, w, h, stride, threshold, nonmax, 0, nullptr, nullptr, &n), "detect (count)");
    std::vector<float> xy((size_t)2 * (n > 0 ? n : 1)), resp((size_t)(n > 0 ? n : 1));
    if (n > 0) check(ctx, fm3d_detect_fast(ctx, px, w, h, stride, threshold, nonmax, n, xy.data(), resp.data(), &n), "detect");
    kpts.clear();
    kpts.reserve(n);
    std::vector<float> k4((size_t)4 * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        cv::KeyPoint kp(xy[2 * i], xy[2 * i + 1], 7.f);
        kp.angle = -1.f;
        kp.response = resp[i];
        kpts.push_back(kp);
        k4[4 * i] = kp.pt.x; k4[4 * i + 1] = kp.pt.y; k4[4 * i + 2] = kp.size; k4[4 * i + 3] = kp.angle;
    }
    if (n == 0) { desc = cv::Mat(); return; }
    if (extractor_type_ == "BRISK") {
        // cv::BRISK::compute erases the keypoints whose pattern would leave the image and returns one CV_8U row of 64
        // bytes per survivor, with KeyPoint::angle set.  FM3D_BRISK_ORIENTATION=0 reproduces OpenCV 2.4, which skips the
        // orientation step for provided keypoints (FAST's angle -1 then means "unrotated").
        const char* e = getenv("FM3D_BRISK_ORIENTATION");
        const int orient = e ? atoi(e) : 1;
        std::vector<uint8_t> rows((size_t)64 * n), kept(n);
        std::vector<float> ang(n);
        check(ctx, fm3d_describe_keypoints_brisk(ctx, px, w, h, stride, k4.data(), n, orient, rows.data(), kept.data(), ang.data()), "compute (BRISK)");
        int m = 0;
        for (int i = 0; i < n; i++) m += kept[i] ? 1 : 0;
        std::vector<cv::KeyPoint> survivors;
        survivors.reserve(m);
        desc = m > 0 ? cv::Mat::zeros(cv::Size(64, m), CV_8U) : cv::Mat();
        for (int i = 0, j = 0; i < n; i++) {
            if (!kept[i]) continue;
            kpts[i].angle = ang[i];
            survivors.push_back(kpts[i]);
            memcpy(desc.ptr<uint8_t>(j++), rows.data() + (size_t)64 * i, 64);
        }
        kpts.swap(survivors);
        return;
    }
    if (extractor_type_ == "ORB") {
        // cv::ORB::compute erases the keypoints within 31 pixels of the border and returns one CV_8U row of 32 bytes per
        // survivor; FAST's angle -1 is used as it is (a rotation by -1 degree).
        std::vector<uint8_t> rows((size_t)32 * n), kept(n);
        check(ctx, fm3d_describe_keypoints_orb(ctx, px, w, h, stride, k4.data(), n, rows.data(), kept.data()), "compute (ORB)");
        int m = 0;
        for (int i = 0; i < n; i++) m += kept[i] ? 1 : 0;
        std::vector<cv::KeyPoint> survivors;
        survivors.reserve(m);
        desc = m > 0 ? cv::Mat::zeros(cv::Size(32, m), CV_8U) : cv::Mat();
        for (int i = 0, j = 0; i < n; i++) {
            if (!kept[i]) continue;
            survivors.push_back(kpts[i]);
            memcpy(desc.ptr<uint8_t>(j++), rows.data() + (size_t)32 * i, 32);
        }
        kpts.swap(survivors);
        return;
    }
    desc = cv::Mat::zeros(cv::Size(128, n), CV_32F);
    if (sift_layers_ != 3 || std::fabs(sift_sigma_ - 1.6) > 1e-12) {
        // the fused base-image kernel of K11 is built for sigma 1.6: other settings take the pyramid builder (octave 0, layer 0)
        std::vector<int32_t> oct(n, 0);
        check(ctx, fm3d_describe_keypoints_sift_oct(ctx, px, w, h, stride, k4.data(), oct.data(), n, sift_layers_, sift_sigma_, desc.ptr<float>()),
              "compute (SIFT, settings of the file)");
        return;
    }
    check(ctx, fm3d_describe_keypoints_sift(ctx, px, w, h, stride, k4.data(), n, desc.ptr<float>()), "compute (SIFT)");
}

void DescriptorsMatcher::knn(const cv::Mat& q, const cv::Mat& t, std::vector<std::vector<cv::DMatch> >& out) {
    fm3d_ctx* ctx = host_ctx();
    const int nq = q.rows, nt = t.rows;
    std::vector<int32_t> idx((size_t)nq * 2, -1);
    std::vector<float> dist((size_t)nq * 2, 0.f);
    if (binary_ && q.depth() != CV_8U) throw std::runtime_error("fm3d: binary extractor needs CV_8U descriptors");
    if (!binary_ && q.depth() != CV_32F) throw std::runtime_error("fm3d: float extractor needs CV_32F descriptors");
    if (host_ctxs().size() > 1 && nq > 0 && nt > 0) {
        knn_sharded(binary_, q, t, idx, dist);
    } else if (binary_) {
        check(ctx, fm3d_match_knn2_hamming(ctx, q.ptr<uint8_t>(), nq, t.ptr<uint8_t>(), nt, q.cols, idx.data(), dist.data()), "knn (hamming)");
    } else {
        check(ctx, fm3d_match_knn2_f32(ctx, q.ptr<float>(), nq, t.ptr<float>(), nt, q.cols, idx.data(), dist.data()), "knn (L2)");
    }
    out.clear();
    out.resize(nq);
    for (int i = 0; i < nq; i++)
        for (int k = 0; k < 2; k++)
            if (idx[2 * i + k] >= 0) out[i].push_back(cv::DMatch(i, idx[2 * i + k], dist[2 * i + k]));
}

void DescriptorsMatcher::crosscompare(std::vector<std::vector<cv::DMatch> >& matchesAB, std::vector<std::vector<cv::DMatch> >& matchesBA,
                                      std::vector<cv::KeyPoint>& kpts_a, std::vector<cv::KeyPoint>& kpts_b,
                                      cv::Mat& da, cv::Mat& db) {
    features(kpts_a, kpts_b, da, db);
    knn(da, db, matchesAB);
    knn(db, da, matchesBA);
}

void DescriptorsMatcher::compare(std::vector<std::vector<cv::DMatch> >& matches, std::vector<cv::KeyPoint>& kpts_a,
                                 std::vector<cv::KeyPoint>& kpts_b, cv::Mat& da, cv::Mat& db) {
    features(kpts_a, kpts_b, da, db);
    knn(da, db, matches_);
    matches = matches_;
}

void DescriptorsMatcher::compareWithNNDR(double epsilon, std::vector<cv::DMatch>& matches, std::vector<cv::KeyPoint>& kpts_a,
                                         std::vector<cv::KeyPoint>& kpts_b, cv::Mat& da, cv::Mat& db) {
    features(kpts_a, kpts_b, da, db);
    fm3d_ctx* ctx = host_ctx();
    const int nq = da.rows, nt = db.rows;
    std::vector<int32_t> qi(nq), ti(nq);
    std::vector<float> d(nq);
    mutual_.assign(nq, 0);
    int n = 0;
    if (host_ctxs().size() > 1 && nq > 0 && nt > 0) {
        // several GPUs: the 2-NN lists of the query shards, then the reference's own loop (descriptorsmatcher.cpp:119-129) on the
        // host -- the comparison in double on the float distances -- and the mutual flag from the role-swapped search
        std::vector<int32_t> ab((size_t)nq * 2, -1), ba((size_t)nt * 2, -1);
        std::vector<float> dab((size_t)nq * 2, 0.f), dba((size_t)nt * 2, 0.f);
        knn_sharded(binary_, da, db, ab, dab);
        knn_sharded(binary_, db, da, ba, dba);
        for (int i = 0; i < nq; i++) {
            if (ab[2 * i] < 0 || ab[2 * i + 1] < 0) continue;
            if ((double)dab[2 * i] <= epsilon * (double)dab[2 * i + 1]) {
                qi[n] = i; ti[n] = ab[2 * i]; d[n] = dab[2 * i];
                mutual_[n] = ba[2 * (size_t)ab[2 * i]] == i ? 1 : 0;
                n++;
            }
        }
    } else if (binary_)
        check(ctx, fm3d_match_nndr_hamming(ctx, da.ptr<uint8_t>(), nq, db.ptr<uint8_t>(), nt, da.cols, epsilon, qi.data(), ti.data(),
                                           d.data(), mutual_.data(), &n), "compareWithNNDR");
    else
        check(ctx, fm3d_match_nndr_f32(ctx, da.ptr<float>(), nq, db.ptr<float>(), nt, da.cols, epsilon, qi.data(), ti.data(), d.data(),
                                       mutual_.data(), &n), "compareWithNNDR");
    mutual_.resize(n);
    for (int i = 0; i < n; i++) matches.push_back(cv::DMatch(qi[i], ti[i], d[i]));  // appended, not cleared (:126)
}

// descriptorsmatcher.cpp:133-174: one keypoint per patch (centre, size = patch edge, angle -1), one
// descriptor row per patch.  ExtractorType SIFT runs on the GPU (K9); the other extractors of the
// reference (SURF / ORB / BRISK / FREAK of OpenCV 2.4) are upstream code this library does not carry.
void DescriptorsMatcher::extractDescriptorsFromPatches(const std::vector<cv::Mat>& patchesVector, cv::Mat& descriptors) {
    if (extractor_type_ != "SIFT" && extractor_type_ != "ORB")
        throw std::runtime_error("fm3d: extractDescriptorsFromPatches runs on the GPU for ExtractorType SIFT and ORB only (settings: " +
                                 extractor_type_ + ")");
    const int n = (int)patchesVector.size();
    if (n == 0) { descriptors = cv::Mat(); return; }
    const int S = patchesVector[0].rows;
    std::vector<uint8_t> packed((size_t)n * S * S);
    for (int k = 0; k < n; k++) {
        const cv::Mat& p = patchesVector[k];
        if (p.rows != S || p.cols != S || p.type() != CV_8UC1) throw std::runtime_error("fm3d: patches must be square CV_8UC1 of one size");
        for (int r = 0; r < S; r++) memcpy(packed.data() + ((size_t)k * S + r) * S, p.ptr<uint8_t>(r), (size_t)S);
    }
    fm3d_ctx* ctx = host_ctx();
    if (extractor_type_ == "ORB") {
        descriptors = cv::Mat::zeros(cv::Size(32, n), CV_8U);
        check(ctx, fm3d_describe_patches_orb(ctx, packed.data(), n, S, descriptors.ptr<uint8_t>()), "extractDescriptorsFromPatches (ORB)");
        return;
    }
    descriptors = cv::Mat::zeros(cv::Size(128, n), CV_32F);
    check(ctx, fm3d_describe_patches_sift(ctx, packed.data(), n, S, descriptors.ptr<float>()), "extractDescriptorsFromPatches");
}

// -------------------------------------------------------------------- SingleCameraTriangulator
SingleCameraTriangulator::SingleCameraTriangulator(cv::FileStorage& settings)
    : ctx_(host_ctx()), patch_eps_(0), patch_cmpp_(0), pyramids_(0), write_patch_files_(getenv("FM3D_NO_PATCH_FILES") == nullptr) {
    std::vector<double> tIC, rIC;
    settings["CameraSettings"]["translationIC"] >> tIC;
    settings["CameraSettings"]["rodriguesIC"] >> rIC;
    if (tIC.size() != 3 || rIC.size() != 3) throw std::runtime_error("fm3d: CameraSettings.translationIC / rodriguesIC missing");
    translation_IC_ = cv::Vec3d(tIC[0], tIC[1], tIC[2]);
    rodrigues_IC_ = cv::Vec3d(rIC[0], rIC[1], rIC[2]);
    cv::Matx33d R;
    rodriguesToMatrix(rodrigues_IC_, R);
    composeTransformation(R, translation_IC_, g_IC_);
    double K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 1}, dist[5];
    K[0] = (double)settings["CameraSettings"]["Fx"]; K[4] = (double)settings["CameraSettings"]["Fy"];
    K[2] = (double)settings["CameraSettings"]["Cx"]; K[5] = (double)settings["CameraSettings"]["Cy"];
    // (k0,k1,p1,p2,k2) of settings.yml -> OpenCV (k1,k2,p1,p2,k3) (singlecameratriangulator.cpp:101-105)
    dist[0] = (double)settings["CameraSettings"]["k0"]; dist[1] = (double)settings["CameraSettings"]["k1"];
    dist[2] = (double)settings["CameraSettings"]["p1"]; dist[3] = (double)settings["CameraSettings"]["p2"];
    dist[4] = (double)settings["CameraSettings"]["k2"];
    settings["CameraSettings"]["zThresholdMin"] >> z_threshold_min_;
    settings["CameraSettings"]["zThresholdMax"] >> z_threshold_max_;
    settings["Neighborhoods"]["pixelsRay"] >> pixels_ray_;
    settings["Neighborhoods"]["pyramids"] >> pyramids_;
    settings["Neighborhoods"]["epsilon"] >> patch_eps_;
    settings["Neighborhoods"]["cmPerPixel"] >> patch_cmpp_;
    const double zmin = z_threshold_min_, zmax = z_threshold_max_;
    for_each_ctx("set_camera", [&](int k) { return fm3d_set_camera(host_ctxs()[k], K, dist, zmin, zmax); });
}

void SingleCameraTriangulator::setImages(const cv::Mat& img1, const cv::Mat& img2) {
    // The reference keeps each image at its own size and step (:116-120); the device pyramids need two CV_8UC1 frames of
    // one size.  Each image is uploaded with ITS OWN row pitch (a ROI or a padded frame has step1() != cols).
    if (img1.empty() || img2.empty() || img1.type() != CV_8UC1 || img2.type() != CV_8UC1)
        throw std::runtime_error("fm3d: setImages needs two non-empty CV_8UC1 images");
    if (img1.cols != img2.cols || img1.rows != img2.rows)
        throw std::runtime_error("fm3d: setImages needs two images of one size (" + std::to_string(img1.cols) + "x" + std::to_string(img1.rows) +
                                 " / " + std::to_string(img2.cols) + "x" + std::to_string(img2.rows) + ")");
    img_1_ = img1; img_2_ = img2;  // shallow, like new cv::Mat(img) in the reference (:118-119)
    const int levels = pyramids_;
    const int s1 = (int)img1.step1(), s2 = (int)img2.step1();   // CV_8UC1: elements = bytes; step1() exists in cv::Mat and in the stand-in
    if (host_ctxs().size() > 1 && host_comm_ok()) {
        // several GPUs: both frames go to the first GPU once and reach the others with one ncclBroadcast each (NVLink);
        // every GPU then builds its own pyramids (K4 is cheaper than shipping them: 1.33 x the frame)
        const int w = img1.cols, h = img1.rows;
        std::vector<unsigned char> t1, t2;
        const size_t bytes = (size_t)w * h;
        std::vector<void*> d1 = replicate(SLOT_IMG1, packed_rows(img1, t1), bytes);
        std::vector<void*> d2 = replicate(SLOT_IMG2, packed_rows(img2, t2), bytes);
        for_each_ctx("set_images", [&](int k) {
            if (int rc = fm3d_set_images_dev(host_ctxs()[k], (const uint8_t*)d1[k], (const uint8_t*)d2[k], w, h, w, levels)) return rc;
            return fm3d_sync(host_ctxs()[k]);
        });
        return;
    }
    for_each_ctx("set_images", [&](int k) {
        return fm3d_set_images2(host_ctxs()[k], img1.data, s1, img2.data, s2, img1.cols, img1.rows, levels);
    });
}

void SingleCameraTriangulator::setg12(const cv::Vec3d& T1, const cv::Vec3d& T2, const cv::Vec3d& rod1, const cv::Vec3d& rod2,
                                      cv::Matx44d& g12) {
    double out[16];
    if (fm3d_compose_g12(T1.val, T2.val, rod1.val, rod2.val, rodrigues_IC_.val, translation_IC_.val, out) != FM3D_OK)
        throw std::runtime_error("fm3d: setg12 failed");
    for (int i = 0; i < 16; i++) g_12_.val[i] = out[i];
    g12 = g_12_;
    for_each_ctx("set_g12", [&](int k) { return fm3d_set_g12(host_ctxs()[k], out); });
}

void SingleCameraTriangulator::setKeypoints(const std::vector<cv::KeyPoint>& kpts1, const std::vector<cv::KeyPoint>& kpts2,
                                            const std::vector<cv::DMatch>& matches) {
    kp1_.resize(kpts1.size() * 2); kp2_.resize(kpts2.size() * 2);
    for (size_t i = 0; i < kpts1.size(); i++) { kp1_[2 * i] = kpts1[i].pt.x; kp1_[2 * i + 1] = kpts1[i].pt.y; }
    for (size_t i = 0; i < kpts2.size(); i++) { kp2_[2 * i] = kpts2[i].pt.x; kp2_[2 * i + 1] = kpts2[i].pt.y; }
    qidx_.resize(matches.size()); tidx_.resize(matches.size());
    for (size_t i = 0; i < matches.size(); i++) { qidx_[i] = matches[i].queryIdx; tidx_[i] = matches[i].trainIdx; }
}

void SingleCameraTriangulator::triangulate(std::vector<cv::Vec3d>& triangulatedPoints, std::vector<bool>& outliersMask) {
    const int n = (int)qidx_.size();
    std::vector<double> xyz_all((size_t)n * 3), xyz((size_t)n * 3);
    std::vector<uint8_t> mask(n);
    int ninl = 0;
    check(ctx_, fm3d_triangulate(ctx_, kp1_.data(), (int)kp1_.size() / 2, kp2_.data(), (int)kp2_.size() / 2, qidx_.data(), tidx_.data(), n,
                                 xyz_all.data(), mask.data(), xyz.data(), nullptr, &ninl), "triangulate");
    triangulatedPoints.clear();                      // cleared (:189)
    for (int i = 0; i < n; i++) {                    // appended, never cleared (:202-209)
        outliersMask.push_back(mask[i] != 0);
        outliers_mask_.push_back(mask[i] != 0);
    }
    for (int i = 0; i < ninl; i++) triangulatedPoints.push_back(cv::Vec3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
}

static void unpack_patches(int n, int S, const std::vector<uint8_t>& p, const std::vector<double>& ip,
                           std::vector<cv::Mat>& patchesVector, std::vector<cv::Mat>& imagePointsVector, bool write_files) {
    patchesVector.clear();
    imagePointsVector.clear();
    for (int f = 0; f < n; f++) {
        cv::Mat patch(cv::Size(S, S), CV_8UC1);
        memcpy(patch.data, p.data() + (size_t)f * S * S, (size_t)S * S);
        cv::Mat pts(cv::Size(1, S * S), CV_64FC2);
        memcpy(pts.data, ip.data() + (size_t)f * S * S * 2, sizeof(double) * 2 * S * S);
        patchesVector.push_back(patch);
        imagePointsVector.push_back(pts);
    }
    // FM3D_PATCH_ATLAS=<file.pgm>: every patch into ONE binary PGM (a grid of ceil(sqrt(n)) patches per row, patch f at row
    // f / cols, column f % cols) instead of n files -- the reference's 20 000 imwrite calls are its I/O bottleneck at C3 size
    if (const char* atlas = getenv("FM3D_PATCH_ATLAS")) {
        if (n > 0) {
            const int cols = (int)std::ceil(std::sqrt((double)n)), rows = (n + cols - 1) / cols;
            cv::Mat sheet = cv::Mat::zeros(cv::Size(cols * S, rows * S), CV_8UC1);
            for (int f = 0; f < n; f++)
                for (int r = 0; r < S; r++)
                    memcpy(sheet.ptr<uint8_t>((f / cols) * S + r) + (size_t)(f % cols) * S, p.data() + ((size_t)f * S + r) * S, (size_t)S);
            cv::imwrite(atlas, sheet);
        }
        return;
    }
    if (write_files)
        for (int f = 0; f < n; f++) cv::imwrite("patch_" + NumberToString<int>(f) + ".pgm", patchesVector[f]);  // (:799-802)
}

void SingleCameraTriangulator::projectReferencePointsToImageWithFrames(const std::vector<cv::Vec3d>& ref,
                                                                       const std::vector<cv::Matx44d>& featureFrames,
                                                                       std::vector<cv::Mat>& patchesVector,
                                                                       std::vector<cv::Mat>& imagePointsVector) {
    const int S = (int)std::sqrt((double)ref.size());
    // the GPU generates the S x S grid analytically; make sure the caller's reference neighbourhood is that grid
    const int Sg = fm3d_patch_size(patch_eps_, patch_cmpp_);
    const double inc = patch_cmpp_ * 0.01;
    bool same = S == Sg && S > 0;
    for (int k = 0; same && k < 3; k++) {
        const int idx = k == 0 ? 0 : (k == 1 ? S * S - 1 : S + 1), i = idx / S, j = idx % S;
        same = ref[idx][0] == -patch_eps_ + inc * i && ref[idx][1] == -patch_eps_ + inc * j && ref[idx][2] == 0;
    }
    if (!same) throw std::runtime_error("fm3d: reference neighbourhood is not the square grid of settings.yml (epsilon, cmPerPixel)");
    const int n = (int)featureFrames.size();
    std::vector<double> frames((size_t)n * 16);
    for (int f = 0; f < n; f++) memcpy(&frames[(size_t)f * 16], featureFrames[f].val, sizeof(double) * 16);
    std::vector<uint8_t> p((size_t)n * S * S);
    std::vector<double> ip((size_t)n * S * S * 2);
    check(ctx_, fm3d_extract_patches(ctx_, frames.data(), n, patch_eps_, patch_cmpp_, p.data(), ip.data()), "extract_patches");
    unpack_patches(n, S, p, ip, patchesVector, imagePointsVector, write_patch_files_);
}

void SingleCameraTriangulator::projectPointsToImage(const IMAGE_ID id, const std::vector<std::vector<cv::Vec3d> >& groups,
                                                    std::vector<cv::Mat>& patchesVector, std::vector<cv::Mat>& imagePointsVector) {
    const int n = (int)groups.size();
    if (n == 0) { patchesVector.clear(); imagePointsVector.clear(); return; }
    const int S = (int)std::sqrt((double)groups[0].size());
    std::vector<double> g((size_t)n * S * S * 3);
    for (int f = 0; f < n; f++)
        for (int k = 0; k < S * S; k++)
            for (int c = 0; c < 3; c++) g[((size_t)f * S * S + k) * 3 + c] = groups[f][k][c];
    std::vector<uint8_t> p((size_t)n * S * S);
    std::vector<double> ip((size_t)n * S * S * 2);
    check(ctx_, fm3d_project_groups(ctx_, id == image1 ? 1 : 2, g.data(), n, S, p.data(), ip.data()), "project_groups");
    unpack_patches(n, S, p, ip, patchesVector, imagePointsVector, write_patch_files_);
}

// ---- the per-evaluation helpers (singlecameratriangulator.cpp:279-339, 341-397, 472-665)
void SingleCameraTriangulator::extractPixelsContour(const cv::Vec3d& point, std::vector<Pixel>& pixels) {
    const int cap = (2 * pixels_ray_ + 1) * (2 * pixels_ray_ + 1);
    std::vector<double> xy((size_t)cap * 2);
    int m = 0;
    check(ctx_, fm3d_disc_pixels(ctx_, point.val, pixels_ray_, xy.data(), cap, &m), "disc_pixels");
    pixels.clear();                                               // (:396)
    pixels.reserve(m);
    for (int i = 0; i < m; i++) { Pixel p = {xy[2 * i], xy[2 * i + 1], 0}; pixels.push_back(p); }
}

int SingleCameraTriangulator::get3dPointsFromImage1Pixels(const cv::Vec3d& point, const cv::Vec3d& normal, const cv::Mat& pixelMat,
                                                          std::vector<cv::Vec3d>& pointsGroup) {
    const int m = pixelMat.rows;                                  // m x 1 CV_64FC2
    if (m == 0) return 0;
    std::vector<double> xyz((size_t)m * 3);
    int info = 0;
    check(ctx_, fm3d_plane_points(ctx_, point.val, normal.val, (const double*)pixelMat.data, m, xyz.data(), &info), "plane_points");
    if (info == -6) throw std::runtime_error("fm3d: NaN plane point (the reference exits with -6)");
    const int cmax = (int)(2 * z_threshold_max_);                 // isInBoundingBox (:646-655)
    for (int i = 0; i < m; i++) {
        const double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
        if (!((X > -cmax && X < cmax) && (Y > -cmax && Y < cmax) && (Z > 0 && Z < cmax))) return -1;
        pointsGroup.push_back(cv::Vec3d(X, Y, Z));
    }
    return 0;
}

void SingleCameraTriangulator::extractPixelsContourAndGet3DPoints(const cv::Vec3d& point, const cv::Vec3d& normal,
                                                                  std::vector<Pixel>& pixels, std::vector<cv::Vec3d>& pointsGroup) {
    extractPixelsContour(point, pixels);
    const int m = (int)pixels.size();
    if (m == 0) return;
    std::vector<double> xy((size_t)m * 2), xyz((size_t)m * 3);
    for (int i = 0; i < m; i++) { xy[2 * i] = pixels[i].x_; xy[2 * i + 1] = pixels[i].y_; }
    int info = 0;
    check(ctx_, fm3d_plane_points(ctx_, point.val, normal.val, xy.data(), m, xyz.data(), &info), "plane_points");
    // no bounding-box gate here (:509-518); the debug drawing into test1.pgm (:520-525) is not reproduced
    for (int i = 0; i < m; i++) pointsGroup.push_back(cv::Vec3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
}

static bool pixel_good(double x, double y, double scale, int cols, int rows) {   // isPixelGood (:657-665)
    return !((x < 0) || (x > ((1 / scale) * cols)) || (y < 0) || (y > ((1 / scale) * rows)));
}

int SingleCameraTriangulator::updateImage1PixelsIntensity(const double scale, std::vector<Pixel>& pixels) {
    const int m = (int)pixels.size();
    if (m == 0) return 0;
    std::vector<double> xy((size_t)m * 2);
    std::vector<float> inten(m);
    for (int i = 0; i < m; i++) { xy[2 * i] = pixels[i].x_; xy[2 * i + 1] = pixels[i].y_; }
    int info = 0;
    check(ctx_, fm3d_sample_pixels(ctx_, 1, 0, scale, 1, xy.data(), m, inten.data(), &info), "sample_pixels");
    for (int i = 0; i < m; i++) {
        if (!pixel_good(pixels[i].x_, pixels[i].y_, scale, img_1_.cols, img_1_.rows)) return -1;
        pixels[i].i_ = inten[i];
    }
    return 0;
}

int SingleCameraTriangulator::projectPointsToImage2(const std::vector<cv::Vec3d>& pointsGroup, const double scale,
                                                    std::vector<Pixel>& pixels) {
    const int m = (int)pointsGroup.size();
    if (m == 0) return 0;
    std::vector<double> xyz((size_t)m * 3), xy2((size_t)m * 2);
    std::vector<float> inten(m);
    for (int i = 0; i < m; i++) for (int c = 0; c < 3; c++) xyz[3 * i + c] = pointsGroup[i][c];
    int info = 0;
    check(ctx_, fm3d_project_to_image2(ctx_, xyz.data(), m, 0, scale, xy2.data(), inten.data(), &info), "project_to_image2");
    for (int i = 0; i < m; i++) {
        if (!pixel_good(xy2[2 * i], xy2[2 * i + 1], scale, img_1_.cols, img_1_.rows)) return -1;
        Pixel p = {xy2[2 * i], xy2[2 * i + 1], inten[i]};
        pixels.push_back(p);
    }
    return 0;
}

void SingleCameraTriangulator::projectPointsAndComputeResidual(const cv::Mat& pointsGroup, cv::Mat& imagePoints1, cv::Mat& imagePoints2,
                                                               std::vector<double>& residualsVector) {
    // single-group variant (:322-339): residual of the ROUNDED pixels, uchar arithmetic promoted to int
    const int n = pointsGroup.cols;
    if (n == 0) return;
    std::vector<double> xyz((size_t)n * 3);
    for (int k = 0; k < n; k++) for (int c = 0; c < 3; c++) xyz[3 * k + c] = pointsGroup.ptr<double>(c)[k];
    imagePoints1 = cv::Mat(n, 1, CV_64FC2);
    imagePoints2 = cv::Mat(n, 1, CV_64FC2);
    std::vector<uint8_t> dummy((size_t)n);
    // image 1: projectPoints with r = t = 0 (n groups of one point); image 2: with g12 (projectPointsToImages, :260-276)
    check(ctx_, fm3d_project_groups(ctx_, 1, xyz.data(), n, 1, dummy.data(), (double*)imagePoints1.data), "project_groups");
    int info = 0;
    check(ctx_, fm3d_project_to_image2(ctx_, xyz.data(), n, 0, 1.0, (double*)imagePoints2.data, nullptr, &info), "project_to_image2");
    // img.at<uchar>(round(y), round(x)) of both images: the bilinear sampler at integer coordinates
    // returns exactly that pixel, so the lookup runs on the GPU like everything else
    std::vector<double> r1((size_t)n * 2), r2((size_t)n * 2);
    for (int k = 0; k < n; k++) {
        r1[2 * k] = std::round(imagePoints1.ptr<double>(k)[0]); r1[2 * k + 1] = std::round(imagePoints1.ptr<double>(k)[1]);
        r2[2 * k] = std::round(imagePoints2.ptr<double>(k)[0]); r2[2 * k + 1] = std::round(imagePoints2.ptr<double>(k)[1]);
    }
    std::vector<float> i1(n), i2(n);
    check(ctx_, fm3d_sample_pixels(ctx_, 1, 0, 1.0, 0, r1.data(), n, i1.data(), &info), "sample_pixels");
    check(ctx_, fm3d_sample_pixels(ctx_, 2, 0, 1.0, 0, r2.data(), n, i2.data(), &info), "sample_pixels");
    for (int k = 0; k < n; k++) residualsVector.push_back((double)((int)i1[k] - (int)i2[k]));   // uchar - uchar, promoted to int
}

void SingleCameraTriangulator::projectPointsAndComputeResidual(const std::vector<cv::Mat>& pointsGroupVector,
                                                               std::vector<cv::Mat>& imagePointsVector1,
                                                               std::vector<cv::Mat>& imagePointsVector2,
                                                               std::vector<std::vector<double> >& residualsVectors) {
    // vector variant (:279-320): bilinear samples of both images at the projected points
    imagePointsVector1.clear();
    imagePointsVector2.clear();
    for (size_t g = 0; g < pointsGroupVector.size(); g++) {
        const cv::Mat& pg = pointsGroupVector[g];
        const int n = pg.cols;
        std::vector<double> xyz((size_t)n * 3);
        for (int k = 0; k < n; k++) for (int c = 0; c < 3; c++) xyz[3 * k + c] = pg.ptr<double>(c)[k];
        cv::Mat ip1(n, 1, CV_64FC2), ip2(n, 1, CV_64FC2);
        std::vector<float> i1(n), i2(n);
        int info = 0;
        if (n > 0) {
            std::vector<uint8_t> dummy((size_t)n);
            check(ctx_, fm3d_project_groups(ctx_, 1, xyz.data(), n, 1, dummy.data(), (double*)ip1.data), "project_groups");
            check(ctx_, fm3d_project_to_image2(ctx_, xyz.data(), n, 0, 1.0, (double*)ip2.data, nullptr, &info), "project_to_image2");
            check(ctx_, fm3d_sample_pixels(ctx_, 1, 0, 1.0, 0, (const double*)ip1.data, n, i1.data(), &info), "sample_pixels");
            check(ctx_, fm3d_sample_pixels(ctx_, 2, 0, 1.0, 0, (const double*)ip2.data, n, i2.data(), &info), "sample_pixels");
        }
        imagePointsVector1.push_back(ip1);
        imagePointsVector2.push_back(ip2);
        std::vector<double> res(n);
        for (int k = 0; k < n; k++) res[k] = (double)i1[k] - (double)i2[k];       // double pixel1 - double pixel2 (:305-309)
        residualsVectors.push_back(res);
    }
}

// ------------------------------------------------------------------------------ NormalOptimizer
NormalOptimizer::NormalOptimizer(const cv::FileStorage settings, SingleCameraTriangulator* sct)
    : sct_(sct), penalty_mode_(FM3D_PENALTY_FABS) {
    // abs() semantics of the penalty wall (normaloptimizer.cpp:126-142): fabs, what today's g++ makes of the
    // source, unless the environment says otherwise -- the unchanged main.cpp has no other way to choose
    if (const char* e = getenv("FM3D_PENALTY")) penalty_mode_ = atoi(e);
    // what is minimised: the reference's SSD (0, default) or the NCC extension (1); see fm3d_cost_mode in fm3d.h
    cost_mode_ = FM3D_COST_SSD;
    if (const char* e = getenv("FM3D_COST")) cost_mode_ = atoi(e);
    settings["Neighborhoods"]["pyramids"] >> pyr_levels_;
    settings["Neighborhoods"]["epsilonLMMIN"] >> epsilon_lmmin_;
    std::vector<double> rIC;
    settings["CameraSettings"]["rodriguesIC"] >> rIC;
    cv::Matx33d R;
    rodriguesToMatrix(cv::Vec3d(rIC[0], rIC[1], rIC[2]), R);
    // gravity = R_IC^-1 (0,0,-1) = -third row of R_IC (normaloptimizer.cpp:171-176)
    gravity_ = cv::Vec3d(-R(2, 0), -R(2, 1), -R(2, 2));
}

cv::Vec3d NormalOptimizer::getGravity() { return gravity_; }

void NormalOptimizer::setImages(const cv::Mat& img1, const cv::Mat& img2) {
    if (sct_ == 0) throw std::runtime_error("fm3d: NormalOptimizer without a SingleCameraTriangulator");  // exit(-1) (:194-197)
    sct_->pyramids_ = pyr_levels_;
    sct_->setImages(img1, img2);  // uploads and builds the pyramids (compute_pyramids, :206-221)
}

void NormalOptimizer::computeOptimizedNormals(std::vector<cv::Vec3d>& points3D, std::vector<cv::Vec3d>& normalsVector) {
    std::vector<cv::Scalar> colors(points3D.size(), cv::Scalar(150, 150, 255));
    computeOptimizedNormals(points3D, normalsVector, colors);
}

void NormalOptimizer::computeOptimizedNormals(std::vector<cv::Vec3d>& points3D, std::vector<cv::Vec3d>& normalsVector,
                                              std::vector<cv::Scalar>&) {
    fm3d_ctx* ctx = sct_->context();
    const int n = (int)points3D.size(), L1 = pyr_levels_ + 1;
    std::vector<double> xyz((size_t)n * 3), normals((size_t)n * 3), cost(n);
    std::vector<int32_t> status(n), nfev((size_t)n * L1), npen(n);
    for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) xyz[3 * i + c] = points3D[i][c];
    // features are independent: contiguous shards, one per GPU context, written into disjoint ranges
    (void)ctx;
    const int G = (int)host_ctxs().size(), r = sct_->pixelsRay(), pm = penalty_mode_;
    const double eps = epsilon_lmmin_;
    const int cm = cost_mode_;
    for_each_ctx("optimize_normals", [&](int k) {
        const int lo = (int)((long long)n * k / G), hi = (int)((long long)n * (k + 1) / G);
        if (hi <= lo) return (int)FM3D_OK;
        if (int rc = fm3d_set_option(host_ctxs()[k], "normals_cost", (double)cm)) return rc;
        return fm3d_optimize_normals(host_ctxs()[k], xyz.data() + 3 * (size_t)lo, hi - lo, r, eps, pm, normals.data() + 3 * (size_t)lo,
                                     status.data() + lo, nfev.data() + (size_t)lo * L1, npen.data() + lo, cost.data() + lo);
    });
    status_.assign(status.begin(), status.end());
    nfev_.assign(n, 0);
    for (int i = 0; i < n; i++) for (int l = 0; l < L1; l++) nfev_[i] += nfev[(size_t)i * L1 + l];
    // failed features are erased from points3D in place; normals are appended (:364-382,447)
    std::vector<cv::Vec3d> kept;
    for (int i = 0; i < n; i++) {
        if (status[i] != FM3D_FEAT_OK) continue;
        kept.push_back(points3D[i]);
        normalsVector.push_back(cv::Vec3d(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]));
    }
    points3D.swap(kept);
}

void NormalOptimizer::computeFeaturesFrames(std::vector<cv::Vec3d>& points3D, std::vector<cv::Vec3d>& normalsVector,
                                            std::vector<cv::Matx44d>& featuresFrames) {
    fm3d_ctx* ctx = sct_->context();
    const int n = (int)std::min(points3D.size(), normalsVector.size());
    std::vector<double> xyz((size_t)n * 3), nrm((size_t)n * 3), fr((size_t)n * 16);
    for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) { xyz[3 * i + c] = points3D[i][c]; nrm[3 * i + c] = normalsVector[i][c]; }
    check(ctx, fm3d_feature_frames(ctx, xyz.data(), nrm.data(), n, gravity_.val, fr.data()), "feature_frames");
    for (int i = 0; i < n; i++) {
        cv::Matx44d F;
        memcpy(F.val, &fr[(size_t)i * 16], sizeof(double) * 16);
        featuresFrames.push_back(F);  // appended (:490)
    }
}

// ----------------------------------------------------------------------- NeighborhoodsGenerator
NeighborhoodsGenerator::NeighborhoodsGenerator(cv::FileStorage settings)
    : epsilon_(0), cm_per_pixel_(0), number_of_angles_(0), number_of_rays_(0), circular_(false) {
    std::string method = (std::string)settings["Neighborhoods"]["method"];
    if (method == "square") {
        settings["Neighborhoods"]["epsilon"] >> epsilon_;
        settings["Neighborhoods"]["cmPerPixel"] >> cm_per_pixel_;
    } else if (method == "circular") {                                    // (:46-66)
        circular_ = true;
        settings["Neighborhoods"]["epsilon"] >> epsilon_;
        settings["Neighborhoods"]["thetas"] >> number_of_angles_;
        settings["Neighborhoods"]["rays"] >> number_of_rays_;
        if (number_of_angles_ < 1 || number_of_rays_ < 1) throw std::runtime_error("fm3d: Neighborhoods.thetas / rays missing");
    } else {
        throw std::runtime_error("fm3d: unsupported method for plane neighborhood extraction: '" + method + "'");  // exit(-10) (:69-73)
    }
}

void NeighborhoodsGenerator::computeCircularNeighborhoodsByNormals(const cv::Mat& points, cv::Mat& normals,
                                                                   std::vector<cv::Mat>& neighborhoodsVector) {
    if (!circular_) throw std::runtime_error("fm3d: NeighborhoodsGenerator was not configured with method: circular");
    fm3d_ctx* ctx = host_ctx();
    const int n = points.cols, S = number_of_angles_ * number_of_rays_;
    if (n == 0) return;
    if (normals.empty()) normals = cv::Mat::zeros(cv::Size(n, 3), CV_64FC1);    // zeros = "use P/|P|" (:166-181)
    std::vector<double> p((size_t)n * 3), nn((size_t)n * 3), out((size_t)n * S * 3);
    for (int f = 0; f < n; f++)
        for (int c = 0; c < 3; c++) { p[3 * f + c] = points.ptr<double>(c)[f]; nn[3 * f + c] = normals.ptr<double>(c)[f]; }
    check(ctx, fm3d_circular_neighborhoods(ctx, p.data(), nn.data(), n, epsilon_, number_of_angles_, number_of_rays_, out.data()),
          "circular_neighborhoods");
    for (int f = 0; f < n; f++) {
        for (int c = 0; c < 3; c++) normals.ptr<double>(c)[f] = nn[3 * f + c];
        cv::Mat nb = cv::Mat::zeros(cv::Size(S, 1), CV_64FC3);
        memcpy(nb.data, &out[(size_t)f * S * 3], sizeof(double) * 3 * (size_t)S);
        neighborhoodsVector.push_back(nb);                                // appended, as the reference (:215)
    }
}

void NeighborhoodsGenerator::computeCircularNeighborhoodByNormal(const cv::Vec3d& point, cv::Vec3d& normal, cv::Mat& neighborhood) {
    if (!circular_) throw std::runtime_error("fm3d: NeighborhoodsGenerator was not configured with method: circular");
    fm3d_ctx* ctx = host_ctx();
    const int S = number_of_angles_ * number_of_rays_;
    neighborhood = cv::Mat::zeros(cv::Size(S, 1), CV_64FC3);
    check(ctx, fm3d_circular_neighborhoods(ctx, point.val, normal.val, 1, epsilon_, number_of_angles_, number_of_rays_,
                                           (double*)neighborhood.data), "circular_neighborhoods");
}

void NeighborhoodsGenerator::getReferenceSquaredNeighborhood(std::vector<cv::Vec3d>& neighborhood) {
    const int S = fm3d_patch_size(epsilon_, cm_per_pixel_);
    const double inc = cm_per_pixel_ * 0.01;
    neighborhood.clear();
    for (int i = 0; i < S; i++)
        for (int j = 0; j < S; j++) neighborhood.push_back(cv::Vec3d(-epsilon_ + inc * i, -epsilon_ + inc * j, 0));
}

void NeighborhoodsGenerator::computeSquareNeighborhoodsByNormals(const std::vector<cv::Matx44d>& featuresFrames,
                                                                 std::vector<std::vector<cv::Vec3d> >& neighborhoodsVector) {
    fm3d_ctx* ctx = host_ctx();
    const int n = (int)featuresFrames.size(), S = fm3d_patch_size(epsilon_, cm_per_pixel_);
    neighborhoodsVector.clear();
    if (n == 0) return;
    std::vector<double> frames((size_t)n * 16), out((size_t)n * S * S * 3);
    for (int f = 0; f < n; f++) memcpy(&frames[(size_t)f * 16], featuresFrames[f].val, sizeof(double) * 16);
    check(ctx, fm3d_square_neighborhoods(ctx, frames.data(), n, epsilon_, cm_per_pixel_, out.data()), "square_neighborhoods");
    neighborhoodsVector.resize(n);
    for (int f = 0; f < n; f++) {
        neighborhoodsVector[f].resize((size_t)S * S);
        for (int k = 0; k < S * S; k++) {
            const double* p = &out[((size_t)f * S * S + k) * 3];
            neighborhoodsVector[f][k] = cv::Vec3d(p[0], p[1], p[2]);
        }
    }
}

void NeighborhoodsGenerator::computeSquareNeighborhoodByNormal(const cv::Matx44d& featureFrame, std::vector<cv::Vec3d>& neighborhood) {
    std::vector<cv::Matx44d> one(1, featureFrame);
    std::vector<std::vector<cv::Vec3d> > out;
    computeSquareNeighborhoodsByNormals(one, out);
    neighborhood = out.empty() ? std::vector<cv::Vec3d>() : out[0];
}
