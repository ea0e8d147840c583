// DescriptorsMatcher -- same public interface as the reference class
// (DescriptorsMatcher/descriptorsmatcher.h:44-79); the matching runs on the GPU through libfm3d.
//
// Keypoint detection and description (descriptorsmatcher.cpp:110-115,176-359: SURF / SIFT / ORB
// factories of OpenCV 2.4 nonfree) are upstream of the hot path and out of scope: features are
// injected with setFeatures() (or taken from the output arguments if the caller pre-filled
// them); the three compare* methods then behave as in the reference, with the exact
// brute-force search replacing FLANN.  extractDescriptorsFromPatches (:133-174) runs on the GPU
// for ExtractorType SIFT (fm3d_describe_patches_sift).
#ifndef FM3D_HOST_DESCRIPTORSMATCHER_H_
#define FM3D_HOST_DESCRIPTORSMATCHER_H_
#include <string>
#include <vector>
#include "../../fm3d_cv.h"

class DescriptorsMatcher {
public:
    DescriptorsMatcher(cv::FileStorage& fs, cv::Mat& frame_a, cv::Mat& frame_b);
    ~DescriptorsMatcher();

    void crosscompare(std::vector<std::vector<cv::DMatch> >& matchesAB, std::vector<std::vector<cv::DMatch> >& matchesBA,
                      std::vector<cv::KeyPoint>& kpts_a, std::vector<cv::KeyPoint>& kpts_b,
                      cv::Mat& completeDescriptors_a, cv::Mat& completeDescriptors_b);
    void compare(std::vector<std::vector<cv::DMatch> >& matches, std::vector<cv::KeyPoint>& kpts_a,
                 std::vector<cv::KeyPoint>& kpts_b, cv::Mat& completeDescriptors_a, cv::Mat& completeDescriptors_b);
    void compareWithNNDR(double epsilon, std::vector<cv::DMatch>& matches, std::vector<cv::KeyPoint>& kpts_a,
                         std::vector<cv::KeyPoint>& kpts_b, cv::Mat& completeDescriptors_a, cv::Mat& completeDescriptors_b);
    void extractDescriptorsFromPatches(const std::vector<cv::Mat>& patchesVector, cv::Mat& descriptors);

    // fm3d extension: the upstream detector/extractor output for both frames
    void setFeatures(const std::vector<cv::KeyPoint>& kpts_a, const cv::Mat& desc_a,
                     const std::vector<cv::KeyPoint>& kpts_b, const cv::Mat& desc_b);
    // fm3d extension: mutual-best flags of the last compareWithNNDR (the fused cross-check)
    const std::vector<unsigned char>& mutualFlags() const { return mutual_; }
    // fm3d extension: the FAST threshold the ADAPTIVE mode ended on for the frame described last (-1: STATIC mode)
    int adaptiveThreshold() const { return adaptive_threshold_used_; }

private:
    void features(std::vector<cv::KeyPoint>& ka, std::vector<cv::KeyPoint>& kb, cv::Mat& da, cv::Mat& db);
    void knn(const cv::Mat& q, const cv::Mat& t, std::vector<std::vector<cv::DMatch> >& out);
    cv::Mat image_a_, image_b_, desc_a_, desc_b_;
    std::vector<cv::KeyPoint> kpts_a_, kpts_b_;
    std::vector<std::vector<cv::DMatch> > matches_;
    std::vector<unsigned char> mutual_;
    void detectAndDescribe(const cv::Mat& image, std::vector<cv::KeyPoint>& kpts, cv::Mat& desc);
    bool binary_, have_features_;
    std::string extractor_type_, detector_type_, detector_mode_;
    int fast_threshold_, fast_nonmax_;
    int adaptive_min_, adaptive_max_, adaptive_iters_;
    int adaptive_threshold_used_ = -1;
    // FeatureOptions.SiftDetector (descriptorsmatcher.cpp:246-251, :306-311); cv::SIFT's defaults where the file is silent
    int sift_nfeatures_ = 0, sift_layers_ = 3;
    double sift_contrast_ = 0.04, sift_edge_ = 10.0, sift_sigma_ = 1.6;
    // FeatureOptions.OrbDetector (descriptorsmatcher.cpp:273-281, :336-342); cv::ORB's defaults where the file is silent
    int orb_nfeatures_ = 500, orb_levels_ = 8;
    double orb_scale_ = 1.2;
};
#endif
