// Test driver: the call sequence of the reference's main.cpp (main.cpp:91-187) on top of the
// fm3d class adapters.  Features come from a file (any upstream detector / extractor), or -- with "-" in place
// of the file and DetectorType FAST + ExtractorType SIFT in the settings -- from the frames themselves, as in
// main.cpp.  Usage: pipeline_main -s settings.yml <features.bin | -> result.bin [circular-settings.yml]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <vector>

#include "DescriptorsMatcher/descriptorsmatcher.h"
#include "Triangulator/neighborhoodsgenerator.h"
#include "Triangulator/normaloptimizer.h"
#include "Triangulator/singlecameratriangulator.h"

static void rd(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char** argv) {
    if (argc < 5 || strcmp(argv[1], "-s") != 0) { std::cerr << "usage: -s <settings.yml> <features.bin> <result.bin> [circular-settings.yml]\n"; return 1; }
    cv::FileStorage fs;
    fs.open(argv[2], cv::FileStorage::READ);
    if (!fs.isOpened()) { std::cerr << "Could not open settings file: " << argv[2] << std::endl; return 1; }
    std::string IMG_1, IMG_2;
    fs["IMAGES"]["img1"] >> IMG_1;
    fs["IMAGES"]["img2"] >> IMG_2;
    cv::Mat img1 = cv::imread(IMG_1, CV_LOAD_IMAGE_GRAYSCALE), img2 = cv::imread(IMG_2, CV_LOAD_IMAGE_GRAYSCALE);
    if (img1.empty() || img2.empty()) { std::cerr << "could not read images\n"; return 1; }

    const bool detect_here = strcmp(argv[3], "-") == 0;
    std::vector<cv::KeyPoint> kpts1, kpts2;
    cv::Mat desc1, desc2;
    std::vector<cv::DMatch> matches;
    DescriptorsMatcher dm(fs, img1, img2);
    if (!detect_here) {
    // upstream features
    FILE* f = fopen(argv[3], "rb");
    if (!f) { std::cerr << "no features file\n"; return 1; }
    int n1, n2, dim;
    rd(f, &n1, 4); rd(f, &n2, 4); rd(f, &dim, 4);
    std::vector<float> k1(2 * n1), k2(2 * n2);
    cv::Mat d1(n1, dim, CV_32FC1), d2(n2, dim, CV_32FC1);
    rd(f, k1.data(), 8 * n1); rd(f, k2.data(), 8 * n2);
    rd(f, d1.data, (size_t)4 * n1 * dim); rd(f, d2.data, (size_t)4 * n2 * dim);
    fclose(f);
    std::vector<cv::KeyPoint> in1(n1), in2(n2);
    for (int i = 0; i < n1; i++) in1[i] = cv::KeyPoint(k1[2 * i], k1[2 * i + 1], 1.f);
    for (int i = 0; i < n2; i++) in2[i] = cv::KeyPoint(k2[2 * i], k2[2 * i + 1], 1.f);

    dm.setFeatures(in1, d1, in2, d2);
    }
    dm.compareWithNNDR(fs["NNDR"]["epsilon"], matches, kpts1, kpts2, desc1, desc2);

    std::vector<double> pos1, pos2;
    fs["IMAGES"]["pos1"] >> pos1;
    fs["IMAGES"]["pos2"] >> pos2;
    cv::Vec3d translation1(pos1[0], pos1[1], pos1[2]), translation2(pos2[0], pos2[1], pos2[2]);
    cv::Vec3d rodrigues1(pos1[3], pos1[4], pos1[5]), rodrigues2(pos2[3], pos2[4], pos2[5]);
    cv::Matx44d g12;
    std::vector<cv::Vec3d> triagulated;
    std::vector<bool> outliersMask;
    SingleCameraTriangulator sct(fs);
    sct.setKeypoints(kpts1, kpts2, matches);
    sct.setg12(translation1, translation2, rodrigues1, rodrigues2, g12);
    sct.triangulate(triagulated, outliersMask);
    std::vector<cv::Vec3d> all_points = triagulated;

    NormalOptimizer no(fs, &sct);
    if (const char* e = getenv("FM3D_PENALTY")) no.setPenaltyMode(atoi(e));
    std::vector<cv::Vec3d> normalsVector;
    std::vector<cv::Scalar> colors(triagulated.size(), cv::Scalar(150, 150, 255));
    no.setImages(img1, img2);
    no.startVisualizerThread();
    no.computeOptimizedNormals(triagulated, normalsVector, colors);
    std::vector<cv::Matx44d> featuresFrames;
    no.computeFeaturesFrames(triagulated, normalsVector, featuresFrames);

    NeighborhoodsGenerator ng(fs);
    std::vector<cv::Vec3d> referenceNeighborhood;
    ng.getReferenceSquaredNeighborhood(referenceNeighborhood);
    std::vector<cv::Mat> patchesVector, imagePointsVector;
    sct.setImages(img1, img2);
    sct.projectReferencePointsToImageWithFrames(referenceNeighborhood, featuresFrames, patchesVector, imagePointsVector);
    std::vector<std::vector<cv::Vec3d> > neighborhoodsVector;
    ng.computeSquareNeighborhoodsByNormals(featuresFrames, neighborhoodsVector);
    no.stopVisualizerThread();

    // results
    FILE* o = fopen(argv[4], "wb");
    int nm = (int)matches.size(), np = (int)all_points.size(), nn = (int)normalsVector.size();
    int S = patchesVector.empty() ? 0 : patchesVector[0].rows;
    fwrite(&nm, 4, 1, o); fwrite(&np, 4, 1, o); fwrite(&nn, 4, 1, o); fwrite(&S, 4, 1, o);
    for (int i = 0; i < nm; i++) { fwrite(&matches[i].queryIdx, 4, 1, o); fwrite(&matches[i].trainIdx, 4, 1, o); fwrite(&matches[i].distance, 4, 1, o); }
    for (int i = 0; i < nm; i++) { unsigned char b = outliersMask[i]; fwrite(&b, 1, 1, o); }
    fwrite(g12.val, 8, 16, o);
    for (int i = 0; i < np; i++) fwrite(all_points[i].val, 8, 3, o);
    for (int i = 0; i < np; i++) fwrite(&no.lastStatus()[i], 4, 1, o);
    for (int i = 0; i < nn; i++) fwrite(triagulated[i].val, 8, 3, o);
    for (int i = 0; i < nn; i++) fwrite(normalsVector[i].val, 8, 3, o);
    for (int i = 0; i < nn; i++) fwrite(featuresFrames[i].val, 8, 16, o);
    for (int i = 0; i < nn; i++) fwrite(patchesVector[i].data, 1, (size_t)S * S, o);
    for (int i = 0; i < nn; i++) fwrite(neighborhoodsVector[i][S * S - 1].val, 8, 3, o);
    cv::Vec3d g = no.getGravity();
    fwrite(g.val, 8, 3, o);
    // evaluateNormal (normaloptimizer.cpp:65-149) by hand through the four public helpers of
    // SingleCameraTriangulator, for the first surviving feature at its refined normal, full resolution
    int hm = 0, rc1 = 0, rc2 = 0, rc3 = 0;
    double hcost = 0;
    if (nn > 0) {
        std::vector<Pixel> px, px2;
        std::vector<cv::Vec3d> pts;
        sct.extractPixelsContour(triagulated[0], px);
        hm = (int)px.size();
        cv::Mat pixelMat(hm, 1, CV_64FC2);
        for (int i = 0; i < hm; i++) { pixelMat.ptr<double>(i)[0] = px[i].x_; pixelMat.ptr<double>(i)[1] = px[i].y_; }
        rc1 = sct.get3dPointsFromImage1Pixels(triagulated[0], normalsVector[0], pixelMat, pts);
        rc2 = sct.updateImage1PixelsIntensity(1.0, px);
        rc3 = sct.projectPointsToImage2(pts, 1.0, px2);
        for (size_t i = 0; i < px2.size() && i < px.size(); i++) { const double d = (double)(px[i].i_ - px2[i].i_); hcost += d * d; }
    }
    fwrite(&hm, 4, 1, o); fwrite(&rc1, 4, 1, o); fwrite(&rc2, 4, 1, o); fwrite(&rc3, 4, 1, o); fwrite(&hcost, 8, 1, o);
    // optional: the circular variant of NeighborhoodsGenerator from a second settings file (method: circular)
    int cs = 0;
    if (argc > 5 && nn > 0) {
        cv::FileStorage fs2;
        fs2.open(argv[5], cv::FileStorage::READ);
        NeighborhoodsGenerator ngc(fs2);
        cv::Mat nb;
        cv::Vec3d nrm(0, 0, 0);                       // "initial guess"
        ngc.computeCircularNeighborhoodByNormal(triagulated[0], nrm, nb);
        cs = nb.cols;
        fwrite(&cs, 4, 1, o);
        fwrite(nrm.val, 8, 3, o);
        fwrite(nb.data, 8, (size_t)3 * cs, o);
    } else {
        fwrite(&cs, 4, 1, o);
    }
    // MOSAIC descriptors of the rectified patches (main.cpp:182-183)
    cv::Mat patchDescriptors;
    try {
        dm.extractDescriptorsFromPatches(patchesVector, patchDescriptors);
    } catch (const std::exception& e) {      // runs for ExtractorType SIFT only (the adapter says so)
        std::cerr << e.what() << std::endl;
        patchDescriptors = cv::Mat();
    }
    int dr = patchDescriptors.rows, dc = patchDescriptors.cols;
    const int pes = dr > 0 ? (int)patchDescriptors.elemSize() : 4;
    int dc_tagged = pes == 1 ? -dc : dc;                 // negative column count: CV_8U rows (binary extractors)
    fwrite(&dr, 4, 1, o); fwrite(&dc_tagged, 4, 1, o);
    if (dr > 0) fwrite(patchDescriptors.data, (size_t)pes, (size_t)dr * dc, o);
    if (detect_here) {
        // what compareWithNNDR returned for the two frames: keypoints (pt, size, angle, response) and descriptors
        const std::vector<cv::KeyPoint>* ks[2] = {&kpts1, &kpts2};
        const cv::Mat* ds[2] = {&desc1, &desc2};
        for (int a = 0; a < 2; a++) {
            int n = (int)ks[a]->size(), cols = ds[a]->cols, es = n > 0 ? (int)ds[a]->elemSize() : 4;
            fwrite(&n, 4, 1, o); fwrite(&cols, 4, 1, o); fwrite(&es, 4, 1, o);
            for (int i = 0; i < n; i++) {
                const cv::KeyPoint& k = (*ks[a])[i];
                const float v[5] = {k.pt.x, k.pt.y, k.size, k.angle, k.response};
                fwrite(v, 4, 5, o);
            }
            if (n > 0) fwrite(ds[a]->data, (size_t)es, (size_t)n * cols, o);
        }
    }
    fclose(o);
    std::cout << nm << " matches, " << np << " inliers, " << nn << " normals, patches " << S << "x" << S << std::endl;
    return 0;
}
