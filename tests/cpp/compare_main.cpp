// Test driver for the raw kNN entry points of DescriptorsMatcher (descriptorsmatcher.cpp:74-105): compare() and
// crosscompare() through the adapter, with injected features or ("-") with features detected from the frames, called
// TWICE with the same output objects (the reference re-detects and overwrites them on every call, :110-115).
// Usage: compare_main -s settings.yml <features.bin | -> out.bin
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "DescriptorsMatcher/descriptorsmatcher.h"

static void rd(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

static void dump(FILE* o, const std::vector<std::vector<cv::DMatch> >& m) {
    int n = (int)m.size();
    fwrite(&n, 4, 1, o);
    for (int i = 0; i < n; i++) {
        int k = (int)m[i].size();
        fwrite(&k, 4, 1, o);
        for (int j = 0; j < k; j++) { fwrite(&m[i][j].queryIdx, 4, 1, o); fwrite(&m[i][j].trainIdx, 4, 1, o); fwrite(&m[i][j].distance, 4, 1, o); }
    }
}

int main(int argc, char** argv) {
    if (argc != 5 || strcmp(argv[1], "-s") != 0) { std::cerr << "usage: -s <settings.yml> <features.bin | -> <out.bin>\n"; return 1; }
    cv::FileStorage fs;
    fs.open(argv[2], cv::FileStorage::READ);
    if (!fs.isOpened()) return 1;
    std::string IMG_1, IMG_2;
    fs["IMAGES"]["img1"] >> IMG_1;
    fs["IMAGES"]["img2"] >> IMG_2;
    cv::Mat img1 = cv::imread(IMG_1, CV_LOAD_IMAGE_GRAYSCALE), img2 = cv::imread(IMG_2, CV_LOAD_IMAGE_GRAYSCALE);
    if (img1.empty() || img2.empty()) return 1;
    try {
        DescriptorsMatcher dm(fs, img1, img2);
        if (strcmp(argv[3], "-") != 0) {
            FILE* f = fopen(argv[3], "rb");
            if (!f) return 1;
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
        std::vector<cv::KeyPoint> ka, kb;
        cv::Mat da, db;
        std::vector<std::vector<cv::DMatch> > m1, m2, ab, ba;
        std::vector<cv::DMatch> nndr;
        dm.compare(m1, ka, kb, da, db);
        // poison the outputs: a second call must overwrite them (stale descriptors of a previous frame must never be matched).
        // Only for detected features: injected ones are shallow cv::Mat copies of the caller's own matrices.
        if (strcmp(argv[3], "-") == 0 && !da.empty()) memset(da.data, 0, da.rows * da.step());
        ka.clear();
        dm.compare(m2, ka, kb, da, db);
        dm.crosscompare(ab, ba, ka, kb, da, db);
        dm.compareWithNNDR(fs["NNDR"]["epsilon"], nndr, ka, kb, da, db);
        const size_t once = nndr.size();
        dm.compareWithNNDR(fs["NNDR"]["epsilon"], nndr, ka, kb, da, db);   // appended, not cleared (:126)
        FILE* o = fopen(argv[4], "wb");
        int na = (int)ka.size(), nb = (int)kb.size(), dim = da.cols, es = (int)da.elemSize();
        fwrite(&na, 4, 1, o); fwrite(&nb, 4, 1, o); fwrite(&dim, 4, 1, o); fwrite(&es, 4, 1, o);
        fwrite(da.data, 1, da.rows * da.step(), o);
        fwrite(db.data, 1, db.rows * db.step(), o);
        dump(o, m1); dump(o, m2); dump(o, ab); dump(o, ba);
        int n1 = (int)once, n2 = (int)nndr.size();
        fwrite(&n1, 4, 1, o); fwrite(&n2, 4, 1, o);
        for (int i = 0; i < n2; i++) { fwrite(&nndr[i].queryIdx, 4, 1, o); fwrite(&nndr[i].trainIdx, 4, 1, o); fwrite(&nndr[i].distance, 4, 1, o); }
        fclose(o);
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 2;
    }
    return 0;
}
