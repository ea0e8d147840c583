// Test driver for the reference's second client, MOSAIC (mosaic.h:47-141, mosaic.cpp:32-73): its constructor runs steps 1-7
// of the pipeline (match, setg12, triangulate, normals, frames, reference neighbourhood, patches) on the four classes.
// Built by __graft_entry__.build() against the reference's own mosaic.h / mosaic.cpp (the latter without its lines 81-89,
// SURVEY D7) and the fm3d adapters.  Usage: ref_mosaic -s settings.yml      (writes patch_<k>.pgm into the CWD, like main)
#include <cstring>
#include <iostream>

#include "mosaic.h"

int main(int argc, char** argv) {
    if (argc != 3 || strcmp(argv[1], "-s") != 0) { std::cerr << "usage: -s <settings.yml>\n"; return 1; }
    cv::FileStorage fs;
    fs.open(argv[2], cv::FileStorage::READ);
    if (!fs.isOpened()) { std::cerr << "Could not open settings file: " << argv[2] << std::endl; return 1; }
    std::string IMG_1, IMG_2;
    fs["IMAGES"]["img1"] >> IMG_1;
    fs["IMAGES"]["img2"] >> IMG_2;
    cv::Mat imgA = cv::imread(IMG_1, CV_LOAD_IMAGE_GRAYSCALE), imgB = cv::imread(IMG_2, CV_LOAD_IMAGE_GRAYSCALE);
    if (imgA.empty() || imgB.empty()) { std::cerr << "could not read images\n"; return 1; }
    std::vector<double> pos1, pos2;
    fs["IMAGES"]["pos1"] >> pos1;
    fs["IMAGES"]["pos2"] >> pos2;
    try {
        MOSAIC mosaic(fs, imgA, imgB, cv::Vec3d(pos1[0], pos1[1], pos1[2]), cv::Vec3d(pos2[0], pos2[1], pos2[2]),
                      cv::Vec3d(pos1[3], pos1[4], pos1[5]), cv::Vec3d(pos2[3], pos2[4], pos2[5]));
        std::cout << "MOSAIC constructed" << std::endl;
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 2;
    }
    return 0;
}
