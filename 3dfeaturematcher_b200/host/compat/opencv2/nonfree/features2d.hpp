// compat: <opencv2/nonfree/features2d.hpp> (cv::SURF / cv::SIFT of OpenCV 2.4): included by main.cpp:13 and
// mosaic.h:38, nothing of it is used once DescriptorsMatcher is the fm3d adapter
