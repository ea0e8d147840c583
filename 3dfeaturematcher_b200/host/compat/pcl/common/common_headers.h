// compat: the two PCL types the reference's main.cpp touches (main.cpp:199-207): a cloud of normals that is filled and
// handed to a viewer.  No PCL functionality.
#ifndef FM3D_COMPAT_PCL_COMMON_HEADERS_H_
#define FM3D_COMPAT_PCL_COMMON_HEADERS_H_
#include <memory>
#include <vector>
namespace pcl {
struct Normal {
    float normal_x, normal_y, normal_z, curvature;
    Normal() : normal_x(0), normal_y(0), normal_z(0), curvature(0) {}
    Normal(float x, float y, float z) : normal_x(x), normal_y(y), normal_z(z), curvature(0) {}
};
struct PointXYZ { float x, y, z; };
struct PointXYZRGB { float x, y, z; unsigned char r, g, b; };
template <typename PointT>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<PointT> > Ptr;
    typedef std::shared_ptr<const PointCloud<PointT> > ConstPtr;
    std::vector<PointT> points;
    unsigned width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    void push_back(const PointT& p) { points.push_back(p); }
};
}  // namespace pcl
#endif
