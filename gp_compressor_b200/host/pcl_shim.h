// pcl_shim.h — the two PCL types the GP path's public API mentions, for builds without PCL.
// Same memory layout as pcl::PointXYZRGB (32 bytes: float x,y,z,1 | b,g,r,a | 12 bytes padding) and the
// members of pcl::PointCloud<T> the reference touches (points, width, height, Ptr/ConstPtr, resize/at).
// In a tree that has PCL, include <pcl/point_cloud.h> and <pcl/point_types.h> instead of this header.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

namespace pcl {

struct alignas(16) PointXYZRGB {
    float x = 0, y = 0, z = 0, w = 1.0f;
    uint8_t b = 0, g = 0, r = 0, a = 255;
    uint8_t pad[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};
static_assert(sizeof(PointXYZRGB) == 32, "PointXYZRGB must be 32 bytes");

template <class PointT>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    std::vector<PointT> points;
    uint32_t width = 0, height = 1;
    size_t size() const { return points.size(); }
    void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
    PointT& at(size_t i) { return points.at(i); }
    const PointT& at(size_t i) const { return points.at(i); }
    typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
    typename std::vector<PointT>::const_iterator end() const { return points.end(); }
};

}  // namespace pcl
