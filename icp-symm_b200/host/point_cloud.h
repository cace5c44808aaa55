// point_cloud.h -- PCL-free stand-ins for the few PCL types the reference's interface exposes.
//
// The reference's public interface (ICP/myicp.h:5,14-16) speaks pcl::PointXYZ and
// pcl::PointCloud<PointT>::Ptr.  PCL is not a dependency of this repository, so the facade uses
// layout-compatible lookalikes: PointXYZ is x,y,z + 4 bytes of padding (16 bytes, as in PCL),
// PointCloud<T> has `points`, `width`, `height` and a shared-ownership `Ptr`.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

namespace icpsymm {

struct alignas(16) PointXYZ {
    float x = 0.f, y = 0.f, z = 0.f, pad = 1.f;
};

struct alignas(16) Normal {
    float normal_x = 0.f, normal_y = 0.f, normal_z = 0.f, curvature = 0.f;
};

template <typename PointT>
struct PointCloud {
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    std::vector<PointT> points;
    uint32_t width = 0, height = 1;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
};

}  // namespace icpsymm
