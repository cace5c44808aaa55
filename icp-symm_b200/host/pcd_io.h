// pcd_io.h -- PCD v0.7 reader / writer.
//
// Replaces pcl::PCDReader::read + pcl::fromPCLPointCloud2 (ICP/myicp.cpp:22-26) and
// pcl::io::savePCDFile (ICP/main.cpp:52, ICP/test.cpp:58).  Formats as evidenced by the
// reference's fixtures: FIELDS x y z (cat.pcd:1-11) and FIELDS x y z label normal_x normal_y
// normal_z curvature (cat_out.pcd:1-11), DATA ascii; DATA binary (uncompressed) is read and
// written too.  Only x, y, z (and optionally the normal fields) survive, like the reference.
#pragma once
#include <string>

#include "point_cloud.h"

namespace icpsymm {

// 0 on success; negative on failure (the reference ignores reader errors, myicp.cpp:23,25,30)
int loadPCDFile(const std::string& path, PointCloud<PointXYZ>& cloud, PointCloud<Normal>* normals = nullptr);
int savePCDFile(const std::string& path, const PointCloud<PointXYZ>& cloud, const PointCloud<Normal>* normals = nullptr,
                bool binary = false);
// test.cpp:33-58: whitespace-separated "x y z" rows -> cloud
int loadXYZText(const std::string& path, PointCloud<PointXYZ>& cloud);

}  // namespace icpsymm
