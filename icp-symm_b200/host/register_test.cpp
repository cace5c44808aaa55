// register_test.cpp -- counterpart of ICP/register-test.cpp:15-67: load two PCD files with the
// same number of points, run the closed-form fit registrateNPoint, print R and T.
#include <cstdio>

#include "pcd_io.h"
#include "regist.h"

int main(int argc, char** argv) {
    const char* a = argc > 1 ? argv[1] : "cat.pcd";
    const char* b = argc > 2 ? argv[2] : "cat_out.pcd";
    icpsymm::PointCloud<icpsymm::PointXYZ> cloud_in, cloud_out;
    if (icpsymm::loadPCDFile(a, cloud_in) != 0) { fprintf(stderr, "Couldn't read file\n"); return -1; }     // :21-25
    if (icpsymm::loadPCDFile(b, cloud_out) != 0) { fprintf(stderr, "Couldn't read file\n"); return -1; }    // :26-30
    std::vector<icpsymm::Point3d> src, dst;                                                                  // :55-60
    for (size_t i = 0; i < cloud_in.points.size(); ++i)
        src.push_back(icpsymm::Point3d(cloud_in.points[i].x, cloud_in.points[i].y, cloud_in.points[i].z));
    for (size_t i = 0; i < cloud_out.points.size(); ++i)
        dst.push_back(icpsymm::Point3d(cloud_out.points[i].x, cloud_out.points[i].y, cloud_out.points[i].z));
    icpsymm::Mat R, T;
    if (!icpsymm::registrateNPoint(src, dst, R, T)) return 1;                                                // :63
    printf("R");
    for (double v : R.data) printf(" %.12g", v);
    printf("\nT");
    for (double v : T.data) printf(" %.12g", v);
    printf("\n");
    return 0;
}
