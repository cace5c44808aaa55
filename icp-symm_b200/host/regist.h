// regist.h -- the reference's closed-form rigid fit (ICP/regist.h:8-72) behind the C ABI.
//
// Same name, argument order and meaning: registrateNPoint(srcPoints, dstPoints, R, T) with
// dst = R * src + T (what the reference's code computes; its comment at regist.h:7 states the
// opposite direction).  cv::Point3d / cv::Mat are replaced by minimal lookalikes because OpenCV
// is not a dependency.  A size mismatch or fewer than 3 pairs only prints in the reference
// (regist.h:9-15); here the call also returns false and leaves R, T untouched.
#pragma once
#include <cstdio>
#include <vector>

#include "../../include/icpsymm.h"

namespace icpsymm {

struct Point3d {
    double x = 0, y = 0, z = 0;
    Point3d() {}
    Point3d(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};

// row-major rows x cols matrix of doubles (CV_64F)
struct Mat {
    int rows = 0, cols = 0;
    std::vector<double> data;
    Mat() {}
    Mat(int r, int c) : rows(r), cols(c), data((size_t)r * c, 0.0) {}
    double& at(int r, int c) { return data[(size_t)r * cols + c]; }
    double at(int r, int c) const { return data[(size_t)r * cols + c]; }
};

inline bool registrateNPoint(std::vector<Point3d>& srcPoints, std::vector<Point3d>& dstPoints, Mat& R, Mat& T,
                             icpsymm_ctx* ctx = nullptr) {
    if (srcPoints.size() != dstPoints.size() || srcPoints.size() < 3 || dstPoints.size() < 3) {
        printf("srcPoints.size():\t%zu", srcPoints.size());
        printf("dstPoints.size():\t%zu", dstPoints.size());
        printf("registrateNPoint points size donot match!");
        return false;
    }
    icpsymm_ctx* own = nullptr;
    if (!ctx) {
        if (icpsymm_create(&own, 0) != ICPSYMM_OK) return false;
        ctx = own;
    }
    static_assert(sizeof(Point3d) == 3 * sizeof(double), "Point3d must be three packed doubles");
    double r[9], t[3];
    int rc = icpsymm_register_npoint(ctx, &srcPoints[0].x, &dstPoints[0].x, srcPoints.size(), r, t);
    if (own) icpsymm_destroy(own);
    if (rc < 0) return false;
    R = Mat(3, 3);
    T = Mat(3, 1);
    for (int i = 0; i < 9; i++) R.data[i] = r[i];
    for (int i = 0; i < 3; i++) T.data[i] = t[i];
    return true;
}

}  // namespace icpsymm
