#include "myicp.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/icpsymm.h"
#include "pcd_io.h"

namespace icpsymm {

// myicp.cpp:6: max_iters(10), diff_threshold(1.f)
MyICP::MyICP(int device) : max_iters(10), diff_threshold(1.f) {
    cloud_src = std::make_shared<PointCloud<PointT>>();
    cloud_tgt = std::make_shared<PointCloud<PointT>>();
    result_.transform = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    status_ = icpsymm_create(&ctx_, device);
    if (status_ != ICPSYMM_OK) ctx_ = nullptr;   // no CPU fallback: every later call reports the error
    // upload, index build and normals of the two clouds run on their own streams and overlap
    else icpsymm_set_option(ctx_, ICPSYMM_OPT_ASYNC_SETUP, 1.0);
}

MyICP::~MyICP() {
    if (ctx_) icpsymm_destroy(ctx_);
}

const char* MyICP::LastError() const { return ctx_ ? icpsymm_last_error(ctx_) : icpsymm_status_string(status_); }

// myicp.cpp:20-31 (which ignores the reader's return codes and always returns 0)
int MyICP::LoadCloud(std::string src_path, std::string tgt_path) {
    int rc = loadPCDFile(src_path, *cloud_src);
    if (rc != 0) { fprintf(stderr, "Couldn't read file %s\n", src_path.c_str()); return rc; }
    rc = loadPCDFile(tgt_path, *cloud_tgt);
    if (rc != 0) { fprintf(stderr, "Couldn't read file %s\n", tgt_path.c_str()); return rc; }
    uploaded_ = false;
    return 0;
}

void MyICP::SetClouds(PointCloud<PointT>::Ptr src, PointCloud<PointT>::Ptr tgt) {
    cloud_src = src;
    cloud_tgt = tgt;
    uploaded_ = false;
}

PointCloud<PointT>::Ptr MyICP::GetSrcCloud() { return this->cloud_src; }
PointCloud<PointT>::Ptr MyICP::GetTgtCloud() { return this->cloud_tgt; }

int MyICP::upload() {
    if (!ctx_) return status_;
    if (uploaded_) return ICPSYMM_OK;
    if (cloud_src->empty() || cloud_tgt->empty()) return status_ = ICPSYMM_ERR_NOT_READY;
    status_ = icpsymm_set_source(ctx_, &cloud_src->points[0].x, cloud_src->size(), sizeof(PointT));
    if (status_ < 0) return status_;
    status_ = icpsymm_set_target(ctx_, &cloud_tgt->points[0].x, cloud_tgt->size(), sizeof(PointT));
    if (status_ < 0) return status_;
    uploaded_ = true;
    return status_;
}

// myicp.cpp:152-172: NormalEstimation, setKSearch(10), viewpoint (0,0,0)
void MyICP::estimateNormals() {
    if (upload() < 0) return;
    status_ = icpsymm_estimate_normals(ctx_, 10, nullptr);
}

void MyICP::SetRelativeTolerance(double tol) {
    if (ctx_) icpsymm_set_option(ctx_, ICPSYMM_OPT_REL_DIFF_TOL, tol);
}

void MyICP::printResult(const char* title) const {
    if (!verbose_) return;
    const float* T = result_.transform.data();
    printf("%s\n", title);
    for (int r = 0; r < 4; r++) printf("%12.6f %12.6f %12.6f %12.6f\n", T[4 * r], T[4 * r + 1], T[4 * r + 2], T[4 * r + 3]);
    printf("  rotation:\n");
    for (int r = 0; r < 3; r++) printf("%12.6f %12.6f %12.6f\n", T[4 * r], T[4 * r + 1], T[4 * r + 2]);
    printf("  translation:\n%12.6f\n%12.6f\n%12.6f\n", T[3], T[7], T[11]);
}

void MyICP::RegisterSymm() { RegisterSymm(max_iters, diff_threshold); }

// myicp.cpp:100-150
void MyICP::RegisterSymm(int iters, float threshold) {
    estimateNormals();                                   // :105
    if (status_ < 0) { fprintf(stderr, "RegisterSymm: %s\n", LastError()); return; }
    icpsymm_stats st;
    float T[16];
    status_ = icpsymm_register_symm(ctx_, iters, threshold, T, &st);   // :117-142
    if (status_ < 0) { fprintf(stderr, "RegisterSymm: %s\n", LastError()); return; }
    memcpy(result_.transform.data(), T, sizeof T);
    result_.iterations = st.iters; result_.converged = st.converged; result_.degenerate = st.degenerate != 0;
    result_.diff = st.diff; result_.rmse = st.rmse; result_.loop_ms = st.loop_ms;
    if (verbose_) printf("iters# %d\ndiff: %g\n", st.iters, st.diff);     // :125-126
    printResult("Result transform:");                    // :146-149
}

void MyICP::RegisterP2P() { RegisterP2P(max_iters, diff_threshold); }

// The reference's RegisterP2P (myicp.cpp:43-59) only applies an identity guess.  This is the
// point-to-point loop it stands for: exact correspondences on the device, closed-form fit of the
// matched pairs (registrateNPoint, regist.h:8-72) on the device, repeat -- icpsymm_register_p2p: nothing but one diff
// and one 3x4 increment per iteration crosses the bus.
void MyICP::RegisterP2P(int iters, float threshold) {
    if (upload() < 0) { fprintf(stderr, "RegisterP2P: %s\n", LastError()); return; }
    icpsymm_stats st;
    status_ = icpsymm_register_p2p(ctx_, iters, threshold, result_.transform.data(), &st);
    if (status_ < 0) { fprintf(stderr, "RegisterP2P: %s\n", LastError()); return; }
    result_.iterations = st.iters; result_.diff = st.diff; result_.converged = st.converged != 0;
    result_.rmse = st.rmse; result_.degenerate = false;
    if (verbose_) printf("iters# %d\ndiff: %g\n", st.iters, st.diff);
    printResult("guess matrix:");                        // myicp.cpp:53
}

int MyICP::ApplyTransformToSource() {
    if (upload() < 0) return status_;
    status_ = icpsymm_get_transformed_source(ctx_, result_.transform.data(), &cloud_src->points[0].x, sizeof(PointT));
    uploaded_ = false;   // the source cloud changed on the host
    return status_;
}

}  // namespace icpsymm
