// myicp.h -- C++ facade with the shape of the reference's class MyICP (ICP/myicp.h:7-36),
// written purely on top of the C ABI in include/icpsymm.h.
//
// Same public methods, same roles: LoadCloud / GetSrcCloud / GetTgtCloud / RegisterP2P /
// RegisterSymm.  Differences, each one a "todo" or a defect of the reference that SURVEY.md
// Appendix A lists: RegisterSymm takes the iteration cap and threshold its comment asks for
// (myicp.h:19; defaults 10 and 1.0f as hard-wired at myicp.cpp:6), LoadCloud reports I/O errors
// instead of always returning 0 (myicp.cpp:23-30), and the resulting 4x4 is kept and returned
// instead of only being printed (myicp.cpp:146-149).
#pragma once
#include <array>
#include <string>

#include "point_cloud.h"

struct icpsymm_ctx;

namespace icpsymm {

typedef PointXYZ PointT;   // myicp.h:5

struct RegistrationResult {
    std::array<float, 16> transform;   // row-major 4x4, source -> target
    int iterations = 0;
    int converged = 0;                 // 1: diff <= threshold, 2: relative tolerance, 0: iteration cap
    bool degenerate = false;
    double diff = 0.0, rmse = 0.0;
    float loop_ms = 0.f;
};

class MyICP {
public:
    explicit MyICP(int device = 0);
    ~MyICP();
    MyICP(const MyICP&) = delete;
    MyICP& operator=(const MyICP&) = delete;

public:
    int LoadCloud(std::string src_path, std::string tgt_path);             // myicp.h:14
    void SetClouds(PointCloud<PointT>::Ptr src, PointCloud<PointT>::Ptr tgt);
    PointCloud<PointT>::Ptr GetSrcCloud();                                  // myicp.h:15
    PointCloud<PointT>::Ptr GetTgtCloud();                                  // myicp.h:16

    void RegisterP2P();                                                     // myicp.h:18 (a stub in the reference)
    void RegisterP2P(int max_iters, float diff_threshold);
    void RegisterSymm();                                                    // myicp.h:19, reference defaults
    void RegisterSymm(int max_iters, float diff_threshold);                 // "todo add params to specify iters & diff"

    const RegistrationResult& GetResult() const { return result_; }
    const std::array<float, 16>& GetTransform() const { return result_.transform; }
    // writes the aligned source back into the cloud GetSrcCloud() returns (the reference never does,
    // myicp.cpp:136 updates only its Eigen copies)
    int ApplyTransformToSource();
    void SetVerbose(bool v) { verbose_ = v; }
    void SetRelativeTolerance(double tol);
    int LastStatus() const { return status_; }
    const char* LastError() const;

private:
    int max_iters;             // myicp.h:23
    float diff_threshold;      // myicp.h:24
    PointCloud<PointT>::Ptr cloud_src, cloud_tgt;   // myicp.h:27
    icpsymm_ctx* ctx_ = nullptr;
    RegistrationResult result_;
    bool verbose_ = true, uploaded_ = false;
    int status_ = 0;

    int upload();
    void estimateNormals();    // myicp.h:30
    void printResult(const char* title) const;
};

}  // namespace icpsymm
