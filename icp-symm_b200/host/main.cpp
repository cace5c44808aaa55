// main.cpp -- headless counterpart of the reference's driver (ICP/main.cpp:5-68): load two PCD
// files, run the symmetric registration, print the transform.  The PCLVisualizer window of the
// reference (main.cpp:13-40) is out of scope; --out writes the aligned source instead.
//   icp_symm_cli [src.pcd tgt.pcd] [--iters N] [--thr X] [--rel-tol X] [--p2p] [--out aligned.pcd] [--quiet]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "myicp.h"
#include "pcd_io.h"

int main(int argc, char** argv) {
    std::string src = "cat.pcd", tgt = "cat_out.pcd", out;      // main.cpp:8
    int iters = 10;
    float thr = 1.f;
    double rel = 0.0;
    bool p2p = false, quiet = false;
    int pos = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--thr") && i + 1 < argc) thr = (float)atof(argv[++i]);
        else if (!strcmp(argv[i], "--rel-tol") && i + 1 < argc) rel = atof(argv[++i]);
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--p2p")) p2p = true;
        else if (!strcmp(argv[i], "--quiet")) quiet = true;
        else if (pos == 0) { src = argv[i]; pos++; }
        else if (pos == 1) { tgt = argv[i]; pos++; }
    }
    icpsymm::MyICP myicp;
    if (myicp.LastStatus() != 0) { fprintf(stderr, "no CUDA device: %s\n", myicp.LastError()); return 2; }
    myicp.SetVerbose(!quiet);
    if (myicp.LoadCloud(src, tgt) != 0) return 1;
    if (rel > 0) myicp.SetRelativeTolerance(rel);
    if (p2p) myicp.RegisterP2P(iters, thr);
    else myicp.RegisterSymm(iters, thr);                        // main.cpp:10
    if (myicp.LastStatus() < 0) return 3;
    const icpsymm::RegistrationResult& r = myicp.GetResult();
    printf("iterations %d converged %d degenerate %d diff %.9g rmse %.9g loop_ms %.3f\n", r.iterations, r.converged,
           (int)r.degenerate, r.diff, r.rmse, r.loop_ms);
    printf("T");
    for (float v : r.transform) printf(" %.9g", v);
    printf("\n");
    if (!out.empty()) {
        if (myicp.ApplyTransformToSource() < 0) return 4;
        if (icpsymm::savePCDFile(out, *myicp.GetSrcCloud()) != 0) return 5;
    }
    return 0;
}
