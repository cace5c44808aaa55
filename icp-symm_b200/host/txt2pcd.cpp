// txt2pcd.cpp -- counterpart of ICP/test.cpp:33-58: "x y z" text rows (za.txt) -> PCD file.
#include <cstdio>

#include "pcd_io.h"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: txt2pcd in.txt out.pcd [--binary]\n"); return 2; }
    icpsymm::PointCloud<icpsymm::PointXYZ> cloud;
    if (icpsymm::loadXYZText(argv[1], cloud) != 0) { fprintf(stderr, "Couldn't read %s\n", argv[1]); return 1; }
    if (icpsymm::savePCDFile(argv[2], cloud, nullptr, argc > 3) != 0) return 1;
    printf("Saved %zu data points to %s.\n", cloud.points.size(), argv[2]);
    return 0;
}
