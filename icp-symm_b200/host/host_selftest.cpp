// host_selftest.cpp -- CPU-only checks of the host logic (PCD I/O); run by tests/test_host.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "pcd_io.h"
#include "../csrc/parts.h"

using namespace icpsymm;

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    PointCloud<PointXYZ> c;
    PointCloud<Normal> n;
    for (int i = 0; i < 1000; i++) {
        PointXYZ p; p.x = std::sin(i * 0.37f) * 123.456f; p.y = 1e-7f * i - 3.0f; p.z = (float)i / 3.0f;
        Normal m; m.normal_x = std::cos(i * 1.f); m.normal_y = 0.25f; m.normal_z = -m.normal_x; m.curvature = 1e-3f * i;
        c.points.push_back(p); n.points.push_back(m);
    }
    for (int binary = 0; binary < 2; binary++)
        for (int with_n = 0; with_n < 2; with_n++) {
            std::string path = dir + "/selftest_" + std::to_string(binary) + std::to_string(with_n) + ".pcd";
            CHECK(savePCDFile(path, c, with_n ? &n : nullptr, binary != 0) == 0);
            PointCloud<PointXYZ> r; PointCloud<Normal> rn;
            CHECK(loadPCDFile(path, r, &rn) == 0);
            CHECK(r.points.size() == c.points.size());
            for (size_t i = 0; i < r.points.size(); i++) {      // bit-exact round trip (ascii uses %.9g)
                CHECK(memcmp(&r.points[i].x, &c.points[i].x, 12) == 0);
                if (with_n) CHECK(memcmp(&rn.points[i], &n.points[i], 16) == 0);
            }
            CHECK(rn.points.size() == (with_n ? c.points.size() : 0));
        }
    // the cat_out.pcd layout: 8 fields incl. an unsigned label (cat_out.pcd:3-6); extras are dropped
    {
        std::string path = dir + "/selftest_fields.pcd";
        FILE* f = fopen(path.c_str(), "w");
        fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z label normal_x normal_y normal_z curvature\n"
                   "SIZE 4 4 4 4 4 4 4 4\nTYPE F F F U F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH 2\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\n"
                   "POINTS 2\nDATA ascii\n-24.415962 3.1901522 30.919439 0 0 0 0 0\n1 2 3 7 0 0 1 0.5\n");
        fclose(f);
        PointCloud<PointXYZ> r; PointCloud<Normal> rn;
        CHECK(loadPCDFile(path, r, &rn) == 0 && r.points.size() == 2);
        CHECK(r.points[0].x == -24.415962f && r.points[1].z == 3.f && rn.points[1].normal_z == 1.f && rn.points[1].curvature == 0.5f);
    }
    PointCloud<PointXYZ> bad;
    CHECK(loadPCDFile(dir + "/does_not_exist.pcd", bad) != 0);   // the reference ignores this (myicp.cpp:23)
    {
        std::string path = dir + "/selftest_trunc.pcd";
        FILE* f = fopen(path.c_str(), "w");
        fprintf(f, "VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 3\nHEIGHT 1\nPOINTS 3\nDATA ascii\n1 2 3\n4 5 6\n");
        fclose(f);
        CHECK(loadPCDFile(path, bad) != 0);                      // truncated file
    }
    // hostile or unsupported headers come back as error codes: no exception, no giant allocation, no zeros
    {
        const char* hdrs[] = {
            "VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 18446744073709551615\nDATA ascii\n1 2 3\n",
            "VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 4000000000\nHEIGHT 4000000000\nDATA binary\n",
            "VERSION 0.7\nFIELDS x y z\nSIZE -4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n1 2 3\n",
            "VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 0 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n1 2 3\n",
            "VERSION 0.7\nFIELDS x y z\nSIZE 2 2 2\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\nabcdef",
            "VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE Q F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n1 2 3\n",
        };
        for (const char* h : hdrs) {
            std::string path = dir + "/selftest_bad.pcd";
            FILE* f = fopen(path.c_str(), "wb");
            fputs(h, f);
            fclose(f);
            PointCloud<PointXYZ> r;
            CHECK(loadPCDFile(path, r) < 0);
        }
        // 8-byte fields are part of the format: double coordinates load as values, not as zeros
        std::string path = dir + "/selftest_f8.pcd";
        FILE* f = fopen(path.c_str(), "wb");
        fputs("VERSION 0.7\nFIELDS x y z\nSIZE 8 8 8\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n", f);
        const double xyz[3] = {1.5, -2.25, 3.0};
        fwrite(xyz, 8, 3, f);
        fclose(f);
        PointCloud<PointXYZ> r;
        CHECK(loadPCDFile(path, r) == 0 && r.points.size() == 1 && r.points[0].x == 1.5f && r.points[0].y == -2.25f);
    }
    {
        std::string path = dir + "/selftest_za.txt";
        FILE* f = fopen(path.c_str(), "w");
        fprintf(f, "-3.03387\t-3.03879\t-2.05533\n-3.03776\t-3.03882\t-2.05536\n");
        fclose(f);
        PointCloud<PointXYZ> t;
        CHECK(loadXYZText(path, t) == 0 && t.points.size() == 2 && t.points[1].y == -3.03882f);
    }
    // the canonical partition (csrc/parts.h): for every tile count the 8 parts are a partition of the tiles, a part's
    // j-th tile enumerates it in increasing order, a run of parts enumerates exactly its tiles in curve order, the work
    // of a run is monotone in the tile count, and the parts are balanced to within one chunk
    {
        uint32_t prev_work[9] = {0};
        for (uint32_t T = 1; T <= 6000; T += (T < 700 ? 1 : 37)) {
            std::vector<int> owner(T, -1);
            uint32_t total = 0, nmin = ~0u, nmax = 0;
            for (uint32_t o = 0; o < ICP_NPARTS; o++) {
                const uint32_t n_o = icp_part_tiles(T, o);
                total += n_o; nmin = n_o < nmin ? n_o : nmin; nmax = n_o > nmax ? n_o : nmax;
                uint32_t last = 0;
                for (uint32_t j = 0; j < n_o; j++) {
                    const uint32_t t = icp_part_tile(T, o, j);
                    CHECK(t < T && owner[t] == -1 && (j == 0 || t > last));
                    owner[t] = (int)o; last = t;
                }
            }
            CHECK(total == T && nmax - nmin <= icp_part_chunk(T));
            for (uint32_t W = 1; W <= 8; W++)
                for (uint32_t r = 0; r < W; r++) {
                    const uint32_t p0 = ICP_NPARTS * r / W, np = ICP_NPARTS * (r + 1) / W - p0;
                    const uint32_t work = icp_shard_work(T, np);
                    uint32_t seen = 0, last = 0;
                    for (uint32_t w = 0; w < work; w++) {
                        const uint32_t t = icp_shard_tile(T, p0, np, w);
                        if (t >= T) continue;
                        CHECK(owner[t] >= (int)p0 && owner[t] < (int)(p0 + np) && (seen == 0 || t > last));
                        seen++; last = t;
                    }
                    uint32_t want = 0;
                    for (uint32_t o = p0; o < p0 + np; o++) want += icp_part_tiles(T, o);
                    CHECK(seen == want);
                }
            for (uint32_t np = 1; np <= 8; np++) { CHECK(icp_shard_work(T, np) >= prev_work[np]); prev_work[np] = icp_shard_work(T, np); }
        }
    }
    printf("host selftest ok\n");
    return 0;
}
