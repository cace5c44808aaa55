#include "pcd_io.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <vector>

namespace icpsymm {

namespace {
struct Field {
    std::string name;
    int size = 4, count = 1;
    char type = 'F';
    size_t offset = 0;
};

// the encodings the PCD format defines: F4 F8, I1 I2 I4 I8, U1 U2 U4 U8
bool supported(const Field& f) {
    if (f.type == 'F') return f.size == 4 || f.size == 8;
    if (f.type == 'I' || f.type == 'U') return f.size == 1 || f.size == 2 || f.size == 4 || f.size == 8;
    return false;
}

double read_scalar(const char* p, const Field& f) {
    switch (f.type) {
        case 'F':
            if (f.size == 4) { float v; memcpy(&v, p, 4); return v; }
            if (f.size == 8) { double v; memcpy(&v, p, 8); return v; }
            break;
        case 'U':
            if (f.size == 1) { uint8_t v; memcpy(&v, p, 1); return v; }
            if (f.size == 2) { uint16_t v; memcpy(&v, p, 2); return v; }
            if (f.size == 4) { uint32_t v; memcpy(&v, p, 4); return v; }
            if (f.size == 8) { uint64_t v; memcpy(&v, p, 8); return (double)v; }
            break;
        case 'I':
            if (f.size == 1) { int8_t v; memcpy(&v, p, 1); return v; }
            if (f.size == 2) { int16_t v; memcpy(&v, p, 2); return v; }
            if (f.size == 4) { int32_t v; memcpy(&v, p, 4); return v; }
            if (f.size == 8) { int64_t v; memcpy(&v, p, 8); return (double)v; }
            break;
    }
    return 0.0;   // unreachable: loadPCDFile rejects every other encoding (-6)
}
}  // namespace

int loadPCDFile(const std::string& path, PointCloud<PointXYZ>& cloud, PointCloud<Normal>* normals) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return -1;
    std::vector<Field> fields;
    size_t npts = 0, width = 0, height = 1;
    bool have_points = false;
    std::string data_kind, line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ls(line);
        std::string key;
        ls >> key;
        if (key == "FIELDS") {
            std::string f;
            while (ls >> f) { Field fd; fd.name = f; fields.push_back(fd); }
        } else if (key == "SIZE") {
            for (auto& f : fields) ls >> f.size;
        } else if (key == "TYPE") {
            for (auto& f : fields) ls >> f.type;
        } else if (key == "COUNT") {
            for (auto& f : fields) ls >> f.count;
        } else if (key == "WIDTH") {
            ls >> width;
        } else if (key == "HEIGHT") {
            ls >> height;
        } else if (key == "POINTS") {
            ls >> npts;
            have_points = true;
        } else if (key == "DATA") {
            ls >> data_kind;
            break;
        }
    }
    if (fields.empty() || data_kind.empty()) return -2;
    if (!in) return -2;
    // the header is untrusted input: reject what the format does not define instead of wrapping sizes or
    // reading unsupported encodings as zeros (MyICP::LoadCloud returns these codes, it never throws)
    for (const auto& f : fields)
        if (f.count < 1 || f.count > 4096 || !supported(f)) return -6;
    if (!have_points) {
        if (height != 0 && width > (size_t)-1 / height) return -6;
        npts = width * height;
    }
    {
        // a point needs at least one byte per field (binary) or two characters per column (ascii): bound npts by
        // the bytes that are actually left in the file before allocating anything
        const std::streampos here = in.tellg();
        in.seekg(0, std::ios::end);
        const std::streampos end = in.tellg();
        in.seekg(here);
        if (here < 0 || end < here) return -2;
        const size_t left = (size_t)(end - here);
        size_t rec = 0;
        for (const auto& f : fields) rec += (size_t)f.size * (size_t)f.count;
        const size_t min_bytes = data_kind == "binary" ? rec : 2 * fields.size() - 1;
        if (min_bytes == 0 || npts > left / min_bytes) return -4;
    }
    int ix = -1, iy = -1, iz = -1, inx = -1, iny = -1, inz = -1, icv = -1;
    size_t off = 0, ncols = 0;
    for (size_t i = 0; i < fields.size(); i++) {
        fields[i].offset = off;
        off += (size_t)fields[i].size * fields[i].count;
        const std::string& n = fields[i].name;
        if (n == "x") ix = (int)i; else if (n == "y") iy = (int)i; else if (n == "z") iz = (int)i;
        else if (n == "normal_x") inx = (int)i; else if (n == "normal_y") iny = (int)i;
        else if (n == "normal_z") inz = (int)i; else if (n == "curvature") icv = (int)i;
        ncols += fields[i].count;
    }
    if (ix < 0 || iy < 0 || iz < 0) return -3;
    const bool want_n = normals && inx >= 0 && iny >= 0 && inz >= 0;
    cloud.points.assign(npts, PointXYZ());
    cloud.width = (uint32_t)(height > 1 ? width : npts);
    cloud.height = (uint32_t)height;
    if (normals) { normals->points.assign(want_n ? npts : 0, Normal()); normals->width = (uint32_t)normals->points.size(); }
    if (data_kind == "ascii") {
        std::vector<size_t> col_of(fields.size());
        size_t c = 0;
        for (size_t i = 0; i < fields.size(); i++) { col_of[i] = c; c += fields[i].count; }
        std::vector<double> vals(ncols);
        size_t got = 0;
        while (got < npts && std::getline(in, line)) {
            std::istringstream ls(line);
            size_t k = 0;
            std::string tok;
            while (k < ncols && ls >> tok) vals[k++] = strtod(tok.c_str(), nullptr);
            if (k < ncols) { if (k == 0) continue; return -4; }
            PointXYZ& p = cloud.points[got];
            p.x = (float)vals[col_of[ix]]; p.y = (float)vals[col_of[iy]]; p.z = (float)vals[col_of[iz]];
            if (want_n) {
                Normal& n = normals->points[got];
                n.normal_x = (float)vals[col_of[inx]]; n.normal_y = (float)vals[col_of[iny]]; n.normal_z = (float)vals[col_of[inz]];
                if (icv >= 0) n.curvature = (float)vals[col_of[icv]];
            }
            got++;
        }
        if (got != npts) return -4;
    } else if (data_kind == "binary") {
        std::vector<char> rec(off);
        for (size_t i = 0; i < npts; i++) {
            in.read(rec.data(), (std::streamsize)off);
            if (!in) return -4;
            PointXYZ& p = cloud.points[i];
            p.x = (float)read_scalar(rec.data() + fields[ix].offset, fields[ix]);
            p.y = (float)read_scalar(rec.data() + fields[iy].offset, fields[iy]);
            p.z = (float)read_scalar(rec.data() + fields[iz].offset, fields[iz]);
            if (want_n) {
                Normal& n = normals->points[i];
                n.normal_x = (float)read_scalar(rec.data() + fields[inx].offset, fields[inx]);
                n.normal_y = (float)read_scalar(rec.data() + fields[iny].offset, fields[iny]);
                n.normal_z = (float)read_scalar(rec.data() + fields[inz].offset, fields[inz]);
                if (icv >= 0) n.curvature = (float)read_scalar(rec.data() + fields[icv].offset, fields[icv]);
            }
        }
    } else {
        return -5;  // binary_compressed is not produced by anything in the reference
    }
    return 0;
}

int savePCDFile(const std::string& path, const PointCloud<PointXYZ>& cloud, const PointCloud<Normal>* normals, bool binary) {
    const bool with_n = normals && normals->points.size() == cloud.points.size() && !cloud.points.empty();
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return -1;
    const size_t n = cloud.points.size();
    fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n");
    if (with_n) fprintf(f, "FIELDS x y z normal_x normal_y normal_z curvature\nSIZE 4 4 4 4 4 4 4\nTYPE F F F F F F F\nCOUNT 1 1 1 1 1 1 1\n");
    else fprintf(f, "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n");
    fprintf(f, "WIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA %s\n", n, n, binary ? "binary" : "ascii");
    for (size_t i = 0; i < n; i++) {
        const PointXYZ& p = cloud.points[i];
        if (binary) {
            float rec[7] = {p.x, p.y, p.z, 0, 0, 0, 0};
            if (with_n) { const Normal& m = normals->points[i]; rec[3] = m.normal_x; rec[4] = m.normal_y; rec[5] = m.normal_z; rec[6] = m.curvature; }
            fwrite(rec, sizeof(float), with_n ? 7 : 3, f);
        } else {
            fprintf(f, "%.9g %.9g %.9g", p.x, p.y, p.z);   // 9 significant digits round-trip fp32 exactly
            if (with_n) { const Normal& m = normals->points[i]; fprintf(f, " %.9g %.9g %.9g %.9g", m.normal_x, m.normal_y, m.normal_z, m.curvature); }
            fputc('\n', f);
        }
    }
    int rc = ferror(f) ? -2 : 0;
    fclose(f);
    return rc;
}

int loadXYZText(const std::string& path, PointCloud<PointXYZ>& cloud) {
    std::ifstream in(path);
    if (!in) return -1;
    cloud.points.clear();
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        PointXYZ p;
        if (ls >> p.x >> p.y >> p.z) cloud.points.push_back(p);
    }
    cloud.width = (uint32_t)cloud.points.size();
    cloud.height = 1;
    return cloud.points.empty() ? -2 : 0;
}

}  // namespace icpsymm
