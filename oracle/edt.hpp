// ORACLE — test infrastructure only (see oracle/README.md).
// CPU restatement of the SFC (safe flight corridor) construction and of the upstream semantics it
// depends on. Citations relative to /root/reference:
//   include/corridor_constructor.hpp:18-44    expandBoxFromPoint
//   include/corridor_constructor.hpp:81-122   isObstacleInBox   (literal triple loop, float samples)
//   include/corridor_constructor.hpp:124-131  isBoxInBoundary
//   include/corridor_constructor.hpp:142-182  setAxisCand
//   include/corridor_constructor.hpp:184-232  expand_box
//   include/corridor_constructor.hpp:234-245  expandSFCFromBox
//   src/multi_sync_simulator.cpp:153-167      setOctomap (maxdist = 1.0, bbx = world box)
// Upstream (source not in the reference tree; restated from the published formats/algorithms,
// SURVEY.md App. C.2-C.3): octomap 1.9 binary ".bt" tree stream, OcTree::coordToKey, and
// dynamicEDT3D's DynamicEDTOctomap(maxdist, tree, bbxMin, bbxMax, false) + getDistance().
// PARITY UNPINNED for the EDT: the upstream brushfire can be inexact at larger distances; this
// restatement computes the exact Euclidean transform clamped at maxDist. The corridor test only
// thresholds at < margin + 0.5*res (sqdist <= 3 cells for the shipped radius) where both agree.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "geom.hpp"

namespace orc {

// ------------------------------------------------------------------------------------------
// .bt reader: returns the finest-resolution (depth 16) keys of all occupied leaves.
// ------------------------------------------------------------------------------------------
struct Key3 { int k[3]; };   // signed key relative to the tree centre (octomap key - 32768)

struct BtTree {
    double res = 0.1;
    std::vector<Key3> occupied;      // finest voxels (coarser leaves expanded)
    size_t n_nodes = 0;
};

inline void bt_recurse(const uint8_t* data, size_t size, size_t& pos, int depth, int cx, int cy,
                       int cz, BtTree& t) {
    // (cx,cy,cz) = min corner of this node in finest-voxel units relative to centre; node edge
    // length = 2^(16-depth) voxels. Each inner node: 2 bytes, child c in bits (2c,2c+1) of the
    // little-endian word: 00 unknown, 01 occupied leaf(bit pattern (0,1) LSB first => value 2),
    // see App. C.3: (1,0) free => value 1, (0,1) occupied => value 2, (1,1) inner => value 3.
    if (pos + 2 > size) throw std::runtime_error("bt: truncated stream");
    unsigned word = data[pos] | (data[pos + 1] << 8);
    pos += 2;
    t.n_nodes++;
    const int half = 1 << (15 - depth);          // child edge length in finest voxels
    unsigned kinds[8];
    for (int c = 0; c < 8; c++) kinds[c] = (word >> (2 * c)) & 3u;
    for (int c = 0; c < 8; c++) {
        if (kinds[c] == 0) continue;
        int ox = cx + ((c & 1) ? half : 0), oy = cy + ((c & 2) ? half : 0), oz = cz + ((c & 4) ? half : 0);
        if (kinds[c] == 3) {
            bt_recurse(data, size, pos, depth + 1, ox, oy, oz, t);
        } else {
            t.n_nodes++;
            if (kinds[c] == 2) {
                for (int x = 0; x < half; x++)
                    for (int y = 0; y < half; y++)
                        for (int z = 0; z < half; z++) t.occupied.push_back(Key3{{ox + x, oy + y, oz + z}});
            }
        }
    }
}

inline BtTree bt_parse(const uint8_t* buf, size_t len) {
    BtTree t;
    size_t pos = 0;
    auto getline = [&](std::string& out) {
        out.clear();
        while (pos < len && buf[pos] != '\n') out.push_back((char)buf[pos++]);
        if (pos < len) pos++;
    };
    std::string line;
    getline(line);
    if (line.rfind("# Octomap OcTree binary file", 0) != 0) throw std::runtime_error("bt: bad magic");
    size_t declared = 0;
    while (pos < len) {
        getline(line);
        if (line.empty() || line[0] == '#') continue;
        if (line.rfind("id ", 0) == 0) { if (line != "id OcTree") throw std::runtime_error("bt: id"); }
        else if (line.rfind("size ", 0) == 0) declared = std::stoul(line.substr(5));
        else if (line.rfind("res ", 0) == 0) t.res = std::stod(line.substr(4));
        else if (line == "data") break;
    }
    if (declared > 0) {
        bt_recurse(buf, len, pos, 0, -32768, -32768, -32768, t);   // root word follows "data\n"
        if (t.n_nodes != declared) throw std::runtime_error("bt: node count mismatch");
    }
    return t;
}

inline BtTree bt_load(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("bt: cannot open " + path);
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    return bt_parse(buf.data(), buf.size());
}

// ------------------------------------------------------------------------------------------
// Distance map (DynamicEDTOctomap semantics)
// ------------------------------------------------------------------------------------------
inline thread_local long long tl_edt_lookups = 0;   // per-thread getDistance() call counter

struct DistMap {
    double res = 0.1;
    int off[3] = {0, 0, 0};     // signed key of cell (0,0,0)
    int size[3] = {0, 0, 0};
    int max_sq = 121;
    std::vector<int> sqdist;    // [x][y][z], clamped at max_sq

    static int coord_to_key(double c, double res) {                 // OcTree::coordToKey, signed
        return (int)std::floor((1.0 / res) * c);
    }
    size_t idx(int x, int y, int z) const { return ((size_t)x * size[1] + y) * size[2] + z; }

    // DynamicEDTOctomap::getDistance: cell distance * resolution (float), -1 outside the map.
    float distance(F3 p) const {
        tl_edt_lookups++;
        int x = coord_to_key((double)p.x, res) - off[0];
        int y = coord_to_key((double)p.y, res) - off[1];
        int z = coord_to_key((double)p.z, res) - off[2];
        if (x >= 0 && x < size[0] && y >= 0 && y < size[1] && z >= 0 && z < size[2]) {
            float cell = (float)std::sqrt((double)sqdist[idx(x, y, z)]);
            return (float)(cell * res);
        }
        return -1.0f;
    }
};

inline DistMap distmap_build(const std::vector<Key3>& occupied, double res, F3 world_min, F3 world_max,
                             float maxdist = 1.0f) {
    DistMap m;
    m.res = res;
    int md = (int)(maxdist / res + 1);          // DynamicEDTOctomap ctor
    m.max_sq = md * md;
    for (int a = 0; a < 3; a++) {
        int lo = DistMap::coord_to_key((double)world_min(a), res);
        int hi = DistMap::coord_to_key((double)world_max(a), res);
        m.off[a] = lo;
        m.size[a] = hi - lo + 1;
    }
    const int sx = m.size[0], sy = m.size[1], sz = m.size[2];
    std::vector<int> a((size_t)sx * sy * sz, m.max_sq), b(a.size());
    std::vector<uint8_t> occ(a.size(), 0);
    for (const Key3& k : occupied) {
        int x = k.k[0] - m.off[0], y = k.k[1] - m.off[1], z = k.k[2] - m.off[2];
        if (x < 0 || x >= sx || y < 0 || y >= sy || z < 0 || z >= sz) continue;
        occ[m.idx(x, y, z)] = 1;
    }
    const int R = md;   // beyond md cells the clamp applies anyway
    // separable exact EDT: pass along z, then y, then x
    for (int x = 0; x < sx; x++) for (int y = 0; y < sy; y++) for (int z = 0; z < sz; z++) {
        int best = m.max_sq;
        for (int dz = -R; dz <= R; dz++) {
            int zz = z + dz;
            if (zz < 0 || zz >= sz) continue;
            if (occ[m.idx(x, y, zz)]) best = std::min(best, dz * dz);
        }
        a[m.idx(x, y, z)] = best;
    }
    for (int x = 0; x < sx; x++) for (int y = 0; y < sy; y++) for (int z = 0; z < sz; z++) {
        int best = m.max_sq;
        for (int dy = -R; dy <= R; dy++) {
            int yy = y + dy;
            if (yy < 0 || yy >= sy) continue;
            best = std::min(best, a[m.idx(x, yy, z)] + dy * dy);
        }
        b[m.idx(x, y, z)] = best;
    }
    for (int x = 0; x < sx; x++) for (int y = 0; y < sy; y++) for (int z = 0; z < sz; z++) {
        int best = m.max_sq;
        for (int dx = -R; dx <= R; dx++) {
            int xx = x + dx;
            if (xx < 0 || xx >= sx) continue;
            best = std::min(best, b[m.idx(xx, y, z)] + dx * dx);
        }
        a[m.idx(x, y, z)] = std::min(best, m.max_sq);
    }
    m.sqdist.swap(a);
    return m;
}

// ------------------------------------------------------------------------------------------
// CorridorConstructor
// ------------------------------------------------------------------------------------------
struct Corridor {
    const DistMap* dm;
    F3 world_min, world_max;
    double res;
    static constexpr double EPS = 1e-9;          // SP_EPSILON
    static constexpr float EPSF = 1e-5f;          // SP_EPSILON_FLOAT (double literal narrowed on use)

    bool obstacle_in_box(const double* box, double margin) const {               // :81-122
        int bs[3];
        for (int i = 0; i < 3; i++) bs[i] = (int)std::round((box[i + 3] - box[i]) / res) + 1;
        int it[3];
        for (it[0] = 0; it[0] < std::max(bs[0], 2); it[0]++)
            for (it[1] = 0; it[1] < std::max(bs[1], 2); it[1]++)
                for (it[2] = 0; it[2] < std::max(bs[2], 2); it[2]++) {
                    float sp[3], delta[3];
                    for (int i = 0; i < 3; i++) {
                        if (bs[i] == 1 && it[i] > 0) sp[i] = (float)box[i];
                        else sp[i] = (float)(box[i] + it[i] * res);
                    }
                    for (int i = 0; i < 3; i++) {
                        if (it[i] == 0 && box[i] > (double)world_min(i) + 1e-5) delta[i] = (float)(-1e-5);
                        else delta[i] = (float)1e-5;
                    }
                    F3 p = f3(sp[0], sp[1], sp[2]) + f3(delta[0], delta[1], delta[2]);
                    float dist = dm->distance(p);
                    if ((double)dist < margin + 0.5 * res - 1e-5) return true;    // :114
                }
        return false;
    }

    bool box_in_boundary(const double* box, double margin) const {               // :124-131
        return box[0] > (double)world_min.x + margin - EPS && box[1] > (double)world_min.y + margin - EPS &&
               box[2] > (double)world_min.z + margin - EPS && box[3] < (double)world_max.x - margin + EPS &&
               box[4] < (double)world_max.y - margin + EPS && box[5] < (double)world_max.z - margin + EPS;
    }

    void axis_cand(const double* box, F3 goal, std::vector<int>& cand) const {    // :142-182
        F3 mid = f3((float)(0.5 * (box[0] + box[3])), (float)(0.5 * (box[1] + box[4])),
                    (float)(0.5 * (box[2] + box[5])));
        F3 delta = goal - mid;
        int offsets[3] = {delta.x > 0 ? 3 : 0, delta.y > 0 ? 3 : 0, delta.z > 0 ? 3 : 0};
        double values[3] = {std::fabs((double)delta.x), std::fabs((double)delta.y), std::fabs((double)delta.z)};
        std::vector<int> order;
        double max_value = -1, min_value = 1e+9;
        for (int i = 0; i < 3; i++) {
            if (values[i] > max_value) { order.insert(order.begin(), i); max_value = values[i]; }
            else if (values[i] < min_value) { order.push_back(i); min_value = values[i]; }
            else order.insert(order.begin() + 1, i);
        }
        cand.assign(6, 0);
        for (int i = 0; i < 3; i++) {
            cand[i] = order[i] + offsets[order[i]];
            cand[5 - i] = order[i] + (3 - offsets[order[i]]);
        }
    }

    void expand_box(const double* initial, F3 goal, double margin, double* out) const {   // :184-232
        std::vector<int> cand;
        axis_cand(initial, goal, cand);
        double box[6], bc[6], bu[6];
        std::memcpy(box, initial, sizeof box);
        int i = -1;
        while (!cand.empty()) {
            std::memcpy(bc, box, sizeof box);
            std::memcpy(bu, box, sizeof box);
            while (!obstacle_in_box(bu, margin) && box_in_boundary(bu, 0)) {
                i++;
                if (i >= (int)cand.size()) i = 0;
                int axis = cand[i];
                std::memcpy(box, bc, sizeof box);
                std::memcpy(bu, bc, sizeof box);
                if (axis < 3) {
                    bu[axis + 3] = bc[axis];
                    bc[axis] = bc[axis] - res;
                    bu[axis] = bc[axis];
                } else {
                    bu[axis - 3] = bc[axis];
                    bc[axis] = bc[axis] + res;
                    bu[axis] = bc[axis];
                }
            }
            // NOTE (reference quirk, :223): when the loop above exits on its very first test
            // (i == -1) the reference erases begin()+(-1), which is undefined behaviour; that can
            // only happen if the seed box is blocked, which expandBoxFromPoint rules out (:35-38).
            if (i < 0) i = 0;
            cand.erase(cand.begin() + i);
            if (i > 0) i--; else i = (int)cand.size() - 1;
        }
        std::memcpy(out, box, sizeof box);
    }

    // returns false if the seed box is blocked (reference throws std::invalid_argument, :35-38)
    bool expand_from_point(F3 point, F3 goal, double radius, F3& bmin, F3& bmax) const {   // :18-44
        double ib[6];
        for (int i = 0; i < 3; i++) {
            double rp = std::round((double)point(i) / res) * res;
            if (std::fabs((double)point(i) - rp) < 0.01) { ib[i] = rp; ib[i + 3] = rp; }
            else {
                ib[i] = std::floor((double)point(i) / res) * res;
                ib[i + 3] = std::ceil((double)point(i) / res) * res;
            }
        }
        if (obstacle_in_box(ib, radius)) return false;
        double eb[6];
        expand_box(ib, goal, radius, eb);                                                   // :234-245
        bmin = f3((float)eb[0], (float)eb[1], (float)eb[2]);
        bmax = f3((float)eb[3], (float)eb[4], (float)eb[5]);
        return true;
    }
};

}  // namespace orc
