// Reader for octomap's binary tree format (".bt", `OcTree::writeBinary`), host side.
// Replaces octomap::OcTree::readBinary as used by MultiSyncSimulator::setOctomap
// (src/multi_sync_simulator.cpp:153-167); octomap itself is an external dependency of the reference.
//
// Stream layout: text header ("# Octomap OcTree binary file", "id OcTree", "size <nodes>", "res <m>", "data"), then
// the tree in depth-first order, 16 bits per inner node: two bits per child (child index = x | y<<1 | z<<2),
// 00 = unknown, 01 = free leaf... as little-endian bit pairs: value 1 free leaf, 2 occupied leaf, 3 inner node whose
// own 16-bit record follows in child order. Depth 16; key 32768 <-> coordinate 0.
// Output: the signed finest-level keys (key - 32768) of every occupied leaf, coarser leaves expanded.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace lscgpu {

struct OccupiedVoxels {
    double res = 0.1;
    std::vector<int32_t> keys;     // xyz triples
    size_t n_nodes = 0;
};

inline bool read_bt_buffer(const uint8_t* buf, size_t len, OccupiedVoxels& out, std::string& err) {
    size_t pos = 0;
    auto next_line = [&](std::string& line) {
        line.clear();
        while (pos < len && buf[pos] != '\n') line.push_back((char)buf[pos++]);
        if (pos < len) pos++;
        if (!line.empty() && line.back() == '\r') line.pop_back();
    };
    std::string line;
    next_line(line);
    if (line.compare(0, 28, "# Octomap OcTree binary file") != 0) { err = "not an octomap binary (.bt) file"; return false; }
    size_t declared = 0;
    bool have_data = false;
    while (pos < len) {
        next_line(line);
        if (line.empty() || line[0] == '#') continue;
        if (line.compare(0, 3, "id ") == 0) {
            if (line != "id OcTree") { err = "unsupported tree id: " + line; return false; }
        } else if (line.compare(0, 5, "size ") == 0) declared = (size_t)std::strtoull(line.c_str() + 5, nullptr, 10);
        else if (line.compare(0, 4, "res ") == 0) out.res = std::strtod(line.c_str() + 4, nullptr);
        else if (line == "data") { have_data = true; break; }
    }
    if (!have_data) { err = "missing data section"; return false; }
    out.keys.clear();
    out.n_nodes = 0;
    if (declared == 0) return true;

    struct Frame { uint16_t word; int child; int depth; int32_t org[3]; };
    std::vector<Frame> stack;
    auto read_word = [&](uint16_t& w) -> bool {
        if (pos + 2 > len) return false;
        w = (uint16_t)(buf[pos] | (buf[pos + 1] << 8));
        pos += 2;
        return true;
    };
    Frame root{};
    root.child = 0; root.depth = 0;
    root.org[0] = root.org[1] = root.org[2] = -32768;
    if (!read_word(root.word)) { err = "truncated tree stream"; return false; }
    out.n_nodes = 1;
    stack.push_back(root);
    while (!stack.empty()) {
        Frame& f = stack.back();
        if (f.child == 8) { stack.pop_back(); continue; }
        const int c = f.child++;
        const unsigned kind = (f.word >> (2 * c)) & 3u;
        if (kind == 0) continue;
        const int32_t edge = 1 << (15 - f.depth);       // child edge in finest voxels
        const int32_t ox = f.org[0] + ((c & 1) ? edge : 0);
        const int32_t oy = f.org[1] + ((c & 2) ? edge : 0);
        const int32_t oz = f.org[2] + ((c & 4) ? edge : 0);
        out.n_nodes++;
        if (kind == 3) {
            if (f.depth + 1 >= 16) { err = "inner node below the finest level"; return false; }
            Frame ch{};
            ch.child = 0; ch.depth = f.depth + 1;
            ch.org[0] = ox; ch.org[1] = oy; ch.org[2] = oz;
            if (!read_word(ch.word)) { err = "truncated tree stream"; return false; }
            stack.push_back(ch);        // invalidates f; not used afterwards
        } else if (kind == 2) {
            if ((size_t)edge * edge * edge > (size_t)1 << 24) { err = "occupied leaf too coarse to expand"; return false; }
            for (int32_t x = 0; x < edge; x++)
                for (int32_t y = 0; y < edge; y++)
                    for (int32_t z = 0; z < edge; z++) {
                        out.keys.push_back(ox + x); out.keys.push_back(oy + y); out.keys.push_back(oz + z);
                    }
        }
    }
    if (out.n_nodes != declared) { err = "node count does not match the header"; return false; }
    return true;
}

inline bool read_bt_file(const char* path, OccupiedVoxels& out, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::vector<uint8_t> buf;
    uint8_t chunk[1 << 16];
    size_t n;
    while ((n = std::fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
    std::fclose(f);
    return read_bt_buffer(buf.data(), buf.size(), out, err);
}

}  // namespace lscgpu
