// C entry points of the host-side goal planner (grid_based_planner.hpp) for the ctypes parity tests
// (tests/test_host_goal.py compares them with the oracle and with the reference's own A* in oracle/_ref).
#include <cstring>

#include "grid_based_planner.hpp"
#include "../csrc/astar_core.cuh"

using namespace DynamicPlanning;

extern "C" int devcore_bucket_sequence(int row_capacity, int* seq, int max_levels);

template <typename I>
static int devcore_astar_t(const int* dim, const unsigned char* grid, const int* start, const int* goal, int* path_out, int max_len,
                           long long* expansions, bool accel) {
    using namespace lscgpu;
    const int H = dim[0], W = dim[1], A = dim[2];
    const size_t cells = (size_t)H * W * A;
    AstarCtx<I> c{};
    c.H = H; c.W = W; c.A = A;
    int seq[kAstarMaxLevels] = {0};
    const int levels = devcore_bucket_sequence(W * A, seq, kAstarMaxLevels);
    c.bkt_seq = seq;
    c.bcap = seq[levels - 1];
    const size_t max_cells = sizeof(I) == 2 ? (size_t)65533 : ((size_t)1 << 30);
    if ((c.bcap) < W * A || cells > max_cells) return -1;
    std::vector<uint8_t> cell(cells);
    for (size_t k = 0; k < cells; k++) cell[k] = grid[k] ? kCellOccupied : 0;
    std::vector<I> g(cells, (I)77), next(cells, (I)55), bkt((size_t)H * c.bcap, (I)12345);
    std::vector<int> head(H, -1), count(H, 0), level(H, 0), min_cell(H, -1), min_g(H, 0);
    std::vector<double> min_f(H, 0.0);
    c.cell = cell.data(); c.g = g.data(); c.next = next.data(); c.bkt = bkt.data();
    std::vector<I> bstamp((size_t)H * c.bcap, (I)999);
    std::vector<int> stamp(H, 0);
    c.bstamp = bstamp.data(); c.stamp = stamp.data();
    std::vector<int> jlo(H, W), jhi(H, -1);
    c.jlo = jlo.data(); c.jhi = jhi.data();
    c.head = head.data(); c.count = count.data(); c.level = level.data(); c.min_cell = min_cell.data(); c.min_g = min_g.data();
    c.min_f = min_f.data();
    c.gi = goal[0]; c.gj = goal[1]; c.gz = goal[2];
    // the accelerators of the shared-memory instantiation: multiply-high divisions and the sqrt table
    unsigned bkt_magic[kAstarMaxLevels] = {0};
    std::vector<double> sqrt_tab;
    if (accel) {
        c.magic_a = astar_magic(A, cells); c.magic_w = astar_magic(W, cells);
        for (int k = 0; k < kAstarMaxLevels; k++) bkt_magic[k] = k < levels ? astar_magic(seq[k], cells) : 0;
        c.bkt_magic = bkt_magic;
        sqrt_tab.resize((size_t)(H - 1) * (H - 1) + (size_t)(W - 1) * (W - 1) + (size_t)(A - 1) * (A - 1) + 1);
        for (size_t k = 0; k < sqrt_tab.size(); k++) sqrt_tab[k] = std::sqrt((double)k);
        c.sqrt_tab = sqrt_tab.data();
    }
    astar_begin(c, start[0], start[1], start[2]);
    int cur = -1, found = 0;
    long long stamp_mismatch = 0;
    while (c.open_size != 0) {
        cur = astar_find_min(c);
        // the expansion in its three pieces; the list-walk re-scan is the statement of the reference, the stamp-based one
        // (what the kernel's warp does) must name the same node every time
        int ci, cj, cz;
        const int cur_g = astar_close(c, cur, ci, cj, cz);
        astar_rescan(c, ci);
        if (count[ci] > 0 && astar_rescan_by_stamp(c, ci) != min_cell[ci]) stamp_mismatch++;
        if (astar_open_neighbours(c, ci, cj, cz, cur_g)) { found = 1; break; }
    }
    if (stamp_mismatch) return -2;
    if (expansions) *expansions = c.expansions;
    if (!found) return 0;
    const int n = (int)g[cur] + 1;
    for (int k = n - 1, p = cur; k >= 0 && p >= 0; k--, p = astar_parent(c, p))
        if (k < max_len) { path_out[3 * k] = p / (W * A); path_out[3 * k + 1] = (p / A) % W; path_out[3 * k + 2] = p % A; }
    return n;
}

extern "C" {

// HashOrderModel against the real std::unordered_map<uint32_t, int> of this toolchain: random insertions / erasures of
// keys below `key_range`, the full iteration order compared after every operation. Returns the number of mismatching
// operations (0 = the model reproduces the container's node order).
int host_hash_order_check(unsigned seed, int key_range, int n_ops) {
    std::unordered_map<uint32_t, int> ref;
    HashOrderModel model;
    model.reset();
    std::vector<int> next(key_range, -1);
    std::vector<char> present(key_range, 0);
    auto key_of = [](int node) { return (uint32_t)(node * 7919u + 13u); };          // any injective key
    unsigned long long state = seed * 6364136223846793005ull + 1442695040888963407ull;
    int bad = 0;
    for (int op = 0; op < n_ops; op++) {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        const int node = (int)((state >> 33) % (unsigned)key_range);
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        const bool erase = present[node] && ((state >> 40) & 3u) != 0;              // erase 3 of 4 times when present
        if (erase) { ref.erase(key_of(node)); model.erase(node, next, key_of); present[node] = 0; }
        else if (!present[node]) { ref[key_of(node)] = node; model.insert(node, next, key_of); present[node] = 1; }
        else continue;
        int c = model.begin();
        bool same = ref.size() == model.size();
        for (auto it = ref.begin(); same && it != ref.end(); ++it) { same = c >= 0 && it->second == c; if (same) c = next[c]; }
        if (same && c >= 0) same = false;
        bad += !same;
    }
    return bad;
}

int host_astar(const int* dim, const unsigned char* grid, const int* start, const int* goal, int* path_out, int max_len,
               long long* expansions) {
    std::vector<uint8_t> g(grid, grid + (size_t)dim[0] * dim[1] * dim[2]);
    AstarExact a;
    const std::vector<GridCell> path = a.plan(g, dim, {start[0], start[1], start[2]}, {goal[0], goal[1], goal[2]});
    if (expansions) *expansions = a.expansions;
    const int n = (int)path.size();
    for (int k = 0; k < n && k < max_len; k++) { path_out[3 * k] = path[k][0]; path_out[3 * k + 1] = path[k][1]; path_out[3 * k + 2] = path[k][2]; }
    return n;
}

// one agent's goalPlanningWithPriority on explicit inputs; sqdist == null: no octomap
int host_goal_plan(int a, int n, const float* pos, const float* desired, const float* prev_traj /*[n][30][3]*/,
                   const float* init_end, const double* radius, const double* downwash, const unsigned char* sqdist,
                   const int* map_size, const int* map_off, double world_res, const float* wmin, const float* wmax,
                   double grid_resolution, double grid_margin, double goal_threshold, double goal_radius,
                   double priority_dist_threshold, float* goal_out, long long* expansions) {
    Param prm = Param::simulationLaunch();
    prm.world_resolution = world_res; prm.grid_resolution = grid_resolution; prm.grid_margin = grid_margin;
    prm.goal_threshold = goal_threshold; prm.goal_radius = goal_radius; prm.priority_dist_threshold = priority_dist_threshold;
    Mission ms;
    ms.world_min = point3d(wmin[0], wmin[1], wmin[2]); ms.world_max = point3d(wmax[0], wmax[1], wmax[2]);
    HostDistMap dm;
    if (sqdist) {
        dm.res = world_res;
        for (int k = 0; k < 3; k++) { dm.size[k] = map_size[k]; dm.off[k] = map_off[k]; }
        dm.sqdist.assign(sqdist, sqdist + (size_t)map_size[0] * map_size[1] * map_size[2]);
    }
    auto P = [](const float* v) { return point3d(v[0], v[1], v[2]); };
    std::vector<GoalObstacle> obs;
    for (int j = 0; j < n; j++) {
        if (j == a) continue;
        GoalObstacle o;
        o.id = j; o.position = P(pos + 3 * j); o.goal_point = P(desired + 3 * j); o.radius = radius[j]; o.downwash = downwash[j];
        o.prev_traj_first_end = P(prev_traj + ((size_t)j * 30 + 5) * 3);
        o.prev_traj_last_end = P(prev_traj + ((size_t)j * 30 + 29) * 3);
        obs.push_back(o);
    }
    const GoalPlanResult r = goalPlanningWithPriority(P(pos + 3 * a), P(desired + 3 * a), P(init_end), radius[a], downwash[a], obs,
                                                      sqdist ? &dm : nullptr, ms, prm);
    goal_out[0] = r.goal.x(); goal_out[1] = r.goal.y(); goal_out[2] = r.goal.z();
    if (expansions) *expansions = r.expansions;
    return r.kind;
}

// The search core the GPU kernel runs (csrc/astar_core.cuh, compiled here as plain C++) on the same inputs as host_astar:
// scratch arrays as the engine lays them out, bucket sequence recorded from libstdc++'s policy like the engine does.
int devcore_bucket_sequence(int row_capacity, int* seq, int max_levels) {
    std::__detail::_Prime_rehash_policy pol;
    size_t bkt = 1, cnt = 0;
    int n = 0;
    seq[n++] = 1;
    while ((int)bkt < row_capacity && n < max_levels) {
        const auto need = pol._M_need_rehash(bkt, cnt, 1);
        if (need.first) { bkt = need.second; seq[n++] = (int)bkt; }
        cnt++;
    }
    return n;
}

// bucket counts a real std::unordered_map<uint32_t, int> of this toolchain moves through while n keys are inserted one by one
int host_unordered_map_bucket_counts(int n, int* out, int max_out) {
    std::unordered_map<uint32_t, int> m;
    int k = 0;
    size_t last = m.bucket_count();
    if (k < max_out) out[k++] = (int)last;
    for (int i = 0; i < n; i++) {
        m.emplace((uint32_t)i * 2654435761u, i);
        if (m.bucket_count() != last) { last = m.bucket_count(); if (k < max_out) out[k++] = (int)last; }
    }
    return k;
}

// index_bits 16: the compact layout the kernel keeps in shared memory (grids below 65 534 cells); 32: the global-memory layout
int devcore_astar(const int* dim, const unsigned char* grid, const int* start, const int* goal, int* path_out, int max_len,
                  long long* expansions, int index_bits) {
    return index_bits == 16 ? devcore_astar_t<uint16_t>(dim, grid, start, goal, path_out, max_len, expansions, true)
                            : devcore_astar_t<int>(dim, grid, start, goal, path_out, max_len, expansions, false);
}

}  // extern "C"
