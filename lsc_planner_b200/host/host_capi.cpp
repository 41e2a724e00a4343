// C entry points of the host-side goal planner (grid_based_planner.hpp) for the ctypes parity tests
// (tests/test_host_goal.py compares them with the oracle and with the reference's own A* in oracle/_ref).
#include <cstring>

#include "grid_based_planner.hpp"

using namespace DynamicPlanning;

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

}  // extern "C"
