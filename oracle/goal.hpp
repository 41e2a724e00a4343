// ORACLE — test infrastructure only (see oracle/README.md). CPU restatement of the reference's goal planning, the
// step before the hot path. Citations relative to /root/reference:
//   src/traj_planner.cpp:540-608            goalPlanningWithPriority
//   src/grid_based_planner.cpp:53-68        plan
//   src/grid_based_planner.cpp:70-90        updateGridInfo
//   src/grid_based_planner.cpp:92-195       updateGridMap (distance field + higher-priority agents)
//   src/grid_based_planner.cpp:197-245     updateGridMission (start cell repair)
//   src/grid_based_planner.cpp:350-407      findLOSFreeGoal
//   src/grid_based_planner.cpp:409-434      castRay
//   src/Astar-3D/isearch.cpp:48-105         startSearch (goal test ignores the altitude, :74)
//   src/Astar-3D/isearch.cpp:107-144        findSuccessors (6-connected: environmentoptions.cpp:13-21)
//   src/Astar-3D/isearch.cpp:177-284        findMin / deleteMin / addOpen (per-row open lists, g-max tie break)
//   src/Astar-3D/astar.cpp:18-30            Euclidean heuristic
// The open list of the reference is one std::unordered_map per grid row; when a row's minimum is re-scanned after a
// pop, ties in (F, g) go to the LAST node in the container's iteration order. That order is implementation defined
// (libstdc++ here); this restatement keeps the same container and the same sequence of insertions and erasures, so
// it reproduces it. tests/test_goal_planning.py pins the search against the reference's own Astar-3D sources compiled
// into oracle/_ref.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <unordered_map>
#include <vector>

#include "edt.hpp"
#include "geom.hpp"

namespace orc {

struct GoalParams {
    double grid_resolution = 0.25, grid_margin = 0.1;     // launch/simulation.launch:87-88
    double goal_threshold = 0.1, goal_radius = 2.0, priority_dist_threshold = 0.4;   // :91-93
    double world_resolution = 0.1;
};

struct AstarNode { int i, j, z; double F, g; int parent; };     // parent: index into the closed vector, -1 = none

// grid[(i * dim1 + j) * dim2 + z] != 0 : occupied. Returns the cell path start..goal (empty: none).
inline std::vector<std::array<int, 3>> astar_search(const std::vector<uint8_t>& grid, const int dim[3], const int start[3],
                                                    const int goal[3], long long* expansions = nullptr) {
    const int H = dim[0], W = dim[1], A = dim[2];
    auto occ = [&](int i, int j, int z) { return grid[((size_t)i * W + j) * A + z] != 0; };
    auto key_of = [&](int i, int j, int z) { return (uint32_t)H * W * z + W * i + j; };       // Node::get_id
    auto close_key = [&](int i, int j, int z) { return (long long)i * W + j + (long long)H * W * z; };
    auto heur = [&](int i, int j, int z) {
        return 1.0 * std::sqrt((double)((goal[0] - i) * (goal[0] - i) + (goal[1] - j) * (goal[1] - j) + (goal[2] - z) * (goal[2] - z)));
    };
    std::vector<std::unordered_map<uint32_t, AstarNode>> open(H);
    std::vector<long long> open_min(H, -1);
    std::unordered_map<long long, int> closed;          // cell -> index into `done`
    std::vector<AstarNode> done;
    int open_size = 0;

    auto add_open = [&](const AstarNode& nn, uint32_t key) {                               // isearch.cpp:244-284
        bool inserted = false;
        auto& row = open[nn.i];
        auto it = row.find(key);
        if (it != row.end()) {
            if (nn.F < it->second.F) { it->second = nn; inserted = true; }
        } else {
            row[key] = nn; inserted = true; ++open_size;
        }
        if (row.size() == 1) open_min[nn.i] = key;
        else {
            const AstarNode mn = row[(uint32_t)open_min[nn.i]];
            if (inserted && nn.F <= mn.F) {
                if (nn.F == mn.F) { if (nn.g >= mn.g) open_min[nn.i] = key; }
                else open_min[nn.i] = key;
            }
        }
    };
    AstarNode cur{start[0], start[1], start[2], 0.0, 0.0, -1};
    cur.F = 1.0 * heur(cur.i, cur.j, cur.z);
    add_open(cur, key_of(cur.i, cur.j, cur.z));
    open_size = 1;                                                                          // isearch.cpp:67
    bool found = false;
    int cur_idx = -1;
    while (open_size != 0) {
        // findMin (:177-207): rows in ascending order; on equal F the later row wins when its g is not smaller
        AstarNode mn{}; mn.F = std::numeric_limits<double>::infinity(); mn.g = 0;
        bool first = true;
        for (int i = 0; i < H; i++) {
            if (open[i].empty()) continue;
            const AstarNode c = open[i][(uint32_t)open_min[i]];
            if (c.F <= mn.F) {
                if (c.F == mn.F && !first) { if (c.g >= mn.g) mn = c; }
                else mn = c;
                first = false;
            }
        }
        cur = mn;
        done.push_back(cur); cur_idx = (int)done.size() - 1;
        closed.insert({close_key(cur.i, cur.j, cur.z), cur_idx});
        // deleteMin (:209-242)
        {
            auto& row = open[cur.i];
            row.erase(key_of(cur.i, cur.j, cur.z));
            AstarNode best{}; best.F = (double)std::numeric_limits<float>::infinity(); best.g = 0;
            bool any = false;
            for (auto it = row.begin(); it != row.end(); ++it) {
                if (it->second.F <= best.F) {
                    if (it->second.F == best.F && any) { if (it->second.g >= best.g) { open_min[cur.i] = it->first; best = it->second; } }
                    else { open_min[cur.i] = it->first; best = it->second; }
                    any = true;
                }
            }
        }
        --open_size;
        if (cur.i == goal[0] && cur.j == goal[1]) { found = true; break; }
        static const int mv[6][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}};   // loop order of :110-112
        for (int s = 0; s < 6; s++) {
            const int ni = cur.i + mv[s][0], nj = cur.j + mv[s][1], nz = cur.z + mv[s][2];
            if (ni < 0 || ni >= H || nj < 0 || nj >= W || nz < 0 || nz > A - 1) continue;
            if (occ(ni, nj, nz)) continue;
            if (closed.find(close_key(ni, nj, nz)) != closed.end()) continue;
            AstarNode nn{ni, nj, nz, 0.0, cur.g + 1.0, cur_idx};
            nn.F = nn.g + 1.0 * heur(ni, nj, nz);
            add_open(nn, key_of(ni, nj, nz));
        }
    }
    if (expansions) *expansions = (long long)done.size();
    std::vector<std::array<int, 3>> path;
    if (!found) return path;
    for (int k = cur_idx; k >= 0; k = done[k].parent) path.push_back({done[k].i, done[k].j, done[k].z});
    std::vector<std::array<int, 3>> fwd(path.rbegin(), path.rend());
    return fwd;
}

// the uninitialised `min.g` of findMin / deleteMin is never read on the first hit (F < inf), which `first` / `any`
// make explicit above.

struct GridPlanner {
    const DistMap* dm;               // null: no octomap
    GoalParams gp;
    F3 world_min, world_max;
    double grid_min[3], grid_max[3];
    int dim[3];
    std::vector<uint8_t> grid;
    std::vector<F3> path;            // plan_result.path
    long long expansions = 0;

    size_t cidx(int i, int j, int k) const { return ((size_t)i * dim[1] + j) * dim[2] + k; }
    F3 cell_point(int i, int j, int k) const {                                              // gridVectorToPoint3D :305-310
        return f3((float)(grid_min[0] + i * gp.grid_resolution), (float)(grid_min[1] + j * gp.grid_resolution),
                  (float)(grid_min[2] + k * gp.grid_resolution));
    }
    void update_info() {                                                                    // :70-90
        const double r = gp.grid_resolution;
        for (int i = 0; i < 3; i++) {
            grid_min[i] = -std::floor((-(double)world_min(i) + 1e-9) / r) * r;
            grid_max[i] = std::floor(((double)world_max(i) + 1e-9) / r) * r;
        }
        for (int i = 0; i < 3; i++) dim[i] = (int)std::round((grid_max[i] - grid_min[i]) / r) + 1;
    }
    // obstacles: every other agent j: position, radius, downwash; `high[j]` = higher priority
    void update_map(double radius, double downwash, int n, int self, const F3* pos, const AgentConst* ac,
                    const std::vector<char>* high) {                                        // :92-195
        grid.assign((size_t)dim[0] * dim[1] * dim[2], 0);
        if (dm) {
            const float margin = (float)gp.grid_margin;
            for (int i = 0; i < dim[0]; i++)
                for (int j = 0; j < dim[1]; j++)
                    for (int k = 0; k < dim[2]; k++) {
                        const float dist = dm->distance(cell_point(i, j, k));
                        if (dist < radius + margin) grid[cidx(i, j, k)] = 1;
                    }
        }
        const double r = gp.grid_resolution;
        for (int j = 0; j < n; j++) {
            if (j == self) continue;
            if (!high || !(*high)[j]) continue;
            const double ox = (double)pos[j].x, oy = (double)pos[j].y, oz = (double)pos[j].z;
            const int oi = (int)std::round((ox - grid_min[0] + 1e-9) / r);
            const int oj = (int)std::round((oy - grid_min[1] + 1e-9) / r);
            const int ok = (int)std::round((oz - grid_min[2] + 1e-9) / r);
            const int size_xy = (int)std::ceil((radius + ac[j].radius) / r);
            const int size_z = (int)std::ceil((radius * downwash + ac[j].radius * ac[j].downwash) / r);
            const double dw = (radius * downwash + ac[j].radius * ac[j].downwash) / (radius + ac[j].radius);
            for (int a = std::max(oi - size_xy, 0); a <= std::min(oi + size_xy, dim[0] - 1); a++)
                for (int b = std::max(oj - size_xy, 0); b <= std::min(oj + size_xy, dim[1] - 1); b++)
                    for (int c = std::max(ok - size_z, 0); c <= std::min(ok + size_z, dim[2] - 1); c++) {
                        const F3 p = cell_point(a, b, c);
                        const double dist = std::sqrt(std::pow((double)p.x - ox, 2) + std::pow((double)p.y - oy, 2) +
                                                      std::pow(((double)p.z - oz) / dw, 2));
                        if (dist < radius + ac[j].radius) grid[cidx(a, b, c)] = 1;
                    }
        }
    }
    void to_cell(F3 p, int out[3]) const {                                                  // point3DToGridVector :329-334
        out[0] = (int)std::round(((double)p.x - grid_min[0]) / gp.grid_resolution);
        out[1] = (int)std::round(((double)p.y - grid_min[1]) / gp.grid_resolution);
        out[2] = (int)std::round(((double)p.z - grid_min[2]) / gp.grid_resolution);
    }
    bool is_occupied(const int c[3]) const {                                                // :257-264
        for (int i = 0; i < 3; i++) if (c[i] < 0 || c[i] > dim[i] - 1) return true;
        return grid[cidx(c[0], c[1], c[2])] != 0;
    }
    // returns false when the start / goal cell lies outside the grid (the reference would index out of bounds)
    bool update_mission(F3 current, F3 goal, int start[3], int goal_c[3]) {                 // :197-245
        to_cell(current, start); to_cell(goal, goal_c);
        for (int i = 0; i < 3; i++) if (start[i] < 0 || start[i] >= dim[i] || goal_c[i] < 0 || goal_c[i] >= dim[i]) return false;
        if (grid[cidx(start[0], start[1], start[2])] != 0) {
            int min_dist = 1000000000, best[3] = {start[0], start[1], start[2]};
            for (int i = -2; i < 3; i++)
                for (int j = -2; j < 3; j++)
                    for (int k = -1; k < 2; k++) {
                        const int c[3] = {start[0] + i, start[1] + j, start[2] + k};
                        if (!is_occupied(c)) {
                            const int d = std::abs(i) + std::abs(j) + std::abs(k);
                            if (d < min_dist) { min_dist = d; best[0] = c[0]; best[1] = c[1]; best[2] = c[2]; }
                        }
                    }
            start[0] = best[0]; start[1] = best[1]; start[2] = best[2];
            if (grid[cidx(start[0], start[1], start[2])] != 0) grid[cidx(start[0], start[1], start[2])] = 0;
        }
        return true;
    }
    // plan (:53-68): fills `path` (world points of the cell path; empty when none)
    void plan(F3 current, F3 goal, double radius, double downwash, int n, int self, const F3* pos, const AgentConst* ac,
              const std::vector<char>* high) {
        update_info();
        update_map(radius, downwash, n, self, pos, ac, high);
        int s[3], g[3];
        path.clear();
        if (!update_mission(current, goal, s, g)) return;
        long long ex = 0;
        const auto cells = astar_search(grid, dim, s, g, &ex);
        expansions += ex;
        for (const auto& c : cells) path.push_back(cell_point(c[0], c[1], c[2]));
    }
    bool cast_ray(F3 a, F3 b, double radius) const {                                        // :409-434
        const double dist = normf(a - b);
        const double thr = std::sqrt(0.25 * dist * dist + radius * radius);
        const double sa = (double)dm->distance(a), sb = (double)dm->distance(b);
        if (sa < radius + 0.5 * gp.world_resolution - 1e-5) return false;
        if (sb < radius + 0.5 * gp.world_resolution - 1e-5) return false;
        if (thr < 1.0 && sa > thr && sb > thr) return true;
        const F3 mid = (a + b) * 0.5f;
        return cast_ray(a, mid, radius) && cast_ray(mid, b, radius);
    }
    F3 los_free_goal(F3 current, F3 goal, double radius) const {                            // :350-407
        F3 los = current;
        std::vector<F3> pts = path;
        pts.push_back(goal);
        for (int i = 0; i < 6; i++) {
            const double ratio = 1.5 - 0.1 * i;
            for (const F3& p : pts) {
                bool safe = true;
                if (dm) safe = cast_ray(current, p, radius * ratio);
                if (safe) los = p; else break;
            }
            if (normf(los - current) > 0.3) break;
        }
        const F3 delta = los - current;
        if (normf(delta) > gp.goal_radius) los = current + normalizedf(delta) * (float)gp.goal_radius;
        return los;
    }
};

struct GoalResult { F3 goal; int mode; int n_high; long long expansions; };   // mode: 0 A*+LOS, 1 retreat

// goalPlanningWithPriority (src/traj_planner.cpp:540-608) for agent `a` of a swarm whose members all run the same planner:
//   pos[j] current positions (= obstacle.pose of the simulator's update()), desired[j] desired goals,
//   prev_traj[j*30 ..] traj_curr of every agent (obs_prev_trajs), init_end = initial_traj[M-1][n] of agent a.
inline GoalResult goal_planning_priority(int a, int n, const F3* pos, const F3* desired, const F3* prev_traj, F3 init_end,
                                         const AgentConst* ac, const DistMap* dm, const GoalParams& gp, F3 world_min,
                                         F3 world_max, const char* in_slack_set = nullptr) {
    GoalResult out{};
    std::vector<char> high(n, 0);
    int closest = -1;
    const double dist_to_goal = normf(pos[a] - desired[a]);
    double min_dist_to_obs = 1e9;
    for (int j = 0; j < n; j++) {
        if (j == a) continue;
        if (in_slack_set && in_slack_set[j]) { high[j] = 1; out.n_high++; continue; }       // traj_planner.cpp:548-551
        const double obs_dist_to_goal = normf(pos[j] - desired[j]);
        const double dist_to_obs = normf(pos[j] - pos[a]);
        if (obs_dist_to_goal < gp.goal_threshold) continue;
        const F3* t = prev_traj + (size_t)j * 30;
        if (dist_to_goal > gp.goal_threshold && dotf(t[29] - t[5], t[5] - pos[a]) > 0) continue;
        if (dist_to_goal < gp.goal_threshold || obs_dist_to_goal < dist_to_goal) {
            if (dist_to_obs < min_dist_to_obs) { min_dist_to_obs = dist_to_obs; closest = j; }
            high[j] = 1; out.n_high++;
        }
    }
    const double dist_keep = gp.priority_dist_threshold + 0.1;
    if (min_dist_to_obs < gp.priority_dist_threshold) {
        out.goal = pos[a] - normalizedf(pos[closest] - pos[a]) * (float)dist_keep;
        out.mode = 1;
        return out;
    }
    GridPlanner g{dm, gp, world_min, world_max};
    g.plan(pos[a], desired[a], ac[a].radius, ac[a].downwash, n, a, pos, ac, &high);
    if (g.path.empty()) g.plan(pos[a], desired[a], ac[a].radius, ac[a].downwash, n, a, pos, ac, nullptr);
    out.goal = g.los_free_goal(init_end, desired[a], ac[a].radius);
    out.mode = 0;
    out.expansions = g.expansions;
    return out;
}

}  // namespace orc
