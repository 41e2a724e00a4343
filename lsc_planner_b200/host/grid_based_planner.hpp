// Goal planning of the drop-in, host side: the reference's default `mode/goal = prior_based`
// (src/traj_planner.cpp:540-608 goalPlanningWithPriority) with its grid planner (src/grid_based_planner.cpp) and the
// A* of src/Astar-3D. This is the step BEFORE the GPU path; it stays on the host (threaded over
// agents by ReplanBatch) when an octomap is loaded. Without an octomap the line-of-sight goal does not depend on the
// A* path at all and the engine computes the goal on the GPU (k_goal_plan, lscgpu_params::goal_mode).
//
// Interface mirrors include/grid_based_planner.hpp: GridBasedPlanner(distmap, mission, param), plan(...),
// findLOSFreeGoal(...), castRay(...).
//
// The search itself is this repository's own statement of the reference's algorithm on flat arrays: dense per-cell
// F / g / parent tables instead of node objects, and — because the reference breaks (F, g) ties inside a grid row by
// the ITERATION ORDER of a std::unordered_map (src/Astar-3D/isearch.cpp:209-242) — an explicit model of that
// container's node order (libstdc++ _Hashtable: one forward list, every bucket a contiguous run, new nodes at the
// front of their bucket's run or of the list, rehash re-threads the list in iteration order), driven by libstdc++'s
// own growth policy object. Keys hash to themselves, so no hashing code is involved.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <unordered_map>      // std::__detail::_Prime_rehash_policy
#include <vector>

#include "sp_const.hpp"

namespace DynamicPlanning {

#define GP_OCCUPIED 1
#define GP_EMPTY 0

// DynamicEDTOctomap::getDistance on the distance field the engine built (lscgpu_get_distmap_sqdist):
// cell distance * resolution as float, -1 outside the map.
struct HostDistMap {
    double res = 0.1;
    int size[3] = {0, 0, 0}, off[3] = {0, 0, 0};
    std::vector<uint8_t> sqdist;
    float getDistance(const point3d& p) const {
        const int x = (int)std::floor((1.0 / res) * (double)p.x()) - off[0];
        const int y = (int)std::floor((1.0 / res) * (double)p.y()) - off[1];
        const int z = (int)std::floor((1.0 / res) * (double)p.z()) - off[2];
        if (x < 0 || x >= size[0] || y < 0 || y >= size[1] || z < 0 || z >= size[2]) return -1.0f;
        const float cell = (float)std::sqrt((double)sqdist[((size_t)x * size[1] + y) * size[2] + z]);
        return (float)(cell * res);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Node order of one std::unordered_map<uint32_t, T> (libstdc++), nodes named by a caller-chosen dense index.
// ---------------------------------------------------------------------------------------------------------------
class HashOrderModel {
public:
    static constexpr int NONE = -2, BEFORE = -1;
    // `next` is shared by all models of one search (one slot per node index); `key_of(index)` = the map key
    void reset() { head_ = -1; count_ = 0; bkt_count_ = 1; bkt_.assign(1, NONE); pol_ = std::__detail::_Prime_rehash_policy(); }
    bool empty() const { return count_ == 0; }
    size_t size() const { return count_; }
    int begin() const { return head_; }

    template <class KeyOf>
    void insert(int node, std::vector<int>& next, const KeyOf& key_of) {
        const auto need = pol_._M_need_rehash(bkt_count_, count_, 1);
        if (need.first) rehash(need.second, next, key_of);
        const size_t b = key_of(node) % bkt_count_;
        if (bkt_[b] != NONE) {
            int& after = bkt_[b] == BEFORE ? head_ : next[bkt_[b]];
            next[node] = after; after = node;
        } else {
            next[node] = head_; head_ = node;
            if (next[node] >= 0) bkt_[key_of(next[node]) % bkt_count_] = node;
            bkt_[b] = BEFORE;
        }
        count_++;
    }
    template <class KeyOf>
    void erase(int node, std::vector<int>& next, const KeyOf& key_of) {
        const size_t b = key_of(node) % bkt_count_;
        int prev = bkt_[b];
        for (int p = prev == BEFORE ? head_ : next[prev]; p != node; p = next[p]) prev = p;
        const int nxt = next[node];
        if (prev == bkt_[b]) {                               // first node of its bucket
            const size_t nb = nxt >= 0 ? key_of(nxt) % bkt_count_ : 0;
            if (nxt < 0 || nb != b) {                        // the bucket becomes empty
                if (nxt >= 0) bkt_[nb] = bkt_[b];
                if (bkt_[b] == BEFORE) head_ = nxt;
                bkt_[b] = NONE;
            }
        } else if (nxt >= 0) {
            const size_t nb = key_of(nxt) % bkt_count_;
            if (nb != b) bkt_[nb] = prev;
        }
        (prev == BEFORE ? head_ : next[prev]) = nxt;
        count_--;
    }

private:
    template <class KeyOf>
    void rehash(size_t n, std::vector<int>& next, const KeyOf& key_of) {
        std::vector<int> nb(n, NONE);
        int p = head_;
        head_ = -1;
        size_t begin_bkt = 0;
        while (p >= 0) {
            const int nxt = next[p];
            const size_t b = key_of(p) % n;
            if (nb[b] == NONE) {
                next[p] = head_; head_ = p; nb[b] = BEFORE;
                if (next[p] >= 0) nb[begin_bkt] = p;
                begin_bkt = b;
            } else {
                int& after = nb[b] == BEFORE ? head_ : next[nb[b]];
                next[p] = after; after = p;
            }
            p = nxt;
        }
        bkt_.swap(nb);
        bkt_count_ = n;
    }
    int head_ = -1;
    size_t count_ = 0, bkt_count_ = 1;
    std::vector<int> bkt_;
    std::__detail::_Prime_rehash_policy pol_;
};

typedef std::array<int, 3> GridCell;

// A* of src/Astar-3D (isearch.cpp:48-284, astar.cpp:18-30) with the options GridBasedPlanner::planAstar passes
// (environmentoptions.cpp:13-21: Euclidean heuristic, 6-connected, unit cost; hweight 1, g-max tie break).
// grid[(i * dim[1] + j) * dim[2] + k] != 0 : occupied.
class AstarExact {
public:
    std::vector<GridCell> plan(const std::vector<uint8_t>& grid, const int dim[3], const GridCell& start, const GridCell& goal) {
        H = dim[0]; W = dim[1]; A = dim[2];
        const size_t cells = (size_t)H * W * A;
        F.assign(cells, 0.0); G.assign(cells, 0.0); parent.assign(cells, -1); state.assign(cells, 0); next.assign(cells, -1);
        rows.resize(H);
        for (auto& r : rows) r.reset();
        row_min.assign(H, -1);
        expansions = 0;
        int open_size = 0;
        auto key_of = [this](int c) {                        // Node::get_id (node.cpp:12-14); c = (i * W + j) * A + z
            const int z = c % A, ij = c / A, j = ij % W, i = ij / W;
            return (uint32_t)H * W * z + W * i + j;
        };
        auto heur = [&](int i, int j, int z) {
            return std::sqrt((double)((goal[0] - i) * (goal[0] - i) + (goal[1] - j) * (goal[1] - j) + (goal[2] - z) * (goal[2] - z)));
        };
        auto add_open = [&](int i, int c, double f, double g, int par) {                 // isearch.cpp:244-284
            bool inserted = false;
            HashOrderModel& row = rows[i];
            if (state[c] == 1) {
                if (f < F[c]) { F[c] = f; G[c] = g; parent[c] = par; inserted = true; }
            } else {
                F[c] = f; G[c] = g; parent[c] = par; state[c] = 1;
                row.insert(c, next, key_of);
                inserted = true; ++open_size;
            }
            if (row.size() == 1) row_min[i] = c;
            else if (inserted && f <= F[row_min[i]]) {
                if (f == F[row_min[i]]) { if (g >= G[row_min[i]]) row_min[i] = c; }
                else row_min[i] = c;
            }
        };
        const int cs = cell(start[0], start[1], start[2]);
        add_open(start[0], cs, heur(start[0], start[1], start[2]), 0.0, -1);
        open_size = 1;
        int cur = -1;
        bool found = false;
        while (open_size != 0) {
            // findMin (:177-207): rows ascending, a later row replaces the incumbent on equal F unless its g is smaller
            cur = -1;
            for (int i = 0; i < H; i++) {
                if (rows[i].empty()) continue;
                const int c = row_min[i];
                if (cur < 0 || F[c] < F[cur] || (F[c] == F[cur] && G[c] >= G[cur])) cur = c;
            }
            const int ci = cur / (W * A), cj = (cur / A) % W, cz = cur % A;
            state[cur] = 2;                                  // closed
            expansions++;
            // deleteMin (:209-242): erase, then re-scan the row in the container's iteration order
            rows[ci].erase(cur, next, key_of);
            int best = -1;
            for (int c = rows[ci].begin(); c >= 0; c = next[c])
                if (best < 0 || F[c] < F[best] || (F[c] == F[best] && G[c] >= G[best])) best = c;
            if (best >= 0) row_min[ci] = best;
            --open_size;
            if (ci == goal[0] && cj == goal[1]) { found = true; break; }                   // altitude ignored (:74)
            static const int mv[6][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}};
            for (int s = 0; s < 6; s++) {
                const int ni = ci + mv[s][0], nj = cj + mv[s][1], nz = cz + mv[s][2];
                if (ni < 0 || ni >= H || nj < 0 || nj >= W || nz < 0 || nz >= A) continue;
                const int c = cell(ni, nj, nz);
                if (grid[c] != 0 || state[c] == 2) continue;
                const double g = G[cur] + 1.0;
                add_open(ni, c, g + heur(ni, nj, nz), g, cur);
            }
        }
        std::vector<GridCell> path;
        if (!found) return path;
        for (int c = cur; c >= 0; c = parent[c]) path.push_back({c / (W * A), (c / A) % W, c % A});
        return std::vector<GridCell>(path.rbegin(), path.rend());
    }
    long long expansions = 0;

private:
    int cell(int i, int j, int z) const { return (i * W + j) * A + z; }
    int H = 0, W = 0, A = 0;
    std::vector<double> F, G;
    std::vector<int> parent, next, row_min;
    std::vector<uint8_t> state;                              // 0 unseen, 1 open, 2 closed
    std::vector<HashOrderModel> rows;
};

// What goal planning reads of a neighbour (dynamic_msgs::Obstacle fields of src/multi_sync_simulator.cpp:269-299)
struct GoalObstacle {
    int id;
    point3d position, goal_point;
    double radius, downwash;
    point3d prev_traj_first_end, prev_traj_last_end;        // obs_prev_trajs[oi][0][n], [M-1][n]
};

// The distance-field part of the occupancy grid (src/grid_based_planner.cpp:109-123) depends only on the map and the
// agent radius: computed once per radius and copied, instead of 18 000 getDistance calls per agent per step.
struct StaticGridCache {
    std::vector<std::pair<double, std::vector<uint8_t>>> by_radius;
    const std::vector<uint8_t>* find(double radius) const {
        for (const auto& e : by_radius) if (e.first == radius) return &e.second;
        return nullptr;
    }
};

class GridBasedPlanner {
public:
    GridBasedPlanner(const HostDistMap* _distmap, const Mission& _mission, const Param& _param,
                     const StaticGridCache* _cache = nullptr)
        : distmap(_distmap), mission(_mission), param(_param), cache(_cache) {}

    // the static occupancy of one radius, for StaticGridCache (same arithmetic as updateGridMap)
    std::vector<uint8_t> staticGrid(double agent_radius) {
        updateGridInfo();
        std::vector<uint8_t> g((size_t)dim[0] * dim[1] * dim[2], GP_EMPTY);
        if (distmap != nullptr) {
            const float grid_margin = (float)param.grid_margin;
            for (int i = 0; i < dim[0]; i++)
                for (int j = 0; j < dim[1]; j++)
                    for (int k = 0; k < dim[2]; k++)
                        if (distmap->getDistance(gridVectorToPoint3D(i, j, k)) < agent_radius + grid_margin) g[at(i, j, k)] = GP_OCCUPIED;
        }
        return g;
    }

    // plan (:53-68). high_priority == nullptr: no agent is an obstacle ("A* without priority")
    const std::vector<point3d>& plan(const point3d& current_position, const point3d& goal_position, double agent_radius,
                                     double agent_downwash, const std::vector<GoalObstacle>& obstacles,
                                     const std::vector<char>* high_priority) {
        updateGridInfo();
        updateGridMap(obstacles, agent_radius, agent_downwash, high_priority);
        path.clear();
        GridCell start = point3DToGridVector(current_position), goal = point3DToGridVector(goal_position);
        for (int i = 0; i < 3; i++)
            if (start[i] < 0 || start[i] >= dim[i] || goal[i] < 0 || goal[i] >= dim[i]) return path;    // reference: out-of-range access
        updateGridMission(start);
        const std::vector<GridCell> cells = astar.plan(grid, dim, start, goal);
        expansions += astar.expansions;
        for (const GridCell& c : cells) path.push_back(gridVectorToPoint3D(c[0], c[1], c[2]));
        return path;
    }

    point3d findLOSFreeGoal(const point3d& current_position, const point3d& goal_position, double agent_radius) const {   // :350-407
        point3d los_free_goal = current_position;
        std::vector<point3d> pts = path;
        pts.push_back(goal_position);
        for (int i = 0; i < 6; i++) {
            const double margin_ratio = 1.5 - 0.1 * i;
            for (const point3d& point : pts) {
                bool is_safe = true;
                if (distmap != nullptr) is_safe = castRay(current_position, point, agent_radius * margin_ratio);
                if (is_safe) los_free_goal = point;
                else break;
            }
            if ((los_free_goal - current_position).norm() > 0.3) break;
        }
        const point3d delta = los_free_goal - current_position;
        if (delta.norm() > param.goal_radius) los_free_goal = current_position + delta.normalized() * (float)param.goal_radius;
        return los_free_goal;
    }

    bool castRay(const point3d& current_position, const point3d& goal_position, double agent_radius) const {            // :409-434
        const double max_dist = 1.0;
        const double dist_to_goal = (current_position - goal_position).norm();
        const double dist_threshold = std::sqrt(0.25 * dist_to_goal * dist_to_goal + agent_radius * agent_radius);
        const double safe_dist_curr = distmap->getDistance(current_position);
        const double safe_dist_goal = distmap->getDistance(goal_position);
        if (safe_dist_curr < agent_radius + 0.5 * param.world_resolution - SP_EPSILON_FLOAT) return false;
        if (safe_dist_goal < agent_radius + 0.5 * param.world_resolution - SP_EPSILON_FLOAT) return false;
        if (dist_threshold < max_dist && safe_dist_curr > dist_threshold && safe_dist_goal > dist_threshold) return true;
        const point3d mid_position = (current_position + goal_position) * 0.5f;
        return castRay(current_position, mid_position, agent_radius) && castRay(mid_position, goal_position, agent_radius);
    }
    long long expansions = 0;

private:
    void updateGridInfo() {                                                                                             // :70-90
        const double r = param.grid_resolution;
        for (int i = 0; i < 3; i++) {
            grid_min[i] = -std::floor((-(double)mission.world_min(i) + SP_EPSILON) / r) * r;
            grid_max[i] = std::floor(((double)mission.world_max(i) + SP_EPSILON) / r) * r;
        }
        for (int i = 0; i < 3; i++) dim[i] = (int)std::round((grid_max[i] - grid_min[i]) / r) + 1;
    }
    point3d gridVectorToPoint3D(int i, int j, int k) const {                                                            // :305-310
        const double r = param.grid_resolution;
        return point3d((float)(grid_min[0] + i * r), (float)(grid_min[1] + j * r), (float)(grid_min[2] + k * r));
    }
    GridCell point3DToGridVector(const point3d& p) const {                                                              // :329-334
        const double r = param.grid_resolution;
        return {(int)std::round(((double)p.x() - grid_min[0]) / r), (int)std::round(((double)p.y() - grid_min[1]) / r),
                (int)std::round(((double)p.z() - grid_min[2]) / r)};
    }
    size_t at(int i, int j, int k) const { return ((size_t)i * dim[1] + j) * dim[2] + k; }
    void updateGridMap(const std::vector<GoalObstacle>& obstacles, double agent_radius, double agent_downwash,
                       const std::vector<char>* high_priority) {                                                        // :92-195
        const std::vector<uint8_t>* cached = cache ? cache->find(agent_radius) : nullptr;
        if (cached && cached->size() == (size_t)dim[0] * dim[1] * dim[2]) grid = *cached;
        else grid = staticGrid(agent_radius);
        const double r = param.grid_resolution;
        for (size_t oi = 0; oi < obstacles.size(); oi++) {
            if (high_priority == nullptr || !(*high_priority)[oi]) continue;
            const GoalObstacle& o = obstacles[oi];
            const double ox = o.position.x(), oy = o.position.y(), oz = o.position.z();
            const int obs_i = (int)std::round((ox - grid_min[0] + SP_EPSILON) / r);
            const int obs_j = (int)std::round((oy - grid_min[1] + SP_EPSILON) / r);
            const int obs_k = (int)std::round((oz - grid_min[2] + SP_EPSILON) / r);
            const int size_xy = (int)std::ceil((agent_radius + o.radius) / r);
            const int size_z = (int)std::ceil((agent_radius * agent_downwash + o.radius * o.downwash) / r);
            const double downwash_total = (agent_radius * agent_downwash + o.radius * o.downwash) / (agent_radius + o.radius);
            for (int i = std::max(obs_i - size_xy, 0); i <= std::min(obs_i + size_xy, dim[0] - 1); i++)
                for (int j = std::max(obs_j - size_xy, 0); j <= std::min(obs_j + size_xy, dim[1] - 1); j++)
                    for (int k = std::max(obs_k - size_z, 0); k <= std::min(obs_k + size_z, dim[2] - 1); k++) {
                        const point3d p = gridVectorToPoint3D(i, j, k);
                        const double dist = std::sqrt(std::pow(p.x() - ox, 2) + std::pow(p.y() - oy, 2) +
                                                      std::pow((p.z() - oz) / downwash_total, 2));
                        if (dist < agent_radius + o.radius) grid[at(i, j, k)] = GP_OCCUPIED;
                    }
        }
    }
    bool isOccupied(const GridCell& c) const {                                                                          // :257-264
        for (int i = 0; i < 3; i++) if (c[i] < 0 || c[i] > dim[i] - 1) return true;
        return grid[at(c[0], c[1], c[2])] == GP_OCCUPIED;
    }
    void updateGridMission(GridCell& start) {                                                                           // :197-245
        if (grid[at(start[0], start[1], start[2])] != GP_OCCUPIED) return;
        int min_dist = (int)SP_INFINITY;
        GridCell closest = start;
        for (int i = -2; i < 3; i++)
            for (int j = -2; j < 3; j++)
                for (int k = 2 - param.world_dimension; k < param.world_dimension - 1; k++) {
                    const GridCell cand = {start[0] + i, start[1] + j, start[2] + k};
                    if (!isOccupied(cand)) {
                        const int dist = std::abs(i) + std::abs(j) + std::abs(k);
                        if (dist < min_dist) { min_dist = dist; closest = cand; }
                    }
                }
        start = closest;
        if (grid[at(start[0], start[1], start[2])] == GP_OCCUPIED) grid[at(start[0], start[1], start[2])] = GP_EMPTY;
    }

    const HostDistMap* distmap;
    const Mission& mission;
    const Param& param;
    const StaticGridCache* cache;
    double grid_min[3] = {0, 0, 0}, grid_max[3] = {0, 0, 0};
    int dim[3] = {0, 0, 0};
    std::vector<uint8_t> grid;
    std::vector<point3d> path;                   // plan_result.path
    AstarExact astar;
};

struct GoalPlanResult {
    point3d goal;
    int kind = 0;                                // 0: A* + line of sight, 1: retreat from the closest higher-priority agent
    long long expansions = 0;
};

// goalPlanningWithPriority (src/traj_planner.cpp:540-608). `obstacles` = every other agent in id order.
inline GoalPlanResult goalPlanningWithPriority(const point3d& current_position, const point3d& desired_goal_position,
                                               const point3d& initial_traj_end, double agent_radius, double agent_downwash,
                                               const std::vector<GoalObstacle>& obstacles, const HostDistMap* distmap,
                                               const Mission& mission, const Param& param,
                                               const StaticGridCache* cache = nullptr) {
    GoalPlanResult out;
    std::vector<char> high(obstacles.size(), 0);
    int closest_obs_id = -1;
    const double dist_to_goal = (current_position - desired_goal_position).norm();
    double min_dist_to_obs = SP_INFINITY;
    for (size_t oi = 0; oi < obstacles.size(); oi++) {
        const GoalObstacle& o = obstacles[oi];
        const double obs_dist_to_goal = (o.position - o.goal_point).norm();
        const double dist_to_obs = (o.position - current_position).norm();
        if (obs_dist_to_goal < param.goal_threshold) continue;
        if (dist_to_goal > param.goal_threshold &&
            (o.prev_traj_last_end - o.prev_traj_first_end).dot(o.prev_traj_first_end - current_position) > 0) continue;
        if (dist_to_goal < param.goal_threshold || obs_dist_to_goal < dist_to_goal) {
            if (dist_to_obs < min_dist_to_obs) { min_dist_to_obs = dist_to_obs; closest_obs_id = (int)oi; }
            high[oi] = 1;
        }
    }
    const double dist_keep = param.priority_dist_threshold + 0.1;
    if (min_dist_to_obs < param.priority_dist_threshold) {
        out.goal = current_position - (obstacles[closest_obs_id].position - current_position).normalized() * (float)dist_keep;
        out.kind = 1;
        return out;
    }
    GridBasedPlanner planner(distmap, mission, param, cache);
    if (planner.plan(current_position, desired_goal_position, agent_radius, agent_downwash, obstacles, &high).empty())
        planner.plan(current_position, desired_goal_position, agent_radius, agent_downwash, obstacles, nullptr);
    out.goal = planner.findLOSFreeGoal(initial_traj_end, desired_goal_position, agent_radius);
    out.expansions = planner.expansions;
    return out;
}

}  // namespace DynamicPlanning
