// Thin C entry point around the REFERENCE's own A* (src/Astar-3D/*.cpp, include/Astar-3D/*.h), compiled from the
// sources where they lie under /root/reference into oracle/_ref/libref_astar.so. Used only by tests/test_oracle_goal.py
// to pin the oracle's restatement of the search (including its tie-breaking, which depends on the iteration order of
// std::unordered_map); never linked into the product. Mirrors GridBasedPlanner::planAstar
// (src/grid_based_planner.cpp:278-295): AstarPlanner::plan(grid, start, goal, default EnvironmentOptions).
// map.h includes "tinyxml2.h" without using it; the build passes -I oracle/ref_stub (an empty header of that name).
#include <array>
#include <vector>

#include "Astar-3D/astarplanner.h"

extern "C" int ref_astar_plan(const int* dim, const unsigned char* grid, const int* start, const int* goal, int* path_out,
                              int max_len) {
    std::vector<std::vector<std::vector<int>>> g(dim[0], std::vector<std::vector<int>>(dim[1], std::vector<int>(dim[2], 0)));
    for (int i = 0; i < dim[0]; i++)
        for (int j = 0; j < dim[1]; j++)
            for (int k = 0; k < dim[2]; k++) g[i][j][k] = grid[((size_t)i * dim[1] + j) * dim[2] + k] ? 1 : 0;
    AstarPlanner planner;
    EnvironmentOptions opt;
    SearchResult sr = planner.plan(g, {start[0], start[1], start[2]}, {goal[0], goal[1], goal[2]}, opt);
    if (!sr.pathfound) return 0;
    int n = 0;
    for (const auto& node : sr.lppath->List) {
        if (n < max_len) { path_out[3 * n] = node.i; path_out[3 * n + 1] = node.j; path_out[3 * n + 2] = node.z; }
        n++;
    }
    return n;
}
