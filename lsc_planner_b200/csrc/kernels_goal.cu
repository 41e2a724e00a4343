// k_goal_astar — goal planning WITH an octomap on the device (sm_100a), one warp per agent.
//
// Replaces, per agent and step, TrajPlanner::goalPlanningWithPriority (src/traj_planner.cpp:540-608):
//   priority rule + retreat point (:553-587)               all lanes over the neighbours, as k_goal_plan
//   GridBasedPlanner::plan (src/grid_based_planner.cpp:53-68): updateGridMap (:92-195: static occupancy from the distance
//   field — precomputed once per radius by k_goal_static_grid — plus the cells within r_i + r_j of every higher-priority
//   agent, z scaled by the pair's downwash), updateGridMission (:197-245), A* (src/Astar-3D, astar_core.cuh), and the
//   re-plan without priorities when no path exists (src/traj_planner.cpp:594-599)
//   findLOSFreeGoal + castRay (src/grid_based_planner.cpp:350-434): 32 path points per round, one per lane
//   clip to goal_radius, getTerminalSegments.
// The search is sequential by nature (its tie-breaking follows the iteration order of one hash container per grid row):
// lane 0 runs it, the warp reduces the row minima; what runs in parallel is the swarm — every agent has its own warp.
// Two instantiations: <uint16_t, shared> keeps the whole search state of the warp (cell bytes, g, list links, the row
// containers' buckets; 137 KB for the shipped 10 m world at grid/resolution 0.25) in shared memory, one warp per SM — a
// dependent access costs ~30 cycles; <int, global> for grids that do not fit runs on per-warp scratch in global memory
// (L2 resident, ~700 cycles per dependent access, but one warp per agent up to 16 per SM). Float arithmetic of octomath::Vector3 and the reference's double expressions
// are written with explicit roundings (no FMA contraction): goals are bit-identical to the oracle / the host planner.
#include <cstdio>
#include <cstdlib>

#include "astar_core.cuh"
#include "kernels.hpp"

namespace lscgpu {

__device__ __forceinline__ F3 f3_normalized_g(F3 v) {
    const double len = sqrt(f3_dot(v, v));
    if (len > 0.0) { const float l = (float)len; v.x = __fdiv_rn(v.x, l); v.y = __fdiv_rn(v.y, l); v.z = __fdiv_rn(v.z, l); }
    return v;
}
__device__ __forceinline__ double f3_norm(F3 v) { return sqrt(f3_dot(v, v)); }

// DynamicEDTOctomap::getDistance: (float)((float)sqrt(sqdist) * res), -1 outside the map
__device__ __forceinline__ float dist_at(const DistMapDev& dm, double inv_res, double res, F3 p) {
    const int x = (int)floor(__dmul_rn(inv_res, (double)p.x)) - dm.off[0];
    const int y = (int)floor(__dmul_rn(inv_res, (double)p.y)) - dm.off[1];
    const int z = (int)floor(__dmul_rn(inv_res, (double)p.z)) - dm.off[2];
    if (x < 0 || x >= dm.size[0] || y < 0 || y >= dm.size[1] || z < 0 || z >= dm.size[2]) return -1.0f;
    const float cell = (float)sqrt((double)dm.sqdist[((size_t)x * dm.size[1] + y) * dm.size[2] + z]);
    return (float)__dmul_rn((double)cell, res);
}

// static occupancy of the planning grid, one table per distinct agent radius (src/grid_based_planner.cpp:109-123)
__global__ void k_goal_static_grid(GoalGridDev g, DistMapDev dm, double world_res, const double* radii, int n_radii,
                                   float grid_margin, uint8_t* out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.cells) return;
    const int z = c % g.dim[2], ij = c / g.dim[2], j = ij % g.dim[1], i = ij / g.dim[1];
    const F3 p{g.axis_pts[i], g.axis_pts[g.dim[0] + j], g.axis_pts[g.dim[0] + g.dim[1] + z]};
    const double inv_res = __ddiv_rn(1.0, world_res);
    const float dist = dist_at(dm, inv_res, world_res, p);
    for (int t = 0; t < n_radii; t++)
        out[(size_t)t * g.cells_pad + c] = ((double)dist < __dadd_rn(radii[t], (double)grid_margin)) ? kCellOccupied : 0;
}
void launch_goal_static_grid(const GoalGridDev& g, const DistMapDev& dm, double world_res, const double* radii_dev, int n_radii,
                             float grid_margin, uint8_t* out, cudaStream_t s) {
    k_goal_static_grid<<<(g.cells + 255) / 256, 256, 0, s>>>(g, dm, world_res, radii_dev, n_radii, grid_margin, out);
}

// castRay (src/grid_based_planner.cpp:409-434) without recursion: depth-first over the bisection tree, left half first;
// the stack holds the right end points still to be reached.
__device__ bool cast_ray(const DistMapDev& dm, double inv_res, double res, F3 a, F3 b, double radius) {
    constexpr int kDepth = 40;
    F3 stack[kDepth];
    int sp = 0;
    const double lim = __dsub_rn(__dadd_rn(radius, __dmul_rn(0.5, res)), 1e-5);
    double sa = (double)dist_at(dm, inv_res, res, a);
    if (sa < lim) return false;
    for (;;) {
        const double dist = f3_norm(f3_sub(a, b));
        const double thr = sqrt(__dadd_rn(__dmul_rn(__dmul_rn(0.25, dist), dist), __dmul_rn(radius, radius)));
        const double sb = (double)dist_at(dm, inv_res, res, b);
        if (sb < lim) return false;
        if (thr < 1.0 && sa > thr && sb > thr) {
            if (sp == 0) return true;
            a = b; sa = sb;                 // the right neighbour starts where this piece ended
            b = stack[--sp];
        } else {
            if (sp == kDepth) return false;
            stack[sp++] = b;
            b = f3_scale(f3_add(a, b), 0.5f);
        }
    }
}

// dynamic shared memory: row table [H] (min_f, head, count, level, min_cell, min_g, stamp, jlo, jhi), then — shared variant — the search
// state of the warp
__host__ __device__ inline size_t goal_rows_bytes(int H) { return ((size_t)H * (sizeof(double) + 8 * sizeof(int)) + 15) / 16 * 16; }
template <typename I>
__host__ __device__ inline size_t goal_state_bytes(const GoalGridDev& g) {
    return g.cells_pad * (1 + 2 * sizeof(I)) + 2 * (((size_t)g.dim[0] * g.bcap * sizeof(I) + 15) / 16 * 16);     // buckets + their order stamps
}
// entries of the table h = sqrt(d2), d2 = squared cell distance to the goal (shared variant)
__host__ __device__ inline int goal_sqrt_entries(const GoalGridDev& g) {
    return (g.dim[0] - 1) * (g.dim[0] - 1) + (g.dim[1] - 1) * (g.dim[1] - 1) + (g.dim[2] - 1) * (g.dim[2] - 1) + 1;
}

template <typename I, bool kShared>
__global__ void __launch_bounds__(32, 1) k_goal_astar(GoalAstarLaunch L) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_seq[kAstarMaxLevels];
    __shared__ unsigned s_magic[kAstarMaxLevels];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const GoalGridDev& G = L.grid;
    const int H = G.dim[0], W = G.dim[1], A = G.dim[2];
    const double inv_res = __ddiv_rn(1.0, L.world_res);
    double* r_min_f = reinterpret_cast<double*>(smem);
    int* r_head = reinterpret_cast<int*>(r_min_f + H);
    int* r_count = r_head + H; int* r_level = r_count + H; int* r_min_cell = r_level + H; int* r_min_g = r_min_cell + H;
    int* r_stamp = r_min_g + H; int* r_jlo = r_stamp + H; int* r_jhi = r_jlo + H;
    uint8_t* cellb; I* gbuf; I* nextb; I* bkt; I* bstamp;
    double* sqrt_tab = nullptr;
    if (kShared) {
        unsigned char* p = smem + goal_rows_bytes(H);
        sqrt_tab = reinterpret_cast<double*>(p); p += ((size_t)goal_sqrt_entries(G) * sizeof(double) + 15) / 16 * 16;
        for (int k = lane; k < goal_sqrt_entries(G); k += 32) sqrt_tab[k] = sqrt((double)k);
        cellb = p; p += G.cells_pad;
        gbuf = reinterpret_cast<I*>(p); p += G.cells_pad * sizeof(I);
        nextb = reinterpret_cast<I*>(p); p += G.cells_pad * sizeof(I);
        bkt = reinterpret_cast<I*>(p); p += ((size_t)H * G.bcap * sizeof(I) + 15) / 16 * 16;
        bstamp = reinterpret_cast<I*>(p);
    } else {
        cellb = L.cell + (size_t)blockIdx.x * G.cells_pad;
        gbuf = reinterpret_cast<I*>(L.gcost) + (size_t)blockIdx.x * G.cells_pad;
        nextb = reinterpret_cast<I*>(L.next) + (size_t)blockIdx.x * G.cells_pad;
        bkt = reinterpret_cast<I*>(L.bkt) + (size_t)blockIdx.x * ((size_t)H * G.bcap);
        bstamp = reinterpret_cast<I*>(L.bstamp) + (size_t)blockIdx.x * ((size_t)H * G.bcap);
    }
    int* path = L.path + (size_t)blockIdx.x * G.cells_pad;
    const GoalLaunch& P = L.g;
    if (lane < kAstarMaxLevels) { s_seq[lane] = G.bkt_seq[lane]; s_magic[lane] = G.bkt_magic[lane]; }
    __syncwarp();

    // agents are handed out dynamically (one atomic per agent): searches differ by orders of magnitude in length, a static
    // split would leave most warps waiting for the unluckiest one
    for (;;) {
        int blk = 0;
        if (lane == 0) blk = atomicAdd(L.next_agent, 1);
        blk = __shfl_sync(FULL, blk, 0);
        if (blk >= L.n) break;
        const int a = L.order ? L.order[L.order_first + blk * L.order_stride] : L.agent_base + blk * L.agent_stride;
        const lscgpu_agent_in& me = P.in[a];
        const F3 pos{me.position[0], me.position[1], me.position[2]};
        const F3 desired{me.goal[0], me.goal[1], me.goal[2]};
        const double dist_to_goal = f3_norm(f3_sub(pos, desired));
        const AgentConstDev ca = P.consts[a];
        const uint8_t* static_occ = G.static_occ + (size_t)ca.sat_index * G.cells_pad;

        // ---- the planning grid starts as the static occupancy of this agent's radius (16 bytes per lane and round)
        {
            const uint4* src = reinterpret_cast<const uint4*>(static_occ);
            uint4* dst = reinterpret_cast<uint4*>(cellb);
            for (size_t k = lane; k < G.cells_pad / 16; k += 32) dst[k] = src[k];
        }
        __syncwarp();

        // ---- priority rule (:553-575); every higher-priority neighbour is stamped into the grid at once
        double best = 1e9;
        int best_j = -1;
        const bool self_reset = P.reset_ever[a] != 0;
        for (int j = lane; j < P.n_agents; j += 32) {
            if (j == a) continue;
            const lscgpu_agent_in& o = P.in[j];
            const F3 op{o.position[0], o.position[1], o.position[2]};
            bool high = false;
            if (self_reset || P.reset_ever[j]) {
                high = true;                                  // obs_slack_indices: high priority, never the retreat target (:548-551)
            } else {
                const F3 og{o.goal[0], o.goal[1], o.goal[2]};
                const double obs_dist_to_goal = f3_norm(f3_sub(op, og));
                const double dist_to_obs = f3_norm(f3_sub(op, pos));
                if (obs_dist_to_goal < P.goal_threshold) continue;
                const float* t = P.prev_traj + (size_t)j * kTrajFloats;
                const F3 first_end{t[15], t[16], t[17]}, last_end{t[87], t[88], t[89]};
                if (dist_to_goal > P.goal_threshold && f3_dot(f3_sub(last_end, first_end), f3_sub(first_end, pos)) > 0.0) continue;
                if (dist_to_goal < P.goal_threshold || obs_dist_to_goal < dist_to_goal) {
                    if (dist_to_obs < best) { best = dist_to_obs; best_j = j; }
                    high = true;
                }
            }
            if (!high) continue;
            // updateGridMap (:125-195): cells within r_i + r_j of the neighbour, z scaled by the pair's downwash
            const AgentConstDev cj = P.consts[j];
            const double ox = (double)op.x, oy = (double)op.y, oz = (double)op.z;
            const int oi = (int)round(__ddiv_rn(__dadd_rn(__dsub_rn(ox, G.gmin[0]), 1e-9), G.res));
            const int oj = (int)round(__ddiv_rn(__dadd_rn(__dsub_rn(oy, G.gmin[1]), 1e-9), G.res));
            const int ok = (int)round(__ddiv_rn(__dadd_rn(__dsub_rn(oz, G.gmin[2]), 1e-9), G.res));
            const double rsum = __dadd_rn(ca.radius, cj.radius);
            const double zsum = __dadd_rn(__dmul_rn(ca.radius, ca.downwash), __dmul_rn(cj.radius, cj.downwash));
            const int size_xy = (int)ceil(__ddiv_rn(rsum, G.res));
            const int size_z = (int)ceil(__ddiv_rn(zsum, G.res));
            const double dw = __ddiv_rn(zsum, rsum);
            for (int i = max(oi - size_xy, 0); i <= min(oi + size_xy, H - 1); i++)
                for (int jj = max(oj - size_xy, 0); jj <= min(oj + size_xy, W - 1); jj++)
                    for (int k = max(ok - size_z, 0); k <= min(ok + size_z, A - 1); k++) {
                        const double dx = __dsub_rn((double)G.axis_pts[i], ox), dy = __dsub_rn((double)G.axis_pts[H + jj], oy);
                        const double dz = __ddiv_rn(__dsub_rn((double)G.axis_pts[H + W + k], oz), dw);
                        const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
                        if (d < rsum) cellb[(i * W + jj) * A + k] = kCellOccupied;
                    }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(FULL, best, o);
            const int oj = __shfl_xor_sync(FULL, best_j, o);
            if (oj >= 0 && (best_j < 0 || ob < best || (ob == best && oj < best_j))) { best = ob; best_j = oj; }
        }
        __syncwarp();

        F3 goal;
        int kind = 0;
        if (best_j >= 0 && best < P.priority_dist_threshold) {
            // retreat from the closest higher-priority agent (:580-587)
            const lscgpu_agent_in& o = P.in[best_j];
            const F3 op{o.position[0], o.position[1], o.position[2]};
            const double dist_keep = P.priority_dist_threshold + 0.1;
            goal = f3_sub(pos, f3_scale(f3_normalized_g(f3_sub(op, pos)), (float)dist_keep));
            kind = 1;
        } else {
            // ---- GridBasedPlanner::plan, with the priorities and — when that finds no path — without them
            int sc[3], gc[3];
            sc[0] = (int)round(__ddiv_rn(__dsub_rn((double)pos.x, G.gmin[0]), G.res));
            sc[1] = (int)round(__ddiv_rn(__dsub_rn((double)pos.y, G.gmin[1]), G.res));
            sc[2] = (int)round(__ddiv_rn(__dsub_rn((double)pos.z, G.gmin[2]), G.res));
            gc[0] = (int)round(__ddiv_rn(__dsub_rn((double)desired.x, G.gmin[0]), G.res));
            gc[1] = (int)round(__ddiv_rn(__dsub_rn((double)desired.y, G.gmin[1]), G.res));
            gc[2] = (int)round(__ddiv_rn(__dsub_rn((double)desired.z, G.gmin[2]), G.res));
            bool in_grid = true;
            for (int k = 0; k < 3; k++)
                if (sc[k] < 0 || sc[k] >= G.dim[k] || gc[k] < 0 || gc[k] >= G.dim[k]) in_grid = false;   // the reference indexes out of range
            int n_path = 0;
            long long expanded = 0;
            for (int attempt = 0; attempt < 2 && in_grid && n_path == 0; attempt++) {
                if (attempt == 1) {
                    const uint4* src = reinterpret_cast<const uint4*>(static_occ);
                    uint4* dst = reinterpret_cast<uint4*>(cellb);
                    for (size_t k = lane; k < G.cells_pad / 16; k += 32) dst[k] = src[k];
                }
                for (int r = lane; r < H; r += 32) { r_head[r] = -1; r_count[r] = 0; r_level[r] = 0; r_stamp[r] = 0; r_jlo[r] = W; r_jhi[r] = -1; }
                __syncwarp();
                AstarCtx<I> c;
                c.cell = cellb; c.g = gbuf; c.next = nextb; c.bkt = bkt; c.bcap = G.bcap;
                c.bstamp = bstamp; c.stamp = r_stamp; c.jlo = r_jlo; c.jhi = r_jhi;
                c.head = r_head; c.count = r_count; c.level = r_level; c.min_cell = r_min_cell; c.min_g = r_min_g; c.min_f = r_min_f;
                c.H = H; c.W = W; c.A = A;
                c.bkt_seq = s_seq;
                c.magic_a = kShared ? G.magic_a : 0u; c.magic_w = kShared ? G.magic_w : 0u;
                c.bkt_magic = kShared ? s_magic : nullptr;
                c.sqrt_tab = sqrt_tab;
                c.gi = gc[0]; c.gj = gc[1]; c.gz = gc[2];
                c.expansions = 0; c.open_size = 0;
#ifdef LSCGPU_GOAL_TIMERS
                for (int k = 0; k < 6; k++) c.tsec[k] = 0;
                const long long t_search = clock64();
#endif
                if (lane == 0) {
                    // updateGridMission (:197-245): an occupied start cell moves to the nearest free cell of its 5 x 5 x 3
                    // neighbourhood (first in scan order among equals) and is cleared if none is free
                    int s0 = sc[0], s1 = sc[1], s2 = sc[2];
                    if (cellb[astar_cell(c, s0, s1, s2)] & kCellOccupied) {
                        int min_dist = 1000000000, b0 = s0, b1 = s1, b2 = s2;
                        for (int i = -2; i < 3; i++)
                            for (int j = -2; j < 3; j++)
                                for (int k = -1; k < 2; k++) {
                                    const int q0 = s0 + i, q1 = s1 + j, q2 = s2 + k;
                                    if (q0 < 0 || q0 >= H || q1 < 0 || q1 >= W || q2 < 0 || q2 >= A) continue;
                                    if (cellb[astar_cell(c, q0, q1, q2)] & kCellOccupied) continue;
                                    const int d = abs(i) + abs(j) + abs(k);
                                    if (d < min_dist) { min_dist = d; b0 = q0; b1 = q1; b2 = q2; }
                                }
                        s0 = b0; s1 = b1; s2 = b2;
                        cellb[astar_cell(c, s0, s1, s2)] &= (uint8_t)~kCellOccupied;
                    }
                    astar_begin(c, s0, s1, s2);
                }
                __syncwarp();
                int cur = -1, found = 0;
                const unsigned NOKEY = 0xffffffffu;
                for (;;) {
                    // findMin (isearch.cpp:177-207) over one slice of rows per lane: lowest F, then highest g, then the later
                    // row. F > 0, so its bit pattern orders like the value: four warp-wide integer reductions (redux.sync)
                    double bf = 0.0; int bg = -1, br = -1;
                    for (int r = lane; r < H; r += 32) {
                        if (r_count[r] == 0) continue;
                        const double f = r_min_f[r]; const int g = r_min_g[r];
                        if (br < 0 || f < bf || (f == bf && g >= bg)) { bf = f; bg = g; br = r; }
                    }
                    {
                        const unsigned hi = br >= 0 ? (unsigned)__double2hiint(bf) : NOKEY, lo = (unsigned)__double2loint(bf);
                        const unsigned mhi = __reduce_min_sync(FULL, hi);
                        if (mhi == NOKEY) break;                             // open list empty: no path
                        bool cand = hi == mhi;
                        const unsigned mlo = __reduce_min_sync(FULL, cand ? lo : NOKEY);
                        cand = cand && lo == mlo;
                        const int mg = __reduce_max_sync(FULL, cand ? bg : -1);
                        cand = cand && bg == mg;
                        br = __reduce_max_sync(FULL, cand ? br : -1);
                    }
                    cur = r_min_cell[br];
#ifdef LSCGPU_GOAL_TIMERS
                    { const long long t = clock64(); if (c.expansions) c.tsec[4] += t - c.tlast; c.tlast = t; }
#endif
                    // (1) lane 0 closes the node and takes it out of its row container
                    int ci = 0, cj = 0, cz = 0, cur_g = 0;
                    if (lane == 0) cur_g = astar_close(c, cur, ci, cj, cz);
                    ci = __shfl_sync(FULL, ci, 0);
                    __syncwarp();
                    // (2) deleteMin's re-scan of the row, by the warp: the open cells of grid row ci ARE the container's
                    // content; among the nodes with the minimal (F, g) the reference takes the LAST in the container's iteration
                    // order = the one whose bucket has the smallest order stamp (astar_core.cuh); only minimal nodes sharing
                    // a bucket make lane 0 walk that bucket's run
                    {
                        const int row_base = ci * W * A, di = gc[0] - ci;
                        // only the span of j in which this row has ever had an open node
                        const int o_lo = r_jlo[ci] * A, o_hi = (r_jhi[ci] + 1) * A;
                        double sf = 0.0; int sg = -1, scell = -1; unsigned sst = 0; bool have_st = false, dup = false;
                        for (int o = o_lo + lane; o < o_hi; o += 32) {
                            const int id = row_base + o;
                            if ((cellb[id] & kCellStateMask) != kCellOpen) continue;
                            const int pg = (int)gbuf[id];
                            const int pj = (int)astar_div((unsigned)o, (unsigned)A, c.magic_a), pz = o - pj * A;
                            const int dj = gc[1] - pj, dz = gc[2] - pz;
                            const double pf = (double)pg + astar_h(c, di * di + dj * dj + dz * dz);
                            if (scell < 0 || pf < sf || (pf == sf && pg > sg)) { sf = pf; sg = pg; scell = id; have_st = false; dup = false; }
                            else if (pf == sf && pg == sg) {
                                if (!have_st) { sst = astar_stamp_of(c, ci, scell); have_st = true; }
                                const unsigned st = astar_stamp_of(c, ci, id);
                                if (st < sst) { scell = id; sst = st; dup = false; }
                                else if (st == sst) dup = true;
                            }
                        }
                        const unsigned hi = scell >= 0 ? (unsigned)__double2hiint(sf) : NOKEY, lo = (unsigned)__double2loint(sf);
                        const unsigned mhi = __reduce_min_sync(FULL, hi);
                        if (mhi != NOKEY) {
                            bool cand = hi == mhi;
                            const unsigned mlo = __reduce_min_sync(FULL, cand ? lo : NOKEY);
                            cand = cand && lo == mlo;
                            const int mg = __reduce_max_sync(FULL, cand ? sg : -1);
                            cand = cand && sg == mg;
                            int mcell;
                            const unsigned cmask = __ballot_sync(FULL, cand);
                            bool walk = false;
                            if (__popc(cmask) == 1 && !__shfl_sync(FULL, (int)dup, __ffs(cmask) - 1)) {
                                mcell = __shfl_sync(FULL, scell, __ffs(cmask) - 1);
                            } else {
                                if (cand && !have_st) sst = astar_stamp_of(c, ci, scell);
                                const unsigned ms = __reduce_min_sync(FULL, cand ? sst : NOKEY);
                                cand = cand && sst == ms;
                                const unsigned m2 = __ballot_sync(FULL, cand);
                                walk = __popc(m2) > 1 || __shfl_sync(FULL, (int)dup, __ffs(m2) - 1) != 0;
                                mcell = __shfl_sync(FULL, scell, __ffs(m2) - 1);
                            }
                            if (lane == 0) {
                                const double mf = __hiloint2double((int)mhi, (int)mlo);
                                if (walk) mcell = astar_last_in_bucket(c, ci, mcell, mf, mg);
                                r_min_cell[ci] = mcell; r_min_f[ci] = mf; r_min_g[ci] = mg;
                            }
                        }
#ifdef LSCGPU_GOAL_TIMERS
                        if (lane == 0) { const long long t = clock64(); c.tsec[2] += t - c.tlast; c.tlast = t; }
#endif
                    }
                    // (3) goal test, neighbours (lane 0 rewrites what the other lanes have just read: order the two)
                    __syncwarp();
                    if (lane == 0) found = astar_open_neighbours(c, ci, cj, cz, cur_g);
#ifdef LSCGPU_GOAL_TIMERS
                    c.tlast = clock64();
#endif
                    found = __shfl_sync(FULL, found, 0);
                    __syncwarp();
                    if (found) break;
                }
#ifdef LSCGPU_GOAL_TIMERS
                if (lane == 0)
                    printf("[goal] agent %d attempt %d expansions %lld cycles %lld | split %lld erase %lld rescan %lld neighbours %lld find-min %lld\n",
                           a, attempt, c.expansions, clock64() - t_search, c.tsec[0], c.tsec[1], c.tsec[2], c.tsec[3], c.tsec[4]);
#endif
                if (lane == 0) {
                    expanded += c.expansions;
                    if (found) {
                        const int n = (int)gbuf[cur] + 1;
                        for (int k = n - 1, p = cur; k >= 0 && p >= 0; k--, p = astar_parent(c, p)) path[k] = p;
                        n_path = n;
                    }
                }
                n_path = __shfl_sync(FULL, n_path, 0);
                __syncwarp();
            }
            if (lane == 0 && L.expansions) atomicAdd(L.expansions, (unsigned long long)expanded);

            // ---- findLOSFreeGoal (:350-407): the farthest path point (then the desired goal) seen from the end of the
            // initial trajectory, with a margin shrinking from 1.5 r to r until the goal is more than 0.3 m away
            const float* pr = P.pred + (size_t)a * kTrajFloats;
            const F3 current{pr[87], pr[88], pr[89]};
            F3 los = current;
            const int n_pts = n_path + 1;
            for (int i = 0; i < 6; i++) {
                const double ratio = __dsub_rn(1.5, __dmul_rn(0.1, (double)i));
                const double rr = __dmul_rn(ca.radius, ratio);
                int first_bad = n_pts;
                for (int base = 0; base < n_pts; base += 32) {
                    const int idx = base + lane;
                    bool safe = true;
                    if (idx < n_pts) {
                        F3 pt = desired;
                        if (idx < n_path) {
                            const int pc = path[idx];
                            pt = F3{G.axis_pts[pc / (W * A)], G.axis_pts[H + (pc / A) % W], G.axis_pts[H + W + pc % A]};
                        }
                        safe = cast_ray(L.dm, inv_res, L.world_res, current, pt, rr);
                    }
                    const unsigned bad = __ballot_sync(FULL, !safe);
                    if (bad) { first_bad = base + __ffs(bad) - 1; break; }
                }
                if (first_bad > 0) {
                    const int idx = first_bad - 1;
                    if (idx < n_path) {
                        const int pc = path[idx];
                        los = F3{G.axis_pts[pc / (W * A)], G.axis_pts[H + (pc / A) % W], G.axis_pts[H + W + pc % A]};
                    } else {
                        los = desired;
                    }
                }
                if (f3_norm(f3_sub(los, current)) > 0.3) break;
            }
            const F3 delta = f3_sub(los, current);
            goal = los;
            if (f3_norm(delta) > P.goal_radius) goal = f3_add(current, f3_scale(f3_normalized_g(delta), (float)P.goal_radius));
        }
        if (lane == 0) {
            P.goal3[(size_t)a * 3] = (double)goal.x; P.goal3[(size_t)a * 3 + 1] = (double)goal.y; P.goal3[(size_t)a * 3 + 2] = (double)goal.z;
            // getTerminalSegments (src/traj_optimizer.cpp:541-548)
            const F3 gd = f3_sub(goal, pos);
            const double ideal = __ddiv_rn(sqrt(f3_dot(gd, gd)), ca.v_nom);
            const double q = __ddiv_rn(__dadd_rn(__dsub_rn(__dmul_rn((double)kM, P.dt), ideal), 1e-9), P.dt);
            int ts = (int)q;
            ts = ts < 1 ? 1 : (ts > kM ? kM : ts);
            P.ts[a] = ts;
            P.goal_kind[a] = kind;
        }
        __syncwarp();
    }
}

// Shared-memory bytes of the <uint16_t, shared> instantiation, or 0 when the grid does not fit (then <int, global> runs)
size_t goal_astar_shared_bytes(const GoalGridDev& g) {
    static const bool allow = [] { const char* v = getenv("LSCGPU_GOAL_SHARED"); return !v || atoi(v) != 0; }();   // 0: A/B against the global variant
    if (!allow || g.cells > 65533) return 0;
    const size_t bytes = goal_rows_bytes(g.dim[0]) + ((size_t)goal_sqrt_entries(g) * sizeof(double) + 15) / 16 * 16 + goal_state_bytes<uint16_t>(g);
    return bytes <= 227 * 1024 - 256 ? bytes : 0;
}
cudaError_t configure_goal_astar(const GoalGridDev& g) {
    const size_t sh = goal_astar_shared_bytes(g);
    if (sh) return cudaFuncSetAttribute(k_goal_astar<uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    return cudaFuncSetAttribute(k_goal_astar<int, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)goal_rows_bytes(g.dim[0]));
}

void launch_goal_astar(const GoalAstarLaunch& L, cudaStream_t s) {
    if (L.n <= 0) return;
    const int blocks = L.n_blocks < L.n ? L.n_blocks : L.n;
    const size_t sh = goal_astar_shared_bytes(L.grid);
    if (sh) k_goal_astar<uint16_t, true><<<blocks, 32, sh, s>>>(L);
    else k_goal_astar<int, false><<<blocks, 32, goal_rows_bytes(L.grid.dim[0]), s>>>(L);
}

}  // namespace lscgpu
