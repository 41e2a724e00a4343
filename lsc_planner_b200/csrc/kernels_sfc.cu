// Distance field (k_edt_*), blocked-voxel summed-volume tables (k_sat_*) and SFC box expansion (k_sfc_expand).
//
// Replaces (reference paths):
//   DynamicEDTOctomap(1.0, tree, bbxMin, bbxMax, false) + update() + getDistance()   upstream dynamicEDT3D, called at
//       src/multi_sync_simulator.cpp:153-167 and include/corridor_constructor.hpp:113
//   CorridorConstructor::expandBoxFromPoint / expandSFCFromBox / expand_box / setAxisCand / isObstacleInBox /
//       isBoxInBoundary                                include/corridor_constructor.hpp:18-44,234-245,184-232,142-182,81-131
//   TrajPlanner::generateFeasibleSFC (window shift)     src/traj_planner.cpp:1451-1491
//
// B200 design. The reference validates every candidate box by sampling the distance field at each lattice point of
// the box (a triple loop, re-run over the WHOLE box up to six extra times per expansion). The set of voxels such a
// scan touches is, per axis, one contiguous run of voxels plus at most one extra voxel (the -1e-5 nudge of the minimum
// face samples the voxel below it), so "is any sample blocked" is a sum over at most 8 axis-aligned voxel boxes.
// We therefore build, once per map and per distinct agent radius, a summed-volume table of the predicate
// "getDistance(voxel) < radius + res/2 - 1e-5" (int32, (sx+1)(sy+1)(sz+1) entries, L2-resident), and every box test
// becomes 64 table reads done by one warp (2 per lane) and a shuffle reduction — O(1) instead of O(volume), with the
// reference's control flow (candidate order, full-box re-tests, boundary tests) kept literally, in the same double /
// float arithmetic, so the boxes are bit-identical.
#include "kernels.hpp"
#include "sfc.cuh"

namespace lscgpu {

// ------------------------------------------------------------------------------------------------------------
// Distance field: exact squared Euclidean cell distance to the nearest occupied cell, clamped at max_sq (=121 for
// maxdist 1.0 / res 0.1), three separable window passes.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_edt_scatter(const int32_t* keys, int n, DistMapDev dm, uint8_t* occ) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int x = keys[3 * t] - dm.off[0], y = keys[3 * t + 1] - dm.off[1], z = keys[3 * t + 2] - dm.off[2];
    if (x < 0 || x >= dm.size[0] || y < 0 || y >= dm.size[1] || z < 0 || z >= dm.size[2]) return;
    occ[((size_t)x * dm.size[1] + y) * dm.size[2] + z] = 1;
}

// axis 2: from occupancy; axis 1/0: from the previous pass
__global__ void k_edt_pass(DistMapDev dm, const uint8_t* src, uint8_t* dst, int axis, int radius, int from_occupancy) {
    const size_t total = (size_t)dm.size[0] * dm.size[1] * dm.size[2];
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int z = (int)(t % dm.size[2]);
    const int y = (int)((t / dm.size[2]) % dm.size[1]);
    const int x = (int)(t / ((size_t)dm.size[2] * dm.size[1]));
    const int pos = axis == 0 ? x : (axis == 1 ? y : z);
    const int len = dm.size[axis];
    const size_t stride = axis == 0 ? (size_t)dm.size[1] * dm.size[2] : (axis == 1 ? (size_t)dm.size[2] : 1);
    int best = dm.max_sq;
    for (int dlt = -radius; dlt <= radius; dlt++) {
        const int q = pos + dlt;
        if (q < 0 || q >= len) continue;
        const uint8_t s = src[t + (ptrdiff_t)dlt * (ptrdiff_t)stride];
        if (from_occupancy) { if (s) best = min(best, dlt * dlt); }
        else best = min(best, (int)s + dlt * dlt);
    }
    dst[t] = (uint8_t)min(best, dm.max_sq);
}

__global__ void k_sat_fill(DistMapDev dm, const int* thresholds, int table) {
    const size_t total = (size_t)dm.size[0] * dm.size[1] * dm.size[2];
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int z = (int)(t % dm.size[2]);
    const int y = (int)((t / dm.size[2]) % dm.size[1]);
    const int x = (int)(t / ((size_t)dm.size[2] * dm.size[1]));
    const int py = dm.size[1] + 1, pz = dm.size[2] + 1;
    const size_t tab = (size_t)(dm.size[0] + 1) * py * pz;
    dm.sat[table * tab + ((size_t)(x + 1) * py + (y + 1)) * pz + (z + 1)] = dm.sqdist[t] <= thresholds[table] ? 1 : 0;
}
// inclusive scan along one axis of the (sx+1)(sy+1)(sz+1) table; one thread per line
__global__ void k_sat_scan(DistMapDev dm, int table, int axis) {
    const int dims[3] = {dm.size[0] + 1, dm.size[1] + 1, dm.size[2] + 1};
    const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= dims[a1] * dims[a2]) return;
    int idx[3];
    idx[a1] = t / dims[a2]; idx[a2] = t % dims[a2];
    const size_t tab = (size_t)dims[0] * dims[1] * dims[2];
    int* S = dm.sat + table * tab;
    int run = 0;
    for (int p = 0; p < dims[axis]; p++) {
        idx[axis] = p;
        const size_t o = ((size_t)idx[0] * dims[1] + idx[1]) * dims[2] + idx[2];
        run += S[o];
        S[o] = run;
    }
}

void launch_edt_build(const int32_t* keys_dev, int n_keys, DistMapDev dm, const int* thresholds_dev, int n_tables,
                      uint8_t* scratch_a, uint8_t* scratch_b, cudaStream_t s) {
    const size_t total = (size_t)dm.size[0] * dm.size[1] * dm.size[2];
    const int blocks = (int)((total + 255) / 256);
    int radius = 1;
    while (radius * radius < dm.max_sq) radius++;
    cudaMemsetAsync(scratch_a, 0, total, s);
    if (n_keys > 0) k_edt_scatter<<<(n_keys + 255) / 256, 256, 0, s>>>(keys_dev, n_keys, dm, scratch_a);
    k_edt_pass<<<blocks, 256, 0, s>>>(dm, scratch_a, scratch_b, 2, radius, 1);
    k_edt_pass<<<blocks, 256, 0, s>>>(dm, scratch_b, scratch_a, 1, radius, 0);
    k_edt_pass<<<blocks, 256, 0, s>>>(dm, scratch_a, dm.sqdist, 0, radius, 0);
    const size_t tab = (size_t)(dm.size[0] + 1) * (dm.size[1] + 1) * (dm.size[2] + 1);
    cudaMemsetAsync(dm.sat, 0, tab * n_tables * sizeof(int), s);
    for (int t = 0; t < n_tables; t++) {
        k_sat_fill<<<blocks, 256, 0, s>>>(dm, thresholds_dev, t);
        for (int axis = 2; axis >= 0; axis--) {
            const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
            const int lines = (dm.size[a1] + 1) * (dm.size[a2] + 1);
            k_sat_scan<<<(lines + 127) / 128, 128, 0, s>>>(dm, t, axis);
        }
    }
}

// CorridorConstructor::expandBoxFromPoint for a batch of seeds (operator-level entry), one warp per seed
__global__ void __launch_bounds__(128) k_sfc_expand(SfcLaunch L) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= L.n) return;
    SfcCtx c;
    sfc_ctx_init(c, L.dm, L.wmin, L.wmax, L.res, threadIdx.x & 31);
    const size_t tab = (size_t)(L.dm.size[0] + 1) * (L.dm.size[1] + 1) * (L.dm.size[2] + 1);
    double face = 0.0;       // lane f < 6: coordinate of face f
    c.sat = L.dm.sat + (size_t)L.sat_index[warp] * tab;
    const F3 p{L.point[3 * warp], L.point[3 * warp + 1], L.point[3 * warp + 2]};
    const F3 g{L.goal[3 * warp], L.goal[3 * warp + 1], L.goal[3 * warp + 2]};
    const bool ok = expand_from_point(c, L.res, p, g, &face);
    if (c.lane < 6) L.box_out[6 * warp + c.lane] = ok ? (float)face : 0.0f;
    if (c.lane == 0) L.ok_out[warp] = ok ? 1 : 0;
}

// the step's new box of one agent per warp; published with a release of the step's epoch
__global__ void __launch_bounds__(128) k_sfc_step(SfcStepLaunch L) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= L.n) return;
    const int a = L.order ? L.order[L.order_first + w * L.order_stride] : L.agent_base + w * L.agent_stride;
    const int epoch = *L.epoch;
    SfcCtx c;
    sfc_ctx_init(c, L.dm, L.wmin, L.wmax, L.res, threadIdx.x & 31);
    double face;
    // a state reset in this step re-arms the corridor (src/traj_planner.cpp:1047-1061); k_predict, which makes the same
    // check, runs beside this kernel and may not have raised init_sfc yet
    bool first = L.init_sfc[a] != 0;
    if (L.planner_seq >= 2) {
        const float* t = L.prev_traj + (size_t)a * kTrajFloats + 18;
        const lscgpu_agent_in& in = L.in[a];
        const F3 dlt = f3_sub(F3{t[0], t[1], t[2]}, F3{in.position[0], in.position[1], in.position[2]});
        first = first || sqrt(f3_dot(dlt, dlt)) > L.reset_threshold;
    }
    // the box grows toward the agent's CURRENT goal: the input goal, or what goal planning made of it (goal_mode 1)
    const lscgpu_agent_in& ia = L.in[a];
    const F3 goal = L.goal3 ? F3{(float)L.goal3[(size_t)a * 3], (float)L.goal3[(size_t)a * 3 + 1], (float)L.goal3[(size_t)a * 3 + 2]}
                            : F3{ia.goal[0], ia.goal[1], ia.goal[2]};
    const bool ok = sfc_agent_box(c, L.dm, L.consts[a].sat_index, L.res, ia, goal, L.prev_traj + (size_t)a * kTrajFloats,
                                  first, face);
    if (c.lane < 6) L.sfc_box_g[(size_t)a * 6 + c.lane] = ok ? (float)face : 0.0f;
    if (c.lane == 0) L.sfc_ok_g[a] = ok ? 1 : 0;
    __threadfence();
    __syncwarp();
    if (c.lane == 0) atomicExch(&L.sfc_ready[a], epoch);
}
void launch_sfc_step(const SfcStepLaunch& L, cudaStream_t s) {
    if (L.n <= 0) return;
    k_sfc_step<<<(L.n * 32 + 127) / 128, 128, 0, s>>>(L);
}

void launch_sfc_expand(const SfcLaunch& L, cudaStream_t s) {
    if (L.n <= 0) return;
    const int blocks = (L.n * 32 + 127) / 128;
    k_sfc_expand<<<blocks, 128, 0, s>>>(L);
}

}  // namespace lscgpu
