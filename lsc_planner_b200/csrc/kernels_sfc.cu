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

// ------------------------------------------------------------------------------------------------------------
// SFC expansion, one warp per seed.
// ------------------------------------------------------------------------------------------------------------
// A box is tracked twice: as the reference's doubles (faces moved by repeated +-res, so the float32 output carries
// exactly the reference's rounding) and as integer lattice planes (plane P <-> coordinate P*res) that drive the voxel
// tests. For |coordinate| < 64 m the reference's float sample `(float)(box + it*res) +- 1e-5f` always lands in voxel
// P (+ nudge) or P-1 (- nudge): float32 rounding there is < 4e-6 per operation, below the 1e-5 nudge, so the integer
// model reproduces OcTree::coordToKey of every sample exactly (the engine rejects larger worlds).
struct LatticeBox {
    double d[6];          // min xyz, max xyz
    int p[6];             // lattice planes of the faces
};

struct SfcCtx {
    DistMapDev dm;
    const int* sat;       // table of this seed's radius
    double res;
    double wmin[3], wmax[3];
    double wmin_nudge[3]; // world_min + 1e-5 (include/corridor_constructor.hpp:104)
    int lane;
};

__device__ __forceinline__ int sat_at(const SfcCtx& c, int x, int y, int z) {
    return c.sat[((size_t)x * (c.dm.size[1] + 1) + y) * (c.dm.size[2] + 1) + z];
}

// number of blocked voxels in the cell box [lo, hi] (inclusive, inside the map): 8 table reads by lanes 0-7
__device__ __forceinline__ int blocked_in_cells(const SfcCtx& c, const int* lo, const int* hi) {
    int v = 0;
    if (c.lane < 8) {
        const int cx = (c.lane & 1) ? hi[0] + 1 : lo[0];
        const int cy = (c.lane & 2) ? hi[1] + 1 : lo[1];
        const int cz = (c.lane & 4) ? hi[2] + 1 : lo[2];
        // inclusion-exclusion: + when the number of "lo" picks is even, i.e. an odd number of "hi+1" picks
        v = (__popc(c.lane) & 1) ? sat_at(c, cx, cy, cz) : -sat_at(c, cx, cy, cz);
    }
    return warp_sum_int(v);
}

// isObstacleInBox(box, margin) (include/corridor_constructor.hpp:81-122): true iff any sampled voxel is blocked or
// outside the map. Per axis the scan visits the voxel of the it == 0 sample (below the min face unless the face sits
// on the world boundary) and the voxels lo+1 .. hi of the it >= 1 samples (voxel lo again for a flat box).
__device__ bool obstacle_in_box(const SfcCtx& c, const LatticeBox& bx) {
    int extra[3], lo[3], hi[3];
    bool outside = false;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int pl = bx.p[i], ph = bx.p[i + 3];
        const bool minus = bx.d[i] > c.wmin_nudge[i];
        extra[i] = (minus ? pl - 1 : pl) - c.dm.off[i];
        if (ph - pl + 1 <= 1) { lo[i] = hi[i] = pl - c.dm.off[i]; }
        else { lo[i] = pl + 1 - c.dm.off[i]; hi[i] = ph - c.dm.off[i]; }
        const int n = c.dm.size[i];
        if (extra[i] < 0 || extra[i] >= n || lo[i] < 0 || hi[i] >= n) outside = true;
    }
    if (outside) return true;       // getDistance = -1 outside the map: always below the threshold
    int sum = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int t = c.lane + 32 * h;
        const int sub = t >> 3, corner = t & 7;
        int coord[3];
        int sign = 1;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const bool single = (sub >> i) & 1;
            const int l = single ? extra[i] : lo[i];
            const int u = single ? extra[i] : hi[i];
            if ((corner >> i) & 1) coord[i] = u + 1;
            else { coord[i] = l; sign = -sign; }
        }
        sum += sign * sat_at(c, coord[0], coord[1], coord[2]);
    }
    sum = warp_sum_int(sum);
    return sum != 0;
}

__device__ __forceinline__ bool box_in_boundary(const SfcCtx& c, const double* b) {
    const double eps = 1e-9;
    return b[0] > c.wmin[0] - eps && b[1] > c.wmin[1] - eps && b[2] > c.wmin[2] - eps &&
           b[3] < c.wmax[0] + eps && b[4] < c.wmax[1] + eps && b[5] < c.wmax[2] + eps;
}

// setAxisCand (include/corridor_constructor.hpp:142-182): faces toward the goal first, largest offset first
__device__ void axis_candidates(const double* box, F3 goal, int* cand) {
    const F3 mid{(float)(0.5 * (box[0] + box[3])), (float)(0.5 * (box[1] + box[4])), (float)(0.5 * (box[2] + box[5]))};
    const F3 dl = f3_sub(goal, mid);
    const float dv[3] = {dl.x, dl.y, dl.z};
    int order[3], n = 0;
    double max_v = -1.0, min_v = 1e9;
    for (int i = 0; i < 3; i++) {
        const double val = fabs((double)dv[i]);
        int at;
        if (val > max_v) { at = 0; max_v = val; }
        else if (val < min_v) { at = n; min_v = val; }
        else at = 1;
        for (int k = n; k > at; k--) order[k] = order[k - 1];
        order[at] = i;
        n++;
    }
    for (int i = 0; i < 3; i++) {
        const int off = dv[order[i]] > 0.0f ? 3 : 0;
        cand[i] = order[i] + off;
        cand[5 - i] = order[i] + (3 - off);
    }
}

// Largest k (power-of-two search, then refinement) such that the next k full round-robin cycles of expand_box are
// guaranteed to pass every slab test: the voxel region all those tests can touch — the committed box grown by k on
// each remaining candidate face, plus the one-voxel rim the +-1e-5 nudges reach — holds no blocked voxel, lies inside
// the map, and the grown box stays inside the world.
__device__ int free_cycles(const SfcCtx& c, const LatticeBox& box, unsigned cand_mask, int k_max) {
    auto passes = [&](int k) -> bool {
        int lo[3], hi[3];
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int kl = (cand_mask >> i) & 1 ? k : 0, kh = (cand_mask >> (i + 3)) & 1 ? k : 0;
            lo[i] = box.p[i] - kl - 1 - c.dm.off[i];
            hi[i] = box.p[i + 3] + kh - c.dm.off[i];
            if (lo[i] < 0 || hi[i] >= c.dm.size[i]) ok = false;
            // in-boundary test of the farthest slabs (1e-9 slack of isBoxInBoundary >> accumulated rounding)
            if (box.d[i] - kl * c.res <= c.wmin[i] - 1e-9 + 1e-11) ok = false;
            if (box.d[i + 3] + kh * c.res >= c.wmax[i] + 1e-9 - 1e-11) ok = false;
        }
        if (!ok) return false;
        return blocked_in_cells(c, lo, hi) == 0;
    };
    if (k_max < 1 || !passes(1)) return 0;
    int k = 1;
    while (2 * k <= k_max && passes(2 * k)) k *= 2;
    for (int step = k / 2; step >= 1; step /= 2)
        if (k + step <= k_max && passes(k + step)) k += step;
    return k;
}

// expandBoxFromPoint + expandSFCFromBox + expand_box: returns false when the seed box is blocked
__device__ bool expand_from_point(const SfcCtx& c, F3 point, F3 goal, double* out) {
    LatticeBox box, bc, bu;
    const float pt[3] = {point.x, point.y, point.z};
    for (int i = 0; i < 3; i++) {
        const double p = (double)pt[i];
        const double ratio = __ddiv_rn(p, c.res);
        const double rp = __dmul_rn(round(ratio), c.res);
        if (fabs(__dsub_rn(p, rp)) < 0.01) {
            box.d[i] = rp; box.d[i + 3] = rp;
            box.p[i] = box.p[i + 3] = (int)round(ratio);
        } else {
            box.d[i] = __dmul_rn(floor(ratio), c.res); box.d[i + 3] = __dmul_rn(ceil(ratio), c.res);
            box.p[i] = (int)floor(ratio); box.p[i + 3] = (int)ceil(ratio);
        }
    }
    if (obstacle_in_box(c, box)) return false;
    int cand[6], n_cand = 6;
    axis_candidates(box.d, goal, cand);
    auto propose = [&](int axis) {       // bc = committed box + one slab on `axis`, bu = that slab
        bu = bc;
        if (axis < 3) {
            bu.d[axis + 3] = bc.d[axis]; bu.p[axis + 3] = bc.p[axis];
            bc.d[axis] = __dsub_rn(bc.d[axis], c.res); bc.p[axis] -= 1;
            bu.d[axis] = bc.d[axis]; bu.p[axis] = bc.p[axis];
        } else {
            bu.d[axis - 3] = bc.d[axis]; bu.p[axis - 3] = bc.p[axis];
            bc.d[axis] = __dadd_rn(bc.d[axis], c.res); bc.p[axis] += 1;
            bu.d[axis] = bc.d[axis]; bu.p[axis] = bc.p[axis];
        }
    };
    int i = -1;
    while (n_cand > 0) {
        bc = box; bu = box;
        bool pending = false;       // true once bu is a proposed slab (the first test after an erase is the whole box)
        int cooldown = 0;
        unsigned mask = 0;
        for (int k = 0; k < n_cand; k++) mask |= 1u << cand[k];
        while (true) {
            if (pending && cooldown == 0) {
                // State: `box` committed, slab of cand[i] proposed. One round-robin cycle = n_cand passed tests moves
                // every candidate face one step and proposes cand[i]'s slab again. Skip k cycles that cannot fail.
                const int k = free_cycles(c, box, mask, 1 << 14);
                if (k > 0) {
                    for (int t = 0; t < n_cand; t++) {
                        const int axis = cand[t];
                        double f = box.d[axis];
                        for (int sidx = 0; sidx < k; sidx++) f = axis < 3 ? __dsub_rn(f, c.res) : __dadd_rn(f, c.res);
                        box.d[axis] = f;
                        box.p[axis] += axis < 3 ? -k : k;
                    }
                    bc = box;
                    propose(cand[i]);
                } else cooldown = n_cand;
            }
            if (cooldown > 0) cooldown--;
            if (obstacle_in_box(c, bu) || !box_in_boundary(c, bu.d)) break;
            i++;
            if (i >= n_cand) i = 0;
            box = bc;
            propose(cand[i]);
            pending = true;
        }
        if (i < 0) i = 0;     // unreachable: the seed box was tested above
        for (int k = i; k < n_cand - 1; k++) cand[k] = cand[k + 1];
        n_cand--;
        if (i > 0) i--; else i = n_cand - 1;
    }
    for (int k = 0; k < 6; k++) out[k] = box.d[k];
    return true;
}

__global__ void __launch_bounds__(128) k_sfc_expand(SfcLaunch L) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= L.n) return;
    SfcCtx c;
    c.dm = L.dm;
    c.res = L.res;
    c.lane = threadIdx.x & 31;
    for (int i = 0; i < 3; i++) {
        c.wmin[i] = (double)L.wmin[i]; c.wmax[i] = (double)L.wmax[i];
        c.wmin_nudge[i] = __dadd_rn(c.wmin[i], 1e-5);
    }
    const size_t tab = (size_t)(L.dm.size[0] + 1) * (L.dm.size[1] + 1) * (L.dm.size[2] + 1);
    double box[6];
    if (L.mode == 1) {
        c.sat = L.dm.sat + (size_t)L.sat_index[warp] * tab;
        const F3 p{L.point[3 * warp], L.point[3 * warp + 1], L.point[3 * warp + 2]};
        const F3 g{L.goal[3 * warp], L.goal[3 * warp + 1], L.goal[3 * warp + 2]};
        const bool ok = expand_from_point(c, p, g, box);
        if (c.lane < 6) L.box_out[6 * warp + c.lane] = ok ? (float)box[c.lane] : 0.0f;
        if (c.lane == 0) L.ok_out[warp] = ok ? 1 : 0;
        return;
    }
    const int a = L.agent_base + warp;
    c.sat = L.dm.sat + (size_t)L.consts[a].sat_index * tab;
    const lscgpu_agent_in& in = L.in[a];
    const F3 g{in.goal[0], in.goal[1], in.goal[2]};
    float* bx = L.boxes + (size_t)a * 30;
    const bool first = L.init_sfc[a] != 0;
    F3 seed;
    if (first) seed = F3{in.position[0], in.position[1], in.position[2]};
    else {
        const float* last = L.prev_traj + (size_t)a * kTrajFloats + (kM * 6 - 1) * 3;   // traj_curr[M-1][n]
        seed = F3{last[0], last[1], last[2]};
    }
    const bool ok = expand_from_point(c, seed, g, box);
    __syncwarp();
    if (first) {
        if (ok && c.lane < 30) bx[c.lane] = (float)box[c.lane % 6];
        if (c.lane == 0) L.init_sfc[a] = 0;
    } else {
        // window shift sfc[m] -> sfc[m-1], then the new box for the last segment
        float keep = 0.0f;
        if (c.lane < 24) keep = bx[c.lane + 6];
        __syncwarp();
        if (c.lane < 24) bx[c.lane] = keep;
        if (ok && c.lane < 6) bx[24 + c.lane] = (float)box[c.lane];
    }
    if (!ok && c.lane == 0) atomicOr(L.flags + a, LSCGPU_FLAG_SFC_SEED_BLOCKED);
}

void launch_sfc_expand(const SfcLaunch& L, cudaStream_t s) {
    if (L.n <= 0) return;
    const int blocks = (L.n * 32 + 127) / 128;
    k_sfc_expand<<<blocks, 128, 0, s>>>(L);
}

}  // namespace lscgpu
