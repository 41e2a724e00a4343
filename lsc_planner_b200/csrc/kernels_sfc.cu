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
// The expansion runs on integer lattice planes (plane P <-> coordinate P*res). For |coordinate| < 64 m the
// reference's float sample `(float)(box + it*res) +- 1e-5f` always lands in voxel P (+ nudge) or P-1 (- nudge): float32
// rounding there is < 4e-6 per operation, below the 1e-5 nudge, so the integer model reproduces OcTree::coordToKey of
// every sample exactly (the engine rejects larger worlds). The reference's doubles (faces moved by repeated +-res) are
// replayed at the end from the number of steps each face took, so the float32 output carries the reference's rounding.
struct IBox { int p[6]; };      // lattice planes: min xyz, max xyz

struct SfcCtx {
    DistMapDev dm;
    const int* sat;       // table of this seed's radius
    int nudge_min[3];     // face plane >= this  <=>  box[i] > world_min + 1e-5   (include/corridor_constructor.hpp:104)
    int bound_min[3];     // face plane >= this  <=>  box[i]   > world_min - 1e-9 (isBoxInBoundary, :124-131)
    int bound_max[3];     // face plane <= this  <=>  box[i+3] < world_max + 1e-9
    int lane;
};

__device__ __forceinline__ int sat_at(const SfcCtx& c, int x, int y, int z) {
    return c.sat[((size_t)x * (c.dm.size[1] + 1) + y) * (c.dm.size[2] + 1) + z];
}

// Voxel cells visited by isObstacleInBox along each axis (include/corridor_constructor.hpp:81-122): the it == 0
// sample (voxel below the min face unless the face sits on the world boundary) and the run lo+1 .. hi of the
// it >= 1 samples (voxel lo again for a flat box). Returns false when a sample leaves the map (getDistance = -1).
__device__ __forceinline__ bool sample_cells(const SfcCtx& c, const IBox& b, int* extra, int* lo, int* hi) {
    bool inside = true;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int pl = b.p[i], ph = b.p[i + 3];
        extra[i] = (pl >= c.nudge_min[i] ? pl - 1 : pl) - c.dm.off[i];
        if (ph - pl + 1 <= 1) { lo[i] = hi[i] = pl - c.dm.off[i]; }
        else { lo[i] = pl + 1 - c.dm.off[i]; hi[i] = ph - c.dm.off[i]; }
        const int n = c.dm.size[i];
        if (extra[i] < 0 || extra[i] >= n || lo[i] < 0 || hi[i] >= n) inside = false;
    }
    return inside;
}

// signed summed-volume terms of sub-boxes [sub0, sub0 + n_sub) of the 8 (extra | run)^3 sub-boxes, 8 corners each
__device__ __forceinline__ int blocked_terms(const SfcCtx& c, const int* extra, const int* lo, const int* hi, int sub0,
                                             int n_sub) {
    int sum = 0;
    for (int sub = sub0; sub < sub0 + n_sub; sub++) {
#pragma unroll
        for (int corner = 0; corner < 8; corner++) {
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
    }
    return sum;
}

__device__ __forceinline__ bool in_boundary(const SfcCtx& c, const IBox& b) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 3; i++) ok = ok && b.p[i] >= c.bound_min[i] && b.p[i + 3] <= c.bound_max[i];
    return ok;
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

// Largest k such that the next k full round-robin cycles of expand_box are guaranteed to pass every slab test: the
// voxel region all those tests can touch — the committed box grown by k on each remaining candidate face, plus the
// one-voxel rim the +-1e-5 nudges reach — holds no blocked voxel, lies inside the map, and the grown box stays inside
// the world. (Conservative: k = 0 merely means the cycles are walked test by test.)
__device__ int free_cycles(const SfcCtx& c, const IBox& box, unsigned cand_mask, int k_max) {
    auto passes = [&](int k) -> bool {
        int lo[3], hi[3];
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int kl = (cand_mask >> i) & 1 ? k : 0, kh = (cand_mask >> (i + 3)) & 1 ? k : 0;
            // lowest voxel any of those tests samples: the one below the (moved) min face, unless that face sits on
            // the world boundary (then the nudge is +1e-5 and the face's own voxel is sampled)
            const int pl = box.p[i] - kl;
            lo[i] = (pl >= c.nudge_min[i] ? pl - 1 : pl) - c.dm.off[i];
            hi[i] = box.p[i + 3] + kh - c.dm.off[i];
            if (lo[i] < 0 || hi[i] >= c.dm.size[i]) ok = false;
            if (box.p[i] - kl < c.bound_min[i] || box.p[i + 3] + kh > c.bound_max[i]) ok = false;
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

// expand_box state: `bc` = committed box plus the proposed slab, `bu` = the box under test, `i` = candidate index
struct Walk {
    IBox box, bc, bu;
    int i;
};
__device__ __forceinline__ int cand_at(unsigned packed, int k) { return (packed >> (3 * k)) & 7; }

// the reference's loop body after a passed test (include/corridor_constructor.hpp:204-221)
__device__ __forceinline__ void advance(Walk& w, unsigned cand, int n_cand) {
    w.i++;
    if (w.i >= n_cand) w.i = 0;
    const int axis = cand_at(cand, w.i);
    w.box = w.bc;
    w.bu = w.bc;
#pragma unroll
    for (int f = 0; f < 3; f++) {
        if (axis == f) { w.bu.p[f + 3] = w.bc.p[f]; w.bc.p[f] -= 1; w.bu.p[f] = w.bc.p[f]; }
        if (axis == f + 3) { w.bu.p[f] = w.bc.p[f + 3]; w.bc.p[f + 3] += 1; w.bu.p[f + 3] = w.bc.p[f + 3]; }
    }
}

// expandBoxFromPoint + expandSFCFromBox + expand_box: returns false when the seed box is blocked.
// Tests are evaluated 8 at a time: lane group g = lane/4 walks g steps ahead assuming the earlier tests pass, its four
// lanes share the 64 table reads of that test; the warp then commits the walk up to the first failing test. The table
// reads of all 8 tests are in flight together, so a step costs ~1/8 of an L2 round trip.
__device__ bool expand_from_point(const SfcCtx& c, double res, F3 point, F3 goal, double* out) {
    double seed_d[6];
    IBox seed;
    const float pt[3] = {point.x, point.y, point.z};
    for (int i = 0; i < 3; i++) {
        const double p = (double)pt[i];
        const double ratio = __ddiv_rn(p, res);
        const double rp = __dmul_rn(round(ratio), res);
        if (fabs(__dsub_rn(p, rp)) < 0.01) {
            seed_d[i] = rp; seed_d[i + 3] = rp;
            seed.p[i] = seed.p[i + 3] = (int)round(ratio);
        } else {
            seed_d[i] = __dmul_rn(floor(ratio), res); seed_d[i + 3] = __dmul_rn(ceil(ratio), res);
            seed.p[i] = (int)floor(ratio); seed.p[i + 3] = (int)ceil(ratio);
        }
    }
    {
        int extra[3], lo[3], hi[3];
        bool blocked = !sample_cells(c, seed, extra, lo, hi);
        if (!blocked) {
            const int part = blocked_terms(c, extra, lo, hi, c.lane >> 2, 1);   // 8 sub-boxes over lane groups ...
            blocked = warp_sum_int((c.lane & 3) == 0 ? part : 0) != 0;            // ... one lane of each group counts
        }
        if (blocked) return false;
    }
    int cand_list[6];
    axis_candidates(seed_d, goal, cand_list);
    unsigned cand = 0;
    for (int k = 0; k < 6; k++) cand |= (unsigned)cand_list[k] << (3 * k);
    int n_cand = 6;
    Walk w;
    w.box = seed; w.i = -1;
    const int g = c.lane >> 2, gl = c.lane & 3;
    while (n_cand > 0) {
        w.bc = w.box; w.bu = w.box;         // the first test after an erase is the whole box
        bool pending = false;
        int cooldown = 0;
        unsigned mask = 0;
        for (int k = 0; k < n_cand; k++) mask |= 1u << cand_at(cand, k);
        while (true) {
            if (pending && cooldown == 0) {
                // `box` committed, slab of cand[i] proposed. One round-robin cycle = n_cand passed tests: every
                // candidate face moves one step and cand[i]'s slab is proposed again. Skip k cycles that cannot fail.
                const int k = free_cycles(c, w.box, mask, 1 << 14);
                if (k > 0) {
                    for (int t = 0; t < n_cand; t++) {
                        const int axis = cand_at(cand, t);
#pragma unroll
                        for (int f = 0; f < 6; f++) if (axis == f) w.box.p[f] += f < 3 ? -k : k;
                    }
                    w.bc = w.box;
                    w.i = w.i == 0 ? n_cand - 1 : w.i - 1;
                    advance(w, cand, n_cand);          // re-propose cand[i]
                } else cooldown = 2 * n_cand;
            }
            // speculative walk: group g tests the box reached after g further passed tests
            Walk mine = w;
            for (int t = 0; t < g; t++) advance(mine, cand, n_cand);
            int extra[3], lo[3], hi[3];
            bool fail = !sample_cells(c, mine.bu, extra, lo, hi) || !in_boundary(c, mine.bu);
            int part = 0;
            if (!fail) part = blocked_terms(c, extra, lo, hi, gl * 2, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            fail = fail || part != 0;
            const unsigned fails = __ballot_sync(0xffffffffu, fail);
            // first failing test (bit 4g of group g)
            int first = 8;
            for (int t = 7; t >= 0; t--) if (fails & (1u << (4 * t))) first = t;
            for (int t = 0; t < first; t++) advance(w, cand, n_cand);
            if (first > 0) pending = true;
            if (cooldown > 0) cooldown = max(cooldown - first, 0);
            if (first < 8) break;
        }
        if (w.i < 0) w.i = 0;     // unreachable: the seed box was tested above
        // erase cand[i]
        {
            unsigned lowbits = cand & ((1u << (3 * w.i)) - 1u);
            unsigned high = cand >> (3 * (w.i + 1));
            cand = lowbits | (high << (3 * w.i));
        }
        n_cand--;
        if (w.i > 0) w.i--; else w.i = n_cand - 1;
    }
    // replay the reference's doubles: every face moved |steps| times by +-res
    if (c.lane < 6) {
        const int f = c.lane;
        int steps = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) if (k == f) steps = f < 3 ? seed.p[k] - w.box.p[k] : w.box.p[k] - seed.p[k];
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) if (k == f) v = seed_d[k];
        for (int sidx = 0; sidx < steps; sidx++) v = f < 3 ? __dsub_rn(v, res) : __dadd_rn(v, res);
        out[0] = v;             // lane f holds face f
    }
    return true;
}

__global__ void __launch_bounds__(128) k_sfc_expand(SfcLaunch L) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= L.n) return;
    SfcCtx c;
    c.dm = L.dm;
    c.lane = threadIdx.x & 31;
    for (int i = 0; i < 3; i++) {
        const double wmin = (double)L.wmin[i], wmax = (double)L.wmax[i];
        c.nudge_min[i] = (int)floor(__ddiv_rn(__dadd_rn(wmin, 1e-5), L.res)) + 1;
        c.bound_min[i] = (int)floor(__ddiv_rn(__dsub_rn(wmin, 1e-9), L.res)) + 1;
        c.bound_max[i] = (int)ceil(__ddiv_rn(__dadd_rn(wmax, 1e-9), L.res)) - 1;
    }
    const size_t tab = (size_t)(L.dm.size[0] + 1) * (L.dm.size[1] + 1) * (L.dm.size[2] + 1);
    double face = 0.0;       // lane f < 6: coordinate of face f
    if (L.mode == 1) {
        c.sat = L.dm.sat + (size_t)L.sat_index[warp] * tab;
        const F3 p{L.point[3 * warp], L.point[3 * warp + 1], L.point[3 * warp + 2]};
        const F3 g{L.goal[3 * warp], L.goal[3 * warp + 1], L.goal[3 * warp + 2]};
        const bool ok = expand_from_point(c, L.res, p, g, &face);
        if (c.lane < 6) L.box_out[6 * warp + c.lane] = ok ? (float)face : 0.0f;
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
    const bool ok = expand_from_point(c, L.res, seed, g, &face);
    __syncwarp();
    if (first) {
        // the first corridor is one box copied to all M segments (src/traj_planner.cpp:1454-1462)
        const float fv = __shfl_sync(0xffffffffu, (float)face, c.lane % 6);
        if (ok && c.lane < 30) bx[c.lane] = fv;
        if (c.lane == 0) L.init_sfc[a] = 0;
    } else {
        // window shift sfc[m] -> sfc[m-1], then the new box for the last segment
        float keep = 0.0f;
        if (c.lane < 24) keep = bx[c.lane + 6];
        __syncwarp();
        if (c.lane < 24) bx[c.lane] = keep;
        if (ok && c.lane < 6) bx[24 + c.lane] = (float)face;
    }
    if (!ok && c.lane == 0) atomicOr(L.flags + a, LSCGPU_FLAG_SFC_SEED_BLOCKED);
}

void launch_sfc_expand(const SfcLaunch& L, cudaStream_t s) {
    if (L.n <= 0) return;
    const int blocks = (L.n * 32 + 127) / 128;
    k_sfc_expand<<<blocks, 128, 0, s>>>(L);
}

}  // namespace lscgpu
