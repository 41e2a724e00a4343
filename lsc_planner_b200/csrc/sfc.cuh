// SFC box expansion against the blocked-voxel summed-volume tables, one warp per seed (device functions shared by
// k_sfc_expand and the SFC warp of k_agent_plan).
//
// Replaces (reference paths):
//   CorridorConstructor::expandBoxFromPoint / expandSFCFromBox / expand_box / setAxisCand / isObstacleInBox /
//       isBoxInBoundary                                include/corridor_constructor.hpp:18-44,234-245,184-232,142-182,81-131
//   TrajPlanner::generateFeasibleSFC (window shift)     src/traj_planner.cpp:1451-1491
#pragma once
#include "kernels.hpp"

namespace lscgpu {

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

// setAxisCand (include/corridor_constructor.hpp:142-182): faces toward the goal first, largest offset first
__device__ __forceinline__ void axis_candidates(const double* box, F3 goal, int* cand) {
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
// the world. (Conservative: k = 0 merely means the cycles are walked test by test.) The predicate is monotone in k (the
// region only grows), so the warp searches 32-ary: every lane tests its own k with its own 8 table reads, all in
// flight together, and a ballot finds the first failure — three rounds (strides 1024, 32, 1) cover any world the engine
// accepts, the first of which fails on the world bounds without touching memory.
static __device__ int free_cycles(const SfcCtx& c, const IBox& box, unsigned cand_mask, int k_max) {
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
        // blocked voxels in the cell box [lo, hi]: inclusion-exclusion over the 8 corners of the summed-volume table
        int v = 0;
#pragma unroll
        for (int corner = 0; corner < 8; corner++) {
            const int cx = (corner & 1) ? hi[0] + 1 : lo[0];
            const int cy = (corner & 2) ? hi[1] + 1 : lo[1];
            const int cz = (corner & 4) ? hi[2] + 1 : lo[2];
            const int t = sat_at(c, cx, cy, cz);
            v += (__popc(corner) & 1) ? t : -t;
        }
        return v == 0;
    };
    int cur = 0;
#pragma unroll 1
    for (int stride = 1024; stride >= 1; stride >>= 5) {
        const int k = cur + (c.lane + 1) * stride;
        const bool ok = k <= k_max && passes(k);
        const unsigned fails = ~__ballot_sync(0xffffffffu, ok);
        const int n_pass = fails ? __ffs(fails) - 1 : 32;        // lanes before the first failure
        cur += n_pass * stride;
    }
    return min(cur, k_max);
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
static __device__ bool expand_from_point(const SfcCtx& c, double res, F3 point, F3 goal, double* out) {
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
        unsigned mask = 0;
        for (int k = 0; k < n_cand; k++) mask |= 1u << cand_at(cand, k);
        // That whole-box test, by all lanes together (its 8 sub-boxes over the lane groups). A failure erases the next
        // candidate without any growth, exactly as the reference's loop does.
        bool box_fail;
        {
            int extra[3], lo[3], hi[3];
            box_fail = !sample_cells(c, w.bu, extra, lo, hi) || !in_boundary(c, w.bu);
            int part = 0;
            if (!box_fail) part = blocked_terms(c, extra, lo, hi, g, 1);
            box_fail = box_fail || warp_sum_int(gl == 0 ? part : 0) != 0;
        }
        if (!box_fail) {
            advance(w, cand, n_cand);           // passed: the first candidate's slab is proposed
            // `box` committed, slab of cand[i] proposed. One round-robin cycle = n_cand passed tests: every candidate
            // face moves one step and cand[i]'s slab is proposed again. Skip the k cycles that cannot fail; the failing
            // test then lies within the next cycle, which one speculative round (8 tests) covers.
            int cooldown = 0;
            while (true) {
                if (cooldown == 0) {
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
                    }
                    cooldown = 2 * n_cand;
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
                cooldown = max(cooldown - first, 0);
                if (first < 8) break;
            }
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


// world-derived lattice limits of the walk
__device__ __forceinline__ void sfc_ctx_init(SfcCtx& c, const DistMapDev& dm, const float* wmin_f, const float* wmax_f, double res, int lane) {
    c.dm = dm;
    c.lane = lane;
    for (int i = 0; i < 3; i++) {
        const double wmin = (double)wmin_f[i], wmax = (double)wmax_f[i];
        c.nudge_min[i] = (int)floor(__ddiv_rn(__dadd_rn(wmin, 1e-5), res)) + 1;
        c.bound_min[i] = (int)floor(__ddiv_rn(__dsub_rn(wmin, 1e-9), res)) + 1;
        c.bound_max[i] = (int)ceil(__ddiv_rn(__dadd_rn(wmax, 1e-9), res)) - 1;
    }
}

// generateFeasibleSFC of one agent by one warp (src/traj_planner.cpp:1451-1491): grows the step's new box — from the
// current position at the first step (:1454-1462), afterwards from traj_curr[M-1][n] (:1471-1480) — toward the current
// goal. Lane f < 6 returns face f in `face`. Returns false when the seed is blocked (the reference throws,
// include/corridor_constructor.hpp:35-38). The persistent window itself is updated by sfc_window_box below.
__device__ __forceinline__ bool sfc_agent_box(SfcCtx& c, const DistMapDev& dm, int sat_index, double res, const lscgpu_agent_in& in,
                                              F3 g /* Agent::current_goal_position */, const float* prev_traj_a, bool first, double& face) {
    const size_t tab = (size_t)(dm.size[0] + 1) * (dm.size[1] + 1) * (dm.size[2] + 1);
    c.sat = dm.sat + (size_t)sat_index * tab;
    F3 seed;
    if (first) seed = F3{in.position[0], in.position[1], in.position[2]};
    else {
        const float* last = prev_traj_a + (kM * 6 - 1) * 3;   // traj_curr[M-1][n]
        seed = F3{last[0], last[1], last[2]};
    }
    face = 0.0;
    const bool ok = expand_from_point(c, res, seed, g, &face);
    __syncwarp();
    return ok;
}

// Element e = m * 6 + f of the agent's window AFTER this step, from the window before it: the first corridor is the new
// box copied to all M segments (src/traj_planner.cpp:1454-1462); afterwards the window shifts, sfc[m] -> sfc[m-1], and
// the new box becomes the last segment's (:1465-1480). A blocked seed leaves the first corridor as it was, and keeps
// the previous last box after the shift. The same rule is applied by k_agent_plan (bounds of the step's QP) and by
// k_commit (persistent window of every replica).
__device__ __forceinline__ float sfc_window_elem(const float* old_win, bool first, bool ok, const float* new_box, int e) {
    const int m = e / 6, f = e - m * 6;
    if (first) return ok ? new_box[f] : old_win[e];
    if (m < kM - 1) return old_win[e + 6];
    return ok ? new_box[f] : old_win[e];
}

}  // namespace lscgpu
