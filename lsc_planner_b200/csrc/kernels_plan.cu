// k_agent_plan — the whole per-agent replan of one synchronous step in ONE thread block (sm_100a):
//   phase 1  LSC construction against every neighbour (exact culling, GJK) — the rows stay in shared memory —
//            while one warp grows the agent's new SFC box against the distance-field tables;
//   phase 2  the Bernstein trajectory QP (qp_core.cuh) priced straight from those shared-memory rows;
//   epilogue result record (trajectory, advanced state, cost, counters) into the step's gather buffer.
// Also k_qp_batch (TrajOptimizer::solve for independent problems, rows from global memory) and k_qp_order (LPT order).
//
// Replaces, per agent (TrajPlanner::planLSC, src/traj_planner.cpp:389-425):
//   generateLSC + normalVectorBetweenPolys        src/traj_planner.cpp:1310-1407,2030-2043
//   closestPointsBetweenPointAndConvexHull / gjk  include/geometry.hpp:364-394, src/openGJK/openGJK.cpp:674-780
//   generateFeasibleSFC                           src/traj_planner.cpp:1442-1491 (sfc.cuh)
//   TrajOptimizer::solve                          src/traj_optimizer.cpp:31-154,239-539 (qp_core.cuh)
//   getStateFromControlPoints at t = dt           include/polynomial.hpp:63-121
//
// Why one kernel: the rows one agent's LSC phase produces (64 B per kept (neighbour, segment) pair, 250-2000 pairs) are
// consumed only by that agent's QP. Kept in shared memory they cost one 29-cycle LDS per pricing pass instead of an L2
// round trip, never touch HBM, and the QP of an agent starts the moment ITS corridors are done instead of after the
// slowest block of a separate corridor kernel. Blocks are issued longest-processing-time first (k_qp_order).
#include <climits>
#include <cstdlib>
#include <string>
#include <type_traits>

#include "gjk.cuh"
#include "kernels.hpp"
#include "qp_core.cuh"
#include "sfc.cuh"

namespace lscgpu {

constexpr int kBatchThreads = 256;       // k_qp_batch
constexpr int kMaxPlanWarps = 16;

// LSC-phase scratch (beside QpShared; the survivor queue aliases QpShared::Q, which the QP only touches afterwards)
struct LscShared {
    float own[kTrajFloats];
    float own_zs[30];
    float4 own_sphere[kM];
    float4 own_tsphere;
    float own_reach[kM];
    float own_reach_max;
    int warp_tot[2][kMaxPlanWarps];
    int sfc_ok;
    float sfc_box[6];           // the box the SFC warp grew in this step
    long long t_sfc, t_lsc;     // cycles until the SFC box was there / the LSC rows were done (diagnostics)
    int sfc_self;               // 1: the block grew the box itself
};

__host__ __device__ constexpr size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
struct PlanSmemLayout {
    size_t qp, lists, lsc, queue, nr, rhs, gate, seg, total;
    __host__ __device__ PlanSmemLayout(int cap, int threads, size_t qp_bytes = sizeof(QpShared)) {
        qp = 0;
        lists = align16(qp + qp_bytes);
        lsc = align16(lists + sizeof(int) * (threads / 32) * kWarpList);
        // survivor queue of the LSC phase: threads * (kM + 1) entries. Up to 256 threads it lives in Q / W (24 KB, untouched
        // until the QP starts); the 512-thread configuration gets its own region
        queue = align16(lsc + sizeof(LscShared));
        nr = align16(queue + (threads > 256 ? sizeof(int2) * (size_t)threads * (kM + 1) : 0));
        rhs = nr + sizeof(float4) * (size_t)cap;
        gate = rhs + sizeof(double2) * 3 * (size_t)cap;
        seg = gate + sizeof(double) * (size_t)cap;
        total = align16(seg + (size_t)cap);
    }
};
size_t agent_plan_smem_bytes(int row_cap, int threads) { return PlanSmemLayout(row_cap, threads).total; }
size_t agent_plan_slack_smem_bytes(int row_cap, int threads) { return PlanSmemLayout(row_cap, threads, sizeof(QpSharedSlack)).total; }

// ------------------------------------------------------------------------------------------------------------
// LSC phase of agent a by the kT threads of the block.
//   Phase A (one neighbour per thread per chunk of kT): exact culling test per (neighbour, segment) from the bounding
//     spheres — 80 B per neighbour, coalesced float4 reads. A pair is dropped only when its LSC rows cannot be violated
//     by any trajectory that respects the velocity / acceleration limits, so dropping is
//     solution-preserving; survivors go to a shared-memory queue.
//   Phase B (whenever the queue holds a full batch, and at the end): one queued (neighbour, segment) hull per thread —
//     GJK in FP64 registers, the LSC rows as one record at the pair's slot of the row source, together with the
//     smallest whitened slack of its rows at the unconstrained QP minimiser x0 (the QP does not look at a pair again
//     until the iterate has travelled that far).
// Returns the number of kept pairs (same value in all participating threads).
// ------------------------------------------------------------------------------------------------------------
// Slack kernels (SH::kE > 0): the pairs against obstacles of the agent's obs_slack_indices — every neighbour once the
// agent itself was reset, else the neighbours that ever were (src/traj_planner.cpp:866-878,1047-1061) — are marked
// kSlackUntouched, and their gate is the distance in the extended whitened space.
template <int kT, class SH>
__device__ __forceinline__ int lsc_phase(const PlanLaunch& L, int a, const RowSrc& rows, LscShared& X, const SH& S,
                                         int2* queue, int& gjk_it) {
    constexpr int kW = kT / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_obs = L.n_agents - 1;
    const AgentConstDev ca = L.consts[a];
    const float ra_f = (float)ca.radius, rdwa_f = (float)(ca.downwash * ca.radius);
    const double dw_self_a = (ca.downwash * ca.radius + ca.downwash * ca.radius) / (ca.radius + ca.radius);
    // queue length and number of survivors so far: uniform values every thread keeps in a register
    int q_count = 0, kept_total = 0, chunk = 0;

    for (int j0 = 0; j0 < n_obs; j0 += kT, chunk++) {
        const int jj = j0 + tid;
        unsigned keep_mask[kM];
        bool keep[kM];
#pragma unroll
        for (int m = 0; m < kM; m++) keep[m] = false;
        if (jj < n_obs) {
            const int j = jj < a ? jj : jj + 1;
            // downwash ratio of the pair in float: the test below is conservative by 1e-4 relative, float rounding is 1e-7
            const float2 rj = L.rdw[j];
            const float inv_dw = (ra_f + rj.x) / (rdwa_f + rj.y);
            const float smax = fmaxf(1.0f, inv_dw);
            const float rho = ra_f + rj.x;
            // coarse pass: the trajectory spheres contain every segment hull, so their gap bounds every segment pair's
            // hull distance from below; with the largest reach of the five segments the same test drops all five at once
            const float4 to = X.own_tsphere;
            const float4 tj = L.tsphere[j];
            const float tx = to.x - tj.x, ty = to.y - tj.y, tz = (to.z - tj.z) * inv_dw;
            const float t_lb = sqrtf(tx * tx + ty * ty + tz * tz) * 0.9999f - smax * (to.w + tj.w);
            if (!(t_lb - rho > 2.0f * smax * X.own_reach_max)) {
#pragma unroll
                for (int m = 0; m < kM; m++) {
                    const float4 so = X.own_sphere[m];
                    const float4 sj = L.sphere[(size_t)m * L.n_pad + j];
                    const float dx = so.x - sj.x, dy = so.y - sj.y, dz = (so.z - sj.z) * inv_dw;
                    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
                    const float d_lb = dist * 0.9999f - smax * (so.w + sj.w);        // lower bound of the hull distance
                    keep[m] = !(d_lb - rho > 2.0f * smax * X.own_reach[m]);
                }
            }
        }
        // order-preserving compaction (warp, then segment, then lane). A survivor's SLOT in the row store is its rank in
        // that order over the whole scan — the same for every block size and however the queue is drained, so results
        // do not depend on the launch configuration. One shared-memory exchange of the per-warp totals per chunk.
        int cnt[kM], wtot = 0;
#pragma unroll
        for (int m = 0; m < kM; m++) {
            keep_mask[m] = __ballot_sync(0xffffffffu, keep[m]);
            cnt[m] = __popc(keep_mask[m]);
            wtot += cnt[m];
        }
        if (lane == 0) X.warp_tot[chunk & 1][warp] = wtot;
        __syncthreads();
        {
            int off = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kW; w++) {
                const int c = X.warp_tot[chunk & 1][w];
                if (w < warp) off += c;
                total += c;
            }
#pragma unroll
            for (int m = 0; m < kM; m++) {
                const int r = off + __popc(keep_mask[m] & ((1u << lane) - 1u));
                if (keep[m]) queue[q_count + r] = make_int2(m * n_obs + jj, kept_total + r);
                off += cnt[m];
            }
            q_count += total;
            kept_total += total;
        }
        __syncthreads();
        // drain full batches (and everything after the last chunk)
        const bool last = j0 + kT >= n_obs;
        while (q_count >= kT || (last && q_count > 0)) {
            const int n_items = min(q_count, kT);
            int p = -1, slot = 0;
            if (tid < n_items) { const int2 it = queue[q_count - n_items + tid]; p = it.x; slot = it.y; }   // batch from the END
            q_count -= n_items;
            if (p >= 0) {
                const int m = p / n_obs, jj2 = p - m * n_obs;
                const int j = jj2 < a ? jj2 : jj2 + 1;
                const AgentConstDev cj = L.consts[j];
                const double downwash = (ca.downwash * ca.radius + cj.downwash * cj.radius) / (ca.radius + cj.radius);
                // pre-scaled z is valid when the pair's ratio equals both agents' own ratio bit for bit
                const bool pre = downwash == dw_self_a &&
                                 downwash == (cj.downwash * cj.radius + cj.downwash * cj.radius) / (cj.radius + cj.radius);
                F3 ow[6], ob[6];
                float obz[6];
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const int cp = m * 6 + i, e = cp * 3;
                    obz[i] = L.predT[(size_t)(e + 2) * L.n_pad + j];
                    ow[i] = F3{X.own[e], X.own[e + 1], pre ? X.own_zs[cp] : downwash_scaled_z(X.own[e + 2], downwash)};
                    ob[i] = F3{L.predT[(size_t)e * L.n_pad + j], L.predT[(size_t)(e + 1) * L.n_pad + j],
                               pre ? L.predZs[(size_t)cp * L.n_pad + j] : downwash_scaled_z(obz[i], downwash)};
                }
                LscSegment seg;
                lsc_segment_scaled(ow, ob, downwash, cj.radius + ca.radius, seg);
                gjk_it += seg.iterations;
                const double ax = (double)seg.normal.x, ay = (double)seg.normal.y, az = (double)seg.normal.z;
                const double an = sqrt(ax * ax + ay * ay + az * az);
                const float inv_an = an > 0.0 ? (float)(1.0 / an) : INFINITY;
                RowRec rec;
                rec.ax = seg.normal.x; rec.ay = seg.normal.y; rec.az = seg.normal.z; rec.inv_an = inv_an;
                double mu_min = INFINITY;
                bool soft = false;
                double isc2 = 0.0;
                if constexpr (SH::kE > 0) {
                    soft = L.reset_ever[a] != 0 || L.reset_ever[j] != 0;
                    isc2 = S.isc_m[m] * S.isc_m[m];
                }
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    // row  a . c_{m,i} >= d_i + a . o_{m,i}      (src/traj_optimizer.cpp:437-466)
                    const double rhs = seg.d[i] + (__dmul_rn(ax, (double)ob[i].x) + __dmul_rn(ay, (double)ob[i].y) +
                                                   __dmul_rn(az, (double)obz[i]));
                    rec.rhs[i] = rhs;
                    if (m == 0 && i < kPhi) continue;
                    const int vi = m * 6 + i;
                    const double slack = ax * S.x[vi] + ay * S.x[kAx + vi] + az * S.x[2 * kAx + vi] - rhs;
                    double mu = an > 0.0 ? slack * (double)inv_an * S.inv_gn[vi] : (slack < 0.0 ? -INFINITY : INFINITY);
                    if constexpr (SH::kE > 0) {
                        if (soft) {
                            const double sc = an > 0.0 ? (double)inv_an * S.inv_gn[vi] : INFINITY;
                            mu = slack * (sc < INFINITY ? sc * rsqrt(1.0 + isc2 * sc * sc) : rsqrt(isc2));
                        }
                    }
                    mu_min = fmin(mu_min, mu);
                }
                rows.store(slot, rec, m, p, mu_min > 0.0 ? mu_min * 0.999999 : mu_min, L.mirror_rows != 0, soft ? kSlackUntouched : 0);
            }
        }
        // the queue tail that stays for the next chunk is only read after that chunk's barriers; nothing to wait for here
    }
    return kept_total;
}

// ------------------------------------------------------------------------------------------------------------
// kSlack: the instantiation with slack coordinates (QpSharedSlack). Which one plans the step is a property of the whole
// swarm — once any agent was reset, every planner's obs_slack_indices holds it — so both kernels are launched every step
// and the one the step does not need returns at once (*L.any_reset, set by k_predict).
template <int kPlanThreads, bool kSfc, bool kSlack>
__global__ void __launch_bounds__(kPlanThreads, kSlack ? (kPlanThreads > 256 ? 1 : 2) : 512 / kPlanThreads) k_agent_plan(PlanLaunch L) {
    using SH = typename std::conditional<kSlack, QpSharedSlack, QpShared>::type;
    if (L.any_reset && (*L.any_reset != 0) != kSlack) return;
    const int row_cap = kSlack ? L.row_cap_slack : L.row_cap;
    constexpr int kPlanWarps = kPlanThreads / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    const PlanSmemLayout lay(row_cap, kPlanThreads, sizeof(SH));
    SH& S = *reinterpret_cast<SH*>(smem + lay.qp);
    int* open_lists = reinterpret_cast<int*>(smem + lay.lists);
    LscShared& X = *reinterpret_cast<LscShared*>(smem + lay.lsc);
    const long long t_start = clock64();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bi = blockIdx.x;
    // scheduling order != data order: blocks are dispatched in blockIdx order, so the agents with the most expensive
    // plan of the previous step go first (longest-processing-time-first)
    const int a = L.order ? L.order[L.order_first + bi * L.order_stride] : L.agent_base + bi * L.agent_stride;
    const QpTablesDev& T = *L.T;
    const int ts = L.ts[a];
    const int n_obs = L.n_agents - 1;

    RowSrc rows;
    rows.s_nr = reinterpret_cast<float4*>(smem + lay.nr);
    rows.s_rhs = reinterpret_cast<double2*>(smem + lay.rhs);
    rows.s_gate = reinterpret_cast<double*>(smem + lay.gate);
    rows.s_seg = smem + lay.seg;
    rows.cap = row_cap;
    rows.g_rows = L.rows + (size_t)bi * L.P_pad;
    rows.g_gate = L.safe + (size_t)bi * L.P_pad;
    rows.g_kept = L.kept + (size_t)bi * L.P_pad;
    rows.n_obs = n_obs;

    // ---- stage: QP tables + x0 (bounds come later, with the SFC box), own prediction ----------------------------
    const double* st = L.state9 + (size_t)a * 9;
    const double* gl = L.goal3 + (size_t)a * 3;
    for (int e = tid; e < kTrajFloats; e += kPlanThreads) X.own[e] = L.pred[(size_t)a * kTrajFloats + e];
    if (tid < 30) X.own_zs[tid] = L.predZs[(size_t)tid * L.n_pad + a];
    if (tid < kM) { X.own_sphere[tid] = L.sphere[(size_t)tid * L.n_pad + a]; X.own_reach[tid] = L.reach[(size_t)a * kM + tid]; }
    if (tid == 32) {
        X.own_tsphere = L.tsphere[a];
        float rm = 0.f;
        for (int m = 0; m < kM; m++) rm = fmaxf(rm, L.reach[(size_t)a * kM + m]);
        X.own_reach_max = rm;
    }
    if (tid == 0) { X.sfc_ok = 1; X.sfc_self = 0; X.t_sfc = 0; }
    qp_stage<kPlanThreads>(S, T, ts, st, gl, nullptr, L.wmin, L.wmax, L.consts[a], L.slack_w);     // ends with a block barrier

    // ---- phase 1: corridors -------------------------------------------------------------------------------------
    int n_kept = 0, gjk_it = 0;
    int2* queue = kPlanThreads > 256 ? reinterpret_cast<int2*>(smem + lay.queue)
                                     : reinterpret_cast<int2*>(S.Q);        // kPlanThreads * (kM + 1) entries <= Q + W
    if (n_obs > 0) n_kept = lsc_phase<kPlanThreads, SH>(L, a, rows, X, S, queue, gjk_it);
    if (tid == 0) X.t_lsc = clock64() - t_start;
    if (kSfc && warp == 0) {
        // The step's new SFC box comes from k_sfc_step, launched at the start of the step beside k_predict. It is
        // normally there by now; a block of the first wave may be early: it waits a bounded time (the walk takes
        // 20-50 us), then grows the box itself — it never depends on another kernel's progress.
        const int epoch = *L.epoch;
        bool have = false;
        const long long t_wait = clock64();
        while (true) {
            int ready = 0;
            if (lane == 0) ready = *(volatile const int*)(L.sfc_ready + a) == epoch;
            ready = __shfl_sync(0xffffffffu, ready, 0);
            if (ready) { have = true; break; }
            if (clock64() - t_wait > L.sfc_wait_cycles) break;
            __nanosleep(512);
        }
        if (have) {
            __threadfence();
            if (lane < 6) X.sfc_box[lane] = __ldcg(L.sfc_box_g + (size_t)a * 6 + lane);
            if (lane == 0) X.sfc_ok = __ldcg(L.sfc_ok_g + a);
        } else {
            SfcCtx c;
            sfc_ctx_init(c, L.dm, L.wmin, L.wmax, L.res, lane);
            double face;
            const F3 cur_goal{(float)L.goal3[(size_t)a * 3], (float)L.goal3[(size_t)a * 3 + 1], (float)L.goal3[(size_t)a * 3 + 2]};
            const bool ok = sfc_agent_box(c, L.dm, L.consts[a].sat_index, L.res, L.in[a], cur_goal, L.prev_traj + (size_t)a * kTrajFloats,
                                          L.init_sfc[a] != 0, face);
            if (lane < 6) X.sfc_box[lane] = ok ? (float)face : 0.0f;
            if (lane == 0) { X.sfc_ok = ok ? 1 : 0; X.sfc_self = 1; }
        }
        if (lane == 0) X.t_sfc = clock64() - t_start;
    }
    if (tid == 0) X.warp_tot[0][0] = n_kept;
    __syncthreads();
    n_kept = X.warp_tot[0][0];
    if (L.counters) {
        const int tot = warp_sum_int(gjk_it);
        if (lane == 0 && tot) atomicAdd(&L.counters->gjk_iterations, (unsigned long long)tot);
        if (tid == 0) atomicAdd(&L.counters->kept_pairs, (unsigned long long)n_kept);
    }
    if (tid == 0) {
        L.kept_count[bi] = n_kept;
        if (L.block_of) L.block_of[a] = bi;
        if (L.kept_step) atomicAdd(L.kept_step, n_kept);
    }
    // bounds: world box, intersected with the agent's SFC window as it stands after this step's new box
    if (tid < 15) {
        const int m = tid / 3, k = tid % 3;
        double lo = (double)L.wmin[k], hi = (double)L.wmax[k];
        if (kSfc) {         // SFC rows == per-variable bounds (src/traj_optimizer.cpp:409-434)
            const float* old_win = L.boxes + (size_t)a * 30;
            const bool first = L.init_sfc[a] != 0, ok = X.sfc_ok != 0;
            lo = fmax(lo, (double)sfc_window_elem(old_win, first, ok, X.sfc_box, m * 6 + k));
            hi = fmin(hi, (double)sfc_window_elem(old_win, first, ok, X.sfc_box, m * 6 + 3 + k));
        }
        S.lb[tid] = lo; S.ub[tid] = hi;
    }
    // warm-start candidates: the bounds / dynamic-limit rows active at this agent's previous solve (none after a failed
    // solve, a state reset, or in the slack instantiation)
    int n_guess = 0;
    if (!kSlack && L.act_prev && !((L.flags ? L.flags[a] : 0) & LSCGPU_FLAG_SLACK_NEEDED)) {
        const unsigned short* ap = L.act_prev + (size_t)a * kActSlots;
        n_guess = min((int)ap[kActSlots - 1], kActSlots - 1);
        if (tid < n_guess) S.act[tid] = ap[tid];
    }
    // Q / W start empty (the queue lived in Q)
    __syncthreads();
    const long long t_lsc = clock64();

    // ---- phase 2: the QP ----------------------------------------------------------------------------------------
#ifdef LSCGPU_QP_SECTION_TIMERS
    long long sec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long* secp = sec;
#else
    long long* secp = nullptr;
#endif
    const QpResultRegs R = qp_solve_core<kPlanThreads>(S, open_lists, rows, n_kept, T.vel_coef, T.acc_coef, L.max_iter, secp,
                                                       L.mirror_rows != 0, n_guess);

    // ---- epilogue -----------------------------------------------------------------------------------------------
    if (L.counters) {
        const unsigned long long ev = (unsigned long long)(warp_sum((double)R.pairs_evaluated) + 0.5);
        if (lane == 0) atomicAdd(&L.counters->rows_priced, 6ull * ev + (warp == 0 ? 414ull * R.passes : 0ull));
    }
    if (warp != 0) return;
    const double cost = qp_objective(S, T, ts, gl, lane);
    if (L.counters && lane == 0) {
        atomicAdd(&L.counters->qp_iterations, (unsigned long long)R.iters);
        atomicAdd(&L.counters->full_passes, R.passes);
        if (n_guess > 0) {
            atomicAdd(&L.counters->warm_tried, 1ull);
            if (R.warm > 0) { atomicAdd(&L.counters->warm_accepted, 1ull); atomicAdd(&L.counters->warm_rows, (unsigned long long)R.warm); }
        }
    }
    // direct exchange: the step's slots live in the parity of its epoch
    const int xparity = L.px.peers ? (*L.epoch & 1) : 0;
    GatherSlot& slot_out = L.out[(size_t)xparity * L.px.slots + L.out_base + bi];
    lscgpu_agent_out& o = slot_out.rec;
    float* tr = &o.traj[0][0][0];
    const bool ok = R.status == LSCGPU_QP_OK;
    {   // next step's warm-start candidates: the fixed rows of the working set, in working-set order
        int id0 = -1, id1 = -1;
        if (ok && !kSlack) {
            if (lane < R.q) id0 = S.act[lane];
            if (lane + 32 < R.q) id1 = S.act[lane + 32];
        }
        const bool k0 = id0 >= 0 && id0 < kFixedRows, k1 = id1 >= 0 && id1 < kFixedRows;
        const unsigned m0 = __ballot_sync(0xffffffffu, k0), m1 = __ballot_sync(0xffffffffu, k1);
        if (k0) slot_out.act[__popc(m0 & ((1u << lane) - 1u))] = (unsigned short)id0;
        if (k1) slot_out.act[__popc(m0) + __popc(m1 & ((1u << lane) - 1u))] = (unsigned short)id1;
        if (lane == 0) slot_out.act[kActSlots - 1] = (unsigned short)(__popc(m0) + __popc(m1));
    }
    // failure: the optimizer keeps its last successful trajectory and cost (src/traj_planner.cpp:1553-1584)
    for (int e = lane; e < kTrajFloats; e += 32) {
        const int axis = e % 3, cp = e / 3;
        tr[e] = ok ? (float)S.x[axis * kAx + cp] : L.prev_traj[(size_t)a * kTrajFloats + e];
    }
    __syncwarp();
    __threadfence_block();
    if (lane < 3) {
        // getStateFromControlPoints at t = dt: segment 1, local time 0 (include/polynomial.hpp:63-121)
        const float inv_dt = (float)(1.0 / T.dt);
        const float c0 = tr[(6 + 0) * 3 + lane], c1 = tr[(6 + 1) * 3 + lane], c2 = tr[(6 + 2) * 3 + lane];
        const float v0 = __fmul_rn(__fmul_rn(__fsub_rn(c1, c0), (float)kN), inv_dt);
        const float v1 = __fmul_rn(__fmul_rn(__fsub_rn(c2, c1), (float)kN), inv_dt);
        o.next_position[lane] = c0;
        o.next_velocity[lane] = v0;
        o.next_acceleration[lane] = __fmul_rn(__fmul_rn(__fsub_rn(v1, v0), (float)(kN - 1)), inv_dt);
    }
    if (lane == 0) {
        const double c_rep = ok ? cost : L.last_cost[a];
        o.agent_id = a;
        o.qp_cost = c_rep;
        L.last_cost[a] = c_rep;
        o.report = LSCGPU_REPORT_SUCCESS;
        o.qp_status = R.status;
        o.qp_iterations = R.iters;
        o.qp_active = R.q;
        int fl = L.flags ? L.flags[a] : 0;
        if (kSfc && !X.sfc_ok) fl |= LSCGPU_FLAG_SFC_SEED_BLOCKED;
        if (R.warm > 0) fl |= LSCGPU_FLAG_WARM_START;
        if (ok && R.worst_slack < -1e-9) fl |= LSCGPU_FLAG_IN_BAND;
        if constexpr (kSlack) {
            fl |= LSCGPU_FLAG_SLACK_MODE;
            bool used = false;
            for (int c = 0; c < S.n_e; c++) used |= S.e[c] < 0.0;
            if (used && ok) fl |= LSCGPU_FLAG_SLACK_USED;
            if (S.overflow) fl |= LSCGPU_FLAG_SLACK_OVERFLOW;
        }
        o.flags = fl;
        o.terminal_segments = ts;
        for (int k = 0; k < 3; k++) o.current_goal[k] = (float)gl[k];
        o.goal_kind = L.goal_kind ? L.goal_kind[a] : 0;
        for (int f = 0; f < 6; f++) o.sfc_box[f] = kSfc ? X.sfc_box[f] : 0.0f;
        o.sfc_in_block = kSfc ? X.sfc_self : 0;
        o.qp_sweeps = (int)R.passes;
        const long long t_end = clock64();
        o.qp_kcycles = (int)((t_end - t_start) >> 10);
        o.qp_price_kcycles = (int)(R.price_cycles >> 10);
        o.lsc_kcycles = (int)((t_lsc - t_start) >> 10);
        o.lsc_pairs_kept = n_kept;
#ifdef LSCGPU_QP_SECTION_TIMERS
        if (L.dbg) {
            for (int i = 1; i < 7; i++) L.dbg[(size_t)bi * 10 + i] = sec[i];
            L.dbg[(size_t)bi * 10] = R.price_cycles;
            L.dbg[(size_t)bi * 10 + 7] = t_lsc - t_start;
            L.dbg[(size_t)bi * 10 + 8] = X.t_sfc;
            L.dbg[(size_t)bi * 10 + 9] = sec[7];
        }
#endif
    }
    if (L.host_out) {
        lscgpu_agent_out* const h = *L.host_out;
        if (h) {
            // the caller's result array is mapped host memory: the record goes there now (posted writes over PCIe), not in a
            // copy after the step
            __syncwarp();
            constexpr int kRecVec = (int)(sizeof(lscgpu_agent_out) / 16);
            static_assert(kRecVec <= 32, "record larger than one vector per lane");
            if (lane < kRecVec) reinterpret_cast<uint4*>(h + a)[lane] = reinterpret_cast<const uint4*>(&o)[lane];
        }
    }
    if (L.px.peers) {
        // The finished slot goes straight into every peer's exchange buffer (stores through the NVLink peer mapping, 16 bytes
        // per lane), then the peer's arrival counter for this rank is bumped: the transfer of one agent's record overlaps the
        // planning of the others, and k_commit on the peer starts as soon as the last one has landed.
        __syncwarp();
        constexpr int kVec = (int)(sizeof(GatherSlot) / 16);
        const uint4* src = reinterpret_cast<const uint4*>(&slot_out);
        uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
        if (lane < kVec) v0 = src[lane];
        if (lane + 32 < kVec) v1 = src[lane + 32];
        static_assert(kVec <= 64, "slot larger than two vectors per lane");
        for (int r = 0; r < L.px.n_ranks; r++) {
            if (r == L.px.rank) continue;
            uint4* dst = reinterpret_cast<uint4*>(L.px.peers[r] + (size_t)xparity * L.px.slots + L.out_base + bi);
            if (lane < kVec) dst[lane] = v0;
            if (lane + 32 < kVec) dst[lane + 32] = v1;
        }
        __threadfence_system();
        __syncwarp();
        if (lane < L.px.n_ranks) atomicAdd_system(L.px.counters(lane) + L.px.rank, 1);
    }
}

// opt in to > 48 KB of dynamic shared memory and the largest shared-memory carve-out (two blocks per SM); per device
template <int kT, bool kSfc, bool kSlack>
static cudaError_t configure_one() {
    cudaError_t rc = cudaFuncSetAttribute(k_agent_plan<kT, kSfc, kSlack>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (rc == cudaSuccess) rc = cudaFuncSetAttribute(k_agent_plan<kT, kSfc, kSlack>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    return rc;
}
cudaError_t configure_qp_batch();
cudaError_t configure_agent_plan() {
    cudaError_t rc = configure_one<256, true, false>();
    if (rc == cudaSuccess) rc = configure_one<256, false, false>();
    if (rc == cudaSuccess) rc = configure_one<128, true, false>();
    if (rc == cudaSuccess) rc = configure_one<128, false, false>();
    if (rc == cudaSuccess) rc = configure_one<512, true, false>();
    if (rc == cudaSuccess) rc = configure_one<512, false, false>();
    if (rc == cudaSuccess) rc = configure_one<256, true, true>();
    if (rc == cudaSuccess) rc = configure_one<256, false, true>();
    if (rc == cudaSuccess) rc = configure_qp_batch();
    return rc;
}

void launch_agent_plan(const PlanLaunch& L, cudaStream_t s) {
    if (L.n_blocks <= 0) return;
    const size_t smem = agent_plan_smem_bytes(L.row_cap, L.threads);
    if (L.threads == 128) {
        if (L.use_sfc) k_agent_plan<128, true, false><<<L.n_blocks, 128, smem, s>>>(L);
        else k_agent_plan<128, false, false><<<L.n_blocks, 128, smem, s>>>(L);
    } else if (L.threads == 512) {
        if (L.use_sfc) k_agent_plan<512, true, false><<<L.n_blocks, 512, smem, s>>>(L);
        else k_agent_plan<512, false, false><<<L.n_blocks, 512, smem, s>>>(L);
    } else {
        if (L.use_sfc) k_agent_plan<256, true, false><<<L.n_blocks, 256, smem, s>>>(L);
        else k_agent_plan<256, false, false><<<L.n_blocks, 256, smem, s>>>(L);
    }
    if (L.any_reset) {
        // the slack instantiation (256 threads, two blocks per SM): returns at once unless some agent was ever reset
        const size_t smem_s = agent_plan_slack_smem_bytes(L.row_cap_slack, 256);
        if (L.use_sfc) k_agent_plan<256, true, true><<<L.n_blocks, 256, smem_s, s>>>(L);
        else k_agent_plan<256, false, true><<<L.n_blocks, 256, smem_s, s>>>(L);
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_qp_batch — TrajOptimizer::solve (src/traj_optimizer.cpp:31-154) for independent problems whose LSC rows arrive
// from the host in the reference's container layout (k_rows_from_lsc); rows are priced from global memory.
// ------------------------------------------------------------------------------------------------------------
// kSlack: problems whose CollisionConstraints carry obs_slack_indices (slot codes set by k_rows_from_lsc); eps_out
// receives the slack variable of every (obstacle, segment).
template <bool kSlack>
__global__ void __launch_bounds__(kBatchThreads, 2) k_qp_batch(QpBatchLaunch L) {
    using SH = typename std::conditional<kSlack, QpSharedSlack, QpShared>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    SH& S = *reinterpret_cast<SH*>(smem);
    int* open_lists = reinterpret_cast<int*>(smem + align16(sizeof(SH)));
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int agent = L.agent_index[b];
    const QpTablesDev& T = *L.T;
    const int ts = L.ts[b];
    const size_t pair0 = (size_t)kPairsPerObs * L.obs_offset[b];
    RowSrc rows;
    rows.s_nr = nullptr; rows.s_rhs = nullptr; rows.s_gate = nullptr; rows.s_seg = nullptr; rows.cap = 0;
    rows.g_rows = L.rows + pair0; rows.g_gate = L.safe + pair0; rows.g_kept = L.kept + pair0;
    rows.n_obs = L.obs_offset[b + 1] - L.obs_offset[b];
    const double* gl = L.goal3 + (size_t)b * 3;
    qp_stage<kBatchThreads>(S, T, ts, L.state9 + (size_t)b * 9, gl, L.boxes ? L.boxes + (size_t)b * 30 : nullptr, L.wmin, L.wmax,
                           L.consts[agent], L.slack_w);
    const QpResultRegs R = qp_solve_core<kBatchThreads>(S, open_lists, rows, L.kept_count[b], T.vel_coef, T.acc_coef, L.max_iter, nullptr,
                                                        /*mirror_rows=*/true);
    if (warp != 0) return;
    const double cost = qp_objective(S, T, ts, gl, lane);
    for (int e = lane; e < kNv; e += 32) L.x_out[(size_t)b * kNv + e] = S.x[e];
    if (lane == 0) {
        L.cost_out[b] = cost; L.iters_out[b] = R.iters;
        int status = R.status;
        if constexpr (kSlack) { if (S.overflow) status |= 0x100; }      // the host maps this to its own error text
        L.status_out[b] = status;
    }
    if constexpr (kSlack) {
        if (L.eps_out) {
            // slot of a pair inside the problem: m * n_b + o  ->  eps_out[(obs_offset[b] + o) * 5 + m]
            const int n_b = rows.n_obs;
            for (int c = lane; c < S.n_e; c += 32) {
                const int slot = S.e_slot[c], m = slot / n_b, o = slot - m * n_b;
                L.eps_out[((size_t)L.obs_offset[b] + o) * kM + m] = S.e[c] * S.e_isc[c];
            }
        }
    }
}

static size_t qp_batch_smem(bool slack) {
    return align16(slack ? sizeof(QpSharedSlack) : sizeof(QpShared)) + sizeof(int) * (kBatchThreads / 32) * kWarpList;
}
cudaError_t configure_qp_batch() {
    return cudaFuncSetAttribute(k_qp_batch<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qp_batch_smem(true));
}
void launch_qp_batch(const QpBatchLaunch& L, cudaStream_t s) {
    if (L.n_problems <= 0) return;
    if (L.slack) k_qp_batch<true><<<L.n_problems, kBatchThreads, qp_batch_smem(true), s>>>(L);
    else k_qp_batch<false><<<L.n_problems, kBatchThreads, qp_batch_smem(false), s>>>(L);
}

// ------------------------------------------------------------------------------------------------------------
// k_qp_order — longest-processing-time-first order of agents a0 .. a0+n-1 from the cycle count each plan recorded in
// its result record at the previous step. Rank sort (position = number of agents with a larger key, ties by id), so
// the permutation is a pure function of the records: every rank of a multi-GPU job derives the same one from its
// replica and takes every G-th entry. It only affects scheduling, never results.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_qp_order(int n, int a0, const lscgpu_agent_out* res, int* order) {
    // 32 agents per block (one per lane); the 8 warps split every tile of keys between them
    constexpr int kTile = 2048;
    __shared__ int tile[kTile];
    __shared__ int part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const int mine = i < n ? res[a0 + i].qp_kcycles : 0;
    int pos = 0;
    for (int j0 = 0; j0 < n; j0 += kTile) {
        const int cnt = min(kTile, n - j0);
        for (int t = threadIdx.x; t < cnt; t += 256) tile[t] = res[a0 + j0 + t].qp_kcycles;
        __syncthreads();
        const int per = (cnt + 7) / 8, t0 = warp * per, t1 = min(cnt, t0 + per);
        for (int t = t0; t < t1; t++) {
            const int other = tile[t];
            pos += (other > mine) || (other == mine && j0 + t < i);
        }
        __syncthreads();
    }
    part[warp][lane] = pos;
    __syncthreads();
    if (warp == 0 && i < n) {
        int p = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) p += part[w][lane];
        order[p] = a0 + i;
    }
}
void launch_qp_order(int n, int a0, const lscgpu_agent_out* res, int* order, cudaStream_t s) {
    if (n > 0) k_qp_order<<<(n + 31) / 32, 256, 0, s>>>(n, a0, res, order);
}

}  // namespace lscgpu
