// k_qp_solve — the Bernstein trajectory QP, one agent per thread block (sm_100a, FP64): warp 0 runs the dual
// active-set iteration, all warps of the block price the rows.
//
// Replaces TrajOptimizer::solve (src/traj_optimizer.cpp:31-154): buildDeq (:239-259), populatebyrow (:261-539) and the
// CPLEX dual-simplex call (:76; IBM ILOG CPLEX 20.1, third party, not in the reference tree).
//
// Formulation (qp_tables.hpp): with the whitened null-space basis of the equality rows the QP is the least-distance
// problem  min |v|^2  s.t.  n_j . v >= -s_j(x0)  in 39 dimensions, x = x0 + (G (+) G (+) G) v. It is solved by a dual
// active-set method (Goldfarb-Idnani with an identity Hessian): start at the unconstrained minimiser, pick a violated
// row, move along the projection z of its normal onto the orthogonal complement of the active normals until the row
// is met or an active multiplier reaches zero (then that row leaves), repeat.
//
// Factorisation: only a THIN orthonormal basis Q (39 x q) of the q active normals, the q x q triangle R
// (N_active = Q R) and its inverse W are kept, in shared memory. The projection is classical Gram-Schmidt against Q with
// a second pass when the first cancelled most of the vector (so it is as accurate as a full QR), the multiplier step
// is the product W d, adding a row appends a column to Q, R and W, dropping a row is a sequence of Givens rotations on
// R's rows and on Q's and W's columns. Work per iteration is O(39 q), q ~ 5-20, instead of the O(39^2) of a full
// orthogonal factor. Shared-memory operands are staged in registers ahead of the FMAs throughout: the update runs in
// ONE warp and is a chain of dependent steps, so exposed load latencies, not throughput, set its speed.
//
// Rows are never assembled as a matrix. Bounds (SFC boxes + world box), velocity and acceleration limits are priced
// from x with per-thread constants held in registers (225 variables / stencils spread over the block). LSC rows come
// from the row store written by k_lsc_build (three non-zeros each) and are priced EVERY iteration over all kept pairs
// (those that survived k_lsc_build's exact culling), so the pivot is always the globally most violated row — the rule
// that keeps the iteration count low (4-20) — but distance-gated: in the whitened space every row normal has unit
// length, so a row's slack cannot fall faster than the iterate travels. Per pair we keep `safe` = (distance travelled
// when it was last evaluated) + (smallest whitened slack of its rows then); a thread reads that one number per pair
// (8 B, coalesced), the pairs whose gate is open are compacted into a shared-memory list, and the list is evaluated
// evenly spread over the block (branch-free, next record prefetched). The solve ends when no evaluated row is violated
// beyond the feasibility tolerance; every skipped row is provably satisfied.
// The whole block prices (per-thread best -> warp redux arg-min -> one shared-memory exchange); warp 0 then does the
// factorisation update while the other warps wait at the barrier: the step time of a swarm is the slowest agent's
// solve, so the design minimises one agent's latency, not aggregate throughput.
#include <cstdlib>

#include "kernels.hpp"

namespace lscgpu {

constexpr int NR = kRed;        // 39
constexpr int LD = 39;          // row pitch of Q and R (odd: row-per-lane accesses are bank-conflict free)
// Primal feasibility tolerance = CPLEX's default EpRHS (the reference sets no tolerance, src/traj_optimizer.cpp:42-54):
// like a dual simplex, a row enters the working set only when violated by more than this; entered rows are then met
// exactly. Trajectories travel as float32, so agents in contact see hulls ~1e-7 closer than r_i + r_j; an exact
// solver would call that infeasible where CPLEX answers "optimal".
constexpr double kFeasTol = 1e-6;
constexpr double kZeroTol = 1e-13;

struct QpShared {
    double Q[NR * LD];          // columns 0..q-1: orthonormal basis of the active normals
    double R[NR * LD];          // upper triangle, row-major
    double W[NR * LD];          // R^-1 (upper triangle, zeros below): the multiplier step is a product, not a substitution
    double x[kNv];
    double nv[NR], z[NR], d[NR], tmp[NR], rr[NR], lam[NR];
    double inv_gn[kAx];
    double lb[15], ub[15], vmax[3], amax[3];
    double travelled;           // path length of the iterate in the whitened space
    double best_mu[16];          // per-warp pricing result
    int best_id[16];
    int stop;                   // 0 run, 1 finished/failed (set by warp 0)
    int open_count[2];          // pairs in the open list of the current / next chunk
    int act[NR];
    double vacc[NR + 1];        // whitened step accumulated by warp 0 since the last block-wide update of x
};

struct Best {
    double mu;
    int id;
};

// slack: a.x - b of the row; scale: 1 / (whitened length of its normal).
// Active rows need no special treatment: they are met exactly (|slack| ~ 1e-13), far inside the tolerance, so they
// can never be selected again while active (checked once per iteration on the winner only).
__device__ __forceinline__ void consider(Best& b, const QpShared& S, int q, double slack, double scale, int id) {
    if (!(slack < -kFeasTol)) return;
    const double mu = scale < INFINITY ? slack * scale : -INFINITY;   // zero normal with positive rhs: infeasible row
    if (mu < b.mu) { b.mu = mu; b.id = id; }
}
// most violated candidate of the warp (smallest mu, ties to the smallest row id): three 32-bit `redux.min` on an
// order-preserving key of the double instead of five shuffle rounds
__device__ __forceinline__ Best warp_argmin(Best b) {
    unsigned long long key = ~0ull;
    if (b.id >= 0) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(b.mu);
        key = bits ^ ((bits >> 63) ? ~0ull : 0x8000000000000000ull);
    }
    const unsigned hi = (unsigned)(key >> 32);
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned lo = hi == mhi ? (unsigned)key : 0xffffffffu;
    const unsigned mlo = __reduce_min_sync(0xffffffffu, lo);
    const bool win = b.id >= 0 && hi == mhi && (unsigned)key == mlo;
    const unsigned wid = __reduce_min_sync(0xffffffffu, win ? (unsigned)b.id : 0xffffffffu);
    Best r{0.0, -1};
    if (wid != 0xffffffffu) {
        const unsigned long long mk = ((unsigned long long)mhi << 32) | mlo;
        r.mu = __longlong_as_double((long long)(mk ^ ((mk >> 63) ? 0x8000000000000000ull : ~0ull)));
        r.id = (int)wid;
    }
    return r;
}

// one (obstacle, segment) pair: up to 6 rows. Returns the smallest whitened slack of the pair's rows. Branch free: the
// six rows are evaluated as six independent chains (all shared-memory loads first), rows that do not exist (initial-state
// control points of segment 0) or are not violated beyond the tolerance become +inf by selects, and only the pair's most
// violated row (smallest i on ties, as a sequential scan would keep) is offered to the thread's running best.
__device__ __forceinline__ double price_pair_vals(Best& best, const QpShared& S, int q, int slot, int m, float4 nr,
                                                  const double* r6) {
    const double ax = (double)nr.x, ay = (double)nr.y, az = (double)nr.z, inv = (double)nr.w;
    const double* xb = S.x + m * 6;
    const double* gb = S.inv_gn + m * 6;
    double slack[6], scale[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        slack[i] = ax * xb[i] + ay * xb[kAx + i] + az * xb[2 * kAx + i] - r6[i];
        scale[i] = inv * gb[i];
    }
    double mu_min = INFINITY, cand_mu = INFINITY;
    int cand_i = -1;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const bool exists = !(i < kPhi && m == 0);
        const bool finite = scale[i] < INFINITY;
        const double mu = finite ? slack[i] * scale[i] : (slack[i] < 0.0 ? -INFINITY : INFINITY);
        mu_min = fmin(mu_min, exists ? mu : INFINITY);
        // candidate value as `consider` computes it: zero normal with positive rhs = infeasible row (-inf)
        const double mu_c = (exists && slack[i] < -kFeasTol) ? (finite ? mu : -INFINITY) : INFINITY;
        if (mu_c < cand_mu) { cand_mu = mu_c; cand_i = i; }
    }
    if (cand_i >= 0 && cand_mu < best.mu) { best.mu = cand_mu; best.id = kFixedRows + slot * 6 + cand_i; }
    return mu_min;
}

// one 64-byte row record: four 16-byte loads
__device__ __forceinline__ void load_pair(const RowRec* rows, int slot, float4& nr, double* r6) {
    const float4* src = reinterpret_cast<const float4*>(rows + slot);
    nr = src[0];
    const double2 a = *reinterpret_cast<const double2*>(src + 1), b = *reinterpret_cast<const double2*>(src + 2),
                  c = *reinterpret_cast<const double2*>(src + 3);
    r6[0] = a.x; r6[1] = a.y; r6[2] = b.x; r6[3] = b.y; r6[4] = c.x; r6[5] = c.y;
}

// The selected row in registers, decoded redundantly by every lane of warp 0 (uniform branches, no shared-memory
// round trip): at most three non-zeros a[t] at variables idx[t] = axis * 30 + var, right-hand side b.
struct RowRegs {
    int nnz;
    int idx[3];
    double a[3];
    double b;
};
__device__ __forceinline__ RowRegs decode_row(int id, int n_obs, const RowRec* rows, const int* kept, double vel_coef,
                                              double acc_coef, const double* lb, const double* ub, const double* vmax,
                                              const double* amax) {
    RowRegs r;
    r.idx[1] = r.idx[2] = 0; r.a[1] = r.a[2] = 0.0;
    if (id < 180) {
        const int var = id >> 1, side = id & 1;
        const int k = var / kAx, m = (var % kAx) / 6;
        r.nnz = 1; r.idx[0] = var;
        if (side == 0) { r.a[0] = 1.0; r.b = lb[m * 3 + k]; }
        else { r.a[0] = -1.0; r.b = -ub[m * 3 + k]; }
    } else if (id < kFixedRows) {
        const int e = id - 180, side = e & 1, idx = e >> 1;
        const int k = idx / 45, rem = idx % 45, m = rem / 9, j = rem % 9;
        const int base = k * kAx + m * 6;
        const double sg = side == 0 ? -1.0 : 1.0;
        if (j < 5) {
            r.nnz = 2; r.idx[0] = base + j + 1; r.idx[1] = base + j;
            r.a[0] = sg * vel_coef; r.a[1] = -sg * vel_coef; r.b = -vmax[k];
        } else {
            const int i = j - 5;
            r.nnz = 3; r.idx[0] = base + i + 2; r.idx[1] = base + i + 1; r.idx[2] = base + i;
            r.a[0] = sg * acc_coef; r.a[1] = -2.0 * sg * acc_coef; r.a[2] = sg * acc_coef; r.b = -amax[k];
        }
    } else {
        const int e = id - kFixedRows, slot = e / 6, i = e % 6;
        const int kp = kept[slot];
        const int m = (kp >= n_obs) + (kp >= 2 * n_obs) + (kp >= 3 * n_obs) + (kp >= 4 * n_obs), vi = m * 6 + i;
        const RowRec& rec = rows[slot];
        r.nnz = 3;
        r.idx[0] = vi; r.idx[1] = kAx + vi; r.idx[2] = 2 * kAx + vi;
        r.a[0] = (double)rec.ax; r.a[1] = (double)rec.ay; r.a[2] = (double)rec.az;
        r.b = rec.rhs[i];
    }
    return r;
}

// remove active row l: delete column l of R, restore the triangle with Givens rotations (rows j, j+1 of R,
// columns j, j+1 of Q)
// With W = R^-1: R' = (G R P)[:q-1] (P deletes column l, G the rotations) gives W' = (P^T W G^T)[:, :q-1]: delete ROW l of W
// and rotate its COLUMNS with the same coefficients.
__device__ __forceinline__ void drop_active(QpShared& S, int& q, int l, int lane) {
    __syncwarp();
    for (int r = lane; r < q; r += 32) {
        for (int j = l; j < q - 1; j++) S.R[r * LD + j] = S.R[r * LD + j + 1];
        S.R[r * LD + q - 1] = 0.0;
    }
    for (int c = lane; c < q; c += 32) {
        for (int r = l; r < q - 1; r++) S.W[r * LD + c] = S.W[(r + 1) * LD + c];
        S.W[(q - 1) * LD + c] = 0.0;
    }
    if (lane == 0)
        for (int j = l; j < q - 1; j++) { S.act[j] = S.act[j + 1]; S.lam[j] = S.lam[j + 1]; }
    q--;
    for (int j = l; j < q; j++) {
        __syncwarp();
        const double a = S.R[j * LD + j], b = S.R[(j + 1) * LD + j];
        __syncwarp();
        if (b == 0.0) continue;
        const double ih = 1.0 / sqrt(a * a + b * b), c = a * ih, s = b * ih;       // entries of R are <= 1: no overflow concerns
        for (int k = j + lane; k < q; k += 32) {
            const double t1 = S.R[j * LD + k], t2 = S.R[(j + 1) * LD + k];
            S.R[j * LD + k] = c * t1 + s * t2;
            S.R[(j + 1) * LD + k] = -s * t1 + c * t2;
        }
        if (lane == 0) S.R[(j + 1) * LD + j] = 0.0;
        for (int r = lane; r <= j; r += 32) {                  // rows below j have zeros in both columns
            const double t1 = S.W[r * LD + j], t2 = S.W[r * LD + j + 1];
            S.W[r * LD + j] = c * t1 + s * t2;
            S.W[r * LD + j + 1] = -s * t1 + c * t2;
        }
        for (int r = lane; r < NR; r += 32) {
            const double t1 = S.Q[r * LD + j], t2 = S.Q[r * LD + j + 1];
            S.Q[r * LD + j] = c * t1 + s * t2;
            S.Q[r * LD + j + 1] = -s * t1 + c * t2;
        }
    }
    __syncwarp();
    for (int r = lane; r <= q; r += 32) S.W[r * LD + q] = 0.0;         // the column that fell off
    __syncwarp();
}

// Fixed rows (ids 0..449) are 90 variable-bound pairs and 135 dynamic-limit stencil pairs = 225 items, spread over the
// threads of the block; each thread keeps the constants of its items in registers.
template <int kItems>
struct FixedItems {
    int base[kItems];       // bounds: variable index; stencils: first variable; -1: none
    int kind[kItems];       // 0 bound, 1 velocity, 2 acceleration
    int id[kItems];         // row id of side 0
    double lo[kItems], hi[kItems], inv[kItems];     // bounds: lb, ub, 1/gnorm; stencils: limit, -, 1/|n|
};

template <int kThreads>
__global__ void __launch_bounds__(kThreads, 512 / kThreads) k_qp_solve(QpLaunch L) {
    constexpr int kWarps = kThreads / 32;
    constexpr int kItems = (225 + kThreads - 1) / kThreads;
    constexpr int kGate = 8;                       // gate values per thread per chunk
    __shared__ QpShared S;
    __shared__ int open_list[kGate * kThreads];
    const long long t_start = clock64();
    // scheduling order != data order: blocks are dispatched in blockIdx order, so the launcher may hand the agents
    // with the most expensive solve of the previous step out first (longest-processing-time-first)
    const int b = L.order ? L.order[L.first + blockIdx.x] : L.first + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool batch = L.obs_offset != nullptr;
    const int agent = L.agent_index ? L.agent_index[b] : L.agent_base + b;
    const int di = batch ? b : agent;                 // index into state9/goal3/ts/boxes
    const int n_obs = batch ? (L.obs_offset[b + 1] - L.obs_offset[b]) : L.n_obs;
    const RowRec* rows = batch ? L.rows + (size_t)kPairsPerObs * L.obs_offset[b] : L.rows + (size_t)b * L.P_pad;
    double* safe = batch ? L.safe + (size_t)kPairsPerObs * L.obs_offset[b] : L.safe + (size_t)b * L.P_pad;
    const int* kept = batch ? L.kept + (size_t)kPairsPerObs * L.obs_offset[b] : L.kept + (size_t)b * L.P_pad;
    const int n_kept = L.kept_count[b];
    const QpTablesDev& T = *L.T;
    const int ts = L.ts[di];
    const double vel_coef = T.vel_coef, acc_coef = T.acc_coef;
    const AgentConstDev& ac = L.consts[agent];

    // ---- stage tables and problem data ------------------------------------------------------------------------
    const double* __restrict__ Gt = &T.G[ts - 1][0][0];      // whitened basis of this ts (read-only, L1-resident)
    for (int e = tid; e < kAx; e += kThreads) S.inv_gn[e] = 1.0 / T.gnorm[ts - 1][e];
    if (tid < 15) {
        const int m = tid / 3, k = tid % 3;
        double lo = (double)L.wmin[k], hi = (double)L.wmax[k];
        if (L.boxes) {      // SFC rows == per-variable bounds (src/traj_optimizer.cpp:409-434)
            const float* bx = L.boxes + (size_t)di * 30 + m * 6;
            lo = fmax(lo, (double)bx[k]);
            hi = fmin(hi, (double)bx[3 + k]);
        }
        S.lb[tid] = lo; S.ub[tid] = hi;
    }
    if (tid < 3) { S.vmax[tid] = ac.vmax[tid]; S.amax[tid] = ac.amax[tid]; }
    const double* st = L.state9 + (size_t)di * 9;
    const double* gl = L.goal3 + (size_t)di * 3;
    for (int e = tid; e < kNv; e += kThreads) {
        const int k = e / kAx, i = e % kAx;
        const double* Xs = T.Xs[ts - 1][i];
        S.x[e] = Xs[0] * st[k] + Xs[1] * st[3 + k] + Xs[2] * st[6 + k] + T.xg[ts - 1][i] * gl[k];
    }
    for (int e = tid; e < NR * LD; e += kThreads) { S.R[e] = 0.0; S.W[e] = 0.0; }
    if (tid == 0) { S.travelled = 0.0; S.stop = 0; }
    __syncthreads();
    FixedItems<kItems> F;
#pragma unroll
    for (int t = 0; t < kItems; t++) {
        const int item = tid + kThreads * t;
        F.base[t] = -1; F.kind[t] = 0; F.id[t] = 0; F.lo[t] = F.hi[t] = F.inv[t] = 0.0;
        if (item < kNv) {
            const int k = item / kAx, mi = item % kAx, m = mi / 6, i = mi % 6;
            if (!(m == 0 && i < kPhi)) {
                F.base[t] = item; F.kind[t] = 0; F.id[t] = item * 2;
                F.lo[t] = S.lb[m * 3 + k]; F.hi[t] = S.ub[m * 3 + k]; F.inv[t] = S.inv_gn[mi];
            }
        } else if (item < kNv + 135) {
            const int idx = item - kNv;
            const int k = idx / 45, rem = idx % 45, m = rem / 9, j = rem % 9;
            const bool vel = j < 5;
            const int i = vel ? j : j - 5;
            const bool skip = vel ? (m == 0 && j < 2) : (m == 0 && i == 0);
            if (!skip) {
                F.base[t] = k * kAx + m * 6 + i; F.kind[t] = vel ? 1 : 2; F.id[t] = 180 + idx * 2;
                F.lo[t] = vel ? S.vmax[k] : S.amax[k]; F.inv[t] = 1.0 / T.dyn_norm[ts - 1][m][j];
            }
        }
    }
    // warp 0: per-lane index constants of the factorisation update (no divisions in the loop)
    const int c_axis0 = lane / kFree, c_col0 = lane % kFree;      // whitened coordinate c = lane
    int q = 0, iters = 0, status = LSCGPU_QP_OK;
    unsigned long long pairs_evaluated = 0, passes = 0;
    long long price_cycles = 0;
#ifdef LSCGPU_QP_SECTION_TIMERS      // build with -DLSCGPU_QP_SECTION_TIMERS and run with LSCGPU_QP_DEBUG=1: cycles per section
    long long sec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tk = 0;
#define TICK() (tk = clock64())
#define TOCK(i) do { const long long now_ = clock64(); sec[i] += now_ - tk; tk = now_; } while (0)
#else
#define TICK() ((void)0)
#define TOCK(i) ((void)0)
#endif

    while (true) {
        // ---- pricing by the whole block ---------------------------------------------------------------------------
        Best best{0.0, -1};
        const long long t_price = clock64();
        if (tid == 0) S.open_count[0] = 0;
        __syncthreads();        // x of the previous update is complete (block-wide update below), list counter zeroed
        const double travelled = S.travelled;
#pragma unroll
        for (int t = 0; t < kItems; t++) {
            if (F.base[t] < 0) continue;
            const double* c = S.x + F.base[t];
            if (F.kind[t] == 0) {
                consider(best, S, q, c[0] - F.lo[t], F.inv[t], F.id[t]);
                consider(best, S, q, F.hi[t] - c[0], F.inv[t], F.id[t] + 1);
            } else {
                const double expr = F.kind[t] == 1 ? vel_coef * (c[1] - c[0]) : acc_coef * (c[2] - 2.0 * c[1] + c[0]);
                consider(best, S, q, F.lo[t] - expr, F.inv[t], F.id[t]);
                consider(best, S, q, F.lo[t] + expr, F.inv[t], F.id[t] + 1);
            }
        }
        // LSC rows, chunk by chunk: (1) every thread reads kGate gate values (coalesced, all loads in flight at once),
        // (2) the pairs whose gate is open are compacted into a shared-memory list, (3) the list is evaluated evenly
        // spread over the block, the next record being loaded while the current one is evaluated.
        for (int s0 = 0, chunk = 0; s0 < n_kept; s0 += kGate * kThreads, chunk++) {
            double sv[kGate];
#pragma unroll
            for (int h = 0; h < kGate; h++) {
                const int si = s0 + kThreads * h + tid;
                sv[h] = si < n_kept ? safe[si] : INFINITY;
            }
            if (chunk > 0) __syncthreads();     // list of the previous chunk consumed, counter of this one zeroed
            int* cnt = &S.open_count[chunk & 1];
#pragma unroll
            for (int h = 0; h < kGate; h++) {
                const bool open = !(sv[h] > travelled);         // may be violated by now
                const unsigned mask = __ballot_sync(0xffffffffu, open);
                if (mask == 0u) continue;
                int base = 0;
                if (lane == 0) base = atomicAdd(cnt, __popc(mask));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (open) open_list[base + __popc(mask & ((1u << lane) - 1u))] = s0 + kThreads * h + tid;
            }
            __syncthreads();
            const int n_open = *cnt;
            if (tid == 0) S.open_count[(chunk + 1) & 1] = 0;
            int idx = tid, si = 0, kp = 0;
            float4 nr; double r6[6];
            if (idx < n_open) { si = open_list[idx]; load_pair(rows, si, nr, r6); kp = kept[si]; }
            while (idx < n_open) {
                const int nidx = idx + kThreads;
                int si2 = 0, kp2 = 0;
                float4 nr2 = nr; double r62[6];
                if (nidx < n_open) { si2 = open_list[nidx]; load_pair(rows, si2, nr2, r62); kp2 = kept[si2]; }
                const int m = (kp >= n_obs) + (kp >= 2 * n_obs) + (kp >= 3 * n_obs) + (kp >= 4 * n_obs);
                const double mu_min = price_pair_vals(best, S, q, si, m, nr, r6);
                // 1e-6 relative margin: the stored 1/|a| is float32, so mu carries ~6e-8 relative error
                safe[si] = travelled + (mu_min > 0.0 ? mu_min * 0.999999 : mu_min);
                pairs_evaluated++;
                idx = nidx; si = si2; kp = kp2; nr = nr2;
#pragma unroll
                for (int i = 0; i < 6; i++) r6[i] = r62[i];
            }
        }
        passes++;
        best = warp_argmin(best);
        if (lane == 0) { S.best_mu[warp] = best.mu; S.best_id[warp] = best.id; }
        __syncthreads();
        {   // every warp reduces the per-warp results again (lane w holds warp w's): same answer in all threads
            Best wb{0.0, -1};
            if (lane < kWarps) { wb.mu = S.best_mu[lane]; wb.id = S.best_id[lane]; }
            best = warp_argmin(wb);
        }
        if (tid == 0) price_cycles += clock64() - t_price;
        if (best.id < 0) break;                 // no row violated beyond the tolerance anywhere: done (block-uniform)

        // ---- factorisation update by warp 0 -----------------------------------------------------------------------
        if (warp == 0) {
            bool done = false;          // set when the solve must stop (failure)
            TICK();
            S.vacc[lane] = 0.0;
            if (lane + 32 < NR + 1) S.vacc[lane + 32] = 0.0;
            __syncwarp();
            do {
                {
                    bool dup = false;
                    for (int k = lane; k < q; k += 32) dup |= S.act[k] == best.id;
                    if (__any_sync(0xffffffffu, dup)) { status = LSCGPU_QP_MAXITER; done = true; break; }   // numerical breakdown
                }
                const RowRegs row = decode_row(best.id, n_obs, rows, kept, vel_coef, acc_coef, S.lb, S.ub, S.vmax, S.amax);
                // whitened normal  nv = (G (+) G (+) G)^T a: lane c owns coordinates c and c + 32 (axis = c / 13)
                double part = 0.0, nv_reg[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int c = lane + 32 * h;
                    nv_reg[h] = 0.0;
                    if (c < NR) {
                        const int k = h == 0 ? c_axis0 : 2, cc = h == 0 ? c_col0 : c - 2 * kFree;
                        double sacc = 0.0;
#pragma unroll
                        for (int t = 0; t < 3; t++) {
                            const int ax_t = row.idx[t] / kAx;
                            if (t < row.nnz && ax_t == k) sacc += row.a[t] * __ldg(Gt + (row.idx[t] - ax_t * kAx) * kFree + cc);
                        }
                        nv_reg[h] = sacc;
                        part += sacc * sacc;
                    }
                }
                const double nrm_len = sqrt(warp_sum(part));
                if (!(nrm_len > 0.0)) { status = LSCGPU_QP_INFEASIBLE; done = true; break; }
                const double inv_len = 1.0 / nrm_len;
                S.nv[lane] = nv_reg[0] * inv_len;
                if (lane + 32 < NR) S.nv[lane + 32] = nv_reg[1] * inv_len;
                // slack of the selected row (normalised); along the step it grows by t |z|^2 (a . G z = |G^T a| nv . z and
                // nv . z = z . z for the projection z of nv), so x itself is only brought up to date once per update
                double slack = -row.b;
#pragma unroll
                for (int t = 0; t < 3; t++) if (t < row.nnz) slack += row.a[t] * S.x[row.idx[t]];
                slack *= inv_len;
                __syncwarp();
                double lam_p = 0.0;
                TOCK(1);
                while (true) {
                    if (++iters > L.max_iter) { status = LSCGPU_QP_MAXITER; done = true; break; }
                    // ---- z = (I - Q Q^T) nv by Gram-Schmidt (second pass when needed); d = Q^T nv --------------------
                    for (int c = lane; c < NR; c += 32) { S.z[c] = S.nv[c]; S.d[c] = 0.0; }
                    __syncwarp();
                    double zz = 1.0;                 // |nv| = 1
                    if (q > 0) {
#pragma unroll 1
                        for (int pass = 0; pass < 2; pass++) {
                            // lane k: column k of Q against z. Three chunks of 13 rows: the 13 column entries are loaded
                            // into registers first (independent shared-memory loads in flight together), then multiplied
                            // against the broadcast z values; without the staging the compiler funnels all 39 loads
                            // through one register pair and exposes the load latency 39 times.
                            if (lane < q) {
                                const double* qc = S.Q + lane;
                                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
                                for (int r0 = 0; r0 < NR; r0 += 13) {
                                    double qv[13];
#pragma unroll
                                    for (int i = 0; i < 13; i++) qv[i] = qc[(r0 + i) * LD];
#pragma unroll
                                    for (int i = 0; i < 13; i++) {
                                        const double pz = S.z[r0 + i];
                                        if (i % 3 == 0) s0 += qv[i] * pz;
                                        else if (i % 3 == 1) s1 += qv[i] * pz;
                                        else s2 += qv[i] * pz;
                                    }
                                }
                                const double sdot = s0 + s1 + s2;
                                S.tmp[lane] = sdot;
                                S.d[lane] += sdot;
                            }
#pragma unroll 1
                            for (int k = lane + 32; k < q; k += 32) {     // q > 32 only
                                double sdot = 0.0;
#pragma unroll 1
                                for (int r = 0; r < NR; r++) sdot += S.Q[r * LD + k] * S.z[r];
                                S.tmp[k] = sdot;
                                S.d[k] += sdot;
                            }
                            __syncwarp();
                            // lane r: rows r and r + 32 of Q against the coefficients, both in one pass over k
                            // (lanes without a second row read row 38 and drop the result)
                            double zp;
                            {
                                const int r2 = min(lane + 32, NR - 1);
                                const double* q1 = S.Q + lane * LD;
                                const double* q2 = S.Q + r2 * LD;
                                double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
                                int k = 0;
                                for (; k + 3 < q; k += 4) {                 // four columns per trip: 12 loads, then 8 FMAs
                                    const double t0 = S.tmp[k], t1 = S.tmp[k + 1], t2 = S.tmp[k + 2], t3 = S.tmp[k + 3];
                                    const double u0 = q1[k], u1 = q1[k + 1], u2 = q1[k + 2], u3 = q1[k + 3];
                                    const double w0 = q2[k], w1 = q2[k + 1], w2 = q2[k + 2], w3 = q2[k + 3];
                                    a0 += u0 * t0; a1 += u1 * t1; b0 += w0 * t0; b1 += w1 * t1;
                                    a0 += u2 * t2; a1 += u3 * t3; b0 += w2 * t2; b1 += w3 * t3;
                                }
                                for (; k < q; k++) { const double t0 = S.tmp[k]; a0 += q1[k] * t0; b0 += q2[k] * t0; }
                                const double z1 = S.z[lane] - (a0 + a1);
                                S.z[lane] = z1;
                                zp = z1 * z1;
                                if (lane + 32 < NR) {
                                    const double z2 = S.z[lane + 32] - (b0 + b1);
                                    S.z[lane + 32] = z2;
                                    zp += z2 * z2;
                                }
                            }
                            const double zz_new = warp_sum(zp);
                            __syncwarp();
                            // "twice is enough": a second pass only when the first one cancelled most of the vector
                            const bool again = zz_new < 0.25 * zz;
                            zz = zz_new;
                            if (!again) break;
                        }
                    }
                    TOCK(2);
                    // rr = R^-1 d (change of the active multipliers per unit step): column-oriented back
                    // substitution, rr[k] lives in lane k's register while q <= 32
                    double t1 = INFINITY;
                    int l = -1;
                    // rr = R^-1 d = W d (change of the active multipliers per unit step): lane k owns row k; the zeros below
                    // the diagonal make the product predicate free; four columns per trip, loads staged before the FMAs
                    for (int k = lane; k < q; k += 32) {
                        const double* wk = S.W + k * LD;
                        double a0 = 0.0, a1 = 0.0;
                        int c = 0;
                        for (; c + 3 < q; c += 4) {
                            const double w0 = wk[c], w1 = wk[c + 1], w2 = wk[c + 2], w3 = wk[c + 3];
                            const double d0 = S.d[c], d1 = S.d[c + 1], d2 = S.d[c + 2], d3 = S.d[c + 3];
                            a0 += w0 * d0; a1 += w1 * d1; a0 += w2 * d2; a1 += w3 * d3;
                        }
                        for (; c < q; c++) a0 += wk[c] * S.d[c];
                        S.rr[k] = a0 + a1;
                    }
                    __syncwarp();
                    if (q <= 32) {
                        const double rk = lane < q ? S.rr[lane] : 0.0;
                        // ratio test over the active multipliers: lane k holds candidate k; warp minimum of the ratio through
                        // its order-preserving 64-bit key (two 32-bit min reductions), ties to the smallest index
                        const bool cand = lane < q && rk > kZeroTol;
                        unsigned long long key = ~0ull;
                        if (cand) {
                            const unsigned long long bits = (unsigned long long)__double_as_longlong(S.lam[lane] / rk);
                            key = bits ^ ((bits >> 63) ? ~0ull : 0x8000000000000000ull);
                        }
                        const unsigned hi = (unsigned)(key >> 32);
                        const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
                        const unsigned lo = hi == mhi ? (unsigned)key : 0xffffffffu;
                        const unsigned mlo = __reduce_min_sync(0xffffffffu, lo);
                        const unsigned win = __ballot_sync(0xffffffffu, cand && hi == mhi && (unsigned)key == mlo);
                        if (win) {
                            l = __ffs(win) - 1;
                            const unsigned long long mk = ((unsigned long long)mhi << 32) | mlo;
                            t1 = __longlong_as_double((long long)(mk ^ ((mk >> 63) ? 0x8000000000000000ull : ~0ull)));
                        }
                    } else {
                        for (int k = lane; k < q; k += 32)
                            if (S.rr[k] > kZeroTol) {
                                const double t = S.lam[k] / S.rr[k];
                                if (t < t1) { t1 = t; l = k; }
                            }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const double ot = __shfl_xor_sync(0xffffffffu, t1, o);
                            const int ol = __shfl_xor_sync(0xffffffffu, l, o);
                            if (ol >= 0 && (l < 0 || ot < t1 || (ot == t1 && ol < l))) { t1 = ot; l = ol; }
                        }
                    }
                    __syncwarp();
                    const bool primal = zz > kZeroTol;
                    double t2 = primal ? -slack / zz : INFINITY;
                    if (t2 < 0.0) t2 = 0.0;
                    const double t = fmin(t1, t2);
                    if (!(t < INFINITY)) { status = LSCGPU_QP_INFEASIBLE; done = true; break; }
                    for (int k = lane; k < q; k += 32) S.lam[k] -= t * S.rr[k];
                    lam_p += t;
                    TOCK(3);
                    if (!primal) { drop_active(S, q, l, lane); TOCK(6); continue; }
                    if (lane == 0) S.travelled += t * sqrt(zz) * (1.0 + 1e-9) + 1e-13;
                    S.vacc[lane] += t * S.z[lane];
                    if (lane + 32 < NR) S.vacc[lane + 32] += t * S.z[lane + 32];
                    slack += t * zz;
                    __syncwarp();
                    TOCK(4);
                    if (t2 <= t1) {
                        // the row becomes active: new basis column z / |z|, new column (d, |z|) of R
                        const double zn = sqrt(zz), izn = 1.0 / zn;
                        for (int r = lane; r < NR; r += 32) S.Q[r * LD + q] = S.z[r] * izn;
                        for (int k = lane; k < q; k += 32) { S.R[k * LD + q] = S.d[k]; S.W[k * LD + q] = -S.rr[k] * izn; }
                        if (lane == 0) {
                            S.R[q * LD + q] = zn;
                            S.W[q * LD + q] = izn;
                            S.act[q] = best.id;
                            S.lam[q] = lam_p;
                        }
                        q++;
                        __syncwarp();
                        TOCK(5);
                        break;
                    }
                    drop_active(S, q, l, lane);
                    TOCK(6);
                }
            } while (false);
            if (done && lane == 0) S.stop = 1;
        }
        __syncthreads();
        const bool stop = S.stop != 0;
        // x += (G (+) G (+) G) vacc by the whole block: element e = axis * 30 + var, 13 products each (the barrier at the
        // top of the next pass makes it visible)
        for (int e = tid; e < kNv; e += kThreads) {
            const int k = e / kAx;
            const double* g = Gt + (e - k * kAx) * kFree;
            const double* va = S.vacc + k * kFree;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int c = 0; c + 1 < kFree; c += 2) { s0 += __ldg(g + c) * va[c]; s1 += __ldg(g + c + 1) * va[c + 1]; }
            s0 += __ldg(g + kFree - 1) * va[kFree - 1];
            S.x[e] += s0 + s1;
        }
        if (stop) break;
    }
    __syncthreads();

    // ---- epilogue ---------------------------------------------------------------------------------------------
    if (L.counters) {
        const unsigned long long ev = warp_sum((double)pairs_evaluated) + 0.5;
        if (lane == 0) atomicAdd(&L.counters->rows_priced, 6ull * ev + (warp == 0 ? 414ull * passes : 0ull));
    }
    if (warp != 0) return;
    // objective as the reference reports it (getObjValue incl. the constant of the terminal cost)
    double jpart = 0.0;
    if (lane < 15) {
        const int k = lane / 5, m = lane % 5;
        const double* c = S.x + k * kAx + m * 6;
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) jpart += T.Qw[i][j] * c[i] * c[j];
        if (m >= kM - ts) { const double e = c[5] - gl[k]; jpart += T.wT * e * e; }
    }
    const double cost = warp_sum(jpart);
    if (L.counters && lane == 0) {
        atomicAdd(&L.counters->qp_iterations, (unsigned long long)iters);
        atomicAdd(&L.counters->full_passes, passes);
    }
    if (L.x_out)
        for (int e = lane; e < kNv; e += 32) L.x_out[(size_t)b * kNv + e] = S.x[e];
    if (L.cost_out && lane == 0) { L.cost_out[b] = cost; L.status_out[b] = status; L.iters_out[b] = iters; }
    if (L.out) {
        lscgpu_agent_out& o = L.out[agent];
        float* tr = &o.traj[0][0][0];
        const bool ok = status == LSCGPU_QP_OK;
        // failure: the optimizer keeps its last successful trajectory and cost (src/traj_planner.cpp:1553-1584)
        for (int e = lane; e < kTrajFloats; e += 32) {
            const int axis = e % 3, cp = e / 3;
            tr[e] = ok ? (float)S.x[axis * kAx + cp] : L.prev_traj[(size_t)agent * kTrajFloats + e];
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
            const double c_rep = ok ? cost : L.last_cost[agent];
            o.qp_cost = c_rep;
            L.last_cost[agent] = c_rep;
            o.report = LSCGPU_REPORT_SUCCESS;
            o.qp_status = status;
            o.qp_iterations = iters;
            o.qp_active = q;
            o.flags = L.flags ? L.flags[agent] : 0;
            o.terminal_segments = ts;
            for (int k = 0; k < 3; k++) o.current_goal[k] = (float)gl[k];
            o.goal_kind = L.goal_kind ? L.goal_kind[agent] : 0;
            o.qp_sweeps = (int)passes;
            o.qp_kcycles = (int)((clock64() - t_start) >> 10);
            o.qp_price_kcycles = (int)(price_cycles >> 10);
            o.lsc_pairs_kept = n_kept;
#ifdef LSCGPU_QP_SECTION_TIMERS
            if (L.dbg) { for (int i = 1; i < 8; i++) L.dbg[(size_t)b * 8 + i] = sec[i]; L.dbg[(size_t)b * 8] = price_cycles; }
#endif
        }
    }
}

// Longest-processing-time-first order of the local agents from the cycle count each solve recorded in its result
// record at the previous step: counting sort on the (clamped) kilo-cycle count, descending. The order inside a bucket
// is arbitrary; it only affects scheduling, never results.
__global__ void __launch_bounds__(1024) k_qp_order(int n_local, int a0, const lscgpu_agent_out* out, int* order) {
    __shared__ int hist[1024];
    __shared__ int warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n_local; i += 1024) atomicAdd(&hist[1023 - min(max(out[a0 + i].qp_kcycles, 0), 1023)], 1);
    __syncthreads();
    const int mine = hist[tid];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int t = warp_tot[lane], inc2 = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += u; }
        warp_tot[lane] = inc2 - t;
    }
    __syncthreads();
    hist[tid] = warp_tot[warp] + incl - mine;       // exclusive prefix
    __syncthreads();
    for (int i = tid; i < n_local; i += 1024) {
        const int pos = atomicAdd(&hist[1023 - min(max(out[a0 + i].qp_kcycles, 0), 1023)], 1);
        order[pos] = i;
    }
}
void launch_qp_order(int n_local, int a0, const lscgpu_agent_out* out, int* order, cudaStream_t s) {
    if (n_local > 0) k_qp_order<<<1, 1024, 0, s>>>(n_local, a0, out, order);
}

void launch_qp_solve(const QpLaunch& L, cudaStream_t s) {
    if (L.n_problems <= 0) return;
    // the step time is the slowest agent's solve, so every agent gets eight warps for pricing (256 threads, two blocks
    // per SM); 128 / 64 are kept for experiments (LSCGPU_QP_THREADS)
    static const int forced = getenv("LSCGPU_QP_THREADS") ? atoi(getenv("LSCGPU_QP_THREADS")) : 0;
    const int threads = forced ? forced : 256;
    if (threads >= 256) k_qp_solve<256><<<L.n_problems, 256, 0, s>>>(L);
    else if (threads == 128) k_qp_solve<128><<<L.n_problems, 128, 0, s>>>(L);
    else k_qp_solve<64><<<L.n_problems, 64, 0, s>>>(L);
}

}  // namespace lscgpu
