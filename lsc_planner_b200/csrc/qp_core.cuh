// qp_core — the Bernstein trajectory QP of ONE agent, solved by one thread block (sm_100a, FP64): all warps price the
// rows, warp 0 runs the dual active-set update. Shared by k_agent_plan (fused LSC build + QP, rows in shared memory)
// and k_qp_batch (operator-level entry, rows in global memory).
//
// Replaces TrajOptimizer::solve (src/traj_optimizer.cpp:31-154): buildDeq (:239-259), populatebyrow (:261-539) and the
// CPLEX dual-simplex call (:76; IBM ILOG CPLEX 20.1, third party, not in the reference tree).
//
// Formulation (qp_tables.hpp): with the whitened null-space basis of the equality rows the QP is the least-distance
// problem  min |v|^2  s.t.  n_j . v >= -s_j(x0)  in 39 dimensions, x = x0 + (G (+) G (+) G) v. Dual active set
// (Goldfarb-Idnani with an identity Hessian): start at the unconstrained minimiser, pick the most violated row, move
// along the projection z of its normal onto the orthogonal complement of the active normals until the row is met or an
// active multiplier reaches zero (then that row leaves), repeat.
//
// Factorisation. For the q active unit normals N (39 x q) the kernel keeps a THIN orthonormal basis Q of span(N) and a
// dense q x q matrix W with N W = Q — no triangular factor:
//   projection   d = Q^T n, z = n - Q d (classical Gram-Schmidt, second pass when the first cancelled most of n)
//   multipliers  rr = W d                      (N rr = Q Q^T n: change of the active multipliers per unit step)
//   add          Q <- [Q, z/|z|],  W <- [[W, -rr/|z|], [0, 1/|z|]]
//   drop l       y = W[l,:]/|W[l,:]| is, in Q coordinates, the direction of span(N) orthogonal to every OTHER active
//                normal. One Householder reflection H with H y = -+e_last: Q <- Q H, W <- W H; the last column of both
//                is deleted and the last row of W moves into row l (the multiplier list is permuted the same way).
// Every step is a matrix-vector product or a rank-one update with independent lanes: no chain of q Givens rotations
// (each a dependent sqrt + divide), no column shifts. tests/qp_kernel_model.py states the same update in numpy and is
// checked against the oracle's full-QR solver on the CPU (tests/test_qp_kernel_model.py).
//
// Scalars. The whitened length of a row normal is known in closed form (bounds: |G_i|; dynamic rows: a table; LSC rows:
// |a| |G_i| because the three axes share one basis), so normalising the entering normal needs no warp reduction; one
// rsqrt(|z|^2) yields |z|, 1/|z| and 1/|z|^2; the ratio test multiplies by a reciprocal.
//
// Pricing. Rows are never assembled as a matrix. Bounds (SFC boxes + world box), velocity and acceleration limits are
// priced from x with per-thread register constants (225 items spread over the block). LSC rows are priced every
// iteration (the pivot is always the globally most violated row, which keeps the iteration count low) but distance
// gated: whitened row normals have unit length, so a row's slack cannot fall faster than the iterate travels; per kept
// pair `gate` = travelled-at-last-evaluation + smallest whitened slack of its rows. Every warp owns slices of 128
// slots: it reads the gates, compacts the open slots into its own shared-memory list (ballot + popc, no block barrier)
// and evaluates the list one pair per lane. The solve ends when no evaluated row is violated beyond the tolerance
// (every skipped row is provably satisfied).
//
// Slack variables (QpSharedT<kE> with kE > 0; src/traj_optimizer.cpp:317-326,383-390,455-457). Once an agent's
// obs_slack_indices is not empty the reference relaxes the LSC rows of those obstacles by one variable eps_{oi,m} <= 0
// per (obstacle, segment) with cost w_s ((M - m)/M) eps^2. In the whitened space that is one extra coordinate
// e_p = sqrt(c_p) eps_p per such pair, objective |v|^2 + |e|^2, row normals (G^T a, -1/sqrt(c_p) at p) and the bound row
// (0, -1/sqrt(c_p) at p). A coordinate is created when a row of its pair first enters the working set (until then
// e_p = 0 and no active normal touches it): Q gets a zero row, nothing else changes — the same thin-Q / W update runs on
// 39 + n_e rows. Up to kE coordinates per agent; an agent that needs more fails like an iteration-limit hit
// (LSCGPU_FLAG_SLACK_OVERFLOW). The pair's slot carries its state in the upper bits of the segment byte: 0 hard rows,
// 31 slack pair without a coordinate yet, 1 + c coordinate c.
#pragma once
#include <cstddef>

#include "kernels.hpp"

namespace lscgpu {

constexpr int NR = kRed;        // 39
// Primal feasibility tolerance = CPLEX's default EpRHS (the reference sets no tolerance, src/traj_optimizer.cpp:42-54):
// like a dual simplex, a row enters the working set only when violated by more than this; entered rows are then met
// exactly. Trajectories travel as float32, so agents in contact see hulls ~1e-7 closer than r_i + r_j; an exact
// solver would call that infeasible where CPLEX answers "optimal".
constexpr double kFeasTol = 1e-6;
constexpr double kZeroTol = 1e-13;
constexpr int kWarpList = 128;  // slots one warp gates (and at most lists) per slice

constexpr int kSlackUntouched = 31;         // slot code of a slack pair that has no coordinate yet
constexpr int kSlackBoundBase = 1 << 30;    // row id of the bound eps <= 0 of slack coordinate c: kSlackBoundBase + c

template <int kE_>
struct QpSharedT {
    static constexpr int kE = kE_;              // slack coordinates an agent can hold (0: no slack variables)
    static constexpr int NRX = NR + kE_;        // rows of Q = dimension of the (extended) whitened space
    static constexpr int LDX = (NR + kE_) | 1;  // row pitch of Q and W (odd: row-per-lane accesses are bank-conflict free)
    double Q[NRX * LDX];        // columns 0..q-1: orthonormal basis of the active normals
    double W[NRX * LDX];        // rows/columns 0..q-1: N W = Q (row k belongs to active row act[k])
    double G[kAx * kFree];      // whitened basis of this agent's terminal-segment count
    double x[kNv];
    double z[NRX + 1], d[NRX + 1], tmp[NRX + 1], rr[NRX + 1], lam[NRX + 1];
    double vacc[NRX + 1];       // whitened step accumulated by warp 0 since the last block-wide update of x
    double inv_gn[kAx];         // 1 / |G row|
    double inv_dyn[kM * 9];     // 1 / whitened length of the velocity (j<5) / acceleration (j>=5) rows
    double lb[15], ub[15], vmax[3], amax[3];
    double travelled;           // path length of the iterate in the whitened space
    double best_mu[16];         // per-warp pricing result
    int best_id[16];
    int act[NRX + 1];
    int stop;                   // 0 run, 1 finished/failed (set by warp 0)
    unsigned long long g_bar;   // mbarrier: completion of the bulk copy of G
    // slack coordinates (kE > 0)
    double e[kE_ > 0 ? kE_ : 1];        // e_c = sqrt(c_p) eps_p  (<= 0 at the solution)
    double e_isc[kE_ > 0 ? kE_ : 1];    // 1 / sqrt(c_p), c_p = w_s (M - m) / M
    int e_slot[kE_ > 0 ? kE_ : 1];      // slot of the pair the coordinate belongs to
    double isc_m[kM];                   // 1 / sqrt(c_p) per segment
    int n_e;                            // coordinates in use
    int overflow;                       // 1: a row needed a coordinate beyond kE
};
using QpShared = QpSharedT<0>;
constexpr int kSlackCoords = 25;            // 39 + 25 = 64 rows: lane r owns rows r and r + 32
using QpSharedSlack = QpSharedT<kSlackCoords>;

// ---- bulk-asynchronous staging (TMA unit, non-tensor form): one thread arms an mbarrier with the byte count and issues
// cp.async.bulk global -> shared; the copy proceeds while the block builds its corridors; consumers wait on the barrier
// phase before the first use. (SASS: UBLKCP + SYNCS.)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_stage_begin(unsigned long long* bar, void* dst_smem, const void* src_global, unsigned bytes) {
    const unsigned b = smem_u32(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_global), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void bulk_stage_wait(unsigned long long* bar) {
    const unsigned b = smem_u32(bar);
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b) : "memory");
    } while (!done);
}

// Where the LSC rows of the agent live: slots [0, cap) in shared memory as structure-of-arrays (16-byte lanes of
// consecutive slots fall into different banks), slots >= cap — and every slot when cap == 0 — in the global row store.
struct RowSrc {
    float4* s_nr;               // [cap]    (a_x, a_y, a_z, 1/|a|)
    double2* s_rhs;             // [3][cap] rhs pairs (0,1), (2,3), (4,5)
    double* s_gate;             // [cap]
    unsigned char* s_seg;       // [cap]    segment m of the slot
    int cap;
    RowRec* g_rows;             // [..] indexed by slot
    double* g_gate;
    int* g_kept;                // dense pair index p = m * n_obs + obstacle of the slot
    int n_obs;

    __device__ __forceinline__ int seg_of_pair(int kp) const {
        return (kp >= n_obs) + (kp >= 2 * n_obs) + (kp >= 3 * n_obs) + (kp >= 4 * n_obs);
    }
    __device__ __forceinline__ double gate(int slot) const { return slot < cap ? s_gate[slot] : g_gate[slot]; }
    __device__ __forceinline__ void set_gate(int slot, double v) const {
        if (slot < cap) s_gate[slot] = v; else g_gate[slot] = v;
    }
    // kCoded: the slot's segment byte / pair word carries a slack code in its upper bits (slack kernels only)
    template <bool kCoded = false>
    __device__ __forceinline__ void load(int slot, float4& nr, double* r6, int& m) const {
        if (slot < cap) {
            nr = s_nr[slot];
            const double2 a = s_rhs[slot], b = s_rhs[cap + slot], c = s_rhs[2 * cap + slot];
            r6[0] = a.x; r6[1] = a.y; r6[2] = b.x; r6[3] = b.y; r6[4] = c.x; r6[5] = c.y;
            m = kCoded ? (s_seg[slot] & 7) : s_seg[slot];
        } else {
            const float4* src = reinterpret_cast<const float4*>(g_rows + slot);
            nr = src[0];
            const double2 a = *reinterpret_cast<const double2*>(src + 1), b = *reinterpret_cast<const double2*>(src + 2),
                          c = *reinterpret_cast<const double2*>(src + 3);
            r6[0] = a.x; r6[1] = a.y; r6[2] = b.x; r6[3] = b.y; r6[4] = c.x; r6[5] = c.y;
            m = seg_of_pair(kCoded ? (g_kept[slot] & 0xffffff) : g_kept[slot]);
        }
    }
    __device__ __forceinline__ int seg(int slot) const {
        return slot < cap ? (s_seg[slot] & 7) : seg_of_pair(g_kept[slot] & 0xffffff);
    }
    __device__ __forceinline__ int code(int slot) const {
        return slot < cap ? (s_seg[slot] >> 3) : ((g_kept[slot] >> 24) & 31);
    }
    __device__ __forceinline__ void set_code(int slot, int code, bool mirror) const {
        if (slot < cap) s_seg[slot] = (unsigned char)((s_seg[slot] & 7) | (code << 3));
        if (slot >= cap || mirror) g_kept[slot] = (g_kept[slot] & 0xffffff) | (code << 24);
    }
    __device__ __forceinline__ void store(int slot, const RowRec& rec, int m, int pair, double gate_v, bool mirror, int code = 0) const {
        if (slot < cap) {
            s_nr[slot] = make_float4(rec.ax, rec.ay, rec.az, rec.inv_an);
            s_rhs[slot] = make_double2(rec.rhs[0], rec.rhs[1]);
            s_rhs[cap + slot] = make_double2(rec.rhs[2], rec.rhs[3]);
            s_rhs[2 * cap + slot] = make_double2(rec.rhs[4], rec.rhs[5]);
            s_gate[slot] = gate_v;
            s_seg[slot] = (unsigned char)(m | (code << 3));
        }
        if (slot >= cap || mirror) {        // mirror: the whole row store also goes to global memory (lscgpu_get_lsc)
            float4* dst = reinterpret_cast<float4*>(g_rows + slot);
            const float4* src = reinterpret_cast<const float4*>(&rec);
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
            g_gate[slot] = gate_v;
            g_kept[slot] = pair | (code << 24);
        }
    }
};

struct Best {
    double mu;
    int id;
};

// a row is a candidate only when violated beyond the tolerance; mu = whitened slack (most negative wins)
__device__ __forceinline__ void consider(Best& b, double slack, double scale, int id) {
    if (!(slack < -kFeasTol)) return;
    const double mu = scale < INFINITY ? slack * scale : -INFINITY;   // zero normal with positive rhs: infeasible row
    if (mu < b.mu) { b.mu = mu; b.id = id; }
}

// 1/x to ~1 ulp for normal x: hardware seed + two Newton steps (a third of the latency of an IEEE division)
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    return bits ^ ((bits >> 63) ? ~0ull : 0x8000000000000000ull);
}
__device__ __forceinline__ double order_unkey(unsigned long long k) {
    return __longlong_as_double((long long)(k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull)));
}

// most violated candidate of the warp (smallest mu, ties to the smallest row id): three 32-bit `redux.min` on an
// order-preserving key of the double instead of five shuffle rounds
__device__ __forceinline__ Best warp_argmin(Best b) {
    const unsigned long long key = b.id >= 0 ? order_key(b.mu) : ~0ull;
    const unsigned hi = (unsigned)(key >> 32);
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned lo = hi == mhi ? (unsigned)key : 0xffffffffu;
    const unsigned mlo = __reduce_min_sync(0xffffffffu, lo);
    const bool win = b.id >= 0 && hi == mhi && (unsigned)key == mlo;
    const unsigned wid = __reduce_min_sync(0xffffffffu, win ? (unsigned)b.id : 0xffffffffu);
    Best r{0.0, -1};
    if (wid != 0xffffffffu) {
        r.mu = order_unkey(((unsigned long long)mhi << 32) | mlo);
        r.id = (int)wid;
    }
    return r;
}

// one (obstacle, segment) pair: up to 6 rows. Returns the smallest whitened slack of the pair's rows. Branch free: the
// six rows are six independent chains, rows that do not exist (initial-state control points of segment 0) or are not
// violated beyond the tolerance become +inf by selects, and only the pair's most violated row (smallest i on ties, as
// a sequential scan would keep) is offered to the thread's running best.
// `code` (slack kernels): 0 hard rows; otherwise the pair's rows read a . c - eps >= rhs with eps = e_c / sqrt(c_p) of
// its coordinate (0 while it has none) and are normalised by the length of the extended normal.
template <class SH>
__device__ __forceinline__ double price_pair(Best& best, const SH& S, int slot, int m, float4 nr, const double* r6, double& raw_min,
                                             int code = 0) {
    const double ax = (double)nr.x, ay = (double)nr.y, az = (double)nr.z, inv = (double)nr.w;
    double eps = 0.0, isc2 = 0.0;
    bool soft = false;
    if constexpr (SH::kE > 0) {
        soft = code != 0;
        if (soft) {
            const double isc = S.isc_m[m];
            isc2 = isc * isc;
            if (code != kSlackUntouched) eps = S.e[code - 1] * isc;
        }
    }
    // x[k][m][0..5] and inv_gn[m][0..5] are 16-byte aligned runs of six doubles
    const double2* xb0 = reinterpret_cast<const double2*>(S.x + m * 6);
    const double2* xb1 = reinterpret_cast<const double2*>(S.x + kAx + m * 6);
    const double2* xb2 = reinterpret_cast<const double2*>(S.x + 2 * kAx + m * 6);
    const double2* gb = reinterpret_cast<const double2*>(S.inv_gn + m * 6);
    double xs[6], ys[6], zs[6], gs[6];
#pragma unroll
    for (int h = 0; h < 3; h++) {
        const double2 a = xb0[h], b = xb1[h], c = xb2[h], g = gb[h];
        xs[2 * h] = a.x; xs[2 * h + 1] = a.y; ys[2 * h] = b.x; ys[2 * h + 1] = b.y;
        zs[2 * h] = c.x; zs[2 * h + 1] = c.y; gs[2 * h] = g.x; gs[2 * h + 1] = g.y;
    }
    double mu_min = INFINITY, cand_mu = INFINITY;
    int cand_i = -1;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double slack = ax * xs[i] + ay * ys[i] + az * zs[i] - r6[i];
        double scale = inv * gs[i];
        if constexpr (SH::kE > 0) {
            if (soft) {
                slack -= eps;
                // 1 / sqrt(|a|^2 |G_i|^2 + 1/c_p) from 1 / (|a| |G_i|)
                scale = scale < INFINITY ? scale * rsqrt(1.0 + isc2 * scale * scale) : rsqrt(isc2);
            }
        }
        const bool exists = !(i < kPhi && m == 0);
        raw_min = fmin(raw_min, exists ? slack : INFINITY);
        const bool finite = scale < INFINITY;
        const double mu = finite ? slack * scale : (slack < 0.0 ? -INFINITY : INFINITY);
        mu_min = fmin(mu_min, exists ? mu : INFINITY);
        const double mu_c = (exists && slack < -kFeasTol) ? (finite ? mu : -INFINITY) : INFINITY;
        if (mu_c < cand_mu) { cand_mu = mu_c; cand_i = i; }
    }
    if (cand_i >= 0 && cand_mu < best.mu) { best.mu = cand_mu; best.id = kFixedRows + slot * 6 + cand_i; }
    return mu_min;
}

// The selected row in registers, decoded redundantly by every lane of warp 0 (uniform branches): at most three
// non-zeros a[t] at variables idx[t] = axis * 30 + var, right-hand side b, inv_len = 1 / (whitened length of the normal).
struct RowRegs {
    int nnz;
    int idx[3];
    double a[3];
    double b, inv_len;
    double nn;          // |normal * inv_len|^2: 1 up to the rounding of inv_len (the float32 1/|a| of an LSC record)
    int ec;             // slack coordinate the row touches (-1: none) ...
    double ea;          // ... and its coefficient there, -1/sqrt(c_p)
};
template <class SH>
__device__ __forceinline__ RowRegs decode_row(int id, const SH& S, const RowSrc& rows, double vel_coef, double acc_coef) {
    RowRegs r;
    r.idx[1] = r.idx[2] = 0; r.a[1] = r.a[2] = 0.0;
    r.ec = -1; r.ea = 0.0;
    if constexpr (SH::kE > 0) {
        if (id >= kSlackBoundBase) {            // eps_c <= 0:  -(1/sqrt(c_p)) e_c >= 0
            const int c = id - kSlackBoundBase;
            r.nnz = 0; r.idx[0] = 0; r.a[0] = 0.0; r.b = 0.0;
            r.ec = c; r.ea = -S.e_isc[c];
            r.inv_len = 1.0 / S.e_isc[c];
            r.nn = 1.0;
            return r;
        }
    }
    if (id < 180) {
        const int var = id >> 1, side = id & 1;
        const int k = var / kAx, mi = var - k * kAx, m = mi / 6;
        r.nnz = 1; r.idx[0] = var;
        if (side == 0) { r.a[0] = 1.0; r.b = S.lb[m * 3 + k]; }
        else { r.a[0] = -1.0; r.b = -S.ub[m * 3 + k]; }
        r.inv_len = S.inv_gn[mi];
        r.nn = 1.0;
    } else if (id < kFixedRows) {
        const int e = id - 180, side = e & 1, idx = e >> 1;
        const int k = idx / 45, rem = idx - k * 45, m = rem / 9, j = rem - m * 9;
        const int base = k * kAx + m * 6;
        const double sg = side == 0 ? -1.0 : 1.0;
        if (j < 5) {
            r.nnz = 2; r.idx[0] = base + j + 1; r.idx[1] = base + j;
            r.a[0] = sg * vel_coef; r.a[1] = -sg * vel_coef; r.b = -S.vmax[k];
        } else {
            const int i = j - 5;
            r.nnz = 3; r.idx[0] = base + i + 2; r.idx[1] = base + i + 1; r.idx[2] = base + i;
            r.a[0] = sg * acc_coef; r.a[1] = -2.0 * sg * acc_coef; r.a[2] = sg * acc_coef; r.b = -S.amax[k];
        }
        r.inv_len = S.inv_dyn[m * 9 + j];
        r.nn = 1.0;
    } else {
        const int e = id - kFixedRows, slot = e / 6, i = e - slot * 6;
        float4 nr; double r6[6]; int m;
        rows.template load<(SH::kE > 0)>(slot, nr, r6, m);
        const int vi = m * 6 + i;
        r.nnz = 3;
        r.idx[0] = vi; r.idx[1] = kAx + vi; r.idx[2] = 2 * kAx + vi;
        r.a[0] = (double)nr.x; r.a[1] = (double)nr.y; r.a[2] = (double)nr.z;
        double rb = r6[0];
#pragma unroll
        for (int t = 1; t < 6; t++) if (t == i) rb = r6[t];
        r.b = rb;
        // 1 / (|a| |G_vi|) with the record's float32 1/|a|: any consistent scale of normal and slack serves the update
        // (the step, the multipliers and the travelled distance are invariant to it); 1/|a| = inf marks a zero normal
        r.inv_len = (double)nr.w * S.inv_gn[vi];
        r.nn = (r.a[0] * r.a[0] + r.a[1] * r.a[1] + r.a[2] * r.a[2]) * ((double)nr.w * (double)nr.w);
        if constexpr (SH::kE > 0) {
            const int code = rows.code(slot);
            if (code != 0) {                    // slack pair (it has a coordinate by now: the caller created it)
                const double isc = S.isc_m[m];
                r.ec = code - 1; r.ea = -isc;
                const double an2 = r.a[0] * r.a[0] + r.a[1] * r.a[1] + r.a[2] * r.a[2];
                const double gn = 1.0 / S.inv_gn[vi];
                r.inv_len = rsqrt(an2 * gn * gn + isc * isc);      // an2 == 0 (zero normal): the row only bounds eps
                r.nn = (an2 * gn * gn + isc * isc) * (r.inv_len * r.inv_len);
            }
        }
    }
    return r;
}

// Remove active row l (see the header comment): one Householder reflection on the columns of Q and W.
// kProj: the projection of the entering normal is carried through the drop instead of being recomputed: with
// Q H = [Q', u] the coefficients become d' = (H^T d)[0 .. q-2] and z' = z + delta u, delta = (H^T d)[q-1] (u is the
// direction the drop frees); returns delta (|z'|^2 = |z|^2 + delta^2).
template <class SH, bool kProj = false>
__device__ __forceinline__ double drop_active(SH& S, int& q, int l, int lane, int nr = NR) {
    constexpr int LD = SH::LDX;
    const int j = q - 1;
    double delta = 0.0;
    __syncwarp();
    if (j == 0 && kProj) {
        // the only active row leaves: u is its basis column
        delta = S.d[0];
        const int nrows = SH::kE > 0 ? nr : NR;
        S.z[lane] += delta * S.Q[lane * LD];
        if (lane + 32 < nrows) S.z[lane + 32] += delta * S.Q[(lane + 32) * LD];
    }
    if (j > 0) {
        const double y0 = lane < q ? S.W[l * LD + lane] : 0.0;
        const double y1 = lane + 32 < q ? S.W[l * LD + lane + 32] : 0.0;
        const double yj = S.W[l * LD + j];
        const double nn = warp_sum(y0 * y0 + y1 * y1);              // > 0: W is invertible
        const double ny = nn * rsqrt(nn);
        const double sg = yj >= 0.0 ? 1.0 : -1.0;
        const double beta = fast_rcp(ny * (ny + fabs(yj)));         // 2 / (v . v),  v = y + sg |y| e_j
        const double v0 = lane == j ? y0 + sg * ny : y0, v1 = lane + 32 == j ? y1 + sg * ny : y1;
        if (lane < q) S.tmp[lane] = v0;
        if (lane + 32 < q) S.tmp[lane + 32] = v1;
        double bvd = 0.0, vj = 0.0;
        if constexpr (kProj) {
            // d <- H^T d = d - beta (v . d) v
            bvd = beta * warp_sum((lane < q ? v0 * S.d[lane] : 0.0) + (lane + 32 < q ? v1 * S.d[lane + 32] : 0.0));
            vj = yj + sg * ny;
            delta = S.d[j] - bvd * vj;
            __syncwarp();
            if (lane < j) S.d[lane] -= bvd * v0;
            if (lane + 32 < j) S.d[lane + 32] -= bvd * v1;
        }
        __syncwarp();
        // Q <- Q - beta (Q v) v^T: lane r owns rows r and r + 32 (lanes without a second row redo row 38 and drop it)
        {
            const int nrows = SH::kE > 0 ? nr : NR;
            const int r2 = min(lane + 32, nrows - 1);
            double* q1 = S.Q + lane * LD;
            double* q2 = S.Q + r2 * LD;
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            int k = 0;
            for (; k + 3 < q; k += 4) {
                const double t0 = S.tmp[k], t1 = S.tmp[k + 1], t2 = S.tmp[k + 2], t3 = S.tmp[k + 3];
                const double u0 = q1[k], u1 = q1[k + 1], u2 = q1[k + 2], u3 = q1[k + 3];
                const double w0 = q2[k], w1 = q2[k + 1], w2 = q2[k + 2], w3 = q2[k + 3];
                a0 += u0 * t0; a1 += u1 * t1; b0 += w0 * t0; b1 += w1 * t1;
                a0 += u2 * t2; a1 += u3 * t3; b0 += w2 * t2; b1 += w3 * t3;
            }
            for (; k < q; k++) { const double t0 = S.tmp[k]; a0 += q1[k] * t0; b0 += q2[k] * t0; }
            const double sa = beta * (a0 + a1), sb = beta * (b0 + b1);
            const bool second = lane + 32 < nrows;
            if constexpr (kProj) {
                // z <- z + delta u, u = column j of Q H
                S.z[lane] += delta * (q1[j] - sa * vj);
                if (second) S.z[lane + 32] += delta * (q2[j] - sb * vj);
            }
            for (k = 0; k < j; k++) {
                const double t0 = S.tmp[k];
                q1[k] -= sa * t0;
                if (second) q2[k] -= sb * t0;
            }
        }
        // W <- W - beta (W v) v^T: lane r owns rows r and r + 32 (< q)
        for (int r = lane; r < q; r += 32) {
            double* wr = S.W + r * LD;
            double a0 = 0.0, a1 = 0.0;
            int k = 0;
            for (; k + 3 < q; k += 4) {
                const double w0 = wr[k], w1 = wr[k + 1], w2 = wr[k + 2], w3 = wr[k + 3];
                const double t0 = S.tmp[k], t1 = S.tmp[k + 1], t2 = S.tmp[k + 2], t3 = S.tmp[k + 3];
                a0 += w0 * t0; a1 += w1 * t1; a0 += w2 * t2; a1 += w3 * t3;
            }
            for (; k < q; k++) a0 += wr[k] * S.tmp[k];
            const double sa = beta * (a0 + a1);
            for (k = 0; k < j; k++) wr[k] -= sa * S.tmp[k];
        }
        __syncwarp();
        if (l != j)
            for (int k = lane; k < j; k += 32) S.W[l * LD + k] = S.W[j * LD + k];
    }
    if (lane == 0 && l != j) { S.act[l] = S.act[j]; S.lam[l] = S.lam[j]; }
    q = j;
    __syncwarp();
    return delta;
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

struct QpResultRegs {
    int q, iters, status;
    unsigned long long pairs_evaluated, passes;
    long long price_cycles;
    int warm;           // rows the warm start put into the working set (0: cold start)
    double worst_slack; // smallest unnormalised slack over the rows of the last pricing pass (<= 0; every row it skipped is
                        // provably satisfied): > -1e-6 by construction, below -1e-9 when a row sits inside the feasibility band
};

#ifdef LSCGPU_QP_SECTION_TIMERS      // build with -DLSCGPU_QP_SECTION_TIMERS and run with LSCGPU_QP_DEBUG=1: cycles per section
#define QP_TICK() (tk = clock64())
#define QP_TOCK(i) do { const long long now_ = clock64(); sec[i] += now_ - tk; tk = now_; } while (0)
#else
#define QP_TICK() ((void)0)
#define QP_TOCK(i) ((void)0)
#endif

// Stage the agent-independent tables and the agent's problem data. Call with all threads; ends with a barrier.
// boxes: this agent's SFC window [5][6] or null. S.x = x0 (equality-constrained minimiser).
template <int kThreads, class SH>
__device__ __forceinline__ void qp_stage(SH& S, const QpTablesDev& T, int ts, const double* st9, const double* gl3,
                                         const float* boxes, const float* wmin, const float* wmax, const AgentConstDev& ac,
                                         double slack_w = 1.0) {
    const int tid = threadIdx.x;
    if constexpr (SH::kE > 0) {
        // c_p = w_s (M - m) / M (src/traj_optimizer.cpp:386-387)
        if (tid < kM) S.isc_m[tid] = rsqrt(slack_w * ((double)(kM - tid) / (double)kM));
        if (tid == 0) { S.n_e = 0; S.overflow = 0; }
    }
    // the whitened basis of this ts (3120 B, contiguous, 16-byte aligned on both sides) travels by one bulk-async copy
    // that completes on S.g_bar; it is first needed by the factorisation update, long after the corridors are built
    static_assert((sizeof(double) * kAx * kFree) % 16 == 0 && offsetof(QpTablesDev, G) % 16 == 0 && offsetof(SH, G) % 16 == 0,
                  "bulk copy of G needs 16-byte alignment and size");
    if (tid == 0) bulk_stage_begin(&S.g_bar, S.G, &T.G[ts - 1][0][0], (unsigned)(sizeof(double) * kAx * kFree));
    for (int e = tid; e < kAx; e += kThreads) S.inv_gn[e] = 1.0 / T.gnorm[ts - 1][e];
    for (int e = tid; e < kM * 9; e += kThreads) S.inv_dyn[e] = 1.0 / T.dyn_norm[ts - 1][e / 9][e % 9];
    if (tid < 15) {
        const int m = tid / 3, k = tid % 3;
        double lo = (double)wmin[k], hi = (double)wmax[k];
        if (boxes) {        // SFC rows == per-variable bounds (src/traj_optimizer.cpp:409-434)
            lo = fmax(lo, (double)boxes[m * 6 + k]);
            hi = fmin(hi, (double)boxes[m * 6 + 3 + k]);
        }
        S.lb[tid] = lo; S.ub[tid] = hi;
    }
    if (tid < 3) { S.vmax[tid] = ac.vmax[tid]; S.amax[tid] = ac.amax[tid]; }
    for (int e = tid; e < kNv; e += kThreads) {
        const int k = e / kAx, i = e % kAx;
        const double* Xs = T.Xs[ts - 1][i];
        S.x[e] = Xs[0] * st9[k] + Xs[1] * st9[3 + k] + Xs[2] * st9[6 + k] + T.xg[ts - 1][i] * gl3[k];
    }
    if (tid == 0) { S.travelled = 0.0; S.stop = 0; }
    __syncthreads();
}

// ---- building blocks of the factorisation update (warp 0; every lane calls them) ---------------------------------
// unit whitened normal nv = (G (+) G (+) G)^T a / |.| of a decoded row: lane c owns coordinates c and c + 32
template <class SH>
__device__ __forceinline__ void row_normal_regs(const SH& S, const RowRegs& row, int lane, int c_axis0, int c_col0, double* nv_reg) {
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
                if (t < row.nnz && ax_t == k) sacc += row.a[t] * S.G[(row.idx[t] - ax_t * kAx) * kFree + cc];
            }
            nv_reg[h] = sacc * row.inv_len;
        } else if (SH::kE > 0 && c - NR == row.ec) {
            nv_reg[h] = row.ea * row.inv_len;
        }
    }
}

// z = (I - Q Q^T) nv by Gram-Schmidt (second pass when the first cancelled most of the vector); d = Q^T nv.
// Writes S.z, S.d (and S.tmp); returns |z|^2. nn = |nv|^2.
template <class SH>
__device__ __forceinline__ double gs_project(SH& S, int q, int nr, const double* nv_reg, double nn, int lane) {
    constexpr int LD = SH::LDX;
    __syncwarp();
    S.z[lane] = nv_reg[0]; S.d[lane] = 0.0;
    if (lane + 32 < nr) { S.z[lane + 32] = nv_reg[1]; S.d[lane + 32] = 0.0; }
    __syncwarp();
    double zz = nn;
    if (q > 0) {
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            // lane k: column k of Q against z. Three chunks of 13 rows: the 13 column entries are loaded
            // into registers first (independent shared-memory loads in flight together), then multiplied
            // against the broadcast z values.
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
                if constexpr (SH::kE > 0) {
#pragma unroll 1
                    for (int r = NR; r < nr; r++) s0 += qc[r * LD] * S.z[r];
                }
                const double sdot = s0 + s1 + s2;
                S.tmp[lane] = sdot;
                S.d[lane] += sdot;
            }
#pragma unroll 1
            for (int k = lane + 32; k < q; k += 32) {     // q > 32 only
                double sdot = 0.0;
#pragma unroll 1
                for (int r = 0; r < nr; r++) sdot += S.Q[r * LD + k] * S.z[r];
                S.tmp[k] = sdot;
                S.d[k] += sdot;
            }
            __syncwarp();
            // lane r: rows r and r + 32 of Q against the coefficients, both in one pass over k
            // (lanes without a second row read the last row and drop the result)
            double zp;
            {
                const int r2 = min(lane + 32, nr - 1);
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
                if (lane + 32 < nr) {
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
    return zz;
}

// rr = W d (change of the active multipliers per unit step): lane k owns row k; four columns per trip, loads staged
// before the FMAs. Ends with a warp barrier.
template <class SH>
__device__ __forceinline__ void w_times_d(SH& S, int q, int lane) {
    constexpr int LD = SH::LDX;
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
}

// the row becomes active: new basis column z / |z|, new column (-rr / |z|, 1 / |z|) of W; q grows by one
template <class SH>
__device__ __forceinline__ void add_column(SH& S, int& q, int nr, double rsq, int id, double lam_new, int lane) {
    constexpr int LD = SH::LDX;
    for (int r = lane; r < nr; r += 32) S.Q[r * LD + q] = S.z[r] * rsq;
    for (int k = lane; k < q; k += 32) { S.W[k * LD + q] = -S.rr[k] * rsq; S.W[q * LD + k] = 0.0; }
    if (lane == 0) {
        S.W[q * LD + q] = rsq;
        S.act[q] = id;
        S.lam[q] = lam_new;
    }
    q++;
    __syncwarp();
}

// x += (G (+) G (+) G) vacc by the whole block: element e = axis * 30 + var, 13 products each
template <int kThreads, class SH>
__device__ __forceinline__ void apply_vacc(SH& S, int tid) {
    for (int e = tid; e < kNv; e += kThreads) {
        const int k = e / kAx;
        const double* g = S.G + (e - k * kAx) * kFree;
        const double* va = S.vacc + k * kFree;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int c = 0; c + 1 < kFree; c += 2) { s0 += g[c] * va[c]; s1 += g[c + 1] * va[c + 1]; }
        s0 += g[kFree - 1] * va[kFree - 1];
        S.x[e] += s0 + s1;
    }
}

// Warm start. S.act[0 .. n_guess) holds candidate rows: the bounds and dynamic limits active at the agent's previous solve
// (the plan keeps its shape relative to the horizon, so most of them are active again). They are factorised at once,
// without pricing, ratio tests or steps, and v = argmin |v|^2 subject to n_k . v = b_k on them is taken as the starting
// point iff every multiplier is >= 0: then (v, candidates) is an S-pair of Goldfarb-Idnani and the iteration converges
// from it to the same unique minimiser. Candidates with a negative multiplier are dropped once (Householder drops) and the
// rest re-checked; otherwise the solve starts cold.
//
// Every such row touches one axis only and the three axes share one whitened basis, so the candidate normals are block
// diagonal: three independent Gram-Schmidt factorisations of at most 13 columns in 13 dimensions, done by warps 0, 1, 2
// side by side (columns of axis k follow those of the axes before it; everything off the diagonal blocks of Q and W is
// zero). Call with all threads of the block; returns the number of active rows (0: cold start), the same in every
// thread. On success S.lam holds the multipliers, S.vacc the whitened step from x0 (the caller applies it to x),
// S.travelled its length.
__device__ __forceinline__ int fixed_row_axis(int id) {
    if (id < 0 || id >= kFixedRows) return -1;
    return id < 180 ? (id >> 1) / kAx : ((id - 180) >> 1) / 45;
}

template <int kThreads, class SH>
__device__ __forceinline__ int qp_warm_start(SH& S, const RowSrc& rows, int n_guess, double vel_coef, double acc_coef) {
    constexpr int LD = SH::LDX;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cand0 = lane < n_guess ? S.act[lane] : -1, cand1 = lane + 32 < n_guess ? S.act[lane + 32] : -1;
    if (tid < 4) S.best_id[4 + tid] = 0;                    // [4..6] failure flags of the axes, [7] result
    __syncthreads();                                        // everybody holds the candidates; S.act is free
    for (int e = tid; e < NR * LD; e += kThreads) { S.Q[e] = 0.0; S.W[e] = 0.0; }
    const int ax0 = fixed_row_axis(cand0), ax1 = fixed_row_axis(cand1);
    unsigned lo[3], hi[3];
    int cnt[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        lo[a] = __ballot_sync(0xffffffffu, ax0 == a); hi[a] = __ballot_sync(0xffffffffu, ax1 == a);
        cnt[a] = __popc(lo[a]) + __popc(hi[a]);
    }
    const int q_all = cnt[0] + cnt[1] + cnt[2];
    __syncthreads();                                        // Q, W zeroed
    if (warp < 3 && q_all > 0 && q_all <= NR) {
        const int k = warp;
        const unsigned mlo = k == 0 ? lo[0] : (k == 1 ? lo[1] : lo[2]), mhi = k == 0 ? hi[0] : (k == 1 ? hi[1] : hi[2]);
        const int nk = k == 0 ? cnt[0] : (k == 1 ? cnt[1] : cnt[2]);
        const int off = k == 0 ? 0 : (k == 1 ? cnt[0] : cnt[0] + cnt[1]);
        const int nlo = __popc(mlo);
        double* zk = S.z + k * kFree;                       // this axis' 13 coordinates of the work vectors
        bool failed = nk > kFree;
        for (int j = 0; j < nk && !failed; j++) {
            // j-th candidate of this axis (candidate order)
            const int src = j < nlo ? __fns(mlo, 0, j + 1) : __fns(mhi, 0, j - nlo + 1);
            const int id = __shfl_sync(0xffffffffu, j < nlo ? cand0 : cand1, src);
            const RowRegs row = decode_row(id, S, rows, vel_coef, acc_coef);
            if (!(row.inv_len < INFINITY)) { failed = true; break; }
            double slack = -row.b;
#pragma unroll
            for (int t = 0; t < 3; t++) if (t < row.nnz) slack += row.a[t] * S.x[row.idx[t]];
            slack *= row.inv_len;
            // unit normal restricted to the axis: lane r < 13 owns coordinate r
            double zr = 0.0;
            if (lane < kFree) {
#pragma unroll
                for (int t = 0; t < 3; t++)
                    if (t < row.nnz) zr += row.a[t] * S.G[(row.idx[t] - k * kAx) * kFree + lane];
                zr *= row.inv_len;
                zk[lane] = zr;
            }
            if (lane < j) S.d[off + lane] = 0.0;
            __syncwarp();
            double zz = row.nn;
            const int c = off + j;
            if (j > 0) {
#pragma unroll 1
                for (int pass = 0; pass < 2; pass++) {
                    if (lane < j) {                         // lane i: column off + i against z
                        const double* qc = S.Q + (k * kFree) * LD + off + lane;
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int r = 0; r + 1 < kFree; r += 2) { s0 += qc[r * LD] * zk[r]; s1 += qc[(r + 1) * LD] * zk[r + 1]; }
                        s0 += qc[(kFree - 1) * LD] * zk[kFree - 1];
                        const double sdot = s0 + s1;
                        S.tmp[off + lane] = sdot;
                        S.d[off + lane] += sdot;
                    }
                    __syncwarp();
                    double zp = 0.0;
                    if (lane < kFree) {                     // lane r: row r of the block against the coefficients
                        const double* qr = S.Q + (k * kFree + lane) * LD + off;
                        double a0 = 0.0, a1 = 0.0;
                        int i = 0;
                        for (; i + 1 < j; i += 2) { a0 += qr[i] * S.tmp[off + i]; a1 += qr[i + 1] * S.tmp[off + i + 1]; }
                        if (i < j) a0 += qr[i] * S.tmp[off + i];
                        zr -= a0 + a1;
                        zk[lane] = zr;
                        zp = zr * zr;
                    }
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) zp += __shfl_xor_sync(0xffffffffu, zp, o);      // lanes 0..15
                    const double zz_new = __shfl_sync(0xffffffffu, zp, 0);
                    __syncwarp();
                    const bool again = zz_new < 0.25 * zz;
                    zz = zz_new;
                    if (!again) break;
                }
            }
            if (!(zz > 1e-10)) { failed = true; break; }    // (numerically) dependent on the columns taken so far
            const double rsq = rsqrt(zz);
            if (lane < j) {                                 // rr = W d inside the block, new column of W
                const double* wk = S.W + (off + lane) * LD + off;
                double a0 = 0.0;
                for (int i = lane; i < j; i++) a0 += wk[i] * S.d[off + i];          // W is upper triangular here
                S.W[(off + lane) * LD + c] = -a0 * rsq;
            }
            if (lane < kFree) S.Q[(k * kFree + lane) * LD + c] = zr * rsq;
            if (lane == 0) { S.W[c * LD + c] = rsq; S.act[c] = id; S.lam[c] = -slack; }      // S.lam: b_k until the multipliers are known
            __syncwarp();
        }
        if (failed && lane == 0) S.best_id[4 + k] = 1;
    }
    __syncthreads();
    if (warp == 0) {
        int q = (q_all > 0 && q_all <= NR && !(S.best_id[4] | S.best_id[5] | S.best_id[6])) ? q_all : 0;
        int result = 0;
        for (int attempt = 0; attempt < 2 && q > 0; attempt++) {
            // y = W^T b (S.tmp), lambda = W y (S.rr)
            for (int j = lane; j < q; j += 32) {
                double s0 = 0.0;
                for (int kk = 0; kk < q; kk++) s0 += S.W[kk * LD + j] * S.lam[kk];
                S.tmp[j] = s0;
            }
            __syncwarp();
            bool neg0 = false, neg1 = false;
            for (int kk = lane; kk < q; kk += 32) {
                const double* wk = S.W + kk * LD;
                double s0 = 0.0;
                for (int j = 0; j < q; j++) s0 += wk[j] * S.tmp[j];
                S.rr[kk] = s0;
                if (s0 < -1e-12) { if (kk < 32) neg0 = true; else neg1 = true; }
            }
            __syncwarp();
            const unsigned m0 = __ballot_sync(0xffffffffu, neg0), m1 = __ballot_sync(0xffffffffu, neg1);
            if (m0 == 0u && m1 == 0u) {
                // accepted: multipliers, v = Q y, its length
                for (int kk = lane; kk < q; kk += 32) S.lam[kk] = fmax(S.rr[kk], 0.0);
                double vp = 0.0;
                for (int r = lane; r < NR; r += 32) {
                    const double* qr = S.Q + r * LD;
                    double s0 = 0.0;
                    for (int j = 0; j < q; j++) s0 += qr[j] * S.tmp[j];
                    S.vacc[r] = s0;
                    vp += s0 * s0;
                }
                const double vv = warp_sum(vp);
                if (lane == 0) S.travelled = sqrt(vv) * (1.0 + 1e-9) + 1e-13;
                result = q;
                break;
            }
            // a few stale candidates are worth a Householder drop each; with many the previous working set says little
            // about this step's and a cold start is cheaper than repairing it
            if (attempt == 1 || __popc(m0) + __popc(m1) > 3) break;
            // drop the candidates with a negative multiplier, highest index first (a drop moves the last row into the
            // hole, and everything behind the hole has been looked at already)
            for (int l = q - 1; l >= 0; l--) {
                const bool neg = l < 32 ? ((m0 >> l) & 1u) : ((m1 >> (l - 32)) & 1u);
                if (neg) drop_active(S, q, l, lane, NR);
            }
        }
        if (lane == 0) S.best_id[7] = result;
    }
    __syncthreads();
    return S.best_id[7];
}

// The active-set solve. Preconditions: qp_stage done (S.x = x0), the row source holds n_kept pairs with their gates.
// open_lists: kThreads / 32 lists of kWarpList ints in shared memory. Every thread returns the same result.
template <int kThreads, class SH>
__device__ __forceinline__ QpResultRegs qp_solve_core(SH& S, int* open_lists, const RowSrc& rows, int n_kept,
                                                      double vel_coef, double acc_coef, int max_iter, long long* sec,
                                                      bool mirror_rows = false, int n_guess = 0) {
    constexpr int kE = SH::kE, NRX = SH::NRX, LD = SH::LDX;
    constexpr int kWarps = kThreads / 32;
    constexpr int kItems = (225 + kThreads - 1) / kThreads;
    constexpr int kGate = kWarpList / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* my_list = open_lists + warp * kWarpList;
#ifdef LSCGPU_QP_SECTION_TIMERS
    long long tk = 0;
#endif
    (void)sec;

    bulk_stage_wait(&S.g_bar);      // G has landed (the copy was issued before the corridor phase)
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
                F.lo[t] = vel ? S.vmax[k] : S.amax[k]; F.inv[t] = S.inv_dyn[m * 9 + j];
            }
        }
    }
    // warp 0: per-lane index constants of the factorisation update (no divisions in the loop)
    const int c_axis0 = lane / kFree, c_col0 = lane - c_axis0 * kFree;      // whitened coordinate c = lane
    QpResultRegs R{0, 0, LSCGPU_QP_OK, 0ull, 0ull, 0ll, 0, 0.0};
    int q = 0, iters = 0, status = LSCGPU_QP_OK;
    double raw_min = 0.0;

    if (kE == 0 && n_guess > 0) {       // block-uniform
        // warm start from the rows in S.act[0 .. n_guess) (qp_warm_start); x is brought up to date by the block
        QP_TICK();
        R.warm = qp_warm_start<kThreads>(S, rows, n_guess, vel_coef, acc_coef);
        if (R.warm > 0) { q = R.warm; apply_vacc<kThreads>(S, tid); }
        QP_TOCK(7);
    }

    while (true) {
        // ---- pricing by the whole block ---------------------------------------------------------------------------
        Best best{0.0, -1};
        const long long t_price = clock64();
        __syncthreads();        // x of the previous update is complete (block-wide update below)
        const double travelled = S.travelled;
        raw_min = 0.0;          // smallest (unnormalised) slack this thread sees in this pass
#pragma unroll
        for (int t = 0; t < kItems; t++) {
            if (F.base[t] < 0) continue;
            const double* c = S.x + F.base[t];
            if (F.kind[t] == 0) {
                consider(best, c[0] - F.lo[t], F.inv[t], F.id[t]);
                consider(best, F.hi[t] - c[0], F.inv[t], F.id[t] + 1);
                raw_min = fmin(raw_min, fmin(c[0] - F.lo[t], F.hi[t] - c[0]));
            } else {
                const double expr = F.kind[t] == 1 ? vel_coef * (c[1] - c[0]) : acc_coef * (c[2] - 2.0 * c[1] + c[0]);
                consider(best, F.lo[t] - expr, F.inv[t], F.id[t]);
                consider(best, F.lo[t] + expr, F.inv[t], F.id[t] + 1);
                raw_min = fmin(raw_min, F.lo[t] - fabs(expr));
            }
        }
        for (int base = warp * kWarpList; base < n_kept; base += kWarps * kWarpList) {
            double gv[kGate];
#pragma unroll
            for (int h = 0; h < kGate; h++) {
                const int slot = base + 32 * h + lane;
                gv[h] = slot < n_kept ? rows.gate(slot) : INFINITY;
            }
            int n_open = 0;
#pragma unroll
            for (int h = 0; h < kGate; h++) {
                const bool open = !(gv[h] > travelled);             // may be violated by now
                const unsigned mask = __ballot_sync(0xffffffffu, open);
                if (open) my_list[n_open + __popc(mask & ((1u << lane) - 1u))] = base + 32 * h + lane;
                n_open += __popc(mask);
            }
            __syncwarp();
            for (int idx = lane; idx < n_open; idx += 32) {
                const int slot = my_list[idx];
                float4 nr; double r6[6]; int m;
                rows.template load<(kE > 0)>(slot, nr, r6, m);
                const double mu_min = price_pair(best, S, slot, m, nr, r6, raw_min, kE > 0 ? rows.code(slot) : 0);
                // 1e-6 relative margin: the stored 1/|a| is float32, so mu carries ~6e-8 relative error
                rows.set_gate(slot, travelled + (mu_min > 0.0 ? mu_min * 0.999999 : mu_min));
                R.pairs_evaluated++;
            }
            __syncwarp();       // the list is rewritten by the next slice
        }
        if constexpr (kE > 0) {
            // bounds eps_c <= 0 of the coordinates in use (whitened slack -e_c)
            if (warp == 0)
                for (int c = lane; c < S.n_e; c += 32) consider(best, -S.e[c] * S.e_isc[c], 1.0 / S.e_isc[c], kSlackBoundBase + c);
        }
        R.passes++;
        best = warp_argmin(best);
        if (lane == 0) { S.best_mu[warp] = best.mu; S.best_id[warp] = best.id; }
        __syncthreads();
        {   // every warp reduces the per-warp results again (lane w holds warp w's): same answer in all threads
            Best wb{0.0, -1};
            if (lane < kWarps) { wb.mu = S.best_mu[lane]; wb.id = S.best_id[lane]; }
            best = warp_argmin(wb);
        }
        if (tid == 0) R.price_cycles += clock64() - t_price;
        if (best.id < 0) break;                 // no row violated beyond the tolerance anywhere: done (block-uniform)

        // ---- factorisation update by warp 0 -----------------------------------------------------------------------
        if (warp == 0) {
            bool done = false;          // set when the solve must stop (failure)
            QP_TICK();
            S.vacc[lane] = 0.0;
            if (lane + 32 < NRX + 1) S.vacc[lane + 32] = 0.0;
            do {
                {
                    bool dup = false;
                    for (int k = lane; k < q; k += 32) dup |= S.act[k] == best.id;
                    if (__any_sync(0xffffffffu, dup)) { status = LSCGPU_QP_MAXITER; done = true; break; }   // numerical breakdown
                }
                if constexpr (kE > 0) {
                    // first row of a slack pair to enter: its slack variable becomes coordinate NR + c (e_c = 0, a zero row
                    // of Q: no active normal touches it yet)
                    if (best.id >= kFixedRows && best.id < kSlackBoundBase) {
                        const int slot = (best.id - kFixedRows) / 6;
                        if (rows.code(slot) == kSlackUntouched) {
                            const int c = S.n_e;
                            if (c >= kE) { if (lane == 0) S.overflow = 1; status = LSCGPU_QP_MAXITER; done = true; break; }
                            const int m = rows.seg(slot);
                            __syncwarp();
                            if (lane == 0) {
                                S.e[c] = 0.0; S.e_isc[c] = S.isc_m[m]; S.e_slot[c] = slot; S.n_e = c + 1;
                                rows.set_code(slot, c + 1, mirror_rows);
                            }
                            for (int k = lane; k < LD; k += 32) S.Q[(NR + c) * LD + k] = 0.0;
                            __syncwarp();
                        }
                    }
                }
                const int nr = kE > 0 ? NR + S.n_e : NR;        // rows of Q in use
                const RowRegs row = decode_row(best.id, S, rows, vel_coef, acc_coef);
                if (!(row.inv_len < INFINITY)) { status = LSCGPU_QP_INFEASIBLE; done = true; break; }
                double nv_reg[2];
                row_normal_regs(S, row, lane, c_axis0, c_col0, nv_reg);
                // slack of the selected row (normalised); along the step it grows by t |z|^2 (a . G z = |G^T a| nv . z and
                // nv . z = z . z for the projection z of nv), so x itself is only brought up to date once per update
                double slack = -row.b;
#pragma unroll
                for (int t = 0; t < 3; t++) if (t < row.nnz) slack += row.a[t] * S.x[row.idx[t]];
                if constexpr (kE > 0) { if (row.ec >= 0) slack += row.ea * S.e[row.ec]; }
                slack *= row.inv_len;
                double lam_p = 0.0, zz = 0.0;
                bool fresh = true;
                QP_TOCK(1);
                while (true) {
                    if (++iters > max_iter) { status = LSCGPU_QP_MAXITER; done = true; break; }
                    if (fresh) zz = gs_project(S, q, nr, nv_reg, row.nn, lane);     // after a drop z, d and |z|^2 are carried over
                    QP_TOCK(2);
                    w_times_d(S, q, lane);
                    // ratio test over the active multipliers: warp minimum of lam / rr through its order-preserving
                    // 64-bit key (two 32-bit min reductions), ties to the smallest index
                    double t1 = INFINITY;
                    int l = -1;
                    {
                        unsigned long long key = ~0ull;
                        int kbest = -1;
                        for (int k = lane; k < q; k += 32) {
                            const double rk = S.rr[k];
                            if (rk > kZeroTol) {
                                const unsigned long long kk = order_key(S.lam[k] * fast_rcp(rk));
                                if (kk < key) { key = kk; kbest = k; }
                            }
                        }
                        const unsigned hi = (unsigned)(key >> 32);
                        const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
                        const unsigned lo = hi == mhi ? (unsigned)key : 0xffffffffu;
                        const unsigned mlo = __reduce_min_sync(0xffffffffu, lo);
                        const bool win = kbest >= 0 && hi == mhi && (unsigned)key == mlo;
                        const unsigned wl = __reduce_min_sync(0xffffffffu, win ? (unsigned)kbest : 0xffffffffu);
                        if (wl != 0xffffffffu) {
                            l = (int)wl;
                            t1 = order_unkey(((unsigned long long)mhi << 32) | mlo);
                        }
                    }
                    const bool primal = zz > kZeroTol;
                    const double rsq = primal ? rsqrt(zz) : 0.0;      // 1/|z|; |z| = zz rsq, 1/|z|^2 = rsq^2
                    double t2 = primal ? -slack * (rsq * rsq) : INFINITY;
                    if (t2 < 0.0) t2 = 0.0;
                    const double t = fmin(t1, t2);
                    if (!(t < INFINITY)) { status = LSCGPU_QP_INFEASIBLE; done = true; break; }
                    for (int k = lane; k < q; k += 32) S.lam[k] -= t * S.rr[k];
                    lam_p += t;
                    QP_TOCK(3);
                    if (!primal) {
                        const double dl = drop_active<SH, true>(S, q, l, lane, nr);
                        zz += dl * dl; fresh = false;
                        QP_TOCK(6);
                        continue;
                    }
                    if (lane == 0) S.travelled += t * (zz * rsq) * (1.0 + 1e-9) + 1e-13;
                    S.vacc[lane] += t * S.z[lane];
                    if (lane + 32 < nr) S.vacc[lane + 32] += t * S.z[lane + 32];
                    slack += t * zz;
                    QP_TOCK(4);
                    if (t2 <= t1) {
                        add_column(S, q, nr, rsq, best.id, lam_p, lane);
                        QP_TOCK(5);
                        break;
                    }
                    {
                        const double dl = drop_active<SH, true>(S, q, l, lane, nr);
                        zz += dl * dl; fresh = false;
                    }
                    QP_TOCK(6);
                }
            } while (false);
            if (done && lane == 0) S.stop = 1;
        }
        __syncthreads();
        const bool stop = S.stop != 0;
        // x += (G (+) G (+) G) vacc by the whole block: element e = axis * 30 + var, 13 products each (the barrier at the
        // top of the next pass makes it visible)
        apply_vacc<kThreads>(S, tid);
        if constexpr (kE > 0) { if (tid < S.n_e) S.e[tid] += S.vacc[NR + tid]; }
        if (stop) break;
    }
    __syncthreads();
    // q, iters and status live in warp 0; hand them to everybody; block minimum of the last pass' smallest slack
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) raw_min = fmin(raw_min, __shfl_xor_sync(0xffffffffu, raw_min, o));
    if (lane == 0) S.best_mu[warp] = raw_min;
    if (tid == 0) { S.act[NRX] = q; S.best_id[0] = iters; S.best_id[1] = status; }
    __syncthreads();
    R.q = S.act[NRX]; R.iters = S.best_id[0]; R.status = S.best_id[1];
    R.worst_slack = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) R.worst_slack = fmin(R.worst_slack, S.best_mu[w]);
    return R;
}

// objective as the reference reports it (getObjValue incl. the constant of the terminal cost); call with a full warp
// (with slack variables: plus their cost sum c_p eps_p^2 = |e|^2)
template <class SH>
__device__ __forceinline__ double qp_objective(const SH& S, const QpTablesDev& T, int ts, const double* gl3, int lane) {
    double jpart = 0.0;
    if constexpr (SH::kE > 0) { if (lane < S.n_e) jpart = S.e[lane] * S.e[lane]; }
    if (lane < 15) {
        const int k = lane / 5, m = lane % 5;
        const double* c = S.x + k * kAx + m * 6;
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) jpart += T.Qw[i][j] * c[i] * c[j];
        if (m >= kM - ts) { const double e = c[5] - gl3[k]; jpart += T.wT * e * e; }
    }
    return warp_sum(jpart);
}

}  // namespace lscgpu
