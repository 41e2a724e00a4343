// k_qp_solve — the Bernstein trajectory QP of one agent per warp (sm_100a, FP64).
//
// Replaces TrajOptimizer::solve (src/traj_optimizer.cpp:31-154): buildDeq (:239-259), populatebyrow (:261-539) and the
// CPLEX dual-simplex call (:76; IBM ILOG CPLEX 20.1, third party, not in the reference tree).
//
// Formulation (qp_tables.hpp): with the whitened null-space basis of the equality rows the QP is the least-distance
// problem  min |v|^2  s.t.  n_j . v >= -s_j(x0)  in 39 dimensions, x = x0 + (G (+) G (+) G) v. It is solved by a dual
// active-set method (Goldfarb-Idnani with an identity Hessian): start at the unconstrained minimiser v = 0, pick a
// violated row, move along the projection of its normal onto the null space of the active normals until the row is
// satisfied or an active multiplier reaches zero (then that row leaves), repeat. The orthogonal factor J (39x39) and
// the triangular factor R of the active normals live in shared memory; adding a row is one Householder reflection
// applied by all 32 lanes (one J row each), dropping a row is a sequence of Givens rotations.
//
// Rows are never assembled as a matrix: bounds (SFC boxes + world box), velocity and acceleration limits are priced
// from x directly; LSC rows are priced from the row store written by k_lsc_build (3 non-zeros each). LSC pricing is
// two-tier: the working set (pairs selected by k_lsc_build, plus every pair found violated later) is priced at every
// iteration; only when nothing in it is violated does the warp sweep ALL pairs of the agent (coalesced, pair index
// fastest across lanes). The solve ends when such a full sweep finds no violated row, so the result satisfies every
// row of the reference's QP and is its unique minimiser.
#include "kernels.hpp"

namespace lscgpu {

constexpr int NR = kRed;        // 39
constexpr int LD = 39;          // row pitch of J and R (odd pitch: row-per-lane accesses are bank-conflict free)
// Primal feasibility tolerance = CPLEX's default EpRHS (the reference sets no tolerance, src/traj_optimizer.cpp:42-54):
// like a dual simplex, a row enters the working set only when violated by more than this; entered rows are then met
// exactly. Trajectories travel as float32, so agents in contact see hulls ~1e-7 closer than r_i + r_j; an exact
// solver would call that infeasible where CPLEX answers "optimal".
constexpr double kFeasTol = 1e-6;
constexpr double kZeroTol = 1e-13;

struct SelectedRow {
    int nnz;
    int idx[3];
    double a[3];
    double b;
};

struct QpShared {
    double J[NR * LD];
    double R[NR * LD];
    double G[kAx * kFree];
    double x[kNv];
    double v[NR], nv[NR], d[NR], z[NR], rr[NR], lam[NR], u[NR];
    double inv_gn[kAx];
    double inv_dyn[45];
    double lb[15], ub[15];      // [m][axis]
    double vmax[3], amax[3];
    int act[NR];
    SelectedRow sel;
};

struct Best {
    double mu;
    int id;
};

__device__ __forceinline__ bool row_is_active(const QpShared& S, int q, int id) {
    for (int k = 0; k < q; k++)
        if (S.act[k] == id) return true;
    return false;
}
// slack: a.x - b of the row; scale: 1 / (whitened length of its normal)
__device__ __forceinline__ void consider(Best& b, const QpShared& S, int q, double slack, double scale, int id) {
    if (!(slack < -kFeasTol)) return;
    const double mu = scale < INFINITY ? slack * scale : -INFINITY;   // zero normal with positive rhs: infeasible row
    if (mu < b.mu && !row_is_active(S, q, id)) { b.mu = mu; b.id = id; }
}
__device__ __forceinline__ Best warp_argmin(Best b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double mu = __shfl_xor_sync(0xffffffffu, b.mu, o);
        const int id = __shfl_xor_sync(0xffffffffu, b.id, o);
        if (id >= 0 && (b.id < 0 || mu < b.mu || (mu == b.mu && id < b.id))) { b.mu = mu; b.id = id; }
    }
    return b;
}

// bounds + dynamic limits (ids 0..449)
__device__ __forceinline__ void price_fixed(Best& best, const QpShared& S, int q, int lane, double vel_coef,
                                            double acc_coef) {
    for (int var = lane; var < kNv; var += 32) {
        const int k = var / kAx, mi = var % kAx, m = mi / 6, i = mi % 6;
        if (m == 0 && i < kPhi) continue;
        const double xv = S.x[var], ig = S.inv_gn[mi];
        consider(best, S, q, xv - S.lb[m * 3 + k], ig, var * 2);
        consider(best, S, q, S.ub[m * 3 + k] - xv, ig, var * 2 + 1);
    }
    for (int idx = lane; idx < 135; idx += 32) {
        const int k = idx / 45, rem = idx % 45, m = rem / 9, j = rem % 9;
        const int base = k * kAx + m * 6;
        double expr, lim;
        if (j < 5) {
            if (m == 0 && j < 2) continue;
            expr = vel_coef * (S.x[base + j + 1] - S.x[base + j]);
            lim = S.vmax[k];
        } else {
            const int i = j - 5;
            if (m == 0 && i == 0) continue;
            expr = acc_coef * (S.x[base + i + 2] - 2.0 * S.x[base + i + 1] + S.x[base + i]);
            lim = S.amax[k];
        }
        const double inv = S.inv_dyn[m * 9 + j];
        consider(best, S, q, lim - expr, inv, kFixedRows - 270 + idx * 2);
        consider(best, S, q, lim + expr, inv, kFixedRows - 270 + idx * 2 + 1);
    }
}

// one (obstacle, segment) pair: up to 6 rows. Returns true when a row of the pair is violated beyond the tolerance.
__device__ __forceinline__ bool price_pair(Best& best, const QpShared& S, int q, int p, int n_obs, const float4* nrm,
                                             const double* rhs, size_t pitch) {
    const int m = p / n_obs;
    const float4 nr = nrm[p];
    const double ax = (double)nr.x, ay = (double)nr.y, az = (double)nr.z, inv = (double)nr.w;
    bool violated = false;
    for (int i = (m == 0 ? kPhi : 0); i < 6; i++) {
        const int vi = m * 6 + i;
        const double slack = ax * S.x[vi] + ay * S.x[kAx + vi] + az * S.x[2 * kAx + vi] - rhs[(size_t)i * pitch + p];
        violated |= slack < -kFeasTol;
        consider(best, S, q, slack, inv * S.inv_gn[vi], kFixedRows + p * 6 + i);
    }
    return violated;
}

__device__ __forceinline__ void decode_row(QpShared& S, int id, int n_obs, const float4* nrm, const double* rhs,
                                           size_t pitch, double vel_coef, double acc_coef) {
    SelectedRow& r = S.sel;
    if (id < 180) {
        const int var = id >> 1, side = id & 1;
        const int k = var / kAx, mi = var % kAx, m = mi / 6;
        r.nnz = 1; r.idx[0] = var;
        if (side == 0) { r.a[0] = 1.0; r.b = S.lb[m * 3 + k]; }
        else { r.a[0] = -1.0; r.b = -S.ub[m * 3 + k]; }
    } else if (id < kFixedRows) {
        const int e = id - 180, side = e & 1, idx = e >> 1;
        const int k = idx / 45, rem = idx % 45, m = rem / 9, j = rem % 9;
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
    } else {
        const int e = id - kFixedRows, p = e / 6, i = e % 6, m = p / n_obs, vi = m * 6 + i;
        const float4 nr = nrm[p];
        r.nnz = 3;
        r.idx[0] = vi; r.idx[1] = kAx + vi; r.idx[2] = 2 * kAx + vi;
        r.a[0] = (double)nr.x; r.a[1] = (double)nr.y; r.a[2] = (double)nr.z;
        r.b = rhs[(size_t)i * pitch + p];
    }
}

__device__ __forceinline__ double selected_slack(const QpShared& S) {
    double s = -S.sel.b;
    for (int t = 0; t < S.sel.nnz; t++) s += S.sel.a[t] * S.x[S.sel.idx[t]];
    return s;
}

// remove active row l: delete column l of R, restore the triangle with Givens rotations (rows j, j+1 of R,
// columns j, j+1 of J)
__device__ __forceinline__ void drop_active(QpShared& S, int& q, int l, int lane) {
    __syncwarp();
    for (int r = lane; r < q; r += 32) {
        for (int j = l; j < q - 1; j++) S.R[r * LD + j] = S.R[r * LD + j + 1];
        S.R[r * LD + q - 1] = 0.0;
    }
    if (lane == 0)
        for (int j = l; j < q - 1; j++) { S.act[j] = S.act[j + 1]; S.lam[j] = S.lam[j + 1]; }
    q--;
    for (int j = l; j < q; j++) {
        __syncwarp();
        const double a = S.R[j * LD + j], b = S.R[(j + 1) * LD + j];
        __syncwarp();
        if (b == 0.0) continue;
        const double h = hypot(a, b), c = a / h, s = b / h;
        for (int k = j + lane; k < q; k += 32) {
            const double t1 = S.R[j * LD + k], t2 = S.R[(j + 1) * LD + k];
            S.R[j * LD + k] = c * t1 + s * t2;
            S.R[(j + 1) * LD + k] = -s * t1 + c * t2;
        }
        for (int r = lane; r < NR; r += 32) {
            const double t1 = S.J[r * LD + j], t2 = S.J[r * LD + j + 1];
            S.J[r * LD + j] = c * t1 + s * t2;
            S.J[r * LD + j + 1] = -s * t1 + c * t2;
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32) k_qp_solve(QpLaunch L) {
    __shared__ QpShared S;
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    const bool batch = L.obs_offset != nullptr;
    const int agent = L.agent_index ? L.agent_index[b] : L.agent_base + b;
    const int di = batch ? b : agent;                 // index into state9/goal3/ts/boxes
    const int n_obs = batch ? (L.obs_offset[b + 1] - L.obs_offset[b]) : L.n_obs;
    const int P = kPairsPerObs * n_obs;
    const size_t pitch = (size_t)L.P_pad;
    const float4* nrm = batch ? L.nrm + (size_t)kPairsPerObs * L.obs_offset[b] : L.nrm + (size_t)b * L.P_pad;
    const double* rhs = batch ? L.rhs + (size_t)kPairsPerObs * L.obs_offset[b] : L.rhs + (size_t)b * 6 * L.P_pad;
    int* cand = L.cand + (size_t)b * L.cand_cap;
    const QpTablesDev& T = *L.T;
    const int ts = L.ts[di];
    const double vel_coef = T.vel_coef, acc_coef = T.acc_coef;
    const AgentConstDev& ac = L.consts[agent];

    // ---- stage tables and problem data ------------------------------------------------------------------------
    for (int e = lane; e < kAx * kFree; e += 32) S.G[e] = (&T.G[ts - 1][0][0])[e];
    for (int e = lane; e < kAx; e += 32) S.inv_gn[e] = 1.0 / T.gnorm[ts - 1][e];
    for (int e = lane; e < 45; e += 32) S.inv_dyn[e] = 1.0 / (&T.dyn_norm[ts - 1][0][0])[e];
    if (lane < 15) {
        const int m = lane / 3, k = lane % 3;
        double lo = (double)L.wmin[k], hi = (double)L.wmax[k];
        if (L.boxes) {      // SFC rows == per-variable bounds (src/traj_optimizer.cpp:409-434)
            const float* bx = L.boxes + (size_t)di * 30 + m * 6;
            lo = fmax(lo, (double)bx[k]);
            hi = fmin(hi, (double)bx[3 + k]);
        }
        S.lb[lane] = lo; S.ub[lane] = hi;
    }
    if (lane < 3) { S.vmax[lane] = ac.vmax[lane]; S.amax[lane] = ac.amax[lane]; }
    const double* st = L.state9 + (size_t)di * 9;
    const double* gl = L.goal3 + (size_t)di * 3;
    for (int e = lane; e < kNv; e += 32) {
        const int k = e / kAx, i = e % kAx;
        const double* Xs = T.Xs[ts - 1][i];
        S.x[e] = Xs[0] * st[k] + Xs[1] * st[3 + k] + Xs[2] * st[6 + k] + T.xg[ts - 1][i] * gl[k];
    }
    for (int e = lane; e < NR * LD; e += 32) { S.J[e] = 0.0; S.R[e] = 0.0; }
    __syncwarp();
    for (int e = lane; e < NR; e += 32) { S.J[e * LD + e] = 1.0; S.v[e] = 0.0; }
    __syncwarp();

    int q = 0, iters = 0, status = LSCGPU_QP_OK;
    int n_work = min(L.cand_count[b], L.cand_cap);
    unsigned long long rows_priced = 0, full_passes = 0;

    while (true) {
        // ---- pricing --------------------------------------------------------------------------------------------
        Best best{0.0, -1};
        price_fixed(best, S, q, lane, vel_coef, acc_coef);
        for (int w = lane; w < n_work; w += 32) price_pair(best, S, q, cand[w], n_obs, nrm, rhs, pitch);
        rows_priced += 414 + 6ull * n_work;
        best = warp_argmin(best);
        if (best.id < 0) {
            // nothing violated among bounds, limits and the working set: sweep every LSC pair of the agent
            full_passes++;
            rows_priced += 6ull * P;
            for (int p0 = 0; p0 < P; p0 += 32) {
                const int p = p0 + lane;
                bool viol = false;
                if (p < P) viol = price_pair(best, S, q, p, n_obs, nrm, rhs, pitch);
                const unsigned mask = __ballot_sync(0xffffffffu, viol);
                if (viol) {
                    const int slot = n_work + __popc(mask & ((1u << lane) - 1u));
                    if (slot < L.cand_cap) cand[slot] = p;
                }
                n_work = min(n_work + __popc(mask), L.cand_cap);
            }
            __syncwarp();
            best = warp_argmin(best);
            if (best.id < 0) break;             // optimal: no violated row anywhere
        }
        if (lane == 0) decode_row(S, best.id, n_obs, nrm, rhs, pitch, vel_coef, acc_coef);
        __syncwarp();
        // whitened normal  nv = (G (+) G (+) G)^T a
        double part = 0.0;
        for (int c = lane; c < NR; c += 32) {
            const int k = c / kFree, cc = c % kFree;
            double s = 0.0;
            for (int t = 0; t < S.sel.nnz; t++)
                if (S.sel.idx[t] / kAx == k) s += S.sel.a[t] * S.G[(S.sel.idx[t] % kAx) * kFree + cc];
            S.nv[c] = s;
            part += s * s;
        }
        const double nrm_len = sqrt(warp_sum(part));
        if (!(nrm_len > 0.0)) { status = LSCGPU_QP_INFEASIBLE; break; }
        __syncwarp();
        for (int c = lane; c < NR; c += 32) S.nv[c] /= nrm_len;
        __syncwarp();
        double lam_p = 0.0;
        bool fail = false;
        while (true) {
            if (++iters > L.max_iter) { status = LSCGPU_QP_MAXITER; fail = true; break; }
            // d = J^T nv
            double zz_part = 0.0;
            for (int c = lane; c < NR; c += 32) {
                double s = 0.0;
                for (int r = 0; r < NR; r++) s += S.J[r * LD + c] * S.nv[r];
                S.d[c] = s;
                if (c >= q) zz_part += s * s;
            }
            const double zz = warp_sum(zz_part);
            __syncwarp();
            // z = J2 d2 (step direction), rr = R^-1 d1 (change of the active multipliers)
            for (int r = lane; r < NR; r += 32) {
                double s = 0.0;
                for (int c = q; c < NR; c++) s += S.J[r * LD + c] * S.d[c];
                S.z[r] = s;
            }
            for (int k = lane; k < q; k += 32) S.rr[k] = S.d[k];
            for (int c = q - 1; c >= 0; c--) {
                __syncwarp();
                const double piv = S.rr[c] / S.R[c * LD + c];
                __syncwarp();
                for (int k = lane; k < c; k += 32) S.rr[k] -= S.R[k * LD + c] * piv;
                if (lane == 0) S.rr[c] = piv;
            }
            __syncwarp();
            // ratio test over the active multipliers
            double t1 = INFINITY;
            int l = -1;
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
            const bool primal = zz > kZeroTol;
            const double slack = selected_slack(S) / nrm_len;
            double t2 = primal ? -slack / zz : INFINITY;
            if (t2 < 0.0) t2 = 0.0;
            const double t = fmin(t1, t2);
            if (!(t < INFINITY)) { status = LSCGPU_QP_INFEASIBLE; fail = true; break; }
            for (int k = lane; k < q; k += 32) S.lam[k] -= t * S.rr[k];
            lam_p += t;
            if (!primal) { drop_active(S, q, l, lane); continue; }
            for (int c = lane; c < NR; c += 32) S.v[c] += t * S.z[c];
            for (int e = lane; e < kNv; e += 32) {
                const int k = e / kAx, i = e % kAx;
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < kFree; c++) s += S.G[i * kFree + c] * S.z[k * kFree + c];
                S.x[e] += t * s;
            }
            __syncwarp();
            if (t2 <= t1) {
                // add the row: Householder reflection H with (J2 H)^T nv = (alpha, 0, ..., 0)
                const double dq = S.d[q];
                const double alpha = dq > 0.0 ? -sqrt(zz) : sqrt(zz);
                if (q < NR - 1) {
                    for (int c = q + lane; c < NR; c += 32) S.u[c] = (c == q) ? dq - alpha : S.d[c];
                    const double uu = zz - dq * dq + (dq - alpha) * (dq - alpha);
                    const double beta = 2.0 / uu;
                    __syncwarp();
                    for (int r = lane; r < NR; r += 32) {
                        double s = 0.0;
                        for (int c = q; c < NR; c++) s += S.J[r * LD + c] * S.u[c];
                        s *= beta;
                        for (int c = q; c < NR; c++) S.J[r * LD + c] -= s * S.u[c];
                    }
                }
                for (int k = lane; k < q; k += 32) S.R[k * LD + q] = S.d[k];
                if (lane == 0) {
                    S.R[q * LD + q] = (q < NR - 1) ? alpha : dq;
                    S.act[q] = best.id;
                    S.lam[q] = lam_p;
                }
                q++;
                __syncwarp();
                break;
            }
            drop_active(S, q, l, lane);
        }
        if (fail) break;
    }
    __syncwarp();

    // ---- epilogue ---------------------------------------------------------------------------------------------
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
        atomicAdd(&L.counters->rows_priced, rows_priced);
        atomicAdd(&L.counters->qp_iterations, (unsigned long long)iters);
        atomicAdd(&L.counters->full_passes, full_passes);
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
        }
    }
}

void launch_qp_solve(const QpLaunch& L, cudaStream_t s) {
    if (L.n_problems <= 0) return;
    k_qp_solve<<<L.n_problems, 32, 0, s>>>(L);
}

}  // namespace lscgpu
