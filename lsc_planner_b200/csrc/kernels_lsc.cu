// k_predict, k_goal_plan, k_commit and the small glue kernels around the per-agent plan (sm_100a).
#include <cstdlib>
#include <string>

#include "gjk.cuh"
#include "kernels.hpp"
#include "sfc.cuh"

namespace lscgpu {

// ------------------------------------------------------------------------------------------------------------
// k_predict — one block of 96 threads per agent; thread e < 90 owns trajectory element e = (m*6+i)*3 + axis.
// Replaces obstaclePredictionWithPrevSol / ...WithCurrVel (src/traj_planner.cpp:829-864,699-712),
// initialTrajPlanningPrevSol / ...CurrVel (:997-1016,1030-1037), obstaclePredictionCheck / initialTrajPlanningCheck
// (:866-878,1047-1061) and getTerminalSegments (src/traj_optimizer.cpp:541-548). Because every agent runs the same
// planner on the same snapshot, agent j's initial trajectory IS the prediction every neighbour makes of j, so it is
// computed once — and so are the two checks: when the observed position is farther than reset_threshold from the start of
// the shifted previous trajectory, prediction and initial trajectory collapse to the position, the agent is marked for
// good (it is in everybody's obs_slack_indices from now on, and everybody is in its own) and its corridor is re-armed.
// ------------------------------------------------------------------------------------------------------------
// getTerminalSegments (src/traj_optimizer.cpp:541-548)
__device__ __forceinline__ int terminal_segments_of(F3 goal, F3 pos, double v_nom, double dt) {
    const F3 gd = f3_sub(goal, pos);
    const double ideal = __ddiv_rn(sqrt(f3_dot(gd, gd)), v_nom);
    const double horizon = __dmul_rn((double)kM, dt);
    const double q = __ddiv_rn(__dadd_rn(__dsub_rn(horizon, ideal), 1e-9), dt);
    int ts = (int)q;
    if (ts < 1) ts = 1;
    if (ts > kM) ts = kM;     // the reference throws here (src/traj_optimizer.cpp:356-358); unreachable for ideal >= 0
    return ts;
}

__global__ void __launch_bounds__(96) k_predict(PredictLaunch L) {
    const int a = blockIdx.x;
    const int e = threadIdx.x;
    const lscgpu_agent_in& in = L.in[a];
    // the check, by every thread alike: start of the shifted previous trajectory against the observed position
    bool reset = false;
    if (L.planner_seq >= 2) {
        const float* t = L.prev_traj + (size_t)a * kTrajFloats + 18;      // traj_curr[1][0]
        const F3 dlt = f3_sub(F3{t[0], t[1], t[2]}, F3{in.position[0], in.position[1], in.position[2]});
        reset = sqrt(f3_dot(dlt, dlt)) > L.reset_threshold;
    }
    if (e < kTrajFloats) {
        const int axis = e % 3, cp = e / 3, m = cp / 6, i = cp % 6;
        float val;
        if (L.planner_seq < 2) {
            // pos + (vel * m_intp) * dt with float3 arithmetic (octomath), m_intp = m + i/n in double
            const double m_intp = (double)m + (double)i / (double)kN;
            const float vel = in.velocity[axis];
            val = __fadd_rn(in.position[axis], __fmul_rn(__fmul_rn(vel, (float)m_intp), (float)L.dt));
        } else if (reset) {
            val = in.position[axis];                                       // :871-875, :1053-1057
        } else {
            const float* t = L.prev_traj + (size_t)a * kTrajFloats;
            val = (m < kM - 1) ? t[((m + 1) * 6 + i) * 3 + axis] : t[(kM * 6 - 1) * 3 + axis];
        }
        L.pred[(size_t)a * kTrajFloats + e] = val;
        L.predT[(size_t)e * L.n_pad + a] = val;
        if (axis == 2) {
            // z pre-scaled by the downwash ratio of a pair of agents with this agent's coefficient (the common case:
            // homogeneous swarm); k_lsc_build uses it whenever the pair's ratio is bitwise the same
            const AgentConstDev& c = L.consts[a];
            const double dw_self = (c.downwash * c.radius + c.downwash * c.radius) / (c.radius + c.radius);
            L.predZs[(size_t)cp * L.n_pad + a] = downwash_scaled_z(val, dw_self);
        }
        if (e < 9) {
            const float s = e < 3 ? in.position[e] : (e < 6 ? in.velocity[e - 3] : in.acceleration[e - 6]);
            L.state9[(size_t)a * 9 + e] = (double)s;
        } else if (e < 12) {
            L.goal3[(size_t)a * 3 + (e - 9)] = (double)in.goal[e - 9];
        }
    }
    __syncthreads();
    const float* o = L.pred + (size_t)a * kTrajFloats;
    if (e == 0) {
        const F3 pos{in.position[0], in.position[1], in.position[2]};
        int fl = 0;
        if (reset) {
            fl |= LSCGPU_FLAG_SLACK_NEEDED;
            L.reset_ever[a] = 1;                  // :870, :1049-1051 (sticky)
            *L.any_reset = 1;
            L.init_sfc[a] = 1;                    // :1059
        }
        L.flags[a] = fl;                          // the SFC warp of k_agent_plan adds its bit in the result record
        const F3 g{in.goal[0], in.goal[1], in.goal[2]};
        L.ts[a] = terminal_segments_of(g, pos, L.consts[a].v_nom, L.dt);
    }
    // reach bounds per (segment, axis): 15 threads, combined per segment below
    __shared__ float s_reach[kM][3], s_dreach[kM][3];
    if (e >= 64 && e < 64 + 3 * kM) {
        const int m = (e - 64) / 3, k = (e - 64) % 3;
        const AgentConstDev& c = L.consts[a];
        const float dtf = (float)L.dt;
        // Reachable offset of any control point of segment m from the current position, per axis, from the velocity
        // AND acceleration rows (src/traj_optimizer.cpp:469-525): the velocity control points v_t satisfy
        // |v_{t+1} - v_t| <= amax dt/(n-1) and |v_t| <= vmax (the first two are fixed by the state), consecutive
        // position control points differ by v_t dt/n, and C1/C2 continuity carries both bounds across segments.
        const float vmax = (float)c.vmax[k], dv = (float)c.amax[k] * dtf / (float)(kN - 1);
        float bound[4 * kM + 1];                          // fully unrolled below: stays in registers
        bound[0] = fabsf(in.velocity[k]);
        bound[1] = fabsf(in.velocity[k] + in.acceleration[k] * dtf / (float)(kN - 1));
#pragma unroll
        for (int t = 2; t <= 4 * kM; t++) bound[t] = fminf(vmax, bound[t - 1] + dv);
        // a feasible trajectory also keeps the fixed first two within vmax only if the state does; use them as is
        float reach = 0.f;
#pragma unroll
        for (int t = 0; t < 5 * kM; t++)
            if (t < 5 * (m + 1)) reach += bound[4 * (t / 5) + (t % 5)] * dtf / (float)kN;
        // Second bound, on x - c directly: both satisfy the same rows, so their velocity control points differ by at
        // most what the acceleration rows let them drift apart, 2 dv per step (and 2 vmax), starting from the mismatch
        // of the first two (zero unless the state was reset: the state IS the previous plan at t = dt). 1 % head room
        // on the limits for the 1e-6 row tolerance and the float32 rounding of c.
        const float v0c = (o[3 + k] - o[k]) * (float)kN / dtf, v1c = (o[6 + k] - o[3 + k]) * (float)kN / dtf;
        float dbound[4 * kM + 1];
        dbound[0] = fabsf(v0c - in.velocity[k]);
        dbound[1] = fabsf(v1c - (in.velocity[k] + in.acceleration[k] * dtf / (float)(kN - 1)));
#pragma unroll
        for (int t = 2; t <= 4 * kM; t++) dbound[t] = fminf(2.02f * vmax, dbound[t - 1] + 2.02f * dv);
        float dreach = fabsf(o[k] - in.position[k]);
#pragma unroll
        for (int t = 0; t < 5 * kM; t++)
            if (t < 5 * (m + 1)) dreach += dbound[4 * (t / 5) + (t % 5)] * dtf / (float)kN;
        s_reach[m][k] = reach; s_dreach[m][k] = dreach;
    }
    __syncthreads();
    if (e >= 32 && e < 32 + kM) {
        // Culling data of segment m. Bounding sphere of the 6 control points, and `reach`: an upper
        // bound of |x - c| for ANY point x the QP may give control point (m,i) and its initial_traj point c: the smaller
        // of |x - p| + |c - p| (p the current position) and a direct bound on x - c (below).
        const int m = e - 32;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        for (int i = 0; i < 6; i++) { cx += o[(m * 6 + i) * 3]; cy += o[(m * 6 + i) * 3 + 1]; cz += o[(m * 6 + i) * 3 + 2]; }
        cx *= (1.f / 6.f); cy *= (1.f / 6.f); cz *= (1.f / 6.f);
        float rad2 = 0.f, far2 = 0.f;
        for (int i = 0; i < 6; i++) {
            const float x = o[(m * 6 + i) * 3], y = o[(m * 6 + i) * 3 + 1], z = o[(m * 6 + i) * 3 + 2];
            rad2 = fmaxf(rad2, (x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz));
            const float dx = x - in.position[0], dy = y - in.position[1], dz = z - in.position[2];
            far2 = fmaxf(far2, dx * dx + dy * dy + dz * dz);
        }
        L.sphere[(size_t)m * L.n_pad + a] = make_float4(cx, cy, cz, sqrtf(rad2) * 1.0001f + 1e-5f);
        float r2 = 0.f, d2 = 0.f;
        for (int k = 0; k < 3; k++) { r2 += s_reach[m][k] * s_reach[m][k]; d2 += s_dreach[m][k] * s_dreach[m][k]; }
        const float r = fminf(sqrtf(r2) + sqrtf(far2), sqrtf(d2));
        L.reach[(size_t)a * kM + m] = r * 1.0001f + 1e-3f;
    }
    if (e == 40) {
        // bounding sphere of the whole trajectory: one 16-byte read per neighbour decides, for most pairs of a swarm
        // that is much larger than an agent's reach, that none of the five segment pairs can matter
        float cx = 0.f, cy = 0.f, cz = 0.f;
        for (int i = 0; i < 30; i++) { cx += o[i * 3]; cy += o[i * 3 + 1]; cz += o[i * 3 + 2]; }
        cx *= (1.f / 30.f); cy *= (1.f / 30.f); cz *= (1.f / 30.f);
        float rad2 = 0.f;
        for (int i = 0; i < 30; i++) {
            const float x = o[i * 3] - cx, y = o[i * 3 + 1] - cy, z = o[i * 3 + 2] - cz;
            rad2 = fmaxf(rad2, x * x + y * y + z * z);
        }
        L.tsphere[a] = make_float4(cx, cy, cz, sqrtf(rad2) * 1.0001f + 1e-5f);
    }
}

void launch_predict(const PredictLaunch& L, cudaStream_t s) { k_predict<<<L.n_agents, 96, 0, s>>>(L); }

// ------------------------------------------------------------------------------------------------------------
// k_goal_plan — goalPlanningWithPriority (src/traj_planner.cpp:540-608) for worlds WITHOUT an octomap, one block of
// 128 threads per agent. Every other agent is an obstacle seen at its current position with its desired goal and its
// previous trajectory (src/multi_sync_simulator.cpp:269-299). Higher-priority agents: closer to their goal than this
// agent (all of them when this agent is at its goal), not at their goal, and not moving away from this agent. If the
// nearest of them is closer than priority_dist_threshold the goal is a retreat point; otherwise the grid planner runs,
// but with distmap_obj == nullptr findLOSFreeGoal accepts every path point (src/grid_based_planner.cpp:350-407), so
// the line-of-sight goal is the desired goal whatever path A* returns: only the clip to goal_radius around the end of
// the initial trajectory remains. Float arithmetic of octomath::Vector3 with explicit roundings.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ F3 f3_normalized(F3 v) {
    const double len = sqrt(f3_dot(v, v));
    if (len > 0.0) { const float l = (float)len; v.x = __fdiv_rn(v.x, l); v.y = __fdiv_rn(v.y, l); v.z = __fdiv_rn(v.z, l); }
    return v;
}

__global__ void __launch_bounds__(128) k_goal_plan(GoalLaunch L) {
    __shared__ double s_dist[4];
    __shared__ int s_idx[4];
    const int a = blockIdx.x, tid = threadIdx.x;
    const lscgpu_agent_in& me = L.in[a];
    const F3 pos{me.position[0], me.position[1], me.position[2]};
    const F3 desired{me.goal[0], me.goal[1], me.goal[2]};
    const F3 to_goal = f3_sub(pos, desired);
    const double dist_to_goal = sqrt(f3_dot(to_goal, to_goal));
    double best = 1e9;                                   // SP_INFINITY
    int best_j = -1;
    const bool self_reset = L.reset_ever[a] != 0;
    for (int j = tid; j < L.n_agents; j += 128) {
        if (j == a) continue;
        if (self_reset || L.reset_ever[j]) continue;      // obs_slack_indices: high priority, never the retreat target (:548-551)
        const lscgpu_agent_in& o = L.in[j];
        const F3 op{o.position[0], o.position[1], o.position[2]}, og{o.goal[0], o.goal[1], o.goal[2]};
        const F3 d1 = f3_sub(op, og), d2 = f3_sub(op, pos);
        const double obs_dist_to_goal = sqrt(f3_dot(d1, d1));
        const double dist_to_obs = sqrt(f3_dot(d2, d2));
        if (obs_dist_to_goal < L.goal_threshold) continue;
        const float* t = L.prev_traj + (size_t)j * kTrajFloats;
        const F3 first_end{t[15], t[16], t[17]}, last_end{t[87], t[88], t[89]};          // [0][n], [M-1][n]
        if (dist_to_goal > L.goal_threshold && f3_dot(f3_sub(last_end, first_end), f3_sub(first_end, pos)) > 0.0) continue;
        if (dist_to_goal < L.goal_threshold || obs_dist_to_goal < dist_to_goal)
            if (dist_to_obs < best) { best = dist_to_obs; best_j = j; }     // ascending j per thread: first minimum kept
    }
    // block argmin, ties to the smallest index (the reference scans obstacles in id order with a strict <)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (oj >= 0 && (best_j < 0 || ob < best || (ob == best && oj < best_j))) { best = ob; best_j = oj; }
    }
    if ((tid & 31) == 0) { s_dist[tid >> 5] = best; s_idx[tid >> 5] = best_j; }
    __syncthreads();
    if (tid != 0) return;
    for (int w = 1; w < 4; w++)
        if (s_idx[w] >= 0 && (best_j < 0 || s_dist[w] < best || (s_dist[w] == best && s_idx[w] < best_j))) { best = s_dist[w]; best_j = s_idx[w]; }
    F3 goal;
    int kind = 0;
    if (best_j >= 0 && best < L.priority_dist_threshold) {
        const lscgpu_agent_in& o = L.in[best_j];
        const F3 op{o.position[0], o.position[1], o.position[2]};
        const double dist_keep = L.priority_dist_threshold + 0.1;
        goal = f3_sub(pos, f3_scale(f3_normalized(f3_sub(op, pos)), (float)dist_keep));
        kind = 1;
    } else {
        const float* p = L.pred + (size_t)a * kTrajFloats;
        const F3 init_end{p[87], p[88], p[89]};
        const F3 delta = f3_sub(desired, init_end);
        goal = desired;
        if (sqrt(f3_dot(delta, delta)) > L.goal_radius) goal = f3_add(init_end, f3_scale(f3_normalized(delta), (float)L.goal_radius));
    }
    L.goal3[(size_t)a * 3] = (double)goal.x; L.goal3[(size_t)a * 3 + 1] = (double)goal.y; L.goal3[(size_t)a * 3 + 2] = (double)goal.z;
    L.ts[a] = terminal_segments_of(goal, pos, L.consts[a].v_nom, L.dt);
    L.goal_kind[a] = kind;
}
void launch_goal_plan(const GoalLaunch& L, cudaStream_t s) { k_goal_plan<<<L.n_agents, 128, 0, s>>>(L); }

// ------------------------------------------------------------------------------------------------------------
// Debug / parity: CollisionConstraints::getLSC layout for one agent.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_lsc_capture(int n_agents, int a, const float* pred, const AgentConstDev* consts, float* normals,
                              double* d) {
    const int jj = blockIdx.x * blockDim.x + threadIdx.x;
    if (jj >= n_agents - 1) return;
    const int j = jj < a ? jj : jj + 1;
    const AgentConstDev ca = consts[a], cj = consts[j];
    const double downwash = (ca.downwash * ca.radius + cj.downwash * cj.radius) / (ca.radius + cj.radius);
    for (int m = 0; m < kM; m++) {
        F3 ow[6], ob[6];
        for (int i = 0; i < 6; i++) {
            const float* po = pred + (size_t)a * kTrajFloats + (m * 6 + i) * 3;
            const float* pj = pred + (size_t)j * kTrajFloats + (m * 6 + i) * 3;
            ow[i] = F3{po[0], po[1], po[2]};
            ob[i] = F3{pj[0], pj[1], pj[2]};
        }
        LscSegment seg;
        lsc_segment(ow, ob, downwash, cj.radius + ca.radius, seg);
        float* no = normals + ((size_t)jj * kM + m) * 3;
        no[0] = seg.normal.x; no[1] = seg.normal.y; no[2] = seg.normal.z;
        for (int i = 0; i < 6; i++) d[((size_t)jj * kM + m) * 6 + i] = seg.d[i];
    }
}
void launch_lsc_capture(int n_agents, int agent, const float* pred, const AgentConstDev* consts, float* normals,
                        double* d, cudaStream_t s) {
    if (n_agents < 2) return;
    k_lsc_capture<<<(n_agents - 1 + 63) / 64, 64, 0, s>>>(n_agents, agent, pred, consts, normals, d);
}

__global__ void k_gjk_batch(int n, const double* hulls, double* v_out, int* iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    D3 P[6];
    for (int i = 0; i < 6; i++) P[i] = D3{hulls[(size_t)t * 18 + 3 * i], hulls[(size_t)t * 18 + 3 * i + 1], hulls[(size_t)t * 18 + 3 * i + 2]};
    D3 v;
    const int it = gjk_origin_hull6(P, v);
    v_out[(size_t)t * 3] = v.x; v_out[(size_t)t * 3 + 1] = v.y; v_out[(size_t)t * 3 + 2] = v.z;
    iters[t] = it;
}
void launch_gjk_batch(int n, const double* hulls, double* v, int* iters, cudaStream_t s) {
    if (n <= 0) return;
    k_gjk_batch<<<(n + 127) / 128, 128, 0, s>>>(n, hulls, v, iters);
}

// ------------------------------------------------------------------------------------------------------------
// Row store from the reference's LSC container (operator-level QP entry, lscgpu_qp_solve_batch).
// Pair layout inside problem b with n_b obstacles: p = m * n_b + o, stored at pair offset 5 * obs_offset[b].
// ------------------------------------------------------------------------------------------------------------
__global__ void k_rows_from_lsc(int n_problems, const int* obs_offset, int total_obs, const float* lsc_normal,
                                const float* lsc_point, const double* lsc_d, RowRec* rows, int* kept,
                                int* kept_count, double* safe, const unsigned char* obs_slack) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;     // global (obstacle, segment)
    if (t >= total_obs * kM) return;
    const int o = t / kM, m = t % kM;
    // find the problem of obstacle o (few problems: linear scan is fine for an operator-level entry)
    int b = 0;
    while (b + 1 < n_problems && obs_offset[b + 1] <= o) b++;
    const int n_b = obs_offset[b + 1] - obs_offset[b];
    const int p = kM * obs_offset[b] + m * n_b + (o - obs_offset[b]);
    // every pair is priced (no culling at the operator level); obstacles of obs_slack_indices: slack code 31 in the upper
    // bits (a slack pair without a coordinate yet, qp_core.cuh)
    kept[p] = (p - kM * obs_offset[b]) | ((obs_slack && obs_slack[o]) ? (31 << 24) : 0);
    safe[p] = -INFINITY;                              // ... starting with the first iteration
    if (o == obs_offset[b] && m == 0) kept_count[b] = kM * n_b;
    const float* nv = lsc_normal + ((size_t)o * kM + m) * 3;
    const double ax = (double)nv[0], ay = (double)nv[1], az = (double)nv[2];
    const double an = sqrt(ax * ax + ay * ay + az * az);
    RowRec& rec = rows[p];
    rec.ax = nv[0]; rec.ay = nv[1]; rec.az = nv[2]; rec.inv_an = an > 0.0 ? (float)(1.0 / an) : INFINITY;
    for (int i = 0; i < 6; i++) {
        const float* pt = lsc_point + (((size_t)o * kM + m) * 6 + i) * 3;
        rec.rhs[i] = lsc_d[((size_t)o * kM + m) * 6 + i] +
            (__dmul_rn(ax, (double)pt[0]) + __dmul_rn(ay, (double)pt[1]) + __dmul_rn(az, (double)pt[2]));
    }
}
void launch_rows_from_lsc(int n_problems, const int* obs_offset, int total_obs, const float* lsc_normal,
                          const float* lsc_point, const double* lsc_d, RowRec* rows, int* kept,
                          int* kept_count, double* safe, cudaStream_t s, const unsigned char* obs_slack) {
    if (total_obs <= 0) return;
    const int n = total_obs * kM;
    k_rows_from_lsc<<<(n + 127) / 128, 128, 0, s>>>(n_problems, obs_offset, total_obs, lsc_normal, lsc_point, lsc_d, rows,
                                                    kept, kept_count, safe, obs_slack);
}

// ------------------------------------------------------------------------------------------------------------
// k_commit — after the (all-gathered) records are complete: one block per slot of the gather buffer. The record of
// agent `agent_id` goes to res[agent_id] (agent order: what the host reads and what the next step's LPT order and
// failure path look at), traj_curr <- new trajectory, the advanced state becomes the input of a device-resident next
// step (MultiSyncSimulator::update(), src/multi_sync_simulator.cpp:203), and the agent's SFC window takes the step's new
// box (generateFeasibleSFC, src/traj_planner.cpp:1451-1491) — on every replica, so any rank can plan the agent next.
// Slots with agent_id < 0 are empty (ranks that own fewer agents).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(128) k_commit(int n_slots, const GatherSlot* gather, lscgpu_agent_out* res,
                                                unsigned short* act_prev, float* prev_traj, lscgpu_agent_in* in,
                                                double* last_cost, float* boxes, int* init_sfc, int* epoch, int* kept_step,
                                                volatile int* kept_host, PeerExchange px, int* done_count,
                                                volatile int* err_host) {
    const int slot = blockIdx.x;
    const int e = threadIdx.x;
    if (slot == 0 && e == 0) {
        if (kept_step) {            // kept pairs of the step just planned: the host picks the next steps' block size from it
            if (kept_host) { *kept_host = *kept_step; __threadfence_system(); }
            *kept_step = 0;
        }
    }
    const int ep = *epoch;          // stable while the kernel runs: the last block to finish bumps it
    if (px.peers) {
        // direct exchange: this step's slots are in the parity of its epoch; wait for the slot's source rank
        gather += (size_t)(ep & 1) * px.slots;
        if (e == 0) {
            const int src = slot / px.block;
            const int expected = (ep - px.base_epoch + 1) * px.planned_by(src);
            volatile const int* cnt = px.counters(px.rank) + src;
            const unsigned long long t0 = global_ns();
            while (*cnt < expected) {
                if (global_ns() - t0 > 2000000000ull) { *err_host = 1; __threadfence_system(); break; }    // never hang the device
                __nanosleep(100);
            }
            __threadfence_system();
        }
        __syncthreads();
    }
    const lscgpu_agent_out& o = gather[slot].rec;
    const int a = o.agent_id;
    if (a >= 0) {
    if (e >= 64 && e < 64 + kActSlots) act_prev[(size_t)a * kActSlots + (e - 64)] = gather[slot].act[e - 64];
    // the record, 16 bytes per thread
    constexpr int kVec = sizeof(lscgpu_agent_out) / 16;
    static_assert(sizeof(lscgpu_agent_out) % 16 == 0, "record must be a multiple of 16 bytes");
    if (e < kVec) reinterpret_cast<uint4*>(res + a)[e] = reinterpret_cast<const uint4*>(&o)[e];
    if (e < kTrajFloats) prev_traj[(size_t)a * kTrajFloats + e] = (&o.traj[0][0][0])[e];
    if (e < 3) {
        in[a].position[e] = o.next_position[e];
        in[a].velocity[e] = o.next_velocity[e];
        in[a].acceleration[e] = o.next_acceleration[e];
    }
    if (e == 0) last_cost[a] = o.qp_cost;
    if (boxes) {
        float* bx = boxes + (size_t)a * 30;
        const bool first = init_sfc[a] != 0, ok = !(o.flags & LSCGPU_FLAG_SFC_SEED_BLOCKED);
        float v = 0.f;
        if (e >= 96 && e < 126) v = sfc_window_elem(bx, first, ok, o.sfc_box, e - 96);
        __syncthreads();
        if (e >= 96 && e < 126) bx[e - 96] = v;
        if (e == 127) init_sfc[a] = 0;
    }
    }
    // the last block to finish closes the step: epoch + 1 (k_sfc_step / k_agent_plan of the next step stamp with it)
    __syncthreads();
    if (e == 0) {
        __threadfence();
        if (atomicAdd(done_count, 1) == (int)gridDim.x - 1) { *done_count = 0; *epoch = ep + 1; }
    }
}
void launch_commit(int n_slots, const GatherSlot* gather, lscgpu_agent_out* res, unsigned short* act_prev, float* prev_traj,
                   lscgpu_agent_in* in, double* last_cost, float* boxes, int* init_sfc, int* epoch, int* kept_step,
                   volatile int* kept_host, PeerExchange px, int* done_count, volatile int* err_host, cudaStream_t s) {
    if (n_slots > 0)
        k_commit<<<n_slots, 128, 0, s>>>(n_slots, gather, res, act_prev, prev_traj, in, last_cost, boxes, init_sfc, epoch, kept_step,
                                         kept_host, px, done_count, err_host);
}

// ------------------------------------------------------------------------------------------------------------
// Safety audit of the planned step (MultiSyncSimulator::savePlanningResult, src/multi_sync_simulator.cpp:446-475).
// k_audit_positions: getFutureStateMsg(t).position of every agent at every recorded sub-time (Bernstein sum in double,
// include/polynomial.hpp:22-45). k_audit_pairs: one block per agent, threads over the other agents, block argmin.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_audit_positions(int n_agents, const float* traj, double dt, int n_samples, double record_time_step,
                                  float* sample_pos) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_agents * n_samples) return;
    const int s = t / n_agents, a = t % n_agents;
    double ft = 0.0;
    for (int k = 0; k < s; k++) ft += record_time_step;        // the reference accumulates future_time the same way
    int m = (int)(ft / dt);
    if (m == kM && ft < kM * dt + 1e-9) m = kM - 1;
    if (m >= kM) m = kM - 1;
    const double tn = ft / dt - m;
    const double binom[6] = {1.0, 5.0, 10.0, 10.0, 5.0, 1.0};
    const float* c = traj + (size_t)a * kTrajFloats + m * 18;
    double x = 0.0, y = 0.0, z = 0.0;
    for (int i = 0; i < 6; i++) {
        const double b = __dmul_rn(__dmul_rn(binom[i], pow(tn, (double)i)), pow(1.0 - tn, (double)(5 - i)));
        x = __dadd_rn(x, __dmul_rn((double)c[3 * i], b));
        y = __dadd_rn(y, __dmul_rn((double)c[3 * i + 1], b));
        z = __dadd_rn(z, __dmul_rn((double)c[3 * i + 2], b));
    }
    float* o = sample_pos + ((size_t)s * n_agents + a) * 3;
    o[0] = (float)x; o[1] = (float)y; o[2] = (float)z;
}

__global__ void __launch_bounds__(128) k_audit_pairs(int n_agents, const float* sample_pos, const AgentConstDev* consts,
                                                     int n_samples, double* ratio, int* closest) {
    __shared__ double s_r[4];
    __shared__ int s_j[4], s_s[4];
    const int a = blockIdx.x, tid = threadIdx.x;
    const AgentConstDev ca = consts[a];
    double best = 1e9;
    int best_j = -1, best_s = 0;
    for (int s = 0; s < n_samples; s++) {
        const float* pa = sample_pos + ((size_t)s * n_agents + a) * 3;
        const F3 p{pa[0], pa[1], pa[2]};
        for (int j = tid; j < n_agents; j += 128) {
            if (j == a) continue;
            const AgentConstDev cj = consts[j];
            const double dw = (ca.downwash * ca.radius + cj.downwash * cj.radius) / (ca.radius + cj.radius);
            const float* pj = sample_pos + ((size_t)s * n_agents + j) * 3;
            F3 d = f3_sub(p, F3{pj[0], pj[1], pj[2]});
            d.z = (float)__ddiv_rn((double)d.z, dw);
            const double r = __ddiv_rn(sqrt(f3_dot(d, d)), ca.radius + cj.radius);
            if (r < best) { best = r; best_j = j; best_s = s; }       // (s, j) ascending per thread: first minimum kept
        }
    }
    // block argmin; ties to the earliest sub-time, then the smallest id (the reference's scan order with a strict <)
    auto better = [](double r1, int s1, int j1, double r0, int s0, int j0) {
        return j1 >= 0 && (j0 < 0 || r1 < r0 || (r1 == r0 && (s1 < s0 || (s1 == s0 && j1 < j0))));
    };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double orr = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o), os = __shfl_xor_sync(0xffffffffu, best_s, o);
        if (better(orr, os, oj, best, best_s, best_j)) { best = orr; best_j = oj; best_s = os; }
    }
    if ((tid & 31) == 0) { s_r[tid >> 5] = best; s_j[tid >> 5] = best_j; s_s[tid >> 5] = best_s; }
    __syncthreads();
    if (tid != 0) return;
    for (int w = 1; w < 4; w++)
        if (better(s_r[w], s_s[w], s_j[w], best, best_s, best_j)) { best = s_r[w]; best_j = s_j[w]; best_s = s_s[w]; }
    ratio[a] = best; closest[a] = best_j;
}
void launch_safety_audit(int n_agents, const float* traj, const AgentConstDev* consts, double dt, int n_samples,
                         double record_time_step, float* sample_pos, double* ratio, int* closest, cudaStream_t s) {
    if (n_agents <= 0 || n_samples <= 0) return;
    const int n = n_agents * n_samples;
    k_audit_positions<<<(n + 127) / 128, 128, 0, s>>>(n_agents, traj, dt, n_samples, record_time_step, sample_pos);
    k_audit_pairs<<<n_agents, 128, 0, s>>>(n_agents, sample_pos, consts, n_samples, ratio, closest);
}

// ------------------------------------------------------------------------------------------------------------
// getTerminalSegments (src/traj_optimizer.cpp:541-548) for the operator-level QP entry: float3 norm of
// goal - position, double division by the nominal velocity.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_terminal_segments(int n, const double* state9, const double* goal3, const int* agent_index,
                                    const AgentConstDev* consts, double dt, int* ts_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const F3 pos{(float)state9[(size_t)b * 9], (float)state9[(size_t)b * 9 + 1], (float)state9[(size_t)b * 9 + 2]};
    const F3 g{(float)goal3[(size_t)b * 3], (float)goal3[(size_t)b * 3 + 1], (float)goal3[(size_t)b * 3 + 2]};
    const F3 gd = f3_sub(g, pos);
    const double ideal = __ddiv_rn(sqrt(f3_dot(gd, gd)), consts[agent_index[b]].v_nom);
    const double q = __ddiv_rn(__dadd_rn(__dsub_rn(__dmul_rn((double)kM, dt), ideal), 1e-9), dt);
    int ts = (int)q;
    if (ts < 1) ts = 1;
    if (ts > kM) ts = kM;
    ts_out[b] = ts;
}
void launch_terminal_segments(int n, const double* state9, const double* goal3, const int* agent_index,
                              const AgentConstDev* consts, double dt, int* ts_out, cudaStream_t s) {
    if (n <= 0) return;
    k_terminal_segments<<<(n + 127) / 128, 128, 0, s>>>(n, state9, goal3, agent_index, consts, dt, ts_out);
}

}  // namespace lscgpu
