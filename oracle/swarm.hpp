// ORACLE — test infrastructure only (see oracle/README.md).
// Synchronous replanning step for a whole swarm, restating the loop glue of the reference
// (citations relative to /root/reference):
//   src/multi_sync_simulator.cpp:190-337   update() snapshot + sequential plan() over agents
//   src/traj_planner.cpp:99-145,344-425    plan / planImpl / planLSC
//   src/traj_planner.cpp:829-864,699-712   obstaclePredictionWithPrevSol / ...WithCurrVel
//   src/traj_planner.cpp:997-1016,1030-1037 initialTrajPlanningPrevSol / ...CurrVel
//   src/traj_planner.cpp:866-878,1047-1061 prediction / initial-trajectory checks: collapse to the current position,
//                                          sticky obs_slack_indices, flag_initialize_sfc re-armed
//   src/traj_optimizer.cpp:317-326,383-390,455-457  slack variables of the obstacles in obs_slack_indices (qp.hpp)
//   src/traj_planner.cpp:1225-1252         generateCollisionConstraints
//   src/traj_planner.cpp:1442-1491         generateFeasibleSFC (window shift + one new box)
//   src/traj_planner.cpp:1548-1585         trajOptimization (failure keeps the optimizer's last
//                                          successful trajectory and still reports SUCCESS)
//   include/polynomial.hpp:63-121          getStateFromControlPoints at t = dt
// Goal planning (src/traj_planner.cpp:477-608) is the step before the path: with goal_mode == 0 current_goal_position is
// an input (goalPlanningWithStaticGoal), with goal_mode == 1 it is computed per agent by goal.hpp
// (goalPlanningWithPriority, the reference's default) from the desired goals.
#pragma once
#include <thread>
#include <vector>

#include "edt.hpp"
#include "geom.hpp"
#include "qp.hpp"

namespace orc {

struct SwarmParams {
    double dt = 0.2, w = 0.01, wT = 1.0, res = 0.1, reset_threshold = 0.15;
    double slack_w = 1.0;                        // opt/slack_collision_weight (src/param.cpp:75)
    int use_octomap = 0;
    float world_min[3] = {-5, -5, 0}, world_max[3] = {5, 5, 2.5f};
};

struct AgentConst { double radius, downwash, vmax[3], amax[3], v_nom; };

}  // namespace orc
#include "goal.hpp"      // needs AgentConst
namespace orc {

enum StepFlags { FLAG_SLACK_NEEDED = 1, FLAG_SFC_SEED_BLOCKED = 2 };

// Receding horizon: segment m of the previous plan is segment m - 1 of this one. Canonical row ids (qp.hpp) of bounds and
// dynamic limits shifted accordingly; -1 when the row has no counterpart (segment 0 before, an initial-state row now, or
// an LSC row, whose slot numbering changes from step to step).
inline int row_segment(int id) {
    if (id < 180) return ((id >> 1) / 6) % 5;
    if (id < 450) return (((id - 180) >> 1) / 9) % 5;
    return -1;
}
inline int shifted_row_id(int id) {
    if (id < 180) {
        const int v = id >> 1, i = v % 6, m = (v / 6) % 5;
        if (m == 0 || (m == 1 && i < 3)) return -1;
        return id - 12;
    }
    if (id < 450) {
        const int v = (id - 180) >> 1, j = v % 9, m = (v / 9) % 5;
        if (m == 0 || (m == 1 && (j < 2 || j == 5))) return -1;
        return id - 18;
    }
    return -1;
}

struct Swarm {
    SwarmParams prm;
    int N = 0;
    std::vector<AgentConst> ac;
    QpTables T;
    const DistMap* dm = nullptr;
    int seq = 0;                                 // planner_seq (lock-step for all agents)
    std::vector<F3> traj;                        // [N][30] traj_curr
    std::vector<F3> pos, vel, acc;               // current state (input of the step)
    std::vector<F3> goal;                        // current_goal_position (input of the step, or output of goal planning)
    std::vector<F3> desired;                     // desired_goal_position (goal_mode == 1)
    int goal_mode = 0;                           // 0 static, 1 prior_based
    GoalParams gp;
    std::vector<int> goal_kind;                  // per agent: 0 A* + line of sight, 1 retreat from a higher-priority agent
    long long astar_expansions = 0;
    std::vector<F3> box_min, box_max;            // [N][5] persistent SFC
    std::vector<int> init_sfc;                   // flag_initialize_sfc
    // Disturbance handling. Every planner runs the same two checks on the same data (the shifted previous trajectory of
    // agent j against j's observed position), so obs_slack_indices of planner i is: every obstacle once i itself was
    // reset (:1049-1051), else the agents j that were ever reset (:869-870). Sticky: the reference never erases it.
    // warm start of the QP (experiment mirror of the kernel's, qp.hpp): previous active rows per agent
    int warm_start = 0;
    std::vector<int> prev_act; std::vector<int> prev_n_act;     // [N][39], [N]; -1: no usable previous solve
    long long warm_accepted = 0, warm_tried = 0;
    std::vector<char> reset_ever;
    std::vector<double> qp_slack_cost; std::vector<int> qp_slack_rows;   // step outputs: slack share of the cost, rows with eps < 0
    // step outputs
    std::vector<double> qp_cost; std::vector<int> qp_status, qp_iters, qp_active, flags;
    std::vector<double> qp_maxviol, qp_kkt;
    std::vector<F3> pred;                        // [N][30] initial_traj == prediction seen by others
    // optional capture of all constraints of the step (tests): [N][N][5]
    bool capture = false;
    std::vector<F3> cap_normal; std::vector<double> cap_d; std::vector<int> cap_gjk;
    long long counters[4] = {0, 0, 0, 0};        // gjk iterations, qp iterations, edt lookups, qp rows

    void init(const SwarmParams& p, int n_agents, const AgentConst* consts) {
        prm = p; N = n_agents; ac.assign(consts, consts + N);
        build_tables(p.dt, p.w, p.wT, T);
        traj.assign((size_t)N * 30, f3(0, 0, 0));                   // traj_planner.cpp:36-39
        pos.assign(N, f3(0, 0, 0)); vel = pos; acc = pos; goal = pos; desired = pos; goal_kind.assign(N, 0);
        box_min.assign((size_t)N * 5, f3(0, 0, 0)); box_max = box_min;
        init_sfc.assign(N, 1);                                      // traj_planner.cpp:49
        reset_ever.assign(N, 0); qp_slack_cost.assign(N, 0); qp_slack_rows.assign(N, 0);
        prev_act.assign((size_t)N * QRED, 0); prev_n_act.assign(N, -1);
        qp_cost.assign(N, 0); qp_status.assign(N, 0); qp_iters.assign(N, 0); qp_active.assign(N, 0);
        qp_maxviol.assign(N, 0); qp_kkt.assign(N, 0); flags.assign(N, 0);
        pred.assign((size_t)N * 30, f3(0, 0, 0));
        seq = 0;
    }

    void predict_all() {
        for (int a = 0; a < N; a++) {
            F3* o = &pred[(size_t)a * 30];
            if (seq < 2) {                                          // :830, :998 (seq already incremented)
                for (int m = 0; m < 5; m++)
                    for (int i = 0; i < 6; i++) {
                        double m_intp = m + (double)i / 5;
                        o[m * 6 + i] = pos[a] + (vel[a] * (float)m_intp) * (float)prm.dt;   // :707-708
                    }
            } else {
                const F3* t = &traj[(size_t)a * 30];
                for (int m = 0; m < 4; m++) for (int i = 0; i < 6; i++) o[m * 6 + i] = t[(m + 1) * 6 + i];
                for (int i = 0; i < 6; i++) o[24 + i] = t[29];                               // :851-855
            }
            if (normf(o[0] - pos[a]) > prm.reset_threshold) {                                // :869, :1048
                flags[a] |= FLAG_SLACK_NEEDED;
                reset_ever[a] = 1;                                                           // :870, :1049-1051
                for (int e = 0; e < 30; e++) o[e] = pos[a];                                  // :871-875, :1053-1057
                init_sfc[a] = 1;                                                             // :1059
            }
        }
    }

    void plan_agent(int a, std::vector<F3>& new_traj, std::vector<LscRows>& rows_buf) {
        const F3* own = &pred[(size_t)a * 30];
        rows_buf.clear();
        long long gjk_it = 0;
        for (int j = 0; j < N; j++) {
            if (j == a) continue;
            LscPair lp;
            const F3* obs = &pred[(size_t)j * 30];
            lsc_pair(own, obs, 5, ac[a].radius, ac[a].downwash, ac[j].radius, ac[j].downwash, lp);
            for (int m = 0; m < 5; m++) {
                LscRows r; r.m = m;
                r.slack = (reset_ever[a] || reset_ever[j]) ? 1 : 0;                          // obs_slack_indices of planner a
                r.a[0] = (double)lp.normal[m].x; r.a[1] = (double)lp.normal[m].y; r.a[2] = (double)lp.normal[m].z;
                for (int i = 0; i < 6; i++) {
                    const F3 o = obs[m * 6 + i];
                    r.rhs[i] = lp.d[m][i] + ((r.a[0] * (double)o.x + r.a[1] * (double)o.y) + r.a[2] * (double)o.z);
                }
                rows_buf.push_back(r);
                gjk_it += lp.gjk_iters[m];
                if (capture) {
                    size_t c = ((size_t)a * N + j) * 5 + m;
                    cap_normal[c] = lp.normal[m]; cap_gjk[c] = lp.gjk_iters[m];
                    for (int i = 0; i < 6; i++) cap_d[c * 6 + i] = lp.d[m][i];
                }
            }
        }
        // SFC (traj_planner.cpp:1451-1491)
        const long long lk0 = tl_edt_lookups;
        if (prm.use_octomap) {
            Corridor cc{dm, f3(prm.world_min[0], prm.world_min[1], prm.world_min[2]),
                        f3(prm.world_max[0], prm.world_max[1], prm.world_max[2]), prm.res};
            F3 bmin, bmax;
            if (init_sfc[a]) {
                if (cc.expand_from_point(pos[a], goal[a], ac[a].radius, bmin, bmax)) {
                    for (int m = 0; m < 5; m++) { box_min[a * 5 + m] = bmin; box_max[a * 5 + m] = bmax; }
                } else flags[a] |= FLAG_SFC_SEED_BLOCKED;
                init_sfc[a] = 0;
            } else {
                for (int m = 1; m < 5; m++) { box_min[a * 5 + m - 1] = box_min[a * 5 + m]; box_max[a * 5 + m - 1] = box_max[a * 5 + m]; }
                if (cc.expand_from_point(traj[(size_t)a * 30 + 29], goal[a], ac[a].radius, bmin, bmax)) {
                    box_min[a * 5 + 4] = bmin; box_max[a * 5 + 4] = bmax;
                } else flags[a] |= FLAG_SFC_SEED_BLOCKED;
            }
        }
        // QP
        QpProblem qp;
        for (int k = 0; k < 3; k++) { qp.s[0][k] = (double)pos[a](k); qp.s[1][k] = (double)vel[a](k); qp.s[2][k] = (double)acc[a](k); qp.goal[k] = (double)goal[a](k); }
        qp.ts = terminal_segments(pos[a], goal[a], ac[a].v_nom, prm.dt);
        if (qp.ts > 5) qp.ts = 5;   // reference throws (traj_optimizer.cpp:356-358); cannot occur for ideal >= 0
        for (int k = 0; k < 3; k++)
            for (int m = 0; m < 5; m++)
                for (int i = 0; i < 6; i++) {
                    int v = k * 30 + m * 6 + i;
                    if (m == 0 && i < 3) { qp.lb[v] = -INFINITY; qp.ub[v] = INFINITY; continue; }
                    double lo = (double)prm.world_min[k], hi = (double)prm.world_max[k];
                    if (prm.use_octomap) {
                        lo = std::max(lo, (double)box_min[a * 5 + m](k));
                        hi = std::min(hi, (double)box_max[a * 5 + m](k));
                    }
                    qp.lb[v] = lo; qp.ub[v] = hi;
                }
        for (int k = 0; k < 3; k++) { qp.vmax[k] = ac[a].vmax[k]; qp.amax[k] = ac[a].amax[k]; }
        qp.rows = rows_buf.data(); qp.n_rows = (int)rows_buf.size();
        QpResult res;
        bool any_slack = false;
        for (const LscRows& r : rows_buf) any_slack |= r.slack != 0;
        qp.slack_w = prm.slack_w;
        if (any_slack) qp_solve_slack(T, qp, res);
        else if (warm_start && prev_n_act[a] > 0 && !(flags[a] & FLAG_SLACK_NEEDED)) {
            int guess[2 * QRED], ng = 0;
            for (int k = 0; k < prev_n_act[a]; k++) {
                const int id = prev_act[(size_t)a * QRED + k];
                if (warm_start & 1) { if (id < 450) guess[ng++] = id; }       // same rows: the plan keeps its shape relative to the horizon
                if ((warm_start & 4) && id >= 450) guess[ng++] = id;           // ... and the same LSC rows (neighbour, segment, point)
                if (warm_start & 2) {
                    const int sh = shifted_row_id(id);
                    if (sh >= 0) guess[ng++] = sh;
                    if (!(warm_start & 1) && row_segment(id) == QM - 1) guess[ng++] = id;
                }
            }
            qp_solve(T, qp, res, 2000, guess, ng);
            __atomic_fetch_add(&warm_tried, 1LL, __ATOMIC_RELAXED);
            if (res.warm_accepted) __atomic_fetch_add(&warm_accepted, 1LL, __ATOMIC_RELAXED);
        } else qp_solve(T, qp, res);
        if (!any_slack && res.status == QP_OK) {
            prev_n_act[a] = res.n_act_ids;
            for (int k = 0; k < res.n_act_ids; k++) prev_act[(size_t)a * QRED + k] = res.act_ids[k];
        } else prev_n_act[a] = -1;
        qp_slack_cost[a] = any_slack ? res.slack_cost : 0.0;
        qp_slack_rows[a] = 0;
        for (double e : res.eps) qp_slack_rows[a] += e < 0.0;
        qp_status[a] = res.status; qp_iters[a] = res.iters; qp_active[a] = res.n_active;
        qp_maxviol[a] = res.max_violation; qp_kkt[a] = res.kkt_stationarity;
        if (res.status == QP_OK) {
            qp_cost[a] = res.cost;
            for (int m = 0; m < 5; m++)
                for (int i = 0; i < 6; i++)
                    new_traj[(size_t)a * 30 + m * 6 + i] =
                        f3((float)res.x[m * 6 + i], (float)res.x[30 + m * 6 + i], (float)res.x[60 + m * 6 + i]);
        }
        __atomic_fetch_add(&counters[0], gjk_it, __ATOMIC_RELAXED);
        __atomic_fetch_add(&counters[1], (long long)res.iters, __ATOMIC_RELAXED);
        __atomic_fetch_add(&counters[2], tl_edt_lookups - lk0, __ATOMIC_RELAXED);
        __atomic_fetch_add(&counters[3], (long long)(450 + 6 * rows_buf.size()), __ATOMIC_RELAXED);
    }

    // goalPlanningWithPriority for agents [a0, a1): every obstacle is another agent of the swarm, seen at its current
    // state with its desired goal and its previous trajectory (src/multi_sync_simulator.cpp:269-299)
    void plan_goals(int a0, int a1, int threads) {
        gp.world_resolution = prm.res;
        const F3 wmin = f3(prm.world_min[0], prm.world_min[1], prm.world_min[2]);
        const F3 wmax = f3(prm.world_max[0], prm.world_max[1], prm.world_max[2]);
        const DistMap* map = prm.use_octomap ? dm : nullptr;
        auto work = [&](int a) {
            std::vector<char> in_set;
            bool any = false;
            for (int j = 0; j < N; j++) any |= reset_ever[j] != 0;
            if (any) { in_set.resize(N); for (int j = 0; j < N; j++) in_set[j] = reset_ever[a] || reset_ever[j]; }
            GoalResult r = goal_planning_priority(a, N, pos.data(), desired.data(), traj.data(), pred[(size_t)a * 30 + 29],
                                                  ac.data(), map, gp, wmin, wmax, any ? in_set.data() : nullptr);
            goal[a] = r.goal; goal_kind[a] = r.mode;
            __atomic_fetch_add(&astar_expansions, r.expansions, __ATOMIC_RELAXED);
        };
        if (threads <= 1) { for (int a = a0; a < a1; a++) work(a); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back([&, t]() { for (int a = a0 + t; a < a1; a += threads) work(a); });
        for (auto& t : th) t.join();
    }

    // one synchronous replanning step for agents [a0, a1)
    void step(int a0, int a1, int threads) {
        std::vector<int> ids;
        for (int a = a0; a < a1; a++) ids.push_back(a);
        step_list(ids.data(), (int)ids.size(), threads, a0, a1);
    }
    // ... for an arbitrary set of agents (a multi-GPU rank's share of the swarm): all of them plan from the same snapshot
    void step_list(const int* ids, int n_ids, int threads, int goal_a0 = -1, int goal_a1 = -1) {
        seq++;                                                         // traj_planner.cpp:127
        std::fill(flags.begin(), flags.end(), 0);
        if (capture) {
            cap_normal.assign((size_t)N * N * 5, f3(0, 0, 0)); cap_d.assign((size_t)N * N * 30, 0.0); cap_gjk.assign((size_t)N * N * 5, 0);
        }
        predict_all();
        if (goal_mode == 1) {                                           // traj_planner.cpp:360-364 (after the initial trajectory)
            if (goal_a0 >= 0) plan_goals(goal_a0, goal_a1, threads);
            else for (int t = 0; t < n_ids; t++) plan_goals(ids[t], ids[t] + 1, 1);
        }
        std::vector<F3> new_traj = traj;
        if (threads <= 1) {
            std::vector<LscRows> rows;
            for (int t = 0; t < n_ids; t++) plan_agent(ids[t], new_traj, rows);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < threads; t++)
                th.emplace_back([&, t]() {
                    std::vector<LscRows> rows;
                    for (int k = t; k < n_ids; k += threads) plan_agent(ids[k], new_traj, rows);
                });
            for (auto& t : th) t.join();
        }
        traj.swap(new_traj);
    }

    // include/polynomial.hpp:63-121 evaluated at current_time = dt: segment 1, t = 0
    void future_state(int a, F3& p, F3& v, F3& ac_) const {
        const F3* c = &traj[(size_t)a * 30 + 6];
        p = c[0];
        F3 v0 = ((c[1] - c[0]) * 5.0f) * (float)std::pow(prm.dt, -1);
        F3 v1 = ((c[2] - c[1]) * 5.0f) * (float)std::pow(prm.dt, -1);
        v = v0;
        ac_ = ((v1 - v0) * 4.0f) * (float)std::pow(prm.dt, -1);
    }
    // getStateFromControlPoints(...).position at an arbitrary time (include/polynomial.hpp:22-45,63-77)
    F3 future_position(int a, double t) const {
        int m = (int)(t / prm.dt);
        if (m == 5 && t < 5 * prm.dt + 1e-9) m = 4;
        const double tn = t / prm.dt - m;
        static const int binom[6] = {1, 5, 10, 10, 5, 1};
        const F3* c = &traj[(size_t)a * 30 + m * 6];
        double x = 0, y = 0, z = 0;
        for (int i = 0; i < 6; i++) {
            const double b = binom[i] * std::pow(tn, i) * std::pow(1 - tn, 5 - i);
            x += c[i].x * b; y += c[i].y * b; z += c[i].z * b;
        }
        return f3((float)x, (float)y, (float)z);
    }
    // minimum-distance audit of MultiSyncSimulator::savePlanningResult (src/multi_sync_simulator.cpp:446-475): for every
    // recorded sub-time of the step and every agent, the smallest downwash-scaled distance to another agent over the sum of
    // the radii (include/util.hpp:225-229). ratio[a] / closest[a]: minimum over the sub-times (first minimum kept).
    void safety_audit(double record_time_step, double time_step, double* ratio, int* closest) const {
        for (int a = 0; a < N; a++) { ratio[a] = 1e9; closest[a] = -1; }
        std::vector<F3> p(N);
        for (double ft = 0; ft < time_step - 1e-5; ft += record_time_step) {
            for (int a = 0; a < N; a++) p[a] = future_position(a, ft);
            for (int a = 0; a < N; a++)
                for (int j = 0; j < N; j++) {
                    if (j == a) continue;
                    const double dw = (ac[a].downwash * ac[a].radius + ac[j].downwash * ac[j].radius) / (ac[a].radius + ac[j].radius);
                    F3 d = p[a] - p[j];
                    d.z = (float)((double)d.z / dw);
                    const double r = normf(d) / (ac[a].radius + ac[j].radius);
                    if (r < ratio[a]) { ratio[a] = r; closest[a] = j; }
                }
        }
    }
    void advance_states() {
        for (int a = 0; a < N; a++) future_state(a, pos[a], vel[a], acc[a]);
    }
};

}  // namespace orc
