// ORACLE — test infrastructure only (see oracle/README.md). C entry points for ctypes.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load
// this library; the product (lsc_planner_b200/) never does.
#include <cstring>
#include <string>

#include "edt.hpp"
#include "geom.hpp"
#include "qp.hpp"
#include "swarm.hpp"

using namespace orc;

extern "C" {

// ---- GJK / LSC ------------------------------------------------------------------------------
int orc_gjk(const double* pts, int np, double* v_out) {
    D3 P[16];
    if (np > 16) np = 16;
    for (int i = 0; i < np; i++) P[i] = D3{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    D3 v;
    int it = gjk_origin_hull(P, np, v);
    v_out[0] = v.x; v_out[1] = v.y; v_out[2] = v.z;
    return it;
}

void orc_lsc_pair(const float* own, const float* obs, double r_i, double dw_i, double r_j, double dw_j,
                  float* normals, double* d, int* iters) {
    LscPair lp;
    lsc_pair(reinterpret_cast<const F3*>(own), reinterpret_cast<const F3*>(obs), 5, r_i, dw_i, r_j, dw_j, lp);
    for (int m = 0; m < 5; m++) {
        normals[3 * m] = lp.normal[m].x; normals[3 * m + 1] = lp.normal[m].y; normals[3 * m + 2] = lp.normal[m].z;
        for (int i = 0; i < 6; i++) d[6 * m + i] = lp.d[m][i];
        iters[m] = lp.gjk_iters[m];
    }
}

// ---- map / EDT / SFC --------------------------------------------------------------------------
struct MapHandle { BtTree tree; DistMap dm; };

void* orc_map_from_bt(const char* path, const float* wmin, const float* wmax, double res) {
    try {
        MapHandle* h = new MapHandle;
        h->tree = bt_load(path);
        h->dm = distmap_build(h->tree.occupied, res, f3(wmin[0], wmin[1], wmin[2]), f3(wmax[0], wmax[1], wmax[2]));
        return h;
    } catch (const std::exception& e) { std::fprintf(stderr, "orc_map_from_bt: %s\n", e.what()); return nullptr; }
}
void* orc_map_from_voxels(const int* keys, int n, const float* wmin, const float* wmax, double res) {
    MapHandle* h = new MapHandle;
    h->tree.res = res;
    for (int i = 0; i < n; i++) h->tree.occupied.push_back(Key3{{keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]}});
    h->dm = distmap_build(h->tree.occupied, res, f3(wmin[0], wmin[1], wmin[2]), f3(wmax[0], wmax[1], wmax[2]));
    return h;
}
void orc_map_info(void* hv, int* size3, int* off3, long long* n_occ, long long* n_nodes) {
    MapHandle* h = (MapHandle*)hv;
    for (int a = 0; a < 3; a++) { size3[a] = h->dm.size[a]; off3[a] = h->dm.off[a]; }
    *n_occ = (long long)h->tree.occupied.size(); *n_nodes = (long long)h->tree.n_nodes;
}
void orc_map_occupied(void* hv, int* keys) {
    MapHandle* h = (MapHandle*)hv;
    for (size_t i = 0; i < h->tree.occupied.size(); i++) for (int a = 0; a < 3; a++) keys[3 * i + a] = h->tree.occupied[i].k[a];
}
void orc_map_sqdist(void* hv, int* out) {
    MapHandle* h = (MapHandle*)hv;
    std::memcpy(out, h->dm.sqdist.data(), h->dm.sqdist.size() * sizeof(int));
}
float orc_map_distance(void* hv, const float* p) { return ((MapHandle*)hv)->dm.distance(f3(p[0], p[1], p[2])); }
void orc_map_free(void* hv) { delete (MapHandle*)hv; }

int orc_sfc_expand(void* hv, const float* wmin, const float* wmax, double res, const float* point,
                   const float* goal, double radius, float* box6, long long* lookups) {
    MapHandle* h = (MapHandle*)hv;
    Corridor cc{&h->dm, f3(wmin[0], wmin[1], wmin[2]), f3(wmax[0], wmax[1], wmax[2]), res};
    long long l0 = tl_edt_lookups;
    F3 bmin, bmax;
    bool ok = cc.expand_from_point(f3(point[0], point[1], point[2]), f3(goal[0], goal[1], goal[2]), radius, bmin, bmax);
    if (lookups) *lookups = tl_edt_lookups - l0;
    if (!ok) return 0;
    box6[0] = bmin.x; box6[1] = bmin.y; box6[2] = bmin.z; box6[3] = bmax.x; box6[4] = bmax.y; box6[5] = bmax.z;
    return 1;
}

// ---- QP ---------------------------------------------------------------------------------------
void* orc_tables_create(double dt, double w, double wT) {
    QpTables* T = new QpTables;
    build_tables(dt, w, wT, *T);
    return T;
}
void orc_tables_free(void* t) { delete (QpTables*)t; }
void orc_tables_get(void* tv, double* Qb /*36*/, double* A17 /*17*30*/, double* Xp /*30*3*/, double* Z /*30*13*/,
                    double* G /*5*30*13*/, double* Xs /*5*30*3*/, double* xg /*5*30*/) {
    QpTables* T = (QpTables*)tv;
    std::memcpy(Qb, T->Qb, sizeof T->Qb); std::memcpy(A17, T->A17, sizeof T->A17);
    std::memcpy(Xp, T->Xp, sizeof T->Xp); std::memcpy(Z, T->Z, sizeof T->Z);
    std::memcpy(G, T->G, sizeof T->G); std::memcpy(Xs, T->Xs, sizeof T->Xs); std::memcpy(xg, T->xg, sizeof T->xg);
}
int orc_terminal_segments(const float* pos, const float* goal, double v_nom, double dt) {
    return terminal_segments(f3(pos[0], pos[1], pos[2]), f3(goal[0], goal[1], goal[2]), v_nom, dt);
}

static void fill_problem(QpProblem& p, std::vector<LscRows>& rows, const double* state9, const double* goal3, int ts,
                         const double* lb, const double* ub, const double* vmax, const double* amax, int n_rows,
                         const int* row_m, const double* row_a, const double* row_rhs) {
    for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) p.s[r][k] = state9[3 * r + k];
    for (int k = 0; k < 3; k++) { p.goal[k] = goal3[k]; p.vmax[k] = vmax[k]; p.amax[k] = amax[k]; }
    p.ts = ts;
    std::memcpy(p.lb, lb, sizeof p.lb); std::memcpy(p.ub, ub, sizeof p.ub);
    rows.resize(n_rows);
    for (int r = 0; r < n_rows; r++) {
        rows[r].m = row_m[r];
        for (int k = 0; k < 3; k++) rows[r].a[k] = row_a[3 * r + k];
        for (int i = 0; i < 6; i++) rows[r].rhs[i] = row_rhs[6 * r + i];
    }
    p.rows = rows.data(); p.n_rows = n_rows;
}

// state9 = pos xyz, vel xyz, acc xyz. Returns status.
int orc_qp_solve(void* tv, const double* state9, const double* goal3, int ts, const double* lb, const double* ub,
                 const double* vmax, const double* amax, int n_rows, const int* row_m, const double* row_a,
                 const double* row_rhs, double* x90, double* info /*cost,iters,n_active,kkt,maxviol*/) {
    QpProblem p; std::vector<LscRows> rows;
    fill_problem(p, rows, state9, goal3, ts, lb, ub, vmax, amax, n_rows, row_m, row_a, row_rhs);
    QpResult r;
    qp_solve(*(QpTables*)tv, p, r);
    std::memcpy(x90, r.x, sizeof r.x);
    info[0] = r.cost; info[1] = r.iters; info[2] = r.n_active; info[3] = r.kkt_stationarity; info[4] = r.max_violation;
    return r.status;
}

// Dense assembly in the reference's row order. Ain must hold max_in*90 doubles. Returns n_in.
int orc_qp_dense(void* tv, const double* state9, const double* goal3, int ts, const double* lb, const double* ub,
                 const double* vmax, const double* amax, int n_rows, const int* row_m, const double* row_a,
                 const double* row_rhs, const float* boxes, double* P, double* q, double* c0, double* Aeq,
                 double* beq, double* Ain, double* bin, int max_in) {
    QpProblem p; std::vector<LscRows> rows;
    fill_problem(p, rows, state9, goal3, ts, lb, ub, vmax, amax, n_rows, row_m, row_a, row_rhs);
    DenseQp D;
    assemble_dense(*(QpTables*)tv, p, boxes, D);
    std::memcpy(P, D.P.data(), D.P.size() * 8); std::memcpy(q, D.qlin.data(), D.qlin.size() * 8);
    *c0 = D.c0;
    std::memcpy(Aeq, D.Aeq.data(), D.Aeq.size() * 8); std::memcpy(beq, D.beq.data(), D.beq.size() * 8);
    if (D.n_in > max_in) return -D.n_in;
    std::memcpy(Ain, D.Ain.data(), D.Ain.size() * 8); std::memcpy(bin, D.bin.data(), D.bin.size() * 8);
    return D.n_in;
}

void orc_set_tier_threshold(double t) { g_tier_threshold = t; }

// The QP with slack variables (qp_solve_slack): row_slack[r] = 1 marks the (obstacle, segment) entries of obstacles in
// obs_slack_indices. eps_out[n_rows]; info = cost, iters, n_active, kkt, maxviol, slack_cost.
int orc_qp_solve_slack(void* tv, const double* state9, const double* goal3, int ts, const double* lb, const double* ub,
                       const double* vmax, const double* amax, int n_rows, const int* row_m, const double* row_a,
                       const double* row_rhs, const int* row_slack, double slack_w, double* x90, double* eps_out, double* info) {
    QpProblem p; std::vector<LscRows> rows;
    fill_problem(p, rows, state9, goal3, ts, lb, ub, vmax, amax, n_rows, row_m, row_a, row_rhs);
    for (int r = 0; r < n_rows; r++) rows[r].slack = row_slack[r];
    p.slack_w = slack_w;
    QpResult r;
    qp_solve_slack(*(QpTables*)tv, p, r);
    std::memcpy(x90, r.x, sizeof r.x);
    for (int i = 0; i < n_rows; i++) eps_out[i] = r.eps[i];
    info[0] = r.cost; info[1] = r.iters; info[2] = r.n_active; info[3] = r.kkt_stationarity; info[4] = r.max_violation;
    info[5] = r.slack_cost;
    return r.status;
}

// ---- swarm ------------------------------------------------------------------------------------
void* orc_swarm_create(int n_agents, double dt, double w, double wT, double res, double reset_threshold,
                       int use_octomap, const float* wmin, const float* wmax, const double* radius,
                       const double* downwash, const double* vmax /*N*3*/, const double* amax /*N*3*/,
                       const double* v_nom) {
    Swarm* s = new Swarm;
    SwarmParams p;
    p.dt = dt; p.w = w; p.wT = wT; p.res = res; p.reset_threshold = reset_threshold; p.use_octomap = use_octomap;
    for (int k = 0; k < 3; k++) { p.world_min[k] = wmin[k]; p.world_max[k] = wmax[k]; }
    std::vector<AgentConst> c(n_agents);
    for (int a = 0; a < n_agents; a++) {
        c[a].radius = radius[a]; c[a].downwash = downwash[a]; c[a].v_nom = v_nom[a];
        for (int k = 0; k < 3; k++) { c[a].vmax[k] = vmax[3 * a + k]; c[a].amax[k] = amax[3 * a + k]; }
    }
    s->init(p, n_agents, c.data());
    return s;
}
void orc_swarm_free(void* sv) { delete (Swarm*)sv; }
void orc_swarm_set_map(void* sv, void* map) { ((Swarm*)sv)->dm = map ? &((MapHandle*)map)->dm : nullptr; }
void orc_swarm_set_capture(void* sv, int on) { ((Swarm*)sv)->capture = on != 0; }
void orc_swarm_set_slack_weight(void* sv, double w) { ((Swarm*)sv)->prm.slack_w = w; }
// active rows of agent a's last successful plain solve (canonical ids); returns the count or -1
int orc_swarm_get_active(void* sv, int a, int* ids) {
    Swarm* s = (Swarm*)sv;
    const int n = s->prev_n_act[a];
    for (int k = 0; k < n; k++) ids[k] = s->prev_act[(size_t)a * QRED + k];
    return n;
}
void orc_swarm_set_warm_start(void* sv, int on) { ((Swarm*)sv)->warm_start = on; }
void orc_swarm_get_warm_stats(void* sv, long long* tried_accepted) { Swarm* s = (Swarm*)sv; tried_accepted[0] = s->warm_tried; tried_accepted[1] = s->warm_accepted; }
// sticky disturbance state (who was ever reset) and the slack outputs of the last step
void orc_swarm_get_reset_ever(void* sv, unsigned char* out) { Swarm* s = (Swarm*)sv; for (int a = 0; a < s->N; a++) out[a] = (unsigned char)s->reset_ever[a]; }
void orc_swarm_set_reset_ever(void* sv, const unsigned char* in) { Swarm* s = (Swarm*)sv; for (int a = 0; a < s->N; a++) s->reset_ever[a] = (char)in[a]; }
void orc_swarm_get_slack(void* sv, double* slack_cost, int* slack_rows) {
    Swarm* s = (Swarm*)sv;
    for (int a = 0; a < s->N; a++) { slack_cost[a] = s->qp_slack_cost[a]; slack_rows[a] = s->qp_slack_rows[a]; }
}
void orc_swarm_set_state(void* sv, const float* pos, const float* vel, const float* acc) {
    Swarm* s = (Swarm*)sv;
    std::memcpy(s->pos.data(), pos, s->N * 12); std::memcpy(s->vel.data(), vel, s->N * 12); std::memcpy(s->acc.data(), acc, s->N * 12);
}
void orc_swarm_set_goals(void* sv, const float* goal) { Swarm* s = (Swarm*)sv; std::memcpy(s->goal.data(), goal, s->N * 12); }
void orc_swarm_set_traj(void* sv, const float* traj, int seq) {
    Swarm* s = (Swarm*)sv; std::memcpy(s->traj.data(), traj, (size_t)s->N * 360); s->seq = seq;
}
void orc_swarm_set_boxes(void* sv, const float* boxes /*N*5*6*/, const int* init_sfc) {
    Swarm* s = (Swarm*)sv;
    for (int i = 0; i < s->N * 5; i++) {
        s->box_min[i] = f3(boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]);
        s->box_max[i] = f3(boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5]);
    }
    for (int a = 0; a < s->N; a++) s->init_sfc[a] = init_sfc[a];
}
// goal planning: mode 0 static (goal = input), 1 prior_based (goal computed from the desired goals)
void orc_swarm_set_goal_mode(void* sv, int mode, double grid_resolution, double grid_margin, double goal_threshold,
                             double goal_radius, double priority_dist_threshold) {
    Swarm* s = (Swarm*)sv;
    s->goal_mode = mode;
    s->gp.grid_resolution = grid_resolution; s->gp.grid_margin = grid_margin; s->gp.goal_threshold = goal_threshold;
    s->gp.goal_radius = goal_radius; s->gp.priority_dist_threshold = priority_dist_threshold;
}
void orc_swarm_set_desired_goals(void* sv, const float* goal) { Swarm* s = (Swarm*)sv; std::memcpy(s->desired.data(), goal, s->N * 12); }
void orc_swarm_get_goals(void* sv, float* goal, int* kind) {
    Swarm* s = (Swarm*)sv;
    std::memcpy(goal, s->goal.data(), s->N * 12);
    if (kind) std::memcpy(kind, s->goal_kind.data(), s->N * 4);
}
long long orc_swarm_astar_expansions(void* sv) { return ((Swarm*)sv)->astar_expansions; }
// A* on an explicit occupancy grid (pins the restatement against oracle/_ref): returns the path length in cells
int orc_astar(const int* dim, const unsigned char* grid, const int* start, const int* goal, int* path_out, int max_len,
              long long* expansions) {
    std::vector<uint8_t> g(grid, grid + (size_t)dim[0] * dim[1] * dim[2]);
    const auto path = astar_search(g, dim, start, goal, expansions);
    const int n = (int)path.size();
    for (int k = 0; k < n && k < max_len; k++) { path_out[3 * k] = path[k][0]; path_out[3 * k + 1] = path[k][1]; path_out[3 * k + 2] = path[k][2]; }
    return n;
}
// one agent's goal planning on explicit inputs
int orc_goal_plan(int a, int n, const float* pos, const float* desired, const float* prev_traj, const float* init_end,
                  const double* radius, const double* downwash, void* map, double world_res, const float* wmin, const float* wmax,
                  double grid_resolution, double grid_margin, double goal_threshold, double goal_radius,
                  double priority_dist_threshold, float* goal_out, long long* expansions) {
    std::vector<AgentConst> ac(n);
    for (int i = 0; i < n; i++) { ac[i] = AgentConst{}; ac[i].radius = radius[i]; ac[i].downwash = downwash[i]; }
    GoalParams gp; gp.grid_resolution = grid_resolution; gp.grid_margin = grid_margin; gp.goal_threshold = goal_threshold;
    gp.goal_radius = goal_radius; gp.priority_dist_threshold = priority_dist_threshold; gp.world_resolution = world_res;
    const GoalResult r = goal_planning_priority(a, n, reinterpret_cast<const F3*>(pos), reinterpret_cast<const F3*>(desired),
                                                reinterpret_cast<const F3*>(prev_traj), f3(init_end[0], init_end[1], init_end[2]),
                                                ac.data(), map ? &((MapHandle*)map)->dm : nullptr, gp, f3(wmin[0], wmin[1], wmin[2]),
                                                f3(wmax[0], wmax[1], wmax[2]));
    goal_out[0] = r.goal.x; goal_out[1] = r.goal.y; goal_out[2] = r.goal.z;
    if (expansions) *expansions = r.expansions;
    return r.mode;
}

void orc_swarm_safety_audit(void* sv, double record_time_step, double time_step, double* ratio, int* closest) {
    ((Swarm*)sv)->safety_audit(record_time_step, time_step, ratio, closest);
}
void orc_swarm_step(void* sv, int a0, int a1, int threads) { ((Swarm*)sv)->step(a0, a1, threads); }
void orc_swarm_step_list(void* sv, const int* ids, int n, int threads) { ((Swarm*)sv)->step_list(ids, n, threads); }
void orc_swarm_advance(void* sv) { ((Swarm*)sv)->advance_states(); }
int orc_swarm_seq(void* sv) { return ((Swarm*)sv)->seq; }
void orc_swarm_get_traj(void* sv, float* out) { Swarm* s = (Swarm*)sv; std::memcpy(out, s->traj.data(), (size_t)s->N * 360); }
void orc_swarm_get_pred(void* sv, float* out) { Swarm* s = (Swarm*)sv; std::memcpy(out, s->pred.data(), (size_t)s->N * 360); }
void orc_swarm_get_state(void* sv, float* pos, float* vel, float* acc) {
    Swarm* s = (Swarm*)sv;
    std::memcpy(pos, s->pos.data(), s->N * 12); std::memcpy(vel, s->vel.data(), s->N * 12); std::memcpy(acc, s->acc.data(), s->N * 12);
}
void orc_swarm_get_boxes(void* sv, float* boxes) {
    Swarm* s = (Swarm*)sv;
    for (int i = 0; i < s->N * 5; i++) {
        boxes[6 * i] = s->box_min[i].x; boxes[6 * i + 1] = s->box_min[i].y; boxes[6 * i + 2] = s->box_min[i].z;
        boxes[6 * i + 3] = s->box_max[i].x; boxes[6 * i + 4] = s->box_max[i].y; boxes[6 * i + 5] = s->box_max[i].z;
    }
}
void orc_swarm_get_qp(void* sv, double* cost, int* status, int* iters, int* n_active, int* flags, double* maxviol, double* kkt) {
    Swarm* s = (Swarm*)sv;
    for (int a = 0; a < s->N; a++) {
        cost[a] = s->qp_cost[a]; status[a] = s->qp_status[a]; iters[a] = s->qp_iters[a]; n_active[a] = s->qp_active[a];
        flags[a] = s->flags[a]; maxviol[a] = s->qp_maxviol[a]; kkt[a] = s->qp_kkt[a];
    }
}
void orc_swarm_get_counters(void* sv, long long* c4) { Swarm* s = (Swarm*)sv; for (int i = 0; i < 4; i++) c4[i] = s->counters[i]; }
void orc_swarm_reset_counters(void* sv) { Swarm* s = (Swarm*)sv; for (int i = 0; i < 4; i++) s->counters[i] = 0; }
// captured constraints of the last step: normals [N][N][5][3], d [N][N][5][6], gjk iters [N][N][5]
void orc_swarm_get_capture(void* sv, float* normals, double* d, int* gjk) {
    Swarm* s = (Swarm*)sv;
    std::memcpy(normals, s->cap_normal.data(), s->cap_normal.size() * 12);
    std::memcpy(d, s->cap_d.data(), s->cap_d.size() * 8);
    std::memcpy(gjk, s->cap_gjk.data(), s->cap_gjk.size() * 4);
}

}  // extern "C"
