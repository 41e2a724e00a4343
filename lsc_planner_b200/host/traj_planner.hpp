// TrajPlanner with the reference's class surface (include/traj_planner.hpp:49-101) on top of the batched C-ABI.
//
// The reference plans agents one after the other: MultiSyncSimulator::plan() loops `agents[qi]->plan(t)`
// (src/multi_sync_simulator.cpp:320-337) and every planner sees the SAME snapshot of the swarm (update() ran before,
// :190-318). Here the N planners of a simulator share one ReplanBatch: the first plan() of a new planner_seq sends
// every agent's state to the GPU (lscgpu_replan_batch) and plans the whole swarm; the other N-1 plan() calls of that
// step just pick up their result. Setters that fed the reference's per-agent copies of the swarm (setObstacles,
// setObsPrevTrajs, setDistMap) are accepted and ignored: the engine already holds every trajectory and the map.
#pragma once
#include <chrono>
#include <cmath>
#include <memory>
#include <vector>

#include <exception>
#include <thread>

#include "grid_based_planner.hpp"
#include "traj_optimizer.hpp"

namespace DynamicPlanning {

inline double nChoosek(int n, int k) {
    if (k > n) return 0;
    double r = 1;
    for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}
inline double getBernsteinBasis(int n, int i, double t) { return nChoosek(n, i) * std::pow(t, i) * std::pow(1 - t, n - i); }
inline point3d getPointFromControlPoints(const std::vector<point3d>& cps, double t) {      // include/polynomial.hpp:26-45
    if (t < 0 - SP_EPSILON || t > 1 + SP_EPSILON) throw std::invalid_argument("[Polynomial] Input of getPointFromControlPoints is out of bound");
    const int n_ctrl = (int)cps.size() - 1;
    double x = 0, y = 0, z = 0;
    for (int i = 0; i < n_ctrl + 1; i++) {
        const double b = getBernsteinBasis(n_ctrl, i, t);
        x += cps[i].x() * b; y += cps[i].y() * b; z += cps[i].z() * b;
    }
    return point3d((float)x, (float)y, (float)z);
}
// include/polynomial.hpp:63-121 (position, velocity, acceleration)
inline State getStateFromControlPoints(const traj_t& control_points, double current_time, int M, int n, double dt) {
    int m = static_cast<int>(current_time / dt);
    if (m == M && current_time < M * dt + SP_EPSILON) m = M - 1;
    else if (m >= M) throw std::invalid_argument("[Polynomial] Input of getOdom is out of bound");
    State state;
    const double tn = current_time / dt - m;
    state.position = getPointFromControlPoints(control_points[m], tn);
    std::vector<point3d> vel(n), acc(n - 1);
    for (int i = 0; i < n; i++) vel[i] = (control_points[m][i + 1] - control_points[m][i]) * (float)n * (float)std::pow(dt, -1);
    state.velocity = getPointFromControlPoints(vel, tn);
    for (int i = 0; i < n - 1; i++) acc[i] = (vel[i + 1] - vel[i]) * (float)(n - 1) * (float)std::pow(dt, -1);
    state.acceleration = getPointFromControlPoints(acc, tn);
    return state;
}

// One per simulator: owns the engine and the host-side in/out records of all agents.
class ReplanBatch {
public:
    ReplanBatch(const Param& _param, const Mission& _mission, int device = 0)
        : param(_param), mission(_mission), engine(createEngine(_param, _mission, device)), in(_mission.qn), out(_mission.qn),
          fresh(_mission.qn, false) {
        // per-phase times come from the cycle counters of the result records (lsc_kcycles / qp_kcycles): the engine
        // keeps its normal scheduling (profiling mode would serialise the kernels of every step)
        sm_clock_hz = 1e3 * std::max(1, lscgpu_sm_clock_khz(engine.get()));
    }
    void setProfiling(bool on) { lscgpu_set_profiling(engine.get(), on ? 1 : 0); }
    double smClockHz() const { return sm_clock_hz; }
    EnginePtr getEngine() const { return engine; }

    void setOctomap(const std::string& file) {      // MultiSyncSimulator::setOctomap (src/multi_sync_simulator.cpp:153-167)
        if (lscgpu_set_octomap_file(engine.get(), file.c_str()) != LSCGPU_OK)
            throw std::invalid_argument(std::string("[lscgpu] ") + lscgpu_last_error());
    }
    // goal: the current goal (GoalMode::STATIC) or the desired goal (GoalMode::PRIORBASED, resolved in ensurePlanned)
    void setInput(int qi, const State& s, const point3d& goal) {
        for (int k = 0; k < 3; k++) {
            in[qi].position[k] = s.position(k); in[qi].velocity[k] = s.velocity(k);
            in[qi].acceleration[k] = s.acceleration(k); in[qi].goal[k] = goal(k);
        }
        fresh[qi] = true;
    }
    bool isFresh(int qi) const { return fresh[qi]; }
    // plans the whole swarm once per planner_seq
    void ensurePlanned(int planner_seq) {
        if (planned_seq >= planner_seq) return;
        for (bool f : fresh) if (!f) throw PlanningReport::WAITFORROSMSG;
        const auto t0 = std::chrono::steady_clock::now();
        goal_seconds = 0;
        if (param.goal_mode == GoalMode::PRIORBASED && param.world_use_octomap && !param.goalPlannerOnDevice(mission.qn)) {
            planGoalsOnHost(planner_seq);
            goal_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        if (lscgpu_replan_batch(engine.get(), in.data(), out.data()) != LSCGPU_OK)
            throw std::invalid_argument(std::string("[lscgpu] ") + lscgpu_last_error());
        wall_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        lscgpu_get_step_stats(engine.get(), &stats);
        astar_expansions += stats.astar_expansions;          // goal planning on the device (k_goal_astar)
        planned_seq = planner_seq;
        std::fill(fresh.begin(), fresh.end(), false);
    }
    const lscgpu_agent_out& result(int qi) const { return out[qi]; }
    double wallSecondsPerAgent() const { return wall_seconds / std::max(1, mission.qn); }
    double goalSecondsPerAgent() const { return goal_seconds / std::max(1, mission.qn); }
    long long astarExpansions() const { return astar_expansions; }
    const lscgpu_step_stats& stepStats() const { return stats; }

private:
    // goalPlanningWithPriority for every agent (src/traj_planner.cpp:540-608) with the grid planner and A* on the host,
    // agents spread over the host threads. Inputs are exactly what MultiSyncSimulator::update() hands every planner
    // (src/multi_sync_simulator.cpp:269-299): the others' current positions, desired goals and previous trajectories
    // (the result records of the previous step; zero before the first, src/traj_planner.cpp:36-39).
    void planGoalsOnHost(int planner_seq) {
        const int N = mission.qn;
        if (distmap.sqdist.empty()) {
            int32_t size[3], off[3]; int64_t n_occ = 0;
            if (lscgpu_get_distmap_info(engine.get(), size, off, &n_occ) != LSCGPU_OK)
                throw std::invalid_argument(std::string("[lscgpu] ") + lscgpu_last_error());
            distmap.res = param.world_resolution;
            for (int k = 0; k < 3; k++) { distmap.size[k] = size[k]; distmap.off[k] = off[k]; }
            distmap.sqdist.resize((size_t)size[0] * size[1] * size[2]);
            if (lscgpu_get_distmap_sqdist(engine.get(), distmap.sqdist.data()) != LSCGPU_OK)
                throw std::invalid_argument(std::string("[lscgpu] ") + lscgpu_last_error());
        }
        if (grid_cache.by_radius.empty()) {
            GridBasedPlanner probe(&distmap, mission, param);
            for (int j = 0; j < N; j++)
                if (!grid_cache.find(mission.agents[j].radius))
                    grid_cache.by_radius.emplace_back(mission.agents[j].radius, probe.staticGrid(mission.agents[j].radius));
        }
        auto P = [](const float* v) { return point3d(v[0], v[1], v[2]); };
        std::vector<GoalObstacle> all(N);
        for (int j = 0; j < N; j++) {
            all[j].id = j; all[j].position = P(in[j].position); all[j].goal_point = P(in[j].goal);
            all[j].radius = mission.agents[j].radius; all[j].downwash = mission.agents[j].downwash;
            all[j].prev_traj_first_end = P(out[j].traj[0][5]); all[j].prev_traj_last_end = P(out[j].traj[4][5]);
        }
        std::vector<point3d> goals(N);
        std::vector<long long> expanded(N, 0);
        const int threads = (int)std::max(1u, std::min((unsigned)N, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        std::vector<std::exception_ptr> errors(threads);      // an exception must not leave a std::thread body
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, t]() {
              try {
                std::vector<GoalObstacle> obstacles;
                for (int a = t; a < N; a += threads) {
                    obstacles.clear();
                    for (int j = 0; j < N; j++) if (j != a) obstacles.push_back(all[j]);
                    const point3d pos = P(in[a].position), vel = P(in[a].velocity);
                    // initial_traj[M-1][n]: straight line at the first step (src/traj_planner.cpp:1030-1037), else the end
                    // of the previous trajectory (:997-1016)
                    const point3d init_end = planner_seq < 2 ? pos + (vel * (float)5.0) * (float)param.dt : P(out[a].traj[4][5]);
                    const GoalPlanResult r = goalPlanningWithPriority(pos, P(in[a].goal), init_end, mission.agents[a].radius,
                                                                      mission.agents[a].downwash, obstacles, &distmap, mission, param, &grid_cache);
                    goals[a] = r.goal; expanded[a] = r.expansions;
                }
              } catch (...) { errors[t] = std::current_exception(); }
            });
        for (auto& th : pool) th.join();
        for (auto& err : errors) if (err) std::rethrow_exception(err);
        for (int a = 0; a < N; a++) {
            for (int k = 0; k < 3; k++) in[a].goal[k] = goals[a](k);
            astar_expansions += expanded[a];
        }
    }

    Param param;
    Mission mission;
    EnginePtr engine;
    HostDistMap distmap;
    StaticGridCache grid_cache;
    double goal_seconds = 0;
    long long astar_expansions = 0;
    std::vector<lscgpu_agent_in> in;
    std::vector<lscgpu_agent_out> out;
    std::vector<bool> fresh;
    int planned_seq = 0;
    double wall_seconds = 0;
    double sm_clock_hz = 1.965e9;
    lscgpu_step_stats stats{};
};

class TrajPlanner {
public:
    TrajPlanner(int _agent_id, const Param& _param, const Mission& _mission, std::shared_ptr<ReplanBatch> _batch)
        : param(_param), mission(_mission), batch(std::move(_batch)) {
        agent = mission.agents[_agent_id];
        M = param.M; n = param.n; dim = param.world_dimension;
        traj_curr.assign(M, std::vector<point3d>(n + 1));               // src/traj_planner.cpp:36-39
        planner_seq = 0;
        planning_report = PlanningReport::Initialized;
        if (param.planner_mode != PlannerMode::LSC) throw std::invalid_argument("[TrajPlanner] Invalid planner mode");
    }

    PlanningReport plan(double /*sim_current_time*/) {
        if (!flag_current_state_updated) return PlanningReport::WAITFORROSMSG;      // src/traj_planner.cpp:101-110
        planner_seq++;                                                                // :127
        goalPlanning();
        batch->setInput(agent.id, agent.current_state, agent.current_goal_position);
        flag_planned = false;
        return planning_report = PlanningReport::SUCCESS;   // provisional; collect() below finalises after the batch ran
    }
    // second half of plan(): called for every agent after all of them pushed their inputs
    PlanningReport collect() {
        if (flag_planned) return planning_report;
        batch->ensurePlanned(planner_seq);
        const lscgpu_agent_out& o = batch->result(agent.id);
        // the reference throws from expandBoxFromPoint when the SFC seed cell is occupied
        // (include/corridor_constructor.hpp:35-38); the engine reports it as a flag
        if (o.flags & LSCGPU_FLAG_SFC_SEED_BLOCKED)
            throw std::invalid_argument("[CorridorConstructor] invalid initial trajectory, obstacle exists at the initial trajectory");
        for (int m = 0; m < M; m++)
            for (int i = 0; i < n + 1; i++) traj_curr[m][i] = point3d(o.traj[m][i][0], o.traj[m][i][1], o.traj[m][i][2]);
        agent.current_goal_position = point3d(o.current_goal[0], o.current_goal[1], o.current_goal[2]);
        current_qp_cost = o.qp_cost;
        qp_status = o.qp_status;
        planning_report = (PlanningReport)o.report;
        const lscgpu_step_stats& st = batch->stepStats();
        const double per = 1e-3 / std::max(1, mission.qn);
        PlanningTimeStatistics t;
        t.obstacle_prediction_time.current = 0.5 * st.ms_predict * per;
        t.initial_traj_planning_time.current = 0.5 * st.ms_predict * per;
        t.goal_planning_time.current = batch->goalSecondsPerAgent();
        // corridors (LSC rows with the SFC box grown beside them) and QP of THIS agent, from its block's cycle counters
        const double corridor_s = 1024.0 * o.lsc_kcycles / batch->smClockHz();
        t.lsc_generation_time.current = corridor_s;
        t.sfc_generation_time.current = param.world_use_octomap ? corridor_s : 0.0;
        t.traj_optimization_time.current = 1024.0 * std::max(0, o.qp_kcycles - o.lsc_kcycles) / batch->smClockHz();
        t.total_planning_time.current = batch->wallSecondsPerAgent();
        planning_time.update(t);
        flag_current_state_updated = false;                                           // :132-136
        flag_planned = true;
        return planning_report;
    }

    // Setters
    void setCurrentState(const State& s) { agent.current_state = s; flag_current_state_updated = true; }
    template <class T> void setDistMap(const T&) {}
    template <class T> void setObstacles(const T&) {}
    void setObsPrevTrajs(const std::vector<traj_t>&) {}
    void updatePlannerState(const PlannerState& s) { planner_state = s; }
    void setStart(const point3d& p) { agent.start_position = p; }
    void setDesiredGoal(const point3d& p) { agent.desired_goal_position = p; }

    // Getters
    point3d getCurrentPosition() const { return agent.current_state.position; }
    State getCurrentStateMsg() const { return agent.current_state; }
    State getFutureStateMsg(double future_time) const { return getStateFromControlPoints(traj_curr, future_time, M, n, param.dt); }
    PlanningTimeStatistics getPlanningTime() const { return planning_time; }
    double getQPCost() const { return current_qp_cost; }
    int getQPStatus() const { return qp_status; }
    PlanningReport getPlanningReport() const { return planning_report; }
    traj_t getTraj() const { return traj_curr; }
    point3d getCurrentGoalPosition() const { return agent.current_goal_position; }
    point3d getDesiredGoalPosition() const { return agent.desired_goal_position; }
    int getPlannerSeq() const { return planner_seq; }
    point3d getNormalVector(int obs_id, int m) const {
        std::vector<float> nr((size_t)(mission.qn - 1) * M * 3);
        std::vector<double> d((size_t)(mission.qn - 1) * M * (n + 1));
        if (lscgpu_get_lsc(batch->getEngine().get(), agent.id, nr.data(), d.data()) != LSCGPU_OK)
            throw std::invalid_argument(std::string("[lscgpu] ") + lscgpu_last_error());
        const int oi = obs_id < agent.id ? obs_id : obs_id - 1;
        return point3d(nr[((size_t)oi * M + m) * 3], nr[((size_t)oi * M + m) * 3 + 1], nr[((size_t)oi * M + m) * 3 + 2]);
    }

private:
    // Goal planning (src/traj_planner.cpp:477-608) is the step BEFORE the path. GoalMode::STATIC: current goal = desired
    // goal. GoalMode::PRIORBASED (the reference's default): the desired goal is handed to the batch, which resolves it for
    // all agents at once — on the GPU without an octomap (k_goal_plan), by the host grid planner with one
    // (grid_based_planner.hpp) — and collect() reads the chosen goal back. ORCA / right-hand modes: not provided.
    void goalPlanning() {
        if (param.goal_mode != GoalMode::STATIC && param.goal_mode != GoalMode::PRIORBASED)
            throw std::invalid_argument("[TrajPlanner] Invalid goal mode");
        agent.current_goal_position = planner_state == PlannerState::GOBACK ? agent.start_position : agent.desired_goal_position;
    }

    Param param;
    Mission mission;
    std::shared_ptr<ReplanBatch> batch;
    Agent agent;
    traj_t traj_curr;
    PlannerState planner_state = PlannerState::GOTO;
    PlanningReport planning_report;
    PlanningTimeStatistics planning_time;
    double current_qp_cost = 0;
    int qp_status = 0;
    int planner_seq, M, n, dim;
    bool flag_current_state_updated = false, flag_planned = true;
};

}  // namespace DynamicPlanning
