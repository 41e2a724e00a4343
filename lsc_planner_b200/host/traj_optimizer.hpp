// TrajOptimizer with the reference's class surface (include/traj_optimizer.hpp:18-28) on top of the C-ABI:
//   TrajOptimizer(param, mission)  — the reference also takes the Bernstein basis B (an Eigen matrix it only feeds to
//                                    buildQBase); the engine builds Q_base itself, so the argument is dropped
//   void solve(const Agent&, const CollisionConstraints&)   throws PlanningReport::QPFAILED like the reference
//   traj_t getTrajectory(); double getQPcost() const; void updateParam(const Param&)
// One optimizer object solves one QP per call (lscgpu_qp_solve_batch with a batch of one). The batched path used by
// the simulator is TrajPlanner / ReplanBatch (traj_planner.hpp).
#pragma once
#include <memory>

#include "../../include/lscgpu.h"
#include "collision_constraints.hpp"
#include "sp_const.hpp"

namespace DynamicPlanning {

inline lscgpu_params toEngineParams(const Param& param, const Mission& mission) {
    lscgpu_params p{};
    p.dt = param.dt;
    p.control_input_weight = param.control_input_weight;
    p.terminal_weight = param.terminal_weight;
    p.world_resolution = param.world_resolution;
    p.reset_threshold = param.multisim_reset_threshold;
    p.world_use_octomap = param.world_use_octomap ? 1 : 0;
    for (int k = 0; k < 3; k++) { p.world_min[k] = mission.world_min(k); p.world_max[k] = mission.world_max(k); }
    p.M = param.M; p.n = param.n; p.phi = param.phi; p.dim = param.world_dimension;
    // prior_based goal planning runs on the GPU inside the step; with goal/planner=host and an octomap the host grid planner
    // computes the goals before the step instead (goal_mode 0 for the engine)
    p.goal_mode = (param.goal_mode == GoalMode::PRIORBASED && (!param.world_use_octomap || param.goalPlannerOnDevice(mission.qn))) ? 1 : 0;
    p.goal_threshold = param.goal_threshold; p.goal_radius = param.goal_radius;
    p.priority_dist_threshold = param.priority_dist_threshold;
    p.grid_resolution = param.grid_resolution; p.grid_margin = param.grid_margin;
    return p;
}

inline std::vector<lscgpu_agent_const> toEngineAgents(const Mission& mission) {
    std::vector<lscgpu_agent_const> out(mission.qn);
    for (int qi = 0; qi < mission.qn; qi++) {
        const Agent& a = mission.agents[qi];
        out[qi].radius = a.radius; out[qi].downwash = a.downwash; out[qi].nominal_velocity = a.nominal_velocity;
        for (int k = 0; k < 3; k++) { out[qi].max_vel[k] = a.max_vel[k]; out[qi].max_acc[k] = a.max_acc[k]; }
    }
    return out;
}

struct EngineDeleter { void operator()(lscgpu_engine* e) const { lscgpu_destroy(e); } };
typedef std::shared_ptr<lscgpu_engine> EnginePtr;

inline EnginePtr createEngine(const Param& param, const Mission& mission, int device = 0) {
    if (std::abs(param.multisim_time_step - param.dt) > SP_EPSILON)
        throw std::invalid_argument("[TrajPlanner] multisim_time_step must be equal to segment time");   // traj_planner.cpp:433-436
    const lscgpu_params p = toEngineParams(param, mission);
    const std::vector<lscgpu_agent_const> ac = toEngineAgents(mission);
    lscgpu_engine* raw = nullptr;
    if (lscgpu_create(&p, mission.qn, ac.data(), device, &raw) != LSCGPU_OK)
        throw std::invalid_argument(std::string("[lscgpu] ") + lscgpu_last_error());
    EnginePtr engine(raw, EngineDeleter());
    // weight of the slack variables a state reset brings into the QPs (src/traj_optimizer.cpp:383-390)
    if (lscgpu_set_slack_collision_weight(raw, param.slack_collision_weight) != LSCGPU_OK)
        throw std::invalid_argument(std::string("[lscgpu] ") + lscgpu_last_error());
    return engine;
}

class TrajOptimizer {
public:
    TrajOptimizer(const Param& _param, const Mission& _mission, EnginePtr _engine = nullptr)
        : param(_param), mission(_mission), engine(std::move(_engine)) {
        M = param.M; n = param.n; phi = param.phi; dim = param.world_dimension;
        if (!engine) engine = createEngine(param, mission);
        trajectory.assign(M, std::vector<point3d>(n + 1));
    }

    void solve(const Agent& agent, const CollisionConstraints& constraints) {
        const int n_obs = (int)constraints.getObsSize();
        const int32_t agent_index = agent.id;
        double state[9], goal[3];
        for (int k = 0; k < 3; k++) {
            state[k] = agent.current_state.position(k);
            state[3 + k] = agent.current_state.velocity(k);
            state[6 + k] = agent.current_state.acceleration(k);
            goal[k] = agent.current_goal_position(k);
        }
        std::vector<float> sfc;
        if (param.world_use_octomap) {
            sfc.resize(M * 6);
            for (int m = 0; m < M; m++) {
                const Box box = constraints.getSFC(m).box;
                for (int k = 0; k < 3; k++) { sfc[m * 6 + k] = box.getBoxMin()(k); sfc[m * 6 + 3 + k] = box.getBoxMax()(k); }
            }
        }
        std::vector<float> normal((size_t)n_obs * M * 3), point((size_t)n_obs * M * (n + 1) * 3);
        std::vector<double> d((size_t)n_obs * M * (n + 1));
        for (int oi = 0; oi < n_obs; oi++)
            for (int m = 0; m < M; m++) {
                for (int i = 0; i < n + 1; i++) {
                    const LSC lsc = constraints.getLSC(oi, m, i);
                    const size_t r = ((size_t)oi * M + m) * (n + 1) + i;
                    for (int k = 0; k < 3; k++) point[r * 3 + k] = lsc.obs_control_point(k);
                    d[r] = lsc.d;
                    if (i == 0) for (int k = 0; k < 3; k++) normal[((size_t)oi * M + m) * 3 + k] = lsc.normal_vector(k);
                }
            }
        const int32_t obs_offset[2] = {0, n_obs};
        double x[LSCGPU_TRAJ_FLOATS], cost = 0;
        int32_t status = 0, iterations = 0;
        // obstacles of obs_slack_indices get slack variables (src/traj_optimizer.cpp:268,317-326,455-457)
        std::vector<uint8_t> obs_slack((size_t)std::max(n_obs, 1), 0);
        for (int oi : constraints.getSlackIndices()) if (oi >= 0 && oi < n_obs) obs_slack[oi] = 1;
        const int rc = lscgpu_qp_solve_batch_slack(engine.get(), 1, &agent_index, state, goal, sfc.empty() ? nullptr : sfc.data(),
                                                   obs_offset, normal.data(), point.data(), d.data(), obs_slack.data(), x, &cost,
                                                   &status, &iterations, nullptr);
        if (rc != LSCGPU_OK) throw std::invalid_argument(std::string("[lscgpu] ") + lscgpu_last_error());
        if (status != LSCGPU_QP_OK) throw PlanningReport::QPFAILED;          // src/traj_optimizer.cpp:143,152
        for (int k = 0; k < dim; k++)
            for (int m = 0; m < M; m++)
                for (int i = 0; i < n + 1; i++) trajectory[m][i](k) = (float)x[k * M * (n + 1) + m * (n + 1) + i];   // :84-86
        current_qp_cost = cost;
    }

    void updateParam(const Param& _param) { param = _param; }
    traj_t getTrajectory() { return trajectory; }
    double getQPcost() const { return current_qp_cost; }

private:
    Param param;
    Mission mission;
    EnginePtr engine;
    traj_t trajectory;
    double current_qp_cost = 0;
    int M, n, phi, dim;
};

}  // namespace DynamicPlanning
