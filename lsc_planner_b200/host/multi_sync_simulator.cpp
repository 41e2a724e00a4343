// ROS-free MultiSyncSimulator: the reference's synchronous simulation loop (src/multi_sync_simulator.cpp:83-147,
// 190-337, 408-633) driving the GPU engine through the TrajPlanner facade. Output files have the reference's formats:
//   result CSV   id,t,px,py,pz,vx,vy,vz,ax,ay,az,planning_time,qp_cost,planning_report,size  per agent (:513-587)
//   summary CSV  start_time,total_flight_time,total_flight_distance,is_collided,safety_ratio_agent,...   (:589-633)
//
//   lsc_sim mission=<file.json> [world/file_name=<map.bt>] [key=value ...] [result=<out.csv>] [summary=<out.csv>]
// keys are the reference's ROS parameter names (launch/simulation.launch); defaults = simulation.launch with
// multisim/max_noise = 0 (the reference's noise is non-deterministic) and world/use_octomap = false unless a map is given.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>

#include "traj_planner.hpp"

using namespace DynamicPlanning;

class MultiSyncSimulator {
public:
    MultiSyncSimulator(const Param& _param, const Mission& _mission) : param(_param), mission(_mission) {
        batch = std::make_shared<ReplanBatch>(param, mission);
        agents.resize(mission.qn);
        for (int qi = 0; qi < mission.qn; qi++) agents[qi] = std::make_unique<TrajPlanner>(qi, param, mission, batch);
        ideal_states.resize(mission.qn);
        parseDisturbances();
    }
    void setOctomap(const std::string& file) { batch->setOctomap(file); }

    void run() {
        for (int iter = 0; iter < param.multisim_max_planner_iteration; iter++) {
            if (isFinished() || iter == param.multisim_max_planner_iteration - 1) { summarizeResult(); break; }
            if (initial_update) initializeTimer(); else doStep();
            update();
            if (!plan()) break;
        }
    }
    std::string result_file, summary_file;
    double total_flight_time = SP_INFINITY, total_distance = 0, safety_ratio_agent = SP_INFINITY;
    bool is_collided = false;
    int qp_failures = 0, n_resets = 0;

private:
    Param param;
    Mission mission;
    std::shared_ptr<ReplanBatch> batch;
    std::vector<std::unique_ptr<TrajPlanner>> agents;
    std::vector<State> ideal_states;
    PlanningTimeStatistics planning_time;
    double sim_start_time = 0, sim_current_time = 0;
    bool initial_update = true;
    std::vector<point3d> last_positions;

    int update_count = 0;
    struct Disturbance { int step, agent; point3d offset; };
    std::vector<Disturbance> disturbances;
    void parseDisturbances() {
        std::stringstream all(param.multisim_disturbance);
        std::string item;
        while (std::getline(all, item, ';')) {
            if (item.empty()) continue;
            Disturbance d{};
            float x = 0, y = 0, z = 0;
            if (std::sscanf(item.c_str(), "%d:%d:%f,%f,%f", &d.step, &d.agent, &x, &y, &z) != 5 || d.agent < 0 || d.agent >= mission.qn)
                throw std::invalid_argument("[MultiSyncSimulator] multisim/disturbance expects <step>:<agent>:<dx>,<dy>,<dz>;...");
            d.offset = point3d(x, y, z);
            disturbances.push_back(d);
        }
    }
    point3d disturbanceOffset(int step, int qi) const {
        point3d o(0, 0, 0);
        for (const Disturbance& d : disturbances) if (d.step == step && d.agent == qi) o = o + d.offset;
        return o;
    }
    void initializeTimer() { sim_start_time = 0; sim_current_time = 0; }
    void doStep() { sim_current_time += param.multisim_time_step; }

    // src/multi_sync_simulator.cpp:190-318: every agent's next initial state is its trajectory at t = time_step —
    // unless its observed position is farther than reset_threshold from where it should be NOW: then the agent restarts at
    // rest from the observed position (:229-246) and the engine's checks take it from there (slack variables, corridor).
    // The reference reads the observed pose from tf; here it is the ideal one plus the offsets of multisim/disturbance.
    void update() {
        for (int qi = 0; qi < mission.qn; qi++) {
            State s;
            if (initial_update) s.position = mission.agents[qi].start_position;        // at rest at the start point
            else {
                const State ideal_curr = agents[qi]->getCurrentStateMsg();
                const point3d real = ideal_curr.position + disturbanceOffset(update_count, qi);
                if ((ideal_curr.position - real).norm() > param.multisim_reset_threshold) {
                    s.position = real;                                                 // velocity, acceleration: zero
                    std::cerr << "[MultiSyncSimulator] agent " << qi << ": diff between ideal and real is too big\n";
                    n_resets++;
                } else s = agents[qi]->getFutureStateMsg(param.multisim_time_step);
            }
            ideal_states[qi] = s;
            agents[qi]->setCurrentState(s);
            agents[qi]->updatePlannerState(PlannerState::GOTO);
        }
        update_count++;
        if (initial_update) {
            last_positions.resize(mission.qn);
            for (int qi = 0; qi < mission.qn; qi++) last_positions[qi] = mission.agents[qi].start_position;
        }
        initial_update = false;
    }

    // :320-337 — the sequential loop over agents; the first collect() plans the whole swarm on the GPU
    bool plan() {
        for (int qi = 0; qi < mission.qn; qi++) {
            PlanningReport r = agents[qi]->plan(sim_current_time);
            if (r == PlanningReport::QPFAILED) return false;
        }
        for (int qi = 0; qi < mission.qn; qi++) {
            PlanningReport r = agents[qi]->collect();
            if (r == PlanningReport::QPFAILED) return false;
            if (agents[qi]->getQPStatus() != LSCGPU_QP_OK) qp_failures++;
        }
        savePlanningResult();
        if (!result_file.empty()) savePlanningResultAsCSV();
        return true;
    }

    bool isFinished() {                                                              // :358-381
        if (initial_update) return false;
        for (int qi = 0; qi < mission.qn; qi++) {
            const double dist = (agents[qi]->getCurrentPosition() - mission.agents[qi].desired_goal_position).norm();
            if (dist > param.goal_threshold) return false;
        }
        total_flight_time = sim_current_time - sim_start_time;
        return true;
    }

    void savePlanningResult() {                                                      // :408-511 (safety audit + timing)
        // minimum distance between agents over the recorded sub-times of the step: O(N^2) pairs, on the device
        // (lscgpu_safety_audit) from the trajectories the engine just committed
        if (mission.qn > 1) {
            std::vector<double> ratio(mission.qn);
            std::vector<int32_t> closest(mission.qn);
            if (lscgpu_safety_audit(batch->getEngine().get(), param.multisim_record_time_step, param.multisim_time_step,
                                    ratio.data(), closest.data()) != LSCGPU_OK)
                throw std::invalid_argument(std::string("[lscgpu] ") + lscgpu_last_error());
            for (int qi = 0; qi < mission.qn; qi++) {
                if (ratio[qi] < safety_ratio_agent) safety_ratio_agent = ratio[qi];
                if (ratio[qi] < 1) is_collided = true;
            }
        }
        for (int qi = 0; qi < mission.qn; qi++) {
            planning_time.update(agents[qi]->getPlanningTime());
            const point3d p = agents[qi]->getFutureStateMsg(param.multisim_time_step).position;
            total_distance += (p - last_positions[qi]).norm();
            last_positions[qi] = p;
        }
    }

    void savePlanningResultAsCSV() {                                                 // :513-587
        std::ofstream csv(result_file, std::ios_base::app);
        if (sim_current_time == sim_start_time) {
            for (int qi = 0; qi < mission.qn; qi++)
                csv << "id,t,px,py,pz,vx,vy,vz,ax,ay,az,planning_time,qp_cost,planning_report,size" << (qi < mission.qn - 1 ? "," : "\n");
        }
        const double record = param.multisim_record_time_step;
        double future_time = 0, t = sim_current_time - sim_start_time;
        while (future_time < param.multisim_time_step) {
            for (int qi = 0; qi < mission.qn; qi++) {
                const State s = agents[qi]->getFutureStateMsg(future_time);
                csv << qi << "," << t << "," << s.position.x() << "," << s.position.y() << "," << s.position.z() << ","
                    << s.velocity.x() << "," << s.velocity.y() << "," << s.velocity.z() << "," << s.acceleration.x() << ","
                    << s.acceleration.y() << "," << s.acceleration.z() << ","
                    << agents[qi]->getPlanningTime().total_planning_time.current << "," << agents[qi]->getQPCost() << ","
                    << agents[qi]->getPlanningReport() << "," << mission.agents[qi].radius << (qi < mission.qn - 1 ? "," : "\n");
            }
            future_time += record;
            t += record;
        }
    }

    void summarizeResult() {                                                         // :383-403, 589-633
        std::cout << "[MultiSyncSimulator] total flight time: " << total_flight_time << "\n"
                  << "[MultiSyncSimulator] total distance: " << total_distance << "\n"
                  << "[MultiSyncSimulator] planning time per agent: " << planning_time.total_planning_time.average << "\n"
                  << "[MultiSyncSimulator] goal planning time per agent: " << planning_time.goal_planning_time.average << "\n"
                  << "[MultiSyncSimulator] safety ratio between agent: " << safety_ratio_agent << "\n"
                  << "[MultiSyncSimulator] is_collided: " << is_collided << " qp_failures: " << qp_failures
                  << " state_resets: " << n_resets << "\n";
        if (summary_file.empty()) return;
        std::ifstream in(summary_file);
        const bool header = !in || in.peek() == std::ifstream::traits_type::eof();
        std::ofstream out(summary_file, std::ios_base::app);
        if (header)
            out << "start_time,total_flight_time,total_flight_distance,is_collided,safety_ratio_agent,"
                << "average_planning_time,min_planning_time,max_planning_time,"
                << "initial_traj_planning_time,obstacle_prediction_time,goal_planning_time,"
                << "lsc_generation_time,sfc_generation_time,traj_optimization_time,"
                << "mission_file_name,world_file_name,planner_mode,prediction_mode,initial_traj_mode,"
                << "slack_mode,goal_mode,world_dimension,dt,horizon,N_constraint_segments\n";
        out << sim_start_time << "," << total_flight_time << "," << total_distance << "," << is_collided << ","
            << safety_ratio_agent << "," << planning_time.total_planning_time.average << ","
            << planning_time.total_planning_time.min << "," << planning_time.total_planning_time.max << ","
            << planning_time.initial_traj_planning_time.average << "," << planning_time.obstacle_prediction_time.average << ","
            << planning_time.goal_planning_time.average << "," << planning_time.lsc_generation_time.average << ","
            << planning_time.sfc_generation_time.average << "," << planning_time.traj_optimization_time.average << ","
            << mission.mission_file_name << "," << mission.world_file_name << ",LSC,previous_solution,previous_solution,none,"
            // Param::getGoalModeStr indexes its table with planner_mode (src/param.cpp:168-171): "static" for the LSC planner
            // whatever mode/goal says; kept so that the column matches the reference's files
            << "static," << param.world_dimension << "," << param.dt << "," << param.horizon << "," << param.N_constraint_segments << "\n";
    }
};

int main(int argc, char** argv) {
    try {
        Param param = Param::simulationLaunch();
        param.multisim_max_noise = 0.0;         // multisim/max_noise=0.02 multisim/noise_seed=<n> for the launch file's value
        param.world_use_octomap = false;
        std::string result, summary, world;
        for (int i = 1; i < argc; i++) {
            const char* eq = std::strchr(argv[i], '=');
            if (!eq) throw std::invalid_argument(std::string("expected key=value, got ") + argv[i]);
            const std::string key(argv[i], eq - argv[i]), val(eq + 1);
            if (key == "result") result = val;
            else if (key == "summary") summary = val;
            else {
                param.set(key, val);
                if (key == "world/file_name") { world = val; param.world_use_octomap = true; }
            }
        }
        Mission mission;
        mission.initialize(param.mission_file_name, param.multisim_max_noise, param.world_dimension, param.world_z_2d, world,
                           param.multisim_noise_seed);
        MultiSyncSimulator sim(param, mission);
        if (param.world_use_octomap) {
            if (world.empty()) throw std::invalid_argument("world/use_octomap needs world/file_name");
            sim.setOctomap(world);
        }
        sim.result_file = result; sim.summary_file = summary;
        if (!result.empty()) std::remove(result.c_str());
        sim.run();
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "lsc_sim: %s\n", e.what());
        return 1;
    } catch (const PlanningReport& r) {
        std::fprintf(stderr, "lsc_sim: planning report %d\n", (int)r);
        return 2;
    }
}
