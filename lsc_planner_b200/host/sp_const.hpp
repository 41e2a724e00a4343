// Host-side types of the drop-in: same names, members and meaning as the reference's
// include/sp_const.hpp, include/param.hpp and include/mission.hpp, without ROS / octomap / Eigen.
#pragma once
#include <cmath>
#include <fstream>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "mini_json.hpp"

namespace DynamicPlanning {

#define SP_EPSILON 1e-9
#define SP_EPSILON_FLOAT 1e-5
#define SP_INFINITY 1e+9

// octomap::point3d (octomath::Vector3): three float32; arithmetic in float, dot()/norm() returned as double
struct point3d {
    float data[3] = {0.f, 0.f, 0.f};
    point3d() = default;
    point3d(float x, float y, float z) { data[0] = x; data[1] = y; data[2] = z; }
    float& x() { return data[0]; } float& y() { return data[1]; } float& z() { return data[2]; }
    float x() const { return data[0]; } float y() const { return data[1]; } float z() const { return data[2]; }
    float& operator()(int i) { return data[i]; }
    float operator()(int i) const { return data[i]; }
    point3d operator+(const point3d& o) const { return point3d(x() + o.x(), y() + o.y(), z() + o.z()); }
    point3d operator-(const point3d& o) const { return point3d(x() - o.x(), y() - o.y(), z() - o.z()); }
    point3d operator*(float s) const { return point3d(x() * s, y() * s, z() * s); }
    double dot(const point3d& o) const { return (double)(x() * o.x() + y() * o.y() + z() * o.z()); }
    double norm() const { return std::sqrt(dot(*this)); }
    point3d normalized() const {
        point3d r(*this);
        const double len = norm();
        if (len > 0) { const float l = (float)len; r.x() /= l; r.y() /= l; r.z() /= l; }
        return r;
    }
};

typedef std::vector<std::vector<point3d>> traj_t;     // [segment m][control point i]

enum class PlannerMode { LSC, BVC, ORCA };
enum class GoalMode { STATIC, ORCA, RIGHTHAND, PRIORBASED };
enum PlannerState { WAIT, GOTO, PATROL, GOBACK };
enum PlanningReport { Initialized, INITTRAJGENERATIONFAILED, CONSTRAINTGENERATIONFAILED, QPFAILED, WAITFORROSMSG, SUCCESS };

struct PlanningTime {
    void update(double time) {
        current = time;
        if (time < min) min = time;
        if (time > max) max = time;
        N_sample++;
        average = (average * (N_sample - 1) + time) / N_sample;
    }
    double current = 0, min = SP_INFINITY, max = 0, average = 0;
    int N_sample = 0;
};

struct PlanningTimeStatistics {
    void update(const PlanningTimeStatistics& t) {
        initial_traj_planning_time.update(t.initial_traj_planning_time.current);
        obstacle_prediction_time.update(t.obstacle_prediction_time.current);
        goal_planning_time.update(t.goal_planning_time.current);
        lsc_generation_time.update(t.lsc_generation_time.current);
        sfc_generation_time.update(t.sfc_generation_time.current);
        traj_optimization_time.update(t.traj_optimization_time.current);
        total_planning_time.update(t.total_planning_time.current);
    }
    PlanningTime initial_traj_planning_time, obstacle_prediction_time, goal_planning_time, lsc_generation_time,
        sfc_generation_time, traj_optimization_time, total_planning_time;
};

struct State { point3d position, velocity, acceleration; };

struct Agent {
    int id = 0, cid = 0;
    State current_state;
    point3d start_position, desired_goal_position, current_goal_position;
    std::vector<double> max_vel, max_acc;
    double radius = 0.15, downwash = 2.0, nominal_velocity = 1.0;
};

// include/param.hpp with the defaults of src/param.cpp:4-107; set(key, value) takes the ROS parameter names.
struct Param {
    bool log = false;
    int world_dimension = 3;
    bool world_use_octomap = false;
    double world_resolution = 0.1, world_z_2d = 1.0;
    int multisim_planning_rate = -1, multisim_qn = 2;
    double multisim_time_step = 0.1;
    double multisim_max_noise = 0.0;
    long long multisim_noise_seed = -1;     // not in the reference: >= 0 seeds Mission::addNoise (the reference uses std::random_device)
    // not in the reference (it listens to tf): externally observed poses for MultiSyncSimulator::update, as
    // "<step>:<agent>:<dx>,<dy>,<dz>;..." — at that update the observed position of the agent is its ideal one + the offset
    std::string multisim_disturbance;
    int multisim_max_planner_iteration = 1000;
    bool multisim_save_result = false, multisim_experiment = false;
    double multisim_record_time_step = 0.1, multisim_reset_threshold = 0.1;
    PlannerMode planner_mode = PlannerMode::LSC;
    GoalMode goal_mode = GoalMode::PRIORBASED;
    double dt = 0.5, horizon = 2.0;
    int n = 5, phi = 3, phi_n = 1, M = 4;
    double control_input_weight = 1, terminal_weight = 1, slack_collision_weight = 1;
    int N_constraint_segments = -1;
    double grid_resolution = 0.3, grid_margin = 0.1;            // src/param.cpp:93-94
    // not in the reference: where prior_based goal planning with an octomap runs. device: on the GPU inside the step
    // (k_goal_astar, one warp per agent); host: grid_based_planner.hpp on the host threads before the step (same goals bit
    // for bit). One A* is a sequential search: a host core runs it ~10x faster than a GPU lane, and the GPU step lasts as
    // long as its longest search (~40 ms in the shipped 10 m world, whatever the swarm size), while the host needs
    // ~0.5 ms of one core per agent (measured, profiles/r02e_goal_crossover.txt: 16 threads beat the device up to 512
    // agents). auto (default): the device only when the swarm is large for the host cores at hand (>= 80 agents per thread,
    // e.g. several ranks sharing one node's cores).
    int goal_planner = 0;               // 0 auto, 1 device, 2 host
    bool goalPlannerOnDevice(int n_agents) const {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        return goal_planner == 1 || (goal_planner == 0 && (unsigned)n_agents >= 80u * hw);
    }
    double goal_threshold = 0.1, goal_radius = 100.0, priority_dist_threshold = 0.4;
    std::string mission_file_name = "default.json", world_file_name = "default.bt", package_path = ".";

    // launch/simulation.launch values (the configuration every shipped launch file uses)
    static Param simulationLaunch() {
        Param p;
        p.world_use_octomap = true; p.multisim_time_step = 0.2; p.multisim_max_noise = 0.02;
        p.multisim_reset_threshold = 0.15; p.dt = 0.2; p.horizon = 1.0; p.control_input_weight = 0.01;
        p.terminal_weight = 1; p.slack_collision_weight = 1e5; p.goal_radius = 2.0; p.grid_resolution = 0.25;
        p.multisim_save_result = true;
        p.finalize();
        return p;
    }
    void finalize() {
        M = (int)std::lround(horizon / dt);                  // src/param.cpp: M = horizon / dt
        if (N_constraint_segments < 0) N_constraint_segments = M;
    }
    void set(const std::string& key, const std::string& v) {
        auto d = [&]() { return std::stod(v); };
        auto i = [&]() { return std::stoi(v); };
        auto b = [&]() { return v == "true" || v == "1"; };
        if (key == "mission") mission_file_name = v;
        else if (key == "log") log = b();
        else if (key == "world/file_name") world_file_name = v;
        else if (key == "world/dimension") world_dimension = i();
        else if (key == "world/use_octomap") world_use_octomap = b();
        else if (key == "world/resolution") world_resolution = d();
        else if (key == "multisim/time_step") multisim_time_step = d();
        else if (key == "multisim/max_noise") multisim_max_noise = d();
        else if (key == "multisim/noise_seed") multisim_noise_seed = std::stoll(v);
        else if (key == "multisim/disturbance") multisim_disturbance = v;
        else if (key == "multisim/max_planner_iteration") multisim_max_planner_iteration = i();
        else if (key == "multisim/save_result") multisim_save_result = b();
        else if (key == "multisim/record_time_step") multisim_record_time_step = d();
        else if (key == "multisim/reset_threshold") multisim_reset_threshold = d();
        else if (key == "mode/planner") {
            if (v != "lsc") throw std::invalid_argument("[Param] only mode/planner = lsc is supported by the GPU path");
        } else if (key == "mode/goal") {
            if (v == "right_hand") goal_mode = GoalMode::RIGHTHAND;
            else if (v == "prior_based") goal_mode = GoalMode::PRIORBASED;
            else if (v == "static") goal_mode = GoalMode::STATIC;
            else throw std::invalid_argument("[Param] Invalid goal mode");
        }
        else if (key == "traj/dt") dt = d();
        else if (key == "traj/horizon") horizon = d();
        else if (key == "traj/n") n = i();
        else if (key == "traj/phi") phi = i();
        else if (key == "traj/phi_n") phi_n = i();
        else if (key == "opt/control_input_weight") control_input_weight = d();
        else if (key == "opt/terminal_weight") terminal_weight = d();
        else if (key == "opt/slack_collision_weight") slack_collision_weight = d();
        else if (key == "opt/N_constraint_segments") N_constraint_segments = i();
        else if (key == "grid/resolution") grid_resolution = d();
        else if (key == "grid/margin") grid_margin = d();
        else if (key == "goal/planner") {
            if (v == "device") goal_planner = 1;
            else if (v == "host") goal_planner = 2;
            else if (v == "auto") goal_planner = 0;
            else throw std::invalid_argument("[Param] goal/planner must be auto, device or host");
        }
        else if (key == "plan/goal_threshold") goal_threshold = d();
        else if (key == "plan/goal_radius") goal_radius = d();
        else if (key == "plan/priority_dist_threshold") priority_dist_threshold = d();
        else throw std::invalid_argument("[Param] unknown parameter: " + key);
        finalize();
    }
};

// include/mission.hpp: src/mission.cpp:20-319 (quadrotors, world[0].dimension, agents[{type,cid,start,goal}])
struct Mission {
    int qn = 0, on = 0;
    std::vector<Agent> agents;
    point3d world_min, world_max;
    std::string mission_file_name, world_file_name;

    // src/mission.cpp:386-395: every desired goal moves by U(0,1) * max_noise per axis (float draws from an mt19937; the
    // reference seeds it from std::random_device, here a seed >= 0 makes the run repeatable)
    void addNoise(double max_noise, int dimension, long long seed = -1) {
        std::random_device rd;
        std::mt19937 gen(seed >= 0 ? (std::mt19937::result_type)seed : rd());
        std::uniform_real_distribution<float> dis(0, 1);
        for (int qi = 0; qi < qn; qi++)
            for (int k = 0; k < dimension; k++) agents[qi].desired_goal_position(k) += dis(gen) * max_noise;
    }

    bool initialize(const std::string& file, double max_noise = 0.0, int world_dimension = 3, double world_z_2d = 1.0,
                    const std::string& world_file = "", long long noise_seed = -1) {
        mission_file_name = file; world_file_name = world_file;
        std::ifstream in(file);
        if (!in) throw std::invalid_argument("[Mission] There is no such file: " + file);
        std::stringstream ss; ss << in.rdbuf();
        const mini_json::Value doc = mini_json::parse(ss.str());
        const mini_json::Value& dim = doc["world"][0]["dimension"];
        world_min = point3d((float)dim[0].number(), (float)dim[1].number(), (float)dim[2].number());
        world_max = point3d((float)dim[3].number(), (float)dim[4].number(), (float)dim[5].number());
        const mini_json::Value& quad = doc["quadrotors"];
        const mini_json::Value& ag = doc["agents"];
        qn = (int)ag.size();
        agents.resize(qn);
        for (int qi = 0; qi < qn; qi++) {
            const mini_json::Value& a = ag[qi];
            const std::string type = a.has("type") ? a["type"].string() : std::string("default");
            if (!quad.has(type)) throw std::invalid_argument("[Mission] unknown quadrotor type: " + type);
            const mini_json::Value& q = quad[type];
            Agent& A = agents[qi];
            A.id = qi;
            A.cid = a.has("cid") ? (int)a["cid"].number() : qi;
            A.start_position = point3d((float)a["start"][0].number(), (float)a["start"][1].number(), (float)a["start"][2].number());
            A.desired_goal_position = point3d((float)a["goal"][0].number(), (float)a["goal"][1].number(), (float)a["goal"][2].number());
            if (world_dimension == 2) { A.start_position.z() = (float)world_z_2d; A.desired_goal_position.z() = (float)world_z_2d; }
            for (int k = 0; k < 3; k++) { A.max_vel.push_back(q["max_vel"][k].number()); A.max_acc.push_back(q["max_acc"][k].number()); }
            A.radius = q["radius"].number();
            A.downwash = q["downwash"].number();
            A.nominal_velocity = q["nominal_velocity"].number();
            A.current_state.position = A.start_position;
            A.current_goal_position = A.desired_goal_position;
        }
        on = doc.has("obstacles") ? (int)doc["obstacles"].size() : 0;
        if (on != 0) throw std::invalid_argument("[Mission] dynamic obstacles are outside the GPU path (all shipped missions have none)");
        addNoise(max_noise, world_dimension, noise_seed);                               // src/mission.cpp:317
        for (Agent& A : agents) A.current_goal_position = A.desired_goal_position;
        return true;
    }
};

}  // namespace DynamicPlanning
