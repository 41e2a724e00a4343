// Host-visible launch wrappers of the engine's kernels (definitions in the .cu files of this directory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_common.cuh"

namespace lscgpu {

// ---- k_predict: initial trajectories / obstacle predictions, terminal segments, widened state -----------------
struct PredictLaunch {
    int n_agents, n_pad;           // n_pad: row length of the transposed prediction table
    int planner_seq;               // already incremented (src/traj_planner.cpp:127)
    double dt, reset_threshold;
    const lscgpu_agent_in* in;     // [N]
    const float* prev_traj;        // [N][90]
    const AgentConstDev* consts;   // [N]
    float* pred;                   // [N][90]    initial_traj, AoS
    float* predT;                  // [90][n_pad] same, element-major (coalesced reads by the LSC kernel)
    float* predZs;                 // [30][n_pad] z of every control point divided by the agent's own downwash ratio
    double* state9;                // [N][9]
    double* goal3;                 // [N][3]
    int* ts;                       // [N]
    int* flags;                    // [N]
    float4* sphere;                // [5][n_pad] bounding sphere (centre xyz, radius) of every segment's control points
    float* reach;                  // [N][5] how far ANY feasible control point of segment m can be from initial_traj's
};
void launch_predict(const PredictLaunch& L, cudaStream_t s);

// ---- k_goal_plan: goalPlanningWithPriority without an octomap ----------------------------------------------
struct GoalLaunch {
    int n_agents;
    double dt, goal_threshold, goal_radius, priority_dist_threshold;
    const lscgpu_agent_in* in;     // [N] goal = DESIRED goal
    const float* prev_traj;        // [N][90] traj_curr of every agent (obs_prev_trajs)
    const float* pred;             // [N][90] initial_traj
    const AgentConstDev* consts;
    double* goal3;                 // [N][3] current goal (overwrites what k_predict stored)
    int* ts;                       // [N]
    int* goal_kind;                // [N] 0 line-of-sight goal, 1 retreat from the closest higher-priority agent
};
void launch_goal_plan(const GoalLaunch& L, cudaStream_t s);

// ---- k_lsc_build ------------------------------------------------------------------------------------------
struct LscLaunch {
    int n_agents, n_pad, a0, n_local;
    const int* order;              // null, or scheduling order of the local agents (see QpLaunch::order)
    int first, count;              // this launch covers positions [first, first + count) of that order
    const float* pred;             // [N][90]
    const float* predT;            // [90][n_pad]
    const float* predZs;           // [30][n_pad]
    const AgentConstDev* consts;
    const float2* rdw;             // [N] (radius, downwash * radius) as float, for the culling pass
    const QpTablesDev* T;
    const double* state9;          // [N][9]
    const double* goal3;
    const int* ts;
    const float4* sphere;          // [5][n_pad]
    const float* reach;            // [N][5]
    RowRec* rows;                  // [n_local][P_pad]   one record per kept pair, kept-list order
    int P_pad;
    int* kept;                     // [n_local][P_pad] pair indices that survive the exact culling test
    int* kept_count;               // [n_local]  (zeroed by the launcher)
    double* safe;                  // [n_local][P_pad] per kept pair: smallest whitened slack of its rows at x0
    StepCounters* counters;
};
void launch_lsc_build(const LscLaunch& L, cudaStream_t s);

// one agent's LSCs recomputed into CollisionConstraints layout (debug / parity)
void launch_lsc_capture(int n_agents, int agent, const float* pred, const AgentConstDev* consts, float* normals,
                        double* d, cudaStream_t s);
void launch_gjk_batch(int n, const double* hulls, double* v, int* iters, cudaStream_t s);
// LSC arrays of the reference container -> row store (operator-level QP entry)
void launch_rows_from_lsc(int n_problems, const int* obs_offset, int total_obs, const float* lsc_normal,
                          const float* lsc_point, const double* lsc_d, RowRec* rows, int* kept,
                          int* kept_count, double* safe, cudaStream_t s);

void launch_terminal_segments(int n, const double* state9, const double* goal3, const int* agent_index,
                              const AgentConstDev* consts, double dt, int* ts_out, cudaStream_t s);

// ---- k_qp_solve -------------------------------------------------------------------------------------------
struct QpLaunch {
    int n_problems;
    const QpTablesDev* T;
    const AgentConstDev* consts;
    const int* order;              // null, or scheduling order: block i solves problem order[first + i]
    int first;                     // ... else problem first + i; n_problems = blocks of this launch
    const int* agent_index;        // null: agent = agent_base + b
    int agent_base;
    const double* state9;          // indexed by agent when agent_index == null, else by problem
    const double* goal3;
    const int* ts;
    const float* boxes;            // [..][5][6] or null (no SFC rows); indexed like state9
    float wmin[3], wmax[3];
    const RowRec* rows;
    const int* obs_offset;         // batch mode: obstacles of problem b = [obs_offset[b], obs_offset[b+1]); null: swarm mode
    int n_obs;                     // swarm mode: N-1
    int P_pad;                     // swarm mode: row pitch per agent
    const int* kept; const int* kept_count;   // pairs to price: swarm mode [b][P_pad]; batch mode at 5*obs_offset[b]
    double* safe;                             // per kept pair: travelled distance up to which it cannot be violated
    int max_iter;
    // outputs
    double* x_out;                 // [n_problems][90] or null
    double* cost_out; int* status_out; int* iters_out;   // batch mode
    lscgpu_agent_out* out;         // swarm mode: indexed by agent
    const float* prev_traj;        // [N][90]  (kept when the QP fails)
    double* last_cost;             // [N]
    const int* flags;              // [N]
    const int* goal_kind;          // [N] or null
    StepCounters* counters;
    long long* dbg;                // null, or [n_problems][8] section cycle counts (LSCGPU_QP_DEBUG)
};
void launch_qp_solve(const QpLaunch& L, cudaStream_t s);
void launch_qp_order(int n_local, int a0, const lscgpu_agent_out* out, int* order, cudaStream_t s);

// commit: every agent's new trajectory becomes traj_curr, advanced state becomes the next resident input
void launch_commit(int n_agents, const lscgpu_agent_out* out, float* prev_traj, lscgpu_agent_in* in, cudaStream_t s);

// safety audit of the planned step (src/multi_sync_simulator.cpp:446-475)
void launch_safety_audit(int n_agents, const float* traj, const AgentConstDev* consts, double dt, int n_samples,
                         double record_time_step, float* sample_pos /*[n_samples][N][3]*/, double* ratio, int* closest,
                         cudaStream_t s);

// ---- distance field / SFC ---------------------------------------------------------------------------------
struct DistMapDev {
    int size[3];                   // cells per axis
    int off[3];                    // signed key of cell 0
    uint8_t* sqdist;               // [x][y][z] squared cell distance clamped at max_sq
    int max_sq;
    int* sat;                      // [n_tables][(sx+1)][(sy+1)][(sz+1)] inclusive-exclusive prefix sums of "blocked"
    int n_tables;
};
void launch_edt_build(const int32_t* keys_dev, int n_keys, DistMapDev dm, const int* thresholds_dev, int n_tables,
                      uint8_t* scratch_a, uint8_t* scratch_b, cudaStream_t s);

struct SfcLaunch {
    int n;                         // seeds
    DistMapDev dm;
    double res;
    float wmin[3], wmax[3];
    // swarm mode (mode 0): seeds from agent inputs / previous trajectories, persistent windows updated in place
    int mode, agent_base, planner_window;
    const lscgpu_agent_in* in;
    const float* prev_traj;
    const AgentConstDev* consts;
    float* boxes;                  // [N][5][6]
    int* init_sfc;                 // [N]
    int* flags;                    // [N]
    // batch mode (mode 1)
    const float* point; const float* goal; const int* sat_index; float* box_out; int* ok_out;
};
void launch_sfc_expand(const SfcLaunch& L, cudaStream_t s);

}  // namespace lscgpu
