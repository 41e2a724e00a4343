// Host-visible launch wrappers of the engine's kernels (definitions in the .cu files of this directory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "astar_core.cuh"
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
    float4* tsphere;               // [n_pad] bounding sphere of ALL control points of the trajectory (coarse culling pass)
    float* reach;                  // [N][5] how far ANY feasible control point of segment m can be from initial_traj's
    // disturbance branch: a state farther than reset_threshold from the start of the shifted previous trajectory collapses
    // the agent's prediction / initial trajectory to the observed position, marks it for good and re-arms its corridor
    unsigned char* reset_ever;     // [N] sticky
    int* any_reset;                // device word
    int* init_sfc;                 // [N] flag_initialize_sfc
};
void launch_predict(const PredictLaunch& L, cudaStream_t s);

// ---- k_goal_plan: goalPlanningWithPriority without an octomap ----------------------------------------------
struct GoalLaunch {
    const unsigned char* reset_ever;   // [N] obstacles of obs_slack_indices are high priority but never the retreat target (:548-551)
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

// ---- k_goal_astar: goalPlanningWithPriority WITH an octomap (grid planner, A*, line-of-sight goal) ---------------
struct GoalGridDev {
    int dim[3];                    // cells per axis (src/grid_based_planner.cpp:70-90)
    int cells;                     // dim[0] * dim[1] * dim[2]
    size_t cells_pad;              // ... rounded up to 16
    double gmin[3], res;           // grid_min, grid/resolution
    const float* axis_pts;         // gridVectorToPoint3D per axis: [dim0] x, [dim1] y, [dim2] z
    const uint8_t* static_occ;     // [distinct radii][cells_pad]: kCellOccupied where getDistance(cell point) < radius + grid/margin
    int bkt_seq[kAstarMaxLevels];  // bucket counts of libstdc++'s unordered_map growth (astar_core.cuh)
    int bcap;                      // bucket capacity of one row container
    unsigned magic_a, magic_w, bkt_magic[kAstarMaxLevels];   // multiply-high reciprocals of dim[2], dim[1] and the bucket counts (astar_magic)
};

// ---- k_agent_plan: LSC construction + SFC window + trajectory QP of one agent per thread block ------------------
struct DistMapDev {
    int size[3];                   // cells per axis
    int off[3];                    // signed key of cell 0
    uint8_t* sqdist;               // [x][y][z] squared cell distance clamped at max_sq
    int max_sq;
    int* sat;                      // [n_tables][(sx+1)][(sy+1)][(sz+1)] inclusive-exclusive prefix sums of "blocked"
    int n_tables;
};

struct PlanLaunch {
    int n_agents, n_pad;
    int n_blocks;                  // blocks of this launch = agents this engine plans in this step
    int threads;                   // 256 (two blocks per SM) or 128 (four)
    long long sfc_wait_cycles;     // how long an early block waits for k_sfc_step's box before growing it itself
    // block i plans agent order[order_first + i * order_stride] (global id), or agent_base + i when order == null.
    // The order is longest-processing-time-first over the agents of the job (k_qp_order); one GPU takes every entry, G
    // ranks deal the entries out round-robin (rank r: order_first = r, order_stride = G) so that every rank gets the
    // same mix of expensive and cheap agents.
    const int* order;
    int order_stride, order_first, agent_base, agent_stride;   // order == null: block i plans agent_base + i * agent_stride
    int* block_of;                 // null, or [N]: block_of[agent] = i (lscgpu_get_lsc finds the agent's row-store row)
    // predictions / culling data (k_predict)
    const float* pred;             // [N][90]
    const float* predT;            // [90][n_pad]
    const float* predZs;           // [30][n_pad]
    const AgentConstDev* consts;
    const float2* rdw;             // [N] (radius, downwash * radius) as float, for the culling pass
    const QpTablesDev* T;
    const double* state9;          // [N][9]
    const double* goal3;           // [N][3]
    const int* ts;                 // [N]
    const float4* sphere;          // [5][n_pad]
    const float4* tsphere;         // [n_pad]
    const float* reach;            // [N][5]
    // row store: slots [0, row_cap) of every agent live in shared memory, the rest (everything when mirror_rows) in
    // the global overflow arrays [n_blocks][P_pad] (row i belongs to block i)
    int row_cap, P_pad, mirror_rows;
    RowRec* rows; int* kept; int* kept_count; double* safe;
    // SFC (world_use_octomap): one warp of the block grows the agent's new box while the others build the LSCs
    int use_sfc;
    DistMapDev dm;
    double res;
    // The step's new SFC box of every agent is grown by k_sfc_step, launched beside k_predict; the block reads it when
    // the agent's flag carries this step's epoch, and grows the box itself when it is not there by the time the LSC
    // rows are done (nobody ever waits for another kernel's progress).
    const int* epoch;              // device step counter (k_commit increments it)
    const int* sfc_ready;          // [N] epoch of the step sfc_box_g / sfc_ok_g belong to
    const float* sfc_box_g;        // [N][6]
    const int* sfc_ok_g;           // [N]
    const lscgpu_agent_in* in;     // [N]
    const float* boxes;            // [N][5][6] persistent windows BEFORE this step (k_commit applies the step's new box)
    const int* init_sfc;           // [N] flag_initialize_sfc
    const int* flags;              // [N] k_predict's flag bits; the SFC warp adds its own in the result record
    float wmin[3], wmax[3];
    int max_iter;
    // outputs
    GatherSlot* out;               // block i writes slot out[out_base + i] (gather buffer: rank-major, carries agent_id)
    PeerExchange px;               // px.peers != null: the block also stores its slot into every peer's buffer
    lscgpu_agent_out* const* host_out;   // null, or device word holding the caller's result array when that is pinned host memory
                                   // mapped into the device (else null): the block stores its record straight into
                                   // host_out[agent] — the device-to-host transfer overlaps the planning of the other agents
    const unsigned short* act_prev;    // null (cold starts), or [N][kActSlots]: rows active at every agent's previous solve
    int out_base;
    const float* prev_traj;        // [N][90]  (kept when the QP fails)
    double* last_cost;             // [N]
    const int* goal_kind;          // [N] or null
    StepCounters* counters;
    int* kept_step;                // sum of the kept pairs of this launch (k_commit hands it to the host: block-size choice)
    // disturbance branch (src/traj_planner.cpp:866-878,1047-1061, src/traj_optimizer.cpp:317-326,383-390,455-457)
    const int* any_reset;          // null: no slack kernel; else device word, != 0 once any agent was ever reset (k_predict)
    const unsigned char* reset_ever;   // [N] sticky: the agent's state was reset at some step
    double slack_w;                // opt/slack_collision_weight
    int row_cap_slack;             // shared-memory row slots of the slack instantiation
    long long* dbg;                // null, or [n_blocks][10] section cycle counts (LSCGPU_QP_DEBUG)
};
void launch_agent_plan(const PlanLaunch& L, cudaStream_t s);

struct GoalAstarLaunch {
    GoalLaunch g;                  // the inputs / outputs k_goal_plan has
    int n;                         // agents this engine plans in this step, mapped to agents as PlanLaunch does
    const int* order; int order_stride, order_first, agent_base, agent_stride;
    DistMapDev dm; double world_res;
    GoalGridDev grid;
    int n_blocks;                  // warps = scratch sets
    // per-warp scratch: path cells [n_blocks][cells_pad]; and, for grids whose search state does not fit shared memory,
    // cell bytes / g / list links [n_blocks][cells_pad] and row buckets [n_blocks][dim0][bcap]
    int* path;
    uint8_t* cell; int* gcost; int* next; int* bkt; int* bstamp;
    unsigned long long* expansions;    // null, or the step's A* expansion counter
    int* next_agent;               // device counter, zero at launch: the next entry of the schedule to plan
};
void launch_goal_astar(const GoalAstarLaunch& L, cudaStream_t s);
size_t goal_astar_shared_bytes(const GoalGridDev& g);      // 0: the grid's search state does not fit one SM's shared memory
cudaError_t configure_goal_astar(const GoalGridDev& g);    // once per grid, before the first launch
void launch_goal_static_grid(const GoalGridDev& g, const DistMapDev& dm, double world_res, const double* radii_dev, int n_radii,
                             float grid_margin, uint8_t* out, cudaStream_t s);
size_t agent_plan_smem_bytes(int row_cap, int threads);
size_t agent_plan_slack_smem_bytes(int row_cap, int threads);
cudaError_t configure_agent_plan();     // once per device, before the first launch

// one agent's LSCs recomputed into CollisionConstraints layout (debug / parity)
void launch_lsc_capture(int n_agents, int agent, const float* pred, const AgentConstDev* consts, float* normals,
                        double* d, cudaStream_t s);
void launch_gjk_batch(int n, const double* hulls, double* v, int* iters, cudaStream_t s);
// LSC arrays of the reference container -> row store (operator-level QP entry)
void launch_rows_from_lsc(int n_problems, const int* obs_offset, int total_obs, const float* lsc_normal,
                          const float* lsc_point, const double* lsc_d, RowRec* rows, int* kept,
                          int* kept_count, double* safe, cudaStream_t s, const unsigned char* obs_slack = nullptr);

void launch_terminal_segments(int n, const double* state9, const double* goal3, const int* agent_index,
                              const AgentConstDev* consts, double dt, int* ts_out, cudaStream_t s);

// ---- k_qp_batch: TrajOptimizer::solve for independent problems (operator-level entry) ---------------------
struct QpBatchLaunch {
    int n_problems;
    const QpTablesDev* T;
    const AgentConstDev* consts;
    const int* agent_index;        // [n_problems] which created agent's limits to use
    const double* state9;          // [n_problems][9]
    const double* goal3;           // [n_problems][3]
    const int* ts;                 // [n_problems]
    const float* boxes;            // [n_problems][5][6] or null (no SFC rows)
    float wmin[3], wmax[3];
    RowRec* rows;                  // pairs of problem b at 5 * obs_offset[b] (k_rows_from_lsc)
    const int* obs_offset;         // [n_problems + 1]
    int* kept; const int* kept_count;
    double* safe;
    int max_iter;
    double* x_out;                 // [n_problems][90]
    double* cost_out; int* status_out; int* iters_out;
    // slack variables (src/traj_optimizer.cpp:317-326,383-390,455-457): the slack instantiation of the kernel
    int slack; double slack_w;
    double* eps_out;               // null or [total_obs][5], zero-filled by the caller
};
void launch_qp_batch(const QpBatchLaunch& L, cudaStream_t s);
// order[0..n) = agents a0 .. a0+n-1 (global ids), most expensive solve of the previous step first; deterministic
// (every rank of a job computes the same permutation from its replica of the result records)
void launch_qp_order(int n, int a0, const lscgpu_agent_out* res, int* order, cudaStream_t s);

// commit: for every filled slot of the gather buffer (agent_id >= 0) the record goes to res[agent_id], the new
// trajectory becomes traj_curr, the advanced state becomes the next resident input
// (and, with an octomap, the agent's SFC window takes the step's new box, src/traj_planner.cpp:1451-1491)
void launch_commit(int n_slots, const GatherSlot* gather, lscgpu_agent_out* res, unsigned short* act_prev, float* prev_traj,
                   lscgpu_agent_in* in, double* last_cost, float* boxes /* null: no octomap */, int* init_sfc, int* epoch,
                   int* kept_step, volatile int* kept_host /* mapped host word or null */, PeerExchange px, int* done_count,
                   volatile int* err_host /* mapped host word: 1 = a peer's records did not arrive in time */, cudaStream_t s);

// safety audit of the planned step (src/multi_sync_simulator.cpp:446-475)
void launch_safety_audit(int n_agents, const float* traj, const AgentConstDev* consts, double dt, int n_samples,
                         double record_time_step, float* sample_pos /*[n_samples][N][3]*/, double* ratio, int* closest,
                         cudaStream_t s);

// ---- distance field / SFC ---------------------------------------------------------------------------------
void launch_edt_build(const int32_t* keys_dev, int n_keys, DistMapDev dm, const int* thresholds_dev, int n_tables,
                      uint8_t* scratch_a, uint8_t* scratch_b, cudaStream_t s);

// k_sfc_step: the step's new SFC box (generateFeasibleSFC) of the agents this engine plans, one warp per agent, in the
// same scheduling order as k_agent_plan's blocks
struct SfcStepLaunch {
    int n;                         // agents
    const int* order; int order_stride, order_first, agent_base, agent_stride;   // as PlanLaunch
    DistMapDev dm;
    double res;
    float wmin[3], wmax[3];
    const lscgpu_agent_in* in; const float* prev_traj; const AgentConstDev* consts; const int* init_sfc;
    int planner_seq; double reset_threshold;       // a reset in this step re-arms the corridor (k_predict runs beside this kernel)
    const int* epoch;
    float* sfc_box_g; int* sfc_ok_g; int* sfc_ready;
    const double* goal3;           // null: the goal is lscgpu_agent_in::goal; else [N][3] the goals goal planning chose (goal_mode 1)
};
void launch_sfc_step(const SfcStepLaunch& L, cudaStream_t s);

struct SfcLaunch {
    int n;                         // seeds
    DistMapDev dm;
    double res;
    float wmin[3], wmax[3];
    const float* point; const float* goal; const int* sat_index; float* box_out; int* ok_out;
};
void launch_sfc_expand(const SfcLaunch& L, cudaStream_t s);

}  // namespace lscgpu
