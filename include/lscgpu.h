/*
 * lscgpu.h — C ABI of the B200 replanning engine (liblscgpu.so).
 *
 * One engine handle plans ALL agents of one synchronous replanning step on one GPU: linear safe
 * corridor (LSC) construction against every neighbour's predicted Bezier hull, safe flight corridor
 * (SFC) box expansion against the octomap distance field, and the Bernstein trajectory QP.
 * Plain pointers and sizes only; no exceptions cross this boundary; every entry point returns
 * LSCGPU_OK (0) or a negative error code and lscgpu_last_error() describes the failure.
 * There is no CPU fallback: every compute entry point fails with LSCGPU_ERR_CUDA when no sm_100
 * device is usable.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * tree qwerty35/lsc_planner @ f4d4caf).
 *
 * Data conventions (same as the reference):
 *   trajectory  float[M=5][n+1=6][3]      traj_t, include/sp_const.hpp:17 (segment, control point, xyz)
 *   QP variable order  k*30 + m*6 + i     src/traj_optimizer.cpp:72-95,277 (axis-major)
 *   status      PlanningReport values      include/sp_const.hpp:80-87 (SUCCESS = 5, QPFAILED = 3)
 */
#ifndef LSCGPU_H
#define LSCGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSCGPU_M 5          /* segments (horizon / dt), launch/simulation.launch:60-64 */
#define LSCGPU_NCP 6        /* control points per segment (n = 5) */
#define LSCGPU_TRAJ_FLOATS 90

enum {
    LSCGPU_OK = 0,
    LSCGPU_ERR_ARG = -1,      /* bad argument / unsupported parameter (reference: std::invalid_argument) */
    LSCGPU_ERR_CUDA = -2,     /* CUDA runtime failure or no usable device */
    LSCGPU_ERR_IO = -3,       /* octomap file unreadable / malformed */
    LSCGPU_ERR_NCCL = -4,
    LSCGPU_ERR_STATE = -5     /* call order violated (e.g. octomap required but not uploaded) */
};

/* PlanningReport, include/sp_const.hpp:80-87 */
enum {
    LSCGPU_REPORT_INITIALIZED = 0,
    LSCGPU_REPORT_INITTRAJGENERATIONFAILED = 1,
    LSCGPU_REPORT_CONSTRAINTGENERATIONFAILED = 2,
    LSCGPU_REPORT_QPFAILED = 3,
    LSCGPU_REPORT_WAITFORROSMSG = 4,
    LSCGPU_REPORT_SUCCESS = 5
};

/* QP solver verdict (finer than the reference, which only knows "IloException") */
enum { LSCGPU_QP_OK = 0, LSCGPU_QP_INFEASIBLE = 1, LSCGPU_QP_MAXITER = 2 };

/* per-agent flag bits of one step */
enum {
    LSCGPU_FLAG_SLACK_NEEDED = 1,      /* |initial_traj start - state| > reset_threshold in this step: the agent's prediction
                                          and initial trajectory were collapsed to its position, its corridor re-armed and
                                          the agent entered everybody's obs_slack_indices for good
                                          (src/traj_planner.cpp:866-878,1047-1061) */
    LSCGPU_FLAG_SFC_SEED_BLOCKED = 2,  /* expandBoxFromPoint would throw (include/corridor_constructor.hpp:35-38) */
    LSCGPU_FLAG_SLACK_MODE = 4,        /* the QP carried slack variables (some agent of the swarm was reset before or in this
                                          step; src/traj_optimizer.cpp:317-326,383-390,455-457) */
    LSCGPU_FLAG_SLACK_USED = 8,        /* ... and at least one of them is < 0 at the solution */
    LSCGPU_FLAG_SLACK_OVERFLOW = 16,   /* the QP needed more slack variables than the kernel holds per agent (25): reported
                                          as LSCGPU_QP_MAXITER, the previous trajectory is kept */
    LSCGPU_FLAG_WARM_START = 32,       /* the QP started from the bound / dynamic-limit rows active at the agent's previous
                                          solve (accepted only when all their multipliers are >= 0; the minimiser is the same) */
    LSCGPU_FLAG_IN_BAND = 64           /* at the returned solution some row is violated by more than 1e-9 (and, by construction,
                                          at most the feasibility tolerance 1e-6, CPLEX's EpRHS): like CPLEX's, the solution is
                                          then defined only up to such rows and may depend on the pivoting order at that level */
};

/* The subset of Param (include/param.hpp, defaults src/param.cpp:4-107) that the hot path reads. */
typedef struct lscgpu_params {
    double dt;                     /* traj/dt              (0.2)  == multisim/time_step */
    double control_input_weight;   /* opt/control_input_weight (0.01) */
    double terminal_weight;        /* opt/terminal_weight  (1.0) */
    double world_resolution;       /* world/resolution     (0.1) */
    double reset_threshold;        /* multisim/reset_threshold (0.15) */
    int world_use_octomap;         /* world/use_octomap: adds the SFC rows (src/traj_optimizer.cpp:409-434) */
    float world_min[3];            /* Mission::world_min / world_max (src/mission.cpp:60-75) */
    float world_max[3];
    int M, n, phi, dim;            /* must be 5, 5, 3, 3 (traj/horizon 1.0, traj/n, traj/phi, world/dimension) */
    /* Goal planning, the step before the path (src/traj_planner.cpp:477-608). 0 = GoalMode::STATIC: lscgpu_agent_in::goal
     * is the current goal. 1 = GoalMode::PRIORBASED (the reference's default) on the GPU: lscgpu_agent_in::goal is the
     * DESIRED goal and the engine derives the current goal: priority rule, retreat point, and — with an octomap — the grid
     * planner (occupancy grid of grid_resolution cells, A* with the reference's tie-breaking, line-of-sight goal by ray
     * casting in the distance field: src/grid_based_planner.cpp:53-433, src/Astar-3D), clip to goal_radius. Without an
     * octomap the line-of-sight goal does not depend on the A* path and only the clip remains. */
    int goal_mode;
    double goal_threshold;             /* plan/goal_threshold (0.1) */
    double goal_radius;                /* plan/goal_radius (2.0) */
    double priority_dist_threshold;    /* plan/priority_dist_threshold (0.4) */
    double grid_resolution;            /* grid/resolution (launch: 0.25; 0 = that): goal_mode 1 with an octomap */
    double grid_margin;                /* grid/margin (0.1): a grid cell is occupied when getDistance < radius + margin */
} lscgpu_params;

/* Constant part of Agent (include/sp_const.hpp:153-165). */
typedef struct lscgpu_agent_const {
    double radius, downwash, nominal_velocity;
    double max_vel[3], max_acc[3];
} lscgpu_agent_const;

/* What MultiSyncSimulator::update() pushes into a TrajPlanner before plan()
 * (src/multi_sync_simulator.cpp:190-318: setCurrentState + the goal chosen by goal planning). */
typedef struct lscgpu_agent_in {
    float position[3], velocity[3], acceleration[3];   /* Agent::current_state */
    float goal[3];                                      /* Agent::current_goal_position */
} lscgpu_agent_in;

/* What the simulator pulls out after plan(): getTraj(), getQPCost(), getPlanningReport(),
 * getFutureStateMsg(dt) (include/traj_planner.hpp:76-101). */
typedef struct lscgpu_agent_out {
    float traj[LSCGPU_M][LSCGPU_NCP][3];
    float next_position[3], next_velocity[3], next_acceleration[3];   /* trajectory evaluated at t = dt */
    int32_t agent_id;           /* the agent this record belongs to (records travel between GPUs in scheduling order) */
    double qp_cost;             /* getObjValue incl. the constant of the terminal cost */
    int32_t report;             /* PlanningReport; the reference reports SUCCESS even when the QP failed and keeps the
                                   previous trajectory (src/traj_planner.cpp:1553-1584) — so does this field */
    int32_t qp_status;          /* LSCGPU_QP_* */
    int32_t qp_iterations;      /* active-set iterations */
    int32_t qp_active;          /* active inequality rows at the solution */
    int32_t flags;              /* LSCGPU_FLAG_* */
    int32_t terminal_segments;  /* getTerminalSegments (src/traj_optimizer.cpp:541-548) */
    int32_t qp_sweeps;          /* verification sweeps over all kept LSC pairs */
    int32_t qp_kcycles;         /* SM clock cycles / 1024 this agent's plan took (corridors + QP: its planning time) */
    int32_t qp_price_kcycles;   /* ... of which: pricing the rows (the rest is the factorisation update) */
    int32_t lsc_pairs_kept;     /* (neighbour, segment) pairs that survived the exact culling test */
    float current_goal[3];      /* Agent::current_goal_position the QP used (getCurrentGoalPosition): the input goal, or the
                                   goal chosen by goal planning when lscgpu_params::goal_mode == 1 */
    int32_t goal_kind;          /* goal planning: 0 line-of-sight goal, 1 retreat from the closest higher-priority agent */
    float sfc_box[6];           /* the SFC box grown in this step (min xyz, max xyz; CorridorConstructor::expandBoxFromPoint):
                                   the last box of the agent's window, or all five at the first step. Zero without octomap
                                   or when the seed was blocked (LSCGPU_FLAG_SFC_SEED_BLOCKED) */
    int32_t lsc_kcycles;        /* ... of qp_kcycles: corridor construction (LSC rows + SFC box) before the QP started */
    int32_t sfc_in_block;       /* 1: the planning block grew the SFC box itself (k_sfc_step's result was not there in time) */
} lscgpu_agent_out;

typedef struct lscgpu_engine lscgpu_engine;

const char* lscgpu_last_error(void);
int lscgpu_version(void);

/* ---- lifetime ---------------------------------------------------------------------------------
 * Replaces: N x TrajPlanner ctor + TrajOptimizer ctor (src/traj_planner.cpp:4-72, src/traj_optimizer.cpp:4-25:
 * Q_base / Aeq_base built once) inside MultiSyncSimulator's ctor (src/multi_sync_simulator.cpp:4-81). */
int lscgpu_create(const lscgpu_params* params, int n_agents, const lscgpu_agent_const* agents, int device,
                  lscgpu_engine** out);
void lscgpu_destroy(lscgpu_engine* e);

/* Replaces MultiSyncSimulator::setOctomap (src/multi_sync_simulator.cpp:153-167): octomap::OcTree::readBinary +
 * DynamicEDTOctomap(1.0, tree, world_min, world_max, false) + update(). The distance field is built on the GPU. */
int lscgpu_set_octomap_file(lscgpu_engine* e, const char* bt_path);
/* Same, from finest-level occupied voxel keys (signed, key - 32768), int32[n][3]. n == 0: empty map. */
int lscgpu_set_octomap_voxels(lscgpu_engine* e, const int32_t* keys, int n);
/* Distance-map geometry and content (DynamicEDTOctomap::getDistance = sqrt(sqdist) * res; -1 outside). */
int lscgpu_get_distmap_info(lscgpu_engine* e, int32_t size[3], int32_t offset[3], int64_t* n_occupied);
int lscgpu_get_distmap_sqdist(lscgpu_engine* e, uint8_t* out /* size[0]*size[1]*size[2], x-major */);

/* ---- multi-GPU ---------------------------------------------------------------------------------
 * No counterpart in the reference (single process; agents planned sequentially, src/multi_sync_simulator.cpp:320-337).
 * Every engine of a job holds a replica of every agent's planner state (previous trajectory, SFC window, last cost).
 *
 * lscgpu_nccl_init: G engines (one per GPU / rank) plan one swarm. Each step all ranks derive the same
 * longest-processing-time-first order of the agents from their replicas, rank r plans entries r, r+G, r+2G, ... of it,
 * one in-place ncclAllGather of the result records (issued by the library on the engine's stream) completes every
 * replica, and every rank returns every agent.
 *
 * lscgpu_set_shard: no communicator; this engine plans agents [a0, a1) only and commits only those. Records, previous
 * trajectories and states of the other agents keep whatever the caller last uploaded: a caller that shards by hand must
 * gather the results itself and refresh every replica (lscgpu_set_prev_traj + the next lscgpu_replan_batch's inputs)
 * before the next step. */
int lscgpu_set_shard(lscgpu_engine* e, int a0, int a1);
int lscgpu_nccl_unique_id(uint8_t id_out[128]);
int lscgpu_nccl_init(lscgpu_engine* e, const uint8_t id[128], int rank, int n_ranks);
/* Direct exchange over NVLink peer memory (optional, after lscgpu_nccl_init; one process per GPU on one node): instead of the
 * all-gather, every planning block stores its finished record straight into every peer's buffer the moment the agent is
 * planned (the transfer overlaps the planning of the other agents) and bumps an arrival counter there; the commit kernel of
 * each rank waits for the counters. lscgpu_p2p_export returns the CUDA IPC handle of this rank's exchange buffer; gather the
 * handles of all ranks in rank order (any out-of-band channel) and pass them to lscgpu_p2p_attach on every rank. Results are
 * identical to the all-gather path. A peer whose records do not arrive within 2 s makes the next synchronising call fail
 * with LSCGPU_ERR_NCCL instead of hanging the device. */
int lscgpu_p2p_export(lscgpu_engine* e, uint8_t handle_out[64]);
int lscgpu_p2p_attach(lscgpu_engine* e, const uint8_t* handles /* [n_ranks][64] */);

/* ---- the replanning step ------------------------------------------------------------------------
 * Replaces the loop `for qi: agents[qi]->plan(sim_current_time)` of MultiSyncSimulator::plan()
 * (src/multi_sync_simulator.cpp:320-337), i.e. per agent TrajPlanner::planLSC (src/traj_planner.cpp:389-425):
 * obstaclePrediction/initialTrajPlanning (prev-solution variants), generateLSC, generateFeasibleSFC,
 * TrajOptimizer::solve. `in` and `out` are host arrays of n_agents elements (all agents, also on a sharded engine:
 * the step ends with the all-gather, so every rank returns every agent). planner_seq is advanced by one. */
int lscgpu_replan_batch(lscgpu_engine* e, const lscgpu_agent_in* in, lscgpu_agent_out* out);
/* When `out` is pinned host memory (cudaHostAlloc / cudaHostRegister; 16-byte aligned) and one engine plans every agent, the
 * planning blocks store their records straight into it as each agent is done (the device-to-host transfer overlaps the
 * planning of the others); otherwise the records are copied after the step. Same bytes either way.
 *
 * The state hand-over of MultiSyncSimulator::update (src/multi_sync_simulator.cpp:203: the next current_state is the planned
 * trajectory at t = dt): in[a].position / velocity / acceleration = out[a].next_*, goals untouched. Host-side convenience. */
int lscgpu_advance_inputs(const lscgpu_agent_out* out, lscgpu_agent_in* in, int n_agents);

/* Device-resident variant used for kernel-level timing: the inputs are the engine's own advanced states
 * (previous trajectories evaluated at t = dt) and the goals of the last lscgpu_replan_batch / lscgpu_set_goals.
 * Nothing crosses PCIe. lscgpu_fetch copies the results of the last step to the host. */
int lscgpu_set_goals(lscgpu_engine* e, const float* goals /* [n_agents][3] */);
int lscgpu_set_states(lscgpu_engine* e, const float* pos, const float* vel, const float* acc /* [n_agents][3] each */);
int lscgpu_replan_resident(lscgpu_engine* e);   /* enqueue only: returns without waiting for the GPU */
int lscgpu_synchronize(lscgpu_engine* e);       /* wait for all enqueued steps; refreshes lscgpu_get_step_stats */
int lscgpu_fetch(lscgpu_engine* e, lscgpu_agent_out* out);

/* TrajPlanner::reset / fresh planners: planner_seq = 0, trajectories zero, flag_initialize_sfc = true. */
int lscgpu_reset(lscgpu_engine* e);
/* Restore persistent planner state (teacher-forced parity runs, checkpoint restore):
 * traj_curr of every agent + planner_seq; SFC windows float[n_agents][5][6] (min xyz, max xyz) + flag_initialize_sfc. */
int lscgpu_set_prev_traj(lscgpu_engine* e, const float* traj /* [n_agents][90] */, int planner_seq);
int lscgpu_set_sfc(lscgpu_engine* e, const float* boxes, const int32_t* flag_initialize_sfc);
int lscgpu_get_sfc(lscgpu_engine* e, float* boxes, int32_t* flag_initialize_sfc);
int lscgpu_get_planner_seq(lscgpu_engine* e);

/* ---- disturbance handling ------------------------------------------------------------------------
 * Replaces obstaclePredictionCheck / initialTrajPlanningCheck (src/traj_planner.cpp:866-878,1047-1061) and the slack
 * variables of TrajOptimizer::populatebyrow (src/traj_optimizer.cpp:317-326,383-390,455-457). The caller reports an
 * externally observed pose the way MultiSyncSimulator::update does (src/multi_sync_simulator.cpp:229-246): the agent's
 * lscgpu_agent_in carries the observed position with zero velocity and acceleration. A step whose position is farther
 * than reset_threshold from the start of the agent's shifted previous trajectory collapses its prediction and initial
 * trajectory to that position, re-arms its corridor and puts the agent into every planner's obs_slack_indices (and every
 * obstacle into its own) for the rest of the mission, as the reference does. From then on every QP of the swarm carries the
 * slack variables eps_{oi,m} <= 0 with cost slack_collision_weight ((M - m) / M) eps^2 (default 1, src/param.cpp:75;
 * launch/simulation.launch sets 100000). qp_cost includes their cost. lscgpu_reset clears the sets.
 * lscgpu_get/set_reset_state: the sticky per-agent "was ever reset" bytes (checkpoint / teacher-forced runs). */
int lscgpu_set_slack_collision_weight(lscgpu_engine* e, double w);
int lscgpu_get_reset_state(lscgpu_engine* e, uint8_t* reset_ever /* [n_agents] */);
int lscgpu_set_reset_state(lscgpu_engine* e, const uint8_t* reset_ever);

/* Constraints of the last step, as CollisionConstraints::getLSC would return them
 * (src/collision_constraints.cpp:362-364): for local agent `agent` and every other agent in id order,
 * normals float[n_agents-1][5][3], margins d double[n_agents-1][5][6]. */
int lscgpu_get_lsc(lscgpu_engine* e, int agent, float* normals, double* d);
/* Same, reading what the planning kernel itself built: after lscgpu_set_capture_rows(e, 1) every step also mirrors the
 * agents' LSC rows (normally shared-memory resident) to global memory; the kept (neighbour, segment) pairs — those the
 * exact culling test could not prove inactive — are decoded from that row store (kept[n_agents-1][5] = 1), the culled
 * ones are recomputed. The agent must have been planned by this engine in the last step. */
int lscgpu_set_capture_rows(lscgpu_engine* e, int on);
int lscgpu_get_lsc_ex(lscgpu_engine* e, int agent, float* normals, double* d, uint8_t* kept);
/* The QP the engine solved for `agent` in the last step as a CPLEX LP file: replaces cplex.exportModel(".../log/QPmodel.lp")
 * (src/traj_optimizer.cpp:62-69,99-101,146-149). Variables, rows and their order are populatebyrow's; with obs_slack_indices
 * not empty the epsilon_slack_<oi>_<m> variables are included. lscgpu_set_lp_dump_dir: every lscgpu_replan_batch writes
 * <dir>/QPmodel_agent<id>_seq<planner_seq>.lp for each agent whose QP failed, as the reference does on an IloException
 * (NULL or "" switches it off). */
int lscgpu_dump_qp_lp(lscgpu_engine* e, int agent, const char* path);
int lscgpu_set_lp_dump_dir(lscgpu_engine* e, const char* dir);
/* initial_traj of every agent of the last step (= the prediction its neighbours used), float[n_agents][90]. */
int lscgpu_get_initial_traj(lscgpu_engine* e, float* out);

/* ---- operator-level entry points (unit parity with the reference's classes) -------------------- */

/* TrajOptimizer::solve (src/traj_optimizer.cpp:31-154) for a batch of independent problems. Inputs per problem b:
 *   state[b][9]   pos xyz, vel xyz, acc xyz  (Agent::current_state)           goal[b][3]  current_goal_position
 *   sfc[b][5][6]  SFC boxes (box_min xyz, box_max xyz) or NULL when !world_use_octomap
 *   obs_offset[b], obs_offset[b+1]  range of this problem's obstacles in the LSC arrays; for obstacle o:
 *   lsc_normal[o][5][3] (LSC::normal_vector), lsc_point[o][5][6][3] (LSC::obs_control_point), lsc_d[o][5][6] (LSC::d)
 *   agent_index[b] selects max_vel/max_acc/nominal_velocity of a created agent.
 * Outputs: x[b][90] in the reference's variable order, cost[b], status[b] (LSCGPU_QP_*), iterations[b]. */
int lscgpu_qp_solve_batch(lscgpu_engine* e, int n_problems, const int32_t* agent_index, const double* state,
                          const double* goal, const float* sfc, const int32_t* obs_offset, const float* lsc_normal,
                          const float* lsc_point, const double* lsc_d, double* x, double* cost, int32_t* status,
                          int32_t* iterations);

/* Same with CollisionConstraints::getSlackIndices (src/traj_optimizer.cpp:268,317-326,383-390,455-457): obs_slack[o] != 0
 * marks the obstacles of obs_slack_indices; their LSC rows are relaxed by eps_{o,m} <= 0 with cost
 * slack_collision_weight ((M - m) / M) eps^2 (lscgpu_set_slack_collision_weight). eps (optional) receives
 * eps[o][5] for every obstacle (0 where the obstacle is outside the set). cost includes the slack cost. With no
 * obstacle marked this is lscgpu_qp_solve_batch. */
int lscgpu_qp_solve_batch_slack(lscgpu_engine* e, int n_problems, const int32_t* agent_index, const double* state,
                                const double* goal, const float* sfc, const int32_t* obs_offset, const float* lsc_normal,
                                const float* lsc_point, const double* lsc_d, const uint8_t* obs_slack, double* x,
                                double* cost, int32_t* status, int32_t* iterations, double* eps);

/* closestPointsBetweenPointAndConvexHull(origin, hull) (include/geometry.hpp:364-394) -> gjk()
 * (src/openGJK/openGJK.cpp:674-780) for n_hulls hulls of 6 points: hulls double[n][6][3] -> v double[n][3]
 * (closest point of the hull to the origin), iterations int32[n]. */
int lscgpu_gjk_batch(lscgpu_engine* e, int n_hulls, const double* hulls, double* v, int32_t* iterations);

/* CorridorConstructor::expandBoxFromPoint (include/corridor_constructor.hpp:18-44) for n seeds:
 * point/goal float[n][3], radius[n] -> box float[n][6], ok int32[n] (0: seed blocked, reference throws). */
int lscgpu_sfc_expand_batch(lscgpu_engine* e, int n, const float* point, const float* goal, const double* radius,
                            float* box, int32_t* ok);

/* Replaces the O(N^2) minimum-distance audit of MultiSyncSimulator::savePlanningResult (src/multi_sync_simulator.cpp:
 * 446-475): for every recorded sub-time t = 0, record_time_step, ... < time_step of the step just planned and every agent,
 * the smallest downwash-scaled distance to another agent (include/util.hpp:225-229) over the sum of the two radii, from
 * the trajectories resident on the device. ratio[a] = minimum over the sub-times and the other agents, closest[a] = that
 * agent (first minimum in time, then id order). safety_ratio_agent = min(ratio), is_collided = any ratio < 1. */
int lscgpu_safety_audit(lscgpu_engine* e, double record_time_step, double time_step, double* ratio, int32_t* closest);

/* ---- instrumentation -------------------------------------------------------------------------- */
typedef struct lscgpu_step_stats {
    /* Totals over the steps enqueued since the previous lscgpu_synchronize / lscgpu_replan_batch return. */
    int32_t steps;
    float ms_total;        /* device time, first kernel of the first step to last kernel of the last (CUDA events on
                              the engine stream) */
    /* per-kernel sums, 0 unless profiling is on: k_predict (+ k_goal_plan), k_agent_plan (LSC + SFC + QP of every agent;
     * lscgpu_agent_out::lsc_kcycles / qp_kcycles split it per agent), the all-gather, k_commit */
    float ms_predict, ms_plan, ms_sfc /* k_sfc_step: runs beside the others unless profiling */, ms_reserved1_, ms_exchange, ms_commit;
    int32_t kernel_launches;        /* kernels of this library launched */
    int64_t lsc_pairs;              /* (agent, neighbour, segment) pairs of the swarm: (N-1) * 5 per local agent */
    int64_t lsc_pairs_kept;         /* pairs that survived the exact culling test, i.e. GJK hull tests actually run */
    int64_t gjk_iterations;         /* GJK outer iterations */
    int64_t qp_rows_priced;         /* inequality rows evaluated by the QP kernel (all local agents) */
    int64_t qp_iterations;          /* active-set iterations (all local agents) */
    int64_t qp_full_passes;         /* verification sweeps over the complete LSC row set (all local agents) */
    float ms_steps;                 /* sum over the steps of (last kernel end - first kernel start): like ms_total without
                                       whatever the caller enqueued on the stream between steps */
    float reserved_;
    /* QP warm starts (all local agents): solves that had candidates from the agent's previous solve, solves whose
     * candidates were accepted as the starting point, and the rows those put into the working set without an iteration */
    int64_t qp_warm_tried, qp_warm_accepted, qp_warm_rows;
    int64_t astar_expansions;       /* goal planning with an octomap (goal_mode 1): nodes the A* searches closed */
} lscgpu_step_stats;
int lscgpu_get_step_stats(lscgpu_engine* e, lscgpu_step_stats* out);
/* enable per-kernel event timing (event records between the kernels of every step; off by default) */
int lscgpu_set_profiling(lscgpu_engine* e, int on);
/* Roofline denominators of `device`, measured now (FMA-chain microbenchmark, ~50 ms): dense FP32 / FP64 FMA TFLOP/s.
 * MEASURED_PEAKS.json carries only HBM and bf16 tensor figures; this path computes in FP64 on CUDA cores. */
int lscgpu_measure_fma_peaks(int device, double* fp32_tflops, double* fp64_tflops);
/* Dependent-issue latencies in SM cycles per operation (one warp, one chain): DFMA, FFMA, shared-memory load, shuffle+DADD
 * (one round of a double warp reduction), rsqrt(double), double division, redux.sync. They bound the QP's active-set update. */
int lscgpu_measure_latencies(int device, double cycles_out[7]);
/* SM clock of the engine's device in kHz (converts the cycle counters of lscgpu_agent_out into seconds) */
int lscgpu_sm_clock_khz(lscgpu_engine* e);
/* the CUDA stream all work of this engine is enqueued on (cudaStream_t as void*) */
void* lscgpu_stream(lscgpu_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* LSCGPU_H */
