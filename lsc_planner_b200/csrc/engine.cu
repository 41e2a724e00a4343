// C ABI of the replanning engine (include/lscgpu.h): device memory, stream, step sequencing, NCCL exchange.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>      // std::__detail::_Prime_rehash_policy (goal_bucket_sequence)
#include <vector>

#include "kernels.hpp"
#include "lp_dump.hpp"
#include "octomap_bt.hpp"

using namespace lscgpu;

// the ctypes / numpy mirrors in lsc_planner_b200/_capi.py assume these layouts (tests/test_capi_load.py)
static_assert(sizeof(lscgpu_params) == 128 && sizeof(lscgpu_agent_in) == 48 && sizeof(lscgpu_agent_out) == 496 &&
              sizeof(lscgpu_agent_const) == 72, "C-ABI struct layout changed: update _capi.py and the tests");

static thread_local std::string g_error;
static int fail(int code, const std::string& msg) { g_error = msg; return code; }

#define CU(expr)                                                                                      \
    do {                                                                                              \
        cudaError_t err__ = (expr);                                                                   \
        if (err__ != cudaSuccess)                                                                     \
            return fail(LSCGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));      \
    } while (0)

// ---- NCCL, bound at run time (libnccl.so.2: the copy torch already loaded, or the system one) ------------------
namespace {
struct NcclId { char internal[128]; };
typedef struct ncclComm* NcclComm;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
bool load_nccl(std::string& err) {
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
    g_nccl.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) {
        err = "libnccl lacks required symbols"; return false;
    }
    g_nccl.handle = h;
    return true;
}
}  // namespace

// pinned / device scratch that grows on demand and lives as long as the engine (operator-level entries, setters)
namespace {
struct Scratch {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        const size_t grow = std::max(need, (size_t)4096);
        cudaError_t rc = cudaMalloc(&p, grow);
        if (rc == cudaSuccess) bytes = grow;
        return rc;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};
// device buffer that frees itself when the function that made it returns (early error returns included)
struct TempBuf {
    void* p = nullptr;
    ~TempBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};
// carve typed arrays out of one scratch allocation (256-byte aligned)
struct Carver {
    size_t off = 0;
    template <class T> size_t take(size_t count) {
        const size_t at = off;
        off = (off + count * sizeof(T) + 255) & ~(size_t)255;
        return at;
    }
};
}  // namespace

struct lscgpu_engine {
    lscgpu_params prm{};
    int N = 0, n_pad = 0;            // agents; n_pad = row pitch of the transposed prediction table
    int a0 = 0, a1 = 0;              // agents this engine plans when no communicator is attached
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_aux = nullptr;     // k_qp_order (needs only the previous step's records) beside k_predict
    cudaEvent_t ev_fork = nullptr, ev_order = nullptr;
    bool lpt_order = true, qp_debug = false, use_graph = true;
    // Block size of k_agent_plan. 128 threads (4 blocks per SM, 256 rows in shared memory) overlaps four agents'
    // latency-bound QP chains per SM; 256 threads (2 blocks, 1024 rows) is faster when the corridors dominate (crowded
    // swarm, ~1000 kept pairs per agent). 0 = choose per step from the kept pairs of the last completed step. Results do
    // not depend on the choice (the row order is canonical).
    int plan_threads = 0;
    int wide_kept = 500;             // mean kept pairs per agent from which the 256-thread configuration is used
    int n_sm = 148;
    int row_cap_forced = -1;
    int* d_kept_step = nullptr;      // device: kept pairs of the step being planned
    int* h_kept_last = nullptr;      // mapped host word: kept pairs of the last committed step
    int* d_kept_last_map = nullptr;  // its device alias
    long long sfc_wait_cycles = 60000;
    bool mirror_rows = false;        // also write every row to the global store (lscgpu_get_lsc reads production rows)
    long long* d_dbg = nullptr;
    float* d_audit_pos = nullptr; double* d_audit_ratio = nullptr; int* d_audit_closest = nullptr;   // lscgpu_safety_audit scratch
    int audit_samples = 0;
    int* d_order = nullptr;          // [N] agents, most expensive plan of the previous step first
    int* d_block_of = nullptr;       // [N] block (= row-store row) that planned the agent in the last step, -1: not planned here
    int* d_epoch = nullptr;          // device step counter (k_commit): stamps the results of k_sfc_step
    int* d_sfc_ready = nullptr; float* d_sfc_box = nullptr; int* d_sfc_ok = nullptr;   // [N] k_sfc_step -> k_agent_plan
    cudaEvent_t ev_sfc = nullptr;
    int planner_seq = 0;
    // disturbance branch: sticky "was ever reset" per agent + "anybody was" (k_predict), slack weight, slack kernel on/off
    unsigned char* d_reset_ever = nullptr;
    int* d_any_reset = nullptr;
    double slack_w = 1.0;            // opt/slack_collision_weight (src/param.cpp:75)
    bool slack_kernel = true;
    int row_cap_slack = 384;
    bool profiling = false;
    int max_iter = 2000;

    std::vector<lscgpu_agent_const> consts_host;
    std::vector<double> radii;       // distinct radii -> blocked-voxel table index

    // device buffers
    QpTablesDev* d_tables = nullptr;
    AgentConstDev* d_consts = nullptr;
    float2* d_rdw = nullptr;         // [N] (radius, downwash * radius) in float: all the culling pass needs
    lscgpu_agent_in* d_in = nullptr;
    GatherSlot* d_gather = nullptr;        // [n_slots] records (+ active rows) in scheduling order, rank-major: the all-gather buffer
    unsigned short* d_act_prev = nullptr;  // [N][kActSlots] rows active at every agent's previous solve (warm-start candidates)
    bool warm_start = true;
    // direct exchange over peer memory (lscgpu_p2p_export / lscgpu_p2p_attach): this rank's exchange buffer, the peers'
    // buffers as mapped here, the device copy of those pointers
    GatherSlot* d_xchg = nullptr;
    std::vector<void*> peer_ptrs;
    GatherSlot** d_peers = nullptr;
    std::string lp_dump_dir;               // not empty: lscgpu_replan_batch writes the LP of every failed QP there
    bool p2p = false;
    int p2p_base_epoch = 0;
    int* d_commit_done = nullptr;          // k_commit's "last block" counter
    int* h_xchg_err = nullptr; int* d_xchg_err_map = nullptr;      // mapped host word: a peer's records did not arrive in time
    lscgpu_agent_out* d_res = nullptr;     // [N] the same records in agent order (k_commit)
    int n_slots = 0;
    float *d_traj = nullptr, *d_pred = nullptr, *d_predT = nullptr, *d_predZs = nullptr, *d_boxes = nullptr;
    double *d_state9 = nullptr, *d_goal3 = nullptr, *d_last_cost = nullptr;
    int *d_ts = nullptr, *d_flags = nullptr, *d_init_sfc = nullptr, *d_goal_kind = nullptr;
    // overflow / mirror row store: one row of P_pad slots per block of k_agent_plan
    RowRec* d_rows = nullptr;
    int P_pad = 0, n_rows_alloc = 0;
    int *d_kept = nullptr, *d_kept_count = nullptr;
    double* d_safe = nullptr;
    float4* d_sphere = nullptr;      // [5][n_pad]
    float4* d_tsphere = nullptr;     // [n_pad]
    float* d_reach = nullptr;        // [N][5]
    StepCounters* d_counters = nullptr;
    Scratch scratch;                 // operator-level entries and setters
    char* h_stage = nullptr; size_t stage_bytes = 0;     // pinned host staging of lscgpu_qp_solve_batch (inputs | results)
    // map
    bool have_map = false;
    DistMapDev dm{};
    int64_t n_occupied = 0;
    // goal planning with an octomap (k_goal_astar): planning grid, static occupancy per radius, per-warp search scratch
    GoalGridDev goal_grid{};
    bool have_goal_grid = false;
    int goal_blocks = 0;
    float* d_goal_axis = nullptr; uint8_t* d_goal_static = nullptr;
    uint8_t* d_goal_cell = nullptr; int* d_goal_g = nullptr; int* d_goal_next = nullptr; int* d_goal_bkt = nullptr; int* d_goal_bstamp = nullptr; int* d_goal_path = nullptr;
    unsigned long long* d_goal_expansions = nullptr;
    int* d_goal_ticket = nullptr;          // next entry of the schedule k_goal_astar hands out
    // zero-copy results: when lscgpu_replan_batch's `out` is pinned host memory the planning blocks store their records
    // straight into it (device word d_host_out = its device alias, null otherwise); h_host_out: pinned staging of that word
    lscgpu_agent_out** d_host_out = nullptr;
    lscgpu_agent_out** h_host_out = nullptr;
    lscgpu_agent_out* host_out_on_device = nullptr;      // value the device word holds
    lscgpu_agent_out* host_out_alias = nullptr;          // device alias of the current call's `out` (null: not mapped)
    bool zero_copy = true;
    // exchange
    NcclComm comm = nullptr;
    int rank = 0, n_ranks = 1, block = 0;
    // the step as a CUDA graph (steps with planner_seq >= 2 are identical launches)
    cudaGraphExec_t graph[3] = {nullptr, nullptr, nullptr};   // [0] 256-thread, [1] 128-thread, [2] 512-thread configuration
    bool graph_failed = false;
    int graph_launches[3] = {0, 0, 0};
    // instrumentation: steps enqueued since the last synchronize
    struct StepEvents { cudaEvent_t ev[7]; };   // begin, predict|, plan|, exchange|, commit|, sfc[ ]sfc
    std::vector<StepEvents> ev_pool;    // per-kernel events of every pending step (profiling mode)
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> step_ev;   // begin / end of every pending step
    int pending = 0, pending_launches = 0;
    int done_steps = 0, done_launches = 0;      // of the batch the last synchronize completed (stats computed lazily)
    bool stats_fresh = true;
    lscgpu_step_stats stats{};

    int n_plan() const { return comm ? std::max(0, (N - rank + n_ranks - 1) / n_ranks) : a1 - a0; }
    int n_plan_max() const { return comm ? (N + n_ranks - 1) / n_ranks : a1 - a0; }
};

static void drop_graph(lscgpu_engine* e) {
    for (auto& g : e->graph) { if (g) cudaGraphExecDestroy(g); g = nullptr; }
}

static void free_rows(lscgpu_engine* e) {
    cudaFree(e->d_rows); cudaFree(e->d_safe);
    cudaFree(e->d_kept); cudaFree(e->d_kept_count);
    e->d_rows = nullptr; e->d_safe = nullptr;
    e->d_kept = nullptr; e->d_kept_count = nullptr;
    e->n_rows_alloc = 0;
}

static int alloc_rows(lscgpu_engine* e) {
    drop_graph(e);
    const int n_rows = std::max(e->n_plan_max(), 1);
    if (n_rows <= e->n_rows_alloc) return LSCGPU_OK;
    free_rows(e);
    const int P = kPairsPerObs * std::max(e->N - 1, 1);
    e->P_pad = (P + 31) / 32 * 32;
    CU(cudaMalloc(&e->d_rows, sizeof(RowRec) * (size_t)n_rows * e->P_pad));
    CU(cudaMalloc(&e->d_safe, sizeof(double) * (size_t)n_rows * e->P_pad));
    CU(cudaMalloc(&e->d_kept, sizeof(int) * (size_t)n_rows * e->P_pad));
    CU(cudaMalloc(&e->d_kept_count, sizeof(int) * (size_t)n_rows));
    CU(cudaMemset(e->d_kept_count, 0, sizeof(int) * (size_t)n_rows));
    e->n_rows_alloc = n_rows;
    return LSCGPU_OK;
}

static int alloc_gather(lscgpu_engine* e, int n_slots) {
    if (n_slots <= e->n_slots) return LSCGPU_OK;
    cudaFree(e->d_gather);
    e->d_gather = nullptr; e->n_slots = 0;
    CU(cudaMalloc(&e->d_gather, sizeof(GatherSlot) * (size_t)n_slots));
    e->n_slots = n_slots;
    return LSCGPU_OK;
}

extern "C" const char* lscgpu_last_error(void) { return g_error.c_str(); }
extern "C" int lscgpu_version(void) { return 210; }     // 2.1: lscgpu_params grew (grid_resolution, grid_margin), lscgpu_step_stats grew (astar_expansions)

extern "C" void lscgpu_destroy(lscgpu_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);      // the aux stream is joined into this one inside every step
    drop_graph(e);
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    for (size_t r = 0; r < e->peer_ptrs.size(); r++)
        if ((int)r != e->rank && e->peer_ptrs[r]) cudaIpcCloseMemHandle(e->peer_ptrs[r]);
    if (e->p2p) e->d_gather = nullptr;                      // it pointed into d_xchg
    cudaFree(e->d_xchg); cudaFree(e->d_peers); cudaFree(e->d_commit_done);
    if (e->h_xchg_err) cudaFreeHost(e->h_xchg_err);
    free_rows(e);
    e->scratch.release();
    cudaFree(e->d_order); cudaFree(e->d_block_of);
    cudaFree(e->d_kept_step); if (e->h_kept_last) cudaFreeHost(e->h_kept_last);
    cudaFree(e->d_epoch); cudaFree(e->d_sfc_ready); cudaFree(e->d_sfc_box); cudaFree(e->d_sfc_ok);
    cudaFree(e->d_reset_ever); cudaFree(e->d_any_reset);
    cudaFree(e->d_host_out); if (e->h_host_out) cudaFreeHost(e->h_host_out);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    cudaFree(e->d_goal_axis); cudaFree(e->d_goal_static); cudaFree(e->d_goal_cell); cudaFree(e->d_goal_g); cudaFree(e->d_goal_next);
    cudaFree(e->d_goal_bkt); cudaFree(e->d_goal_bstamp); cudaFree(e->d_goal_path); cudaFree(e->d_goal_expansions); cudaFree(e->d_goal_ticket);
    if (e->ev_sfc) cudaEventDestroy(e->ev_sfc);
    cudaFree(e->d_rdw); cudaFree(e->d_audit_pos); cudaFree(e->d_audit_ratio); cudaFree(e->d_audit_closest); cudaFree(e->d_dbg);
    cudaFree(e->d_tables); cudaFree(e->d_consts); cudaFree(e->d_in); cudaFree(e->d_gather); cudaFree(e->d_res); cudaFree(e->d_traj);
    cudaFree(e->d_act_prev);
    cudaFree(e->d_pred); cudaFree(e->d_predT); cudaFree(e->d_predZs); cudaFree(e->d_boxes); cudaFree(e->d_state9); cudaFree(e->d_goal3);
    cudaFree(e->d_last_cost); cudaFree(e->d_ts); cudaFree(e->d_flags); cudaFree(e->d_init_sfc); cudaFree(e->d_goal_kind); cudaFree(e->d_counters);
    cudaFree(e->dm.sqdist); cudaFree(e->dm.sat); cudaFree(e->d_sphere); cudaFree(e->d_tsphere); cudaFree(e->d_reach);
    for (auto& se : e->ev_pool) for (auto& ev : se.ev) if (ev) cudaEventDestroy(ev);
    for (auto& pr : e->step_ev) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    if (e->ev_begin) cudaEventDestroy(e->ev_begin);
    if (e->ev_end) cudaEventDestroy(e->ev_end);
    if (e->ev_order) cudaEventDestroy(e->ev_order);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->stream_aux) cudaStreamDestroy(e->stream_aux);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static int reset_state(lscgpu_engine* e) {
    const size_t N = e->N;
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemsetAsync(e->d_traj, 0, sizeof(float) * N * kTrajFloats, e->stream));      // src/traj_planner.cpp:36-39
    CU(cudaMemsetAsync(e->d_boxes, 0, sizeof(float) * N * 30, e->stream));
    CU(cudaMemsetAsync(e->d_last_cost, 0, sizeof(double) * N, e->stream));
    CU(cudaMemsetAsync(e->d_in, 0, sizeof(lscgpu_agent_in) * N, e->stream));
    CU(cudaMemsetAsync(e->d_res, 0, sizeof(lscgpu_agent_out) * N, e->stream));
    // agent_id = -1: empty slot. Not with the direct exchange: a peer that is already in its next step may be writing into
    // this buffer; its unused slots were marked empty when the buffer was made and are never written
    if (!e->p2p) CU(cudaMemsetAsync(e->d_gather, 0xff, sizeof(GatherSlot) * (size_t)e->n_slots, e->stream));
    CU(cudaMemsetAsync(e->d_act_prev, 0, sizeof(unsigned short) * kActSlots * N, e->stream));     // no candidates: cold starts
    CU(cudaMemsetAsync(e->d_block_of, 0xff, sizeof(int) * N, e->stream));
    CU(cudaMemsetAsync(e->d_reset_ever, 0, N, e->stream));                              // obs_slack_indices start empty
    CU(cudaMemsetAsync(e->d_any_reset, 0, sizeof(int), e->stream));
    std::vector<int> ones(N, 1);                                                         // flag_initialize_sfc, :49
    CU(cudaMemcpyAsync(e->d_init_sfc, ones.data(), sizeof(int) * N, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->planner_seq = 0;
    e->pending = 0; e->pending_launches = 0;
    if (e->h_kept_last) *e->h_kept_last = 0;
    return LSCGPU_OK;
}

extern "C" int lscgpu_create(const lscgpu_params* p, int n_agents, const lscgpu_agent_const* agents, int device,
                             lscgpu_engine** out) {
    if (!p || !agents || !out || n_agents < 1) return fail(LSCGPU_ERR_ARG, "null argument or n_agents < 1");
    if (p->goal_mode != 0 && p->goal_mode != 1) return fail(LSCGPU_ERR_ARG, "goal_mode must be 0 (static) or 1 (prior_based)");
    if (p->goal_mode == 1 && p->world_use_octomap && (p->grid_resolution < 0 || p->grid_margin < 0))
        return fail(LSCGPU_ERR_ARG, "grid_resolution and grid_margin must not be negative");
    if (p->M != 5 || p->n != 5 || p->phi != 3 || p->dim != 3)
        return fail(LSCGPU_ERR_ARG, "only M=5 (horizon/dt), n=5, phi=3, dim=3 is supported (launch/simulation.launch)");
    if (!(p->dt > 0) || !(p->world_resolution > 0)) return fail(LSCGPU_ERR_ARG, "dt and world_resolution must be positive");
    for (int k = 0; k < 3; k++)
        if (!(p->world_min[k] < p->world_max[k])) return fail(LSCGPU_ERR_ARG, "empty world box");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(LSCGPU_ERR_CUDA, "no CUDA device: the engine has no CPU path");
    if (device < 0 || device >= n_dev) return fail(LSCGPU_ERR_ARG, "device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    // the library carries sm_100a SASS only (arch-specific, not forward compatible)
    if (prop.major != 10 || prop.minor != 0)
        return fail(LSCGPU_ERR_CUDA, "device is not sm_100 (B200); the kernels are built for sm_100a only");

    lscgpu_engine* e = new lscgpu_engine;
    e->prm = *p; e->N = n_agents; e->device = device;
    if (!(e->prm.grid_resolution > 0)) e->prm.grid_resolution = 0.25;      // launch/simulation.launch:87
    e->n_sm = prop.multiProcessorCount;
    e->n_pad = (n_agents + 31) / 32 * 32;
    e->a0 = 0; e->a1 = n_agents; e->block = n_agents;
    e->consts_host.assign(agents, agents + n_agents);
    if (const char* s = getenv("LSCGPU_MAX_ITER")) e->max_iter = atoi(s);

    auto bail = [&](int code) { lscgpu_destroy(e); return code; };
#define CUB(expr)                                                                                         \
    do {                                                                                                  \
        cudaError_t err__ = (expr);                                                                       \
        if (err__ != cudaSuccess) {                                                                       \
            g_error = std::string(#expr) + ": " + cudaGetErrorString(err__);                              \
            return bail(LSCGPU_ERR_CUDA);                                                                 \
        }                                                                                                 \
    } while (0)
    CUB(configure_agent_plan());
    CUB(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&e->stream_aux, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&e->ev_order, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&e->ev_sfc, cudaEventDisableTiming));
    if (const char* v = getenv("LSCGPU_LPT_ORDER")) e->lpt_order = atoi(v) != 0;
    if (const char* v = getenv("LSCGPU_GRAPH")) e->use_graph = atoi(v) != 0;
    if (const char* v = getenv("LSCGPU_PLAN_THREADS")) { const int t = atoi(v); e->plan_threads = (t == 128 || t == 256 || t == 512) ? t : 0; }
    if (const char* v = getenv("LSCGPU_WIDE_KEPT")) e->wide_kept = atoi(v);
    if (const char* v = getenv("LSCGPU_ROW_CAP")) e->row_cap_forced = std::min(std::max(atoi(v), 0), 2560) / 32 * 32;
    if (const char* v = getenv("LSCGPU_SFC_WAIT_CYCLES")) e->sfc_wait_cycles = atoll(v);
    if (const char* v = getenv("LSCGPU_SLACK_KERNEL")) e->slack_kernel = atoi(v) != 0;     // 0: measure the step without its launch
    if (const char* v = getenv("LSCGPU_WARM_START")) e->warm_start = atoi(v) != 0;         // 0: every QP starts cold
    e->qp_debug = getenv("LSCGPU_QP_DEBUG") != nullptr;
    CUB(cudaEventCreate(&e->ev_begin));
    CUB(cudaEventCreate(&e->ev_end));

    // constant QP tables (once; the reference builds Q_base/Aeq_base once per TrajOptimizer)
    QpTablesDev* T = new QpTablesDev;
    try {
        build_qp_tables(p->dt, p->control_input_weight, p->terminal_weight, *T);
    } catch (const std::exception& ex) {
        delete T;
        g_error = std::string("QP tables: ") + ex.what();
        return bail(LSCGPU_ERR_ARG);
    }
    CUB(cudaMalloc(&e->d_tables, sizeof(QpTablesDev)));
    CUB(cudaMemcpy(e->d_tables, T, sizeof(QpTablesDev), cudaMemcpyHostToDevice));
    delete T;

    std::vector<AgentConstDev> cd(n_agents);
    for (int a = 0; a < n_agents; a++) {
        const lscgpu_agent_const& c = agents[a];
        if (!(c.radius > 0) || !(c.nominal_velocity > 0)) { g_error = "agent radius and nominal_velocity must be positive"; return bail(LSCGPU_ERR_ARG); }
        cd[a].radius = c.radius; cd[a].downwash = c.downwash; cd[a].v_nom = c.nominal_velocity;
        for (int k = 0; k < 3; k++) { cd[a].vmax[k] = c.max_vel[k]; cd[a].amax[k] = c.max_acc[k]; }
        size_t t = 0;
        while (t < e->radii.size() && e->radii[t] != c.radius) t++;
        if (t == e->radii.size()) e->radii.push_back(c.radius);
        cd[a].sat_index = (int)t; cd[a].pad = 0;
    }
    if (e->radii.size() > 16) { g_error = "more than 16 distinct agent radii"; return bail(LSCGPU_ERR_ARG); }
    const size_t N = n_agents;
    CUB(cudaMalloc(&e->d_consts, sizeof(AgentConstDev) * N));
    CUB(cudaMemcpy(e->d_consts, cd.data(), sizeof(AgentConstDev) * N, cudaMemcpyHostToDevice));
    {
        std::vector<float2> rdw(N);
        for (size_t a = 0; a < N; a++) rdw[a] = make_float2((float)cd[a].radius, (float)(cd[a].downwash * cd[a].radius));
        CUB(cudaMalloc(&e->d_rdw, sizeof(float2) * N));
        CUB(cudaMemcpy(e->d_rdw, rdw.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
    }
    CUB(cudaMalloc(&e->d_in, sizeof(lscgpu_agent_in) * N));
    CUB(cudaMalloc(&e->d_res, sizeof(lscgpu_agent_out) * N));
    CUB(cudaMalloc(&e->d_act_prev, sizeof(unsigned short) * kActSlots * N));
    CUB(cudaMalloc(&e->d_order, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_block_of, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_kept_step, sizeof(int)));
    CUB(cudaMemset(e->d_kept_step, 0, sizeof(int)));
    CUB(cudaHostAlloc(&e->h_kept_last, sizeof(int), cudaHostAllocMapped));
    if (e->h_kept_last) *e->h_kept_last = 0;
    CUB(cudaHostGetDevicePointer(&e->d_kept_last_map, e->h_kept_last, 0));
    CUB(cudaMalloc(&e->d_epoch, sizeof(int)));
    CUB(cudaMemset(e->d_epoch, 0, sizeof(int)));
    CUB(cudaMalloc(&e->d_commit_done, sizeof(int)));
    CUB(cudaMemset(e->d_commit_done, 0, sizeof(int)));
    CUB(cudaHostAlloc(&e->h_xchg_err, sizeof(int), cudaHostAllocMapped));
    *e->h_xchg_err = 0;
    CUB(cudaHostGetDevicePointer(&e->d_xchg_err_map, e->h_xchg_err, 0));
    CUB(cudaMalloc(&e->d_reset_ever, N));
    CUB(cudaMalloc(&e->d_any_reset, sizeof(int)));
    CUB(cudaMalloc(&e->d_sfc_ready, sizeof(int) * N));
    CUB(cudaMemset(e->d_sfc_ready, 0xff, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_sfc_box, sizeof(float) * 6 * N));
    CUB(cudaMalloc(&e->d_sfc_ok, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_traj, sizeof(float) * N * kTrajFloats));
    CUB(cudaMalloc(&e->d_pred, sizeof(float) * N * kTrajFloats));
    CUB(cudaMalloc(&e->d_predT, sizeof(float) * (size_t)kTrajFloats * e->n_pad));
    CUB(cudaMemset(e->d_predT, 0, sizeof(float) * (size_t)kTrajFloats * e->n_pad));
    CUB(cudaMalloc(&e->d_predZs, sizeof(float) * (size_t)30 * e->n_pad));
    CUB(cudaMemset(e->d_predZs, 0, sizeof(float) * (size_t)30 * e->n_pad));
    CUB(cudaMalloc(&e->d_sphere, sizeof(float4) * (size_t)kM * e->n_pad));
    CUB(cudaMemset(e->d_sphere, 0, sizeof(float4) * (size_t)kM * e->n_pad));
    CUB(cudaMalloc(&e->d_tsphere, sizeof(float4) * (size_t)e->n_pad));
    CUB(cudaMemset(e->d_tsphere, 0, sizeof(float4) * (size_t)e->n_pad));
    CUB(cudaMalloc(&e->d_reach, sizeof(float) * N * kM));
    CUB(cudaMalloc(&e->d_boxes, sizeof(float) * N * 30));
    CUB(cudaMalloc(&e->d_state9, sizeof(double) * N * 9));
    CUB(cudaMalloc(&e->d_goal3, sizeof(double) * N * 3));
    CUB(cudaMalloc(&e->d_last_cost, sizeof(double) * N));
    CUB(cudaMalloc(&e->d_ts, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_flags, sizeof(int) * N));
    CUB(cudaMemset(e->d_flags, 0, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_goal_kind, sizeof(int) * N));
    CUB(cudaMemset(e->d_goal_kind, 0, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_init_sfc, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_host_out, sizeof(void*)));
    CUB(cudaMemset(e->d_host_out, 0, sizeof(void*)));
    CUB(cudaHostAlloc(&e->h_host_out, sizeof(void*), cudaHostAllocDefault));
    *e->h_host_out = nullptr;
    if (const char* v = getenv("LSCGPU_ZERO_COPY")) e->zero_copy = atoi(v) != 0;      // 0: always copy the results after the step
    CUB(cudaMalloc(&e->d_counters, sizeof(StepCounters)));
    CUB(cudaMemset(e->d_counters, 0, sizeof(StepCounters)));
    int rc = alloc_gather(e, n_agents);
    if (rc != LSCGPU_OK) return bail(rc);
    rc = alloc_rows(e);
    if (rc != LSCGPU_OK) return bail(rc);
    rc = reset_state(e);
    if (rc != LSCGPU_OK) return bail(rc);
    *out = e;
    return LSCGPU_OK;
#undef CUB
}

// ---- octomap ---------------------------------------------------------------------------------------------------
static int coord_to_key(double c, double res) { return (int)std::floor((1.0 / res) * c); }   // OcTree::coordToKey - 32768


// ---- planning grid of goal planning (goal_mode 1 with an octomap) ----------------------------------------------------
// The bucket counts a std::unordered_map with the default load factor moves through while it grows one element at a
// time, recorded from libstdc++'s own policy object (astar_core.cuh explains why the search needs them).
static int goal_bucket_sequence(int row_capacity, int* seq) {
    std::__detail::_Prime_rehash_policy pol;
    size_t bkt = 1, cnt = 0;
    int n = 0;
    seq[n++] = 1;
    while ((int)bkt < row_capacity) {
        if (n == kAstarMaxLevels) return -1;
        const auto need = pol._M_need_rehash(bkt, cnt, 1);
        if (need.first) { bkt = need.second; seq[n++] = (int)bkt; }
        cnt++;
    }
    for (int k = n; k < kAstarMaxLevels; k++) seq[k] = seq[n - 1];
    return n;
}

// GridBasedPlanner::updateGridInfo (src/grid_based_planner.cpp:70-90) + the static part of updateGridMap (:109-123), once
// per map; scratch for one search per warp.
static int setup_goal_grid(lscgpu_engine* e) {
    cudaFree(e->d_goal_axis); cudaFree(e->d_goal_static); cudaFree(e->d_goal_cell); cudaFree(e->d_goal_g); cudaFree(e->d_goal_next);
    cudaFree(e->d_goal_bkt); cudaFree(e->d_goal_bstamp); cudaFree(e->d_goal_path);
    e->d_goal_bstamp = nullptr;
    e->d_goal_axis = nullptr; e->d_goal_static = nullptr; e->d_goal_cell = nullptr; e->d_goal_g = nullptr; e->d_goal_next = nullptr;
    e->d_goal_bkt = nullptr; e->d_goal_path = nullptr;
    e->have_goal_grid = false;
    GoalGridDev& g = e->goal_grid;
    g = GoalGridDev{};
    const double r = e->prm.grid_resolution;
    double gmax[3];
    for (int k = 0; k < 3; k++) {
        g.gmin[k] = -std::floor((-(double)e->prm.world_min[k] + 1e-9) / r) * r;
        gmax[k] = std::floor(((double)e->prm.world_max[k] + 1e-9) / r) * r;
        g.dim[k] = (int)std::round((gmax[k] - g.gmin[k]) / r) + 1;
        if (g.dim[k] < 1) return fail(LSCGPU_ERR_ARG, "empty planning grid");
    }
    g.res = r;
    if (g.dim[0] > 2048) return fail(LSCGPU_ERR_ARG, "planning grid has more than 2048 rows along x (grid_resolution too fine for this world)");
    const size_t cells = (size_t)g.dim[0] * g.dim[1] * g.dim[2];
    if (cells > ((size_t)1 << 26)) return fail(LSCGPU_ERR_ARG, "planning grid too large");
    g.cells = (int)cells;
    g.cells_pad = (cells + 15) / 16 * 16;
    if (goal_bucket_sequence(g.dim[1] * g.dim[2], g.bkt_seq) < 0) return fail(LSCGPU_ERR_ARG, "planning grid rows too long");
    g.bcap = g.bkt_seq[kAstarMaxLevels - 1];
    g.magic_a = astar_magic((unsigned)g.dim[2], cells); g.magic_w = astar_magic((unsigned)g.dim[1], cells);
    for (int k = 0; k < kAstarMaxLevels; k++) g.bkt_magic[k] = astar_magic((unsigned)g.bkt_seq[k], cells);
    std::vector<float> axis((size_t)g.dim[0] + g.dim[1] + g.dim[2]);
    {
        size_t o = 0;
        for (int k = 0; k < 3; k++)
            for (int i = 0; i < g.dim[k]; i++) {
                volatile double prod = i * r;                     // gridVectorToPoint3D (:305-310): product and sum rounded separately
                axis[o++] = (float)(g.gmin[k] + prod);
            }
    }
    CU(cudaMalloc(&e->d_goal_axis, sizeof(float) * axis.size()));
    CU(cudaMemcpyAsync(e->d_goal_axis, axis.data(), sizeof(float) * axis.size(), cudaMemcpyHostToDevice, e->stream));
    g.axis_pts = e->d_goal_axis;
    const int n_radii = (int)e->radii.size();
    CU(cudaMalloc(&e->d_goal_static, g.cells_pad * (size_t)n_radii));
    CU(cudaMemsetAsync(e->d_goal_static, 0, g.cells_pad * (size_t)n_radii, e->stream));
    g.static_occ = e->d_goal_static;
    TempBuf radii_buf;
    CU(radii_buf.alloc(sizeof(double) * n_radii));
    CU(cudaMemcpyAsync(radii_buf.p, e->radii.data(), sizeof(double) * n_radii, cudaMemcpyHostToDevice, e->stream));
    launch_goal_static_grid(g, e->dm, e->prm.world_resolution, radii_buf.as<double>(), n_radii, (float)e->prm.grid_margin, e->d_goal_static, e->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    // One search per warp. Small grids keep the warp's search state in shared memory (one warp per SM, or more when several
    // fit); larger ones run on per-warp scratch in global memory, as many warps as the device keeps resident, bounded by
    // 4 GB of scratch.
    CU(configure_goal_astar(g));
    const size_t sh = goal_astar_shared_bytes(g);
    int blocks;
    if (sh) {
        blocks = std::min(e->N, e->n_sm * std::max(1, (int)((227 * 1024) / (sh + 1024))));
    } else {
        const size_t per_warp = g.cells_pad * (1 + 3 * sizeof(int)) + 2 * (size_t)g.dim[0] * g.bcap * sizeof(int);
        blocks = std::min(e->N, e->n_sm * 16);
        blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)blocks, ((size_t)4 << 30) / per_warp));
        CU(cudaMalloc(&e->d_goal_cell, g.cells_pad * (size_t)blocks));
        CU(cudaMalloc(&e->d_goal_g, sizeof(int) * g.cells_pad * (size_t)blocks));
        CU(cudaMalloc(&e->d_goal_next, sizeof(int) * g.cells_pad * (size_t)blocks));
        CU(cudaMalloc(&e->d_goal_bkt, sizeof(int) * (size_t)g.dim[0] * g.bcap * (size_t)blocks));
        CU(cudaMalloc(&e->d_goal_bstamp, sizeof(int) * (size_t)g.dim[0] * g.bcap * (size_t)blocks));
    }
    e->goal_blocks = blocks;
    CU(cudaMalloc(&e->d_goal_path, sizeof(int) * g.cells_pad * (size_t)blocks));
    if (!e->d_goal_ticket) CU(cudaMalloc(&e->d_goal_ticket, sizeof(int)));
    if (!e->d_goal_expansions) {
        CU(cudaMalloc(&e->d_goal_expansions, sizeof(unsigned long long)));
        CU(cudaMemset(e->d_goal_expansions, 0, sizeof(unsigned long long)));
    }
    e->have_goal_grid = true;
    return LSCGPU_OK;
}

static int build_map(lscgpu_engine* e, const int32_t* keys, int n) {
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    drop_graph(e);                          // the step graph carries the table pointers
    cudaFree(e->dm.sqdist); cudaFree(e->dm.sat);
    e->dm = DistMapDev{};
    e->have_map = false;
    const double res = e->prm.world_resolution;
    const float maxdist = 1.0f;                                       // src/multi_sync_simulator.cpp:160
    const int md = (int)(maxdist / res + 1);                          // DynamicEDTOctomap ctor
    if (md * md > 255) return fail(LSCGPU_ERR_ARG, "world_resolution too fine: clamped squared distance exceeds 8 bits");
    e->dm.max_sq = md * md;
    size_t total = 1, tab = 1;
    for (int k = 0; k < 3; k++) {
        const int lo = coord_to_key((double)e->prm.world_min[k], res), hi = coord_to_key((double)e->prm.world_max[k], res);
        e->dm.off[k] = lo; e->dm.size[k] = hi - lo + 1;
        total *= (size_t)e->dm.size[k]; tab *= (size_t)(e->dm.size[k] + 1);
    }
    if (tab > (size_t)1 << 31) return fail(LSCGPU_ERR_ARG, "world too large for the 32-bit blocked-voxel tables");
    for (int k = 0; k < 3; k++)
        if (std::fabs(e->prm.world_min[k]) > 60.f || std::fabs(e->prm.world_max[k]) > 60.f)
            return fail(LSCGPU_ERR_ARG, "SFC kernel supports worlds within +-60 m (float32 sample rounding must stay below the 1e-5 nudge)");
    // blocked predicate of isObstacleInBox (include/corridor_constructor.hpp:113-114), per distinct radius:
    // getDistance = (float)((float)sqrt(sq) * res) < radius + 0.5 res - 1e-5  -> largest blocked squared distance
    std::vector<int> thr(e->radii.size());
    for (size_t t = 0; t < e->radii.size(); t++) {
        int last = -1;
        for (int sq = 0; sq <= e->dm.max_sq; sq++) {
            const float cell = (float)std::sqrt((double)sq);
            const float dist = (float)(cell * res);
            if ((double)dist < e->radii[t] + 0.5 * res - 1e-5) last = sq; else break;
        }
        thr[t] = last;
    }
    e->dm.n_tables = (int)thr.size();
    TempBuf keys_buf, thr_buf, sa, sb;              // freed on every return path
    CU(cudaMalloc(&e->dm.sqdist, total));
    CU(cudaMalloc(&e->dm.sat, tab * sizeof(int) * thr.size()));
    CU(sa.alloc(total)); CU(sb.alloc(total));
    CU(thr_buf.alloc(sizeof(int) * thr.size()));
    CU(cudaMemcpyAsync(thr_buf.p, thr.data(), sizeof(int) * thr.size(), cudaMemcpyHostToDevice, e->stream));
    if (n > 0) {
        CU(keys_buf.alloc(sizeof(int32_t) * 3 * (size_t)n));
        CU(cudaMemcpyAsync(keys_buf.p, keys, sizeof(int32_t) * 3 * (size_t)n, cudaMemcpyHostToDevice, e->stream));
    }
    launch_edt_build(keys_buf.as<int32_t>(), n, e->dm, thr_buf.as<int>(), e->dm.n_tables, sa.as<uint8_t>(), sb.as<uint8_t>(), e->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    int64_t occ = 0;
    for (int i = 0; i < n; i++) {
        bool in = true;
        for (int k = 0; k < 3; k++) { const int c = keys[3 * i + k] - e->dm.off[k]; if (c < 0 || c >= e->dm.size[k]) in = false; }
        occ += in;
    }
    e->n_occupied = occ;
    e->have_map = true;
    if (e->prm.goal_mode == 1 && e->prm.world_use_octomap) return setup_goal_grid(e);
    return LSCGPU_OK;
}

extern "C" int lscgpu_set_octomap_voxels(lscgpu_engine* e, const int32_t* keys, int n) {
    if (!e || n < 0 || (n > 0 && !keys)) return fail(LSCGPU_ERR_ARG, "bad voxel list");
    return build_map(e, keys, n);
}

extern "C" int lscgpu_set_octomap_file(lscgpu_engine* e, const char* path) {
    if (!e || !path) return fail(LSCGPU_ERR_ARG, "null argument");
    OccupiedVoxels vox;
    std::string err;
    if (!read_bt_file(path, vox, err)) return fail(LSCGPU_ERR_IO, err);
    if (std::fabs(vox.res - e->prm.world_resolution) > 1e-9)
        return fail(LSCGPU_ERR_ARG, "octomap resolution differs from world_resolution");
    return build_map(e, vox.keys.data(), (int)(vox.keys.size() / 3));
}

extern "C" int lscgpu_get_distmap_info(lscgpu_engine* e, int32_t size[3], int32_t offset[3], int64_t* n_occupied) {
    if (!e || !e->have_map) return fail(LSCGPU_ERR_STATE, "no octomap uploaded");
    for (int k = 0; k < 3; k++) { size[k] = e->dm.size[k]; offset[k] = e->dm.off[k]; }
    if (n_occupied) *n_occupied = e->n_occupied;
    return LSCGPU_OK;
}
extern "C" int lscgpu_get_distmap_sqdist(lscgpu_engine* e, uint8_t* out) {
    if (!e || !e->have_map) return fail(LSCGPU_ERR_STATE, "no octomap uploaded");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpy(out, e->dm.sqdist, (size_t)e->dm.size[0] * e->dm.size[1] * e->dm.size[2], cudaMemcpyDeviceToHost));
    return LSCGPU_OK;
}

// ---- sharding / NCCL ---------------------------------------------------------------------------------------------
extern "C" int lscgpu_set_shard(lscgpu_engine* e, int a0, int a1) {
    if (!e || a0 < 0 || a1 > e->N || a0 > a1) return fail(LSCGPU_ERR_ARG, "bad shard range");
    if (e->comm) return fail(LSCGPU_ERR_STATE, "the partition is fixed by lscgpu_nccl_init");
    if (e->pending) return fail(LSCGPU_ERR_STATE, "lscgpu_synchronize first");
    CU(cudaSetDevice(e->device));
    e->a0 = a0; e->a1 = a1;
    // slots of the gather buffer beyond the shard must read "empty"
    CU(cudaMemsetAsync(e->d_gather, 0xff, sizeof(GatherSlot) * (size_t)e->n_slots, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return alloc_rows(e);
}

extern "C" int lscgpu_nccl_unique_id(uint8_t id_out[128]) {
    std::string err;
    if (!load_nccl(err)) return fail(LSCGPU_ERR_NCCL, err);
    NcclId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return fail(LSCGPU_ERR_NCCL, std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    std::memcpy(id_out, id.internal, 128);
    return LSCGPU_OK;
}

extern "C" int lscgpu_nccl_init(lscgpu_engine* e, const uint8_t id_bytes[128], int rank, int n_ranks) {
    if (!e || !id_bytes || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(LSCGPU_ERR_ARG, "bad rank / n_ranks");
    if (e->comm) return fail(LSCGPU_ERR_STATE, "a communicator is already attached to this engine");
    if (e->pending) return fail(LSCGPU_ERR_STATE, "lscgpu_synchronize first");
    std::string err;
    if (!load_nccl(err)) return fail(LSCGPU_ERR_NCCL, err);
    CU(cudaSetDevice(e->device));
    NcclId id;
    std::memcpy(id.internal, id_bytes, 128);
    NcclComm comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, n_ranks, id, rank);
    if (rc != 0) return fail(LSCGPU_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    e->comm = comm;
    e->rank = rank; e->n_ranks = n_ranks;
    e->block = (e->N + n_ranks - 1) / n_ranks;           // slots per rank in the gather buffer; the last may stay empty
    e->a0 = 0; e->a1 = e->N;                             // every rank may plan any agent: the LPT order is dealt out
    const int r1 = alloc_gather(e, e->block * n_ranks);
    if (r1 != LSCGPU_OK) return r1;
    CU(cudaMemset(e->d_gather, 0xff, sizeof(GatherSlot) * (size_t)e->n_slots));
    return alloc_rows(e);
}

// Direct exchange over NVLink peer memory. lscgpu_p2p_export makes this rank's exchange buffer (two parities of the gather
// buffer + one arrival counter per source rank) and returns its CUDA IPC handle; the caller hands every rank's handle to
// lscgpu_p2p_attach (any out-of-band channel), which maps the peers' buffers and switches the step from ncclAllGather to
// stores by the planning blocks (k_agent_plan epilogue, k_commit).
extern "C" int lscgpu_p2p_export(lscgpu_engine* e, uint8_t handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    if (!e || !handle_out) return fail(LSCGPU_ERR_ARG, "null argument");
    if (!e->comm) return fail(LSCGPU_ERR_STATE, "lscgpu_nccl_init first (it fixes rank, n_ranks and the slot layout)");
    if (e->pending) return fail(LSCGPU_ERR_STATE, "lscgpu_synchronize first");
    if (e->n_ranks > 32) return fail(LSCGPU_ERR_ARG, "direct exchange supports up to 32 ranks");
    CU(cudaSetDevice(e->device));
    if (!e->d_xchg) {
        const int slots = e->block * e->n_ranks;
        const size_t bytes = PeerExchange::counters_offset(slots) + sizeof(int) * (size_t)e->n_ranks;
        CU(cudaMalloc(&e->d_xchg, bytes));
        CU(cudaMemset(e->d_xchg, 0xff, sizeof(GatherSlot) * 2 * (size_t)slots));
        CU(cudaMemset((char*)e->d_xchg + PeerExchange::counters_offset(slots), 0, sizeof(int) * (size_t)e->n_ranks));
        CU(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, e->d_xchg));
    std::memcpy(handle_out, &h, 64);
    return LSCGPU_OK;
}

extern "C" int lscgpu_p2p_attach(lscgpu_engine* e, const uint8_t* handles) {
    if (!e || !handles) return fail(LSCGPU_ERR_ARG, "null argument");
    if (!e->d_xchg) return fail(LSCGPU_ERR_STATE, "lscgpu_p2p_export first");
    if (e->p2p) return fail(LSCGPU_ERR_STATE, "the direct exchange is already attached");
    if (e->pending) return fail(LSCGPU_ERR_STATE, "lscgpu_synchronize first");
    CU(cudaSetDevice(e->device));
    e->peer_ptrs.assign(e->n_ranks, nullptr);
    for (int r = 0; r < e->n_ranks; r++) {
        if (r == e->rank) { e->peer_ptrs[r] = e->d_xchg; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        const cudaError_t rc = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (rc != cudaSuccess) {
            for (int q = 0; q < r; q++) if (q != e->rank && e->peer_ptrs[q]) cudaIpcCloseMemHandle(e->peer_ptrs[q]);
            e->peer_ptrs.clear();
            return fail(LSCGPU_ERR_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(rc));
        }
        e->peer_ptrs[r] = p;
    }
    CU(cudaMalloc(&e->d_peers, sizeof(void*) * (size_t)e->n_ranks));
    CU(cudaMemcpy(e->d_peers, e->peer_ptrs.data(), sizeof(void*) * (size_t)e->n_ranks, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(&e->p2p_base_epoch, e->d_epoch, sizeof(int), cudaMemcpyDeviceToHost));
    drop_graph(e);
    cudaFree(e->d_gather);                                  // the step now writes into the exchange buffer
    e->d_gather = e->d_xchg;
    e->p2p = true;
    return LSCGPU_OK;
}

// ---- the step ----------------------------------------------------------------------------------------------------
// Kernels of one step on the engine stream `s` (k_qp_order forks onto the aux stream and is joined before the plan):
//   k_qp_order || k_predict [-> k_goal_plan] -> k_agent_plan -> [ncclAllGather] -> k_commit
// `ev` (profiling) gets an event after each stage. Callable under stream capture.
static int enqueue_step_kernels(lscgpu_engine* e, int planner_seq, int threads, cudaEvent_t* ev, int* launches_out) {
    cudaStream_t s = e->stream;
    int launches = 0;
    const int n_plan = e->n_plan();
    const bool dealt = e->comm != nullptr;
    // scheduling order of this step's blocks from the cost of the previous step's plans (records in d_res)
    const int n_order = dealt ? e->N : e->a1 - e->a0;
    const bool ordered = e->lpt_order && planner_seq > 1 && n_order > 1;
    const bool use_sfc = e->prm.world_use_octomap != 0;
    const bool side = !e->profiling;
    cudaStream_t so = side ? e->stream_aux : s;
    if (side && (ordered || (use_sfc && e->prm.goal_mode != 1))) { CU(cudaEventRecord(e->ev_fork, s)); CU(cudaStreamWaitEvent(so, e->ev_fork, 0)); }
    if (ordered) {
        launch_qp_order(n_order, dealt ? 0 : e->a0, e->d_res, e->d_order, so); launches++;
        if (side) CU(cudaEventRecord(e->ev_order, so));
    }
    // the step's new SFC boxes depend on the step's inputs only: grown beside k_predict and the LSC rows — unless the goals
    // are planned in this step (goal_mode 1): the box grows toward the agent's CURRENT goal, so k_sfc_step follows the
    // goal kernel on the engine stream
    const bool sfc_after_goal = use_sfc && e->prm.goal_mode == 1;
    auto launch_sfc = [&](cudaStream_t st, const double* goal3) -> int {
        SfcStepLaunch sl{};
        sl.n = n_plan;
        sl.order = ordered ? e->d_order : nullptr;
        sl.order_first = dealt ? e->rank : 0; sl.order_stride = dealt ? e->n_ranks : 1;
        sl.agent_base = dealt ? e->rank : e->a0; sl.agent_stride = dealt ? e->n_ranks : 1;
        sl.dm = e->dm; sl.res = e->prm.world_resolution;
        for (int k = 0; k < 3; k++) { sl.wmin[k] = e->prm.world_min[k]; sl.wmax[k] = e->prm.world_max[k]; }
        sl.in = e->d_in; sl.prev_traj = e->d_traj; sl.consts = e->d_consts; sl.init_sfc = e->d_init_sfc;
        sl.planner_seq = planner_seq; sl.reset_threshold = e->prm.reset_threshold;
        sl.epoch = e->d_epoch; sl.sfc_box_g = e->d_sfc_box; sl.sfc_ok_g = e->d_sfc_ok; sl.sfc_ready = e->d_sfc_ready;
        sl.goal3 = goal3;
        if (ev) CU(cudaEventRecord(ev[5], st));
        launch_sfc_step(sl, st); launches++;
        if (ev) CU(cudaEventRecord(ev[6], st));
        return LSCGPU_OK;
    };
    if (use_sfc && n_plan > 0 && !sfc_after_goal) {
        if (const int rc = launch_sfc(so, nullptr)) return rc;
        if (side) CU(cudaEventRecord(e->ev_sfc, so));
    }
    PredictLaunch pl{};
    pl.n_agents = e->N; pl.n_pad = e->n_pad; pl.planner_seq = planner_seq;
    pl.dt = e->prm.dt; pl.reset_threshold = e->prm.reset_threshold;
    pl.in = e->d_in; pl.prev_traj = e->d_traj; pl.consts = e->d_consts;
    pl.pred = e->d_pred; pl.predT = e->d_predT; pl.predZs = e->d_predZs; pl.state9 = e->d_state9; pl.goal3 = e->d_goal3;
    pl.ts = e->d_ts; pl.flags = e->d_flags; pl.sphere = e->d_sphere; pl.tsphere = e->d_tsphere; pl.reach = e->d_reach;
    pl.reset_ever = e->d_reset_ever; pl.any_reset = e->d_any_reset; pl.init_sfc = e->d_init_sfc;
    launch_predict(pl, s); launches++;
    bool order_waited = false;
    if (e->prm.goal_mode == 1) {
        GoalLaunch gl{};
        gl.reset_ever = e->d_reset_ever;
        gl.n_agents = e->N; gl.dt = e->prm.dt; gl.goal_threshold = e->prm.goal_threshold; gl.goal_radius = e->prm.goal_radius;
        gl.priority_dist_threshold = e->prm.priority_dist_threshold;
        gl.in = e->d_in; gl.prev_traj = e->d_traj; gl.pred = e->d_pred; gl.consts = e->d_consts;
        gl.goal3 = e->d_goal3; gl.ts = e->d_ts; gl.goal_kind = e->d_goal_kind;
        if (!use_sfc) {
            launch_goal_plan(gl, s); launches++;
        } else if (n_plan > 0) {
            // with an octomap: grid planner + A* + line-of-sight goal, for the agents this engine plans, in their order
            if (ordered && !e->profiling) { CU(cudaStreamWaitEvent(s, e->ev_order, 0)); order_waited = true; }
            GoalAstarLaunch al{};
            al.g = gl; al.n = n_plan;
            al.order = ordered ? e->d_order : nullptr;
            al.order_first = dealt ? e->rank : 0; al.order_stride = dealt ? e->n_ranks : 1;
            al.agent_base = dealt ? e->rank : e->a0; al.agent_stride = dealt ? e->n_ranks : 1;
            al.dm = e->dm; al.world_res = e->prm.world_resolution; al.grid = e->goal_grid; al.n_blocks = e->goal_blocks;
            al.cell = e->d_goal_cell; al.gcost = e->d_goal_g; al.next = e->d_goal_next; al.bkt = e->d_goal_bkt; al.bstamp = e->d_goal_bstamp; al.path = e->d_goal_path;
            al.expansions = e->d_goal_expansions;
            al.next_agent = e->d_goal_ticket;
            CU(cudaMemsetAsync(e->d_goal_ticket, 0, sizeof(int), s));
            launch_goal_astar(al, s); launches++;
            if (const int rc = launch_sfc(s, e->d_goal3)) return rc;
        }
    }
    if (ev) CU(cudaEventRecord(ev[1], s));
    if (ordered && !e->profiling && !order_waited) CU(cudaStreamWaitEvent(s, e->ev_order, 0));

    PlanLaunch L{};
    L.n_agents = e->N; L.n_pad = e->n_pad; L.n_blocks = n_plan;
    L.threads = threads; L.sfc_wait_cycles = e->sfc_wait_cycles;
    L.order = ordered ? e->d_order : nullptr;
    L.order_first = dealt ? e->rank : 0; L.order_stride = dealt ? e->n_ranks : 1;
    // without an order (first step): rank r of a dealt job takes agents r, r + G, ... through an identity "order"
    L.agent_base = e->a0;
    L.agent_stride = dealt ? e->n_ranks : 1;
    if (dealt) L.agent_base = e->rank;
    L.pred = e->d_pred; L.predT = e->d_predT; L.predZs = e->d_predZs; L.consts = e->d_consts; L.rdw = e->d_rdw; L.T = e->d_tables;
    L.state9 = e->d_state9; L.goal3 = e->d_goal3; L.ts = e->d_ts; L.sphere = e->d_sphere; L.tsphere = e->d_tsphere; L.reach = e->d_reach;
    L.row_cap = e->row_cap_forced >= 0 ? e->row_cap_forced : (threads == 128 ? 256 : (threads == 512 ? 2048 : 1024));
    L.P_pad = e->P_pad; L.mirror_rows = e->mirror_rows ? 1 : 0;
    L.kept_step = e->d_kept_step;
    L.rows = e->d_rows; L.kept = e->d_kept; L.kept_count = e->d_kept_count; L.safe = e->d_safe;
    L.block_of = e->mirror_rows ? e->d_block_of : nullptr;
    L.use_sfc = e->prm.world_use_octomap ? 1 : 0;
    L.dm = e->dm; L.res = e->prm.world_resolution;
    L.epoch = e->d_epoch; L.sfc_ready = e->d_sfc_ready; L.sfc_box_g = e->d_sfc_box; L.sfc_ok_g = e->d_sfc_ok;
    L.in = e->d_in; L.boxes = e->d_boxes; L.init_sfc = e->d_init_sfc; L.flags = e->d_flags;
    for (int k = 0; k < 3; k++) { L.wmin[k] = e->prm.world_min[k]; L.wmax[k] = e->prm.world_max[k]; }
    L.max_iter = e->max_iter;
    L.out = e->d_gather; L.out_base = dealt ? e->rank * e->block : 0;
    L.act_prev = e->warm_start ? e->d_act_prev : nullptr;
    PeerExchange px{};
    if (e->p2p) {
        px.peers = e->d_peers; px.n_ranks = e->n_ranks; px.rank = e->rank; px.slots = e->block * e->n_ranks; px.block = e->block;
        px.n_agents = e->N; px.base_epoch = e->p2p_base_epoch;
    }
    L.px = px;
    L.host_out = e->d_host_out;
    L.prev_traj = e->d_traj; L.last_cost = e->d_last_cost; L.goal_kind = e->d_goal_kind;
    L.counters = e->d_counters;
    L.any_reset = e->slack_kernel ? e->d_any_reset : nullptr; L.reset_ever = e->d_reset_ever;
    L.slack_w = e->slack_w; L.row_cap_slack = e->row_cap_slack;
    if (e->qp_debug) { if (!e->d_dbg) CU(cudaMalloc(&e->d_dbg, sizeof(long long) * 10 * (size_t)e->N)); L.dbg = e->d_dbg; }
    if (n_plan > 0) { launch_agent_plan(L, s); launches += L.any_reset ? 2 : 1; }
    if (ev) CU(cudaEventRecord(ev[2], s));

    if (dealt && e->n_ranks > 1 && !e->p2p) {
        // in-place all-gather: every rank's block of records lands in every replica (with the direct exchange the planning
        // blocks have already stored them into every peer's buffer)
        const size_t bytes = sizeof(GatherSlot) * (size_t)e->block;
        const int rc = g_nccl.AllGather((const char*)e->d_gather + bytes * e->rank, e->d_gather, bytes, /*ncclInt8*/ 0, e->comm, s);
        if (rc != 0) return fail(LSCGPU_ERR_NCCL, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    }
    if (ev) CU(cudaEventRecord(ev[3], s));
    const int n_slots = dealt ? e->block * e->n_ranks : n_plan;
    if (use_sfc && n_plan > 0 && side && !sfc_after_goal) CU(cudaStreamWaitEvent(s, e->ev_sfc, 0));     // k_commit rewrites what k_sfc_step reads
    launch_commit(n_slots, e->d_gather, e->d_res, e->d_act_prev, e->d_traj, e->d_in, e->d_last_cost,
                  use_sfc ? e->d_boxes : nullptr, e->d_init_sfc, e->d_epoch, e->d_kept_step, e->d_kept_last_map, px, e->d_commit_done,
                  e->d_xchg_err_map, s); launches++;
    if (ev) CU(cudaEventRecord(ev[4], s));
    *launches_out = launches;
    return LSCGPU_OK;
}

static int step_device(lscgpu_engine* e) {
    if (e->prm.world_use_octomap && !e->have_map)
        return fail(LSCGPU_ERR_STATE, "world_use_octomap is set but no octomap was uploaded (lscgpu_set_octomap_*)");
    if (e->prm.world_use_octomap && e->prm.goal_mode == 1 && !e->have_goal_grid)
        return fail(LSCGPU_ERR_STATE, "goal planning grid missing (octomap upload failed?)");
    cudaStream_t s = e->stream;
    const int seq = e->planner_seq + 1;                     // src/traj_planner.cpp:127; committed once the step is enqueued
    const bool prof = e->profiling;
    cudaEvent_t* ev = nullptr;
    if (prof) {
        if ((int)e->ev_pool.size() <= e->pending) {
            lscgpu_engine::StepEvents se{};
            for (auto& x : se.ev) CU(cudaEventCreate(&x));
            e->ev_pool.push_back(se);
        }
        ev = e->ev_pool[e->pending].ev;
    }
    if ((int)e->step_ev.size() <= e->pending) {
        std::pair<cudaEvent_t, cudaEvent_t> pr;
        CU(cudaEventCreate(&pr.first)); CU(cudaEventCreate(&pr.second));
        e->step_ev.push_back(pr);
    }
    if (e->pending == 0 && !e->stats_fresh) {
        // nobody asked for the statistics of the previous batch: its events are about to be reused, drop its counters too
        CU(cudaMemsetAsync(e->d_counters, 0, sizeof(StepCounters), s));
        e->stats = lscgpu_step_stats{};
        e->stats_fresh = true;
    }
    if (e->pending == 0) CU(cudaEventRecord(e->ev_begin, s));
    CU(cudaEventRecord(e->step_ev[e->pending].first, s));
    if (prof) CU(cudaEventRecord(ev[0], s));
    int launches = 0;
    // Steps from the second on are the same launches with the same arguments: one CUDA graph, instantiated at the
    // first such step and re-launched afterwards (profiling and debug modes enqueue the kernels directly).
    int threads = e->plan_threads;
    if (threads == 0) {
        const int n_plan = std::max(e->n_plan(), 1);
        const int kept_last = *(volatile int*)e->h_kept_last;          // of the last step the GPU has completed
        // one wave of wide blocks (two per SM) holds every agent: the step is the slowest agent's chain, and eight warps
        // shorten it (multi-GPU shares); otherwise four narrow blocks per SM overlap more QP chains
        threads = (n_plan <= 2 * e->n_sm || kept_last / n_plan >= e->wide_kept) ? 256 : 128;
        // at most one agent per SM (small shares of a multi-GPU job): sixteen warps per agent
        if (n_plan <= e->n_sm) threads = 512;
    }
    const int gi = threads == 128 ? 1 : (threads == 512 ? 2 : 0);
    const bool graphable = e->use_graph && !e->graph_failed && !prof && !e->qp_debug && seq >= 2;
    if (graphable && !e->graph[gi]) {
        cudaGraph_t g = nullptr;
        int rc = LSCGPU_OK;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = enqueue_step_kernels(e, seq, threads, nullptr, &launches);
            const cudaError_t ce = cudaStreamEndCapture(s, &g);
            if (rc == LSCGPU_OK && ce == cudaSuccess && g &&
                cudaGraphInstantiate(&e->graph[gi], g, nullptr, nullptr, 0) == cudaSuccess) {
                e->graph_launches[gi] = launches;
            } else {
                e->graph[gi] = nullptr; e->graph_failed = true;
                cudaGetLastError();             // clear the capture error; the kernels are enqueued directly below
            }
            if (g) cudaGraphDestroy(g);
        } else {
            e->graph_failed = true;
            cudaGetLastError();
        }
    }
    if (graphable && e->graph[gi]) {
        CU(cudaGraphLaunch(e->graph[gi], s));
        launches = e->graph_launches[gi];
    } else {
        const int rc = enqueue_step_kernels(e, seq, threads, ev, &launches);
        if (rc != LSCGPU_OK) return rc;
    }
    CU(cudaEventRecord(e->step_ev[e->pending].second, s));
    CU(cudaEventRecord(e->ev_end, s));
    CU(cudaGetLastError());
    e->planner_seq = seq;
    e->pending++;
    e->pending_launches += launches;
    return LSCGPU_OK;
}

// wait for the enqueued steps; the statistics of the batch are computed when somebody asks for them
static int finish_steps(lscgpu_engine* e) {
    CU(cudaStreamSynchronize(e->stream));
    if (e->h_xchg_err && *(volatile int*)e->h_xchg_err) {
        *e->h_xchg_err = 0;
        return fail(LSCGPU_ERR_NCCL, "direct exchange: a peer's records did not arrive within 2 s (ranks out of step?)");
    }
    if (e->pending > 0) {
        e->done_steps = e->pending; e->done_launches = e->pending_launches;
        e->stats_fresh = false;
    }
    e->pending = 0; e->pending_launches = 0;
    return LSCGPU_OK;
}

static int compute_stats(lscgpu_engine* e) {
    if (e->stats_fresh) return LSCGPU_OK;
    lscgpu_step_stats& st = e->stats;
    st = lscgpu_step_stats{};
    StepCounters c;
    CU(cudaMemcpyAsync(&c, e->d_counters, sizeof c, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemsetAsync(e->d_counters, 0, sizeof(StepCounters), e->stream));
    unsigned long long astar = 0;
    if (e->d_goal_expansions) {
        CU(cudaMemcpyAsync(&astar, e->d_goal_expansions, sizeof astar, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemsetAsync(e->d_goal_expansions, 0, sizeof astar, e->stream));
    }
    CU(cudaStreamSynchronize(e->stream));
    st.astar_expansions = (int64_t)astar;
    st.steps = e->done_steps;
    CU(cudaEventElapsedTime(&st.ms_total, e->ev_begin, e->ev_end));
    for (int i = 0; i < e->done_steps; i++) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e->step_ev[i].first, e->step_ev[i].second));
        st.ms_steps += ms;
    }
    if (e->profiling) {
        for (int i = 0; i < e->done_steps && i < (int)e->ev_pool.size(); i++) {
            cudaEvent_t* ev = e->ev_pool[i].ev;
            float ms[4];
            for (int k = 0; k < 4; k++) CU(cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]));
            st.ms_predict += ms[0]; st.ms_plan += ms[1]; st.ms_exchange += ms[2]; st.ms_commit += ms[3];
            if (e->prm.world_use_octomap && e->n_plan() > 0) {      // serialised in front of k_predict in this mode
                float sf = 0.f;
                CU(cudaEventElapsedTime(&sf, ev[5], ev[6]));
                st.ms_sfc += sf; st.ms_predict -= sf;
            }
        }
    }
    if (e->d_dbg && e->qp_debug) {
        // section cycle counts of the last step (build with -DLSCGPU_QP_SECTION_TIMERS): slowest block and swarm mean
        const int nl = e->n_plan();
        std::vector<long long> h((size_t)nl * 10);
        CU(cudaMemcpy(h.data(), e->d_dbg, sizeof(long long) * 10 * nl, cudaMemcpyDeviceToHost));
        int worst = 0; long long wt = -1; long long tot[10] = {0};
        for (int i = 0; i < nl; i++) {
            long long t = 0;
            for (int k = 0; k < 10; k++) { if (k < 8 || k == 9) t += h[(size_t)i * 10 + k]; tot[k] += h[(size_t)i * 10 + k]; }
            if (t > wt) { wt = t; worst = i; }
        }
        const char* names[10] = {"price", "normal", "gs", "ratio", "step", "add", "drop", "corridor", "sfc-warp", "warm-start"};
        std::string line = "[qp dbg] kcycles, slowest block " + std::to_string(worst) + ":";
        for (int k = 0; k < 10; k++) line += std::string(" ") + names[k] + " " + std::to_string(h[(size_t)worst * 10 + k] >> 10);
        line += " | mean:";
        for (int k = 0; k < 10; k++) line += std::string(" ") + names[k] + " " + std::to_string((tot[k] / std::max(nl, 1)) >> 10);
        fprintf(stderr, "%s\n", line.c_str());
    }
    st.kernel_launches = e->done_launches;
    st.lsc_pairs = (int64_t)e->done_steps * e->n_plan() * (e->N - 1) * kPairsPerObs;
    st.lsc_pairs_kept = (int64_t)c.kept_pairs;
    st.gjk_iterations = (int64_t)c.gjk_iterations;
    st.qp_rows_priced = (int64_t)c.rows_priced;
    st.qp_iterations = (int64_t)c.qp_iterations;
    st.qp_full_passes = (int64_t)c.full_passes;
    st.qp_warm_tried = (int64_t)c.warm_tried; st.qp_warm_accepted = (int64_t)c.warm_accepted; st.qp_warm_rows = (int64_t)c.warm_rows;
    e->stats_fresh = true;
    return LSCGPU_OK;
}

static int dump_qp_lp(lscgpu_engine* e, int agent, const char* path);

// Points the device word the planning blocks read at `alias` (null: they keep their records on the device). Enqueued on the
// engine stream, only when the value changes.
static int set_host_out(lscgpu_engine* e, lscgpu_agent_out* alias) {
    if (alias == e->host_out_on_device) return LSCGPU_OK;
    CU(cudaStreamSynchronize(e->stream));            // the staging word may still be in flight
    *e->h_host_out = alias;
    CU(cudaMemcpyAsync(e->d_host_out, e->h_host_out, sizeof(void*), cudaMemcpyHostToDevice, e->stream));
    e->host_out_on_device = alias;
    return LSCGPU_OK;
}

extern "C" int lscgpu_replan_batch(lscgpu_engine* e, const lscgpu_agent_in* in, lscgpu_agent_out* out) {
    if (!e || !in || !out) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    // `out` in pinned host memory (cudaHostAlloc / cudaHostRegister, 16-byte aligned), one engine planning every agent: the
    // planning blocks write their records into it as they finish; otherwise one copy after the step
    lscgpu_agent_out* alias = nullptr;
    if (e->zero_copy && !e->comm && e->a0 == 0 && e->a1 == e->N && ((uintptr_t)out & 15) == 0) {
        // looked up on every call (sub-microsecond): the same address may be pinned in one call and pageable in the next
        cudaPointerAttributes at{};
        e->host_out_alias = nullptr;
        if (cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            e->host_out_alias = (lscgpu_agent_out*)at.devicePointer;
        else cudaGetLastError();
        alias = e->host_out_alias;
    }
    if (const int rs = set_host_out(e, alias)) return rs;
    CU(cudaMemcpyAsync(e->d_in, in, sizeof(lscgpu_agent_in) * (size_t)e->N, cudaMemcpyHostToDevice, e->stream));
    const int rc = step_device(e);
    if (rc != LSCGPU_OK) return rc;
    if (!alias) CU(cudaMemcpyAsync(out, e->d_res, sizeof(lscgpu_agent_out) * (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
    const int rf = finish_steps(e);
    if (rf != LSCGPU_OK || e->lp_dump_dir.empty()) return rf;
    // as the reference on an IloException (src/traj_optimizer.cpp:99-101): the model of every failed solve goes to a file
    for (int a = 0; a < e->N; a++)
        if (out[a].qp_status != LSCGPU_QP_OK) {
            const std::string path = e->lp_dump_dir + "/QPmodel_agent" + std::to_string(a) + "_seq" + std::to_string(e->planner_seq) + ".lp";
            const int rd = dump_qp_lp(e, a, path.c_str());
            if (rd != LSCGPU_OK) return rd;
        }
    return LSCGPU_OK;
}

// MultiSyncSimulator::update's hand-over of the planned state (src/multi_sync_simulator.cpp:203: the next current_state is
// the trajectory at t = dt): in[a].position / velocity / acceleration = out[a].next_*; goals are left as they are. Host only.
extern "C" int lscgpu_advance_inputs(const lscgpu_agent_out* out, lscgpu_agent_in* in, int n_agents) {
    if (!out || !in || n_agents < 0) return fail(LSCGPU_ERR_ARG, "null argument");
    for (int a = 0; a < n_agents; a++) {
        std::memcpy(in[a].position, out[a].next_position, sizeof(float) * 3);
        std::memcpy(in[a].velocity, out[a].next_velocity, sizeof(float) * 3);
        std::memcpy(in[a].acceleration, out[a].next_acceleration, sizeof(float) * 3);
    }
    return LSCGPU_OK;
}

extern "C" int lscgpu_replan_resident(lscgpu_engine* e) {
    if (!e) return fail(LSCGPU_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    if (const int rs = set_host_out(e, nullptr)) return rs;
    return step_device(e);
}

extern "C" int lscgpu_synchronize(lscgpu_engine* e) {
    if (!e) return fail(LSCGPU_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    return finish_steps(e);
}

extern "C" int lscgpu_fetch(lscgpu_engine* e, lscgpu_agent_out* out) {
    if (!e || !out) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(out, e->d_res, sizeof(lscgpu_agent_out) * (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return LSCGPU_OK;
}

namespace {
__global__ void k_set_goals(int n, const float* goals, lscgpu_agent_in* in) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n * 3) in[t / 3].goal[t % 3] = goals[t];
}
__global__ void k_set_states(int n, const float* pos, const float* vel, const float* acc, lscgpu_agent_in* in) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n * 3) {
        in[t / 3].position[t % 3] = pos[t];
        in[t / 3].velocity[t % 3] = vel[t];
        in[t / 3].acceleration[t % 3] = acc[t];
    }
}
}  // namespace

extern "C" int lscgpu_set_goals(lscgpu_engine* e, const float* goals) {
    if (!e || !goals) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    const size_t bytes = sizeof(float) * 3 * (size_t)e->N;
    CU(cudaStreamSynchronize(e->stream));
    CU(e->scratch.reserve(bytes));
    float* d = (float*)e->scratch.p;
    CU(cudaMemcpyAsync(d, goals, bytes, cudaMemcpyHostToDevice, e->stream));
    k_set_goals<<<(e->N * 3 + 127) / 128, 128, 0, e->stream>>>(e->N, d, e->d_in);
    CU(cudaStreamSynchronize(e->stream));
    return LSCGPU_OK;
}

extern "C" int lscgpu_set_states(lscgpu_engine* e, const float* pos, const float* vel, const float* acc) {
    if (!e || !pos || !vel || !acc) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    const size_t n3 = 3 * (size_t)e->N;
    CU(cudaStreamSynchronize(e->stream));
    CU(e->scratch.reserve(sizeof(float) * 3 * n3));
    float* d = (float*)e->scratch.p;
    CU(cudaMemcpyAsync(d, pos, sizeof(float) * n3, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d + n3, vel, sizeof(float) * n3, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d + 2 * n3, acc, sizeof(float) * n3, cudaMemcpyHostToDevice, e->stream));
    k_set_states<<<(e->N * 3 + 127) / 128, 128, 0, e->stream>>>(e->N, d, d + n3, d + 2 * n3, e->d_in);
    CU(cudaStreamSynchronize(e->stream));
    return LSCGPU_OK;
}

extern "C" int lscgpu_reset(lscgpu_engine* e) {
    if (!e) return fail(LSCGPU_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    return reset_state(e);
}

extern "C" int lscgpu_set_prev_traj(lscgpu_engine* e, const float* traj, int planner_seq) {
    if (!e || !traj || planner_seq < 0) return fail(LSCGPU_ERR_ARG, "bad argument");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(e->d_traj, traj, sizeof(float) * (size_t)e->N * kTrajFloats, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->planner_seq = planner_seq;
    return LSCGPU_OK;
}

extern "C" int lscgpu_set_sfc(lscgpu_engine* e, const float* boxes, const int32_t* init) {
    if (!e || !boxes || !init) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(e->d_boxes, boxes, sizeof(float) * (size_t)e->N * 30, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->d_init_sfc, init, sizeof(int) * (size_t)e->N, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return LSCGPU_OK;
}

extern "C" int lscgpu_get_sfc(lscgpu_engine* e, float* boxes, int32_t* init) {
    if (!e || !boxes) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(boxes, e->d_boxes, sizeof(float) * (size_t)e->N * 30, cudaMemcpyDeviceToHost, e->stream));
    if (init) CU(cudaMemcpyAsync(init, e->d_init_sfc, sizeof(int) * (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return LSCGPU_OK;
}

extern "C" int lscgpu_get_planner_seq(lscgpu_engine* e) { return e ? e->planner_seq : -1; }

extern "C" int lscgpu_set_slack_collision_weight(lscgpu_engine* e, double w) {
    if (!e || !(w > 0.0)) return fail(LSCGPU_ERR_ARG, "slack_collision_weight must be positive");
    if (e->pending) return fail(LSCGPU_ERR_STATE, "lscgpu_synchronize first");
    e->slack_w = w;
    drop_graph(e);                          // the step graph carries the weight as a kernel argument
    return LSCGPU_OK;
}

extern "C" int lscgpu_get_reset_state(lscgpu_engine* e, uint8_t* reset_ever) {
    if (!e || !reset_ever) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(reset_ever, e->d_reset_ever, (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return LSCGPU_OK;
}

extern "C" int lscgpu_set_reset_state(lscgpu_engine* e, const uint8_t* reset_ever) {
    if (!e || !reset_ever) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    int any = 0;
    for (int a = 0; a < e->N; a++) any |= reset_ever[a] != 0;
    CU(cudaMemcpyAsync(e->d_reset_ever, reset_ever, (size_t)e->N, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->d_any_reset, &any, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return LSCGPU_OK;
}

extern "C" int lscgpu_set_capture_rows(lscgpu_engine* e, int on) {
    if (!e) return fail(LSCGPU_ERR_ARG, "null engine");
    if (e->pending) return fail(LSCGPU_ERR_STATE, "lscgpu_synchronize first");
    e->mirror_rows = on != 0;
    drop_graph(e);
    return LSCGPU_OK;
}

// Rows of `agent` as k_agent_plan built them in the last step (needs lscgpu_set_capture_rows before that step): the kept
// (neighbour, segment) pairs decoded from the mirrored row store — normal = the record's a, d_i = rhs_i - a . o_{m,i} —
// and, for the pairs the exact culling test dropped, the values k_lsc_capture recomputes. kept_out (optional) marks which
// is which.
extern "C" int lscgpu_get_lsc_ex(lscgpu_engine* e, int agent, float* normals, double* d, uint8_t* kept_out) {
    if (!e || !normals || !d || agent < 0 || agent >= e->N) return fail(LSCGPU_ERR_ARG, "bad argument");
    if (e->N < 2) return LSCGPU_OK;
    CU(cudaSetDevice(e->device));
    const size_t n_obs = e->N - 1;
    CU(cudaStreamSynchronize(e->stream));
    Carver cv;
    const size_t o_n = cv.take<float>(n_obs * 15), o_d = cv.take<double>(n_obs * 30);
    CU(e->scratch.reserve(cv.off));
    float* dn = (float*)((char*)e->scratch.p + o_n); double* dd = (double*)((char*)e->scratch.p + o_d);
    launch_lsc_capture(e->N, agent, e->d_pred, e->d_consts, dn, dd, e->stream);
    CU(cudaMemcpyAsync(normals, dn, sizeof(float) * n_obs * 15, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(d, dd, sizeof(double) * n_obs * 30, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (kept_out) std::memset(kept_out, 0, n_obs * kPairsPerObs);
    if (!e->mirror_rows) {
        if (kept_out) return fail(LSCGPU_ERR_STATE, "lscgpu_set_capture_rows(e, 1) before the step to read the production rows");
        return LSCGPU_OK;
    }
    int bi = -1;
    CU(cudaMemcpy(&bi, e->d_block_of + agent, sizeof(int), cudaMemcpyDeviceToHost));
    if (bi < 0) return fail(LSCGPU_ERR_STATE, "the agent was not planned by this engine in the last step");
    int n_kept = 0;
    CU(cudaMemcpy(&n_kept, e->d_kept_count + bi, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<RowRec> rows(n_kept);
    std::vector<int> kept(n_kept);
    std::vector<float> pred((size_t)e->N * kTrajFloats);
    if (n_kept > 0) {
        CU(cudaMemcpy(rows.data(), e->d_rows + (size_t)bi * e->P_pad, sizeof(RowRec) * n_kept, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(kept.data(), e->d_kept + (size_t)bi * e->P_pad, sizeof(int) * n_kept, cudaMemcpyDeviceToHost));
    }
    CU(cudaMemcpy(pred.data(), e->d_pred, sizeof(float) * pred.size(), cudaMemcpyDeviceToHost));
    for (int sidx = 0; sidx < n_kept; sidx++) {
        const int p = kept[sidx] & 0xffffff;          // upper bits: slack code of the slot (qp_core.cuh)
        const int m = p / (int)n_obs, jj = p % (int)n_obs;
        const int j = jj < agent ? jj : jj + 1;
        const RowRec& r = rows[sidx];
        float* no = normals + ((size_t)jj * kM + m) * 3;
        no[0] = r.ax; no[1] = r.ay; no[2] = r.az;
        const double ax = r.ax, ay = r.ay, az = r.az;
        for (int i = 0; i < 6; i++) {
            const float* o = pred.data() + (size_t)j * kTrajFloats + (m * 6 + i) * 3;
            const double t0 = ax * (double)o[0], t1 = ay * (double)o[1], t2 = az * (double)o[2];
            d[((size_t)jj * kM + m) * 6 + i] = r.rhs[i] - ((t0 + t1) + t2);
        }
        if (kept_out) kept_out[(size_t)jj * kM + m] = 1;
    }
    return LSCGPU_OK;
}

extern "C" int lscgpu_get_lsc(lscgpu_engine* e, int agent, float* normals, double* d) {
    return lscgpu_get_lsc_ex(e, agent, normals, d, nullptr);
}

// The QP the engine solved for `agent` in the last step, as a CPLEX LP file (what the reference exports to log/QPmodel.lp
// when a solve fails or param.log is set, src/traj_optimizer.cpp:62-69,99-101): state, goal, terminal segments, SFC window
// and predictions are the step's own (device copies), the LSC rows are recomputed from the predictions.
static int dump_qp_lp(lscgpu_engine* e, int agent, const char* path) {
    const size_t n_obs = (size_t)e->N - 1;
    CU(cudaStreamSynchronize(e->stream));
    std::vector<float> normals(std::max<size_t>(n_obs, 1) * 15), pred((size_t)e->N * kTrajFloats), boxes(30);
    std::vector<double> d(std::max<size_t>(n_obs, 1) * 30);
    if (n_obs > 0) {
        const int rc = lscgpu_get_lsc_ex(e, agent, normals.data(), d.data(), nullptr);
        if (rc != LSCGPU_OK) return rc;
    }
    CU(cudaMemcpy(pred.data(), e->d_pred, sizeof(float) * pred.size(), cudaMemcpyDeviceToHost));
    std::vector<float> points(std::max<size_t>(n_obs, 1) * 90);
    for (size_t o = 0; o < n_obs; o++) {
        const size_t j = (int)o < agent ? o : o + 1;
        std::memcpy(points.data() + o * 90, pred.data() + j * kTrajFloats, sizeof(float) * 90);
    }
    std::vector<unsigned char> ever((size_t)e->N), slack(std::max<size_t>(n_obs, 1), 0);
    CU(cudaMemcpy(ever.data(), e->d_reset_ever, (size_t)e->N, cudaMemcpyDeviceToHost));
    bool any = false;
    for (int a = 0; a < e->N; a++) any |= ever[a] != 0;
    for (size_t o = 0; o < n_obs; o++) slack[o] = ever[agent] || ever[(int)o < agent ? o : o + 1];
    LpProblem p{};
    p.dt = e->prm.dt; p.w = e->prm.control_input_weight; p.wT = e->prm.terminal_weight;
    CU(cudaMemcpy(p.state, e->d_state9 + (size_t)agent * 9, sizeof(double) * 9, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(p.goal, e->d_goal3 + (size_t)agent * 3, sizeof(double) * 3, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&p.ts, e->d_ts + agent, sizeof(int), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++) { p.wmin[k] = e->prm.world_min[k]; p.wmax[k] = e->prm.world_max[k]; }
    if (e->prm.world_use_octomap) {
        CU(cudaMemcpy(boxes.data(), e->d_boxes + (size_t)agent * 30, sizeof(float) * 30, cudaMemcpyDeviceToHost));
        p.boxes = boxes.data();
    }
    p.n_obs = (int)n_obs; p.normals = normals.data(); p.points = points.data(); p.d = d.data();
    p.slack = any ? slack.data() : nullptr; p.slack_w = e->slack_w;
    const lscgpu_agent_const& c = e->consts_host[agent];
    for (int k = 0; k < 3; k++) { p.vmax[k] = c.max_vel[k]; p.amax[k] = c.max_acc[k]; }
    std::string err;
    if (!write_qp_lp(path, p, err)) return fail(LSCGPU_ERR_IO, err);
    return LSCGPU_OK;
}

extern "C" int lscgpu_dump_qp_lp(lscgpu_engine* e, int agent, const char* path) {
    if (!e || !path || agent < 0 || agent >= e->N) return fail(LSCGPU_ERR_ARG, "bad argument");
    if (e->planner_seq < 1) return fail(LSCGPU_ERR_STATE, "no step has been planned yet");
    CU(cudaSetDevice(e->device));
    return dump_qp_lp(e, agent, path);
}

extern "C" int lscgpu_set_lp_dump_dir(lscgpu_engine* e, const char* dir) {
    if (!e) return fail(LSCGPU_ERR_ARG, "null engine");
    e->lp_dump_dir = dir ? dir : "";
    return LSCGPU_OK;
}

extern "C" int lscgpu_get_initial_traj(lscgpu_engine* e, float* out) {
    if (!e || !out) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(out, e->d_pred, sizeof(float) * (size_t)e->N * kTrajFloats, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return LSCGPU_OK;
}

// ---- operator-level entries ------------------------------------------------------------------------------------------
// Device scratch comes from the engine's pool (one allocation that only grows): a batch-of-one TrajOptimizer::solve
// pays four copies in, one kernel chain and one copy out, no cudaMalloc.
static int qp_solve_batch_impl(lscgpu_engine* e, int nb, const int32_t* agent_index, const double* state,
                               const double* goal, const float* sfc, const int32_t* obs_offset,
                               const float* lsc_normal, const float* lsc_point, const double* lsc_d,
                               const uint8_t* obs_slack, double slack_w, double* x,
                               double* cost, int32_t* status, int32_t* iterations, double* eps) {
    if (!e || nb < 0 || !agent_index || !state || !goal || !obs_offset || !x || !cost || !status || !iterations)
        return fail(LSCGPU_ERR_ARG, "null argument");
    if (nb == 0) return LSCGPU_OK;
    for (int b = 0; b < nb; b++) {
        if (agent_index[b] < 0 || agent_index[b] >= e->N) return fail(LSCGPU_ERR_ARG, "agent_index out of range");
        if (obs_offset[b + 1] < obs_offset[b]) return fail(LSCGPU_ERR_ARG, "obs_offset must be non-decreasing");
    }
    if (obs_offset[0] != 0) return fail(LSCGPU_ERR_ARG, "obs_offset[0] must be 0");
    const int total_obs = obs_offset[nb];
    if (total_obs > 0 && (!lsc_normal || !lsc_point || !lsc_d)) return fail(LSCGPU_ERR_ARG, "null LSC arrays");
    CU(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    CU(cudaStreamSynchronize(s));                 // the pool may still be in use by an enqueued step's caller
    const size_t pairs = (size_t)total_obs * kPairsPerObs, pp = std::max<size_t>(pairs, 1);
    Carver cv;
    // results first, contiguous: x | cost | status | iterations come back in one copy
    const size_t o_x = cv.off; cv.off += sizeof(double) * kNv * nb;
    const size_t o_cost = cv.off; cv.off += sizeof(double) * nb;
    const size_t o_status = cv.off; cv.off += sizeof(int) * nb;
    const size_t o_iters = cv.off; cv.off += sizeof(int) * nb;
    const size_t out_bytes = cv.off;
    cv.off = (cv.off + 255) & ~(size_t)255;
    // inputs next, contiguous as well: packed into one pinned staging buffer on the host and uploaded by ONE copy (a
    // batch-of-one TrajOptimizer::solve would otherwise pay seven pageable copies of a few bytes each); the kept counters
    // travel with them as zeros
    const bool slack = obs_slack != nullptr && total_obs > 0;
    const size_t in_begin = cv.off;
    const size_t o_ai = cv.take<int>(nb), o_off = cv.take<int>(nb + 1), o_kc = cv.take<int>(nb);
    const size_t o_state = cv.take<double>((size_t)nb * 9), o_goal = cv.take<double>((size_t)nb * 3);
    const size_t o_sfc = cv.take<float>((size_t)nb * 30);
    const size_t o_n = cv.take<float>(pp * 3), o_p = cv.take<float>(pp * 18), o_d = cv.take<double>(pp * 6);
    const size_t o_slack = cv.take<unsigned char>(std::max(total_obs, 1));
    const size_t in_bytes = cv.off - in_begin;
    const size_t o_ts = cv.take<int>(nb);
    const size_t o_rows = cv.take<RowRec>(pp), o_kept = cv.take<int>(pp), o_safe = cv.take<double>(pp);
    const size_t o_eps = cv.take<double>(pp);
    CU(e->scratch.reserve(cv.off));
    if (e->stage_bytes < in_bytes + out_bytes) {
        if (e->h_stage) cudaFreeHost(e->h_stage);
        e->h_stage = nullptr; e->stage_bytes = 0;
        const size_t want = std::max<size_t>(2 * (in_bytes + out_bytes), 1 << 16);
        CU(cudaHostAlloc((void**)&e->h_stage, want, cudaHostAllocDefault));
        e->stage_bytes = want;
    }
    char* base = (char*)e->scratch.p;
    auto at = [&](size_t o) { return base + o; };
    char* hin = e->h_stage;                           // mirrors [in_begin, in_begin + in_bytes) of the device scratch
    auto hat = [&](size_t o) { return hin + (o - in_begin); };
    std::memcpy(hat(o_ai), agent_index, sizeof(int) * nb);
    std::memcpy(hat(o_off), obs_offset, sizeof(int) * (nb + 1));
    std::memset(hat(o_kc), 0, sizeof(int) * nb);
    std::memcpy(hat(o_state), state, sizeof(double) * 9 * nb);
    std::memcpy(hat(o_goal), goal, sizeof(double) * 3 * nb);
    if (sfc) std::memcpy(hat(o_sfc), sfc, sizeof(float) * 30 * nb);
    if (pairs) {
        std::memcpy(hat(o_n), lsc_normal, sizeof(float) * pairs * 3);
        std::memcpy(hat(o_p), lsc_point, sizeof(float) * pairs * 18);
        std::memcpy(hat(o_d), lsc_d, sizeof(double) * pairs * 6);
    }
    if (slack) std::memcpy(hat(o_slack), obs_slack, (size_t)total_obs);
    CU(cudaMemcpyAsync(at(in_begin), hin, in_bytes, cudaMemcpyHostToDevice, s));
    if (slack) CU(cudaMemsetAsync(at(o_eps), 0, sizeof(double) * pairs, s));
    launch_rows_from_lsc(nb, (int*)at(o_off), total_obs, (float*)at(o_n), (float*)at(o_p), (double*)at(o_d), (RowRec*)at(o_rows),
                         (int*)at(o_kept), (int*)at(o_kc), (double*)at(o_safe), s, slack ? (unsigned char*)at(o_slack) : nullptr);
    launch_terminal_segments(nb, (double*)at(o_state), (double*)at(o_goal), (int*)at(o_ai), e->d_consts, e->prm.dt, (int*)at(o_ts), s);
    QpBatchLaunch ql{};
    ql.n_problems = nb; ql.T = e->d_tables; ql.consts = e->d_consts; ql.agent_index = (int*)at(o_ai);
    ql.state9 = (double*)at(o_state); ql.goal3 = (double*)at(o_goal); ql.ts = (int*)at(o_ts);
    ql.boxes = sfc ? (float*)at(o_sfc) : nullptr;
    for (int k = 0; k < 3; k++) { ql.wmin[k] = e->prm.world_min[k]; ql.wmax[k] = e->prm.world_max[k]; }
    ql.rows = (RowRec*)at(o_rows); ql.obs_offset = (int*)at(o_off);
    ql.kept = (int*)at(o_kept); ql.kept_count = (int*)at(o_kc); ql.safe = (double*)at(o_safe); ql.max_iter = e->max_iter;
    ql.x_out = (double*)at(o_x); ql.cost_out = (double*)at(o_cost); ql.status_out = (int*)at(o_status); ql.iters_out = (int*)at(o_iters);
    ql.slack = slack ? 1 : 0; ql.slack_w = slack_w; ql.eps_out = slack ? (double*)at(o_eps) : nullptr;
    launch_qp_batch(ql, s);
    CU(cudaGetLastError());
    char* hb = e->h_stage + in_bytes;                 // results land in the pinned staging buffer, behind the inputs
    CU(cudaMemcpyAsync(hb, base, out_bytes, cudaMemcpyDeviceToHost, s));
    if (eps) {
        if (slack) CU(cudaMemcpyAsync(eps, at(o_eps), sizeof(double) * pairs, cudaMemcpyDeviceToHost, s));
        else std::memset(eps, 0, sizeof(double) * pairs);
    }
    CU(cudaStreamSynchronize(s));
    std::memcpy(x, hb + o_x, sizeof(double) * kNv * nb);
    std::memcpy(cost, hb + o_cost, sizeof(double) * nb);
    std::memcpy(status, hb + o_status, sizeof(int) * nb);
    std::memcpy(iterations, hb + o_iters, sizeof(int) * nb);
    for (int b = 0; b < nb; b++) status[b] &= 0xff;       // bit 8: more slack variables needed than the kernel holds (reported as MAXITER)
    return LSCGPU_OK;
}

extern "C" int lscgpu_qp_solve_batch(lscgpu_engine* e, int nb, const int32_t* agent_index, const double* state,
                                     const double* goal, const float* sfc, const int32_t* obs_offset,
                                     const float* lsc_normal, const float* lsc_point, const double* lsc_d, double* x,
                                     double* cost, int32_t* status, int32_t* iterations) {
    return qp_solve_batch_impl(e, nb, agent_index, state, goal, sfc, obs_offset, lsc_normal, lsc_point, lsc_d, nullptr, 1.0, x, cost,
                               status, iterations, nullptr);
}

extern "C" int lscgpu_qp_solve_batch_slack(lscgpu_engine* e, int nb, const int32_t* agent_index, const double* state,
                                           const double* goal, const float* sfc, const int32_t* obs_offset,
                                           const float* lsc_normal, const float* lsc_point, const double* lsc_d,
                                           const uint8_t* obs_slack, double* x, double* cost, int32_t* status,
                                           int32_t* iterations, double* eps) {
    if (!e || !obs_offset || nb < 0) return fail(LSCGPU_ERR_ARG, "null argument");
    bool any = false;
    if (obs_slack) for (int o = 0; o < obs_offset[nb]; o++) any |= obs_slack[o] != 0;
    // the reference adds the variables only when the set is not empty (src/traj_optimizer.cpp:317)
    return qp_solve_batch_impl(e, nb, agent_index, state, goal, sfc, obs_offset, lsc_normal, lsc_point, lsc_d, any ? obs_slack : nullptr,
                               e->slack_w, x, cost, status, iterations, eps);
}

extern "C" int lscgpu_gjk_batch(lscgpu_engine* e, int n, const double* hulls, double* v, int32_t* iterations) {
    if (!e || n < 0 || !hulls || !v || !iterations) return fail(LSCGPU_ERR_ARG, "null argument");
    if (n == 0) return LSCGPU_OK;
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    Carver cv;
    const size_t o_h = cv.take<double>((size_t)n * 18), o_v = cv.take<double>((size_t)n * 3), o_i = cv.take<int>(n);
    CU(e->scratch.reserve(cv.off));
    char* base = (char*)e->scratch.p;
    double* dh = (double*)(base + o_h); double* dv = (double*)(base + o_v); int* di = (int*)(base + o_i);
    CU(cudaMemcpyAsync(dh, hulls, sizeof(double) * 18 * n, cudaMemcpyHostToDevice, e->stream));
    launch_gjk_batch(n, dh, dv, di, e->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(v, dv, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(iterations, di, sizeof(int) * n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return LSCGPU_OK;
}

extern "C" int lscgpu_sfc_expand_batch(lscgpu_engine* e, int n, const float* point, const float* goal,
                                       const double* radius, float* box, int32_t* ok) {
    if (!e || n < 0 || !point || !goal || !radius || !box || !ok) return fail(LSCGPU_ERR_ARG, "null argument");
    if (!e->have_map) return fail(LSCGPU_ERR_STATE, "no octomap uploaded");
    if (n == 0) return LSCGPU_OK;
    std::vector<int> sat(n);
    for (int i = 0; i < n; i++) {
        size_t t = 0;
        while (t < e->radii.size() && e->radii[t] != radius[i]) t++;
        if (t == e->radii.size()) return fail(LSCGPU_ERR_ARG, "radius does not belong to any created agent");
        sat[i] = (int)t;
    }
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    Carver cv;
    const size_t o_p = cv.take<float>((size_t)n * 3), o_g = cv.take<float>((size_t)n * 3), o_b = cv.take<float>((size_t)n * 6);
    const size_t o_s = cv.take<int>(n), o_k = cv.take<int>(n);
    CU(e->scratch.reserve(cv.off));
    char* base = (char*)e->scratch.p;
    float* dp = (float*)(base + o_p); float* dg = (float*)(base + o_g); float* db = (float*)(base + o_b);
    int* ds = (int*)(base + o_s); int* dk = (int*)(base + o_k);
    CU(cudaMemcpyAsync(dp, point, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(dg, goal, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(ds, sat.data(), sizeof(int) * n, cudaMemcpyHostToDevice, e->stream));
    SfcLaunch sl{};
    sl.n = n; sl.dm = e->dm; sl.res = e->prm.world_resolution;
    for (int k = 0; k < 3; k++) { sl.wmin[k] = e->prm.world_min[k]; sl.wmax[k] = e->prm.world_max[k]; }
    sl.point = dp; sl.goal = dg; sl.sat_index = ds; sl.box_out = db; sl.ok_out = dk;
    launch_sfc_expand(sl, e->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(box, db, sizeof(float) * 6 * n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(ok, dk, sizeof(int) * n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return LSCGPU_OK;
}

extern "C" int lscgpu_safety_audit(lscgpu_engine* e, double record_time_step, double time_step, double* ratio, int32_t* closest) {
    if (!e || !ratio || !closest) return fail(LSCGPU_ERR_ARG, "null argument");
    if (!(record_time_step > 0.0) || !(time_step > 0.0)) return fail(LSCGPU_ERR_ARG, "record_time_step and time_step must be positive");
    CU(cudaSetDevice(e->device));
    int n_samples = 0;
    for (double ft = 0; ft < time_step - 1e-5; ft += record_time_step) n_samples++;      // src/multi_sync_simulator.cpp:447
    if (n_samples > 4096) return fail(LSCGPU_ERR_ARG, "record_time_step too small");
    if (n_samples > e->audit_samples) {             // scratch kept for the engine's lifetime
        cudaFree(e->d_audit_pos); e->d_audit_pos = nullptr;
        CU(cudaMalloc(&e->d_audit_pos, sizeof(float) * 3 * (size_t)e->N * n_samples));
        e->audit_samples = n_samples;
    }
    if (!e->d_audit_ratio) {
        CU(cudaMalloc(&e->d_audit_ratio, sizeof(double) * e->N));
        CU(cudaMalloc(&e->d_audit_closest, sizeof(int) * e->N));
    }
    launch_safety_audit(e->N, e->d_traj, e->d_consts, e->prm.dt, n_samples, record_time_step, e->d_audit_pos, e->d_audit_ratio,
                        e->d_audit_closest, e->stream);
    CU(cudaMemcpyAsync(ratio, e->d_audit_ratio, sizeof(double) * e->N, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(closest, e->d_audit_closest, sizeof(int) * e->N, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaGetLastError());
    return LSCGPU_OK;
}

extern "C" int lscgpu_get_step_stats(lscgpu_engine* e, lscgpu_step_stats* out) {
    if (!e || !out) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    const int rc = compute_stats(e);
    if (rc != LSCGPU_OK) return rc;
    *out = e->stats;
    return LSCGPU_OK;
}
extern "C" int lscgpu_set_profiling(lscgpu_engine* e, int on) {
    if (!e) return fail(LSCGPU_ERR_ARG, "null engine");
    if (e->pending) return fail(LSCGPU_ERR_STATE, "lscgpu_synchronize first");
    const int rc = compute_stats(e);        // close the books of the finished batch under the mode it ran in
    if (rc != LSCGPU_OK) return rc;
    e->profiling = on != 0;
    return LSCGPU_OK;
}
extern "C" int lscgpu_sm_clock_khz(lscgpu_engine* e) {
    if (!e) return -1;
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device) != cudaSuccess) return -1;
    return khz;
}
extern "C" void* lscgpu_stream(lscgpu_engine* e) { return e ? (void*)e->stream : nullptr; }
