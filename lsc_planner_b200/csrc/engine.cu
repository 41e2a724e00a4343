// C ABI of the replanning engine (include/lscgpu.h): device memory, stream, step sequencing, NCCL exchange.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.hpp"
#include "octomap_bt.hpp"

using namespace lscgpu;

// the ctypes / numpy mirrors in lsc_planner_b200/_capi.py assume these layouts (tests/test_capi_load.py)
static_assert(sizeof(lscgpu_params) == 112 && sizeof(lscgpu_agent_in) == 48 && sizeof(lscgpu_agent_out) == 464 &&
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

struct lscgpu_engine {
    lscgpu_params prm{};
    int N = 0, n_pad = 0;            // agents; n_pad = row pitch of the transposed prediction table
    int a0 = 0, a1 = 0;              // local shard
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_sfc = nullptr;     // k_sfc_expand runs beside k_lsc_build (independent until k_qp_solve)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t stream_aux = nullptr;     // k_qp_order (needs only the previous step's records) beside k_predict
    cudaEvent_t ev_order = nullptr;
    bool overlap_sfc = true, lpt_order = true, qp_debug = false;
    static constexpr int kMaxGroups = 8;
    int pipeline_groups = 2;
    cudaStream_t stream_grp[kMaxGroups] = {};
    cudaEvent_t ev_lsc[kMaxGroups] = {}, ev_qp[kMaxGroups] = {};
    long long* d_dbg = nullptr;
    float* d_audit_pos = nullptr; double* d_audit_ratio = nullptr; int* d_audit_closest = nullptr;   // lscgpu_safety_audit scratch
    int audit_samples = 0;
    int* d_order = nullptr;          // [n_local] local agents, most expensive QP of the previous step first
    int planner_seq = 0;
    bool profiling = false;
    int max_iter = 2000;

    std::vector<lscgpu_agent_const> consts_host;
    std::vector<double> radii;       // distinct radii -> blocked-voxel table index

    // device buffers
    QpTablesDev* d_tables = nullptr;
    AgentConstDev* d_consts = nullptr;
    float2* d_rdw = nullptr;         // [N] (radius, downwash * radius) in float: all the culling pass of k_lsc_build needs
    lscgpu_agent_in* d_in = nullptr;
    lscgpu_agent_out* d_out = nullptr;     // [n_out] gather buffer
    int n_out = 0;
    float *d_traj = nullptr, *d_pred = nullptr, *d_predT = nullptr, *d_predZs = nullptr, *d_boxes = nullptr;
    double *d_state9 = nullptr, *d_goal3 = nullptr, *d_last_cost = nullptr;
    int *d_ts = nullptr, *d_flags = nullptr, *d_init_sfc = nullptr, *d_goal_kind = nullptr;
    // row store of the local shard
    RowRec* d_rows = nullptr;
    int P_pad = 0, n_rows_alloc = 0;
    int *d_kept = nullptr, *d_kept_count = nullptr;
    double* d_safe = nullptr;
    float4* d_sphere = nullptr;      // [5][n_pad]
    float* d_reach = nullptr;        // [N][5]
    StepCounters* d_counters = nullptr;
    // map
    bool have_map = false;
    DistMapDev dm{};
    int64_t n_occupied = 0;
    // exchange
    NcclComm comm = nullptr;
    int rank = 0, n_ranks = 1, block = 0;
    // instrumentation: steps enqueued since the last synchronize
    struct StepEvents { cudaEvent_t ev[9]; };   // begin, predict|, sfc[ (side stream) ]sfc, lsc|, qp[ ]qp, exchange|, commit|
    std::vector<StepEvents> ev_pool;    // per-kernel events of every pending step (profiling mode)
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> step_ev;   // begin / end of every pending step
    int pending = 0, pending_launches = 0;
    lscgpu_step_stats stats{};
};

static void free_rows(lscgpu_engine* e) {
    cudaFree(e->d_rows); cudaFree(e->d_safe);
    cudaFree(e->d_kept); cudaFree(e->d_kept_count); cudaFree(e->d_order);
    e->d_order = nullptr;
    e->d_rows = nullptr; e->d_safe = nullptr;
    e->d_kept = nullptr; e->d_kept_count = nullptr;
    e->n_rows_alloc = 0;
}

static int alloc_rows(lscgpu_engine* e) {
    const int n_local = e->a1 - e->a0;
    if (n_local <= e->n_rows_alloc) return LSCGPU_OK;
    free_rows(e);
    const int P = kPairsPerObs * std::max(e->N - 1, 1);
    e->P_pad = (P + 31) / 32 * 32;
    CU(cudaMalloc(&e->d_rows, sizeof(RowRec) * (size_t)n_local * e->P_pad));
    CU(cudaMalloc(&e->d_safe, sizeof(double) * (size_t)n_local * e->P_pad));
    CU(cudaMalloc(&e->d_kept, sizeof(int) * (size_t)n_local * e->P_pad));
    CU(cudaMalloc(&e->d_kept_count, sizeof(int) * (size_t)n_local));
    CU(cudaMalloc(&e->d_order, sizeof(int) * (size_t)n_local));
    e->n_rows_alloc = n_local;
    return LSCGPU_OK;
}

extern "C" const char* lscgpu_last_error(void) { return g_error.c_str(); }
extern "C" int lscgpu_version(void) { return 100; }

extern "C" void lscgpu_destroy(lscgpu_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);      // every side stream is joined into this one at the end of a step
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    free_rows(e);
    cudaFree(e->d_rdw); cudaFree(e->d_audit_pos); cudaFree(e->d_audit_ratio); cudaFree(e->d_audit_closest); cudaFree(e->d_dbg);
    cudaFree(e->d_tables); cudaFree(e->d_consts); cudaFree(e->d_in); cudaFree(e->d_out); cudaFree(e->d_traj);
    cudaFree(e->d_pred); cudaFree(e->d_predT); cudaFree(e->d_predZs); cudaFree(e->d_boxes); cudaFree(e->d_state9); cudaFree(e->d_goal3);
    cudaFree(e->d_last_cost); cudaFree(e->d_ts); cudaFree(e->d_flags); cudaFree(e->d_init_sfc); cudaFree(e->d_goal_kind); cudaFree(e->d_counters);
    cudaFree(e->dm.sqdist); cudaFree(e->dm.sat); cudaFree(e->d_sphere); cudaFree(e->d_reach);
    for (auto& se : e->ev_pool) for (auto& ev : se.ev) if (ev) cudaEventDestroy(ev);
    for (auto& pr : e->step_ev) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    if (e->ev_begin) cudaEventDestroy(e->ev_begin);
    if (e->ev_end) cudaEventDestroy(e->ev_end);
    for (int g = 0; g < lscgpu_engine::kMaxGroups; g++) {
        if (e->ev_lsc[g]) cudaEventDestroy(e->ev_lsc[g]);
        if (e->ev_qp[g]) cudaEventDestroy(e->ev_qp[g]);
        if (e->stream_grp[g]) cudaStreamDestroy(e->stream_grp[g]);
    }
    if (e->ev_order) cudaEventDestroy(e->ev_order);
    if (e->stream_aux) cudaStreamDestroy(e->stream_aux);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->stream_sfc) cudaStreamDestroy(e->stream_sfc);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static int reset_state(lscgpu_engine* e) {
    const size_t N = e->N;
    CU(cudaMemsetAsync(e->d_traj, 0, sizeof(float) * N * kTrajFloats, e->stream));      // src/traj_planner.cpp:36-39
    CU(cudaMemsetAsync(e->d_boxes, 0, sizeof(float) * N * 30, e->stream));
    CU(cudaMemsetAsync(e->d_last_cost, 0, sizeof(double) * N, e->stream));
    CU(cudaMemsetAsync(e->d_in, 0, sizeof(lscgpu_agent_in) * N, e->stream));
    CU(cudaMemsetAsync(e->d_out, 0, sizeof(lscgpu_agent_out) * (size_t)e->n_out, e->stream));
    std::vector<int> ones(N, 1);                                                         // flag_initialize_sfc, :49
    CU(cudaMemcpyAsync(e->d_init_sfc, ones.data(), sizeof(int) * N, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->planner_seq = 0;
    return LSCGPU_OK;
}

extern "C" int lscgpu_create(const lscgpu_params* p, int n_agents, const lscgpu_agent_const* agents, int device,
                             lscgpu_engine** out) {
    if (!p || !agents || !out || n_agents < 1) return fail(LSCGPU_ERR_ARG, "null argument or n_agents < 1");
    if (p->goal_mode != 0 && p->goal_mode != 1) return fail(LSCGPU_ERR_ARG, "goal_mode must be 0 (static) or 1 (prior_based)");
    if (p->goal_mode == 1 && p->world_use_octomap)
        return fail(LSCGPU_ERR_ARG, "goal_mode 1 (prior_based on the GPU) needs world_use_octomap = 0: with an octomap the goals "
                                    "come from the host grid planner (lsc_planner_b200/host/grid_based_planner.hpp)");
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
    if (prop.major < 10) return fail(LSCGPU_ERR_CUDA, "device is not sm_100 class; kernels are built for sm_100a only");

    lscgpu_engine* e = new lscgpu_engine;
    e->prm = *p; e->N = n_agents; e->device = device;
    e->n_pad = (n_agents + 31) / 32 * 32;
    e->a0 = 0; e->a1 = n_agents; e->block = n_agents; e->n_out = n_agents;
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
    CUB(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&e->stream_sfc, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&e->stream_aux, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&e->ev_order, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    if (const char* v = getenv("LSCGPU_OVERLAP_SFC")) e->overlap_sfc = atoi(v) != 0;
    if (const char* v = getenv("LSCGPU_LPT_ORDER")) e->lpt_order = atoi(v) != 0;
    e->qp_debug = getenv("LSCGPU_QP_DEBUG") != nullptr;
    if (const char* v = getenv("LSCGPU_PIPELINE_GROUPS")) e->pipeline_groups = std::min(std::max(atoi(v), 1), (int)lscgpu_engine::kMaxGroups);
    for (int g = 0; g < lscgpu_engine::kMaxGroups; g++) {
        CUB(cudaStreamCreateWithFlags(&e->stream_grp[g], cudaStreamNonBlocking));
        CUB(cudaEventCreateWithFlags(&e->ev_lsc[g], cudaEventDisableTiming));
        CUB(cudaEventCreateWithFlags(&e->ev_qp[g], cudaEventDisableTiming));
    }
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
    CUB(cudaMalloc(&e->d_out, sizeof(lscgpu_agent_out) * N));
    CUB(cudaMalloc(&e->d_traj, sizeof(float) * N * kTrajFloats));
    CUB(cudaMalloc(&e->d_pred, sizeof(float) * N * kTrajFloats));
    CUB(cudaMalloc(&e->d_predT, sizeof(float) * (size_t)kTrajFloats * e->n_pad));
    CUB(cudaMemset(e->d_predT, 0, sizeof(float) * (size_t)kTrajFloats * e->n_pad));
    CUB(cudaMalloc(&e->d_predZs, sizeof(float) * (size_t)30 * e->n_pad));
    CUB(cudaMemset(e->d_predZs, 0, sizeof(float) * (size_t)30 * e->n_pad));
    CUB(cudaMalloc(&e->d_sphere, sizeof(float4) * (size_t)kM * e->n_pad));
    CUB(cudaMemset(e->d_sphere, 0, sizeof(float4) * (size_t)kM * e->n_pad));
    CUB(cudaMalloc(&e->d_reach, sizeof(float) * N * kM));
    CUB(cudaMalloc(&e->d_boxes, sizeof(float) * N * 30));
    CUB(cudaMalloc(&e->d_state9, sizeof(double) * N * 9));
    CUB(cudaMalloc(&e->d_goal3, sizeof(double) * N * 3));
    CUB(cudaMalloc(&e->d_last_cost, sizeof(double) * N));
    CUB(cudaMalloc(&e->d_ts, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_flags, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_goal_kind, sizeof(int) * N));
    CUB(cudaMemset(e->d_goal_kind, 0, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_init_sfc, sizeof(int) * N));
    CUB(cudaMalloc(&e->d_counters, sizeof(StepCounters)));
    CUB(cudaMemset(e->d_counters, 0, sizeof(StepCounters)));
    int rc = alloc_rows(e);
    if (rc != LSCGPU_OK) return bail(rc);
    rc = reset_state(e);
    if (rc != LSCGPU_OK) return bail(rc);
    *out = e;
    return LSCGPU_OK;
#undef CUB
}

// ---- octomap ---------------------------------------------------------------------------------------------------
static int coord_to_key(double c, double res) { return (int)std::floor((1.0 / res) * c); }   // OcTree::coordToKey - 32768

static int build_map(lscgpu_engine* e, const int32_t* keys, int n) {
    CU(cudaSetDevice(e->device));
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
    int32_t* d_keys = nullptr; int* d_thr = nullptr; uint8_t *sa = nullptr, *sb = nullptr;
    CU(cudaMalloc(&e->dm.sqdist, total));
    CU(cudaMalloc(&e->dm.sat, tab * sizeof(int) * thr.size()));
    CU(cudaMalloc(&sa, total)); CU(cudaMalloc(&sb, total));
    CU(cudaMalloc(&d_thr, sizeof(int) * thr.size()));
    CU(cudaMemcpyAsync(d_thr, thr.data(), sizeof(int) * thr.size(), cudaMemcpyHostToDevice, e->stream));
    if (n > 0) {
        CU(cudaMalloc(&d_keys, sizeof(int32_t) * 3 * (size_t)n));
        CU(cudaMemcpyAsync(d_keys, keys, sizeof(int32_t) * 3 * (size_t)n, cudaMemcpyHostToDevice, e->stream));
    }
    launch_edt_build(d_keys, n, e->dm, d_thr, e->dm.n_tables, sa, sb, e->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(d_keys); cudaFree(d_thr); cudaFree(sa); cudaFree(sb);
    int64_t occ = 0;
    for (int i = 0; i < n; i++) {
        bool in = true;
        for (int k = 0; k < 3; k++) { const int c = keys[3 * i + k] - e->dm.off[k]; if (c < 0 || c >= e->dm.size[k]) in = false; }
        occ += in;
    }
    e->n_occupied = occ;
    e->have_map = true;
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
    if (e->comm) return fail(LSCGPU_ERR_STATE, "shard is fixed by lscgpu_nccl_init");
    CU(cudaSetDevice(e->device));
    e->a0 = a0; e->a1 = a1;
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
    std::string err;
    if (!load_nccl(err)) return fail(LSCGPU_ERR_NCCL, err);
    CU(cudaSetDevice(e->device));
    NcclId id;
    std::memcpy(id.internal, id_bytes, 128);
    const int rc = g_nccl.CommInitRank(&e->comm, n_ranks, id, rank);
    if (rc != 0) { e->comm = nullptr; return fail(LSCGPU_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?")); }
    e->rank = rank; e->n_ranks = n_ranks;
    e->block = (e->N + n_ranks - 1) / n_ranks;           // contiguous blocks; the last ranks may own fewer (or no) agents
    e->a0 = std::min(e->N, rank * e->block);
    e->a1 = std::min(e->N, e->a0 + e->block);
    if (e->block * n_ranks > e->n_out) {
        cudaFree(e->d_out);
        e->n_out = e->block * n_ranks;
        CU(cudaMalloc(&e->d_out, sizeof(lscgpu_agent_out) * (size_t)e->n_out));
        CU(cudaMemset(e->d_out, 0, sizeof(lscgpu_agent_out) * (size_t)e->n_out));
    }
    return alloc_rows(e);
}

// ---- the step ----------------------------------------------------------------------------------------------------
static int step_device(lscgpu_engine* e) {
    if (e->prm.world_use_octomap && !e->have_map)
        return fail(LSCGPU_ERR_STATE, "world_use_octomap is set but no octomap was uploaded (lscgpu_set_octomap_*)");
    cudaStream_t s = e->stream;
    const int n_local = e->a1 - e->a0;
    e->planner_seq++;                                       // src/traj_planner.cpp:127
    int launches = 0;
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
    if (e->pending == 0) CU(cudaEventRecord(e->ev_begin, s));
    CU(cudaEventRecord(e->step_ev[e->pending].first, s));
    if (prof) CU(cudaEventRecord(ev[0], s));
    if (n_local > 0) CU(cudaMemsetAsync(e->d_kept_count, 0, sizeof(int) * n_local, s));
    CU(cudaMemsetAsync(e->d_flags, 0, sizeof(int) * (size_t)e->N, s));      // k_predict and k_sfc_expand OR their bits in
    const bool do_sfc = e->prm.world_use_octomap && n_local > 0;
    const bool side = e->overlap_sfc && !e->profiling;      // profiling mode serialises the kernels
    // scheduling order of this step's LSC / QP blocks from the cost of the previous step's solves (still in d_out)
    const bool ordered = e->lpt_order && e->planner_seq > 1 && n_local > 1;
    if (side && (do_sfc || ordered)) CU(cudaEventRecord(e->ev_fork, s));   // k_sfc_expand / k_qp_order need only the inputs
    if (ordered) {
        cudaStream_t so = side ? e->stream_aux : s;
        if (side) CU(cudaStreamWaitEvent(so, e->ev_fork, 0));
        launch_qp_order(n_local, e->a0, e->d_out, e->d_order, so); launches++;
        if (side) CU(cudaEventRecord(e->ev_order, so));
    }

    PredictLaunch pl{};
    pl.n_agents = e->N; pl.n_pad = e->n_pad; pl.planner_seq = e->planner_seq;
    pl.dt = e->prm.dt; pl.reset_threshold = e->prm.reset_threshold;
    pl.in = e->d_in; pl.prev_traj = e->d_traj; pl.consts = e->d_consts;
    pl.pred = e->d_pred; pl.predT = e->d_predT; pl.predZs = e->d_predZs; pl.state9 = e->d_state9; pl.goal3 = e->d_goal3;
    pl.ts = e->d_ts; pl.flags = e->d_flags; pl.sphere = e->d_sphere; pl.reach = e->d_reach;
    launch_predict(pl, s); launches++;
    if (e->prm.goal_mode == 1) {
        GoalLaunch gl{};
        gl.n_agents = e->N; gl.dt = e->prm.dt; gl.goal_threshold = e->prm.goal_threshold; gl.goal_radius = e->prm.goal_radius;
        gl.priority_dist_threshold = e->prm.priority_dist_threshold;
        gl.in = e->d_in; gl.prev_traj = e->d_traj; gl.pred = e->d_pred; gl.consts = e->d_consts;
        gl.goal3 = e->d_goal3; gl.ts = e->d_ts; gl.goal_kind = e->d_goal_kind;
        launch_goal_plan(gl, s); launches++;
    }
    if (prof) CU(cudaEventRecord(ev[1], s));

    // k_sfc_expand depends on the step's inputs only (state, goal, previous trajectory, its own windows) and k_qp_solve is
    // its only consumer: it runs on a side stream beside k_predict and k_lsc_build
    cudaStream_t ss = side ? e->stream_sfc : s;
    if (do_sfc) {
        if (side) CU(cudaStreamWaitEvent(ss, e->ev_fork, 0));
        if (prof) CU(cudaEventRecord(ev[2], ss));
        SfcLaunch sl{};
        sl.n = n_local; sl.dm = e->dm; sl.res = e->prm.world_resolution;
        for (int k = 0; k < 3; k++) { sl.wmin[k] = e->prm.world_min[k]; sl.wmax[k] = e->prm.world_max[k]; }
        sl.mode = 0; sl.agent_base = e->a0;
        sl.in = e->d_in; sl.prev_traj = e->d_traj; sl.consts = e->d_consts;
        sl.boxes = e->d_boxes; sl.init_sfc = e->d_init_sfc; sl.flags = e->d_flags;
        launch_sfc_expand(sl, ss); launches++;
        if (prof) CU(cudaEventRecord(ev[3], ss));
        if (side) CU(cudaEventRecord(e->ev_join, ss));
    } else if (prof) {
        CU(cudaEventRecord(ev[2], s)); CU(cudaEventRecord(ev[3], s));
    }

    // LSC + QP, pipelined over groups of agents in scheduling order: k_lsc_build of group g+1 runs beside k_qp_solve of
    // group g (own stream per group), so the long solves of the few crowded agents (first in the LPT order) overlap
    // with the corridor construction of everybody else. Profiling mode serialises everything (one group, per-kernel
    // events); results do not depend on the grouping.
    int groups = 1;
    if (!prof && ordered && e->pipeline_groups > 1 && n_local >= 128 * e->pipeline_groups) groups = e->pipeline_groups;
    LscLaunch ll{};
    ll.n_agents = e->N; ll.n_pad = e->n_pad; ll.a0 = e->a0; ll.n_local = n_local;
    ll.order = ordered ? e->d_order : nullptr;
    ll.pred = e->d_pred; ll.predT = e->d_predT; ll.predZs = e->d_predZs; ll.consts = e->d_consts; ll.rdw = e->d_rdw; ll.T = e->d_tables;
    ll.state9 = e->d_state9; ll.goal3 = e->d_goal3; ll.ts = e->d_ts;
    ll.sphere = e->d_sphere; ll.reach = e->d_reach;
    ll.rows = e->d_rows; ll.P_pad = e->P_pad;
    ll.kept = e->d_kept; ll.kept_count = e->d_kept_count; ll.safe = e->d_safe;
    ll.counters = e->d_counters;
    QpLaunch ql{};
    ql.T = e->d_tables; ql.consts = e->d_consts;
    ql.order = ordered ? e->d_order : nullptr;
    ql.agent_index = nullptr; ql.agent_base = e->a0;
    ql.state9 = e->d_state9; ql.goal3 = e->d_goal3; ql.ts = e->d_ts;
    ql.boxes = e->prm.world_use_octomap ? e->d_boxes : nullptr;
    for (int k = 0; k < 3; k++) { ql.wmin[k] = e->prm.world_min[k]; ql.wmax[k] = e->prm.world_max[k]; }
    ql.rows = e->d_rows; ql.obs_offset = nullptr; ql.n_obs = e->N - 1; ql.P_pad = e->P_pad;
    ql.kept = e->d_kept; ql.kept_count = e->d_kept_count; ql.safe = e->d_safe; ql.max_iter = e->max_iter;
    ql.out = e->d_out; ql.prev_traj = e->d_traj; ql.last_cost = e->d_last_cost; ql.flags = e->d_flags;
    ql.goal_kind = e->d_goal_kind;
    ql.counters = e->d_counters;
    if (e->qp_debug) { if (!e->d_dbg) CU(cudaMalloc(&e->d_dbg, sizeof(long long) * 8 * (size_t)e->N)); ql.dbg = e->d_dbg; }
    if (groups == 1) {
        if (ordered && side) CU(cudaStreamWaitEvent(s, e->ev_order, 0));
        if (n_local > 0 && e->N > 1) { ll.first = 0; ll.count = n_local; launch_lsc_build(ll, s); launches++; }
        if (prof) CU(cudaEventRecord(ev[4], s));
        if (do_sfc && side) CU(cudaStreamWaitEvent(s, e->ev_join, 0));
        if (prof) CU(cudaEventRecord(ev[5], s));
        if (n_local > 0) { ql.first = 0; ql.n_problems = n_local; launch_qp_solve(ql, s); launches++; }
    } else {
        // group sizes: the first (most expensive) groups are the smallest, so their solves start early
        int first = 0;
        CU(cudaEventRecord(e->ev_lsc[0], s));           // predictions (and the order) are ready
        for (int g = 0; g < groups; g++) {
            const int rest = n_local - first;
            const int count = g == groups - 1 ? rest : std::max(64, (int)(n_local * (g + 1.0) / (groups * (groups + 1) / 2.0)));
            const int cnt = std::min(count, rest);
            if (cnt <= 0) { CU(cudaEventRecord(e->ev_qp[g], s)); continue; }
            cudaStream_t sg = e->stream_grp[g];
            CU(cudaStreamWaitEvent(sg, e->ev_lsc[0], 0));
            if (ordered && side) CU(cudaStreamWaitEvent(sg, e->ev_order, 0));
            ll.first = first; ll.count = cnt; launch_lsc_build(ll, sg); launches++;
            if (do_sfc && side) CU(cudaStreamWaitEvent(sg, e->ev_join, 0));
            ql.first = first; ql.n_problems = cnt; launch_qp_solve(ql, sg); launches++;
            CU(cudaEventRecord(e->ev_qp[g], sg));
            first += cnt;
        }
        for (int g = 0; g < groups; g++) CU(cudaStreamWaitEvent(s, e->ev_qp[g], 0));
    }
    if (prof) CU(cudaEventRecord(ev[6], s));

    if (e->comm && e->n_ranks > 1) {
        // in-place all-gather: every rank's block of results lands in every replica
        const size_t bytes = sizeof(lscgpu_agent_out) * (size_t)e->block;
        const int rc = g_nccl.AllGather((const char*)e->d_out + bytes * e->rank, e->d_out, bytes, /*ncclInt8*/ 0, e->comm, s);
        if (rc != 0) return fail(LSCGPU_ERR_NCCL, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    }
    if (prof) CU(cudaEventRecord(ev[7], s));
    launch_commit(e->N, e->d_out, e->d_traj, e->d_in, s); launches++;
    if (prof) CU(cudaEventRecord(ev[8], s));
    CU(cudaEventRecord(e->step_ev[e->pending].second, s));
    CU(cudaEventRecord(e->ev_end, s));
    CU(cudaGetLastError());
    e->pending++;
    e->pending_launches += launches;
    return LSCGPU_OK;
}

static int finish_steps(lscgpu_engine* e) {
    CU(cudaStreamSynchronize(e->stream));
    lscgpu_step_stats& st = e->stats;
    st = lscgpu_step_stats{};
    if (e->pending == 0) return LSCGPU_OK;
    StepCounters c;
    CU(cudaMemcpyAsync(&c, e->d_counters, sizeof c, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemsetAsync(e->d_counters, 0, sizeof(StepCounters), e->stream));
    CU(cudaStreamSynchronize(e->stream));
    st.steps = e->pending;
    CU(cudaEventElapsedTime(&st.ms_total, e->ev_begin, e->ev_end));
    for (int i = 0; i < e->pending; i++) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e->step_ev[i].first, e->step_ev[i].second));
        st.ms_steps += ms;
    }
    if (e->profiling) {
        for (int i = 0; i < e->pending && i < (int)e->ev_pool.size(); i++) {
            cudaEvent_t* ev = e->ev_pool[i].ev;
            float ms[8];
            for (int k = 0; k < 8; k++) CU(cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]));
            // ev: begin, predict|, sfc[ ]sfc (side stream when overlapped), lsc|, join|, qp|, exchange|, commit|
            st.ms_predict += ms[0]; st.ms_sfc += ms[2]; st.ms_qp += ms[5]; st.ms_exchange += ms[6]; st.ms_commit += ms[7];
            float lsc = 0.f;
            CU(cudaEventElapsedTime(&lsc, ev[3], ev[4]));
            st.ms_lsc += lsc;
        }
    }
    if (e->d_dbg && e->qp_debug) {
        const int nl = e->a1 - e->a0;
        std::vector<long long> h((size_t)nl * 8);
        CU(cudaMemcpy(h.data(), e->d_dbg, sizeof(long long) * 8 * nl, cudaMemcpyDeviceToHost));
        int worst = 0; long long wt = -1; long long tot[8] = {0};
        for (int i = 0; i < nl; i++) { long long t = 0; for (int k = 0; k < 8; k++) { t += h[(size_t)i * 8 + k]; tot[k] += h[(size_t)i * 8 + k]; } if (t > wt) { wt = t; worst = i; } }
        fprintf(stderr, "[qp dbg] worst local agent %d kcycles: price %lld nv %lld gs %lld solve %lld xupd %lld add %lld drop %lld | mean: price %lld nv %lld gs %lld solve %lld xupd %lld add %lld drop %lld\n", worst,
                h[(size_t)worst*8]>>10, h[(size_t)worst*8+1]>>10, h[(size_t)worst*8+2]>>10, h[(size_t)worst*8+3]>>10, h[(size_t)worst*8+4]>>10, h[(size_t)worst*8+5]>>10, h[(size_t)worst*8+6]>>10,
                tot[0]/nl>>10, tot[1]/nl>>10, tot[2]/nl>>10, tot[3]/nl>>10, tot[4]/nl>>10, tot[5]/nl>>10, tot[6]/nl>>10);
    }
    st.kernel_launches = e->pending_launches;
    st.lsc_pairs = (int64_t)e->pending * (e->a1 - e->a0) * (e->N - 1) * kPairsPerObs;
    st.lsc_pairs_kept = (int64_t)c.kept_pairs;
    st.gjk_iterations = (int64_t)c.gjk_iterations;
    st.qp_rows_priced = (int64_t)c.rows_priced;
    st.qp_iterations = (int64_t)c.qp_iterations;
    st.qp_full_passes = (int64_t)c.full_passes;
    e->pending = 0; e->pending_launches = 0;
    return LSCGPU_OK;
}

extern "C" int lscgpu_replan_batch(lscgpu_engine* e, const lscgpu_agent_in* in, lscgpu_agent_out* out) {
    if (!e || !in || !out) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(e->d_in, in, sizeof(lscgpu_agent_in) * (size_t)e->N, cudaMemcpyHostToDevice, e->stream));
    const int rc = step_device(e);
    if (rc != LSCGPU_OK) return rc;
    CU(cudaMemcpyAsync(out, e->d_out, sizeof(lscgpu_agent_out) * (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
    return finish_steps(e);
}

extern "C" int lscgpu_replan_resident(lscgpu_engine* e) {
    if (!e) return fail(LSCGPU_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
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
    CU(cudaMemcpyAsync(out, e->d_out, sizeof(lscgpu_agent_out) * (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
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
    float* d = nullptr;
    const size_t bytes = sizeof(float) * 3 * (size_t)e->N;
    CU(cudaMalloc(&d, bytes));
    CU(cudaMemcpyAsync(d, goals, bytes, cudaMemcpyHostToDevice, e->stream));
    k_set_goals<<<(e->N * 3 + 127) / 128, 128, 0, e->stream>>>(e->N, d, e->d_in);
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(d);
    return LSCGPU_OK;
}

extern "C" int lscgpu_set_states(lscgpu_engine* e, const float* pos, const float* vel, const float* acc) {
    if (!e || !pos || !vel || !acc) return fail(LSCGPU_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    float* d = nullptr;
    const size_t n3 = 3 * (size_t)e->N;
    CU(cudaMalloc(&d, sizeof(float) * 3 * n3));
    CU(cudaMemcpyAsync(d, pos, sizeof(float) * n3, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d + n3, vel, sizeof(float) * n3, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d + 2 * n3, acc, sizeof(float) * n3, cudaMemcpyHostToDevice, e->stream));
    k_set_states<<<(e->N * 3 + 127) / 128, 128, 0, e->stream>>>(e->N, d, d + n3, d + 2 * n3, e->d_in);
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(d);
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

extern "C" int lscgpu_get_lsc(lscgpu_engine* e, int agent, float* normals, double* d) {
    if (!e || !normals || !d || agent < 0 || agent >= e->N) return fail(LSCGPU_ERR_ARG, "bad argument");
    if (e->N < 2) return LSCGPU_OK;
    CU(cudaSetDevice(e->device));
    const size_t n_obs = e->N - 1;
    float* dn = nullptr; double* dd = nullptr;
    CU(cudaMalloc(&dn, sizeof(float) * n_obs * 15));
    CU(cudaMalloc(&dd, sizeof(double) * n_obs * 30));
    launch_lsc_capture(e->N, agent, e->d_pred, e->d_consts, dn, dd, e->stream);
    CU(cudaMemcpyAsync(normals, dn, sizeof(float) * n_obs * 15, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(d, dd, sizeof(double) * n_obs * 30, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(dn); cudaFree(dd);
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
namespace {
struct DevBuf {
    std::vector<void*> ptrs;
    ~DevBuf() { for (void* p : ptrs) cudaFree(p); }
    template <class T>
    cudaError_t get(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t rc = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (rc == cudaSuccess) ptrs.push_back(p);
        *out = (T*)p;
        return rc;
    }
};
}  // namespace

extern "C" int lscgpu_qp_solve_batch(lscgpu_engine* e, int nb, const int32_t* agent_index, const double* state,
                                     const double* goal, const float* sfc, const int32_t* obs_offset,
                                     const float* lsc_normal, const float* lsc_point, const double* lsc_d, double* x,
                                     double* cost, int32_t* status, int32_t* iterations) {
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
    DevBuf B;
    int *d_ai, *d_off, *d_ts, *d_status, *d_iters, *d_kept, *d_kc;
    double *d_state, *d_goal, *d_d, *d_x, *d_cost, *d_safe;
    float *d_sfc = nullptr, *d_n, *d_p;
    RowRec* d_rows;
    const size_t pairs = (size_t)total_obs * kPairsPerObs;
    CU(B.get(&d_ai, nb)); CU(B.get(&d_off, nb + 1)); CU(B.get(&d_ts, nb)); CU(B.get(&d_safe, pairs));
    CU(B.get(&d_status, nb)); CU(B.get(&d_iters, nb));
    CU(B.get(&d_kept, pairs)); CU(B.get(&d_kc, nb));
    CU(cudaMemsetAsync(d_kc, 0, sizeof(int) * nb, s));
    CU(B.get(&d_state, (size_t)nb * 9)); CU(B.get(&d_goal, (size_t)nb * 3)); CU(B.get(&d_d, pairs * 6));
    CU(B.get(&d_x, (size_t)nb * kNv)); CU(B.get(&d_cost, nb));
    CU(B.get(&d_n, pairs * 3)); CU(B.get(&d_p, pairs * 18)); CU(B.get(&d_rows, pairs));
    CU(cudaMemcpyAsync(d_ai, agent_index, sizeof(int) * nb, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_off, obs_offset, sizeof(int) * (nb + 1), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_state, state, sizeof(double) * 9 * nb, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_goal, goal, sizeof(double) * 3 * nb, cudaMemcpyHostToDevice, s));
    if (sfc) {
        CU(B.get(&d_sfc, (size_t)nb * 30));
        CU(cudaMemcpyAsync(d_sfc, sfc, sizeof(float) * 30 * nb, cudaMemcpyHostToDevice, s));
    }
    if (pairs) {
        CU(cudaMemcpyAsync(d_n, lsc_normal, sizeof(float) * pairs * 3, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(d_p, lsc_point, sizeof(float) * pairs * 18, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(d_d, lsc_d, sizeof(double) * pairs * 6, cudaMemcpyHostToDevice, s));
    }
    launch_rows_from_lsc(nb, d_off, total_obs, d_n, d_p, d_d, d_rows, d_kept, d_kc, d_safe, s);
    launch_terminal_segments(nb, d_state, d_goal, d_ai, e->d_consts, e->prm.dt, d_ts, s);
    QpLaunch ql{};
    ql.n_problems = nb; ql.T = e->d_tables; ql.consts = e->d_consts; ql.agent_index = d_ai; ql.agent_base = 0;
    ql.state9 = d_state; ql.goal3 = d_goal; ql.ts = d_ts; ql.boxes = d_sfc;
    for (int k = 0; k < 3; k++) { ql.wmin[k] = e->prm.world_min[k]; ql.wmax[k] = e->prm.world_max[k]; }
    ql.rows = d_rows; ql.obs_offset = d_off; ql.n_obs = 0; ql.P_pad = (int)pairs;
    ql.kept = d_kept; ql.kept_count = d_kc; ql.safe = d_safe; ql.max_iter = e->max_iter;
    ql.x_out = d_x; ql.cost_out = d_cost; ql.status_out = d_status; ql.iters_out = d_iters;
    ql.out = nullptr; ql.counters = nullptr;
    launch_qp_solve(ql, s);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(x, d_x, sizeof(double) * kNv * nb, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(cost, d_cost, sizeof(double) * nb, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(status, d_status, sizeof(int) * nb, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(iterations, d_iters, sizeof(int) * nb, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return LSCGPU_OK;
}

extern "C" int lscgpu_gjk_batch(lscgpu_engine* e, int n, const double* hulls, double* v, int32_t* iterations) {
    if (!e || n < 0 || !hulls || !v || !iterations) return fail(LSCGPU_ERR_ARG, "null argument");
    if (n == 0) return LSCGPU_OK;
    CU(cudaSetDevice(e->device));
    DevBuf B;
    double *dh, *dv; int* di;
    CU(B.get(&dh, (size_t)n * 18)); CU(B.get(&dv, (size_t)n * 3)); CU(B.get(&di, n));
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
    DevBuf B;
    float *dp, *dg, *db; int *ds, *dk;
    CU(B.get(&dp, (size_t)n * 3)); CU(B.get(&dg, (size_t)n * 3)); CU(B.get(&db, (size_t)n * 6));
    CU(B.get(&ds, n)); CU(B.get(&dk, n));
    CU(cudaMemcpyAsync(dp, point, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(dg, goal, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(ds, sat.data(), sizeof(int) * n, cudaMemcpyHostToDevice, e->stream));
    SfcLaunch sl{};
    sl.n = n; sl.dm = e->dm; sl.res = e->prm.world_resolution;
    for (int k = 0; k < 3; k++) { sl.wmin[k] = e->prm.world_min[k]; sl.wmax[k] = e->prm.world_max[k]; }
    sl.mode = 1; sl.point = dp; sl.goal = dg; sl.sat_index = ds; sl.box_out = db; sl.ok_out = dk;
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
    *out = e->stats;
    return LSCGPU_OK;
}
extern "C" int lscgpu_set_profiling(lscgpu_engine* e, int on) {
    if (!e) return fail(LSCGPU_ERR_ARG, "null engine");
    if (e->pending) return fail(LSCGPU_ERR_STATE, "lscgpu_synchronize first");
    e->profiling = on != 0;
    return LSCGPU_OK;
}
extern "C" void* lscgpu_stream(lscgpu_engine* e) { return e ? (void*)e->stream : nullptr; }
