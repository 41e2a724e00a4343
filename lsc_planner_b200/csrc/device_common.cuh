// Shared device-side declarations of the replanning engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lscgpu.h"
#include "qp_tables.hpp"

namespace lscgpu {

// ---- swarm geometry of one step ----------------------------------------------------------------
constexpr int kTrajFloats = 90;       // [m][i][xyz]
constexpr int kPairsPerObs = 5;       // one LSC normal per (obstacle, segment)

// Per-agent constants, device copy of lscgpu_agent_const.
struct AgentConstDev {
    double radius, downwash, v_nom;
    double vmax[3], amax[3];
    int sat_index;        // which blocked-voxel table (one per distinct radius) the SFC kernel uses
    int pad;
};

// LSC row storage of ONE agent: one 64-byte record per KEPT (obstacle, segment) pair, in kept-list order (slot s),
// so that pricing a pair is two fully used 32-byte sectors:
//   a = LSC normal with z un-scaled (float32, widened to double on use), inv_an = 1/|a|,
//   rhs[i] = d_i + a . o_{m,i}  (i = 0..5)      ->  row:  a . c_{m,i} >= rhs[i]
// kept[s] = dense pair index p = m * n_obs + obstacle (gives the segment m of the slot).
struct __align__(16) RowRec {
    float ax, ay, az, inv_an;
    double rhs[6];
};
static_assert(sizeof(RowRec) == 64, "RowRec must be 64 bytes");

// canonical inequality row ids inside the QP kernel:
//   [0,180)    variable bounds  ((k*5+m)*6+i)*2 + side          side 0: x >= lb, side 1: x <= ub
//   [180,450)  dynamic limits   180 + ((k*5+m)*9+j)*2 + side    j<5 velocity, j>=5 acceleration
//   [450, ..)  LSC              450 + s*6 + i      (s = slot in the agent's kept list)
constexpr int kFixedRows = 450;

struct StepCounters {
    unsigned long long rows_priced;
    unsigned long long qp_iterations;
    unsigned long long full_passes;
    unsigned long long gjk_iterations;
    unsigned long long kept_pairs;
    unsigned long long warm_tried, warm_accepted, warm_rows;     // QP warm starts: attempted, accepted, rows they activated
};

// One slot of the step's gather buffer (what travels between the GPUs of a job): the public result record plus the
// bounds / dynamic-limit rows active at the solution, the next step's warm-start candidates (replicated like the record:
// any rank may plan the agent next).
constexpr int kActSlots = 40;         // [0, 39): row ids, [39]: count
struct __align__(16) GatherSlot {
    lscgpu_agent_out rec;
    unsigned short act[kActSlots];
    int pad[4];
};
static_assert(sizeof(GatherSlot) % 16 == 0, "gather slot must be a multiple of 16 bytes");

// Direct exchange over NVLink peer memory (lscgpu_p2p_attach) instead of the all-gather. Every rank owns one exchange
// buffer: two parities of the gather buffer (a step writes the parity of its epoch, so a rank that is one step ahead never
// overwrites slots a slower peer is still committing) followed by one arrival counter per source rank. A planning block
// stores its finished slot into every peer's buffer through the peer mapping, fences, and bumps the peer's counter for this
// rank; k_commit waits until the counter of a slot's source rank has reached the count this step brings it to.
struct PeerExchange {
    GatherSlot* const* peers;      // null: off. Device array [n_ranks]: every rank's exchange buffer as mapped on this device
    int n_ranks, rank;
    int slots;                     // slots per parity (= block * n_ranks)
    int block;                     // slots per rank
    int n_agents;
    int base_epoch;                // epoch of the first step that uses the exchange
    __host__ __device__ static size_t counters_offset(int slots) { return (sizeof(GatherSlot) * 2 * (size_t)slots + 255) & ~(size_t)255; }
    __device__ int* counters(int r) const {
        return reinterpret_cast<int*>(reinterpret_cast<char*>(peers[r]) + counters_offset(slots));
    }
    __device__ int planned_by(int r) const { return (n_agents - r + n_ranks - 1) / n_ranks; }
};

// ---- small float3/double3 helpers with explicit IEEE roundings ----------------------------------
// The reference's geometry is octomap::point3d = float32 with float arithmetic.
// Where the rounding sequence matters for parity (predictions, normals, margins, terminal-segment count)
// we use the _rn intrinsics so that nvcc cannot contract a*b+c into an FMA.
struct F3 { float x, y, z; };
struct D3 { double x, y, z; };

__device__ __forceinline__ F3 f3_sub(F3 a, F3 b) { return F3{__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)}; }
__device__ __forceinline__ F3 f3_add(F3 a, F3 b) { return F3{__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)}; }
__device__ __forceinline__ F3 f3_scale(F3 a, float s) { return F3{__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)}; }
// octomath::Vector3::dot: float products and sums, returned as double
__device__ __forceinline__ double f3_dot(F3 a, F3 b) {
    return (double)__fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ double d3_dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 d3_sub(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 d3_cross(D3 a, D3 b) {
    return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace lscgpu
