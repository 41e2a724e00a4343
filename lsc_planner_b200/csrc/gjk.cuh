// GJK distance origin <-> convex hull of 6 points, FP64, one hull per thread, and the LSC (linear safe
// corridor) row construction built on it.
//
// Replaces (reference paths):
//   gjk() + support()                      src/openGJK/openGJK.cpp:633-655,674-780  (outer loop semantics: start
//                                          vertex P[0], strict-improvement support scan, exit tests eps_rel = 1e-10 /
//                                          eps_tot = 1e-12, at most 25 iterations, stop on a full tetrahedron)
//   closestPointsBetweenPointAndConvexHull include/geometry.hpp:364-394
//   normalVectorBetweenPolys               src/traj_planner.cpp:2030-2043
//   generateLSC                            src/traj_planner.cpp:1310-1407 (downwash scaling, margins d_i)
// The closest-point-of-simplex step is NOT openGJK's signed-volume recursion (GPLv3, not restated): it scores the
// features of the (<= 4 vertex) simplex that contain the newest vertex — that vertex, its open edges, its open faces —
// and keeps the nearest one, which is exact for any simplex, including the degenerate (collinear / coincident control
// points) hulls that straight-line predictions produce at the first replanning step.
#pragma once
#include "device_common.cuh"

namespace lscgpu {

struct SimplexD {
    D3 p[4];
    int n;
};

struct NearestFeature {
    D3 v;
    double n2;
    unsigned keep;
};

__device__ __forceinline__ void feature_try(NearestFeature& best, D3 v, unsigned keep) {
    const double n2 = d3_dot(v, v);
    if (n2 < best.n2) { best.v = v; best.n2 = n2; best.keep = keep; }
}

__device__ __forceinline__ void feature_edge(NearestFeature& best, D3 a, D3 b, unsigned keep) {
    const D3 ab = d3_sub(b, a);
    const double den = d3_dot(ab, ab);
    if (den <= 0.0) return;
    const double t = -d3_dot(a, ab) / den;
    if (t <= 0.0 || t >= 1.0) return;           // the end points are scored as vertices
    feature_try(best, D3{a.x + t * ab.x, a.y + t * ab.y, a.z + t * ab.z}, keep);
}

__device__ __forceinline__ void feature_face(NearestFeature& best, D3 a, D3 b, D3 c, unsigned keep) {
    const D3 nrm = d3_cross(d3_sub(b, a), d3_sub(c, a));
    const double nn = d3_dot(nrm, nrm);
    if (nn <= 0.0) return;                       // sliver: its edges cover it
    const double s = d3_dot(nrm, a) / nn;
    const D3 q{nrm.x * s, nrm.y * s, nrm.z * s}; // foot of the origin on the face plane
    const D3 qa = d3_sub(a, q), qb = d3_sub(b, q), qc = d3_sub(c, q);
    if (d3_dot(d3_cross(qb, qc), nrm) <= 0.0) return;
    if (d3_dot(d3_cross(qc, qa), nrm) <= 0.0) return;
    if (d3_dot(d3_cross(qa, qb), nrm) <= 0.0) return;
    feature_try(best, q, keep);
}

// true when the origin lies inside (or on) a non-flat tetrahedron
__device__ __forceinline__ bool tetra_contains_origin(const D3* p) {
#pragma unroll
    for (int f = 0; f < 4; f++) {
        // face f omits vertex f
        const D3 a = p[(f + 1) & 3], b = p[(f + 2) & 3], c = p[(f + 3) & 3], d = p[f];
        const D3 nrm = d3_cross(d3_sub(b, a), d3_sub(c, a));
        const double side_o = -d3_dot(nrm, a);
        const double side_d = d3_dot(nrm, d3_sub(d, a));
        if (side_d == 0.0) return false;
        if (side_o != 0.0 && ((side_o > 0.0) != (side_d > 0.0))) return false;
    }
    return true;
}

__device__ __forceinline__ void simplex_nearest(SimplexD& s, D3& v) {
    const int n = s.n;
    if (n == 4 && tetra_contains_origin(s.p)) { v = D3{0.0, 0.0, 0.0}; return; }
    NearestFeature best;
    best.n2 = INFINITY; best.keep = 0u; best.v = D3{0.0, 0.0, 0.0};
    // The newest vertex w strictly improves on the previous closest point (the outer loop only adds it
    // when v.v - v.w exceeds the tolerances), so the closest point of the enlarged simplex lies in the relative interior
    // of a feature that CONTAINS w: vertex w, the edges (a, w), the faces (a, b, w). Features of the old simplex alone
    // cannot win and are not evaluated (7 instead of 14 candidates for a tetrahedron).
    // (the outer loop keeps the newest vertex at index 0)
    feature_try(best, s.p[0], 1u);
#pragma unroll
    for (int a = 1; a < 4; a++)
        if (a < n) feature_edge(best, s.p[0], s.p[a], 1u | (1u << a));
#pragma unroll
    for (int a = 1; a < 4; a++)
#pragma unroll
        for (int b = a + 1; b < 4; b++)
            if (b < n) feature_face(best, s.p[0], s.p[a], s.p[b], 1u | (1u << a) | (1u << b));
    int m = 0;
#pragma unroll
    for (int a = 0; a < 4; a++)
        if (a < n && (best.keep & (1u << a))) {
            const D3 t = s.p[a];
#pragma unroll
            for (int k = 0; k < 4; k++) if (k == m) s.p[k] = t;
            m++;
        }
    s.n = m;
    v = best.v;
}

__device__ __forceinline__ D3 pick6(const D3* P, int idx) {
    D3 w = P[0];
#pragma unroll
    for (int i = 1; i < 6; i++) if (i == idx) w = P[i];
    return w;
}

// Returns the number of outer iterations; v = closest point of conv(P) to the origin.
__device__ __forceinline__ int gjk_origin_hull6(const D3* P, D3& v) {
    const double eps_rel = 1e-10, eps_tot = 1e-12;
    SimplexD s;
    s.n = 1; s.p[0] = P[0];
    s.p[1] = s.p[2] = s.p[3] = D3{0.0, 0.0, 0.0};
    v = P[0];
    int sup = 0, k = 0;
    double norm2_max = 0.0;
#pragma unroll 1
    do {
        k++;
        double best = -d3_dot(pick6(P, sup), v);
        int better = -1;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const double sc = -d3_dot(P[i], v);
            if (sc > best) { best = sc; better = i; }
        }
        if (better >= 0) sup = better;
        const D3 w = pick6(P, sup);
        const double vv = d3_dot(v, v);
        const double exceed = vv - d3_dot(v, w);
        if (exceed <= eps_rel * vv || exceed < eps_tot) break;
        if (vv < eps_rel * eps_rel) break;
        s.p[3] = s.p[2]; s.p[2] = s.p[1]; s.p[1] = s.p[0]; s.p[0] = w;       // newest vertex first (s.n <= 3 here)
        s.n++;
        simplex_nearest(s, v);
#pragma unroll
        for (int t = 0; t < 4; t++)
            if (t < s.n) norm2_max = fmax(norm2_max, d3_dot(s.p[t], s.p[t]));
        if (d3_dot(v, v) <= eps_tot * eps_tot * norm2_max) break;
    } while (s.n != 4 && k != 25);
    return k;
}

// One (agent, obstacle, segment) LSC: unit normal in downwash-scaled space, margins d_i, un-scaled normal.
//   own/obs: the 6 control points of initial_traj / obstacle prediction of this segment (world coordinates)
struct LscSegment {
    F3 normal;         // as stored by setLSC (z divided by the downwash ratio again, src/traj_planner.cpp:1403)
    double d[6];
    int iterations;
};

// coordinateTransform (include/util.hpp:231-240): z divided by the pair's downwash ratio, in double, stored as float
__device__ __forceinline__ float downwash_scaled_z(float z, double downwash) { return (float)__ddiv_rn((double)z, downwash); }

// own_s/obs_s: the 6 control points with z ALREADY downwash-scaled
__device__ __forceinline__ void lsc_segment_scaled(const F3* own_s, const F3* obs_s, double downwash, double collision_dist,
                                                   LscSegment& out) {
    F3 rel_f[6];
    D3 rel[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        rel_f[i] = f3_sub(own_s[i], obs_s[i]);
        rel[i] = D3{(double)rel_f[i].x, (double)rel_f[i].y, (double)rel_f[i].z};
    }
    D3 v;
    out.iterations = gjk_origin_hull6(rel, v);
    F3 nrm{(float)v.x, (float)v.y, (float)v.z};                // closest_point2 = origin + float3(v)
    const double len = sqrt(f3_dot(nrm, nrm));                 // octomath normalized(): divide by (float)norm()
    if (len > 0.0) {
        const float l = (float)len;
        nrm.x = __fdiv_rn(nrm.x, l); nrm.y = __fdiv_rn(nrm.y, l); nrm.z = __fdiv_rn(nrm.z, l);
    }
#pragma unroll
    for (int i = 0; i < 6; i++) out.d[i] = 0.5 * (collision_dist + f3_dot(rel_f[i], nrm));
    nrm.z = (float)__ddiv_rn((double)nrm.z, downwash);
    out.normal = nrm;
}

__device__ __forceinline__ void lsc_segment(const F3* own, const F3* obs, double downwash, double collision_dist,
                                            LscSegment& out) {
    F3 a[6], o[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        a[i] = own[i]; o[i] = obs[i];
        a[i].z = downwash_scaled_z(a[i].z, downwash);
        o[i].z = downwash_scaled_z(o[i].z, downwash);
    }
    lsc_segment_scaled(a, o, downwash, collision_dist, out);
}

}  // namespace lscgpu
