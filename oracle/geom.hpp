// ORACLE — test infrastructure only (see oracle/README.md). CPU restatement of the reference's
// LSC construction. Nothing under oracle/ is linked into, imported by, or executed from the product
// path (lsc_planner_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker / reported baseline.
//
// Follows (citations relative to /root/reference):
//   src/traj_planner.cpp:1310-1407   generateLSC
//   src/traj_planner.cpp:2030-2043   normalVectorBetweenPolys
//   include/geometry.hpp:364-394     closestPointsBetweenPointAndConvexHull
//   include/util.hpp:191-203,231-240 point3DsToArray, coordinateTransform
//   src/openGJK/openGJK.cpp:674-780  gjk() outer loop: start vertex, exit tests, 25-iteration cap
// The GJK sub-algorithm (closest point of a 1/2/3-simplex to the origin) is this repository's own
// statement (candidate-feature enumeration), not the signed-volume code of openGJK; the result is
// the same mathematical point (the unique min-norm point of the hull). tests/test_oracle_gjk.py pins
// it against the reference's openGJK.cpp compiled into oracle/_ref.
//
// Float semantics of octomap::point3d (octomath::Vector3, upstream): three
// float32; +,-,*(float),/ in float; dot()/norm_sq() evaluated in float, returned as double;
// norm() = sqrt(double); normalize() divides each component by (float)norm().
// Build with -ffp-contract=off so that no FMA contraction changes those roundings.
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

struct F3 {
    float x, y, z;
    float operator()(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline F3 f3(float x, float y, float z) { return F3{x, y, z}; }
inline F3 operator+(F3 a, F3 b) { return F3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline F3 operator-(F3 a, F3 b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline F3 operator*(F3 a, float s) { return F3{a.x * s, a.y * s, a.z * s}; }
inline double dotf(F3 a, F3 b) { return (double)((a.x * b.x + a.y * b.y) + a.z * b.z); }
inline double normf(F3 a) { return std::sqrt(dotf(a, a)); }
inline F3 normalizedf(F3 a) {
    double len = normf(a);
    if (len > 0) {
        float l = (float)len;
        a.x /= l; a.y /= l; a.z /= l;
    }
    return a;
}

// ---------------------------------------------------------------------------------------------
// GJK: min-norm point of conv{P_0..P_{np-1}} (double), i.e. distance origin <-> hull.
// ---------------------------------------------------------------------------------------------
struct D3 { double x, y, z; };
inline D3 dsub(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double ddot(D3 a, D3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline D3 dcross(D3 a, D3 b) {
    return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

struct Simplex { D3 p[4]; int n; };

// Candidate-feature enumeration. Every candidate is a point of the simplex, so the minimum-norm
// candidate is the closest point as soon as the true closest feature is among the candidates; all
// features (vertices, clamped edges, face interiors) are enumerated. `keep` returns which vertices
// support the winner (bit mask) so the simplex can be reduced.
struct Cand { D3 v; double n2; unsigned keep; };

inline void cand_consider(Cand& best, D3 v, unsigned keep) {
    double n2 = ddot(v, v);
    if (n2 < best.n2) { best.v = v; best.n2 = n2; best.keep = keep; }
}

inline void cand_edge(Cand& best, const D3* p, int a, int b) {
    D3 ab = dsub(p[b], p[a]);
    double den = ddot(ab, ab);
    if (den <= 0.0) return;                       // coincident vertices: vertex candidates cover it
    double t = -ddot(p[a], ab) / den;
    if (t <= 0.0 || t >= 1.0) return;              // clamped ends are the vertex candidates
    D3 v{p[a].x + t * ab.x, p[a].y + t * ab.y, p[a].z + t * ab.z};
    cand_consider(best, v, (1u << a) | (1u << b));
}

inline void cand_face(Cand& best, const D3* p, int a, int b, int c) {
    D3 ab = dsub(p[b], p[a]), ac = dsub(p[c], p[a]);
    D3 nrm = dcross(ab, ac);
    double nn = ddot(nrm, nrm);
    if (nn <= 0.0) return;                        // degenerate face: edges cover it
    // projection of the origin on the plane, q = nrm * (nrm.a)/nn ; inside test by barycentrics
    double s = ddot(nrm, p[a]) / nn;
    D3 q{nrm.x * s, nrm.y * s, nrm.z * s};
    D3 qa = dsub(p[a], q), qb = dsub(p[b], q), qc = dsub(p[c], q);
    double wa = ddot(dcross(qb, qc), nrm);
    double wb = ddot(dcross(qc, qa), nrm);
    double wc = ddot(dcross(qa, qb), nrm);
    if (wa <= 0.0 || wb <= 0.0 || wc <= 0.0) return;
    cand_consider(best, q, (1u << a) | (1u << b) | (1u << c));
}

// returns true when the origin is inside the (non-degenerate) tetrahedron
inline bool origin_in_tetra(const D3* p) {
    const int f[4][4] = {{0, 1, 2, 3}, {0, 3, 1, 2}, {0, 2, 3, 1}, {1, 3, 2, 0}};
    for (int k = 0; k < 4; k++) {
        D3 a = p[f[k][0]], b = p[f[k][1]], c = p[f[k][2]], d = p[f[k][3]];
        D3 nrm = dcross(dsub(b, a), dsub(c, a));
        double so = -ddot(nrm, a);               // side of the origin
        double sd = ddot(nrm, dsub(d, a));       // side of the opposite vertex
        if (sd == 0.0) return false;             // flat tetrahedron
        if ((so > 0.0) != (sd > 0.0) && so != 0.0) return false;
    }
    return true;
}

inline void simplex_closest(Simplex& s, D3& v) {
    Cand best; best.n2 = INFINITY; best.keep = 0; best.v = D3{0, 0, 0};
    const int n = s.n;
    if (n == 4 && origin_in_tetra(s.p)) { v = D3{0, 0, 0}; return; }
    for (int a = 0; a < n; a++) cand_consider(best, s.p[a], 1u << a);
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++) cand_edge(best, s.p, a, b);
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++)
            for (int c = b + 1; c < n; c++) cand_face(best, s.p, a, b, c);
    int m = 0;
    for (int a = 0; a < n; a++)
        if (best.keep & (1u << a)) s.p[m++] = s.p[a];
    s.n = m;
    v = best.v;
}

// openGJK.cpp:674-780 outer loop. Returns the number of iterations used.
inline int gjk_origin_hull(const D3* P, int np, D3& v) {
    const double eps_rel = 1e-10, eps_rel2 = eps_rel * eps_rel, eps_tot = 1e-12;   // :40-41
    Simplex s; s.n = 1; s.p[0] = P[0];
    v = P[0];                                                                     // :710-717
    int sup = 0;
    double norm2Wmax = 0;
    int k = 0;
    do {
        k++;
        // support of the hull in direction -v; strict improvement over the previous support (:633-655)
        double maxs = -ddot(P[sup], v);
        int better = -1;
        for (int i = 0; i < np; i++) {
            double sc = -ddot(P[i], v);
            if (sc > maxs) { maxs = sc; better = i; }
        }
        if (better != -1) sup = better;
        D3 w = P[sup];
        double vv = ddot(v, v);
        double exceed = vv - ddot(v, w);                                          // :741
        if (exceed <= eps_rel * vv || exceed < eps_tot) break;                    // :742
        if (vv < eps_rel2) break;                                                 // :746
        s.p[s.n++] = w;                                                           // :752-755
        simplex_closest(s, v);
        for (int j = 0; j < s.n; j++) {                                           // :761-766
            double t = ddot(s.p[j], s.p[j]);
            if (t > norm2Wmax) norm2Wmax = t;
        }
        if (ddot(v, v) <= eps_tot * eps_tot * norm2Wmax) break;                   // :768
    } while (s.n != 4 && k != 25);                                                // :773
    return k;
}

// ---------------------------------------------------------------------------------------------
// LSC for one (agent, neighbour) pair: M normals, M*(n+1) margins.
// own/obs: [M][6] control points of initial_traj / obs_pred_traj (float3).
// ---------------------------------------------------------------------------------------------
struct LscPair {
    F3 normal[5];        // un-scaled back to world coordinates (traj_planner.cpp:1403)
    double d[5][6];      // safety margins (traj_planner.cpp:1389-1394)
    int gjk_iters[5];
};

inline double pair_downwash(double r_i, double dw_i, double r_j, double dw_j) {
    return (dw_i * r_i + dw_j * r_j) / (r_i + r_j);                               // :1339-1341
}

inline void lsc_pair(const F3* own, const F3* obs, int M, double r_i, double dw_i, double r_j,
                     double dw_j, LscPair& out) {
    const double downwash = pair_downwash(r_i, dw_i, r_j, dw_j);
    for (int m = 0; m < M; m++) {
        F3 a_t[6], o_t[6];
        D3 rel[6];
        for (int i = 0; i < 6; i++) {
            a_t[i] = own[m * 6 + i]; o_t[i] = obs[m * 6 + i];
            a_t[i].z = (float)((double)a_t[i].z / downwash);                      // util.hpp:235
            o_t[i].z = (float)((double)o_t[i].z / downwash);
            F3 r = a_t[i] - o_t[i];                                               // :2036
            rel[i] = D3{(double)r.x, (double)r.y, (double)r.z};                   // util.hpp:199
        }
        D3 v;
        out.gjk_iters[m] = gjk_origin_hull(rel, 6, v);
        F3 cp2 = f3(0, 0, 0) + f3((float)v.x, (float)v.y, (float)v.z);             // geometry.hpp:390
        F3 nrm = normalizedf(cp2);                                                // :2041
        const double collision_dist = r_j + r_i;                                  // :1391
        for (int i = 0; i < 6; i++)
            out.d[m][i] = 0.5 * (collision_dist + dotf(a_t[i] - o_t[i], nrm));    // :1392-1393
        nrm.z = (float)((double)nrm.z / downwash);                                // :1403
        out.normal[m] = nrm;
    }
}

}  // namespace orc
