// The trajectory QP of one agent as a CPLEX LP file, host side: what the reference writes with cplex.exportModel when a
// solve fails or param.log is set (src/traj_optimizer.cpp:62-69,99-101,146-149 -> log/QPmodel.lp). Variables, rows and
// their order are populatebyrow's (src/traj_optimizer.cpp:261-539): x_m_i / y_m_i / z_m_i (+ epsilon_slack_oi_m when
// obs_slack_indices is not empty), rows c1.. = 45 Aeq_base rows (axis-major), SFC rows, LSC rows, dynamic limits, 6 stop
// rows; objective = jerk + terminal (+ slack) with the constant of the terminal cost. Numbers are printed with 15
// significant digits, as CPLEX does.
#pragma once
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "qp_tables.hpp"

namespace lscgpu {

struct LpProblem {
    double dt, w, wT;
    double state[9];                // pos xyz, vel xyz, acc xyz
    double goal[3];
    int ts;                         // terminal segments
    float wmin[3], wmax[3];
    const float* boxes;             // null or [5][6] SFC window (min xyz, max xyz)
    int n_obs;
    const float* normals;           // [n_obs][5][3]
    const float* points;            // [n_obs][5][6][3]  obstacle control points
    const double* d;                // [n_obs][5][6]
    const unsigned char* slack;     // null or [n_obs]: member of obs_slack_indices
    double slack_w;
    double vmax[3], amax[3];
};

namespace detail {
inline std::string lp_var(int k, int m, int i) {
    char buf[32];
    std::snprintf(buf, sizeof buf, "%c_%d_%d", "xyz"[k], m, i);
    return buf;
}
inline void lp_term(std::string& s, double c, const std::string& v, bool& first) {
    if (c == 0.0) return;
    char buf[64];
    const double a = std::fabs(c);
    if (a == 1.0) std::snprintf(buf, sizeof buf, "%s%s", c < 0 ? (first ? "- " : " - ") : (first ? "" : " + "), v.c_str());
    else std::snprintf(buf, sizeof buf, "%s%.15g %s", c < 0 ? (first ? "- " : " - ") : (first ? "" : " + "), a, v.c_str());
    s += buf;
    first = false;
}
}  // namespace detail

inline bool write_qp_lp(const char* path, const LpProblem& p, std::string& err) {
    using namespace detail;
    FILE* f = std::fopen(path, "w");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    double Qb[6][6];
    jerk_cost_block(p.dt, Qb);
    bool any_slack = false;
    if (p.slack) for (int o = 0; o < p.n_obs; o++) any_slack |= p.slack[o] != 0;
    std::fprintf(f, "\\ENCODING=ISO-8859-1\n\\Problem name: lscgpu\n\nMinimize\n obj1: ");
    // linear part and constant of the terminal cost  wT (c_{m,5} - g)^2  (src/traj_optimizer.cpp:352-372)
    std::string s;
    bool first = true;
    double c0 = 0.0;
    for (int k = 0; k < 3; k++)
        for (int m = kM - p.ts; m < kM; m++) { lp_term(s, -2.0 * p.wT * p.goal[k], lp_var(k, m, kN), first); c0 += p.wT * p.goal[k] * p.goal[k]; }
    if (!first) s += " + ";
    s += "[ ";
    first = true;
    for (int k = 0; k < 3; k++)
        for (int m = 0; m < kM; m++)
            for (int i = 0; i < 6; i++)
                for (int j = i; j < 6; j++) {
                    double c = p.w * Qb[i][j] * (i == j ? 1.0 : 2.0);
                    if (i == j && j == kN && m >= kM - p.ts) c += p.wT;
                    c *= 2.0;                                                      // "[ ... ] / 2"
                    if (c == 0.0) continue;
                    const std::string v = i == j ? lp_var(k, m, i) + " ^2" : lp_var(k, m, i) + " * " + lp_var(k, m, j);
                    lp_term(s, c, v, first);
                }
    if (any_slack)                                                                 // :383-390 (every obstacle gets its variables)
        for (int o = 0; o < p.n_obs; o++)
            for (int m = 0; m < kM; m++) {
                char v[48];
                std::snprintf(v, sizeof v, "epsilon_slack_%d_%d ^2", o, m);
                lp_term(s, 2.0 * p.slack_w * ((double)(kM - m) / kM), v, first);
            }
    char tail[64];
    std::snprintf(tail, sizeof tail, " ] / 2 + %.15g\n", c0);
    s += tail;
    std::fputs(s.c_str(), f);
    std::fprintf(f, "Subject To\n");
    int row = 0;
    auto emit = [&](const std::string& lhs, const char* op, double rhs) { std::fprintf(f, " c%d: %s %s %.15g\n", ++row, lhs.c_str(), op, rhs); };
    // equality rows: Aeq_base x = deq, axis-major (:396-407); the stop rows come last (:529-536)
    const Mat A = equality_rows(p.dt);
    for (int k = 0; k < 3; k++)
        for (int r = 0; r < 15; r++) {
            std::string lhs; bool fst = true;
            for (int c = 0; c < kAx; c++) lp_term(lhs, A(r, c), lp_var(k, c / 6, c % 6), fst);
            emit(lhs, "=", r < 3 ? p.state[3 * r + k] : 0.0);
        }
    // SFC rows (:409-434)
    if (p.boxes)
        for (int m = 0; m < kM; m++)
            for (int k = 0; k < 3; k++)
                for (int side = 0; side < 2; side++)
                    for (int i = 0; i < 6; i++) {
                        if (m == 0 && i < kPhi) continue;
                        std::string lhs; bool fst = true;
                        lp_term(lhs, side == 0 ? 1.0 : -1.0, lp_var(k, m, i), fst);
                        emit(lhs, ">=", side == 0 ? (double)p.boxes[m * 6 + k] : -(double)p.boxes[m * 6 + 3 + k]);
                    }
    // LSC rows (:436-466): n . (c - o) - d [- eps] >= 0
    for (int o = 0; o < p.n_obs; o++)
        for (int m = 0; m < kM; m++)
            for (int i = 0; i < 6; i++) {
                if (m == 0 && i < kPhi) continue;
                const float* nv = p.normals + ((size_t)o * kM + m) * 3;
                const float* pt = p.points + (((size_t)o * kM + m) * 6 + i) * 3;
                std::string lhs; bool fst = true;
                double rhs = p.d[((size_t)o * kM + m) * 6 + i];
                for (int k = 0; k < 3; k++) { lp_term(lhs, (double)nv[k], lp_var(k, m, i), fst); rhs += (double)nv[k] * (double)pt[k]; }
                if (p.slack && p.slack[o]) {
                    char v[48];
                    std::snprintf(v, sizeof v, "epsilon_slack_%d_%d", o, m);
                    lp_term(lhs, -1.0, v, fst);
                }
                if (fst) lhs = "0 " + lp_var(0, m, i);
                emit(lhs, ">=", rhs);
            }
    // dynamic limits (:469-525)
    const double vc = std::pow(p.dt, -1) * kN, ac = std::pow(p.dt, -2) * kN * (kN - 1);
    for (int k = 0; k < 3; k++)
        for (int m = 0; m < kM; m++) {
            for (int i = 0; i < kN; i++) {
                if (m == 0 && i < kPhi - 1) continue;
                for (int side = 0; side < 2; side++) {
                    const double sg = side == 0 ? 1.0 : -1.0;
                    std::string lhs; bool fst = true;
                    lp_term(lhs, sg * vc, lp_var(k, m, i + 1), fst); lp_term(lhs, -sg * vc, lp_var(k, m, i), fst);
                    emit(lhs, "<=", p.vmax[k]);
                }
            }
            for (int i = 0; i < kN - 1; i++) {
                if (m == 0 && i < kPhi - 2) continue;
                for (int side = 0; side < 2; side++) {
                    const double sg = side == 0 ? 1.0 : -1.0;
                    std::string lhs; bool fst = true;
                    lp_term(lhs, sg * ac, lp_var(k, m, i + 2), fst); lp_term(lhs, -2.0 * sg * ac, lp_var(k, m, i + 1), fst);
                    lp_term(lhs, sg * ac, lp_var(k, m, i), fst);
                    emit(lhs, "<=", p.amax[k]);
                }
            }
        }
    for (int k = 0; k < 3; k++)
        for (int i = 1; i < kPhi; i++) {
            std::string lhs; bool fst = true;
            lp_term(lhs, 1.0, lp_var(k, kM - 1, kN), fst); lp_term(lhs, -1.0, lp_var(k, kM - 1, kN - i), fst);
            emit(lhs, "=", 0.0);
        }
    std::fprintf(f, "Bounds\n");
    for (int k = 0; k < 3; k++)
        for (int m = 0; m < kM; m++)
            for (int i = 0; i < 6; i++) {
                if (m == 0 && i < kPhi) std::fprintf(f, "      %s Free\n", lp_var(k, m, i).c_str());       // :291-293
                else std::fprintf(f, " %.15g <= %s <= %.15g\n", (double)p.wmin[k], lp_var(k, m, i).c_str(), (double)p.wmax[k]);
            }
    if (any_slack)
        for (int o = 0; o < p.n_obs; o++)
            for (int m = 0; m < kM; m++) std::fprintf(f, " -infinity <= epsilon_slack_%d_%d <= 0\n", o, m);
    std::fprintf(f, "End\n");
    std::fclose(f);
    return true;
}

}  // namespace lscgpu
