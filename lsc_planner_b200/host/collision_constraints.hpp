// Constraint containers handed from the corridor construction to the optimizer: same class surface as the reference's
// include/collision_constraints.hpp / src/collision_constraints.cpp:17-59,81-88,333-397 (visualisation converters and
// the dead two-box SFC::update code are not part of the path).
#pragma once
#include <set>
#include <stdexcept>
#include <vector>

#include "sp_const.hpp"

namespace DynamicPlanning {

// LSC = {c in R^3 | (c - c_obs).dot(normal_vector) - d > 0}
class LSC {
public:
    LSC() = default;
    LSC(const point3d& _obs_control_point, const point3d& _normal_vector, double _d)
        : obs_control_point(_obs_control_point), normal_vector(_normal_vector), d(_d) {}
    point3d obs_control_point;
    point3d normal_vector;
    double d = 0;
};
typedef std::vector<LSC> LSCs;

class Box {
public:
    Box() = default;
    Box(const point3d& _box_min, const point3d& _box_max) : box_min(_box_min), box_max(_box_max) {}

    // six axis half-spaces anchored at the origin (src/collision_constraints.cpp:37-59)
    LSCs convertToLSCs(int dim) const {
        const point3d zero(0, 0, 0);
        LSCs lscs(2 * dim);
        for (int k = 0; k < dim; k++) {
            point3d nv(0, 0, 0);
            nv(k) = 1;
            lscs[2 * k] = LSC(zero, nv, box_min(k));
            nv(k) = -1;
            lscs[2 * k + 1] = LSC(zero, nv, -box_max(k));
        }
        return lscs;
    }
    bool isPointInBox(const point3d& p) const {
        return p.x() > box_min.x() - SP_EPSILON_FLOAT && p.y() > box_min.y() - SP_EPSILON_FLOAT &&
               p.z() > box_min.z() - SP_EPSILON_FLOAT && p.x() < box_max.x() + SP_EPSILON_FLOAT &&
               p.y() < box_max.y() + SP_EPSILON_FLOAT && p.z() < box_max.z() + SP_EPSILON_FLOAT;
    }
    point3d getBoxMin() const { return box_min; }
    point3d getBoxMax() const { return box_max; }

private:
    point3d box_min, box_max;
};

class SFC {
public:
    Box box;
    LSCs lscs;      // always empty on the live path of the reference
    LSCs convertToLSCs(int dim) const {
        LSCs out = box.convertToLSCs(dim);
        out.insert(out.end(), lscs.begin(), lscs.end());
        return out;
    }
    bool update(const Box& _box) { box = _box; lscs.clear(); return true; }
};

typedef std::vector<std::vector<std::vector<LSC>>> RSFCs;   // [obs_idx][segment_idx][control_point_idx]
typedef std::vector<SFC> SFCs;                               // [segment_idx]

class CollisionConstraints {
public:
    CollisionConstraints() = default;

    // keeps the SFC windows across replans (src/collision_constraints.cpp:349-350)
    void initialize(int _N_obs, int _M, int _n, double _dt, std::set<int> _obs_slack_indices) {
        N_obs = _N_obs; M = _M; n = _n; dt = _dt;
        obs_slack_indices = std::move(_obs_slack_indices);
        lscs.assign(N_obs, std::vector<std::vector<LSC>>(M, std::vector<LSC>(n + 1)));
        sfcs.resize(M);
    }
    LSC getLSC(int oi, int m, int i) const { return lscs[oi][m][i]; }
    SFC getSFC(int m) const { return sfcs[m]; }
    size_t getObsSize() const { return lscs.size(); }
    std::set<int> getSlackIndices() const { return obs_slack_indices; }

    void setLSC(int oi, int m, const std::vector<point3d>& obs_control_points, const point3d& normal_vector,
                const std::vector<double>& ds) {
        for (int i = 0; i < n + 1; i++) lscs[oi][m][i] = LSC(obs_control_points[i], normal_vector, ds[i]);
    }
    void setLSC(int oi, int m, const std::vector<point3d>& obs_control_points, const point3d& normal_vector, double d) {
        for (int i = 0; i < n + 1; i++) lscs[oi][m][i] = LSC(obs_control_points[i], normal_vector, d);
    }
    void setSFC(int m, const SFC& sfc) { sfcs[m] = sfc; }
    void setSFC(int m, const Box& box) { sfcs[m].update(box); }

private:
    RSFCs lscs;
    SFCs sfcs;
    std::set<int> obs_slack_indices;
    int N_obs = 0, M = 0, n = 0;
    double dt = 0;
};

}  // namespace DynamicPlanning
