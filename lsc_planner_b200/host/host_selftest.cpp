// CPU-only checks of the host-side mirror (no engine calls): mission loader, parameters, constraint containers,
// Bernstein evaluation. Exit code 0 = all passed.
#include <cstdio>
#include <cstdlib>

#include "traj_planner.hpp"

using namespace DynamicPlanning;

#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main(int argc, char** argv) {
    CHECK(argc >= 2);
    Mission ms;
    ms.initialize(argv[1]);
    CHECK(ms.qn == 20 && ms.on == 0);
    CHECK(ms.world_min.x() == -10.f && ms.world_max.z() == 2.5f);
    CHECK(ms.agents[0].radius == 0.15 && ms.agents[0].downwash == 2.0 && ms.agents[3].max_acc[2] == 2.0);
    CHECK(std::fabs(ms.agents[0].start_position.norm() - std::sqrt(65.0)) < 1e-5);     // r = 8 circle at z = 1

    Param p = Param::simulationLaunch();
    CHECK(p.M == 5 && p.n == 5 && p.phi == 3 && p.N_constraint_segments == 5 && p.dt == 0.2);
    p.set("traj/horizon", "2.0");
    CHECK(p.M == 10);
    bool threw = false;
    try { p.set("mode/planner", "orca"); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);
    {   // where prior_based goal planning with an octomap runs (not in the reference): goal/planner=auto|device|host
        Param g = Param::simulationLaunch();
        CHECK(!g.goalPlannerOnDevice(10));                        // auto: a small swarm goes to the host threads
        CHECK(g.goalPlannerOnDevice(80 * 4096));                  // ... a swarm that is large for the host cores does not
        g.set("goal/planner", "device"); CHECK(g.goalPlannerOnDevice(1));
        CHECK(toEngineParams(g, ms).goal_mode == 1 && toEngineParams(g, ms).grid_resolution == 0.25);
        g.set("goal/planner", "host"); CHECK(!g.goalPlannerOnDevice(1000000));
        CHECK(toEngineParams(g, ms).goal_mode == 0);              // the host computes the goals; the engine takes them as given
        bool bad = false;
        try { g.set("goal/planner", "gpu"); } catch (const std::invalid_argument&) { bad = true; }
        CHECK(bad);
    }
    const lscgpu_params ep = toEngineParams(Param::simulationLaunch(), ms);
    CHECK(ep.M == 5 && ep.world_min[0] == -10.f && ep.control_input_weight == 0.01);

    CollisionConstraints cc;
    cc.initialize(3, 5, 5, 0.2, {});
    CHECK(cc.getObsSize() == 3);
    std::vector<point3d> obs(6, point3d(1, 2, 3));
    cc.setLSC(1, 2, obs, point3d(0, 1, 0), std::vector<double>{1, 2, 3, 4, 5, 6});
    CHECK(cc.getLSC(1, 2, 4).d == 5 && cc.getLSC(1, 2, 4).normal_vector.y() == 1.f);
    cc.setSFC(0, Box(point3d(-1, -2, 0), point3d(1, 2, 2)));
    cc.initialize(2, 5, 5, 0.2, {});                       // SFC windows survive re-initialisation
    CHECK(cc.getSFC(0).box.getBoxMax().y() == 2.f);
    const LSCs faces = cc.getSFC(0).convertToLSCs(3);
    CHECK(faces.size() == 6 && faces[0].d == -1.0 && faces[1].d == -1.0 && faces[3].normal_vector.y() == -1.f);
    CHECK(cc.getSFC(0).box.isPointInBox(point3d(0, 0, 1)) && !cc.getSFC(0).box.isPointInBox(point3d(0, 3, 1)));

    // straight-line trajectory: constant velocity 1 m/s along x
    traj_t tr(5, std::vector<point3d>(6));
    for (int m = 0; m < 5; m++) for (int i = 0; i < 6; i++) tr[m][i] = point3d((float)(0.2 * (m + i / 5.0)), 0, 1);
    const State s = getStateFromControlPoints(tr, 0.2, 5, 5, 0.2);
    CHECK(std::fabs(s.position.x() - 0.2f) < 1e-6 && std::fabs(s.velocity.x() - 1.f) < 1e-5 && std::fabs(s.acceleration.x()) < 1e-4);
    const State s2 = getStateFromControlPoints(tr, 0.5, 5, 5, 0.2);
    CHECK(std::fabs(s2.position.x() - 0.5f) < 1e-6);
    std::puts("host selftest ok");
    return 0;
}
