// Thin C entry point around the REFERENCE's own GJK, compiled from the sources where they lie
// under /root/reference (src/openGJK/openGJK.cpp, include/openGJK/openGJK.hpp) into oracle/_ref/.
// Used only by tests/test_oracle_gjk.py to pin the oracle's GJK restatement; never shipped
// (openGJK is GPLv3) and never linked into the product. Mirrors the call made by
// include/geometry.hpp:364-394 (closestPointsBetweenPointAndConvexHull): body 1 = hull, body 2 = point.
#include "openGJK/openGJK.hpp"

extern "C" double ref_gjk_point_hull(const double* hull, int np, const double* point, double* v_out) {
    struct simplex s;
    struct bd bd1, bd2;
    bd1.numpoints = np;
    for (int i = 0; i < np; i++) bd1.coord.push_back({{hull[3 * i], hull[3 * i + 1], hull[3 * i + 2]}});
    bd2.numpoints = 1;
    bd2.coord.push_back({{point[0], point[1], point[2]}});
    double v[3];
    double dd = gjk(bd1, bd2, &s, v);
    v_out[0] = v[0]; v_out[1] = v[1]; v_out[2] = v[2];
    return dd;
}
