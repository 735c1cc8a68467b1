// The reference's un-wired harness BeamSimulatorTester::launchRotationTest
// [REF src/beamsimulatortester.cpp:8-41,110-163] written against the drop-in
// header: 3 axes x 3 quarter turns x 1..5 sub-beams of a unit cantilever, each
// checked against the closed form the reference left to a human (P L^3 / 3 EIz).
// Also runs the shipped fixture scenario (README screenshot).  Exit code 0 = all pass.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "beamsimulator.hpp"

namespace {
struct V3 { double x, y, z; };
V3 rot(const V3 &axis, double ang, const V3 &v)
{   // Rodrigues (axis is a unit coordinate axis here)
    const double c = std::cos(ang), s = std::sin(ang);
    const V3 k = axis;
    const V3 kxv{k.y * v.z - k.z * v.y, k.z * v.x - k.x * v.z, k.x * v.y - k.y * v.x};
    const double kv = k.x * v.x + k.y * v.y + k.z * v.z;
    return {v.x * c + kxv.x * s + k.x * kv * (1 - c), v.y * c + kxv.y * s + k.y * kv * (1 - c),
            v.z * c + kxv.z * s + k.z * kv * (1 - c)};
}

SimulationJob cantilever(size_t nBeams, const V3 &X, const V3 &Y)
{
    const float forceMagnitude = 1000, E = 210, nu = 0.3f, b = 0.1f, h = 0.2f;
    SimulationJob j;
    const size_t nNodes = nBeams + 1;
    j.nodes.resize((std::ptrdiff_t)nNodes, 3);
    const double len = std::sqrt(X.x * X.x + X.y * X.y + X.z * X.z);
    for (size_t i = 0; i < nNodes; ++i) {
        j.nodes((std::ptrdiff_t)i, 0) = i * (X.x / len / nBeams);
        j.nodes((std::ptrdiff_t)i, 1) = i * (X.y / len / nBeams);
        j.nodes((std::ptrdiff_t)i, 2) = i * (X.z / len / nBeams);
    }
    j.elements.resize((std::ptrdiff_t)nBeams, 2);
    j.elementalNormals.resize((std::ptrdiff_t)nBeams, 3);
    for (size_t e = 0; e < nBeams; ++e) {
        j.elements((std::ptrdiff_t)e, 0) = (int)e; j.elements((std::ptrdiff_t)e, 1) = (int)e + 1;
        j.elementalNormals((std::ptrdiff_t)e, 0) = Y.x; j.elementalNormals((std::ptrdiff_t)e, 1) = Y.y;
        j.elementalNormals((std::ptrdiff_t)e, 2) = Y.z;
    }
    j.fixedNodes.resize(1); j.fixedNodes(0) = 0;
    const std::ptrdiff_t last = (std::ptrdiff_t)nBeams;
    j.nodalForces = {{last, 0, Y.x * forceMagnitude}, {last, 1, Y.y * forceMagnitude}, {last, 2, Y.z * forceMagnitude}};
    j.beamDimensions.assign(nBeams, BeamDimensions(b, h));
    j.beamMaterial.assign(nBeams, BeamMaterial(nu, E));
    return j;
}
}  // namespace

int main()
{
    mbfea_options o;
    mbfea_options_default(&o);
    o.tol = 1e-13;
    BeamSimulator sim(o);
    sim.setResultsElementalForcesCSVFilepath("");
    sim.setResultsNodalDisplacementCSVFilepath("");
    int fails = 0, runs = 0;
    double axis[3] = {0, 0, 1};
    do {
        const float rotationAngle = (float)M_PI_2;
        const int steps = (int)std::ceil(2 * M_PI / rotationAngle);
        for (int step = 1; step < steps; ++step) {
            const V3 ax{axis[0], axis[1], axis[2]};
            const V3 X = rot(ax, step * (double)rotationAngle, {1, 0, 0}), Y = rot(ax, step * (double)rotationAngle, {0, 1, 0});
            for (size_t nBeams = 1; nBeams < 6; ++nBeams) {
                sim.setSimulation(cantilever(nBeams, X, Y));
                const SimulationResults r = sim.executeSimulation();
                const std::ptrdiff_t tip = (std::ptrdiff_t)nBeams;
                const double defl = r.nodalDisplacements(tip, 0) * Y.x + r.nodalDisplacements(tip, 1) * Y.y +
                                    r.nodalDisplacements(tip, 2) * Y.z;
                const double want = 1000.0 / (3.0 * 14000000.0);
                const bool ok = std::fabs(defl - want) <= 1e-9 * want && std::fabs(r.edgeForces[5](0) + 1000.0) < 1e-5;
                fails += !ok; ++runs;
                if (!ok) std::printf("FAIL axis(%g,%g,%g) step %d beams %zu: tip %.15g want %.15g Mz %.9g\n", axis[0], axis[1],
                                     axis[2], step, nBeams, defl, want, r.edgeForces[5](0));
            }
        }
    } while (std::next_permutation(axis, axis + 3));

    // shipped fixture, README screenshot scenario
    SimulationJob j;
    j.nodes.resize(3, 3);
    for (int i = 0; i < 3; ++i) { j.nodes(i, 0) = i; j.nodes(i, 1) = 0; j.nodes(i, 2) = 0; }
    j.elements.resize(2, 2); j.elements(0, 0) = 0; j.elements(0, 1) = 1; j.elements(1, 0) = 1; j.elements(1, 1) = 2;
    j.elementalNormals.resize(2, 3);
    for (int e = 0; e < 2; ++e) { j.elementalNormals(e, 0) = 0; j.elementalNormals(e, 1) = 1; j.elementalNormals(e, 2) = 0; }
    j.fixedNodes.resize(1); j.fixedNodes(0) = 0;
    j.nodalForces = {{2, 1, 100.0}};
    j.beamDimensions.assign(2, BeamDimensions(0.01f, 0.01f));
    j.beamMaterial.assign(2, BeamMaterial(0.3f, 800));
    sim.setSimulation(j);
    const SimulationResults r = sim.executeSimulation();
    const double want[4] = {-200, 100, -100, 0};
    for (int i = 0; i < 4; ++i)
        if (std::fabs(r.edgeForces[5](i) - want[i]) > 1e-6) { ++fails; std::printf("FAIL Mz[%d]=%.12g\n", i, r.edgeForces[5](i)); }
    ++runs;
    // error conventions
    j.fixedNodes.resize(3); j.fixedNodes(0) = 1; j.fixedNodes(1) = 2; j.fixedNodes(2) = 3;
    try { sim.setSimulation(j); ++fails; std::printf("FAIL: out-of-range fixed vertex accepted\n"); }
    catch (const std::runtime_error &e) { if (std::string(e.what()) != "Vertex index is outside of valid range.") ++fails; }
    std::printf("tester_main: %d runs, %d failures\n", runs, fails);
    return fails ? 1 : 0;
}
