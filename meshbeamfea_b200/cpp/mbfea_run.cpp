// mbfea_run -- command-line front end over the C ABI: runs one of the reference
// application's scenario files (config.json: plyFilename, fixedVertices, forces
// [REF src/utilities.hpp:50-95] + its ASCII PLY edge mesh [REF src/edgemesh.cpp:115-135])
// through libmbfea and writes the two result files the reference writes
// [REF src/beamsimulator.cpp:46-54; src/gui.cpp:363-370].  The reference itself is GUI-only.
//
//   mbfea_run config.json [--ply mesh.ply] [--out DIR] [--tol 1e-8] [--device N] [--quiet]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/mbfea.h"

int main(int argc, char **argv)
{
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s config.json [--ply mesh.ply] [--out DIR] [--tol T] [--device N] [--quiet]\n", argv[0]);
        return 2;
    }
    const char *cfg = argv[1], *ply = nullptr;
    std::string out = "Results";
    mbfea_options o;
    mbfea_options_default(&o);
    bool quiet = false;
    for (int i = 2; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--ply") && i + 1 < argc) ply = argv[++i];
        else if (!std::strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!std::strcmp(argv[i], "--tol") && i + 1 < argc) o.tol = std::atof(argv[++i]);
        else if (!std::strcmp(argv[i], "--device") && i + 1 < argc) o.device = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--quiet")) quiet = true;
        else { std::fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    mbfea_ctx *c = nullptr;
    int rc = mbfea_create(&c, &o);
    if (rc != MBFEA_OK) { std::fprintf(stderr, "mbfea_run: %s\n", mbfea_last_error(nullptr)); return 1; }
    auto die = [&](const char *what, int code) {
        std::fprintf(stderr, "mbfea_run: %s: %s (%s)\n", what, mbfea_last_error(c), mbfea_status_string(code));
        mbfea_destroy(c);
        return 1;
    };
    if ((rc = mbfea_load_scenario(c, cfg, ply)) != MBFEA_OK) return die("loading the scenario", rc);
    rc = mbfea_execute(c);
    if (rc != MBFEA_OK && rc != MBFEA_ENOCONV) return die("simulation", rc);
    const int solve_rc = rc;
    std::string mk = "mkdir -p '" + out + "'";
    if (std::system(mk.c_str()) != 0) { std::fprintf(stderr, "mbfea_run: cannot create %s\n", out.c_str()); mbfea_destroy(c); return 1; }
    // file names of GUI::setSimulation [REF src/gui.cpp:363-370]
    const std::string nd = out + "/nodal_displacements.csv", ef = out + "/elemental_displacements.csv";
    if ((rc = mbfea_write_displacements_csv(c, nd.c_str())) != MBFEA_OK) return die("writing displacements", rc);
    if ((rc = mbfea_write_element_forces_csv(c, ef.c_str())) != MBFEA_OK) return die("writing beam forces", rc);
    mbfea_stats s;
    mbfea_get_stats(c, &s);
    double lo[6], hi[6];
    mbfea_get_force_ranges(c, lo, hi);
    if (!quiet) {
        std::printf("vertices %lld  beams %lld  fixed %lld  free DOF %lld\n", (long long)s.n_vertices, (long long)s.n_beams,
                    (long long)s.n_fixed, (long long)(6 * s.n_block_rows_global));
        std::printf("CG iterations %lld  residual %.3e (recomputed %.3e)  %s\n", (long long)s.iterations, s.rel_residual,
                    s.true_rel_residual, solve_rc == MBFEA_OK ? "converged" : "NOT converged");
        std::printf("ms: assemble %.3f  precondition %.3f  solve %.3f  recover %.3f\n", s.ms_assemble, s.ms_precond, s.ms_pcg,
                    s.ms_recover);
        const char *names[6] = {"N", "Ty", "Tz", "Mx", "My", "Mz"};  // [REF src/gui.cpp:185-186]
        for (int k = 0; k < 6; ++k) std::printf("%-2s in [%.9g, %.9g]\n", names[k], lo[k], hi[k]);
        std::printf("wrote %s and %s\n", nd.c_str(), ef.c_str());
    }
    mbfea_destroy(c);
    return solve_rc == MBFEA_OK ? 0 : 3;
}
