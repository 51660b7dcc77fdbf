// main.cpp -- command line front-end, same flags as the reference executable
// (aquagpusph/ArgumentsManager.cpp:39-191, main.cpp:108-200):
//   AQUAgpusph -i Main.xml -d 2|3 [-l LEVELS] [-q QUEUES] [-v] [-h]
// plus  --root DIR (folder holding resources/), --resolve OUT.xml (write the
// flattened problem and exit: no GPU needed), --steps N (run N steps and stop).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "aquahost.h"

static void usage()
{
    printf("Usage: AQUAgpusph-b200 [Option]...\n"
           "  -l, --log-level=LEVEL  00 debug .. 33 error, one digit per process\n"
           "  -i, --input=INPUT      XML definition input file (Input.xml by default)\n"
           "  -q, --queues=QUEUES    accepted for compatibility (one CUDA stream is used)\n"
           "  -d, --dimensions=DIMS  2 or 3 (3 by default)\n"
           "      --root=DIR         folder that contains resources/ (or AQUAGPUSPH_ROOT)\n"
           "      --resolve=OUT      write the resolved problem as one XML and exit\n"
           "      --steps=N          run N time steps instead of the <Timing> criteria\n"
           "  -v, --version          show the version\n"
           "  -h, --help             show this help\n");
}

int main(int argc, char** argv)
{
    std::string input = "Input.xml", root, resolve, level = "1";
    int dims = 3, steps = -1;
    auto value = [&](int& i, const char* shortf, const char* longf, std::string& out) {
        const std::string a = argv[i];
        if (a == shortf || a == longf) {
            if (i + 1 >= argc) {
                fprintf(stderr, "Option %s requires an argument\n", a.c_str());
                exit(EXIT_FAILURE);
            }
            out = argv[++i];
            return true;
        }
        const std::string pre = std::string(longf) + "=";
        if (!a.compare(0, pre.size(), pre)) {
            out = a.substr(pre.size());
            return true;
        }
        return false;
    };
    for (int i = 1; i < argc; i++) {
        std::string v;
        if (value(i, "-i", "--input", input)) continue;
        if (value(i, "-l", "--log-level", level)) continue;
        if (value(i, "-q", "--queues", v)) continue;
        if (value(i, "-d", "--dimensions", v)) { dims = atoi(v.c_str()); continue; }
        if (value(i, "", "--root", root)) continue;
        if (value(i, "", "--resolve", resolve)) continue;
        if (value(i, "", "--steps", v)) { steps = atoi(v.c_str()); continue; }
        if (!strcmp(argv[i], "-v") || !strcmp(argv[i], "--version")) {
            printf("AQUAgpusph-b200 (sm_100a host for the AQUAgpusph 5.0.4 XML API)\n");
            return EXIT_SUCCESS;
        }
        if (!strcmp(argv[i], "-h") || !strcmp(argv[i], "--help")) {
            usage();
            return EXIT_SUCCESS;
        }
        fprintf(stderr, "Unknown option %s\n", argv[i]);
        usage();
        return EXIT_FAILURE;
    }
    if (dims != 2 && dims != 3) {
        fprintf(stderr, "Only 2D and 3D simulations can be considered\n");
        return EXIT_FAILURE;
    }
    aqh_set_log_level(level.empty() ? 1 : level[0] - '0');
    aqh_sim* sim = nullptr;
    if (!resolve.empty()) {
        if (aqh_parse(input.c_str(), dims, root.c_str(), &sim) ||
            aqh_write_resolved(sim, resolve.c_str())) {
            fprintf(stderr, "ERROR: %s\n", aqh_last_error());
            return EXIT_FAILURE;
        }
        printf("%d tools written to %s\n", aqh_n_tools(sim), resolve.c_str());
        aqh_destroy(sim);
        return EXIT_SUCCESS;
    }
    if (aqh_load(input.c_str(), dims, -1, root.c_str(), 0, 1, &sim)) {
        fprintf(stderr, "ERROR: %s\n", aqh_last_error());
        return EXIT_FAILURE;
    }
    const int rc = steps >= 0 ? (aqh_step(sim, steps) || aqh_sync(sim)) : aqh_run(sim);
    if (rc)
        fprintf(stderr, "ERROR: %s\n", aqh_last_error());
    float t = 0.f;
    unsigned iter = 0;
    aqh_scalar_get(sim, "t", &t, sizeof(t));
    aqh_scalar_get(sim, "iter", &iter, sizeof(iter));
    printf("Simulation finished: iter = %u, t = %g s, %llu CUDA kernels launched\n", iter, t,
           (unsigned long long)aqh_launch_count(sim));
    aqh_destroy(sim);
    return rc ? EXIT_FAILURE : EXIT_SUCCESS;
}
