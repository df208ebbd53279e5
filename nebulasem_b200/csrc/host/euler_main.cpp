// euler_main.cpp -- `euler ./controls`: drop-in for the reference solver binary (apps/euler/euler.cpp:296-302).
// Reads the same controls / grid / field files from the current directory, runs the time loop on the GPU through
// the C ABI and writes the same rho/U/T/p<k>.{bin,txt} dumps.  One process per GPU/partition.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "nsem_host.h"

int main(int argc, char* argv[]) {
    if (argc < 2 || !std::strcmp(argv[1], "-h")) {
        std::printf("Usage:\n  %s <inputfile>\nOptions:\n  -h          --  Display this message\n\n", argv[0]);
        return argc < 2 ? 1 : 0;
    }
    try {
        nsemh::EulerSolver s;
        std::string ctl = argv[1];
        std::string dir = ".";
        const size_t slash = ctl.find_last_of('/');
        if (slash != std::string::npos && ctl.substr(slash + 1) == "controls") dir = ctl.substr(0, slash);
        else if (ctl != "controls" && ctl != "./controls") {
            std::fprintf(stderr, "euler: the input file must be named `controls`\n");
            return 1;
        }
        s.read_controls(dir);
        const int step = (int)(s.start_step / s.write_interval);
        s.load_mesh(step);
        s.read_fields(step);
        s.setup();
        const char* dev = std::getenv("NSEM_DEVICE");
        s.attach_device(dev ? std::atoi(dev) : 0);
        s.run();
        std::printf("Exiting application run with 1 processes\n");
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "euler: %s\n", e.what());
        return 1;
    }
}
