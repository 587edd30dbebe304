// picsp_b200_run <input.ini> [--out file.h5] [--steps N] [--device D] [--quiet]
// Drop-in for `./picsp <input.ini>` (src/main.cpp:336-344): same ini keys, same stdout lines, same HDF5 layout.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "../../../include/picsp_b200.h"
#include "../../../include/picsp_b200_host.h"

int main(int argc, char *argv[]) {
    if (argc < 2) {
        std::cout << "ERROR, at least one argument expected (the input file)." << std::endl;   // main.cpp:340-343
        return EXIT_FAILURE;
    }
    const char *out = nullptr;
    int steps = -1, device = 0, quiet = 0;
    for (int a = 2; a < argc; a++) {
        if (!std::strcmp(argv[a], "--out") && a + 1 < argc) out = argv[++a];
        else if (!std::strcmp(argv[a], "--steps") && a + 1 < argc) steps = std::atoi(argv[++a]);
        else if (!std::strcmp(argv[a], "--device") && a + 1 < argc) device = std::atoi(argv[++a]);
        else if (!std::strcmp(argv[a], "--quiet")) quiet = 1;
    }
    return picsp_host_run(argv[1], out, steps, quiet, device) == PICSP_OK ? 0 : EXIT_FAILURE;
}
