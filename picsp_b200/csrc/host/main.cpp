// picsp_b200_run <input.ini> [--out file.h5] [--steps N] [--device D] [--quiet]
// Drop-in for `./picsp <input.ini>` (src/main.cpp:336-344): same ini keys, same stdout lines, same HDF5 layout.
// Launched once per GPU with RANK / WORLD_SIZE / LOCAL_RANK in the environment (e.g. `torchrun --no-python
// --nproc-per-node N picsp_b200_run input.ini`) it runs sharded over the N GPUs of the node and writes ONE file.
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
    const char *er = std::getenv("RANK"), *ew = std::getenv("WORLD_SIZE"), *el = std::getenv("LOCAL_RANK");
    const int rank = er ? std::atoi(er) : 0, nranks = ew ? std::atoi(ew) : 1;
    if (nranks > 1) {
        if (el) device = std::atoi(el);
        return picsp_host_run_ranked(argv[1], out, steps, quiet, device, rank, nranks) == PICSP_OK ? 0 : EXIT_FAILURE;
    }
    return picsp_host_run(argv[1], out, steps, quiet, device) == PICSP_OK ? 0 : EXIT_FAILURE;
}
