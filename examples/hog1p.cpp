// Hog1p 5-gene-state model with a time-varying signal coefficient (reaction 2): BASELINE config 2.
// Counterpart of the reference's examples/hog1p.cpp (X0 = 0, t_f = 180, fsp_tol = 1e-4).
#include "example_common.h"
int main(int argc, char *argv[]) { return run_fsp_example(argc, argv, "hog1p", nullptr); }
