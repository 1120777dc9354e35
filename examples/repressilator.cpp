// Repressilator (3 species, 6 reactions, time-invariant): BASELINE config 1.
// Counterpart of the reference's examples/repressilator.cpp (X0 = (21,0,0), t_f = 10, fsp_tol = 1e-4, CVODE rtol 1e-4).
#include "example_common.h"
int main(int argc, char *argv[]) { return run_fsp_example(argc, argv, "repressilator", "repressilator_custom"); }
