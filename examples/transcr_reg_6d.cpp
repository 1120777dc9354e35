// 6-species transcription regulation with three time-varying reactions: BASELINE config 3.
// Counterpart of the reference's examples/transcr_reg_6d.cpp (X0 = (2,6,0,2,0,0), t_f = 300, fsp_tol = 1e-4, CVODE).
#include "example_common.h"
int main(int argc, char *argv[]) { return run_fsp_example(argc, argv, "transcr_reg_6d", nullptr); }
