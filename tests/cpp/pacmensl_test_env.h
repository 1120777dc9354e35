// Test environment: initialise the library once (device selection, NCCL world when launched multi-rank);
// the counterpart of the reference's tests/pacmensl_test_env.h.
#pragma once
#include "mini_gtest.h"
#include "pacmensl_all.h"

int main(int argc, char *argv[]) {
  int ierr = pacmensl::PACMENSLInit(&argc, &argv, nullptr);
  if (ierr) {
    std::printf("PACMENSLInit failed (no CUDA device?)\n");
    return 2;
  }
  int rc = ::testing::RunAllTests(argc > 1 ? argv[1] : nullptr);
  pacmensl::PACMENSLFinalize();
  return rc;
}
