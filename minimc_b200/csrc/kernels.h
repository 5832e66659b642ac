// Host-side launch interface of kernels.cu (internal; the public boundary is
// include/minimc_b200.h).
#pragma once

#include <cuda_runtime.h>

#include "../../include/minimc_b200.h"
#include "world_blob.h"

namespace mmc {

constexpr int kThreadsPerBlock = 256;

struct LaunchConfig {
  int blocks;
};

cudaError_t launch_fixed_source(
    const LaunchConfig& cfg, const char* world_d, const RunSpec& run, const double* bounds_d, BankSite* site_scratch,
    uint2* pending_scratch, unsigned long long* next_history, unsigned long long* scores,
    unsigned long long* square_scores, mmc_counters* counters, cudaStream_t stream);

cudaError_t launch_trace(
    const char* world_d, const RunSpec& run, BankSite* site_scratch, mmc_event_record* records, unsigned long long cap,
    unsigned long long* n_records, mmc_counters* counters, cudaStream_t stream);

cudaError_t launch_test_math(int fn, const double* x_d, double* out0_d, double* out1_d, size_t n, cudaStream_t stream);

cudaError_t launch_test_geometry(
    const char* world_d, size_t n, const double* pos_d, const double* dir_d, int32_t* cell_d, int32_t* surface_d,
    double* distance_d, cudaStream_t stream);

// occupancy query for the fused kernel
int max_blocks_per_sm(int tracking, bool continuous_energy, size_t smem);

}  // namespace mmc
