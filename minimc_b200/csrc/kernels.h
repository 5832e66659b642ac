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

// Fission-bank plumbing of one k-eigenvalue generation (device pointers).
struct GenerationIO {
  const BankSite* in = nullptr;              // source bank of this generation, [n_histories]
  BankSite* out = nullptr;                   // unordered fission bank, [capacity]
  unsigned long long capacity = 0;
  unsigned long long* n_out = nullptr;       // sites claimed so far
  uint32_t* child_count = nullptr;           // [n_histories] secondaries of each source particle
  unsigned long long* child_start = nullptr; // [n_histories] where that particle's run starts in `out`
};

// generation == nullptr: fixed source; otherwise one k-eigenvalue generation over generation->in
cudaError_t launch_fixed_source(
    const LaunchConfig& cfg, const char* world_d, const RunSpec& run, const double* bounds_d, BankSite* site_scratch,
    uint2* pending_scratch, unsigned long long* next_history, unsigned long long* scores,
    unsigned long long* square_scores, mmc_counters* counters, const GenerationIO* generation, cudaStream_t stream);

cudaError_t launch_source_bank(const RunSpec& run, BankSite* bank, cudaStream_t stream);
uint32_t bank_scan_blocks(uint64_t n_parents);
cudaError_t launch_order_bank(
    const uint32_t* child_count, const unsigned long long* child_start, uint64_t n_parents, unsigned long long* block_sums,
    const BankSite* unordered, BankSite* ordered, cudaStream_t stream);
cudaError_t launch_resample_bank(
    const BankSite* slice, uint64_t slice_first, uint64_t slice_n, uint64_t m_total, uint64_t n_total, uint64_t first_out,
    uint64_t n_out, BankSite* next, unsigned long long* errors, cudaStream_t stream);

cudaError_t launch_trace(
    const char* world_d, const RunSpec& run, BankSite* site_scratch, mmc_event_record* records, unsigned long long cap,
    unsigned long long* n_records, mmc_counters* counters, cudaStream_t stream);

cudaError_t launch_test_math(int fn, const double* x_d, double* out0_d, double* out1_d, size_t n, cudaStream_t stream);

cudaError_t launch_test_geometry(
    const char* world_d, size_t n, const double* pos_d, const double* dir_d, int32_t* cell_d, int32_t* surface_d,
    double* distance_d, cudaStream_t stream);

// occupancy query for the fused kernel
int max_blocks_per_sm(int tracking, bool continuous_energy, bool generation, size_t smem);

}  // namespace mmc
