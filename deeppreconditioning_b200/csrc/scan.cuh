// scan.cuh — exclusive prefix sum of int32 (row counts -> rowptr, histograms -> offsets). Integer work: exact.
#pragma once

#include "common.cuh"

namespace dp {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;                            // per thread
constexpr int kScanSpan = kScanThreads * kScanItems;     // items per CTA

size_t scan_workspace_bytes(long long m);
// out[i] = sum_{j<i} in[j], i in [0, m). in == out allowed. ws: scan_workspace_bytes(m).
int exclusive_scan_i32(const int* in, int* out, long long m, void* ws, cudaStream_t stream);

}  // namespace dp
