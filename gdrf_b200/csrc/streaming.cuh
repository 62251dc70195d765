// Row gather for streaming / mini-batch inference.  The reference draws a multiset of
// observation indices on the host every sub-epoch and hands `xs[selection, ...]`, `ws[selection, ...]` to svi.step
// (gdrf/train_script.py:442-460); with the data set resident in HBM that fancy-index becomes one HBM-bound pass:
// 2 * (4 D + 4 V) bytes per selected row, one warp per row, 16-byte accesses on the count rows when V % 4 == 0.
#pragma once
#include "common.cuh"

namespace gdrf {

template <bool VEC4>
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ xs, const int* __restrict__ ws,
                                                     const long long* __restrict__ index, long long n_sel,
                                                     long long n_rows, int D, int V, float* __restrict__ xs_out,
                                                     int* __restrict__ ws_out, int* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n_sel; r += nwarps) {
    long long src = index[r];
    if (src < 0) src += n_rows;                       // numpy / torch negative indexing
    const bool ok = src >= 0 && src < n_rows;
    if (!ok && lane == 0 && status) atomicCAS(status, 0, (int)(r < 0x7ffffffe ? r + 1 : 0x7fffffff));
    if (lane < D) xs_out[r * D + lane] = ok ? xs[src * D + lane] : 0.f;
    if (VEC4) {
      const int4* s = reinterpret_cast<const int4*>(ws + src * V);
      int4* d = reinterpret_cast<int4*>(ws_out + r * V);
      for (int v = lane; v < V / 4; v += 32) d[v] = ok ? __ldg(s + v) : make_int4(0, 0, 0, 0);
    } else {
      for (int v = lane; v < V; v += 32) ws_out[r * V + v] = ok ? ws[src * V + v] : 0;
    }
  }
}

}  // namespace gdrf
