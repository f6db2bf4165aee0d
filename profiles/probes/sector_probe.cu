// How fast can an SM pull W-byte pieces that sit 512 B apart?  (the access pattern of a
// channel slab of a (b, n, n, 128) fp32 tensor).  Each warp issues U independent 128-bit loads
// per lane per iteration; lanes are grouped W/16 per piece.  Prints GB/s of useful bytes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int U>
__global__ void probe(const float4* __restrict__ src, long long n_pieces, int lanes_per_piece,
                      int pieces_total_per_row, float4* __restrict__ sink) {
  // piece p lives at byte p_row * 512 + slab * W  (slab fixed per CTA-group to mimic the kernel)
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int ppw = 32 / lanes_per_piece;          // pieces per warp instruction
  const int sub = lane % lanes_per_piece, pl = lane / lanes_per_piece;
  float4 acc = make_float4(0, 0, 0, 0);
  for (long long base = warp * ppw * U; base < n_pieces; base += nwarps * ppw * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      long long p = base + (long long)u * ppw + pl;
      if (p >= n_pieces) p = n_pieces - 1;
      const long long row = p / pieces_total_per_row, slab = p % pieces_total_per_row;
      // 512 B rows = 32 float4; slab-major order inside the launch: consecutive p of one warp
      // are consecutive ROWS of the same slab (stride 512 B)
      (void)slab;
      v[u] = __ldg(src + row * 32 + (p % pieces_total_per_row) * lanes_per_piece + sub);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x == 123.456f) sink[0] = acc;
}

// slab-major variant: warp walks rows of ONE slab (what a (graph, slab) CTA does)
template <int U>
__global__ void probe_slab(const float4* __restrict__ src, long long n_rows, int lanes_per_piece,
                           float4* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int slabs = 32 / lanes_per_piece;        // slabs per 512 B row
  const int ppw = 32 / lanes_per_piece;
  const int sub = lane % lanes_per_piece, pl = lane / lanes_per_piece;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  // work unit = (chunk of 4096 rows, slab); a warp processes ppw*U rows of its slab per iteration
  const long long chunk_rows = 2048;
  const long long units = (n_rows / chunk_rows) * slabs;
  float4 acc = make_float4(0, 0, 0, 0);
  for (long long unit = warp / 8; unit < units; unit += nwarps / 8) {   // 8 warps share a unit
    const long long chunk = unit / slabs; const int slab = (int)(unit % slabs);
    const int w8 = (int)(warp % 8);
    for (long long r0 = w8 * ppw * U; r0 < chunk_rows; r0 += 8 * ppw * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        long long r = chunk * chunk_rows + r0 + (long long)u * ppw + pl;
        v[u] = __ldg(src + r * 32 + slab * lanes_per_piece + sub);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
  }
  if (acc.x == 123.456f) sink[0] = acc;
}

int main() {
  const long long n_rows = 1LL << 21;            // 2M rows x 512 B = 1 GiB
  float4 *src, *sink;
  cudaMalloc(&src, n_rows * 512);
  cudaMalloc(&sink, 64);
  cudaMemset(src, 0, n_rows * 512);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int W : {32, 64, 128, 512}) {
    const int lpp = W / 16;
    for (int blocks_per_sm : {1, 2}) {
      const int grid = 148 * blocks_per_sm;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        probe_slab<8><<<grid, 256>>>(src, n_rows, lpp, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("piece %3d B, stride 512 B, %d CTA/SM x 8 warps x 8 loads: %7.1f us  %7.1f GB/s\n", W,
             blocks_per_sm, ms * 1e3, (double)n_rows * 512 / (ms * 1e-3) / 1e9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
