// How fast can W warps of one SM write 512 B rows that are 8 KB apart?  (pad zero-fill pattern)
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void zero_rows(float* out, long long n_rows, long long stride_floats) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = warp; r < n_rows; r += nwarps) {
    float* o = out + r * stride_floats + lane * 4;
    if (MODE == 0) *reinterpret_cast<float4*>(o) = z;
    if (MODE == 1) __stcs(reinterpret_cast<float4*>(o), z);
    if (MODE == 2) __stwt(reinterpret_cast<float4*>(o), z);
    if (MODE == 3) {
      if (lane < 16) asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(out + r * stride_floats + lane * 8), "f"(0.f) : "memory");
    }
  }
}

int main() {
  const long long n_rows = 1 << 17;            // 128K rows
  const long long stride = 2048;               // floats: 8 KB between rows -> 1 GiB span
  float* out;
  cudaMalloc(&out, n_rows * stride * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"st", "st.cs", "st.wt", "st.v8 (16 lanes)"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {1, 2, 4, 8, 16, 32}) {
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) zero_rows<0><<<148, warps * 32>>>(out, n_rows, stride);
        if (mode == 1) zero_rows<1><<<148, warps * 32>>>(out, n_rows, stride);
        if (mode == 2) zero_rows<2><<<148, warps * 32>>>(out, n_rows, stride);
        if (mode == 3) zero_rows<3><<<148, warps * 32>>>(out, n_rows, stride);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      const double per_store_ns = ms * 1e6 / ((double)n_rows / (148.0 * warps));
      printf("%-18s %2d warps/SM: %8.1f us  %7.1f GB/s  %6.1f ns per store per warp\n", names[mode], warps,
             ms * 1e3, n_rows * 512.0 / (ms * 1e-3) / 1e9, per_store_ns);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
