// MaskedTensor kernels: 2-FWL contraction (CUDA-core path), masked pooling, masked fill.
//
// Layout is the reference's: (b, n, n, dense) with the dense (channel) axis contiguous
// (backend/MaTensor.py:34-111).  Every kernel maps lanes to channels, so each warp
// request is one contiguous 128 B line.
#include <math.h>

#include "common.cuh"

namespace pgh {

// ------------------------------------------------------------------ mamamm, algo 0
// One CTA per (graph b, 32-channel slab).  lane = channel; each warp owns 4x4 output
// tiles (i, k) in registers and walks j.  Exact fp32 (sequential j order).
constexpr int kTI = 4, kTK = 4;

__global__ void __launch_bounds__(256)
mamamm_simt_kernel(const float* __restrict__ A, long long sAi, long long sAj,
                   const float* __restrict__ B, long long sBj, long long sBk,
                   const unsigned char* __restrict__ mask, const int* __restrict__ ext, int n_i,
                   int n_j, int n_k, int dense, float* __restrict__ out) {
  const int b = blockIdx.x;
  // valid extents: operand pads are zero, so the j loop can stop at the valid extent
  const int nj_valid = ext ? min(max(ext[3 * b + 1], 0), n_j) : n_j;
  const int ch = blockIdx.y * 32 + (threadIdx.x & 31);
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const bool ch_ok = ch < dense;
  const float* Ab = A + (size_t)b * n_i * n_j * dense + ch;
  const float* Bb = B + (size_t)b * n_j * n_k * dense + ch;
  float* Ob = out + (size_t)b * n_i * n_k * dense + ch;
  const unsigned char* Mb = mask + (size_t)b * n_i * n_k;
  const int tiles_i = (n_i + kTI - 1) / kTI, tiles_k = (n_k + kTK - 1) / kTK;
  for (int tile = warp; tile < tiles_i * tiles_k; tile += nwarps) {
    const int i0 = (tile / tiles_k) * kTI, k0 = (tile % tiles_k) * kTK;
    float acc[kTI][kTK];
#pragma unroll
    for (int x = 0; x < kTI; ++x)
#pragma unroll
      for (int y = 0; y < kTK; ++y) acc[x][y] = 0.f;
    if (ch_ok) {
      for (int j = 0; j < nj_valid; ++j) {
        float a[kTI], bb[kTK];
#pragma unroll
        for (int x = 0; x < kTI; ++x)
          a[x] = (i0 + x < n_i) ? __ldg(Ab + ((size_t)(i0 + x) * sAi + (size_t)j * sAj) * dense) : 0.f;
#pragma unroll
        for (int y = 0; y < kTK; ++y)
          bb[y] = (k0 + y < n_k) ? __ldg(Bb + ((size_t)j * sBj + (size_t)(k0 + y) * sBk) * dense) : 0.f;
#pragma unroll
        for (int x = 0; x < kTI; ++x)
#pragma unroll
          for (int y = 0; y < kTK; ++y) acc[x][y] = fmaf(a[x], bb[y], acc[x][y]);
      }
#pragma unroll
      for (int x = 0; x < kTI; ++x)
#pragma unroll
        for (int y = 0; y < kTK; ++y)
          if (i0 + x < n_i && k0 + y < n_k) {
            const size_t o = (size_t)(i0 + x) * n_k + (k0 + y);
            Ob[o * dense] = Mb[o] ? acc[x][y] : 0.f;
          }
    }
  }
}

// ------------------------------------------------------------------ masked pooling
// data (b, n1, n2, dense).  One thread per (b, kept index, channel); the reduced extent
// is walked with stride so that a warp always touches one contiguous line.
template <int AGGR>
__global__ void masked_pool_kernel(const float* __restrict__ data,
                                   const unsigned char* __restrict__ mask, int n1, int n2,
                                   int dense, int red, float* __restrict__ out,
                                   unsigned char* __restrict__ out_mask, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % dense);
  long long rest = idx / dense;
  // kept extent: red==1 -> n2, red==2 -> n1, red==3 -> 1
  const int nkeep = red == 1 ? n2 : red == 2 ? n1 : 1;
  const int keep = (int)(rest % nkeep);
  const long long b = rest / nkeep;
  const int nred = red == 1 ? n1 : red == 2 ? n2 : n1 * n2;
  float acc = AGGR == PGH_MAX ? -INFINITY : AGGR == PGH_MIN ? INFINITY : 0.f;
  int cnt = 0;
  for (int r = 0; r < nred; ++r) {
    long long pos;  // position in (n1, n2)
    if (red == 1) pos = (long long)r * n2 + keep;
    else if (red == 2) pos = (long long)keep * n2 + r;
    else pos = r;
    pos += b * n1 * n2;
    if (mask[pos]) {
      const float v = __ldg(data + pos * dense + ch);
      ++cnt;
      if (AGGR == PGH_MAX) acc = fmaxf(acc, v);
      else if (AGGR == PGH_MIN) acc = fminf(acc, v);
      else acc += v;
    }
  }
  if (cnt == 0) acc = 0.f;
  else if (AGGR == PGH_MEAN) acc = acc / (float)cnt;
  if ((AGGR == PGH_MAX || AGGR == PGH_MIN) && isinf(acc)) acc = 0.f;  // filterinf, MaTensor.py:8-31
  out[idx] = acc;
  if (out_mask && ch == 0) out_mask[rest] = cnt > 0;
}

template <int AGGR>
__global__ void masked_pool_bwd_kernel(const float* __restrict__ data,
                                       const unsigned char* __restrict__ mask,
                                       const float* __restrict__ outp,
                                       const float* __restrict__ g_out, int n1, int n2,
                                       int dense, int red, float* __restrict__ g_data,
                                       long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % dense);
  long long rest = idx / dense;
  const int nkeep = red == 1 ? n2 : red == 2 ? n1 : 1;
  const int keep = (int)(rest % nkeep);
  const long long b = rest / nkeep;
  const int nred = red == 1 ? n1 : red == 2 ? n2 : n1 * n2;
  const float g = g_out[idx];
  const float o = (AGGR == PGH_MAX || AGGR == PGH_MIN) ? outp[idx] : 0.f;
  int cnt = 0;
  for (int r = 0; r < nred; ++r) {
    long long pos = red == 1 ? (long long)r * n2 + keep : red == 2 ? (long long)keep * n2 + r : r;
    pos += b * n1 * n2;
    if (mask[pos]) {
      if (AGGR == PGH_MAX || AGGR == PGH_MIN) cnt += (__ldg(data + pos * dense + ch) == o) ? 1 : 0;
      else ++cnt;
    }
  }
  const float w = (AGGR == PGH_SUM) ? g : (cnt > 0 ? g / (float)cnt : 0.f);
  for (int r = 0; r < nred; ++r) {
    long long pos = red == 1 ? (long long)r * n2 + keep : red == 2 ? (long long)keep * n2 + r : r;
    pos += b * n1 * n2;
    float v = 0.f;
    if (mask[pos]) {
      if (AGGR == PGH_MAX || AGGR == PGH_MIN) v = (__ldg(data + pos * dense + ch) == o) ? w : 0.f;
      else v = w;
    }
    g_data[pos * dense + ch] = v;
  }
}

__global__ void masked_fill_kernel(const float* __restrict__ data,
                                   const unsigned char* __restrict__ mask, long long rows,
                                   int dense, float value, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * dense) return;
  out[idx] = mask[idx / dense] ? data[idx] : value;
}

int mamamm_tc_launch(const float* A, int trans_a, const float* B, int trans_b,
                     const unsigned char* mask, const int* ext, int64_t b, int64_t n_i,
                     int64_t n_j, int64_t n_k, int64_t dense, float* out, cudaStream_t s);

// ext[b] = (1 + last row with a valid entry, 1 + last column with a valid entry) of a
// (b, n1, n2) mask; one CTA per graph
__global__ void mask_extents_kernel(const unsigned char* __restrict__ mask, int n1, int n2,
                                    int* __restrict__ ext) {
  __shared__ int rmax, cmax;
  if (threadIdx.x == 0) { rmax = 0; cmax = 0; }
  __syncthreads();
  const unsigned char* m = mask + (size_t)blockIdx.x * n1 * n2;
  int r = 0, c = 0;
  for (int p = threadIdx.x; p < n1 * n2; p += blockDim.x)
    if (m[p]) { r = max(r, p / n2 + 1); c = max(c, p % n2 + 1); }
  atomicMax(&rmax, r);
  atomicMax(&cmax, c);
  __syncthreads();
  if (threadIdx.x == 0) { ext[2 * blockIdx.x] = rmax; ext[2 * blockIdx.x + 1] = cmax; }
}

}  // namespace pgh

using namespace pgh;

extern "C" int pgh_mask_extents(const uint8_t* mask, int64_t b, int64_t n1, int64_t n2,
                                int32_t* ext, void* stream) {
  if (!mask || !ext) return arg_error("mask_extents: null pointer");
  if (b <= 0) return 0;
  mask_extents_kernel<<<(unsigned)b, 256, 0, as_stream(stream)>>>(mask, (int)n1, (int)n2, ext);
  return check_launch("mask_extents");
}

extern "C" int pgh_mamamm_f32(const float* A, int trans_a, const float* B, int trans_b,
                              const uint8_t* mask, const int32_t* ext, int64_t b, int64_t n_i,
                              int64_t n_j, int64_t n_k, int64_t dense, int algo, float* out,
                              void* stream) {
  if (!A || !B || !mask || !out) return arg_error("mamamm: null pointer");
  if (b < 0 || n_i <= 0 || n_j <= 0 || n_k <= 0 || dense <= 0) return arg_error("mamamm: sizes");
  if (b == 0) return 0;
  cudaStream_t s = as_stream(stream);
  if (algo == 1) return mamamm_tc_launch(A, trans_a, B, trans_b, mask, ext, b, n_i, n_j, n_k, dense, out, s);
  if (algo != 0) return arg_error("mamamm: algo");
  // A' (b, n_i, n_j): stored (b, n_i, n_j) or, transposed, (b, n_j, n_i)
  const long long sAi = trans_a ? 1 : n_j, sAj = trans_a ? n_i : 1;
  const long long sBj = trans_b ? 1 : n_k, sBk = trans_b ? n_j : 1;
  dim3 grid((unsigned)b, (unsigned)((dense + 31) / 32));
  mamamm_simt_kernel<<<grid, 256, 0, s>>>(A, sAi, sAj, B, sBj, sBk, mask, ext, (int)n_i,
                                          (int)n_j, (int)n_k, (int)dense, out);
  return check_launch("mamamm_simt");
}

extern "C" int pgh_masked_pool_f32(const float* data, const uint8_t* mask, int64_t b, int64_t n1,
                                   int64_t n2, int64_t dense, int red_dims, int aggr, float* out,
                                   uint8_t* out_mask, void* stream) {
  if (!data || !mask || !out) return arg_error("masked_pool: null pointer");
  if (red_dims < 1 || red_dims > 3) return arg_error("masked_pool: red_dims");
  const int64_t nkeep = red_dims == 1 ? n2 : red_dims == 2 ? n1 : 1;
  const long long total = (long long)b * nkeep * dense;
  if (total <= 0) return 0;
  cudaStream_t s = as_stream(stream);
  const unsigned nb = blocks_for(total, 256);
#define PGH_POOL(AG) masked_pool_kernel<AG><<<nb, 256, 0, s>>>(data, mask, (int)n1, (int)n2, (int)dense, red_dims, out, out_mask, total)
  switch (aggr) {
    case PGH_SUM: PGH_POOL(PGH_SUM); break;
    case PGH_MEAN: PGH_POOL(PGH_MEAN); break;
    case PGH_MAX: PGH_POOL(PGH_MAX); break;
    case PGH_MIN: PGH_POOL(PGH_MIN); break;
    default: return arg_error("masked_pool: aggr");
  }
#undef PGH_POOL
  return check_launch("masked_pool");
}

extern "C" int pgh_masked_pool_bwd_f32(const float* data, const uint8_t* mask, const float* out,
                                       const float* g_out, int64_t b, int64_t n1, int64_t n2,
                                       int64_t dense, int red_dims, int aggr, float* g_data,
                                       void* stream) {
  if (!data || !mask || !g_out || !g_data) return arg_error("masked_pool_bwd: null pointer");
  if (red_dims < 1 || red_dims > 3) return arg_error("masked_pool_bwd: red_dims");
  if ((aggr == PGH_MAX || aggr == PGH_MIN) && !out) return arg_error("masked_pool_bwd: out needed");
  const int64_t nkeep = red_dims == 1 ? n2 : red_dims == 2 ? n1 : 1;
  const long long total = (long long)b * nkeep * dense;
  if (total <= 0) return 0;
  cudaStream_t s = as_stream(stream);
  const unsigned nb = blocks_for(total, 256);
#define PGH_POOLB(AG) masked_pool_bwd_kernel<AG><<<nb, 256, 0, s>>>(data, mask, out, g_out, (int)n1, (int)n2, (int)dense, red_dims, g_data, total)
  switch (aggr) {
    case PGH_SUM: PGH_POOLB(PGH_SUM); break;
    case PGH_MEAN: PGH_POOLB(PGH_MEAN); break;
    case PGH_MAX: PGH_POOLB(PGH_MAX); break;
    case PGH_MIN: PGH_POOLB(PGH_MIN); break;
    default: return arg_error("masked_pool_bwd: aggr");
  }
#undef PGH_POOLB
  return check_launch("masked_pool_bwd");
}

extern "C" int pgh_masked_fill_f32(const float* data, const uint8_t* mask, int64_t rows,
                                   int64_t dense, float value, float* out, void* stream) {
  if (!data || !mask || !out) return arg_error("masked_fill: null pointer");
  if (rows * dense <= 0) return 0;
  masked_fill_kernel<<<blocks_for(rows * dense, 256), 256, 0, as_stream(stream)>>>(
      data, mask, rows, (int)dense, value, out);
  return check_launch("masked_fill");
}
