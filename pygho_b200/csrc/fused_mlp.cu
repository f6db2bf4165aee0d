// Fused BatchNorm(training) + activation around the tuplewise Linear layers.
//
// The reference's MLP block is Linear -> BatchNorm1d over ALL tuples of the batch -> SiLU
// (honn/utils.py:46-61, 85-142).  With stock ATen kernels one block costs 5 full passes over
// the (tuples x channels) activation forward and 7 backward (plus a separate column reduce for
// the Linear bias gradient).  These kernels do it in 3 + 5 passes, all HBM-streaming:
//
//   forward   stats(y) -> mean, rstd            (1 read)
//             z = act(gamma * (y - mean) * rstd + beta)          (1 read, 1 write)
//   backward  reduce: sum dyh, sum dyh*xh  with dyh = dz * act'(yh) recomputed   (2 reads)
//             dy = gamma*rstd*(dyh - mean(dyh) - xh*mean(dyh*xh)), column sums of dy (2 reads, 1 write)
//
// Thread layout everywhere: a thread owns 4 consecutive channels (one 128-bit lane of the
// row) and walks rows with stride, so every warp request is a contiguous 512 B piece of a row
// and per-channel accumulators live in registers.  Partial sums go to a workspace and are
// combined by a tiny second kernel in a fixed order (deterministic, no atomics).
#include <math.h>

#include "common.cuh"

namespace pgh {

constexpr int kBnThreads = 256;

// run-time tuning (pgh_set_tuning): [2] partial-reduction CTAs per SM (default 4), [3] rows in
// flight per thread in the backward kernels (2 or 4), [4] 1 = the apply passes walk the rows in
// the opposite direction of the reduce pass before them (the tail of that pass is still in the
// 126 MB L2), [5] elements in flight per thread of the forward apply (1, 2 or 4; default 2).
// Sweep on B200 (profiles/r1_bn_sweep.txt): 2 rows in flight and 4 CTAs/SM are best, the
// reverse walk buys nothing (the L2 does not keep the tail of a streaming pass)
static int bn_tune(int key, int dflt) { return g_tune[key] > 0 ? g_tune[key] : dflt; }

struct BnGeom {
  int c4;        // float4 columns per row
  int ty;        // row lanes per block
  int blocks;
  long long rows_per_block;
};

static BnGeom bn_geom(int64_t rows, int64_t C) {
  BnGeom g;
  g.c4 = (int)(C / 4);
  g.ty = kBnThreads / g.c4;
  if (g.ty < 1) g.ty = 1;
  const long long max_blocks = (long long)kSMs * bn_tune(2, 4);
  long long rpb = (rows + max_blocks - 1) / max_blocks;
  const long long min_rows = 8LL * g.ty;
  if (rpb < min_rows) rpb = min_rows;
  rpb = (rpb + g.ty - 1) / g.ty * g.ty;
  g.rows_per_block = rpb;
  g.blocks = (int)((rows + rpb - 1) / rpb);
  return g;
}

template <int ACT>
__device__ __forceinline__ float act_fwd(float x) {
  if (ACT == 1) return x / (1.f + __expf(-x));
  if (ACT == 2) return fmaxf(x, 0.f);
  return x;
}

template <int ACT>
__device__ __forceinline__ float act_grad(float x) {
  if (ACT == 1) {
    const float s = 1.f / (1.f + __expf(-x));
    return s * (1.f + x * (1.f - s));
  }
  if (ACT == 2) return x > 0.f ? 1.f : 0.f;
  return 1.f;
}

// block-level reduction over the ty row lanes of two float4 accumulators; result valid for ty==0
__device__ __forceinline__ void reduce_rows(float4& a, float4& b, float4* sm, int c4, int ty_n) {
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (ty_n == 1) return;
  sm[(ty * c4 + tx) * 2] = a;
  sm[(ty * c4 + tx) * 2 + 1] = b;
  __syncthreads();
  if (ty == 0) {
    for (int t = 1; t < ty_n; ++t) {
      const float4 x = sm[(t * c4 + tx) * 2], y = sm[(t * c4 + tx) * 2 + 1];
      a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
      b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w;
    }
  }
}

// ---- forward statistics: shifted sums (shift = row 0) to avoid cancellation -------------
__global__ void bn_stats_partial_kernel(const float* __restrict__ y, long long rows, int C,
                                        long long rows_per_block, float* __restrict__ part) {
  extern __shared__ float4 sm[];
  const int tx = threadIdx.x, c4 = blockDim.x, ty_n = blockDim.y;
  const float4 shift = __ldg(reinterpret_cast<const float4*>(y) + tx);
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  const float4* base = reinterpret_cast<const float4*>(y) + tx;
#pragma unroll 4
  for (long long r = r0 + threadIdx.y; r < r1; r += ty_n) {
    const float4 v = __ldg(base + r * c4);
    const float dx = v.x - shift.x, dy = v.y - shift.y, dz = v.z - shift.z, dw = v.w - shift.w;
    s.x += dx; s.y += dy; s.z += dz; s.w += dw;
    q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
  }
  reduce_rows(s, q, sm, c4, ty_n);
  if (threadIdx.y == 0) {
    float4* p = reinterpret_cast<float4*>(part + (size_t)blockIdx.x * 2 * C);
    p[tx] = s;
    p[c4 + tx] = q;
  }
}

// Second stage of every reduction: block = 32 channels x 32 partial lanes; lane ty sums the
// partials ty, ty+32, ... in double, then the 32 lanes are combined in a fixed order.
__device__ __forceinline__ void combine_partials(const float* __restrict__ part, int blocks,
                                                 size_t block_stride, int off0, int off1, int c,
                                                 bool ok, double& s, double& q) {
  __shared__ double sm[2][32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  double a = 0.0, b2 = 0.0;
  if (ok)
    for (int b = ty; b < blocks; b += 32) {
      a += (double)part[(size_t)b * block_stride + off0 + c];
      if (off1 >= 0) b2 += (double)part[(size_t)b * block_stride + off1 + c];
    }
  sm[0][ty][tx] = a;
  sm[1][ty][tx] = b2;
  __syncthreads();
  s = 0.0;
  q = 0.0;
  if (ty == 0)
    for (int t = 0; t < 32; ++t) { s += sm[0][t][tx]; q += sm[1][t][tx]; }
}

__global__ void bn_stats_final_kernel(const float* __restrict__ y, const float* __restrict__ part,
                                      int blocks, long long rows, int C, float eps, float momentum,
                                      float* __restrict__ mean, float* __restrict__ rstd,
                                      float* __restrict__ running_mean,
                                      float* __restrict__ running_var) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, q;
  combine_partials(part, blocks, (size_t)2 * C, 0, C, c, c < C, s, q);
  if (threadIdx.y != 0 || c >= C) return;
  const double n = (double)rows;
  const double ms = s / n;                       // mean of (y - shift)
  double var = q / n - ms * ms;                  // biased variance
  if (var < 0.0) var = 0.0;
  const float m = (float)(ms + (double)y[c]);
  mean[c] = m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = rows > 1 ? var * n / (n - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ---- forward apply ------------------------------------------------------------------------
template <int ACT, int UN>
__global__ void bn_act_fwd_kernel(const float4* __restrict__ y, const float4* __restrict__ mean,
                                  const float4* __restrict__ rstd, const float4* __restrict__ gamma,
                                  const float4* __restrict__ beta, long long n4, int c4, int rev,
                                  const float4* __restrict__ residual, float4* __restrict__ z) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * UN) {
    float4 v[UN], res[UN];
    long long idx[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long i = min(i0 + u * stride, n4 - 1);      // clamped: loads are unconditional
      idx[u] = rev ? n4 - 1 - i : i;
      v[u] = __ldg(y + idx[u]);
      if (residual) res[u] = __ldg(residual + idx[u]);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (i0 + u * stride >= n4) break;
      const int c = (int)(idx[u] % c4);
      const float4 m = __ldg(mean + c), r = __ldg(rstd + c);
      const float4 g = gamma ? __ldg(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 b = beta ? __ldg(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 o;
      o.x = act_fwd<ACT>(fmaf((v[u].x - m.x) * r.x, g.x, b.x));
      o.y = act_fwd<ACT>(fmaf((v[u].y - m.y) * r.y, g.y, b.y));
      o.z = act_fwd<ACT>(fmaf((v[u].z - m.z) * r.z, g.z, b.z));
      o.w = act_fwd<ACT>(fmaf((v[u].w - m.w) * r.w, g.w, b.w));
      if (residual) { o.x += res[u].x; o.y += res[u].y; o.z += res[u].z; o.w += res[u].w; }
      z[idx[u]] = o;
    }
  }
}

// ---- backward reduce: sum dyh and sum dyh * xh -------------------------------------------
template <int ACT, int UN>
__global__ void bn_act_bwd_reduce_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         long long rows, int C, long long rows_per_block,
                                         float* __restrict__ part) {
  extern __shared__ float4 sm[];
  const int tx = threadIdx.x, c4 = blockDim.x, ty_n = blockDim.y;
  const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + tx);
  const float4 rs = __ldg(reinterpret_cast<const float4*>(rstd) + tx);
  const float4 g = gamma ? __ldg(reinterpret_cast<const float4*>(gamma) + tx) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 bt = beta ? __ldg(reinterpret_cast<const float4*>(beta) + tx) : make_float4(0.f, 0.f, 0.f, 0.f);
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  const float4* yb = reinterpret_cast<const float4*>(y) + tx;
  const float4* db = reinterpret_cast<const float4*>(dz) + tx;
  for (long long rb = r0 + threadIdx.y; rb < r1; rb += (long long)ty_n * UN) {
    float4 vv[UN], dd[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {                 // clamped rows: all 2*UN loads issue together
      const long long r = min(rb + (long long)u * ty_n, r1 - 1);
      vv[u] = __ldg(yb + r * c4);
      dd[u] = __ldg(db + r * c4);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (rb + (long long)u * ty_n >= r1) break;
      const float4 v = vv[u], d = dd[u];
      const float xx = (v.x - m.x) * rs.x, xy = (v.y - m.y) * rs.y, xz = (v.z - m.z) * rs.z, xw = (v.w - m.w) * rs.w;
      const float gx = d.x * act_grad<ACT>(fmaf(xx, g.x, bt.x)), gy = d.y * act_grad<ACT>(fmaf(xy, g.y, bt.y));
      const float gz = d.z * act_grad<ACT>(fmaf(xz, g.z, bt.z)), gw = d.w * act_grad<ACT>(fmaf(xw, g.w, bt.w));
      s.x += gx; s.y += gy; s.z += gz; s.w += gw;
      q.x = fmaf(gx, xx, q.x); q.y = fmaf(gy, xy, q.y); q.z = fmaf(gz, xz, q.z); q.w = fmaf(gw, xw, q.w);
    }
  }
  reduce_rows(s, q, sm, c4, ty_n);
  if (threadIdx.y == 0) {
    float4* p = reinterpret_cast<float4*>(part + (size_t)blockIdx.x * 2 * C);
    p[tx] = s;
    p[c4 + tx] = q;
  }
}

// writes dbeta = S1, dgamma = S2 and the per-channel means used by the apply pass
__global__ void bn_bwd_final_kernel(const float* __restrict__ part, int blocks, long long rows, int C,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    float* __restrict__ m1, float* __restrict__ m2) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, q;
  combine_partials(part, blocks, (size_t)2 * C, 0, C, c, c < C, s, q);
  if (threadIdx.y != 0 || c >= C) return;
  if (dbeta) dbeta[c] = (float)s;
  if (dgamma) dgamma[c] = (float)q;
  m1[c] = (float)(s / (double)rows);
  m2[c] = (float)(q / (double)rows);
}

// ---- backward apply: dy and the column sums of dy (the Linear bias gradient) ------------------
template <int ACT, int UN>
__global__ void bn_act_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                        const float* __restrict__ m1, const float* __restrict__ m2,
                                        long long rows, int C, long long rows_per_block, int rev,
                                        float* __restrict__ dy, float* __restrict__ part) {
  extern __shared__ float4 sm[];
  const int tx = threadIdx.x, c4 = blockDim.x, ty_n = blockDim.y;
  const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + tx);
  const float4 rs = __ldg(reinterpret_cast<const float4*>(rstd) + tx);
  const float4 g = gamma ? __ldg(reinterpret_cast<const float4*>(gamma) + tx) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 bt = beta ? __ldg(reinterpret_cast<const float4*>(beta) + tx) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 a1 = __ldg(reinterpret_cast<const float4*>(m1) + tx);
  const float4 a2 = __ldg(reinterpret_cast<const float4*>(m2) + tx);
  const float4 sc = make_float4(g.x * rs.x, g.y * rs.y, g.z * rs.z, g.w * rs.w);
  // rev: CTA 0 takes the LAST row slab (the rows the reduce pass touched most recently)
  const long long slab = rev ? (long long)(gridDim.x - 1 - blockIdx.x) : (long long)blockIdx.x;
  const long long r0 = slab * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), unused = s;
  const float4* yb = reinterpret_cast<const float4*>(y) + tx;
  const float4* db = reinterpret_cast<const float4*>(dz) + tx;
  float4* ob = reinterpret_cast<float4*>(dy) + tx;
  for (long long rb = r0 + threadIdx.y; rb < r1; rb += (long long)ty_n * UN) {
    float4 vv[UN], dd[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long r = min(rb + (long long)u * ty_n, r1 - 1);
      vv[u] = __ldg(yb + r * c4);
      dd[u] = __ldg(db + r * c4);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long r = rb + (long long)u * ty_n;
      if (r >= r1) break;
      const float4 v = vv[u], d = dd[u];
      const float xx = (v.x - m.x) * rs.x, xy = (v.y - m.y) * rs.y, xz = (v.z - m.z) * rs.z, xw = (v.w - m.w) * rs.w;
      const float gx = d.x * act_grad<ACT>(fmaf(xx, g.x, bt.x)), gy = d.y * act_grad<ACT>(fmaf(xy, g.y, bt.y));
      const float gz = d.z * act_grad<ACT>(fmaf(xz, g.z, bt.z)), gw = d.w * act_grad<ACT>(fmaf(xw, g.w, bt.w));
      float4 o;
      o.x = sc.x * (gx - a1.x - xx * a2.x);
      o.y = sc.y * (gy - a1.y - xy * a2.y);
      o.z = sc.z * (gz - a1.z - xz * a2.z);
      o.w = sc.w * (gw - a1.w - xw * a2.w);
      ob[r * c4] = o;
      s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
    }
  }
  if (part) {
    reduce_rows(s, unused, sm, c4, ty_n);
    if (threadIdx.y == 0) reinterpret_cast<float4*>(part + (size_t)slab * C)[tx] = s;
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ part, int blocks, int C,
                                    float* __restrict__ out) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, q;
  combine_partials(part, blocks, (size_t)C, 0, -1, c, c < C, s, q);
  if (threadIdx.y == 0 && c < C) out[c] = (float)s;
}

static bool bn_ok(int64_t rows, int64_t C) {
  return rows > 0 && C >= 4 && C % 4 == 0 && C / 4 <= kBnThreads;
}

}  // namespace pgh

using namespace pgh;

extern "C" size_t pgh_bn_ws_bytes(int64_t rows, int64_t C) {
  if (!bn_ok(rows, C)) return 256;
  const BnGeom g = bn_geom(rows, C);
  return (size_t)g.blocks * 2 * C * sizeof(float) + 2 * C * sizeof(float) + 256;
}

extern "C" int pgh_bn_stats_f32(const float* y, int64_t rows, int64_t C, float eps, float momentum,
                                float* mean, float* rstd, float* running_mean, float* running_var,
                                void* ws, size_t ws_bytes, void* stream) {
  if (!y || !mean || !rstd || !ws) return arg_error("bn_stats: null pointer");
  if (!bn_ok(rows, C)) return arg_error("bn_stats: need rows > 0, C % 4 == 0, C <= 1024");
  if ((reinterpret_cast<uintptr_t>(y) & 15)) return arg_error("bn_stats: y must be 16-byte aligned");
  const BnGeom g = bn_geom(rows, C);
  if (ws_bytes < pgh_bn_ws_bytes(rows, C)) return arg_error("bn_stats: workspace too small");
  cudaStream_t s = as_stream(stream);
  float* part = reinterpret_cast<float*>(ws);
  const size_t smem = (size_t)g.c4 * g.ty * 2 * sizeof(float4);
  bn_stats_partial_kernel<<<g.blocks, dim3(g.c4, g.ty), smem, s>>>(y, rows, (int)C, g.rows_per_block, part);
  bn_stats_final_kernel<<<blocks_for(C, 32), dim3(32, 32), 0, s>>>(y, part, g.blocks, rows, (int)C, eps, momentum,
                                                            mean, rstd, running_mean, running_var);
  return check_launch("bn_stats");
}

extern "C" int pgh_bn_act_res_fwd_f32(const float* y, const float* mean, const float* rstd,
                                      const float* gamma, const float* beta, int64_t rows,
                                      int64_t C, int act, const float* residual, float* z,
                                      void* stream) {
  if (!y || !mean || !rstd || !z) return arg_error("bn_act_fwd: null pointer");
  if (!bn_ok(rows, C)) return arg_error("bn_act_fwd: need rows > 0, C % 4 == 0, C <= 1024");
  if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z) |
       reinterpret_cast<uintptr_t>(residual)) & 15)
    return arg_error("bn_act_fwd: tensors must be 16-byte aligned");
  const long long n4 = rows * (C / 4);
  long long nb = (n4 + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  cudaStream_t s = as_stream(stream);
  const int rev = bn_tune(4, 0) == 1;
#define PGH_FWD_U(A, U) bn_act_fwd_kernel<A, U><<<(unsigned)nb, 256, 0, s>>>(                        \
      (const float4*)y, (const float4*)mean, (const float4*)rstd, (const float4*)gamma,             \
      (const float4*)beta, n4, (int)(C / 4), rev, (const float4*)residual, (float4*)z)
#define PGH_FWD(A)                                                                                   \
  do {                                                                                               \
    const int un_ = bn_tune(5, 2);                                                                   \
    if (un_ >= 4) PGH_FWD_U(A, 4); else if (un_ == 2) PGH_FWD_U(A, 2); else PGH_FWD_U(A, 1);        \
  } while (0)
  if (act == 1) PGH_FWD(1); else if (act == 2) PGH_FWD(2); else if (act == 0) PGH_FWD(0);
  else return arg_error("bn_act_fwd: act");
#undef PGH_FWD_U
#undef PGH_FWD
  return check_launch("bn_act_fwd");
}

extern "C" int pgh_bn_act_fwd_f32(const float* y, const float* mean, const float* rstd,
                                  const float* gamma, const float* beta, int64_t rows, int64_t C,
                                  int act, float* z, void* stream) {
  return pgh_bn_act_res_fwd_f32(y, mean, rstd, gamma, beta, rows, C, act, nullptr, z, stream);
}

extern "C" int pgh_bn_act_bwd_f32(const float* dz, const float* y, const float* mean, const float* rstd,
                                  const float* gamma, const float* beta, int64_t rows, int64_t C,
                                  int act, float* dy, float* dgamma, float* dbeta, float* dbias,
                                  void* ws, size_t ws_bytes, void* stream) {
  if (!dz || !y || !mean || !rstd || !dy || !ws) return arg_error("bn_act_bwd: null pointer");
  if (!bn_ok(rows, C)) return arg_error("bn_act_bwd: need rows > 0, C % 4 == 0, C <= 1024");
  if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dy)) & 15)
    return arg_error("bn_act_bwd: tensors must be 16-byte aligned");
  if (act < 0 || act > 2) return arg_error("bn_act_bwd: act");
  const BnGeom g = bn_geom(rows, C);
  if (ws_bytes < pgh_bn_ws_bytes(rows, C)) return arg_error("bn_act_bwd: workspace too small");
  cudaStream_t s = as_stream(stream);
  float* part = reinterpret_cast<float*>(ws);
  float* m1 = part + (size_t)g.blocks * 2 * C;
  float* m2 = m1 + C;
  const size_t smem = (size_t)g.c4 * g.ty * 2 * sizeof(float4);
  const dim3 blk(g.c4, g.ty);
  const int rev = bn_tune(4, 0) == 1;
#define PGH_BWD_U(A, U)                                                                              \
  bn_act_bwd_reduce_kernel<A, U><<<g.blocks, blk, smem, s>>>(dz, y, mean, rstd, gamma, beta, rows,  \
                                                              (int)C, g.rows_per_block, part);       \
  bn_bwd_final_kernel<<<blocks_for(C, 32), dim3(32, 32), 0, s>>>(part, g.blocks, rows, (int)C, dgamma,      \
                                                          dbeta, m1, m2);                            \
  bn_act_bwd_apply_kernel<A, U><<<g.blocks, blk, smem, s>>>(dz, y, mean, rstd, gamma, beta, m1, m2, \
                                                             rows, (int)C, g.rows_per_block, rev,    \
                                                             dy, dbias ? part : nullptr)
#define PGH_BWD(A)                                                                                   \
  do {                                                                                               \
    if (bn_tune(3, 2) >= 4) { PGH_BWD_U(A, 4); } else { PGH_BWD_U(A, 2); }                           \
  } while (0)
  if (act == 1) { PGH_BWD(1); } else if (act == 2) { PGH_BWD(2); } else { PGH_BWD(0); }
#undef PGH_BWD_U
#undef PGH_BWD
  if (dbias) colsum_final_kernel<<<blocks_for(C, 32), dim3(32, 32), 0, s>>>(part, g.blocks, (int)C, dbias);
  return check_launch("bn_act_bwd");
}
