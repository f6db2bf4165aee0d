// 2-FWL contraction of MaskedTensors (reference backend/Mamamm.py:7-64), algo 4:
//
//   out[b,i,k,c] = mask[b,i,k] ? sum_j A'[b,i,j,c] * B'[b,j,k,c] : 0        (per channel c)
//
// The work is one tiny (n x n x n, n ~ 23 on ZINC) product per (graph, channel): 0.5 GFLOP of
// useful math against 180 MB of compulsory traffic at b = 128, n <= 40, d = 128.  It is an HBM
// problem, not a tensor-core problem, so this kernel keeps the reference's channel-last layout
// (no transposes into MMA operand layouts) and does exact fp32 FMAs from shared memory:
//
//   warp 15    TMA producer: one thread walks a dynamic queue of (graph, 16-channel slab, row
//              pass) units; for every chunk of JC contraction indices it issues 4-D
//              cp.async.bulk.tensor boxes (16 channels x JC x 8 rows / columns) of A' and B' into a
//              shared-memory ring -- only boxes that intersect the graph's valid extents are
//              fetched, the ring keeps filling across unit boundaries
//   warps 0-14 consumers: a thread owns a 4 (rows) x 4 (columns) x 4 (channels) register tile;
//              four lanes cover the 16 channels of a cell, the two lane groups of a quarter
//              warp share the A' rows (broadcast) and take the even / odd columns of an 8-column
//              block, whose cells are an odd number of 64-byte cells apart in both operand
//              layouts (JC = 9 when B' is contraction-major): all LDS.128 are conflict free
//   epilogue   mask select + 128-bit stores of the valid region; the pad region of the output is
//              zero-filled in whole 512-byte rows, dealt over the units of the graph
//
// Sum order is j ascending with fmaf, the same as the CUDA-core kernel (algo 0): bit-identical.
// Operand pads must be zero (MaskedTensor keeps them at padvalue 0), as for the other kernels.
#include <cuda.h>

#include "common.cuh"

namespace pgh {

constexpr int kMsCH = 16;                    // channels per slab: one 64-byte cell per (row, col)
constexpr int kMsCell = kMsCH * 4;           // bytes
constexpr int kMsRB = 8;                     // rows / columns per TMA box
constexpr int kMsConsumerWarps = 15;          // + 1 producer warp = 16 warps: 128 registers per thread
constexpr int kMsConsumers = kMsConsumerWarps * 32;
constexpr int kMsThreads = kMsConsumers + 32;
constexpr int kMsMaxPairs = kMsConsumers / 8;   // a pair = 4 rows x 8 columns of output
constexpr int kMsMaxStages = 8;
constexpr size_t kMsSmemCap = 227 * 1024;

struct MsParams {
  int n_i, n_j, n_k, dense;       // tensor extents of A' (n_i x n_j) and B' (n_j x n_k)
  int nslab, npass, rows_per_pass;
  int nab_max, nkb_max;           // boxes of A' / B' per stage
  int stages, stage_bytes;
  long long units;                // b * nslab * npass
};

struct MsMeta {
  int item, slab, pass, chunk, nchunks, ei, ek, rows;
};

__device__ __forceinline__ uint32_t ms_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void ms_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// bounded wait: a lost arrival traps (reported as a launch error) instead of hanging the GPU
__device__ __forceinline__ void ms_mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < (1u << 14); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void ms_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ms_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void ms_tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                               int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ float4 ms_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ void ms_fma4(float4& acc, const float4& a, const float4& b) {
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.y = fmaf(a.y, b.y, acc.y);
  acc.z = fmaf(a.z, b.z, acc.z);
  acc.w = fmaf(a.w, b.w, acc.w);
}

// AK / BK: the operand's contraction index is its fast (inner) spatial dim as stored:
//   A' = A   stored (b, n_i, n_j, C) -> AK;   A' = A^T stored (b, n_j, n_i, C) -> !AK
//   B' = B   stored (b, n_j, n_k, C) -> !BK;  B' = B^T stored (b, n_k, n_j, C) -> BK
// Box of a contraction-major operand: (16 ch, JC, 8 rows): cell(r, jj) = r * JC + jj;
// of a row-major one: (16 ch, 8 rows, JC): cell(r, jj) = jj * 8 + r.
template <bool AK, bool BK>
__global__ void __launch_bounds__(kMsThreads, 1)
mamamm_smem_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const unsigned char* __restrict__ mask, const int* __restrict__ ext, MsParams P,
                   float* __restrict__ out, unsigned int* __restrict__ counters) {
  constexpr int JC = BK ? 9 : 8;
  constexpr int kBoxCells = kMsRB * JC;
  constexpr int kBoxBytes = kBoxCells * kMsCell;
  extern __shared__ unsigned char ms_smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[kMsMaxStages], empty_bar[kMsMaxStages];
  __shared__ MsMeta meta[kMsMaxStages];

  const uint32_t ring = (ms_u32(ms_smem_raw) + 127u) & ~127u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < P.stages; ++s) {
      ms_mbar_init(ms_u32(&full_bar[s]), 1);
      ms_mbar_init(ms_u32(&empty_bar[s]), kMsConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int upi = P.nslab * P.npass;          // units per graph

  if (warp == kMsConsumerWarps) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (;;) {
        const unsigned int u = atomicAdd(&counters[0], 1u);
        if ((long long)u >= P.units) break;
        const int item = (int)(u / (unsigned)upi), rem = (int)(u % (unsigned)upi);
        const int slab = rem / P.npass, pass = rem % P.npass;
        int ei = P.n_i, ej = P.n_j, ek = P.n_k;
        if (ext) {
          ei = min(max(ext[3 * item + 0], 0), P.n_i);
          ej = min(max(ext[3 * item + 1], 0), P.n_j);
          ek = min(max(ext[3 * item + 2], 0), P.n_k);
        }
        const int r0 = pass * P.rows_per_pass;
        const int rows = max(min(ei - r0, P.rows_per_pass), 0);
        const int nab = (rows + kMsRB - 1) / kMsRB, nkb = (ek + kMsRB - 1) / kMsRB;
        const int nch = (rows > 0 && ek > 0) ? max((ej + JC - 1) / JC, 1) : 0;
        const int c0 = slab * kMsCH;
        for (int ch = 0; ch < max(nch, 1); ++ch) {
          ms_mbar_wait(ms_u32(&empty_bar[s]), ph ^ 1u);
          meta[s] = MsMeta{item, slab, pass, ch, nch, ei, ek, rows};
          const uint32_t fb = ms_u32(&full_bar[s]);
          if (nch == 0) {
            ms_mbar_arrive(fb);
          } else {
            ms_mbar_expect_tx(fb, (uint32_t)((nab + nkb) * kBoxBytes));
            const uint32_t sa = ring + (uint32_t)s * (uint32_t)P.stage_bytes;
            const uint32_t sb = sa + (uint32_t)P.nab_max * kBoxBytes;
            const int j0 = ch * JC;
            for (int ib = 0; ib < nab; ++ib) {
              const int row = r0 + ib * kMsRB;
              if (AK) ms_tma_load_4d(sa + ib * kBoxBytes, &map_a, fb, c0, j0, row, item);
              else    ms_tma_load_4d(sa + ib * kBoxBytes, &map_a, fb, c0, row, j0, item);
            }
            for (int kb = 0; kb < nkb; ++kb) {
              const int col = kb * kMsRB;
              if (BK) ms_tma_load_4d(sb + kb * kBoxBytes, &map_b, fb, c0, j0, col, item);
              else    ms_tma_load_4d(sb + kb * kBoxBytes, &map_b, fb, c0, col, j0, item);
            }
          }
          if (++s == P.stages) { s = 0; ph ^= 1u; }
        }
      }
      // end marker for the consumers
      ms_mbar_wait(ms_u32(&empty_bar[s]), ph ^ 1u);
      meta[s] = MsMeta{-1, 0, 0, 0, 0, 0, 0, 0};
      ms_mbar_arrive(ms_u32(&full_bar[s]));
      // the last CTA to run dry re-arms the queue for the next launch
      __threadfence();
      if (atomicAdd(&counters[1], 1u) == gridDim.x - 1) {
        counters[0] = 0;
        counters[1] = 0;
        __threadfence();
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int grp = tid >> 2, l4 = tid & 3, pair = grp >> 1, par = grp & 1;
  float4 acc[4][4];
  int s = 0;
  uint32_t ph = 0;
  bool active = false;
  int i0 = 0, kb = 0;
  uint32_t offA = 0, offB = 0;
  for (;;) {
    ms_mbar_wait(ms_u32(&full_bar[s]), ph);
    const MsMeta m = meta[s];
    if (m.item < 0) break;
    if (m.chunk == 0) {
      const int nkb = (m.ek + kMsRB - 1) / kMsRB;
      const int it = nkb > 0 ? pair / nkb : 0;
      kb = nkb > 0 ? pair - it * nkb : 0;
      i0 = it * 4;
      active = m.nchunks > 0 && i0 < m.rows;
      const int rr0 = i0 & 7;
      offA = (uint32_t)((i0 >> 3) * kBoxBytes + (AK ? rr0 * JC : rr0) * kMsCell + l4 * 16);
      offB = (uint32_t)(P.nab_max * kBoxBytes + kb * kBoxBytes + (BK ? par * JC : par) * kMsCell + l4 * 16);
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (active) {
      const uint32_t base = ring + (uint32_t)s * (uint32_t)P.stage_bytes;
      const uint32_t pa = base + offA, pb = base + offB;
#pragma unroll(1)
      for (int jj = 0; jj < JC; ++jj) {
        float4 a[4], bv[4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
          a[x] = ms_lds128(pa + (uint32_t)((AK ? x * JC + jj : jj * kMsRB + x) * kMsCell));
#pragma unroll
        for (int y = 0; y < 4; ++y)
          bv[y] = ms_lds128(pb + (uint32_t)((BK ? 2 * y * JC + jj : jj * kMsRB + 2 * y) * kMsCell));
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) ms_fma4(acc[x][y], a[x], bv[y]);
      }
    }
    __syncwarp();
    if (lane == 0) ms_mbar_arrive(ms_u32(&empty_bar[s]));
    if (++s == P.stages) { s = 0; ph ^= 1u; }

    if (m.chunk + 1 >= m.nchunks) {
      // ---------------------------------------------------------------- epilogue of the unit
      const long long cell0 = (long long)m.item * P.n_i * P.n_k;
      if (active) {
        const int r0 = m.pass * P.rows_per_pass;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int i = r0 + i0 + x;
#pragma unroll
          for (int y = 0; y < 4; ++y) {
            const int k = kb * kMsRB + 2 * y + par;
            if (i < m.ei && k < m.ek) {
              const long long cell = cell0 + (long long)i * P.n_k + k;
              const float4 v = mask[cell] ? acc[x][y] : make_float4(0.f, 0.f, 0.f, 0.f);
              *reinterpret_cast<float4*>(out + cell * P.dense + m.slab * kMsCH + l4 * 4) = v;
            }
          }
        }
      }
      // pads of the output: whole rows of `dense` zeros, dealt over the units of the graph
      const int cells = P.n_i * P.n_k, me = m.slab * P.npass + m.pass;
      const int c4n = P.dense >> 2;
      for (int q = me + upi * warp; q < cells; q += upi * kMsConsumerWarps) {
        const int i = q / P.n_k, k = q - i * P.n_k;
        if (i >= m.ei || k >= m.ek) {
          float4* row = reinterpret_cast<float4*>(out + (cell0 + q) * P.dense);
          for (int c4 = lane; c4 < c4n; c4 += 32) row[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
}

typedef CUresult (*PFN_msEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                      const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_msEncodeTiled ms_get_encode() {
  static PFN_msEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_msEncodeTiled>(p);
  }
  return fn;
}

// (b, d1, d2, C) fp32, channel-last.  kfast: the contraction index is d2 -> box (16, JC, 8, 1),
// otherwise it is d1 -> box (16, 8, JC, 1).  Out-of-bounds parts of a box are zero-filled.
static bool ms_make_map(CUtensorMap* map, const float* base, int64_t b, int64_t d1, int64_t d2,
                        int64_t C, bool kfast, int JC) {
  PFN_msEncodeTiled enc = ms_get_encode();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)d2, (cuuint64_t)d1, (cuuint64_t)b};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)d2 * C * 4, (cuuint64_t)d1 * d2 * C * 4};
  const cuuint32_t box[4] = {(cuuint32_t)kMsCH, (cuuint32_t)(kfast ? JC : kMsRB),
                             (cuuint32_t)(kfast ? kMsRB : JC), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// self-resetting work queues, rotated per launch so that launches on different streams (or
// concurrent branches of one captured graph) never share one
constexpr int kMsQueues = 64;
__device__ unsigned int g_ms_counters[kMsQueues * 2];

static bool ms_plan(int64_t b, int64_t n_i, int64_t n_j, int64_t n_k, int64_t dense, bool bk, MsParams& P) {
  if (dense % kMsCH != 0 || n_i <= 0 || n_j <= 0 || n_k <= 0) return false;
  if (n_i > 32768 || n_j > 32768 || n_k > 32768 || dense > 32768) return false;
  const int JC = bk ? 9 : 8;
  const int nkb = (int)((n_k + kMsRB - 1) / kMsRB);
  if (nkb > kMsMaxPairs) return false;
  const int itiles = (int)((n_i + 3) / 4);
  int itp = kMsMaxPairs / nkb;
  if (itp > itiles) itp = itiles;
  P.n_i = (int)n_i; P.n_j = (int)n_j; P.n_k = (int)n_k; P.dense = (int)dense;
  P.nslab = (int)(dense / kMsCH);
  P.rows_per_pass = itp * 4;
  P.npass = (itiles + itp - 1) / itp;
  P.nab_max = (P.rows_per_pass + kMsRB - 1) / kMsRB;
  P.nkb_max = nkb;
  P.stage_bytes = (P.nab_max + P.nkb_max) * kMsRB * JC * kMsCell;
  const long long room = (long long)kMsSmemCap - 1024 - 128;   // static barriers / meta, alignment
  long long st = room / P.stage_bytes;
  if (st < 2) return false;
  P.stages = (int)(st > kMsMaxStages ? kMsMaxStages : st);
  const long long units = (long long)b * P.nslab * P.npass;
  if (units <= 0 || units > 0x7fffffffLL - 4096 || (long long)n_i * n_k > 0x7fffffffLL) return false;
  P.units = units;
  return true;
}

template <bool AK, bool BK>
static int ms_launch_t(const CUtensorMap& ma, const CUtensorMap& mb, const unsigned char* mask,
                       const int* ext, const MsParams& P, float* out, unsigned int* counters,
                       cudaStream_t s) {
  const size_t smem = (size_t)P.stages * P.stage_bytes + 128;
  static bool attr_set = false;
  if (!attr_set) {
    PGH_CUDA(cudaFuncSetAttribute(mamamm_smem_kernel<AK, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(kMsSmemCap - 1024)));
    attr_set = true;
  }
  const unsigned grid = (unsigned)(P.units < kSMs ? P.units : kSMs);
  mamamm_smem_kernel<AK, BK><<<grid, kMsThreads, smem, s>>>(ma, mb, mask, ext, P, out, counters);
  return check_launch("mamamm_smem");
}

// returns -1 when the shape is not supported (the caller falls back to another algo)
int mamamm_smem_launch(const float* A, int trans_a, const float* B, int trans_b,
                       const unsigned char* mask, const int* ext, int64_t b, int64_t n_i, int64_t n_j,
                       int64_t n_k, int64_t dense, float* out, cudaStream_t s) {
  const bool ak = !trans_a, bk = trans_b != 0;
  MsParams P;
  if (!ms_plan(b, n_i, n_j, n_k, dense, bk, P)) return -1;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(out)) & 15)
    return -1;
  const int JC = bk ? 9 : 8;
  CUtensorMap ma, mb;
  // A' = A: stored (b, n_i, n_j, C), contraction fast; A' = A^T: stored (b, n_j, n_i, C)
  if (!ms_make_map(&ma, A, b, trans_a ? n_j : n_i, trans_a ? n_i : n_j, dense, ak, JC)) return -1;
  // B' = B: stored (b, n_j, n_k, C), contraction slow; B' = B^T: stored (b, n_k, n_j, C)
  if (!ms_make_map(&mb, B, b, trans_b ? n_k : n_j, trans_b ? n_j : n_k, dense, bk, JC)) return -1;
  static unsigned int* counters = nullptr;
  static unsigned int next_queue = 0;
  if (!counters) {
    void* p = nullptr;
    PGH_CUDA(cudaGetSymbolAddress(&p, g_ms_counters));
    counters = static_cast<unsigned int*>(p);
  }
  unsigned int* q = counters + 2 * (next_queue++ % kMsQueues);
  if (ak && bk) return ms_launch_t<true, true>(ma, mb, mask, ext, P, out, q, s);
  if (ak && !bk) return ms_launch_t<true, false>(ma, mb, mask, ext, P, out, q, s);
  if (!ak && bk) return ms_launch_t<false, true>(ma, mb, mask, ext, P, out, q, s);
  return ms_launch_t<false, false>(ma, mb, mask, ext, P, out, q, s);
}

}  // namespace pgh
