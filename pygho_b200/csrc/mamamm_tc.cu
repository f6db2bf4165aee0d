// 2-FWL contraction on the 5th-generation tensor cores (tcgen05, TF32 in / FP32 accumulate).
//
//   out[b,i,k,c] = mask[b,i,k] * sum_j A'[b,i,j,c] * B'[b,j,k,c]
//
// The channel axis c is the contiguous one in memory and is a *batch* axis of the GEMM, so
// the operands cannot be fed to UMMA in their global layout: a CTA takes one graph b and a
// slab of 8 channels (one full 32 B sector per (i,j) position), reads the slab once with
// 128-bit loads and transposes it on the fly into 8 per-channel K-major / no-swizzle UMMA
// tiles in shared memory (core matrix = 8 rows x 16 B).  One elected thread then issues
// M=128 x N=n_k(pad 16) x K=8 tcgen05.mma instructions, accumulators live in TMEM
// (4 channels x N columns per round, 256 columns per CTA so two CTAs share an SM), and the
// epilogue reads them back with tcgen05.ld (lane = output row), gathers 4 channels per
// thread and writes masked 16 B pieces straight to global memory.
//
// Rows >= n_i of the M=128 tile and columns >= n_k of the N tile read whatever follows the
// tile in shared memory: they only produce accumulator rows/columns that are never read.
// The K padding (n_j -> multiple of 8) is zero-filled in both operands.
//
// HBM-bound by design (13 flop/B vs a ridge of ~200 flop/B): what matters is that every
// byte is read once, in full sectors, and that loads of one CTA overlap the MMA/epilogue of
// the other CTA on the SM.
#include "common.cuh"

namespace pgh {

constexpr int kCS = 8;            // channels per CTA
constexpr int kTcThreads = 256;
constexpr int kCoreBytes = 128;   // one 8x4 fp32 core matrix
constexpr int kTmemCols = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout:
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64))
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// kind::tf32 instruction descriptor: D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2,
// both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // bounded spin: a lost arrival traps instead of hanging the GPU
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "r"(taddr));
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2);
  v[3] = __uint_as_float(r3); v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5);
  v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

struct TcGeom {
  int n_i, n_j, n_k;          // padded (tensor) extents of A' (n_i x n_j) and B' (n_j x n_k)
  int sbo;                    // bytes between 8-row groups (= k_pad_max/4 core matrices)
  int ts_a, ts_b;             // bytes per channel tile (odd multiple of 16 B: bank spread)
  int off_b;                  // byte offset of the B tiles
  int off_mask;               // byte offset of the staged (n_i x n_k) mask tile
  long long sa_i, sa_j;       // element strides of A' in units of `dense` floats
  long long sb_j, sb_k;
};

// position (row, kcol) of a K-major no-swizzle tile -> byte offset
__device__ __forceinline__ int tile_off(int row, int kcol, int sbo) {
  return (row >> 3) * sbo + (kcol >> 2) * kCoreBytes + (row & 7) * 16 + (kcol & 3) * 4;
}

constexpr int kLU = 10;  // 128-bit loads in flight per thread in the load phase

struct OperandView {
  const float* src;      // + 4 * half
  unsigned char* tiles;  // + (4 * half) * ts
  int n_rows, ts;
  long long s_row, s_k;
};

// Fill the 8 channel tiles of BOTH operands for rows < n_rows and K columns < k_pad (zeros at
// kcol >= nj).  A warp-wide unit is a 4-row x 4-kcol patch of one operand: lane = (r2, kq, h)
// with h = channel half, kq = kcol & 3, r2 = row & 3, which puts the 32 lanes of every STS on
// 32 different banks (bank = kq + 4*r2 + 16*h) while each lane pair still reads one full 32 B
// sector from global memory.  Units of A and B are interleaved over the warps and up to kLU
// loads per thread are issued before the first store, so a typical graph (n ~ 23: 72 units,
// 9 per warp) costs ONE global-memory latency for the whole load phase.
__device__ __forceinline__ void load_tiles(const OperandView& va, const OperandView& vb, int nj,
                                           int k_pad, int dense, int sbo, int warp, int lane) {
  const int h = lane & 1, kq = (lane >> 1) & 3, r2 = lane >> 3;
  const int kquads = k_pad >> 2;
  const int units_a = kquads * ((va.n_rows + 3) >> 2);
  const int units = units_a + kquads * ((vb.n_rows + 3) >> 2);
  constexpr int kWarps = kTcThreads / 32;
  // (patch row, patch column) of unit = warp + k * kWarps, advanced without divisions
  const int dq = kWarps / kquads, dr = kWarps % kquads;
  for (int u0 = warp; u0 < units; u0 += kLU * kWarps) {
    float4 v[kLU];
    int off[kLU];   // byte offset from smem base of channel tile (4*h), -1 = nothing to store
    int unit = u0;
    bool is_b = unit >= units_a;
    int local = is_b ? unit - units_a : unit;
    int pq = local / kquads, pr = local - pq * kquads;
#pragma unroll
    for (int u = 0; u < kLU; ++u) {
      const bool in_range = unit < units;
      const OperandView& w = is_b ? vb : va;
      const int row = pq * 4 + r2, kcol = pr * 4 + kq;
      const bool live = in_range && row < w.n_rows;
      off[u] = live ? (int)(w.tiles - va.tiles) + (4 * h) * w.ts + tile_off(row, kcol, sbo) : -1;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live && kcol < nj)
        v[u] = __ldg(reinterpret_cast<const float4*>(
            w.src + 4 * h + ((size_t)row * w.s_row + (size_t)kcol * w.s_k) * dense));
      // next unit of this warp
      unit += kWarps;
      if (!is_b && unit >= units_a) {
        is_b = true;
        local = unit - units_a;
        pq = local / kquads;
        pr = local - pq * kquads;
      } else {
        pq += dq;
        pr += dr;
        if (pr >= kquads) { pr -= kquads; ++pq; }
      }
    }
#pragma unroll
    for (int u = 0; u < kLU; ++u) {
      if (off[u] >= 0) {
        const int ts = (u0 + u * kWarps) >= units_a ? vb.ts : va.ts;
        unsigned char* t = va.tiles + off[u];
        *reinterpret_cast<float*>(t) = v[u].x;
        *reinterpret_cast<float*>(t + ts) = v[u].y;
        *reinterpret_cast<float*>(t + 2 * ts) = v[u].z;
        *reinterpret_cast<float*>(t + 3 * ts) = v[u].w;
      }
    }
  }
}

template <int CH>
__device__ __forceinline__ void epilogue_round(uint32_t tmem_base, int warp, int n_pad, int n_k_valid,
                                               bool row_ok, const unsigned char* mrow,
                                               float* orow, int dense) {
  const int kchunks = (n_k_valid + 7) / 8;
  // warps w and w+4 share TMEM quarter (w & 3); they take alternate k-chunks
  for (int kc = (warp >> 2); kc < kchunks; kc += 2) {
    float v[CH][8];
#pragma unroll
    for (int cc = 0; cc < CH; ++cc)
      tmem_ld8(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cc * n_pad + kc * 8), v[cc]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row_ok) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int k = kc * 8 + kk;
        if (k < n_k_valid) {
          const bool m = mrow[k] != 0;
          float* o = orow + (size_t)k * dense;
#pragma unroll
          for (int q = 0; q < CH / 4; ++q) {
            const float4 w = m ? make_float4(v[4 * q][kk], v[4 * q + 1][kk], v[4 * q + 2][kk], v[4 * q + 3][kk])
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(o + 4 * q) = w;
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kTcThreads, 2)
mamamm_tc_kernel(const float* __restrict__ A, const float* __restrict__ B,
                 const unsigned char* __restrict__ mask, const int* __restrict__ ext, int dense,
                 TcGeom g, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar_storage;
  __shared__ uint32_t tmem_base_holder;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slabs = dense / kCS;
  const int b = blockIdx.x / slabs;
  const int c0 = (blockIdx.x % slabs) * kCS;
  const uint32_t bar = smem_u32(&mbar_storage);

  // valid extents of this graph (operand pads are zero, so only [0,ni) x [0,nj) x [0,nk) matters)
  int ni = g.n_i, nj = g.n_j, nk = g.n_k;
  if (ext) {
    ni = min(max(__ldg(ext + 3 * b), 0), g.n_i);
    nj = min(max(__ldg(ext + 3 * b + 1), 0), g.n_j);
    nk = min(max(__ldg(ext + 3 * b + 2), 0), g.n_k);
  }
  float* ob = out + (size_t)b * g.n_i * g.n_k * dense + c0;

  // zeros outside the valid rectangle (fire-and-forget stores, overlap with everything below)
  {
    const int h = tid & 1;
    const bool empty = ni == 0 || nj == 0 || nk == 0;
    const int step = kTcThreads / 2, di = step / g.n_k, dk = step - di * g.n_k;
    int p = tid >> 1;
    int i = p / g.n_k, k = p - i * g.n_k;
    for (; p < g.n_i * g.n_k; p += step) {
      if (empty || i >= ni || k >= nk)
        *reinterpret_cast<float4*>(ob + (size_t)p * dense + 4 * h) = make_float4(0.f, 0.f, 0.f, 0.f);
      i += di;
      k += dk;
      if (k >= g.n_k) { k -= g.n_k; ++i; }
    }
    if (empty) return;
  }

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_holder)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  const int k_pad = (nj + 7) & ~7;            // K steps of 8, zero padded
  const int n_pad = max((nk + 15) & ~15, 16); // UMMA N (multiple of 16 at M = 128)
  const int ch_round = (kCS * n_pad <= kTmemCols) ? 8 : 4;

  // ---- load phase: global (b, i, j, c0..c0+7) -> 8 per-channel UMMA tiles ----------------
  // the mask rows are staged too (the epilogue must not pay a global-load latency per store);
  // their loads are issued first so that they overlap the operand loads
  {
    const unsigned char* mb = mask + (size_t)b * g.n_i * g.n_k;
    unsigned char mv[8];
    const int total = ni * g.n_k;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = tid + q * kTcThreads;
      mv[q] = t < total ? __ldg(mb + t) : 0;
    }
    OperandView va{A + (size_t)b * g.n_i * g.n_j * dense + c0, smem, ni, g.ts_a, g.sa_i, g.sa_j};
    OperandView vb{B + (size_t)b * g.n_j * g.n_k * dense + c0, smem + g.off_b, nk, g.ts_b, g.sb_k, g.sb_j};
    load_tiles(va, vb, nj, k_pad, dense, g.sbo, warp, lane);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = tid + q * kTcThreads;
      if (t < total) smem[g.off_mask + t] = mv[q];
    }
    for (int t = tid + 8 * kTcThreads; t < total; t += kTcThreads) smem[g.off_mask + t] = __ldg(mb + t);
  }
  // make the generic-proxy smem writes visible to the tensor-core (async) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_holder;

  const uint32_t idesc = umma_idesc_tf32(128, n_pad);
  const uint32_t smem_base = smem_u32(smem);
  const int rounds = kCS / ch_round;
  const int row = (warp & 3) * 32 + lane;               // TMEM lane == output row i
  const bool row_ok = row < ni;
  const unsigned char* mrow = smem + g.off_mask + (row_ok ? row : 0) * g.n_k;
  float* orow = ob + ((size_t)(row_ok ? row : 0) * g.n_k) * dense;

  for (int r = 0; r < rounds; ++r) {
    if (tid == 0) {
      for (int cc = 0; cc < ch_round; ++cc) {
        const int ch = r * ch_round + cc;
        const uint32_t ta = smem_base + ch * g.ts_a;
        const uint32_t tb = smem_base + g.off_b + ch * g.ts_b;
        const uint32_t td = tmem_base + (uint32_t)(cc * n_pad);
        for (int ks = 0; ks < k_pad / 8; ++ks) {
          // one K=8 step = 2 core matrices along K
          const uint64_t da = umma_desc(ta + ks * 2 * kCoreBytes, kCoreBytes, g.sbo);
          const uint64_t db = umma_desc(tb + ks * 2 * kCoreBytes, kCoreBytes, g.sbo);
          umma_tf32(td, da, db, idesc, ks > 0 ? 1u : 0u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                   : "memory");
    }
    mbar_wait(bar, (uint32_t)(r & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ch_round == 8)
      epilogue_round<8>(tmem_base, warp, n_pad, nk, row_ok, mrow, orow, dense);
    else
      epilogue_round<4>(tmem_base, warp, n_pad, nk, row_ok, mrow, orow + r * 4, dense);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                 : "memory");
  }
}

int mamamm_tc_launch(const float* A, int trans_a, const float* B, int trans_b,
                     const unsigned char* mask, const int* ext, int64_t b, int64_t n_i,
                     int64_t n_j, int64_t n_k, int64_t dense, float* out, cudaStream_t s) {
  if (dense % kCS != 0) {
    set_error("mamamm algo=1 needs dense %% 8 == 0");
    return -2;
  }
  if (n_i > 128 || n_k > 64 || n_j > 128) {
    set_error("mamamm algo=1 supports n_i <= 128, n_k <= 64, n_j <= 128");
    return -2;
  }
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) |
       reinterpret_cast<uintptr_t>(out)) & 15) {
    set_error("mamamm algo=1 needs 16-byte aligned tensors");
    return -2;
  }
  TcGeom g;
  g.n_i = (int)n_i; g.n_j = (int)n_j; g.n_k = (int)n_k;
  const int k_pad = (int)((n_j + 7) / 8 * 8);
  const int n_pad = (int)((n_k + 15) / 16 * 16);
  g.sbo = k_pad / 4 * kCoreBytes;
  const int groups_a = (int)((n_i + 7) / 8), groups_b = (int)((n_k + 7) / 8);
  // +16 B: four tile strides = 16 banks, so the two channel halves never share a bank
  g.ts_a = groups_a * g.sbo + 16;
  g.ts_b = groups_b * g.sbo + 16;
  g.off_b = (kCS * g.ts_a + 127) / 128 * 128;
  g.sa_i = trans_a ? 1 : n_j;  g.sa_j = trans_a ? n_i : 1;
  g.sb_j = trans_b ? 1 : n_k;  g.sb_k = trans_b ? n_j : 1;
  // the M=128 / N=n_pad tiles read past the stored rows: keep every read inside the buffer
  const int end_a = (kCS - 1) * g.ts_a + 16 * g.sbo;
  const int end_b = g.off_b + (kCS - 1) * g.ts_b + (n_pad / 8) * g.sbo;
  int total = g.off_b + kCS * g.ts_b;
  if (end_a > total) total = end_a;
  if (end_b > total) total = end_b;
  total = (total + 127) / 128 * 128;
  g.off_mask = total;
  total += (int)((n_i * n_k + 127) / 128 * 128);
  if (total > 227 * 1024) {
    set_error("mamamm algo=1: tiles (%d bytes) exceed shared memory", total);
    return -2;
  }
  static int configured = 0;
  if (configured < total) {
    PGH_CUDA(cudaFuncSetAttribute(mamamm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, total));
    configured = total;
  }
  const unsigned grid = (unsigned)(b * (dense / kCS));
  mamamm_tc_kernel<<<grid, kTcThreads, total, s>>>(A, B, mask, ext, (int)dense, g, out);
  return check_launch("mamamm_tc");
}

}  // namespace pgh
