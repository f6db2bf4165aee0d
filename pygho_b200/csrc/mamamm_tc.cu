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
template <int kWarps>
__device__ __forceinline__ void load_tiles(const OperandView& va, const OperandView& vb, int nj,
                                           int k_pad, int dense, int sbo, int warp, int lane) {
  const int h = lane & 1, kq = (lane >> 1) & 3, r2 = lane >> 3;
  const int kquads = k_pad >> 2;
  const int units_a = kquads * ((va.n_rows + 3) >> 2);
  const int units = units_a + kquads * ((vb.n_rows + 3) >> 2);
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
    load_tiles<kTcThreads / 32>(va, vb, nj, k_pad, dense, g.sbo, warp, lane);
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


// ---------------------------------------------------------------------------------------
// Pipelined variant (algo 2): one persistent CTA per SM, three warp roles, the same tile
// layout and MMA shape as above.
//   * producer warps (8 = two groups of 4, group s owns shared-memory stage s and every
//     second work item): 128-bit loads of the two operand slabs, transposed on the fly into
//     the per-channel K-major tiles exactly like the kernel above (load_tiles), plus the
//     item's mask tile; while one group waits for its loads or stores its tiles the other
//     group's loads are in flight, so the load stream does not drain between items.
//     (A 4-byte cp.async scatter was measured first: LDGSTS.32 with 32 distinct destination
//     rows costs ~46 cycles per warp instruction and starves the whole LSU,
//     profiles/r1_mamamm_pipe.md.)
//   * MMA warp (one elected thread): waits for a full stage and a free accumulator, issues
//     the tcgen05.mma K-steps of 4 or 8 channels (one "round") and commits to the
//     accumulator-full barrier; after the last round of an item the commit also releases
//     the shared-memory stage.
//   * epilogue warps (8, two per TMEM lane quarter): zero the pad positions of the item,
//     pull their mask row out of the stage into a 64-bit register, then per round
//     tcgen05.ld -> masked 16/32 B stores and hand the accumulator back.
// Two shared-memory stages and two TMEM accumulators (2 x 256 columns) decouple the roles.
constexpr int kTraceItems = 32;
__device__ unsigned long long* g_trace = nullptr;
__device__ long long g_trace_words = 0;
// profiling hook (pgh_debug_trace): 64-bit %globaltimer stamps of CTA 0, laid out as
// [role 0..2][item slot 0..kTraceItems-1][event 0..3]
__device__ __forceinline__ void trace(unsigned long long* t, int role, int q, int ev) {
  if (t && q < kTraceItems) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    t[(role * kTraceItems + q) * 4 + ev] = now;
  }
}

constexpr int kPipeProducers = 8;   // warps 0..7: two groups of kPipeGroup warps
constexpr int kPipeGroup = 4;
constexpr int kPipeEpi = 8;         // warps 9..16 (warp 8 issues the MMAs)
constexpr int kPipeThreads = (kPipeProducers + 1 + kPipeEpi) * 32;
constexpr int kPipeAccCols = 256;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

struct PipeItem {
  int b, c0, ni, nj, nk;
  bool empty;
};

__device__ __forceinline__ PipeItem pipe_item(int it, int slabs, const int* __restrict__ ext,
                                              const TcGeom& g) {
  PipeItem I;
  I.b = it / slabs;
  I.c0 = (it - I.b * slabs) * kCS;
  I.ni = g.n_i; I.nj = g.n_j; I.nk = g.n_k;
  if (ext) {
    I.ni = min(max(__ldg(ext + 3 * I.b), 0), g.n_i);
    I.nj = min(max(__ldg(ext + 3 * I.b + 1), 0), g.n_j);
    I.nk = min(max(__ldg(ext + 3 * I.b + 2), 0), g.n_k);
  }
  I.empty = I.ni == 0 || I.nj == 0 || I.nk == 0;
  return I;
}

template <int CH>
__device__ __forceinline__ void pipe_epilogue_round(uint32_t tacc, int quarter, int alt, int n_pad,
                                                    int nk, bool row_ok, unsigned long long mbits,
                                                    float* orow, int dense) {
  const int kchunks = (nk + 7) / 8;
  for (int kc = alt; kc < kchunks; kc += 2) {
    float v[CH][8];
#pragma unroll
    for (int cc = 0; cc < CH; ++cc)
      tmem_ld8(tacc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cc * n_pad + kc * 8), v[cc]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row_ok) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int k = kc * 8 + kk;
        if (k < nk) {
          const bool m = (mbits >> k) & 1ull;
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

__global__ void __launch_bounds__(kPipeThreads, 1)
mamamm_tc_pipe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                      const unsigned char* __restrict__ mask, const int* __restrict__ ext,
                      int dense, int n_items, int stage_bytes, int mask_off, TcGeom g,
                      float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  // barriers: full[2], empty[2] (smem stages), tfull[2], tempty[2] (TMEM accumulators)
  __shared__ __align__(8) unsigned long long bars[8];
  __shared__ uint32_t tmem_base_holder;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slabs = dense / kCS;
  const uint32_t bar0 = smem_u32(bars);
#define PGH_BAR(i) (bar0 + 8u * (uint32_t)(i))
  if (tid == 0) {
    mbar_init(PGH_BAR(0), kPipeGroup * 32);      // full: the stage's producer group
    mbar_init(PGH_BAR(1), kPipeGroup * 32);
    mbar_init(PGH_BAR(2), 1 + kPipeEpi);         // empty: MMA commit + mask read by the epilogue
    mbar_init(PGH_BAR(3), 1 + kPipeEpi);
    mbar_init(PGH_BAR(4), 1);                    // tfull: MMA commit
    mbar_init(PGH_BAR(5), 1);
    mbar_init(PGH_BAR(6), kPipeEpi);             // tempty: epilogue warps
    mbar_init(PGH_BAR(7), kPipeEpi);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kPipeProducers) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_holder)),
                 "r"(2 * kPipeAccCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_holder;
  const uint32_t smem_base = smem_u32(smem);
  unsigned long long* tr =
      (blockIdx.x == 0 && lane == 0 && g_trace_words >= 3 * kTraceItems * 4) ? g_trace : nullptr;

  if (warp < kPipeProducers) {
    // ------------------------------------------------------------------ producers
    const int grp = warp / kPipeGroup, gw = warp % kPipeGroup, gtid = gw * 32 + lane;
    unsigned char* stage = smem + (size_t)grp * stage_bytes;
    int q = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const PipeItem I = pipe_item(it, slabs, ext, g);
      if (I.empty) continue;
      if ((q & 1) == grp) {
        if (gw == 0) trace(tr, 0, q, 0);
        mbar_wait(PGH_BAR(2 + grp), (uint32_t)(((q >> 1) & 1) ^ 1));    // stage free?
        if (gw == 0) trace(tr, 0, q, 1);
        const int k_pad = (I.nj + 7) & ~7;
        const unsigned char* mb = mask + (size_t)I.b * g.n_i * g.n_k;
        constexpr int kMV = 16;
        unsigned char mv[kMV];
        const int total = I.ni * g.n_k;
#pragma unroll
        for (int m = 0; m < kMV; ++m) {
          const int t = gtid + m * kPipeGroup * 32;
          mv[m] = t < total ? __ldg(mb + t) : 0;
        }
        OperandView va{A + (size_t)I.b * g.n_i * g.n_j * dense + I.c0, stage, I.ni, g.ts_a, g.sa_i, g.sa_j};
        OperandView vb{B + (size_t)I.b * g.n_j * g.n_k * dense + I.c0, stage + g.off_b, I.nk, g.ts_b, g.sb_k, g.sb_j};
        load_tiles<kPipeGroup>(va, vb, I.nj, k_pad, dense, g.sbo, gw, lane);
        if (gw == 0) trace(tr, 0, q, 2);
#pragma unroll
        for (int m = 0; m < kMV; ++m) {
          const int t = gtid + m * kPipeGroup * 32;
          if (t < total) stage[mask_off + t] = mv[m];
        }
        for (int t = gtid + kMV * kPipeGroup * 32; t < total; t += kPipeGroup * 32)
          stage[mask_off + t] = __ldg(mb + t);
        // generic-proxy writes -> visible to the tensor-core (async) proxy, then publish
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(PGH_BAR(grp));
        if (gw == 0) trace(tr, 0, q, 3);
      }
      ++q;
    }
  } else if (warp == kPipeProducers) {
    // ------------------------------------------------------------------ MMA issue
    if (lane == 0) {
      int q = 0, rr = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const PipeItem I = pipe_item(it, slabs, ext, g);
        if (I.empty) continue;
        const int s = q & 1;
        const int ksteps = ((I.nj + 7) & ~7) / 8;
        const int n_pad = max((I.nk + 15) & ~15, 16);
        const int ch_round = (kCS * n_pad <= kPipeAccCols) ? 8 : 4;
        const uint32_t idesc = umma_idesc_tf32(128, n_pad);
        const uint32_t st = smem_base + (uint32_t)(s * stage_bytes);
        trace(tr, 1, q, 0);
        mbar_wait(PGH_BAR(s), (uint32_t)((q >> 1) & 1));              // operands landed
        trace(tr, 1, q, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int r = 0; r < kCS / ch_round; ++r, ++rr) {
          const int a = rr & 1;
          mbar_wait(PGH_BAR(6 + a), (uint32_t)(((rr >> 1) & 1) ^ 1)); // accumulator drained
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (r == 0) trace(tr, 1, q, 2);
          for (int cc = 0; cc < ch_round; ++cc) {
            const int ch = r * ch_round + cc;
            // descriptors of K step 0; a K=8 step = 2 core matrices = +256 B = +16 in the
            // (address >> 4) start field, which never carries out of its 14 bits here
            uint64_t da = umma_desc(st + (uint32_t)(ch * g.ts_a), kCoreBytes, g.sbo);
            uint64_t db = umma_desc(st + (uint32_t)(g.off_b + ch * g.ts_b), kCoreBytes, g.sbo);
            const uint32_t td = tmem_base + (uint32_t)(a * kPipeAccCols + cc * n_pad);
            umma_tf32(td, da, db, idesc, 0u);
            for (int ks = 1; ks < ksteps; ++ks) {
              da += (2 * kCoreBytes) >> 4;
              db += (2 * kCoreBytes) >> 4;
              umma_tf32(td, da, db, idesc, 1u);
            }
          }
          umma_commit(PGH_BAR(4 + a));                                 // accumulator ready
        }
        umma_commit(PGH_BAR(2 + s));                                   // stage may be refilled
        trace(tr, 1, q, 3);
        ++q;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp & 3;                     // TMEM lanes this warp may read
    const int alt = (warp - (kPipeProducers + 1)) >> 2;
    const int etid = tid - (kPipeProducers + 1) * 32; // 0 .. 255
    const int row = quarter * 32 + lane;              // TMEM lane == output row i
    int rr = 0, q = 0;
    unsigned long long* etr = (quarter == 0 && alt == 0) ? tr : nullptr;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const PipeItem I = pipe_item(it, slabs, ext, g);
      trace(etr, 2, q, 0);
      float* ob = out + (size_t)I.b * g.n_i * g.n_k * dense + I.c0;
      // zeros outside the valid rectangle
      {
        const int h = etid & 1;
        const int step = kPipeEpi * 32 / 2, di = step / g.n_k, dk = step - di * g.n_k;
        int p = etid >> 1;
        int i = p / g.n_k, k = p - i * g.n_k;
        for (; p < g.n_i * g.n_k; p += step) {
          if (I.empty || i >= I.ni || k >= I.nk)
            *reinterpret_cast<float4*>(ob + (size_t)p * dense + 4 * h) = make_float4(0.f, 0.f, 0.f, 0.f);
          i += di;
          k += dk;
          if (k >= g.n_k) { k -= g.n_k; ++i; }
        }
      }
      if (I.empty) continue;
      const int s = q & 1;
      const bool row_ok = row < I.ni;
      // the mask row travels with the operands: read it out of the stage, then let it go
      mbar_wait(PGH_BAR(s), (uint32_t)((q >> 1) & 1));
      unsigned long long mbits = 0ull;
      if (row_ok) {
        const unsigned char* mrow = smem + (size_t)s * stage_bytes + mask_off + row * g.n_k;
        for (int k = 0; k < I.nk; ++k) mbits |= (unsigned long long)(mrow[k] != 0) << k;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(PGH_BAR(2 + s));
      trace(etr, 2, q, 1);
      const int n_pad = max((I.nk + 15) & ~15, 16);
      const int ch_round = (kCS * n_pad <= kPipeAccCols) ? 8 : 4;
      float* orow = ob + ((size_t)(row_ok ? row : 0) * g.n_k) * dense;
      for (int r = 0; r < kCS / ch_round; ++r, ++rr) {
        const int a = rr & 1;
        mbar_wait(PGH_BAR(4 + a), (uint32_t)((rr >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (r == 0) trace(etr, 2, q, 2);
        const uint32_t tacc = tmem_base + (uint32_t)(a * kPipeAccCols);
        if (quarter * 32 < I.ni) {
          if (ch_round == 8)
            pipe_epilogue_round<8>(tacc, quarter, alt, n_pad, I.nk, row_ok, mbits, orow, dense);
          else
            pipe_epilogue_round<4>(tacc, quarter, alt, n_pad, I.nk, row_ok, mbits, orow + r * 4, dense);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(PGH_BAR(6 + a));
      }
      trace(etr, 2, q, 3);
      ++q;
    }
  }
#undef PGH_BAR
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == kPipeProducers) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(2 * kPipeAccCols)
                 : "memory");
  }
}

int mamamm_tc_launch(const float* A, int trans_a, const float* B, int trans_b,
                     const unsigned char* mask, const int* ext, int64_t b, int64_t n_i,
                     int64_t n_j, int64_t n_k, int64_t dense, int pipelined, float* out,
                     cudaStream_t s) {
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
  // two stages of (operand tiles + mask tile) fit: persistent, warp-specialised pipeline
  const int stage = total + (int)((n_i * n_k + 127) / 128 * 128);
  if (pipelined && 2 * stage <= 227 * 1024) {
    static int configured_pipe = 0;
    if (configured_pipe < 2 * stage) {
      PGH_CUDA(cudaFuncSetAttribute(mamamm_tc_pipe_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * stage));
      configured_pipe = 2 * stage;
    }
    const int64_t n_items = b * (dense / kCS);
    const unsigned grid = (unsigned)(n_items < kSMs ? n_items : kSMs);
    mamamm_tc_pipe_kernel<<<grid, kPipeThreads, 2 * stage, s>>>(A, B, mask, ext, (int)dense,
                                                                (int)n_items, stage, total, g, out);
    return check_launch("mamamm_tc_pipe");
  }
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

extern "C" int pgh_debug_trace(void* device_buf, int64_t n_words) {
  unsigned long long* p = static_cast<unsigned long long*>(device_buf);
  long long n = device_buf ? (long long)n_words : 0;
  PGH_CUDA(cudaMemcpyToSymbol(pgh::g_trace, &p, sizeof(p)));
  PGH_CUDA(cudaMemcpyToSymbol(pgh::g_trace_words, &n, sizeof(n)));
  return 0;
}
