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
#include <stdlib.h>

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
  // try_wait with a suspend-time hint: the waiting warp is parked by the hardware until the
  // phase completes (or the hint expires) instead of spinning -- a spinning waiter competes
  // for issue slots with the working warps of its scheduler, which made every role of the
  // pipelined kernel several times slower.  Bounded: a lost arrival traps instead of hanging.
  for (uint32_t spin = 0; spin < (1u << 12); ++spin) {
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
// Pipelined variant (algo 2): one persistent CTA per SM, four warp roles, the same tile
// layout and MMA shape as above.  What the one-CTA-per-item kernel cannot do is keep enough
// bytes in flight: a loaded HBM round trip is ~2 us on B200, so an SM needs >= 64 KB of
// outstanding loads *all the time* to see its share of the bandwidth, while a register-
// staged load phase holds ~40 KB for a fraction of the CTA's life (profiles/r1_mamamm_pipe.md).
//   * copy warps (3): walk the CTA's work items (graph, 8-channel slab) and streams both
//     operand slabs global -> shared with 16-byte cp.async (LDGSTS.128, one instruction per
//     16 positions) into a byte ring of ~120 KB: no registers, no waiting, 3-4 typical items
//     in flight.  Completion is signalled per item by cp.async.mbarrier.arrive.
//   * transposer warps (8): wait for an item to land, re-lay it from the ring into the 8
//     per-channel K-major UMMA tiles (LDS.128 + 4 STS.32 in the bank-conflict-free lane
//     mapping of load_tiles), stage the mask tile, hand the ring bytes back.
//   * MMA warp: warp-uniform control flow, one elected lane issues the tcgen05.mma K-steps
//     of 4 or 8 channels (one "round") into one of two TMEM accumulators and commits.
//   * epilogue warps (8, two per TMEM lane quarter): mask row -> 64-bit register, then per
//     round tcgen05.ld -> masked 32-byte (256-bit) stores and hand the accumulator back.
//     The warps whose quarter has no valid row of the item write the zeros of its pad
//     positions afterwards (whole dense rows, shared among the slab items of a graph).
constexpr int kTraceItems = 32;
__device__ unsigned long long* g_trace = nullptr;
__device__ long long g_trace_words = 0;
// profiling hook (pgh_debug_trace): 64-bit %globaltimer stamps of CTA 0, laid out as
// [role 0..3][item slot 0..kTraceItems-1][event 0..3]
__device__ __forceinline__ void trace(unsigned long long* t, int role, int q, int ev) {
  if (t && q < kTraceItems) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    t[(role * kTraceItems + q) * 4 + ev] = now;
  }
}

// Roles are laid out by scheduler (warp & 3 = SM sub-partition): the five roles run five
// different loops, and a sub-partition whose warps all run the same loop keeps it in its
// instruction cache.  The epilogue has to sit on all four (a warp can only read the TMEM lanes
// of its own quarter), but for graphs of <= 32 nodes only quarter 0 has rows.
//   warps 0..7   epilogue (quarter = warp & 3, two per quarter)
//   warps 8..19  scheduler 1, 2: transposers; scheduler 3: copy; scheduler 0: MMA + 2 transposers
constexpr int kPipeCopy = 3;
constexpr int kPipeXpose = 8;
constexpr int kPipeEpi = 8;
constexpr int kPipeThreads = 20 * 32;
enum PipeRole { kRoleEpi, kRoleXpose, kRoleCopy, kRoleMma };
constexpr int kPipeAccCols = 256;
constexpr int kRingSlots = 8;       // items in flight in the staging ring (barrier pairs)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(saddr)
               : "memory");
  return v;
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

// staging of one item in the ring: operand A rows at [0, ni * pitch), operand B rows after
// them; a row holds k_pad positions of 32 B (8 channels) plus 16 B of padding, which spreads
// the 16 B pieces of a column-wise copy over all banks
struct PipeStage {
  int pitch, off_b, bytes;
};
__device__ __forceinline__ PipeStage pipe_stage(const PipeItem& I) {
  PipeStage S;
  S.pitch = ((I.nj + 7) & ~7) * 32 + 16;
  S.off_b = I.ni * S.pitch;
  S.bytes = ((I.ni + I.nk) * S.pitch + 127) & ~127;
  return S;
}
// every role replays the same allocation rule, so ring offsets need no communication
__device__ __forceinline__ int ring_place(int& head, int bytes, int cap) {
  const int off = (head + bytes > cap) ? 0 : head;
  head = off + bytes;
  return off;
}

// global -> ring: positions are walked along the dimension that is contiguous in global
// memory (stride one position = dense floats); copy warp cw takes every kPipeCopy-th outer
// row, lane = (position & 15, 16-byte half).  One warp issues ~0.2 instructions per cycle
// on dependent address arithmetic, hence several copy warps and a two-add inner loop.
__device__ __forceinline__ void pipe_copy_operand(const float* __restrict__ src, uint32_t stg,
                                                  int n_rows, int nj, long long s_row,
                                                  long long s_k, int dense, int pitch, int cw,
                                                  int lane) {
  const int h = lane & 1, p = lane >> 1;
  const bool k_inner = (s_k == 1);
  const int n_in = k_inner ? nj : n_rows, n_out = k_inner ? n_rows : nj;
  const size_t pos_bytes = (size_t)dense * 4;
  const size_t out_bytes = (size_t)(k_inner ? s_row : s_k) * pos_bytes;
  const uint32_t in_step = k_inner ? 32u : (uint32_t)pitch, out_step = k_inner ? (uint32_t)pitch : 32u;
  const char* gp_row = reinterpret_cast<const char*>(src) + h * 16 + (size_t)p * pos_bytes +
                       (size_t)cw * out_bytes;
  uint32_t dst_row = stg + (uint32_t)h * 16u + (uint32_t)p * in_step + (uint32_t)cw * out_step;
  const size_t g16 = 16 * pos_bytes, g_out = kPipeCopy * out_bytes;
  const uint32_t s16 = 16u * in_step, s_out = kPipeCopy * out_step;
  const bool p0 = p < n_in, p1 = p + 16 < n_in;
  for (int o = cw; o < n_out; o += kPipeCopy) {
    if (p0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_row), "l"(gp_row) : "memory");
    if (p1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_row + s16), "l"(gp_row + g16) : "memory");
    if (n_in > 32) {
      const char* gp = gp_row + 2 * g16;
      uint32_t dst = dst_row + 2 * s16;
      for (int t = p + 32; t < n_in; t += 16) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gp) : "memory");
        gp += g16;
        dst += s16;
      }
    }
    gp_row += g_out;
    dst_row += s_out;
  }
}

// ring -> the 8 channel tiles of both operands.  A warp takes a patch row (4 tile rows of one
// operand) and walks its K quads; lane = (r2, kq, h) as in load_tiles, so each LDS.128
// quarter-warp reads 128 contiguous bytes and every STS.32 hits 32 different banks.
template <int kWarps>
__device__ __forceinline__ void pipe_transpose(uint32_t stg_a, uint32_t stg_b, int pitch,
                                               unsigned char* tiles_a, unsigned char* tiles_b,
                                               int rows_a, int rows_b, int ts_a, int ts_b, int nj,
                                               int k_pad, int sbo, int warp, int lane) {
  const int h = lane & 1, kq = (lane >> 1) & 3, r2 = lane >> 3;
  const int kquads = k_pad >> 2;
  const int npa = (rows_a + 3) >> 2, npb = (rows_b + 3) >> 2;
  for (int pp = warp; pp < npa + npb; pp += kWarps) {
    const bool is_b = pp >= npa;
    const int row = (is_b ? pp - npa : pp) * 4 + r2;
    if (row >= (is_b ? rows_b : rows_a)) continue;
    const int ts = is_b ? ts_b : ts_a;
    uint32_t src = (is_b ? stg_b : stg_a) + (uint32_t)(row * pitch + kq * 32 + h * 16);
    unsigned char* dst = (is_b ? tiles_b : tiles_a) + (4 * h) * ts + (row >> 3) * sbo + (row & 7) * 16 + kq * 4;
    const int nq_live = (nj - kq + 3) >> 2;            // quads whose kcol = 4 pr + kq < nj
#pragma unroll 2
    for (int pr = 0; pr < kquads; ++pr) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pr < nq_live) v = lds128(src);
      *reinterpret_cast<float*>(dst) = v.x;
      *reinterpret_cast<float*>(dst + ts) = v.y;
      *reinterpret_cast<float*>(dst + 2 * ts) = v.z;
      *reinterpret_cast<float*>(dst + 3 * ts) = v.w;
      src += 128u;
      dst += kCoreBytes;
    }
  }
}

// zeros outside the valid rectangle of one item.  The 32-byte pieces of one slab would cost
// one L2 request each; instead the `slabs` items of a graph share the pad positions among them
// (position p = i * n_k + k belongs to slab p % slabs) and write whole dense rows, 128 B per
// request.  The 32 lanes test 32 candidate positions at once (one division per lane, in
// parallel), then the warp visits them and stores one 512 B row per pad position.
__device__ __forceinline__ void pipe_zero_item(const PipeItem& I, const TcGeom& g, int slabs,
                                               int dense, float* __restrict__ out, int zr, int nz,
                                               int lane) {
  const int ni = I.empty ? 0 : I.ni, nk = I.nk;
  const int sidx = I.c0 / kCS;
  const int total = g.n_i * g.n_k;
  float* og = out + (size_t)I.b * total * dense + lane * 4;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  // candidate m of this slab is position sidx + m * slabs; warp zr takes m = zr (mod nz)
  for (int m0 = zr; sidx + m0 * slabs < total; m0 += 32 * nz) {
    const int p = sidx + (m0 + lane * nz) * slabs;
    const int i = p / g.n_k, k = p - i * g.n_k;
    const bool pad = p < total && (i >= ni || k >= nk);
    unsigned todo = __ballot_sync(0xffffffffu, pad);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      float* o = og + (size_t)(sidx + (m0 + src * nz) * slabs) * dense;
      for (int c = lane * 4; c < dense; c += 128) *reinterpret_cast<float4*>(o + (c - lane * 4)) = z;
    }
  }
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}

__device__ __forceinline__ void st256(float* p, const float4& a, const float4& b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y),
               "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}

template <int CH, bool WIDE>
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
      float* o = orow + (size_t)(kc * 8) * dense;
      const unsigned have = (unsigned)(mbits >> (kc * 8)) & 0xffu;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        if (kc * 8 + kk < nk) {
          const bool m = (have >> kk) & 1u;
          float4 w[CH / 4];
#pragma unroll
          for (int q = 0; q < CH / 4; ++q)
            w[q] = m ? make_float4(v[4 * q][kk], v[4 * q + 1][kk], v[4 * q + 2][kk], v[4 * q + 3][kk])
                     : make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (CH == 8 && WIDE) {
            st256(o, w[0], w[1]);
          } else {
#pragma unroll
            for (int q = 0; q < CH / 4; ++q) *reinterpret_cast<float4*>(o + 4 * q) = w[q];
          }
        }
        o += dense;
      }
    }
  }
}

template <bool WIDE>
__global__ void __launch_bounds__(kPipeThreads, 1)
mamamm_tc_pipe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                      const unsigned char* __restrict__ mask, const int* __restrict__ ext,
                      int dense, int n_items, int mask_off, int mask_stride, int ring_off, int ring_bytes,
                      TcGeom g,
                      int dbg, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  // barriers: [0,8) ring full, [8,16) ring freed, 16 tiles full, 17 tiles empty,
  //           18,19 accumulator full, 20,21 accumulator empty, 22..24 mask tile full
  __shared__ __align__(8) unsigned long long bars[2 * kRingSlots + 9];
  __shared__ uint32_t tmem_base_holder;
  __shared__ int ring_meta[kPipeCopy * 2 * kRingSlots];   // per copy warp: (offset, bytes) of items in flight

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int role = kRoleEpi, ridx = warp;            // ridx: index of the warp inside its role
  if (warp >= kPipeEpi) {
    const int sch = warp & 3, j = (warp - kPipeEpi) >> 2;
    if (sch == 1 || sch == 2) { role = kRoleXpose; ridx = (sch - 1) * 3 + j; }
    else if (sch == 3) { role = kRoleCopy; ridx = j; }
    else if (j == 0) { role = kRoleMma; ridx = 0; }
    else { role = kRoleXpose; ridx = 5 + j; }
  }
  const int slabs = dense / kCS;
  const uint32_t bar0 = smem_u32(bars);
#define PGH_BAR(i) (bar0 + 8u * (uint32_t)(i))
  constexpr int kFull = 0, kFreed = kRingSlots, kTilesFull = 2 * kRingSlots,
                kTilesEmpty = kTilesFull + 1, kAccFull = kTilesFull + 2, kAccEmpty = kTilesFull + 4,
                kMaskFull = kTilesFull + 6;
  if (tid == 0) {
    for (int k = 0; k < kRingSlots; ++k) {
      mbar_init(PGH_BAR(kFull + k), kPipeCopy * 32);   // cp.async completion of every copy lane
      mbar_init(PGH_BAR(kFreed + k), kPipeXpose);      // transposer warps are done reading
    }
    mbar_init(PGH_BAR(kTilesFull), kPipeXpose * 32);
    mbar_init(PGH_BAR(kTilesEmpty), 1);                // MMA commit
    mbar_init(PGH_BAR(kAccFull), 1);
    mbar_init(PGH_BAR(kAccFull + 1), 1);
    mbar_init(PGH_BAR(kAccEmpty), kPipeEpi);
    mbar_init(PGH_BAR(kAccEmpty + 1), kPipeEpi);
    for (int k = 0; k < 3; ++k) mbar_init(PGH_BAR(kMaskFull + k), kPipeXpose * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (role == kRoleMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_holder)),
                 "r"(2 * kPipeAccCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring_base = smem_base + (uint32_t)ring_off;
  unsigned long long* tr =
      (blockIdx.x == 0 && lane == 0 && g_trace_words >= 4 * kTraceItems * 4) ? g_trace : nullptr;

  if (role == kRoleCopy) {
    // ------------------------------------------------------------------ copy warps
    // every copy warp replays the same ring bookkeeping (private copy of the in-flight list)
    const int cw = ridx;
    int* meta = ring_meta + cw * 2 * kRingSlots;
    int q = 0, oldest = 0, head = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const PipeItem I = pipe_item(it, slabs, ext, g);
      if (I.empty) continue;
      if (cw == 0) trace(tr, 0, q, 0);
      const PipeStage S = pipe_stage(I);
      const int off = ring_place(head, S.bytes, ring_bytes);
      // wait until no item still in the ring overlaps [off, off + bytes)
      for (;;) {
        bool clash = (q - oldest) >= kRingSlots;
        for (int j = oldest; j < q && !clash; ++j) {
          const int o = meta[2 * (j % kRingSlots)], sz = meta[2 * (j % kRingSlots) + 1];
          clash = !(o + sz <= off || off + S.bytes <= o);
        }
        if (!clash) break;
        mbar_wait(PGH_BAR(kFreed + oldest % kRingSlots), (uint32_t)((oldest / kRingSlots) & 1));
        ++oldest;
      }
      if (cw == 0) trace(tr, 0, q, 1);
      const uint32_t stg = ring_base + (uint32_t)off;
      if (!(dbg & 4)) {
      pipe_copy_operand(A + (size_t)I.b * g.n_i * g.n_j * dense + I.c0, stg, I.ni, I.nj, g.sa_i,
                        g.sa_j, dense, S.pitch, cw, lane);
      pipe_copy_operand(B + (size_t)I.b * g.n_j * g.n_k * dense + I.c0, stg + (uint32_t)S.off_b,
                        I.nk, I.nj, g.sb_k, g.sb_j, dense, S.pitch, cw, lane);
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(
                       PGH_BAR(kFull + q % kRingSlots))
                   : "memory");
      __syncwarp();
      meta[2 * (q % kRingSlots)] = off;
      meta[2 * (q % kRingSlots) + 1] = S.bytes;
      __syncwarp();
      if (cw == 0) trace(tr, 0, q, 2);
      ++q;
    }
  } else if (role == kRoleXpose) {
    // ------------------------------------------------------------------ transposers
    const int tw = ridx, ttid = tw * 32 + lane;
    int q = 0, head = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const PipeItem I = pipe_item(it, slabs, ext, g);
      if (I.empty) continue;
      const PipeStage S = pipe_stage(I);
      const int off = ring_place(head, S.bytes, ring_bytes);
      // mask bytes of the item: issued now, consumed after the operands have landed
      const unsigned char* mb = mask + (size_t)I.b * g.n_i * g.n_k;
      // warp tw takes rows tw, tw + 8, ...: lane = column k (and k + 32); the bytes are
      // requested now, balloted into one 64-bit word per row after the operands have landed
      constexpr int kMR = 4;                    // rows per warp held in registers (n_i <= 32)
      unsigned char mlo[kMR], mhi[kMR];
#pragma unroll
      for (int m = 0; m < kMR; ++m) {
        const int r = tw + m * kPipeXpose;
        mlo[m] = (r < I.ni && lane < I.nk) ? __ldg(mb + r * g.n_k + lane) : 0;
        mhi[m] = (r < I.ni && lane + 32 < I.nk) ? __ldg(mb + r * g.n_k + lane + 32) : 0;
      }
      if (tw == 0) trace(tr, 1, q, 0);
      mbar_wait(PGH_BAR(kFull + q % kRingSlots), (uint32_t)((q / kRingSlots) & 1));   // landed
      if (tw == 0) trace(tr, 1, q, 1);
      mbar_wait(PGH_BAR(kTilesEmpty), (uint32_t)((q & 1) ^ 1));                       // tiles free
      if (tw == 0) trace(tr, 1, q, 2);
      const uint32_t stg = ring_base + (uint32_t)off;
      pipe_transpose<kPipeXpose>(stg, stg + (uint32_t)S.off_b, S.pitch, smem, smem + g.off_b, I.ni,
                                 I.nk, g.ts_a, g.ts_b, I.nj, (I.nj + 7) & ~7, g.sbo, tw, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(PGH_BAR(kFreed + q % kRingSlots));                   // ring bytes back
      // three mask tiles in rotation: the epilogue reads tile q % 3 at the start of item q,
      // and the MMAs of item q + 2 cannot be issued before it has drained an accumulator of
      // item q, so tile q % 3 is free again when item q + 3 is transposed
      unsigned long long* mtile = reinterpret_cast<unsigned long long*>(smem + mask_off + (q % 3) * mask_stride);
#pragma unroll
      for (int m = 0; m < kMR; ++m) {
        const int r = tw + m * kPipeXpose;
        const unsigned lo = __ballot_sync(0xffffffffu, mlo[m] != 0), hi = __ballot_sync(0xffffffffu, mhi[m] != 0);
        if (r < I.ni && lane == 0) mtile[r] = ((unsigned long long)hi << 32) | lo;
      }
      for (int r = tw + kMR * kPipeXpose; r < I.ni; r += kPipeXpose) {     // graphs of > 32 nodes
        const unsigned lo = __ballot_sync(0xffffffffu, lane < I.nk && __ldg(mb + r * g.n_k + lane) != 0);
        const unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < I.nk && __ldg(mb + r * g.n_k + lane + 32) != 0);
        if (lane == 0) mtile[r] = ((unsigned long long)hi << 32) | lo;
      }
      // the epilogue may be two items behind, which a one-bit phase cannot express on the
      // tiles barrier: each mask tile has its own barrier (one phase per three items)
      mbar_arrive(PGH_BAR(kMaskFull + q % 3));
      // generic-proxy writes -> visible to the tensor-core (async) proxy, then publish
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(PGH_BAR(kTilesFull));
      if (tw == 0) trace(tr, 1, q, 3);
      ++q;
    }
  } else if (role == kRoleMma) {
    // ------------------------------------------------------------------ MMA issue
    // the whole warp runs the (warp-uniform) control flow; values that feed the descriptors
    // are broadcast from lane 0 so that they live in uniform registers
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_holder, 0);
    int q = 0, rr = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const PipeItem I = pipe_item(it, slabs, ext, g);
      const int nj = __shfl_sync(0xffffffffu, I.nj, 0), nk = __shfl_sync(0xffffffffu, I.nk, 0);
      const int ni = __shfl_sync(0xffffffffu, I.ni, 0);
      if (ni == 0 || nj == 0 || nk == 0) continue;
      const int ksteps = ((nj + 7) & ~7) / 8;
      const int n_pad = max((nk + 15) & ~15, 16);
      const int ch_round = (kCS * n_pad <= kPipeAccCols) ? 8 : 4;
      const uint32_t idesc = umma_idesc_tf32(128, n_pad);
      trace(tr, 2, q, 0);
      mbar_wait(PGH_BAR(kTilesFull), (uint32_t)(q & 1));                // tiles written
      trace(tr, 2, q, 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int r = 0; r < kCS / ch_round; ++r, ++rr) {
        const int a = rr & 1;
        mbar_wait(PGH_BAR(kAccEmpty + a), (uint32_t)(((rr >> 1) & 1) ^ 1));   // accumulator drained
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (r == 0) trace(tr, 2, q, 2);
        if (elect_one()) {
          for (int cc = 0; cc < ch_round; ++cc) {
            const int ch = r * ch_round + cc;
            // descriptors of K step 0; a K=8 step = 2 core matrices = +256 B = +16 in the
            // (address >> 4) start field, which never carries out of its 14 bits here
            uint64_t da = umma_desc(smem_base + (uint32_t)(ch * g.ts_a), kCoreBytes, g.sbo);
            uint64_t db = umma_desc(smem_base + (uint32_t)(g.off_b + ch * g.ts_b), kCoreBytes, g.sbo);
            const uint32_t td = tmem_base + (uint32_t)(a * kPipeAccCols + cc * n_pad);
            umma_tf32(td, da, db, idesc, 0u);
            for (int ks = 1; ks < ksteps; ++ks) {
              da += (2 * kCoreBytes) >> 4;
              db += (2 * kCoreBytes) >> 4;
              umma_tf32(td, da, db, idesc, 1u);
            }
          }
          umma_commit(PGH_BAR(kAccFull + a));                             // accumulator ready
          if (r == kCS / ch_round - 1) umma_commit(PGH_BAR(kTilesEmpty)); // tiles may be rewritten
        }
        __syncwarp();
      }
      trace(tr, 2, q, 3);
      ++q;
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const uint32_t tmem_base = tmem_base_holder;
    const int quarter = warp & 3;                     // TMEM lanes this warp may read
    const int alt = warp >> 2;
    const int row = quarter * 32 + lane;              // TMEM lane == output row i
    int rr = 0, q = 0;
    unsigned long long* etr = (quarter == 0 && alt == 0) ? tr : nullptr;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const PipeItem I = pipe_item(it, slabs, ext, g);
      // the zeros of the item's pad positions are written by the epilogue warps whose TMEM
      // quarter holds no valid row of this item (6 of 8 for graphs of <= 32 nodes; all of
      // them if every quarter is busy), after their barrier duties for the item: they stay
      // in step with the pipeline but never hold it up
      const int busy_q = I.empty ? 0 : (I.ni + 31) >> 5;
      const int nz = busy_q < 4 ? 2 * (4 - busy_q) : kPipeEpi;
      const int zr = busy_q >= 4 ? warp : (quarter >= busy_q ? (quarter - busy_q) * 2 + alt : -1);
      if (I.empty) {
        if (!(dbg & 1)) pipe_zero_item(I, g, slabs, dense, out, zr, nz, lane);
        continue;
      }
      trace(etr, 3, q, 0);
      float* ob = out + (size_t)I.b * g.n_i * g.n_k * dense + I.c0;
      const bool row_ok = row < I.ni;
      // the mask tile is written with the operand tiles: take this row's bits
      mbar_wait(PGH_BAR(kMaskFull + q % 3), (uint32_t)((q / 3) & 1));
      unsigned long long mbits = 0ull;
      if (row_ok)
        mbits = reinterpret_cast<const unsigned long long*>(smem + mask_off + (q % 3) * mask_stride)[row];
      trace(etr, 3, q, 1);
      const int n_pad = max((I.nk + 15) & ~15, 16);
      const int ch_round = (kCS * n_pad <= kPipeAccCols) ? 8 : 4;
      float* orow = ob + ((size_t)(row_ok ? row : 0) * g.n_k) * dense;
      for (int r = 0; r < kCS / ch_round; ++r, ++rr) {
        const int a = rr & 1;
        mbar_wait(PGH_BAR(kAccFull + a), (uint32_t)((rr >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (r == 0) trace(etr, 3, q, 2);
        const uint32_t tacc = tmem_base + (uint32_t)(a * kPipeAccCols);
        if (quarter * 32 < I.ni && !(dbg & 2)) {
          if (ch_round == 8)
            pipe_epilogue_round<8, WIDE>(tacc, quarter, alt, n_pad, I.nk, row_ok, mbits, orow, dense);
          else
            pipe_epilogue_round<4, WIDE>(tacc, quarter, alt, n_pad, I.nk, row_ok, mbits, orow + r * 4, dense);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(PGH_BAR(kAccEmpty + a));
      }
      trace(etr, 3, q, 3);
      if (zr >= 0 && !(dbg & 1)) pipe_zero_item(I, g, slabs, dense, out, zr, nz, lane);
      ++q;
    }
  }
#undef PGH_BAR
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (role == kRoleMma) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_holder),
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
  // persistent pipeline: one set of tiles + mask tile, the rest of shared memory is the
  // staging ring, which must hold at least the largest possible item
  {
    const int mask_off = total;
    const int mask_stride = (int)((n_i * 8 + 127) / 128 * 128);        // one 64-bit word per row
    const int ring_off = total + 3 * mask_stride;
    const int ring_bytes = (227 * 1024 - 1024 - ring_off) & ~127;    // 1 KB: static shared memory
    const int64_t max_item = ((n_i + n_k) * (int64_t)(k_pad * 32 + 16) + 127) / 128 * 128;
    if (pipelined && ring_bytes >= max_item) {
      const int smem_bytes = ring_off + ring_bytes;
      const bool wide = (reinterpret_cast<uintptr_t>(out) & 31) == 0;
      static bool configured_pipe = false;
      if (!configured_pipe) {
        PGH_CUDA(cudaFuncSetAttribute(mamamm_tc_pipe_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        PGH_CUDA(cudaFuncSetAttribute(mamamm_tc_pipe_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        configured_pipe = true;
      }
      const int64_t n_items = b * (dense / kCS);
      const unsigned grid = (unsigned)(n_items < kSMs ? n_items : kSMs);
      // timing experiments only (results are wrong): 1 = no zero fill, 2 = no epilogue, 4 = no copies
      static const int dbg = [] { const char* e = getenv("PYGHO_B200_PIPE_DEBUG"); return e ? atoi(e) : 0; }();
      if (wide)
        mamamm_tc_pipe_kernel<true><<<grid, kPipeThreads, smem_bytes, s>>>(
            A, B, mask, ext, (int)dense, (int)n_items, mask_off, mask_stride, ring_off, ring_bytes, g, dbg,
            out);
      else
        mamamm_tc_pipe_kernel<false><<<grid, kPipeThreads, smem_bytes, s>>>(
            A, B, mask, ext, (int)dense, (int)n_items, mask_off, mask_stride, ring_off, ring_bytes, g, dbg,
            out);
      return check_launch("mamamm_tc_pipe");
    }
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

namespace pgh {
// the same buffer, for kernels in other translation units (passed to them as a kernel argument)
unsigned long long* g_trace_host = nullptr;
long long g_trace_host_words = 0;
}

extern "C" int pgh_debug_trace(void* device_buf, int64_t n_words) {
  unsigned long long* p = static_cast<unsigned long long*>(device_buf);
  long long n = device_buf ? (long long)n_words : 0;
  pgh::g_trace_host = p;
  pgh::g_trace_host_words = n;
  PGH_CUDA(cudaMemcpyToSymbol(pgh::g_trace, &p, sizeof(p)));
  PGH_CUDA(cudaMemcpyToSymbol(pgh::g_trace_words, &n, sizeof(n)));
  return 0;
}
