// 2-FWL contraction of MaskedTensors (reference backend/Mamamm.py:7-64), algo 4:
//
//   out[b,i,k,c] = mask[b,i,k] ? sum_j A'[b,i,j,c] * B'[b,j,k,c] : 0        (per channel c)
//
// The work is one tiny (n x n x n, n ~ 23 on ZINC) product per (graph, channel): 0.5 GFLOP of
// useful math against 180 MB of compulsory traffic at b = 128, n <= 40, d = 128.  It is an HBM
// problem, not a tensor-core problem, so this kernel keeps the reference's channel-last layout
// (no transposes into MMA operand layouts) and does exact fp32 FMAs from shared memory:
//
//   warp 7     TMA producer: one thread walks a dynamic queue of (graph, 16-channel slab, row
//              pass) units; for every chunk of JC contraction indices it issues 4-D
//              cp.async.bulk.tensor boxes (16 channels x JC x 8 rows / columns) of A' and B' into a
//              shared-memory ring -- only boxes that intersect the graph's valid extents are
//              fetched, the ring keeps filling across unit boundaries.  The same thread has the
//              bulk-copy engine write the unit's share of the output's pad region from a block of
//              zeros in shared memory (cp.async.bulk shared -> global, one run per output row)
//   warps 0-6  consumers: a thread owns a 4 (rows) x 4 (columns) register tile of 4 (or, for
//              small graphs, 2) channels; 4 (8) lanes cover the 16 channels of a cell, the two
//              lane groups of a quarter (half) warp share the A' rows (broadcast) and take the
//              even / odd columns of an 8-column block, whose cells are an odd number of 64-byte
//              cells apart in both operand layouts (JC = 5 when B' is contraction-major, else 4):
//              all shared-memory loads are conflict free
//   epilogue   mask select (bytes prefetched towards L1 at the unit's first chunk) + vector stores
//              of the valid region
// Two CTAs of 8 warps per SM (113 KB of shared memory each): one CTA's start-up, epilogue and
// queue latency overlap the other's FMAs (profiles/r2_mamamm_smem.md).
//
// Sum order is j ascending with fmaf, the same as the CUDA-core kernel (algo 0): bit-identical.
// Operand pads must be zero (MaskedTensor keeps them at padvalue 0), as for the other kernels.
#include <cuda.h>

#include "common.cuh"

namespace pgh {

constexpr int kMsCH = 16;                    // channels per slab: one 64-byte cell per (row, col)
constexpr int kMsCell = kMsCH * 4;           // bytes
constexpr int kMsRB = 8;                     // rows / columns per TMA box
constexpr int kMsConsumerWarps = 7;          // + producer warp = 8 warps, two CTAs per SM
constexpr int kMsProducerWarp = kMsConsumerWarps;
constexpr int kMsConsumers = kMsConsumerWarps * 32;
constexpr int kMsThreads = kMsConsumers + 32;
constexpr int kMsMaxPairs = kMsConsumers / 8;   // a pair = 4 rows x 8 columns of output
constexpr int kMsMaxStages = 8;
constexpr size_t kMsSmemCap = 113 * 1024;    // dynamic + static per CTA: two CTAs per SM
constexpr int kMsZeroBytes = 4096;           // source of the pad-fill bulk stores
constexpr int kMsQueues = 1;                 // work queues per launch (16 measured slower: the unit order is (graph, slab, pass)-major, a queue per residue gets all the small second passes)

struct MsParams {
  int n_i, n_j, n_k, dense;       // tensor extents of A' (n_i x n_j) and B' (n_j x n_k)
  int nslab, npass;               // units per graph = nslab * npass
  int stages, stage_bytes;
  long long units;                // b * nslab * npass
  unsigned long long* trace;      // profiling hook (pgh_debug_trace): [count, (tag, clock) ...] of CTA 0
  long long trace_words;
  int dbg;                        // profiling only (pgh_set_tuning key 7): 1 no FMAs, 2 no loads, 4 no stores, 8 no pad fill, 16 never two channels per lane, 32 no mask loads, 64 no tile stores
};

// what the producer tells the consumers about a ring stage
struct MsMeta {
  int item, slab, chunk, nchunks, ei, ek, r0, rows, nab;
};

// rows of output one pass covers for a graph whose B' has nkb column blocks: as many 4-row
// tiles as the consumer threads hold next to each other (host and device must agree)
__host__ __device__ __forceinline__ int ms_rows_per_pass(int nkb) {
  const int itp = nkb > 0 ? kMsMaxPairs / nkb : kMsMaxPairs;
  return 4 * (itp > 0 ? itp : 1);
}

__device__ __forceinline__ uint32_t ms_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void ms_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// bounded wait: a lost arrival traps (reported as a launch error) instead of hanging the GPU
__device__ __forceinline__ void ms_mbar_wait(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t done;
    if (hint_ns) {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}\n"
          : "=r"(done)
          : "r"(bar), "r"(parity), "r"(hint_ns)
          : "memory");
    } else {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}\n"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
    }
    if (done) return;
  }
  __trap();
}
// one lane waits for the warp: 32 lanes of 7 warps polling one barrier cost ~900 cycles per ring
// stage (profiles/r2_mamamm_smem.md); the lane's acquire is passed on by the warp barrier
__device__ __forceinline__ void ms_warp_wait(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  if ((threadIdx.x & 31) == 0) ms_mbar_wait(bar, parity, hint_ns);
  __syncwarp();
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
// profiling hook: CTA 0's producer thread (region 0) and first consumer thread (region 1) each
// append (tag, clock64) pairs to their own half of the buffer; word 0 / 1 = the regions' counts
__device__ __forceinline__ void ms_trace(const MsParams& P, int region, unsigned long long tag) {
  if (P.trace && blockIdx.x == 0) {
    const long long half = (P.trace_words - 2) / 2;
    const unsigned long long i = P.trace[region];
    if ((long long)(2 * i + 2) <= half) {
      P.trace[2 + region * half + 2 * i] = tag;
      P.trace[2 + region * half + 2 * i + 1] = (unsigned long long)clock64();
      P.trace[region] = i + 1;
    }
  }
}

struct MsRing {
  uint32_t ring, full, empty;     // shared-memory addresses: stage 0, full_bar[0], empty_bar[0]
  int s;
  uint32_t ph, hint;
};

template <int V> struct MsVec;
template <> struct MsVec<4> {
  float v[4];
  __device__ __forceinline__ void load(uint32_t addr) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
  }
};
template <> struct MsVec<2> {
  float v[2];
  __device__ __forceinline__ void load(uint32_t addr) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr));
  }
};

// One unit (graph, 16-channel slab, row pass) on the consumer warps: all its chunks, then the
// epilogue.  L lanes cover the 16 channels of a cell (L = 4: float4 per lane, L = 8: float2); the
// caller has already waited for the unit's first chunk.
template <bool AK, bool BK, int L>
__device__ __forceinline__ void ms_unit(const MsMeta& m, MsRing& R, const MsParams& P,
                                        const unsigned char* __restrict__ mask, float* __restrict__ out) {
  constexpr int JC = BK ? 5 : 4;
  constexpr int kBoxBytes = kMsRB * JC * kMsCell;
  constexpr int V = kMsCH / L;                          // channels per lane
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid / L, l = tid % L, pair = grp >> 1, par = grp & 1;
  const int nkb = (m.ek + kMsRB - 1) / kMsRB;
  const int it = nkb > 0 ? pair / nkb : 0;
  const int kb = nkb > 0 ? pair - it * nkb : 0;
  const int i0 = it * 4;
  const bool active = i0 < m.rows;
  const int rr0 = i0 & 7;
  const uint32_t offA = (uint32_t)((i0 >> 3) * kBoxBytes + (AK ? rr0 * JC : rr0) * kMsCell + l * V * 4);
  const uint32_t offB =
      (uint32_t)((m.nab + kb) * kBoxBytes + (BK ? par * JC : par) * kMsCell + l * V * 4);
  const int r0 = m.r0;
  // first cell of this thread's tile (row r0 + i0, column kb * 8 + par)
  const long long cellb = ((long long)m.item * P.n_i + r0 + i0) * P.n_k + kb * kMsRB + par;
  if (active && l == 0) {
    // the epilogue's mask bytes: start them towards L1 now, no register held
#pragma unroll
    for (int x = 0; x < 4; ++x)
      if (r0 + i0 + x < m.ei) asm volatile("prefetch.global.L1 [%0];" ::"l"(mask + cellb + (long long)x * P.n_k));
  }
  float acc[4][4][V];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
      for (int c = 0; c < V; ++c) acc[x][y][c] = 0.f;

  const int nst = m.nchunks;
  const bool tr = tid == 0 && P.trace != nullptr;
  for (int ch = 0; ch < nst; ++ch) {
    if (ch > 0) ms_warp_wait(R.full + 8u * R.s, R.ph, R.hint);
    if (tr) ms_trace(P, 1, (2ull << 56) | ((unsigned long long)m.item << 16) | (m.slab << 8) | ch);
    if (active && !(P.dbg & 1)) {
      const uint32_t base = R.ring + (uint32_t)R.s * (uint32_t)P.stage_bytes;
      const uint32_t pa = base + offA, pb = base + offB;
#pragma unroll(L == 8 ? JC : 1)
      for (int jj = 0; jj < JC; ++jj) {
        MsVec<V> a[4], bv[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) a[x].load(pa + (uint32_t)((AK ? x * JC + jj : jj * kMsRB + x) * kMsCell));
#pragma unroll
        for (int y = 0; y < 4; ++y)
          bv[y].load(pb + (uint32_t)((BK ? 2 * y * JC + jj : jj * kMsRB + 2 * y) * kMsCell));
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int c = 0; c < V; ++c) acc[x][y][c] = fmaf(a[x].v[c], bv[y].v[c], acc[x][y][c]);
      }
    }
    __syncwarp();
    if (tr) ms_trace(P, 1, (3ull << 56) | ((unsigned long long)m.item << 16) | (m.slab << 8) | ch);
    if (lane == 0) ms_mbar_arrive(R.empty + 8u * R.s);
    if (++R.s == P.stages) { R.s = 0; R.ph ^= 1u; }
  }
  if (tr) ms_trace(P, 1, (4ull << 56) | ((unsigned long long)m.item << 16) | (m.slab << 8));

  if (active && !(P.dbg & 4)) {
    // ------------------------------------------------------------------ epilogue of the unit
    const unsigned char* mp = mask + cellb;
    float* op = out + cellb * P.dense + m.slab * kMsCH + l * V;
    const int ni = m.ei - (r0 + i0), nk = m.ek - (kb * kMsRB + par);    // valid rows / columns from here
    unsigned int mk = 0;
    if (P.dbg & 32) {
      mk = 0xFFFFu;
    } else {
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y)
          if (x < ni && 2 * y < nk && mp[(long long)x * P.n_k + 2 * y]) mk |= 1u << (x * 4 + y);
    }
    if (P.dbg & 64) {          // profiling: mask loads only, one store
      if (mk == 0x12345u) *op = 1.f;
      return;
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      float* orow = op + (long long)x * P.n_k * P.dense;
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        if (x < ni && 2 * y < nk) {
          const bool on = (mk >> (x * 4 + y)) & 1u;
          float* o = orow + (long long)(2 * y) * P.dense;
          if (V == 4)
            *reinterpret_cast<float4*>(o) = on ? make_float4(acc[x][y][0], acc[x][y][1], acc[x][y][V - 2], acc[x][y][V - 1])
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
          else
            *reinterpret_cast<float2*>(o) = on ? make_float2(acc[x][y][0], acc[x][y][1]) : make_float2(0.f, 0.f);
        }
      }
    }
  }
  if (tr) ms_trace(P, 1, (5ull << 56) | ((unsigned long long)m.item << 16) | (m.slab << 8));
}

// AK / BK: the operand's contraction index is its fast (inner) spatial dim as stored:
//   A' = A   stored (b, n_i, n_j, C) -> AK;   A' = A^T stored (b, n_j, n_i, C) -> !AK
//   B' = B   stored (b, n_j, n_k, C) -> !BK;  B' = B^T stored (b, n_k, n_j, C) -> BK
// Box of a contraction-major operand: (16 ch, JC, 8 rows): cell(r, jj) = r * JC + jj;
// of a row-major one: (16 ch, 8 rows, JC): cell(r, jj) = jj * 8 + r.
template <bool AK, bool BK>
__global__ void __launch_bounds__(kMsThreads, 2)
mamamm_smem_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const unsigned char* __restrict__ mask, const int* __restrict__ ext, MsParams P,
                   float* __restrict__ out, unsigned int* __restrict__ counters) {
  constexpr int JC = BK ? 5 : 4;
  constexpr int kBoxBytes = kMsRB * JC * kMsCell;
  extern __shared__ unsigned char ms_smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[kMsMaxStages], empty_bar[kMsMaxStages];
  __shared__ MsMeta meta[kMsMaxStages];

  // [zeros: kMsZeroBytes][ring: stages x stage_bytes], 128-byte aligned
  const uint32_t zeros = (ms_u32(ms_smem_raw) + 127u) & ~127u;
  const uint32_t ring = zeros + kMsZeroBytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int o = tid * 16; o < kMsZeroBytes; o += kMsThreads * 16)
    asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(zeros + o), "f"(0.f) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> bulk-copy reads

  if (tid == 0) {
    for (int s = 0; s < P.stages; ++s) {
      ms_mbar_init(ms_u32(&full_bar[s]), 1);
      ms_mbar_init(ms_u32(&empty_bar[s]), kMsConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int upi = P.nslab * P.npass;          // units per graph
  const uint32_t hint = (uint32_t)P.dbg >> 8;  // suspend-time hint of the barrier waits, ns (0: none)

  if (warp == kMsProducerWarp) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const long long row_bytes = (long long)P.n_k * P.dense * 4;
      ms_trace(P, 0, 6ull << 56);
      // kMsQueues work queues (same-address atomics are served one at a time by the L2, ~14 ns
      // each: one queue for all CTAs cost 20 us per launch): queue q holds the units q, q + NQ,
      // q + 2 NQ ... and is shared by the CTAs with blockIdx % NQ == q.  The first unit of a CTA
      // is static; the queue is read two units ahead and the extents one unit ahead, so neither
      // the atomic's nor the loads' round trip sits between two units' boxes
      const unsigned int q = blockIdx.x % kMsQueues;
      const unsigned int ctas_q = (gridDim.x - q + kMsQueues - 1) / kMsQueues;
      unsigned int* qc = counters + 32 * q;
      unsigned int u = q + kMsQueues * (blockIdx.x / kMsQueues);
      unsigned int u1 = q + kMsQueues * (ctas_q + atomicAdd(qc, 1u));
      int e0 = P.n_i, e1 = P.n_j, e2 = P.n_k;
      if (ext && (long long)u < P.units) {
        const int it = (int)(u / (unsigned)upi);
        e0 = ext[3 * it + 0]; e1 = ext[3 * it + 1]; e2 = ext[3 * it + 2];
      }
      while ((long long)u < P.units) {
        const unsigned int u2 = q + kMsQueues * (ctas_q + atomicAdd(qc, 1u));
        int f0 = P.n_i, f1 = P.n_j, f2 = P.n_k;
        if (ext && (long long)u1 < P.units) {
          const int it = (int)(u1 / (unsigned)upi);
          f0 = ext[3 * it + 0]; f1 = ext[3 * it + 1]; f2 = ext[3 * it + 2];
        }
        const int item = (int)(u / (unsigned)upi), rem = (int)(u % (unsigned)upi);
        const int slab = rem / P.npass, pass = rem % P.npass;
        const int ei = min(max(e0, 0), P.n_i), ej = min(max(e1, 0), P.n_j), ek = min(max(e2, 0), P.n_k);
        const int nkb = (ek + kMsRB - 1) / kMsRB;
        const int r0 = pass * ms_rows_per_pass(nkb);
        const int rows = max(min(ei - r0, ms_rows_per_pass(nkb)), 0);
        const int nab = (rows + kMsRB - 1) / kMsRB;
        const int nch = (rows > 0 && ek > 0) ? max((ej + JC - 1) / JC, 1) : 0;
        const int c0 = slab * kMsCH;
        for (int ch = 0; ch < nch; ++ch) {
          ms_mbar_wait(ms_u32(&empty_bar[s]), ph ^ 1u, hint);
          ms_trace(P, 0, (1ull << 56) | ((unsigned long long)item << 16) | (slab << 8) | ch);
          meta[s] = MsMeta{item, slab, ch, nch, ei, ek, r0, rows, nab};
          const uint32_t fb = ms_u32(&full_bar[s]);
          if (P.dbg & 2) {
            ms_mbar_arrive(fb);
          } else {
            ms_mbar_expect_tx(fb, (uint32_t)((nab + nkb) * kBoxBytes));
            const uint32_t sa = ring + (uint32_t)s * (uint32_t)P.stage_bytes;
            const uint32_t sb = sa + (uint32_t)nab * kBoxBytes;
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
        // this unit's share of the output's pad region (rows >= ei, and the columns >= ek of the
        // rows below): whole runs of zeros by the bulk-copy engine, rows dealt over the units
        if (!(P.dbg & 8)) {
          char* obase = reinterpret_cast<char*>(out) + (long long)item * P.n_i * row_bytes;
          for (int i = rem; i < P.n_i; i += upi) {
            const int k0 = i < ei ? ek : 0;
            char* dst = obase + i * row_bytes + (long long)k0 * P.dense * 4;
            long long left = (long long)(P.n_k - k0) * P.dense * 4;
            while (left > 0) {
              const uint32_t nb = (uint32_t)(left < kMsZeroBytes ? left : kMsZeroBytes);
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(zeros),
                           "r"(nb)
                           : "memory");
              dst += nb;
              left -= nb;
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ms_trace(P, 0, (7ull << 56) | ((unsigned long long)item << 16) | (slab << 8));
        u = u1; u1 = u2;
        e0 = f0; e1 = f1; e2 = f2;
        ms_trace(P, 0, (8ull << 56));
      }
      // end marker for the consumers
      ms_mbar_wait(ms_u32(&empty_bar[s]), ph ^ 1u, hint);
      meta[s] = MsMeta{-1, 0, 0, 0, 0, 0, 0, 0, 0};
      ms_mbar_arrive(ms_u32(&full_bar[s]));
      ms_trace(P, 0, 9ull << 56);
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      ms_trace(P, 0, 10ull << 56);
      // the last CTA to run dry re-arms the queue for the next launch
      __threadfence();
      if (atomicAdd(&counters[32 * kMsQueues], 1u) == gridDim.x - 1) {
        for (int i = 0; i <= kMsQueues; ++i) counters[32 * i] = 0;
        __threadfence();
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  MsRing R;
  R.ring = ring; R.full = ms_u32(&full_bar[0]); R.empty = ms_u32(&empty_bar[0]);
  R.s = 0; R.ph = 0; R.hint = hint;
  for (;;) {
    ms_warp_wait(R.full + 8u * R.s, R.ph, hint);
    const MsMeta m = meta[R.s];
    if (m.item < 0) break;
    // units that need at most half of the tile slots are spread over twice the threads (two
    // channels per lane instead of four): twice the warps to hide the shared-memory latency
    const int need = ((m.rows + 3) >> 2) * ((m.ek + kMsRB - 1) / kMsRB);
    if (need * 16 <= kMsConsumers && !(P.dbg & 16))
      ms_unit<AK, BK, 8>(m, R, P, mask, out);
    else
      ms_unit<AK, BK, 4>(m, R, P, mask, out);
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

extern unsigned long long* g_trace_host;   // mamamm_tc.cu (pgh_debug_trace)
extern long long g_trace_host_words;

// self-resetting work-queue sets ([kMsQueues counters + 1 exit count], one 128-byte line each),
// rotated per launch so that launches on different streams (or concurrent branches of one
// captured graph) never share one
constexpr int kMsQueueSets = 32;
constexpr int kMsQueueSetWords = 32 * (kMsQueues + 1);
__device__ unsigned int g_ms_counters[kMsQueueSets * kMsQueueSetWords];

static bool ms_plan(int64_t b, int64_t n_i, int64_t n_j, int64_t n_k, int64_t dense, bool bk, MsParams& P) {
  if (dense % kMsCH != 0 || n_i <= 0 || n_j <= 0 || n_k <= 0) return false;
  if (n_i > 32768 || n_j > 32768 || n_k > 32768 || dense > 32768) return false;
  const int JC = bk ? 5 : 4;
  const int nkb_max = (int)((n_k + kMsRB - 1) / kMsRB);
  if (nkb_max > kMsMaxPairs) return false;
  const int ni4 = (int)((n_i + 3) / 4) * 4;
  P.n_i = (int)n_i; P.n_j = (int)n_j; P.n_k = (int)n_k; P.dense = (int)dense;
  P.nslab = (int)(dense / kMsCH);
  // a graph with nkb column blocks is cut into passes of ms_rows_per_pass(nkb) rows: the widest
  // graphs need the most passes, and a stage must hold the boxes of any (rows, nkb) combination
  P.npass = (ni4 + ms_rows_per_pass(nkb_max) - 1) / ms_rows_per_pass(nkb_max);
  int boxes = 0;
  for (int nkb = 1; nkb <= nkb_max; ++nkb) {
    const int rows = ms_rows_per_pass(nkb) < ni4 ? ms_rows_per_pass(nkb) : ni4;
    const int nb = (rows + kMsRB - 1) / kMsRB + nkb;
    if (nb > boxes) boxes = nb;
  }
  P.stage_bytes = boxes * kMsRB * JC * kMsCell;
  const long long room = (long long)kMsSmemCap - 2048 - 128 - kMsZeroBytes;   // static + reserved, alignment
  long long st = room / P.stage_bytes;
  if (st < 3) return false;
  P.stages = (int)(st > kMsMaxStages ? kMsMaxStages : st);
  const long long units = (long long)b * P.nslab * P.npass;
  if (units <= 0 || units > 0x7fffffffLL - (1 << 20) || (long long)n_i * n_k > 0x7fffffffLL) return false;
  P.units = units;
  P.dbg = g_tune[7];
  P.trace = g_trace_host;
  P.trace_words = g_trace_host_words;
  return true;
}

template <bool AK, bool BK>
static int ms_launch_t(const CUtensorMap& ma, const CUtensorMap& mb, const unsigned char* mask,
                       const int* ext, const MsParams& P, float* out, unsigned int* counters,
                       cudaStream_t s) {
  const size_t smem = (size_t)P.stages * P.stage_bytes + 128 + kMsZeroBytes;
  static bool attr_set = false;
  if (!attr_set) {
    PGH_CUDA(cudaFuncSetAttribute(mamamm_smem_kernel<AK, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(kMsSmemCap - 2048)));
    PGH_CUDA(cudaFuncSetAttribute(mamamm_smem_kernel<AK, BK>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  const unsigned grid = (unsigned)(P.units < 2 * kSMs ? P.units : 2 * kSMs);
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
  const int JC = bk ? 5 : 4;
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
  unsigned int* q = counters + kMsQueueSetWords * (next_queue++ % kMsQueueSets);
  if (ak && bk) return ms_launch_t<true, true>(ma, mb, mask, ext, P, out, q, s);
  if (ak && !bk) return ms_launch_t<true, false>(ma, mb, mask, ext, P, out, q, s);
  if (!ak && bk) return ms_launch_t<false, true>(ma, mb, mask, ext, P, out, q, s);
  return ms_launch_t<false, false>(ma, mb, mask, ext, P, out, q, s);
}

}  // namespace pgh
