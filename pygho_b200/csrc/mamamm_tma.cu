// 2-FWL contraction with TMA-staged operands (algo 3): no thread ever touches an operand.
//
//   out[b,i,k,c] = mask[b,i,k] * sum_j A'[b,i,j,c] * B'[b,j,k,c]      (backend/Mamamm.py:7-64)
//
// The channel axis c is contiguous in memory and is a batch axis of the contraction.  algo 1 / 2
// (mamamm_tc.cu) copy 8-channel slabs to shared memory and let transposer warps re-lay them into
// per-channel K-major tiles (their serial transposer -> MMA chain is the 40 us floor measured in
// round 1).  Here the raw TMA box IS the UMMA operand:
//
//   * a work item is (graph b, FOUR channels c0..c0+3, 32-row tile of i).  A TMA box
//     (4 channels x 8 j x R rows) lands in shared memory as [row][8 j][4 c] = exactly the canonical
//     MN-major / no-swizzle UMMA layout with the MN index m = 4 * row + c (16-byte groups of
//     4 consecutive MN elements, 8 K-positions 16 bytes apart, SBO = 128 bytes between rows): the
//     tensor core contracts over j with M = (32 i) x (4 c) and N = (k) x (4 c').  Operand B uses
//     the tensor-map dimension order (c, j, k), so its box lands as [k][8 j][4 c] although memory
//     is [j][k][c]; a transposed operand (the backward products) is just another stride order in
//     the tensor map -- no kernel code.
//   * the accumulator holds all 4 x 4 channel pairs; the epilogue keeps the diagonal c == c'
//     (lane = 4 * i + c reads column 4 * k + c).  4x redundant tensor-core work is free here: the
//     contraction has 13 flop/byte against a ridge of ~170.
//   * roles: warp 0 = TMA producer (one thread, 2 * ceil(n_j / 8) bulk-tensor loads per item into
//     a 4-stage ring), warp 1 = MMA issuer (one thread, one tcgen05.mma.kind::tf32 per 8 j),
//     warps 4-7 = epilogue (tcgen05.ld, diagonal select, mask, stores).  Double-buffered TMEM
//     accumulators (2 x 256 columns).
//   * per-graph valid extents: box heights come from a menu of tensor maps (8, 16, ... rows), so
//     only ceil8 of the valid rows / columns of every graph is fetched; `out` is zero-filled by a
//     memset node and only the valid rectangle is written.
//
// Limits: dense % 4 == 0, n_j <= 64, n_k <= 64 (box menu), any n_i (32-row tiles); 2-4 ring stages
// depending on the tile size.
#include <cuda.h>

#include "common.cuh"

namespace pgh {

constexpr int kMtThreads = 256;
constexpr int kMtStages = 4;
constexpr int kMtMaxRowsA = 32;            // i rows per M = 128 tile (4 channels each)
constexpr int kMtMaxK = 64;                // n_j, n_k limit
constexpr int kMtMapsA = kMtMaxRowsA / 8;  // box heights 8, 16, 24, 32
constexpr int kMtMapsB = kMtMaxK / 8;      // box heights 8 .. 64
constexpr int kMtAccCols = 256;

__device__ __forceinline__ uint32_t mt_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// MN-major SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14), leading byte offset>>4 [16,30) = stride between 8-wide K blocks,
// stride byte offset>>4 [32,46) = stride between 16-byte MN groups, version = 1 [46,48)
__device__ __forceinline__ uint64_t mt_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// kind::tf32, D = F32, A and B MN-major (bits 15, 16), N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t mt_idesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mt_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mt_mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void mt_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mt_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mt_tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                               int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void mt_umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mt_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void mt_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct MtMaps {
  CUtensorMap a[kMtMapsA];       // box (4 c, 8 j, 8 * (idx + 1) rows of i, 1)
  CUtensorMap b[kMtMapsB];       // box (4 c, 8 j, 8 * (idx + 1) rows of k, 1)
};

struct MtItem {
  int b, c0, i0, ni, nj, nk;     // ni = valid rows of this 32-row tile
  bool empty;
};

__device__ __forceinline__ MtItem mt_item(int it, int groups, int mtiles, const int* __restrict__ ext,
                                          int n_i, int n_j, int n_k) {
  MtItem I;
  const int per_graph = groups * mtiles;
  I.b = it / per_graph;
  const int r = it - I.b * per_graph;
  const int mt = r / groups;
  I.c0 = (r - mt * groups) * 4;
  I.i0 = mt * kMtMaxRowsA;
  int vi = n_i, vj = n_j, vk = n_k;
  if (ext) {
    vi = min(max(__ldg(ext + 3 * I.b), 0), n_i);
    vj = min(max(__ldg(ext + 3 * I.b + 1), 0), n_j);
    vk = min(max(__ldg(ext + 3 * I.b + 2), 0), n_k);
  }
  I.ni = min(max(vi - I.i0, 0), kMtMaxRowsA);
  I.nj = vj;
  I.nk = vk;
  I.empty = I.ni == 0 || vj == 0 || vk == 0;
  return I;
}

struct MtBars {
  unsigned long long full[kMtStages], empty[kMtStages], tfull[2], tempty[2];
};

__global__ void __launch_bounds__(kMtThreads, 1)
mamamm_tma_kernel(const __grid_constant__ MtMaps maps, const unsigned char* __restrict__ mask,
                  const int* __restrict__ ext, int n_items, int groups, int mtiles, int n_i, int n_j,
                  int n_k, int dense, int stage_bytes, int off_b, int n_stages, float* __restrict__ out) {
  extern __shared__ unsigned char mt_smem_raw[];
  __shared__ MtBars bars;
  __shared__ uint32_t tmem_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = (mt_smem_u32(mt_smem_raw) + 127u) & ~127u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMtStages; ++s) {
      mt_mbar_init(mt_smem_u32(&bars.full[s]), 1);
      mt_mbar_init(mt_smem_u32(&bars.empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mt_mbar_init(mt_smem_u32(&bars.tfull[a]), 1);
      mt_mbar_init(mt_smem_u32(&bars.tempty[a]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     mt_smem_u32(&tmem_holder)),
                 "r"(2 * kMtAccCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const MtItem I = mt_item(it, groups, mtiles, ext, n_i, n_j, n_k);
        if (I.empty) continue;
        const int kb = (I.nj + 7) >> 3;
        const int ra = (I.ni + 7) & ~7, rb = (I.nk + 7) & ~7;          // box heights
        mt_mbar_wait(mt_smem_u32(&bars.empty[stage]), phase ^ 1u);
        const uint32_t full = mt_smem_u32(&bars.full[stage]);
        const uint32_t sa = ring + (uint32_t)stage * stage_bytes, sb = sa + (uint32_t)off_b;
        mt_mbar_expect_tx(full, (uint32_t)(kb * (ra + rb) * 128));
        const CUtensorMap* ma = &maps.a[(ra >> 3) - 1];
        const CUtensorMap* mb = &maps.b[(rb >> 3) - 1];
        for (int jb = 0; jb < kb; ++jb) {
          mt_tma_load_4d(sa + (uint32_t)(jb * ra * 128), ma, full, I.c0, jb * 8, I.i0, I.b);
          mt_tma_load_4d(sb + (uint32_t)(jb * rb * 128), mb, full, I.c0, jb * 8, 0, I.b);
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int q = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const MtItem I = mt_item(it, groups, mtiles, ext, n_i, n_j, n_k);
        if (I.empty) continue;
        const int kb = (I.nj + 7) >> 3;
        const int ra = (I.ni + 7) & ~7, rb = (I.nk + 7) & ~7;
        const int buf = q & 1;
        const uint32_t aphase = (uint32_t)(q >> 1) & 1u;
        mt_mbar_wait(mt_smem_u32(&bars.tempty[buf]), aphase ^ 1u);
        mt_mbar_wait(mt_smem_u32(&bars.full[stage]), phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = ring + (uint32_t)stage * stage_bytes, sb = sa + (uint32_t)off_b;
        const uint32_t idesc = mt_idesc(4 * rb);
        const uint32_t acc = tmem_base + (uint32_t)(buf * kMtAccCols);
        for (int jb = 0; jb < kb; ++jb)
          mt_umma(acc, mt_desc(sa + (uint32_t)(jb * ra * 128), (uint32_t)(ra * 128), 128u),
                  mt_desc(sb + (uint32_t)(jb * rb * 128), (uint32_t)(rb * 128), 128u), idesc,
                  (uint32_t)(jb != 0));
        mt_commit(mt_smem_u32(&bars.empty[stage]));
        mt_commit(mt_smem_u32(&bars.tfull[buf]));
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        ++q;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: lane = 4 * i_local + c ; column = 4 * k + c' ; keep c' == c =====
    const int qd = warp - 4;
    const int i_loc = qd * 8 + (lane >> 2), c = lane & 3;
    int q = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const MtItem I = mt_item(it, groups, mtiles, ext, n_i, n_j, n_k);
      if (I.empty) continue;
      const int buf = q & 1;
      const uint32_t aphase = (uint32_t)(q >> 1) & 1u;
      mt_mbar_wait(mt_smem_u32(&bars.tfull[buf]), aphase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (qd * 8 < I.ni) {                               // this warp's 8 rows hold valid rows
        const bool row_ok = i_loc < I.ni;
        const size_t pos0 = ((size_t)I.b * n_i + (size_t)(I.i0 + (row_ok ? i_loc : 0))) * n_k;
        const unsigned char* mrow = mask + pos0;
        float* orow = out + pos0 * dense + I.c0 + c;
        const uint32_t tacc = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * kMtAccCols);
        for (int k0 = 0; k0 < I.nk; k0 += 8) {           // 32 accumulator columns = 8 k x 4 c'
          uint32_t r[32];
          mt_tmem_ld32(tacc + (uint32_t)(4 * k0), r);
          if (row_ok) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const int k = k0 + kk;
              if (k < I.nk) {
                const uint32_t v = c == 0 ? r[4 * kk] : c == 1 ? r[4 * kk + 1] : c == 2 ? r[4 * kk + 2]
                                                                                        : r[4 * kk + 3];
                orow[(size_t)k * dense] = mrow[k] ? __uint_as_float(v) : 0.f;
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mt_mbar_arrive(mt_smem_u32(&bars.tempty[buf]));
      ++q;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(2 * kMtAccCols)
                 : "memory");
}

typedef CUresult (*PFN_mtEncode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);

static PFN_mtEncode mt_encode_fn() {
  static PFN_mtEncode fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_mtEncode>(p);
  }
  return fn;
}

// 4-D view (c, j, row, b) of an operand: `s_j` / `s_row` are the strides of the contraction index
// and of the free index in POSITIONS (one position = `dense` floats); rows beyond `n_rows` and
// columns beyond `n_j` are zero-filled by the TMA unit
static bool mt_make_map(CUtensorMap* map, const float* base, int64_t dense, int64_t n_j, int64_t s_j,
                        int64_t n_rows, int64_t s_row, int64_t batch, int64_t graph_positions,
                        int box_rows) {
  PFN_mtEncode enc = mt_encode_fn();
  if (!enc) return false;
  const cuuint64_t ps = (cuuint64_t)dense * sizeof(float);
  const cuuint64_t dims[4] = {(cuuint64_t)dense, (cuuint64_t)n_j, (cuuint64_t)n_rows, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)s_j * ps, (cuuint64_t)s_row * ps, (cuuint64_t)graph_positions * ps};
  const cuuint32_t box[4] = {4, 8, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int mamamm_tma_supported(int64_t n_i, int64_t n_j, int64_t n_k, int64_t dense) {
  return dense % 4 == 0 && n_j >= 1 && n_j <= kMtMaxK && n_k >= 1 && n_k <= kMtMaxK && n_i >= 1;
}

int mamamm_tma_launch(const float* A, int trans_a, const float* B, int trans_b,
                      const unsigned char* mask, const int* ext, int64_t b, int64_t n_i, int64_t n_j,
                      int64_t n_k, int64_t dense, float* out, cudaStream_t s) {
  if (!mamamm_tma_supported(n_i, n_j, n_k, dense)) {
    set_error("mamamm algo=3 supports dense %% 4 == 0, n_j <= 64, n_k <= 64");
    return -2;
  }
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) {
    set_error("mamamm algo=3 needs 16-byte aligned operands");
    return -2;
  }
  // A' (n_i x n_j): stored (n_i, n_j) or, transposed, (n_j, n_i); B' (n_j x n_k) likewise
  const int64_t sa_i = trans_a ? 1 : n_j, sa_j = trans_a ? n_i : 1;
  const int64_t sb_j = trans_b ? 1 : n_k, sb_k = trans_b ? n_j : 1;
  MtMaps maps;
  for (int m = 0; m < kMtMapsA; ++m)
    if (!mt_make_map(&maps.a[m], A, dense, n_j, sa_j, n_i, sa_i, b, n_i * n_j, 8 * (m + 1))) {
      set_error("mamamm algo=3: cuTensorMapEncodeTiled failed (A)");
      return -3;
    }
  for (int m = 0; m < kMtMapsB; ++m)
    if (!mt_make_map(&maps.b[m], B, dense, n_j, sb_j, n_k, sb_k, b, n_j * n_k, 8 * (m + 1))) {
      set_error("mamamm algo=3: cuTensorMapEncodeTiled failed (B)");
      return -3;
    }
  const int kb_max = (int)((n_j + 7) / 8);
  const int rb_max = (int)((n_k + 7) / 8 * 8);
  // operand A of a stage: kb blocks of <= 32 rows, read as M = 128 (32 rows) by the MMA whatever
  // the box height: reserve full blocks; operand B: kb blocks of rb_max rows
  const int off_b = kb_max * kMtMaxRowsA * 128;
  const int stage_bytes = off_b + kb_max * rb_max * 128;
  int n_stages = (int)((216 * 1024 - 128 - 4096) / stage_bytes);
  if (n_stages > kMtStages) n_stages = kMtStages;
  if (n_stages < 2) {
    set_error("mamamm algo=3: two tile stages do not fit in shared memory");
    return -2;
  }
  const size_t smem = (size_t)n_stages * stage_bytes + 128 + 4096;   // + alignment + M-tile over-read
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(mamamm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("mamamm algo=3: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_smem = smem;
  }
  const int groups = (int)(dense / 4);
  const int mtiles = (int)((n_i + kMtMaxRowsA - 1) / kMtMaxRowsA);
  const long long n_items = (long long)b * groups * mtiles;
  if (n_items > 0x7fffffffLL) {
    set_error("mamamm algo=3: too many work items");
    return -2;
  }
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)b * n_i * n_k * dense * sizeof(float), s);
  if (e != cudaSuccess) {
    set_error("mamamm algo=3: %s", cudaGetErrorString(e));
    return (int)e;
  }
  const int grid = (int)(n_items < kSMs ? n_items : kSMs);
  mamamm_tma_kernel<<<grid, kMtThreads, smem, s>>>(maps, mask, ext, (int)n_items, groups, mtiles, (int)n_i,
                                                   (int)n_j, (int)n_k, (int)dense, stage_bytes, off_b, n_stages, out);
  return check_launch("mamamm_tma");
}

}  // namespace pgh
