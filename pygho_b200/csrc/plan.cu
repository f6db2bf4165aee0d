// Device-side index-plan builders (integer work, HBM/latency bound, cub for sort/scan).
//
// The reference builds its plans with torch.argsort / searchsorted / cumsum /
// repeat_interleave / unique (backend/Spspmm.py:57-222, SpTensor.py:167-197), each a
// separate ATen launch with host synchronisation in between.  Here every stage is a
// stream-ordered kernel; data-dependent sizes land in device counters that the host
// reads once.  All orders are canonical (stable sorts), so plans are deterministic.
#include <cub/cub.cuh>

#include "common.cuh"

namespace pgh {

constexpr int kT = 256;

struct RowSel {
  int n;
  int rows[8];
  long long dims[8];
};

__global__ void pack_keys_kernel(const long long* __restrict__ ind, long long ld, RowSel sel,
                                 int bits, long long nnz, long long* __restrict__ key,
                                 int* __restrict__ info) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  long long k = 0;
  int neg = 0, big = 0;
  for (int r = 0; r < sel.n; ++r) {
    const long long v = ind[(long long)sel.rows[r] * ld + i];
    neg |= v < 0;
    big |= (sel.n > 1) && (v >> bits) != 0;
    k = (sel.n > 1) ? ((k << bits) | v) : v;
  }
  key[i] = k;
  if (info) {
    if (neg) atomicAdd(info + 0, 1);
    if (big && !neg) atomicAdd(info + 1, 1);
  }
}

__global__ void pack_tight_kernel(const long long* __restrict__ ind, long long ld, RowSel sel,
                                  long long nnz, long long* __restrict__ key,
                                  int* __restrict__ info) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  long long k = 0;
  int bad = 0;
  for (int r = 0; r < sel.n; ++r) {
    const long long v = ind[(long long)sel.rows[r] * ld + i];
    bad |= (v < 0) || (v >= sel.dims[r]);
    k = k * sel.dims[r] + v;
  }
  key[i] = k;
  if (info && bad) atomicAdd(info + 0, 1);
}

__global__ void unpack_keys_kernel(const long long* __restrict__ key, long long n, int sd,
                                   int bits, long long* __restrict__ out, long long ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long k = key[i];
  if (sd == 1) {
    out[i] = k;
    return;
  }
  const long long field = (1ll << bits) - 1;
  for (int r = 0; r < sd; ++r) out[(long long)r * ld + i] = (k >> (bits * (sd - 1 - r))) & field;
}

__global__ void unpack_tight_kernel(const long long* __restrict__ key, long long n, RowSel sel,
                                    long long* __restrict__ out, long long ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long k = key[i];
  for (int r = sel.n - 1; r >= 0; --r) {
    const long long s = sel.dims[r];
    out[(long long)r * ld + i] = (r == 0) ? k : k % s;
    k /= s;
  }
}

__global__ void iota_kernel(int* __restrict__ p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int)i;
}

__global__ void head_flags_kernel(const long long* __restrict__ key, long long n,
                                  int* __restrict__ flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// seg (inclusive scan of head flags, in place) -> seg-1; heads write their key; last writes count
__global__ void unique_finish_kernel(const long long* __restrict__ key, long long n,
                                     int* __restrict__ seg, long long* __restrict__ ukey,
                                     int* __restrict__ count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = seg[i] - 1;
  seg[i] = s;
  if (i == 0 || key[i] != key[i - 1]) ukey[s] = key[i];
  if (i == n - 1) *count = s + 1;
}

// rowptr[r] = first position whose key >= r ; one thread per position boundary
__global__ void rowptr_kernel(const int* __restrict__ key, long long n, long long n_rows,
                              int* __restrict__ rowptr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const long long prev = (i == 0) ? -1 : (long long)key[i - 1];
  const long long cur = (i == n) ? n_rows : (long long)key[i];
  for (long long r = prev + 1; r <= cur && r <= n_rows; ++r) rowptr[r] = (int)i;
}

__device__ __forceinline__ long long lower_bound_ll(const long long* __restrict__ a, long long n,
                                                    long long q) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (a[mid] < q) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ long long upper_bound_ll(const long long* __restrict__ a, long long n,
                                                    long long q) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (a[mid] <= q) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void match_kernel(const long long* __restrict__ sorted, long long n,
                             const long long* __restrict__ q, long long m,
                             int* __restrict__ lo, long long* __restrict__ cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const long long l = lower_bound_ll(sorted, n, q[i]);
  const long long u = upper_bound_ll(sorted, n, q[i]);
  lo[i] = (int)l;
  cnt[i] = u - l;
}

__global__ void set_last_kernel(const long long* __restrict__ excl, const long long* __restrict__ cnt,
                                long long m, long long* __restrict__ off) {
  // off[0..m) already holds the exclusive scan; write the total
  if (threadIdx.x == 0 && blockIdx.x == 0) off[m] = (m > 0) ? excl[m - 1] + cnt[m - 1] : 0;
}

__global__ void expand_kernel(const long long* __restrict__ off, const int* __restrict__ lo,
                              const int* __restrict__ perm2, long long m, long long total,
                              int* __restrict__ c, int* __restrict__ d) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long q = upper_bound_ll(off, m + 1, t) - 1;
  const long long j = (long long)lo[q] + (t - off[q]);
  c[t] = (int)q;
  d[t] = perm2 ? perm2[j] : (int)j;
}

struct PairSel {
  int n1, n2;
  int rows1[8], rows2[8];
};

__global__ void pair_keys_kernel(const long long* __restrict__ ind1, long long ld1,
                                 const long long* __restrict__ ind2, long long ld2, PairSel sel,
                                 const int* __restrict__ c, const int* __restrict__ d,
                                 long long total, int bits, long long* __restrict__ key,
                                 int* __restrict__ info) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long p = c[t], q = d[t];
  long long k = 0;
  int big = 0;
  const bool single = (sel.n1 + sel.n2) == 1;
  for (int r = 0; r < sel.n1; ++r) {
    const long long v = ind1[(long long)sel.rows1[r] * ld1 + p];
    big |= !single && (v >> bits) != 0;
    k = single ? v : ((k << bits) | v);
  }
  for (int r = 0; r < sel.n2; ++r) {
    const long long v = ind2[(long long)sel.rows2[r] * ld2 + q];
    big |= !single && (v >> bits) != 0;
    k = single ? v : ((k << bits) | v);
  }
  key[t] = k;
  if (info && big) atomicAdd(info + 1, 1);
}

__global__ void lookup_kernel(const long long* __restrict__ tkeys, long long nt,
                              const long long* __restrict__ keys, long long n,
                              int* __restrict__ pos) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long k = keys[i];
  const long long l = lower_bound_ll(tkeys, nt, k);
  pos[i] = (l < nt && tkeys[l] == k) ? (int)l : -1;
}

__global__ void keep_flags_kernel(const int* __restrict__ a, const int* __restrict__ map,
                                  long long n, int* __restrict__ a_mapped,
                                  int* __restrict__ flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int v = a[i];
  if (map && v >= 0) v = map[v];
  a_mapped[i] = v;
  flag[i] = v >= 0 ? 1 : 0;
}

__global__ void compact_scatter_kernel(const int* __restrict__ a_mapped,
                                       const int* __restrict__ c, const int* __restrict__ d,
                                       const int* __restrict__ excl, long long n,
                                       int* __restrict__ oa, int* __restrict__ oc,
                                       int* __restrict__ od, int* __restrict__ count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = a_mapped[i];
  const int o = excl[i];
  if (v >= 0) {
    oa[o] = v;
    oc[o] = c[i];
    od[o] = d[i];
  }
  if (i == n - 1) *count = o + (v >= 0 ? 1 : 0);
}

__global__ void i64_to_i32_kernel(const long long* __restrict__ s, long long n,
                                  int* __restrict__ dst, int* __restrict__ info) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long v = s[i];
  dst[i] = (int)v;
  if (info && (v < 0 || v > 0x7fffffffll)) atomicAdd(info + 0, 1);
}

__global__ void i32_to_i64_kernel(const int* __restrict__ s, long long n,
                                  long long* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = s[i];
}

__global__ void gather_i32_kernel(const int* __restrict__ s, const int* __restrict__ idx,
                                  long long n, int* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = s[idx[i]];
}

__global__ void gather_i64_i32_kernel(const long long* __restrict__ s,
                                      const int* __restrict__ idx, long long n,
                                      int* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (int)s[idx[i]];
}

__global__ void check_sorted_kernel(const long long* __restrict__ key, long long n, int strict,
                                    int* __restrict__ info) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 < n && (key[i] > key[i + 1] || (strict && key[i] == key[i + 1])))
    atomicAdd(info + 0, 1);
}

static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace pgh

using namespace pgh;

static int fill_rowsel(RowSel& s, const int32_t* rows, const int64_t* dims, int n) {
  if (n < 1 || n > 8) return arg_error("1..8 index rows supported");
  s.n = n;
  for (int i = 0; i < n; ++i) {
    s.rows[i] = rows[i];
    s.dims[i] = dims ? dims[i] : 0;
  }
  return 0;
}

extern "C" int pgh_pack_keys(const int64_t* ind, int64_t ld, const int32_t* rows_host,
                             int n_rows_sel, int bits, int64_t nnz, int64_t* key, int32_t* info,
                             void* stream) {
  RowSel s;
  if (int e = fill_rowsel(s, rows_host, nullptr, n_rows_sel)) return e;
  if (n_rows_sel > 1 && (bits < 1 || bits * n_rows_sel > 63)) return arg_error("pack_keys: bits");
  if (nnz <= 0) return 0;
  pack_keys_kernel<<<blocks_for(nnz, kT), kT, 0, as_stream(stream)>>>(
      (const long long*)ind, ld, s, bits, nnz, (long long*)key, info);
  return check_launch("pack_keys");
}

extern "C" int pgh_pack_tight(const int64_t* ind, int64_t ld, const int32_t* rows_host,
                              const int64_t* dims_host, int n_rows_sel, int64_t nnz, int64_t* key,
                              int32_t* info, void* stream) {
  RowSel s;
  if (int e = fill_rowsel(s, rows_host, dims_host, n_rows_sel)) return e;
  if (nnz <= 0) return 0;
  pack_tight_kernel<<<blocks_for(nnz, kT), kT, 0, as_stream(stream)>>>(
      (const long long*)ind, ld, s, nnz, (long long*)key, info);
  return check_launch("pack_tight");
}

extern "C" int pgh_unpack_keys(const int64_t* key, int64_t n, int sd, int bits, int64_t* out,
                               int64_t ld, void* stream) {
  if (sd < 1 || sd > 8) return arg_error("unpack_keys: sd");
  if (n <= 0) return 0;
  unpack_keys_kernel<<<blocks_for(n, kT), kT, 0, as_stream(stream)>>>(
      (const long long*)key, n, sd, bits, (long long*)out, ld);
  return check_launch("unpack_keys");
}

extern "C" size_t pgh_sort_ws_bytes(int64_t n) {
  if (n <= 0) return 256;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned long long*)nullptr,
                                  (unsigned long long*)nullptr, (const int*)nullptr,
                                  (int*)nullptr, (int)n, 0, 64);
  return align_up(tmp) + align_up(sizeof(int) * (size_t)n);
}

extern "C" int pgh_sort_keys_perm(const int64_t* key_in, int64_t n, int end_bit, int64_t* key_out,
                                  int32_t* perm_out, void* ws, size_t ws_bytes, void* stream) {
  if (n <= 0) return 0;
  if (n > 0x7fffffff) return arg_error("sort: n too large");
  if (end_bit < 1) end_bit = 1;
  if (end_bit > 64) end_bit = 64;
  cudaStream_t s = as_stream(stream);
  const size_t iota_bytes = align_up(sizeof(int) * (size_t)n);
  if (ws_bytes < iota_bytes + 256) return arg_error("sort: workspace too small");
  int* iota = reinterpret_cast<int*>(ws);
  void* tmp = reinterpret_cast<char*>(ws) + iota_bytes;
  size_t tmp_bytes = ws_bytes - iota_bytes;
  iota_kernel<<<blocks_for(n, kT), kT, 0, s>>>(iota, n);
  PGH_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const unsigned long long*)key_in,
                                           (unsigned long long*)key_out, (const int*)iota,
                                           (int*)perm_out, (int)n, 0, end_bit, s));
  return check_launch("sort_keys_perm");
}

// ------------------------------------------------------------------ one-call CSR regrouping
// Everything plans.TriplePlan needs from a reference-format plan (3, T) int64 in ONE host call:
// int32 copies of a / c / d and, for each requested grouping (bit 1 = by a, 2 = by c, 4 = by d),
// rowptr + the two other index arrays in that grouping's (stable) order.  ~14 launches issued from
// C instead of ~30 Python-level operations per plan: the host side of feeding a batch is
// launch-overhead bound (profiles/r1_host_profile.txt).
__global__ void gather2_i32_kernel(const int* __restrict__ s1, const int* __restrict__ s2,
                                   const int* __restrict__ idx, long long n, int* __restrict__ d1,
                                   int* __restrict__ d2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = idx[i];
  d1[i] = s1[p];
  d2[i] = s2[p];
}

static int bit_length(int64_t v) {
  int b = 0;
  while (v > 0) { ++b; v >>= 1; }
  return b < 1 ? 1 : b;
}

// All three groupings with ONE radix sort: key = (grouping << bits) | index value over the 3 T
// entries of the plan, value = entry id; each grouping's T entries come out contiguous and stably
// ordered.  A radix sort of 55 k elements is all fixed cost (~10 us per pass): 3 passes instead of 9.
__global__ void acd_compose_kernel(const long long* __restrict__ acd, long long T, int bits,
                                   int* __restrict__ idx32, unsigned int* __restrict__ keys,
                                   int* __restrict__ vals) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * T) return;
  const int g = (int)(i / T);
  const int v = (int)acd[i];
  idx32[i] = v;
  keys[i] = ((unsigned int)g << bits) | (unsigned int)v;
  vals[i] = (int)(i - (long long)g * T);
}

struct Regroup3 {
  int n_rows[3];
  int* rowptr[3];
  const int* f[3];
  const int* g[3];
  int* fo[3];
  int* go[3];
};

__global__ void rowptr3_kernel(const unsigned int* __restrict__ ks, long long T, unsigned int mask,
                               Regroup3 R) {
  const int grp = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > T) return;
  const unsigned int* key = ks + grp * T;
  const long long n_rows = R.n_rows[grp];
  const long long prev = (i == 0) ? -1 : (long long)(key[i - 1] & mask);
  const long long cur = (i == T) ? n_rows : (long long)(key[i] & mask);
  for (long long r = prev + 1; r <= cur && r <= n_rows; ++r) R.rowptr[grp][r] = (int)i;
}

__global__ void gather3_kernel(const int* __restrict__ perm, long long T, Regroup3 R) {
  const int grp = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const int p = perm[grp * T + i];
  R.fo[grp][i] = R.f[grp][p];
  R.go[grp][i] = R.g[grp][p];
}

extern "C" size_t pgh_acd_regroup_ws_bytes(int64_t T) {
  if (T <= 0) return 256;
  size_t tmp = 0, tmp3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned int*)nullptr, (unsigned int*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)T, 0, 32);
  if (3 * T <= 0x7fffffff)
    cub::DeviceRadixSort::SortPairs(nullptr, tmp3, (const unsigned int*)nullptr, (unsigned int*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, (int)(3 * T), 0, 32);
  const size_t one = align_up(tmp) + 3 * align_up(sizeof(int) * (size_t)T);
  const size_t all = align_up(tmp3) + 4 * align_up(sizeof(int) * (size_t)(3 * T));
  return one > all ? one : all;
}

extern "C" int pgh_acd_regroup(const int64_t* acd, int64_t T, int64_t n_out, int64_t n_a, int64_t n_b,
                               int which, int32_t* idx32, int32_t* rowptr_a, int32_t* first_a,
                               int32_t* second_a, int32_t* rowptr_c, int32_t* first_c,
                               int32_t* second_c, int32_t* rowptr_d, int32_t* first_d,
                               int32_t* second_d, void* ws, size_t ws_bytes, void* stream) {
  if (T < 0 || n_out < 0 || n_a < 0 || n_b < 0) return arg_error("acd_regroup: sizes");
  if (T > 0x7fffffff) return arg_error("acd_regroup: T too large");
  if (T > 0 && (!acd || !idx32)) return arg_error("acd_regroup: null pointer");
  cudaStream_t s = as_stream(stream);
  int32_t* a32 = idx32;
  int32_t* c32 = idx32 + T;
  int32_t* d32 = idx32 + 2 * T;
  if (ws_bytes < pgh_acd_regroup_ws_bytes(T)) return arg_error("acd_regroup: workspace too small");
  int64_t n_max = n_out > n_a ? n_out : n_a;
  if (n_b > n_max) n_max = n_b;
  const int bits = bit_length(n_max);
  if (which == 7 && T > 0 && 3 * T <= 0x7fffffff && bits <= 30 && n_max < 0x7fffffff) {
    // one composite-key sort for the three groupings
    if (!rowptr_a || !rowptr_c || !rowptr_d || !first_a || !second_a || !first_c || !second_c ||
        !first_d || !second_d)
      return arg_error("acd_regroup: null output");
    const size_t arr3 = align_up(sizeof(int) * (size_t)(3 * T));
    char* p = static_cast<char*>(ws);
    unsigned int* keys = reinterpret_cast<unsigned int*>(p);
    unsigned int* ks3 = reinterpret_cast<unsigned int*>(p + arr3);
    int* vals = reinterpret_cast<int*>(p + 2 * arr3);
    int* perm3 = reinterpret_cast<int*>(p + 3 * arr3);
    void* tmp3 = p + 4 * arr3;
    size_t tb = ws_bytes - 4 * arr3;
    acd_compose_kernel<<<blocks_for(3 * T, kT), kT, 0, s>>>((const long long*)acd, T, bits, idx32, keys, vals);
    PGH_CUDA(cub::DeviceRadixSort::SortPairs(tmp3, tb, (const unsigned int*)keys, ks3, (const int*)vals,
                                             perm3, (int)(3 * T), 0, bits + 2, s));
    Regroup3 R;
    R.n_rows[0] = (int)n_out; R.n_rows[1] = (int)n_a; R.n_rows[2] = (int)n_b;
    R.rowptr[0] = rowptr_a; R.rowptr[1] = rowptr_c; R.rowptr[2] = rowptr_d;
    R.f[0] = c32; R.g[0] = d32; R.fo[0] = first_a; R.go[0] = second_a;
    R.f[1] = a32; R.g[1] = d32; R.fo[1] = first_c; R.go[1] = second_c;
    R.f[2] = a32; R.g[2] = c32; R.fo[2] = first_d; R.go[2] = second_d;
    const unsigned int mask = (1u << bits) - 1u;
    rowptr3_kernel<<<dim3(blocks_for(T + 1, kT), 3), kT, 0, s>>>(ks3, T, mask, R);
    gather3_kernel<<<dim3(blocks_for(T, kT), 3), kT, 0, s>>>(perm3, T, R);
    return check_launch("acd_regroup");
  }
  if (T > 0) i64_to_i32_kernel<<<blocks_for(3 * T, kT), kT, 0, s>>>((const long long*)acd, 3 * T, idx32, nullptr);
  const size_t arr = align_up(sizeof(int) * (size_t)(T > 0 ? T : 1));
  int* iota = reinterpret_cast<int*>(ws);
  int* ks = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + arr);
  int* perm = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + 2 * arr);
  void* tmp = reinterpret_cast<char*>(ws) + 3 * arr;
  size_t tmp_bytes = ws_bytes - 3 * arr;
  if (T > 0 && which) iota_kernel<<<blocks_for(T, kT), kT, 0, s>>>(iota, T);
  struct G { int bit; const int32_t* key; int64_t n_rows; const int32_t* f; const int32_t* g; int32_t *rp, *fo, *go; };
  const G groups[3] = {{1, a32, n_out, c32, d32, rowptr_a, first_a, second_a},
                       {2, c32, n_a, a32, d32, rowptr_c, first_c, second_c},
                       {4, d32, n_b, a32, c32, rowptr_d, first_d, second_d}};
  for (const G& g : groups) {
    if (!(which & g.bit)) continue;
    if (!g.rp || (T > 0 && (!g.fo || !g.go))) return arg_error("acd_regroup: null output");
    if (T > 0) {
      size_t tb = tmp_bytes;
      PGH_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, (const unsigned int*)g.key, (unsigned int*)ks,
                                               (const int*)iota, perm, (int)T, 0, bit_length(g.n_rows), s));
    }
    rowptr_kernel<<<blocks_for(T + 1, kT), kT, 0, s>>>(ks, T, g.n_rows, g.rp);
    if (T > 0) gather2_i32_kernel<<<blocks_for(T, kT), kT, 0, s>>>(g.f, g.g, perm, T, g.fo, g.go);
  }
  return check_launch("acd_regroup");
}

// ------------------------------------------------------------------ embedding plan in one call
// plans.EmbeddingPlan: idx32, the stable sort of the positions by index value, and the row
// pointers of the reduction tree (all shapes static: V + ceil(size / chunk) rows per level).  One
// 32-bit radix sort and one small kernel per level instead of ~35 torch operations per table.
template <typename T>
__global__ void emb_convert_kernel(const T* __restrict__ idx, long long n, int* __restrict__ idx32,
                                   int* __restrict__ iota) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  idx32[i] = (int)idx[i];
  iota[i] = (int)i;
}

// One level of the tree.  bounds (V + 1, non-decreasing, bounds[V] = size): where each index value's
// range starts in the current level.  level (rows + 1) = sorted union of bounds[0..V) and the cuts
// 0, chunk, 2 chunk, ... (< size), then size; next (V + 1) = position of each value's first row.
__global__ void emb_level_kernel(const int* __restrict__ bounds, int V, int size, int chunk,
                                 int* __restrict__ level, int* __restrict__ next) {
  const int n_cuts = (size + chunk - 1) / chunk;
  const int rows = V + n_cuts;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { level[rows] = bounds[V]; next[V] = rows; }
  if (i < V) {
    const int b = bounds[i];
    const int below_cuts = min(n_cuts, (b + chunk - 1) / chunk);        // cuts < b
    level[i + below_cuts] = b;                     // ties among equal values: any order, same content
    int lo = 0, hi = i;                            // first u with bounds[u] == b
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (bounds[mid] < b) lo = mid + 1; else hi = mid;
    }
    next[i] = lo + below_cuts;
  } else if (i < rows) {
    const int j = i - V, cut = j * chunk;
    int lo = 0, hi = V;                            // number of bounds[0..V) <= cut
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (bounds[mid] <= cut) lo = mid + 1; else hi = mid;
    }
    level[j + lo] = cut;
  }
}

extern "C" size_t pgh_embedding_plan_ws_bytes(int64_t n) {
  if (n <= 0) return 256;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned int*)nullptr, (unsigned int*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)n, 0, 32);
  return align_up(tmp) + 2 * align_up(sizeof(int) * (size_t)n) + align_up(sizeof(int) * 2);
}

// levels: concatenation of the level arrays (sizes as computed by the caller with the same rule:
// while size > max(2 V, 4 chunk): rows = V + ceil(size / chunk), array of rows + 1; then V + 1);
// bounds_ws: 2 * (V + 1) ints of scratch.
extern "C" int pgh_embedding_plan(const void* idx, int idx_is_i64, int64_t n, int64_t V, int64_t chunk,
                                  int32_t* idx32, int32_t* perm, int32_t* levels, int64_t levels_len,
                                  int32_t* bounds_ws, void* ws, size_t ws_bytes, void* stream) {
  if (!idx || !idx32 || !perm || !levels || !bounds_ws || !ws) return arg_error("embedding_plan: null pointer");
  if (n <= 0 || n > 0x7fffffff || V <= 0 || V > (1 << 24) || chunk < 2 || chunk > (1 << 20))
    return arg_error("embedding_plan: sizes");
  if (ws_bytes < pgh_embedding_plan_ws_bytes(n)) return arg_error("embedding_plan: workspace too small");
  cudaStream_t s = as_stream(stream);
  const size_t arr = align_up(sizeof(int) * (size_t)n);
  char* p = static_cast<char*>(ws);
  int* iota = reinterpret_cast<int*>(p);
  int* ks = reinterpret_cast<int*>(p + arr);
  void* tmp = p + 2 * arr;
  size_t tb = ws_bytes - 2 * arr;
  if (idx_is_i64)
    emb_convert_kernel<long long><<<blocks_for(n, kT), kT, 0, s>>>((const long long*)idx, n, idx32, iota);
  else
    emb_convert_kernel<int><<<blocks_for(n, kT), kT, 0, s>>>((const int*)idx, n, idx32, iota);
  PGH_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, (const unsigned int*)idx32, (unsigned int*)ks,
                                           (const int*)iota, perm, (int)n, 0, bit_length(V), s));
  int* bounds = bounds_ws;
  int* next = bounds_ws + (V + 1);
  rowptr_kernel<<<blocks_for(n + 1, kT), kT, 0, s>>>(ks, n, V, bounds);
  int64_t size = n, off = 0;
  const int64_t stop = (2 * V > 4 * chunk) ? 2 * V : 4 * chunk;
  while (size > stop) {
    const int64_t rows = V + (size + chunk - 1) / chunk;
    if (off + rows + 1 > levels_len) return arg_error("embedding_plan: levels buffer too small");
    emb_level_kernel<<<blocks_for(rows, kT), kT, 0, s>>>(bounds, (int)V, (int)size, (int)chunk,
                                                        levels + off, next);
    off += rows + 1;
    size = rows;
    int* t = bounds; bounds = next; next = t;
  }
  if (off + V + 1 > levels_len) return arg_error("embedding_plan: levels buffer too small");
  PGH_CUDA(cudaMemcpyAsync(levels + off, bounds, sizeof(int) * (size_t)(V + 1), cudaMemcpyDeviceToDevice, s));
  return check_launch("embedding_plan");
}

// ------------------------------------------------------------------ row-wise merge of two CSRs
// Row r of the result = the entries of g1's row r followed by those of g2's row r; first indices
// are remapped to stride * first + off (plans.merge_groups: the one-launch SSWL gradient).  Filler
// entries (behind rowptr[n_rows] of their grouping) land, as zeros, behind the result's last row.
__device__ __forceinline__ int row_of_entry(const int* __restrict__ rowptr, int n_rows, int t) {
  int lo = 0, hi = n_rows;                    // largest r with rowptr[r] <= t (n_rows: filler)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (rowptr[mid] <= t) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void merge_rowptr_kernel(const int* __restrict__ rp1, const int* __restrict__ rp2,
                                    long long n1, int* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n1) out[i] = rp1[i] + rp2[i];
}

__global__ void merge_entries_kernel(const int* __restrict__ rp1, const int* __restrict__ f1,
                                     const int* __restrict__ s1, int T1,
                                     const int* __restrict__ rp2, const int* __restrict__ f2,
                                     const int* __restrict__ s2, int T2, int n_rows, int stride,
                                     int off1, int off2, int* __restrict__ fo, int* __restrict__ so) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)T1 + T2) return;
  if (i < T1) {
    const int t = (int)i, r = row_of_entry(rp1, n_rows, t);
    const bool real = t < rp1[n_rows];
    const int pos = t + rp2[r];
    fo[pos] = real ? f1[t] * stride + off1 : 0;
    so[pos] = real ? s1[t] : 0;
  } else {
    const int t = (int)(i - T1), r = row_of_entry(rp2, n_rows, t);
    const bool real = t < rp2[n_rows];
    const int pos = t + rp1[real ? r + 1 : n_rows];
    fo[pos] = real ? f2[t] * stride + off2 : 0;
    so[pos] = real ? s2[t] : 0;
  }
}

extern "C" int pgh_merge_groups_i32(const int32_t* rowptr1, const int32_t* first1, const int32_t* second1,
                                    int64_t T1, const int32_t* rowptr2, const int32_t* first2,
                                    const int32_t* second2, int64_t T2, int64_t n_rows, int stride,
                                    int off1, int off2, int32_t* rowptr_out, int32_t* first_out,
                                    int32_t* second_out, void* stream) {
  if (!rowptr1 || !rowptr2 || !rowptr_out || n_rows < 0 || T1 < 0 || T2 < 0 || T1 + T2 > 0x7fffffff)
    return arg_error("merge_groups: arguments");
  if (T1 + T2 > 0 && (!first_out || !second_out || (T1 > 0 && (!first1 || !second1)) ||
                      (T2 > 0 && (!first2 || !second2))))
    return arg_error("merge_groups: null pointer");
  cudaStream_t s = as_stream(stream);
  merge_rowptr_kernel<<<blocks_for(n_rows + 1, kT), kT, 0, s>>>(rowptr1, rowptr2, n_rows + 1, rowptr_out);
  if (T1 + T2 > 0)
    merge_entries_kernel<<<blocks_for(T1 + T2, kT), kT, 0, s>>>(
        rowptr1, first1, second1, (int)T1, rowptr2, first2, second2, (int)T2, (int)n_rows, stride, off1,
        off2, first_out, second_out);
  return check_launch("merge_groups");
}

// 32-bit keys: what every CSR builder sorts (row ids) -- half the key traffic of the 64-bit entry
// point and no conversion kernels around it.  Workspace: pgh_sort_ws_bytes(n).
extern "C" int pgh_sort_i32_perm(const int32_t* key_in, int64_t n, int end_bit, int32_t* key_out,
                                 int32_t* perm_out, void* ws, size_t ws_bytes, void* stream) {
  if (n <= 0) return 0;
  if (n > 0x7fffffff) return arg_error("sort: n too large");
  if (!key_in || !key_out || !perm_out || !ws) return arg_error("sort: null pointer");
  if (end_bit < 1) end_bit = 1;
  if (end_bit > 32) end_bit = 32;
  cudaStream_t s = as_stream(stream);
  const size_t iota_bytes = align_up(sizeof(int) * (size_t)n);
  if (ws_bytes < iota_bytes + 256) return arg_error("sort: workspace too small");
  int* iota = reinterpret_cast<int*>(ws);
  void* tmp = reinterpret_cast<char*>(ws) + iota_bytes;
  size_t tmp_bytes = ws_bytes - iota_bytes;
  iota_kernel<<<blocks_for(n, kT), kT, 0, s>>>(iota, n);
  PGH_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const unsigned int*)key_in,
                                           (unsigned int*)key_out, (const int*)iota, (int*)perm_out,
                                           (int)n, 0, end_bit, s));
  return check_launch("sort_i32_perm");
}

extern "C" size_t pgh_unique_ws_bytes(int64_t n) {
  if (n <= 0) return 256;
  size_t tmp = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tmp, (const int*)nullptr, (int*)nullptr, (int)n);
  return align_up(tmp);
}

extern "C" int pgh_unique_sorted(const int64_t* key_sorted, int64_t n, int64_t* ukey, int32_t* seg,
                                 int32_t* count_dev, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = as_stream(stream);
  if (n <= 0) {
    PGH_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(int), s));
    return 0;
  }
  head_flags_kernel<<<blocks_for(n, kT), kT, 0, s>>>((const long long*)key_sorted, n, seg);
  size_t tmp = ws_bytes;
  PGH_CUDA(cub::DeviceScan::InclusiveSum(ws, tmp, (const int*)seg, (int*)seg, (int)n, s));
  unique_finish_kernel<<<blocks_for(n, kT), kT, 0, s>>>((const long long*)key_sorted, n, seg,
                                                        (long long*)ukey, count_dev);
  return check_launch("unique_sorted");
}

extern "C" int pgh_rowptr_from_sorted(const int32_t* key_sorted, int64_t n, int64_t n_rows,
                                      int32_t* rowptr, void* stream) {
  if (n_rows < 0 || n < 0) return arg_error("rowptr: sizes");
  rowptr_kernel<<<blocks_for(n + 1, kT), kT, 0, as_stream(stream)>>>(key_sorted, n, n_rows, rowptr);
  return check_launch("rowptr_from_sorted");
}

extern "C" size_t pgh_match_ws_bytes(int64_t m) {
  if (m <= 0) return 256;
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const long long*)nullptr, (long long*)nullptr, (int)m);
  return align_up(tmp) + align_up(sizeof(long long) * (size_t)m);
}

extern "C" int pgh_match_ranges(const int64_t* sorted_keys, int64_t n, const int64_t* queries,
                                int64_t m, int32_t* lo, int64_t* off, void* ws, size_t ws_bytes,
                                void* stream) {
  cudaStream_t s = as_stream(stream);
  if (m <= 0) {
    PGH_CUDA(cudaMemsetAsync(off, 0, sizeof(int64_t), s));
    return 0;
  }
  const size_t cnt_bytes = align_up(sizeof(long long) * (size_t)m);
  if (ws_bytes < cnt_bytes + 256) return arg_error("match: workspace too small");
  long long* cnt = reinterpret_cast<long long*>(ws);
  void* tmp = reinterpret_cast<char*>(ws) + cnt_bytes;
  size_t tmp_bytes = ws_bytes - cnt_bytes;
  match_kernel<<<blocks_for(m, kT), kT, 0, s>>>((const long long*)sorted_keys, n,
                                               (const long long*)queries, m, lo, cnt);
  PGH_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (const long long*)cnt, (long long*)off,
                                         (int)m, s));
  set_last_kernel<<<1, 32, 0, s>>>((const long long*)off, cnt, m, (long long*)off);
  return check_launch("match_ranges");
}

extern "C" int pgh_expand_pairs(const int64_t* off, const int32_t* lo, const int32_t* perm2,
                                int64_t m, int64_t total, int32_t* c, int32_t* d, void* stream) {
  if (total <= 0) return 0;
  expand_kernel<<<blocks_for(total, kT), kT, 0, as_stream(stream)>>>((const long long*)off, lo,
                                                                     perm2, m, total, c, d);
  return check_launch("expand_pairs");
}

extern "C" int pgh_pair_keys(const int64_t* ind1, int64_t ld1, int sd1, int dim1,
                             const int64_t* ind2, int64_t ld2, int sd2, int dim2, const int32_t* c,
                             const int32_t* d, int64_t total, int bits, int64_t* key, int32_t* info,
                             void* stream) {
  if (sd1 < 1 || sd2 < 1 || sd1 + sd2 - 2 > 8 || sd1 + sd2 - 2 < 1)
    return arg_error("pair_keys: sparse dims");
  PairSel sel;
  sel.n1 = sel.n2 = 0;
  for (int r = 0; r < sd1; ++r) if (r != dim1) sel.rows1[sel.n1++] = r;
  for (int r = 0; r < sd2; ++r) if (r != dim2) sel.rows2[sel.n2++] = r;
  if (total <= 0) return 0;
  pair_keys_kernel<<<blocks_for(total, kT), kT, 0, as_stream(stream)>>>(
      (const long long*)ind1, ld1, (const long long*)ind2, ld2, sel, c, d, total, bits,
      (long long*)key, info);
  return check_launch("pair_keys");
}

extern "C" int pgh_lookup_sorted(const int64_t* tkeys, int64_t nt, const int64_t* keys, int64_t n,
                                 int32_t* pos, void* stream) {
  if (n <= 0) return 0;
  lookup_kernel<<<blocks_for(n, kT), kT, 0, as_stream(stream)>>>((const long long*)tkeys, nt,
                                                                 (const long long*)keys, n, pos);
  return check_launch("lookup_sorted");
}

extern "C" size_t pgh_compact_ws_bytes(int64_t n) {
  if (n <= 0) return 256;
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int*)nullptr, (int*)nullptr, (int)n);
  return align_up(tmp) + 2 * align_up(sizeof(int) * (size_t)n);
}

extern "C" int pgh_compact_triples(const int32_t* a, const int32_t* map, const int32_t* c,
                                   const int32_t* d, int64_t n, int32_t* oa, int32_t* oc,
                                   int32_t* od, int32_t* count_dev, void* ws, size_t ws_bytes,
                                   void* stream) {
  cudaStream_t s = as_stream(stream);
  if (n <= 0) {
    PGH_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(int), s));
    return 0;
  }
  const size_t arr = align_up(sizeof(int) * (size_t)n);
  if (ws_bytes < 2 * arr + 256) return arg_error("compact: workspace too small");
  int* a_mapped = reinterpret_cast<int*>(ws);
  int* flag = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + arr);
  void* tmp = reinterpret_cast<char*>(ws) + 2 * arr;
  size_t tmp_bytes = ws_bytes - 2 * arr;
  keep_flags_kernel<<<blocks_for(n, kT), kT, 0, s>>>(a, map, n, a_mapped, flag);
  PGH_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (const int*)flag, flag, (int)n, s));
  compact_scatter_kernel<<<blocks_for(n, kT), kT, 0, s>>>(a_mapped, c, d, flag, n, oa, oc, od,
                                                          count_dev);
  return check_launch("compact_triples");
}

extern "C" int pgh_i64_to_i32(const int64_t* src, int64_t n, int32_t* dst, int32_t* info,
                              void* stream) {
  if (n <= 0) return 0;
  i64_to_i32_kernel<<<blocks_for(n, kT), kT, 0, as_stream(stream)>>>((const long long*)src, n, dst, info);
  return check_launch("i64_to_i32");
}

extern "C" int pgh_i32_to_i64(const int32_t* src, int64_t n, int64_t* dst, void* stream) {
  if (n <= 0) return 0;
  i32_to_i64_kernel<<<blocks_for(n, kT), kT, 0, as_stream(stream)>>>(src, n, (long long*)dst);
  return check_launch("i32_to_i64");
}

extern "C" int pgh_gather_i32(const int32_t* src, const int32_t* idx, int64_t n, int32_t* dst,
                              void* stream) {
  if (n <= 0) return 0;
  gather_i32_kernel<<<blocks_for(n, kT), kT, 0, as_stream(stream)>>>(src, idx, n, dst);
  return check_launch("gather_i32");
}

extern "C" int pgh_gather_i64_as_i32(const int64_t* src, const int32_t* idx, int64_t n,
                                     int32_t* dst, void* stream) {
  if (n <= 0) return 0;
  gather_i64_i32_kernel<<<blocks_for(n, kT), kT, 0, as_stream(stream)>>>((const long long*)src, idx, n, dst);
  return check_launch("gather_i64_as_i32");
}

extern "C" int pgh_unpack_tight(const int64_t* key, int64_t n, const int64_t* dims_host, int sd,
                                int64_t* out, int64_t ld, void* stream) {
  RowSel s;
  int32_t rows[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  if (int e = fill_rowsel(s, rows, dims_host, sd)) return e;
  if (n <= 0) return 0;
  unpack_tight_kernel<<<blocks_for(n, kT), kT, 0, as_stream(stream)>>>((const long long*)key, n, s,
                                                                       (long long*)out, ld);
  return check_launch("unpack_tight");
}

extern "C" int pgh_check_sorted_i64(const int64_t* key, int64_t n, int strict, int32_t* info,
                                    void* stream) {
  if (n <= 1) return 0;
  check_sorted_kernel<<<blocks_for(n, kT), kT, 0, as_stream(stream)>>>((const long long*)key, n, strict, info);
  return check_launch("check_sorted");
}
