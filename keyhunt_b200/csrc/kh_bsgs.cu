// kh_bsgs.cu — BSGS half of the C ABI: baby-step table + 3-tier bloom build, giant-step walk, tier-2/3
// refinement, all resident in HBM (no host round trip between a tier-1 positive and the verified key).
//
// Kernels:
//   kh_baby_kernel    walk of (i+1)*G, i < m: atomicOr into the 3 x 256 bloom shards + bP table fill
//                     (thread_bPload keyhunt.cpp:5284-5472)
//   kh_bitonic_step   sorts the bP table by (6-byte key, index)     (bsgs_sort keyhunt.cpp:4412)
//   kh_giant_kernel   walk of Q - (base + (2g+1)m)G with the tier-1 probe fused in
//                     (thread_process_bsgs keyhunt.cpp:4644-4823)
//   kh_refine_kernel  one warp per tier-1 positive: lanes = the 32 tier-2 sub-steps, then the 32
//                     tier-3 sub-steps, bP lookup and key verification
//                     (bsgs_secondcheck :5151, bsgs_thirdcheck :5186, bsgs_searchbinary :4510)
//   kh_amp_kernel     BSGS_AMP2 / BSGS_AMP3 tables (keyhunt.cpp:1818-1842)
#include <algorithm>

#include "kh_ctx.cuh"

using namespace kh;

// CTA shape of the walk kernels of this file (A/B knobs; T is a multiple of 256 whatever the shape, see kh_pick_T)
#ifndef KH_BSGS_BLOCK
#define KH_BSGS_BLOCK 256
#endif
#define KH_BLOCK KH_BSGS_BLOCK
#ifndef KH_GIANT_MINBLOCKS
#define KH_GIANT_MINBLOCKS (512 / KH_BSGS_BLOCK)
#endif
#ifndef KH_BABY_MINBLOCKS
#define KH_BABY_MINBLOCKS (512 / KH_BSGS_BLOCK)
#endif

__device__ __forceinline__ void kh_stage_table_b(uint32_t *smem, const uint32_t *gtab) {
  const uint4 *src = reinterpret_cast<const uint4 *>(gtab);
  uint4 *dst = reinterpret_cast<uint4 *>(smem);
  for (int i = threadIdx.x; i < KH_TAB_WORDS / 4; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

__global__ void __launch_bounds__(KH_BLOCK, KH_BABY_MINBLOCKS) kh_baby_kernel(WalkParams wp, BsgsTables bt, BabyBins bins) {
  extern __shared__ __align__(16) uint32_t kh_smem_tab[];
  kh_stage_table_b(kh_smem_tab, wp.gtab);
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= wp.T) return;
  BabyEmit emit(bt, bins);
  walk_batches(wp, kh_smem_tab, t, emit);
}

// applies the binned records (emit.cuh BabyBins).  CTAs are dispatched in index order = (shard, pass, bucket of the shard, block),
// so the resident CTAs work on ONE slice of ONE tier-1 shard (a pass covers the bit positions [pass, pass+1) * pass_bits of the
// shard: <= 48 MB, L2-resident; a shard of -k 512 is one slice, one of -k 4096 six) and on a few 16 MB regions of the prefix bitmap
__global__ void __launch_bounds__(256) kh_baby_apply(BabyBins bins, BsgsTables bt, uint32_t blocks_per_bucket, uint32_t n_pass, uint64_t pass_bits) {
  constexpr uint32_t BPS = 1u << (KH_BABY_BUCKET_BITS - 8);      // buckets per shard
  uint32_t idx = blockIdx.x;
  const uint32_t blk = idx % blocks_per_bucket; idx /= blocks_per_bucket;
  const uint32_t bis = idx % BPS; idx /= BPS;
  const uint32_t pass = idx % n_pass, shard = idx / n_pass;
  const uint32_t bucket = shard * BPS + bis;
  const uint32_t slot = blk * 256u + threadIdx.x;
  uint32_t n = bins.count[bucket];
  if (n > bins.cap) n = bins.cap;
  if (slot >= n) return;
  const uint64_t at = (uint64_t)bucket * bins.cap + slot;
  const uint64_t a = bins.a[at], b = bins.b[at];
  if (n_pass == 1) {
    bloom_set(bt.tier[0], shard, a, b);
  } else {
    const BloomDev &bl = bt.tier[0];
    uint8_t *bf = bl.bf + (uint64_t)shard * bl.stride;
    const uint64_t lo = (uint64_t)pass * pass_bits;
    uint64_t x = a;
#pragma unroll 1
    for (uint32_t i = 0; i < bl.hashes; i++) {
      const uint64_t r = bloom_mod(x, bl.bits, bl.magic);
      if (r - lo < pass_bits) {
        const uint64_t byte = r >> 3;
        atomicOr(reinterpret_cast<uint32_t *>(bf + (byte & ~3ULL)), 1u << (8u * (uint32_t)(byte & 3) + (uint32_t)(r & 7)));
      }
      x += b;
    }
  }
  if (bt.pre_k && pass == 0) {
    const uint64_t idx2 = ((uint64_t)bucket << (bt.pre_k - KH_BABY_BUCKET_BITS)) | bins.lo[at];
    atomicOr(bt.pre + (idx2 >> 5), 1u << (uint32_t)(idx2 & 31));
  }
}

__global__ void __launch_bounds__(KH_BLOCK, KH_GIANT_MINBLOCKS) kh_giant_kernel(WalkParams wp, GiantParams gp) {
  extern __shared__ __align__(16) uint32_t kh_smem_tab[];
  kh_stage_table_b(kh_smem_tab, wp.gtab);
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= wp.T) return;
  GiantEmit emit(gp);
  walk_batches(wp, kh_smem_tab, t, emit);
}

// ---- bP table sort -----------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t bp_key48(const BpEntry &e) {
  return ((uint64_t)e.value[0] << 40) | ((uint64_t)e.value[1] << 32) | ((uint64_t)e.value[2] << 24) |
         ((uint64_t)e.value[3] << 16) | ((uint64_t)e.value[4] << 8) | (uint64_t)e.value[5];
}
__global__ void kh_bp_pad(BpEntry *tab, uint64_t m3, uint64_t n2) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + m3;
  if (i >= n2) return;
  BpEntry e;
  for (int k = 0; k < 6; k++) e.value[k] = 0xFF;
  e.pad[0] = e.pad[1] = 0;
  e.index = ~0ULL;   // sentinels sort after every real entry
  tab[i] = e;
}
__global__ void kh_bitonic_step(BpEntry *tab, uint64_t n2, uint64_t j, uint64_t k) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const uint64_t ixj = i ^ j;
  if (ixj <= i) return;
  BpEntry a = tab[i], b = tab[ixj];
  const uint64_t ka = bp_key48(a), kb = bp_key48(b);
  const bool a_gt_b = (ka != kb) ? (ka > kb) : (a.index > b.index);
  const bool ascending = (i & k) == 0;
  if (a_gt_b == ascending) { tab[i] = b; tab[ixj] = a; }
}

// sets *flag when two byte ranges differ (kh_bsgs_import)
__global__ void kh_diff_kernel(const uint8_t *a, const uint8_t *b, uint64_t n, uint32_t *flag) {
  const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  bool diff = false;
  for (uint64_t i = i0; i < n && i < i0 + 16; i++) diff |= (a[i] != b[i]);
  if (diff) atomicOr(flag, 1u);
}

// order-independent digest of a byte range (kh_bsgs_digest): sum over its 4-byte words of mix(word, position) mod 2^64.
// Lets a test compare two builds of tables that are too big to bring to the host (7.7 GB tier 1, 64 GB prefix bitmap).
__global__ void __launch_bounds__(256) kh_digest_kernel(const uint32_t *w, uint64_t n_words, uint64_t salt, unsigned long long *out) {
  uint64_t acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v = w[i];
    if (v) {
      uint64_t h = (i + salt) * 0x9E3779B97F4A7C15ull + v;
      h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
      acc += h;
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, (unsigned long long)acc);
}

// ---- AMP tables: entry i (<32) = -(2i+1)*m2*G, entry 32+i = -(2i+1)*m3*G ------------------------------
__global__ void __launch_bounds__(64) kh_amp_kernel(uint32_t *aux, uint64_t m2, uint64_t m3) {
  const uint32_t i = threadIdx.x;
  if (i >= 64) return;
  u256 k, z, mm;
  u256_set_u64(z, 0);
  u256_set_u64(mm, i < 32 ? m2 : m3);
  u256_add_mul64(k, z, mm, 2ull * (i & 31) + 1ull);
  ge p;
  ge_mul_g(p, k);
  ge_neg(p, p);
#pragma unroll
  for (int l = 0; l < 8; l++) { aux[16 * i + l] = p.x.v[l]; aux[16 * i + 8 + l] = p.y.v[l]; }
}

// ---- refinement ---------------------------------------------------------------------------------------
struct RefineParams {
  BsgsTables bt;
  const uint32_t *aux;      // AMP2 | AMP3
  const GiantCand *cands;
  uint32_t n_cands;
  uint32_t base_check;      // server variant: a window's base point equal to the target is a hit (bsgsd.cpp:2544)
  uint64_t steps_per_window;// giant steps between window bases (= aux)
  ge q;                     // target public key
  u256 start;               // range start (base key of window 0)
  uint32_t *found;          // [0] flag
  u256 *found_key;
  const uint32_t *comb;     // fixed-base comb of G (ec.cuh ge_mul_g_comb)
};

__device__ __forceinline__ void load_aux_point(ge &p, const uint32_t *aux, int i) {
#pragma unroll
  for (int l = 0; l < 8; l++) { p.x.v[l] = aux[16 * i + l]; p.y.v[l] = aux[16 * i + 8 + l]; }
  p.inf = 0;
}
__device__ __forceinline__ bool tier_check(const BloomDev &bl, const fe &x) {
  uint32_t w[8];
  fe_to_le_words(w, x);
  const uint64_t a = xxh64_32(w, KH_BLOOM_SEED);
  const uint64_t b = xxh64_32(w, a);
  return bloom_test(bl, x.v[7] >> 24, a, b);
}
// S = Q - key*G  (AddDirect(Q, Negation(ComputePublicKey(key))), keyhunt.cpp:5163-5171)
__device__ __noinline__ void q_minus_key(ge &s, const ge &q, const u256 &key, const uint32_t *comb) {
  ge bp;
  ge_mul_g_comb(bp, key, comb);
  ge_neg(bp, bp);
  ge_add_direct(s, q, bp);
}
__device__ __noinline__ bool key_matches(const ge &q, const u256 &key, const uint32_t *comb) {
  ge p;
  ge_mul_g_comb(p, key, comb);
  return !p.inf && fe_eq(p.x, q.x);
}

__global__ void __launch_bounds__(32) kh_refine_kernel(RefineParams rp) {
  const uint32_t cand = blockIdx.x;
  if (cand >= rp.n_cands) return;
  const uint32_t lane = threadIdx.x;
  const GiantCand c = rp.cands[cand];
  const uint64_t g = c.batch * KH_GRP + c.idx;            // global giant-step number (window*aux + a)
  // base2 = start + g*2m   (bsgs_secondcheck keyhunt.cpp:5159-5161, windows being contiguous)
  u256 two_m, base2;
  u256_set_u64(two_m, rp.bt.m);
  { u256 z; u256_set_u64(z, 0); u256_add_mul64(two_m, z, two_m, 2); }
  u256_add_mul64(base2, rp.start, two_m, g);
  // The reference's BSGS server compares base_key*G with the target before it walks a window (bsgsd.cpp:2528-2563).
  // That case is giant step 0 of the window meeting baby step m (Q - (base+m)G = -mG), so it always arrives here as a
  // tier-1 positive; the checks below cannot confirm it (they do not see offsets 0..2m of a window), this one does.
  if (rp.base_check && (g % rp.steps_per_window) == 0) {
    ge bp;
    ge_mul_g_comb(bp, base2, rp.comb);
    if (!bp.inf && fe_eq(bp.x, rp.q.x) && fe_eq(bp.y, rp.q.y)) {
      if (lane == 0 && atomicCAS(rp.found, 0u, 1u) == 0u) *rp.found_key = base2;
      return;
    }
  }
  ge S;
  q_minus_key(S, rp.q, base2, rp.comb);
  // tier 2: lane i2 tests S + AMP2[i2]
  ge amp, P2;
  load_aux_point(amp, rp.aux, (int)lane);
  ge_add_direct(P2, S, amp);
  const bool pos2 = tier_check(rp.bt.tier[1], P2.x);
  uint32_t mask2 = __ballot_sync(0xFFFFFFFFu, pos2);
  while (mask2) {
    const uint32_t i2 = __ffs(mask2) - 1;
    mask2 &= mask2 - 1;
    // base3 = base2 + i2*2*m2   (bsgs_thirdcheck keyhunt.cpp:5194-5196)
    u256 two_m2, base3;
    u256_set_u64(two_m2, rp.bt.m2);
    { u256 z; u256_set_u64(z, 0); u256_add_mul64(two_m2, z, two_m2, 2); }
    u256_add_mul64(base3, base2, two_m2, (uint64_t)i2);
    ge S3, P3;
    q_minus_key(S3, rp.q, base3, rp.comb);
    load_aux_point(amp, rp.aux, 32 + (int)lane);
    ge_add_direct(P3, S3, amp);
    // calcualteindex(i3) = (2*i3+1)*m3   (keyhunt.cpp:7859)
    u256 calc, m3v, key;
    u256_set_u64(m3v, rp.bt.m3);
    { u256 z; u256_set_u64(z, 0); u256_add_mul64(calc, z, m3v, 2ull * lane + 1ull); }
    bool ok = false;
    if (tier_check(rp.bt.tier[2], P3.x)) {
      // bsgs_searchbinary on X bytes 16..21; every entry sharing the 6-byte key is tried
      const uint64_t want = ((uint64_t)P3.x.v[3] << 16) | (P3.x.v[2] >> 16);
      uint64_t lo = 0, hi = rp.bt.m3;
      while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (bp_key48(rp.bt.table[mid]) < want) lo = mid + 1; else hi = mid;
      }
      while (!ok && lo < rp.bt.m3 && bp_key48(rp.bt.table[lo]) == want) {
        const uint64_t j1 = rp.bt.table[lo].index + 1;
        u256 t;
        u256_add_mul64(t, base3, calc, 1);
        u256_add_u64(key, t, j1);                                   // keyhunt.cpp:5212-5219
        if (key_matches(rp.q, key, rp.comb)) ok = true;
        else { u256_sub_u64(key, t, j1); if (key_matches(rp.q, key, rp.comb)) ok = true; }   // :5221-5228
        lo++;
      }
    } else if (fe_eq(S3.x, amp.x)) {                                // keyhunt.cpp:5238-5243
      u256_add_mul64(key, base3, calc, 1);
      ok = true;
    }
    const uint32_t okmask = __ballot_sync(0xFFFFFFFFu, ok);
    if (okmask) {
      if (lane == (uint32_t)(__ffs(okmask) - 1)) {
        if (atomicCAS(rp.found, 0u, 1u) == 0u) *rp.found_key = key;
      }
      return;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
static int ilog2_exact(uint64_t n) {
  int lg = 0;
  while (lg < 63 && (1ULL << lg) < n) lg++;
  return ((1ULL << lg) == n) ? lg : -1;
}
static void free_bsgs(kh_ctx *c) {
  for (int t = 0; t < 3; t++) { if (c->d_tier[t]) cudaFree(c->d_tier[t]); c->d_tier[t] = nullptr; }
  if (c->d_bsgs_pre) cudaFree(c->d_bsgs_pre);
  c->d_bsgs_pre = nullptr; c->bsgs_pre_k = 0;
  if (c->d_bptable) cudaFree(c->d_bptable);
  if (c->d_aux_tab) cudaFree(c->d_aux_tab);
  c->d_bptable = nullptr; c->d_aux_tab = nullptr;
  c->have_bsgs = false;
}
static void fill_tables(kh_ctx *c, BsgsTables &bt) {
  for (int t = 0; t < 3; t++) {
    bt.tier[t].bf = c->d_tier[t];
    bt.tier[t].bits = c->bsgs.tier[t].bits;
    bt.tier[t].magic = (~0ULL) / c->bsgs.tier[t].bits;
    bt.tier[t].stride = c->tier_stride[t];
    bt.tier[t].hashes = c->bsgs.tier[t].hashes;
    bt.tier[t].pad = 0;
  }
  bt.table = c->d_bptable;
  bt.m = c->bsgs.m; bt.m2 = c->bsgs.m2; bt.m3 = c->bsgs.m3;
  bt.pre = c->d_bsgs_pre; bt.pre_k = c->bsgs_pre_k; bt.pad = 0;
}

extern "C" {

int kh_bsgs_build(kh_ctx *c, uint64_t n, uint32_t k) {
  if (!c) return KH_EINVAL;
  cudaSetDevice(c->device);
  // -n must have an exact square root (keyhunt.cpp:1474) that is a multiple of 1024 (:1509)
  const int lg = ilog2_exact(n);
  if (lg < 20 || (lg & 1) || k < 1) return kh_fail(c, KH_EINVAL, "bsgs n must be 2^even >= 2^20 and k >= 1");
  free_bsgs(c);
  kh_bsgs_desc d;
  memset(&d, 0, sizeof(d));
  d.m = (1ULL << (lg / 2)) * (uint64_t)k;                               // keyhunt.cpp:1557
  d.m2 = d.m / 32 + ((d.m % 32) ? 1 : 0);                                // :1561-1566
  d.m3 = d.m2 / 32 + ((d.m2 % 32) ? 1 : 0);                              // :1578-1583
  d.aux = n / d.m;                                                       // :1591-1603
  if (d.aux == 0) return kh_fail(c, KH_EINVAL, "k too large for n");
  d.n = (n % d.m) ? d.m * d.aux : n;
  const uint64_t items[3] = {
      (d.m / 256 > 10000) ? (d.m / 256 + ((d.m % 256) ? 1 : 0)) : 1000,  // :1633-1661
      (d.m2 / 256 > 1000) ? (d.m2 / 256 + ((d.m2 % 256) ? 1 : 0)) : 1000,
      (d.m3 / 256 > 1000) ? (d.m3 / 256 + ((d.m3 % 256) ? 1 : 0)) : 1000};
  for (int t = 0; t < 3; t++)
    if (kh_bloom_params(items[t] <= 10000 ? 10000 : items[t], &d.tier[t]) != KH_OK) return kh_fail(c, KH_EINVAL, "bloom sizing");
  c->bsgs = d;
  for (int t = 0; t < 3; t++) {
    c->tier_stride[t] = ((d.tier[t].bytes + 15) / 16) * 16;
    KH_CUDA(c, cudaMalloc(&c->d_tier[t], (size_t)c->tier_stride[t] * 256));
    KH_CUDA(c, cudaMemsetAsync(c->d_tier[t], 0, (size_t)c->tier_stride[t] * 256, c->stream));
  }
  uint64_t n2 = 1;
  while (n2 < d.m3) n2 <<= 1;
  KH_CUDA(c, cudaMalloc(&c->d_bptable, (size_t)n2 * sizeof(BpEntry)));
  KH_CUDA(c, cudaMalloc(&c->d_aux_tab, 64 * 16 * sizeof(uint32_t)));

  // baby steps: point p = (p+1)*G
  const uint64_t n_batches = (d.m + KH_GRP - 1) / KH_GRP;
  const uint64_t T = kh_pick_T(c, n_batches);
  int rc = kh_ensure_walk_buffers(c, T);
  if (rc) return rc;
  WalkSetup ws;
  memset(&ws, 0, sizeof(ws));
  u256_set_u64(ws.s, 1);
  u256_set_u64(ws.k0, 1);
  ws.q.inf = 1; ws.neg = 0; ws.T = T; ws.first_batch = 0;
  rc = kh_run_setup(c, ws);
  if (rc) return rc;
  // prefix bitmap over the baby points: 256 bits per point (0.4 % fill) when HBM allows, never more than 3/4 of what is
  // free now (the walk buffers are already allocated), never less than 8 bits per point (then it is not worth a probe)
  if (c->bsgs_prefilter) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    uint32_t k = 20;
    while (k < 48 && (1ull << k) < 256ull * d.m) k++;
    while (k > 20 && (1ull << (k - 3)) > (uint64_t)free_b / 4 * 3) k--;
    if (const char *e = getenv("KH_BSGS_PRE_LOG2")) { const uint32_t cap = (uint32_t)atoi(e); if (cap >= 20 && cap < k) k = cap; }   // experiments
    if ((1ull << k) >= 8ull * d.m && cudaMalloc(&c->d_bsgs_pre, (size_t)1 << (k - 3)) == cudaSuccess) {
      KH_CUDA(c, cudaMemsetAsync(c->d_bsgs_pre, 0, (size_t)1 << (k - 3), c->stream));
      c->bsgs_pre_k = k;
    } else {
      cudaGetLastError();
      c->d_bsgs_pre = nullptr;
    }
  }
  BsgsTables bt;
  fill_tables(c, bt);
  WalkParams wp;
  wp.gtab = c->d_gtab; wp.centers = c->d_centers; wp.scratch = c->d_scratch;
  wp.T = T; wp.n_batches = n_batches; wp.steps = (uint32_t)c->steps_per_launch; wp.pad = 0;
  // binned build (emit.cuh BabyBins) for tables that do not fit L2 anyway: one step of the walk per launch, then its buckets
  // are applied.  The record buffers (20 B per point of one step, ~12.7 GB for 606,208 walkers) live only during the build.
  BabyBins bins;
  memset(&bins, 0, sizeof(bins));
  void *bin_mem = nullptr;
  uint32_t blocks_per_bucket = 0;
  const uint32_t n_buckets = 1u << KH_BABY_BUCKET_BITS;
  // a tier-1 shard is applied in slices of <= 48 MB so that the slice being written stays in L2 (KH_BABY_SLICE_KB: tests, experiments)
  uint64_t slice_bytes = 48ull << 20;
  if (const char *e = getenv("KH_BABY_SLICE_KB")) { const long v = atol(e); if (v >= 1) slice_bytes = (uint64_t)v << 10; }
  uint32_t n_pass = (uint32_t)((d.tier[0].bytes + slice_bytes - 1) / slice_bytes);
  if (n_pass < 1) n_pass = 1;
  if (n_pass > 64) n_pass = 64;

  if (c->bsgs_binned_build && (c->bsgs_binned_build == 2 || d.m >= (1ull << 26)) && (c->bsgs_pre_k == 0 || (c->bsgs_pre_k >= KH_BABY_BUCKET_BITS + 8 && c->bsgs_pre_k - KH_BABY_BUCKET_BITS <= 32))) {
    const uint64_t per_launch = std::min<uint64_t>(T * (uint64_t)KH_GRP, d.m);
    const uint64_t avg = per_launch / n_buckets;
    const uint64_t cap = avg + avg / 32 + 4096;                     // ~6 sigma above the mean for uniformly distributed X
    const size_t recs = (size_t)cap * n_buckets;
    if (cap < (1ull << 31) && cudaMalloc(&bin_mem, recs * 20 + n_buckets * sizeof(uint32_t) + 64) == cudaSuccess) {
      bins.a = static_cast<uint64_t *>(bin_mem);
      bins.b = bins.a + recs;
      bins.lo = reinterpret_cast<uint32_t *>(bins.b + recs);
      bins.count = bins.lo + recs;
      bins.cap = (uint32_t)cap;
      blocks_per_bucket = (uint32_t)((cap + 255) / 256);
      while (n_pass > 1 && (uint64_t)n_buckets * blocks_per_bucket * n_pass >= (1ull << 31)) n_pass--;   // one grid
      if ((uint64_t)n_buckets * blocks_per_bucket >= (1ull << 31)) { cudaFree(bin_mem); bin_mem = nullptr; memset(&bins, 0, sizeof(bins)); }   // (never with real T)
      else wp.steps = 1;
    } else {
      cudaGetLastError();
      bin_mem = nullptr;
    }
  }
  const uint64_t pass_bits = (d.tier[0].bits + n_pass - 1) / n_pass;
  kh_time_begin(c);
  uint64_t launches = 0;
  for (uint64_t base = 0; base < n_batches; base += (uint64_t)wp.steps * T) {
    wp.batch_base = base;
    if (bins.cap) cudaMemsetAsync(bins.count, 0, n_buckets * sizeof(uint32_t), c->stream);
    kh_baby_kernel<<<(unsigned)(T / KH_BLOCK), KH_BLOCK, KH_TAB_WORDS * sizeof(uint32_t), c->stream>>>(wp, bt, bins);
    launches++;
    if (bins.cap) {
      kh_baby_apply<<<n_buckets * blocks_per_bucket * n_pass, 256, 0, c->stream>>>(bins, bt, blocks_per_bucket, n_pass, pass_bits);
      launches++;
    }
  }
  c->stats.walk_ms += kh_time_end(c);
  if (bin_mem) cudaFree(bin_mem);
  c->stats.walk_launches += launches;
  c->stats.points += d.m;
  c->stats.walker_threads = T;
  KH_CUDA(c, cudaGetLastError());

  // sort + AMP tables
  kh_time_begin(c);
  uint64_t other = 0;
  if (n2 > d.m3) { kh_bp_pad<<<(unsigned)((n2 - d.m3 + 255) / 256), 256, 0, c->stream>>>(c->d_bptable, d.m3, n2); other++; }
  for (uint64_t kk = 2; kk <= n2; kk <<= 1)
    for (uint64_t j = kk >> 1; j > 0; j >>= 1) {
      kh_bitonic_step<<<(unsigned)((n2 + 255) / 256), 256, 0, c->stream>>>(c->d_bptable, n2, j, kk);
      other++;
    }
  kh_amp_kernel<<<1, 64, 0, c->stream>>>(c->d_aux_tab, d.m2, d.m3);
  other++;
  c->stats.aux_ms += kh_time_end(c);
  c->stats.other_launches += other;
  KH_CUDA(c, cudaGetLastError());
  c->have_bsgs = true;
  return KH_OK;
}

int kh_bsgs_describe(kh_ctx *c, kh_bsgs_desc *out) {
  if (!c || !out) return KH_EINVAL;
  if (!c->have_bsgs) return kh_fail(c, KH_ESTATE, "kh_bsgs_build has not run");
  *out = c->bsgs;
  return KH_OK;
}

static int bsgs_region(kh_ctx *c, int tier, int shard, uint8_t **ptr, uint64_t *len) {
  if (!c->have_bsgs) return kh_fail(c, KH_ESTATE, "kh_bsgs_build has not run");
  if (tier == 0) { *ptr = reinterpret_cast<uint8_t *>(c->d_bptable); *len = c->bsgs.m3 * sizeof(BpEntry); return KH_OK; }
  if (tier < 1 || tier > 3 || shard < 0 || shard > 255) return kh_fail(c, KH_EINVAL, "bad tier/shard");
  *ptr = c->d_tier[tier - 1] + (uint64_t)shard * c->tier_stride[tier - 1];
  *len = c->bsgs.tier[tier - 1].bytes;
  return KH_OK;
}

int kh_bsgs_export(kh_ctx *c, int tier, int shard, void *dst, uint64_t cap) {
  if (!c || !dst) return KH_EINVAL;
  cudaSetDevice(c->device);
  uint8_t *p; uint64_t len;
  int rc = bsgs_region(c, tier, shard, &p, &len);
  if (rc) return rc;
  if (cap < len) return kh_fail(c, KH_EINVAL, "buffer too small (%llu < %llu)", (unsigned long long)cap, (unsigned long long)len);
  KH_CUDA(c, cudaMemcpyAsync(dst, p, len, cudaMemcpyDeviceToHost, c->stream));
  KH_CUDA(c, cudaStreamSynchronize(c->stream));
  return KH_OK;
}

int kh_bsgs_digest(kh_ctx *c, int tier, uint64_t *out) {
  if (!c || !out) return KH_EINVAL;
  cudaSetDevice(c->device);
  if (!c->have_bsgs) return kh_fail(c, KH_ESTATE, "kh_bsgs_build has not run");
  const uint8_t *p; uint64_t len;
  if (tier == 0) { p = reinterpret_cast<const uint8_t *>(c->d_bptable); len = c->bsgs.m3 * sizeof(BpEntry); }
  else if (tier >= 1 && tier <= 3) { p = c->d_tier[tier - 1]; len = (uint64_t)c->tier_stride[tier - 1] * 256; }
  else if (tier == 4) { p = reinterpret_cast<const uint8_t *>(c->d_bsgs_pre); len = c->bsgs_pre_k ? (1ull << (c->bsgs_pre_k - 3)) : 0; }
  else return kh_fail(c, KH_EINVAL, "bad tier");
  unsigned long long *d_out = nullptr;
  KH_CUDA(c, cudaMalloc(&d_out, sizeof(unsigned long long)));
  cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), c->stream);
  if (len) kh_digest_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(reinterpret_cast<const uint32_t *>(p), len / 4, (uint64_t)tier << 56, d_out);
  unsigned long long h = 0;
  cudaMemcpyAsync(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
  cudaStreamSynchronize(c->stream);
  cudaFree(d_out);
  KH_CUDA(c, cudaGetLastError());
  *out = (uint64_t)h + len;       // the length takes part: an absent bitmap (len 0) differs from an empty one
  return KH_OK;
}

int kh_bsgs_import(kh_ctx *c, int tier, int shard, const void *src, uint64_t n) {
  if (!c || !src) return KH_EINVAL;
  cudaSetDevice(c->device);
  uint8_t *p; uint64_t len;
  int rc = bsgs_region(c, tier, shard, &p, &len);
  if (rc) return rc;
  if (n != len) return kh_fail(c, KH_EINVAL, "length mismatch (%llu != %llu)", (unsigned long long)n, (unsigned long long)len);
  if (tier == 1 && c->bsgs_pre_k) {
    // the prefix bitmap describes the baby points kh_bsgs_build walked: it stays valid only while tier 1 holds exactly
    // what was built (the normal case: files written by a build with the same n / k); anything else switches it off
    uint8_t *tmp = nullptr;
    uint32_t *d_flag = nullptr, h_flag = 0;
    KH_CUDA(c, cudaMalloc(&tmp, len));
    KH_CUDA(c, cudaMalloc(&d_flag, sizeof(uint32_t)));
    cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), c->stream);
    cudaMemcpyAsync(tmp, src, len, cudaMemcpyHostToDevice, c->stream);
    kh_diff_kernel<<<(unsigned)((len + 256 * 16 - 1) / (256 * 16)), 256, 0, c->stream>>>(p, tmp, len, d_flag);
    cudaMemcpyAsync(&h_flag, d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    if (h_flag) { cudaMemcpyAsync(p, tmp, len, cudaMemcpyDeviceToDevice, c->stream); cudaStreamSynchronize(c->stream); c->bsgs_pre_k = 0; }
    cudaFree(tmp); cudaFree(d_flag);
    KH_CUDA(c, cudaGetLastError());
    return KH_OK;
  }
  KH_CUDA(c, cudaMemcpyAsync(p, src, len, cudaMemcpyHostToDevice, c->stream));
  KH_CUDA(c, cudaStreamSynchronize(c->stream));
  return KH_OK;
}

int kh_bsgs_search(kh_ctx *c, const uint8_t pub_xy_be[64], const uint8_t start_be[32], const uint8_t end_be[32],
                   uint8_t found_key_be[32], int *found) {
  if (!c || !pub_xy_be || !start_be || !end_be || !found_key_be || !found) return KH_EINVAL;
  if (!c->have_bsgs) return kh_fail(c, KH_ESTATE, "kh_bsgs_search before kh_bsgs_build");
  cudaSetDevice(c->device);
  *found = 0;
  const kh_bsgs_desc &d = c->bsgs;
  u256 start, end;
  u256_from_be(start, start_be);
  u256_from_be(end, end_be);
  // windows: base_w = start + w*2n while base_w < end (keyhunt.cpp:4603-4617)
  u256 diff;
  if (kh_sub8(diff.v, end.v, start.v)) return kh_fail(c, KH_EINVAL, "end < start");
  bool zero = true;
  for (int i = 0; i < 8; i++) zero &= (diff.v[i] == 0);
  if (zero) return KH_OK;
  for (int i = 4; i < 8; i++) if (diff.v[i]) return kh_fail(c, KH_EINVAL, "range wider than 2^128 is not supported");
  const unsigned __int128 width = ((unsigned __int128)diff.v[3] << 96) | ((unsigned __int128)diff.v[2] << 64) |
                                  ((unsigned __int128)diff.v[1] << 32) | diff.v[0];
  const unsigned __int128 step = (unsigned __int128)2 * d.n;
  const unsigned __int128 windows128 = (width + step - 1) / step;
  if (windows128 > ((unsigned __int128)1 << 50)) return kh_fail(c, KH_EINVAL, "too many windows");
  const uint64_t windows = (uint64_t)windows128;
  const uint64_t cycles = d.aux / 1024 + ((d.aux % 1024) ? 1 : 0);       // keyhunt.cpp:4583
  // giant steps of consecutive windows are one arithmetic progression; window w covers steps
  // [w*aux, w*aux + cycles*1024)
  const uint64_t n_steps = (windows - 1) * d.aux + cycles * 1024;
  const uint64_t n_batches = (n_steps + KH_GRP - 1) / KH_GRP;
  const uint64_t T = kh_pick_T(c, n_batches);
  int rc = kh_ensure_walk_buffers(c, T);
  if (rc) return rc;

  WalkSetup ws;
  memset(&ws, 0, sizeof(ws));
  u256 mm, z;
  u256_set_u64(mm, d.m);
  u256_set_u64(z, 0);
  u256_add_mul64(ws.s, z, mm, 2);                 // step scalar 2m, walked downwards
  u256_add_mul64(ws.k0, start, mm, 1);            // giant step 0 is Q - (start + m)G
  ws.neg = 1; ws.T = T; ws.first_batch = 0;
  for (int i = 0; i < 8; i++) {
    const uint8_t *px = pub_xy_be + 4 * (7 - i), *py = pub_xy_be + 32 + 4 * (7 - i);
    ws.q.x.v[i] = ((uint32_t)px[0] << 24) | ((uint32_t)px[1] << 16) | ((uint32_t)px[2] << 8) | px[3];
    ws.q.y.v[i] = ((uint32_t)py[0] << 24) | ((uint32_t)py[1] << 16) | ((uint32_t)py[2] << 8) | py[3];
  }
  ws.q.inf = 0;
  rc = kh_run_setup(c, ws);
  if (rc) return rc;

  // candidate queue + result
  // candidate queue, counters ([0] candidate count, [1] found flag) and the found key live in the context: a server
  // answers thousands of requests with the same three buffers
  const uint32_t cap = 1u << 16;
  if (!c->d_giant_cands) KH_CUDA(c, cudaMalloc(&c->d_giant_cands, cap * sizeof(GiantCand)));
  if (!c->d_giant_cnt) KH_CUDA(c, cudaMalloc(&c->d_giant_cnt, 4 * sizeof(uint32_t)));
  if (!c->d_giant_key) KH_CUDA(c, cudaMalloc(&c->d_giant_key, sizeof(u256)));
  GiantCand *d_cands = static_cast<GiantCand *>(c->d_giant_cands);
  uint32_t *d_cnt = c->d_giant_cnt;
  u256 *d_key = static_cast<u256 *>(c->d_giant_key);
  KH_CUDA(c, cudaMemsetAsync(d_cnt, 0, 4 * sizeof(uint32_t), c->stream));

  BsgsTables bt;
  fill_tables(c, bt);
  GiantParams gp;
  gp.tier1 = bt.tier[0]; gp.cands = d_cands; gp.count = d_cnt; gp.cap = cap; gp.n_steps = n_steps;
  gp.pre = c->d_bsgs_pre; gp.pre_k = c->bsgs_pre_k;
  RefineParams rp;
  rp.bt = bt; rp.aux = c->d_aux_tab; rp.cands = d_cands; rp.n_cands = 0; rp.base_check = (uint32_t)c->bsgs_base_check; rp.steps_per_window = d.aux;
  rp.q = ws.q; rp.start = start; rp.found = d_cnt + 1; rp.found_key = d_key; rp.comb = c->d_comb;

  WalkParams wp;
  wp.gtab = c->d_gtab; wp.centers = c->d_centers; wp.scratch = c->d_scratch;
  wp.T = T; wp.n_batches = n_batches; wp.pad = 0;
  // a found key ends the search (keyhunt.cpp:4644 `bsgs_found[k] == 0`): keep one launch to about 2^29 giant steps
  // (~50 ms) so that the check between launches is fine-grained
  {
    const uint64_t per_step = T * (uint64_t)KH_GRP;
    uint64_t steps = ((1ULL << 29) + per_step - 1) / per_step;
    if (steps < 1) steps = 1;
    if (steps > (uint64_t)c->steps_per_launch) steps = (uint64_t)c->steps_per_launch;
    wp.steps = (uint32_t)steps;
  }

  int result = KH_OK;
  uint64_t steps_done = 0;
  for (uint64_t base = 0; base < n_batches && !*found; base += (uint64_t)wp.steps * T) {
    wp.batch_base = base;
    kh_time_begin(c);
    kh_giant_kernel<<<(unsigned)(T / KH_BLOCK), KH_BLOCK, KH_TAB_WORDS * sizeof(uint32_t), c->stream>>>(wp, gp);
    uint32_t h[2] = {0, 0};
    cudaMemcpyAsync(h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    c->stats.walk_ms += kh_time_end(c);
    c->stats.walk_launches += 1;
    const uint64_t covered = std::min<uint64_t>(n_batches - base, (uint64_t)wp.steps * T) * KH_GRP;
    steps_done += covered;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { result = kh_fail(c, KH_ENODEV, "giant kernel: %s", cudaGetErrorString(e)); break; }
    if (h[0]) {
      uint32_t nc = h[0];
      if (nc > cap) { nc = cap; c->overflowed = true; }
      c->stats.tier1_positives += h[0];
      rp.n_cands = nc;
      kh_time_begin(c);
      kh_refine_kernel<<<nc, 32, 0, c->stream>>>(rp);
      cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t), c->stream);
      uint32_t f = 0;
      cudaMemcpyAsync(&f, d_cnt + 1, sizeof(f), cudaMemcpyDeviceToHost, c->stream);
      c->stats.aux_ms += kh_time_end(c);
      c->stats.other_launches += 1;
      e = cudaGetLastError();
      if (e != cudaSuccess) { result = kh_fail(c, KH_ENODEV, "refine kernel: %s", cudaGetErrorString(e)); break; }
      if (f) {
        u256 key;
        cudaMemcpyAsync(&key, d_key, sizeof(key), cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        u256_to_be(found_key_be, key);
        *found = 1;
      }
    }
  }
  c->stats.points += steps_done;
  c->stats.walker_threads = T;
  if (result == KH_OK && c->overflowed) {
    c->overflowed = false;
    // a found key stands (it was verified on the device); only a search that ended empty-handed may have lost its hit
    if (!*found) return kh_fail(c, KH_EOVERFLOW, "tier-1 candidate queue overflowed");
    c->err = "warning: tier-1 candidate queue overflowed (the key was found nevertheless)";
  }
  return result;
}

}  // extern "C"
