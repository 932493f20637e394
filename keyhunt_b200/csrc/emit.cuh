// emit.cuh — what happens to each point the walk produces: hash -> bloom probe -> sorted-table
// search -> hit record.  Device form of loop C of thread_process (keyhunt.cpp:3475-3830), of the
// tier-1 probe of thread_process_bsgs (:4819-4823) and of the per-point body of thread_bPload
// (:5394-5443).
#pragma once
#include <stdint.h>

#include "bloom.cuh"
#include "walk.cuh"

namespace kh {

// raw device-side records; the host turns them into kh_hit (private key, n-k fix-up, public key)
enum { KH_KIND_COMP02 = 0, KH_KIND_COMP03 = 1, KH_KIND_UNCOMP = 2, KH_KIND_ETH = 3, KH_KIND_XPOINT = 4 };

struct RawHit {
  uint64_t batch;
  uint32_t idx;
  uint32_t kind;
  uint32_t h[5];   // matched 20 bytes as LE words
  uint32_t variant; // with -e: the reference's candidate index l (0..11 BTC, 0..5 ETH, 0..2 xpoint), else 0
};

struct HitSink {
  RawHit *hits;
  uint32_t *count;   // total hits produced (may exceed cap; the excess is dropped and reported)
  uint32_t cap;
  uint32_t pad;
};

KH_HD uint32_t kh_atomic_inc(uint32_t *p) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, 1u);
#else
  return (*p)++;
#endif
}
KH_HD void sink_push(const HitSink &s, uint64_t batch, uint32_t idx, uint32_t kind, const uint32_t h[5], uint32_t variant = 0) {
  uint32_t slot = kh_atomic_inc(s.count);
  if (slot < s.cap) {
    RawHit r;
    r.batch = batch; r.idx = idx; r.kind = kind; r.variant = variant;
#pragma unroll
    for (int i = 0; i < 5; i++) r.h[i] = h[i];
    s.hits[slot] = r;
  }
}

// ---- scan modes ------------------------------------------------------------------------------------
enum { KH_SCAN_XPOINT = 0, KH_SCAN_COMP = 1, KH_SCAN_UNCOMP = 2, KH_SCAN_BOTH = 3, KH_SCAN_ETH = 4 };

struct ScanTargets {
  BloomDev bloom;
  const uint32_t *table;   // N x 5 big-endian-packed words, ascending
  uint64_t n;
  HitSink sink;
  // -m vanity: instead of bloom + table, `van_n` closed intervals [A, B] of hash160 values (vanityrmdmatch
  // keyhunt.cpp:6677).  van = 2048-word bitmap over the first two digest bytes (set where some interval reaches),
  // then van_n x (A[5], B[5]) as big-endian-packed words.
  const uint32_t *van;
  uint32_t van_n;
  // exact pre-filter in front of the bloom: 2^pre_k bits (16 <= pre_k <= 32) indexed by the first pre_k bits of the 20-byte
  // record, set for every target (no false negatives, so bloom AND table give the same hits with or without it).  Sized
  // so that a warp rarely holds any passing lane (<= 0.4 % fill up to 2^24 targets): then the two XXH64 + the bloom probes
  // (~170 instructions per record, more for the warp-divergent tail of a half-full filter) are skipped for the whole
  // warp.  Always present: "prefilter" = 0 uploads an all-ones bitmap.
  uint32_t pre_k;
  const uint32_t *pre;
  // message schedules of the second SHA-256 block of the uncompressed key (hash.cuh KH_SHA_UNC2_TAB): KH_SHA2TAB_WORDS words in
  // global memory; the scan kernels that hash uncompressed keys copy them into shared memory
  const uint32_t *sha2;
};
KH_HD bool prefilter_pass(const ScanTargets &tg, uint32_t first_word_be) {
  const uint32_t idx = first_word_be >> (32 - tg.pre_k);
  return (kh_ld_u32(tg.pre + (idx >> 5)) >> (idx & 31)) & 1u;
}

// vanityrmdmatch (keyhunt.cpp:6677-6703): the reference pre-filters with a bloom over the first
// vanity_rmd_minimun_bytes_check_length bytes of every lower limit; every value inside an interval shares those
// bytes with its lower limit, so the filter never changes the answer and the interval test alone is exact.
KH_HD bool vanity_match(const uint32_t *van, uint32_t n, const uint32_t w_le[5]) {
  uint32_t d[5];
#pragma unroll
  for (int i = 0; i < 5; i++) d[i] = bswap32(w_le[i]);
  const uint32_t p16 = d[0] >> 16;
  if (!((kh_ld_u32(van + (p16 >> 5)) >> (p16 & 31)) & 1u)) return false;
  const uint32_t *iv = van + 2048;
  for (uint32_t i = 0; i < n; i++, iv += 10)
    if (cmp20(iv, d) >= 0 && cmp20(iv + 5, d) <= 0) return true;     // A <= d && d <= B
  return false;
}

// endomorphism constants exactly as the reference's -e sets them (keyhunt.cpp:930-931), little-endian limbs
#define KH_BETA  {0x719501EEu, 0xC1396C28u, 0x12F58995u, 0x9CF04975u, 0xAC3434E9u, 0x6E64479Eu, 0x657C0710u, 0x7AE96A2Bu}
#define KH_BETA2 {0x8E6AFA40u, 0x3EC693D6u, 0xED0A766Au, 0x630FB68Au, 0x53CBCB16u, 0x919BB861u, 0x9A83F8EFu, 0x851695D4u}

// VANITY: the membership test is the interval match of -m vanity instead of bloom + table; a compile-time switch, because a
// run-time branch in probe() costs the C2 kernel 1 % (A/B: 2,272 -> 2,250 Mkeys/s)
template <int KIND, bool ENDO = false, bool VANITY = false>
struct ScanEmit {
  static constexpr bool NEED_Y = (KIND == KH_SCAN_UNCOMP || KIND == KH_SCAN_BOTH || KIND == KH_SCAN_ETH);  // keyhunt.cpp:3294
#ifndef KH_OUTLINE_MUL
#define KH_OUTLINE_MUL 1
#endif
  static constexpr bool OUTLINE_MUL = KH_OUTLINE_MUL && (KIND != KH_SCAN_XPOINT || ENDO);
  // the rare-branch final reduction of the multiplier (fe.cuh KH_RARE_REDUCE) pays everywhere (x-only walk +4.2 %, giant +2.6 %, compress /
  // uncompress / ETH +1 %) except in the C2 kernel, which loses 0.5 % to it (A/B profiles/r02_ab_rare_reduce.txt): that kernel keeps the
  // straight-line form
#ifndef KH_RARE_REDUCE_BOTH
#define KH_RARE_REDUCE_BOTH 0
#endif
  static constexpr int RARE_REDUCE = (KIND == KH_SCAN_BOTH) ? (KH_RARE_REDUCE && KH_RARE_REDUCE_BOTH) : KH_RARE_REDUCE;
  // squarings (the 255 of the inversion, one per point in the walk) through a dedicated squaring (fe.cuh fe_sqr_ol / fe_sqr_sel) instead of
  // the shared multiplier: x-only walk +1.2 %, giant +0.9 %, compress +0.9 %, uncompress +1.3 %, ETH +0.8 %; the C2 kernel loses 0.3 % to the
  // second out-of-line copy and keeps the single multiplier (A/B profiles/r02_ab_inv_sqr.txt, r02_ab_hash_sqr_keccak_peel.txt)
#ifndef KH_INV_SQR
#define KH_INV_SQR 1
#endif
#ifndef KH_INV_SQR_HASH
#define KH_INV_SQR_HASH 1
#endif
#ifndef KH_INV_SQR_BOTH
#define KH_INV_SQR_BOTH 0
#endif
  static constexpr bool INV_SQR = !OUTLINE_MUL ? (KH_INV_SQR != 0) : ((KIND == KH_SCAN_BOTH) ? (KH_INV_SQR_BOTH != 0) : (KH_INV_SQR_HASH != 0));
  static constexpr bool PAIRS = (KIND == KH_SCAN_XPOINT) && !ENDO;
  static constexpr bool SHA2TAB = KH_SHA_UNC2_TAB && (KIND == KH_SCAN_UNCOMP || KIND == KH_SCAN_BOTH);   // kernels that stage the table
  const ScanTargets &tg;
  const uint32_t *sha2;      // the schedule table where the hash jobs read it (shared memory on the device), nullptr = compute
  KH_HDM explicit ScanEmit(const ScanTargets &t, const uint32_t *sha2tab = nullptr) : tg(t), sha2(sha2tab) {}

  KH_HDM void probe(const uint32_t h[5], uint32_t kind, uint64_t batch, uint32_t idx, uint32_t variant = 0) {
    if (VANITY) {                                      // -m vanity (keyhunt.cpp:4129, :4192, :4259)
      if (vanity_match(tg.van, tg.van_n, h)) sink_push(tg.sink, batch, idx, kind, h, variant);
      return;
    }
    if (!prefilter_pass(tg, bswap32(h[0]))) return;
    if (bloom_check20(tg.bloom, h)) {                  // keyhunt.cpp:3621
      if (table_contains(tg.table, tg.n, h))           // keyhunt.cpp:3623
        sink_push(tg.sink, batch, idx, kind, h, variant);
    }
  }
  // -e (FLAGENDOMORPHISM): the candidates x, beta*x, beta^2*x of one point (keyhunt.cpp:3408-3473) through the same
  // hash/probe code; `variant` is the reference's candidate index l that selects the fix-up on the host
  // (:3557-3617 compress, :3643-3686 uncompress, :3704-3749 ETH, :3769-3807 xpoint).
  KH_HDM void point_endo(const fe &x, const fe &y, uint64_t batch, uint32_t idx) {
    uint32_t h[5];
    const fe b1 = {KH_BETA}, b2 = {KH_BETA2};
#pragma unroll 1
    for (int v = 0; v < 3; v++) {
      fe xv = x;
      if (v == 1) fe_mul_sel<OUTLINE_MUL, RARE_REDUCE>(xv, x, b1);
      if (v == 2) fe_mul_sel<OUTLINE_MUL, RARE_REDUCE>(xv, x, b2);
      if (KIND == KH_SCAN_XPOINT) {
#pragma unroll
        for (int i = 0; i < 5; i++) h[i] = bswap32(xv.v[7 - i]);
        probe(h, KH_KIND_XPOINT, batch, idx, (uint32_t)v);
      }
      if (KIND == KH_SCAN_ETH) {
        // slot 2v: (xv, y) — except slot 4, where the reference hashes the BETA point again (keyhunt.cpp:3534);
        // slot 2v+1: (xv, -y)
        fe xe = xv;
        if (v == 2) fe_mul_sel<OUTLINE_MUL, RARE_REDUCE>(xe, x, b1);
        eth_address(h, xe, y);
        probe(h, KH_KIND_ETH, batch, idx, (uint32_t)(2 * v));
        fe ny;                                           // negated on the spot: one fe less alive across the candidate loop
        fe_neg(ny, y);
        eth_address(h, xv, ny);
        probe(h, KH_KIND_ETH, batch, idx, (uint32_t)(2 * v + 1));
      }
      if (KIND == KH_SCAN_COMP || KIND == KH_SCAN_UNCOMP || KIND == KH_SCAN_BOTH) {
        // one loop, one call site (= one copy of the hash code in the hot loop): jobs 0,1 = prefixes 02,03 of xv,
        // jobs 2,3 = uncompressed (xv, y) and (xv, -y)
        const int j0 = (KIND == KH_SCAN_UNCOMP) ? 2 : 0, j1 = (KIND == KH_SCAN_COMP) ? 2 : 4;
#pragma unroll 1
        for (int j = j0; j < j1; j++) {
          fe yy = y;
          if (NEED_Y && j == 3) fe_neg(yy, y);
          hash160_job<NEED_Y>(h, j < 2 ? j : 2, xv, yy, sha2);
          probe(h, j < 2 ? (uint32_t)j : (uint32_t)KH_KIND_UNCOMP, batch, idx, j < 2 ? (uint32_t)(2 * v + j) : (uint32_t)(6 + 2 * v + (j - 2)));
        }
      }
    }
  }
  // xpoint: both points of a +-e pair, bloom probes of the two overlapped (keyhunt.cpp:3810-3821 twice)
  KH_HDM void pair(const fe &xa, uint32_t ia, const fe &xb, uint32_t ib, uint64_t batch) {
    uint32_t ha[5], hb[5];
#pragma unroll
    for (int i = 0; i < 5; i++) { ha[i] = bswap32(xa.v[7 - i]); hb[i] = bswap32(xb.v[7 - i]); }
    bool okA = prefilter_pass(tg, xa.v[7]), okB = prefilter_pass(tg, xb.v[7]);
    if (!(okA || okB)) return;
    const uint64_t aA = xxh64_20(ha, KH_BLOOM_SEED), aB = xxh64_20(hb, KH_BLOOM_SEED);
    const uint64_t bA = xxh64_20(ha, aA), bB = xxh64_20(hb, aB);
    bloom_test_pair(tg.bloom, 0, aA, bA, 0, aB, bB, okA, okB);
    if (okA && table_contains(tg.table, tg.n, ha)) sink_push(tg.sink, batch, ia, KH_KIND_XPOINT, ha);
    if (okB && table_contains(tg.table, tg.n, hb)) sink_push(tg.sink, batch, ib, KH_KIND_XPOINT, hb);
  }
  KH_HDM void point(const fe &x, const fe &y, uint64_t batch, uint32_t idx) {
    if (ENDO) { point_endo(x, y, batch, idx); return; }
    uint32_t h[5];
    if (KIND == KH_SCAN_XPOINT) {                      // keyhunt.cpp:3810-3821 (first 20 bytes of X)
#pragma unroll
      for (int i = 0; i < 5; i++) h[i] = bswap32(x.v[7 - i]);
      probe(h, KH_KIND_XPOINT, batch, idx);
    }
    if (KIND == KH_SCAN_COMP || KIND == KH_SCAN_UNCOMP || KIND == KH_SCAN_BOTH) {
      // jobs 0,1 = prefixes 02,03 (hashed for EVERY X, keyhunt.cpp:3493-3494), job 2 = uncompressed (:3519);
      // one static copy of the hash + probe code, looped
      const int j0 = (KIND == KH_SCAN_UNCOMP) ? 2 : 0, j1 = (KIND == KH_SCAN_COMP) ? 2 : 3;
#pragma unroll 1
      for (int job = j0; job < j1; job++) {
        hash160_job<NEED_Y>(h, job, x, y, sha2);
        probe(h, (uint32_t)job, batch, idx);   // KH_KIND_COMP02 = 0, COMP03 = 1, UNCOMP = 2
      }
    }
    if (KIND == KH_SCAN_ETH) {                           // keyhunt.cpp:3540
      eth_address(h, x, y);
      probe(h, KH_KIND_ETH, batch, idx);
    }
  }
};

// ---- BSGS ------------------------------------------------------------------------------------------
struct BpEntry {          // == struct bsgs_xvalue (keyhunt.cpp:132): 6 key bytes, 2 pad, u64 index
  uint8_t value[6];
  uint8_t pad[2];
  uint64_t index;
};

struct BsgsTables {
  BloomDev tier[3];       // bloom_bP, bloom_bPx2nd, bloom_bPx3rd: 256 shards each, shard = X[0]
  BpEntry *table;         // m3 entries
  uint64_t m, m2, m3;
  // exact prefix bitmap over the baby points' X (same idea as ScanTargets::pre): 2^pre_k bits indexed by the first pre_k
  // bits of X, set for every baby point; a giant step whose bit is clear cannot be a baby point, so its tier-1 probes
  // (~4 random 64-byte HBM atoms) are replaced by this one.  pre_k = 0: off.
  uint32_t *pre;
  uint32_t pre_k;
  uint32_t pad;
};
// One random 4-byte read of the baby-point prefix bitmap (64 GB at -k 512) per giant step.  ncu (profiles/r02_giant_*): L1 asks
// L2 for ONE 32-byte sector per probe, but L2 looks up and fetches all FOUR sectors of the 128-byte line from DRAM
// (lts__t_sectors_srcunit_tex_op_read = 4 x lts__t_requests; dram__sectors_read = the same) — the 145 B per giant step are L2
// sector promotion, not page-table reads.  With the .L2::64B qualifier L2 fetches two sectors: 81 B per giant step instead of
// 145 B (21.9 GB instead of 39.0 GB per 2^28 steps) — and the kernel takes exactly as long (13.74 vs 13.78 ms): it is bound by
// the EC arithmetic, not by this traffic.  The smaller fetch is kept for the HBM power it saves.  KH_PRE_LD selects the flavour:
//   0 = ld.global.nc (LDG.CONSTANT)   1 = ld.global.nc.L2::64B   2 = ld.global.L2::64B   3 = ld.global.cv
//   4 = ld.global.cg                  5 = atom.global.or.b32 with 0 (L2 atomic unit works on 32-byte sectors)
//   6 = ld.global.nc.L1::no_allocate.L2::64B
#ifndef KH_PRE_LD
#define KH_PRE_LD 1
#endif
KH_HD uint32_t kh_ld_probe32(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
  uint32_t v;
#if KH_PRE_LD == 0
  v = __ldg(p);
#elif KH_PRE_LD == 1
  asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
#elif KH_PRE_LD == 2
  asm volatile("ld.global.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
#elif KH_PRE_LD == 3
  asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
#elif KH_PRE_LD == 4
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
#elif KH_PRE_LD == 5
  asm volatile("atom.global.or.b32 %0, [%1], 0;" : "=r"(v) : "l"(p) : "memory");
#else
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
#endif
  return v;
#else
  return *p;
#endif
}
KH_HD uint64_t bsgs_pre_index(const fe &x, uint32_t k) { return (((uint64_t)x.v[7] << 32) | x.v[6]) >> (64 - k); }
KH_HD bool bsgs_pre_test(const uint32_t *pre, uint32_t k, const fe &x) {
  const uint64_t idx = bsgs_pre_index(x, k);
  return (kh_ld_probe32(pre + (idx >> 5)) >> (uint32_t)(idx & 31)) & 1u;
}

// Binned table build.  A baby point sets 20 bits of ONE tier-1 shard (shard = first byte of X, ~30 MB at -k 512) and one bit of
// the prefix bitmap (indexed by the first pre_k bits of X).  Done directly that is 21 random read-modify-writes over 72 GB per
// point (0.99 G baby steps/s, bound by DRAM transactions).  Binned: the walk only appends (a, b, low index bits) of each
// point to the bucket of the first KH_BABY_BUCKET_BITS bits of its X; a second kernel then applies the buckets in order, so
// that at any moment the SMs work on a few consecutive buckets = ONE bloom shard (L2-resident) and a few 16 MB bitmap regions.
// The OR-ed image is the same whatever the order (byte-identical files, tests/test_gpu_bsgs.py).
#define KH_BABY_BUCKET_BITS 12
struct BabyBins {
  uint64_t *a, *b;        // XXH64(X, seed), XXH64(X, a)
  uint32_t *lo;           // bitmap index below the bucket bits
  uint32_t *count;        // records per bucket (may exceed cap: the excess took the direct path)
  uint32_t cap;           // records a bucket holds; 0 = binning off (everything direct)
  uint32_t pad;
};

// baby steps: point p = batch*1024 + idx is (p+1)*G            (thread_bPload keyhunt.cpp:5394-5443)
struct BabyEmit {
  static constexpr bool NEED_Y = false;
  static constexpr bool OUTLINE_MUL = false;
  static constexpr int RARE_REDUCE = KH_RARE_REDUCE;
  static constexpr bool INV_SQR = KH_INV_SQR != 0;
  static constexpr bool PAIRS = true;
  KH_HDM void pair(const fe &xa, uint32_t ia, const fe &xb, uint32_t ib, uint64_t batch) {
    const fe dummy = xa;
    point(xa, dummy, batch, ia);
    point(xb, dummy, batch, ib);
  }
  const BsgsTables &bt;
  BabyBins bins;
  KH_HDM explicit BabyEmit(const BsgsTables &b) : bt(b) { bins.a = bins.b = nullptr; bins.lo = nullptr; bins.count = nullptr; bins.cap = 0; bins.pad = 0; }
  KH_HDM BabyEmit(const BsgsTables &b, const BabyBins &bn) : bt(b), bins(bn) {}
  KH_HDM void point(const fe &x, const fe &, uint64_t batch, uint32_t idx) {
    const uint64_t p = batch * KH_GRP + idx;
    if (p >= bt.m) return;
    // binned: claim the record slot first — the counter's round trip to L2 is hidden behind the two XXH64 below
    const uint32_t bucket = x.v[7] >> (32 - KH_BABY_BUCKET_BITS);
    uint32_t slot = 0xFFFFFFFFu;
    if (bins.cap) slot = kh_atomic_inc(bins.count + bucket);
    uint32_t w[8];
    fe_to_le_words(w, x);
    const uint64_t a = xxh64_32(w, KH_BLOOM_SEED);
    const uint64_t b = xxh64_32(w, a);
    const uint32_t shard = x.v[7] >> 24;               // first byte of the big-endian X
    if (p < bt.m3) {
      BpEntry en;
      // X bytes 16..21 = limb 3 (bytes 16..19) and the top half of limb 2 (bytes 20,21)
      en.value[0] = (uint8_t)(x.v[3] >> 24); en.value[1] = (uint8_t)(x.v[3] >> 16);
      en.value[2] = (uint8_t)(x.v[3] >> 8);  en.value[3] = (uint8_t)(x.v[3]);
      en.value[4] = (uint8_t)(x.v[2] >> 24); en.value[5] = (uint8_t)(x.v[2] >> 16);
      en.pad[0] = 0; en.pad[1] = 0;
      en.index = p;
      bt.table[p] = en;
      bloom_set(bt.tier[2], shard, a, b);
    }
    if (p < bt.m2) bloom_set(bt.tier[1], shard, a, b);
    if (slot < bins.cap) {                                // binned: append, the apply kernel sets the bits (cap = 0: never)
      const uint64_t at = (uint64_t)bucket * bins.cap + slot;
      bins.a[at] = a; bins.b[at] = b;
      bins.lo[at] = bt.pre_k ? (uint32_t)(bsgs_pre_index(x, bt.pre_k) & ((1ull << (bt.pre_k - KH_BABY_BUCKET_BITS)) - 1)) : 0u;
      return;
    }
    bloom_set(bt.tier[0], shard, a, b);
    if (bt.pre_k) {
      const uint64_t idx = bsgs_pre_index(x, bt.pre_k);
#ifdef __CUDA_ARCH__
      atomicOr(bt.pre + (idx >> 5), 1u << (uint32_t)(idx & 31));
#else
      bt.pre[idx >> 5] |= 1u << (uint32_t)(idx & 31);
#endif
    }
  }
};

// giant steps: tier-1 probe; positives go to a candidate queue     (keyhunt.cpp:4819-4823)
struct GiantCand {
  uint64_t batch;
  uint32_t idx;
  uint32_t pad;
};
struct GiantParams {
  BloomDev tier1;
  GiantCand *cands;
  uint32_t *count;
  uint32_t cap;
  uint32_t pre_k;          // BsgsTables::pre / pre_k (0 = off)
  uint64_t n_steps;        // giant steps >= n_steps are outside the reference's walk and are skipped
  const uint32_t *pre;
};
#ifndef KH_GIANT_OUTLINE
#define KH_GIANT_OUTLINE 0
#endif
struct GiantEmit {
  static constexpr bool NEED_Y = false;
  static constexpr bool OUTLINE_MUL = KH_GIANT_OUTLINE != 0;
  static constexpr int RARE_REDUCE = KH_RARE_REDUCE;
  static constexpr bool INV_SQR = KH_INV_SQR != 0;
  static constexpr bool PAIRS = true;
  KH_HDM void push(uint64_t batch, uint32_t idx) {
    uint32_t slot = kh_atomic_inc(gp.count);
    if (slot < gp.cap) { GiantCand c; c.batch = batch; c.idx = idx; c.pad = 0; gp.cands[slot] = c; }
  }
  // two giant steps at once: 2 x 2 XXH64 chains interleaved, tier-1 probes of both in flight together
  KH_HDM void pair(const fe &xa, uint32_t ia, const fe &xb, uint32_t ib, uint64_t batch) {
    bool okA = (batch * KH_GRP + ia) < gp.n_steps, okB = (batch * KH_GRP + ib) < gp.n_steps;
    if (gp.pre_k) {
      okA = okA && bsgs_pre_test(gp.pre, gp.pre_k, xa);
      okB = okB && bsgs_pre_test(gp.pre, gp.pre_k, xb);
      if (!(okA || okB)) return;
    }
    uint32_t wa[8], wb[8];
    fe_to_le_words(wa, xa);
    fe_to_le_words(wb, xb);
    const uint64_t aA = xxh64_32(wa, KH_BLOOM_SEED), aB = xxh64_32(wb, KH_BLOOM_SEED);
    const uint64_t bA = xxh64_32(wa, aA), bB = xxh64_32(wb, aB);
    bloom_test_pair(gp.tier1, xa.v[7] >> 24, aA, bA, xb.v[7] >> 24, aB, bB, okA, okB);
    if (okA) push(batch, ia);
    if (okB) push(batch, ib);
  }
  const GiantParams &gp;
  KH_HDM explicit GiantEmit(const GiantParams &g) : gp(g) {}
  KH_HDM void point(const fe &x, const fe &, uint64_t batch, uint32_t idx) {
    if (batch * KH_GRP + idx >= gp.n_steps) return;
    if (gp.pre_k && !bsgs_pre_test(gp.pre, gp.pre_k, x)) return;
    uint32_t w[8];
    fe_to_le_words(w, x);
    const uint64_t a = xxh64_32(w, KH_BLOOM_SEED);
    const uint64_t b = xxh64_32(w, a);
    if (bloom_test(gp.tier1, x.v[7] >> 24, a, b)) {
      uint32_t slot = kh_atomic_inc(gp.count);
      if (slot < gp.cap) { GiantCand c; c.batch = batch; c.idx = idx; c.pad = 0; gp.cands[slot] = c; }
    }
  }
};

}  // namespace kh
