// bloom.cuh — bit-exact device restatement of the reference's libbloom fork (bloom/bloom.cpp).
//
//   bloom_check  bloom.cpp:189-212   a = XXH64(buf,len,0x59f2815b16f81798); b = XXH64(buf,len,a);
//                                    bit_i = (a + b*i mod 2^64) % bits ; early-out on the first clear bit
//   bloom_add    bloom.cpp:122-146   same hashing, sets the bits (atomicOr here: order-free, so the
//                                    final image equals the reference's whatever the thread order)
//   test_bit     bloom.cpp:70-92     byte = bit >> 3, mask = 1 << (bit & 7)
//
// `% bits` for an arbitrary (non power-of-two) 64-bit `bits` is done exactly with a host-precomputed
// reciprocal magic = floor((2^64-1)/bits): q = mulhi64(x, magic) is floor(x/bits) or one less, so one
// conditional subtraction finishes the remainder.
#pragma once
#include <stdint.h>

#include "hash.cuh"

namespace kh {

#define KH_BLOOM_SEED 0x59f2815b16f81798ULL

struct BloomDev {
  uint8_t *bf;        // shard 0; shard s starts at bf + s*stride
  uint64_t bits;      // bits per shard (bloom->bits)
  uint64_t magic;     // floor((2^64-1)/bits)
  uint64_t stride;    // bytes between shards (bytes rounded up to a multiple of 16)
  uint32_t hashes;    // bloom->hashes
  uint32_t pad;
};

KH_HD uint64_t kh_umul64hi(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
KH_HD uint64_t bloom_mod(uint64_t x, uint64_t bits, uint64_t magic) {
  uint64_t q = kh_umul64hi(x, magic);
  uint64_t r = x - q * bits;
  if (r >= bits) r -= bits;
  return r;
}
// single-byte bloom probe.  KH_PROBE_LD selects the load flavour (A/B on the HBM-resident BSGS bloom):
//   0 = ld.global.nc (LDG.CONSTANT, L1-allocating)   1 = plain ld.global   2 = ld.global.cg (L2 only, no L1)
//   3 = ld.global.nc.L1::no_allocate
#ifndef KH_PROBE_LD
#define KH_PROBE_LD 0
#endif
KH_HD uint8_t kh_ld_u8(const uint8_t *p) {
#ifdef __CUDA_ARCH__
#if KH_PROBE_LD == 0
  return __ldg(p);
#elif KH_PROBE_LD == 1
  return *p;
#elif KH_PROBE_LD == 2
  uint32_t v;
  asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return (uint8_t)v;
#else
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return (uint8_t)v;
#endif
#else
  return *p;
#endif
}

// membership test with hashes (a, b) already computed
KH_HD bool bloom_test(const BloomDev &bl, uint32_t shard, uint64_t a, uint64_t b) {
  const uint8_t *bf = bl.bf + (uint64_t)shard * bl.stride;
  uint64_t x = a;
#pragma unroll 1
  for (uint32_t i = 0; i < bl.hashes; i++) {
    uint64_t r = bloom_mod(x, bl.bits, bl.magic);
    if (!((kh_ld_u8(bf + (r >> 3)) >> (r & 7)) & 1)) return false;
    x += b;  // (a + b*i) in wrapping u64
  }
  return true;
}
KH_HD void bloom_set(const BloomDev &bl, uint32_t shard, uint64_t a, uint64_t b) {
  uint8_t *bf = bl.bf + (uint64_t)shard * bl.stride;
  uint64_t x = a;
#pragma unroll 1
  for (uint32_t i = 0; i < bl.hashes; i++) {
    uint64_t r = bloom_mod(x, bl.bits, bl.magic);
    uint64_t byte = r >> 3;
#ifdef __CUDA_ARCH__
    // shard bases are 16-byte aligned, so the containing 32-bit word is addressable
    uint32_t *word = reinterpret_cast<uint32_t *>(bf + (byte & ~3ULL));
    atomicOr(word, 1u << (8u * (uint32_t)(byte & 3) + (uint32_t)(r & 7)));
#else
    bf[byte] |= (uint8_t)(1u << (r & 7));
#endif
    x += b;
  }
}

// Two membership tests at once.  The first KH_BLOOM_PAR bit positions of BOTH keys are loaded before any
// of them is examined (2*KH_BLOOM_PAR independent loads in flight instead of one dependent load at a time);
// the rare survivors finish in a joint loop.  The result is the same AND over all `hashes` bits as
// bloom_check's early-exit loop (bloom.cpp:189-212), only the evaluation order differs.
#ifndef KH_BLOOM_PAR
#define KH_BLOOM_PAR 2
#endif
KH_HD void bloom_test_pair(const BloomDev &bl, uint32_t shardA, uint64_t aA, uint64_t bA, uint32_t shardB, uint64_t aB,
                           uint64_t bB, bool &okA, bool &okB) {
  const uint8_t *bfA = bl.bf + (uint64_t)shardA * bl.stride;
  const uint8_t *bfB = bl.bf + (uint64_t)shardB * bl.stride;
  uint64_t xA = aA, xB = aB;
  uint8_t vA[KH_BLOOM_PAR], vB[KH_BLOOM_PAR];
  uint32_t sA[KH_BLOOM_PAR], sB[KH_BLOOM_PAR];
#pragma unroll
  for (int k = 0; k < KH_BLOOM_PAR; k++) {
    const uint64_t rA = bloom_mod(xA, bl.bits, bl.magic), rB = bloom_mod(xB, bl.bits, bl.magic);
    vA[k] = okA ? kh_ld_u8(bfA + (rA >> 3)) : (uint8_t)0;
    vB[k] = okB ? kh_ld_u8(bfB + (rB >> 3)) : (uint8_t)0;
    sA[k] = (uint32_t)(rA & 7); sB[k] = (uint32_t)(rB & 7);
    xA += bA; xB += bB;
  }
#pragma unroll
  for (int k = 0; k < KH_BLOOM_PAR; k++) { okA = okA && ((vA[k] >> sA[k]) & 1); okB = okB && ((vB[k] >> sB[k]) & 1); }
#pragma unroll 1
  for (uint32_t i = KH_BLOOM_PAR; i < bl.hashes && (okA || okB); i++) {
    if (okA) { const uint64_t r = bloom_mod(xA, bl.bits, bl.magic); okA = (kh_ld_u8(bfA + (r >> 3)) >> (r & 7)) & 1; xA += bA; }
    if (okB) { const uint64_t r = bloom_mod(xB, bl.bits, bl.magic); okB = (kh_ld_u8(bfB + (r >> 3)) >> (r & 7)) & 1; xB += bB; }
  }
}

KH_HD bool bloom_check20(const BloomDev &bl, const uint32_t w[5]) {
  uint64_t a = xxh64_20(w, KH_BLOOM_SEED);
  uint64_t b = xxh64_20(w, a);
  return bloom_test(bl, 0, a, b);
}
KH_HD void bloom_add20(const BloomDev &bl, const uint32_t w[5]) {
  uint64_t a = xxh64_20(w, KH_BLOOM_SEED);
  uint64_t b = xxh64_20(w, a);
  bloom_set(bl, 0, a, b);
}

// ---- sorted 20-byte table (struct address_value keyhunt.cpp:247; searchbinary :3065) -----------------
// Device layout: N records of five BIG-endian-packed words, so numeric word order == memcmp order.
// Any correct binary search returns the same found/not-found answer as the reference's.
KH_HD uint32_t kh_ld_u32(const uint32_t *p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
KH_HD int cmp20(const uint32_t *rec, const uint32_t key_be[5]) {  // sign of (key - rec)
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint32_t r = kh_ld_u32(rec + i);
    if (key_be[i] < r) return -1;
    if (key_be[i] > r) return 1;
  }
  return 0;
}
KH_HD bool table_contains(const uint32_t *table, uint64_t n, const uint32_t w_le[5]) {
  uint32_t key[5];
#pragma unroll
  for (int i = 0; i < 5; i++) key[i] = bswap32(w_le[i]);
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    uint64_t mid = lo + ((hi - lo) >> 1);
    int c = cmp20(table + 5 * mid, key);
    if (c == 0) return true;
    if (c < 0) hi = mid; else lo = mid + 1;
  }
  return false;
}

}  // namespace kh
