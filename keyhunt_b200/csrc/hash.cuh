// hash.cuh — per-thread SHA-256, RIPEMD-160, Keccak-256 and XXH64 for sm_100a.
//
// Replaces (reference file:line):
//   sha256sse_1B / sha256sse_2B / _sha256sse::Transform   hash/sha256_sse.cpp:426, :491, :95
//   ripemd160sse_32 / ripemd160sse::Transform              hash/ripemd160_sse.cpp:323, :92
//   KEYBUFFPREFIX / KEYBUFFUNCOMP message packers           secp256k1/SECP256K1.cpp:1187, :992
//   generate_binaddress_eth -> KECCAK_256 -> keccakf1600    keyhunt.cpp:5663, sha3/sha3.c:414, sha3/keccak.c:144
//   XXH64 (len 20 and 32 only)                              xxhash/xxhash.h:2512
// The reference runs 4 keys per SSE register; here every thread hashes its own key and the warp is
// the SIMD unit.  All message words are built straight from the field limbs (no byte buffer).
//
// Digest conventions: hash160 / ETH-address results are returned as five LITTLE-ENDIAN 32-bit words
// w[0..4] (bytes 4k..4k+3 of the 20-byte string = LE(w[k])), the layout XXH64 consumes directly.
#pragma once
#include <stdint.h>

#include "fe.cuh"

namespace kh {

KH_HD uint32_t rotr32(uint32_t x, int n) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(x, x, n);
#else
  return (x >> n) | (x << (32 - n));
#endif
}
KH_HD uint32_t rotl32(uint32_t x, int n) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l(x, x, n);
#else
  return (x << n) | (x >> (32 - n));
#endif
}
// (hi:lo) >> 8, low word:  (hi << 24) | (lo >> 8)
KH_HD uint32_t shr8_pair(uint32_t hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, 8);
#else
  return (hi << 24) | (lo >> 8);
#endif
}
KH_HD uint32_t bswap32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __byte_perm(x, 0, 0x0123);
#else
  return (x >> 24) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24);
#endif
}
KH_HD uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }


// ---- pipe balancing ---------------------------------------------------------------------------------
// SHA-256 / RIPEMD-160 are pure ALU-pipe code (SHF, LOP3, IADD3) and the hash-heavy scan kernels run the
// ALU pipe at 82-87 % of peak while the FMA-heavy pipe is ~25 % busy (ncu, profiles/r01_v3_*).  kh_addf(a, b)
// computes a + b as IMAD a*1+b with the 1 taken from constant memory (so ptxas cannot fold it back into an
// IADD), which moves the addition onto the FMA-heavy pipe.  It is applied to the SHA-256 rounds and schedule
// only: measured on B200 (kh_hash_peak) SHA-256 alone gains 15 % (13.9 -> 16.0 G compressions/s), whereas
// RIPEMD-160 (more adds than logic ops) LOSES 37 % when its adds are moved, so it keeps plain adds.
// Fused kernels: comp +3 %, uncomp +6 %, both +4.5 %.
#ifndef KH_FMA_ADDS
#define KH_FMA_ADDS 1
#endif
#if defined(__CUDACC__)
static __constant__ uint32_t kh_c_one = 1u;
#endif
KH_HD uint32_t kh_addf(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__) && KH_FMA_ADDS
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(kh_c_one), "r"(b));
  return r;
#else
  return a + b;
#endif
}

// KH_SHA_SIGMA_WIDE (A/B knob): rotations as wide multiplies on the FMA-heavy pipe.  x * 2^(32-n) is a 64-bit
// value whose halves hold the two parts of rotr(x, n) in disjoint bits, so rotr = hi ^ lo and x >> n = hi; the XOR of the
// halves folds into 3-input LOP3s.  Per function: Sigma 3 SHF + 1 LOP3 -> 3 IMAD.WIDE + 3 LOP3 (-1 ALU op), sigma
// 3 SHF + 1 LOP3 -> 3 IMAD.WIDE + 2 LOP3 (-2 ALU ops).  1 = the message-schedule sigmas only, 2 = Sigmas as well.
#ifndef KH_SHA_SIGMA_WIDE
#define KH_SHA_SIGMA_WIDE 0
#endif
#if defined(__CUDA_ARCH__) && KH_SHA_SIGMA_WIDE
__device__ __forceinline__ void kh_mulw(uint32_t &lo, uint32_t &hi, uint32_t x, uint32_t m) {
  asm("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x), "r"(m));
}
__device__ __forceinline__ uint32_t kh_sig3w(uint32_t x, int r1, int r2, int r3) {      // rotr r1 ^ rotr r2 ^ rotr r3
  uint32_t l1, h1, l2, h2, l3, h3;
  kh_mulw(l1, h1, x, 1u << (32 - r1)); kh_mulw(l2, h2, x, 1u << (32 - r2)); kh_mulw(l3, h3, x, 1u << (32 - r3));
  return (l1 ^ h1 ^ l2) ^ (h2 ^ l3 ^ h3);
}
__device__ __forceinline__ uint32_t kh_sig2sw(uint32_t x, int r1, int r2, int s) {      // rotr r1 ^ rotr r2 ^ (x >> s)
  uint32_t l1, h1, l2, h2, l3, h3;
  kh_mulw(l1, h1, x, 1u << (32 - r1)); kh_mulw(l2, h2, x, 1u << (32 - r2)); kh_mulw(l3, h3, x, 1u << (32 - s));
  return (l1 ^ h1 ^ l2) ^ (h2 ^ h3);
}
#define KH_SHA_s0(x) kh_sig2sw(x, 7, 18, 3)
#define KH_SHA_s1(x) kh_sig2sw(x, 17, 19, 10)
#if KH_SHA_SIGMA_WIDE >= 2
#define KH_SHA_S0(x) kh_sig3w(x, 2, 13, 22)
#define KH_SHA_S1(x) kh_sig3w(x, 6, 11, 25)
#else
#define KH_SHA_S0(x) (rotr32(x, 2) ^ rotr32(x, 13) ^ rotr32(x, 22))
#define KH_SHA_S1(x) (rotr32(x, 6) ^ rotr32(x, 11) ^ rotr32(x, 25))
#endif
#else
#define KH_SHA_S0(x) (rotr32(x, 2) ^ rotr32(x, 13) ^ rotr32(x, 22))
#define KH_SHA_S1(x) (rotr32(x, 6) ^ rotr32(x, 11) ^ rotr32(x, 25))
#define KH_SHA_s0(x) (rotr32(x, 7) ^ rotr32(x, 18) ^ ((x) >> 3))
#define KH_SHA_s1(x) (rotr32(x, 17) ^ rotr32(x, 19) ^ ((x) >> 10))
#endif
#define KH_SHA_CH(x, y, z) (((x) & (y)) ^ (~(x) & (z)))
#define KH_SHA_MAJ(x, y, z) (((x) & (y)) ^ ((x) & (z)) ^ ((y) & (z)))
#define KH_SHA_RND(a, b, c, d, e, f, g, h, k, wv)                                   \
  {                                                                                 \
    uint32_t t1 = kh_addf(kh_addf(h + (k) + (wv), KH_SHA_S1(e)), KH_SHA_CH(e, f, g)); \
    uint32_t t2 = kh_addf(KH_SHA_S0(a), KH_SHA_MAJ(a, b, c));                       \
    d = kh_addf(d, t1);                                                             \
    h = kh_addf(t1, t2);                                                            \
  }

#define KH_RMD_F0(x, y, z) ((x) ^ (y) ^ (z))
#define KH_RMD_F1(x, y, z) (((x) & (y)) | (~(x) & (z)))
#define KH_RMD_F2(x, y, z) (((x) | ~(y)) ^ (z))
#define KH_RMD_F3(x, y, z) (((x) & (z)) | ((y) & ~(z)))
#define KH_RMD_F4(x, y, z) ((x) ^ ((y) | ~(z)))
#define KH_RMD_STEP(F, a, b, c, d, e, xv, k, s)              \
  {                                                          \
    a = rotl32(a + F(b, c, d) + (xv) + (k), s) + e;          \
    c = rotl32(c, 10);                                       \
  }

#define KH_SHA_K_LIST                                                                               \
  0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, \
  0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, \
  0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, \
  0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, \
  0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, \
  0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, \
  0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, \
  0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u
#if defined(__CUDACC__)
static __constant__ uint32_t kh_sha_k_dev[64] = {KH_SHA_K_LIST};
#endif
static const uint32_t kh_sha_k_host[64] = {KH_SHA_K_LIST};
KH_HD uint32_t sha_k(int i) {
#ifdef __CUDA_ARCH__
  return kh_sha_k_dev[i];
#else
  return kh_sha_k_host[i];
#endif
}

#include "hash_rounds.inc"

// one SHA-256 compression; w[] is consumed (used as the circular schedule)
// KH_SHA_ROLLED: 4 trips of 16 rounds (round constants from constant memory) instead of 64 unrolled
// rounds: 3.5x less code.  The scan kernels are instruction-fetch sensitive (ncu: stall no_instruction),
// so the hot loop must stay inside the SM instruction cache.
#ifndef KH_SHA_ROLLED
#define KH_SHA_ROLLED 1
#endif
// shape (KH_SHA_SPECIAL): 0 = any block, 1 = the 33-byte compressed-key block, 2 = the second block of the 65-byte key; the
// first schedule expansion then drops the zero words and folds sigma(constant) (gen_rounds.py emit_special)
#ifndef KH_SHA_SPECIAL
#define KH_SHA_SPECIAL 1
#endif
// The second block of the 65-byte uncompressed key holds ONE byte of data (the last byte of Y, then 0x80, zeros and the length
// 0x208): its whole 64-word message schedule is a function of that byte.  KH_SHA_UNC2_TAB = 1: the scan kernels keep the 256
// schedules in shared memory (sha_unc2_table_fill; rows padded to 68 words so that the rows of the lanes of a quarter-warp fall on
// different 16-byte bank groups) and `row` replaces the 48 schedule expansions of that block by 16 LDS.128 (~320 ALU-pipe
// instructions less per uncompressed key, on kernels that are ALU-pipe bound).  row == nullptr: the schedule is computed.
#ifndef KH_SHA_UNC2_TAB
#define KH_SHA_UNC2_TAB 1
#endif
#define KH_SHA2TAB_STRIDE 68
#define KH_SHA2TAB_WORDS (256 * KH_SHA2TAB_STRIDE)
KH_HD void sha_load16(uint32_t w[16], const uint32_t *p) {
#if defined(__CUDA_ARCH__)
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  const uint4 a = q[0], b = q[1], c = q[2], d = q[3];
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
  w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w; w[12] = d.x; w[13] = d.y; w[14] = d.z; w[15] = d.w;
#else
  for (int i = 0; i < 16; i++) w[i] = p[i];
#endif
}
KH_HD void sha256_compress(uint32_t st[8], uint32_t w[16], int shape = 0, const uint32_t *row = nullptr) {
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#if KH_SHA_ROLLED
#pragma unroll 1
  for (int kb = 0; kb < 64; kb += 16) {
    if (KH_SHA_UNC2_TAB && row) { sha_load16(w, row + kb); }
    else if (KH_SHA_SPECIAL && kb == 16 && shape == 1) { KH_SHA256_EXPAND16_COMP33(w); }
    else if (KH_SHA_SPECIAL && kb == 16 && shape == 2) { KH_SHA256_EXPAND16_UNC2(w); }
    else if (kb) { KH_SHA256_EXPAND16(w); }
    KH_SHA256_ROUNDS16(a, b, c, d, e, f, g, h, w, kb);
  }
#else
  KH_SHA256_ROUNDS(a, b, c, d, e, f, g, h, w);
#endif
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
// schedule table of the second block of the uncompressed key: row v = W[0..63] for the last Y byte v
KH_HD void sha_unc2_table_row(uint32_t *row, uint32_t v) {
  uint32_t w[64];
  w[0] = (v << 24) | 0x00800000u;
  for (int i = 1; i < 15; i++) w[i] = 0;
  w[15] = 0x208u;
  for (int i = 16; i < 64; i++) {
    const uint32_t x = w[i - 15], y = w[i - 2];
    w[i] = w[i - 16] + w[i - 7] + (rotr32(x, 7) ^ rotr32(x, 18) ^ (x >> 3)) + (rotr32(y, 17) ^ rotr32(y, 19) ^ (y >> 10));
  }
  for (int i = 0; i < 64; i++) row[i] = w[i];
  for (int i = 64; i < KH_SHA2TAB_STRIDE; i++) row[i] = 0;
}
KH_HD void sha256_init(uint32_t st[8]) {
  st[0] = 0x6a09e667u; st[1] = 0xbb67ae85u; st[2] = 0x3c6ef372u; st[3] = 0xa54ff53au;
  st[4] = 0x510e527fu; st[5] = 0x9b05688cu; st[6] = 0x1f83d9abu; st[7] = 0x5be0cd19u;
}

// RIPEMD-160 of the 32-byte SHA-256 digest held as 8 big-endian state words; out = 5 LE words
KH_HD void ripemd160_of_sha(uint32_t out[5], const uint32_t sha[8]) {
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = bswap32(sha[i]);
  x[8] = 0x00000080u;
  x[9] = 0; x[10] = 0; x[11] = 0; x[12] = 0; x[13] = 0;
  x[14] = 256u;  // bit length, little-endian 64-bit
  x[15] = 0;
  const uint32_t h0 = 0x67452301u, h1 = 0xEFCDAB89u, h2 = 0x98BADCFEu, h3 = 0x10325476u, h4 = 0xC3D2E1F0u;
  uint32_t al = h0, bl = h1, cl = h2, dl = h3, el = h4;
  uint32_t ar = h0, br = h1, cr = h2, dr = h3, er = h4;
  KH_RMD160_ROUNDS(al, bl, cl, dl, el, ar, br, cr, dr, er, x);
  out[0] = h1 + cl + dr;
  out[1] = h2 + dl + er;
  out[2] = h3 + el + ar;
  out[3] = h4 + al + br;
  out[4] = h0 + bl + cr;
}

// hash160 of the compressed public key prefix||X  (GetHash160_fromX, SECP256K1.cpp:1207)
KH_HD void hash160_compressed(uint32_t out[5], uint32_t prefix, const fe &x) {
  uint32_t w[16], st[8];
  w[0] = (prefix << 24) | (x.v[7] >> 8);
#pragma unroll
  for (int i = 1; i < 8; i++) w[i] = shr8_pair(x.v[8 - i], x.v[7 - i]);
  w[8] = (x.v[0] << 24) | 0x00800000u;
  w[9] = 0; w[10] = 0; w[11] = 0; w[12] = 0; w[13] = 0; w[14] = 0;
  w[15] = 0x108u;
  sha256_init(st);
  sha256_compress(st, w);
  ripemd160_of_sha(out, st);
}

// hash160 of the uncompressed public key 04||X||Y  (GetHash160(...,false,...), SECP256K1.cpp:1045)
KH_HD void hash160_uncompressed(uint32_t out[5], const fe &x, const fe &y) {
  uint32_t w[16], st[8];
  w[0] = 0x04000000u | (x.v[7] >> 8);
#pragma unroll
  for (int i = 1; i < 8; i++) w[i] = shr8_pair(x.v[8 - i], x.v[7 - i]);
  w[8] = shr8_pair(x.v[0], y.v[7]);
#pragma unroll
  for (int i = 1; i < 8; i++) w[8 + i] = shr8_pair(y.v[8 - i], y.v[7 - i]);
  sha256_init(st);
  sha256_compress(st, w);
  w[0] = (y.v[0] << 24) | 0x00800000u;
#pragma unroll
  for (int i = 1; i < 15; i++) w[i] = 0;
  w[15] = 0x208u;
  sha256_compress(st, w);
  ripemd160_of_sha(out, st);
}

// One hash160 "job" with a single static copy of the SHA-256 and RIPEMD-160 bodies, whatever the job:
// job 0 / 1 = compressed key with prefix 02 / 03 (one block), job 2 = uncompressed key 04||X||Y (two
// blocks; only when WITH_UNCOMP).  The scan kernel loops over jobs with `#pragma unroll 1`, so its hot
// loop holds each hash body once (instruction-cache footprint, see sha256_compress).
template <bool WITH_UNCOMP>
KH_HD void hash160_job(uint32_t out[5], int job, const fe &x, const fe &y, const uint32_t *sha2tab = nullptr) {
  uint32_t w[16], st[8];
  sha256_init(st);
  const bool unc = WITH_UNCOMP && (job == 2);
  const int nblk = unc ? 2 : 1;
#pragma unroll 1
  for (int blk = 0; blk < nblk; blk++) {
    if (blk == 0) {
      const uint32_t pre = unc ? 4u : (2u + (uint32_t)job);
      w[0] = (pre << 24) | (x.v[7] >> 8);
#pragma unroll
      for (int i = 1; i < 8; i++) w[i] = shr8_pair(x.v[8 - i], x.v[7 - i]);
      if (unc) {
        w[8] = shr8_pair(x.v[0], y.v[7]);
#pragma unroll
        for (int i = 1; i < 8; i++) w[8 + i] = shr8_pair(y.v[8 - i], y.v[7 - i]);
      } else {
        w[8] = (x.v[0] << 24) | 0x00800000u;
        w[9] = 0; w[10] = 0; w[11] = 0; w[12] = 0; w[13] = 0; w[14] = 0;
        w[15] = 0x108u;
      }
    } else if (!(KH_SHA_UNC2_TAB && sha2tab)) {
      w[0] = (y.v[0] << 24) | 0x00800000u;
#pragma unroll
      for (int i = 1; i < 15; i++) w[i] = 0;
      w[15] = 0x208u;
    }
    sha256_compress(st, w, blk ? 2 : (unc ? 0 : 1), (WITH_UNCOMP && blk && sha2tab) ? sha2tab + (y.v[0] & 0xFFu) * KH_SHA2TAB_STRIDE : nullptr);
  }
  ripemd160_of_sha(out, st);
}

// ---- Keccak-256 of X||Y (64 bytes), Ethereum address = bytes 12..31 ---------------------------------
#define KH_KECCAK_RC_LIST                                                                          \
  0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,      \
  0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,      \
  0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,      \
  0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,      \
  0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,      \
  0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL
#if defined(__CUDACC__)
static __constant__ uint64_t kh_keccak_rc_dev[24] = {KH_KECCAK_RC_LIST};
#endif
static const uint64_t kh_keccak_rc_host[24] = {KH_KECCAK_RC_LIST};
KH_HD uint64_t keccak_rc(int r) {
#ifdef __CUDA_ARCH__
  return kh_keccak_rc_dev[r];
#else
  return kh_keccak_rc_host[r];
#endif
}

KH_HD void keccak_round(uint64_t s[25], uint64_t rc) {
  uint64_t c0 = s[0] ^ s[5] ^ s[10] ^ s[15] ^ s[20];
  uint64_t c1 = s[1] ^ s[6] ^ s[11] ^ s[16] ^ s[21];
  uint64_t c2 = s[2] ^ s[7] ^ s[12] ^ s[17] ^ s[22];
  uint64_t c3 = s[3] ^ s[8] ^ s[13] ^ s[18] ^ s[23];
  uint64_t c4 = s[4] ^ s[9] ^ s[14] ^ s[19] ^ s[24];
  uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1),
           d4 = c3 ^ rotl64(c0, 1);
#pragma unroll
  for (int y = 0; y < 25; y += 5) { s[y] ^= d0; s[y + 1] ^= d1; s[y + 2] ^= d2; s[y + 3] ^= d3; s[y + 4] ^= d4; }
  // rho + pi
  uint64_t t = s[1], u;
#define KH_RP(j, n) u = s[j]; s[j] = rotl64(t, n); t = u;
  KH_RP(10, 1) KH_RP(7, 3) KH_RP(11, 6) KH_RP(17, 10) KH_RP(18, 15) KH_RP(3, 21) KH_RP(5, 28) KH_RP(16, 36)
  KH_RP(8, 45) KH_RP(21, 55) KH_RP(24, 2) KH_RP(4, 14) KH_RP(15, 27) KH_RP(23, 41) KH_RP(19, 56) KH_RP(13, 8)
  KH_RP(12, 25) KH_RP(2, 43) KH_RP(20, 62) KH_RP(14, 18) KH_RP(22, 39) KH_RP(9, 61) KH_RP(6, 20) KH_RP(1, 44)
#undef KH_RP
  // chi
#pragma unroll
  for (int y = 0; y < 25; y += 5) {
    uint64_t a0 = s[y], a1 = s[y + 1], a2 = s[y + 2], a3 = s[y + 3], a4 = s[y + 4];
    s[y] = a0 ^ (~a1 & a2); s[y + 1] = a1 ^ (~a2 & a3); s[y + 2] = a2 ^ (~a3 & a4);
    s[y + 3] = a3 ^ (~a4 & a0); s[y + 4] = a4 ^ (~a0 & a1);
  }
  s[0] ^= rc;
}
// KH_KECCAK_PEEL: the first and the last round outside the rolled loop.  The compiler then sees the 17 lanes that are zero or
// constant when the first round starts (the message is 64 bytes: 8 lanes, the padding 2) and that only 3 of the 25 lanes of the
// last round are output, and drops the work on them
#ifndef KH_KECCAK_PEEL
#define KH_KECCAK_PEEL 1
#endif
KH_HD void keccak_f1600(uint64_t s[25]) {
#if KH_KECCAK_PEEL
  keccak_round(s, keccak_rc(0));
#pragma unroll 1
  for (int r = 1; r < 23; r++) keccak_round(s, keccak_rc(r));
  keccak_round(s, keccak_rc(23));
#else
#pragma unroll 1
  for (int r = 0; r < 24; r++) keccak_round(s, keccak_rc(r));
#endif
}

// generate_binaddress_eth (keyhunt.cpp:5663): Keccak-256 (0x01 padding, rate 136) of X||Y
KH_HD void eth_address(uint32_t out[5], const fe &x, const fe &y) {
  uint64_t s[25];
  // lane i = little-endian u64 of message bytes 8i..8i+7; the message is big-endian X then Y
#pragma unroll
  for (int i = 0; i < 4; i++) {
    s[i] = ((uint64_t)bswap32(x.v[6 - 2 * i]) << 32) | bswap32(x.v[7 - 2 * i]);
    s[4 + i] = ((uint64_t)bswap32(y.v[6 - 2 * i]) << 32) | bswap32(y.v[7 - 2 * i]);
  }
  s[8] = 0x01ULL;
#pragma unroll
  for (int i = 9; i < 25; i++) s[i] = 0;
  s[16] = 0x8000000000000000ULL;
  keccak_f1600(s);
  out[0] = (uint32_t)(s[1] >> 32);
  out[1] = (uint32_t)s[2];
  out[2] = (uint32_t)(s[2] >> 32);
  out[3] = (uint32_t)s[3];
  out[4] = (uint32_t)(s[3] >> 32);
}

// ---- XXH64 (xxhash.h:2512) specialised for 20- and 32-byte inputs given as LE 32-bit words --------
#define KH_XP1 0x9E3779B185EBCA87ULL
#define KH_XP2 0xC2B2AE3D27D4EB4FULL
#define KH_XP3 0x165667B19E3779F9ULL
#define KH_XP4 0x85EBCA77C2B2AE63ULL
#define KH_XP5 0x27D4EB2F165667C5ULL
KH_HD uint64_t xxh_round(uint64_t acc, uint64_t in) {
  acc += in * KH_XP2;
  acc = rotl64(acc, 31);
  return acc * KH_XP1;
}
KH_HD uint64_t xxh_avalanche(uint64_t h) {
  h ^= h >> 33; h *= KH_XP2; h ^= h >> 29; h *= KH_XP3; h ^= h >> 32;
  return h;
}
KH_HD uint64_t xxh64_20(const uint32_t w[5], uint64_t seed) {
  uint64_t h = seed + KH_XP5 + 20ULL;
  uint64_t k = ((uint64_t)w[1] << 32) | w[0];
  h ^= xxh_round(0, k); h = rotl64(h, 27) * KH_XP1 + KH_XP4;
  k = ((uint64_t)w[3] << 32) | w[2];
  h ^= xxh_round(0, k); h = rotl64(h, 27) * KH_XP1 + KH_XP4;
  h ^= (uint64_t)w[4] * KH_XP1; h = rotl64(h, 23) * KH_XP2 + KH_XP3;
  return xxh_avalanche(h);
}
KH_HD uint64_t xxh_merge(uint64_t acc, uint64_t v) {
  v = xxh_round(0, v);
  acc ^= v;
  return acc * KH_XP1 + KH_XP4;
}
KH_HD uint64_t xxh64_32(const uint32_t w[8], uint64_t seed) {
  uint64_t v1 = seed + KH_XP1 + KH_XP2, v2 = seed + KH_XP2, v3 = seed, v4 = seed - KH_XP1;
  v1 = xxh_round(v1, ((uint64_t)w[1] << 32) | w[0]);
  v2 = xxh_round(v2, ((uint64_t)w[3] << 32) | w[2]);
  v3 = xxh_round(v3, ((uint64_t)w[5] << 32) | w[4]);
  v4 = xxh_round(v4, ((uint64_t)w[7] << 32) | w[6]);
  uint64_t h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
  h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
  h += 32ULL;
  return xxh_avalanche(h);
}

// the byte string Int::Get32Bytes(x) as eight LE words (word k = bytes 4k..4k+3)
KH_HD void fe_to_le_words(uint32_t w[8], const fe &x) {
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = bswap32(x.v[7 - i]);
}

}  // namespace kh
