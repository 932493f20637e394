// fe.cuh — secp256k1 field arithmetic for sm_100a, 8 x 32-bit limbs in registers.
//
// Replaces the reference's 4x64-bit Int::ModMulK1 / ModSquareK1 / ModAdd / ModSub / ModNeg / ModInv
// (secp256k1/IntMod.cpp:855, :977, :41, :72, :102, :382).  Values are ALWAYS canonical (< P) on
// entry and exit, so serialised X/Y bytes equal the reference's Int::Get32Bytes output.
//
// Multiplication is operand-scanning schoolbook in the "even/odd column" arrangement: every
// 32x32->64 product is one (mad.lo.cc, madc.hi.cc) pair that lands on two ADJACENT limbs of one
// of two accumulators, so ptxas fuses the pair into a single IMAD.WIDE.U32(.X) on the FMA-heavy
// pipe and the carry chain rides on the multiplies (64 IMAD.WIDE per 256x256 product instead of
// 128 IMAD + 128 IMAD.HI).  Reduction folds the high 256 bits by 2^32+977 (P = 2^256-2^32-977).
//
// The limb-level primitives (kh_add8 / kh_sub8 / kh_mad_row) have a PTX carry-chain body for the
// device and a portable body for the host; the host body exists ONLY so tests/devsim can unit-test
// this exact limb algorithm on a machine without a GPU.  It is never part of the product path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KH_HD __host__ __device__ __forceinline__
#define KH_HDM __host__ __device__ __forceinline__   // member functions
#else
#define KH_HD static inline
#define KH_HDM inline
#endif

namespace kh {

struct fe {
  uint32_t v[8];  // little-endian limbs
};

// ---- limb primitives ---------------------------------------------------------------------------
// r = a + b, returns carry (0/1)
KH_HD uint32_t kh_add8(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
  uint32_t cf;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, 0;\n\t"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(cf)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
  cf = (uint32_t)c;
#endif
  return cf;
}
// r = a - b, returns borrow mask (0xFFFFFFFF on borrow, else 0)
KH_HD uint32_t kh_sub8(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
  uint32_t m;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;\n\t"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(m)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
  uint64_t bw = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t d = (uint64_t)a[i] - b[i] - bw;
    r[i] = (uint32_t)d;
    bw = (d >> 32) & 1;
  }
  m = bw ? 0xFFFFFFFFu : 0u;
#endif
  return m;
}
// acc[0..7] += (a0,a2,a4,a6) * b as four adjacent 64-bit lanes; the carry-out is added into acc[8]
KH_HD void kh_mad_row(uint32_t *acc, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;\n\t"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7]), "+r"(acc[8])
      : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
#else
  const uint32_t a[4] = {a0, a2, a4, a6};
  uint64_t c = 0;
  for (int k = 0; k < 4; k++) {
    uint64_t p = (uint64_t)a[k] * b;
    uint64_t t = (uint64_t)acc[2 * k] + (uint32_t)p + c;
    acc[2 * k] = (uint32_t)t;
    c = t >> 32;
    t = (uint64_t)acc[2 * k + 1] + (p >> 32) + c;
    acc[2 * k + 1] = (uint32_t)t;
    c = t >> 32;
  }
  acc[8] += (uint32_t)c;
#endif
}
// A wide multiply-add that neither takes nor produces a carry (IMAD.WIDE.U32 with RZ as addend) costs the multiplier about half of
// a link of a carry chain (IMAD.WIDE.U32.X: 13.1 against 7.25 T/s, kh_pipe_peak), and a row that lands on accumulators that are still
// zero needs no carries at all: its four products occupy four disjoint 64-bit lanes.  KH_PLAIN_HEAD = 1 writes those rows as
// plain products (8 of the 64 products of fe_mul_wide, 4 of the 8 of fe_reduce_wide, 7 of the 28 off-diagonal ones of fe_sqr_wide).
#ifndef KH_PLAIN_HEAD
#define KH_PLAIN_HEAD 1
#endif
// acc[0..2N-1] = x[k]*y[k] as N adjacent 64-bit lanes (acc was zero)
template <int N>
KH_HD void kh_mul_lanes(uint32_t *acc, const uint32_t *x, const uint32_t *y) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    const uint64_t p = (uint64_t)x[k] * y[k];
    acc[2 * k] = (uint32_t)p;
    acc[2 * k + 1] = (uint32_t)(p >> 32);
  }
}
// acc[0..2N-1] += x[k]*y[k] as N adjacent 64-bit lanes (k < N <= 4); the carry-out is added into acc[2N]
template <int N>
KH_HD void kh_mad_chain(uint32_t *acc, const uint32_t *x, const uint32_t *y) {
#ifdef __CUDA_ARCH__
  if (N == 1) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;\n\t"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]) : "r"(x[0]), "r"(y[0]));
  } else if (N == 2) {
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %8, %2;\n\tmadc.hi.cc.u32 %3, %6, %8, %3;\n\taddc.u32 %4, %4, 0;\n\t"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4])
        : "r"(x[0]), "r"(x[1]), "r"(y[0]), "r"(y[1]));
  } else if (N == 3) {
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\tmadc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %11, %2;\n\tmadc.hi.cc.u32 %3, %8, %11, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %12, %4;\n\tmadc.hi.cc.u32 %5, %9, %12, %5;\n\taddc.u32 %6, %6, 0;\n\t"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(y[0]), "r"(y[1]), "r"(y[2]));
  } else {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %14, %2;\n\tmadc.hi.cc.u32 %3, %10, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %15, %4;\n\tmadc.hi.cc.u32 %5, %11, %15, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %16, %6;\n\tmadc.hi.cc.u32 %7, %12, %16, %7;\n\taddc.u32 %8, %8, 0;\n\t"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8])
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]));
  }
#else
  uint64_t c = 0;
  for (int k = 0; k < N; k++) {
    uint64_t p = (uint64_t)x[k] * y[k];
    uint64_t t = (uint64_t)acc[2 * k] + (uint32_t)p + c;
    acc[2 * k] = (uint32_t)t; c = t >> 32;
    t = (uint64_t)acc[2 * k + 1] + (p >> 32) + c;
    acc[2 * k + 1] = (uint32_t)t; c = t >> 32;
  }
  acc[2 * N] += (uint32_t)c;
#endif
}
// r[0..15] = 2*t[0..15] + sum a[i]^2 * 2^(64 i)      (t < 2^511)
KH_HD void kh_double_add_squares(uint32_t r[16], const uint32_t t[16], const uint32_t a[8]) {
  uint32_t d[16];
  d[0] = t[0] << 1;
#pragma unroll
  for (int i = 1; i < 16; i++) d[i] = (t[i] << 1) | (t[i - 1] >> 31);
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\tmadc.hi.cc.u32 %1, %16, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %2;\n\tmadc.hi.cc.u32 %3, %17, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %4;\n\tmadc.hi.cc.u32 %5, %18, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %6;\n\tmadc.hi.cc.u32 %7, %19, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %8;\n\tmadc.hi.cc.u32 %9, %20, %20, %9;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %10;\n\tmadc.hi.cc.u32 %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %12;\n\tmadc.hi.cc.u32 %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %14;\n\tmadc.hi.u32 %15, %23, %23, %15;\n\t"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7]),
        "+r"(d[8]), "+r"(d[9]), "+r"(d[10]), "+r"(d[11]), "+r"(d[12]), "+r"(d[13]), "+r"(d[14]), "+r"(d[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t p = (uint64_t)a[i] * a[i];
    uint64_t s = (uint64_t)d[2 * i] + (uint32_t)p + c;
    d[2 * i] = (uint32_t)s; c = s >> 32;
    s = (uint64_t)d[2 * i + 1] + (p >> 32) + c;
    d[2 * i + 1] = (uint32_t)s; c = s >> 32;
  }
#endif
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = d[i];
}
// r[0..15] = e[0..15] + (o[0..14] << 32)   (final combination of the even/odd accumulators)
KH_HD void kh_combine_eo(uint32_t r[16], const uint32_t e[17], const uint32_t o[17]) {
  r[0] = e[0];
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %15, %30;\n\t"
      "addc.cc.u32 %1, %16, %31;\n\t"
      "addc.cc.u32 %2, %17, %32;\n\t"
      "addc.cc.u32 %3, %18, %33;\n\t"
      "addc.cc.u32 %4, %19, %34;\n\t"
      "addc.cc.u32 %5, %20, %35;\n\t"
      "addc.cc.u32 %6, %21, %36;\n\t"
      "addc.cc.u32 %7, %22, %37;\n\t"
      "addc.cc.u32 %8, %23, %38;\n\t"
      "addc.cc.u32 %9, %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32 %14, %29, %44;\n\t"
      : "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]),
        "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15])
      : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]),
        "r"(e[9]), "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]),
        "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]),
        "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]));
#else
  uint64_t c = 0;
  for (int i = 1; i < 16; i++) { c += (uint64_t)e[i] + o[i - 1]; r[i] = (uint32_t)c; c >>= 32; }
#endif
}
KH_HD uint32_t kh_umulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// ---- basic helpers -------------------------------------------------------------------------------
KH_HD void fe_set_zero(fe &r) {
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
}
KH_HD void fe_set_u32(fe &r, uint32_t x) {
  fe_set_zero(r);
  r.v[0] = x;
}
KH_HD bool fe_is_zero(const fe &a) {
  return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0;
}
KH_HD bool fe_eq(const fe &a, const fe &b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
  return d == 0;
}

// ---- add / sub / neg (mod P, canonical) ----------------------------------------------------------
// P = 2^256 - C, C = 2^32 + 977
KH_HD void fe_sub(fe &r, const fe &a, const fe &b) {
  uint32_t t[8];
  uint32_t m = kh_sub8(t, a.v, b.v);
  const uint32_t c[8] = {m & 977u, m & 1u, 0, 0, 0, 0, 0, 0};  // on borrow: + P  ==  - C (mod 2^256)
  kh_sub8(r.v, t, c);
}
// value = t + cf*2^256 (known < 2P)  ->  r = value mod P
KH_HD void fe_final_reduce(fe &r, const uint32_t t[8], uint32_t cf) {
  const uint32_t c[8] = {977u, 1u, 0, 0, 0, 0, 0, 0};
  uint32_t u[8];
  uint32_t k = kh_add8(u, t, c);
  bool take = (cf | k) != 0;  // value >= P  <=>  value + C >= 2^256
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = take ? u[i] : t[i];
}
KH_HD void fe_add(fe &r, const fe &a, const fe &b) {
  uint32_t t[8];
  uint32_t cf = kh_add8(t, a.v, b.v);
  fe_final_reduce(r, t, cf);
}
KH_HD void fe_neg(fe &r, const fe &a) {  // 0 -> 0
  fe z;
  fe_set_zero(z);
  fe_sub(r, z, a);
}

// ---- 256 x 256 -> 512 ------------------------------------------------------------------------------
KH_HD void fe_mul_wide(uint32_t r[16], const fe &a, const fe &b) {
  // e[k] <-> column k ; o[k] <-> column k+1
  uint32_t e[17], o[17];
#pragma unroll
  for (int i = 0; i < 17; i++) { e[i] = 0; o[i] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    // row i (even): even j -> even column -> e[i+j] ; odd j -> odd column -> o[i+j-1]
    if (KH_PLAIN_HEAD && i == 0) {   // both accumulators are still zero: eight carry-free products
      const uint32_t xe[4] = {a.v[0], a.v[2], a.v[4], a.v[6]}, xo[4] = {a.v[1], a.v[3], a.v[5], a.v[7]}, y[4] = {b.v[0], b.v[0], b.v[0], b.v[0]};
      kh_mul_lanes<4>(e, xe, y);
      kh_mul_lanes<4>(o, xo, y);
    } else {
      kh_mad_row(e + i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
      kh_mad_row(o + i, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
    }
    // row i+1 (odd): odd j -> even column -> e[i+1+j] ; even j -> odd column -> o[i+j]
    kh_mad_row(e + i + 2, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
    kh_mad_row(o + i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1]);
  }
  kh_combine_eo(r, e, o);
}

#ifndef KH_RARE_REDUCE
#define KH_RARE_REDUCE 1
#endif
#ifndef KH_FOLD_IN_O
#define KH_FOLD_IN_O 1
#endif
// ---- 512 -> 256 (mod P), canonical ------------------------------------------------------------------
template <int RR = KH_RARE_REDUCE>
KH_HD void fe_reduce_wide(fe &r, const uint32_t w[16]) {
  // t = lo + hi*977 + (hi << 32), 10 limbs
  uint32_t e[9], o[9];
  // KH_FOLD_IN_O: the "hi << 32" term (hi limb i at limb i+1 = o[i]) is what the o accumulator starts from, so that it rides in the
  // carry chain of the odd products instead of costing an 8-limb addition of its own (the four odd products are then links of a
  // chain, not plain ones)
  constexpr bool FOLD_IN_O = KH_FOLD_IN_O != 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { e[i] = w[i]; o[i] = FOLD_IN_O ? w[8 + i] : 0u; }
  e[8] = 0; o[8] = 0;
  kh_mad_row(e, w[8], w[10], w[12], w[14], 977u);  // hi limbs 0,2,4,6 -> columns 0,2,4,6
  if (KH_PLAIN_HEAD && !FOLD_IN_O) {               // o is zero: four carry-free products
    const uint32_t xo[4] = {w[9], w[11], w[13], w[15]}, y[4] = {977u, 977u, 977u, 977u};
    kh_mul_lanes<4>(o, xo, y);
  } else
  kh_mad_row(o, w[9], w[11], w[13], w[15], 977u);  // hi limbs 1,3,5,7 -> columns 1,3,5,7 (o[k] <-> limb k+1)
  uint32_t t[8], top0, top1;
  // limbs 1..8 : e[1..8] + o[0..7] ; limb 9 : o[8] + carry
  uint32_t s[8];
  uint32_t cf = kh_add8(s, e + 1, o);
  top1 = o[8] + cf;
  if (!FOLD_IN_O) {                                // += hi << 32 (limbs 1..8)
    cf = kh_add8(s, s, w + 8);
    top1 += cf;
  }
  t[0] = e[0];
#pragma unroll
  for (int i = 0; i < 7; i++) t[1 + i] = s[i];
  top0 = s[7];
  // second fold: top = top0 + top1*2^32 (< 2^35): t += top*977 + (top << 32)
  if (RR) {
  // one 8-limb addition instead of two (top*977 and top << 32 are added together first: three small limbs), and the final
  // conditional subtraction of P as a branch that is practically never taken: the folded value t + cf*2^256 is >= P only if it
  // overflowed 2^256 (cf) or its limbs 2..7 are all ones — probability ~2^-190 on field elements that come out of a multiplication
  // — so the common case tests that (3 LOP3 + 1 ISETP) instead of running the 8 additions + 8 selects of fe_final_reduce.  The result
  // is canonical either way (tests/test_gpu_field.py forces both cases on the device).
  uint32_t g[8];
  {
    const uint32_t lo977 = top0 * 977u, hi977 = kh_umulhi(top0, 977u) + top1 * 977u;   // top*977 < 2^45
    // g = top*977 + (top << 32): limb0 = lo977, limb1 = hi977 + top0, limb2 = top1 + carry, limb3 = carry
    const uint32_t g1 = hi977 + top0;
    const uint32_t c1 = (g1 < hi977) ? 1u : 0u;
    const uint32_t g2 = top1 + c1;
    g[0] = lo977; g[1] = g1; g[2] = g2; g[3] = (g2 < c1) ? 1u : 0u; g[4] = 0; g[5] = 0; g[6] = 0; g[7] = 0;
  }
  const uint32_t cf2 = kh_add8(t, t, g);
  const uint32_t ones = t[2] & t[3] & t[4] & t[5] & t[6] & t[7];
#if defined(__CUDA_ARCH__)
  if (__builtin_expect((cf2 != 0) | (ones == 0xFFFFFFFFu), 0)) {
#else
  if ((cf2 != 0) | (ones == 0xFFFFFFFFu)) {
#endif
    fe_final_reduce(r, t, cf2);
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  } else {
  const uint32_t f1[8] = {top0 * 977u, kh_umulhi(top0, 977u) + top1 * 977u, top1, 0, 0, 0, 0, 0};
  const uint32_t f2[8] = {0, top0, 0, 0, 0, 0, 0, 0};
  uint32_t cfa = kh_add8(t, t, f1);
  uint32_t cfb = kh_add8(t, t, f2);
  fe_final_reduce(r, t, cfa | cfb);
  }
}

template <int RR = KH_RARE_REDUCE>
KH_HD void fe_mul(fe &r, const fe &a, const fe &b) {
  uint32_t w[16];
  fe_mul_wide(w, a, b);
  fe_reduce_wide<RR>(r, w);
}
// a^2 with 36 instead of 64 wide multiplies: off-diagonal products once (even/odd columns as in
// fe_mul_wide), doubled, plus the eight squares.  Used where the FMA-heavy pipe is the bottleneck
// (xpoint / BSGS walks: ncu shows it 71 % busy with IMAD.WIDE).
KH_HD void fe_sqr_wide(uint32_t r[16], const fe &a) {
  uint32_t e[17], o[17];
#pragma unroll
  for (int i = 0; i < 17; i++) { e[i] = 0; o[i] = 0; }
  const uint32_t *v = a.v;
  // row i: a_i * a_j for j > i ; column i+j odd -> o[i+j-1], even -> e[i+j]
  if (KH_PLAIN_HEAD) {   // the first chain of each accumulator lands on zeros: seven carry-free products
    { const uint32_t x[4] = {v[0], v[0], v[0], v[0]}, y[4] = {v[1], v[3], v[5], v[7]}; kh_mul_lanes<4>(o + 0, x, y); }
    { const uint32_t x[3] = {v[0], v[0], v[0]}, y[3] = {v[2], v[4], v[6]}; kh_mul_lanes<3>(e + 2, x, y); }
  } else {
  { const uint32_t x[4] = {v[0], v[0], v[0], v[0]}, y[4] = {v[1], v[3], v[5], v[7]}; kh_mad_chain<4>(o + 0, x, y); }
  { const uint32_t x[3] = {v[0], v[0], v[0]}, y[3] = {v[2], v[4], v[6]}; kh_mad_chain<3>(e + 2, x, y); }
  }
  { const uint32_t x[3] = {v[1], v[1], v[1]}, y[3] = {v[2], v[4], v[6]}; kh_mad_chain<3>(o + 2, x, y); }
  { const uint32_t x[3] = {v[1], v[1], v[1]}, y[3] = {v[3], v[5], v[7]}; kh_mad_chain<3>(e + 4, x, y); }
  { const uint32_t x[3] = {v[2], v[2], v[2]}, y[3] = {v[3], v[5], v[7]}; kh_mad_chain<3>(o + 4, x, y); }
  { const uint32_t x[2] = {v[2], v[2]}, y[2] = {v[4], v[6]}; kh_mad_chain<2>(e + 6, x, y); }
  { const uint32_t x[2] = {v[3], v[3]}, y[2] = {v[4], v[6]}; kh_mad_chain<2>(o + 6, x, y); }
  { const uint32_t x[2] = {v[3], v[3]}, y[2] = {v[5], v[7]}; kh_mad_chain<2>(e + 8, x, y); }
  { const uint32_t x[2] = {v[4], v[4]}, y[2] = {v[5], v[7]}; kh_mad_chain<2>(o + 8, x, y); }
  { const uint32_t x[1] = {v[4]}, y[1] = {v[6]}; kh_mad_chain<1>(e + 10, x, y); }
  { const uint32_t x[1] = {v[5]}, y[1] = {v[6]}; kh_mad_chain<1>(o + 10, x, y); }
  { const uint32_t x[1] = {v[5]}, y[1] = {v[7]}; kh_mad_chain<1>(e + 12, x, y); }
  { const uint32_t x[1] = {v[6]}, y[1] = {v[7]}; kh_mad_chain<1>(o + 12, x, y); }
  uint32_t t[16];
  kh_combine_eo(t, e, o);
  kh_double_add_squares(r, t, v);
}
template <int RR = KH_RARE_REDUCE>
KH_HD void fe_sqr(fe &r, const fe &a) {
  uint32_t w[16];
  fe_sqr_wide(w, a);
  fe_reduce_wide<RR>(r, w);
}

// Out-of-line multiply: the hash-heavy scan kernels are instruction-fetch bound (ncu: stall no_instruction,
// GPC instruction-cache requests at >90 % of peak), so their EC loop calls ONE shared copy of the multiplier
// instead of inlining ~2.4 KB of code at each of its eight call sites.  Operands travel by value (registers).
#if defined(__CUDACC__)
template <int RR = KH_RARE_REDUCE>
static __device__ __noinline__ fe fe_mul_ol(fe a, fe b) {
  fe r;
  fe_mul<RR>(r, a, b);
  return r;
}
#endif
// RR: the form of the final reduction (KH_RARE_REDUCE); a kernel picks ONE value for every multiplication it contains, so that it
// holds one out-of-line copy
template <bool OUTLINE, int RR = KH_RARE_REDUCE>
KH_HD void fe_mul_sel(fe &r, const fe &a, const fe &b) {
#if defined(__CUDA_ARCH__)
  if (OUTLINE) { r = fe_mul_ol<RR>(a, b); return; }
#endif
  fe_mul<RR>(r, a, b);
}

// r = a^(2^n)
KH_HD void fe_sqr_n(fe &r, const fe &a, int n) {
  r = a;
#pragma unroll 1
  for (int i = 0; i < n; i++) fe_sqr(r, r);
}

// a^(P-2): 255 squarings + 15 multiplications (addition chain over the run-lengths of P-2:
// 223 ones, 0, 22 ones, 0000, 1, 0, 11, 0, 1).  inv(0) = 0, like Int::ModInv's "no inverse" result.
// Cold code (once per 1024 points): on the device every multiply goes through the shared out-of-line copy
// so that the inversion does not flush the hot loop out of the instruction cache.
template <int RR = KH_RARE_REDUCE>
KH_HD void fe_mul_cold(fe &r, const fe &a, const fe &b) {
#if defined(__CUDA_ARCH__)
  r = fe_mul_ol<RR>(a, b);
#else
  fe_mul<RR>(r, a, b);
#endif
}
// SQR: the 255 squarings of the inversion through a dedicated out-of-line squaring (44 instead of 72 wide multiplies each: 2.8 % of
// all wide multiplies of the x-only walk) — for the kernels where the multiplier, not the instruction cache, is the bound
#if defined(__CUDACC__)
template <int RR = KH_RARE_REDUCE>
static __device__ __noinline__ fe fe_sqr_ol(fe a) {
  fe r;
  fe_sqr<RR>(r, a);
  return r;
}
#endif
// squaring in a kernel that calls out-of-line copies: the shared multiplier (SQR = false: one copy of code) or the squaring
template <bool SQR, int RR = KH_RARE_REDUCE>
KH_HD void fe_sqr_sel(fe &r, const fe &a) {
#if defined(__CUDA_ARCH__)
  if (SQR) r = fe_sqr_ol<RR>(a); else r = fe_mul_ol<RR>(a, a);
#else
  if (SQR) fe_sqr<RR>(r, a); else fe_mul<RR>(r, a, a);
#endif
}
template <int RR = KH_RARE_REDUCE, bool SQR = false>
KH_HD void fe_sqr_n_cold(fe &r, const fe &a, int n) {
  r = a;
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#if defined(__CUDA_ARCH__)
    if (SQR) r = fe_sqr_ol<RR>(r); else r = fe_mul_ol<RR>(r, r);
#else
    if (SQR) fe_sqr<RR>(r, r); else fe_mul<RR>(r, r, r);
#endif
  }
}
template <int RR = KH_RARE_REDUCE, bool SQR = false>
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static
#endif
void fe_inv(fe &r, const fe &a) {
  fe x2, x3, x6, x9, x11, x22, x44, x88, x176, x220, x223, t;
  fe_mul_cold<RR>(x2, a, a); fe_mul_cold<RR>(x2, x2, a);
  fe_mul_cold<RR>(x3, x2, x2); fe_mul_cold<RR>(x3, x3, a);
  fe_sqr_n_cold<RR, SQR>(x6, x3, 3); fe_mul_cold<RR>(x6, x6, x3);
  fe_sqr_n_cold<RR, SQR>(x9, x6, 3); fe_mul_cold<RR>(x9, x9, x3);
  fe_sqr_n_cold<RR, SQR>(x11, x9, 2); fe_mul_cold<RR>(x11, x11, x2);
  fe_sqr_n_cold<RR, SQR>(x22, x11, 11); fe_mul_cold<RR>(x22, x22, x11);
  fe_sqr_n_cold<RR, SQR>(x44, x22, 22); fe_mul_cold<RR>(x44, x44, x22);
  fe_sqr_n_cold<RR, SQR>(x88, x44, 44); fe_mul_cold<RR>(x88, x88, x44);
  fe_sqr_n_cold<RR, SQR>(x176, x88, 88); fe_mul_cold<RR>(x176, x176, x88);
  fe_sqr_n_cold<RR, SQR>(x220, x176, 44); fe_mul_cold<RR>(x220, x220, x44);
  fe_sqr_n_cold<RR, SQR>(x223, x220, 3); fe_mul_cold<RR>(x223, x223, x3);
  fe_sqr_n_cold<RR, SQR>(t, x223, 23); fe_mul_cold<RR>(t, t, x22);
  fe_sqr_n_cold<RR, SQR>(t, t, 5); fe_mul_cold<RR>(t, t, a);
  fe_sqr_n_cold<RR, SQR>(t, t, 3); fe_mul_cold<RR>(t, t, x2);
  fe_sqr_n_cold<RR, SQR>(t, t, 2); fe_mul_cold<RR>(r, t, a);
}

// Inversion with operands in registers: fe_inv takes references, which makes its arguments address-taken locals of the
// caller (stack traffic at every use inside the walk's hot loop); the by-value wrapper keeps them in registers.
#if defined(__CUDACC__)
static __device__ __noinline__ fe fe_inv_ol(fe a) {
  fe r;
  fe_inv(r, a);
  return r;
}
#endif
KH_HD void fe_inv_reg(fe &r, const fe &a) {
#if defined(__CUDA_ARCH__)
  r = fe_inv_ol(a);
#else
  fe_inv(r, a);
#endif
}

// ---- (de)serialisation -------------------------------------------------------------------------------
// 32-byte big-endian string (Int::Get32Bytes, Int.cpp:308) <-> limbs
KH_HD void fe_from_be(fe &r, const uint8_t *b) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t *p = b + 4 * (7 - i);
    r.v[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
  }
}
KH_HD void fe_to_be(uint8_t *b, const fe &a) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint8_t *p = b + 4 * (7 - i);
    p[0] = (uint8_t)(a.v[i] >> 24); p[1] = (uint8_t)(a.v[i] >> 16); p[2] = (uint8_t)(a.v[i] >> 8); p[3] = (uint8_t)a.v[i];
  }
}

}  // namespace kh
