// ec.cuh — secp256k1 group operations used OUTSIDE the batch inner loop (set-up of the G-multiple
// tables and thread start points, hit derivation, BSGS refinement).  The inner loop itself
// (affine P +- i*G with a shared Montgomery inverse) lives in walk.cuh.
//
// Replaces Secp256K1::ComputePublicKey / ScalarBaseMultiplication (SECP256K1.cpp:205, :702),
// AddDirect (:455), DoubleDirect (:589), Negation (:316).  The reference uses a wNAF-7 table; here
// every thread that needs k*G does a plain Jacobian double-and-add (set-up cost only, off the hot
// path) and one field inversion.
#pragma once
#include "fe.cuh"

namespace kh {

struct ge {   // affine point; inf != 0 means the point at infinity
  fe x, y;
  uint32_t inf;
};
struct gej {  // Jacobian
  fe x, y, z;
  uint32_t inf;
};

struct u256 {  // plain 256-bit integer (scalars / private keys), little-endian limbs
  uint32_t v[8];
};

// Generator (SECP256K1.cpp:161-162), limbs little-endian
#define KH_GX {0x16F81798u, 0x59F2815Bu, 0x2DCE28D9u, 0x029BFCDBu, 0xCE870B07u, 0x55A06295u, 0xF9DCBBACu, 0x79BE667Eu}
#define KH_GY {0xFB10D4B8u, 0x9C47D08Fu, 0xA6855419u, 0xFD17B448u, 0x0E1108A8u, 0x5DA4FBFCu, 0x26A3C465u, 0x483ADA77u}

KH_HD void ge_set_g(ge &g) {
  const uint32_t gx[8] = KH_GX, gy[8] = KH_GY;
#pragma unroll
  for (int i = 0; i < 8; i++) { g.x.v[i] = gx[i]; g.y.v[i] = gy[i]; }
  g.inf = 0;
}

// ---- plain 256-bit integers ------------------------------------------------------------------------
KH_HD void u256_from_be(u256 &r, const uint8_t *b) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t *p = b + 4 * (7 - i);
    r.v[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
  }
}
KH_HD void u256_to_be(uint8_t *b, const u256 &a) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint8_t *p = b + 4 * (7 - i);
    p[0] = (uint8_t)(a.v[i] >> 24); p[1] = (uint8_t)(a.v[i] >> 16); p[2] = (uint8_t)(a.v[i] >> 8); p[3] = (uint8_t)a.v[i];
  }
}
// r = a + b*m (mod 2^256), m 64-bit
KH_HD void u256_add_mul64(u256 &r, const u256 &a, const u256 &b, uint64_t m) {
  const uint32_t m0 = (uint32_t)m, m1 = (uint32_t)(m >> 32);
  uint32_t t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = 0;
  // t = b * m (truncated), schoolbook with 64-bit temporaries (set-up path only)
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (uint64_t)b.v[i] * m0; t[i] = (uint32_t)c; c >>= 32; }
  c = 0;
#pragma unroll
  for (int i = 0; i < 7; i++) { c += (uint64_t)b.v[i] * m1 + t[i + 1]; t[i + 1] = (uint32_t)c; c >>= 32; }
  c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + t[i]; r.v[i] = (uint32_t)c; c >>= 32; }
}
// curve order n (SECP256K1.cpp:164)
#define KH_N {0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}
// r = n - a   (Int::Neg + Add(order), keyhunt.cpp:3632-3633)
KH_HD void u256_neg_mod_n(u256 &r, const u256 &a) {
  const uint32_t n[8] = KH_N;
  kh_sub8(r.v, n, a.v);
}

// ---- Jacobian arithmetic (a = 0) ---------------------------------------------------------------------
KH_HD void gej_set_inf(gej &r) { r.inf = 1; fe_set_zero(r.x); fe_set_zero(r.y); fe_set_zero(r.z); }

#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static
#endif
void gej_double(gej &r, const gej &p) {
  if (p.inf || fe_is_zero(p.y)) { gej_set_inf(r); return; }
  fe a, b, c, d, e, f, t;
  fe_sqr(a, p.x);
  fe_sqr(b, p.y);
  fe_sqr(c, b);
  fe_add(t, p.x, b); fe_sqr(t, t); fe_sub(t, t, a); fe_sub(t, t, c); fe_add(d, t, t);
  fe_add(e, a, a); fe_add(e, e, a);
  fe_sqr(f, e);
  gej o; o.inf = 0;
  fe_mul(o.z, p.y, p.z); fe_add(o.z, o.z, o.z);
  fe_sub(o.x, f, d); fe_sub(o.x, o.x, d);
  fe_sub(t, d, o.x); fe_mul(o.y, e, t);
  fe_add(c, c, c); fe_add(c, c, c); fe_add(c, c, c);
  fe_sub(o.y, o.y, c);
  r = o;
}

#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static
#endif
void gej_add_ge(gej &r, const gej &p, const ge &q) {
  if (q.inf) { r = p; return; }
  if (p.inf) { r.x = q.x; r.y = q.y; fe_set_u32(r.z, 1); r.inf = 0; return; }
  fe z2, u2, s2, h, rr, h2, h3, t;
  fe_sqr(z2, p.z);
  fe_mul(u2, q.x, z2);
  fe_mul(s2, q.y, p.z); fe_mul(s2, s2, z2);
  fe_sub(h, u2, p.x);
  fe_sub(rr, s2, p.y);
  if (fe_is_zero(h)) {
    if (fe_is_zero(rr)) { gej_double(r, p); return; }
    gej_set_inf(r); return;
  }
  fe_sqr(h2, h); fe_mul(h3, h2, h);
  gej o; o.inf = 0;
  fe_mul(t, p.x, h2);
  fe_sqr(o.x, rr); fe_sub(o.x, o.x, h3); fe_sub(o.x, o.x, t); fe_sub(o.x, o.x, t);
  fe_sub(t, t, o.x); fe_mul(o.y, rr, t);
  fe_mul(t, p.y, h3); fe_sub(o.y, o.y, t);
  fe_mul(o.z, p.z, h);
  r = o;
}

KH_HD void gej_to_ge(ge &r, const gej &p) {
  if (p.inf) { r.inf = 1; fe_set_zero(r.x); fe_set_zero(r.y); return; }
  fe zi, zi2, zi3;
  fe_inv(zi, p.z); fe_sqr(zi2, zi); fe_mul(zi3, zi2, zi);
  fe_mul(r.x, p.x, zi2); fe_mul(r.y, p.y, zi3);
  r.inf = 0;
}

// r = k * base  (double-and-add, MSB first; k is any 256-bit integer)
KH_HD void ge_scalar_mul(ge &r, const ge &base, const u256 &k) {
  gej acc;
  gej_set_inf(acc);
#pragma unroll 1
  for (int i = 255; i >= 0; i--) {
    gej_double(acc, acc);
    if ((k.v[i >> 5] >> (i & 31)) & 1) gej_add_ge(acc, acc, base);
  }
  gej_to_ge(r, acc);
}
KH_HD void ge_mul_g(ge &r, const u256 &k) {
  ge g;
  ge_set_g(g);
  ge_scalar_mul(r, g, k);
}
// k*G by a fixed-base comb: comb[(w*256 + d)*16 ..] = d * 2^(8w) * G as 8 x-limbs + 8 y-limbs (d = 1..255, w = 0..31; built
// once per context by kh_comb_kernel, 512 KB, L2-resident).  At most 32 mixed additions + one inversion instead of 256
// doublings + ~128 additions: the set-up of a walk (one k*G per walker thread) is 5x shorter, which is what a short
// kh_bsgs_search call or a server request consists of.  comb == nullptr: the plain double-and-add (host test build).
#define KH_COMB_WORDS (32 * 256 * 16)
KH_HD void ge_mul_g_comb(ge &r, const u256 &k, const uint32_t *comb) {
  if (!comb) { ge_mul_g(r, k); return; }
  gej acc;
  gej_set_inf(acc);
#pragma unroll 1
  for (int w = 0; w < 32; w++) {
    const uint32_t d = (k.v[w >> 2] >> (8 * (w & 3))) & 0xFFu;
    if (!d) continue;
    const uint32_t *e = comb + (size_t)(w * 256 + d) * 16;
    ge q;
#pragma unroll
    for (int l = 0; l < 8; l++) { q.x.v[l] = e[l]; q.y.v[l] = e[8 + l]; }
    q.inf = 0;
    gej_add_ge(acc, acc, q);
  }
  gej_to_ge(r, acc);
}
KH_HD void ge_neg(ge &r, const ge &p) { r.x = p.x; fe_neg(r.y, p.y); r.inf = p.inf; }
// full affine addition with every special case (used for start points: Q + k*G)
KH_HD void ge_add(ge &r, const ge &p, const ge &q) {
  gej j;
  j.x = p.x; j.y = p.y; fe_set_u32(j.z, 1); j.inf = p.inf;
  gej s;
  gej_add_ge(s, j, q);
  gej_to_ge(r, s);
}

// Secp256K1::AddDirect (SECP256K1.cpp:455) exactly as the reference computes it: no special cases,
// and an unusable difference (dx = 0) yields inv = 0 and therefore the same deterministic garbage
// the reference produces.  Used by the BSGS tier checks so that they agree with the reference even
// on its degenerate inputs.
KH_HD void ge_add_direct(ge &r, const ge &p1, const ge &p2) {
  fe dy, dx, s, p, t;
  fe_sub(dy, p2.y, p1.y);
  fe_sub(dx, p2.x, p1.x);
  fe_inv(dx, dx);
  fe_mul(s, dy, dx);
  fe_sqr(p, s);
  ge o;
  fe_sub(o.x, p, p1.x);
  fe_sub(o.x, o.x, p2.x);
  fe_sub(t, p2.x, o.x);
  fe_mul(o.y, t, s);
  fe_sub(o.y, o.y, p2.y);
  o.inf = 0;
  r = o;
}
// r = a + b, r = a - b for a small b (keys near a base key)
KH_HD void u256_add_u64(u256 &r, const u256 &a, uint64_t b) {
  uint32_t bb[8] = {(uint32_t)b, (uint32_t)(b >> 32), 0, 0, 0, 0, 0, 0};
  kh_add8(r.v, a.v, bb);
}
KH_HD void u256_sub_u64(u256 &r, const u256 &a, uint64_t b) {
  uint32_t bb[8] = {(uint32_t)b, (uint32_t)(b >> 32), 0, 0, 0, 0, 0, 0};
  kh_sub8(r.v, a.v, bb);
}
KH_HD void u256_set_u64(u256 &r, uint64_t b) {
  r.v[0] = (uint32_t)b; r.v[1] = (uint32_t)(b >> 32);
#pragma unroll
  for (int i = 2; i < 8; i++) r.v[i] = 0;
}

// r = a*b mod n  (Int::ModMulK1order, IntMod.cpp:1111) — only for the -e hit fix-ups (k*lambda), cold
KH_HD void u256_mulmod_n(u256 &r, const u256 &a, const u256 &b) {
  const uint32_t n[8] = KH_N;
  uint32_t t[16];
#pragma unroll
  for (int i = 0; i < 16; i++) t[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) { c += (uint64_t)a.v[i] * b.v[j] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
    t[i + 8] = (uint32_t)c;
  }
  u256 acc;
  for (int i = 0; i < 8; i++) acc.v[i] = 0;
  for (int bit = 511; bit >= 0; bit--) {
    const uint32_t top = acc.v[7] >> 31;
    for (int i = 7; i > 0; i--) acc.v[i] = (acc.v[i] << 1) | (acc.v[i - 1] >> 31);
    acc.v[0] = (acc.v[0] << 1) | ((t[bit >> 5] >> (bit & 31)) & 1u);
    uint32_t d[8];
    const uint32_t borrow = kh_sub8(d, acc.v, n);
    if (top || !borrow) { for (int i = 0; i < 8; i++) acc.v[i] = d[i]; }
  }
  r = acc;
}
// lambda, lambda^2 as the reference's -e sets them (keyhunt.cpp:928-929)
#define KH_LAMBDA  {0x1B23BD72u, 0xDF02967Cu, 0x20816678u, 0x122E22EAu, 0x8812645Au, 0xA5261C02u, 0xC05C30E0u, 0x5363AD4Cu}
#define KH_LAMBDA2 {0xB51283CEu, 0xE0CFC810u, 0x8EC739C2u, 0xA880B9FCu, 0x77ED9BA4u, 0x5AD9E3FDu, 0x3FA3CF1Fu, 0xAC9C52B3u}

}  // namespace kh
