/* oracle/kh_oracle.c — TEST INFRASTRUCTURE ONLY (see kh_oracle.h).  PARITY: pinned.
 *
 * Plain-C CPU restatement of the reference's key-range search hot path.  Every function cites the
 * reference file:line (relative to /root/reference) it follows.  Differences from the reference
 * that do not change results:
 *   - field results are always canonical (< P); the reference's ModMulK1 is lazily reduced
 *     (IntMod.cpp:912-913), equal as residues.
 *   - Int::ModInv (DRS62 xgcd, IntMod.cpp:382) is restated as Fermat a^(P-2); same value, inv(0)=0.
 *   - ComputePublicKey (wNAF-7, SECP256K1.cpp:702) is restated as Jacobian double-and-add.
 *   - the per-batch ComputePublicKey of thread_process (keyhunt.cpp:3352) is replaced by the
 *     "next start point" addition the reference itself computes at :3840-3855 (same point).
 */
#define _GNU_SOURCE
#include "kh_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;          /* little-endian limbs, always < P */
typedef struct { fe x, y; } ge;                /* affine point */
typedef struct { fe x, y, z; int inf; } gej;   /* Jacobian */

/* SECP256K1.cpp:155-164 */
static const fe FE_P = {{0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
static const fe FE_N = {{0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
static const ge GE_G = {{{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}},
                        {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}}};
#define K1C 0x1000003D1ULL /* 2^256 mod P */
/* endomorphism constants exactly as the -e option sets them (keyhunt.cpp:928-931) */
static const fe SC_LAMBDA = {{0xDF02967C1B23BD72ULL, 0x122E22EA20816678ULL, 0xA5261C028812645AULL, 0x5363AD4CC05C30E0ULL}};
static const fe SC_LAMBDA2 = {{0xE0CFC810B51283CEULL, 0xA880B9FC8EC739C2ULL, 0x5AD9E3FD77ED9BA4ULL, 0xAC9C52B33FA3CF1FULL}};
static const fe FE_BETA = {{0xC1396C28719501EEULL, 0x9CF0497512F58995ULL, 0x6E64479EAC3434E9ULL, 0x7AE96A2B657C0710ULL}};
static const fe FE_BETA2 = {{0x3EC693D68E6AFA40ULL, 0x630FB68AED0A766AULL, 0x919BB86153CBCB16ULL, 0x851695D49A83F8EFULL}};

/* ------------------------------------------------------------------------------------------------
 * 256-bit helpers (Int.cpp)
 * ---------------------------------------------------------------------------------------------- */
static void fe_from_be(fe *r, const uint8_t b[32]) { /* Int::Set32Bytes Int.cpp:297 */
  for (int i = 0; i < 4; i++) {
    uint64_t v = 0;
    for (int j = 0; j < 8; j++) v = (v << 8) | b[(3 - i) * 8 + j];
    r->l[i] = v;
  }
}
static void fe_to_be(uint8_t b[32], const fe *a) { /* Int::Get32Bytes Int.cpp:308 */
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) b[(3 - i) * 8 + j] = (uint8_t)(a->l[i] >> (56 - 8 * j));
}
static int u256_cmp(const fe *a, const fe *b) {
  for (int i = 3; i >= 0; i--) {
    if (a->l[i] < b->l[i]) return -1;
    if (a->l[i] > b->l[i]) return 1;
  }
  return 0;
}
static int u256_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static uint64_t u256_add(fe *r, const fe *a, const fe *b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; r->l[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
static uint64_t u256_sub(fe *r, const fe *a, const fe *b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - borrow;
    r->l[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  return borrow;
}
static void u256_set_u64(fe *r, uint64_t v) { r->l[0] = v; r->l[1] = r->l[2] = r->l[3] = 0; }
/* r = a * b (b 64-bit), truncated to 256 bits (callers keep values small) */
static void u256_mul_u64(fe *r, const fe *a, uint64_t b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a->l[i] * b; r->l[i] = (uint64_t)c; c >>= 64; }
}

/* r = a*b mod n  (Int::ModMulK1order IntMod.cpp:1111; value only, restated as product + binary reduction) */
static void sc_mul(fe *r, const fe *a, const fe *b) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a->l[i] * b->l[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
    t[i + 4] = (uint64_t)c;
  }
  fe acc; u256_set_u64(&acc, 0);
  for (int bit = 511; bit >= 0; bit--) {
    uint64_t top = acc.l[3] >> 63;
    for (int i = 3; i > 0; i--) acc.l[i] = (acc.l[i] << 1) | (acc.l[i - 1] >> 63);
    acc.l[0] = (acc.l[0] << 1) | ((t[bit / 64] >> (bit % 64)) & 1);
    if (top || u256_cmp(&acc, &FE_N) >= 0) u256_sub(&acc, &acc, &FE_N);
  }
  *r = acc;
}

/* ------------------------------------------------------------------------------------------------
 * Field arithmetic mod P
 * ---------------------------------------------------------------------------------------------- */
static void fe_add(fe *r, const fe *a, const fe *b) { /* Int::ModAdd IntMod.cpp:41 */
  uint64_t c = u256_add(r, a, b);
  if (c || u256_cmp(r, &FE_P) >= 0) u256_sub(r, r, &FE_P);
}
static void fe_sub(fe *r, const fe *a, const fe *b) { /* Int::ModSub IntMod.cpp:72 */
  if (u256_sub(r, a, b)) u256_add(r, r, &FE_P);
}
static void fe_neg(fe *r, const fe *a) { /* Int::ModNeg IntMod.cpp:102 */
  if (u256_is_zero(a)) { *r = *a; return; }
  u256_sub(r, &FE_P, a);
}
static void fe_reduce512(fe *r, const uint64_t t[8]) { /* fold by 0x1000003D1 twice, IntMod.cpp:886-913 */
  u128 c = 0;
  uint64_t lo[4];
  for (int i = 0; i < 4; i++) { c += (u128)t[4 + i] * K1C + t[i]; lo[i] = (uint64_t)c; c >>= 64; }
  uint64_t top = (uint64_t)c;
  c = (u128)top * K1C + lo[0];
  lo[0] = (uint64_t)c; c >>= 64;
  for (int i = 1; i < 4; i++) { c += lo[i]; lo[i] = (uint64_t)c; c >>= 64; }
  if (c) { /* wrapped past 2^256 once more: 2^256 == K1C (mod P) */
    c = (u128)lo[0] + K1C;
    lo[0] = (uint64_t)c; c >>= 64;
    for (int i = 1; i < 4; i++) { c += lo[i]; lo[i] = (uint64_t)c; c >>= 64; }
  }
  for (int i = 0; i < 4; i++) r->l[i] = lo[i];
  if (u256_cmp(r, &FE_P) >= 0) u256_sub(r, r, &FE_P);
}
static void fe_mul(fe *r, const fe *a, const fe *b) { /* Int::ModMulK1 IntMod.cpp:855 */
  uint64_t t[8] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a->l[i] * b->l[j] + t[i + j];
      t[i + j] = (uint64_t)c;
      c >>= 64;
    }
    t[i + 4] = (uint64_t)c;
  }
  fe_reduce512(r, t);
}
static void fe_sqr(fe *r, const fe *a) { fe_mul(r, a, a); } /* Int::ModSquareK1 IntMod.cpp:977 */
static void fe_inv(fe *r, const fe *a) { /* Int::ModInv IntMod.cpp:382 — value a^(P-2); 0 -> 0 */
  /* P-2 = 0xFFFFFFFF FFFFFFFF FFFFFFFF FFFFFFFF FFFFFFFF FFFFFFFF FFFFFFFE FFFFFC2D */
  fe e = FE_P, acc, base = *a;
  e.l[0] -= 2;
  u256_set_u64(&acc, 1);
  for (int i = 0; i < 256; i++) {
    if ((e.l[i / 64] >> (i % 64)) & 1) fe_mul(&acc, &acc, &base);
    fe_sqr(&base, &base);
  }
  *r = acc;
}

/* ------------------------------------------------------------------------------------------------
 * Group law
 * ---------------------------------------------------------------------------------------------- */
static void ge_add_direct(ge *r, const ge *p1, const ge *p2) { /* Secp256K1::AddDirect SECP256K1.cpp:455 */
  fe dy, dx, s, p, t;
  fe_sub(&dy, &p2->y, &p1->y);
  fe_sub(&dx, &p2->x, &p1->x);
  fe_inv(&dx, &dx);
  fe_mul(&s, &dy, &dx);
  fe_sqr(&p, &s);
  ge out;
  fe_sub(&out.x, &p, &p1->x);
  fe_sub(&out.x, &out.x, &p2->x);
  fe_sub(&t, &p2->x, &out.x);
  fe_mul(&out.y, &t, &s);
  fe_sub(&out.y, &out.y, &p2->y);
  *r = out;
}
static void ge_double_direct(ge *r, const ge *p) { /* Secp256K1::DoubleDirect SECP256K1.cpp:589 */
  fe s, a, t, x2;
  fe_sqr(&x2, &p->x);
  fe_add(&a, &x2, &x2);
  fe_add(&a, &a, &x2);          /* 3x^2 */
  fe_add(&t, &p->y, &p->y);     /* 2y */
  fe_inv(&t, &t);
  fe_mul(&s, &a, &t);
  ge out;
  fe_sqr(&out.x, &s);
  fe_sub(&out.x, &out.x, &p->x);
  fe_sub(&out.x, &out.x, &p->x);
  fe_sub(&t, &p->x, &out.x);
  fe_mul(&out.y, &t, &s);
  fe_sub(&out.y, &out.y, &p->y);
  *r = out;
}
static void ge_neg(ge *r, const ge *p) { r->x = p->x; fe_neg(&r->y, &p->y); } /* Negation SECP256K1.cpp:316 */

static void gej_double(gej *r, const gej *p) {
  if (p->inf || u256_is_zero(&p->y)) { r->inf = 1; return; }
  fe a, b, c, d, e, f, t;
  fe_sqr(&a, &p->x);
  fe_sqr(&b, &p->y);
  fe_sqr(&c, &b);
  fe_add(&t, &p->x, &b); fe_sqr(&t, &t); fe_sub(&t, &t, &a); fe_sub(&t, &t, &c); fe_add(&d, &t, &t);
  fe_add(&e, &a, &a); fe_add(&e, &e, &a);
  fe_sqr(&f, &e);
  gej o; o.inf = 0;
  fe_mul(&o.z, &p->y, &p->z); fe_add(&o.z, &o.z, &o.z);
  fe_sub(&o.x, &f, &d); fe_sub(&o.x, &o.x, &d);
  fe_sub(&t, &d, &o.x); fe_mul(&o.y, &e, &t);
  fe_add(&c, &c, &c); fe_add(&c, &c, &c); fe_add(&c, &c, &c);
  fe_sub(&o.y, &o.y, &c);
  *r = o;
}
static void gej_add_ge(gej *r, const gej *p, const ge *q) {
  if (p->inf) { r->x = q->x; r->y = q->y; u256_set_u64(&r->z, 1); r->inf = 0; return; }
  fe z2, u2, s2, h, rr, h2, h3, t;
  fe_sqr(&z2, &p->z);
  fe_mul(&u2, &q->x, &z2);
  fe_mul(&s2, &q->y, &p->z); fe_mul(&s2, &s2, &z2);
  fe_sub(&h, &u2, &p->x);
  fe_sub(&rr, &s2, &p->y);
  if (u256_is_zero(&h)) {
    if (u256_is_zero(&rr)) { gej_double(r, p); return; }
    r->inf = 1; return;
  }
  fe_sqr(&h2, &h); fe_mul(&h3, &h2, &h);
  gej o; o.inf = 0;
  fe_mul(&t, &p->x, &h2);
  fe_sqr(&o.x, &rr); fe_sub(&o.x, &o.x, &h3); fe_sub(&o.x, &o.x, &t); fe_sub(&o.x, &o.x, &t);
  fe_sub(&t, &t, &o.x); fe_mul(&o.y, &rr, &t);
  fe_mul(&t, &p->y, &h3); fe_sub(&o.y, &o.y, &t);
  fe_mul(&o.z, &p->z, &h);
  *r = o;
}
static void gej_to_ge(ge *r, const gej *p) {
  if (p->inf) { memset(r, 0, sizeof(*r)); return; }
  fe zi, zi2, zi3;
  fe_inv(&zi, &p->z); fe_sqr(&zi2, &zi); fe_mul(&zi3, &zi2, &zi);
  fe_mul(&r->x, &p->x, &zi2); fe_mul(&r->y, &p->y, &zi3);
}
static void ge_scalar_mul(ge *r, const ge *base, const fe *k) { /* ComputePublicKey SECP256K1.cpp:205 */
  gej acc; acc.inf = 1;
  for (int i = 255; i >= 0; i--) {
    gej_double(&acc, &acc);
    if ((k->l[i / 64] >> (i % 64)) & 1) gej_add_ge(&acc, &acc, base);
  }
  gej_to_ge(r, &acc);
}

/* ------------------------------------------------------------------------------------------------
 * The 1024-point group (CPU_GRP_SIZE keyhunt.cpp:299); Gn table init_generator keyhunt.cpp:5266
 * ---------------------------------------------------------------------------------------------- */
#define GRP 1024
#define HALF 512
typedef struct { ge gn[HALF]; ge g2n; } gtable;

static void gtable_init(gtable *t, const ge *base) { /* Gn[i]=(i+1)*base, _2Gn = 1024*base */
  t->gn[0] = *base;
  ge_double_direct(&t->gn[1], base);
  for (int i = 2; i < HALF; i++) ge_add_direct(&t->gn[i], &t->gn[i - 1], base);
  ge_double_direct(&t->g2n, &t->gn[HALF - 1]);
}

/* one batch around centre c: pts[i] for i in 0..1023 (pts[512]=c); returns the next centre.
 * keyhunt.cpp:3355-3461 (+ :3840-3855), IntGroup::ModInv IntGroup.cpp:36-57 */
static void batch_around(const gtable *t, const ge *c, int with_y, ge *pts, ge *next) {
  fe dx[HALF + 1], subp[HALF + 1], inv, nv, dy, s, p;
  for (int i = 0; i < HALF; i++) fe_sub(&dx[i], &t->gn[i].x, &c->x);
  fe_sub(&dx[HALF], &t->g2n.x, &c->x);
  subp[0] = dx[0];
  for (int i = 1; i <= HALF; i++) fe_mul(&subp[i], &subp[i - 1], &dx[i]);
  fe_inv(&inv, &subp[HALF]);
  for (int i = HALF; i > 0; i--) {
    fe_mul(&nv, &subp[i - 1], &inv);
    fe_mul(&inv, &inv, &dx[i]);
    dx[i] = nv;
  }
  dx[0] = inv;
  pts[HALF] = *c;
  for (int i = 0; i < HALF; i++) {
    const ge *g = &t->gn[i];
    /* P = c + (i+1)G */
    if (i < HALF - 1) {
      ge *pp = &pts[HALF + i + 1];
      fe_sub(&dy, &g->y, &c->y);
      fe_mul(&s, &dy, &dx[i]);
      fe_sqr(&p, &s);
      fe_sub(&pp->x, &p, &c->x); fe_sub(&pp->x, &pp->x, &g->x);
      if (with_y) { fe_sub(&pp->y, &g->x, &pp->x); fe_mul(&pp->y, &pp->y, &s); fe_sub(&pp->y, &pp->y, &g->y); }
      else pp->y = c->y;
    }
    /* P = c - (i+1)G */
    ge *pn = &pts[HALF - i - 1];
    fe_neg(&dy, &g->y); fe_sub(&dy, &dy, &c->y);
    fe_mul(&s, &dy, &dx[i]);
    fe_sqr(&p, &s);
    fe_sub(&pn->x, &p, &c->x); fe_sub(&pn->x, &pn->x, &g->x);
    if (with_y) { fe_sub(&pn->y, &g->x, &pn->x); fe_mul(&pn->y, &pn->y, &s); fe_add(&pn->y, &pn->y, &g->y); }
    else pn->y = c->y;
  }
  if (next) {
    fe_sub(&dy, &t->g2n.y, &c->y);
    fe_mul(&s, &dy, &dx[HALF]);
    fe_sqr(&p, &s);
    ge n;
    fe_sub(&n.x, &p, &c->x); fe_sub(&n.x, &n.x, &t->g2n.x);
    fe_sub(&n.y, &t->g2n.x, &n.x); fe_mul(&n.y, &n.y, &s); fe_sub(&n.y, &n.y, &t->g2n.y);
    *next = n;
  }
}

/* ------------------------------------------------------------------------------------------------
 * SHA-256 (hash/sha256.cpp), RIPEMD-160 (hash/ripemd160.cpp), Keccak-256 (sha3/), XXH64 (xxhash.h)
 * ---------------------------------------------------------------------------------------------- */
static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#define ROR32(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
#define ROL32(x, n) (((x) << (n)) | ((x) >> (32 - (n))))
static void sha256_block(uint32_t st[8], const uint8_t blk[64]) {
  uint32_t w[64], a, b, c, d, e, f, g, h;
  for (int i = 0; i < 16; i++)
    w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = ROR32(w[i - 15], 7) ^ ROR32(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = ROR32(w[i - 2], 17) ^ ROR32(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  a = st[0]; b = st[1]; c = st[2]; d = st[3]; e = st[4]; f = st[5]; g = st[6]; h = st[7];
  for (int i = 0; i < 64; i++) {
    uint32_t t1 = h + (ROR32(e, 6) ^ ROR32(e, 11) ^ ROR32(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i];
    uint32_t t2 = (ROR32(a, 2) ^ ROR32(a, 13) ^ ROR32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
void kho_sha256(const uint8_t *in, uint64_t len, uint8_t out[32]) {
  uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  uint64_t off = 0;
  for (; off + 64 <= len; off += 64) sha256_block(st, in + off);
  uint8_t tail[128];
  uint64_t rem = len - off;
  memset(tail, 0, sizeof(tail));
  memcpy(tail, in + off, rem);
  tail[rem] = 0x80;
  uint64_t tl = (rem < 56) ? 64 : 128;
  uint64_t bits = len * 8;
  for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
  sha256_block(st, tail);
  if (tl == 128) sha256_block(st, tail + 64);
  for (int i = 0; i < 8; i++) { out[4 * i] = st[i] >> 24; out[4 * i + 1] = st[i] >> 16; out[4 * i + 2] = st[i] >> 8; out[4 * i + 3] = st[i]; }
}

static const uint8_t RMD_RL[80] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 7, 4, 13, 1, 10, 6, 15, 3, 12, 0, 9, 5, 2, 14, 11, 8,
                                   3, 10, 14, 4, 9, 15, 8, 1, 2, 7, 0, 6, 13, 11, 5, 12, 1, 9, 11, 10, 0, 8, 12, 4, 13, 3, 7, 15, 14, 5, 6, 2,
                                   4, 0, 5, 9, 7, 12, 2, 10, 14, 1, 3, 8, 11, 6, 15, 13};
static const uint8_t RMD_RR[80] = {5, 14, 7, 0, 9, 2, 11, 4, 13, 6, 15, 8, 1, 10, 3, 12, 6, 11, 3, 7, 0, 13, 5, 10, 14, 15, 8, 12, 4, 9, 1, 2,
                                   15, 5, 1, 3, 7, 14, 6, 9, 11, 8, 12, 2, 10, 0, 4, 13, 8, 6, 4, 1, 3, 11, 15, 0, 5, 12, 2, 13, 9, 7, 10, 14,
                                   12, 15, 10, 4, 1, 5, 8, 7, 6, 2, 13, 14, 0, 3, 9, 11};
static const uint8_t RMD_SL[80] = {11, 14, 15, 12, 5, 8, 7, 9, 11, 13, 14, 15, 6, 7, 9, 8, 7, 6, 8, 13, 11, 9, 7, 15, 7, 12, 15, 9, 11, 7, 13, 12,
                                   11, 13, 6, 7, 14, 9, 13, 15, 14, 8, 13, 6, 5, 12, 7, 5, 11, 12, 14, 15, 14, 15, 9, 8, 9, 14, 5, 6, 8, 6, 5, 12,
                                   9, 15, 5, 11, 6, 8, 13, 12, 5, 12, 13, 14, 11, 8, 5, 6};
static const uint8_t RMD_SR[80] = {8, 9, 9, 11, 13, 15, 15, 5, 7, 7, 8, 11, 14, 14, 12, 6, 9, 13, 15, 7, 12, 8, 9, 11, 7, 7, 12, 7, 6, 15, 13, 11,
                                   9, 7, 15, 11, 8, 6, 6, 14, 12, 13, 5, 14, 13, 13, 7, 5, 15, 5, 8, 11, 14, 14, 6, 14, 6, 9, 12, 9, 12, 5, 15, 8,
                                   8, 5, 12, 9, 12, 5, 14, 6, 8, 13, 6, 5, 15, 13, 11, 11};
static const uint32_t RMD_KL[5] = {0x00000000, 0x5A827999, 0x6ED9EBA1, 0x8F1BBCDC, 0xA953FD4E};
static const uint32_t RMD_KR[5] = {0x50A28BE6, 0x5C4DD124, 0x6D703EF3, 0x7A6D76E9, 0x00000000};
static uint32_t rmd_f(int j, uint32_t x, uint32_t y, uint32_t z) {
  switch (j / 16) {
    case 0: return x ^ y ^ z;
    case 1: return (x & y) | (~x & z);
    case 2: return (x | ~y) ^ z;
    case 3: return (x & z) | (y & ~z);
    default: return x ^ (y | ~z);
  }
}
static void rmd160_block(uint32_t h[5], const uint8_t blk[64]) {
  uint32_t X[16];
  for (int i = 0; i < 16; i++)
    X[i] = (uint32_t)blk[4 * i] | ((uint32_t)blk[4 * i + 1] << 8) | ((uint32_t)blk[4 * i + 2] << 16) | ((uint32_t)blk[4 * i + 3] << 24);
  uint32_t al = h[0], bl = h[1], cl = h[2], dl = h[3], el = h[4];
  uint32_t ar = h[0], br = h[1], cr = h[2], dr = h[3], er = h[4], t;
  for (int j = 0; j < 80; j++) {
    t = ROL32(al + rmd_f(j, bl, cl, dl) + X[RMD_RL[j]] + RMD_KL[j / 16], RMD_SL[j]) + el;
    al = el; el = dl; dl = ROL32(cl, 10); cl = bl; bl = t;
    t = ROL32(ar + rmd_f(79 - j, br, cr, dr) + X[RMD_RR[j]] + RMD_KR[j / 16], RMD_SR[j]) + er;
    ar = er; er = dr; dr = ROL32(cr, 10); cr = br; br = t;
  }
  t = h[1] + cl + dr; h[1] = h[2] + dl + er; h[2] = h[3] + el + ar; h[3] = h[4] + al + br; h[4] = h[0] + bl + cr; h[0] = t;
}
void kho_ripemd160(const uint8_t *in, uint64_t len, uint8_t out[20]) {
  uint32_t h[5] = {0x67452301, 0xEFCDAB89, 0x98BADCFE, 0x10325476, 0xC3D2E1F0};
  uint64_t off = 0;
  for (; off + 64 <= len; off += 64) rmd160_block(h, in + off);
  uint8_t tail[128];
  uint64_t rem = len - off;
  memset(tail, 0, sizeof(tail));
  memcpy(tail, in + off, rem);
  tail[rem] = 0x80;
  uint64_t tl = (rem < 56) ? 64 : 128;
  uint64_t bits = len * 8;
  for (int i = 0; i < 8; i++) tail[tl - 8 + i] = (uint8_t)(bits >> (8 * i));
  rmd160_block(h, tail);
  if (tl == 128) rmd160_block(h, tail + 64);
  for (int i = 0; i < 5; i++) { out[4 * i] = h[i]; out[4 * i + 1] = h[i] >> 8; out[4 * i + 2] = h[i] >> 16; out[4 * i + 3] = h[i] >> 24; }
}

static void hash160(const uint8_t *msg, uint64_t len, uint8_t out[20]) {
  uint8_t d[32];
  kho_sha256(msg, len, d);
  kho_ripemd160(d, 32, out);
}
static void hash160_comp_fe(int prefix, const fe *x, uint8_t out[20]) { /* GetHash160_fromX SECP256K1.cpp:1207 */
  uint8_t m[33];
  m[0] = (uint8_t)prefix;
  fe_to_be(m + 1, x);
  hash160(m, 33, out);
}
static void hash160_uncomp_ge(const ge *p, uint8_t out[20]) { /* GetHash160(...,false,...) SECP256K1.cpp:1045 */
  uint8_t m[65];
  m[0] = 0x04;
  fe_to_be(m + 1, &p->x);
  fe_to_be(m + 33, &p->y);
  hash160(m, 65, out);
}

static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800AULL, 0x800000008000000AULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KECCAK_PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
#define ROL64(x, n) (((x) << (n)) | ((x) >> (64 - (n))))
static void keccakf(uint64_t st[25]) { /* keccakf1600 sha3/keccak.c:144 */
  uint64_t bc[5], t;
  for (int r = 0; r < 24; r++) {
    for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
    for (int i = 0; i < 5; i++) {
      t = bc[(i + 4) % 5] ^ ROL64(bc[(i + 1) % 5], 1);
      for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
    }
    t = st[1];
    for (int i = 0; i < 24; i++) {
      int j = KECCAK_PILN[i];
      bc[0] = st[j];
      st[j] = ROL64(t, KECCAK_ROTC[i]);
      t = bc[0];
    }
    for (int j = 0; j < 25; j += 5) {
      for (int i = 0; i < 5; i++) bc[i] = st[j + i];
      for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
    }
    st[0] ^= KECCAK_RC[r];
  }
}
static void eth_addr_ge(const ge *p, uint8_t out[20]) { /* generate_binaddress_eth keyhunt.cpp:5663; KECCAK_256_Final sha3.c:414 */
  uint8_t m[64];
  uint64_t st[25];
  fe_to_be(m, &p->x);
  fe_to_be(m + 32, &p->y);
  memset(st, 0, sizeof(st));
  for (int i = 0; i < 8; i++) {
    uint64_t v = 0;
    for (int j = 7; j >= 0; j--) v = (v << 8) | m[8 * i + j];
    st[i] = v;
  }
  st[8] ^= 0x01ULL;                    /* Keccak (pre-NIST) domain padding, rate 136 bytes = 17 lanes */
  st[16] ^= 0x8000000000000000ULL;
  keccakf(st);
  uint8_t d[32];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) d[8 * i + j] = (uint8_t)(st[i] >> (8 * j));
  memcpy(out, d + 12, 20);
}

#define XP1 0x9E3779B185EBCA87ULL
#define XP2 0xC2B2AE3D27D4EB4FULL
#define XP3 0x165667B19E3779F9ULL
#define XP4 0x85EBCA77C2B2AE63ULL
#define XP5 0x27D4EB2F165667C5ULL
static uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t xxh_round(uint64_t acc, uint64_t in) { acc += in * XP2; acc = ROL64(acc, 31); return acc * XP1; }
static uint64_t xxh_merge(uint64_t acc, uint64_t v) { v = xxh_round(0, v); acc ^= v; return acc * XP1 + XP4; }
uint64_t kho_xxh64(const void *buf, uint64_t len, uint64_t seed) { /* XXH64 xxhash.h:2512 / :2469 / :2334 */
  const uint8_t *p = (const uint8_t *)buf, *end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
    do {
      v1 = xxh_round(v1, rd64(p)); v2 = xxh_round(v2, rd64(p + 8));
      v3 = xxh_round(v3, rd64(p + 16)); v4 = xxh_round(v4, rd64(p + 24));
      p += 32;
    } while (p + 32 <= end);
    h = ROL64(v1, 1) + ROL64(v2, 7) + ROL64(v3, 12) + ROL64(v4, 18);
    h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
  } else {
    h = seed + XP5;
  }
  h += len;
  while (p + 8 <= end) { h ^= xxh_round(0, rd64(p)); h = ROL64(h, 27) * XP1 + XP4; p += 8; }
  if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * XP1; h = ROL64(h, 23) * XP2 + XP3; p += 4; }
  while (p < end) { h ^= (*p) * XP5; h = ROL64(h, 11) * XP1; p++; }
  h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
  return h;
}

/* ------------------------------------------------------------------------------------------------
 * Bloom filter (bloom/bloom.cpp)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t entries, bits, bytes; uint32_t hashes; uint8_t *bf; } obloom;

void *kho_bloom_new(uint64_t entries) { /* bloom_init2 bloom.cpp:154-187, called with the double literal 0.000001 (keyhunt.cpp:7620) */
  long double error = 0.000001;
  if (entries < 1000) return NULL;
  obloom *b = (obloom *)calloc(1, sizeof(obloom));
  long double num = -logl(error);               /* C++ overload log(long double) */
  long double denom = 0.480453013918201;
  double bpe = (double)(num / denom);           /* struct field is a double (bloom.h:45) */
  long double allbits = (long double)entries * bpe;
  b->entries = entries;
  b->bits = (uint64_t)allbits;
  b->bytes = b->bits / 8 + ((b->bits % 8) ? 1 : 0);
  b->hashes = (uint8_t)ceil(0.693147180559945 * bpe);
  b->bf = (uint8_t *)calloc(b->bytes, 1);
  return b;
}
void kho_bloom_free(void *h) { if (h) { free(((obloom *)h)->bf); free(h); } }
void kho_bloom_desc(void *h, uint64_t *entries, uint64_t *bits, uint64_t *bytes, uint32_t *hashes) {
  obloom *b = (obloom *)h;
  *entries = b->entries; *bits = b->bits; *bytes = b->bytes; *hashes = b->hashes;
}
uint8_t *kho_bloom_data(void *h) { return ((obloom *)h)->bf; }
static int bloom_check_add(obloom *b, const void *buf, int len, int add) { /* bloom.cpp:122-146 / :189-212 */
  uint64_t a = kho_xxh64(buf, (uint64_t)len, 0x59f2815b16f81798ULL);
  uint64_t bb = kho_xxh64(buf, (uint64_t)len, a);
  uint32_t hits = 0;
  for (uint32_t i = 0; i < b->hashes; i++) {
    uint64_t x = (a + bb * i) % b->bits;       /* wrapping u64, then reduce */
    uint8_t mask = (uint8_t)(1u << (x & 7));
    if (b->bf[x >> 3] & mask) hits++;
    else if (add) b->bf[x >> 3] |= mask;
    else return 0;
  }
  return hits == b->hashes;
}
int kho_bloom_add(void *h, const void *buf, int len) { return bloom_check_add((obloom *)h, buf, len, 1); }
int kho_bloom_check(void *h, const void *buf, int len) { return bloom_check_add((obloom *)h, buf, len, 0); }

/* ------------------------------------------------------------------------------------------------
 * Exported primitive wrappers
 * ---------------------------------------------------------------------------------------------- */
void kho_fe_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { fe x, y, r; fe_from_be(&x, a); fe_from_be(&y, b); fe_mul(&r, &x, &y); fe_to_be(out, &r); }
void kho_fe_sqr(const uint8_t a[32], uint8_t out[32]) { fe x, r; fe_from_be(&x, a); fe_sqr(&r, &x); fe_to_be(out, &r); }
void kho_fe_inv(const uint8_t a[32], uint8_t out[32]) { fe x, r; fe_from_be(&x, a); fe_inv(&r, &x); fe_to_be(out, &r); }
void kho_fe_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { fe x, y, r; fe_from_be(&x, a); fe_from_be(&y, b); fe_add(&r, &x, &y); fe_to_be(out, &r); }
void kho_fe_sub(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { fe x, y, r; fe_from_be(&x, a); fe_from_be(&y, b); fe_sub(&r, &x, &y); fe_to_be(out, &r); }
void kho_fe_neg(const uint8_t a[32], uint8_t out[32]) { fe x, r; fe_from_be(&x, a); fe_neg(&r, &x); fe_to_be(out, &r); }
void kho_pubkey(const uint8_t key[32], uint8_t xy[64]) { fe k; ge p; fe_from_be(&k, key); ge_scalar_mul(&p, &GE_G, &k); fe_to_be(xy, &p.x); fe_to_be(xy + 32, &p.y); }
static void ge_from_be(ge *p, const uint8_t xy[64]) { fe_from_be(&p->x, xy); fe_from_be(&p->y, xy + 32); }
static void ge_to_be(uint8_t xy[64], const ge *p) { fe_to_be(xy, &p->x); fe_to_be(xy + 32, &p->y); }
void kho_add_direct(const uint8_t a[64], const uint8_t b[64], uint8_t out[64]) { ge p, q, r; ge_from_be(&p, a); ge_from_be(&q, b); ge_add_direct(&r, &p, &q); ge_to_be(out, &r); }
void kho_hash160_comp(int prefix, const uint8_t x[32], uint8_t out[20]) { fe v; fe_from_be(&v, x); hash160_comp_fe(prefix, &v, out); }
void kho_hash160_uncomp(const uint8_t xy[64], uint8_t out[20]) { ge p; ge_from_be(&p, xy); hash160_uncomp_ge(&p, out); }
void kho_eth_addr(const uint8_t xy[64], uint8_t out[20]) { ge p; ge_from_be(&p, xy); eth_addr_ge(&p, out); }

void kho_batch_points(const uint8_t base_key[32], const uint8_t stride_be[32], int with_y, uint8_t *out) {
  fe key, stride, t;
  ge sg, c;
  gtable *tab = (gtable *)malloc(sizeof(gtable));
  ge *pts = (ge *)malloc(sizeof(ge) * GRP);
  fe_from_be(&key, base_key);
  fe_from_be(&stride, stride_be);
  ge_scalar_mul(&sg, &GE_G, &stride);
  gtable_init(tab, &sg);
  u256_mul_u64(&t, &stride, HALF);
  u256_add(&key, &key, &t);
  ge_scalar_mul(&c, &GE_G, &key);            /* keyhunt.cpp:3349-3353 */
  batch_around(tab, &c, with_y, pts, NULL);
  for (int i = 0; i < GRP; i++) ge_to_be(out + 64 * i, &pts[i]);
  free(pts); free(tab);
}

/* ------------------------------------------------------------------------------------------------
 * Targets: bloom + sorted table (readFileAddress keyhunt.cpp:7033.., _sort :4307, searchbinary :3065)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t N; uint8_t *table; obloom *bloom;
  /* vanity (-m vanity): interval pairs instead of the table, bloom over the first vmin bytes of every A limit */
  uint64_t vn; uint8_t *va, *vb; int vmin; obloom *vbloom;
} otargets;
static int cmp20(const void *a, const void *b) { return memcmp(a, b, 20); }
void *kho_targets_new(const uint8_t *raw20, uint64_t N) {
  otargets *t = (otargets *)calloc(1, sizeof(otargets));
  t->N = N;
  t->table = (uint8_t *)malloc(N ? N * 20 : 20);
  memcpy(t->table, raw20, N * 20);
  t->bloom = (obloom *)kho_bloom_new(N <= 10000 ? 10000 : N); /* initBloomFilter keyhunt.cpp:7608 (no -z) */
  for (uint64_t i = 0; i < N; i++) kho_bloom_add(t->bloom, raw20 + 20 * i, 20);
  qsort(t->table, N, 20, cmp20);
  return t;
}
void kho_targets_free(void *h) {
  otargets *t = (otargets *)h;
  if (t) { kho_bloom_free(t->bloom); kho_bloom_free(t->vbloom); free(t->table); free(t->va); free(t->vb); free(t); }
}
void *kho_targets_bloom(void *h) { return ((otargets *)h)->bloom; }
const uint8_t *kho_targets_table(void *h, uint64_t *N) { *N = ((otargets *)h)->N; return ((otargets *)h)->table; }
int kho_searchbinary(void *h, const uint8_t data[20]) { /* keyhunt.cpp:3065-3089, same probe sequence */
  otargets *t = (otargets *)h;
  int64_t half, min = 0, max = (int64_t)t->N, current = 0;
  int r = 0;
  half = (int64_t)t->N;
  while (!r && half >= 1) {
    half = (max - min) / 2;
    int rc = memcmp(data, t->table + 20 * (current + half), 20);
    if (rc == 0) r = 1;
    else {
      if (rc < 0) max = max - half; else min = min + half;
      current = min;
    }
  }
  return r;
}

/* ------------------------------------------------------------------------------------------------
 * Vanity targets (-m vanity): addvanity keyhunt.cpp:6739-6860, b58tobin base58/base58.c:39-110,
 * processOneVanity / readFileVanity :6971-7030, vanityrmdmatch :6677
 * ---------------------------------------------------------------------------------------------- */
static const char B58_DIGITS[] = "123456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz";
/* the reference's decoder: the number fills the WHOLE binsz-byte buffer big-endian (right-aligned); *binsz becomes
 * (bytes after the leading zero bytes) + (number of leading '1' digits); 0 = invalid digit or overflow */
int kho_b58tobin(uint8_t *bin, uint64_t *binsz, const char *b58, uint64_t b58sz) {
  uint64_t n = *binsz, i = 0, zeros = 0;
  if (!b58sz) b58sz = strlen(b58);
  memset(bin, 0, n);
  while (i < b58sz && b58[i] == '1') { zeros++; i++; }
  for (; i < b58sz; i++) {
    const char *d = (b58[i] & 0x80) ? NULL : strchr(B58_DIGITS, b58[i]);
    if (!d || !b58[i]) return 0;
    uint32_t carry = (uint32_t)(d - B58_DIGITS);
    for (uint64_t j = n; j-- > 0;) { uint32_t t = (uint32_t)bin[j] * 58u + carry; bin[j] = (uint8_t)t; carry = t >> 8; }
    if (carry) return 0;
  }
  uint64_t lead = 0;
  while (lead < n && !bin[lead]) lead++;
  *binsz = n - lead + zeros;
  return 1;
}
/* one side of addvanity: pad the prefix with `fill` until it decodes to more than 25 bytes; every 25-byte decode
 * contributes bytes 1..20 (the hash160 between version byte and checksum) */
static int vanity_side(const char *target, char fill, uint8_t *out, int max_r) {
  char copy[64];
  uint8_t raw[50];
  int size = (int)strlen(target), j = 0;
  memset(copy, 0, sizeof(copy));
  memcpy(copy, target, (size_t)size);
  uint64_t len;
  do {
    len = 50;
    kho_b58tobin(raw, &len, copy, (uint64_t)size);           /* a failed decode leaves len = 50 and ends the loop */
    if (len < 25) copy[size++] = fill;
    if (len == 25) {
      uint64_t l2 = 25;
      kho_b58tobin(raw, &l2, copy, (uint64_t)size);          /* second decode into exactly 25 bytes: version|hash160|checksum */
      if (j < max_r) memcpy(out + 20 * j, raw + 1, 20);
      j++;
      copy[size++] = fill;
    }
  } while (len <= 25 && size < 60);
  return j;
}
/* returns r = number of [A,B] pairs for this prefix (0 = not added), *min_bytes lowered like :6833-6836 */
int kho_addvanity(const char *target, uint8_t *A, uint8_t *B, int max_r, int *min_bytes) {
  if (strlen(target) >= 30) return 0;
  uint8_t a[20 * 16], b[20 * 16];
  int na = vanity_side(target, '1', a, 16), nb = vanity_side(target, 'z', b, 16);
  if (na < 1 || nb < 1) return 0;
  int r = na < nb ? na : nb;
  if (r > 16) r = 16;
  if (r > max_r) r = max_r;
  for (int j = 0; j < r; j++) {
    int same = 0;
    while (same < 20 && a[20 * j + same] == b[20 * j + same]) same++;
    if (min_bytes && same < *min_bytes) *min_bytes = same;
    memcpy(A + 20 * j, a + 20 * j, 20);
    memcpy(B + 20 * j, b + 20 * j, 20);
  }
  return r;
}
/* targets handle for a vanity search: n interval pairs (flattened 20-byte limits), bloom of n entries over the first
 * min_bytes bytes of every A limit (processOneVanity :6971) */
void *kho_targets_new_vanity(const uint8_t *A, const uint8_t *B, uint64_t n, int min_bytes) {
  otargets *t = (otargets *)calloc(1, sizeof(otargets));
  t->vn = n; t->vmin = min_bytes;
  t->va = (uint8_t *)malloc(n ? n * 20 : 20); t->vb = (uint8_t *)malloc(n ? n * 20 : 20);
  memcpy(t->va, A, n * 20); memcpy(t->vb, B, n * 20);
  t->vbloom = (obloom *)kho_bloom_new(n <= 10000 ? 10000 : n);     /* initBloomFilterMapped -> initBloomFilter :7608 */
  for (uint64_t i = 0; i < n; i++) kho_bloom_add(t->vbloom, A + 20 * i, min_bytes);
  return t;
}

/* ------------------------------------------------------------------------------------------------
 * Scan (thread_process keyhunt.cpp:3265-3861)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  otargets *t; int mode, crypto, search, endo;
  fe start, stride; const gtable *tab;
  uint64_t batch0, batch1;
  kho_hit *hits; uint64_t nhits, cap;
} scan_job;

static void job_push_v(scan_job *j, const fe *key, const uint8_t m[20], int kind, uint64_t index, int variant);
static void job_push(scan_job *j, const fe *key, const uint8_t m[20], int kind, uint64_t index) { job_push_v(j, key, m, kind, index, 0); }
static void job_push_v(scan_job *j, const fe *key, const uint8_t m[20], int kind, uint64_t index, int variant) {
  if (j->nhits == j->cap) {
    j->cap = j->cap ? j->cap * 2 : 16;
    j->hits = (kho_hit *)realloc(j->hits, j->cap * sizeof(kho_hit));
  }
  kho_hit *h = &j->hits[j->nhits++];
  memset(h, 0, sizeof(*h));
  fe_to_be(h->key_be, key);
  memcpy(h->matched, m, 20);
  h->kind = (uint8_t)kind;
  h->pad[0] = (uint8_t)variant;
  h->index = index;
}
static int probe(otargets *t, const uint8_t h[20]) { /* bloom_check then searchbinary, keyhunt.cpp:3621-3624 */
  if (t->vn) {                                      /* vanityrmdmatch keyhunt.cpp:6677-6703 */
    if (!kho_bloom_check(t->vbloom, h, t->vmin)) return 0;
    for (uint64_t i = 0; i < t->vn; i++)
      if (memcmp(t->va + 20 * i, h, 20) <= 0 && memcmp(t->vb + 20 * i, h, 20) >= 0) return 1;
    return 0;
  }
  return kho_bloom_check(t->bloom, h, 20) && kho_searchbinary(t, h);
}

/* -e: the six (three for xpoint) endomorphic candidates of one point and their hit fix-ups, exactly as
 * thread_process does them (keyhunt.cpp:3408-3473 candidates, :3483-3516/:3525-3536 hashes, :3557-3617,
 * :3643-3686, :3704-3749, :3769-3807 fix-ups) — including the ETH slot 4 that hashes the beta point again
 * (:3534) and therefore reports a key that does not own the address. variant = the reference's index l. */
static void scan_point_endo(scan_job *j, const ge *p, const fe *key, uint64_t index) {
  uint8_t h[20], h2[20];
  ge q[3];
  q[0] = *p;
  q[1].y = p->y; fe_mul(&q[1].x, &p->x, &FE_BETA);
  q[2].y = p->y; fe_mul(&q[2].x, &p->x, &FE_BETA2);
  const fe *lam[3] = {NULL, &SC_LAMBDA, &SC_LAMBDA2};
  if (j->mode == KHO_MODE_XPOINT) {
    for (int v = 0; v < 3; v++) {
      uint8_t x[32];
      fe_to_be(x, &q[v].x);
      if (probe(j->t, x)) { fe kk = *key; if (v) sc_mul(&kk, &kk, lam[v]); job_push_v(j, &kk, x, KHO_HIT_XPOINT, index, v); }
    }
    return;
  }
  if (j->crypto == KHO_CRYPTO_ETH) {
    for (int l = 0; l < 6; l++) {
      ge c = q[l / 2];
      if (l == 4) c = q[1];                                   /* :3534 hashes endomorphism_beta again */
      if (l & 1) { if (l == 5) c = q[2]; fe_neg(&c.y, &c.y); }
      eth_addr_ge(&c, h);
      if (!probe(j->t, h)) continue;
      fe kk = *key; ge pub;
      if (l >= 2) sc_mul(&kk, &kk, lam[l / 2]);
      ge_scalar_mul(&pub, &GE_G, &kk);
      eth_addr_ge(&pub, h2);
      if (memcmp(h, h2, 20) != 0) u256_sub(&kk, &FE_N, &kk);
      job_push_v(j, &kk, h, KHO_HIT_ETH, index, l);
    }
    return;
  }
  ge pub0;                                                    /* publickey = ComputePublicKey(keyfound) before any lambda */
  int have_pub0 = 0;
  if (j->search == KHO_SEARCH_COMPRESS || j->search == KHO_SEARCH_BOTH) {
    for (int l = 0; l < 6; l++) {
      hash160_comp_fe((l & 1) ? 0x03 : 0x02, &q[l / 2].x, h);
      if (!probe(j->t, h)) continue;
      fe kk = *key;
      if (!have_pub0) { ge_scalar_mul(&pub0, &GE_G, &kk); have_pub0 = 1; }
      int odd = (int)(pub0.y.l[0] & 1);
      if (l >= 2) sc_mul(&kk, &kk, lam[l / 2]);
      if (((l & 1) == 0 && odd) || ((l & 1) == 1 && !odd)) u256_sub(&kk, &FE_N, &kk);
      job_push_v(j, &kk, h, (l & 1) ? KHO_HIT_COMP03 : KHO_HIT_COMP02, index, l);
    }
  }
  if (j->search == KHO_SEARCH_UNCOMPRESS || j->search == KHO_SEARCH_BOTH) {
    for (int l = 6; l < 12; l++) {
      ge c = q[(l - 6) / 2];
      if (l & 1) fe_neg(&c.y, &c.y);
      hash160_uncomp_ge(&c, h);
      if (!probe(j->t, h)) continue;
      fe kk = *key; ge pub;
      if (l >= 8) sc_mul(&kk, &kk, lam[(l - 6) / 2]);
      ge_scalar_mul(&pub, &GE_G, &kk);
      hash160_uncomp_ge(&pub, h2);
      if (memcmp(h, h2, 20) != 0) u256_sub(&kk, &FE_N, &kk);
      job_push_v(j, &kk, h, KHO_HIT_UNCOMP, index, l);
    }
  }
}

static void *scan_worker(void *arg) {
  scan_job *j = (scan_job *)arg;
  ge *pts = (ge *)malloc(sizeof(ge) * GRP);
  int need_y = (j->crypto == KHO_CRYPTO_ETH) ||
               (j->mode != KHO_MODE_XPOINT && (j->search == KHO_SEARCH_UNCOMPRESS || j->search == KHO_SEARCH_BOTH)); /* :3294 */
  /* centre of the first batch: start + (batch0*1024 + 512)*stride */
  fe k, t;
  u256_mul_u64(&t, &j->stride, HALF);
  fe off;
  { /* off = batch0*1024*stride */
    fe s1024; u256_mul_u64(&s1024, &j->stride, GRP);
    u256_mul_u64(&off, &s1024, j->batch0);
  }
  u256_add(&k, &j->start, &off);
  fe base = k;                       /* key of pts[0] in this batch */
  u256_add(&k, &k, &t);
  ge c, next;
  ge_scalar_mul(&c, &GE_G, &k);
  fe s1024; u256_mul_u64(&s1024, &j->stride, GRP);
  for (uint64_t b = j->batch0; b < j->batch1; b++) {
    batch_around(j->tab, &c, need_y, pts, &next);
    for (int i = 0; i < GRP; i++) {
      uint8_t h[20], h2[20];
      fe key, ik;
      u256_mul_u64(&ik, &j->stride, (uint64_t)i);
      u256_add(&key, &base, &ik);                                  /* keyfound = k*stride + key_mpz :3625-3627 */
      uint64_t index = b * GRP + (uint64_t)i;
      if (j->endo) { scan_point_endo(j, &pts[i], &key, index); continue; }
      if (j->mode == KHO_MODE_XPOINT) {                            /* :3810-3821 */
        uint8_t x[32];
        fe_to_be(x, &pts[i].x);
        if (probe(j->t, x)) job_push(j, &key, x, KHO_HIT_XPOINT, index);
      } else if (j->crypto == KHO_CRYPTO_ETH) {                    /* :3540, :3752-3763 */
        eth_addr_ge(&pts[i], h);
        if (probe(j->t, h)) job_push(j, &key, h, KHO_HIT_ETH, index);
      } else {
        if (j->search == KHO_SEARCH_COMPRESS || j->search == KHO_SEARCH_BOTH) { /* :3493-3494, :3620-3638 */
          for (int l = 0; l < 2; l++) {
            hash160_comp_fe(l == 0 ? 0x02 : 0x03, &pts[i].x, h);
            if (probe(j->t, h)) {
              /* fix-up: recompute the real compressed hash160 of keyfound; if it is not the matched
               * one the target belongs to n - keyfound (:3629-3634) */
              ge pub; fe kk = key;
              ge_scalar_mul(&pub, &GE_G, &kk);
              hash160_comp_fe((pub.y.l[0] & 1) ? 0x03 : 0x02, &pub.x, h2);
              if (memcmp(h, h2, 20) != 0) u256_sub(&kk, &FE_N, &kk);
              job_push(j, &kk, h, l == 0 ? KHO_HIT_COMP02 : KHO_HIT_COMP03, index);
            }
          }
        }
        if (j->search == KHO_SEARCH_UNCOMPRESS || j->search == KHO_SEARCH_BOTH) { /* :3519, :3689-3698 */
          hash160_uncomp_ge(&pts[i], h);
          if (probe(j->t, h)) job_push(j, &key, h, KHO_HIT_UNCOMP, index);
        }
      }
    }
    c = next;
    u256_add(&base, &base, &s1024);
  }
  free(pts);
  return NULL;
}

int64_t kho_scan_ex(void *targets, int mode, int crypto, int search, int endo, const uint8_t start[32], const uint8_t stride_be[32],
                    uint64_t n_points, kho_hit *hits, uint64_t max_hits, int nthreads) {
  if (n_points % GRP) return -1;
  uint64_t nb = n_points / GRP;
  if (nthreads < 1) nthreads = 1;
  if ((uint64_t)nthreads > nb) nthreads = (int)(nb ? nb : 1);
  gtable *tab = (gtable *)malloc(sizeof(gtable));
  fe stride; ge sg;
  fe_from_be(&stride, stride_be);
  ge_scalar_mul(&sg, &GE_G, &stride);
  gtable_init(tab, &sg);
  scan_job *jobs = (scan_job *)calloc((size_t)nthreads, sizeof(scan_job));
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int i = 0; i < nthreads; i++) {
    scan_job *j = &jobs[i];
    j->t = (otargets *)targets; j->mode = mode; j->crypto = crypto; j->search = search; j->endo = endo;
    fe_from_be(&j->start, start); j->stride = stride; j->tab = tab;
    j->batch0 = nb * (uint64_t)i / (uint64_t)nthreads;
    j->batch1 = nb * (uint64_t)(i + 1) / (uint64_t)nthreads;
    if (nthreads == 1) scan_worker(j); else pthread_create(&th[i], NULL, scan_worker, j);
  }
  int64_t total = 0;
  for (int i = 0; i < nthreads; i++) {
    if (nthreads > 1) pthread_join(th[i], NULL);
    for (uint64_t h = 0; h < jobs[i].nhits; h++) {
      if ((uint64_t)total < max_hits) hits[total] = jobs[i].hits[h];
      total++;
    }
    free(jobs[i].hits);
  }
  free(jobs); free(th); free(tab);
  return total;
}

int64_t kho_scan(void *targets, int mode, int crypto, int search, const uint8_t start[32], const uint8_t stride_be[32],
                 uint64_t n_points, kho_hit *hits, uint64_t max_hits, int nthreads) {
  return kho_scan_ex(targets, mode, crypto, search, 0, start, stride_be, n_points, hits, max_hits, nthreads);
}

/* ------------------------------------------------------------------------------------------------
 * BSGS (keyhunt.cpp:1450-1842 parameters, :5284 thread_bPload, :4412 bsgs_sort, :4549 search)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t n, m, m2, m3, aux;
  obloom *tier[3][256];
  kho_bp_entry *table;
  gtable gs;            /* GSn[i] = -(i+1)*2m*G, _2GSn = -2048*m*G  (:1803-1816) */
  ge amp2[32], amp3[32];/* -(2i+1)*m2*G, -(2i+1)*m3*G (:1818-1842) */
} obsgs;

static int cmp_bp(const void *a, const void *b) {
  const kho_bp_entry *x = (const kho_bp_entry *)a, *y = (const kho_bp_entry *)b;
  int r = memcmp(x->value, y->value, 6);
  if (r) return r;
  return (x->index > y->index) - (x->index < y->index);   /* canonical tie order */
}

typedef struct { obsgs *b; uint64_t from, to; pthread_mutex_t *mtx; } bp_job;
static void *bp_worker(void *arg) { /* thread_bPload keyhunt.cpp:5284-5472 */
  bp_job *j = (bp_job *)arg;
  obsgs *b = j->b;
  gtable *tab = (gtable *)malloc(sizeof(gtable));
  ge *pts = (ge *)malloc(sizeof(ge) * GRP);
  gtable_init(tab, &GE_G);
  fe km; u256_set_u64(&km, j->from + 1 + HALF);
  ge c, next;
  ge_scalar_mul(&c, &GE_G, &km);
  uint64_t i_counter = j->from;
  uint64_t nb = (j->to - j->from + GRP - 1) / GRP;
  for (uint64_t s = 0; s < nb; s++) {
    batch_around(tab, &c, 0, pts, &next);
    for (int q = 0; q < GRP; q++, i_counter++) {
      uint8_t x[32];
      fe_to_be(x, &pts[q].x);
      int shard = x[0];
      pthread_mutex_lock(&j->mtx[shard]);
      if (i_counter < b->m3) {
        memcpy(b->table[i_counter].value, x + 16, 6);
        b->table[i_counter].index = i_counter;
        kho_bloom_add(b->tier[2][shard], x, 32);
      }
      if (i_counter < b->m2) kho_bloom_add(b->tier[1][shard], x, 32);
      if (i_counter < j->to) kho_bloom_add(b->tier[0][shard], x, 32);
      pthread_mutex_unlock(&j->mtx[shard]);
    }
    c = next;
  }
  free(pts); free(tab);
  return NULL;
}

void *kho_bsgs_new(uint64_t n, uint32_t k, int nthreads) {
  /* n must be 2^even (exact sqrt, :1474) and sqrt(n) a multiple of 1024 (:1509) */
  int lg = 0;
  while (lg < 63 && (1ULL << lg) < n) lg++;
  if ((1ULL << lg) != n || (lg & 1) || lg < 20 || k < 1) return NULL;
  obsgs *b = (obsgs *)calloc(1, sizeof(obsgs));
  uint64_t m = (1ULL << (lg / 2)) * k;                       /* :1557 */
  b->m = m;
  b->m2 = m / 32 + ((m % 32) ? 1 : 0);                       /* :1561-1566 */
  b->m3 = b->m2 / 32 + ((b->m2 % 32) ? 1 : 0);               /* :1578-1583 */
  b->aux = n / m;                                            /* :1591-1603 */
  b->n = (n % m) ? m * b->aux : n;
  uint64_t items1 = (m / 256 > 10000) ? (m / 256 + ((m % 256) ? 1 : 0)) : 1000;          /* :1633-1641 */
  uint64_t items2 = (b->m2 / 256 > 1000) ? (b->m2 / 256 + ((b->m2 % 256) ? 1 : 0)) : 1000;
  uint64_t items3 = (b->m3 / 256 > 1000) ? (b->m3 / 256 + ((b->m3 % 256) ? 1 : 0)) : 1000;
  uint64_t items[3] = {items1, items2, items3};
  for (int t = 0; t < 3; t++)
    for (int s = 0; s < 256; s++)
      b->tier[t][s] = (obloom *)kho_bloom_new(items[t] <= 10000 ? 10000 : items[t]);    /* initBloomFilter :7608 */
  b->table = (kho_bp_entry *)calloc(b->m3, sizeof(kho_bp_entry));
  /* giant-step tables */
  fe s; ge p, np;
  u256_set_u64(&s, 2); u256_mul_u64(&s, &s, m);              /* 2m */
  ge_scalar_mul(&p, &GE_G, &s); ge_neg(&np, &p);
  gtable_init(&b->gs, &np);
  u256_set_u64(&s, b->m2); ge_scalar_mul(&p, &GE_G, &s); ge_neg(&b->amp2[0], &p);
  u256_set_u64(&s, 2 * b->m2); ge_scalar_mul(&p, &GE_G, &s); ge_neg(&np, &p);
  for (int i = 1; i < 32; i++) ge_add_direct(&b->amp2[i], &b->amp2[i - 1], &np);
  u256_set_u64(&s, b->m3); ge_scalar_mul(&p, &GE_G, &s); ge_neg(&b->amp3[0], &p);
  u256_set_u64(&s, 2 * b->m3); ge_scalar_mul(&p, &GE_G, &s); ge_neg(&np, &p);
  for (int i = 1; i < 32; i++) ge_add_direct(&b->amp3[i], &b->amp3[i - 1], &np);
  /* baby steps */
  if (nthreads < 1) nthreads = 1;
  uint64_t nbatch = m / GRP;
  if ((uint64_t)nthreads > nbatch) nthreads = (int)nbatch;
  pthread_mutex_t mtx[256];
  for (int i = 0; i < 256; i++) pthread_mutex_init(&mtx[i], NULL);
  bp_job *jobs = (bp_job *)calloc((size_t)nthreads, sizeof(bp_job));
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int i = 0; i < nthreads; i++) {
    jobs[i].b = b; jobs[i].mtx = mtx;
    jobs[i].from = (nbatch * (uint64_t)i / (uint64_t)nthreads) * GRP;
    jobs[i].to = (nbatch * (uint64_t)(i + 1) / (uint64_t)nthreads) * GRP;
    pthread_create(&th[i], NULL, bp_worker, &jobs[i]);
  }
  for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  free(jobs); free(th);
  qsort(b->table, b->m3, sizeof(kho_bp_entry), cmp_bp);      /* bsgs_sort :4412 (ties canonicalised) */
  return b;
}
void kho_bsgs_free(void *h) {
  obsgs *b = (obsgs *)h;
  if (!b) return;
  for (int t = 0; t < 3; t++) for (int s = 0; s < 256; s++) kho_bloom_free(b->tier[t][s]);
  free(b->table); free(b);
}
void kho_bsgs_params(void *h, uint64_t *n, uint64_t *m, uint64_t *m2, uint64_t *m3, uint64_t *aux) {
  obsgs *b = (obsgs *)h; *n = b->n; *m = b->m; *m2 = b->m2; *m3 = b->m3; *aux = b->aux;
}
void *kho_bsgs_bloom(void *h, int tier, int shard) { return ((obsgs *)h)->tier[tier - 1][shard]; }
const kho_bp_entry *kho_bsgs_table(void *h) { return ((obsgs *)h)->table; }

/* bsgs_searchbinary keyhunt.cpp:4510, but checking every entry that shares the 6-byte key */
static int bp_lookup_all(obsgs *b, const uint8_t x[32], uint64_t *idx, int max) {
  int64_t lo = 0, hi = (int64_t)b->m3;
  while (lo < hi) { int64_t mid = (lo + hi) / 2; if (memcmp(b->table[mid].value, x + 16, 6) < 0) lo = mid + 1; else hi = mid; }
  int cnt = 0;
  while (lo < (int64_t)b->m3 && memcmp(b->table[lo].value, x + 16, 6) == 0 && cnt < max) idx[cnt++] = b->table[lo++].index;
  return cnt;
}
/* key = base + k*mult (256-bit) */
static void key_add_mul(fe *r, const fe *base, uint64_t mult, uint64_t k) {
  u128 v = (u128)mult * k;
  fe t; t.l[0] = (uint64_t)v; t.l[1] = (uint64_t)(v >> 64); t.l[2] = t.l[3] = 0;
  u256_add(r, base, &t);
}
static int bsgs_thirdcheck(obsgs *b, const ge *Q, const fe *base_key, fe *priv) { /* :5186-5248; base_key = a*2*m2 + start_range (:5194-5196) */
  ge bp, nbp, S, P3;
  ge_scalar_mul(&bp, &GE_G, base_key); ge_neg(&nbp, &bp);
  ge_add_direct(&S, Q, &nbp);
  for (int i = 0; i < 32; i++) {
    ge_add_direct(&P3, &S, &b->amp3[i]);
    uint8_t x[32];
    fe_to_be(x, &P3.x);
    fe calc; u256_set_u64(&calc, (2ULL * (uint64_t)i + 1) * b->m3);            /* calcualteindex :7859 */
    if (kho_bloom_check(b->tier[2][x[0]], x, 32)) {
      uint64_t idx[8];
      int c = bp_lookup_all(b, x, idx, 8);
      for (int q = 0; q < c; q++) {
        fe cand, j1; ge pub;
        u256_set_u64(&j1, idx[q] + 1);
        u256_add(&cand, &calc, &j1); u256_add(&cand, &cand, base_key);         /* :5212-5219 */
        ge_scalar_mul(&pub, &GE_G, &cand);
        if (u256_cmp(&pub.x, &Q->x) == 0) { *priv = cand; return 1; }
        u256_sub(&cand, &calc, &j1); u256_add(&cand, &cand, base_key);         /* :5221-5228 */
        ge_scalar_mul(&pub, &GE_G, &cand);
        if (u256_cmp(&pub.x, &Q->x) == 0) { *priv = cand; return 1; }
      }
    } else if (u256_cmp(&S.x, &b->amp3[i].x) == 0) {                           /* :5238-5243 */
      u256_add(priv, &calc, base_key);
      return 1;
    }
  }
  return 0;
}
static int bsgs_secondcheck(obsgs *b, const ge *Q, const fe *start_range, uint64_t a, fe *priv) { /* :5151-5184 */
  fe base_key; key_add_mul(&base_key, start_range, 2 * b->m, a);               /* :5159-5161 */
  ge bp, nbp, S, P2;
  ge_scalar_mul(&bp, &GE_G, &base_key); ge_neg(&nbp, &bp);
  ge_add_direct(&S, Q, &nbp);
  for (int i = 0; i < 32; i++) {
    ge_add_direct(&P2, &S, &b->amp2[i]);
    uint8_t x[32];
    fe_to_be(x, &P2.x);
    if (kho_bloom_check(b->tier[1][x[0]], x, 32)) {
      fe base3; key_add_mul(&base3, &base_key, 2 * b->m2, (uint64_t)i);
      if (bsgs_thirdcheck(b, Q, &base3, priv)) return 1;
    }
  }
  return 0;
}

int kho_bsgs_search(void *h, const uint8_t pub_xy[64], const uint8_t start_be[32], const uint8_t end_be[32],
                    uint8_t found_key[32], uint64_t *giant_steps, uint64_t *tier1_positives) {
  return kho_bsgs_search_ex(h, pub_xy, start_be, end_be, 0, found_key, giant_steps, tier1_positives);
}

/* base_check != 0: the server variant of the loop (bsgsd.cpp:2528-2563) first compares every window's base point
 * with the target, which finds a key equal to a window base (the giant-step path itself cannot see offsets 0..2m
 * of a window) */
int kho_bsgs_search_ex(void *h, const uint8_t pub_xy[64], const uint8_t start_be[32], const uint8_t end_be[32], int base_check,
                       uint8_t found_key[32], uint64_t *giant_steps, uint64_t *tier1_positives) {
  obsgs *b = (obsgs *)h;
  ge Q; ge_from_be(&Q, pub_xy);
  fe base_key, end, step;
  fe_from_be(&base_key, start_be); fe_from_be(&end, end_be);
  u256_set_u64(&step, b->n); u256_add(&step, &step, &step);                    /* BSGS_STEP = 2N :1626 */
  uint64_t cycles = b->aux / 1024 + ((b->aux % 1024) ? 1 : 0);                 /* :4583 */
  ge *pts = (ge *)malloc(sizeof(ge) * GRP);
  uint64_t gs = 0, pos = 0;
  int found = 0;
  while (!found && u256_cmp(&base_key, &end) < 0) {                            /* :4616 */
    /* startP = Q + (order - base_key - 1025*m)*G  (:4635-4642) */
    fe km, t;
    u256_sub(&km, &FE_N, &base_key);
    u256_set_u64(&t, 1025); u256_mul_u64(&t, &t, b->m);
    u256_sub(&km, &km, &t);
    ge aux, c, next;
    if (base_check) {                                                            /* bsgsd.cpp:2544 base_point.equals */
      ge bp;
      ge_scalar_mul(&bp, &GE_G, &base_key);
      if (u256_cmp(&bp.x, &Q.x) == 0 && u256_cmp(&bp.y, &Q.y) == 0) { fe_to_be(found_key, &base_key); found = 1; break; }
    }
    ge_scalar_mul(&aux, &GE_G, &km);
    ge_add_direct(&c, &Q, &aux);
    for (uint64_t j = 0; j < cycles && !found; j++) {
      batch_around(&b->gs, &c, 0, pts, &next);
      for (int i = 0; i < GRP && !found; i++) {
        uint8_t x[32];
        fe_to_be(x, &pts[i].x);
        gs++;
        if (kho_bloom_check(b->tier[0][x[0]], x, 32)) {                        /* :4821 */
          pos++;
          fe priv;
          if (bsgs_secondcheck(b, &Q, &base_key, j * 1024 + (uint64_t)i, &priv)) {
            fe_to_be(found_key, &priv);
            found = 1;
          }
        }
      }
      c = next;
    }
    u256_add(&base_key, &base_key, &step);
  }
  free(pts);
  if (giant_steps) *giant_steps = gs;
  if (tier1_positives) *tier1_positives = pos;
  return found;
}
