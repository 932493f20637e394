// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// White-box harness around the UNMODIFIED reference objects (compiled by oracle/Makefile from
// /root/reference into oracle/_ref/obj_pic/).  It exposes the reference's own primitives behind the
// same C signatures as our restatement (kh_oracle.h, prefix kho_ -> khr_) so tests can pin the
// restatement against the real thing, and so golden vectors can be generated from the reference
// itself (tests/golden/make_golden.py).  Nothing here is shipped or measured.
//
// Reference entry points used (file:line under /root/reference):
//   Int::ModMulK1 IntMod.cpp:855, Int::ModSquareK1 :977, Int::ModInv :382, IntGroup::ModInv IntGroup.cpp:36
//   Secp256K1::ComputePublicKey SECP256K1.cpp:205, AddDirect :455, DoubleDirect :589, Negation :316
//   Secp256K1::GetHash160_fromX SECP256K1.cpp:1207, GetHash160 (4-lane) :1045, (scalar) :1134
//   KECCAK_256 path keyhunt.cpp:5647 (SHA3_256_Init/Update sha3.c:306, KECCAK_256_Final sha3.c:414)
//   XXH64 xxhash.h:2512, bloom_init2 bloom.cpp:154, bloom_add :215, bloom_check :189
//   batch geometry: thread_process keyhunt.cpp:3348-3461 (restated below with the reference's classes)
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "secp256k1/SECP256k1.h"
#include "secp256k1/IntGroup.h"
#include "bloom/bloom.h"
#include "sha3/sha3.h"
#include "hash/sha256.h"
#include "hash/ripemd160.h"
#define XXH_STATIC_LINKING_ONLY
#include "xxhash/xxhash.h"
extern "C" {
#include "base58/libbase58.h"
}

static Secp256K1 *g_secp = nullptr;

static void ensure_init() {
  if (!g_secp) {
    g_secp = new Secp256K1();
    g_secp->Init();
  }
}

static void canon(Int &a) {  // reference results may be lazily reduced (>= P); canonicalise
  ensure_init();
  while (a.IsGreaterOrEqual(&g_secp->P)) a.Sub(&g_secp->P);
}

extern "C" {

void khr_init(void) { ensure_init(); }

void khr_fe_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  ensure_init();
  Int x, y, r;
  x.Set32Bytes((unsigned char *)a);
  y.Set32Bytes((unsigned char *)b);
  r.ModMulK1(&x, &y);
  canon(r);
  r.Get32Bytes(out);
}

void khr_fe_sqr(const uint8_t a[32], uint8_t out[32]) {
  ensure_init();
  Int x, r;
  x.Set32Bytes((unsigned char *)a);
  r.ModSquareK1(&x);
  canon(r);
  r.Get32Bytes(out);
}

void khr_fe_inv(const uint8_t a[32], uint8_t out[32]) {
  ensure_init();
  Int x;
  x.Set32Bytes((unsigned char *)a);
  x.ModInv();
  canon(x);
  x.Get32Bytes(out);
}

// Int::ModAdd(a, b) / ModSub(a, b) / ModNeg (IntMod.cpp:51, :97, :105)
void khr_fe_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  ensure_init();
  Int x, y, r;
  x.Set32Bytes((unsigned char *)a);
  y.Set32Bytes((unsigned char *)b);
  r.ModAdd(&x, &y);
  canon(r);
  r.Get32Bytes(out);
}
void khr_fe_sub(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  ensure_init();
  Int x, y, r;
  x.Set32Bytes((unsigned char *)a);
  y.Set32Bytes((unsigned char *)b);
  r.ModSub(&x, &y);
  canon(r);
  r.Get32Bytes(out);
}
void khr_fe_neg(const uint8_t a[32], uint8_t out[32]) {
  ensure_init();
  Int x;
  x.Set32Bytes((unsigned char *)a);
  x.ModNeg();
  canon(x);
  x.Get32Bytes(out);
}

void khr_pubkey(const uint8_t key[32], uint8_t xy[64]) {
  ensure_init();
  Int k;
  k.Set32Bytes((unsigned char *)key);
  Point p = g_secp->ComputePublicKey(&k);
  p.x.Get32Bytes(xy);
  p.y.Get32Bytes(xy + 32);
}

void khr_hash160_comp(int prefix, const uint8_t x[32], uint8_t out[20]) {
  ensure_init();
  Int k0, k1, k2, k3;
  uint8_t h1[20], h2[20], h3[20];
  k0.Set32Bytes((unsigned char *)x);
  k1.Set(&k0); k2.Set(&k0); k3.Set(&k0);
  g_secp->GetHash160_fromX(P2PKH, (unsigned char)prefix, &k0, &k1, &k2, &k3, out, h1, h2, h3);
}

void khr_hash160_uncomp(const uint8_t xy[64], uint8_t out[20]) {
  ensure_init();
  Point p0, p1, p2, p3;
  uint8_t h1[20], h2[20], h3[20];
  p0.x.Set32Bytes((unsigned char *)xy);
  p0.y.Set32Bytes((unsigned char *)xy + 32);
  p0.z.SetInt32(1);
  p1.Set(p0); p2.Set(p0); p3.Set(p0);
  g_secp->GetHash160(P2PKH, false, p0, p1, p2, p3, out, h1, h2, h3);
}

// scalar single-key path used by hit fix-up / writekey (SECP256K1.cpp:1134)
void khr_hash160_scalar(int compressed, const uint8_t xy[64], uint8_t out[20]) {
  ensure_init();
  Point p;
  p.x.Set32Bytes((unsigned char *)xy);
  p.y.Set32Bytes((unsigned char *)xy + 32);
  p.z.SetInt32(1);
  g_secp->GetHash160(P2PKH, compressed != 0, p, out);
}

void khr_eth_addr(const uint8_t xy[64], uint8_t out[20]) {  // keyhunt.cpp:5647-5669
  uint8_t buf[64];
  memcpy(buf, xy, 64);
  SHA3_256_CTX ctx;
  SHA3_256_Init(&ctx);
  SHA3_256_Update(&ctx, buf, 64);
  KECCAK_256_Final(buf, &ctx);
  memcpy(out, buf + 12, 20);
}

uint64_t khr_xxh64(const void *buf, uint64_t len, uint64_t seed) { return XXH64(buf, (size_t)len, seed); }
// base58/base58.c:39 — the decoder addvanity (keyhunt.cpp:6739) leans on, with its whole-buffer / length conventions
int khr_b58tobin(uint8_t *bin, uint64_t *binsz, const char *b58, uint64_t b58sz) {
  size_t n = (size_t)*binsz;
  bool ok = b58tobin(bin, &n, b58, (size_t)b58sz);
  *binsz = n;
  return ok ? 1 : 0;
}

void khr_sha256(const uint8_t *in, uint64_t len, uint8_t out[32]) { sha256((uint8_t *)in, (size_t)len, out); }

// ---- bloom (bloom.cpp) -------------------------------------------------------------------------
void *khr_bloom_new(uint64_t entries) {  // error literal as in initBloomFilter keyhunt.cpp:7620
  struct bloom *b = (struct bloom *)calloc(1, sizeof(struct bloom));
  if (bloom_init2(b, entries, 0.000001) != 0) { free(b); return nullptr; }
  return b;
}
void khr_bloom_free(void *h) { if (h) { bloom_free((struct bloom *)h); free(h); } }
void khr_bloom_desc(void *h, uint64_t *entries, uint64_t *bits, uint64_t *bytes, uint32_t *hashes) {
  struct bloom *b = (struct bloom *)h;
  *entries = b->entries; *bits = b->bits; *bytes = b->bytes; *hashes = b->hashes;
}
uint8_t *khr_bloom_data(void *h) { return ((struct bloom *)h)->bf; }
int khr_bloom_add(void *h, const void *buf, int len) { return bloom_add((struct bloom *)h, buf, len); }
int khr_bloom_check(void *h, const void *buf, int len) { return bloom_check((struct bloom *)h, buf, len); }
int khr_sizeof_bloom(void) { return (int)sizeof(struct bloom); }

// ---- the 1024-point batch of thread_process (keyhunt.cpp:3348-3461), with the reference classes --
// out: 1024 x 64 bytes (X||Y big-endian); pts[i] <-> key base + i*stride.  with_y=0 leaves Y of the
// non-centre points as the reference leaves them (stale centre Y) — callers only compare X then.
void khr_batch_points(const uint8_t base_key[32], const uint8_t stride_be[32], int with_y, uint8_t *out) {
  ensure_init();
  const int GRP = 1024, half = GRP / 2, hLength = half - 1;
  Int stride, key, tmp;
  stride.Set32Bytes((unsigned char *)stride_be);
  key.Set32Bytes((unsigned char *)base_key);
  // init_generator keyhunt.cpp:5266
  std::vector<Point> Gn(half);
  Point G = g_secp->ComputePublicKey(&stride);
  Point g; g.Set(G);
  Gn[0] = g;
  g = g_secp->DoubleDirect(g);
  Gn[1] = g;
  for (int i = 2; i < half; i++) { g = g_secp->AddDirect(g, G); Gn[i] = g; }
  Point _2Gn = g_secp->DoubleDirect(Gn[half - 1]);

  std::vector<Int> dx(half + 1);
  IntGroup grp(half + 1);
  grp.Set(dx.data());
  std::vector<Point> pts(GRP);
  Int dy, dyn, _s, _p;
  Point pp, pn;

  tmp.SetInt32(half);
  tmp.Mult(&stride);
  key.Add(&tmp);
  Point startP = g_secp->ComputePublicKey(&key);
  int i;
  for (i = 0; i < hLength; i++) dx[i].ModSub(&Gn[i].x, &startP.x);
  dx[i].ModSub(&Gn[i].x, &startP.x);
  dx[i + 1].ModSub(&_2Gn.x, &startP.x);
  grp.ModInv();
  pts[half] = startP;
  for (i = 0; i < hLength; i++) {
    pp = startP; pn = startP;
    dy.ModSub(&Gn[i].y, &pp.y);
    _s.ModMulK1(&dy, &dx[i]);
    _p.ModSquareK1(&_s);
    pp.x.ModNeg(); pp.x.ModAdd(&_p); pp.x.ModSub(&Gn[i].x);
    if (with_y) { pp.y.ModSub(&Gn[i].x, &pp.x); pp.y.ModMulK1(&_s); pp.y.ModSub(&Gn[i].y); }
    dyn.Set(&Gn[i].y); dyn.ModNeg(); dyn.ModSub(&pn.y);
    _s.ModMulK1(&dyn, &dx[i]);
    _p.ModSquareK1(&_s);
    pn.x.ModNeg(); pn.x.ModAdd(&_p); pn.x.ModSub(&Gn[i].x);
    if (with_y) { pn.y.ModSub(&Gn[i].x, &pn.x); pn.y.ModMulK1(&_s); pn.y.ModAdd(&Gn[i].y); }
    pts[half + (i + 1)] = pp;
    pts[half - (i + 1)] = pn;
  }
  pn = startP;
  dyn.Set(&Gn[i].y); dyn.ModNeg(); dyn.ModSub(&pn.y);
  _s.ModMulK1(&dyn, &dx[i]);
  _p.ModSquareK1(&_s);
  pn.x.ModNeg(); pn.x.ModAdd(&_p); pn.x.ModSub(&Gn[i].x);
  if (with_y) { pn.y.ModSub(&Gn[i].x, &pn.x); pn.y.ModMulK1(&_s); pn.y.ModAdd(&Gn[i].y); }
  pts[0] = pn;
  for (int j = 0; j < GRP; j++) {
    canon(pts[j].x); canon(pts[j].y);
    pts[j].x.Get32Bytes(out + 64 * j);
    pts[j].y.Get32Bytes(out + 64 * j + 32);
  }
}

// affine add of two arbitrary points (AddDirect), for BSGS tier checks
void khr_add_direct(const uint8_t a[64], const uint8_t b[64], uint8_t out[64]) {
  ensure_init();
  Point p, q;
  p.x.Set32Bytes((unsigned char *)a); p.y.Set32Bytes((unsigned char *)a + 32); p.z.SetInt32(1);
  q.x.Set32Bytes((unsigned char *)b); q.y.Set32Bytes((unsigned char *)b + 32); q.z.SetInt32(1);
  Point r = g_secp->AddDirect(p, q);
  canon(r.x); canon(r.y);
  r.x.Get32Bytes(out); r.y.Get32Bytes(out + 32);
}

}  // extern "C"
