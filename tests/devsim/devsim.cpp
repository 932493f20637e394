// tests/devsim/devsim.cpp — TEST INFRASTRUCTURE ONLY.
//
// Compiles the product's device headers (keyhunt_b200/csrc/*.cuh) for the HOST with plain g++ so the
// limb algorithms, message packers, batch geometry, bloom arithmetic and emit logic can be unit
// tested against the oracle on a machine that has no GPU (`pytest -m "not gpu"`).  The PTX bodies are
// replaced by the portable bodies that sit beside them in the same functions; everything else is
// the very code the kernels run.  This library is never loaded by keyhunt_b200 and is not a fallback:
// the product path requires libkh_b200.so + a GPU.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../keyhunt_b200/csrc/emit.cuh"
#include "../../keyhunt_b200/csrc/setup.cuh"
#include "../../keyhunt_b200/csrc/plan.hpp"

using namespace kh;

extern "C" {

void ds_fe_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { fe x, y, r; fe_from_be(x, a); fe_from_be(y, b); fe_mul(r, x, y); fe_to_be(out, r); }
void ds_fe_sqr(const uint8_t a[32], uint8_t out[32]) { fe x, r; fe_from_be(x, a); fe_sqr(r, x); fe_to_be(out, r); }
void ds_fe_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { fe x, y, r; fe_from_be(x, a); fe_from_be(y, b); fe_add(r, x, y); fe_to_be(out, r); }
void ds_fe_sub(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { fe x, y, r; fe_from_be(x, a); fe_from_be(y, b); fe_sub(r, x, y); fe_to_be(out, r); }
void ds_fe_neg(const uint8_t a[32], uint8_t out[32]) { fe x, r; fe_from_be(x, a); fe_neg(r, x); fe_to_be(out, r); }
// (hi * 2^256 + lo) mod P for any 512-bit value (the KH_FE_REDUCE_WIDE op of kh_selftest_fe)
void ds_fe_reduce_wide(const uint8_t hi[32], const uint8_t lo[32], uint8_t out[32]) {
  fe h, l, r; fe_from_be(h, hi); fe_from_be(l, lo);
  uint32_t w[16];
  for (int i = 0; i < 8; i++) { w[i] = l.v[i]; w[8 + i] = h.v[i]; }
  fe_reduce_wide(r, w); fe_to_be(out, r);
}
void ds_fe_inv(const uint8_t a[32], uint8_t out[32]) { fe x, r; fe_from_be(x, a); fe_inv(r, x); fe_to_be(out, r); }
// the other form of the multiplier's final reduction (fe.cuh KH_RARE_REDUCE; the kernels contain both)
void ds_fe_mul_alt(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { fe x, y, r; fe_from_be(x, a); fe_from_be(y, b); fe_mul<!KH_RARE_REDUCE>(r, x, y); fe_to_be(out, r); }
void ds_fe_sqr_alt(const uint8_t a[32], uint8_t out[32]) { fe x, r; fe_from_be(x, a); fe_sqr<!KH_RARE_REDUCE>(r, x); fe_to_be(out, r); }
void ds_fe_inv_alt(const uint8_t a[32], uint8_t out[32]) { fe x, r; fe_from_be(x, a); fe_inv<!KH_RARE_REDUCE>(r, x); fe_to_be(out, r); }
void ds_fe_reduce_wide_alt(const uint8_t hi[32], const uint8_t lo[32], uint8_t out[32]) {
  fe h, l, r; fe_from_be(h, hi); fe_from_be(l, lo);
  uint32_t w[16];
  for (int i = 0; i < 8; i++) { w[i] = l.v[i]; w[8 + i] = h.v[i]; }
  fe_reduce_wide<!KH_RARE_REDUCE>(r, w); fe_to_be(out, r);
}

void ds_pubkey(const uint8_t key[32], uint8_t xy[64]) {
  u256 k; u256_from_be(k, key);
  ge p; ge_mul_g(p, k);
  fe_to_be(xy, p.x); fe_to_be(xy + 32, p.y);
}
static void words_to_bytes(uint8_t out[20], const uint32_t w[5]) {
  for (int i = 0; i < 5; i++) { out[4 * i] = (uint8_t)w[i]; out[4 * i + 1] = (uint8_t)(w[i] >> 8); out[4 * i + 2] = (uint8_t)(w[i] >> 16); out[4 * i + 3] = (uint8_t)(w[i] >> 24); }
}
void ds_hash160_comp(int prefix, const uint8_t x[32], uint8_t out[20]) { fe v; fe_from_be(v, x); uint32_t h[5]; hash160_compressed(h, (uint32_t)prefix, v); words_to_bytes(out, h); }
void ds_hash160_uncomp(const uint8_t xy[64], uint8_t out[20]) { fe x, y; fe_from_be(x, xy); fe_from_be(y, xy + 32); uint32_t h[5]; hash160_uncompressed(h, x, y); words_to_bytes(out, h); }
void ds_eth_addr(const uint8_t xy[64], uint8_t out[20]) { fe x, y; fe_from_be(x, xy); fe_from_be(y, xy + 32); uint32_t h[5]; eth_address(h, x, y); words_to_bytes(out, h); }
// hash160 of the uncompressed key through the job form with the schedule table of the second block (hash.cuh KH_SHA_UNC2_TAB)
void ds_hash160_uncomp_tab(const uint8_t xy[64], uint8_t out[20]) {
  static std::vector<uint32_t> tab;
  if (tab.empty()) { tab.resize(KH_SHA2TAB_WORDS); for (uint32_t v = 0; v < 256; v++) sha_unc2_table_row(&tab[(size_t)v * KH_SHA2TAB_STRIDE], v); }
  fe x, y; fe_from_be(x, xy); fe_from_be(y, xy + 32);
  uint32_t h[5];
  hash160_job<true>(h, 2, x, y, tab.data());
  words_to_bytes(out, h);
}
uint64_t ds_xxh64_20(const uint8_t b[20], uint64_t seed) { uint32_t w[5]; memcpy(w, b, 20); return xxh64_20(w, seed); }
uint64_t ds_xxh64_32(const uint8_t b[32], uint64_t seed) { uint32_t w[8]; memcpy(w, b, 32); return xxh64_32(w, seed); }
uint64_t ds_bloom_mod(uint64_t x, uint64_t bits) { return bloom_mod(x, bits, (~0ULL) / bits); }

// ---- scan emulation: what the host API + kernels do, with T walker "threads" run one after another ----
struct DsHit { uint64_t index; uint32_t kind; uint8_t matched[20]; uint32_t variant; };

static void make_walk(const WalkSetup &ws, std::vector<uint32_t> &gtab, std::vector<uint32_t> &centers, std::vector<kh_u4> &scratch) {
  gtab.resize(KH_TAB_WORDS);
  for (uint32_t e = 0; e < KH_TAB_ENTRIES; e++) setup_table_entry(&gtab[16 * e], ws, e);
  centers.resize(16 * ws.T);
  // two-level set-up exactly as kh_run_setup does it: row bases by scalar multiplication, the rest of a row by setup_row_fill
  std::vector<uint32_t> offs(16 * (KH_SETUP_ROW - 1));
  for (uint32_t j = 1; j < KH_SETUP_ROW; j++) setup_row_offset(&offs[16 * (j - 1)], ws, j);
  for (uint64_t t = 0; t < ws.T; t += KH_SETUP_ROW) {
    fe cx, cy;
    setup_center(cx, cy, ws, t);
    for (int l = 0; l < 8; l++) { centers[l * ws.T + t] = cx.v[l]; centers[(8 + l) * ws.T + t] = cy.v[l]; }
  }
  std::vector<fe> pre(KH_SETUP_ROW);
  for (uint64_t row = 0; row * KH_SETUP_ROW < ws.T; row++) setup_row_fill(centers.data(), offs.data(), ws, row, pre.data());
  if (getenv("KH_DEVSIM_CHECK_SETUP")) {     // the two-level centres equal the directly computed ones
    for (uint64_t t = 0; t < ws.T; t++) {
      fe cx, cy;
      if (!setup_center(cx, cy, ws, t)) continue;
      for (int l = 0; l < 8; l++)
        if (centers[l * ws.T + t] != cx.v[l] || centers[(8 + l) * ws.T + t] != cy.v[l]) { fprintf(stderr, "devsim: two-level centre %llu differs\n", (unsigned long long)t); abort(); }
    }
  }
  scratch.resize((size_t)1024 * ws.T);
}

// centres of T walkers set up in two levels (setup_row_offset / setup_row_fill) against setup_center for every walker:
// returns the number of walkers that differ (walkers at infinity are compared by their flag only); *n_inf = walkers at infinity
int ds_setup_centres(const uint8_t k0[32], const uint8_t s[32], uint64_t T, int neg, const uint8_t *q_xy, uint64_t *n_inf) {
  WalkSetup ws;
  memset(&ws, 0, sizeof(ws));
  u256_from_be(ws.k0, k0); u256_from_be(ws.s, s);
  ws.neg = neg ? 1 : 0; ws.T = T; ws.first_batch = 0; ws.n_batches = 0; ws.comb = nullptr;
  ws.q.inf = 1;
  if (q_xy) { fe_from_be(ws.q.x, q_xy); fe_from_be(ws.q.y, q_xy + 32); ws.q.inf = 0; }
  std::vector<uint32_t> gtab, centers;
  std::vector<kh_u4> scratch;
  std::vector<uint32_t> offs(16 * (KH_SETUP_ROW - 1));
  centers.resize(16 * T);
  for (uint32_t j = 1; j < KH_SETUP_ROW; j++) setup_row_offset(&offs[16 * (j - 1)], ws, j);
  for (uint64_t t = 0; t < T; t += KH_SETUP_ROW) {
    fe cx, cy;
    setup_center(cx, cy, ws, t);
    for (int l = 0; l < 8; l++) { centers[l * T + t] = cx.v[l]; centers[(8 + l) * T + t] = cy.v[l]; }
  }
  std::vector<fe> pre(KH_SETUP_ROW);
  uint64_t flagged = 0;
  for (uint64_t row = 0; row * KH_SETUP_ROW < T; row++) if (!setup_row_fill(centers.data(), offs.data(), ws, row, pre.data())) flagged++;
  int bad = 0;
  uint64_t inf = 0;
  for (uint64_t t = 0; t < T; t++) {
    fe cx, cy;
    if (!setup_center(cx, cy, ws, t)) { inf++; continue; }
    for (int l = 0; l < 8; l++) if (centers[l * T + t] != cx.v[l] || centers[(8 + l) * T + t] != cy.v[l]) { bad++; break; }
  }
  if (n_inf) *n_inf = inf;
  if (inf && !flagged && (inf > 1 || T % KH_SETUP_ROW != 1)) {
    // a walker at infinity that is not a row base must have been reported by its row
    bool only_bases = true;
    for (uint64_t t = 0; t < T; t++) { fe cx, cy; if (!setup_center(cx, cy, ws, t) && t % KH_SETUP_ROW) only_bases = false; }
    if (!only_bases) bad += 1000000;
  }
  return bad;
}

// the plan of the last emulated scan (plan.hpp): out[0] = segments, out[1] = collapsed batches, out[2] = distinct T values
static std::vector<ScanSegment> g_last_plan;
void ds_last_plan(uint64_t out[3]) {
  out[0] = g_last_plan.size(); out[1] = 0;
  std::vector<uint64_t> ts;
  for (const ScanSegment &sg : g_last_plan) { out[1] += sg.collapsed; ts.push_back(sg.T); }
  std::sort(ts.begin(), ts.end());
  out[2] = (uint64_t)(std::unique(ts.begin(), ts.end()) - ts.begin());
}
// 1 if plan.hpp's constant really is 1024^-1 mod n
int ds_plan_selfcheck(void) {
  u256 one, z;
  u256_set_u64(one, 1); u256_set_u64(z, 0);
  const ScanGeometry g = plan_geometry(z, one);
  u256 k;
  u256_set_u64(k, 1024);
  u256 r;
  u256_mulmod_n(r, g.inv1024, k);
  for (int i = 1; i < 8; i++) if (r.v[i]) return 0;
  return r.v[0] == 1;
}

// -m vanity for the next ds_scan calls: van = 2048-word prefix bitmap + n x (A[5], B[5]) big-endian words (ScanTargets::van); n = 0 switches it off
static const uint32_t *g_van = nullptr;
static uint32_t g_van_n = 0;
void ds_set_vanity(const uint32_t *van, uint32_t n) { g_van = van; g_van_n = n; }

// kind: KH_SCAN_*; table20: N sorted 20-byte records; bloom image as bytes
int64_t ds_scan(int kind, const uint8_t *table20, uint64_t n_targets, const uint8_t *bloom_bytes, uint64_t bloom_bits,
                uint32_t bloom_hashes, const uint8_t start[32], const uint8_t stride[32], uint64_t n_batches, uint64_t T_cap,
                uint32_t steps_per_launch, DsHit *out, uint32_t max_hits, int endo) {
  WalkSetup ws;
  memset(&ws, 0, sizeof(ws));
  u256_from_be(ws.s, stride);
  u256_from_be(ws.k0, start);
  ws.q.inf = 1; ws.neg = 0;
  // what kh_scan does: the batches as segments (almost always one), T walkers at most
  if (!plan_scan(ws.k0, ws.s, n_batches, T_cap, 1, g_last_plan)) return -2;
  std::vector<uint32_t> gtab, centers;
  std::vector<kh_u4> scratch;

  std::vector<uint32_t> table(5 * n_targets);
  for (uint64_t i = 0; i < n_targets; i++)
    for (int k = 0; k < 5; k++) {
      const uint8_t *p = table20 + 20 * i + 4 * k;
      table[5 * i + k] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
  std::vector<RawHit> raw(max_hits);
  uint32_t count = 0;
  ScanTargets tg;
  tg.bloom.bf = const_cast<uint8_t *>(bloom_bytes);
  tg.bloom.bits = bloom_bits; tg.bloom.magic = (~0ULL) / bloom_bits; tg.bloom.stride = 0; tg.bloom.hashes = bloom_hashes;
  tg.table = table.data(); tg.n = n_targets;
  tg.sink.hits = raw.data(); tg.sink.count = &count; tg.sink.cap = max_hits;
  tg.van = g_van; tg.van_n = g_van_n; tg.pre = nullptr; tg.pre_k = 0;
  // the prefix bitmap exactly as kh_set_targets sizes and fills it (kh_scan.cu)
  std::vector<uint32_t> bm;
  {
    uint32_t k = 16;
    while (k < 32 && (1ull << k) < 256ull * n_targets) k++;
    bm.assign((size_t)1 << (k - 5), 0u);
    for (uint64_t i = 0; i < n_targets; i++) { const uint32_t idx = table[5 * i] >> (32 - k); bm[idx >> 5] |= 1u << (idx & 31); }
    tg.pre = bm.data(); tg.pre_k = k;
  }

  for (const ScanSegment &sg : g_last_plan) {
  const uint64_t T = sg.T;
  ws.T = T; ws.first_batch = sg.first; ws.n_batches = sg.end;
  make_walk(ws, gtab, centers, scratch);
  WalkParams wp;
  wp.gtab = gtab.data(); wp.centers = centers.data(); wp.scratch = scratch.data();
  wp.T = T; wp.n_batches = sg.end; wp.steps = steps_per_launch; wp.pad = 0;
  for (uint64_t base = sg.first; base < sg.end; base += (uint64_t)steps_per_launch * T) {
    wp.batch_base = base;
    for (uint64_t t = 0; t < T; t++) {
      switch (kind + (endo ? 8 : 0) + (g_van_n ? 16 : 0)) {
        case 16 + KH_SCAN_COMP:   { ScanEmit<KH_SCAN_COMP, false, true> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case 16 + KH_SCAN_UNCOMP: { ScanEmit<KH_SCAN_UNCOMP, false, true> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case 16 + KH_SCAN_BOTH:   { ScanEmit<KH_SCAN_BOTH, false, true> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case 24 + KH_SCAN_COMP:   { ScanEmit<KH_SCAN_COMP, true, true> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case 24 + KH_SCAN_UNCOMP: { ScanEmit<KH_SCAN_UNCOMP, true, true> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case 24 + KH_SCAN_BOTH:   { ScanEmit<KH_SCAN_BOTH, true, true> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case KH_SCAN_XPOINT: { ScanEmit<KH_SCAN_XPOINT> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case KH_SCAN_COMP:   { ScanEmit<KH_SCAN_COMP> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case KH_SCAN_UNCOMP: { ScanEmit<KH_SCAN_UNCOMP> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case KH_SCAN_BOTH:   { ScanEmit<KH_SCAN_BOTH> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case KH_SCAN_ETH:    { ScanEmit<KH_SCAN_ETH> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case 8 + KH_SCAN_XPOINT: { ScanEmit<KH_SCAN_XPOINT, true> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case 8 + KH_SCAN_COMP:   { ScanEmit<KH_SCAN_COMP, true> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case 8 + KH_SCAN_UNCOMP: { ScanEmit<KH_SCAN_UNCOMP, true> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case 8 + KH_SCAN_BOTH:   { ScanEmit<KH_SCAN_BOTH, true> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        case 8 + KH_SCAN_ETH:    { ScanEmit<KH_SCAN_ETH, true> e(tg); walk_batches(wp, gtab.data(), t, e); break; }
        default: return -1;
      }
    }
  }
  }
  uint32_t n = count < max_hits ? count : max_hits;
  for (uint32_t i = 0; i < n; i++) {
    out[i].index = raw[i].batch * KH_GRP + raw[i].idx;
    out[i].kind = raw[i].kind;
    words_to_bytes(out[i].matched, raw[i].h);
    out[i].variant = raw[i].variant;
  }
  return count;
}

// all X (and Y) of the first n_batches batches, in range order: out[(b*1024+i)*64]
struct DumpEmit {
  static constexpr bool NEED_Y = true;
  static constexpr bool OUTLINE_MUL = false;
  static constexpr int RARE_REDUCE = KH_RARE_REDUCE;
  static constexpr bool INV_SQR = true;
  static constexpr bool PAIRS = false;
  void pair(const fe &, uint32_t, const fe &, uint32_t, uint64_t) {}
  uint8_t *out;
  void point(const fe &x, const fe &y, uint64_t batch, uint32_t idx) {
    uint8_t *p = out + (batch * KH_GRP + idx) * 64;
    fe_to_be(p, x); fe_to_be(p + 32, y);
  }
};
void ds_walk_dump(const uint8_t start[32], const uint8_t stride[32], uint64_t n_batches, uint64_t T_cap, uint32_t steps_per_launch, uint8_t *out) {
  WalkSetup ws;
  memset(&ws, 0, sizeof(ws));
  u256_from_be(ws.s, stride);
  u256_from_be(ws.k0, start);
  ws.q.inf = 1; ws.neg = 0;
  if (!plan_scan(ws.k0, ws.s, n_batches, T_cap, 1, g_last_plan)) return;
  std::vector<uint32_t> gtab, centers;
  std::vector<kh_u4> scratch;
  DumpEmit e; e.out = out;
  for (const ScanSegment &sg : g_last_plan) {
    ws.T = sg.T; ws.first_batch = sg.first; ws.n_batches = sg.end;
    make_walk(ws, gtab, centers, scratch);
    WalkParams wp;
    wp.gtab = gtab.data(); wp.centers = centers.data(); wp.scratch = scratch.data();
    wp.T = sg.T; wp.n_batches = sg.end; wp.steps = steps_per_launch; wp.pad = 0;
    for (uint64_t base = sg.first; base < sg.end; base += (uint64_t)steps_per_launch * sg.T) {
      wp.batch_base = base;
      for (uint64_t t = 0; t < sg.T; t++) walk_batches(wp, gtab.data(), t, e);
    }
  }
}

}  // extern "C"
