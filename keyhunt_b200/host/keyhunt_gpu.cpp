// keyhunt_gpu.cpp — host driver with keyhunt's command-line surface on top of the C ABI (include/keyhunt_b200.h).
//
// It mirrors the L5 layer of the reference (keyhunt.cpp main() :643-3013): getopt flags, range set-up, target
// file loading, the range cursor handed out in chunks of N_SEQUENTIAL_MAX keys, the hit output formats
// (writekey :6891, writekeyeth :6925, BSGS "Key found privkey" :4837-4840) and the `-S` BSGS table files.
// Everything from "scan this chunk" downwards is the GPU library; this file contains no curve or hash code on
// the hot path (the only host arithmetic is 256-bit add/compare for the cursor, SHA-256 for base58check /
// file checksums, and one modular square root per BSGS public key to decompress it).
//
//   -t N   number of GPUs to use (the reference's worker-thread count); default 1
// Modes outside the GPU path (minikeys, pub2rmd), -R random and the mmap'd bloom/ptable flags are
// parsed; the former are refused, the latter accepted and ignored (tables live in HBM).  -e is supported;
// -B sequential | backward | both are the window pickers over kh_bsgs_search.
#include <getopt.h>
#include <inttypes.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/keyhunt_b200.h"

// ---------------------------------------------------------------------------------------------------
// 256-bit big-endian integers (only what the cursor needs)
// ---------------------------------------------------------------------------------------------------
struct U256 { uint8_t b[32]; };
static U256 u_zero() { U256 r; memset(r.b, 0, 32); return r; }
static U256 u_from_u64(uint64_t v) { U256 r = u_zero(); for (int i = 0; i < 8; i++) r.b[31 - i] = (uint8_t)(v >> (8 * i)); return r; }
static int u_cmp(const U256 &a, const U256 &b) { return memcmp(a.b, b.b, 32); }
static bool u_is_zero(const U256 &a) { for (int i = 0; i < 32; i++) if (a.b[i]) return false; return true; }
static U256 u_add(const U256 &a, const U256 &b) { U256 r; int c = 0; for (int i = 31; i >= 0; i--) { int s = a.b[i] + b.b[i] + c; r.b[i] = (uint8_t)s; c = s >> 8; } return r; }
static U256 u_sub(const U256 &a, const U256 &b) { U256 r; int c = 0; for (int i = 31; i >= 0; i--) { int s = a.b[i] - b.b[i] - c; r.b[i] = (uint8_t)s; c = (s < 0); } return r; }
static U256 u_shl1(const U256 &a) { U256 r; int c = 0; for (int i = 31; i >= 0; i--) { int s = (a.b[i] << 1) | c; r.b[i] = (uint8_t)s; c = s >> 8; } return r; }
static U256 u_mul_u64(const U256 &a, uint64_t m) { U256 r = u_zero(), t = a; for (int i = 0; i < 64; i++) { if ((m >> i) & 1) r = u_add(r, t); t = u_shl1(t); } return r; }
static bool u_from_hex(U256 &r, const char *s) {
  if (s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) s += 2;
  size_t n = strlen(s);
  if (n == 0 || n > 64) return false;
  r = u_zero();
  for (size_t i = 0; i < n; i++) {
    char c = s[n - 1 - i];
    int v = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : -1;
    if (v < 0) return false;
    r.b[31 - i / 2] |= (uint8_t)(v << (4 * (i & 1)));
  }
  return true;
}
static std::string u_hex(const U256 &a) {  // Int::GetBase16: lower case, no leading zeros
  static const char *d = "0123456789abcdef";
  std::string s;
  for (int i = 0; i < 32; i++) { s.push_back(d[a.b[i] >> 4]); s.push_back(d[a.b[i] & 15]); }
  size_t p = s.find_first_not_of('0');
  return p == std::string::npos ? "0" : s.substr(p);
}
static std::string hex_of(const uint8_t *p, size_t n) {
  static const char *d = "0123456789abcdef";
  std::string s;
  for (size_t i = 0; i < n; i++) { s.push_back(d[p[i] >> 4]); s.push_back(d[p[i] & 15]); }
  return s;
}
static const uint8_t ORDER_N[32] = {0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFE,
                                    0xBA, 0xAE, 0xDC, 0xE6, 0xAF, 0x48, 0xA0, 0x3B, 0xBF, 0xD2, 0x5E, 0x8C, 0xD0, 0x36, 0x41, 0x41};

// ---------------------------------------------------------------------------------------------------
// SHA-256 (address checksum, table-file checksums) and base58 — cold host code
// ---------------------------------------------------------------------------------------------------
static void sha256_host(const uint8_t *in, size_t len, uint8_t out[32]) {
  static const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
      0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
      0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
      0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
      0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
      0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  auto ror = [](uint32_t x, int n) { return (x >> n) | (x << (32 - n)); };
  auto block = [&](const uint8_t *p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = ror(w[i - 15], 7) ^ ror(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = ror(w[i - 2], 17) ^ ror(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t t1 = hh + (ror(e, 6) ^ ror(e, 11) ^ ror(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
      uint32_t t2 = (ror(a, 2) ^ ror(a, 13) ^ ror(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  };
  size_t off = 0;
  for (; off + 64 <= len; off += 64) block(in + off);
  uint8_t tail[128];
  size_t rem = len - off;
  memset(tail, 0, sizeof(tail));
  memcpy(tail, in + off, rem);
  tail[rem] = 0x80;
  size_t tl = rem < 56 ? 64 : 128;
  uint64_t bits = (uint64_t)len * 8;
  for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
  block(tail);
  if (tl == 128) block(tail + 64);
  for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
}
static const char *B58 = "123456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz";
static std::string b58enc(const uint8_t *in, size_t len) {
  std::vector<uint8_t> d(len * 2, 0);
  size_t zeros = 0;
  while (zeros < len && in[zeros] == 0) zeros++;
  size_t dl = 0;
  for (size_t i = zeros; i < len; i++) {
    int carry = in[i];
    for (size_t j = 0; j < dl; j++) { carry += d[j] << 8; d[j] = carry % 58; carry /= 58; }
    while (carry) { d[dl++] = carry % 58; carry /= 58; }
  }
  std::string s(zeros, '1');
  for (size_t j = 0; j < dl; j++) s.push_back(B58[d[dl - 1 - j]]);
  return s;
}
static bool b58dec25(const char *s, uint8_t out[25]) {
  uint8_t acc[40];
  memset(acc, 0, sizeof(acc));
  for (const char *p = s; *p; p++) {
    const char *q = strchr(B58, *p);
    if (!q) return false;
    int carry = (int)(q - B58);
    for (int i = 39; i >= 0; i--) { carry += 58 * acc[i]; acc[i] = (uint8_t)carry; carry >>= 8; }
    if (carry) return false;
  }
  for (int i = 0; i < 15; i++) if (acc[i]) return false;
  memcpy(out, acc + 15, 25);
  return true;
}
static std::string rmd160_to_address(const uint8_t h[20]) {  // rmd160toaddress_dst keyhunt.cpp:3016
  uint8_t d[25], c1[32], c2[32];
  d[0] = 0;
  memcpy(d + 1, h, 20);
  sha256_host(d, 21, c1);
  sha256_host(c1, 32, c2);
  memcpy(d + 21, c2, 4);
  return b58enc(d, 25);
}

// ---------------------------------------------------------------------------------------------------
// secp256k1 field (host, __int128) — only to decompress BSGS public keys (ParsePublicKeyHex SECP256K1.cpp:327)
// ---------------------------------------------------------------------------------------------------
typedef unsigned __int128 u128;
struct F { uint64_t l[4]; };
static const F FP = {{0xFFFFFFFEFFFFFC2FULL, ~0ULL, ~0ULL, ~0ULL}};
static int f_cmp(const F &a, const F &b) { for (int i = 3; i >= 0; i--) { if (a.l[i] != b.l[i]) return a.l[i] < b.l[i] ? -1 : 1; } return 0; }
static void f_subp(F &a) { u128 br = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)a.l[i] - FP.l[i] - (uint64_t)br; a.l[i] = (uint64_t)d; br = (d >> 64) & 1; } }
static F f_mul(const F &a, const F &b) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 4; i++) { u128 c = 0; for (int j = 0; j < 4; j++) { c += (u128)a.l[i] * b.l[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; } t[i + 4] = (uint64_t)c; }
  const uint64_t C = 0x1000003D1ULL;
  u128 c = 0; F r;
  for (int i = 0; i < 4; i++) { c += (u128)t[4 + i] * C + t[i]; r.l[i] = (uint64_t)c; c >>= 64; }
  uint64_t top = (uint64_t)c;
  c = (u128)top * C + r.l[0]; r.l[0] = (uint64_t)c; c >>= 64;
  for (int i = 1; i < 4; i++) { c += r.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
  if (c) { c = (u128)r.l[0] + C; r.l[0] = (uint64_t)c; c >>= 64; for (int i = 1; i < 4; i++) { c += r.l[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
  if (f_cmp(r, FP) >= 0) f_subp(r);
  return r;
}
static F f_from_be(const uint8_t *b) { F r; for (int i = 0; i < 4; i++) { uint64_t v = 0; for (int j = 0; j < 8; j++) v = (v << 8) | b[(3 - i) * 8 + j]; r.l[i] = v; } return r; }
static void f_to_be(uint8_t *b, const F &a) { for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[(3 - i) * 8 + j] = (uint8_t)(a.l[i] >> (56 - 8 * j)); }
static bool decompress_pub(const uint8_t x_be[32], int odd, uint8_t y_be[32]) {
  F x = f_from_be(x_be);
  if (f_cmp(x, FP) >= 0) return false;
  F x3 = f_mul(f_mul(x, x), x);
  F seven = {{7, 0, 0, 0}};
  u128 c = 0; F rhs;
  for (int i = 0; i < 4; i++) { c += (u128)x3.l[i] + seven.l[i]; rhs.l[i] = (uint64_t)c; c >>= 64; }
  if (c || f_cmp(rhs, FP) >= 0) f_subp(rhs);
  // y = rhs^((P+1)/4) ; P+1 does not overflow 256 bits
  F e = FP;
  e.l[0] += 1;
  for (int i = 0; i < 4; i++) e.l[i] = (e.l[i] >> 2) | (i < 3 ? (e.l[i + 1] << 62) : 0);
  F acc = {{1, 0, 0, 0}}, base = rhs;
  for (int i = 0; i < 256; i++) { if ((e.l[i / 64] >> (i % 64)) & 1) acc = f_mul(acc, base); base = f_mul(base, base); }
  if (f_cmp(f_mul(acc, acc), rhs) != 0) return false;   // not on the curve
  if ((int)(acc.l[0] & 1) != odd) { F n = FP; u128 br = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)n.l[i] - acc.l[i] - (uint64_t)br; n.l[i] = (uint64_t)d; br = (d >> 64) & 1; } acc = n; }
  f_to_be(y_be, acc);
  return true;
}

// ---------------------------------------------------------------------------------------------------
// options / globals (names follow keyhunt.cpp:409-641)
// ---------------------------------------------------------------------------------------------------
static int FLAGMODE = KH_MODE_ADDRESS, FLAGCRYPTO = 0, FLAGSEARCH = KH_SEARCH_BOTH, NGPUS = 1, KFACTOR = 1;
static int FLAGQUIET = 0, FLAGMATRIX = 0, FLAGSAVEREADFILE = 0, FLAGSKIPCHECKSUM = 0, FLAGBITRANGE = 0, FLAGRANGE = 0, FLAG_N = 0;
static int FLAGBLOOMMULTIPLIER = 1, OUTPUTSECONDS = 30, FLAGENDOMORPHISM = 0, FLAGBSGSMODE = 0;
static const char *fileName = "addresses.txt", *str_N = nullptr;
static U256 n_range_start, n_range_end, stride_v;
static uint64_t N_SEQUENTIAL_MAX = 0x100000000ULL;
static std::mutex write_random, write_keys;
static std::atomic<uint64_t> total_points{0};

static void die(const char *fmt, const char *a = "") { fprintf(stderr, fmt, a); fprintf(stderr, "\n"); exit(EXIT_FAILURE); }

static int validate_nk(uint64_t n, uint64_t k) {   // util.c:358-389
  if (n < (1ULL << 20)) { fprintf(stderr, "[E] n must be at least 2^20 (0x100000)\n"); return 0; }
  if (n & (n - 1)) { fprintf(stderr, "[E] n must be a power of two\n"); return 0; }
  int bits = 0;
  for (uint64_t t = n; t > 1; t >>= 1) bits++;
  if (bits % 2 || bits < 20 || bits > 64) { fprintf(stderr, "[E] invalid n 0x%" PRIx64 "\n", n); return 0; }
  uint64_t kmax = 1ULL << ((bits - 20) / 2);
  if (k > kmax) { fprintf(stderr, "[E] k value %" PRIu64 " is too large for n 0x%" PRIx64 " (max %" PRIu64 ")\n", k, n, kmax); return 0; }
  return 1;
}

static std::string trim(const char *s) { std::string t(s); size_t a = t.find_first_not_of(" \t\r\n"), b = t.find_last_not_of(" \t\r\n"); return a == std::string::npos ? "" : t.substr(a, b - a + 1); }
static bool is_hex(const std::string &s) { if (s.empty()) return false; for (char c : s) if (!isxdigit((unsigned char)c)) return false; return true; }
static bool hex2bin(const std::string &s, uint8_t *out, size_t n) { if (s.size() < 2 * n) return false; for (size_t i = 0; i < n; i++) { unsigned v; if (sscanf(s.c_str() + 2 * i, "%2x", &v) != 1) return false; out[i] = (uint8_t)v; } return true; }

// forceReadFileAddress / Eth / XPoint (keyhunt.cpp:7239, :7312, :7392) -> raw 20-byte records
static std::vector<uint8_t> load_targets(const char *fn) {
  FILE *f = fopen(fn, "r");
  if (!f) { fprintf(stderr, "[E] Error opening the file %s\n", fn); exit(EXIT_FAILURE); }
  std::vector<uint8_t> recs;
  char line[1024];
  while (fgets(line, sizeof(line), f)) {
    std::string s = trim(line);
    if (s.empty()) continue;
    uint8_t raw[80];
    bool ok = false;
    if (FLAGMODE == KH_MODE_XPOINT) {
      std::string tok = s.substr(0, s.find_first_of(" \t"));
      if (is_hex(tok)) {
        if (tok.size() == 64 && hex2bin(tok, raw, 32)) { recs.insert(recs.end(), raw, raw + 20); ok = true; }
        else if (tok.size() == 66 && hex2bin(tok.substr(2), raw, 32)) { recs.insert(recs.end(), raw, raw + 20); ok = true; }
        else if (tok.size() == 130 && hex2bin(tok, raw, 65)) { recs.insert(recs.end(), raw + 2, raw + 22); ok = true; }   // reference quirk :7463
      }
    } else if (FLAGCRYPTO == KH_CRYPTO_ETH) {
      if (s.size() == 40 && is_hex(s) && hex2bin(s, raw, 20)) { recs.insert(recs.end(), raw, raw + 20); ok = true; }
      else if (s.size() == 42 && is_hex(s.substr(2)) && hex2bin(s.substr(2), raw, 20)) { recs.insert(recs.end(), raw, raw + 20); ok = true; }
    } else {
      if (s.size() == 40 && is_hex(s) && hex2bin(s, raw, 20)) { recs.insert(recs.end(), raw, raw + 20); ok = true; }
      else if (s.size() > 20 && s.size() < 40 && b58dec25(s.c_str(), raw)) { recs.insert(recs.end(), raw + 1, raw + 21); ok = true; }
    }
    if (!ok) fprintf(stderr, "[I] Ommiting invalid line %s\n", s.c_str());
  }
  fclose(f);
  return recs;
}

// ---------------------------------------------------------------------------------------------------
// -m vanity targets: addvanity (keyhunt.cpp:6739-6860).  A base58 prefix becomes one [A, B] interval of hash160 values per
// address length it can have: the prefix is padded with '1' (lowest digit) for A and with 'z' (highest) for B until the
// decode is 25 bytes long, and bytes 1..20 of each 25-byte decode are the limits.
// ---------------------------------------------------------------------------------------------------
static std::vector<uint8_t> vanity_A, vanity_B;      // flattened 20-byte limits, pair i = [A_i, B_i]
static int vanity_targets = 0;
// the reference decoder's conventions (base58/base58.c:39): the value fills a binsz-byte big-endian buffer; the reported
// length is (bytes after the leading zero bytes) + (number of leading '1' digits); false = bad digit / does not fit
static bool b58_fixed(const std::string &t, size_t binsz, std::vector<uint8_t> &bin, size_t &reported) {
  bin.assign(binsz, 0);
  size_t i = 0, ones = 0;
  while (i < t.size() && t[i] == '1') { ones++; i++; }
  for (; i < t.size(); i++) {
    const char *d = strchr(B58, t[i]);
    if (!d || !t[i]) return false;
    unsigned carry = (unsigned)(d - B58);
    for (size_t j = binsz; j-- > 0;) { unsigned v = bin[j] * 58u + carry; bin[j] = (uint8_t)v; carry = v >> 8; }
    if (carry) return false;
  }
  size_t lead = 0;
  while (lead < binsz && !bin[lead]) lead++;
  reported = binsz - lead + ones;
  return true;
}
static std::vector<std::vector<uint8_t>> vanity_limits(const std::string &prefix, char fill) {
  std::vector<std::vector<uint8_t>> out;
  std::string t = prefix;
  std::vector<uint8_t> bin;
  for (;;) {
    size_t len = 50;
    if (!b58_fixed(t, 50, bin, len)) len = 50;             // a failed decode leaves the length at 50 and ends the loop (:6767)
    if (len > 25 || t.size() > 48) break;
    if (len == 25) {
      size_t l2;
      b58_fixed(t, 25, bin, l2);                            // decoded again into exactly 25 bytes: version | hash160 | checksum
      out.emplace_back(bin.begin() + 1, bin.begin() + 21);
    }
    t.push_back(fill);
  }
  return out;
}
static bool is_b58_string(const std::string &s) { for (char ch : s) if (!ch || !strchr(B58, ch)) return false; return true; }
static int add_vanity(const std::string &target) {
  if (target.size() >= 30) return 0;
  auto a = vanity_limits(target, '1'), b = vanity_limits(target, 'z');
  const size_t r = std::min(a.size(), b.size());
  for (size_t j = 0; j < r; j++) { vanity_A.insert(vanity_A.end(), a[j].begin(), a[j].end()); vanity_B.insert(vanity_B.end(), b[j].begin(), b[j].end()); }
  if (r) vanity_targets++;
  return (int)r;
}
// readFileVanity (keyhunt.cpp:6990): one prefix per line; a missing file is fine when -v gave targets
static void read_vanity_file(const char *fn) {
  FILE *f = fopen(fn, "r");
  if (f) {
    char line[100];
    while (fgets(line, sizeof(line), f)) {
      std::string s = trim(line);
      if (s.empty() || s.size() >= 36) continue;
      if (is_b58_string(s)) add_vanity(s);
      else fprintf(stderr, "[E] the string \"%s\" is not valid Base58, omiting it\n", s.c_str());
    }
    fclose(f);
  }
  if (vanity_targets == 0) { fprintf(stderr, "[E] There aren't any vanity targets\n[E] Unenexpected error\n"); exit(EXIT_FAILURE); }
}

// ---------------------------------------------------------------------------------------------------
// hit output (writekey keyhunt.cpp:6891, writekeyeth :6925)
// ---------------------------------------------------------------------------------------------------
static void write_scan_hit(kh_ctx *c, const kh_hit &h) {
  kh_keyinfo ki;
  if (kh_derive(c, h.key_be, 1, &ki) != KH_OK) die("[E] %s", kh_last_error(c));
  U256 k; memcpy(k.b, h.key_be, 32);
  std::string hexkey = u_hex(k);
  std::lock_guard<std::mutex> g(write_keys);
  FILE *keys = fopen("KEYFOUNDKEYFOUND.txt", "a+");
  if (FLAGCRYPTO == KH_CRYPTO_ETH && FLAGMODE == KH_MODE_ADDRESS) {
    std::string addr = "0x" + hex_of(ki.eth, 20);
    if (keys) { fprintf(keys, "Private Key: %s\naddress: %s\n", hexkey.c_str(), addr.c_str()); fclose(keys); }
    printf("\n Hit!!!! Private Key: %s\naddress: %s\n", hexkey.c_str(), addr.c_str());
  } else {
    bool compressed = (h.kind == KH_HIT_COMP02 || h.kind == KH_HIT_COMP03);
    std::string pub;
    const uint8_t *rmd;
    if (compressed) { pub = std::string((ki.pub_y[31] & 1) ? "03" : "02") + hex_of(ki.pub_x, 32); rmd = ki.h160_comp; }
    else { pub = "04" + hex_of(ki.pub_x, 32) + hex_of(ki.pub_y, 32); rmd = ki.h160_uncomp; }
    std::string addr = rmd160_to_address(rmd), hexrmd = hex_of(rmd, 20);
    if (FLAGMODE == KH_MODE_VANITY) {                                                    // writevanitykey keyhunt.cpp:6705
      if (keys) fclose(keys);
      keys = fopen("VANITYKEYFOUND.txt", "a+");
      if (keys) { fprintf(keys, "Vanity Private Key: %s\npubkey: %s\nAddress %s\nrmd160 %s\n", hexkey.c_str(), pub.c_str(), addr.c_str(), hexrmd.c_str()); fclose(keys); }
      printf("\nVanity Private Key: %s\npubkey: %s\nAddress %s\nrmd160 %s\n", hexkey.c_str(), pub.c_str(), addr.c_str(), hexrmd.c_str());
      fflush(stdout);
      return;
    }
    if (keys) { fprintf(keys, "Private Key: %s\npubkey: %s\nAddress %s\nrmd160 %s\n", hexkey.c_str(), pub.c_str(), addr.c_str(), hexrmd.c_str()); fclose(keys); }
    printf("\nHit! Private Key: %s\npubkey: %s\nAddress %s\nrmd160 %s\n", hexkey.c_str(), pub.c_str(), addr.c_str(), hexrmd.c_str());
  }
  fflush(stdout);
}

// one host thread per GPU: the loop of thread_process (keyhunt.cpp:3309-3858) with the batches on the device.
// A worker of the reference claims ONE chunk of N_SEQUENTIAL_MAX keys per turn; one such chunk (the reference's usual
// -n 0x100000 .. 0x1000000) is far too little work for a GPU call (606,208 walkers x 1024 keys fill one step), so a
// worker claims a RUN of contiguous chunks under the same lock and scans it with one kh_scan.  With stride 1 a run of
// chunks is one arithmetic progression and every claimed chunk is still scanned whole (the overshoot rule of
// keyhunt.cpp:3314-3324), so the scanned key set and the hit records are exactly the reference's.  With a stride the
// reference still advances its cursor by N_SEQUENTIAL_MAX per chunk (:3323), so consecutive chunks are NOT one
// progression and are scanned one by one.
static int n_scan_workers = 1;
static const uint64_t RUN_POINTS = 1ULL << 32;                             // keys per kh_scan the workers aim at
static void scan_worker(kh_ctx *c, int id) {
  const bool stride_one = (u_cmp(stride_v, u_from_u64(1)) == 0);
  for (;;) {
    U256 key;
    uint64_t chunks = 0;
    {
      std::lock_guard<std::mutex> g(write_random);
      if (u_cmp(n_range_start, n_range_end) >= 0) break;                  // keyhunt.cpp:3314
      key = n_range_start;
      uint64_t want = 1;
      if (stride_one && N_SEQUENTIAL_MAX < RUN_POINTS) {
        want = RUN_POINTS / N_SEQUENTIAL_MAX;
        // leave the other GPUs their share when what is left of the range is small
        const U256 left = u_sub(n_range_end, n_range_start);
        bool small = true;
        for (int i = 0; i < 24; i++) small = small && left.b[i] == 0;
        if (small) {
          uint64_t l = 0;
          for (int i = 24; i < 32; i++) l = (l << 8) | left.b[i];
          const uint64_t left_chunks = l / N_SEQUENTIAL_MAX + ((l % N_SEQUENTIAL_MAX) ? 1 : 0);
          const uint64_t share = (left_chunks + (uint64_t)n_scan_workers - 1) / (uint64_t)n_scan_workers;
          if (share < want) want = share ? share : 1;
        }
      }
      while (chunks < want && u_cmp(n_range_start, n_range_end) < 0) {     // every claim is the reference's: test, then Add(N_SEQUENTIAL_MAX) (:3323)
        if (!FLAGQUIET) { if (FLAGMATRIX) printf("Base key: %s thread %i\n", u_hex(n_range_start).c_str(), id); else printf("\rBase key: %s     \r", u_hex(n_range_start).c_str()); }
        n_range_start = u_add(n_range_start, u_from_u64(N_SEQUENTIAL_MAX));
        chunks++;
      }
      if (!FLAGQUIET) fflush(stdout);
    }
    const uint64_t n_points = chunks * N_SEQUENTIAL_MAX;
    if (kh_scan(c, key.b, stride_v.b, n_points) != KH_OK) die("[E] %s", kh_last_error(c));
    total_points += n_points;
    kh_hit hits[64];
    int n = 0;
    do {
      int rc = kh_poll_hits(c, hits, 64, &n);
      if (rc != KH_OK && rc != KH_EOVERFLOW) die("[E] %s", kh_last_error(c));
      if (rc == KH_EOVERFLOW) fprintf(stderr, "\n[W] more hits in the chunks at %s than the device hit buffer holds: the excess was DROPPED; use a smaller -n\n", u_hex(key).c_str());
      for (int i = 0; i < n; i++) write_scan_hit(c, hits[i]);
    } while (n == 64);
  }
}

// ---------------------------------------------------------------------------------------------------
// BSGS table files in the reference's formats (keyhunt.cpp:2504-2652 write, :1983-2236 read)
// ---------------------------------------------------------------------------------------------------
static void blm_header(uint8_t hdr[112], const kh_bloom_desc &d) {
  memset(hdr, 0, 112);
  memcpy(hdr + 0, &d.entries, 8); memcpy(hdr + 8, &d.bits, 8); memcpy(hdr + 16, &d.bytes, 8);
  hdr[24] = (uint8_t)d.hashes;
  long double err = 0.000001;                       // as bloom_init2 receives it
  memcpy(hdr + 32, &err, 10);
  hdr[48] = 1; hdr[49] = 2; hdr[50] = 201;           // ready, BLOOM_VERSION_MAJOR, BLOOM_VERSION_MINOR (bloom.cpp:35-36)
  double bpe = (double)(-logl(err) / (long double)0.480453013918201);
  memcpy(hdr + 56, &bpe, 8);
}
static bool bsgs_save(kh_ctx *c, const kh_bsgs_desc &d) {
  const struct { int tier; int num; uint64_t m; } F3[3] = {{1, 4, d.m}, {2, 6, d.m2}, {3, 7, d.m3}};
  for (auto &t : F3) {
    char fn[128];
    snprintf(fn, sizeof(fn), "keyhunt_bsgs_%d_%" PRIu64 ".blm", t.num, t.m);
    FILE *f = fopen(fn, "wb");
    if (!f) return false;
    printf("[+] Writing bloom filter to file %s ", fn); fflush(stdout);
    std::vector<uint8_t> bf(d.tier[t.tier - 1].bytes);
    uint8_t hdr[112], sum[32];
    blm_header(hdr, d.tier[t.tier - 1]);
    for (int s = 0; s < 256; s++) {
      if (kh_bsgs_export(c, t.tier, s, bf.data(), bf.size()) != KH_OK) die("[E] %s", kh_last_error(c));
      sha256_host(bf.data(), bf.size(), sum);
      fwrite(hdr, 112, 1, f); fwrite(bf.data(), bf.size(), 1, f); fwrite(sum, 32, 1, f); fwrite(sum, 32, 1, f);
      if (s % 64 == 0) { printf("."); fflush(stdout); }
    }
    fclose(f);
    printf(" Done!\n");
  }
  char fn[128];
  snprintf(fn, sizeof(fn), "keyhunt_bsgs_2_%" PRIu64 ".tbl", d.m3);
  FILE *f = fopen(fn, "wb");
  if (!f) return false;
  std::vector<uint8_t> tab(d.m3 * 16);
  uint8_t sum[32];
  if (kh_bsgs_export(c, 0, 0, tab.data(), tab.size()) != KH_OK) die("[E] %s", kh_last_error(c));
  sha256_host(tab.data(), tab.size(), sum);
  fwrite(tab.data(), tab.size(), 1, f); fwrite(sum, 32, 1, f);
  fclose(f);
  printf("[+] Writing bP Table to file %s .. Done!\n", fn);
  return true;
}
static bool bsgs_load(kh_ctx *c, const kh_bsgs_desc &d) {   // all four files must exist and verify, else rebuild
  const struct { int tier; int num; uint64_t m; } F3[3] = {{1, 4, d.m}, {2, 6, d.m2}, {3, 7, d.m3}};
  for (auto &t : F3) {
    char fn[128];
    snprintf(fn, sizeof(fn), "keyhunt_bsgs_%d_%" PRIu64 ".blm", t.num, t.m);
    FILE *f = fopen(fn, "rb");
    if (!f) return false;
    const uint64_t nb = d.tier[t.tier - 1].bytes;
    std::vector<uint8_t> rec(112 + nb + 64);
    uint8_t sum[32];
    for (int s = 0; s < 256; s++) {
      if (fread(rec.data(), rec.size(), 1, f) != 1) { fclose(f); return false; }
      uint64_t bytes; memcpy(&bytes, rec.data() + 16, 8);
      if (bytes != nb) { fclose(f); return false; }
      if (!FLAGSKIPCHECKSUM) { sha256_host(rec.data() + 112, nb, sum); if (memcmp(sum, rec.data() + 112 + nb, 32) || memcmp(sum, rec.data() + 144 + nb, 32)) { fprintf(stderr, "[E] Error checksum file mismatch! %s\n", fn); exit(EXIT_FAILURE); } }
      if (kh_bsgs_import(c, t.tier, s, rec.data() + 112, nb) != KH_OK) die("[E] %s", kh_last_error(c));
    }
    fclose(f);
    printf("[+] Reading bloom filter from file %s .... Done!\n", fn);
  }
  char fn[128];
  snprintf(fn, sizeof(fn), "keyhunt_bsgs_2_%" PRIu64 ".tbl", d.m3);
  FILE *f = fopen(fn, "rb");
  if (!f) return false;
  std::vector<uint8_t> tab(d.m3 * 16 + 32);
  uint8_t sum[32];
  if (fread(tab.data(), tab.size(), 1, f) != 1) { fclose(f); return false; }
  fclose(f);
  if (!FLAGSKIPCHECKSUM) { sha256_host(tab.data(), d.m3 * 16, sum); if (memcmp(sum, tab.data() + d.m3 * 16, 32)) { fprintf(stderr, "[E] Error checksum file mismatch! %s\n", fn); exit(EXIT_FAILURE); } }
  if (kh_bsgs_import(c, 0, 0, tab.data(), d.m3 * 16) != KH_OK) die("[E] %s", kh_last_error(c));
  printf("[+] Reading bP Table from file %s .... Done!\n", fn);
  return true;
}


// ---------------------------------------------------------------------------------------------------
// scan-mode target cache `data_<first 4 bytes of sha256(target file)>.dat` (writeFileIfNeeded keyhunt.cpp:7756,
// readFileAddress :7043-7206): 32 B sha256(bf) | struct bloom (112 B) | bf | 32 B sha256(table) | u64 size | table
// ---------------------------------------------------------------------------------------------------
static bool sha256_of_file(const char *fn, uint8_t out[32]) {
  FILE *f = fopen(fn, "rb");
  if (!f) return false;
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  fclose(f);
  sha256_host(buf.data(), buf.size(), out);
  return true;
}
static std::string dat_name(const char *targets) {
  uint8_t sum[32];
  if (!sha256_of_file(targets, sum)) { fprintf(stderr, "[E] sha256_file error\n"); exit(EXIT_FAILURE); }
  return "data_" + hex_of(sum, 4) + ".dat";
}
static bool dat_read(const std::string &fn, kh_bloom_desc &d, std::vector<uint8_t> &bf, std::vector<uint8_t> &table) {
  FILE *f = fopen(fn.c_str(), "rb");
  if (!f) return false;
  printf("[+] Reading file %s\n", fn.c_str());
  uint8_t bsum[32], dsum[32], hdr[112], sum[32];
  uint64_t size = 0;
  bool ok = fread(bsum, 1, 32, f) == 32 && fread(hdr, 1, 112, f) == 112;
  if (ok) {
    memcpy(&d.entries, hdr, 8); memcpy(&d.bits, hdr + 8, 8); memcpy(&d.bytes, hdr + 16, 8);
    d.hashes = hdr[24]; d.pad = 0;
    ok = d.bytes < (1ULL << 40) && d.bits && d.hashes;
  }
  if (ok) { bf.resize(d.bytes); ok = fread(bf.data(), 1, d.bytes, f) == d.bytes; }
  if (ok) ok = fread(dsum, 1, 32, f) == 32 && fread(&size, 1, 8, f) == 8 && size % 20 == 0 && size < (1ULL << 40);
  if (ok) { table.resize(size); ok = size == 0 || fread(table.data(), 1, size, f) == size; }
  fclose(f);
  if (!ok) { fprintf(stderr, "[E] Error reading file %s\n", fn.c_str()); return false; }
  if (!FLAGSKIPCHECKSUM) {
    sha256_host(bf.data(), bf.size(), sum);
    if (memcmp(sum, bsum, 32)) { fprintf(stderr, "[E] Error checksum mismatch (bloom) in %s\n", fn.c_str()); return false; }
    sha256_host(table.data(), table.size(), sum);
    if (memcmp(sum, dsum, 32)) { fprintf(stderr, "[E] Error checksum mismatch (data) in %s\n", fn.c_str()); return false; }
  }
  printf("[+] Bloom filter for %" PRIu64 " elements.\n", d.entries);
  return true;
}
static void dat_write(const std::string &fn, kh_ctx *c) {
  kh_bloom_desc d;
  if (kh_get_bloom(c, &d, NULL, 0) != KH_OK) die("[E] %s", kh_last_error(c));
  std::vector<uint8_t> bf(d.bytes);
  uint64_t n = 0;
  if (kh_get_bloom(c, &d, bf.data(), bf.size()) != KH_OK || kh_get_table(c, NULL, 0, &n) != KH_OK) die("[E] %s", kh_last_error(c));
  std::vector<uint8_t> table(n * 20 + 1);
  if (kh_get_table(c, table.data(), n, &n) != KH_OK) die("[E] %s", kh_last_error(c));
  FILE *f = fopen(fn.c_str(), "wb");
  if (!f) return;
  uint64_t size = n * 20;
  printf("[D] size data %" PRIu64 "\n[+] Writing file %s ", size, fn.c_str());
  uint8_t sum[32], hdr[112];
  sha256_host(bf.data(), bf.size(), sum);
  fwrite(sum, 1, 32, f);
  blm_header(hdr, d);
  fwrite(hdr, 1, 112, f);
  fwrite(bf.data(), 1, bf.size(), f);
  sha256_host(table.data(), size, sum);
  fwrite(sum, 1, 32, f);
  fwrite(&size, 1, 8, f);
  fwrite(table.data(), 1, size, f);
  fclose(f);
  printf("........\n");
}

#ifdef KH_BSGSD
#include "bsgsd_serve.hpp"
#endif

// ---------------------------------------------------------------------------------------------------
static void menu() {
#ifdef KH_BSGSD
  printf("\nUsage: keyhunt-b200-bsgsd [-k factor] [-n N] [-t gpus] [-i ip] [-p port] [-S] [-6]\n"
         "GPU (B200) drop-in for keyhunt's bsgsd: builds (or reads) the BSGS tables once, keeps them in HBM and answers\n"
         "\"<pubkey> <from>:<to>\" lines / HTTP POST JSON requests on ip:port (default 127.0.0.1:8080).\n"
         "-S reads / writes the reference's keyhunt_bsgs_*.blm / .tbl files (the reference server always does).\n");
  exit(EXIT_FAILURE);
#endif
  printf("\nUsage: keyhunt-b200 -m address|rmd160|xpoint|bsgs|vanity -f file [-v prefix] [-r A:B | -b bits] [-l compress|uncompress|both] [-c btc|eth]\n"
         "                    [-k factor] [-n N] [-t gpus] [-I stride] [-s seconds] [-q] [-M] [-S] [-6] [-z mult]\n"
         "GPU (B200) drop-in for keyhunt's key-range search; same flags, -t selects the number of GPUs.\n");
  exit(EXIT_FAILURE);
}

int main(int argc, char **argv) {
  static struct option long_options[] = {
      {"mapped", optional_argument, 0, 0}, {"mapped-size", required_argument, 0, 0}, {"mapped-chunks", required_argument, 0, 0},
      {"bloom-file", required_argument, 0, 0}, {"load-bloom", no_argument, 0, 0}, {"ptable", required_argument, 0, 0},
      {"ptable-size", required_argument, 0, 0}, {"load-ptable", no_argument, 0, 0}, {"ptable-cache", no_argument, 0, 0},
      {"bloom-bytes", required_argument, 0, 0}, {"create-mapped", optional_argument, 0, 0}, {"tmpdir", required_argument, 0, 0},
      {"bsgs-block-count", required_argument, 0, 0}, {"bsgs-block-size", required_argument, 0, 0}, {"rmd-batch-size", required_argument, 0, 0},
      {0, 0, 0, 0}};
  printf("[+] Version 0.2.230519 Satoshi Quest (keyhunt-b200: CUDA sm_100a back end)\n");
  const char *range_arg = nullptr, *str_stride = nullptr;
  int bitrange = 0, c, oi = 0;
#ifdef KH_BSGSD
  const char *listen_ip = "127.0.0.1";
  int listen_port = 8080;
  FLAGMODE = KH_MODE_BSGS;
  // The reference server always reads/writes the .blm/.tbl files (bsgsd.cpp:238 FLAGSAVEREADFILE = 1).  Here the
  // tables are rebuilt in HBM faster than a disk can deliver them (-k 4096: 62 GB in 18 s), so files are opt-in: -S.
#endif
  stride_v = u_from_u64(1);
  while ((c = getopt_long(argc, argv, "deh6MqRSB:b:c:C:E:f:I:i:k:l:m:N:n:p:r:s:t:v:G:8:z:", long_options, &oi)) != -1) {
    switch (c) {
      case 0: fprintf(stderr, "[I] --%s accepted and ignored: bloom filters and the bP table are resident in GPU memory\n", long_options[oi].name); break;
      case 'h': menu(); break;
      case '6': FLAGSKIPCHECKSUM = 1; break;
      case 'M': FLAGMATRIX = 1; printf("[+] Matrix screen\n"); break;
      case 'q': FLAGQUIET = 1; printf("[+] Quiet thread output\n"); break;
      case 'S': FLAGSAVEREADFILE = 1; break;
      case 'R': die("[E] -R (random mode) uses the OS RNG and is not reproducible; not supported by the GPU back end");
      case 'e': FLAGENDOMORPHISM = 1; printf("[+] Endomorphism enabled\n"); break;
      case 'B':   // bsgs_modes keyhunt.cpp:423; random / dance draw from the OS RNG, ggsb / angrygiant re-order one window's giant steps
        if (!strcmp(optarg, "sequential")) FLAGBSGSMODE = 0;
        else if (!strcmp(optarg, "backward")) FLAGBSGSMODE = 1;
        else if (!strcmp(optarg, "both")) FLAGBSGSMODE = 2;
        else die("[E] -B %s is not supported by the GPU back end (sequential, backward, both)", optarg);
        break;
      case 'b': bitrange = atoi(optarg); if (bitrange > 0 && bitrange <= 256) FLAGBITRANGE = 1; else fprintf(stderr, "[E] invalid bits param: %s.\n", optarg); break;
      case 'c': if (!strcmp(optarg, "btc")) FLAGCRYPTO = KH_CRYPTO_BTC; else if (!strcmp(optarg, "eth")) { FLAGCRYPTO = KH_CRYPTO_ETH; printf("[+] Setting search for ETH adddress.\n"); } else die("[E] Unknow crypto value %s", optarg); break;
      case 'f': fileName = optarg; break;
      case 'I': str_stride = optarg; break;
      case 'k': KFACTOR = atoi(optarg); if (KFACTOR <= 0) KFACTOR = 1; printf("[+] K factor %i\n", KFACTOR); break;
      case 'l':
        if (!strcmp(optarg, "uncompress")) { FLAGSEARCH = KH_SEARCH_UNCOMPRESS; printf("[+] Search uncompress only\n"); }
        else if (!strcmp(optarg, "compress")) { FLAGSEARCH = KH_SEARCH_COMPRESS; printf("[+] Search compress only\n"); }
        else if (!strcmp(optarg, "both")) { FLAGSEARCH = KH_SEARCH_BOTH; printf("[+] Search both compress and uncompress\n"); }
        break;
      case 'm':
        if (!strcmp(optarg, "xpoint")) { FLAGMODE = KH_MODE_XPOINT; printf("[+] Mode xpoint\n"); }
        else if (!strcmp(optarg, "address")) { FLAGMODE = KH_MODE_ADDRESS; printf("[+] Mode address\n"); }
        else if (!strcmp(optarg, "bsgs")) { FLAGMODE = KH_MODE_BSGS; }
        else if (!strcmp(optarg, "rmd160")) { FLAGMODE = KH_MODE_RMD160; FLAGCRYPTO = KH_CRYPTO_BTC; printf("[+] Mode rmd160\n"); }
        else if (!strcmp(optarg, "vanity")) { FLAGMODE = KH_MODE_VANITY; printf("[+] Mode vanity\n"); }
        else die("[E] mode %s is not part of the GPU path (address, rmd160, xpoint, bsgs, vanity)", optarg);
        break;
      case 'n': FLAG_N = 1; str_N = optarg; break;
      case 'r': range_arg = optarg; FLAGRANGE = 1; break;
      case 's': OUTPUTSECONDS = atoi(optarg); if (OUTPUTSECONDS < 0) OUTPUTSECONDS = 30; if (!OUTPUTSECONDS) printf("[+] Turn off stats output\n"); else printf("[+] Stats output every %d seconds\n", OUTPUTSECONDS); break;
      case 't': NGPUS = atoi(optarg); if (NGPUS <= 0) NGPUS = 1; printf("[+] GPUs : %d\n", NGPUS); break;
      case 'z': FLAGBLOOMMULTIPLIER = atoi(optarg); if (FLAGBLOOMMULTIPLIER <= 0) FLAGBLOOMMULTIPLIER = 1; printf("[+] Bloom Size Multiplier %i\n", FLAGBLOOMMULTIPLIER); break;
#ifdef KH_BSGSD
      case 'p': listen_port = atoi(optarg); break;
      case 'i': listen_ip = optarg; break;
#else
      case 'p': case 'i': break;
#endif
      case 'v':                                                                     // keyhunt.cpp:1083-1099
        if (is_b58_string(optarg)) { if (add_vanity(optarg) > 0) printf("[+] Added Vanity search : %s\n", optarg); else printf("[+] Vanity search \"%s\" was NOT Added\n", optarg); }
        else fprintf(stderr, "[+] The string \"%s\" is not Valid Base58\n", optarg);
        break;
      case 'd': case 'C': case 'E': case 'N': case 'G': case '8': break;   // accepted, no effect here
      default: menu();
    }
  }
  // -n is validated in every mode (keyhunt.cpp:1173-1183)
  uint64_t nk_n = 0x100000000000ULL;
  if (FLAG_N) nk_n = (str_N[0] == '0' && (str_N[1] == 'x' || str_N[1] == 'X')) ? strtoull(str_N + 2, NULL, 16) : strtoull(str_N, NULL, 10);
  if (!validate_nk(nk_n, (uint64_t)KFACTOR)) exit(EXIT_FAILURE);
  if (str_stride) {
    if (FLAGMODE == KH_MODE_BSGS) die("[E] Stride doesn't work with BSGS");
    if (str_stride[0] == '0' && str_stride[1] == 'x') { if (!u_from_hex(stride_v, str_stride + 2)) die("[E] bad stride"); }
    else stride_v = u_from_u64(strtoull(str_stride, NULL, 10));
    printf("[+] Stride : %s\n", str_stride);
  }
  if (FLAGMODE == KH_MODE_ADDRESS && FLAGCRYPTO == 0) { FLAGCRYPTO = KH_CRYPTO_BTC; printf("[+] Setting search for btc adddress\n"); }
  if (FLAGMODE == KH_MODE_BSGS) printf("[+] Mode BSGS %s\n", FLAGBSGSMODE == 0 ? "sequential" : FLAGBSGSMODE == 1 ? "backward" : "both");
  if (FLAGMODE == KH_MODE_BSGS && FLAGENDOMORPHISM) die("[E] Endomorphism doesn't work with BSGS");

  // range (keyhunt.cpp:1221-1262, :854-873)
  U256 order; memcpy(order.b, ORDER_N, 32);
  if (FLAGRANGE) {
    std::string r(range_arg);
    size_t p = r.find(':');
    if (p == std::string::npos) die("[E] bad -r, expected A:B");
    if (!u_from_hex(n_range_start, r.substr(0, p).c_str()) || !u_from_hex(n_range_end, r.substr(p + 1).c_str())) die("[E] bad -r, expected hex A:B");
    if (u_is_zero(n_range_start)) n_range_start = u_from_u64(1);
    if (u_cmp(n_range_start, n_range_end) == 0) die("[E] Start and End range can't be the same");
    if (u_cmp(n_range_start, n_range_end) > 0) { fprintf(stderr, "[W] Opps, start range can't be great than end range. Swapping them\n"); std::swap(n_range_start, n_range_end); }
    if (u_cmp(n_range_end, order) > 0) die("[E] Start and End range can't be great than N");
  } else if (FLAGBITRANGE) {
    n_range_start = u_from_u64(1);
    for (int i = 0; i < bitrange - 1; i++) n_range_start = u_shl1(n_range_start);
    n_range_end = (bitrange == 256) ? order : u_shl1(n_range_start);
    if (u_cmp(n_range_end, order) > 0) n_range_end = order;
  } else {
#ifdef KH_BSGSD
    n_range_start = u_from_u64(1); n_range_end = order;   // every request carries its own range
#else
    die("[E] a range is required: -r A:B or -b bits (random start needs the OS RNG and is not reproducible)");
#endif
  }

  int ndev = NGPUS;
  std::vector<kh_ctx *> gpus;
  for (int g = 0; g < ndev; g++) {
    kh_ctx *ctx = nullptr;
    if (kh_create(&ctx, g) != KH_OK) { if (g == 0) die("[E] no usable CUDA device (this build has no CPU fallback)"); break; }
    gpus.push_back(ctx);
  }
  {
    char name[128]; int sms = 0; uint64_t mem = 0;
    kh_device_info(gpus[0], name, sizeof(name), &sms, &mem);
    printf("[+] %zu x %s (%d SMs, %.0f GB)\n", gpus.size(), name, sms, (double)mem / 1e9);
  }
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);

  if (FLAGMODE != KH_MODE_BSGS) {
    if (FLAG_N) {
      N_SEQUENTIAL_MAX = nk_n;
      if (N_SEQUENTIAL_MAX < 1024 || N_SEQUENTIAL_MAX % 1024) { fprintf(stderr, "[I] n value need to be equal or great than 1024 and multiplier of 1024, back to defaults\n"); N_SEQUENTIAL_MAX = 0x100000000ULL; }
    }
    printf("[+] N = 0x%" PRIx64 "\n", N_SEQUENTIAL_MAX);
    if (FLAGBITRANGE) printf("[+] Bit Range %i\n", bitrange); else printf("[+] Range \n");
    printf("[+] -- from : 0x%s\n[+] -- to   : 0x%s\n", u_hex(n_range_start).c_str(), u_hex(n_range_end).c_str());
    if (FLAGMODE == KH_MODE_VANITY) {
      FLAGCRYPTO = KH_CRYPTO_BTC;
      read_vanity_file(fileName);
      for (kh_ctx *g : gpus) {
        if (kh_set_option(g, "endomorphism", FLAGENDOMORPHISM) != KH_OK) die("[E] %s", kh_last_error(g));
        if (kh_set_option(g, "hit_capacity", 1 << 22) != KH_OK) die("[E] %s", kh_last_error(g));   // short prefixes match often
        if (kh_set_vanity(g, FLAGSEARCH, vanity_A.data(), vanity_B.data(), vanity_A.size() / 20) != KH_OK) die("[E] %s", kh_last_error(g));
      }
    }
    std::vector<uint8_t> recs, cached_bf;
    kh_bloom_desc d;
    bool from_cache = false;
    if (FLAGMODE != KH_MODE_VANITY) {
    std::string cache = FLAGSAVEREADFILE ? dat_name(fileName) : std::string();
    if (FLAGSAVEREADFILE && dat_read(cache, d, cached_bf, recs)) from_cache = true;   // the reference's -S cache (bloom image reused bit for bit)
    else recs = load_targets(fileName);
    const uint64_t N = recs.size() / 20;
    printf("[+] Allocating memory for %" PRIu64 " elements: %.2f MB\n", N, (double)(20.0 * N / 1048576.0));
    if (!from_cache) {
      uint64_t items = (N <= 10000) ? 10000 : (uint64_t)FLAGBLOOMMULTIPLIER * N;   // initBloomFilter keyhunt.cpp:7608
      kh_bloom_params(items, &d);
      printf("[+] Bloom filter for %" PRIu64 " elements.\n", N);
    }
    printf("[+] Loading data to the bloomfilter total: %.2f MB\n", (double)d.bytes / 1048576.0);
    for (kh_ctx *g : gpus) {
      if (kh_set_option(g, "endomorphism", FLAGENDOMORPHISM) != KH_OK) die("[E] %s", kh_last_error(g));
      if (kh_set_targets(g, FLAGMODE, FLAGCRYPTO, FLAGSEARCH, recs.data(), N, &d, from_cache ? cached_bf.data() : NULL) != KH_OK) die("[E] %s", kh_last_error(g));
    }
    if (FLAGSAVEREADFILE && !from_cache) dat_write(cache, gpus[0]);
    printf("[+] Sorting data ... done! %" PRIu64 " values were loaded and sorted\n", N);
    }
    fflush(stdout);
    std::vector<std::thread> th;
    n_scan_workers = (int)gpus.size();
    for (size_t g = 0; g < gpus.size(); g++) th.emplace_back(scan_worker, gpus[g], (int)g);
    for (auto &t : th) t.join();
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double secs = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    uint64_t shown = total_points.load();                                                                // keyhunt.cpp:2883-2891
    if (FLAGENDOMORPHISM) shown *= (FLAGMODE == KH_MODE_XPOINT) ? 3 : 6;
    else if (FLAGSEARCH == KH_SEARCH_COMPRESS) shown *= 2;
    printf("\r[+] Total %" PRIu64 " keys in %.0f seconds: ~%.0f Mkeys/s (%.0f keys/s)\n", shown, secs, shown / secs / 1e6, shown / secs);
    printf("\nEnd\n");
  } else {
    // ---- BSGS ----------------------------------------------------------------------------------------
    std::vector<std::vector<uint8_t>> pubs;   // 64-byte X||Y
    std::vector<bool> pub_compressed;
#ifndef KH_BSGSD
    FILE *f = fopen(fileName, "rb");
    if (!f) { fprintf(stderr, "[E] Can't open file %s\n", fileName); exit(EXIT_FAILURE); }
    printf("[+] Opening file %s\n", fileName);
    char line[1024];
    while (fgets(line, sizeof(line), f)) {
      std::string s = trim(line);
      if (s.size() < 66) continue;
      std::string tok = s.substr(0, s.find_first_of(" \t"));
      uint8_t raw[65], xy[64];
      if (tok.size() == 66 && is_hex(tok) && hex2bin(tok, raw, 33) && (raw[0] == 2 || raw[0] == 3) && decompress_pub(raw + 1, raw[0] & 1, xy + 32)) {
        memcpy(xy, raw + 1, 32); pubs.emplace_back(xy, xy + 64); pub_compressed.push_back(true);
      } else if (tok.size() == 130 && is_hex(tok) && hex2bin(tok, raw, 65) && raw[0] == 4) {
        pubs.emplace_back(raw + 1, raw + 65); pub_compressed.push_back(false);
      } else printf("Invalid length: %s\n", tok.c_str());
    }
    fclose(f);
    if (pubs.empty()) die("[E] The file don't have any valid publickeys");
    printf("[+] Added %zu points from file\n", pubs.size());
    printf("[+] Range \n[+] -- from : 0x%s\n[+] -- to   : 0x%s\n", u_hex(n_range_start).c_str(), u_hex(n_range_end).c_str());
#endif
    kh_bsgs_desc d;
    // build (or load with -S) on every GPU: each holds its own replica of the tables
    for (size_t g = 0; g < gpus.size(); g++) {
      if (kh_bsgs_build(gpus[g], nk_n, (uint32_t)KFACTOR) != KH_OK) die("[E] %s", kh_last_error(gpus[g]));
      kh_bsgs_describe(gpus[g], &d);
      if (g == 0) {
        printf("[+] N = 0x%" PRIx64 "\n", d.n);
        printf("[+] Bloom filter for %" PRIu64 " elements : %.2f MB\n", d.m, 256.0 * d.tier[0].bytes / 1048576.0);
        printf("[+] Bloom filter for %" PRIu64 " elements : %.2f MB\n", d.m2, 256.0 * d.tier[1].bytes / 1048576.0);
        printf("[+] Bloom filter for %" PRIu64 " elements : %.2f MB\n", d.m3, 256.0 * d.tier[2].bytes / 1048576.0);
        printf("[+] Allocating %.2f MB for %" PRIu64 " bP Points\n", (double)(d.m3 * 16) / 1048576.0, d.m3);
      }
      if (FLAGSAVEREADFILE) {
        if (bsgs_load(gpus[g], d)) { if (g == 0) printf("[+] tables loaded from files\n"); }
        else if (g == 0) bsgs_save(gpus[g], d);
      }
    }
#ifdef KH_BSGSD
    {
      for (kh_ctx *g : gpus) kh_set_option(g, "bsgs_base_check", 1);
      const int rc = bsgsd_serve(gpus, d, listen_ip, listen_port);
      for (kh_ctx *g : gpus) kh_destroy(g);
      return rc;
    }
#endif
    // Window pickers (-B).  A segment {from, to} stands for the windows based at from, from+2N, ... while below `to`
    // (what one kh_bsgs_search call walks); the pickers differ only in how the range is cut into segments:
    //   sequential (thread_process_bsgs :4549)          windows aligned at the range start
    //   backward   (thread_process_bsgs_backward :5953) windows aligned at the range END, taken downwards; a last one
    //                                                   clipped to the range start when the width is not a multiple of 2N
    //   both       (thread_process_bsgs_both :6211)     the reference flips rand()%2 between the two cursors per window; here
    //                                                   the lower half of the windows comes from the bottom cursor and the
    //                                                   upper half from the top one, one of the coverings it can produce
    // Windows of a segment are dealt to the GPUs in contiguous blocks; each key is searched until found.
    const U256 two_n = u_mul_u64(u_from_u64(d.n), 2);
    std::unique_ptr<std::atomic<int>[]> found(new std::atomic<int>[pubs.size()]);
    for (size_t k = 0; k < pubs.size(); k++) found[k] = 0;
    std::atomic<int> nfound{0};
    const U256 width = u_sub(n_range_end, n_range_start);
    for (int i = 0; i < 16; i++) if (width.b[i]) die("[E] BSGS ranges wider than 2^128 keys are not supported");
    unsigned __int128 w128 = 0;
    for (int i = 16; i < 32; i++) w128 = (w128 << 8) | width.b[i];
    const unsigned __int128 step128 = (unsigned __int128)2 * d.n;
    if ((w128 + step128 - 1) / step128 > ((unsigned __int128)1 << 50)) die("[E] more than 2^50 BSGS windows in the range: raise -n / -k");
    const uint64_t q_full = (uint64_t)(w128 / step128), windows = (uint64_t)((w128 + step128 - 1) / step128);
    const bool ragged = (w128 % step128) != 0;
    struct Seg { U256 from, to; uint64_t windows; };
    std::vector<Seg> segs;
    if (FLAGBSGSMODE == 0) segs.push_back({n_range_start, n_range_end, windows});
    else if (FLAGBSGSMODE == 1) {
      if (q_full) segs.push_back({u_sub(n_range_end, u_mul_u64(two_n, q_full)), n_range_end, q_full});
      if (ragged) segs.push_back({n_range_start, u_add(n_range_start, u_from_u64(1)), 1});
    } else {
      const uint64_t bottom = (windows + 1) / 2, top = windows - bottom;
      U256 bt = u_add(n_range_start, u_mul_u64(two_n, bottom));
      if (u_cmp(bt, n_range_end) > 0) bt = n_range_end;
      segs.push_back({n_range_start, bt, bottom});
      if (top) segs.push_back({u_sub(n_range_end, u_mul_u64(two_n, top)), n_range_end, top});
    }
    auto worker = [&](size_t g) {
      for (const Seg &sg : segs) {
        const uint64_t base = sg.windows / gpus.size(), rem = sg.windows % gpus.size();
        const uint64_t first = g * base + std::min<uint64_t>(g, rem), cnt = base + (g < rem ? 1 : 0);
        if (!cnt) continue;
        const U256 from = u_add(sg.from, u_mul_u64(two_n, first));
        U256 to = u_add(from, u_mul_u64(two_n, cnt));
        if (first + cnt == sg.windows) to = sg.to;                 // the last window of a segment keeps the reference's overshoot past `to`
        for (size_t k = 0; k < pubs.size(); k++) {
          if (found[k].load()) continue;
          uint8_t key[32]; int fnd = 0;
          if (!FLAGQUIET) { printf("[+] Thread 0x%s \n", u_hex(from).c_str()); fflush(stdout); }
          if (kh_bsgs_search(gpus[g], pubs[k].data(), from.b, to.b, key, &fnd) != KH_OK) die("[E] %s", kh_last_error(gpus[g]));
          if (fnd && !found[k].exchange(1)) {
            U256 kk; memcpy(kk.b, key, 32);
            std::string pubhex = pub_compressed[k] ? (std::string((pubs[k][63] & 1) ? "03" : "02") + hex_of(pubs[k].data(), 32)) : ("04" + hex_of(pubs[k].data(), 64));
            std::lock_guard<std::mutex> lk(write_keys);
            printf("[+] Thread Key found privkey %s   \n[+] Publickey %s\n", u_hex(kk).c_str(), pubhex.c_str());
            FILE *fk = fopen("KEYFOUNDKEYFOUND.txt", "a");
            if (fk) { fprintf(fk, "Key found privkey %s\nPublickey %s\n", u_hex(kk).c_str(), pubhex.c_str()); fclose(fk); }
            nfound++;
          }
        }
      }
    };
    std::vector<std::thread> th;
    for (size_t g = 0; g < gpus.size(); g++) th.emplace_back(worker, g);
    for (auto &t : th) t.join();
    if (nfound.load() == (int)pubs.size()) { printf("All points were found\n"); for (kh_ctx *g : gpus) kh_destroy(g); exit(EXIT_FAILURE); }   // sic: keyhunt.cpp:4855-4858
    printf("\nEnd\n");
  }
  for (kh_ctx *g : gpus) kh_destroy(g);
  return 0;
}
