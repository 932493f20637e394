// plan.hpp — host-side planning of a scan: which 1024-point batches have no shared inverse, and how the batches are dealt to
// walker threads so that the kernels never meet one in the middle of a walk.  Host code only (kh_scan.cu and tests/devsim).
//
// The walk (walk.cuh) inverts the 513 differences dx_e = (e*S).x - C.x of a batch together, entry 0 being the walker's hop
// W = T*1024*S.  A zero difference makes the shared product zero:
//   * C = +-e*S, e = 1..512 ("collapsed" batch; only for ranges that touch key 0 mod n).  The reference has exactly the same
//     weakness: IntGroup::ModInv then yields zeros and the 1023 non-centre points of the batch are deterministic garbage
//     (SURVEY App. B.11).  The kernels produce the SAME garbage (tests/test_devsim.py compares it with the oracle), but the
//     walker's next centre would be garbage too — so a collapsed batch is walked alone, as the last step of its walker.
//   * C = +-W (the hop; e.g. `-r 200:...` puts batch T-1 exactly on W).  The reference has no hop, its batch is fine: the
//     range is cut at that batch and the rest is walked with another T (another W), so the coincidence never happens.
// Both cases are predicted exactly here: the centre of batch b is (k0 + (1024 b + 512) s) G, so C = m*S  <=>
// b = (m - 512 - k0/s) / 1024 (mod n), for the 1026 values m = +-1..+-512, +-1024 T.  The kernels stay free of any test (an
// in-kernel test cost the C2 kernel 3-5 %: the hash kernels sit at the 128-register cap).
// BSGS giant walks are not planned: their centres depend on the unknown key; a batch collapses there only when the key
// sits exactly on the giant-step grid, which the reference's batch does not survive either.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "ec.cuh"

namespace kh {

struct ScanSegment {
  uint64_t first, end;   // batches [first, end) of the scan
  uint64_t T;            // walker threads (multiple of t_align)
  uint32_t collapsed;    // 1: the single batch of this segment has no shared inverse
};

namespace plan_detail {
static inline bool u256_is_zero(const u256 &a) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.v[i]; return o == 0; }
static inline int u256_cmp(const u256 &a, const u256 &b) { for (int i = 7; i >= 0; i--) if (a.v[i] != b.v[i]) return a.v[i] < b.v[i] ? -1 : 1; return 0; }
static inline u256 order_n() { u256 n = {KH_N}; return n; }
static inline u256 mod_n(const u256 &a) { u256 n = order_n(), d; if (!kh_sub8(d.v, a.v, n.v)) return d; return a; }   // a < 2^256 < 2n
static inline u256 addmod(const u256 &a, const u256 &b) {   // a, b < n
  u256 r, n = order_n(), d;
  const uint32_t cf = kh_add8(r.v, a.v, b.v);
  if (cf || !kh_sub8(d.v, r.v, n.v)) { if (cf) kh_sub8(d.v, r.v, n.v); return d; }
  return r;
}
static inline u256 negmod(const u256 &a) { if (u256_is_zero(a)) return a; u256 r; u256_neg_mod_n(r, a); return r; }
static inline u256 submod(const u256 &a, const u256 &b) { return addmod(a, negmod(b)); }
static inline u256 mulmod(const u256 &a, const u256 &b) { u256 r; u256_mulmod_n(r, a, b); return r; }
static inline u256 from_u64(uint64_t v) { u256 r; u256_set_u64(r, v); return r; }
static inline u256 invmod(const u256 &a) {   // a^(n-2) mod n
  u256 e = order_n(), two = from_u64(2), r = from_u64(1), base = a;
  kh_sub8(e.v, e.v, two.v);
  for (int i = 0; i < 256; i++) {
    if ((e.v[i >> 5] >> (i & 31)) & 1) r = mulmod(r, base);
    base = mulmod(base, base);
  }
  return r;
}
// value < limit ?  (value is a residue mod n, limit a 64-bit count)
static inline bool below(const u256 &v, uint64_t limit, uint64_t &out) {
  for (int i = 2; i < 8; i++) if (v.v[i]) return false;
  out = ((uint64_t)v.v[1] << 32) | v.v[0];
  return out < limit;
}
}  // namespace plan_detail

// Degenerate batches of the scan k0 + i*s, batches [0, n_batches): collapsed ones (sorted) into `collapsed`; returns false if the
// stride is 0 mod n.  hop_batch(T, sign) is then answered by plan_hop().
struct ScanGeometry {
  u256 c;        // (k0 / s + 512) mod n
  u256 inv1024;  // 1024^-1 mod n
  bool ok;
};
static inline ScanGeometry plan_geometry(const u256 &k0, const u256 &s) {
  using namespace plan_detail;
  ScanGeometry g;
  const u256 sm = mod_n(s);
  g.ok = !u256_is_zero(sm);
  if (!g.ok) return g;
  u256 q = mod_n(k0);
  const u256 one = from_u64(1);
  if (u256_cmp(sm, one) != 0) {
    // the inverse of the stride costs ~450 slow modular multiplications: remembered per host thread (one thread per GPU)
    static thread_local u256 last_s = {{0, 0, 0, 0, 0, 0, 0, 0}}, last_inv = {{0, 0, 0, 0, 0, 0, 0, 0}};
    if (u256_cmp(last_s, sm) != 0) { last_inv = invmod(sm); last_s = sm; }
    q = mulmod(q, last_inv);
  }
  g.c = addmod(q, from_u64(512));
  // 1024^-1 mod n (little-endian limbs; checked against 1024 * x = 1 mod n by tests/test_devsim.py through ds_plan_selfcheck)
  {
    const u256 i1024 = {{0x5DDCE6D4u, 0x9D01C8F4u, 0xDD1ADFEAu, 0x9AA7F950u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x4FBFFFFFu}};
    g.inv1024 = i1024;
  }
  return g;
}
// batch whose centre is m*S (m a residue mod n), if it lies in [0, limit)
static inline bool plan_batch_of(const ScanGeometry &g, const u256 &m, uint64_t limit, uint64_t &b) {
  using namespace plan_detail;
  return below(mulmod(submod(m, g.c), g.inv1024), limit, b);
}
static inline void plan_collapsed(const ScanGeometry &g, uint64_t n_batches, std::vector<uint64_t> &out) {
  using namespace plan_detail;
  out.clear();
  // b(m) = (m - c) / 1024 is an arithmetic progression in m: one multiplication, then additions
  for (int sign = 0; sign < 2; sign++) {
    const u256 step = sign ? negmod(g.inv1024) : g.inv1024;
    u256 b = mulmod(submod(sign ? negmod(from_u64(1)) : from_u64(1), g.c), g.inv1024);
    for (int e = 1; e <= 512; e++) {
      uint64_t v;
      if (below(b, n_batches, v)) out.push_back(v);
      b = addmod(b, step);
    }
  }
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
}

// Cuts batches [0, n_batches) into segments.  t_cap = the most walker threads a segment may use, t_align = their granularity
// (256 on the device).  Every segment is free of hop coincidences for its T, and every collapsed batch is a segment of its own.
static inline bool plan_scan(const u256 &k0, const u256 &s, uint64_t n_batches, uint64_t t_cap, uint64_t t_align, std::vector<ScanSegment> &out) {
  using namespace plan_detail;
  out.clear();
  const ScanGeometry g = plan_geometry(k0, s);
  if (!g.ok) return false;
  std::vector<uint64_t> col;
  plan_collapsed(g, n_batches, col);
  auto pick_T = [&](uint64_t batches, uint64_t below_T) {       // largest usable T for `batches`, strictly below `below_T` if given
    uint64_t need = ((batches + t_align - 1) / t_align) * t_align;
    uint64_t T = std::min(need, (t_cap / t_align) * t_align);
    if (T < t_align) T = t_align;
    if (below_T && T >= below_T) T = (below_T > t_align) ? below_T - t_align : T + t_align;   // no smaller T exists: take a larger one
    return T;
  };
  uint64_t b0 = 0;
  size_t ci = 0;
  uint64_t avoid_T = 0;
  while (b0 < n_batches) {
    while (ci < col.size() && col[ci] < b0) ci++;
    const uint64_t stop = (ci < col.size()) ? col[ci] : n_batches;   // next collapsed batch (or the end)
    if (stop == b0) {                                                  // a collapsed batch: alone, last step of its walker
      ScanSegment sg = {b0, b0 + 1, t_align, 1u};
      out.push_back(sg);
      b0++; ci++; avoid_T = 0;
      continue;
    }
    // the hop: smallest batch in [b0, stop) whose centre is +-W for this T
    uint64_t T = pick_T(stop - b0, avoid_T);
    uint64_t hop = stop;
    for (int tries = 0; tries < 8; tries++) {
      hop = stop;
      const u256 w = mulmod(from_u64(T), from_u64(1024));
      for (int sign = 0; sign < 2; sign++) {
        uint64_t b;
        if (plan_batch_of(g, sign ? negmod(w) : w, stop, b) && b >= b0 && b < hop) hop = b;
      }
      if (hop != b0) break;                                            // the first batch itself sits on +-W: another T
      T = (T > t_align) ? T - t_align : T + (uint64_t)(tries + 2) * t_align;
    }
    if (hop == b0) return false;                                       // cannot happen: at most two T values coincide per batch
    ScanSegment sg = {b0, hop, T, 0u};
    out.push_back(sg);
    avoid_T = (hop < stop) ? T : 0;                                    // the cut batch starts the next segment with a different W
    b0 = hop;
  }
  return true;
}

}  // namespace kh
