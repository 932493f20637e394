// setup.cuh — one-off set-up of a walk: the G-multiple table and the per-thread start centres.
//
// Replaces init_generator (keyhunt.cpp:5266: Gn[i] = (i+1)*stride*G, _2Gn), the GSn/_2GSn tables of
// BSGS (keyhunt.cpp:1803-1816: GSn[i] = -(i+1)*2m*G), and the per-batch / per-window start points
// (keyhunt.cpp:3349-3353, :4634-4642, :5314-5315).  Every table entry and every centre is one
// independent scalar multiplication, so set-up is a tiny data-parallel kernel.
//
// A walk is described by
//   s    : the step scalar (stride for scans, 2m for giant steps, 1 for baby steps)
//   neg  : S = -(s*G) instead of s*G   (giant steps walk downwards)
//   k0   : scalar of point index 0     (range start ; base+m ; 1)
//   Q    : optional base point (BSGS target public key), else the walk starts from infinity
// point(index p) = Q + sgn*(k0 + p*s)*G ; batch b covers p in [1024b, 1024b+1024) and its centre is
// p = 1024b + 512 ; thread t starts on batch `first_batch + t` and hops W = T*1024*S per step.
#pragma once
#include "walk.cuh"

namespace kh {

struct WalkSetup {
  u256 s;            // step scalar
  u256 k0;           // scalar of point index 0
  ge q;              // base point (q.inf != 0: none)
  uint32_t neg;      // 1: negate s*G
  uint32_t pad;
  uint64_t T;        // walker threads
  uint64_t first_batch;
  uint64_t n_batches; // walkers whose first batch is >= n_batches are idle (0 = every walker is live)
  const uint32_t *comb; // fixed-base comb of G (ec.cuh ge_mul_g_comb), nullptr = plain double-and-add
};

// table entry e (0 = W = T*1024*S ; e >= 1 : e*S)  ->  16 words (x limbs, y limbs)
KH_HD void setup_table_entry(uint32_t out[16], const WalkSetup &ws, uint32_t e) {
  u256 zero, k;
#pragma unroll
  for (int i = 0; i < 8; i++) zero.v[i] = 0;
  const uint64_t mult = (e == 0) ? ws.T * (uint64_t)KH_GRP : (uint64_t)e;
  u256_add_mul64(k, zero, ws.s, mult);
  ge p;
  ge_mul_g_comb(p, k, ws.comb);
  if (ws.neg) ge_neg(p, p);
#pragma unroll
  for (int i = 0; i < 8; i++) { out[i] = p.x.v[i]; out[8 + i] = p.y.v[i]; }
}

// centre of walker thread t: Q +- (k0 + (1024*(first_batch+t) + 512)*s)*G ; returns false if it is
// the point at infinity (cannot happen for keys in [1, n-1] away from the range ends)
KH_HD bool setup_center(fe &cx, fe &cy, const WalkSetup &ws, uint64_t t) {
  u256 k;
  u256_add_mul64(k, ws.k0, ws.s, (ws.first_batch + t) * (uint64_t)KH_GRP + (uint64_t)KH_HALF);
  ge p;
  ge_mul_g_comb(p, k, ws.comb);
  if (ws.neg) ge_neg(p, p);
  if (!ws.q.inf) { ge r; ge_add(r, ws.q, p); p = r; }
  cx = p.x; cy = p.y;
  return p.inf == 0;
}

}  // namespace kh
