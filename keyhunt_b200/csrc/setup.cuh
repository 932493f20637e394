// setup.cuh — one-off set-up of a walk: the G-multiple table and the per-thread start centres.
//
// Replaces init_generator (keyhunt.cpp:5266: Gn[i] = (i+1)*stride*G, _2Gn), the GSn/_2GSn tables of
// BSGS (keyhunt.cpp:1803-1816: GSn[i] = -(i+1)*2m*G), and the per-batch / per-window start points
// (keyhunt.cpp:3349-3353, :4634-4642, :5314-5315).  Every table entry is one independent scalar multiplication.  The T start
// centres are an arithmetic progression of points, centre_t = centre_0 + t*D with D = 1024*S, and are set up in two levels: walker
// t = 64*i + j is A_i + B_j, where the row bases A_i = centre_{64 i} and the 63 offsets B_j = j*D are scalar multiplications
// (T/64 + 63 of them instead of T) and the 63 sums of a row are affine additions that share ONE field inversion (setup_row_fill).
// That is ~10 instead of ~630 field multiplications per walker: the set-up of 606,208 walkers takes 0.2 ms instead of 4 ms, which
// is what a short kh_bsgs_search call (a server request) mostly consisted of.
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
  if (p.inf) { fe_set_zero(cx); fe_set_zero(cy); }     // (0, 0) is not on the curve: setup_row_fill recognises a row base at infinity
  return p.inf == 0;
}


// ---- two-level set-up of the centres -----------------------------------------------------------------------------
#define KH_SETUP_ROW 64
// offset j (1..63) of a row: B_j = +-(1024*j*s)*G  ->  16 words (x limbs, y limbs)
KH_HD void setup_row_offset(uint32_t out[16], const WalkSetup &ws, uint32_t j) {
  u256 zero, k;
#pragma unroll
  for (int i = 0; i < 8; i++) zero.v[i] = 0;
  u256_add_mul64(k, zero, ws.s, (uint64_t)j * KH_GRP);
  ge p;
  ge_mul_g_comb(p, k, ws.comb);
  if (ws.neg) ge_neg(p, p);
#pragma unroll
  for (int i = 0; i < 8; i++) { out[i] = p.x.v[i]; out[8 + i] = p.y.v[i]; }
}
// Centres of the walkers 64*row + 1 .. 64*row + 63 (those below T) from the row base, which setup_center has already put into
// `centers`: 63 affine additions A + B_j behind one shared inversion.  A zero difference (A = +-B_j: only when the range touches
// the keys +-1024*j*s, or with a live walker at infinity) sends the whole row through setup_center instead, so every special case
// is decided by the code that decided it before; so does a row whose base is the point at infinity (stored as (0, 0)).  `pre`: 63 field elements of scratch.  Returns false if a LIVE walker of the row
// sits at infinity.
KH_HD bool setup_row_fill(uint32_t *centers, const uint32_t *offs, const WalkSetup &ws, uint64_t row, fe *pre) {
  const uint64_t t0 = row * KH_SETUP_ROW;
  if (t0 + 1 >= ws.T) return true;
  const int nj = (int)((ws.T - t0 - 1 < KH_SETUP_ROW - 1) ? (ws.T - t0 - 1) : (KH_SETUP_ROW - 1));
  fe ax, ay;
#pragma unroll
  for (int l = 0; l < 8; l++) { ax.v[l] = centers[(uint64_t)l * ws.T + t0]; ay.v[l] = centers[(uint64_t)(8 + l) * ws.T + t0]; }
  fe acc;
  fe_set_zero(acc);
#pragma unroll 1
  for (int j = 1; j <= nj; j++) {
    fe bx, dx;
#pragma unroll
    for (int l = 0; l < 8; l++) bx.v[l] = offs[16 * (j - 1) + l];
    fe_sub(dx, bx, ax);
    if (j == 1) acc = dx; else fe_mul_cold(acc, acc, dx);
    pre[j - 1] = acc;
  }
  bool ok = true;
  if (fe_is_zero(acc) || (fe_is_zero(ax) && fe_is_zero(ay))) {
#pragma unroll 1
    for (int j = 1; j <= nj; j++) {
      const uint64_t t = t0 + (uint64_t)j;
      fe cx, cy;
      if (!setup_center(cx, cy, ws, t) && (ws.n_batches == 0 || ws.first_batch + t < ws.n_batches)) ok = false;
#pragma unroll
      for (int l = 0; l < 8; l++) { centers[(uint64_t)l * ws.T + t] = cx.v[l]; centers[(uint64_t)(8 + l) * ws.T + t] = cy.v[l]; }
    }
    return ok;
  }
  fe inv;
  fe_inv(inv, acc);
#pragma unroll 1
  for (int j = nj; j >= 1; j--) {
    fe bx, by, dinv, dy, sl, x3, y3;
#pragma unroll
    for (int l = 0; l < 8; l++) { bx.v[l] = offs[16 * (j - 1) + l]; by.v[l] = offs[16 * (j - 1) + 8 + l]; }
    if (j > 1) {
      fe dx;
      fe_mul_cold(dinv, pre[j - 2], inv);
      fe_sub(dx, bx, ax);
      fe_mul_cold(inv, inv, dx);
    } else {
      dinv = inv;
    }
    fe_sub(dy, by, ay);
    fe_mul_cold(sl, dy, dinv);
    fe_mul_cold(x3, sl, sl);
    fe_sub(x3, x3, ax);
    fe_sub(x3, x3, bx);
    fe_sub(y3, ax, x3);
    fe_mul_cold(y3, y3, sl);
    fe_sub(y3, y3, ay);
    const uint64_t t = t0 + (uint64_t)j;
#pragma unroll
    for (int l = 0; l < 8; l++) { centers[(uint64_t)l * ws.T + t] = x3.v[l]; centers[(uint64_t)(8 + l) * ws.T + t] = y3.v[l]; }
  }
  return ok;
}

}  // namespace kh
