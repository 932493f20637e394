// walk.cuh — the per-thread batch walk: the device form of the inner loops of thread_process
// (keyhunt.cpp:3348-3461, :3840-3855), thread_process_bsgs (:4644-4716, :4863-4877) and
// thread_bPload (:5317-5458).
//
// Geometry (identical to the reference's CPU_GRP_SIZE = 1024 group): a batch has a
// centre C; its 1024 points are pts[512] = C, pts[512 +- i] = C +- i*S (i = 1..511) and
// pts[0] = C - 512*S, where S is the "stride point" (stride*G for scans, -2m*G for BSGS giant steps).
// The 512 differences Gx[i] - Cx are inverted together with ONE field inversion (Montgomery's trick,
// IntGroup::ModInv IntGroup.cpp:36-57); +i and -i share their inverse.
//
// What differs from the reference (results are identical):
//   * The reference recomputes every batch centre with a full scalar multiplication
//     (keyhunt.cpp:3352).  Here thread t owns batches t, t+T, t+2T, ... and moves its centre with one
//     more affine addition of W = T*1024*S whose inverse rides in the same batched inversion (table
//     entry 0), so a batch costs no scalar multiplication at all.
//   * The 513 prefix products live in a per-thread column of a global scratch array laid out
//     [entry][thread] in 16-byte words, so every warp access is one fully coalesced 512-byte line.
//   * The G-multiple table (513 x 64 B) is staged in shared memory once per CTA; all threads of a
//     warp read the same entry at the same time (a broadcast, conflict-free).
#pragma once
#include <stdint.h>

#include "ec.cuh"

namespace kh {

#define KH_GRP 1024
#define KH_HALF 512
#define KH_TAB_ENTRIES 513          // entry 0 = W, entries 1..512 = 1*S .. 512*S
#define KH_TAB_WORDS (KH_TAB_ENTRIES * 16)

struct kh_u4 {  // 16-byte word (uint4 on the device)
  uint32_t x, y, z, w;
};

struct WalkParams {
  const uint32_t *gtab;   // global copy of the table: entry e -> 16 words: x limbs 0..7, y limbs 8..15
  uint32_t *centers;      // [16][T]: limb l of thread t at centers[l*T + t] (x limbs 0..7, y limbs 8..15)
  kh_u4 *scratch;         // [1024][T]: prefix product of entry e, half h at scratch[(2e+h)*T + t]
  uint64_t T;             // walker threads
  uint64_t batch_base;    // batch handled by thread 0 in step 0 of this launch
  uint64_t n_batches;     // batches in the whole range; batch b is processed iff b < n_batches
  uint32_t steps;         // steps of this launch
  uint32_t pad;
};

KH_HD void tab_load(fe &gx, fe &gy, const uint32_t *tab, int e) {
  const uint32_t *p = tab + 16 * e;
#if defined(__CUDA_ARCH__)
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = q[0], b = q[1], c = q[2], d = q[3];
  gx.v[0] = a.x; gx.v[1] = a.y; gx.v[2] = a.z; gx.v[3] = a.w; gx.v[4] = b.x; gx.v[5] = b.y; gx.v[6] = b.z; gx.v[7] = b.w;
  gy.v[0] = c.x; gy.v[1] = c.y; gy.v[2] = c.z; gy.v[3] = c.w; gy.v[4] = d.x; gy.v[5] = d.y; gy.v[6] = d.z; gy.v[7] = d.w;
#else
  for (int i = 0; i < 8; i++) { gx.v[i] = p[i]; gy.v[i] = p[8 + i]; }
#endif
}
KH_HD void tab_load_x(fe &gx, const uint32_t *tab, int e) {
  const uint32_t *p = tab + 16 * e;
#if defined(__CUDA_ARCH__)
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = q[0], b = q[1];
  gx.v[0] = a.x; gx.v[1] = a.y; gx.v[2] = a.z; gx.v[3] = a.w; gx.v[4] = b.x; gx.v[5] = b.y; gx.v[6] = b.z; gx.v[7] = b.w;
#else
  for (int i = 0; i < 8; i++) gx.v[i] = p[i];
#endif
}
KH_HD void scratch_store(kh_u4 *s, uint64_t T, uint64_t t, int e, const fe &a) {
  kh_u4 lo = {a.v[0], a.v[1], a.v[2], a.v[3]}, hi = {a.v[4], a.v[5], a.v[6], a.v[7]};
#if defined(__CUDA_ARCH__)
  reinterpret_cast<uint4 *>(s)[(uint64_t)(2 * e) * T + t] = make_uint4(lo.x, lo.y, lo.z, lo.w);
  reinterpret_cast<uint4 *>(s)[(uint64_t)(2 * e + 1) * T + t] = make_uint4(hi.x, hi.y, hi.z, hi.w);
#else
  s[(uint64_t)(2 * e) * T + t] = lo;
  s[(uint64_t)(2 * e + 1) * T + t] = hi;
#endif
}
KH_HD void scratch_load(fe &a, const kh_u4 *s, uint64_t T, uint64_t t, int e) {
#if defined(__CUDA_ARCH__)
  uint4 lo = reinterpret_cast<const uint4 *>(s)[(uint64_t)(2 * e) * T + t];
  uint4 hi = reinterpret_cast<const uint4 *>(s)[(uint64_t)(2 * e + 1) * T + t];
#else
  kh_u4 lo = s[(uint64_t)(2 * e) * T + t], hi = s[(uint64_t)(2 * e + 1) * T + t];
#endif
  a.v[0] = lo.x; a.v[1] = lo.y; a.v[2] = lo.z; a.v[3] = lo.w; a.v[4] = hi.x; a.v[5] = hi.y; a.v[6] = hi.z; a.v[7] = hi.w;
}

// Walks `steps` batches for walker thread t.  `tab` is the table in shared memory (device) or plain
// memory (host test build).  For each point calls emit.point(x, y, batch, idx) where key index in the
// range is batch*1024 + idx; y is valid only if Emit::NEED_Y.
template <class Emit>
KH_HD void walk_batches(const WalkParams &wp, const uint32_t *tab, uint64_t t, Emit &emit) {
  constexpr bool OL = Emit::OUTLINE_MUL;   // one shared multiplier copy in instruction-fetch-bound kernels
  constexpr int RR = Emit::RARE_REDUCE;    // form of the multiplier's final reduction (fe.cuh KH_RARE_REDUCE), one per kernel
  fe px, py;
#pragma unroll
  for (int l = 0; l < 8; l++) { px.v[l] = wp.centers[(uint64_t)l * wp.T + t]; py.v[l] = wp.centers[(uint64_t)(8 + l) * wp.T + t]; }

#pragma unroll 1
  for (uint32_t step = 0; step < wp.steps; step++) {
    const uint64_t batch = wp.batch_base + (uint64_t)step * wp.T + t;
    if (batch >= wp.n_batches) break;

    // ---- forward pass: prefix products of dx_e = tab[e].x - px ------------------------------------
    fe acc;
#pragma unroll 1
    for (int e = 0; e < KH_TAB_ENTRIES; e++) {
      fe gx, dx;
      tab_load_x(gx, tab, e);
      fe_sub(dx, gx, px);
      if (e == 0) acc = dx; else fe_mul_sel<OL, RR>(acc, acc, dx);
      if (e < KH_TAB_ENTRIES - 1) scratch_store(wp.scratch, wp.T, t, e, acc);
    }
    fe inv;
    fe_inv<RR, Emit::INV_SQR>(inv, acc);   // one inversion per 1024 points (+ the centre move)

    // ---- backward pass: peel the inverses off and produce the points ------------------------------
#pragma unroll 1
    for (int e = KH_TAB_ENTRIES - 1; e >= 0; e--) {
      fe gx, gy, dinv;
      tab_load(gx, gy, tab, e);
      if (e > 0) {
        fe pre, dx;
        scratch_load(pre, wp.scratch, wp.T, t, e - 1);
        fe_mul_sel<OL, RR>(dinv, pre, inv);       // 1/dx_e
        fe_sub(dx, gx, px);
        fe_mul_sel<OL, RR>(inv, inv, dx);         // 1/(dx_0 ... dx_{e-1})
      } else {
        dinv = inv;
      }
      if (Emit::PAIRS && e != 0 && e != KH_HALF) {
        // x-only emitters take C+e*S and C-e*S together: two independent multiply chains (ILP) and, in the
        // emitter, the bloom probes of both points in flight at the same time (memory-level parallelism)
        fe dyp, dym, sp, sm, xp, xm, c;
        fe_sub(dyp, gy, py);
        fe_add(dym, gy, py);
        fe_mul_sel<OL, RR>(sp, dyp, dinv);
        fe_mul_sel<OL, RR>(sm, dym, dinv);
        if (OL) { fe_sqr_sel<Emit::INV_SQR, RR>(xp, sp); fe_sqr_sel<Emit::INV_SQR, RR>(xm, sm); } else { fe_sqr<RR>(xp, sp); fe_sqr<RR>(xm, sm); }
        fe_add(c, px, gx);
        fe_sub(xp, xp, c);
        fe_sub(xm, xm, c);
        emit.pair(xp, (uint32_t)(KH_HALF + e), xm, (uint32_t)(KH_HALF - e), batch);
        continue;
      }
#pragma unroll 1
      for (int sgn = 0; sgn < 2; sgn++) {
        if (e == KH_HALF && sgn == 0) continue;        // +512*S belongs to the next batch (pts[0] there)
        fe x3, y3;
        uint32_t idx;
        bool do_emit = true;
        if (e == 0 && sgn == 0) {                       // the centre itself
          x3 = px; y3 = py; idx = KH_HALF;
        } else {
          fe s, dy, s2;
          if (sgn == 0 || e == 0) fe_sub(dy, gy, py);   // C + e*S  (and the centre move C + W)
          else fe_add(dy, gy, py);                      // C - e*S : slope is -(gy+py)/dx, its sign is irrelevant for x
          fe_mul_sel<OL, RR>(s, dy, dinv);
          if (OL) fe_sqr_sel<Emit::INV_SQR, RR>(s2, s); else fe_sqr<RR>(s2, s);   // dedicated squaring where the FMA-heavy pipe is the bound
          fe_sub(x3, s2, px);
          fe_sub(x3, x3, gx);
          if (e == 0) {                                 // new centre: always needs y
            fe_sub(y3, gx, x3); fe_mul_sel<OL, RR>(y3, y3, s); fe_sub(y3, y3, gy);
            px = x3; py = y3;
            do_emit = false;
            idx = 0;
          } else {
            if (Emit::NEED_Y) {
              if (sgn == 0) { fe_sub(y3, gx, x3); fe_mul_sel<OL, RR>(y3, y3, s); fe_sub(y3, y3, gy); }   // s*(gx-x3) - gy
              else          { fe_sub(y3, x3, gx); fe_mul_sel<OL, RR>(y3, y3, s); fe_add(y3, y3, gy); }   // s'*(x3-gx) + gy, s' = -s
            } else {
              y3 = py;
            }
            idx = (sgn == 0) ? (uint32_t)(KH_HALF + e) : (uint32_t)(KH_HALF - e);
          }
        }
        if (do_emit) emit.point(x3, y3, batch, idx);
      }
    }
  }
#pragma unroll
  for (int l = 0; l < 8; l++) { wp.centers[(uint64_t)l * wp.T + t] = px.v[l]; wp.centers[(uint64_t)(8 + l) * wp.T + t] = py.v[l]; }
}

}  // namespace kh
