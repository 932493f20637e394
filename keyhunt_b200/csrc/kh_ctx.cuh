// kh_ctx.cuh — host-side context shared by the scan and BSGS translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/keyhunt_b200.h"
#include "emit.cuh"
#include "setup.cuh"

// walker-thread counts are multiples of this, whatever CTA shape a kernel uses (128 or 256 threads)
#define KH_T_ALIGN 256

struct kh_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  cudaDeviceProp prop;

  // options
  int threads_per_sm = 4096;       // walker threads per SM: 8 waves of 2 resident CTAs; measured +4..6 % over one wave (less launch tail, phases decorrelate)
  int steps_per_launch = 16;
  uint32_t hit_capacity = 1u << 16;
  int endomorphism = 0;            // -e: test beta*x and beta^2*x of every point too
  int prefilter = 1;               // exact prefix bitmap in front of the bloom for target sets of <= 65,536 records
  int bsgs_prefilter = 1;          // build the baby-point prefix bitmap when HBM allows (kh_bsgs_build)
  int bsgs_base_check = 0;         // server variant of the BSGS search (bsgsd.cpp:2544)
  int bsgs_binned_build = 1;       // kh_bsgs_build bins the bloom / bitmap updates by X prefix so that they hit L2 (emit.cuh BabyBins)

  // walk state
  uint64_t T_alloc = 0;            // walker threads the buffers are sized for
  uint32_t *d_gtab = nullptr;      // KH_TAB_WORDS
  uint32_t *d_rowoffs = nullptr;   // 63 x 16 words: the row offsets of the two-level centre set-up (setup.cuh)
  uint32_t *d_centers = nullptr;   // 16*T
  kh::kh_u4 *d_scratch = nullptr;  // 1024*T
  uint32_t *d_flags = nullptr;     // [0] = set-up error flag (centre at infinity)
  uint32_t *d_sha2 = nullptr;      // 256 x 68 words: SHA-256 schedules of the uncompressed key's second block (hash.cuh)
  uint32_t *d_comb = nullptr;      // fixed-base comb of G: 32 x 256 points (ec.cuh ge_mul_g_comb), built on first use

  // scan targets
  bool have_targets = false;
  int mode = 0, crypto = 0, search = 0, scan_kind = 0;
  kh_bloom_desc bloom_desc{};
  uint8_t *d_bloom = nullptr;
  uint32_t *d_table = nullptr;     // N x 5 BE words
  size_t table_alloc = 0, bloom_alloc = 0;   // bytes behind d_table / d_bloom
  uint64_t n_targets = 0;
  std::vector<uint8_t> h_table20;  // sorted records (host copy for kh_get_table)
  uint32_t *d_pre = nullptr;       // exact prefix bitmap in front of the bloom (ScanTargets::pre), 2^pre_k bits
  uint32_t pre_k = 0;
  uint32_t *d_vanity = nullptr;    // -m vanity: prefix bitmap + interval limits (ScanTargets::van)
  uint32_t n_vanity = 0;

  // hits
  kh::RawHit *d_hits = nullptr;
  uint32_t *d_hit_count = nullptr;
  uint32_t hits_alloc = 0;
  struct PendingScan { uint8_t start[32]; uint8_t stride[32]; };
  std::vector<kh_hit> ready_hits;  // converted, not yet polled
  bool overflowed = false;

  // bsgs
  bool have_bsgs = false;
  kh_bsgs_desc bsgs{};
  uint8_t *d_tier[3] = {nullptr, nullptr, nullptr};
  uint64_t tier_stride[3] = {0, 0, 0};
  kh::BpEntry *d_bptable = nullptr;
  uint32_t *d_aux_tab = nullptr;   // AMP2/AMP3 + helper points for refinement
  uint32_t *d_bsgs_pre = nullptr;  // exact prefix bitmap over the baby points' X (BsgsTables::pre), 2^bsgs_pre_k bits
  uint32_t bsgs_pre_k = 0;
  void *d_giant_cands = nullptr;   // tier-1 positives of one launch (kh::GiantCand[65536])
  uint32_t *d_giant_cnt = nullptr; // [0] candidate count, [1] found flag
  void *d_giant_key = nullptr;     // found key (kh::u256)

  kh_stats stats{};
};

static inline int kh_fail(kh_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

#define KH_CUDA(c, call)                                                                             \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return kh_fail((c), (e_ == cudaErrorMemoryAllocation) ? KH_ENOMEM : KH_ENODEV, "%s: %s (%s:%d)", #call, \
                     cudaGetErrorString(e_), __FILE__, __LINE__);                                    \
  } while (0)

// implemented in kh_scan.cu, used by kh_bsgs.cu
int kh_ensure_walk_buffers(kh_ctx *c, uint64_t T);
uint64_t kh_pick_T(kh_ctx *c, uint64_t n_batches);
int kh_run_setup(kh_ctx *c, const kh::WalkSetup &ws);   // (fills in ws.comb itself)
int kh_ensure_comb(kh_ctx *c);
void kh_time_begin(kh_ctx *c);
double kh_time_end(kh_ctx *c);  // ms since kh_time_begin on the context stream (synchronises)
