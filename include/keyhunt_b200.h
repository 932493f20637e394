/* keyhunt_b200.h — C ABI of libkh_b200.so: the B200-native replacement for keyhunt's per-thread
 * key-range search workers.
 *
 * The reference (naanprofit/keyhunt) has no FFI boundary: its workers are pthread entry points that
 * read process globals.  This header is the boundary a maintainer binds instead of spawning those
 * threads (see INTEGRATION.md for the exact patch against keyhunt.cpp).  Each entry point names the
 * reference code it replaces (file:line relative to the reference root).
 *
 * Conventions: extern "C", opaque handle, plain pointers + sizes, caller owns every host buffer, the
 * library owns device memory.  Return 0 on success, a negative KH_E* code on failure (message via
 * kh_last_error).  One host thread per context (= per GPU); calls on one context are not re-entrant.
 * 256-bit integers cross the boundary as 32-byte BIG-ENDIAN strings (Int::Get32Bytes, Int.cpp:308).
 * There is no CPU fallback: kh_create fails when no CUDA device is usable.
 */
#ifndef KEYHUNT_B200_H
#define KEYHUNT_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KH_OK 0
#define KH_ENODEV (-1)   /* no usable CUDA device / CUDA runtime error */
#define KH_EINVAL (-2)   /* bad argument (range not a multiple of 1024, n not 2^even, ...) */
#define KH_ENOMEM (-3)   /* device or host allocation failed */
#define KH_ESTATE (-4)   /* call out of order (scan before set_targets, search before build, ...) */
#define KH_EOVERFLOW (-5)/* more hits than the hit buffer holds (hits were dropped) */

typedef struct kh_ctx kh_ctx;

/* FLAGMODE / FLAGCRYPTO / FLAGSEARCH values of the reference (keyhunt.cpp:55-68) */
enum { KH_MODE_XPOINT = 0, KH_MODE_ADDRESS = 1, KH_MODE_BSGS = 2, KH_MODE_RMD160 = 3, KH_MODE_VANITY = 6 };
enum { KH_CRYPTO_BTC = 1, KH_CRYPTO_ETH = 2 };
enum { KH_SEARCH_UNCOMPRESS = 0, KH_SEARCH_COMPRESS = 1, KH_SEARCH_BOTH = 2 };
/* which derived value matched */
enum { KH_HIT_COMP02 = 0, KH_HIT_COMP03 = 1, KH_HIT_UNCOMP = 2, KH_HIT_ETH = 3, KH_HIT_XPOINT = 4 };

/* context: cudaSetDevice, stream, hit buffer.  (replaces nothing: the reference is one process) */
int kh_create(kh_ctx **out, int device_ordinal);
void kh_destroy(kh_ctx *ctx);
const char *kh_last_error(kh_ctx *ctx);
/* options: "threads_per_sm" (walker threads per SM), "steps_per_launch", "hit_capacity",
 * "endomorphism" (1 = the reference's -e: also test beta*x and beta^2*x of every point, keyhunt.cpp:3408-3473),
 * "prefilter" (default 1; 0 = no prefix bitmap in front of the bloom: kh_set_targets normally also builds an exact bitmap over
 *  the first 16..32 bits of every target record, which lets whole warps skip the two XXH64 + bloom probes; hits are identical
 *  either way, the bloom image and table returned by kh_get_bloom / kh_get_table are unaffected),
 * "bsgs_prefilter" (default 1: kh_bsgs_build also fills an exact bitmap over the first bits of every baby point's X, up to
 *  3/4 of the free HBM, which answers most tier-1 probes with one memory access; found keys are identical),
 * "bsgs_binned_build" (default 1: kh_bsgs_build bins the bloom / bitmap updates of tables with >= 2^26 baby steps by the
 *  first bits of X so that they hit L2 instead of 21 random HBM read-modify-writes per baby step; 2 = always, 0 = never; the
 *  tables are byte-identical either way),
 * "bsgs_base_check" (1 = kh_bsgs_search behaves like the reference SERVER's loop, which also reports a key equal to
 *  the base key of a 2N window, bsgsd.cpp:2544; 0 = keyhunt.cpp's thread_process_bsgs, the default) */
int kh_set_option(kh_ctx *ctx, const char *name, int64_t value);

/* bloom_init2 sizing (bloom/bloom.cpp:154-187) with error = 0.000001 (keyhunt.cpp:7620) */
typedef struct {
  uint64_t entries, bits, bytes;
  uint32_t hashes;
  uint32_t pad;
} kh_bloom_desc;
int kh_bloom_params(uint64_t entries, kh_bloom_desc *out);

/* ---- scan modes: replaces thread_process (keyhunt.cpp:3265-3861) --------------------------------- */

/* Target set = what readFileAddress leaves in `bloom` + `addressTable` (keyhunt.cpp:7033-7470,
 * _sort :4307): N raw 20-byte records in any order.  The library sorts them, uploads them and sets
 * the bloom bits on the device (bit-identical to bloom_add per record).  If bloom_bits != NULL it is
 * used as the bloom image instead (desc must describe it).  entries = N<=10000 ? 10000 : N unless
 * desc overrides it (keyhunt.cpp:7608). */
int kh_set_targets(kh_ctx *ctx, int mode, int crypto, int search, const uint8_t *records20, uint64_t n_records,
                   const kh_bloom_desc *desc, const uint8_t *bloom_bits);
/* -m vanity: replaces processOneVanity / readFileVanity (keyhunt.cpp:6971-7030) + vanityrmdmatch (:6677) inside
 * thread_process_vanity (:3867).  The caller passes what addvanity (:6739) produced: n_pairs interval limits
 * vanity_rmd_limit_values_A/B[i][j], flattened, 20 bytes each; a hash160 h matches when A <= h <= B (memcmp order) for
 * some pair.  Candidates per point follow `search` exactly like -m rmd160 (both compressed prefixes / uncompressed),
 * hits come back through kh_poll_hits with the same n-k fix-up, and "endomorphism" applies.  BTC only. */
int kh_set_vanity(kh_ctx *ctx, int search, const uint8_t *limits_a20, const uint8_t *limits_b20, uint64_t n_pairs);
/* read back the device-resident bloom image / sorted table (parity checks, -S style persistence) */
int kh_get_bloom(kh_ctx *ctx, kh_bloom_desc *desc, uint8_t *dst, uint64_t cap_bytes);
int kh_get_table(kh_ctx *ctx, uint8_t *dst20, uint64_t cap_records, uint64_t *n_records);

/* Scans the keys start + i*stride, i in [0, n_points); n_points must be a multiple of 1024 (one
 * reference batch).  This is what a worker does for its claimed chunk (keyhunt.cpp:3321-3324,
 * :3348-3856) — the caller keeps the range cursor and the chunk overshoot rule.
 * Blocking; hits accumulate in the context until polled. */
int kh_scan(kh_ctx *ctx, const uint8_t start_be[32], const uint8_t stride_be[32], uint64_t n_points);

typedef struct {
  uint8_t key_be[32];    /* reported private key, after the n-k fix-up of keyhunt.cpp:3629-3635 */
  uint8_t pub_x[32];     /* public key of key_be */
  uint8_t pub_y[32];
  uint8_t matched[20];   /* the 20 bytes found in the table */
  uint8_t kind;          /* KH_HIT_* */
  uint8_t pad[3];        /* pad[0]: with "endomorphism" the reference's candidate index l (0..11 BTC, 0..5 ETH, 0..2 xpoint) */
  uint64_t index;        /* point index inside the scanned range */
} kh_hit;
/* drains the hits of the scans since the last poll, ascending (index, kind); *n = number written.
 * Returns KH_EOVERFLOW if the device hit buffer overflowed. */
int kh_poll_hits(kh_ctx *ctx, kh_hit *out, int max, int *n);

/* per-key derivation on the device (hit fix-up / writekey: keyhunt.cpp:3629, :6891, :6925):
 * public key, both hash160 forms, ETH address */
typedef struct {
  uint8_t pub_x[32], pub_y[32];
  uint8_t h160_comp[20], h160_uncomp[20], eth[20];
  uint8_t pad[4];
} kh_keyinfo;
int kh_derive(kh_ctx *ctx, const uint8_t *keys_be, uint64_t n_keys, kh_keyinfo *out);

/* ---- BSGS: replaces thread_bPload (keyhunt.cpp:5284), bsgs_sort (:4412), thread_process_bsgs (:4549),
 *      bsgs_secondcheck (:5151), bsgs_thirdcheck (:5186), bsgs_searchbinary (:4510) -------------------- */
typedef struct {
  uint64_t n, m, m2, m3, aux;          /* BSGS_N (rounded), bsgs_m, bsgs_m2, bsgs_m3, bsgs_aux */
  kh_bloom_desc tier[3];               /* per-shard descriptors of bloom_bP / bloom_bPx2nd / bloom_bPx3rd */
} kh_bsgs_desc;
/* n = -n value (2^even >= 2^20), k = -k factor.  Builds the 3 x 256 bloom shards and the sorted bP
 * table on the device; everything stays resident in HBM. */
int kh_bsgs_build(kh_ctx *ctx, uint64_t n, uint32_t k);
int kh_bsgs_describe(kh_ctx *ctx, kh_bsgs_desc *out);
/* tier 1..3: bf bytes of one shard; tier 0: the bP table (m3 x 16-byte struct bsgs_xvalue, shard ignored) */
int kh_bsgs_export(kh_ctx *ctx, int tier, int shard, void *dst, uint64_t cap_bytes);
int kh_bsgs_import(kh_ctx *ctx, int tier, int shard, const void *src, uint64_t len_bytes);
/* order-independent 64-bit digest, computed on the device, of the bP table (tier 0), of all 256 shards of bloom tier 1..3, or of
 * the baby-point prefix bitmap (tier 4; "bsgs_prefilter").  Two builds hold the same bytes iff (up to 2^-64) the digests agree: the
 * way to compare tables that are too big to export (7.7 GB tier 1 and a 64 GB bitmap at -k 512). */
int kh_bsgs_digest(kh_ctx *ctx, int tier, uint64_t *out);
/* sequential search (-B sequential) of [start, end) for one public key, in windows of 2n keys like
 * thread_process_bsgs; *found = 1 and the key when found. */
int kh_bsgs_search(kh_ctx *ctx, const uint8_t pub_xy_be[64], const uint8_t start_be[32], const uint8_t end_be[32],
                   uint8_t found_key_be[32], int *found);

/* ---- measurement -------------------------------------------------------------------------------- */
typedef struct {
  double walk_ms;          /* device time of the walk kernels (CUDA events on the context stream) */
  double setup_ms;         /* device time of table/centre set-up kernels */
  double aux_ms;           /* BSGS refine / sort / bloom-build kernels */
  uint64_t walk_launches;  /* kernels launched */
  uint64_t other_launches;
  uint64_t points;         /* points (scan) or giant steps (bsgs) or baby steps (build) walked */
  uint64_t walker_threads; /* T of the last walk */
  uint64_t tier1_positives;/* BSGS: tier-1 bloom positives sent to refinement */
  uint64_t collapsed_batches; /* 1024-point batches whose shared inverse did not exist (centre = +-e*stride*G, i.e. the range
                                 touches key 0 mod n); the reference's IntGroup::ModInv yields garbage for such a batch
                                 (SURVEY App. B.11), here it reports nothing and is counted */
} kh_stats;
int kh_get_stats(kh_ctx *ctx, kh_stats *out, int reset);
int kh_device_info(kh_ctx *ctx, char *name, int name_cap, int *sm_count, uint64_t *hbm_bytes);
/* integer-pipe micro-benchmarks (thread-level ops/s over the whole chip), the measured denominators of
 * the integer roofline: [0] IADD3, [1] LOP3, [2] SHF, [3] IMAD, [4] IMAD.WIDE.U32.X (fe_mul row),
 * [5] LOP3+IMAD issued together (both integer pipes) */
int kh_int_peak(kh_ctx *ctx, double out_ops_per_s[6]);
/* more pipe micro-benchmarks (thread-level ops/s over the whole chip), the evidence behind the choice of multiplier:
 * [0] IMAD.WIDE.U32 without carry chain, [1] IMAD.HI.U32, [2] DFMA (FP64 pipe), [3] DADD, [4] DFMA + IMAD.WIDE issued
 * together (do the FP64 and FMA-heavy pipes overlap?), [5] IMAD.WIDE.U32.X + IADD3 together (multiplier + carry work),
 * [6] FFMA, [7] IMAD.WIDE.U32.X per second when issued in the x-only walk's own mix (16 wide multiply-adds in carry chains per
 * 32 ALU-pipe ops, the ratio ncu measures in kh_scan_kernel<XPOINT>): the practical ceiling of the EC-bound kernels,
 * [8] IMAD.WIDE.U32 without carry + LOP3 issued together (total ops/s), [9..15] reserved (0) */
int kh_pipe_peak(kh_ctx *ctx, double out_ops_per_s[16]);
/* hash micro-benchmarks in isolation: [0] SHA-256 compressions/s, [1] RIPEMD-160 blocks/s, at blocks_per_sm CTAs of 256 */
int kh_hash_peak(kh_ctx *ctx, int blocks_per_sm, double out_blocks_per_s[2]);

/* ---- device-side known-answer test of the field layer ---------------------------------------------
 * Runs one field operation per element ON THE GPU with the very functions the kernels use (the PTX carry-chain bodies of
 * fe.cuh) — parity tests feed it the reference-generated vectors of Int::ModMulK1 (IntMod.cpp:855), ModSquareK1 (:977),
 * ModInv (:382), ModAdd (:41), ModSub (:72), ModNeg (:102) and forced edge operands.  a_be / b_be / out_be: n x 32-byte
 * big-endian values (b is ignored by unary ops; operands of the modular ops must be < P like everywhere on the path).
 * KH_FE_REDUCE_WIDE reduces the 512-bit value a*2^256 + b, any a and b. */
enum { KH_FE_MUL = 0, KH_FE_SQR = 1, KH_FE_INV = 2, KH_FE_ADD = 3, KH_FE_SUB = 4, KH_FE_NEG = 5, KH_FE_MUL_OUTLINE = 6,
       KH_FE_MULWIDE_LO = 7, KH_FE_MULWIDE_HI = 8, KH_FE_SQRWIDE_LO = 9, KH_FE_SQRWIDE_HI = 10, KH_FE_REDUCE_WIDE = 11,
       /* the same operations with the OTHER form of the multiplier's final reduction: the kernels contain both (the C2 kernel keeps
        * the straight-line conditional subtraction, the others test "is it >= P at all" first and branch) */
       KH_FE_MUL_ALT = 12, KH_FE_SQR_ALT = 13, KH_FE_INV_ALT = 14, KH_FE_MUL_OUTLINE_ALT = 15, KH_FE_REDUCE_WIDE_ALT = 16,
       KH_FE_INV_SQR = 17 /* the inversion as the x-only walks run it: its 255 squarings through the dedicated squaring */ };
int kh_selftest_fe(kh_ctx *ctx, int op, const uint8_t *a_be, const uint8_t *b_be, uint64_t n, uint8_t *out_be);

#ifdef __cplusplus
}
#endif
#endif
