/* oracle/kh_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, 64-bit limbs + unsigned __int128) of the reference's bounded key-range
 * search hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this; the product (keyhunt_b200/) never links, imports or executes anything under oracle/.
 *
 * PARITY: pinned.  Every function below is checked (tests/test_oracle_vs_ref.py, run in the build
 * container where /root/reference exists) against the reference's own object code through
 * oracle/_ref/libkh_ref.so (same signatures, prefix khr_), against black-box runs of the
 * unmodified reference binary oracle/_ref/keyhunt on the reference's fixture files, and against
 * the committed golden vectors under tests/golden/ (generated FROM the reference by
 * tests/golden/make_golden.py).
 *
 * All multi-byte integers cross this API as 32-byte BIG-ENDIAN strings, the reference's
 * Int::Get32Bytes serialisation (secp256k1/Int.cpp:308).
 */
#ifndef KH_ORACLE_H
#define KH_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- L0 field (IntMod.cpp:855 ModMulK1, :977 ModSquareK1, :382 ModInv); canonical outputs ------ */
void kho_fe_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
void kho_fe_sqr(const uint8_t a[32], uint8_t out[32]);
void kho_fe_inv(const uint8_t a[32], uint8_t out[32]);      /* inv(0) = 0 like Int::ModInv */
void kho_fe_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);   /* Int::ModAdd IntMod.cpp:51 */
void kho_fe_sub(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);   /* Int::ModSub IntMod.cpp:97 */
void kho_fe_neg(const uint8_t a[32], uint8_t out[32]);                        /* Int::ModNeg IntMod.cpp:105, canonical */

/* ---- L1 group (SECP256K1.cpp:205 ComputePublicKey, :455 AddDirect) ---------------------------- */
void kho_pubkey(const uint8_t key[32], uint8_t xy[64]);
void kho_add_direct(const uint8_t a[64], const uint8_t b[64], uint8_t out[64]);
/* the 1024-point batch of thread_process (keyhunt.cpp:3348-3461): out[i] = (base+i*stride)*G     */
void kho_batch_points(const uint8_t base_key[32], const uint8_t stride[32], int with_y, uint8_t *out);

/* ---- L2 hashes -------------------------------------------------------------------------------- */
void kho_sha256(const uint8_t *in, uint64_t len, uint8_t out[32]);
void kho_ripemd160(const uint8_t *in, uint64_t len, uint8_t out[20]);
void kho_hash160_comp(int prefix, const uint8_t x[32], uint8_t out[20]);   /* SECP256K1.cpp:1207 */
void kho_hash160_uncomp(const uint8_t xy[64], uint8_t out[20]);            /* SECP256K1.cpp:1045 */
void kho_eth_addr(const uint8_t xy[64], uint8_t out[20]);                  /* keyhunt.cpp:5663   */
uint64_t kho_xxh64(const void *buf, uint64_t len, uint64_t seed);          /* xxhash.h:2512      */

/* ---- L3 bloom (bloom.cpp:154 bloom_init2, :215 bloom_add, :189 bloom_check) -------------------- */
void *kho_bloom_new(uint64_t entries);                    /* error = (long double)0.000001 */
void kho_bloom_free(void *h);
void kho_bloom_desc(void *h, uint64_t *entries, uint64_t *bits, uint64_t *bytes, uint32_t *hashes);
uint8_t *kho_bloom_data(void *h);
int kho_bloom_add(void *h, const void *buf, int len);
int kho_bloom_check(void *h, const void *buf, int len);

/* ---- L4 scan (thread_process keyhunt.cpp:3265) ------------------------------------------------- */
enum { KHO_MODE_XPOINT = 0, KHO_MODE_ADDRESS = 1, KHO_MODE_RMD160 = 2 };
enum { KHO_CRYPTO_BTC = 0, KHO_CRYPTO_ETH = 1 };
enum { KHO_SEARCH_UNCOMPRESS = 0, KHO_SEARCH_COMPRESS = 1, KHO_SEARCH_BOTH = 2 }; /* keyhunt.cpp:66-68 */
enum { KHO_HIT_COMP02 = 0, KHO_HIT_COMP03 = 1, KHO_HIT_UNCOMP = 2, KHO_HIT_ETH = 3, KHO_HIT_XPOINT = 4 };

typedef struct {
  uint8_t key_be[32];   /* reported private key (after the n-k fix-up of keyhunt.cpp:3629-3635)    */
  uint8_t matched[20];  /* the 20 bytes that matched the table                                    */
  uint8_t kind;         /* KHO_HIT_*                                                               */
  uint8_t pad[3];       /* pad[0] = with -e: the reference's candidate index l (0..11 BTC, 0..5 ETH, 0..2 xpoint) */
  uint64_t index;       /* point index inside the scanned range (key = start + index*stride)      */
} kho_hit;

/* targets: N raw 20-byte records (any order); builds bloom (entries = N<=10000?10000:N,
 * keyhunt.cpp:7608) and the ascending table (_sort keyhunt.cpp:4307)                             */
void *kho_targets_new(const uint8_t *raw20, uint64_t N);
void kho_targets_free(void *t);
void *kho_targets_bloom(void *t);                       /* a kho_bloom handle */
const uint8_t *kho_targets_table(void *t, uint64_t *N); /* sorted 20-byte records */
int kho_searchbinary(void *t, const uint8_t data[20]);  /* keyhunt.cpp:3065 */

/* vanity targets (-m vanity): kho_addvanity restates addvanity (keyhunt.cpp:6739) for one base58 prefix and writes its
 * r interval pairs (20-byte limits, flattened) to A/B, lowering *min_bytes like :6833; kho_targets_new_vanity builds the
 * handle kho_scan takes (mode KHO_MODE_RMD160, BTC): membership = vanityrmdmatch (:6677) */
int kho_b58tobin(uint8_t *bin, uint64_t *binsz, const char *b58, uint64_t b58sz);   /* base58/base58.c:39 */
int kho_addvanity(const char *target, uint8_t *A, uint8_t *B, int max_r, int *min_bytes);
void *kho_targets_new_vanity(const uint8_t *A, const uint8_t *B, uint64_t n, int min_bytes);

/* scans keys start + i*stride, i in [0, n_points) (n_points multiple of 1024); returns #hits
 * (writes at most max_hits, in ascending index order).  nthreads>1 splits batches over pthreads. */
int64_t kho_scan(void *targets, int mode, int crypto, int search, const uint8_t start[32],
                 const uint8_t stride[32], uint64_t n_points, kho_hit *hits, uint64_t max_hits,
                 int nthreads);

/* the same with FLAGENDOMORPHISM (-e, keyhunt.cpp:3408-3473, :3556-3618): endo != 0 tests the six (xpoint: three)
 * endomorphic candidates of every point and applies the reference's lambda / negation fix-ups */
int64_t kho_scan_ex(void *targets, int mode, int crypto, int search, int endo, const uint8_t start[32],
                    const uint8_t stride[32], uint64_t n_points, kho_hit *hits, uint64_t max_hits, int nthreads);

/* ---- BSGS (keyhunt.cpp:1450-1842 setup, :5284 thread_bPload, :4549 thread_process_bsgs) -------- */
typedef struct {
  uint8_t value[6];   /* X bytes [16..21] */
  uint8_t pad[2];
  uint64_t index;     /* baby index i (key i+1) */
} kho_bp_entry;       /* == struct bsgs_xvalue keyhunt.cpp:132, 16 bytes */

void *kho_bsgs_new(uint64_t n, uint32_t k, int nthreads);   /* NULL if n not an even power of two etc. */
void kho_bsgs_free(void *b);
void kho_bsgs_params(void *b, uint64_t *n, uint64_t *m, uint64_t *m2, uint64_t *m3, uint64_t *aux);
void *kho_bsgs_bloom(void *b, int tier /*1,2,3*/, int shard /*0..255*/); /* a kho_bloom handle */
const kho_bp_entry *kho_bsgs_table(void *b);            /* m3 entries, ascending (value, index) */
/* sequential search of [start,end) in windows of 2n keys (thread_process_bsgs, -B sequential);
 * returns 1 and the key when found. giant_steps (may be NULL) counts tier-1 probes done. */
int kho_bsgs_search(void *b, const uint8_t pub_xy[64], const uint8_t start[32], const uint8_t end[32],
                    uint8_t found_key[32], uint64_t *giant_steps, uint64_t *tier1_positives);

/* the same loop as the reference's BSGS server runs it (bsgsd.cpp:2462): base_check != 0 adds the base-point
 * comparison of bsgsd.cpp:2544 in front of every window */
int kho_bsgs_search_ex(void *b, const uint8_t pub_xy[64], const uint8_t start[32], const uint8_t end[32], int base_check,
                       uint8_t found_key[32], uint64_t *giant_steps, uint64_t *tier1_positives);

#ifdef __cplusplus
}
#endif
#endif
