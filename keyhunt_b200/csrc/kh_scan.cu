// kh_scan.cu — context, target set, scan kernels and the scan half of the C ABI (include/keyhunt_b200.h).
//
// Kernels (all sm_100a, integer pipes only — there is no dense contraction on this path):
//   kh_setup_kernel / kh_setup_fill_kernel   table entries + per-thread start centres (setup.cuh: two levels)
//   kh_scan_kernel<K> the batch walk of walk.cuh fused with hash -> bloom -> table probe (emit.cuh)
//   kh_bloom_build    bloom_add of every target record (atomicOr)
//   kh_derive_kernel  private key -> public key, hash160 (both forms), ETH address
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>

#include "plan.hpp"
#include "scan_kernel.cuh"

using namespace kh;

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
// set-up, phase 1 (setup.cuh): the 513 table entries, the 63 row offsets B_j and the row bases (walkers 0, 64, 128, ...), one scalar
// multiplication each
__global__ void __launch_bounds__(128) kh_setup_kernel(WalkSetup ws, uint32_t *gtab, uint32_t *centers, uint32_t *rowoffs, uint32_t *flags) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_rows = (ws.T + KH_SETUP_ROW - 1) / KH_SETUP_ROW;
  if (i < KH_TAB_ENTRIES) {
    uint32_t e[16];
    setup_table_entry(e, ws, (uint32_t)i);
#pragma unroll
    for (int k = 0; k < 16; k++) gtab[16 * i + k] = e[k];
  } else if (i < KH_TAB_ENTRIES + (KH_SETUP_ROW - 1)) {
    const uint32_t j = (uint32_t)(i - KH_TAB_ENTRIES) + 1;
    uint32_t e[16];
    setup_row_offset(e, ws, j);
#pragma unroll
    for (int k = 0; k < 16; k++) rowoffs[16 * (j - 1) + k] = e[k];
  } else if (i < KH_TAB_ENTRIES + (KH_SETUP_ROW - 1) + n_rows) {
    const uint64_t t = (i - KH_TAB_ENTRIES - (KH_SETUP_ROW - 1)) * KH_SETUP_ROW;
    fe cx, cy;
    // (an idle walker — its first batch lies beyond the segment — may sit anywhere, infinity included)
    if (!setup_center(cx, cy, ws, t) && (ws.n_batches == 0 || ws.first_batch + t < ws.n_batches)) atomicOr(flags, 1u);
#pragma unroll
    for (int l = 0; l < 8; l++) { centers[(uint64_t)l * ws.T + t] = cx.v[l]; centers[(uint64_t)(8 + l) * ws.T + t] = cy.v[l]; }
  }
}
// set-up, phase 2: one thread per row of 64 walkers, 63 affine additions behind one inversion (setup_row_fill)
__global__ void __launch_bounds__(64) kh_setup_fill_kernel(WalkSetup ws, uint32_t *centers, const uint32_t *rowoffs, uint32_t *flags) {
  const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row * KH_SETUP_ROW >= ws.T) return;
  fe pre[KH_SETUP_ROW - 1];
  if (!setup_row_fill(centers, rowoffs, ws, row, pre)) atomicOr(flags, 1u);
}

// fixed-base comb of G (ec.cuh): entry (w, d) = d * 2^(8w) * G, one plain scalar multiplication each, once per context
__global__ void __launch_bounds__(128) kh_comb_kernel(uint32_t *comb) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 32 * 256) return;
  const uint32_t w = i >> 8, d = i & 255;
  u256 k;
#pragma unroll
  for (int l = 0; l < 8; l++) k.v[l] = 0;
  k.v[w >> 2] = d << (8 * (w & 3));
  ge p;
  ge_mul_g(p, k);
#pragma unroll
  for (int l = 0; l < 8; l++) { comb[(size_t)i * 16 + l] = p.x.v[l]; comb[(size_t)i * 16 + 8 + l] = p.y.v[l]; }
}

// exact prefix bitmap over the first k bits of every target record (ScanTargets::pre)
__global__ void kh_pre_build(uint32_t *pre, uint32_t k, const uint32_t *table_be, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t idx = table_be[5 * i] >> (32 - k);
  atomicOr(pre + (idx >> 5), 1u << (idx & 31));
}

// the uploaded 20-byte records -> 5 big-endian-packed words each (numeric order = memcmp order), in place
__global__ void kh_table_pack(uint32_t *table, uint64_t n_words) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_words) table[i] = bswap32(table[i]);
}

__global__ void kh_bloom_build(BloomDev bl, const uint32_t *table_be, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[5];
#pragma unroll
  for (int k = 0; k < 5; k++) w[k] = bswap32(table_be[5 * i + k]);
  bloom_add20(bl, w);
}

struct DevKeyInfo {
  uint32_t x[8], y[8], hc[5], hu[5], eth[5], inf;
};
__global__ void __launch_bounds__(64) kh_derive_kernel(const u256 *keys, uint64_t n, DevKeyInfo *out, const uint32_t *comb) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge p;
  ge_mul_g_comb(p, keys[i], comb);
  DevKeyInfo r;
#pragma unroll
  for (int k = 0; k < 8; k++) { r.x[k] = p.x.v[k]; r.y[k] = p.y.v[k]; }
  r.inf = p.inf;
  hash160_compressed(r.hc, 2u + (p.y.v[0] & 1u), p.x);
  hash160_uncompressed(r.hu, p.x, p.y);
  eth_address(r.eth, p.x, p.y);
  out[i] = r;
}

// ---------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------
void kh_time_begin(kh_ctx *c) { cudaEventRecord(c->ev0, c->stream); }
double kh_time_end(kh_ctx *c) {
  cudaEventRecord(c->ev1, c->stream);
  cudaEventSynchronize(c->ev1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  return (double)ms;
}

uint64_t kh_pick_T(kh_ctx *c, uint64_t n_batches) {
  uint64_t tmax = (uint64_t)c->sm_count * (uint64_t)c->threads_per_sm;
  tmax = (tmax / KH_T_ALIGN) * KH_T_ALIGN;
  if (tmax < KH_T_ALIGN) tmax = KH_T_ALIGN;
  uint64_t need = ((n_batches + KH_T_ALIGN - 1) / KH_T_ALIGN) * KH_T_ALIGN;
  return need < tmax ? need : tmax;
}

int kh_ensure_walk_buffers(kh_ctx *c, uint64_t T) {
  if (!c->d_gtab) KH_CUDA(c, cudaMalloc(&c->d_gtab, KH_TAB_WORDS * sizeof(uint32_t)));
  if (!c->d_rowoffs) KH_CUDA(c, cudaMalloc(&c->d_rowoffs, (KH_SETUP_ROW - 1) * 16 * sizeof(uint32_t)));
  if (!c->d_flags) {
    KH_CUDA(c, cudaMalloc(&c->d_flags, 16 * sizeof(uint32_t)));
    KH_CUDA(c, cudaMemsetAsync(c->d_flags, 0, 16 * sizeof(uint32_t), c->stream));
  }
  if (T > c->T_alloc) {
    if (c->d_centers) cudaFree(c->d_centers);
    if (c->d_scratch) cudaFree(c->d_scratch);
    c->d_centers = nullptr; c->d_scratch = nullptr; c->T_alloc = 0;
    KH_CUDA(c, cudaMalloc(&c->d_centers, 16 * T * sizeof(uint32_t)));
    KH_CUDA(c, cudaMalloc(&c->d_scratch, (size_t)1024 * T * sizeof(kh_u4)));
    c->T_alloc = T;
  }
  return KH_OK;
}

int kh_ensure_comb(kh_ctx *c) {
  if (c->d_comb) return KH_OK;
  KH_CUDA(c, cudaMalloc(&c->d_comb, (size_t)KH_COMB_WORDS * sizeof(uint32_t)));
  kh_time_begin(c);
  kh_comb_kernel<<<(32 * 256 + 127) / 128, 128, 0, c->stream>>>(c->d_comb);
  c->stats.setup_ms += kh_time_end(c);
  c->stats.other_launches += 1;
  KH_CUDA(c, cudaGetLastError());
  return KH_OK;
}

int kh_run_setup(kh_ctx *c, const WalkSetup &ws_in) {
  int rcc = kh_ensure_comb(c);
  if (rcc) return rcc;
  WalkSetup ws = ws_in;
  ws.comb = c->d_comb;
  const uint64_t n_rows = (ws.T + KH_SETUP_ROW - 1) / KH_SETUP_ROW;
  const uint64_t n = KH_TAB_ENTRIES + (KH_SETUP_ROW - 1) + n_rows;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  kh_time_begin(c);
  kh_setup_kernel<<<blocks, 128, 0, c->stream>>>(ws, c->d_gtab, c->d_centers, c->d_rowoffs, c->d_flags);
  kh_setup_fill_kernel<<<(unsigned)((n_rows + 63) / 64), 64, 0, c->stream>>>(ws, c->d_centers, c->d_rowoffs, c->d_flags);
  KH_CUDA(c, cudaGetLastError());
  uint32_t flag = 0;
  KH_CUDA(c, cudaMemcpyAsync(&flag, c->d_flags, sizeof(flag), cudaMemcpyDeviceToHost, c->stream));
  c->stats.setup_ms += kh_time_end(c);
  c->stats.other_launches += 2;
  KH_CUDA(c, cudaGetLastError());
  if (flag) {
    cudaMemsetAsync(c->d_flags, 0, sizeof(uint32_t), c->stream);
    return kh_fail(c, KH_EINVAL, "a start point is the point at infinity (range touches key 0 mod n)");
  }
  return KH_OK;
}

static int ensure_hit_buffer(kh_ctx *c) {
  if (c->d_hits && c->hits_alloc == c->hit_capacity) return KH_OK;
  if (c->d_hits) cudaFree(c->d_hits);
  c->d_hits = nullptr;
  KH_CUDA(c, cudaMalloc(&c->d_hits, (size_t)c->hit_capacity * sizeof(RawHit)));
  if (!c->d_hit_count) KH_CUDA(c, cudaMalloc(&c->d_hit_count, 4 * sizeof(uint32_t)));
  KH_CUDA(c, cudaMemsetAsync(c->d_hit_count, 0, 4 * sizeof(uint32_t), c->stream));
  c->hits_alloc = c->hit_capacity;
  return KH_OK;
}

static void words_to_bytes20(uint8_t out[20], const uint32_t w[5]) {
  for (int i = 0; i < 5; i++) { out[4 * i] = (uint8_t)w[i]; out[4 * i + 1] = (uint8_t)(w[i] >> 8); out[4 * i + 2] = (uint8_t)(w[i] >> 16); out[4 * i + 3] = (uint8_t)(w[i] >> 24); }
}
static void limbs_to_be(uint8_t out[32], const uint32_t v[8]) {
  for (int i = 0; i < 8; i++) { uint8_t *p = out + 4 * (7 - i); p[0] = (uint8_t)(v[i] >> 24); p[1] = (uint8_t)(v[i] >> 16); p[2] = (uint8_t)(v[i] >> 8); p[3] = (uint8_t)v[i]; }
}

static int derive_dev(kh_ctx *c, const std::vector<u256> &keys, std::vector<DevKeyInfo> &out) {
  const uint64_t n = keys.size();
  out.resize(n);
  if (!n) return KH_OK;
  {
    int rcc = kh_ensure_comb(c);
    if (rcc) return rcc;
  }
  u256 *d_keys = nullptr;
  DevKeyInfo *d_out = nullptr;
  KH_CUDA(c, cudaMalloc(&d_keys, n * sizeof(u256)));
  if (cudaMalloc(&d_out, n * sizeof(DevKeyInfo)) != cudaSuccess) { cudaFree(d_keys); return kh_fail(c, KH_ENOMEM, "cudaMalloc derive"); }
  cudaMemcpyAsync(d_keys, keys.data(), n * sizeof(u256), cudaMemcpyHostToDevice, c->stream);
  kh_time_begin(c);
  kh_derive_kernel<<<(unsigned)((n + 63) / 64), 64, 0, c->stream>>>(d_keys, n, d_out, c->d_comb);
  cudaMemcpyAsync(out.data(), d_out, n * sizeof(DevKeyInfo), cudaMemcpyDeviceToHost, c->stream);
  c->stats.aux_ms += kh_time_end(c);
  c->stats.other_launches += 1;
  cudaError_t e = cudaGetLastError();
  cudaFree(d_keys); cudaFree(d_out);
  if (e != cudaSuccess) return kh_fail(c, KH_ENODEV, "derive kernel: %s", cudaGetErrorString(e));
  return KH_OK;
}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int kh_create(kh_ctx **out, int device_ordinal) {
  if (!out) return KH_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device_ordinal < 0 || device_ordinal >= ndev) return KH_ENODEV;
  if (cudaSetDevice(device_ordinal) != cudaSuccess) return KH_ENODEV;
  kh_ctx *c = new kh_ctx();
  c->device = device_ordinal;
  if (cudaGetDeviceProperties(&c->prop, device_ordinal) != cudaSuccess) { delete c; return KH_ENODEV; }
  c->sm_count = c->prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return KH_ENODEV; }
  // KH_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity (A/B knob): measured on B200 it changes neither the
  // 243 B of DRAM reads per BSGS giant step nor any throughput, so the driver default is left alone unless asked.
  if (const char *g = getenv("KH_L2_FETCH")) {
    const size_t gran = (size_t)atoi(g);
    if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
  }
  cudaEventCreate(&c->ev0);
  cudaEventCreate(&c->ev1);
  *out = c;
  return KH_OK;
}

void kh_destroy(kh_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  void *ptrs[] = {c->d_gtab, c->d_centers, c->d_scratch, c->d_flags, c->d_bloom, c->d_table, c->d_hits, c->d_hit_count,
                  c->d_tier[0], c->d_tier[1], c->d_tier[2], c->d_bptable, c->d_aux_tab, c->d_vanity, c->d_pre, c->d_giant_cands, c->d_giant_cnt, c->d_giant_key, c->d_bsgs_pre, c->d_comb, c->d_rowoffs, c->d_sha2};
  for (void *p : ptrs) if (p) cudaFree(p);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  delete c;
}

const char *kh_last_error(kh_ctx *c) { return c ? c->err.c_str() : "no context"; }

int kh_set_option(kh_ctx *c, const char *name, int64_t value) {
  if (!c || !name) return KH_EINVAL;
  if (!strcmp(name, "threads_per_sm")) {
    if (value < 32 || value > 32768) return kh_fail(c, KH_EINVAL, "threads_per_sm out of range");
    c->threads_per_sm = (int)value;
  } else if (!strcmp(name, "steps_per_launch")) {
    if (value < 1 || value > (1 << 20)) return kh_fail(c, KH_EINVAL, "steps_per_launch out of range");
    c->steps_per_launch = (int)value;
  } else if (!strcmp(name, "endomorphism")) {      // -e (FLAGENDOMORPHISM, keyhunt.cpp:924)
    c->endomorphism = value ? 1 : 0;
  } else if (!strcmp(name, "prefilter")) {         // exact prefix bitmap in front of the bloom (takes effect at the next kh_set_targets)
    c->prefilter = value ? 1 : 0;
  } else if (!strcmp(name, "bsgs_prefilter")) {    // baby-point prefix bitmap in front of the tier-1 bloom (next kh_bsgs_build)
    c->bsgs_prefilter = value ? 1 : 0;
  } else if (!strcmp(name, "bsgs_binned_build")) { // 0 = every baby step sets its bits directly (A/B, and the small-table path)
    if (value < 0 || value > 2) return kh_fail(c, KH_EINVAL, "bsgs_binned_build is 0, 1 (tables of 2^26 baby steps and more) or 2 (always)");
    c->bsgs_binned_build = (int)value;
  } else if (!strcmp(name, "bsgs_base_check")) {   // the reference SERVER's search loop (bsgsd.cpp:2544)
    c->bsgs_base_check = value ? 1 : 0;
  } else if (!strcmp(name, "hit_capacity")) {
    if (value < 16 || value > (1 << 26)) return kh_fail(c, KH_EINVAL, "hit_capacity out of range");
    c->hit_capacity = (uint32_t)value;
  } else {
    return kh_fail(c, KH_EINVAL, "unknown option %s", name);
  }
  return KH_OK;
}

int kh_bloom_params(uint64_t entries, kh_bloom_desc *out) {  // bloom_init2, bloom/bloom.cpp:154-187
  if (!out || entries < 1000) return KH_EINVAL;
  const long double error = 0.000001;            // double literal widened, as initBloomFilter passes it (keyhunt.cpp:7620)
  const long double num = -logl(error);
  const long double denom = 0.480453013918201;   // ln(2)^2
  const double bpe = (double)(num / denom);      // struct bloom::bpe is a double
  const long double allbits = (long double)entries * bpe;
  out->entries = entries;
  out->bits = (uint64_t)allbits;
  out->bytes = out->bits / 8 + ((out->bits % 8) ? 1 : 0);
  out->hashes = (uint8_t)ceil(0.693147180559945 * bpe);
  out->pad = 0;
  return KH_OK;
}

int kh_set_targets(kh_ctx *c, int mode, int crypto, int search, const uint8_t *records20, uint64_t n, const kh_bloom_desc *desc,
                   const uint8_t *bloom_bits) {
  if (!c) return KH_EINVAL;
  cudaSetDevice(c->device);
  if (!records20 && n) return kh_fail(c, KH_EINVAL, "records20 is NULL");
  int kind;
  if (mode == KH_MODE_XPOINT) kind = KH_SCAN_XPOINT;
  else if (mode == KH_MODE_ADDRESS || mode == KH_MODE_RMD160) {
    if (crypto == KH_CRYPTO_ETH) {
      if (mode != KH_MODE_ADDRESS) return kh_fail(c, KH_EINVAL, "ETH needs -m address");
      kind = KH_SCAN_ETH;
    } else if (crypto == KH_CRYPTO_BTC) {
      if (search == KH_SEARCH_COMPRESS) kind = KH_SCAN_COMP;
      else if (search == KH_SEARCH_UNCOMPRESS) kind = KH_SCAN_UNCOMP;
      else if (search == KH_SEARCH_BOTH) kind = KH_SCAN_BOTH;
      else return kh_fail(c, KH_EINVAL, "bad search type %d", search);
    } else return kh_fail(c, KH_EINVAL, "bad crypto %d", crypto);
  } else return kh_fail(c, KH_EINVAL, "mode %d is not a scan mode", mode);

  kh_bloom_desc d;
  if (desc) d = *desc;
  else if (kh_bloom_params(n <= 10000 ? 10000 : n, &d) != KH_OK) return kh_fail(c, KH_EINVAL, "bloom sizing failed");
  if (d.bits == 0 || d.hashes == 0 || d.bytes < (d.bits + 7) / 8) return kh_fail(c, KH_EINVAL, "bad bloom descriptor");

  // sorted table (_sort keyhunt.cpp:4307: ascending memcmp order).  The reference hands over an already sorted addressTable: then
  // the caller's buffer goes to the device as it is (the words are packed there, kh_table_pack) and the host only keeps a copy
  bool sorted = true;
  for (uint64_t i = 1; i < n && sorted; i++) sorted = memcmp(records20 + 20 * (i - 1), records20 + 20 * i, 20) <= 0;
  c->h_table20.resize(20 * n);
  if (sorted) {
    if (n) memcpy(c->h_table20.data(), records20, 20 * n);
  } else {
    std::vector<uint64_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return memcmp(records20 + 20 * a, records20 + 20 * b, 20) < 0; });
    for (uint64_t i = 0; i < n; i++) memcpy(&c->h_table20[20 * i], records20 + 20 * order[i], 20);
  }
  // device buffers are kept across calls when the sizes repeat (a caller that re-sends the same target set every step)
  c->have_targets = false;
  const size_t table_bytes = std::max<size_t>(5 * n, 8) * sizeof(uint32_t);
  if (!c->d_table || c->table_alloc != table_bytes) {
    if (c->d_table) { cudaFree(c->d_table); c->d_table = nullptr; }
    KH_CUDA(c, cudaMalloc(&c->d_table, table_bytes));
    c->table_alloc = table_bytes;
  }
  if (n) {
    KH_CUDA(c, cudaMemcpyAsync(c->d_table, sorted ? records20 : c->h_table20.data(), 20 * n, cudaMemcpyHostToDevice, c->stream));
    kh_table_pack<<<(unsigned)((5 * n + 255) / 256), 256, 0, c->stream>>>(c->d_table, 5 * n);
    c->stats.other_launches += 1;
  }
  const size_t bloom_alloc = (size_t)((d.bytes + 15) / 16) * 16;
  if (!c->d_bloom || c->bloom_alloc != bloom_alloc) {
    if (c->d_bloom) { cudaFree(c->d_bloom); c->d_bloom = nullptr; }
    KH_CUDA(c, cudaMalloc(&c->d_bloom, bloom_alloc));
    c->bloom_alloc = bloom_alloc;
  }
  KH_CUDA(c, cudaMemsetAsync(c->d_bloom, 0, bloom_alloc, c->stream));
  if (bloom_bits) {
    KH_CUDA(c, cudaMemcpyAsync(c->d_bloom, bloom_bits, d.bytes, cudaMemcpyHostToDevice, c->stream));
  } else if (n) {
    BloomDev bl;
    bl.bf = c->d_bloom; bl.bits = d.bits; bl.magic = (~0ULL) / d.bits; bl.stride = 0; bl.hashes = d.hashes; bl.pad = 0;
    kh_time_begin(c);
    kh_bloom_build<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(bl, c->d_table, n);
    c->stats.aux_ms += kh_time_end(c);
    c->stats.other_launches += 1;
  }
  // prefix bitmap (ScanTargets::pre): >= 256 bits per target (fill <= 0.4 %) between 2^16 and 2^32 bits (8 KB .. 512 MB),
  // filled on the device from the table that was just uploaded
  {
    uint32_t k = 16;
    if (c->prefilter) while (k < 32 && (1ull << k) < 256ull * n) k++;
    const size_t pre_bytes = (size_t)1 << (k - 3);
    if (!c->d_pre || c->pre_k != k) {
      if (c->d_pre) { cudaFree(c->d_pre); c->d_pre = nullptr; }
      KH_CUDA(c, cudaMalloc(&c->d_pre, pre_bytes));
    }
    KH_CUDA(c, cudaMemsetAsync(c->d_pre, c->prefilter ? 0x00 : 0xFF, pre_bytes, c->stream));
    if (c->prefilter && n) {
      kh_pre_build<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_pre, k, c->d_table, n);
      c->stats.other_launches += 1;
    }
    c->pre_k = k;
  }
  KH_CUDA(c, cudaStreamSynchronize(c->stream));
  KH_CUDA(c, cudaGetLastError());
  c->bloom_desc = d;
  c->n_targets = n;
  c->n_vanity = 0;
  c->mode = mode; c->crypto = crypto; c->search = search; c->scan_kind = kind;
  c->have_targets = true;
  return KH_OK;
}

int kh_set_vanity(kh_ctx *c, int search, const uint8_t *a20, const uint8_t *b20, uint64_t n) {
  if (!c) return KH_EINVAL;
  cudaSetDevice(c->device);
  if (!a20 || !b20 || n == 0) return kh_fail(c, KH_EINVAL, "There aren't any vanity targets");
  if (n > (1u << 20)) return kh_fail(c, KH_EINVAL, "too many vanity intervals");
  int kind;
  if (search == KH_SEARCH_COMPRESS) kind = KH_SCAN_COMP;
  else if (search == KH_SEARCH_UNCOMPRESS) kind = KH_SCAN_UNCOMP;
  else if (search == KH_SEARCH_BOTH) kind = KH_SCAN_BOTH;
  else return kh_fail(c, KH_EINVAL, "bad search type %d", search);
  std::vector<uint32_t> van(2048 + 10 * n, 0u);
  auto pack = [](uint32_t *dst, const uint8_t *p) {
    for (int k = 0; k < 5; k++) dst[k] = ((uint32_t)p[4 * k] << 24) | ((uint32_t)p[4 * k + 1] << 16) | ((uint32_t)p[4 * k + 2] << 8) | p[4 * k + 3];
  };
  for (uint64_t i = 0; i < n; i++) {
    const uint8_t *a = a20 + 20 * i, *b = b20 + 20 * i;
    pack(&van[2048 + 10 * i], a);
    pack(&van[2048 + 10 * i + 5], b);
    if (memcmp(a, b, 20) > 0) continue;                     // empty interval: nothing can match it
    const uint32_t pa = ((uint32_t)a[0] << 8) | a[1], pb = ((uint32_t)b[0] << 8) | b[1];
    for (uint32_t p = pa; p <= pb; p++) van[p >> 5] |= 1u << (p & 31);
  }
  if (c->d_vanity) { cudaFree(c->d_vanity); c->d_vanity = nullptr; }
  c->have_targets = false;
  KH_CUDA(c, cudaMalloc(&c->d_vanity, van.size() * sizeof(uint32_t)));
  KH_CUDA(c, cudaMemcpyAsync(c->d_vanity, van.data(), van.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  KH_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_vanity = (uint32_t)n;
  c->n_targets = 0;
  c->h_table20.clear();
  memset(&c->bloom_desc, 0, sizeof(c->bloom_desc));
  c->mode = KH_MODE_VANITY; c->crypto = KH_CRYPTO_BTC; c->search = search; c->scan_kind = kind;
  c->have_targets = true;
  return KH_OK;
}

int kh_get_bloom(kh_ctx *c, kh_bloom_desc *desc, uint8_t *dst, uint64_t cap) {
  if (!c) return KH_EINVAL;
  if (!c->have_targets) return kh_fail(c, KH_ESTATE, "no targets set");
  cudaSetDevice(c->device);
  if (desc) *desc = c->bloom_desc;
  if (dst) {
    if (cap < c->bloom_desc.bytes) return kh_fail(c, KH_EINVAL, "buffer too small");
    KH_CUDA(c, cudaMemcpyAsync(dst, c->d_bloom, c->bloom_desc.bytes, cudaMemcpyDeviceToHost, c->stream));
    KH_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return KH_OK;
}

int kh_get_table(kh_ctx *c, uint8_t *dst20, uint64_t cap_records, uint64_t *n_records) {
  if (!c) return KH_EINVAL;
  if (!c->have_targets) return kh_fail(c, KH_ESTATE, "no targets set");
  if (n_records) *n_records = c->n_targets;
  if (dst20) {
    if (cap_records < c->n_targets) return kh_fail(c, KH_EINVAL, "buffer too small");
    // read back from the device copy so the caller sees what the kernels see
    std::vector<uint32_t> packed(5 * c->n_targets);
    cudaSetDevice(c->device);
    if (c->n_targets) {
      KH_CUDA(c, cudaMemcpyAsync(packed.data(), c->d_table, packed.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
      KH_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    for (uint64_t i = 0; i < 5 * c->n_targets; i++) {
      dst20[4 * i] = (uint8_t)(packed[i] >> 24); dst20[4 * i + 1] = (uint8_t)(packed[i] >> 16);
      dst20[4 * i + 2] = (uint8_t)(packed[i] >> 8); dst20[4 * i + 3] = (uint8_t)packed[i];
    }
  }
  return KH_OK;
}

}  // extern "C"

template <int KIND>
static cudaError_t launch_scan(kh_ctx *c, const WalkParams &wp, const ScanTargets &tg) {
  if (c->endomorphism) return kh_launch_scan_kernel<KIND, true, false>(c, wp, tg);
  return kh_launch_scan_kernel<KIND, false, false>(c, wp, tg);
}

extern "C" {

int kh_scan(kh_ctx *c, const uint8_t start_be[32], const uint8_t stride_be[32], uint64_t n_points) {
  if (!c || !start_be || !stride_be) return KH_EINVAL;
  if (!c->have_targets) return kh_fail(c, KH_ESTATE, "kh_scan before kh_set_targets");
  if (n_points == 0 || (n_points % KH_GRP) != 0) return kh_fail(c, KH_EINVAL, "n_points must be a positive multiple of 1024");
  cudaSetDevice(c->device);
  const uint64_t n_batches = n_points / KH_GRP;
  int rc = ensure_hit_buffer(c);
  if (rc) return rc;

  WalkSetup ws;
  memset(&ws, 0, sizeof(ws));
  u256_from_be(ws.s, stride_be);
  u256_from_be(ws.k0, start_be);
  bool stride_zero = true;
  for (int i = 0; i < 8; i++) stride_zero &= (ws.s.v[i] == 0);
  if (stride_zero) return kh_fail(c, KH_EINVAL, "stride is zero");
  const uint64_t T_cap = kh_pick_T(c, n_batches);
  // scalars are plain 256-bit integers on the device (u256_add_mul64 works mod 2^256, k*G needs no reduction mod n):
  // the largest one a scan forms is start + stride*(n_points + T*1024) (the hop W = T*1024*S); refuse ranges where
  // that wraps instead of walking wrong points
  {
    unsigned __int128 carry = 0;
    const uint64_t reach = n_points + (T_cap + 4 * KH_T_ALIGN) * (uint64_t)KH_GRP;
    for (int i = 0; i < 8; i++) {
      carry += (unsigned __int128)ws.s.v[i] * reach + ws.k0.v[i];
      carry >>= 32;
    }
    if (carry != 0) return kh_fail(c, KH_EINVAL, "start + stride*n_points reaches 2^256");
  }
  // the batches of this scan as segments that the kernels can walk without ever meeting a zero difference in the middle of a
  // walk (plan.hpp): almost always ONE segment with T = T_cap walkers
  std::vector<ScanSegment> plan;
  if (!plan_scan(ws.k0, ws.s, n_batches, T_cap, KH_T_ALIGN, plan)) return kh_fail(c, KH_EINVAL, "stride is 0 mod n");
  uint64_t T_max = 0;
  for (const ScanSegment &sg : plan) T_max = std::max(T_max, sg.T);
  rc = kh_ensure_walk_buffers(c, T_max);
  if (rc) return rc;

  ScanTargets tg;
  const bool vanity = (c->mode == KH_MODE_VANITY);
  tg.van = vanity ? c->d_vanity : nullptr; tg.van_n = vanity ? c->n_vanity : 0u;
  tg.pre = c->d_pre; tg.pre_k = c->pre_k;
  if (!c->d_sha2) {                // SHA-256 schedules of the uncompressed key's second block (hash.cuh KH_SHA_UNC2_TAB), once per context
    std::vector<uint32_t> tab(KH_SHA2TAB_WORDS);
    for (uint32_t v = 0; v < 256; v++) sha_unc2_table_row(&tab[(size_t)v * KH_SHA2TAB_STRIDE], v);
    KH_CUDA(c, cudaMalloc(&c->d_sha2, tab.size() * sizeof(uint32_t)));
    KH_CUDA(c, cudaMemcpyAsync(c->d_sha2, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    KH_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  tg.sha2 = c->d_sha2;
  tg.bloom.bf = c->d_bloom; tg.bloom.bits = c->bloom_desc.bits; tg.bloom.magic = vanity ? 0ull : (~0ULL) / c->bloom_desc.bits;
  tg.bloom.stride = 0; tg.bloom.hashes = c->bloom_desc.hashes; tg.bloom.pad = 0;
  tg.table = c->d_table; tg.n = c->n_targets;
  tg.sink.hits = c->d_hits; tg.sink.count = c->d_hit_count; tg.sink.cap = c->hits_alloc; tg.sink.pad = 0;

  uint64_t launches = 0;
  for (const ScanSegment &sg : plan) {
    const uint64_t T = sg.T;
    ws.q.inf = 1; ws.neg = 0; ws.T = T; ws.first_batch = sg.first; ws.n_batches = sg.end;
    rc = kh_run_setup(c, ws);
    if (rc) return rc;
    WalkParams wp;
    wp.gtab = c->d_gtab; wp.centers = c->d_centers; wp.scratch = c->d_scratch;
    wp.T = T; wp.n_batches = sg.end; wp.steps = (uint32_t)c->steps_per_launch; wp.pad = 0;
    kh_time_begin(c);
    for (uint64_t base = sg.first; base < sg.end; base += (uint64_t)wp.steps * T) {
      wp.batch_base = base;
      cudaError_t e;
      if (vanity) e = kh_launch_vanity(c, c->scan_kind, wp, tg);
      else switch (c->scan_kind) {
        case KH_SCAN_XPOINT: e = launch_scan<KH_SCAN_XPOINT>(c, wp, tg); break;
        case KH_SCAN_COMP: e = launch_scan<KH_SCAN_COMP>(c, wp, tg); break;
        case KH_SCAN_UNCOMP: e = launch_scan<KH_SCAN_UNCOMP>(c, wp, tg); break;
        case KH_SCAN_BOTH: e = launch_scan<KH_SCAN_BOTH>(c, wp, tg); break;
        default: e = launch_scan<KH_SCAN_ETH>(c, wp, tg); break;
      }
      if (e != cudaSuccess) return kh_fail(c, KH_ENODEV, "scan launch: %s", cudaGetErrorString(e));
      launches++;
    }
    c->stats.walk_ms += kh_time_end(c);
    {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return kh_fail(c, KH_ENODEV, "scan kernel: %s", cudaGetErrorString(e));
    }
    c->stats.collapsed_batches += sg.collapsed;
    c->stats.walker_threads = std::max<uint64_t>(T, (&sg == &plan[0]) ? 0 : c->stats.walker_threads);
  }
  c->stats.walk_launches += launches;
  c->stats.points += n_points;

  // collect raw hits of this scan and convert them while start/stride are at hand
  uint32_t count = 0;
  KH_CUDA(c, cudaMemcpyAsync(&count, c->d_hit_count, sizeof(count), cudaMemcpyDeviceToHost, c->stream));
  KH_CUDA(c, cudaStreamSynchronize(c->stream));
  if (count) {
    if (count > c->hits_alloc) { c->overflowed = true; count = c->hits_alloc; }
    std::vector<RawHit> raw(count);
    KH_CUDA(c, cudaMemcpyAsync(raw.data(), c->d_hits, count * sizeof(RawHit), cudaMemcpyDeviceToHost, c->stream));
    KH_CUDA(c, cudaMemsetAsync(c->d_hit_count, 0, sizeof(uint32_t), c->stream));
    KH_CUDA(c, cudaStreamSynchronize(c->stream));
    std::sort(raw.begin(), raw.end(), [](const RawHit &a, const RawHit &b) {
      const uint64_t ia = a.batch * KH_GRP + a.idx, ib = b.batch * KH_GRP + b.idx;
      if (ia != ib) return ia < ib;
      return a.kind != b.kind ? a.kind < b.kind : a.variant < b.variant;
    });
    // keyfound = index*stride + start (keyhunt.cpp:3625-3627)
    std::vector<u256> keys(count);
    for (uint32_t i = 0; i < count; i++) {
      u256_add_mul64(keys[i], ws.k0, ws.s, raw[i].batch * KH_GRP + raw[i].idx);
      const uint32_t order[8] = KH_N;            // a chunk that runs past n: report the key mod n (< 2^256 < 2n: one subtraction)
      uint32_t d[8];
      if (!kh_sub8(d, keys[i].v, order)) for (int l = 0; l < 8; l++) keys[i].v[l] = d[l];
    }
    std::vector<DevKeyInfo> info;
    rc = derive_dev(c, keys, info);
    if (rc) return rc;
    auto negate = [&](uint32_t i) { u256 nk; u256_neg_mod_n(nk, keys[i]); keys[i] = nk; };
    if (!c->endomorphism) {
      // compressed matches are parity-blind: if the real prefix differs from the matched one the
      // target's key is n - k (keyhunt.cpp:3629-3634)
      for (uint32_t i = 0; i < count; i++) {
        const uint32_t odd = info[i].y[0] & 1u;
        if ((raw[i].kind == KH_KIND_COMP02 && odd) || (raw[i].kind == KH_KIND_COMP03 && !odd)) negate(i);
      }
    } else {
      // -e fix-ups, candidate index l = raw.variant: multiply by lambda^(l/2) (ModMulK1order), then
      //   compress  (keyhunt.cpp:3566-3613): negate by the parity of the ORIGINAL key's Y against the matched prefix
      //   uncompress (:3652-3682) / ETH (:3714-3744): recompute the hash of the candidate key, negate on mismatch
      //   xpoint    (:3782-3806): no negation
      const u256 lam[3] = {{{1, 0, 0, 0, 0, 0, 0, 0}}, {KH_LAMBDA}, {KH_LAMBDA2}};
      std::vector<uint32_t> recheck;
      for (uint32_t i = 0; i < count; i++) {
        const uint32_t l = raw[i].variant, kind = raw[i].kind;
        const uint32_t v = (kind == KH_KIND_UNCOMP) ? (l - 6) / 2 : (kind == KH_KIND_XPOINT ? l : l / 2);
        const uint32_t odd = info[i].y[0] & 1u;
        if (v) { u256 t; u256_mulmod_n(t, keys[i], lam[v]); keys[i] = t; }
        if (kind == KH_KIND_COMP02 || kind == KH_KIND_COMP03) {
          if ((kind == KH_KIND_COMP02 && odd) || (kind == KH_KIND_COMP03 && !odd)) negate(i);
        } else if (kind == KH_KIND_UNCOMP || kind == KH_KIND_ETH) {
          recheck.push_back(i);
        }
      }
      if (!recheck.empty()) {
        std::vector<u256> k2(recheck.size());
        for (size_t j = 0; j < recheck.size(); j++) k2[j] = keys[recheck[j]];
        std::vector<DevKeyInfo> i2;
        rc = derive_dev(c, k2, i2);
        if (rc) return rc;
        for (size_t j = 0; j < recheck.size(); j++) {
          const uint32_t i = recheck[j];
          const uint32_t *hh = (raw[i].kind == KH_KIND_ETH) ? i2[j].eth : i2[j].hu;
          if (memcmp(hh, raw[i].h, 20) != 0) negate(i);
        }
      }
    }
    rc = derive_dev(c, keys, info);       // public keys of the keys that are reported
    if (rc) return rc;
    for (uint32_t i = 0; i < count; i++) {
      kh_hit h;
      memset(&h, 0, sizeof(h));
      u256_to_be(h.key_be, keys[i]);
      limbs_to_be(h.pub_x, info[i].x);
      limbs_to_be(h.pub_y, info[i].y);
      words_to_bytes20(h.matched, raw[i].h);
      h.kind = (uint8_t)raw[i].kind;
      h.pad[0] = (uint8_t)raw[i].variant;
      h.index = raw[i].batch * KH_GRP + raw[i].idx;
      c->ready_hits.push_back(h);
    }
  }
  return KH_OK;
}

int kh_poll_hits(kh_ctx *c, kh_hit *out, int max, int *n) {
  if (!c || !n || (max > 0 && !out)) return KH_EINVAL;
  int k = 0;
  while (k < max && k < (int)c->ready_hits.size()) { out[k] = c->ready_hits[k]; k++; }
  c->ready_hits.erase(c->ready_hits.begin(), c->ready_hits.begin() + k);
  *n = k;
  if (c->overflowed) { c->overflowed = false; return kh_fail(c, KH_EOVERFLOW, "device hit buffer overflowed; hits were dropped"); }
  return KH_OK;
}

int kh_derive(kh_ctx *c, const uint8_t *keys_be, uint64_t n, kh_keyinfo *out) {
  if (!c || (!keys_be && n) || (!out && n)) return KH_EINVAL;
  cudaSetDevice(c->device);
  std::vector<u256> keys(n);
  for (uint64_t i = 0; i < n; i++) u256_from_be(keys[i], keys_be + 32 * i);
  std::vector<DevKeyInfo> info;
  int rc = derive_dev(c, keys, info);
  if (rc) return rc;
  for (uint64_t i = 0; i < n; i++) {
    memset(&out[i], 0, sizeof(kh_keyinfo));
    limbs_to_be(out[i].pub_x, info[i].x);
    limbs_to_be(out[i].pub_y, info[i].y);
    words_to_bytes20(out[i].h160_comp, info[i].hc);
    words_to_bytes20(out[i].h160_uncomp, info[i].hu);
    words_to_bytes20(out[i].eth, info[i].eth);
  }
  return KH_OK;
}

int kh_get_stats(kh_ctx *c, kh_stats *out, int reset) {
  if (!c || !out) return KH_EINVAL;
  *out = c->stats;
  if (reset) c->stats = kh_stats{};
  return KH_OK;
}

int kh_device_info(kh_ctx *c, char *name, int name_cap, int *sm_count, uint64_t *hbm_bytes) {
  if (!c) return KH_EINVAL;
  if (name && name_cap > 0) { strncpy(name, c->prop.name, (size_t)name_cap - 1); name[name_cap - 1] = 0; }
  if (sm_count) *sm_count = c->sm_count;
  if (hbm_bytes) *hbm_bytes = (uint64_t)c->prop.totalGlobalMem;
  return KH_OK;
}

}  // extern "C"
