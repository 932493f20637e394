// kh_peak.cu — integer-pipe micro-benchmarks: the measured denominators of the integer roofline
// (SURVEY §8d asks for an IMAD/IADD3/LOP3 micro-benchmark on the box next to MEASURED_PEAKS.json).
// Each kernel runs 16 independent dependency chains per thread so that the pipes, not latency, bound it.
#include "kh_ctx.cuh"
#include "hash.cuh"

#define PEAK_ITERS 4096
#define PEAK_CHAINS 16

template <int KIND>
__global__ void __launch_bounds__(256) kh_peak_kernel(uint32_t *out, uint32_t seed) {
  uint32_t a[PEAK_CHAINS], b[PEAK_CHAINS];
#pragma unroll
  for (int i = 0; i < PEAK_CHAINS; i++) { a[i] = seed + threadIdx.x * 977u + i; b[i] = seed * 31u + blockIdx.x + i * 7u; }
  const uint32_t x = seed ^ (threadIdx.x * 2654435761u);
  uint32_t row[4][9];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 9; k++) row[i][k] = x + 9 * i + k;
#pragma unroll 1
  for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; i++) {
      if (KIND == 0) {          // IADD3 : a = a + b + it
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b[i]), "r"(it));
      } else if (KIND == 1) {   // LOP3  : a = (a & b) ^ it
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(a[i]) : "r"(b[i]), "r"(it));
      } else if (KIND == 2) {   // SHF   : a = rotl(a, 7) (funnel shift with itself)
        asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
      } else if (KIND == 3) {   // IMAD  : a = a * b + it
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(it));
      } else if (KIND == 4) {   // IMAD.WIDE.U32.X : one row of fe_mul_wide (4 wide multiply-adds with carry) per step
        if (i < 4) kh::kh_mad_row(row[i], a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3], row[i][0] | 1u);
      } else {                  // MIX   : one ALU-pipe op (LOP3) + one FMA-pipe op (IMAD) per chain step
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(a[i]) : "r"(b[i]), "r"(it));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(a[i]), "r"(it));
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < PEAK_CHAINS; i++) r ^= a[i] ^ b[i];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 9; k++) r ^= row[i][k];
  if (r == 0x12345678u) out[0] = r;   // keeps the chains alive
}

template <int KIND>
static double run_peak(kh_ctx *c, uint32_t *d_out, int ops_per_step) {
  const int blocks = c->sm_count * 8;
  kh_peak_kernel<KIND><<<blocks, 256, 0, c->stream>>>(d_out, 12345u);   // warm-up
  double best = 0;
  for (int rep = 0; rep < 3; rep++) {
    kh_time_begin(c);
    kh_peak_kernel<KIND><<<blocks, 256, 0, c->stream>>>(d_out, 12345u + rep);
    const double ms = kh_time_end(c);
    const double ops = (double)blocks * 256.0 * PEAK_ITERS * PEAK_CHAINS * ops_per_step;
    best = std::max(best, ops / (ms * 1e-3));
  }
  return best;
}

extern "C" int kh_int_peak(kh_ctx *c, double out_ops_per_s[6]) {
  if (!c || !out_ops_per_s) return KH_EINVAL;
  cudaSetDevice(c->device);
  uint32_t *d_out = nullptr;
  KH_CUDA(c, cudaMalloc(&d_out, 64));
  out_ops_per_s[0] = run_peak<0>(c, d_out, 1);
  out_ops_per_s[1] = run_peak<1>(c, d_out, 1);
  out_ops_per_s[2] = run_peak<2>(c, d_out, 1);
  out_ops_per_s[3] = run_peak<3>(c, d_out, 1);
  out_ops_per_s[4] = run_peak<4>(c, d_out, 1);   // 16 IMAD.WIDE per loop trip = 16 "chain steps"
  out_ops_per_s[5] = run_peak<5>(c, d_out, 2);
  cudaFree(d_out);
  c->stats.other_launches += 24;
  KH_CUDA(c, cudaGetLastError());
  return KH_OK;
}

// ---- more pipes: what a different multiplier could be built from ("the multiplier") -------------------
// KIND 0 IMAD.WIDE.U32 (64-bit accumulate, no carry chain)   1 IMAD.HI.U32   2 DFMA   3 DADD
//      4 DFMA + IMAD.WIDE per chain step (FP64 pipe and FMA-heavy pipe together)
//      5 fe_mul row (4 x IMAD.WIDE.U32.X) + 4 IADD3 (multiplier + the carry/reduction work that shares issue slots)
//      6 FFMA
template <int KIND>
__global__ void __launch_bounds__(256) kh_pipe_kernel(uint32_t *out, uint32_t seed) {
  uint64_t w[PEAK_CHAINS];
  uint32_t a[PEAK_CHAINS], b[PEAK_CHAINS];
  double d[PEAK_CHAINS];
  float f[PEAK_CHAINS];
  uint32_t f_as_u[PEAK_CHAINS];
#pragma unroll
  for (int i = 0; i < PEAK_CHAINS; i++) {
    f_as_u[i] = seed * 3u + i;
    a[i] = seed + threadIdx.x * 977u + i; b[i] = (seed * 31u + blockIdx.x + i * 7u) | 1u;
    w[i] = ((uint64_t)a[i] << 32) | b[i];
    d[i] = 1.0 + 1e-9 * (double)(a[i] & 1023u);
    f[i] = 1.0f + 1e-6f * (float)(a[i] & 1023u);
  }
  const double dm = 1.0 + 1e-12 * (double)(seed & 7u), da = 1e-15 * (double)(threadIdx.x & 3u);
  const float fm = 1.0f + 1e-7f * (float)(seed & 7u), fa = 1e-9f * (float)(threadIdx.x & 3u);
  uint32_t row[4][9];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 9; k++) row[i][k] = seed + 9 * i + k + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; i++) {
      if (KIND == 0) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));
      } else if (KIND == 1) {
        asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(it));
      } else if (KIND == 2) {
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(da));
      } else if (KIND == 3) {
        asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(da));
      } else if (KIND == 4) {
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(da));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));
      } else if (KIND == 5) {
        if (i < 4) kh::kh_mad_row(row[i], a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3], row[i][0] | 1u);
        if (i >= 4 && i < 8) {
#pragma unroll
          for (int k = 0; k < 4; k++) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(b[4 * (i - 4) + k]) : "r"(a[i]), "r"(it));
        }
      } else if (KIND == 8) {   // IMAD.WIDE.U32 without carries + LOP3: do the FMA-heavy and the ALU pipe overlap when no carry is involved?
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(f_as_u[i]) : "r"(b[i]), "r"(it));
      } else if (KIND == 9) {   // "head" form: a wide multiply-add that produces a carry but takes none, its carry collected by one IADD3.X
        uint32_t lo = (uint32_t)w[i], hi = (uint32_t)(w[i] >> 32);
        asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo), "+r"(hi), "+r"(f_as_u[i]) : "r"(a[i]), "r"(b[i]));
        w[i] = ((uint64_t)hi << 32) | lo;
      } else if (KIND == 10) {  // carry-free wide multiply-add + one plain IADD3 (the comparison for KIND 9)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(f_as_u[i]) : "r"(b[i]), "r"(it));
      } else if (KIND == 11) {  // two-link chain: head + one IMAD.WIDE.U32.X + the IADD3.X that collects the carry (2 wide ops + 1 ALU op)
        if (i < 8) {
          uint32_t lo = (uint32_t)w[2 * i], hi = (uint32_t)(w[2 * i] >> 32), lo2 = (uint32_t)w[2 * i + 1], hi2 = (uint32_t)(w[2 * i + 1] >> 32);
          asm volatile("mad.lo.cc.u32 %0, %5, %6, %0;\n\tmadc.hi.cc.u32 %1, %5, %6, %1;\n\tmadc.lo.cc.u32 %2, %5, %7, %2;\n\tmadc.hi.cc.u32 %3, %5, %7, %3;\n\taddc.u32 %4, %4, 0;"
                       : "+r"(lo), "+r"(hi), "+r"(lo2), "+r"(hi2), "+r"(f_as_u[i]) : "r"(a[i]), "r"(b[i]), "r"(b[i + 8]));
          w[2 * i] = ((uint64_t)hi << 32) | lo; w[2 * i + 1] = ((uint64_t)hi2 << 32) | lo2;
        }
      } else if (KIND == 7) {
        // the x-only walk's own mix: per point 224 wide multiply-adds in carry chains and ~470 ALU-pipe ops (ncu), i.e.
        // 16 IMAD.WIDE.U32.X + 34 ALU ops per trip here (half LOP3, half carry-free IADD3)
        if (i < 4) kh::kh_mad_row(row[i], a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3], row[i][0] | 1u);
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(b[i]) : "r"(a[i]), "r"(it));
        asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b[i]), "r"(it));
      } else {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fm), "f"(fa));
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < PEAK_CHAINS; i++) r ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)__double2loint(d[i]) ^ __float_as_uint(f[i]) ^ f_as_u[i];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 9; k++) r ^= row[i][k];
  if (r == 0x12345678u) out[0] = r;
}

template <int KIND>
static double run_pipe(kh_ctx *c, uint32_t *d_out, int ops_per_step) {
  const int blocks = c->sm_count * 8;
  kh_pipe_kernel<KIND><<<blocks, 256, 0, c->stream>>>(d_out, 12345u);
  double best = 0;
  for (int rep = 0; rep < 3; rep++) {
    kh_time_begin(c);
    kh_pipe_kernel<KIND><<<blocks, 256, 0, c->stream>>>(d_out, 12345u + rep);
    const double ms = kh_time_end(c);
    const double ops = (double)blocks * 256.0 * PEAK_ITERS * PEAK_CHAINS * ops_per_step;
    best = std::max(best, ops / (ms * 1e-3));
  }
  return best;
}

extern "C" int kh_pipe_peak(kh_ctx *c, double out[16]) {
  if (!c || !out) return KH_EINVAL;
  cudaSetDevice(c->device);
  uint32_t *d_out = nullptr;
  KH_CUDA(c, cudaMalloc(&d_out, 64));
  out[0] = run_pipe<0>(c, d_out, 1);
  out[1] = run_pipe<1>(c, d_out, 1);
  out[2] = run_pipe<2>(c, d_out, 1);
  out[3] = run_pipe<3>(c, d_out, 1);
  out[4] = run_pipe<4>(c, d_out, 2);
  out[5] = run_pipe<5>(c, d_out, 2);   // 16 IMAD.WIDE.X + 16 IADD3 per loop trip
  out[6] = run_pipe<6>(c, d_out, 1);
  out[7] = run_pipe<7>(c, d_out, 1);
  out[8] = run_pipe<8>(c, d_out, 2);   // IMAD.WIDE.U32 (no carry) + LOP3 together, total ops/s
  out[9] = run_pipe<9>(c, d_out, 2);    // head-form IMAD.WIDE (carry out only) + IADD3.X, total ops/s
  out[10] = run_pipe<10>(c, d_out, 2);  // carry-free IMAD.WIDE + IADD3, total ops/s
  out[11] = run_pipe<11>(c, d_out, 1) * (24.0 / 16.0);   // 8 x (head + .X + IADD3.X) per trip = 24 ops, total ops/s
  for (int i = 12; i < 16; i++) out[i] = 0;   // trips x 16 per second = IMAD.WIDE.U32.X per second inside the walk's mix (16 per trip)
  cudaFree(d_out);
  c->stats.other_launches += 48;
  KH_CUDA(c, cudaGetLastError());
  return KH_OK;
}

// ---- hash micro-benchmarks: SHA-256 compressions / RIPEMD-160 blocks per second in isolation -----------
template <int WHICH>
__global__ void __launch_bounds__(256) kh_hash_bench(uint32_t *out, uint32_t seed, int iters) {
  uint32_t st[8], w[16], h[5];
#pragma unroll
  for (int i = 0; i < 8; i++) st[i] = seed + threadIdx.x * 31u + i;
#pragma unroll
  for (int i = 0; i < 5; i++) h[i] = 0;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (WHICH == 0) {
#pragma unroll
      for (int i = 0; i < 16; i++) w[i] = st[i & 7] ^ (uint32_t)(it + i);
      kh::sha256_compress(st, w);
    } else {
      kh::ripemd160_of_sha(h, st);
#pragma unroll
      for (int i = 0; i < 5; i++) st[i] ^= h[i];
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= st[i];
  if (r == 0x12345678u) out[0] = r;
}

extern "C" int kh_hash_peak(kh_ctx *c, int blocks_per_sm, double out_blocks_per_s[2]) {
  if (!c || !out_blocks_per_s || blocks_per_sm < 1) return KH_EINVAL;
  cudaSetDevice(c->device);
  uint32_t *d_out = nullptr;
  KH_CUDA(c, cudaMalloc(&d_out, 64));
  const int blocks = c->sm_count * blocks_per_sm, iters = 2000;
  for (int which = 0; which < 2; which++) {
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
      kh_time_begin(c);
      if (which == 0) kh_hash_bench<0><<<blocks, 256, 0, c->stream>>>(d_out, 7u + rep, iters);
      else kh_hash_bench<1><<<blocks, 256, 0, c->stream>>>(d_out, 7u + rep, iters);
      const double ms = kh_time_end(c);
      if (rep) best = std::max(best, (double)blocks * 256.0 * iters / (ms * 1e-3));
    }
    out_blocks_per_s[which] = best;
  }
  cudaFree(d_out);
  KH_CUDA(c, cudaGetLastError());
  return KH_OK;
}
