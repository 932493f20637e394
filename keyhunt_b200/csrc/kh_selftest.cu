// kh_selftest.cu — device-side known-answer entry for the field layer (fe.cuh).
//
// The device bodies of fe.cuh are PTX carry chains (mad.lo.cc / madc.hi.cc / addc / subc) that share no code with the
// portable bodies the host test build compiles, so the reference-generated vectors for Int::ModMulK1 / ModSquareK1 /
// ModInv / ModAdd / ModSub / ModNeg (secp256k1/IntMod.cpp:855, :977, :382, :41, :72, :102; tests/golden/primitives.json)
// and forced edge operands (second-fold carry of fe_reduce_wide, the take path of fe_final_reduce, borrow with b = 0,
// inv(0) = 0) must be run THROUGH THE GPU.  One thread per vector; every op is the exact inline function the walk uses,
// plus the out-of-line multiplier copy the hash kernels call.
#include "kh_ctx.cuh"

using namespace kh;

__global__ void __launch_bounds__(128) kh_selftest_fe_kernel(int op, const uint32_t *a_in, const uint32_t *b_in, uint64_t n, uint32_t *out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe a, b, r;
#pragma unroll
  for (int l = 0; l < 8; l++) { a.v[l] = a_in[8 * i + l]; b.v[l] = b_in[8 * i + l]; }
  fe_set_zero(r);
  uint32_t w[16];
  switch (op) {
    case KH_FE_MUL: fe_mul(r, a, b); break;
    case KH_FE_SQR: fe_sqr(r, a); break;
    case KH_FE_INV: fe_inv(r, a); break;
    case KH_FE_ADD: fe_add(r, a, b); break;
    case KH_FE_SUB: fe_sub(r, a, b); break;
    case KH_FE_NEG: fe_neg(r, a); break;
    case KH_FE_MUL_OUTLINE: r = fe_mul_ol(a, b); break;
    case KH_FE_MULWIDE_LO: case KH_FE_MULWIDE_HI:
      fe_mul_wide(w, a, b);
#pragma unroll
      for (int l = 0; l < 8; l++) r.v[l] = (op == KH_FE_MULWIDE_LO) ? w[l] : w[8 + l];
      break;
    case KH_FE_SQRWIDE_LO: case KH_FE_SQRWIDE_HI:
      fe_sqr_wide(w, a);
#pragma unroll
      for (int l = 0; l < 8; l++) r.v[l] = (op == KH_FE_SQRWIDE_LO) ? w[l] : w[8 + l];
      break;
    case KH_FE_REDUCE_WIDE:          // (a * 2^256 + b) mod P for ANY 512-bit value: forces the rare folds directly
#pragma unroll
      for (int l = 0; l < 8; l++) { w[l] = b.v[l]; w[8 + l] = a.v[l]; }
      fe_reduce_wide(r, w);
      break;
    // the other form of the multiplier's final reduction (fe.cuh KH_RARE_REDUCE): the kernels use both (emit.cuh RARE_REDUCE)
    case KH_FE_MUL_ALT: fe_mul<!KH_RARE_REDUCE>(r, a, b); break;
    case KH_FE_SQR_ALT: fe_sqr<!KH_RARE_REDUCE>(r, a); break;
    case KH_FE_INV_ALT: fe_inv<!KH_RARE_REDUCE>(r, a); break;
    case KH_FE_INV_SQR: fe_inv<KH_RARE_REDUCE, true>(r, a); break;      // the x-only walks' inversion: dedicated out-of-line squaring
    case KH_FE_MUL_OUTLINE_ALT: r = fe_mul_ol<!KH_RARE_REDUCE>(a, b); break;
    case KH_FE_REDUCE_WIDE_ALT:
#pragma unroll
      for (int l = 0; l < 8; l++) { w[l] = b.v[l]; w[8 + l] = a.v[l]; }
      fe_reduce_wide<!KH_RARE_REDUCE>(r, w);
      break;
    default: break;
  }
#pragma unroll
  for (int l = 0; l < 8; l++) out[8 * i + l] = r.v[l];
}

extern "C" int kh_selftest_fe(kh_ctx *c, int op, const uint8_t *a_be, const uint8_t *b_be, uint64_t n, uint8_t *out_be) {
  if (!c || !a_be || !b_be || !out_be) return KH_EINVAL;
  if (op < KH_FE_MUL || op > KH_FE_INV_SQR) return kh_fail(c, KH_EINVAL, "unknown field op %d", op);
  if (n == 0) return KH_OK;
  cudaSetDevice(c->device);
  std::vector<uint32_t> ha(8 * n), hb(8 * n), ho(8 * n);
  for (uint64_t i = 0; i < n; i++) {
    u256 x, y;
    u256_from_be(x, a_be + 32 * i);
    u256_from_be(y, b_be + 32 * i);
    for (int l = 0; l < 8; l++) { ha[8 * i + l] = x.v[l]; hb[8 * i + l] = y.v[l]; }
  }
  uint32_t *d = nullptr;
  const size_t bytes = 8 * n * sizeof(uint32_t);
  KH_CUDA(c, cudaMalloc(&d, 3 * bytes));
  cudaMemcpyAsync(d, ha.data(), bytes, cudaMemcpyHostToDevice, c->stream);
  cudaMemcpyAsync(d + 8 * n, hb.data(), bytes, cudaMemcpyHostToDevice, c->stream);
  kh_selftest_fe_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(op, d, d + 8 * n, n, d + 16 * n);
  cudaMemcpyAsync(ho.data(), d + 16 * n, bytes, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(d);
  c->stats.other_launches += 1;
  if (e != cudaSuccess) return kh_fail(c, KH_ENODEV, "selftest kernel: %s", cudaGetErrorString(e));
  for (uint64_t i = 0; i < n; i++) {
    u256 r;
    for (int l = 0; l < 8; l++) r.v[l] = ho[8 * i + l];
    u256_to_be(out_be + 32 * i, r);
  }
  return KH_OK;
}
