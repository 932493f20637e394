// scan_kernel.cuh — the scan kernel template, shared by kh_scan.cu (table searches) and kh_vanity.cu (-m vanity
// instantiations, a separate translation unit so that the two compile in parallel).
#pragma once
#include "kh_ctx.cuh"

#ifndef KH_BLOCK
#define KH_BLOCK 256
#endif
// two CTAs of 256 threads per SM (128 registers per thread)
#ifndef KH_SCAN_MINBLOCKS
#define KH_SCAN_MINBLOCKS (512 / KH_BLOCK)
#endif
// the x-only walk has no hash state to keep alive: its CTA shape is a knob of its own (A/B: registers vs resident warps)
#ifndef KH_XPOINT_BLOCK
#define KH_XPOINT_BLOCK KH_BLOCK
#endif
#ifndef KH_XPOINT_MINBLOCKS
#define KH_XPOINT_MINBLOCKS (512 / KH_XPOINT_BLOCK)
#endif
template <int KIND, bool ENDO>
struct ScanShape {
  static constexpr bool XP = (KIND == kh::KH_SCAN_XPOINT) && !ENDO;
  static constexpr int BLOCK = XP ? KH_XPOINT_BLOCK : KH_BLOCK;
  static constexpr int MINBLOCKS = XP ? KH_XPOINT_MINBLOCKS : KH_SCAN_MINBLOCKS;
};

__device__ __forceinline__ void kh_stage_table(uint32_t *smem, const uint32_t *gtab) {
  const uint4 *src = reinterpret_cast<const uint4 *>(gtab);
  uint4 *dst = reinterpret_cast<uint4 *>(smem);
  for (int i = threadIdx.x; i < KH_TAB_WORDS / 4; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// dynamic shared memory of a scan kernel: the G-multiple table, plus the SHA-256 schedule table of the uncompressed key's second
// block for the kinds that hash uncompressed keys (hash.cuh KH_SHA_UNC2_TAB): 32.8 KB + 69.6 KB, two CTAs per SM
template <int KIND, bool ENDO, bool VANITY = false>
constexpr size_t kh_scan_smem_bytes() {
  return (size_t)(KH_TAB_WORDS + (kh::ScanEmit<KIND, ENDO, VANITY>::SHA2TAB ? KH_SHA2TAB_WORDS : 0)) * sizeof(uint32_t);
}

template <int KIND, bool ENDO, bool VANITY = false>
__global__ void __launch_bounds__((ScanShape<KIND, ENDO>::BLOCK), (ScanShape<KIND, ENDO>::MINBLOCKS)) kh_scan_kernel(kh::WalkParams wp, kh::ScanTargets tg) {
  extern __shared__ __align__(16) uint32_t kh_smem_tab[];
  const uint32_t *sha2 = nullptr;
  if (kh::ScanEmit<KIND, ENDO, VANITY>::SHA2TAB) {
    const uint4 *src = reinterpret_cast<const uint4 *>(tg.sha2);
    uint4 *dst = reinterpret_cast<uint4 *>(kh_smem_tab + KH_TAB_WORDS);
    for (int i = threadIdx.x; i < KH_SHA2TAB_WORDS / 4; i += blockDim.x) dst[i] = src[i];
    sha2 = kh_smem_tab + KH_TAB_WORDS;
  }
  kh_stage_table(kh_smem_tab, wp.gtab);
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= wp.T) return;
  kh::ScanEmit<KIND, ENDO, VANITY> emit(tg, sha2);
  kh::walk_batches(wp, kh_smem_tab, t, emit);
}

// launches one scan kernel instantiation (opts in to more than 48 KB of dynamic shared memory where the kind needs it)
template <int KIND, bool ENDO, bool VANITY>
static cudaError_t kh_launch_scan_kernel(kh_ctx *c, const kh::WalkParams &wp, const kh::ScanTargets &tg) {
  constexpr size_t smem = kh_scan_smem_bytes<KIND, ENDO, VANITY>();
  if (smem > 48 * 1024) {                            // (per device and per instantiation; a few microseconds in front of a launch)
    cudaError_t e = cudaFuncSetAttribute(kh_scan_kernel<KIND, ENDO, VANITY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  constexpr int BLOCK = ScanShape<KIND, ENDO>::BLOCK;
  kh_scan_kernel<KIND, ENDO, VANITY><<<(unsigned)(wp.T / BLOCK), BLOCK, smem, c->stream>>>(wp, tg);
  return cudaGetLastError();
}

// kh_vanity.cu: launches kh_scan_kernel<KIND, endo, true> for KIND = COMP / UNCOMP / BOTH
cudaError_t kh_launch_vanity(kh_ctx *c, int kind, const kh::WalkParams &wp, const kh::ScanTargets &tg);
