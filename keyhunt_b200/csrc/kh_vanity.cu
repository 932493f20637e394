// kh_vanity.cu — the -m vanity instantiations of the scan kernel (thread_process_vanity keyhunt.cpp:3867): the same walk and
// hash jobs as -m rmd160, with the interval match of emit.cuh (vanity_match) as the membership test.  A separate
// translation unit so that it compiles next to kh_scan.cu.
#include "scan_kernel.cuh"

using namespace kh;

template <int KIND>
static cudaError_t launch_vanity(kh_ctx *c, const WalkParams &wp, const ScanTargets &tg) {
  if (c->endomorphism) return kh_launch_scan_kernel<KIND, true, true>(c, wp, tg);
  return kh_launch_scan_kernel<KIND, false, true>(c, wp, tg);
}

cudaError_t kh_launch_vanity(kh_ctx *c, int kind, const WalkParams &wp, const ScanTargets &tg) {
  switch (kind) {
    case KH_SCAN_COMP: return launch_vanity<KH_SCAN_COMP>(c, wp, tg);
    case KH_SCAN_UNCOMP: return launch_vanity<KH_SCAN_UNCOMP>(c, wp, tg);
    case KH_SCAN_BOTH: return launch_vanity<KH_SCAN_BOTH>(c, wp, tg);
    default: return cudaErrorInvalidValue;
  }
}
