// evs_scan_generic.cu -- the any-d fallback of the GEMV scan (scalar coalesced loads).
#include "evs_scan_launch.cuh"

namespace evs {

cudaError_t launch_scan_generic_any(const ScanArgs& a, ScanPlan* plan, cudaStream_t st) {
    return a.is_bf16 ? launch_scan_generic<__nv_bfloat16>(a, plan, st) : launch_scan_generic<float>(a, plan, st);
}

}  // namespace evs
