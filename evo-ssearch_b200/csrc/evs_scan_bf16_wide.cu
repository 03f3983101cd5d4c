// evs_scan_bf16_wide.cu -- instantiations of the GEMV scan kernels (evs_scan.cuh) for __nv_bfloat16 rows, 3, 4 16-byte vectors per lane.
#include "evs_scan_launch.cuh"

namespace evs {

cudaError_t launch_scan_bf16_wide(const ScanArgs& a, ScanPlan* plan, cudaStream_t st) {
    switch (plan->nv) {
        case 3: return launch_scan_nq<__nv_bfloat16, 3>(a, plan, st);
        case 4: return launch_scan_nq<__nv_bfloat16, 4>(a, plan, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace evs
