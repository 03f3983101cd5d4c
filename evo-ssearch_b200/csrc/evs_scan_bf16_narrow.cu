// evs_scan_bf16_narrow.cu -- instantiations of the GEMV scan kernels (evs_scan.cuh) for __nv_bfloat16 rows, 1, 2 16-byte vectors per lane.
#include "evs_scan_launch.cuh"

namespace evs {

cudaError_t launch_scan_bf16_narrow(const ScanArgs& a, ScanPlan* plan, cudaStream_t st) {
    switch (plan->nv) {
        case 1: return launch_scan_nq<__nv_bfloat16, 1>(a, plan, st);
        case 2: return launch_scan_nq<__nv_bfloat16, 2>(a, plan, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace evs
