// evs_scan_f32_wide.cu -- instantiations of the GEMV scan kernels (evs_scan.cuh) for float rows, 6, 8 16-byte vectors per lane.
#include "evs_scan_launch.cuh"

namespace evs {

cudaError_t launch_scan_f32_wide(const ScanArgs& a, ScanPlan* plan, cudaStream_t st) {
    switch (plan->nv) {
        case 6: return launch_scan_nq<float, 6>(a, plan, st);
        case 8: return launch_scan_nq<float, 8>(a, plan, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace evs
