// evs_scan_f32_narrow.cu -- instantiations of the GEMV scan kernels (evs_scan.cuh) for float rows, 1, 2, 3 16-byte vectors per lane.
#include "evs_scan_launch.cuh"

namespace evs {

cudaError_t launch_scan_f32_narrow(const ScanArgs& a, ScanPlan* plan, cudaStream_t st) {
    switch (plan->nv) {
        case 1: return launch_scan_nq<float, 1>(a, plan, st);
        case 2: return launch_scan_nq<float, 2>(a, plan, st);
        case 3: return launch_scan_nq<float, 3>(a, plan, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace evs
