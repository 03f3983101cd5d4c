// evs_scan_f32_512.cu -- instantiations of the GEMV scan kernels (evs_scan.cuh) for float rows, 4 16-byte vectors per lane.
#include "evs_scan_launch.cuh"

namespace evs {

cudaError_t launch_scan_f32_512(const ScanArgs& a, ScanPlan* plan, cudaStream_t st) {
    switch (plan->nv) {
        case 4: return launch_scan_nq<float, 4>(a, plan, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace evs
