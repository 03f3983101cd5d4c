// micro-probe: per-warp issue cost of fp64 ops on this GPU (one block, W warps, independent chains)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, float* fin, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    float f0 = fin[threadIdx.x], f1 = f0 + 1.f, f2 = f0 + 2.f, f3 = f0 + 3.f;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (OP == 0) { a0 = fma(a0, 1.0000001, 0.5); a1 = fma(a1, 1.0000001, 0.5); a2 = fma(a2, 1.0000001, 0.5); a3 = fma(a3, 1.0000001, 0.5); }
        if (OP == 1) { a0 += (double)f0; a1 += (double)f1; a2 += (double)f2; a3 += (double)f3; f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f; }
        if (OP == 2) { f0 = fmaf(f0, 1.0000001f, 0.5f); f1 = fmaf(f1, 1.0000001f, 0.5f); f2 = fmaf(f2, 1.0000001f, 0.5f); f3 = fmaf(f3, 1.0000001f, 0.5f); }
        if (OP == 3) { a0 = (a0 > a1) ? a0 + 1.0 : a0; a1 = (a1 > a2) ? a1 : a1 + 1.0; a2 = (a2 > a3) ? a2 + 1.0 : a2; a3 = (a3 > a0) ? a3 : a3 + 1.0; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a0 + a1 + a2 + a3 + f0 + f1 + f2 + f3;
    if (threadIdx.x == 0) out[1024] = (double)(t1 - t0);
}
int main() {
    double* out; float* fin; cudaMalloc(&out, 1025 * 8); cudaMalloc(&fin, 4096); cudaMemset(fin, 0, 4096);
    const char* names[] = {"DFMA x4", "F2F.F64.F32+DADD x4 (+FADD x4)", "FFMA x4", "DSETP+DADD x4"};
    for (int w : {1, 4, 32}) for (int op = 0; op < 4; op++) {
        int iters = 2000; double cyc;
        if (op == 0) k<0><<<1, 32 * w>>>(out, fin, iters); if (op == 1) k<1><<<1, 32 * w>>>(out, fin, iters);
        if (op == 2) k<2><<<1, 32 * w>>>(out, fin, iters); if (op == 3) k<3><<<1, 32 * w>>>(out, fin, iters);
        cudaDeviceSynchronize(); cudaMemcpy(&cyc, out + 1024, 8, cudaMemcpyDeviceToHost);
        printf("warps=%2d %-34s %.1f cycles per loop iteration (4 chains)\n", w, names[op], cyc / iters);
    }
    return 0;
}
