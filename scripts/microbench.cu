// Micro-benchmarks behind DESIGN.md's model of the FP64 cell-step kernel on B200
// (sm_100a): FP64 FMA dependent latency, FP64 issue cadence against occupancy
// and ILP, whether integer work issues in the shadow of FP64 instructions, and
// load latencies.  Build + run (one GPU):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/microbench scripts/microbench.cu
//   build/microbench > gpurun_out/microbench.txt
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

// ILP independent chains of dependent DFMAs per thread
template <int ILP, int MIXI>
__global__ void k_dfma(double* out, long long* cyc, int iters, double a, double b) {
    double x[ILP];
    int n[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = a + i + threadIdx.x; n[i] = threadIdx.x + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                x[i] = fma(x[i], b, a);
                if (MIXI) {
#pragma unroll
                    for (int m = 0; m < MIXI; m++) n[i] = n[i] * 3 + it;  // IMAD, independent of the FP64 chain
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0; int ns = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) { s += x[i]; ns += n[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + ns;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, int MIXI>
void run_dfma(int warps_per_smsp, int sms) {
    int threads = warps_per_smsp * 4 * 32;      // one CTA per SM
    int iters = 2000;
    double* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(double) * threads * sms));
    CK(cudaMalloc(&cyc, sizeof(long long)));
    k_dfma<ILP, MIXI><<<sms, threads>>>(out, cyc, 10, 1.0, 0.999);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_dfma<ILP, MIXI><<<sms, threads>>>(out, cyc, iters, 1.0, 0.999);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long c; CK(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
    double dfma_per_warp = (double)iters * 8 * ILP;
    // cycles per DFMA warp-instruction per SMSP
    double cyc_per = (double)c / (dfma_per_warp * warps_per_smsp);
    printf("dfma ilp=%d mix_imad=%d warps/smsp=%d : %.2f cycles/DFMA/SMSP (%.2f cycles per dependent step), %.1f Gthread-instr/s\n",
           ILP, MIXI, warps_per_smsp, cyc_per, (double)c / (iters * 8.0),
           dfma_per_warp * threads * sms / (ms * 1e-3) / 1e9);
    CK(cudaFree(out)); CK(cudaFree(cyc));
}

// Pointer chase: latency of dependent loads over a footprint
__global__ void k_chase(const unsigned int* p, int steps, unsigned int* out, long long* cyc) {
    unsigned int i = 0;
    for (int k = 0; k < 64; k++) i = p[i];      // warm
    long long t0 = clock64();
    for (int k = 0; k < steps; k++) i = p[i];
    long long t1 = clock64();
    *out = i; *cyc = t1 - t0;
}
__global__ void k_chase_cg(const unsigned int* p, int steps, unsigned int* out, long long* cyc) {
    unsigned int i = 0;
    for (int k = 0; k < 64; k++) i = __ldcg(p + i);
    long long t0 = clock64();
    for (int k = 0; k < steps; k++) i = __ldcg(p + i);
    long long t1 = clock64();
    *out = i; *cyc = t1 - t0;
}

void run_chase(size_t bytes, bool cg, const char* label) {
    size_t n = bytes / 4;
    unsigned int* h = (unsigned int*)malloc(bytes);
    size_t stride = 64;     // 256 B: a new line every step
    for (size_t i = 0; i < n; i++) h[i] = (unsigned int)((i + stride) % n);
    unsigned int *d, *out; long long* cyc;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, 8));
    CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
    int steps = 4096;
    if (cg) k_chase_cg<<<1, 1>>>(d, steps, out, cyc); else k_chase<<<1, 1>>>(d, steps, out, cyc);
    CK(cudaDeviceSynchronize());
    if (cg) k_chase_cg<<<1, 1>>>(d, steps, out, cyc); else k_chase<<<1, 1>>>(d, steps, out, cyc);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    printf("load latency %-28s footprint %8zu KB: %.0f cycles\n", label, bytes / 1024, (double)c / steps);
    CK(cudaFree(d)); CK(cudaFree(out)); CK(cudaFree(cyc)); free(h);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int sms = p.multiProcessorCount;
    run_dfma<1, 0>(1, sms);
    run_dfma<2, 0>(1, sms);
    run_dfma<4, 0>(1, sms);
    run_dfma<8, 0>(1, sms);
    run_dfma<1, 0>(2, sms);
    run_dfma<1, 0>(3, sms);
    run_dfma<1, 0>(4, sms);
    run_dfma<1, 0>(5, sms);
    run_dfma<1, 0>(6, sms);
    run_dfma<1, 0>(8, sms);
    run_dfma<2, 0>(2, sms);
    run_dfma<2, 0>(4, sms);
    run_dfma<4, 0>(4, sms);
    // integer work in the shadow of FP64
    run_dfma<4, 1>(4, sms);
    run_dfma<4, 2>(4, sms);
    run_dfma<2, 1>(4, sms);
    run_dfma<1, 1>(4, sms);
    run_dfma<1, 1>(8, sms);
    run_chase(16 << 10, false, "ld.global (L1 hit)");
    run_chase(4 << 20, false, "ld.global (L2 hit)");
    run_chase(4 << 20, true, "ld.global.cg (L2 hit)");
    run_chase(64 << 20, true, "ld.global.cg (L2, far?)");
    run_chase(1024ull << 20, true, "ld.global.cg (DRAM)");
    return 0;
}
