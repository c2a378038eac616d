// Micro-probe: issue rate of FFMA vs FFMA2 (packed fp32x2) on sm_100a, for the operand patterns the depthwise kernel uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/fma_probe tools/probes/fma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int MODE, int NACC>
__global__ void probe(float* out, long long* cyc, int iters, float seed) {
    // MODE 0: scalar FFMA acc[j] = v*w[j%7] + acc[j];  1: FFMA2 same pattern;  2: FFMA2 acc[j] = acc[j]*w + v (acc as multiplicand)
    // 3: FFMA2 with all-distinct operands per instruction
    uint64_t acc[NACC], w[8], v[4];
    float facc[NACC], fw[8], fv[4];
    for (int j = 0; j < NACC; ++j) { facc[j] = seed * j; acc[j] = (uint64_t)__float_as_uint(seed * j) * 0x100000001ull; }
    for (int j = 0; j < 8; ++j) { fw[j] = seed + j; w[j] = (uint64_t)__float_as_uint(seed + j) * 0x100000001ull; }
    for (int j = 0; j < 4; ++j) { fv[j] = seed - j; v[j] = (uint64_t)__float_as_uint(seed - j) * 0x100000001ull; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) {
                if (MODE == 0) facc[j] = ffma1(fv[r], fw[(j + r) % 7], facc[j]);
                if (MODE == 1) acc[j] = ffma2(v[r], w[(j + r) % 7], acc[j]);
                if (MODE == 2) acc[j] = ffma2(acc[j], w[(j + r) % 7], v[r]);
                if (MODE == 3) acc[j] = ffma2(v[(j + r) % 4], w[(j + r) % 7], acc[j]);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < NACC; ++j) s += (MODE == 0) ? facc[j] : __uint_as_float((uint32_t)acc[j]) + __uint_as_float((uint32_t)(acc[j] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int NACC>
void run(const char* name, int warps_per_sm) {
    float* out; long long* cyc;
    const int blocks = 148, thr = warps_per_sm * 32, iters = 2000;
    cudaMalloc(&out, blocks * thr * 4); cudaMalloc(&cyc, blocks * 8);
    probe<MODE, NACC><<<blocks, thr>>>(out, cyc, iters, 1.0f);
    cudaDeviceSynchronize();
    probe<MODE, NACC><<<blocks, thr>>>(out, cyc, iters, 1.0f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    const double instr_per_smsp = (double)iters * 4 * NACC * warps_per_sm / 4.0;
    const double fma_per_clk_sm = (double)iters * 4 * NACC * warps_per_sm * 32 * (MODE == 0 ? 1 : 2) / avg;
    printf("%-34s warps/SM=%2d NACC=%2d: %.2f cycles per warp-instr per SMSP, %.1f FMA/clk/SM (err=%s)\n", name, warps_per_sm, NACC, avg / instr_per_smsp,
           fma_per_clk_sm, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16}) {
        if (w == 4) { run<0, 28>("FFMA  acc=v*w+acc", 4); run<1, 28>("FFMA2 acc=v*w+acc", 4); run<2, 28>("FFMA2 acc=acc*w+v", 4); run<3, 28>("FFMA2 distinct v,w", 4); }
        if (w == 8) { run<0, 28>("FFMA  acc=v*w+acc", 8); run<1, 28>("FFMA2 acc=v*w+acc", 8); run<2, 28>("FFMA2 acc=acc*w+v", 8); run<3, 28>("FFMA2 distinct v,w", 8); run<1, 8>("FFMA2 acc=v*w+acc", 8); }
        if (w == 16) { run<0, 28>("FFMA  acc=v*w+acc", 16); run<1, 28>("FFMA2 acc=v*w+acc", 16); }
    }
    return 0;
}
