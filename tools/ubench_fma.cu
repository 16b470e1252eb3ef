// ubench_fma.cu -- what the FP32 pipes of one B200 SM deliver for the instruction forms the DWT tap loops can use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_fma tools/ubench_fma.cu
// Prints, per variant: FMA/clk/SM (clock64-based, per-SM cycle counts) and TFMA/s chip-wide (CUDA events).
#include <cstdio>
#include <cuda_runtime.h>

struct Taps {
    float k[32];
};

__device__ __forceinline__ unsigned long long pk(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void fma2(unsigned long long& acc, unsigned long long a, unsigned long long b)
{
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

// V0: scalar FFMA, multiplier from the constant bank, NCH independent chains
// V1: FFMA2  acc(c0,c1) += x(c0,c1) * k.F32         (column-pair form)
// V2: FFMA2  acc(lo,hi) += x.F32 * (kl,kh)          (filter-pair form)
// V3: scalar FFMA with three register operands
// V4: V2 plus one LDS.128 per 7 FFMA2 (smem co-issue)
// V5: V0 plus one LDS.128 per 14 FFMA
// V6: V2 with one ALU instruction (XOR) per FFMA2 -- does an FFMA2 hold the scheduler's issue port for two cycles?
// V7: scalar FFMA (constant-bank multiplier) with one XOR per TWO FFMA: the same FMA and XOR counts as V6
// V8: V2 with one XOR per TWO FFMA2 (the ratio of the DWT kernels: ~112 FFMA2 to ~100 other instructions is 1:1,
//     V6; with the bookkeeping halved it would be V8)
template <int V, int NCH>
__global__ void __launch_bounds__(256) kern(const __grid_constant__ Taps t, float* out, int iters, long long* cyc)
{
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    float x[NCH], a[NCH], b[NCH];
    unsigned long long acc2[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        x[i] = threadIdx.x * 0.001f + i;
        a[i] = 0.f;
        b[i] = out[i];
        acc2[i] = 0ull;
    }
    float4 ld = make_float4(0, 0, 0, 0);
    unsigned z[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) z[i] = threadIdx.x + i;
    long long c0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 14; j++) {
            if (V == 0 || V == 5) {
#pragma unroll
                for (int i = 0; i < NCH; i++) a[i] = fmaf(x[i], t.k[j], a[i]);
                if (V == 5 && j == 0) {
                    float4 q = sm[(threadIdx.x + it) & 1023];
                    ld.x += q.x; ld.y += q.y; ld.z += q.z; ld.w += q.w;
                }
            } else if (V == 1) {
#pragma unroll
                for (int i = 0; i < NCH; i++) fma2(acc2[i], pk(x[i], b[i]), pk(t.k[j], t.k[j]));
            } else if (V == 2 || V == 4) {
#pragma unroll
                for (int i = 0; i < NCH; i++) fma2(acc2[i], pk(x[i], x[i]), pk(t.k[j], t.k[j + 16]));
                if (V == 4 && (j == 0 || j == 7)) {
                    float4 q = sm[(threadIdx.x + it + j) & 1023];
                    ld.x += q.x; ld.y += q.y; ld.z += q.z; ld.w += q.w;
                }
            } else if (V == 6 || V == 8) {
#pragma unroll
                for (int i = 0; i < NCH; i++) {
                    fma2(acc2[i], pk(x[i], x[i]), pk(t.k[j], t.k[j + 16]));
                    if (V == 6 || (i & 1)) asm volatile("add.s32 %0, %0, %1;" : "+r"(z[i]) : "r"(z[(i + 3) % NCH]));
                }
            } else if (V == 7) {
#pragma unroll
                for (int i = 0; i < NCH; i++) {
                    a[i] = fmaf(x[i], t.k[j], a[i]);
                    b[i] = fmaf(x[i], t.k[j + 16], b[i]);
                    asm volatile("add.s32 %0, %0, %1;" : "+r"(z[i]) : "r"(z[(i + 3) % NCH]));
                }
            } else if (V == 3) {
#pragma unroll
                for (int i = 0; i < NCH; i++) a[i] = fmaf(x[i], b[i], a[i]);
            }
        }
    }
    long long c1 = clock64();
    float s = ld.x + ld.y + ld.z + ld.w;
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        float2 f = *reinterpret_cast<float2*>(&acc2[i]);
        s += a[i] + f.x + f.y + (float)z[i] + ((V == 7) ? b[i] : 0.f);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <int V, int NCH>
void run(const char* name, int ctas_per_sm, int threads)
{
    Taps t;
    for (int i = 0; i < 32; i++) t.k[i] = 1e-3f * (i + 1);
    const int nsm = 148, grid = nsm * ctas_per_sm, iters = 4000;
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * grid * threads);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    kern<V, NCH><<<grid, threads>>>(t, out, 10, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<V, NCH><<<grid, threads>>>(t, out, iters, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[grid];
    cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < grid; i++) mean += h[i];
    mean /= grid;
    const double fma_per_thread = (double)iters * 14 * NCH * ((V == 1 || V == 2 || V == 4 || V == 6 || V == 7 || V == 8) ? 2 : 1);
    const double per_sm = fma_per_thread * threads * ctas_per_sm;
    printf("%-34s ctas/sm=%d thr=%d chains=%d : %7.1f FMA/clk/SM   %6.2f TFMA/s  (%.3f ms, %.0f cyc)  err=%s\n", name,
           ctas_per_sm, threads, NCH, per_sm / mean, per_sm * nsm / (ms * 1e-3) / 1e12, ms, mean,
           cudaGetErrorString(cudaGetLastError()));
    delete[] h;
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    run<0, 8>("FFMA R,R,c[],R", 2, 256);
    run<0, 8>("FFMA R,R,c[],R", 4, 256);
    run<3, 8>("FFMA R,R,R,R", 2, 256);
    run<1, 8>("FFMA2 pair x pair, k.F32", 2, 256);
    run<1, 8>("FFMA2 pair x pair, k.F32", 1, 128);
    run<2, 8>("FFMA2 x.F32 * (kl,kh)", 2, 256);
    run<2, 4>("FFMA2 x.F32 * (kl,kh)", 2, 256);
    run<2, 2>("FFMA2 x.F32 * (kl,kh)", 2, 256);
    run<2, 2>("FFMA2 x.F32 * (kl,kh)", 1, 128);
    run<2, 8>("FFMA2 x.F32 * (kl,kh)", 1, 128);
    run<4, 8>("FFMA2 + LDS.128 per 7", 2, 256);
    run<5, 8>("FFMA + LDS.128 per 14", 2, 256);
    run<6, 8>("FFMA2 + XOR 1:1", 2, 256);
    run<6, 8>("FFMA2 + XOR 1:1", 1, 128);
    run<6, 8>("FFMA2 + XOR 1:1", 1, 256);
    run<8, 8>("FFMA2 + XOR 2:1", 2, 256);
    run<8, 8>("FFMA2 + XOR 2:1", 1, 128);
    run<7, 8>("2 FFMA + XOR (same work as 1:1)", 2, 256);
    run<7, 8>("2 FFMA + XOR (same work as 1:1)", 1, 128);
    return 0;
}
