// microbench.cu -- issue-rate probes for the SPH inner loop on sm_100a (B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
// Prints lane-ops per clock per SM for scalar vs packed (f32x2) FP32 ops, FMNMX/FSETP (ALU
// pipe), a mixed FMA+ALU stream and broadcast LDS.128.  Results are recorded in DESIGN.md.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;  // independent chains per thread

#define DEF_KERNEL(NAME, DECL, BODY, FINAL)                                            \
    __global__ void __launch_bounds__(1024) NAME(float* out, long long* cycles, float a, float b) { \
        DECL;                                                                          \
        __syncthreads();                                                               \
        long long t0 = clock64();                                                      \
        _Pragma("unroll 4")                                                            \
        for (int it = 0; it < ITERS; it++) {                                           \
            BODY;                                                                      \
        }                                                                              \
        long long t1 = clock64();                                                      \
        __syncthreads();                                                               \
        FINAL;                                                                         \
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;                            \
    }

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

// scalar fma
#define D_F float r0 = a, r1 = a + 1, r2 = a + 2, r3 = a + 3, r4 = a + 4, r5 = a + 5, r6 = a + 6, r7 = a + 7
#define FIN_F out[blockIdx.x * blockDim.x + threadIdx.x] = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7
#define OP_FMA(i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(r##i) : "f"(a), "f"(b));
#define OP_ADD(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r##i) : "f"(b));
#define OP_MUL(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r##i) : "f"(b));
#define OP_MIN(i) asm volatile("min.f32 %0, %0, %1;" : "+f"(r##i) : "f"(b));
#define OP_SETP(i) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(r##i) : "f"(b), "f"(a));
#define OP_MIX(i) asm volatile("fma.rn.f32 %0, %1, %2, %0; min.f32 %0, %0, %1;" : "+f"(r##i) : "f"(a), "f"(b));
DEF_KERNEL(k_fma, D_F, REP8(OP_FMA), FIN_F)
DEF_KERNEL(k_add, D_F, REP8(OP_ADD), FIN_F)
DEF_KERNEL(k_mul, D_F, REP8(OP_MUL), FIN_F)
DEF_KERNEL(k_min, D_F, REP8(OP_MIN), FIN_F)
DEF_KERNEL(k_setp, D_F, REP8(OP_SETP), FIN_F)
DEF_KERNEL(k_mix, D_F, REP8(OP_MIX), FIN_F)

// packed f32x2
#define D_P unsigned long long r0, r1, r2, r3, r4, r5, r6, r7, pa, pb; \
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a));  \
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));  \
    r0 = r1 = r2 = r3 = r4 = r5 = r6 = r7 = pa
#define FIN_P out[blockIdx.x * blockDim.x + threadIdx.x] = (float)((r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7) & 0xffff)
#define OP_FMA2(i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(r##i) : "l"(pa), "l"(pb));
#define OP_ADD2(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r##i) : "l"(pb));
#define OP_MUL2(i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(r##i) : "l"(pb));
#define OP_MIX2(i) asm volatile("{.reg .f32 lo, hi; fma.rn.f32x2 %0, %1, %2, %0; mov.b64 {lo, hi}, %0; min.f32 lo, lo, hi; mov.b64 %0, {lo, hi};}" : "+l"(r##i) : "l"(pa), "l"(pb));
DEF_KERNEL(k_fma2, D_P, REP8(OP_FMA2), FIN_P)
DEF_KERNEL(k_add2, D_P, REP8(OP_ADD2), FIN_P)
DEF_KERNEL(k_mul2, D_P, REP8(OP_MUL2), FIN_P)
DEF_KERNEL(k_mix2, D_P, REP8(OP_MIX2), FIN_P)

// broadcast LDS.128: every lane reads the same float4
__global__ void __launch_bounds__(1024) k_lds(float* out, long long* cycles, float a, float b) {
    __shared__ float4 s[256];
    if (threadIdx.x < 256) s[threadIdx.x] = make_float4(a, b, a, b);
    __syncthreads();
    float acc = 0;
    const volatile float4* vs = s;
    const int off = (int)a;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const volatile float4* q = &vs[(it * 8 + k + off) & 255];
            float x, y, z, w;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"((unsigned)__cvta_generic_to_shared((const void*)q)));
            acc += x;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename K>
void run(const char* name, K kernel, double lane_ops_per_iter, int threads) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    kernel<<<148, threads>>>(out, cyc, 1.0001f, 0.9999f);
    kernel<<<148, threads>>>(out, cyc, 1.0001f, 0.9999f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; i++) mean += h[i];
    mean /= 148;
    double ops = (double)ITERS * lane_ops_per_iter * threads;
    printf("%-28s threads=%4d  cycles=%9.0f  lane-ops/clk/SM=%7.1f  (%s)\n", name, threads, mean,
           ops / mean, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int threads : {256, 1024}) {
        run("fma.f32 (3-reg)", k_fma, CH, threads);
        run("add.f32", k_add, CH, threads);
        run("mul.f32", k_mul, CH, threads);
        run("min.f32 (ALU pipe)", k_min, CH, threads);
        run("setp+selp", k_setp, CH * 2, threads);
        run("fma + min interleaved", k_mix, CH * 2, threads);
        run("fma.f32x2 (2 flop-lanes)", k_fma2, CH * 2, threads);
        run("add.f32x2", k_add2, CH * 2, threads);
        run("mul.f32x2", k_mul2, CH * 2, threads);
        run("fma.f32x2 + min", k_mix2, CH * 3, threads);
        run("LDS.128 broadcast (loads)", k_lds, 8, threads);
    }
    return 0;
}
