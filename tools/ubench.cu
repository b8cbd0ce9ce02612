// ubench.cu -- issue-rate microbenchmarks that decide the shape of the SLIC assignment kernel:
// scalar FADD/FMUL/FFMA vs the sm_100 packed f32x2 forms, FSETP+SEL, FMNMX, LDS broadcast, REDUX.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run on the B200.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

__device__ __forceinline__ unsigned long long pk(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed)
{
    float a[UNROLL], b = seed + threadIdx.x * 1e-3f, c = seed * 0.5f;
    unsigned long long A[UNROLL], B = pk(b, c), Cc = pk(c, b);
    __shared__ float sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    int ia[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) { a[u] = seed + u; A[u] = pk(seed + u, seed - u); ia[u] = u + threadIdx.x; }
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (MODE == 0) a[u] = __fadd_rn(a[u], b);
            if (MODE == 1) a[u] = __fmul_rn(a[u], b);
            if (MODE == 2) a[u] = __fmaf_rn(a[u], b, c);
            if (MODE == 3) A[u] = add2(A[u], B);
            if (MODE == 4) A[u] = mul2(A[u], B);
            if (MODE == 5) A[u] = fma2(A[u], B, Cc);
            if (MODE == 6) { bool p = a[u] < b; a[u] = p ? a[u] + 1.0f : b; ia[u] = p ? i : ia[u]; }   // FSETP + FSEL/FADD + SEL
            if (MODE == 7) a[u] = fminf(a[u], b + (float)i);
            if (MODE == 8) a[u] += sm[(i + u) & 63];                       // uniform-address LDS + FADD
            if (MODE == 9) ia[u] = __reduce_add_sync(0xffffffffu, ia[u]);  // REDUX
            if (MODE == 10) { a[u] = __fadd_rn(a[u], b); ia[u] = ia[u] * 3 + i; }   // FADD + IMAD mix
            if (MODE == 11) { A[u] = add2(A[u], B); ia[u] = (ia[u] ^ i) + u; }      // FADD2 + int ALU mix
        }
    }
    float s = 0;
#pragma unroll
    for (int u = 0; u < UNROLL; u++) s += a[u] + (float)(A[u] & 0xffff) + ia[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_inner)
{
    float* out;
    int blocks = 148 * 8;
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)blocks * 8 * ITERS * UNROLL * ops_per_inner;
    printf("%-28s %8.3f ms  %8.1f Gwarp-instr/s  (%.2f per clk per SM at 1.9 GHz)\n", name, ms, warp_instr / ms / 1e6,
           warp_instr / ms / 1e6 / 148 / 1.9);
    cudaFree(out);
}

int main()
{
    run<0>("FADD", 1); run<1>("FMUL", 1); run<2>("FFMA", 1);
    run<3>("FADD2 (f32x2)", 1); run<4>("FMUL2 (f32x2)", 1); run<5>("FFMA2 (f32x2)", 1);
    run<6>("FSETP+FSEL+SEL (3 instr)", 3); run<7>("FMNMX(+I2F+FADD)", 1); run<8>("LDS bcast + FADD", 2);
    run<9>("REDUX", 1); run<10>("FADD + IMAD", 2); run<11>("FADD2 + 2 int ALU", 3);
    return 0;
}
