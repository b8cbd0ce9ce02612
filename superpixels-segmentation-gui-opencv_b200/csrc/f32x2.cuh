// f32x2.cuh -- the sm_100 packed float32x2 instructions (FADD2 / FMUL2 / FFMA2) as inline PTX: same rounding per element as
// the scalar forms, half the issue slots.  Shared by the tiled SLIC and LSC assignment kernels.
#pragma once

namespace spx {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// add of two PRODUCTS.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false (scalar
// mul.rn/add.rn are left alone, as the PTX ISA promises); a flush-to-zero mismatch between the two stops it.  The
// operands here are squares of differences of an 8-bit value and a float32 mean (zero or >= 2^-48, never subnormal),
// so .ftz does not change any result.  tests/test_abi.py checks the SASS for stray FFMA2.
__device__ __forceinline__ u64 add2_products(u64 a, u64 b)
{
    u64 r;
    asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 bcast2(float v) { return pack2(v, v); }   // ptxas folds this into a scalar operand

}  // namespace spx
