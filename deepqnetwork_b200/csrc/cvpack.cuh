// Packed conv weights of the image-resident kernels (cvgemm.cuh, cvtower.cuh): layout and the per-element pack routines,
// shared by pack_all_kernel and by the optimizer tail (optim.cu), which re-packs the conv weights it has just updated.
#pragma once
#include "common.cuh"

namespace cv {

__device__ __forceinline__ float pk_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }  // x - tf32(x)

// ---- the 3-term product, mixed precision (NTERMS == 3) ------------------------------------------------------------------
// a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with hi = the fp32 word as TF32 (the tensor core reads its top 19 bits) and
// lo = x - tf32(x).  The two cross terms are 2^-11 of the product, so they only need ~8 bits of their own: they run as
// kind::f16 MMAs on BF16 copies (K = 16 per instruction) while hi*hi stays a kind::tf32 MMA (K = 8).  A k-block of 32
// then takes 4 + 2 + 2 = 8 tensor instructions instead of 12; every tcgen05.mma of these kernels costs ~40 cycles whatever
// its N (the 128 x 32 B read of the A operand), so the k-loops get a third shorter.  Error of a product: 2^-19 relative
// (bf16 rounding of a cross-term operand) instead of 2^-22 (the dropped lo*lo), still far below what the tensor core's
// truncating fp32 accumulation contributes.
//
// packed weights of one layer: per k-block of 32 and N = cout rows, N * 256 bytes:
//   [0, 128 N)      hi, fp32, dense canonical K-major UMMA layout (8-row x 16-byte core matrices, LBO 128, SBO 1024):
//                   element (n, k) at (n/8)*1024 + (k/4)*128 + (n%8)*16 + (k%4)*4
//   [128 N, 192 N)  bf16(x), same canonical layout for 16-bit elements (LBO 128, SBO 512):
//                   element (n, k) at (n/8)*512 + (k/8)*128 + (n%8)*16 + (k%8)*2
//   [192 N, 256 N)  bf16(lo)
constexpr uint32_t PK_LBO = 128, PK_SBO = 1024, PK_SBO16 = 512;
__host__ __device__ __forceinline__ uint32_t pk_off32(int n, int kin) { return (uint32_t)((n >> 3) * PK_SBO + (kin >> 2) * PK_LBO + (n & 7) * 16 + (kin & 3) * 4); }
__host__ __device__ __forceinline__ uint32_t pk_off16(int n, int kin) { return (uint32_t)((n >> 3) * PK_SBO16 + (kin >> 3) * PK_LBO + (n & 7) * 16 + (kin & 7) * 2); }
__device__ __forceinline__ uint16_t bf16_bits(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }
// two consecutive k as one 32-bit cell of a 16-bit operand: the even k in the low half
__device__ __forceinline__ uint32_t bf16_pair(float k_even, float k_odd) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(k_even, k_odd);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// one weight into its k-block (blk = first byte of the block, N rows)
__device__ __forceinline__ void pk_store(uint8_t* blk, int N, int n, int kin, float v) {
    *reinterpret_cast<float*>(blk + pk_off32(n, kin)) = v;
    const uint32_t o = pk_off16(n, kin);
    *reinterpret_cast<uint16_t*>(blk + 128 * N + o) = bf16_bits(v);
    *reinterpret_cast<uint16_t*>(blk + 192 * N + o) = bf16_bits(pk_lo(v));
}
// four consecutive k (kin % 4 == 0) of row n: one 16-byte and two 8-byte stores
__device__ __forceinline__ void pk_store4(uint8_t* blk, int N, int n, int kin, const float4& r) {
    *reinterpret_cast<float4*>(blk + pk_off32(n, kin)) = r;
    const uint32_t o = pk_off16(n, kin);
    *reinterpret_cast<uint2*>(blk + 128 * N + o) = make_uint2(bf16_pair(r.x, r.y), bf16_pair(r.z, r.w));
    *reinterpret_cast<uint2*>(blk + 192 * N + o) = make_uint2(bf16_pair(pk_lo(r.x), pk_lo(r.y)), bf16_pair(pk_lo(r.z), pk_lo(r.w)));
}
// eight consecutive k (kin % 8 == 0) of row n: two 16-byte fp32 pieces and one 16-byte piece per bf16 tile (whole core-matrix rows)
__device__ __forceinline__ void pk_store8(uint8_t* blk, int N, int n, int kin, const float4& r0, const float4& r1) {
    *reinterpret_cast<float4*>(blk + pk_off32(n, kin)) = r0;
    *reinterpret_cast<float4*>(blk + pk_off32(n, kin + 4)) = r1;
    const uint32_t o = pk_off16(n, kin);
    *reinterpret_cast<uint4*>(blk + 128 * N + o) = make_uint4(bf16_pair(r0.x, r0.y), bf16_pair(r0.z, r0.w), bf16_pair(r1.x, r1.y), bf16_pair(r1.z, r1.w));
    *reinterpret_cast<uint4*>(blk + 192 * N + o) = make_uint4(bf16_pair(pk_lo(r0.x), pk_lo(r0.y)), bf16_pair(pk_lo(r0.z), pk_lo(r0.w)),
                                                              bf16_pair(pk_lo(r1.x), pk_lo(r1.y)), bf16_pair(pk_lo(r1.z), pk_lo(r1.w)));
}
// Layer 0 of the fused tower (cvtower.cuh) multiplies the uint8 frame values themselves (exact in bf16) with w / 255 split
// into three bf16 pieces (24 bits): per k-block [b1 | b2 | b3 | unused], each N * 64 bytes in the 16-bit layout above.
__device__ __forceinline__ void pk_store_u8(uint8_t* blk, int N, int n, int kin, float v) {
    const float w = v / 255.0f;
    const float b1 = __bfloat162float(__float2bfloat16_rn(w));
    const float r1 = w - b1;
    const float b2 = __bfloat162float(__float2bfloat16_rn(r1));
    const float r2 = r1 - b2;
    const uint32_t o = pk_off16(n, kin);
    *reinterpret_cast<uint16_t*>(blk + o) = bf16_bits(b1);
    *reinterpret_cast<uint16_t*>(blk + 64 * N + o) = bf16_bits(b2);
    *reinterpret_cast<uint16_t*>(blk + 128 * N + o) = bf16_bits(r2);
}

}  // namespace cv
