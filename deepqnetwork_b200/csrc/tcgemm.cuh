// tcgen05 (5th-gen tensor core) implicit GEMM for sm_100a: C[M,N] = A[M,K] * B[K,N] in TF32 with fp32 accumulation
// in TMEM, either single pass (NTERMS=1) or the 3-term split A*B + A*B_lo + A_lo*B (NTERMS=3) that recovers
// fp32-grade products.  It takes the same operand / epilogue functors as the FFMA kernel in igemm.cuh, so every
// layer (conv im2col forward, dgrad, wgrad, dense layers, transposed operands) runs on it unchanged.
//
// Structure of one CTA (one 128 x BN output tile, 544 threads):
//   warps 0-15 producers, then epilogue.  Both operands are staged K-major in the canonical no-swizzle UMMA layout
//              (8-row x 16-byte core matrices; LBO padded to 144 bytes so every store pattern below is bank-conflict
//              free).  Operands that are contiguous along K in global memory are gathered with 16-byte cp.async
//              (zero-fill for padding / ragged edges), S stages deep: one global chunk is exactly one core-matrix
//              row.  Operands contiguous along M/N (TF-layout weights [in,out], transposed activations) are loaded
//              16 bytes at a time into registers two k-blocks ahead and transposed by 4-byte shared stores
//              (kind::tf32 has no 16-byte-granular MN-major layout).  For NTERMS=3 the low parts (x - tf32(x)) go
//              into the *_lo tiles: from registers for the transposed operands, by re-reading its own chunks for
//              the cp.async ones.
//   warp 16    allocates TMEM, and one elected lane issues tcgen05.mma (kind::tf32, M=128, N=BN, K=8) per k-step,
//              releasing smem stages with tcgen05.commit -> mbarrier.
//   epilogue   tcgen05.ld 32x32b: thread (warp%4, lane) owns accumulator row 32*(warp%4)+lane, warp/4 selects a quarter
//              of the columns; values go through the epilogue functor (bias/ReLU/gating/partials/...).
#pragma once
#include "igemm.cuh"

namespace tc {

constexpr int BM = 128;       // UMMA M
constexpr int BK = 32;        // floats per k-block (4 MMAs of K=8)
constexpr int PRODUCERS = 512;  // 16 producer warps: the per-k-block address/convert work is split thin enough that
                                 // a single warp's serial instruction stream no longer bounds the k-loop
constexpr int THREADS = PRODUCERS + 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const uint32_t n = valid ? 16u : 0u;  // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared-memory matrix descriptor, no-swizzle canonical layouts (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);           // start address        bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;  // leading byte offset  bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;  // stride byte offset   bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    return d;                                           // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

// instruction descriptor: kind::tf32, fp32 accumulate, M=128, N (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t make_idesc(int n, bool a_mn_major, bool b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                        // c_format  = F32
    d |= 2u << 7;                        // a_format  = TF32
    d |= 2u << 10;                       // b_format  = TF32
    d |= (a_mn_major ? 1u : 0u) << 15;   // a_major
    d |= (b_mn_major ? 1u : 0u) << 16;   // b_major
    d |= (uint32_t)(n >> 3) << 17;       // n_dim
    d |= (uint32_t)(BM >> 4) << 24;      // m_dim
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}


// Sum of the same columns of several accumulators (the issuing warps' partial sums, added in index order): all loads are issued
// before ONE tcgen05.wait::ld -- a load + wait pair per accumulator paid a tensor-memory round trip each, and the epilogue of a
// k-split tile is nothing but such round trips.
__device__ __forceinline__ void tmem_ld16_sum2(uint32_t a0, uint32_t a1, float (&v)[16]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(a0), "r"(a1));
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float s = __uint_as_float(r[i]);
        s += __uint_as_float(r[16 + i]);
        v[i] = s;
    }
}
__device__ __forceinline__ void tmem_ld16_sum4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, float (&v)[16]) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%64];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%65];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47}, [%66];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%67];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3));
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float s = __uint_as_float(r[i]);
        s += __uint_as_float(r[16 + i]);
        s += __uint_as_float(r[32 + i]);
        s += __uint_as_float(r[48 + i]);
        v[i] = s;
    }
}
__device__ __forceinline__ void tmem_ld32_sum2(uint32_t a0, uint32_t a1, float (&v)[32]) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%64];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%65];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(a0), "r"(a1));
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float s = __uint_as_float(r[i]);
        s += __uint_as_float(r[32 + i]);
        v[i] = s;
    }
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major no-swizzle tile [R rows x 32 k]: element (r,k) at (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4.
// LBO = 144 (128-byte core matrix + 16 bytes) and SBO = 8*LBO + 16 are padded so that both producer patterns are
// bank-conflict free while their GLOBAL side stays coalesced:
//   K-contiguous operand : 8 lanes x 16 B = one 128-byte row piece -> chunk stride LBO: banks 4*c (mod 32), distinct
//   MN-contiguous operand: 8 lanes x 16 B along M/N at 4 consecutive k -> 4-byte stores at word
//                          292*(r/8) + 4*(r%8) + k%4 == 4*(r/8) + 16*((r/4)%2) + 4*j + k%4 (mod 32), distinct
constexpr uint32_t LBO = 144, SBO = 8 * LBO + 16;
template <int R>
struct Tile {
    static constexpr int BYTES = (R / 8) * SBO;
    static constexpr int CHUNKS = R * 8;  // 16-byte pieces
    __device__ static __forceinline__ uint32_t off(int r, int k) {
        return (uint32_t)((r >> 3) * SBO + (k >> 2) * LBO + (r & 7) * 16 + (k & 3) * 4);
    }
    __device__ static __forceinline__ uint64_t desc(uint32_t tile_addr, int ks) { return make_desc(tile_addr + ks * 2 * LBO, LBO, SBO); }
};
// Producer work assignment (PRODUCERS threads, tile of R rows x 32 k = R*8 chunks of 16 bytes):
//   K-contiguous operand : thread owns ONE row and RUN consecutive chunks of it (one row state, contiguous global
//                          bytes); 8/RUN neighbouring lanes cover the 128 bytes of a row
//   MN-contiguous operand: 8 lanes walk 128 bytes along M/N, 4 consecutive k per warp; all chunks of a thread share k
template <bool KCONTIG, int R>
struct Assign {
    static constexpr int RUN = KCONTIG ? (R * 8 >= PRODUCERS ? R * 8 / PRODUCERS : 1) : 1;   // chunks per row run
    static constexpr int N = KCONTIG ? RUN : (R >= 64 ? R / 64 : 1);                           // chunks per thread
    // chunk i of thread tid -> (row, k offset inside the k-block) or row < 0 if the thread has no i-th chunk
    __device__ static __forceinline__ void rk(int tid, int i, int& r, int& k) {
        if (KCONTIG) {
            constexpr int TPR = 8 / RUN;  // threads per row
            r = tid / TPR;
            k = ((tid % TPR) * RUN + i) * 4;
            if (r >= R) r = -1;
        } else {
            const int w = tid >> 5;
            const int g = (w >> 3) + 2 * i;            // group of 8 row-quads
            r = g < R / 32 ? (g * 8 + (tid & 7)) * 4 : -1;
            k = (w & 7) * 4 + ((tid >> 3) & 3);
        }
    }
};

template <int BN, int NTERMS>
struct SmemPlan {
    static constexpr int A_BYTES = Tile<BM>::BYTES, B_BYTES = Tile<BN>::BYTES;
    static constexpr int STAGE = (A_BYTES + B_BYTES) * (NTERMS == 3 ? 2 : 1);
    static constexpr int MAX_STAGES = (212 * 1024) / STAGE;
    static constexpr int STAGES = MAX_STAGES > 5 ? 5 : (MAX_STAGES < 3 ? 3 : MAX_STAGES);
    static constexpr int BYTES = STAGES * STAGE + 256;
};

// epilogue functors may ask for a shared-memory transpose so that global accesses run along n (row-major outputs whose
// row pitch is much larger than the tile: dense wgrad)
template <class E, class = void> struct wants_coalesce { static constexpr bool value = false; };
template <class E> struct wants_coalesce<E, decltype((void)E::COALESCE)> { static constexpr bool value = E::COALESCE; };

__device__ __forceinline__ float lo_part(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int BN, int NTERMS, bool SPLIT, class AOp, class BOp, class Epi>
__global__ void __launch_bounds__(THREADS, 1) tcgemm_kernel(const AOp a, const BOp b, const Epi epi, int M, int N, int K,
                                                            int kchunk, const float* __restrict__ zeros,
                                                            unsigned long long* __restrict__ dbg) {
    using Plan = SmemPlan<BN, NTERMS>;
    constexpr int S = Plan::STAGES;
    using TA = Tile<BM>;
    using TB = Tile<BN>;
    using AA = Assign<AOp::KCONTIG, BM>;
    using AB = Assign<BOp::KCONTIG, BN>;
    constexpr int LA = AA::N, LB = AB::N;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t bars = sbase + S * Plan::STAGE;  // full[S], empty[S], accum, tmem slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
    const uint32_t accum_bar = bars + 8u * (2 * S);
    const uint32_t tmem_slot = bars + 8u * (2 * S + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.z;
    // n-tile is the fastest-varying block index: CTAs that run together share the same A rows (L2) and, for the dense
    // wgrad, touch whole contiguous parameter rows (DRAM page locality of the fused optimizer epilogue)
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = SPLIT ? z * kchunk : 0;
    const int kend = SPLIT ? min(K, kbeg + kchunk) : K;
    const int nkb = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), PRODUCERS / 32); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    if (warp == PRODUCERS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    auto stage_a = [&](int s) { return sbase + s * Plan::STAGE; };
    auto stage_alo = [&](int s) { return stage_a(s) + Plan::A_BYTES; };
    auto stage_b = [&](int s) { return stage_a(s) + Plan::A_BYTES * (NTERMS == 3 ? 2 : 1); };
    auto stage_blo = [&](int s) { return stage_b(s) + Plan::B_BYTES; };

    if (warp < PRODUCERS / 32) {
        // ---------------- producers ----------------
        int a_r[LA], a_k[LA], b_r[LB], b_k[LB];
        uint32_t a_off[LA], b_off[LB];
        typename AOp::Row a_row[LA];
        typename BOp::Row b_row[LB];
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            AA::rk(tid, i, a_r[i], a_k[i]);
            a_off[i] = a_r[i] >= 0 ? TA::off(a_r[i], a_k[i]) : 0;
            a_row[i] = a.row(min(m0 + max(a_r[i], 0), M - 1), z);
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            AB::rk(tid, i, b_r[i], b_k[i]);
            b_off[i] = b_r[i] >= 0 ? TB::off(b_r[i], b_k[i]) : 0;
            b_row[i] = b.row(min(n0 + max(b_r[i], 0), N - 1), z);
        }
        // Operand tiles go global -> registers -> shared, PF k-blocks ahead.  (A producer-side fence.proxy.async compiles
        // to MEMBAR.ALL.CTA, which waits for every load/async copy the thread still has in flight and serialised the
        // k-loop on memory latency; the fence therefore lives on the consumer side, after the barrier acquire.)
        constexpr int PF = 4;  // k-blocks of loads in flight per thread
        float4 a_pf[PF][LA], b_pf[PF][LB];
        auto load_regs = [&](int kb, float4 (&ap)[LA], float4 (&bp)[LB]) {
            const int k0 = kbeg + kb * BK;
#pragma unroll
            for (int i = 0; i < LA; ++i) {
                const int r = m0 + a_r[i], k = k0 + a_k[i];
                const float* src = (kb < nkb && a_r[i] >= 0 && r < M && k < kend) ? a.addr4(a_row[i], r, k, z) : nullptr;
                // padding / ragged edges read a 16-byte block of zeros: the select is on the ADDRESS, so nothing depends
                // on the loaded value until the shared-memory store PF k-blocks later (a select on the value would
                // stall this thread for a full memory round trip every k-block)
                ap[i] = __ldg(reinterpret_cast<const float4*>(src ? src : zeros));
            }
#pragma unroll
            for (int i = 0; i < LB; ++i) {
                const int r = n0 + b_r[i], k = k0 + b_k[i];
                const float* src = (kb < nkb && b_r[i] >= 0 && r < N && k < kend) ? b.addr4(b_row[i], r, k, z) : nullptr;
                bp[i] = __ldg(reinterpret_cast<const float4*>(src ? src : zeros));
            }
        };
        // K-contiguous chunk: 4 consecutive k of one row = one 16-byte core-matrix row
        auto sts_v = [&](uint32_t addr, const float4& v) {
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        };
        // MN-contiguous chunk: rows r..r+3 of column k (row stride inside a core matrix is 16 bytes; r%8 is 0 or 4)
        auto sts_t = [&](uint32_t addr, const float4& v) {
            asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+16], %2;\n\tst.shared.f32 [%0+32], %3;\n\tst.shared.f32 [%0+48], %4;"
                         ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        };
        auto store_regs = [&](int s, const float4 (&ap)[LA], const float4 (&bp)[LB]) {
#pragma unroll
            for (int i = 0; i < LA; ++i) {
                if (a_r[i] >= 0) {
                    const float4 v = ap[i];
                    // the tensor core reads the upper 19 bits of each word (tf32); what it drops goes into the lo tile
                    const float4 l = make_float4(lo_part(v.x), lo_part(v.y), lo_part(v.z), lo_part(v.w));
                    if (AOp::KCONTIG) { sts_v(stage_a(s) + a_off[i], v); if (NTERMS == 3) sts_v(stage_alo(s) + a_off[i], l); }
                    else { sts_t(stage_a(s) + a_off[i], v); if (NTERMS == 3) sts_t(stage_alo(s) + a_off[i], l); }
                }
            }
#pragma unroll
            for (int i = 0; i < LB; ++i) {
                if (b_r[i] >= 0) {
                    const float4 v = bp[i];
                    const float4 l = make_float4(lo_part(v.x), lo_part(v.y), lo_part(v.z), lo_part(v.w));
                    if (BOp::KCONTIG) { sts_v(stage_b(s) + b_off[i], v); if (NTERMS == 3) sts_v(stage_blo(s) + b_off[i], l); }
                    else { sts_t(stage_b(s) + b_off[i], v); if (NTERMS == 3) sts_t(stage_blo(s) + b_off[i], l); }
                }
            }
        };
        auto block = [&](int kb, float4 (&ap)[LA], float4 (&bp)[LB]) {
            const int s = kb % S;
            if (dbg && tid == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0 && kb < 64) dbg[kb * 8 + 0] = clock64();
            if (kb >= S) mbar_wait(empty_bar(s), ((kb / S) - 1) & 1);  // the MMAs of block kb-S are done with this stage
            if (dbg && tid == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0 && kb < 64) dbg[kb * 8 + 1] = clock64();
            store_regs(s, ap, bp);
            if (dbg && tid == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0 && kb < 64) dbg[kb * 8 + 2] = clock64();
            // Publish the tile: release-arrive on the stage barrier (one arrival per producer warp).  The generic->async
            // proxy fence is executed by the MMA thread after it acquires the barrier (see the issuer loop): a fence
            // here would compile to MEMBAR.ALL.CTA and drain this thread's prefetch loads every k-block.
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(s));
            load_regs(kb + PF, ap, bp);
            if (dbg && tid == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0 && kb < 64) dbg[kb * 8 + 3] = clock64();
        };
#pragma unroll
        for (int p = 0; p < PF; ++p) load_regs(p, a_pf[p], b_pf[p]);
        for (int kb = 0; kb < nkb; kb += PF) {
#pragma unroll
            for (int p = 0; p < PF; ++p)
                if (kb + p < nkb) block(kb + p, a_pf[p], b_pf[p]);
        }
        // ---------------- epilogue ----------------
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
        const int sub = warp & 3, quarter = warp >> 2;
        const int m = m0 + sub * 32 + lane;
        constexpr int QCOLS = BN / 4;  // columns per warp: 8, 16 or 32
        const uint32_t trow = tmem_base + ((uint32_t)(sub * 32) << 16);
        if constexpr (wants_coalesce<Epi>::value) {
            // accumulators -> smem tile [128][BN+4] (stage buffers are idle now) -> rows handed to whole warps
            constexpr int PITCH = BN + 4;
            float* tile = reinterpret_cast<float*>(smem + (sbase - smem_u32(smem)));
            const int row = sub * 32 + lane;
            if (QCOLS == 8) {
                float v[8];
                if (nkb > 0) tmem_ld8(trow + (uint32_t)(quarter * 8), v);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; i += 4)
                    *reinterpret_cast<float4*>(tile + row * PITCH + quarter * 8 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
                for (int c = 0; c < QCOLS; c += 16) {
                    float v[16];
                    if (nkb > 0) tmem_ld16(trow + (uint32_t)(quarter * QCOLS + c), v);
                    else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(tile + row * PITCH + quarter * QCOLS + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(PRODUCERS) : "memory");
            constexpr int LANES_PER_ROW = BN / 4;            // float4 per row
            constexpr int ROWS_PER_PASS = PRODUCERS / LANES_PER_ROW;
            const int c4 = (tid % LANES_PER_ROW) * 4;
            const float step = epi.step_size();
            constexpr int RB = 4;  // rows per batch: all their read-modify-write loads are issued before any store
            for (int r0 = tid / LANES_PER_ROW; r0 < BM; r0 += ROWS_PER_PASS * RB) {
                typename Epi::State st[RB];
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const int r = r0 + j * ROWS_PER_PASS;
                    if (r < BM && m0 + r < M && n0 + c4 < N) epi.load(m0 + r, n0 + c4, st[j]);
                }
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const int r = r0 + j * ROWS_PER_PASS;
                    if (r < BM && m0 + r < M && n0 + c4 < N)
                        epi.apply(m0 + r, n0 + c4, *reinterpret_cast<const float4*>(tile + r * PITCH + c4), st[j], step);
                }
            }
        } else if (QCOLS == 8) {
            const int col = quarter * 8;
            float v[8];
            if (nkb > 0) tmem_ld8(trow + (uint32_t)col, v);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.f;
            }
            if (m < M && n0 + col < N) epi.template store<8>(m, n0 + col, v, z);
        } else {
#pragma unroll
            for (int c = 0; c < QCOLS; c += 16) {
                const int col = quarter * QCOLS + c;
                float v[16];
                if (nkb > 0) tmem_ld16(trow + (uint32_t)col, v);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                if (m < M && n0 + col < N) epi.template store<16>(m, n0 + col, v, z);
            }
        }
        tc_fence_before();
    } else {
        // ---------------- MMA issuer ----------------
        // Every lane computes the (warp-uniform) descriptors so they live in uniform registers; one elected lane
        // issues.  A divergent `if (lane == 0)` around the whole loop makes the compiler wrap each tcgen05.mma in a
        // waterfall loop and the issue loop, not the tensor core, becomes the bottleneck.
        const uint32_t idesc = make_idesc(BN, false, false);  // both operands are staged K-major
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        const uint64_t da0 = TA::desc(stage_a(0), 0), db0 = TB::desc(stage_b(0), 0);
        const uint64_t dal0 = TA::desc(stage_alo(0), 0), dbl0 = TB::desc(stage_blo(0), 0);
        constexpr uint64_t STAGE16 = Plan::STAGE >> 4, KS16 = (2 * LBO) >> 4;  // descriptor address units (16 bytes)
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S;
            if (dbg && lane == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0 && kb < 64) dbg[kb * 8 + 4] = clock64();
            mbar_wait(full_bar(s), (kb / S) & 1);   // acquire: the producers' shared-memory stores are visible here
            if (dbg && lane == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0 && kb < 64) dbg[kb * 8 + 5] = clock64();
            fence_proxy_async();                      // ... and from here on to the tensor core's async proxy
            tc_fence_after();
            if (dbg && lane == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0 && kb < 64) dbg[kb * 8 + 6] = clock64();
            const uint64_t so = (uint64_t)s * STAGE16;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                const uint64_t o = so + ks * KS16;
                const uint32_t acc = (kb | ks) != 0;
                if (leader) {
                    if (NTERMS == 3) {
                        mma_tf32(tmem_base, da0 + o, dbl0 + o, idesc, acc);
                        mma_tf32(tmem_base, dal0 + o, db0 + o, idesc, 1);
                        mma_tf32(tmem_base, da0 + o, db0 + o, idesc, 1);
                    } else {
                        mma_tf32(tmem_base, da0 + o, db0 + o, idesc, acc);
                    }
                }
            }
            if (leader) {
                mma_commit(empty_bar(s));                    // frees the stage when these MMAs retire
                if (kb == nkb - 1) mma_commit(accum_bar);    // accumulator complete
            }
            if (dbg && lane == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0 && kb < 64) dbg[kb * 8 + 7] = clock64();
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == PRODUCERS / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

template <int BN, int NTERMS, bool SPLIT, class AOp, class BOp, class Epi>
static void launch(dqn_learner* h, const AOp& a, const BOp& b, const Epi& e, int M, int N, int K, int Z, int kchunk) {
    using Plan = SmemPlan<BN, NTERMS>;
    auto kern = tcgemm_kernel<BN, NTERMS, SPLIT, AOp, BOp, Epi>;
    static bool configured = false;  // one per template instantiation
    if (!configured) {
        DQN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::BYTES + 1024));
        configured = true;
    }
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), Z);
    launch_kernel(false, kern, grid, dim3(THREADS), (size_t)Plan::BYTES + 1024, h->stream, a, b, e, M, N, K, kchunk, (const float*)h->zeros16, h->tc_dbg);
    h->launches++;
}

}  // namespace tc
