// tcgen05 (5th-gen tensor core) implicit GEMM for sm_100a: C[M,N] = A[M,K] * B[K,N] in TF32 with fp32 accumulation
// in TMEM, either single pass (NTERMS=1) or the 3-term split A*B + A*B_lo + A_lo*B (NTERMS=3) that recovers
// fp32-grade products.  It takes the same operand / epilogue functors as the FFMA kernel in igemm.cuh, so every
// layer (conv im2col forward, dgrad, wgrad, dense layers, transposed operands) runs on it unchanged.
//
// Structure of one CTA (one 128 x BN output tile, 288 threads):
//   warps 0-7  producers, then epilogue.  Both operands are staged K-major in the canonical no-swizzle UMMA layout
//              (8-row x 16-byte core matrices; LBO padded to 144 bytes so every store pattern below is bank-conflict
//              free).  Operands that are contiguous along K in global memory are gathered with 16-byte cp.async
//              (zero-fill for padding / ragged edges), S stages deep: one global chunk is exactly one core-matrix
//              row.  Operands contiguous along M/N (TF-layout weights [in,out], transposed activations) are loaded
//              16 bytes at a time into registers two k-blocks ahead and transposed by 4-byte shared stores
//              (kind::tf32 has no 16-byte-granular MN-major layout).  For NTERMS=3 the low parts (x - tf32(x)) go
//              into the *_lo tiles: from registers for the transposed operands, by re-reading its own chunks for
//              the cp.async ones.
//   warp 8     allocates TMEM, and one elected lane issues tcgen05.mma (kind::tf32, M=128, N=BN, K=8) per k-step,
//              releasing smem stages with tcgen05.commit -> mbarrier.
//   epilogue   tcgen05.ld 32x32b.x16: thread (warp%4, lane) owns accumulator row 32*(warp%4)+lane, warps 4-7 take the
//              upper half of the columns; values go through the epilogue functor (bias/ReLU/gating/partials/...).
#pragma once
#include "igemm.cuh"

namespace tc {

constexpr int BM = 128;       // UMMA M
constexpr int BK = 32;        // floats per k-block (4 MMAs of K=8)
constexpr int PRODUCERS = 256;
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
// chunk f of a tile -> (first row, first k) of the 16 global bytes it covers
//   K-contiguous operand : 4 consecutive k of one row; 8 consecutive lanes walk the 128 bytes of one row
//   MN-contiguous operand: 4 consecutive rows at one k; 8 consecutive lanes walk 128 bytes along M/N, 4 k per warp
template <bool KCONTIG, int R>
__device__ __forceinline__ void chunk_rk(int f, int& r, int& k) {
    if (KCONTIG) { k = (f & 7) * 4; r = f >> 3; }
    else {
        const int rest = f >> 5;
        r = ((rest % (R / 32)) * 8 + (f & 7)) * 4;
        k = (rest / (R / 32)) * 4 + ((f >> 3) & 3);
    }
}

template <int BN, int NTERMS>
struct SmemPlan {
    static constexpr int A_BYTES = Tile<BM>::BYTES, B_BYTES = Tile<BN>::BYTES;
    static constexpr int STAGE = (A_BYTES + B_BYTES) * (NTERMS == 3 ? 2 : 1);
    static constexpr int MAX_STAGES = (212 * 1024) / STAGE;
    static constexpr int STAGES = MAX_STAGES > 5 ? 5 : (MAX_STAGES < 3 ? 3 : MAX_STAGES);
    static constexpr int BYTES = STAGES * STAGE + 256;
};

__device__ __forceinline__ float lo_part(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int BN, int NTERMS, bool SPLIT, class AOp, class BOp, class Epi>
__global__ void __launch_bounds__(THREADS, 1) tcgemm_kernel(const AOp a, const BOp b, const Epi epi, int M, int N, int K,
                                                            int kchunk, const float* __restrict__ zeros16) {
    using Plan = SmemPlan<BN, NTERMS>;
    constexpr int S = Plan::STAGES;
    using TA = Tile<BM>;
    using TB = Tile<BN>;
    constexpr int LA = TA::CHUNKS / PRODUCERS;
    constexpr int LB = (TB::CHUNKS + PRODUCERS - 1) / PRODUCERS;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t bars = sbase + S * Plan::STAGE;  // full[S], empty[S], accum, tmem slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
    const uint32_t accum_bar = bars + 8u * (2 * S);
    const uint32_t tmem_slot = bars + 8u * (2 * S + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.z;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = SPLIT ? z * kchunk : 0;
    const int kend = SPLIT ? min(K, kbeg + kchunk) : K;
    const int nkb = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), PRODUCERS / 32); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    if (warp == 8) {
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

    if (warp < 8) {
        // ---------------- producers ----------------
        int a_r[LA], a_k[LA], b_r[LB], b_k[LB];
        uint32_t a_off[LA], b_off[LB];
        typename AOp::Row a_row[LA];
        typename BOp::Row b_row[LB];
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            chunk_rk<AOp::KCONTIG, BM>(tid + i * PRODUCERS, a_r[i], a_k[i]);
            a_off[i] = TA::off(a_r[i], a_k[i]);
            a_row[i] = a.row(min(m0 + a_r[i], M - 1), z);
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            const int f = tid + i * PRODUCERS;
            if (f < TB::CHUNKS) {
                chunk_rk<BOp::KCONTIG, BN>(f, b_r[i], b_k[i]);
                b_off[i] = TB::off(b_r[i], b_k[i]);
            } else { b_r[i] = -1; b_k[i] = 0; b_off[i] = 0; }
            b_row[i] = b.row(min(n0 + max(b_r[i], 0), N - 1), z);
        }
        // K-contiguous operands: asynchronous 16-byte copies straight into the stage
        auto issue_async = [&](int kb) {
            const int s = kb % S, k0 = kbeg + kb * BK;
            if (AOp::KCONTIG) {
                const uint32_t ta = stage_a(s);
#pragma unroll
                for (int i = 0; i < LA; ++i) {
                    const int r = m0 + a_r[i], k = k0 + a_k[i];
                    const float* src = (r < M && k < kend) ? a.addr4(a_row[i], r, k, z) : nullptr;
                    cp_async16(ta + a_off[i], src ? src : zeros16, src != nullptr);
                }
            }
            if (BOp::KCONTIG) {
                const uint32_t tb = stage_b(s);
#pragma unroll
                for (int i = 0; i < LB; ++i) {
                    if (b_r[i] >= 0) {
                        const int r = n0 + b_r[i], k = k0 + b_k[i];
                        const float* src = (r < N && k < kend) ? b.addr4(b_row[i], r, k, z) : nullptr;
                        cp_async16(tb + b_off[i], src ? src : zeros16, src != nullptr);
                    }
                }
            }
        };
        // MN-contiguous operands: 16-byte loads into registers, two k-blocks ahead (two named register sets so the
        // compiler keeps them in registers)
        constexpr int NA = AOp::KCONTIG ? 1 : LA, NB = BOp::KCONTIG ? 1 : LB;
        float4 a_pf0[NA], a_pf1[NA], b_pf0[NB], b_pf1[NB];
        auto load_regs = [&](int kb, float4 (&ap)[NA], float4 (&bp)[NB]) {
            const int k0 = kbeg + kb * BK;
            if (!AOp::KCONTIG) {
#pragma unroll
                for (int i = 0; i < LA; ++i) {
                    const int r = m0 + a_r[i], k = k0 + a_k[i];
                    const float* src = (kb < nkb && r < M && k < kend) ? a.addr4(a_row[i], r, k, z) : nullptr;
                    ap[i] = src ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (!BOp::KCONTIG) {
#pragma unroll
                for (int i = 0; i < LB; ++i) {
                    const int r = n0 + b_r[i], k = k0 + b_k[i];
                    const float* src = (kb < nkb && b_r[i] >= 0 && r < N && k < kend) ? b.addr4(b_row[i], r, k, z) : nullptr;
                    bp[i] = src ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        // transpose store: rows r..r+3 of column k (row stride inside a core matrix is 16 bytes; r%8 is 0 or 4)
        auto sts_t = [&](uint32_t addr, const float4& v) {
            asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+16], %2;\n\tst.shared.f32 [%0+32], %3;\n\tst.shared.f32 [%0+48], %4;"
                         ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        };
        auto store_regs = [&](int s, const float4 (&ap)[NA], const float4 (&bp)[NB]) {
            if (!AOp::KCONTIG) {
#pragma unroll
                for (int i = 0; i < LA; ++i) {
                    const float4 v = ap[i];
                    sts_t(stage_a(s) + a_off[i], v);
                    if (NTERMS == 3) sts_t(stage_alo(s) + a_off[i], make_float4(lo_part(v.x), lo_part(v.y), lo_part(v.z), lo_part(v.w)));
                }
            }
            if (!BOp::KCONTIG) {
#pragma unroll
                for (int i = 0; i < LB; ++i) {
                    if (b_r[i] >= 0) {
                        const float4 v = bp[i];
                        sts_t(stage_b(s) + b_off[i], v);
                        if (NTERMS == 3) sts_t(stage_blo(s) + b_off[i], make_float4(lo_part(v.x), lo_part(v.y), lo_part(v.z), lo_part(v.w)));
                    }
                }
            }
        };
        auto split_lo = [&](uint32_t hi_addr, uint32_t lo_addr) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(hi_addr));
            // the tensor core reads the upper 19 bits of each operand (tf32); what it drops goes into the lo tile
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr), "f"(lo_part(v.x)), "f"(lo_part(v.y)),
                         "f"(lo_part(v.z)), "f"(lo_part(v.w)) : "memory");
        };
        // one k-block: registers -> stage, wait for the cp.async part, low parts, hand the stage to the MMA warp, refill
        auto block = [&](int kb, float4 (&ap)[NA], float4 (&bp)[NB]) {
            const int s = kb % S;
            // stage s is free: its previous user (block kb-S) was waited for at iteration kb-S+2
            store_regs(s, ap, bp);
            load_regs(kb + 2, ap, bp);
            cp_async_wait<S - 3>();  // this thread's cp.async chunks of block kb have landed
            if (NTERMS == 3) {
                if (AOp::KCONTIG) {
#pragma unroll
                    for (int i = 0; i < LA; ++i) split_lo(stage_a(s) + a_off[i], stage_alo(s) + a_off[i]);
                }
                if (BOp::KCONTIG) {
#pragma unroll
                    for (int i = 0; i < LB; ++i)
                        if (b_r[i] >= 0) split_lo(stage_b(s) + b_off[i], stage_blo(s) + b_off[i]);
                }
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(s));  // one arrival per producer warp
            // refill the stage block kb-2 used: its MMAs were committed a whole iteration ago, so this wait does not
            // put the producers in lock-step with the tensor core (one stage of slack)
            const int nxt = kb + S - 2;
            if (nxt < nkb) {
                if (kb >= 2) mbar_wait(empty_bar((kb - 2) % S), ((kb - 2) / S) & 1);
                issue_async(nxt);
            }
            cp_async_commit();
        };
        for (int kb = 0; kb < S - 2; ++kb) {
            if (kb < nkb) issue_async(kb);
            cp_async_commit();
        }
        load_regs(0, a_pf0, b_pf0);
        load_regs(1, a_pf1, b_pf1);
        for (int kb = 0; kb < nkb; kb += 2) {
            block(kb, a_pf0, b_pf0);
            if (kb + 1 < nkb) block(kb + 1, a_pf1, b_pf1);
        }
        // ---------------- epilogue ----------------
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
        const int sub = warp & 3, half = warp >> 2;
        const int m = m0 + sub * 32 + lane;
        constexpr int HALF_COLS = BN / 2;
#pragma unroll
        for (int c = 0; c < HALF_COLS; c += 16) {
            const int col = half * HALF_COLS + c;
            float v[16];
            if (nkb > 0) tmem_ld16(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)col, v);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
            if (m < M && n0 + col < N) epi.template store<16>(m, n0 + col, v, z);
        }
        tc_fence_before();
    } else {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = make_idesc(BN, false, false);  // both operands are staged K-major
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S;
            mbar_wait(full_bar(s), (kb / S) & 1);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint64_t da = TA::desc(stage_a(s), ks), db = TB::desc(stage_b(s), ks);
                    if (NTERMS == 3) {
                        const uint64_t dal = TA::desc(stage_alo(s), ks), dbl = TB::desc(stage_blo(s), ks);
                        mma_tf32(tmem_base, da, dbl, idesc, (kb | ks) != 0);
                        mma_tf32(tmem_base, dal, db, idesc, 1);
                        mma_tf32(tmem_base, da, db, idesc, 1);
                    } else {
                        mma_tf32(tmem_base, da, db, idesc, (kb | ks) != 0);
                    }
                }
                mma_commit(empty_bar(s));                    // frees the stage when these MMAs retire
                if (kb == nkb - 1) mma_commit(accum_bar);    // accumulator complete
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 8) {
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
    dim3 grid(ceil_div(M, BM), ceil_div(N, BN), Z);
    kern<<<grid, THREADS, Plan::BYTES + 1024, h->stream>>>(a, b, e, M, N, K, kchunk, h->zeros16);
    DQN_CUDA_OK(cudaGetLastError());
    h->launches++;
}

}  // namespace tc
