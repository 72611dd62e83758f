// tcgen05 (5th-gen tensor core) implicit GEMM for sm_100a: C[M,N] = A[M,K] * B[K,N] in TF32 with fp32 accumulation
// in TMEM, either single pass (NTERMS=1) or the 3-term split A*B + A*B_lo + A_lo*B (NTERMS=3) that recovers
// fp32-grade products.  It takes the same operand / epilogue functors as the FFMA kernel in igemm.cuh, so every
// layer (conv im2col forward, dgrad, wgrad, dense layers, transposed operands) runs on it unchanged.
//
// Structure of one CTA (one 128 x BN output tile, 288 threads):
//   warps 0-7  producers, then epilogue.  Operand tiles are gathered straight from global/L2 into shared memory with
//              16-byte cp.async (zero-fill for padding / ragged edges), S stages deep.  A 16-byte global chunk is
//              exactly one row of a UMMA core matrix in either canonical no-swizzle layout: K-major for operands
//              that are contiguous along K, MN-major for operands contiguous along M/N, so no transposes are needed.
//              For NTERMS=3 each thread then rewrites its own chunks' low parts (x - tf32(x)) into the *_lo tiles.
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

// byte offset of 16-byte chunk inside a tile
//   K-major  tile [R rows x 32 k]: chunk (r, c=k/4)   -> (r/8)*1024 + c*128 + (r%8)*16       LBO=128, SBO=1024
//   MN-major tile [R mn   x 32 k]: chunk (g=mn/4, k)  -> (k/8)*(R*32) + g*128 + (k%8)*16    LBO=R*32, SBO=128
template <bool KCONTIG, int R>
struct TileMap {
    static constexpr int CHUNKS = R * 8;
    // chunk f of the tile -> (first row, k offset); consecutive f walk 16-byte pieces of one 128-byte core matrix
    __device__ static __forceinline__ void rk(int f, int& r, int& k) {
        if (KCONTIG) { r = f % R; k = (f / R) * 4; }
        else { k = (f & 7) + 8 * ((f >> 3) / (R / 4)); r = ((f >> 3) % (R / 4)) * 4; }
    }
    __device__ static __forceinline__ uint32_t off(int r, int k) {
        if (KCONTIG) return (uint32_t)((r >> 3) * 1024 + (k >> 2) * 128 + (r & 7) * 16);
        return (uint32_t)((k >> 3) * (R * 32) + (r >> 2) * 128 + (k & 7) * 16);
    }
    __device__ static __forceinline__ uint64_t desc(uint32_t tile_addr, int ks) {
        if (KCONTIG) return make_desc(tile_addr + ks * 256, 128, 1024);
        return make_desc(tile_addr + ks * (R * 32), R * 32, 128);
    }
};

template <int BN, int NTERMS>
struct SmemPlan {
    static constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
    static constexpr int STAGE = (A_BYTES + B_BYTES) * (NTERMS == 3 ? 2 : 1);
    static constexpr int MAX_STAGES = (200 * 1024) / STAGE;
    static constexpr int STAGES = MAX_STAGES > 4 ? 4 : (MAX_STAGES < 2 ? 2 : MAX_STAGES);
    static constexpr int BYTES = STAGES * STAGE + 256;
};

template <int BN, int NTERMS, bool SPLIT, class AOp, class BOp, class Epi>
__global__ void __launch_bounds__(THREADS, 1) tcgemm_kernel(const AOp a, const BOp b, const Epi epi, int M, int N, int K,
                                                            int kchunk, const float* __restrict__ zeros16) {
    using Plan = SmemPlan<BN, NTERMS>;
    constexpr int S = Plan::STAGES;
    using AMap = TileMap<AOp::KCONTIG, BM>;
    using BMap = TileMap<BOp::KCONTIG, BN>;
    constexpr int LA = AMap::CHUNKS / PRODUCERS;
    constexpr int LB = (BMap::CHUNKS + PRODUCERS - 1) / PRODUCERS;
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
        for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), PRODUCERS); mbar_init(empty_bar(s), 1); }
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
            AMap::rk(tid + i * PRODUCERS, a_r[i], a_k[i]);
            a_off[i] = AMap::off(a_r[i], a_k[i]);
            a_row[i] = a.row(min(m0 + a_r[i], M - 1), z);
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            const int f = tid + i * PRODUCERS;
            if (f < BMap::CHUNKS) {
                BMap::rk(f, b_r[i], b_k[i]);
                b_off[i] = BMap::off(b_r[i], b_k[i]);
            } else { b_r[i] = -1; b_k[i] = 0; b_off[i] = 0; }
            b_row[i] = b.row(min(n0 + max(b_r[i], 0), N - 1), z);
        }
        auto issue = [&](int kb) {
            const int s = kb % S, k0 = kbeg + kb * BK;
            const uint32_t ta = stage_a(s), tb = stage_b(s);
#pragma unroll
            for (int i = 0; i < LA; ++i) {
                const int r = m0 + a_r[i], k = k0 + a_k[i];
                const float* src = (r < M && k < kend) ? a.addr4(a_row[i], r, k, z) : nullptr;
                cp_async16(ta + a_off[i], src ? src : zeros16, src != nullptr);
            }
#pragma unroll
            for (int i = 0; i < LB; ++i) {
                if (b_r[i] >= 0) {
                    const int r = n0 + b_r[i], k = k0 + b_k[i];
                    const float* src = (r < N && k < kend) ? b.addr4(b_row[i], r, k, z) : nullptr;
                    cp_async16(tb + b_off[i], src ? src : zeros16, src != nullptr);
                }
            }
        };
        auto split_lo = [&](uint32_t hi_addr, uint32_t lo_addr) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(hi_addr));
            // the tensor core reads the upper 19 bits of each operand (tf32); what it drops goes into the lo tile
            v.x -= __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            v.y -= __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            v.z -= __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            v.w -= __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        };
        for (int kb = 0; kb < S - 1; ++kb) {
            if (kb < nkb) issue(kb);
            cp_async_commit();
        }
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S;
            cp_async_wait<S - 2>();  // this thread's chunks of block kb have landed
            if (NTERMS == 3) {
#pragma unroll
                for (int i = 0; i < LA; ++i) split_lo(stage_a(s) + a_off[i], stage_alo(s) + a_off[i]);
#pragma unroll
                for (int i = 0; i < LB; ++i)
                    if (b_r[i] >= 0) split_lo(stage_b(s) + b_off[i], stage_blo(s) + b_off[i]);
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
            mbar_arrive(full_bar(s));
            const int nxt = kb + S - 1;
            if (nxt < nkb) {
                if (kb >= 1) mbar_wait(empty_bar((kb - 1) % S), ((kb - 1) / S) & 1);  // MMAs of block kb-1 are done with the stage
                issue(nxt);
            }
            cp_async_commit();
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
        const uint32_t idesc = make_idesc(BN, !AOp::KCONTIG, !BOp::KCONTIG);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S;
            mbar_wait(full_bar(s), (kb / S) & 1);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint64_t da = AMap::desc(stage_a(s), ks), db = BMap::desc(stage_b(s), ks);
                    if (NTERMS == 3) {
                        const uint64_t dal = AMap::desc(stage_alo(s), ks), dbl = BMap::desc(stage_blo(s), ks);
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
