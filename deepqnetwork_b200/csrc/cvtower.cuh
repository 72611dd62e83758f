// The whole conv stack of a tower (conv1 -> conv2 -> conv3, rl/deep_q_model.py:292-309) as ONE image-resident kernel.
//
// Why (profiles/README.md r02): as three launches of cvconv_kernel the conv forward is a 50 us chain of
// latency-bound kernels (18 + 15 + 24 us for 96 one-CTA images whose k-loops add up to ~20 us): every layer pays a launch
// ramp, a prologue (barriers, TMEM allocation, tables), an image load from L2 and a drained pipeline, and conv1 reads an
// f32 copy of the batch that the gather has to write first.  Here one CTA walks ONE image through all three layers:
//   * layer 0 reads the uint8 frame stack itself (8-byte cp.async into a zero-padded byte image; a k-block = one kernel
//     row = 32 contiguous bytes per output pixel, two 128-bit shared loads), cast / 255 (deep_q_model.py:226-227, exactly
//     rounded) happens in the A converter -- the f32 copy of the batch is not needed any more;
//   * the epilogue of layer l writes bias + ReLU both to global memory (the backward pass needs the activations) and, as
//     the zero-padded, stride-phase-split input image of layer l+1, to shared memory: activations never make the round
//     trip through L2 between layers;
//   * the packed weights of the three layers stream through ONE TMA ring whose producer runs ahead across the layer
//     boundaries; barriers, tensor memory and the converter / MMA / TMA roles are set up once.
//   * layer 0 multiplies the frame BYTES (exact in bf16) with the three bf16 pieces of w / 255 (cvpack.cuh): 6 MMAs per
//     k-block, no cast / 255 and no hi / lo split in the converter; layers 1, 2 use the mixed 3-term product (8 MMAs);
//   * SEVERAL MMA-issuing warps.  One thread issues a tcgen05.mma every ~53 cycles whatever its shape
//     (scratch/ubench/mma_issue.cu: N = 32 / 64 MMAs from one thread 53 cycles each, from two threads 26.6 / 32.3, from four
//     16.1 / 32.2 = the tensor pipe's own N / 2), so a single issuer leaves the pipe 40-70 % idle at N <= 64.  Layer 0 (four
//     bands of 128 output pixels, N = 32): the k-blocks are walked band-interleaved and issuing warp w owns band w (its own
//     accumulator); layers 1, 2 (one band, N = 64): warp w takes the k-blocks kb % 4 == w, each into its OWN accumulator
//     (deterministic order inside each), and the epilogue adds the four in index order.
// Operands and converter roles are those of cvconv_kernel (cvgemm.cuh).
//
// Shared memory (C2: 89 x 80 x 4 frames): ring 4 x 16 KB | barriers, tables 8 KB | image A 80.4 KB (conv2 input) |
// image B 51.9 KB (the uint8 frames, 31.5 KB; re-used as conv3's input once conv1 is done) = 205 KB.
// Tensor memory: 256 accumulator columns (4 bands x 32 or 4 k-quarters x 64) + 4 A stages x 64 columns = 512.
#pragma once
#include "cvgemm.cuh"

namespace cv {

struct TowerSet {  // one tower's operands
    const uint8_t* in;       // [images][ih][iw][4] uint8
    const float* packed[3];  // packed conv weights (pack_all_kernel)
    const float* bias[3];
    float* act[3];           // post-ReLU NHWC outputs
    int act_rows[3];         // ... written for the first act_rows[l] images only (the backward pass reads the training rows)
    float* pk;               // conv3 output as the dense forward's operand tiles (EpiBiasReluPack layout), or null
    int pk_rows;
};

struct TowerArgs {
    TowerSet a, b;
    int images_a;
    ImgGeom g[3];
    int img_off[3];   // byte offset of layer l's input image inside the image area
    int img_zero[3];  // bytes of that image (zero-filled before the previous layer's epilogue writes into it)
    int u8_row;       // bytes per padded row of the layer-0 byte image
    int rt0[3];       // first row-table band of layer l
    int ext2;         // layer 2 streams its weights through 4 extra stages in image A (free during its k-loop)
};

constexpr int TW_THREADS = 544;     // warps 0-7 converters / epilogue, warps 8-11 MMA issuers, warp 12 TMA, warps 13-16 writers
constexpr int TW_S = 4;             // ring depth (B tiles in shared memory, A tiles in tensor memory).  == the number of issuing warps:
                                    // warp i (layer 0) / warp i % 2 (layers 1, 2) waits on the stages s % 4 == i / s % 2 == i only and so
                                    // sees every phase of their barriers (a waiter that skips phases can mistake an older phase of the
                                    // same parity for the one it wants)
constexpr int TW_STAGE = 16384;     // ring slot: one k-block of packed weights of a 64-column layer
constexpr int TW_DCOLS = 256;       // accumulator columns (4 bands x 32, or 4 k-quarters x 64); A stages follow
constexpr int TW_TAB = 8192;        // barriers 512 | k-block offsets 3 x 256 | bias 3 x 256 | row tables 6 x 128 x 8
constexpr int TW_MAXBANDS = 6;

__device__ __forceinline__ void tma_bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

template <int NTERMS>
__global__ void __launch_bounds__(TW_THREADS, 1) cvtower_kernel(const __grid_constant__ TowerArgs t, const float* __restrict__ zeros,
                                                                unsigned long long* __restrict__ dbg) {
    constexpr int ACOLS = BK * (NTERMS == 3 ? 2 : 1);
    const bool cta0 = dbg && blockIdx.x == 0;  // in-kernel timeline (dqn_debug_timeline)
#define TW_STAMP(cond, row, col) do { if (cta0 && (cond) && (row) < 128) dbg[(row) * 8 + (col)] = clock64(); } while (0)
    TW_STAMP(threadIdx.x == 0, 127, 0);
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int S = TW_S;
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t ring = sbase;
    const uint32_t bars = ring + (uint32_t)S * TW_STAGE;
    const uint32_t kbtab = bars + 512, bias_s = bars + 1280, rowtab = bars + 2048, imgs = bars + TW_TAB;
    auto afull_bar = [&](int s) { return bars + 8u * s; };          // A tile of a step is in tensor memory (4 converter warps)
    auto aempty_bar = [&](int s) { return bars + 8u * (5 + s); };   // ... and consumed (this CTA's MMAs: tcgen05.commit)
    // weight ring of layers 1, 2: slots 0-3 = the ring, slots 4-7 (layer 2 only, t.ext2) = 4 more stages in image A, which is free
    // during layer 2's k-loop.  One numbering of uses per slot across the two layers: parity of use u is u & 1.
    auto bfull_bar = [&](int s) { return bars + 8u * (40 + s); };   // weight tile landed (all slices, from every CTA of the cluster)
    auto bempty_bar = [&](int s) { return bars + 8u * (48 + s); };  // ... and consumed by EVERY CTA of the cluster (multicast commits)
    auto accum_bar = [&](int l, int b) { return bars + 8u * (20 + l * 4 + b); };  // each used once
    const uint32_t w0full = bars + 8u * 32;                          // layer 0's weights (all k-blocks) are resident
    const uint32_t tmem_slot = bars + 8u * 34;
    auto imgready_bar = [&](int l) { return bars + 8u * (35 + l); };  // layer l's output is complete in shared memory (8 epilogue warps)
    const uint32_t wdone = bars + 8u * 38;                            // the writers have copied layer 0's output out of image A
    uint32_t crank, csize;  // thread-block cluster: the CTAs of a cluster (same tower) share one weight stream by TMA multicast
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool set_b = (int)blockIdx.x >= t.images_a;
    const int image = set_b ? blockIdx.x - t.images_a : blockIdx.x;
    const TowerSet& ts_ = set_b ? t.b : t.a;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(afull_bar(s), 4);
            mbar_init(aempty_bar(s), 1);
        }
        for (int s = 0; s < 8; ++s) {
            mbar_init(bfull_bar(s), 1);
            mbar_init(bempty_bar(s), csize);
        }
        mbar_init(w0full, 1);
        for (int l = 0; l < 3; ++l) mbar_init(imgready_bar(l), 8);
        mbar_init(wdone, 4);
        // a band's accumulator(s) are complete: one issuing warp per band (4 bands), or the two k-halves of a single band
        for (int l = 0; l < 3; ++l)
            for (int b = 0; b < 4; ++b) mbar_init(accum_bar(l, b), t.g[l].nbands == 1 ? 4 : 1);
        fence_barrier_init();
    }
    if (tid < 192) {
        // byte offset of k-block kb's 32 K elements from the pixel a thread owns.  Layer 0 (4 channels, 8 x 8 kernel): a
        // k-block is kernel row kb of the byte image; deeper layers: one tap, 32 channels (cvconv_kernel)
        const int l = tid >> 6, kb = tid & 63;
        const ImgGeom& g = t.g[l];
        if (kb < g.nkb) {
            int off;
            if (l == 0) off = kb * t.u8_row;
            else {
                const int kk = kb * 32, tap = kk / g.cin, c = kk - tap * g.cin;
                const int kh = tap / g.k, kw = tap - kh * g.k;
                off = (kh * g.wcols + (kw % g.s) * g.wps + kw / g.s) * g.pitch + c * 4;
            }
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(kbtab + 4u * tid), "r"(off) : "memory");
        }
    }
    for (int l = 0; l < 3; ++l) {
        const ImgGeom& g = t.g[l];
        for (int e = tid; e < g.nbands * 128; e += TW_THREADS) {
            const int band = e >> 7, m = e & 127;
            const int a0 = band * g.th, na = min(g.th, g.oh - a0);
            const int ai = m / g.ow, bi = m - ai * g.ow;
            // {byte offset of the row's input pixel, where its output goes}: layers 0, 1 -> byte offset of the pixel in the next
            // layer's resident image; layer 2 -> output pixel index; -1: rows past the band (any finite data, never stored)
            int off = 0, dst = -1;
            if (ai < na) {
                const int a = a0 + ai;
                off = l == 0 ? a * g.s * t.u8_row + bi * g.s * g.cin : (a * g.s * g.wcols + bi) * g.pitch;
                dst = a * g.ow + bi;
                if (l < 2) {
                    const ImgGeom& n = t.g[l + 1];
                    const int r = a + n.pt, col = bi + n.pl;
                    dst = (r * n.wcols + (col % n.s) * n.wps + col / n.s) * n.pitch;
                }
            }
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rowtab + 8u * ((t.rt0[l] + band) * 128 + m)), "r"(off), "r"(dst) : "memory");
        }
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    if (csize > 1) {  // every CTA's barriers exist before a peer multicasts into them
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    auto tmem_a = [&](int s) { return tmem_base + (uint32_t)(TW_DCOLS + s * ACOLS); };
    auto stage_b = [&](int s) { return s < S ? ring + (uint32_t)s * TW_STAGE : imgs + t.img_off[1] + (uint32_t)(s - S) * TW_STAGE; };
    // weight-ring slot and use number of step j (layer-local) of layer l (1 or 2)
    const int n1 = t.g[1].nkb * t.g[1].nbands, nst2 = t.ext2 ? 2 * S : S;
    auto b_slot = [&](int l, int j) { return l == 1 ? j % S : j % nst2; };
    auto b_use = [&](int l, int j) { return l == 1 ? j / S : (j % nst2 < S ? n1 / S : 0) + j / nst2; };
    TW_STAMP(tid == 0, 127, 1);

    if (warp < 8) {
        auto zero_fill = [&](uint32_t base, int bytes) {
            for (int o = tid * 16; o < bytes; o += 256 * 16)
                asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + (uint32_t)o), "r"(0) : "memory");
        };
        pdl_wait();  // the batch and the packed weights are written by stream predecessors
        TW_STAMP(tid == 0, 127, 2);
        {
            // ---- the uint8 frames -> zero-padded byte image (8-byte cp.async; pl * cin and iw * cin are multiples of 8) ----
            const ImgGeom& g = t.g[0];
            const uint32_t img = imgs + t.img_off[0];
            const uint8_t* src_img = ts_.in + (size_t)image * g.ih * g.iw * g.cin;
            const int cpr = t.u8_row / 8, first = g.pl * g.cin / 8;
            const int ncopy = min(g.iw * g.cin, t.u8_row - g.pl * g.cin) / 8;
            for (int q = tid; q < g.hp * cpr; q += 256) {
                const int iyp = q / cpr, c = q - iyp * cpr;
                const int iy = iyp - g.pt, cc = c - first;
                const bool ok = (unsigned)iy < (unsigned)g.ih && (unsigned)cc < (unsigned)ncopy;
                const void* src = ok ? (const void*)(src_img + (size_t)iy * g.iw * g.cin + cc * 8) : (const void*)zeros;
                const uint32_t n = ok ? 8u : 0u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(img + (uint32_t)(iyp * t.u8_row + c * 8)), "l"(src), "r"(n) : "memory");
            }
            if (tid < 192) {  // biases of the three layers
                const int l = tid >> 6, n = tid & 63;
                if (n < t.g[l].cout) asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4u * tid), "f"(ts_.bias[l][n]) : "memory");
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        TW_STAMP(tid == 0, 127, 3);
        const int grp = warp & 3, cg = warp >> 2;  // TMEM lane quarter, converter group (alternate k-blocks)
        const int m = grp * 32 + lane;             // row of the band tile
        const uint32_t lane_base = (uint32_t)(grp * 32) << 16;

        // ---- epilogue of one band: bias + ReLU -> global NHWC, -> next layer's image in shared memory / dense operand tiles ----
        auto epilogue = [&](int l, int band) {
            const ImgGeom& g = t.g[l];
            const int BN = g.cout, HC = BN >> 1;
            mbar_wait(accum_bar(l, band), 0);
            tc_fence_after();
            int opix;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(opix) : "r"(rowtab + 8u * ((t.rt0[l] + band) * 128 + m) + 4u));
            const bool ksplit = g.nbands == 1;  // two accumulators (even / odd k-blocks) to add
            const uint32_t trow = tmem_base + lane_base + (uint32_t)(band * BN + cg * HC);
            TW_STAMP(tid == 0, 100 + l * 4 + band, 0);
            // layers 0, 1: the row table holds the pixel's place in the next layer's image; the last layer's output is staged in
            // image A (dead by now) for the writer warps, one pixel every cout * 4 + 16 bytes
            const bool to_smem = opix >= 0;
            const uint32_t sdst = l == 2 ? imgs + t.img_off[1] + (uint32_t)(opix * (BN * 4 + 16)) : imgs + t.img_off[l + 1] + (uint32_t)opix;
            for (int c = 0; c < HC; c += 16) {
                float v[16];
                const uint32_t t0 = trow + (uint32_t)c;
                if (ksplit) tmem_ld16_sum4(t0, t0 + BN, t0 + 2 * BN, t0 + 3 * BN, v);  // the four issuing warps' accumulators (k-blocks kb % 4 == a), index order
                else tmem_ld16(t0, v);
                TW_STAMP(tid == 0, 100 + l * 4 + band, 1 + (c >> 4));
                if (opix < 0) continue;
                const int n0 = cg * HC + c;
                const uint32_t sb = bias_s + 4u * (l * 64 + n0);
                const int k0 = opix * BN + n0;  // flat index (tf.layers.flatten of NHWC) of this run of 16 channels
                // the dense forward's B operand (fcfwd.cuh): k-block k0 / 32 of pk_rows batch rows, packed layout of cvpack.cuh
                uint8_t* blk = (l == 2 && ts_.pk) ? reinterpret_cast<uint8_t*>(ts_.pk) + (size_t)(k0 >> 5) * ts_.pk_rows * 256 : nullptr;
#pragma unroll
                for (int j = 0; j < 16; j += 8) {
                    float4 b0, b1;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w) : "r"(sb + 4u * j));
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b1.x), "=f"(b1.y), "=f"(b1.z), "=f"(b1.w) : "r"(sb + 4u * j + 16u));
                    const float4 r0 = make_float4(fmaxf(v[j] + b0.x, 0.f), fmaxf(v[j + 1] + b0.y, 0.f), fmaxf(v[j + 2] + b0.z, 0.f), fmaxf(v[j + 3] + b0.w, 0.f));
                    const float4 r1 = make_float4(fmaxf(v[j + 4] + b1.x, 0.f), fmaxf(v[j + 5] + b1.y, 0.f), fmaxf(v[j + 6] + b1.z, 0.f), fmaxf(v[j + 7] + b1.w, 0.f));
                    if (to_smem) {
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sdst + 4u * (n0 + j)), "f"(r0.x), "f"(r0.y), "f"(r0.z), "f"(r0.w) : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sdst + 4u * (n0 + j + 4)), "f"(r1.x), "f"(r1.y), "f"(r1.z), "f"(r1.w) : "memory");
                    }
                    if (blk) pk_store8(blk, ts_.pk_rows, image, (k0 & 31) + j, r0, r1);
                }
            }
            tc_fence_before();
            TW_STAMP(tid == 0, 100 + l * 4 + band, 3);
        };

        int step = cg, step_end = 0;
        // the converter loop of one layer (its k-blocks band-interleaved: step = kb * nbands + band); U8: layer 0 (byte image)
        auto convert_layer = [&](auto u8tag, int l) {
            constexpr bool U8 = decltype(u8tag)::value;
            const ImgGeom& g = t.g[l];
            const int nb = g.nbands;  // 1 or 4
            const uint32_t ibase = imgs + t.img_off[l];
            const int step0 = step_end;
            step_end += g.nkb * nb;
            for (; step < step_end; step += 2) {
                const int ls = step - step0;
                const int kb = nb == 1 ? ls : ls >> 2, band = nb == 1 ? 0 : ls & 3;
                const int s = step % S, ph = (step / S) & 1;
                uint32_t roff, koff;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(roff) : "r"(rowtab + 8u * ((t.rt0[l] + band) * 128 + m)));
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(koff) : "r"(kbtab + 4u * (l * 64 + kb)));
                if (step >= S) mbar_wait(aempty_bar(s), ph ^ 1);  // TMEM stage s is free again
                TW_STAMP(grp == 0 && lane == 0, step, 0);
                const uint32_t src = ibase + roff + koff;
                const uint32_t ta = tmem_a(s) + lane_base;
                if constexpr (U8) {
                    // the 32 frame bytes of this kernel row as 16 bf16 pairs (integers 0..255 are exact in bf16)
                    uint32_t x[16];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t w[4];
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(src + 16u * c));
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            // byte b -> float: 0x4B0000bb is 2^23 + b; minus 2^23 leaves b exactly (full-rate ops, no I2F);
                            // its low 16 bits are zero, so the bf16 is the upper half-word
                            const float f0 = __uint_as_float(__byte_perm(w[q], 0x4B000000u, 0x7540)) - 8388608.0f;
                            const float f1 = __uint_as_float(__byte_perm(w[q], 0x4B000000u, 0x7541)) - 8388608.0f;
                            const float f2 = __uint_as_float(__byte_perm(w[q], 0x4B000000u, 0x7542)) - 8388608.0f;
                            const float f3 = __uint_as_float(__byte_perm(w[q], 0x4B000000u, 0x7543)) - 8388608.0f;
                            x[8 * c + 2 * q] = __byte_perm(__float_as_uint(f0), __float_as_uint(f1), 0x7632);
                            x[8 * c + 2 * q + 1] = __byte_perm(__float_as_uint(f2), __float_as_uint(f3), 0x7632);
                        }
                    }
                    tmem_st16u(ta, x);
                } else {
                    float hi[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(hi[4 * c]), "=f"(hi[4 * c + 1]), "=f"(hi[4 * c + 2]), "=f"(hi[4 * c + 3]) : "r"(src + 16u * c));
                    if (NTERMS == 3) a_stage_store3(ta, hi); else ts::tmem_st32(ta, hi);
                }
                ts::tmem_st_wait();
                tc_fence_before();
                TW_STAMP(grp == 0 && lane == 0, step, 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(afull_bar(s));
            }
        };
        for (int l = 0; l < 3; ++l) {
            if (l == 0) convert_layer(std::true_type{}, l); else convert_layer(std::false_type{}, l);
            if (l == 2) pdl_launch_dependents();
            if (l == 0) {
                // all MMAs of layer 0 have completed => its weights (resident in image A) and the frame bytes (image B) are
                // dead: the two areas become the zero-padded inputs of conv2 / conv3, whose interiors the epilogues fill
                for (int band = 0; band < t.g[0].nbands; ++band) mbar_wait(accum_bar(0, band), 0);
                zero_fill(imgs + t.img_off[1], t.img_zero[1]);
                zero_fill(imgs + t.img_off[2], t.img_zero[2]);
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            if (l == 2) mbar_wait(wdone, 0);  // (long since) the writers are done with layer 0's output in image A: it stages layer 2's
            for (int band = 0; band < t.g[l].nbands; ++band) epilogue(l, band);
            __syncwarp();
            if (lane == 0) mbar_arrive(imgready_bar(l));  // the writer warps copy this layer's output to global memory
            TW_STAMP(tid == 0, 120 + l, 0);
            if (l < 2) asm volatile("bar.sync 1, 256;" ::: "memory");  // next layer's image complete, accumulators drained
        }
    } else if (warp < 12) {
        // ---------------- MMA issuers (warps 8-11) ----------------
        // a layer of 4 bands: warp w owns band w (steps kb * 4 + w); a layer of one band: warp w owns the k-blocks kb % 4 == w,
        // each with its own accumulator; every warp walks ITS steps in order (per issuing warp a step costs ~8 x 53 cycles of
        // issue + ~500 of barrier waits and commits: four of them keep the tensor pipe, 8 x 32 cycles per step, busy)
        const int w = warp - 8;
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        int step0 = 0;  // first step of the layer
        for (int l = 0; l < 3; ++l) {
            const int BN = t.g[l].cout, nkb = t.g[l].nkb, nb = t.g[l].nbands;
            constexpr int stride = 4;  // distance between this warp's steps (== TW_S: the warp owns stage w)
            {
                const uint32_t idesc = make_idesc(BN, false, false), idesc16 = make_idesc_bf16(BN);
                const uint32_t tmem_d = tmem_base + (uint32_t)(w * BN);
                const int last = step0 + nkb * nb;
                if (l == 0) mbar_wait(w0full, 0);
                for (int step = step0 + w; step < last; step += stride) {
                    const int s = step % S, ph = (step / S) & 1;
                    const int bs = l > 0 ? b_slot(l, step - step0) : 0, bph = l > 0 ? b_use(l, step - step0) & 1 : 0;
                    mbar_wait(afull_bar(s), ph);
                    if (l > 0) mbar_wait(bfull_bar(bs), bph);
                    TW_STAMP(lane == 0, step, 3);
                    tc_fence_after();
                    const uint32_t ta = tmem_a(s);
                    const bool first = step < step0 + stride;
                    if (leader) {
                        if (l == 0) {
                            // frame bytes (exact bf16) x the bf16 pieces of w / 255, smallest piece first: 2 x NP instructions
                            constexpr int NP = NTERMS == 3 ? 3 : 2;
                            const uint32_t wk = imgs + t.img_off[1] + (uint32_t)((step - step0) >> 2) * (192u * BN);  // k-block's weights
#pragma unroll
                            for (int p = NP - 1; p >= 0; --p) {
                                const uint64_t d = make_desc(wk + (uint32_t)(p * 64 * BN), PK_LBO, PK_SBO16);
#pragma unroll
                                for (int j = 0; j < 2; ++j)
                                    mma_bf16_ts(tmem_d, ta + j * 8, d + j * ((uint64_t)(2 * PK_LBO) >> 4), idesc16, !(first && p == NP - 1 && j == 0));
                            }
                        } else if (NTERMS == 3) mma_kblock3(tmem_d, ta, stage_b(bs), BN, idesc, idesc16, first);
                        else {
                            const uint64_t db = make_desc(stage_b(bs), PK_LBO, PK_SBO);
#pragma unroll
                            for (int ks = 0; ks < BK / 8; ++ks)
                                ts::mma_tf32_ts(tmem_d, ta + ks * 8, db + ((uint64_t)(ks * 2 * PK_LBO) >> 4), idesc, !(first && ks == 0));
                        }
                        mma_commit(aempty_bar(s));
                        if (l > 0) {  // the weight stage is free once EVERY CTA of the cluster has consumed it
                            if (csize > 1) mma_commit_mc(bempty_bar(bs), cmask); else mma_commit(bempty_bar(bs));
                        }
                        if (step + stride >= last) mma_commit(accum_bar(l, nb == 1 ? 0 : w));
                    }
                    TW_STAMP(lane == 0, step, 4);
                    __syncwarp();
                }
            }
            step0 += nkb * nb;
        }
        tc_fence_before();
    } else if (warp >= 13) {
        // ---------------- writers (warps 13-16): layer outputs shared memory -> global NHWC, whole 128-byte lines, off the critical
        // path (layer l's output is layer l+1's resident input image; the last layer's is staged in image A) ----------------
        const int wt = tid - 13 * 32;
        for (int l = 0; l < 3; ++l) {
            const ImgGeom& g = t.g[l];
            mbar_wait(imgready_bar(l), 0);
            if (image < ts_.act_rows[l]) {
                const int cpp = g.cout >> 2, npix = g.oh * g.ow;  // 16-byte chunks per pixel
                float* og = ts_.act[l] + (size_t)image * npix * g.cout;
                for (int q = wt; q < npix * cpp; q += 128) {
                    const int p = q / cpp, c = q - p * cpp;
                    uint32_t src;
                    if (l == 2) src = imgs + t.img_off[1] + (uint32_t)(p * (g.cout * 4 + 16));
                    else {
                        const ImgGeom& n = t.g[l + 1];
                        const int a = p / g.ow, b = p - a * g.ow;
                        const int r = a + n.pt, col = b + n.pl;
                        if (r >= n.hp || col >= n.wcols) continue;  // (never read by the next layer: not resident) -- cannot happen for SAME
                        src = imgs + t.img_off[l + 1] + (uint32_t)((r * n.wcols + (col % n.s) * n.wps + col / n.s) * n.pitch);
                    }
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src + 16u * c));
                    *reinterpret_cast<float4*>(og + (size_t)p * g.cout + c * 4) = v;
                }
            }
            if (l == 0) { __syncwarp(); if (lane == 0) mbar_arrive(wdone); }
        }
    } else {
        // ---------------- weight TMA producer (warp 12, one lane): runs ahead across the layer boundaries ----------------
        if (lane == 0) {
            pdl_wait();
            {
                // layer 0: all k-blocks (three bf16 pieces each) become resident in image A; the CTAs of a cluster fetch
                // alternate k-blocks and multicast them
                const uint32_t stride = (uint32_t)t.g[0].cout * 256u, bytes = stride / 4 * 3;
                mbar_expect_tx(w0full, bytes * t.g[0].nkb);
                for (int kb = crank; kb < t.g[0].nkb; kb += csize) {
                    const uint32_t dst = imgs + t.img_off[1] + (uint32_t)kb * bytes;
                    const float* src = ts_.packed[0] + (size_t)kb * (stride / 4);
                    if (csize > 1) tma_bulk_g2s_mc(dst, src, bytes, w0full, cmask); else tma_bulk_g2s(dst, src, bytes, w0full);
                }
            }
            for (int l = 1; l < 3; ++l) {
                const uint32_t bytes = (uint32_t)t.g[l].cout * 256u;  // one k-block of packed weights
                const uint32_t slice = bytes / csize;                 // this CTA fetches slice `crank` for the whole cluster
                const float* src = ts_.packed[l];
                for (int j = 0; j < t.g[l].nkb * t.g[l].nbands; ++j) {
                    const int s = b_slot(l, j), u = b_use(l, j);
                    if (u >= 1) mbar_wait(bempty_bar(s), (u - 1) & 1);
                    else if (s >= S) {
                        // first use of an extra stage: image A is free once every MMA of layer 1 has completed (its converters have
                        // read their last pixel) and the writer warps have copied layer 0's output out of it
                        mbar_wait(accum_bar(1, 0), 0);
                        mbar_wait(wdone, 0);
                    }
                    mbar_expect_tx(bfull_bar(s), bytes);
                    const uint32_t dst = stage_b(s) + crank * slice;
                    const float* sp = src + (size_t)(j / t.g[l].nbands) * (bytes / 4) + (size_t)crank * (slice / 4);
                    if (csize > 1) tma_bulk_g2s_mc(dst, sp, slice, bfull_bar(s), cmask); else tma_bulk_g2s(dst, sp, slice, bfull_bar(s));
                }
            }
        }
    }
    __syncthreads();
    if (csize > 1) {  // no CTA leaves while a peer may still signal its barriers
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
    TW_STAMP(tid == 0, 127, 4);
#undef TW_STAMP
}

// geometry of the fused tower for this network, or false if a layer is outside its coverage (the per-layer kernels run then)
static inline bool tower_plan(const ConvGeom* conv, TowerArgs& t, size_t& smem_bytes) {
    int bands = 0;
    for (int l = 0; l < 3; ++l) {
        ImgGeom& q = t.g[l];
        // 4 bands: one issuing warp per band; 1 band: two issuing warps, each >= 1 k-block and its own accumulator
        if (!img_geom(conv[l], q) || (q.nbands != 1 && q.nbands != 4) || q.nbands * q.cout > TW_DCOLS || q.nkb < 2) return false;
        if (q.nbands == 1 && (4 * q.cout > TW_DCOLS || q.nkb < 4)) return false;
        if (l > 0 && (conv[l].cin % 32 != 0 || conv[l].cin != conv[l - 1].cout || conv[l].ih != conv[l - 1].oh || conv[l].iw != conv[l - 1].ow)) return false;
        t.rt0[l] = bands;
        bands += q.nbands;
    }
    if (bands > TW_MAXBANDS) return false;
    for (int l = 1; l < 3; ++l)  // layer l-1's whole output is resident as layer l's image (the writer warps copy it from there)
        if (t.g[l].hp < conv[l].ih + conv[l].pt || t.g[l].wcols < conv[l].iw + conv[l].pl) return false;
    if ((t.g[0].nbands * t.g[0].nkb) % TW_S != 0 || (t.g[1].nbands * t.g[1].nkb) % TW_S != 0) return false;  // issuing warp w keeps stage w
    const ConvGeom& g0 = conv[0];
    if (g0.cin != 4 || g0.k != 8 || g0.s * g0.cin != 16) return false;  // a k-block = one kernel row = 32 bytes, pixels 16 bytes apart
    t.u8_row = ((g0.ow - 1) * g0.s + g0.k) * g0.cin;
    if (t.u8_row % 16 != 0 || (g0.pl * g0.cin) % 8 != 0 || (g0.iw * g0.cin) % 8 != 0) return false;
    const int img0 = ((t.g[0].hp * t.u8_row + 127) / 128) * 128;
    const int imgA = t.g[1].img_bytes, imgB = img0 > t.g[2].img_bytes ? img0 : t.g[2].img_bytes;
    t.img_off[0] = imgA; t.img_off[1] = 0; t.img_off[2] = imgA;
    t.img_zero[0] = 0; t.img_zero[1] = t.g[1].img_bytes; t.img_zero[2] = t.g[2].img_bytes;
    if (t.g[0].nkb * t.g[0].cout * 192 > imgA) return false;  // layer 0's weights are resident in image A during its k-loop
    t.ext2 = imgA >= TW_S * TW_STAGE && t.g[2].nbands == 1 && t.g[1].nbands == 1;  // (launch_tower clears it for clusters)
    smem_bytes = (size_t)TW_S * TW_STAGE + TW_TAB + imgA + imgB + 1024;
    static_assert(TW_DCOLS + TW_S * 64 <= 512, "tensor memory budget");
    return smem_bytes <= 227 * 1024;
}

template <int NTERMS>
static void launch_tower(dqn_learner* h, const TowerArgs& t, size_t smem_bytes, int images) {
    auto kern = cvtower_kernel<NTERMS>;
    static size_t configured = 0;
    if (smem_bytes > configured) {
        DQN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        configured = smem_bytes;
    }
    // clusters of up to 4 images of the same tower share the weight stream (TMA multicast): L2 -> SM traffic / 4
    int csize = 1;
    const int images_b = images - t.images_a;
    for (int c = h->tower_cluster; c > 1; c >>= 1)
        if (t.images_a % c == 0 && images_b % c == 0 && (64 * 256) % c == 0) { csize = c; break; }
    TowerArgs ta = t;
    if (csize > 1 || !h->tower_ext2) ta.ext2 = 0;  // a peer's image A is not known to be free when this CTA's is
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(images); cfg.blockDim = dim3(TW_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = h->stream;
    cudaLaunchAttribute at[3];
    int n = 0;
    at[n].id = cudaLaunchAttributePriority; at[n].val.priority = g_launch_priority; ++n;
    if (h->pdl) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
    at[n].id = cudaLaunchAttributeClusterDimension; at[n].val.clusterDim.x = csize; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1; ++n;
    cfg.attrs = at; cfg.numAttrs = n;
    DQN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, (const float*)h->zeros16, h->tc_dbg));
    h->launches++;
}

}  // namespace cv
