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
// Operands, roles and the 3-term TF32 split are those of cvconv_kernel (cvgemm.cuh); results are bit-identical to the
// per-layer kernels (same products, same accumulation order).
//
// Shared memory (C2: 89 x 80 x 4 frames): ring 5 x 16 KB | barriers, tables 8 KB | image A 80.4 KB (conv2 input) |
// image B 51.9 KB (the uint8 frames, 31.5 KB; re-used as conv3's input once conv1 is done) = 221 KB.
#pragma once
#include "cvgemm.cuh"

namespace cv {

struct TowerSet {  // one tower's operands
    const uint8_t* in;       // [images][ih][iw][4] uint8
    const float* packed[3];  // packed conv weights (pack_all_kernel)
    const float* bias[3];
    float* act[3];           // post-ReLU NHWC outputs
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
    int S;            // ring depth (B tiles in shared memory, A tiles in tensor memory)
};

constexpr int TW_THREADS = 320;     // warps 0-7 converters / epilogue, warp 8 MMA, warp 9 TMA
constexpr int TW_STAGE = 16384;     // ring slot: hi + lo tile of a 64-column layer
constexpr int TW_DCOLS = 128;       // accumulator columns (bands x cout <= 128), A stages follow
constexpr int TW_TAB = 8192;        // barriers 256 | k-block offsets 3 x 256 | bias 3 x 256 | row tables 6 x 128 x 8
constexpr int TW_MAXBANDS = 6;

template <int NTERMS>
__global__ void __launch_bounds__(TW_THREADS, 1) cvtower_kernel(const __grid_constant__ TowerArgs t, const float* __restrict__ zeros,
                                                                unsigned long long* __restrict__ dbg) {
    constexpr int ACOLS = BK * (NTERMS == 3 ? 2 : 1);
    const bool cta0 = dbg && blockIdx.x == 0;  // in-kernel timeline (dqn_debug_timeline)
#define TW_STAMP(cond, row, col) do { if (cta0 && (cond) && (row) < 128) dbg[(row) * 8 + (col)] = clock64(); } while (0)
    TW_STAMP(threadIdx.x == 0, 127, 0);
    extern __shared__ __align__(1024) uint8_t smem[];
    const int S = t.S;
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t ring = sbase;
    const uint32_t bars = ring + (uint32_t)S * TW_STAGE;
    const uint32_t kbtab = bars + 256, bias_s = bars + 1024, rowtab = bars + 2048, imgs = bars + TW_TAB;
    auto afull_bar = [&](int s) { return bars + 8u * s; };
    auto bfull_bar = [&](int s) { return bars + 8u * (6 + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (12 + s); };
    auto accum_bar = [&](int l, int b) { return bars + 8u * (18 + l * 4 + b); };  // each used once
    const uint32_t tmem_slot = bars + 248;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool set_b = (int)blockIdx.x >= t.images_a;
    const int image = set_b ? blockIdx.x - t.images_a : blockIdx.x;
    const TowerSet& ts_ = set_b ? t.b : t.a;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(afull_bar(s), 4);
            mbar_init(bfull_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 12; ++b) mbar_init(accum_bar(0, b), 1);
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
            int off = 0, opix = -1;  // rows past the band: any finite data, never stored
            if (ai < na) {
                const int a = a0 + ai;
                off = l == 0 ? a * g.s * t.u8_row + bi * g.s * g.cin : (a * g.s * g.wcols + bi) * g.pitch;
                opix = a * g.ow + bi;
            }
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rowtab + 8u * ((t.rt0[l] + band) * 128 + m)), "r"(off), "r"(opix) : "memory");
        }
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    auto tmem_a = [&](int s) { return tmem_base + (uint32_t)(TW_DCOLS + s * ACOLS); };
    auto stage_b = [&](int s) { return ring + (uint32_t)s * TW_STAGE; };
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
            zero_fill(imgs + t.img_off[1], t.img_zero[1]);  // SAME padding of conv2's input
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
            const uint32_t trow = tmem_base + lane_base + (uint32_t)(band * BN + cg * HC);
            float* og = ts_.act[l] + ((size_t)image * (g.oh * g.ow) + (opix < 0 ? 0 : opix)) * BN;
            uint32_t sdst = 0;
            bool to_smem = false;
            if (l < 2 && opix >= 0) {
                const ImgGeom& n = t.g[l + 1];
                const int a = opix / g.ow, b = opix - a * g.ow;
                const int r = a + n.pt, col = b + n.pl;
                to_smem = r < n.hp && col < n.wcols;
                const int phase = col % n.s, j = col / n.s;
                sdst = imgs + t.img_off[l + 1] + (uint32_t)((r * n.wcols + phase * n.wps + j) * n.pitch);
            }
            for (int c = 0; c < HC; c += 16) {
                float v[16];
                tmem_ld16(trow + (uint32_t)c, v);
                if (opix < 0) continue;
                const int n0 = cg * HC + c;
                const uint32_t sb = bias_s + 4u * (l * 64 + n0);
                float* pt = nullptr;
                if (l == 2 && ts_.pk) {
                    const int k0 = opix * BN + n0;  // flat index (tf.layers.flatten of NHWC) of this run of 16 channels
                    pt = ts_.pk + (size_t)(k0 >> 5) * (2 * ts_.pk_rows * 32) + (((image >> 3) * 1024 + ((k0 & 31) >> 2) * 128 + (image & 7) * 16) >> 2);
                }
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 b;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(sb + 4u * j));
                    const float4 r = make_float4(fmaxf(v[j] + b.x, 0.f), fmaxf(v[j + 1] + b.y, 0.f), fmaxf(v[j + 2] + b.z, 0.f), fmaxf(v[j + 3] + b.w, 0.f));
                    *reinterpret_cast<float4*>(og + n0 + j) = r;
                    if (to_smem)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sdst + 4u * (n0 + j)), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w) : "memory");
                    if (pt) {
                        const float4 lo = make_float4(lo_part(r.x), lo_part(r.y), lo_part(r.z), lo_part(r.w));
                        *reinterpret_cast<float4*>(pt + (j >> 2) * 32) = r;
                        *reinterpret_cast<float4*>(pt + (j >> 2) * 32 + ts_.pk_rows * 32) = lo;
                    }
                }
            }
            tc_fence_before();
        };

        int step = cg, s = cg % S, ph = 0, step_end = 0;
        // the converter loop of one band; U8: layer 0 (byte image, cast / 255 here)
        auto convert_band = [&](auto u8tag, int l, int band) {
            constexpr bool U8 = decltype(u8tag)::value;
            const ImgGeom& g = t.g[l];
            uint32_t roff;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(roff) : "r"(rowtab + 8u * ((t.rt0[l] + band) * 128 + m)));
            const uint32_t base = imgs + t.img_off[l] + roff;
            const int step0 = step_end;
            step_end += g.nkb;
            for (; step < step_end; step += 2) {
                const int kb = step - step0;
                if (step >= S) mbar_wait(empty_bar(s), ph ^ 1);  // TMEM stage s is free again
                TW_STAMP(grp == 0 && lane == 0, step, 0);
                uint32_t koff;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(koff) : "r"(kbtab + 4u * (l * 64 + kb)));
                const uint32_t src = base + koff;
                float hi[32];
                if constexpr (U8) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t w[4];
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(src + 16u * c));
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            hi[16 * c + 4 * q] = u8_to_unit(w[q] & 255u);
                            hi[16 * c + 4 * q + 1] = u8_to_unit((w[q] >> 8) & 255u);
                            hi[16 * c + 4 * q + 2] = u8_to_unit((w[q] >> 16) & 255u);
                            hi[16 * c + 4 * q + 3] = u8_to_unit(w[q] >> 24);
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(hi[4 * c]), "=f"(hi[4 * c + 1]), "=f"(hi[4 * c + 2]), "=f"(hi[4 * c + 3]) : "r"(src + 16u * c));
                }
                const uint32_t ta = tmem_a(s) + lane_base;
                ts::tmem_st32(ta, hi);
                if (NTERMS == 3) {
                    float lo[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) lo[j] = lo_part(hi[j]);
                    ts::tmem_st32(ta + BK, lo);
                }
                ts::tmem_st_wait();
                tc_fence_before();
                TW_STAMP(grp == 0 && lane == 0, step, 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(afull_bar(s));
                s += 2;
                if (s >= S) { s -= S; ph ^= 1; }
            }
        };
        for (int l = 0; l < 3; ++l) {
            const int nb = t.g[l].nbands;
            for (int band = 0; band < nb; ++band) {
                if (l == 0) convert_band(std::true_type{}, l, band); else convert_band(std::false_type{}, l, band);
                if (l == 2 && band == nb - 1) pdl_launch_dependents();
                // the previous band's accumulator drains while the tensor core works on this one
                if (band > 0) epilogue(l, band - 1);
            }
            if (l == 0) {
                // all MMAs of layer 0 have completed => every converter has read its last byte of the frame image: its
                // space becomes conv3's input (zero padding first; conv2's epilogue fills the interior)
                mbar_wait(accum_bar(0, nb - 1), 0);
                zero_fill(imgs + t.img_off[2], t.img_zero[2]);
            }
            epilogue(l, nb - 1);
            TW_STAMP(tid == 0, 120 + l, 0);
            if (l < 2) asm volatile("bar.sync 1, 256;" ::: "memory");  // next layer's image complete, accumulators drained
        }
    } else if (warp == 8) {
        // ---------------- MMA issuer ----------------
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        int nsteps = 0;
        for (int l = 0; l < 3; ++l) nsteps += t.g[l].nbands * t.g[l].nkb;
        int s = 0, ph = 0, step = 0;
        bool ready = false;
        for (int l = 0; l < 3; ++l) {
            const int BN = t.g[l].cout, nkb = t.g[l].nkb;
            const uint32_t idesc = make_idesc(BN, false, false);
            const uint32_t b_tile = (uint32_t)BN * 128u;
            for (int band = 0; band < t.g[l].nbands; ++band) {
                const uint32_t tmem_d = tmem_base + (uint32_t)(band * BN);
                for (int kb = 0; kb < nkb; ++kb, ++step) {
                    if (!ready) {
                        mbar_wait(afull_bar(s), ph);
                        mbar_wait(bfull_bar(s), ph);
                    }
                    TW_STAMP(lane == 0, step, 3);
                    tc_fence_after();
                    const int s1 = s + 1 == S ? 0 : s + 1, ph1 = s + 1 == S ? ph ^ 1 : ph;
                    ready = step + 1 < nsteps && mbar_test(afull_bar(s1), ph1) && mbar_test(bfull_bar(s1), ph1);
                    const uint64_t db = make_desc(stage_b(s), PK_LBO, PK_SBO);
                    const uint64_t dbl = make_desc(stage_b(s) + b_tile, PK_LBO, PK_SBO);
                    const uint32_t ta = tmem_a(s);
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint64_t o = (uint64_t)(ks * 2 * PK_LBO) >> 4;
                        const uint32_t acc = (kb | ks) != 0;
                        if (leader) {
                            if (NTERMS == 3) {
                                ts::mma_tf32_ts(tmem_d, ta + ks * 8, dbl + o, idesc, acc);      // A_hi * B_lo
                                ts::mma_tf32_ts(tmem_d, ta + BK + ks * 8, db + o, idesc, 1);   // A_lo * B_hi
                                ts::mma_tf32_ts(tmem_d, ta + ks * 8, db + o, idesc, 1);        // A_hi * B_hi
                            } else {
                                ts::mma_tf32_ts(tmem_d, ta + ks * 8, db + o, idesc, acc);
                            }
                        }
                    }
                    if (leader) {
                        mma_commit(empty_bar(s));
                        if (kb == nkb - 1) mma_commit(accum_bar(l, band));
                    }
                    TW_STAMP(lane == 0, step, 4);
                    __syncwarp();
                    s = s1; ph = ph1;
                }
            }
        }
        tc_fence_before();
    } else {
        // ---------------- weight TMA producer (warp 9, one lane): runs ahead across the layer boundaries ----------------
        if (lane == 0) {
            pdl_wait();
            int s = 0, ph = 0, step = 0;
            for (int l = 0; l < 3; ++l) {
                const uint32_t bytes = (uint32_t)t.g[l].cout * 256u;  // hi + lo tile of one k-block
                const float* src = ts_.packed[l];
                for (int band = 0; band < t.g[l].nbands; ++band)
                    for (int kb = 0; kb < t.g[l].nkb; ++kb, ++step) {
                        if (step >= S) mbar_wait(empty_bar(s), ph ^ 1);
                        mbar_expect_tx(bfull_bar(s), bytes);
                        tma_bulk_g2s(stage_b(s), src + (size_t)kb * (bytes / 4), bytes, bfull_bar(s));
                        if (++s == S) { s = 0; ph ^= 1; }
                    }
            }
        }
    }
    __syncthreads();
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
        if (!img_geom(conv[l], q) || q.nbands > 4 || q.nbands * q.cout > TW_DCOLS) return false;
        if (l > 0 && (conv[l].cin % 32 != 0 || conv[l].cin != conv[l - 1].cout || conv[l].ih != conv[l - 1].oh || conv[l].iw != conv[l - 1].ow)) return false;
        t.rt0[l] = bands;
        bands += q.nbands;
    }
    if (bands > TW_MAXBANDS) return false;
    const ConvGeom& g0 = conv[0];
    if (g0.cin != 4 || g0.k != 8 || g0.s * g0.cin != 16) return false;  // a k-block = one kernel row = 32 bytes, pixels 16 bytes apart
    t.u8_row = ((g0.ow - 1) * g0.s + g0.k) * g0.cin;
    if (t.u8_row % 16 != 0 || (g0.pl * g0.cin) % 8 != 0 || (g0.iw * g0.cin) % 8 != 0) return false;
    const int img0 = ((t.g[0].hp * t.u8_row + 127) / 128) * 128;
    const int imgA = t.g[1].img_bytes, imgB = img0 > t.g[2].img_bytes ? img0 : t.g[2].img_bytes;
    t.img_off[0] = imgA; t.img_off[1] = 0; t.img_off[2] = imgA;
    t.img_zero[0] = 0; t.img_zero[1] = t.g[1].img_bytes; t.img_zero[2] = t.g[2].img_bytes;
    for (int S = 6; S >= 3; --S) {
        const size_t need = (size_t)S * TW_STAGE + TW_TAB + imgA + imgB + 1024;
        if (need <= 227 * 1024 && TW_DCOLS + S * 64 <= 512) { t.S = S; smem_bytes = need; return true; }
    }
    return false;
}

template <int NTERMS>
static void launch_tower(dqn_learner* h, const TowerArgs& t, size_t smem_bytes, int images) {
    auto kern = cvtower_kernel<NTERMS>;
    static size_t configured = 0;
    if (smem_bytes > configured) {
        DQN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        configured = smem_bytes;
    }
    launch_kernel(h->pdl, kern, dim3(images), dim3(TW_THREADS), smem_bytes, h->stream, t, (const float*)h->zeros16, h->tc_dbg);
    h->launches++;
}

}  // namespace cv
