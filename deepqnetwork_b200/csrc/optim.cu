// Optimizer, step bookkeeping and target sync over the contiguous parameter slab.
//
// Replaces (reference file:line):
//   rl/deep_q_model.py:384-392  tf.train.AdamOptimizer(lr) / MomentumOptimizer(lr, momentum, use_nesterov=True)
//                               .minimize(loss, global_step=train/step)
//   rl/deep_q_model.py:258-263,160-165  copy_online_to_target (one device copy of the whole slab)
// TF kernels restated: ApplyAdam  lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g*g-v)(1-b2);
// var -= m*lr_t/(sqrt(v)+eps).  ApplyMomentum(nesterov): acc = acc*mom + g; var -= g*lr + acc*mom*lr.
// All parameters live in one slab, so "multi-tensor apply" is a single elementwise pass.
#include "learner.h"
#include "cvpack.cuh"
#include <algorithm>

// Streaming Adam over one slab range, one-wave grid-stride grid (8 CTAs per SM).  Launch shapes measured inside the
// whole step (tools/cupti_trace): this one 5090 steps/s; one 1024-float4 chunk per short CTA (yields the SMs to the
// higher-priority chain, but then this HBM-bound pass itself finishes last) 4690; the same plus a 48 KB shared-memory
// reservation to cap residency 4720.
__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v,
                                                   const float4* __restrict__ g, size_t n4, float lr,
                                                   const dqn_learner::Scalars* __restrict__ sc, int prev_betas) {
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    // prev_betas: the pass was deferred past the step bookkeeping (it belongs to the step before the last advance)
    const float b1p = prev_betas ? sc->b1p_prev : sc->b1p, b2p = prev_betas ? sc->b2p_prev : sc->b2p;
    const float lr_t = lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
    const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 gg = g[i];
        float4 mm = m[i], vv = v[i], pp = p[i];
#define ADAM1(c)                                             \
        mm.c = mm.c + (gg.c - mm.c) * omb1;                  \
        vv.c = vv.c + (gg.c * gg.c - vv.c) * omb2;           \
        pp.c = pp.c - (mm.c * lr_t) / (sqrtf(vv.c) + eps);
        ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
        m[i] = mm; v[i] = vv; p[i] = pp;
    }
}

__global__ void __launch_bounds__(256) momentum_kernel(float4* __restrict__ p, float4* __restrict__ acc,
                                                       const float4* __restrict__ g, size_t n4, float lr, float mom) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 gg = g[i];
        float4 aa = acc[i], pp = p[i];
#define MOM1(c)                                  \
        aa.c = aa.c * mom + gg.c;                \
        pp.c = pp.c - (gg.c * lr + aa.c * mom * lr);
        MOM1(x) MOM1(y) MOM1(z) MOM1(w)
#undef MOM1
        acc[i] = aa; p[i] = pp;
    }
}

static void optimizer_range(dqn_learner* h, int64_t beg, int64_t end, bool prev_betas = false) {
    if (end <= beg) return;
    const size_t n4 = (size_t)(end - beg) / 4, o4 = (size_t)beg / 4;
    // deferred pass (prev_betas): it runs under the NEXT step's conv forward, whose CTAs must get onto the SMs at once:
    // one float4 per thread, so every CTA is gone after a microsecond instead of holding its thread slots for the whole pass
    const int blocks = prev_betas ? (int)((n4 + 255) / 256) : (int)std::min<size_t>((n4 + 255) / 256, (size_t)h->sm_count * 8);
    if (h->cfg.use_momentum)
        launch_kernel(false, momentum_kernel, dim3(blocks), dim3(256), 0, h->stream, (float4*)h->online + o4, (float4*)h->slot_m + o4,
                      (const float4*)h->grads + o4, n4, (float)h->cfg.learning_rate, (float)h->cfg.momentum);
    else
        launch_kernel(false, adam_kernel, dim3(blocks), dim3(256), 0, h->stream, (float4*)h->online + o4, (float4*)h->slot_m + o4,
                      (float4*)h->slot_v + o4, (const float4*)h->grads + o4, n4, (float)h->cfg.learning_rate, h->d_sc, prev_betas ? 1 : 0);
    h->launches++;
}

// optimizer for the dense hidden kernels only (98.6% of the parameters): issued from inside the backward pass as soon
// as their gradient exists and nothing reads the old weights any more
void launch_optimizer_fc(dqn_learner* h) {
    const NetDesc& net = h->net;
    const int64_t fcw = (int64_t)net.flat * net.hidden;
    optimizer_range(h, net.off_fc_w[0], net.off_fc_w[0] + fcw);
    if (net.dueling) optimizer_range(h, net.off_fc_w[1], net.off_fc_w[1] + fcw);
}

void launch_optimizer_fc_stream(dqn_learner* h, int st, bool prev_betas) {
    const NetDesc& net = h->net;
    const int64_t fcw = (int64_t)net.flat * net.hidden;
    optimizer_range(h, net.off_fc_w[st], net.off_fc_w[st] + fcw, prev_betas);
}

// Dense hidden layers: weight gradient AND optimizer in one pass (the default for a whole step on one GPU).
//   dW[k][n] = sum_b act3[b][k] * dH[b][n]   (K = batch = 32 MACs per element: plain fp32 FFMA, exact like the reference)
// is formed in registers from shared-memory tiles and consumed on the spot by Adam / Nesterov momentum, so the 31.5 MB
// gradient of the dense kernels never makes its HBM round trip and the kernel is bound by the unavoidable
// read(w, m, v) + write(w, m, v) = 6 x 4 bytes per parameter.
// Persistent CTAs (2 per SM): a CTA owns one 256-column n tile, keeps the dH tile [32 b][256 n] resident in shared
// memory and walks 16-row k sub-tiles with a grid stride.  Per sub-tile the optimizer state of the thread's 4 x 4
// outputs (w, m, v: 12 float4) is requested FIRST, the 32-deep fmaf chains run while those loads are in flight (and
// while the act3 tile of the next sub-tile is fetched into a register), then Adam / Nesterov consumes both.  Two CTAs
// x 256 threads x 192 bytes keep ~98 KB of HBM requests in flight per SM.  WRITE_G also stores dW (parity path).
template <bool WRITE_G>
__global__ void __launch_bounds__(256, 2) fc_wgrad_adam_kernel(const float* __restrict__ act3, const float* __restrict__ dhid,
                                                            int Bn, int flat, int NH, int hid,
                                                            float* __restrict__ w0, float* __restrict__ w1,
                                                            float* __restrict__ m0, float* __restrict__ m1,
                                                            float* __restrict__ v0, float* __restrict__ v1,
                                                            float* __restrict__ g0, float* __restrict__ g1,
                                                            float lr, float mom, int use_momentum,
                                                            const dqn_learner::Scalars* __restrict__ sc) {
    __shared__ __align__(16) float sa[2][32][16];  // act3 tiles [b][k], double-buffered
    __shared__ __align__(16) float sd[32][256];    // dH tile    [b][n], resident
    const int tid = threadIdx.x;
    const int nt = blockIdx.y * 256;
    const int tn = (tid & 63) * 4, tk = (tid >> 6) * 4;
    const int st = nt >= hid;                     // a 256-wide n tile never straddles the two streams (hid % 256 == 0)
    const int ncol = nt + tn - (st ? hid : 0);
    float* W = st ? w1 : w0; float* Mo = st ? m1 : m0; float* V = st ? v1 : v0; float* G = st ? g1 : g0;
    const int nsub = flat / 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // 32 b x 64 float4 of dH
        const int f = tid + i * 256, b = f >> 6, q = f & 63;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < Bn) v = *reinterpret_cast<const float4*>(dhid + (size_t)b * NH + nt + q * 4);
        *reinterpret_cast<float4*>(&sd[b][q * 4]) = v;
    }
    const int ab = tid >> 2, aq = tid & 3;  // threads 0..127 fetch the act3 tile: 32 b x 4 float4
    auto load_a = [&](int sub) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < 128 && ab < Bn && sub < nsub) v = *reinterpret_cast<const float4*>(act3 + (size_t)ab * flat + sub * 16 + aq * 4);
        return v;
    };
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float lr_t = use_momentum ? lr : lr * sqrtf(1.0f - sc->b2p) / (1.0f - sc->b1p);
    const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
    float4 anext = load_a(blockIdx.x);
    int buf = 0;
    for (int sub = blockIdx.x; sub < nsub; sub += gridDim.x, buf ^= 1) {
        // optimizer state first: these loads fly under the fmaf chains
        float4 mm[4], pp[4], vv[4];
        size_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[i] = (size_t)(sub * 16 + tk + i) * hid + ncol;
            mm[i] = __ldcs(reinterpret_cast<const float4*>(Mo + o[i]));
            pp[i] = __ldcs(reinterpret_cast<const float4*>(W + o[i]));
            if (!use_momentum) vv[i] = __ldcs(reinterpret_cast<const float4*>(V + o[i]));
        }
        if (tid < 128) *reinterpret_cast<float4*>(&sa[buf][ab][aq * 4]) = anext;
        __syncthreads();  // sa[buf] (and, first time, sd) visible; everyone is done with sa[buf] of two iterations ago
        anext = load_a(sub + gridDim.x);
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
        for (int b = 0; b < 32; ++b) {  // ascending sample order, one fmaf chain per element (deterministic)
            const float4 a = *reinterpret_cast<const float4*>(&sa[buf][b][tk]);
            const float4 d = *reinterpret_cast<const float4*>(&sd[b][tn]);
            const float av[4] = {a.x, a.y, a.z, a.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], dv[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 gg = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            if (use_momentum) {
#define MOM1(c) mm[i].c = mm[i].c * mom + gg.c; pp[i].c = pp[i].c - (gg.c * lr + mm[i].c * mom * lr);
                MOM1(x) MOM1(y) MOM1(z) MOM1(w)
#undef MOM1
            } else {
#define ADAM1(c) mm[i].c = mm[i].c + (gg.c - mm[i].c) * omb1; vv[i].c = vv[i].c + (gg.c * gg.c - vv[i].c) * omb2; \
                 pp[i].c = pp[i].c - (mm[i].c * lr_t) / (sqrtf(vv[i].c) + eps);
                ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
                __stcs(reinterpret_cast<float4*>(V + o[i]), vv[i]);
            }
            __stcs(reinterpret_cast<float4*>(Mo + o[i]), mm[i]);
            *reinterpret_cast<float4*>(W + o[i]) = pp[i];
            if (WRITE_G) *reinterpret_cast<float4*>(G + o[i]) = gg;
        }
    }
}

bool fc_wgrad_adam_supported(const dqn_learner* h) {
    const NetDesc& net = h->net;
    return net.flat % 16 == 0 && net.hidden % 256 == 0;
}

void launch_fc_wgrad_adam(dqn_learner* h, int rows, bool write_g) {
    const NetDesc& net = h->net;
    DQN_REQUIRE(rows <= 32, "fused dense wgrad + optimizer pass: at most 32 samples");
    float *W = h->online, *M = h->slot_m, *V = h->slot_v, *G = h->grads;
    const int ntiles = net.NH / 256;
    int per_tile = (2 * h->sm_count) / ntiles;  // two resident CTAs per SM, split evenly over the n tiles
    if (per_tile < 1) per_tile = 1;
    if (per_tile > net.flat / 16) per_tile = net.flat / 16;
    dim3 grid(per_tile, ntiles);
    auto go = [&](auto kern) {
        launch_kernel(false, kern, grid, dim3(256), 0, h->stream, (const float*)h->acts_on.act[2], (const float*)h->d_hid, rows,
                      net.flat, net.NH, net.hidden, W + net.off_fc_w[0], W + net.off_fc_w[1], M + net.off_fc_w[0],
                      M + net.off_fc_w[1], V + net.off_fc_w[0], V + net.off_fc_w[1], G + net.off_fc_w[0], G + net.off_fc_w[1],
                      (float)h->cfg.learning_rate, (float)h->cfg.momentum, h->cfg.use_momentum, (const dqn_learner::Scalars*)h->d_sc);
    };
    if (write_g) go(fc_wgrad_adam_kernel<true>); else go(fc_wgrad_adam_kernel<false>);
    h->launches++;
}

// Up to three slab ranges in ONE launch, followed by the step bookkeeping (global_step += 1, beta powers,
// RNG counter: AdamOptimizer._finish / rl/deep_q_model.py:391) done by whichever CTA finishes last -- the betas are
// read by every CTA before its ticket is taken, so the update cannot race with them.
struct OptRanges { long long beg4[3], n4[3]; int count; };

// The conv kernels this launch updates are re-packed on the spot for the image-resident conv kernels (cvgemm.cuh:
// [k-block][hi, lo][rows x 32 k] tiles in the dense K-major UMMA layout; dgrad: per stride class, flipped / transposed),
// which saves the separate pack launch at the very end of every step.
struct TailPack {
    int n;
    struct L { long long w0; int KN, N, k, s, cin, cout; float* fwd; float* dg; float* u8; } l[3];
};
__device__ __forceinline__ void tail_pack4(const TailPack::L& q, int e, const float4& w) {
    const float v[4] = {w.x, w.y, w.z, w.w};
    const int kk = e / q.N, n = e - kk * q.N;  // HWIO: e = kk * cout + n, the four values are n .. n+3 of one kk
    {
        const int kb = kk >> 5, kin = kk & 31;
        uint8_t* blk = reinterpret_cast<uint8_t*>(q.fwd) + (size_t)kb * q.N * 256;
#pragma unroll
        for (int j = 0; j < 4; ++j) cv::pk_store(blk, q.N, n + j, kin, v[j]);
        if (q.u8) {
            uint8_t* b8 = reinterpret_cast<uint8_t*>(q.u8) + (size_t)kb * q.N * 256;
#pragma unroll
            for (int j = 0; j < 4; ++j) cv::pk_store_u8(b8, q.N, n + j, kin, v[j]);
        }
    }
    if (q.dg) {
        const int tap = kk / q.cin, ci = kk - tap * q.cin;
        const int kh = tap / q.k, kw = tap - kh * q.k;
        const int kt = q.k / q.s, Kc = kt * kt * q.cout;
        const int cls = (kh % q.s) * q.s + (kw % q.s);
        const int th = kt - 1 - kh / q.s, tw = kt - 1 - kw / q.s;
        const int k2 = (th * kt + tw) * q.cout + n;  // co = n .. n+3 -> four consecutive k of the dgrad operand
        const int kb = cls * (Kc >> 5) + (k2 >> 5), kin = k2 & 31;
        uint8_t* blk = reinterpret_cast<uint8_t*>(q.dg) + (size_t)kb * q.cin * 256;
        cv::pk_store4(blk, q.cin, ci, kin, w);
    }
}

__global__ void __launch_bounds__(256) optimizer_tail_kernel(float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v,
                                                             const float4* __restrict__ g, OptRanges rg, float lr, float mom,
                                                             int use_momentum, dqn_learner::Scalars* sc,
                                                             unsigned int* __restrict__ ticket, const TailPack tp,
                                                             float* host_mirror, int mirror_words, unsigned int join_expected) {
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float lr_t = use_momentum ? lr : lr * sqrtf(1.0f - sc->b2p) / (1.0f - sc->b1p);
    const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
    // the (up to three) parameter ranges as ONE index space: a thread that is active in several ranges would otherwise walk
    // their load -> update -> store latency chains one after the other (the lowest-numbered CTAs set the kernel's duration)
    const size_t n0 = rg.count > 0 ? (size_t)rg.n4[0] : 0, n1 = rg.count > 1 ? (size_t)rg.n4[1] : 0, n2 = rg.count > 2 ? (size_t)rg.n4[2] : 0;
    {
        for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n0 + n1 + n2; j += (size_t)gridDim.x * blockDim.x) {
            const int q = j < n0 ? 0 : j < n0 + n1 ? 1 : 2;
            const size_t o = (size_t)rg.beg4[q], i = j - (q == 0 ? 0 : q == 1 ? n0 : n0 + n1);
            const float4 gg = g[o + i];
            float4 mm = m[o + i], pp = p[o + i];
            if (use_momentum) {
#define MOM1(c) mm.c = mm.c * mom + gg.c; pp.c = pp.c - (gg.c * lr + mm.c * mom * lr);
                MOM1(x) MOM1(y) MOM1(z) MOM1(w)
#undef MOM1
            } else {
                float4 vv = v[o + i];
#define ADAM1(c) mm.c = mm.c + (gg.c - mm.c) * omb1; vv.c = vv.c + (gg.c * gg.c - vv.c) * omb2; \
                 pp.c = pp.c - (mm.c * lr_t) / (sqrtf(vv.c) + eps);
                ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
                v[o + i] = vv;
            }
            m[o + i] = mm; p[o + i] = pp;
            const long long e = (long long)(o + i) * 4;
            for (int l = 0; l < tp.n; ++l)
                if (e >= tp.l[l].w0 && e < tp.l[l].w0 + tp.l[l].KN) tail_pack4(tp.l[l], (int)(e - tp.l[l].w0), pp);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            // this kernel is done; join_expected = 2 when the fused dense kernel runs beside it (learner.h)
            if (step_join_arrive(ticket, join_expected)) step_bookkeeping(sc, ticket, use_momentum, host_mirror, mirror_words);
        }
    }
}

// skip_fc: the two dense hidden kernels were already updated (inside backward, or in the wgrad epilogue)
void launch_optimizer(dqn_learner* h, bool skip_fc) {
    const NetDesc& net = h->net;
    OptRanges rg; rg.count = 0;
    auto add = [&](int64_t beg, int64_t end) {
        if (end <= beg) return;
        rg.beg4[rg.count] = beg / 4; rg.n4[rg.count] = (end - beg) / 4; rg.count++;
    };
    const int64_t fcw = (int64_t)net.flat * net.hidden;
    if (!skip_fc) add(0, net.slab);
    else {
        add(0, net.off_fc_w[0]);
        if (net.dueling) { add(net.off_fc_w[0] + fcw, net.off_fc_w[1]); add(net.off_fc_w[1] + fcw, net.slab); }
        else add(net.off_fc_w[0] + fcw, net.slab);
    }
    TailPack tp; tp.n = 0;
    if (h->pack_in_tail && h->cfg.math_mode != DQN_MATH_FP32 && h->packed[0]) {
        for (int l = 0; l < 3; ++l) {
            const ConvGeom& g = net.conv[l];
            const int K = g.k * g.k * g.cin, N = g.cout;
            if (K % 32 != 0 || N % 4 != 0) { tp.n = 0; break; }
            tp.l[tp.n++] = {net.off_cw[l], K * N, N, g.k, g.s, g.cin, g.cout, h->packed[0] + h->pk_off[l],
                            (l > 0 && h->packed_dg && dgrad_pack_ok(h, l)) ? h->packed_dg + h->pkd_off[l] : nullptr,
                            l == 0 ? h->packed[0] + h->pk_off[3] : nullptr};
        }
    }
    long long most = 1;
    for (int q = 0; q < rg.count; ++q) most = std::max(most, rg.n4[q]);
    const int blocks = (int)std::min<long long>((most + 255) / 256, (long long)h->sm_count * 8);
    launch_kernel(false, optimizer_tail_kernel, dim3(blocks), dim3(256), 0, h->stream, (float4*)h->online, (float4*)h->slot_m,
                  (float4*)h->slot_v, (const float4*)h->grads, rg, (float)h->cfg.learning_rate, (float)h->cfg.momentum,
                  h->cfg.use_momentum, h->d_sc, h->opt_ticket, tp, h->tail_mirror, h->tail_mirror_words, h->rows_joined ? 2u : 1u);
    h->rows_joined = false;
    h->launches++;
    if (!tp.n) launch_pack_conv(h, 0);  // (option off: separate pack launch) the conv kernels read the weights pre-split and pre-tiled
}

void launch_tree_update(dqn_learner* h, int n, const long long* d_tree_idx, const float* d_scores, int store_losses);

// the step bookkeeping now rides on the optimizer launch; what is left is the optional priority write-back
void launch_post_step(dqn_learner* h, int per_update) {
    // replay_sampler.update_sum_tree(tree_idxes, _make_priority(losses))  (deep_q_agent.py:341-347)
    if (per_update) launch_tree_update(h, h->B, h->b_tree_idx, h->b_newprio, 1);
}

void launch_copy_network(dqn_learner* h) {
    DQN_CUDA_OK(cudaMemcpyAsync(h->target, h->online, (size_t)h->net.slab * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    if (h->packed[0] && h->cfg.math_mode != DQN_MATH_FP32)
        DQN_CUDA_OK(cudaMemcpyAsync(h->packed[1], h->packed[0], (size_t)h->pk_floats * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
}
