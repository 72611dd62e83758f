// Q-network tower forward and backward (fp32 implicit-GEMM path).
//
// Replaces the TF graph built by rl/deep_q_model.py:266-354 (_make_q_network: 3x conv 'same'+ReLU, NHWC flatten,
// dense 512 ReLU, optional dueling streams) and its autodiff backward (rl/deep_q_model.py:391 optimizer.minimize).
// Layout: activations NHWC f32, conv kernels HWIO (== row-major [KH*KW*Cin, Cout]), dense kernels [in,out] --
// the TF layouts, so tensors copy 1:1 to/from checkpoints.
#include "tsgemm.cuh"
#include "cvgemm.cuh"
#include "cvtower.cuh"
#include "fcadam.cuh"
#include "fcfwd.cuh"
#include <algorithm>

static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

void net_describe(NetDesc& net, const dqn_config& cfg) {
    net.H = cfg.input_height; net.W = cfg.input_width; net.C = cfg.input_channels;
    net.A = cfg.num_actions; net.hidden = cfg.num_hidden; net.dueling = cfg.use_dueling ? 1 : 0;
    net.NH = net.hidden * (net.dueling ? 2 : 1);
    const int maps[3] = {32, 64, 64};
    const int ks_ref[3] = {8, 4, 4}, ks_nat[3] = {8, 4, 3};
    const int st[3] = {4, 2, 1};
    const bool same = cfg.conv_variant == DQN_CONV_REFERENCE;
    int ih = net.H, iw = net.W, cin = net.C;
    for (int l = 0; l < 3; ++l) {
        ConvGeom& g = net.conv[l];
        g.ih = ih; g.iw = iw; g.cin = cin; g.cout = maps[l]; g.s = st[l];
        g.k = same ? ks_ref[l] : ks_nat[l];
        if (same) {  // TF 'SAME': out = ceil(in/s); pad_total = max((out-1)*s + k - in, 0); before = pad_total/2
            g.oh = (ih + g.s - 1) / g.s; g.ow = (iw + g.s - 1) / g.s;
            int ph = (g.oh - 1) * g.s + g.k - ih; if (ph < 0) ph = 0;
            int pw = (g.ow - 1) * g.s + g.k - iw; if (pw < 0) pw = 0;
            g.pt = ph / 2; g.pl = pw / 2;
        } else {
            g.oh = (ih - g.k) / g.s + 1; g.ow = (iw - g.k) / g.s + 1; g.pt = g.pl = 0;
        }
        g.cin_log2 = ilog2(cin);
        g.m_k = fastdiv_magic(g.k); g.m_ow = fastdiv_magic(g.ow); g.m_hw = fastdiv_magic(g.oh * g.ow);
        g.kt = g.k / g.s; g.m_kt = fastdiv_magic(g.kt); g.m_s = fastdiv_magic(g.s);
        DQN_REQUIRE((1 << g.cin_log2) == cin && cin % 4 == 0, "channel counts must be powers of two >= 4");
        DQN_REQUIRE(g.k % g.s == 0, "kernel size must be a multiple of the stride");
        ih = g.oh; iw = g.ow; cin = g.cout;
    }
    net.flat = ih * iw * cin;
    DQN_REQUIRE(net.hidden % 8 == 0 && net.flat % 16 == 0, "hidden/flat alignment");
    int64_t off = 0;
    auto add = [&](const std::string& name, int64_t count) {
        TensorInfo t; t.name = name; t.offset = off; t.count = count;
        net.tensors.push_back(t);
        int64_t o = off;
        off += (count + 3) / 4 * 4;  // keep every tensor 16-byte aligned
        return o;
    };
    net.tensors.clear();
    for (int l = 0; l < 3; ++l) {
        const ConvGeom& g = net.conv[l];
        std::string nm = l == 0 ? "conv2d" : ("conv2d_" + std::to_string(l));
        net.off_cw[l] = add(nm + "/kernel", (int64_t)g.k * g.k * g.cin * g.cout);
        net.off_cb[l] = add(nm + "/bias", g.cout);
    }
    // creation order of rl/deep_q_model.py:312-343: dense (adv hidden), dense_1 (adv), dense_2 (val hidden), dense_3 (val)
    const int streams = net.dueling ? 2 : 1;
    for (int s = 0; s < streams; ++s) {
        const int outs = s == 0 ? net.A : 1;
        std::string n0 = (2 * s) == 0 ? "dense" : ("dense_" + std::to_string(2 * s));
        std::string n1 = "dense_" + std::to_string(2 * s + 1);
        net.off_fc_w[s] = add(n0 + "/kernel", (int64_t)net.flat * net.hidden);
        net.off_fc_b[s] = add(n0 + "/bias", net.hidden);
        net.off_hd_w[s] = add(n1 + "/kernel", (int64_t)net.hidden * outs);
        net.off_hd_b[s] = add(n1 + "/bias", outs);
    }
    if (!net.dueling) { net.off_fc_w[1] = net.off_fc_w[0]; net.off_fc_b[1] = net.off_fc_b[0]; net.off_hd_w[1] = net.off_hd_w[0]; net.off_hd_b[1] = net.off_hd_b[0]; }
    net.slab = off;
}

// ---- tile selection: largest tile that still yields >= one CTA per SM --------------------------------
template <bool SPLIT, class AOp, class BOp, class Epi>
static void igemm_auto(dqn_learner* h, const AOp& a, const BOp& b, const Epi& e, int M, int N, int K, int Z, int kchunk) {
    const int sms = h->sm_count;
    auto tiles = [&](int bm, int bn) { return (int64_t)ceil_div(M, bm) * ceil_div(N, bn) * Z; };
    if (N % 64 == 0) {
        if (M > 64 && tiles(128, 64) >= sms) return launch_igemm<TileCfg<128, 64, 16, 8, 4>, SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
        if (M > 32 && tiles(64, 64) >= sms) return launch_igemm<TileCfg<64, 64, 16, 4, 4>, SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
        return launch_igemm<TileCfg<32, 64, 16, 4, 4>, SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
    }
    if (M > 64 && tiles(128, 32) >= sms) return launch_igemm<TileCfg<128, 32, 16, 8, 4>, SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
    if (M > 32 && tiles(64, 32) >= sms) return launch_igemm<TileCfg<64, 32, 16, 4, 4>, SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
    return launch_igemm<TileCfg<32, 32, 16, 4, 4>, SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
}

// ---- math-mode dispatch: tcgen05 tensor cores (TF32 x1 / x3) or the FFMA kernel ------------------------------------
template <int BN, bool SPLIT, class AOp, class BOp, class Epi>
static void tc_go(dqn_learner* h, const AOp& a, const BOp& b, const Epi& e, int M, int N, int K, int Z, int kchunk, int nclass = 0) {
    if constexpr (ts::capable<AOp>::value) {
        if (h->ts_gemm || nclass > 0) {  // A through tensor memory
            if (h->cfg.math_mode == DQN_MATH_TC3X) ts::launch<BN, 3, SPLIT>(h, a, b, e, M, N, K, Z, kchunk, nclass);
            else ts::launch<BN, 1, SPLIT>(h, a, b, e, M, N, K, Z, kchunk, nclass);
            return;
        }
    }
    DQN_REQUIRE(nclass == 0, "class + split launches need the tensor-memory kernel");
    if (h->cfg.math_mode == DQN_MATH_TC3X) tc::launch<BN, 3, SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
    else tc::launch<BN, 1, SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
}

// bn_hint: force the N tile (0 = choose).  A wide tile reads the 128-row operand once -- what an HBM-streamed A wants --
// while the automatic choice trades re-reads of an L2-resident A for more CTAs.
template <bool SPLIT, class AOp, class BOp, class Epi>
static void gemm_auto(dqn_learner* h, const AOp& a, const BOp& b, const Epi& e, int M, int N, int K, int Z, int kchunk,
                      int bn_hint = 0, int nclass = 0) {
    if (h->cfg.math_mode == DQN_MATH_FP32 || (N % 32) != 0) return igemm_auto<SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
    if (bn_hint == 128 && N % 128 == 0) return tc_go<128, SPLIT>(h, a, b, e, M, N, K, Z, kchunk, nclass);
    if (bn_hint == 64 && N % 64 == 0) return tc_go<64, SPLIT>(h, a, b, e, M, N, K, Z, kchunk, nclass);
    if (bn_hint == 32) return tc_go<32, SPLIT>(h, a, b, e, M, N, K, Z, kchunk, nclass);
    const int64_t mt = (int64_t)ceil_div(M, tc::BM) * Z;
    // widest N tile that still gives about one CTA per SM; narrower tiles re-read A from L2 but fill the chip
    if (N % 128 == 0 && mt * (N / 128) >= h->sm_count) return tc_go<128, SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
    if (N % 64 == 0 && mt * (N / 64) >= h->sm_count) return tc_go<64, SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
    return tc_go<32, SPLIT>(h, a, b, e, M, N, K, Z, kchunk);
}

__global__ void u8_to_f32_kernel(const uint32_t* __restrict__ src, float4* __restrict__ dst, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t w = src[i];
        dst[i] = make_float4(u8_to_unit(w & 255u), u8_to_unit((w >> 8) & 255u), u8_to_unit((w >> 16) & 255u), u8_to_unit(w >> 24));
    }
}

// cast + divide by 255 (rl/deep_q_model.py:226-227) once per batch for the cp.async loaders of the tcgen05 path
void launch_u8_to_f32(dqn_learner* h, const uint8_t* src, float* dst, size_t n) {
    const size_t n4 = n / 4;
    int blocks = (int)std::min<size_t>((n4 + 255) / 256, (size_t)h->sm_count * 8);
    launch_kernel(false, u8_to_f32_kernel, dim3(blocks), dim3(256), 0, h->stream, reinterpret_cast<const uint32_t*>(src),
                  reinterpret_cast<float4*>(dst), n4);
    h->launches++;
}

// ---- dense hidden layer: reduce split-K partials, bias, ReLU, then the head dot products and dueling combine
// One CTA per batch row, one thread per hidden column: sum the split-K partials (independent loads), bias, ReLU,
// then A (+1) dot products reduced across the CTA.  AMAX = compile-time bound on the action count.
template <int AMAX>
__global__ void __launch_bounds__(1024) fc_reduce_heads_kernel(const float* __restrict__ part, int splits, int rows, int NH,
                                                               int hidden, int A, int dueling,
                                                               const float* __restrict__ b0, const float* __restrict__ b1,
                                                               const float* __restrict__ hw0, const float* __restrict__ hb0,
                                                               const float* __restrict__ hw1, const float* __restrict__ hb1,
                                                               float* __restrict__ hid, float* __restrict__ q) {
    const int row = blockIdx.x;
    float acc[AMAX + 1];
#pragma unroll
    for (int a = 0; a <= AMAX; ++a) acc[a] = 0.f;
    for (int col = threadIdx.x; col < NH; col += blockDim.x) {
        const float* p = part + (size_t)row * NH + col;
        const size_t zs = (size_t)rows * NH;
        float s = 0.f;
        int z = 0;
        for (; z + 4 <= splits; z += 4) {  // four partial loads in flight; summation order stays z ascending
            const float v0 = p[(size_t)z * zs], v1 = p[(size_t)(z + 1) * zs], v2 = p[(size_t)(z + 2) * zs], v3 = p[(size_t)(z + 3) * zs];
            s += v0; s += v1; s += v2; s += v3;
        }
        for (; z < splits; ++z) s += p[(size_t)z * zs];
        const int st = col >= hidden, j = st ? col - hidden : col;
        s += st ? b1[j] : b0[j];
        const float hv = fmaxf(s, 0.f);  // tf.nn.relu (deep_q_model.py:316,325,339)
        hid[(size_t)row * NH + col] = hv;
        if (!st) {
#pragma unroll
            for (int a = 0; a < AMAX; ++a)
                if (a < A) acc[a] = fmaf(hv, hw0[j * A + a], acc[a]);
        } else {
            acc[AMAX] = fmaf(hv, hw1[j], acc[AMAX]);
        }
    }
    __shared__ float red[32][AMAX + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int a = 0; a <= AMAX; ++a) {
        float v = acc[a];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[w][a] = v;
    }
    __syncthreads();
    if (w == 0) {
        // lane a finishes output a (a < A: advantage / Q head, a == A: value head)
        const int a = lane;
        float v = 0.f;
        if (a <= A) {
            const int src = a < A ? a : AMAX;
            for (int ww = 0; ww < nw; ++ww) v += red[ww][src];
            v += a < A ? hb0[a] : (dueling ? hb1[0] : 0.f);
        }
        if (dueling) {
            float sum = a < A ? v : 0.f;
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum / (float)A;  // tf.reduce_mean(advantage, axis=1) (deep_q_model.py:332-334)
            const float val = __shfl_sync(0xffffffffu, v, A);
            if (a < A) q[(size_t)row * A + a] = val + (v - mean);
        } else if (a < A) {
            q[(size_t)row * A + a] = v;
        }
    }
}

static void launch_fc_reduce_heads(dqn_learner* h, const float* P, int splits, int rows, TowerActs& acts) {
    const NetDesc& net = h->net;
    const int threads = std::min(1024, ((net.NH + 31) / 32) * 32);
#define FC_HEADS(AMAX)                                                                                                  \
    launch_kernel(false, fc_reduce_heads_kernel<AMAX>, dim3(rows), dim3(threads), 0, h->stream, acts.fc_part, splits, rows, net.NH, net.hidden, net.A, \
        net.dueling, P + net.off_fc_b[0], P + net.off_fc_b[1], P + net.off_hd_w[0], P + net.off_hd_b[0],               \
        P + net.off_hd_w[1], P + net.off_hd_b[1], acts.hid, acts.q)
    if (net.A <= 4) FC_HEADS(4);
    else if (net.A <= 8) FC_HEADS(8);
    else if (net.A <= 16) FC_HEADS(16);
    else FC_HEADS(31);
#undef FC_HEADS
    DQN_CUDA_OK(cudaGetLastError());
    h->launches++;
}

static int fc_kchunk() { return 256; }

// Pull the dense hidden kernels (2 towers x 31.5 MB at C2) into L2 while the conv towers run: HBM is idle then, a third of the
// SMs too, and the dense forward / dgrad that follow are streams of exactly these bytes.  One bulk prefetch per 16 KB.
__global__ void l2_prefetch_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, const float* __restrict__ d,
                                   unsigned chunks_each) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4 * chunks_each) return;
    const unsigned w = i / chunks_each, k = i - w * chunks_each;
    const float* base = w == 0 ? a : w == 1 ? b : w == 2 ? c : d;
    if (!base) return;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + (size_t)k * 4096), "r"(16384) : "memory");
}

// forked at the start of a step (own branch, lowest priority); nothing waits for it
void launch_dense_prefetch(dqn_learner* h, bool with_target) {
    const NetDesc& net = h->net;
    const size_t fcw = (size_t)net.flat * net.hidden;
    if (!h->fc_prefetch || fcw % 4096 != 0) return;
    const unsigned chunks = (unsigned)(fcw / 4096);
    const float* t0 = with_target ? h->target + net.off_fc_w[0] : nullptr;
    const float* t1 = with_target && net.dueling ? h->target + net.off_fc_w[1] : nullptr;
    cudaStream_t sv;
    side_begin(h, sv, 6, 0);
    launch_kernel(false, l2_prefetch_kernel, dim3(ceil_div(4 * (int)chunks, 128)), dim3(128), 0, h->stream, (const float*)(h->online + net.off_fc_w[0]),
                  net.dueling ? (const float*)(h->online + net.off_fc_w[1]) : (const float*)nullptr, t0, t1, chunks);
    h->launches++;
    side_end(h, sv);
}

// ---- image-resident conv path (cvgemm.cuh) -------------------------------------------------------------------------
void alloc_packed_conv(dqn_learner* h) {
    int64_t off = 0;
    for (int l = 0; l < 3; ++l) { h->pk_off[l] = off; off += (int64_t)cv::packed_floats(h->net.conv[l]); }
    h->pk_off[3] = off; off += (int64_t)cv::packed_floats(h->net.conv[0]);
    h->pk_floats = off;
    for (int t = 0; t < 2; ++t) DQN_CUDA_OK(cudaMalloc(&h->packed[t], (size_t)off * sizeof(float)));
    for (int t = 0; t < 2; ++t) DQN_CUDA_OK(cudaMalloc(&h->act3_pk[t], (size_t)(t == 0 ? 64 : 32) * h->net.flat * 2 * sizeof(float)));
    int64_t offd = 0;
    for (int l = 1; l < 3; ++l) { h->pkd_off[l] = offd; offd += (int64_t)cv::packed_floats(h->net.conv[l]); }
    DQN_CUDA_OK(cudaMalloc(&h->packed_dg, (size_t)offd * sizeof(float)));
}

bool dgrad_pack_ok(const dqn_learner* h, int layer) {
    cv::ImgGeom q; cv::Bands bd;
    return cv::dgrad_geom(h->net.conv[layer], q, bd);
}

void launch_pack_conv(dqn_learner* h, int tower) {
    if (h->cfg.math_mode == DQN_MATH_FP32 || !h->packed[tower]) return;
    const float* P = tower ? h->target : h->online;
    cv::PackJobs jobs; jobs.n = 0;
    int most = 0;
    for (int l = 0; l < 3; ++l) {
        const ConvGeom& g = h->net.conv[l];
        const int K = g.k * g.k * g.cin, N = g.cout;
        if (K % 32 != 0) continue;
        jobs.j[jobs.n++] = {P + h->net.off_cw[l], h->packed[tower] + h->pk_off[l], K, N, g.k, g.s, g.cin, g.cout, 0};
        if (l == 0) jobs.j[jobs.n++] = {P + h->net.off_cw[l], h->packed[tower] + h->pk_off[3], K, N, g.k, g.s, g.cin, g.cout, 2};
        most = std::max(most, K * N);
        cv::ImgGeom q; cv::Bands bd;
        if (tower == 0 && l > 0 && cv::dgrad_geom(g, q, bd))
            jobs.j[jobs.n++] = {P + h->net.off_cw[l], h->packed_dg + h->pkd_off[l], K, N, g.k, g.s, g.cin, g.cout, 1};
    }
    if (!jobs.n) return;
    launch_kernel(false, cv::pack_all_kernel, dim3(ceil_div(most, 256), jobs.n), dim3(256), 0, h->stream, jobs);
    h->launches++;
}

template <class Epi>
static bool conv_image_resident(dqn_learner* h, const float* in, const float* packed, const Epi& e, const ConvGeom& g, int images) {
    cv::ImgGeom q; cv::Bands bd;
    if (!h->img_conv || !cv::img_geom(g, q) || q.nbands > 4 || q.nbands == 3 || q.nkb < 4) return false;
    cv::forward_bands(q, bd);
    const bool x3 = h->cfg.math_mode == DQN_MATH_TC3X;
    if (g.cout == 64) { if (x3) cv::launch_conv<64, 3>(h, in, packed, e, q, bd, images); else cv::launch_conv<64, 1>(h, in, packed, e, q, bd, images); }
    else              { if (x3) cv::launch_conv<32, 3>(h, in, packed, e, q, bd, images); else cv::launch_conv<32, 1>(h, in, packed, e, q, bd, images); }
    return true;
}

static void tower_fc_heads(dqn_learner* h, const float* P, int rows, TowerActs& acts, bool tg);

// Both towers' conv stacks as ONE launch per layer (grid = online images + target images): three launches instead of
// six, and the target tower no longer trails the online one by a launch latency.  Returns false (nothing launched) if a
// layer is outside the image-resident kernel's coverage.
static bool g_pack_ok(const dqn_learner* h, int rows_on, int rows_tg) {
    return rows_on == 64 && rows_tg == 32 && h->net.conv[2].cout == 64 && ff::supported(h, rows_on) && ff::supported(h, rows_tg);
}

bool towers_conv_forward_merged(dqn_learner* h, const float* x_on, int rows_on, const float* x_tg, int rows_tg) {
    const NetDesc& net = h->net;
    if (h->cfg.math_mode == DQN_MATH_FP32 || !h->img_conv || !h->merge_towers) return false;
    cv::ImgGeom q[3]; cv::Bands bd[3];
    for (int l = 0; l < 3; ++l) {
        if (!cv::img_geom(net.conv[l], q[l]) || q[l].nbands > 4 || q[l].nbands == 3 || q[l].nkb < 4) return false;
        cv::forward_bands(q[l], bd[l]);
    }
    const bool x3 = h->cfg.math_mode == DQN_MATH_TC3X;
    const float *in_on = x_on, *in_tg = x_tg;
    static const char* kName[3] = {"conv1_fwd", "conv2_fwd", "conv3_fwd"};
    // conv3 also writes the dense layer's operand tiles when the TMA dense forward will consume them
    const bool pack = h->fc_tma_fwd && g_pack_ok(h, rows_on, rows_tg);
    h->act3_pk_valid[0] = h->act3_pk_valid[1] = false;
    for (int l = 0; l < 3; ++l) {
        const ConvGeom& g = net.conv[l];
        prof_mark(h, kName[l]);
        const bool pk = pack && l == 2;
        EpiBiasReluPack e_on{h->acts_on.act[l], h->online + net.off_cb[l], g.cout, pk ? h->act3_pk[0] : nullptr, rows_on, g.oh * g.ow};
        EpiBiasReluPack e_tg{h->acts_tg.act[l], h->target + net.off_cb[l], g.cout, pk ? h->act3_pk[1] : nullptr, rows_tg, g.oh * g.ow};
        if (pk) h->act3_pk_valid[0] = h->act3_pk_valid[1] = true;
        const float* pk_on = h->packed[0] + h->pk_off[l];
        const float* pk_tg = h->packed[1] + h->pk_off[l];
        if (g.cout == 64) {
            if (x3) cv::launch_conv<64, 3>(h, in_on, pk_on, e_on, q[l], bd[l], rows_on, in_tg, pk_tg, &e_tg, rows_tg);
            else cv::launch_conv<64, 1>(h, in_on, pk_on, e_on, q[l], bd[l], rows_on, in_tg, pk_tg, &e_tg, rows_tg);
        } else {
            if (x3) cv::launch_conv<32, 3>(h, in_on, pk_on, e_on, q[l], bd[l], rows_on, in_tg, pk_tg, &e_tg, rows_tg);
            else cv::launch_conv<32, 1>(h, in_on, pk_on, e_on, q[l], bd[l], rows_on, in_tg, pk_tg, &e_tg, rows_tg);
        }
        in_on = h->acts_on.act[l]; in_tg = h->acts_tg.act[l];
    }
    return true;
}

// Conv stacks of both towers (or of one: rows_tg == 0) as ONE launch of the fused tower kernel (cvtower.cuh), reading the
// uint8 batch itself.  Returns false (nothing launched) if the option is off or a layer is outside its coverage.
bool tower_fused_ready(const dqn_learner* h) {
    if (h->cfg.math_mode == DQN_MATH_FP32 || !h->img_conv || !h->tower_fused || !h->packed[0]) return false;
    cv::TowerArgs t; size_t bytes;
    return cv::tower_plan(h->net.conv, t, bytes);
}

bool towers_conv_forward_fused(dqn_learner* h, const float* P_on, const uint8_t* s_on, int rows_on, TowerActs& acts_on,
                               const uint8_t* s_tg, int rows_tg, int act_keep_rows) {
    const NetDesc& net = h->net;
    if (!tower_fused_ready(h)) return false;
    if (P_on != h->online && P_on != h->target) return false;  // packed weights exist for the handle's two towers only
    cv::TowerArgs t; size_t bytes;
    if (!cv::tower_plan(net.conv, t, bytes)) return false;
    const bool pack = rows_tg > 0 && h->fc_tma_fwd && g_pack_ok(h, rows_on, rows_tg);
    h->act3_pk_valid[0] = h->act3_pk_valid[1] = false;
    const int tw = P_on == h->online ? 0 : 1;
    t.a.in = s_on; t.b.in = s_tg; t.images_a = rows_on;
    for (int l = 0; l < 3; ++l) {
        const int64_t po = h->pk_off[l == 0 ? 3 : l];  // layer 0: the uint8 form (frame bytes x bf16 pieces of w / 255)
        t.a.packed[l] = h->packed[tw] + po; t.a.bias[l] = P_on + net.off_cb[l]; t.a.act[l] = acts_on.act[l];
        t.b.packed[l] = h->packed[1] + po; t.b.bias[l] = h->target + net.off_cb[l]; t.b.act[l] = h->acts_tg.act[l];
    }
    // activations to global memory: a training step reads those of the online tower's s rows only (weight gradients, ReLU
    // gates); act3 of every row is needed where the dense forward does not take the packed tiles
    const int keep = act_keep_rows < 0 ? rows_on : act_keep_rows;
    for (int l = 0; l < 3; ++l) {
        t.a.act_rows[l] = (l == 2 && !pack) ? rows_on : keep;
        t.b.act_rows[l] = (l == 2 && !pack) ? rows_tg : (act_keep_rows < 0 ? rows_tg : 0);
    }
    t.a.pk = pack ? h->act3_pk[0] : nullptr; t.a.pk_rows = rows_on;
    t.b.pk = pack ? h->act3_pk[1] : nullptr; t.b.pk_rows = rows_tg;
    if (pack) h->act3_pk_valid[0] = h->act3_pk_valid[1] = true;
    prof_mark(h, "conv_tower_fwd");
    if (h->cfg.math_mode == DQN_MATH_TC3X) cv::launch_tower<3>(h, t, bytes, rows_on + rows_tg);
    else cv::launch_tower<1>(h, t, bytes, rows_on + rows_tg);
    return true;
}

void tower_fc_forward(dqn_learner* h, const float* P, int rows, TowerActs& acts) {
    tower_fc_heads(h, P, rows, acts, h->prof_tag[0] == 't');
}

void tower_forward(dqn_learner* h, const float* P, const uint8_t* d_states, const float* x_f32, int rows, TowerActs& acts) {
    const NetDesc& net = h->net;
    const bool tcm = h->cfg.math_mode != DQN_MATH_FP32;
    const bool tg = h->prof_tag[0] == 't';
    // one tower (dqn_predict: the actor's get_action / get_actions): the fused conv tower on the uint8 rows, no f32 copy
    if (tcm && (P == h->online || P == h->target) && towers_conv_forward_fused(h, P, d_states, rows, acts, nullptr, 0)) {
        tower_fc_heads(h, P, rows, acts, tg);
        return;
    }
    DQN_REQUIRE(!tcm || x_f32 != nullptr, "tensor-core path needs the f32 input batch");
    const void* in = tcm ? (const void*)x_f32 : (const void*)d_states;
    static const char* kConvName[2][3] = {{"on_conv1_fwd", "on_conv2_fwd", "on_conv3_fwd"}, {"tg_conv1_fwd", "tg_conv2_fwd", "tg_conv3_fwd"}};
    for (int l = 0; l < 3; ++l) {
        const ConvGeom& g = net.conv[l];
        prof_mark(h, kConvName[tg][l]);
        OpConvFwdA a{in, g, (l == 0 && !tcm) ? 1 : 0};
        OpColContig b{P + net.off_cw[l], g.cout};
        EpiBiasRelu e{acts.act[l], P + net.off_cb[l], g.cout};
        const int tw = P == h->online ? 0 : 1;
        if (!(tcm && (P == h->online || P == h->target) &&
              conv_image_resident(h, (const float*)in, h->packed[tw] + h->pk_off[l], e, g, rows)))
            gemm_auto<false>(h, a, b, e, rows * g.oh * g.ow, g.cout, g.k * g.k * g.cin, 1, 0);
        in = acts.act[l];
    }
    tower_fc_heads(h, P, rows, acts, tg);
}

// dense hidden layer(s) (split-K partials) + bias / ReLU / heads / dueling combine of one tower
static void tower_fc_heads(dqn_learner* h, const float* P, int rows, TowerActs& acts, bool tg) {
    const NetDesc& net = h->net;
    const bool tcm = h->cfg.math_mode != DQN_MATH_FP32;
    int kchunk = fc_kchunk();
    int splits = ceil_div(net.flat, kchunk);
    if (h->defer_start && P == h->online) {
        // the deferred optimizer pass over the dense kernels (issued at the start of this graph) must have landed
        main_wait_side(h, 2);
        main_wait_side(h, 5);
    }
    prof_mark(h, tg ? "tg_fc_fwd" : "on_fc_fwd");
    if (tcm && h->fc_tma_fwd && (P == h->online || P == h->target) && ff::supported(h, rows)) {
        const int tw = P == h->online ? 0 : 1;
        splits = ff::launch(h, P, acts.act[2], rows, acts.fc_part, h->fc_splits, h->act3_pk_valid[tw] ? h->act3_pk[tw] : nullptr);
        h->act3_pk_valid[tw] = false;
    } else if (tcm && rows % 32 == 0) {
        // swap-AB: the 128-row MMA tile spans hidden units (weights stream as the A operand), the batch is N
        const int mt = ceil_div(net.NH, tc::BM);
        int want = std::min(std::max(1, h->sm_count / mt), h->fc_splits);
        kchunk = ceil_div(ceil_div(net.flat, want), tc::BK) * tc::BK;
        splits = ceil_div(net.flat, kchunk);
        DQN_REQUIRE(splits <= h->fc_splits, "fc split workspace too small");
        OpFcWeightFwd a{P + net.off_fc_w[0], P + net.off_fc_w[1], net.hidden};
        OpRowMajorK b{acts.act[2], net.flat};
        EpiPartialT e{acts.fc_part, rows, net.NH};
        // the weights are the HBM-streamed operand: one N tile over the whole batch, so every byte of W is read once
        gemm_auto<true>(h, a, b, e, net.NH, rows, net.flat, splits, kchunk, rows % 64 == 0 ? 64 : 32);
    } else {
        DQN_REQUIRE(splits <= h->fc_splits, "fc split workspace too small");
        OpRowMajorK a{acts.act[2], net.flat};
        OpFcWeightFwd b{P + net.off_fc_w[0], P + net.off_fc_w[1], net.hidden};
        EpiPartial e{acts.fc_part, rows, net.NH};
        igemm_auto<true>(h, a, b, e, rows, net.NH, net.flat, splits, kchunk);
    }
    prof_mark(h, tg ? "tg_fc_reduce_heads" : "on_fc_reduce_heads");
    launch_fc_reduce_heads(h, P, splits, rows, acts);
}

// debug: conv layer `layer` of the online tower alone on `rows` images (dqn_debug_timeline, mode 16 + layer);
// layer 3 = the dense forward + heads
void debug_conv_layer(dqn_learner* h, int layer, int rows) {
    const NetDesc& net = h->net;
    if (layer == 3) {
        // conv stacks of both towers first (they write the packed activation tiles when the TMA dense forward is on), un-stamped
        unsigned long long* dbg = h->tc_dbg; h->tc_dbg = nullptr;
        towers_conv_forward_merged(h, h->x_f32, rows, h->x_f32 + (size_t)(rows / 2) * h->state_bytes, rows / 2);
        h->tc_dbg = dbg;
        tower_fc_heads(h, h->online, rows, h->acts_on, false);
        return;
    }
    if (layer >= 5 && layer <= 7) {  // image-resident weight gradient of conv layer (layer - 5) on the buffers of the last training step
        const int l = layer - 5;
        const ConvGeom& g = net.conv[l];
        cv::ImgGeom q;
        DQN_REQUIRE(cv::img_geom(g, q), "debug wgrad geometry");
        const void* xin = l == 0 ? (const void*)h->b_states : (const void*)h->acts_on.act[l - 1];
        float* part = h->wg_part + h->wg_off[l];
        const bool ok = g.cout == 64 ? cv::launch_wgrad<64, 3>(h, xin, l == 0, h->d_act[l], part, q, rows) : cv::launch_wgrad<32, 3>(h, xin, l == 0, h->d_act[l], part, q, rows);
        DQN_REQUIRE(ok, "debug wgrad launch");
        return;
    }
    if (layer == 4) {
        // rows % 3 == 0: both towers as in a training step (2/3 online images, 1/3 target images)
        const int r_on = rows % 3 == 0 ? rows / 3 * 2 : rows, r_tg = rows - r_on;
        DQN_REQUIRE(towers_conv_forward_fused(h, h->online, h->b_states, r_on, h->acts_on, r_tg ? h->b_states + (size_t)(r_on / 2) * h->state_bytes : nullptr, r_tg),
                    "fused conv tower not available");
        return;
    }
    const ConvGeom& g = net.conv[layer];
    const float* P = h->online;
    const void* in = layer == 0 ? (const void*)h->x_f32 : (const void*)h->acts_on.act[layer - 1];
    OpConvFwdA a{in, g, 0};
    OpColContig b{P + net.off_cw[layer], g.cout};
    EpiBiasRelu e{h->acts_on.act[layer], P + net.off_cb[layer], g.cout};
    if (!conv_image_resident(h, (const float*)in, h->packed[0] + h->pk_off[layer], e, g, rows))
        gemm_auto<false>(h, a, b, e, rows * g.oh * g.ow, g.cout, g.k * g.k * g.cin, 1, 0);
}

// ---- backward ------------------------------------------------------------------------------------------
// conv wgrad with an extra all-ones row appended to A^T so the same GEMM also yields the bias gradient
struct OpConvWgradABias {
    static constexpr bool KCONTIG = false;
    OpConvWgradA base; int mreal; const float* ones;  // ones -> {1,0,0,0}
    typedef OpConvWgradA::Row Row;
    __device__ __forceinline__ Row row(int m, int z) const { return base.row(m < mreal ? m : 0, z); }
    __device__ __forceinline__ float4 load4(const Row& r, int m, int k, int z) const {
        if (m >= mreal) return make_float4(m == mreal ? 1.f : 0.f, 0.f, 0.f, 0.f);
        return base.load4(r, m, k, z);
    }
    __device__ __forceinline__ const float* addr4(const Row& r, int m, int k, int z) const {
        if (m >= mreal) return m == mreal ? ones : nullptr;
        return base.addr4(r, m, k, z);
    }
};

// grads[w_off + i] = sum_z part[z][i] for i < mreal*N ; grads[b_off + n] = sum_z part[z][mreal][n]
// CTA = 64 float4 outputs x 4 z-groups: each group adds its contiguous quarter of the partials in z order (8 loads in
// flight), the four group sums are then added in group order (deterministic).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int Z, int mreal, int N,
                                                           float* __restrict__ gw, float* __restrict__ gb) {
    __shared__ float4 s_sum[4][64];
    const int total4 = (mreal + 1) * N / 4;
    const size_t zstride = (size_t)(mreal + 4) * N;
    const int e = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int i = blockIdx.x * 64 + e;
    const int zq = (Z + 3) / 4, zb = grp * zq, ze = min(Z, zb + zq);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < total4) {
        const float* p = part + (size_t)i * 4;
        int z = zb;
        for (; z + 8 <= ze; z += 8) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(p + (size_t)(z + j) * zstride);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s.x += v[j].x; s.y += v[j].y; s.z += v[j].z; s.w += v[j].w; }
        }
        for (; z < ze; ++z) {
            const float4 v = *reinterpret_cast<const float4*>(p + (size_t)z * zstride);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    s_sum[grp][e] = s;
    __syncthreads();
    if (grp == 0 && i < total4) {
#pragma unroll
        for (int q = 1; q < 4; ++q) { const float4 v = s_sum[q][e]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
        const int el = i * 4;
        if (el < mreal * N) *reinterpret_cast<float4*>(gw + el) = s;
        else *reinterpret_cast<float4*>(gb + (el - mreal * N)) = s;
    }
}

// out = (sum_z part[z]) * (gate > 0): split-K consumer of the dense dgrad (ReLU backward of conv3 fused in)
__global__ void __launch_bounds__(256) sum_gate_kernel(const float* __restrict__ part, int Z, int total4,
                                                       const float* __restrict__ gate, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    float4 s = *reinterpret_cast<const float4*>(part + (size_t)i * 4);
    for (int z = 1; z < Z; ++z) {
        const float4 v = *reinterpret_cast<const float4*>(part + ((size_t)z * total4 + i) * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const float4 g = *reinterpret_cast<const float4*>(gate + (size_t)i * 4);
    *reinterpret_cast<float4*>(out + (size_t)i * 4) =
        make_float4(g.x > 0.f ? s.x : 0.f, g.y > 0.f ? s.y : 0.f, g.z > 0.f ? s.z : 0.f, g.w > 0.f ? s.w : 0.f);
}

void tower_backward(dqn_learner* h, const uint8_t* d_states, const float* x_f32, int rows) {
    const NetDesc& net = h->net;
    const bool tcm = h->cfg.math_mode != DQN_MATH_FP32;
    const float* P = h->online;
    float* G = h->grads;
    TowerActs& acts = h->acts_on;
    h->rows_joined = false;
    // Weight gradients do not feed anything before the optimizer, so each one is forked onto the side stream as soon
    // as its dY exists, while the main stream continues down the dgrad chain.
    cudaStream_t saved;
    auto fc_wgrad = [&]() {
        // dense hidden: dW = X^T dH (K = batch).  Fused mode applies Adam/momentum in the epilogue (dW never lands
        // in HBM); otherwise dW goes to the gradient slab (data-parallel all-reduce, gradient read-back).
        side_begin(h, saved);
        prof_mark(h, "fc_wgrad");
        OpColContig a{acts.act[2], net.flat};
        OpColContig b{h->d_hid, net.NH};
        if (h->fuse_fc_adam) {
            float* W = h->online;
            EpiFcWgradAdam e{W + net.off_fc_w[0], W + net.off_fc_w[1], h->slot_m + net.off_fc_w[0], h->slot_m + net.off_fc_w[1],
                             h->slot_v + net.off_fc_w[0], h->slot_v + net.off_fc_w[1], net.hidden, h->d_sc,
                             (float)h->cfg.learning_rate, (float)h->cfg.momentum, h->cfg.use_momentum};
            gemm_auto<false>(h, a, b, e, net.flat, net.NH, rows, 1, 0);
        } else {
            EpiFcWgrad e{G + net.off_fc_w[0], G + net.off_fc_w[1], net.hidden};
            gemm_auto<false>(h, a, b, e, net.flat, net.NH, rows, 1, 0);
        }
        // split (data-parallel) step: the dense-kernel gradients are final here, long before the conv chain is done
        if (h->phase_ev_armed) DQN_CUDA_OK(cudaEventRecord(h->phase_ev[0], h->stream));
        side_end(h, saved);
    };
    const bool ffma_adam = h->early_fc_adam && !h->fuse_fc_adam && h->fc_ffma_adam && rows <= 32 && fc_wgrad_adam_supported(h);
    // default for a whole step on one GPU: tcgen05 wgrad + optimizer in one pass with TMA-staged state (fcadam.cuh)
    const bool rows_adam = h->early_fc_adam && !h->fuse_fc_adam && !ffma_adam && !h->fc_grads_needed && tcm && h->fc_rows_ctas > 0 &&
                           fa::rows_supported(h, rows);
    const bool tma_adam = rows_adam || (h->early_fc_adam && !h->fuse_fc_adam && !ffma_adam && !h->fc_grads_needed && tcm &&
                                        h->fc_tma_adam && fa::supported(h, rows));
    // unfused + optimizer inside backward: one wgrad launch per hidden stream, each on its own side stream, so that the
    // optimizer pass of stream 0 (HBM-bound) overlaps the weight gradient of stream 1 (tensor-bound)
    const bool split_fc = h->early_fc_adam && !h->fuse_fc_adam && !ffma_adam && !tma_adam && tcm && net.dueling && h->fc_split_streams;
    auto fc_wgrad_stream = [&](int st) {
        side_begin(h, saved, st == 0 ? 0 : 5, 0);
        prof_mark(h, st == 0 ? "fc_wgrad" : "fc_wgrad1");
        OpColContig a{acts.act[2], net.flat};
        OpColContig b{h->d_hid + (size_t)st * net.hidden, net.NH};
        EpiFcWgrad e{G + net.off_fc_w[st], G + net.off_fc_w[st], net.hidden};
        gemm_auto<false>(h, a, b, e, net.flat, net.hidden, rows, 1, 0);
        side_end(h, saved);
    };
    if (split_fc) { fc_wgrad_stream(0); fc_wgrad_stream(1); }
    else if (!h->fuse_fc_adam && !ffma_adam && !tma_adam) fc_wgrad();  // unfused: independent of the dgrad below, start it first
    // fused dense wgrad+optimizer writing to the scratch copy: it only needs dH and act3, so its side stream forks HERE, in front
    // of the dense dgrad; the kernel itself is issued right behind the dgrad launch (both become ready at the same instant and
    // the high-priority dgrad grid is dispatched first, the persistent low-priority grid takes the SMs that are left)
    const bool rows_early = rows_adam && h->w_scratch_on && h->w_scratch != nullptr && overlap_enabled(h);
    cudaStream_t early_saved = nullptr;
    if (rows_early) { side_begin(h, early_saved, 0); side_end(h, early_saved); }
    // dense hidden dgrad, gated by conv3's ReLU: d_act[2] = (dH W^T) * (act3 > 0)
    cudaEvent_t dgrad_read_w = nullptr;
    {
        prof_mark(h, "fc_dgrad");
        if (tcm && rows % 32 == 0) {
            // swap-AB (weights = the 128-row operand, K-major as stored, streamed through the cp.async ring) and
            // split-K over the hidden units so that 2 x flat/128 = 120 CTAs pull W from HBM; the partial sums are
            // added and gated by a small consumer kernel
            const int Z = 2, kchunk = net.NH / Z;
            DQN_REQUIRE((int64_t)Z * net.flat <= (int64_t)h->fc_splits * net.NH && kchunk % 32 == 0, "dense dgrad split workspace");
            OpFcWeightBwd a{P + net.off_fc_w[0], P + net.off_fc_w[1], net.hidden};
            OpRowMajorK b{h->d_hid, net.NH};
            EpiPartialT e{acts.fc_part, rows, net.flat};  // the forward split-K workspace is idle during backward
            gemm_auto<true>(h, a, b, e, net.flat, rows, net.NH, Z, kchunk, 32);
            // the GEMM above is the last reader of the old dense weights: the in-place fused dense wgrad+optimizer kernel may start
            // behind it, next to the sum+gate kernel
            if (h->fc_after_dgrad && rows_adam && !rows_early && !h->profile) dgrad_read_w = fork_event_here(h);
            prof_mark(h, "fc_dgrad_reduce");
            const int total4 = rows * net.flat / 4;
            launch_kernel(false, sum_gate_kernel, dim3(ceil_div(total4, 256)), dim3(256), 0, h->stream, acts.fc_part, Z, total4,
                          acts.act[2], h->d_act[2]);
            h->launches++;
        } else {
            OpRowMajorK a{h->d_hid, net.NH};
            OpFcWeightBwd b{P + net.off_fc_w[0], P + net.off_fc_w[1], net.hidden};
            EpiGateStore e{h->d_act[2], acts.act[2], net.flat};
            igemm_auto<false>(h, a, b, e, rows, net.flat, net.NH, 1, 0);
        }
    }
    if (h->fuse_fc_adam) fc_wgrad();  // fused: it overwrites W, so it starts only after the dgrad has read W
    if (split_fc && h->defer_capture) {
        h->fc_adam_done = true;  // applied at the start of the next step's graph (or by the flush at the end of dqn_train_steps)
    } else if (split_fc) {
        for (int st = 0; st < 2; ++st) {
            side_begin(h, saved, st == 0 ? 0 : 5, 0);  // after this stream's wgrad, and after the dgrad above has read the old W
            prof_mark(h, st == 0 ? "optimizer_fc" : "optimizer_fc1");
            launch_optimizer_fc_stream(h, st);
            side_end(h, saved);
        }
        h->fc_adam_done = true;
    } else if (rows_early) {
        // side stream 0 already waits for the head backward only (fork above): launch there, then, once the dense dgrad on the
        // main stream has read the old weights, copy the updated dense kernels from the scratch over the slab
        cudaStream_t main_s = h->stream;
        const int sp = g_launch_priority;
        h->stream = h->side_stream; g_launch_priority = h->prio_lo;
        prof_mark(h, "fc_wgrad_adam");
        fa::launch_rows(h, rows, std::min(h->fc_rows_ctas, h->sm_count));
        h->stream = main_s; g_launch_priority = sp;
        cudaEvent_t e = h->fork_ev[21];  // dgrad (+ its reduce) done
        DQN_CUDA_OK(cudaEventRecord(e, main_s));
        DQN_CUDA_OK(cudaStreamWaitEvent(h->side_stream, e, 0));
        const size_t fcw = (size_t)net.flat * net.hidden * sizeof(float);
        for (int st = 0; st < 2; ++st)
            DQN_CUDA_OK(cudaMemcpyAsync(h->online + net.off_fc_w[st], h->w_scratch + (size_t)st * net.flat * net.hidden, fcw,
                                        cudaMemcpyDeviceToDevice, h->side_stream));
        h->fc_adam_done = true;
    } else if (h->early_fc_adam && !h->fuse_fc_adam) {
        // the dense kernels hold 98.6% of the parameters: update them now on the first side stream (after their wgrad,
        // and after the dgrad above has read the old W) while the main stream walks the conv dgrad chain
        side_begin_at(h, saved, 0, 0, dgrad_read_w);
        if (tma_adam) {
            prof_mark(h, "fc_wgrad_adam");
            if (rows_adam) {
                // (tail_join) the optimizer tail will not wait for this kernel: they share the end-of-step join
                const bool join = h->tail_join && !(h->w_scratch_on && h->w_scratch);
                fa::launch_rows(h, rows, std::min(h->fc_rows_ctas, h->sm_count), join);
                h->rows_joined = join;
                if (h->w_scratch_on && h->w_scratch) {  // (not forked early: single-stream / profile mode) the copy follows at once
                    const size_t fcw = (size_t)net.flat * net.hidden * sizeof(float);
                    for (int st = 0; st < 2; ++st)
                        DQN_CUDA_OK(cudaMemcpyAsync(h->online + net.off_fc_w[st], h->w_scratch + (size_t)st * net.flat * net.hidden, fcw,
                                                    cudaMemcpyDeviceToDevice, h->stream));
                }
            } else fa::launch(h, rows);
        } else if (ffma_adam) {
            prof_mark(h, "fc_wgrad_adam");
            launch_fc_wgrad_adam(h, rows, h->fc_grads_needed);
        } else {
            prof_mark(h, "optimizer_fc");
            launch_optimizer_fc(h);
        }
        side_end(h, saved);
        h->fc_adam_done = true;
    }
    for (int l = 2; l >= 0; --l) {
        const ConvGeom& g = net.conv[l];
        float* part = h->wg_part + h->wg_off[l];
        const int mreal = g.k * g.k * g.cin, N = g.cout, K = rows * g.oh * g.ow;
        // wgrad (+bias) split over the pixel dimension
        {
            // needs d_act[l], which the main stream has just produced; one side stream per layer, so that a layer's
            // weight gradient never queues behind the previous layer's
            const int wstream = l == 2 ? 1 : l == 1 ? 3 : 4;
            side_begin(h, saved, wstream, 1);
            const int mt = ceil_div(mreal + 4, tcm ? tc::BM : 64);
            // conv1's wgrad runs last, next to the persistent dense wgrad+optimizer kernel: size it for the SMs that one leaves
            int sms_avail = h->sm_count;
            if (l == 0 && rows_adam && h->wg1_fit) sms_avail = std::max(32, h->sm_count - h->fc_rows_ctas);
            int Z = ((tcm ? 1 : 2) * sms_avail) / (mt * ceil_div(N, 64));
            if (Z < 1) Z = 1; if (Z > 64) Z = 64;
            int kchunk = ceil_div(ceil_div(K, Z), 32) * 32;
            Z = ceil_div(K, kchunk);
            DQN_REQUIRE((int64_t)Z * (mreal + 4) * N <= h->wg_part_floats, "wgrad workspace too small");
            static const char* kW[3] = {"conv1_wgrad", "conv2_wgrad", "conv3_wgrad"};
            prof_mark(h, kW[l]);
            const void* in0 = tcm ? (const void*)x_f32 : (const void*)d_states;
            bool done = false;
            cv::ImgGeom q;
            if (tcm && h->img_conv && (l > 0 || h->img_wgrad1) && rows <= 64 && cv::img_geom(g, q)) {
                // image-resident kernel: one partial per image (Z = rows), summed in image order by the reduce below;
                // conv1 reads the uint8 frame stack itself
                const void* xin = l == 0 ? (const void*)d_states : (const void*)acts.act[l - 1];
                const bool u8 = l == 0;
                const bool x3 = h->cfg.math_mode == DQN_MATH_TC3X;
                if (N == 64) done = x3 ? cv::launch_wgrad<64, 3>(h, xin, u8, h->d_act[l], part, q, rows) : cv::launch_wgrad<64, 1>(h, xin, u8, h->d_act[l], part, q, rows);
                else if (N == 32) done = x3 ? cv::launch_wgrad<32, 3>(h, xin, u8, h->d_act[l], part, q, rows) : cv::launch_wgrad<32, 1>(h, xin, u8, h->d_act[l], part, q, rows);
                if (done) Z = rows;
            }
            if (!done) {
                if (l == 0 && tcm && !h->xf_valid) {  // the fused tower did not need the f32 copy of the batch; this kernel does
                    launch_u8_to_f32(h, d_states, h->x_f32, (size_t)rows * h->state_bytes);
                    h->xf_valid = true;
                }
                OpConvWgradABias a{OpConvWgradA{l == 0 ? in0 : (const void*)acts.act[l - 1], g, (l == 0 && !tcm) ? 1 : 0}, mreal, h->ones_row};
                OpColContig b{h->d_act[l], N};
                EpiPartial e{part, mreal + 4, N};
                gemm_auto<true>(h, a, b, e, mreal + 4, N, K, Z, kchunk);
            }
            static const char* kR[3] = {"conv1_wgrad_reduce", "conv2_wgrad_reduce", "conv3_wgrad_reduce"};
            prof_mark(h, kR[l]);
            launch_kernel(false, wgrad_reduce_kernel, dim3(ceil_div((mreal + 1) * N / 4, 64)), dim3(256), 0, h->stream, part, Z, mreal,
                          N, G + net.off_cw[l], G + net.off_cb[l]);
            h->launches++;
            side_end(h, saved);
        }
        if (l == 0) break;  // no gradient flows into the input image
        // dgrad into the previous activation, gated by its ReLU
        {
            prof_mark(h, l == 2 ? "conv3_dgrad" : "conv2_dgrad");
            const int cl2 = ilog2(g.cout);
            DgradClass c{g};
            const int Zc = g.s * g.s;
            const int cy = ceil_div(g.ih, g.s), cx = ceil_div(g.iw, g.s);
            const int Mc = rows * cy * cx;  // upper bound over classes
            OpConvDgradA a{h->d_act[l], c, cl2, rows};
            OpConvDgradB b{P + net.off_cw[l], g, cl2};
            const int Kd = (g.k / g.s) * (g.k / g.s) * g.cout;
            const int ctas = ceil_div(Mc, tc::BM) * ceil_div(g.cin, 32) * Zc;
            cv::ImgGeom q; cv::Bands bd;
            if (tcm && h->img_conv && h->img_dgrad && cv::dgrad_geom(g, q, bd) && q.nkb >= 4) {
                // image-resident: one CTA per image holds dY, the stride-parity classes are its accumulators
                EpiGatePix e{h->d_act[l - 1], acts.act[l - 1], g.cin};
                const float* pk = h->packed_dg + h->pkd_off[l];
                const bool x3 = h->cfg.math_mode == DQN_MATH_TC3X;
                // the stride classes of one image are independent accumulators: two per CTA while that still fits one wave
                const int bpc = (bd.n >= 4 && 2 * rows <= h->sm_count) ? 2 : 4;
                if (g.cin == 64) { if (x3) cv::launch_conv<64, 3>(h, h->d_act[l], pk, e, q, bd, rows, nullptr, nullptr, (const EpiGatePix*)nullptr, 0, bpc); else cv::launch_conv<64, 1>(h, h->d_act[l], pk, e, q, bd, rows, nullptr, nullptr, (const EpiGatePix*)nullptr, 0, bpc); }
                else             { if (x3) cv::launch_conv<32, 3>(h, h->d_act[l], pk, e, q, bd, rows, nullptr, nullptr, (const EpiGatePix*)nullptr, 0, bpc); else cv::launch_conv<32, 1>(h, h->d_act[l], pk, e, q, bd, rows, nullptr, nullptr, (const EpiGatePix*)nullptr, 0, bpc); }
            } else if (tcm && h->ts_gemm && 2 * ctas <= h->sm_count && Kd % 64 == 0 && g.cin % 32 == 0 &&
                (int64_t)2 * rows * g.ih * g.iw * g.cin <= (int64_t)h->fc_splits * rows * net.NH) {
                // too few tiles to fill the chip (conv3: 60): split K in two as well; the halves are added and gated by
                // the consumer kernel.  Workspace: the forward split-K buffer (idle here).
                EpiDgradPartial e{acts.fc_part, c, rows, Zc, (size_t)rows * g.ih * g.iw * g.cin};
                gemm_auto<true>(h, a, b, e, Mc, g.cin, Kd, 2 * Zc, Kd / 2, 32, Zc);
                prof_mark(h, l == 2 ? "conv3_dgrad_reduce" : "conv2_dgrad_reduce");
                const int total4 = rows * g.ih * g.iw * g.cin / 4;
                launch_kernel(false, sum_gate_kernel, dim3(ceil_div(total4, 256)), dim3(256), 0, h->stream, acts.fc_part, 2, total4,
                              acts.act[l - 1], h->d_act[l - 1]);
                h->launches++;
            } else {
                EpiDgradGate e{h->d_act[l - 1], acts.act[l - 1], c, rows};
                gemm_auto<false>(h, a, b, e, Mc, g.cin, Kd, Zc, 0);
            }
        }
    }
    if (!h->rows_joined) main_wait_side(h, 0);  // all weight gradients are in the slab (and the dense kernels updated) before the
    main_wait_side(h, 1);                       // optimizer pass over the remaining tensors runs (rows_joined: finish_step joins stream 0)
    main_wait_side(h, 3);
    main_wait_side(h, 4);
    if (split_fc) main_wait_side(h, 5);
}

// ---- GEMM self-test: plain matrices through either engine (dqn_selftest_gemm) ----------------------------------------
// A is [M][K] (a_kcontig) or [K][M]; B is [N][K] (b_kcontig) or [K][N]; C is [M][N].
int selftest_gemm(dqn_learner* h, int mode, int M, int N, int K, int a_kcontig, int b_kcontig, const float* dA,
                  const float* dB, float* dC) {
    const int saved = h->cfg.math_mode;
    h->cfg.math_mode = mode;
    EpiStore e{dC, N};
    try {
        if (a_kcontig && b_kcontig) gemm_auto<false>(h, OpRowMajorK{dA, K}, OpRowMajorK{dB, K}, e, M, N, K, 1, 0);
        else if (a_kcontig && !b_kcontig) gemm_auto<false>(h, OpRowMajorK{dA, K}, OpColContig{dB, N}, e, M, N, K, 1, 0);
        else if (!a_kcontig && b_kcontig) gemm_auto<false>(h, OpColContig{dA, M}, OpRowMajorK{dB, K}, e, M, N, K, 1, 0);
        else gemm_auto<false>(h, OpColContig{dA, M}, OpColContig{dB, N}, e, M, N, K, 1, 0);
    } catch (...) {
        h->cfg.math_mode = saved;
        throw;
    }
    h->cfg.math_mode = saved;
    return 0;
}
