// Fused TD epilogue and head backward.
//
// Replaces (reference file:line):
//   rl/deep_q_model.py:247-255  max_q_values: reduce_max(target_q * one_hot(argmax(online_q))) (double) / reduce_max(target_q)
//   rl/deep_q_model.py:106-114  host target y = rewards(u8) + continues(bool) * discount * max_q, NumPy float64, fed as f32
//   rl/deep_q_model.py:361-378  q_a = sum(Q * one_hot(a)); abs / huber(delta=1) / is_w * squared_difference; reduce_mean
//   rl/deep_q_agent.py:527-534  _make_priority: min(|delta| + 0.01, 1.0) ** per_a  (float32)
// and the autodiff of the loss w.r.t. the head inputs (dueling combine, the two tiny head matmuls, hidden ReLU).
#include "learner.h"
#include <string.h>

struct TdArgs {
    int n, A, use_double, use_per, train;
    double gamma, grad_scale;
    float per_a;
    const float *q_on, *q_on_next, *q_tg_next;
    const uint8_t *actions, *rewards, *cont;
    const float* isw;
    float *o_y, *o_maxq, *o_losses, *o_dq, *o_newprio;
    dqn_learner::Scalars* sc;
};

// TD target, loss and dL/dq of sample i (the arithmetic of deep_q_model.py:247-255,106-114,361-378)
__device__ __forceinline__ void td_sample(const TdArgs& t, int i, float& y, float& maxq, float& out_l, float& dq, float& newprio, float& li) {
    const int A = t.A;
    const float* qt = t.q_tg_next + (size_t)i * A;
    if (t.use_double) {
        const float* qo = t.q_on_next + (size_t)i * A;
        int best = 0;
        float bv = qo[0];
        for (int a = 1; a < A; ++a) if (qo[a] > bv) { bv = qo[a]; best = a; }  // tf.argmax: first maximum
        // reduce_max over {target_q[a*] * 1, target_q[a] * 0 ...}: the masked zeros take part in the max
        maxq = qt[best];
        if (A > 1) maxq = fmaxf(maxq, 0.f);
    } else {
        maxq = qt[0];
        for (int a = 1; a < A; ++a) maxq = fmaxf(maxq, qt[a]);
    }
    const double y64 = (double)t.rewards[i] + ((t.cont[i] ? 1.0 : 0.0) * t.gamma) * (double)maxq;
    y = (float)y64;  // fed through a tf.float32 placeholder (deep_q_model.py:206)
    const float qa = t.q_on[(size_t)i * A + t.actions[i]];
    const float delta = y - qa;
    const float ad = fabsf(delta);
    if (t.use_per) {
        const float w = t.isw[i];
        li = w * (delta * delta);  // is_weights * squared_difference (deep_q_model.py:374)
        out_l = ad;                // self.losses = abs_losses (:375)
        dq = (float)(-2.0 * (double)w * (double)delta * t.grad_scale / (double)t.n);
    } else {
        const float quad = fminf(ad, 1.0f);  // tf.losses.huber_loss, delta=1
        li = 0.5f * quad * quad + (ad - quad);
        out_l = li;
        const float e = -delta;  // predictions - labels
        dq = (float)((double)fminf(fmaxf(e, -1.0f), 1.0f) * t.grad_scale / (double)t.n);
    }
    newprio = t.use_per ? powf(fminf(ad + 0.01f, 1.0f), t.per_a) : 0.f;
}

// outputs + scalar loss, by ONE CTA (fixed reduction order): thread i handles samples i, i + blockDim, ...
__device__ __forceinline__ void td_block(const TdArgs& t, float* s_dq_out) {
    __shared__ double s_red[32];
    double lsum = 0.0;
    for (int i = threadIdx.x; i < t.n; i += blockDim.x) {
        float y, maxq, out_l, dq, newprio, li;
        td_sample(t, i, y, maxq, out_l, dq, newprio, li);
        t.o_y[i] = y; t.o_maxq[i] = maxq; t.o_losses[i] = out_l;
        if (t.train) {
            t.o_dq[i] = dq;
            if (t.use_per) t.o_newprio[i] = newprio;
        }
        if (s_dq_out) s_dq_out[i] = dq;
        lsum += (double)li;
    }
    for (int o = 16; o; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
        const float loss = (float)(tot / (double)t.n);
        t.sc->loss = loss;
        if (t.train) t.sc->loss_sum += loss;
    }
}

// one thread per sample; a single CTA so the scalar loss is reduced in a fixed order
__global__ void __launch_bounds__(1024) td_epilogue_kernel(const TdArgs t) { td_block(t, nullptr); }

static TdArgs make_td_args(dqn_learner* h, int n, const float* q_on, const float* q_on_next, const float* q_tg_next,
                           const uint8_t* actions, const uint8_t* rewards, const uint8_t* cont, const float* isw, int train,
                           double grad_scale) {
    TdArgs t;
    t.n = n; t.A = h->net.A; t.use_double = h->cfg.use_double; t.use_per = h->cfg.use_per; t.train = train;
    t.gamma = h->cfg.discount_rate; t.grad_scale = grad_scale; t.per_a = (float)h->cfg.per_a;
    t.q_on = q_on; t.q_on_next = q_on_next; t.q_tg_next = q_tg_next;
    t.actions = actions; t.rewards = rewards; t.cont = cont; t.isw = isw;
    t.o_y = h->b_y; t.o_maxq = h->b_maxq; t.o_losses = h->b_losses_out; t.o_dq = h->b_dq; t.o_newprio = h->b_newprio;
    t.sc = h->d_sc;
    return t;
}

void launch_td_epilogue(dqn_learner* h, int n, const float* q_on, const float* q_on_next, const float* q_tg_next,
                        const uint8_t* actions, const uint8_t* rewards, const uint8_t* cont, const float* isw,
                        int train, double grad_scale) {
    int threads = ((n + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    const TdArgs t = make_td_args(h, n, q_on, q_on_next, q_tg_next, actions, rewards, cont, isw, train, grad_scale);
    launch_kernel(false, td_epilogue_kernel, dim3(1), dim3(threads), 0, h->stream, t);
    h->launches++;
}

// dH (gated by the hidden ReLU), head weight/bias grads and hidden bias grads.  CTA = 32 hidden columns x 32 samples:
// every (sample, column) is one thread (coalesced along the column), the sums over the batch are then formed from
// shared memory by one thread per (column, output) in ascending sample order with the same fmaf chain a serial walk
// would use (deterministic).
template <int AMAX>
__global__ void __launch_bounds__(1024) head_backward_kernel(int n, int NH, int hidden, int A, int dueling,
                                                             const float* __restrict__ hid, const float* __restrict__ dq,
                                                             const uint8_t* __restrict__ actions,
                                                             const float* __restrict__ hw0, const float* __restrict__ hw1,
                                                             float* __restrict__ d_hid, float* __restrict__ g_hw0,
                                                             float* __restrict__ g_hb0, float* __restrict__ g_hw1,
                                                             float* __restrict__ g_hb1, float* __restrict__ g_fb0,
                                                             float* __restrict__ g_fb1, const TdArgs td) {
    extern __shared__ float sm[];
    float* s_dh = sm;                 // [32][33]
    float* s_hv = sm + 32 * 33;       // [32][33]
    float* s_dq = sm + 2 * 32 * 33;   // [n]
    uint8_t* s_act = reinterpret_cast<uint8_t*>(s_dq + n);  // [n]
    if (td.n > 0) {
        // fused TD epilogue: every CTA derives dL/dq of all samples itself (a few hundred flops) instead of waiting for a
        // separate one-CTA kernel; CTA 0 also publishes the targets / losses / priorities and the scalar loss
        if (blockIdx.x == 0) {
            td_block(td, s_dq);
        } else {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                float y, maxq, out_l, dqi, newprio, li;
                td_sample(td, i, y, maxq, out_l, dqi, newprio, li);
                s_dq[i] = dqi;
            }
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) s_act[i] = actions[i];
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) { s_dq[i] = dq[i]; s_act[i] = actions[i]; }
    }
    __syncthreads();
    const float invA = 1.0f / (float)A;
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + c;           // NH is a multiple of 32
    const int st = col >= hidden, j = st ? col - hidden : col;
    float w[AMAX];
#pragma unroll
    for (int a = 0; a < AMAX; ++a) w[a] = (!st && a < A) ? hw0[j * A + a] : 0.f;
    const float w1 = st ? hw1[j] : 0.f;
    float acc = 0.f;  // r == 0: hidden bias grad; r == a+1: head weight grad for output a (value stream: r == 1)
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + r;
        float hv = 0.f, dh = 0.f;
        if (i < n) {
            hv = hid[(size_t)i * NH + col];
            const float d = s_dq[i];
            float g;
            if (!st) {
                const int ai = s_act[i];
                g = 0.f;
#pragma unroll
                for (int a = 0; a < AMAX; ++a)
                    if (a < A) {
                        // dQ = one_hot(a_i) * dq ; dueling: dA = dQ - mean_a(dQ)
                        const float da = dueling ? d * ((a == ai ? 1.f : 0.f) - invA) : (a == ai ? d : 0.f);
                        g = fmaf(da, w[a], g);
                    }
            } else {
                g = d * w1;  // dV = sum_a dQ = dq
            }
            dh = hv > 0.f ? g : 0.f;
            d_hid[(size_t)i * NH + col] = dh;
        }
        s_dh[r * 33 + c] = dh;
        s_hv[r * 33 + c] = hv;
        __syncthreads();
        const int cnt = min(32, n - i0);
        if (r == 0) {
            for (int q = 0; q < cnt; ++q) acc += s_dh[q * 33 + c];
        } else if (!st && r <= A) {
            const int a = r - 1;
            for (int q = 0; q < cnt; ++q) {
                const float d = s_dq[i0 + q];
                const int ai = s_act[i0 + q];
                const float da = dueling ? d * ((a == ai ? 1.f : 0.f) - invA) : (a == ai ? d : 0.f);
                acc = fmaf(s_hv[q * 33 + c], da, acc);
            }
        } else if (st && r == 1) {
            for (int q = 0; q < cnt; ++q) acc = fmaf(s_hv[q * 33 + c], s_dq[i0 + q], acc);
        }
        __syncthreads();
    }
    if (r == 0) { if (st) g_fb1[j] = acc; else g_fb0[j] = acc; }
    else if (!st && r <= A) g_hw0[j * A + (r - 1)] = acc;
    else if (st && r == 1) g_hw1[j] = acc;
    if (blockIdx.x == 0 && threadIdx.x <= A) {  // head biases
        const int a = threadIdx.x;
        float s = 0.f;
        if (a < A) {
            for (int i = 0; i < n; ++i) {
                const float d = s_dq[i];
                s += dueling ? d * ((a == (int)s_act[i] ? 1.f : 0.f) - invA) : (a == (int)s_act[i] ? d : 0.f);
            }
            g_hb0[a] = s;
        } else if (dueling) {
            for (int i = 0; i < n; ++i) s += s_dq[i];
            g_hb1[0] = s;
        }
    }
}

void launch_head_backward(dqn_learner* h, int n, const uint8_t* actions, const TdFuse* fuse) {
    TdArgs td; memset(&td, 0, sizeof(td));
    if (fuse) td = make_td_args(h, n, fuse->q_on, fuse->q_on_next, fuse->q_tg_next, actions, fuse->rewards, fuse->cont, fuse->isw, 1, fuse->grad_scale);
    const NetDesc& net = h->net;
    const float* P = h->online;
    float* G = h->grads;
    const size_t smem = (size_t)2 * 32 * 33 * 4 + (size_t)n * 4 + ((n + 3) / 4) * 4;
    DQN_REQUIRE(net.NH % 32 == 0 && net.A <= 31, "head backward tiling");
#define HEAD_BWD(AMAX)                                                                                                  \
    launch_kernel(false, head_backward_kernel<AMAX>, dim3(net.NH / 32), dim3(1024), smem, h->stream,                    \
        n, net.NH, net.hidden, net.A, net.dueling, h->acts_on.hid, h->b_dq, actions, P + net.off_hd_w[0],              \
        P + net.off_hd_w[1], h->d_hid, G + net.off_hd_w[0], G + net.off_hd_b[0], G + net.off_hd_w[1],                  \
        G + net.off_hd_b[1], G + net.off_fc_b[0], G + net.off_fc_b[1], td)
    if (net.A <= 4) HEAD_BWD(4);
    else if (net.A <= 8) HEAD_BWD(8);
    else if (net.A <= 16) HEAD_BWD(16);
    else HEAD_BWD(31);
#undef HEAD_BWD
    DQN_CUDA_OK(cudaGetLastError());
    h->launches++;
}
