// Fused TD epilogue and head backward.
//
// Replaces (reference file:line):
//   rl/deep_q_model.py:247-255  max_q_values: reduce_max(target_q * one_hot(argmax(online_q))) (double) / reduce_max(target_q)
//   rl/deep_q_model.py:106-114  host target y = rewards(u8) + continues(bool) * discount * max_q, NumPy float64, fed as f32
//   rl/deep_q_model.py:361-378  q_a = sum(Q * one_hot(a)); abs / huber(delta=1) / is_w * squared_difference; reduce_mean
//   rl/deep_q_agent.py:527-534  _make_priority: min(|delta| + 0.01, 1.0) ** per_a  (float32)
// and the autodiff of the loss w.r.t. the head inputs (dueling combine, the two tiny head matmuls, hidden ReLU).
#include "learner.h"

// one thread per sample; a single CTA so the scalar loss is reduced in a fixed order
__global__ void __launch_bounds__(1024) td_epilogue_kernel(
    int n, int A, int use_double, int use_per, double gamma, float per_a, double grad_scale, int train,
    const float* __restrict__ q_on, const float* __restrict__ q_on_next, const float* __restrict__ q_tg_next,
    const uint8_t* __restrict__ actions, const uint8_t* __restrict__ rewards, const uint8_t* __restrict__ cont,
    const float* __restrict__ isw, float* __restrict__ o_y, float* __restrict__ o_maxq, float* __restrict__ o_losses,
    float* __restrict__ o_dq, float* __restrict__ o_newprio, dqn_learner::Scalars* sc) {
    __shared__ double s_red[32];
    double lsum = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float* qt = q_tg_next + (size_t)i * A;
        float maxq;
        if (use_double) {
            const float* qo = q_on_next + (size_t)i * A;
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
        const double y64 = (double)rewards[i] + ((cont[i] ? 1.0 : 0.0) * gamma) * (double)maxq;
        const float y = (float)y64;  // fed through a tf.float32 placeholder (deep_q_model.py:206)
        const float qa = q_on[(size_t)i * A + actions[i]];
        const float delta = y - qa;
        const float ad = fabsf(delta);
        float li, out_l, dq;
        if (use_per) {
            const float w = isw[i];
            li = w * (delta * delta);  // is_weights * squared_difference (deep_q_model.py:374)
            out_l = ad;                // self.losses = abs_losses (:375)
            dq = (float)(-2.0 * (double)w * (double)delta * grad_scale / (double)n);
        } else {
            const float quad = fminf(ad, 1.0f);  // tf.losses.huber_loss, delta=1
            li = 0.5f * quad * quad + (ad - quad);
            out_l = li;
            const float e = -delta;  // predictions - labels
            dq = (float)((double)fminf(fmaxf(e, -1.0f), 1.0f) * grad_scale / (double)n);
        }
        o_y[i] = y; o_maxq[i] = maxq; o_losses[i] = out_l;
        if (train) {
            o_dq[i] = dq;
            if (use_per) o_newprio[i] = powf(fminf(ad + 0.01f, 1.0f), per_a);
        }
        lsum += (double)li;
    }
    for (int o = 16; o; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
        const float loss = (float)(t / (double)n);
        sc->loss = loss;
        if (train) sc->loss_sum += loss;
    }
}

void launch_td_epilogue(dqn_learner* h, int n, const float* q_on, const float* q_on_next, const float* q_tg_next,
                        const uint8_t* actions, const uint8_t* rewards, const uint8_t* cont, const float* isw,
                        int train, double grad_scale) {
    int threads = ((n + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    launch_kernel(false, td_epilogue_kernel, dim3(1), dim3(threads), 0, h->stream, n, h->net.A, h->cfg.use_double, h->cfg.use_per,
                  h->cfg.discount_rate, (float)h->cfg.per_a, grad_scale, train, q_on, q_on_next, q_tg_next, actions, rewards, cont,
                  isw, h->b_y, h->b_maxq, h->b_losses_out, h->b_dq, h->b_newprio, h->d_sc);
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
                                                             float* __restrict__ g_fb1) {
    extern __shared__ float sm[];
    float* s_dh = sm;                 // [32][33]
    float* s_hv = sm + 32 * 33;       // [32][33]
    float* s_dq = sm + 2 * 32 * 33;   // [n]
    uint8_t* s_act = reinterpret_cast<uint8_t*>(s_dq + n);  // [n]
    for (int i = threadIdx.x; i < n; i += blockDim.x) { s_dq[i] = dq[i]; s_act[i] = actions[i]; }
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

void launch_head_backward(dqn_learner* h, int n, const uint8_t* actions) {
    const NetDesc& net = h->net;
    const float* P = h->online;
    float* G = h->grads;
    const size_t smem = (size_t)2 * 32 * 33 * 4 + (size_t)n * 4 + ((n + 3) / 4) * 4;
    DQN_REQUIRE(net.NH % 32 == 0 && net.A <= 31, "head backward tiling");
#define HEAD_BWD(AMAX)                                                                                                  \
    launch_kernel(false, head_backward_kernel<AMAX>, dim3(net.NH / 32), dim3(1024), smem, h->stream,                    \
        n, net.NH, net.hidden, net.A, net.dueling, h->acts_on.hid, h->b_dq, actions, P + net.off_hd_w[0],              \
        P + net.off_hd_w[1], h->d_hid, G + net.off_hd_w[0], G + net.off_hd_b[0], G + net.off_hd_w[1],                  \
        G + net.off_hd_b[1], G + net.off_fc_b[0], G + net.off_fc_b[1])
    if (net.A <= 4) HEAD_BWD(4);
    else if (net.A <= 8) HEAD_BWD(8);
    else if (net.A <= 16) HEAD_BWD(16);
    else HEAD_BWD(31);
#undef HEAD_BWD
    DQN_CUDA_OK(cudaGetLastError());
    h->launches++;
}
