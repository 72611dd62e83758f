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
#include <algorithm>

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v,
                                                   const float4* __restrict__ g, size_t n4, float lr,
                                                   const dqn_learner::Scalars* __restrict__ sc) {
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float lr_t = lr * sqrtf(1.0f - sc->b2p) / (1.0f - sc->b1p);
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

static void optimizer_range(dqn_learner* h, int64_t beg, int64_t end) {
    if (end <= beg) return;
    const size_t n4 = (size_t)(end - beg) / 4, o4 = (size_t)beg / 4;
    const int blocks = (int)std::min<size_t>((n4 + 255) / 256, (size_t)h->sm_count * 8);
    if (h->cfg.use_momentum)
        momentum_kernel<<<blocks, 256, 0, h->stream>>>((float4*)h->online + o4, (float4*)h->slot_m + o4,
                                                       (const float4*)h->grads + o4, n4, (float)h->cfg.learning_rate,
                                                       (float)h->cfg.momentum);
    else
        adam_kernel<<<blocks, 256, 0, h->stream>>>((float4*)h->online + o4, (float4*)h->slot_m + o4, (float4*)h->slot_v + o4,
                                                   (const float4*)h->grads + o4, n4, (float)h->cfg.learning_rate, h->d_sc);
    DQN_CUDA_OK(cudaGetLastError());
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

// skip_fc: the two dense hidden kernels were already updated in the wgrad epilogue (EpiFcWgradAdam)
void launch_optimizer(dqn_learner* h, bool skip_fc) {
    const NetDesc& net = h->net;
    if (!skip_fc) return optimizer_range(h, 0, net.slab);
    const int64_t fcw = (int64_t)net.flat * net.hidden;
    optimizer_range(h, 0, net.off_fc_w[0]);
    if (net.dueling) {
        optimizer_range(h, net.off_fc_w[0] + fcw, net.off_fc_w[1]);
        optimizer_range(h, net.off_fc_w[1] + fcw, net.slab);
    } else {
        optimizer_range(h, net.off_fc_w[0] + fcw, net.slab);
    }
}

// after the optimizer: global_step += 1, beta powers advance (AdamOptimizer._finish), RNG counter advances
__global__ void advance_step_kernel(dqn_learner::Scalars* sc, int adam) {
    sc->step += 1;
    if (adam) { sc->b1p *= 0.9f; sc->b2p *= 0.999f; }
    sc->rng_counter += 1;
}

void launch_tree_update(dqn_learner* h, int n, const long long* d_tree_idx, const float* d_scores, int store_losses);

void launch_post_step(dqn_learner* h, int per_update) {
    advance_step_kernel<<<1, 1, 0, h->stream>>>(h->d_sc, h->cfg.use_momentum ? 0 : 1);
    DQN_CUDA_OK(cudaGetLastError());
    h->launches++;
    // replay_sampler.update_sum_tree(tree_idxes, _make_priority(losses))  (deep_q_agent.py:341-347)
    if (per_update) launch_tree_update(h, h->B, h->b_tree_idx, h->b_newprio, 1);
}

void launch_copy_network(dqn_learner* h) {
    DQN_CUDA_OK(cudaMemcpyAsync(h->target, h->online, (size_t)h->net.slab * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
}
