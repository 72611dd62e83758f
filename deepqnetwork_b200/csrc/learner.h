// Internal state of a dqn_learner handle.
#pragma once
#include "common.cuh"
#include "../../include/dqn_b200.h"
#include <vector>
#include <string>
#include <utility>

struct ConvGeom {
    int ih, iw, cin, oh, ow, cout, k, s, pt, pl;
    int cin_log2;
    // ceil(2^32/d) magics: x/d == __umulhi(x, magic) for x*d < 2^32 (all index ranges here)
    unsigned m_k, m_ow, m_hw, m_kt, m_s;
    int kt;  // k / s: taps per axis that reach one stride-parity class
};
__host__ __device__ inline unsigned fastdiv_magic(unsigned d) { return d <= 1 ? 0u : (unsigned)((0x100000000ull + d - 1) / d); }
__device__ __forceinline__ int fdiv(int x, unsigned magic, int d) { return d <= 1 ? x : (int)__umulhi((unsigned)x, magic); }

struct TensorInfo {
    std::string name;
    int64_t offset;  // floats from slab base (16-byte aligned)
    int64_t count;
};

struct NetDesc {
    int H, W, C, A, hidden, dueling, NH;  // NH = hidden * (dueling ? 2 : 1)
    ConvGeom conv[3];
    int flat;
    int64_t off_cw[3], off_cb[3];
    int64_t off_fc_w[2], off_fc_b[2];  // hidden layers: [0]=advantage (or the only) stream, [1]=value stream
    int64_t off_hd_w[2], off_hd_b[2];  // heads: [0]=advantage/Q head [hidden,A], [1]=value head [hidden,1]
    int64_t slab;                      // padded float count of one tower
    std::vector<TensorInfo> tensors;   // TF creation order
};

// activations of one tower forward over `rows` samples
struct TowerActs {
    float* act[3];   // post-ReLU conv outputs, NHWC
    float* fc_part;  // split-K partials [splits][rows][NH]
    float* hid;      // post-ReLU hidden [rows][NH]
    float* q;        // [rows][A]
};

struct dqn_learner {
    dqn_config cfg;
    NetDesc net;
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // second stream + events: independent kernels of one step (target tower vs online tower, wgrads vs the dgrad
    // chain) are forked onto it so that the small batch-32 grids overlap; inside a captured graph these become
    // parallel branches
    cudaStream_t side_stream = nullptr, side2_stream = nullptr, side3_stream = nullptr, side4_stream = nullptr, side5_stream = nullptr, side6_stream = nullptr;
    cudaEvent_t fork_ev[24] = {};
    // data-parallel split step: recorded inside phase 0 -- [0] the dense-kernel gradients (98.6 % of the slab) are complete,
    // [1] the whole gradient slab is complete -- so that an external all-reduce can start on the big bucket while the conv
    // backward chain is still running (dqn_phase_wait)
    cudaEvent_t phase_ev[2] = {nullptr, nullptr};
    bool phase_ev_armed = false;  // set while phase 0 of a split step is being issued
    int fork_used = 0;
    bool overlap = true;
    int prio_hi = 0, prio_lo = 0;  // device stream-priority range (hi is numerically lower)
    int saved_prio = 0;
    bool use_prio = true;       // option: launch priorities by branch
    bool pdl = true;            // programmatic dependent launch for the kernels that call pdl_wait()
    bool ts_gemm = true;        // tcgen05 GEMMs take A from tensor memory where the operand functor allows it
    bool fuse_enabled = false;  // allow the fused dense-wgrad+optimizer epilogue in dqn_train_step(s)
    bool fuse_fc_adam = false;
    bool early_fc_adam = false;  // set per launch sequence: optimizer for the dense kernels runs inside backward (side stream)
    bool fc_adam_done = false;
    bool fc_ffma_adam = false;    // option: dense wgrad + optimizer as one fp32 FFMA pass (no gradient round trip); measured
                                  // slower than tensor-core wgrad + streaming Adam (instruction-bound: 61 us vs 17 + 45 us)
    bool fc_split_streams = true; // dense wgrad + optimizer per hidden stream on two side streams (pipelines HBM- and tensor-bound halves)
    int fc_tma_ctas = 0;          // > 0: the persistent form of that pass with this many CTAs (S-stage TMA ring of optimizer state)
    float* act3_pk[2] = {nullptr, nullptr};  // conv3 output of the online (64 rows) / target (32 rows) tower as the dense layer's
                                             // packed B tiles (EpiBiasReluPack -> fcfwd.cuh); valid for the launch that follows
    bool act3_pk_valid[2] = {false, false};
    void* ff_maps[2] = {nullptr, nullptr};  // ff::WMaps of the online / target dense kernels (fcfwd.cuh), built on first use
    int fc_fwd_share = 0;         // option: dense forward CTAs split 2 : 1 between the online (64 rows) and the target (32 rows) launch
                                  // (measured: no difference, 5925 steps/s either way inside the +-4 % instance-to-instance spread)
    bool fc_tma_fwd = true;       // dense forward with TMA-streamed weights and packed activation tiles (fcfwd.cuh)
    bool keep_grads = false;      // option: dqn_train_batch leaves the dense-kernel gradients in the slab (slot 3 read-back, parity
                                  // tests); off: the fused dense wgrad+optimizer kernel runs there too and never writes them
    bool wg1_fit = false;         // option: conv1 wgrad split count sized for the SMs the persistent dense kernel leaves free (measured -4 %)
    unsigned int* fa_tile_ctr = nullptr;  // dynamic tile counter of the fused dense kernel (= opt_ticket + 1, zeroed by the optimizer tail)
    void* fa_row_maps = nullptr;  // fa::RowMaps
    float* w_scratch = nullptr;   // [2][flat][hidden]: the fused dense kernel writes the updated dense weights here ...
    bool w_scratch_on = false;    // ... (option w_scratch) so that it can start BEFORE the dense dgrad has read the old ones; copied over the slab
                                  // afterwards.  Measured (profiles/r02_trace_w_scratch.txt): the pass then runs 82-135 us instead of 101-157, but its
                                  // 100 persistent CTAs starve the conv wgrads (conv3 wgrad 18 -> 55 us) and the two 15.7 MB copies run as SM kernels
                                  // (30 us) in front of the optimizer tail: 5780 vs 5950-6060 steps/s.  Off.
    int fc_rows_ctas = 100;       // dense wgrad + optimizer as the row-contiguous persistent kernel (fcadam.cuh) on this many SMs;
                                  // 0: tcgen05 wgrad into the gradient slab + streaming Adam.  clamped to the SM count.  With 148 the step time is BIMODAL (5400-5600 vs 6150 steps/s per learner instance): the kernel
                                  // becomes ready at the same instant as conv3's dgrad + wgrad (96 CTAs); if its persistent CTAs are dispatched first
                                  // they hold every SM and the critical conv chain waits 40 us (CUPTI traces of both modes).  With <= 100 CTAs the
                                  // chain always finds SMs and dynamic tile claims keep the pass balanced: 84-100 -> 6030-6180 in both modes
    int l2_keep = 0;              // L2 residency hints on the TMA copies (fcadam.cuh): bit 0 = dense-kernel m, v evict_last; bit 1 = online dense
                                  // weights evict_last (optimizer pass + dense forward); bit 2 = target dense weights evict_first
    void* fa_maps = nullptr;      // fa::StateMaps (TMA tensor maps of w/m/v of the dense kernels), built on first use
    bool fc_tma_adam = false;     // option: tcgen05 dense wgrad + optimizer in one pass, optimizer state staged by TMA (fcadam.cuh);
                                  // measured 67 us alone vs 23 + 40 us for wgrad + streaming Adam (too few bytes in flight per CTA)
    bool fc_grads_needed = false; // set per launch sequence: that pass also stores dW (seam-2 / parity path)  // set per launch sequence: dense wgrad applies the optimizer in its epilogue
    std::string err;
    int64_t launches = 0;

    // parameters
    float *online = nullptr, *target = nullptr, *slot_m = nullptr, *slot_v = nullptr, *grads = nullptr;
    float* slab_base = nullptr;  // the five slabs above are carved from this one allocation
    // device scalars: [0]=step (int64), game_count, beta powers, PER stats
    struct Scalars {
        long long step;
        long long game_count;
        float b1p, b2p;
        float stat_min, stat_max, stat_avg, stat_total;  // raw leaf stats (SumTree.get_*)
        float last_max_weight;                            // deep_q_agent.py:510
        float per_b;                                      // beta used by the last sample
        float loss;                                       // scalar loss of the last step
        float loss_sum;                                   // running sum for the report line
        int has_max_weight;
        unsigned long long rng_counter;
        float b1p_prev, b2p_prev;  // beta powers of the step before the last advance (deferred dense-kernel optimizer pass)
    };
    unsigned int* opt_ticket = nullptr;  // last-CTA-done counter of optimizer_tail_kernel
    Scalars* d_sc = nullptr;
    Scalars* h_sc = nullptr;  // pinned mirror for reads
    float* tail_mirror = nullptr;  // set while dqn_train_batch captures its graph: optimizer tail writes scalars+losses to h_sc
    int tail_mirror_words = 0;
    // host-owned sizes mirrored on the device so that captured graphs see current values
    struct DevSizes { long long tree_size; long long ring_len; };
    DevSizes* d_dev = nullptr;
    int64_t host_step = 0;     // host mirror of Scalars::step (cadence decisions without a device read)
    bool has_max_weight = false;
    bool external_normaliser = false;  // option: dqn_set_per_normaliser is the only writer of last_max_weight

    // replay ring (SoA, HBM)
    int64_t cap = 0, ring_size = 0, ring_cur = 0;
    bool ring_full = false;
    size_t state_bytes = 0;
    uint8_t *r_states = nullptr, *r_next = nullptr, *r_actions = nullptr, *r_rewards = nullptr, *r_cont = nullptr;
    __half* r_losses = nullptr;
    // frame-de-duplicated layout (cfg.frame_dedup): frames stored once, transitions hold 8 frame slots
    bool dedup = false;
    int64_t fcap = 0, f_written = 0;   // frame ring slots, frames appended so far (slot = count % fcap)
    size_t frame_bytes = 0;            // H * W
    uint8_t* r_frames = nullptr;       // [fcap][H][W]
    uint32_t* r_fidx = nullptr;        // [cap][8]: s channels 0..3, s' channels 0..3

    // sum tree
    float* tree = nullptr;       // 2*cap-1
    uint32_t* tree_data = nullptr;  // cap
    int64_t tree_size = 0, tree_write = 0;
    bool tree_max_reached = false;

    // train batch (device): rows [0,B) = s, [B,2B) = s'
    int B = 0, max_rows = 0;
    uint8_t* b_states = nullptr;  // sampled batch [2*max_rows][S]
    uint8_t *b_actions = nullptr, *b_rewards = nullptr, *b_cont = nullptr;
    // host-fed batch (train_batch / predict / get_losses), kept apart from the sampled one
    uint8_t* in_states = nullptr;
    uint8_t *in_actions = nullptr, *in_rewards = nullptr, *in_cont = nullptr;
    float* in_isw = nullptr;
    size_t in_small_bytes = 0;         // in_actions/in_rewards/in_cont/in_isw are carved from one device block
    uint8_t* h_in_small = nullptr;     // pinned mirror of that block (also stages the per-step tree write-back)
    cudaEvent_t stage_ev = nullptr;    // last copy out of h_in_small
    long long* b_rows = nullptr;       // uniform-sampler rows
    __half* b_losses = nullptr;
    long long* b_tree_idx = nullptr;   // [B]
    long long* b_data_idx = nullptr;
    long long* b_mem_idx = nullptr;
    double* b_prio = nullptr;          // f32 tree values widened, as the reference's float array holds them
    double* b_isw64 = nullptr;
    float* b_isw = nullptr;            // fed as f32
    double* b_u = nullptr;             // uniforms
    float *b_y = nullptr, *b_maxq = nullptr, *b_losses_out = nullptr, *b_dq = nullptr, *b_newprio = nullptr;

    // The sampled batch is double-buffered: inside the graph of step i the batch of step i+1 is sampled and gathered
    // (side branch, right after the priority write-back of step i) into the other set, so sampling leaves the critical
    // path.  The b_* pointers above always alias sets[cur]; select_set() switches them.
    struct BatchSet {
        uint8_t *states, *actions, *rewards, *cont;
        __half* losses;
        long long *tree_idx, *data_idx, *mem_idx, *rows;
        double *prio, *isw64, *u;
        float *isw, *x_f32;
        long long* blk;  // tree_idx | prio | isw64 | data_idx | mem_idx | u carved from one block: one D2H for the seam outputs
    } sets[2] = {};
    uint8_t* h_stage2 = nullptr;   // pinned: uniforms in, {tree_idx, prio, isw} out of dqn_per_sample
    cudaEvent_t stage2_ev = nullptr;
    int cur_set = 0;
    bool pf_valid = false;    // sets[pf_set] holds the batch the next fused step consumes (sampled ahead)
    int pf_set = 0;
    bool pf_capture = false;  // set while capturing a pipelined step graph: step_core appends the look-ahead sample
    int sample_ahead = 0;     // the sampler reads step / rng counter + sample_ahead (1 for the look-ahead sample)
    bool prefetch = true;     // option

    // activations / workspaces
    TowerActs acts_on, acts_tg;
    float *d_hid = nullptr, *d_act[3] = {nullptr, nullptr, nullptr};
    float* wg_part = nullptr;  // split-K partials for conv wgrad
    int64_t wg_part_floats = 0;   // per layer (64 partials of (mreal + 4) x cout)
    int64_t wg_off[3] = {0, 0, 0};  // each layer has its own slice: the three weight gradients run on parallel streams
    int fc_splits = 1;

    // tcgen05 path helpers
    unsigned long long* tc_dbg = nullptr;  // optional in-kernel timeline capture (dqn_debug_timeline)
    float* zeros16 = nullptr;   // 16 zero bytes (cp.async dummy source)
    float* ones_row = nullptr;  // {1,0,0,0}: the bias row appended to conv wgrad A^T
    float* x_f32 = nullptr;     // u8 batch converted to f32 (x/255) for cp.async loaders: [2*max_rows][S]
    // image-resident conv kernels (cvgemm.cuh): pre-split, pre-tiled conv weights per tower ([0]=online, [1]=target),
    // kept in sync with the parameter slabs by launch_pack_conv (after every optimizer step / copy / set_param)
    float* packed[2] = {nullptr, nullptr};
    int64_t pk_off[4] = {0, 0, 0, 0};  // [3]: layer 0 in the uint8 form of the fused tower (cvgemm.cuh pk_store_u8)
    float* packed_dg = nullptr;            // online tower, dgrad operands (flipped / transposed, per stride class) of conv2, conv3
    int64_t pkd_off[3] = {0, 0, 0};
    int64_t pk_floats = 0;
    bool fuse_td = true;        // option: TD epilogue inside the head backward kernel (training steps)
    bool pack_in_tail = true;   // option: optimizer_tail_kernel re-packs the conv weights it updates (no separate pack launch)
    bool img_wgrad1 = true;     // option: conv1 wgrad through the image-resident kernel on the uint8 frames
    bool merge_towers = true;   // option: conv layers of the online and the target tower in one launch each
    bool img_dgrad = true;      // option: conv dgrads through the image-resident kernel as well
    bool img_conv = true;       // option: conv forward through the image-resident kernel where the geometry allows
    bool tower_fused = true;    // option: the conv stack of both towers as one fused kernel on the uint8 batch (cvtower.cuh)
    int tail_join = 1;          // option: the optimizer tail does not wait for the fused dense wgrad+optimizer kernel; the two run side by side and
                                // the end-of-step bookkeeping (step counter, beta powers, counters, host mirror) is done by the last CTA of whichever
                                // finishes second (step_join_arrive / step_bookkeeping below).  Pure rescheduling: bit-identical results
    int fc_after_dgrad = 0;     // option (measured: +0.3 % alone, -4 % together with tail_join -- the persistent grid then takes its SMs before conv3's dgrad and wgrad get theirs): the fused dense kernel's branch forks behind the dense dgrad GEMM (the last reader of the old dense
                                // weights), not behind the sum+gate kernel that follows it on the main stream
    bool rows_joined = false;   // set per launch sequence: the fused dense kernel takes part in the end-of-step join
    int fc_rows_pf = 0;         // option: the fused dense kernel prefetches the tile fc_rows_pf x CTAs claims ahead into L2 (0 = off; measured: 6550-6630 vs 6650-6670 steps/s without -- the pass is not latency-bound in situ)
    bool fc_prefetch = true;    // option: the dense hidden kernels are prefetched into L2 under the conv towers (nn.cu launch_dense_prefetch)
    bool tower_ext2 = false;    // option: conv3 of the fused tower streams its weights through 8 stages (4 of them in image A, free by then).
                                // Measured SLOWER (6537-6567 vs 6637-6651 steps/s): the layer's k-loop is not bound by the depth of the weight
                                // ring but by shared-memory bandwidth -- per k-block 16 KB of TMA writes + 16 KB of converter reads + 16 KB of
                                // MMA operand reads = 384 cycles at 128 B/clk, against 256 cycles of tensor time
    int tower_cluster = 1;      // option: CTAs (images) per thread-block cluster of the fused tower sharing one multicast weight stream (1, 2, 4, 8).
                                // Measured 6605-6633 steps/s without clusters, 6515-6589 with 4, 6549 with 8: the weight stream of a CTA is bound by the
                                // round trip of its 4 stages, not by L2 bandwidth, and a cluster adds the skew of its slowest CTA to every stage
    bool xf_valid = false;      // x_f32 holds the f32 copy of the batch the current step trains on
    bool xf_from_gather = false;  // the gather that filled b_states also wrote x_f32 (consumed by the next step_core)
    // staging
    uint8_t* dev_stage = nullptr;
    size_t dev_stage_bytes = 0;
    uint8_t* stats_scratch = nullptr;

    // graph for the fused step
    cudaGraphExec_t step_graph = nullptr;
    // pipelined step consuming sets[k], sampling into sets[1-k]; [k][1] additionally starts with the deferred optimizer
    // pass over the dense kernels of the PREVIOUS step (it runs under this step's conv forward instead of crowding the
    // previous step's backward), [k][0] is the first step after anything else touched the handle
    cudaGraphExec_t pipe_graph[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    bool defer_adam = false;         // option (bit-identical results); measured 3% slower: the 60 us pass does not fit under the 40 us conv forward
    bool defer_capture = false;      // capturing a deferred-optimizer graph: backward leaves the dense kernels to the next step
    bool defer_start = false;        // capturing variant [k][1]: step_core opens with the deferred pass
    bool fc_adam_pending = false;    // host: dense-kernel gradients of the last step are in the slab, not yet applied
    bool graph_valid = false;
    // graph of one host-fed step (dqn_train_batch), keyed by the batch size
    cudaGraphExec_t batch_graph = nullptr;
    int batch_graph_n = 0;
    bool graph_valid_batch = false, batch_graph_enabled = true;
    // graph of one fused step fed HOST draws (dqn_train_step with h_u / h_rows: the drop-in agent's call): the draws ride a
    // pinned staging block (fixed address -> memcpy node), scalars + losses come back through the optimizer tail's mirror
    cudaGraphExec_t ustep_graph = nullptr, ustep_graph_alt = nullptr;
    bool graph_valid_u = false, graph_valid_u_alt = false;
    // asynchronous form (dqn_train_step_async / dqn_train_step_result): two slots {pinned draws, pinned result mirror, graph, event}
    uint8_t* h_stage2_alt = nullptr;
    Scalars* h_sc_alt = nullptr;
    cudaEvent_t ustep_ev[2] = {nullptr, nullptr};
    int ustep_slot = 0, ustep_read = 0, ustep_pending = 0;
    int64_t ustep_graph_launches = 0;
    int64_t batch_graph_launches = 0;
    int64_t graph_launches = 0;

    // per-segment CUDA-event timing (dqn_set_profile)
    int profile = 0;
    bool prof_capturing = false;
    cudaGraphExec_t prof_graph = nullptr;
    const char* prof_tag = "";  // "on_" / "tg_": which tower the forward marks belong to
    std::vector<cudaEvent_t> prof_events;
    std::vector<const char*> prof_marks;
    size_t prof_used = 0;
    std::vector<std::pair<const char*, float>> prof_acc;
    int64_t prof_steps = 0;
};

// ---- kernels / launchers (defined across the .cu files) ----
void prof_mark(dqn_learner* h, const char* name);
// stream fork/join helpers (api.cu). side_begin: side stream waits for everything issued on the main stream so far
// and becomes the launch stream; side_end: back to the main stream; main_wait_side / side_wait_main: cross edges.
bool overlap_enabled(const dqn_learner* h);
// prio_class: launch priority of the branch (common.cuh g_launch_priority): 0 = work with slack (dense wgrad, optimizer),
// 1 = needed soon (conv wgrads, look-ahead sampling), 2 = critical path (same as the main stream)
void side_begin(dqn_learner* h, cudaStream_t& saved, int which = 0, int prio_class = 0);
void side_end(dqn_learner* h, cudaStream_t saved);
cudaEvent_t fork_event_here(dqn_learner* h);  // an event recorded on the current stream at this point of the launch sequence (nullptr: no overlap)
void side_begin_at(dqn_learner* h, cudaStream_t& saved, int which, int prio_class, cudaEvent_t after);  // side_begin forking at `after` (nullptr: here)
void main_wait_side(dqn_learner* h, int which = 0);
void launch_optimizer_fc(dqn_learner* h);
void launch_optimizer_fc_stream(dqn_learner* h, int st, bool prev_betas = false);  // one hidden stream's dense kernel
bool fc_wgrad_adam_supported(const dqn_learner* h);
void launch_fc_wgrad_adam(dqn_learner* h, int rows, bool write_g);  // optim.cu: dense wgrad + optimizer in one FFMA pass  // api.cu: CUDA-event boundary, active only in profile mode
void net_describe(NetDesc& net, const dqn_config& cfg);

// replay.cu
// with_f32: also write the batch as f32/255 (the fused step feeds the tensor-core loaders from it)
void launch_gather(dqn_learner* h, const long long* d_rows, int n, bool with_f32 = false);
void launch_fill_synthetic(dqn_learner* h, int64_t start, int64_t n, uint64_t seed, float* d_prio);
// stacked [n][H][W][4] states of ring rows -> dst (frame_dedup: assembled from the frame ring); next: the s' stacks
void launch_gather_states(dqn_learner* h, const long long* d_rows, int n, bool next, uint8_t* dst);

void launch_preprocess(dqn_learner* h, const uint8_t* d_obs, int raw_h, int raw_w, int top, int stride, int oh, int ow, int64_t n,
                       uint8_t* d_out);

// sumtree.cu
// u_src: uniforms read from there instead of b_u (may be pinned host memory); host_out: pinned host block that also
// receives the {tree_idx | priority | is_weight} columns (stride max_rows, 8-byte entries)
// raw_probe: the "uniforms" are the probes themselves (SumTree.get_with_info(s))
void launch_tree_sample(dqn_learner* h, int batch, int device_rng, double per_b, const double* u_src = nullptr,
                        void* host_out = nullptr, bool raw_probe = false);
void launch_tree_update(dqn_learner* h, int n, const long long* d_tree_idx, const float* d_scores, int store_losses);
void launch_tree_add(dqn_learner* h, int64_t n, const float* d_scores, const uint32_t* d_data, int64_t write0);
// d_out4 == nullptr: refresh the handle's PER state (stat_*, last_max_weight; deep_q_agent.py:505-510 cadence);
// otherwise {min, max, avg, total} go to d_out4 and the handle's state is left alone (read-only getters)
void launch_tree_stats(dqn_learner* h, double per_b_override, float* d_out4 = nullptr);

// nn.cu
// x_f32: the same rows as d_states already converted to f32 (used by the tcgen05 path; may be null in FP32 mode)
void tower_forward(dqn_learner* h, const float* params, const uint8_t* d_states, const float* x_f32, int rows, TowerActs& acts);
void tower_backward(dqn_learner* h, const uint8_t* d_states, const float* x_f32, int rows);
bool towers_conv_forward_merged(dqn_learner* h, const float* x_on, int rows_on, const float* x_tg, int rows_tg);
bool tower_fused_ready(const dqn_learner* h);
bool towers_conv_forward_fused(dqn_learner* h, const float* P_on, const uint8_t* s_on, int rows_on, TowerActs& acts_on,
                               const uint8_t* s_tg, int rows_tg, int act_keep_rows = -1);  // -1: activations of every image are stored
void tower_fc_forward(dqn_learner* h, const float* params, int rows, TowerActs& acts);  // dense + heads after the conv stack
void alloc_packed_conv(dqn_learner* h);          // sizes + device buffers of the packed conv weights
void launch_pack_conv(dqn_learner* h, int tower);
bool dgrad_pack_ok(const dqn_learner* h, int layer);  // nn.cu: layer has a packed dgrad operand (cv::dgrad_geom)  // re-pack one tower's conv weights from its parameter slab
void launch_u8_to_f32(dqn_learner* h, const uint8_t* src, float* dst, size_t n);
void launch_dense_prefetch(dqn_learner* h, bool with_target);
int selftest_gemm(dqn_learner* h, int mode, int M, int N, int K, int a_kcontig, int b_kcontig, const float* dA, const float* dB, float* dC);

// heads.cu
void launch_td_epilogue(dqn_learner* h, int n, const float* q_on, const float* q_on_next, const float* q_tg_next,
                        const uint8_t* actions, const uint8_t* rewards, const uint8_t* cont, const float* isw,
                        int train, double grad_scale);
// fuse != nullptr: the TD epilogue of a training step is computed inside the head backward kernel (no separate launch)
struct TdFuse { const float *q_on, *q_on_next, *q_tg_next; const uint8_t *rewards, *cont; const float* isw; double grad_scale; };
void launch_head_backward(dqn_learner* h, int n, const uint8_t* actions, const TdFuse* fuse = nullptr);

// optim.cu
void launch_optimizer(dqn_learner* h, bool skip_fc);
void launch_post_step(dqn_learner* h, int per_update);
void launch_copy_network(dqn_learner* h);

#ifdef __CUDACC__
// End-of-step bookkeeping, run by exactly ONE thread per step.  opt_ticket layout: [0] CTAs of the optimizer tail that are done,
// [1] tile counter of the fused dense wgrad+optimizer kernel, [2] CTAs of that kernel that are done, [3] kernels that are done.
// The last CTA of the optimizer tail and (option tail_join) the last CTA of the fused dense kernel arrive at [3]; the arrival that
// completes `expected` runs step_bookkeeping.  Both kernels read the beta powers when they start, and nothing else reads the
// scalars before the step's graph has ended, so it does not matter which of the two it is.
__device__ __forceinline__ bool step_join_arrive(unsigned int* ticket, unsigned int expected) {
    __threadfence();
    return atomicAdd(ticket + 3, 1u) == expected - 1u;
}
__device__ __forceinline__ void step_bookkeeping(dqn_learner::Scalars* sc, unsigned int* ticket, int use_momentum, float* host_mirror,
                                                 int mirror_words) {
    ticket[0] = 0; ticket[1] = 0; ticket[2] = 0; ticket[3] = 0;  // ready for the next step (no memset nodes in the graph)
    sc->step += 1;
    sc->b1p_prev = sc->b1p; sc->b2p_prev = sc->b2p;
    if (!use_momentum) { sc->b1p *= 0.9f; sc->b2p *= 0.999f; }
    sc->rng_counter += 1;
    // host-fed step (dqn_train_batch / dqn_train_step with host draws): the scalar block and the per-sample losses behind it go
    // straight into their pinned, device-mapped mirror -- the step needs no read-back copy operation after its last kernel
    if (host_mirror) {
        __threadfence();
        const volatile float* src = reinterpret_cast<const volatile float*>(sc);
        for (int i = 0; i < mirror_words; ++i) host_mirror[i] = src[i];
        __threadfence_system();
    }
}
#endif
