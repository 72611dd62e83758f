/*
 * dqn_b200.h -- C ABI of the B200-native DQN learner (libdqn_b200.so).
 *
 * Drop-in boundary for the training-step hot path of reynkun/DeepQNetwork.
 * Every entry point cites the reference interface it replaces (file:line into
 * the reference tree).  Plain pointers and sizes only; all functions return 0
 * on success and a non-zero status on failure, with the message available from
 * dqn_last_error().  CUDA errors are fatal for the handle: there is no CPU
 * fallback anywhere behind this interface.
 *
 * Pointers named h_* are HOST pointers (pageable or pinned); the library does
 * the host<->device copies on the handle's stream.  One handle = one CUDA
 * device + one stream; a handle is not thread-safe.
 */
#ifndef DQN_B200_H
#define DQN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dqn_learner dqn_learner;

/* conv stack variants */
#define DQN_CONV_REFERENCE 0 /* rl/deep_q_model.py:278-283: k 8/4/4, s 4/2/1, padding 'same' */
#define DQN_CONV_NATURE    1 /* labelled extra: k 8/4/3, s 4/2/1, 'valid'       */

/* matmul arithmetic modes */
#define DQN_MATH_FP32  0 /* FFMA fp32 on CUDA cores                                   */
#define DQN_MATH_TC3X  1 /* tcgen05 kind::tf32, 3-term split (fp32-grade accuracy)    */
#define DQN_MATH_TC1X  2 /* tcgen05 kind::tf32 single pass (reported separately)      */

typedef struct dqn_config {
    int32_t struct_size;      /* = sizeof(dqn_config); guards ABI drift */
    int32_t device;           /* CUDA ordinal */
    /* model: rl/deep_q_model.py:42-59 class attrs + :24-37 DEFAULT_OPTIONS */
    int32_t input_height, input_width, input_channels;
    int32_t num_actions;      /* conf['action_space'], deep_q_model.py:71 */
    int32_t num_hidden;       /* NUM_HIDDEN = 512 */
    int32_t conv_variant;     /* DQN_CONV_* */
    int32_t use_dueling, use_double, use_per, use_momentum;
    double  learning_rate;    /* 0.00025 */
    double  momentum;         /* 0.95, nesterov */
    double  discount_rate;    /* 0.99 */
    /* agent: rl/deep_q_agent.py:33-72 */
    int32_t batch_size;       /* 32 */
    int64_t replay_capacity;  /* replay_max_memory_length */
    double  per_a, per_b_start, per_b_end;
    int64_t per_anneal_steps;
    int32_t use_per_annealing;
    int64_t per_calculate_steps;  /* 5000 */
    int64_t copy_network_steps;   /* 10000 */
    int32_t max_rows;         /* largest n for predict/get_losses/train_batch (0 -> max(2*batch,256)) */
    int32_t math_mode;        /* DQN_MATH_* */
    uint64_t seed;            /* device Philox seed for throughput-mode sampling */
    /* Replay layout.  0: the reference's layout, full stacked s and s' per transition (rl/data/replay_memory.py:26-39,
     * 2 x H*W*4 bytes per row).  1: frame-de-duplicated -- every H x W frame is stored once in a frame ring and a
     * transition holds the 8 frame slots of its s / s' stacks (the 4-frame deque of rl/game_runner.py:139-148,236-244
     * re-uses 3 of 4 frames from one step to the next): H*W + 32 bytes per transition, so that the 10M-transition replay
     * of BASELINE.json configs[2] fits HBM.  The gathered batch, and everything downstream of it, is byte-identical. */
    int32_t frame_dedup;
    int64_t frame_capacity;   /* frame ring slots (frame_dedup only); 0 -> replay_capacity * 17/16 + 64 */
} dqn_config;

/* ---- lifecycle ------------------------------------------------------- */
/* replaces DeepQModel(conf)+Model.open (rl/deep_q_model.py:62-72, rl/model/model.py:86-99)
 * and DeepQAgent._train_init_memory (rl/deep_q_agent.py:197-237): allocates params
 * (online, target, optimizer slots), the replay ring, the sum tree and workspaces in HBM. */
int dqn_create(const dqn_config* cfg, dqn_learner** out);
int dqn_destroy(dqn_learner* h);
const char* dqn_last_error(const dqn_learner* h); /* h may be NULL: last create error */
int dqn_sync(dqn_learner* h);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the private stream */
int dqn_set_stream(dqn_learner* h, void* cuda_stream);

/* ---- parameters: names are the TF variable names of rl/deep_q_model.py:292-343
 * relative to the tower scope ("conv2d/kernel", "dense_1/bias", ...). tower: 0 online, 1 target.
 * slot: 0 value, 1 Adam m / Momentum accumulator, 2 Adam v.  Replaces tf.train.Saver access
 * (rl/model/model.py:60-83) at tensor granularity. */
int dqn_num_params(const dqn_learner* h, int32_t* n_tensors, int64_t* n_floats);
int dqn_param_info(const dqn_learner* h, int32_t i, const char** name, int64_t* count, int64_t* offset);
int dqn_get_param(dqn_learner* h, int32_t tower, int32_t slot, const char* name, float* h_out, int64_t count);
int dqn_set_param(dqn_learner* h, int32_t tower, int32_t slot, const char* name, const float* h_in, int64_t count);
/* raw device pointers, for torch.distributed all-reduce of the gradient slab and zero-copy views */
int dqn_device_ptr(dqn_learner* h, const char* what, void** ptr, int64_t* bytes);
int dqn_get_step(dqn_learner* h, int64_t* step);          /* DeepQModel.get_training_step, deep_q_model.py:168 */
int dqn_set_step(dqn_learner* h, int64_t step);
int dqn_get_game_count(dqn_learner* h, int64_t* n);        /* deep_q_model.py:176-188 */
int dqn_set_game_count(dqn_learner* h, int64_t n);
int dqn_get_beta_powers(dqn_learner* h, float* b1p, float* b2p);
int dqn_set_beta_powers(dqn_learner* h, float b1p, float b2p);
int dqn_copy_network(dqn_learner* h);                      /* DeepQModel.copy_network, deep_q_model.py:160-165 */
/* {min, max, avg, total, last_max_weight, per_b, last loss, running loss sum}: the PER fields of the
 * [train] report line (rl/deep_q_agent.py:386-392) */
int dqn_get_stats(dqn_learner* h, float* out8);

/* ---- replay ring: rl/data/replay_memory.py -------------------------- */
/* ReplayMemory.append x n (replay_memory.py:54-67). When use_per, also SumTree.add(loss, cur_idx)
 * per row first, as ReplaySamplerPriority.append does (replay_sampler_priority.py:22-24).
 * h_losses may be NULL (no PER / uniform sampler). states are [n,H,W,C] u8. */
int dqn_replay_append(dqn_learner* h, int64_t n, const uint8_t* h_states, const uint8_t* h_actions,
                      const uint8_t* h_rewards, const uint8_t* h_next_states, const uint8_t* h_continues,
                      const float* h_losses);
/* ReplayMemory.get for a list of rows (replay_memory.py:78-91); any out pointer may be NULL */
int dqn_replay_read(dqn_learner* h, int64_t n, const int64_t* h_rows, uint8_t* h_states, uint8_t* h_actions,
                    uint8_t* h_rewards, uint8_t* h_next_states, uint8_t* h_continues, uint16_t* h_losses_f16);
int dqn_replay_info(dqn_learner* h, int64_t* size, int64_t* cur_idx, int32_t* is_full); /* __len__, cur_idx, is_full */
/* frame_dedup layout: store n new H x W frames ([n,H,W] u8) in the frame ring; *first_slot receives the slot of the first
 * one, the others follow consecutively modulo frame_capacity.  A frame slot is reused after frame_capacity appends: the
 * caller keeps frame_capacity >= replay_capacity + the frames a transition can lag behind the newest one. */
int dqn_frames_append(dqn_learner* h, int64_t n, const uint8_t* h_frames, int64_t* first_slot);
/* frame_dedup layout: ReplayMemory.append x n with states given as frame slots: h_frame_slots is [n][8] u32 =
 * {s channel 0..3, s' channel 0..3}.  Same tree / loss-column semantics as dqn_replay_append. */
int dqn_replay_append_indexed(dqn_learner* h, int64_t n, const uint32_t* h_frame_slots, const uint8_t* h_actions,
                              const uint8_t* h_rewards, const uint8_t* h_continues, const float* h_losses);
int dqn_frames_info(dqn_learner* h, int64_t* frame_capacity, int64_t* frames_written);
/* GameEnvironment.preprocess_observation (rl/game_environment.py:110-116,134-140,154-160) for n raw RGB observations
 * ([n, raw_h, raw_w, 3] u8): rows top .. raw_h - bottom and all columns at the given stride, "grey" = dot(rgb, [0.299,
 * 0.587, 0.144]) in float64 with the reference's contraction, C cast to uint8 (values above 255 wrap).  h_frames
 * ([n, oh, ow] u8) may be NULL when to_frame_ring != 0: the frames then go straight into the frame ring of a frame_dedup
 * handle (*first_slot = slot of the first one).  SURVEY.md 8f-2. */
int dqn_preprocess(dqn_learner* h, int64_t n, const uint8_t* h_obs, int32_t raw_h, int32_t raw_w, int32_t top, int32_t bottom,
                   int32_t stride, uint8_t* h_frames, int32_t to_frame_ring, int64_t* first_slot);
int dqn_replay_clear(dqn_learner* h);                                                    /* replay_memory.py:46-52 */
/* bench/test fill on the device (SURVEY.md 8d "synthetic inputs"): n rows from Philox(seed);
 * when use_per the tree is built with the semantics of n sequential SumTree.add calls. */
int dqn_replay_fill_synthetic(dqn_learner* h, int64_t n, uint64_t seed);

/* ---- sum tree: rl/data/sum_tree.py ---------------------------------- */
int dqn_tree_add(dqn_learner* h, int64_t n, const float* h_scores, const uint32_t* h_data);       /* SumTree.add :23-42 */
int dqn_tree_update(dqn_learner* h, int64_t n, const int64_t* h_tree_idx, const float* h_scores,   /* update_score :45-52, in order */
                    int32_t store_losses);                                                         /* + replay.losses[mem]=p (replay_sampler_priority.py:46-50) */
/* SumTree.get_with_info for the B stratified probes of ReplaySamplerPriority.sample_memories
 * (replay_sampler_priority.py:30-35); u are the B uniform draws (random.random()). Outputs may be NULL. */
int dqn_tree_sample(dqn_learner* h, int32_t batch, const double* h_u, int64_t* h_tree_idx, int64_t* h_data_idx,
                    double* h_priorities, int64_t* h_mem_idx);
/* SumTree.get_with_info(s) (sum_tree.py:76-83) for n caller-computed probes s (rounded to f32 like the reference's
 * NumPy-2 arithmetic): same outputs as dqn_tree_sample */
int dqn_tree_retrieve(dqn_learner* h, int32_t n, const double* h_scores, int64_t* h_tree_idx, int64_t* h_data_idx,
                      double* h_priorities, int64_t* h_mem_idx);
/* read-only (does not touch the PER normaliser or the report statistics of the handle) */
int dqn_tree_stats(dqn_learner* h, float* mn, float* mx, float* avg, float* total);               /* :139-165 */
int dqn_tree_read(dqn_learner* h, float* h_tree /*2N-1*/, uint32_t* h_data /*N*/, int64_t* size, int64_t* write);
/* SumTree.size / .write / .total (sum_tree.py:18-21,139-141) without moving the tree: size and write are host-side
 * bookkeeping, total is the 4-byte root; any pointer may be NULL */
int dqn_tree_info(dqn_learner* h, int64_t* size, int64_t* write, float* total);
/* SumTree.tree[tree_idx] and SumTree.get_data(tree_idx) (sum_tree.py:62-73) for n nodes; h_data entries of inner
 * nodes are 0; either output may be NULL */
int dqn_tree_get(dqn_learner* h, int64_t n, const int64_t* h_tree_idx, float* h_scores, uint32_t* h_data);

/* ---- samplers (seam 3): fill the device-resident train batch --------- */
/* ReplaySamplerPriority.sample_memories (replay_sampler_priority.py:27-43) + the IS-weight maths of
 * DeepQAgent._sample_memories (deep_q_agent.py:500-524). refresh_stats!=0 recomputes avg/min/max and
 * last_max_weight first.  Returns once tree_idx / priorities / is_weights are on the host; the gather of the sampled
 * rows into the device batch is stream-ordered behind that (every later call on this handle sees it; code that reads
 * the batch through dqn_device_ptr on another stream calls dqn_sync first). */
int dqn_per_sample(dqn_learner* h, const double* h_u, int32_t refresh_stats, double per_b,
                   int64_t* h_tree_idx, double* h_priorities, double* h_is_weights);
/* ReplaySampler.sample_memories (replay_sampler.py:21-32) with the caller's randint draws */
int dqn_uniform_sample(dqn_learner* h, const int64_t* h_rows);
/* read back the device train batch (parity/debug only); any pointer may be NULL */
int dqn_batch_read(dqn_learner* h, uint8_t* h_states, uint8_t* h_actions, uint8_t* h_rewards,
                   uint8_t* h_next_states, uint8_t* h_continues, uint16_t* h_losses_f16);

/* ---- model (seam 2): host-buffer calls ------------------------------- */
/* DeepQModel.train (deep_q_model.py:75-103). h_is_weights NULL unless use_per. */
int dqn_train_batch(dqn_learner* h, int32_t n, const uint8_t* h_states, const uint8_t* h_actions,
                    const uint8_t* h_rewards, const uint8_t* h_continues, const uint8_t* h_next_states,
                    const float* h_is_weights, int64_t* step, float* h_losses, float* loss);
/* DeepQModel.predict (deep_q_model.py:118-131) */
int dqn_predict(dqn_learner* h, int32_t n, const uint8_t* h_states, int32_t use_target, float* h_q);
/* DeepQModel.get_losses (deep_q_model.py:144-157) */
int dqn_get_losses(dqn_learner* h, int32_t n, const uint8_t* h_states, const uint8_t* h_actions,
                   const uint8_t* h_rewards, const uint8_t* h_continues, const uint8_t* h_next_states,
                   float* h_losses);
/* debug: Q-values/targets of the last train step (q_online[B,A], y[B], max_q[B]); grads when kept */
int dqn_debug_read(dqn_learner* h, const char* what, float* h_out, int64_t count);

/* ---- fused step: DeepQAgent._train_run_train (deep_q_agent.py:326-349) ------------------------
 * sample -> gather -> target -> forward/backward -> optimizer -> priority write-back, all on device.
 * Parity runs hand in the reference's draws: h_u = B random.random() values (PER) or h_rows = B randint rows
 * (uniform sampler); both NULL -> device Philox. Outputs may be NULL (then there is no D2H and no sync). */
int dqn_train_step(dqn_learner* h, const double* h_u, const int64_t* h_rows, int64_t* step, float* h_losses, float* loss);
/* The same step, asynchronous: _train_run_train returns nothing (rl/deep_q_agent.py:326-349; the losses only feed the log line
 * of _train_report_stats :220-236), so the caller may queue step i+1 -- its B host draws are copied from h_u / h_rows before the
 * call returns -- while step i runs, and collect (step, losses[B], loss) afterwards.  At most two steps in flight; results come
 * back in launch order.  dqn_train_step_result blocks until the oldest unread step has finished. */
int dqn_train_step_async(dqn_learner* h, const double* h_u, const int64_t* h_rows);
int dqn_train_step_result(dqn_learner* h, int64_t* step, float* h_losses, float* loss);
/* n back-to-back fused steps (device RNG), target sync every copy_network_steps and PER stats
 * refresh every per_calculate_steps included; returns after enqueueing (call dqn_sync). */
int dqn_train_steps(dqn_learner* h, int64_t n);
/* data-parallel split step (new capability; the reference is single-device).  phase 0 = sample the
 * local replay shard, forward, backward, priority write-back of the sampled rows: the gradient of (grad_scale x the local mean loss) is left in the "grads" slab
 * (dqn_device_ptr) for an external all-reduce; phase 4 = the same on the batch already in the device buffers (after
 * dqn_uniform_sample / dqn_per_sample: equivalence tests); phase 1 = optimizer on the (reduced) slab + target sync cadence.  grad_scale = 1 / world_size with a SUM all-reduce gives the
 * gradient of the global-batch mean. */
int dqn_train_step_phase(dqn_learner* h, int32_t phase, double grad_scale);
/* make a caller's CUDA stream wait for a point inside the phase 0 just issued: "fc_grads" = the gradients of the two
 * dense hidden kernels (98.6 % of the slab) are complete -- they are produced first, so their all-reduce can overlap
 * the conv backward chain -- or "grads" = the whole slab is complete */
int dqn_phase_wait(dqn_learner* h, const char* what, void* cuda_stream);
/* overwrite the PER normaliser last_max_weight (rl/deep_q_agent.py:510) -- data-parallel ranks install the maximum
 * over all shards so that importance weights share one scale */
int dqn_set_per_normaliser(dqn_learner* h, float last_max_weight);
/* execution options (ablation / tests); the full table with defaults is in INTEGRATION.md section 4.  Every option keeps
 * the results within the parity bars.  "keep_grads" (default 0): dqn_train_batch leaves the gradients of the dense
 * hidden kernels readable through dqn_get_param(slot 3); with 0 they are consumed by the fused dense wgrad+optimizer
 * kernel and slot 3 holds stale values for those two tensors.  "repack": call after writing weights through
 * dqn_device_ptr. */
int dqn_set_option(dqn_learner* h, const char* name, int32_t value);
/* per-kernel launch counter since creation (bench.py reports gpu_launches from this) */
int dqn_launch_count(const dqn_learner* h, int64_t* n);
/* CUDA-event time per step segment (sample, gather, forward_online, forward_target, td_epilogue, backward,
 * optimizer, post_step), accumulated over dqn_train_steps calls while enabled (steps run un-graphed) */
int dqn_set_profile(dqn_learner* h, int32_t enabled);
int dqn_get_profile(dqn_learner* h, int32_t* n, const char** names, float* ms, int32_t cap, int64_t* steps);

/* GEMM engine self-test: C[M,N] = A*B on host matrices through the FFMA kernel (mode DQN_MATH_FP32) or the
 * tcgen05 kernel (DQN_MATH_TC3X / DQN_MATH_TC1X). A is [M][K] if a_kcontig else [K][M]; B is [N][K] if b_kcontig
 * else [K][N]. */
int dqn_selftest_gemm(dqn_learner* h, int32_t mode, int32_t M, int32_t N, int32_t K, int32_t a_kcontig,
                      int32_t b_kcontig, const float* h_A, const float* h_B, float* h_C);

/* GEMM engine micro-benchmark on device-resident operands: average ms per launch over `iters` launches */
int dqn_bench_gemm(dqn_learner* h, int32_t mode, int32_t M, int32_t N, int32_t K, int32_t a_kcontig, int32_t b_kcontig,
                   int32_t iters, float* ms_per_launch);

/* debug: in-kernel clock64 timeline of the k-loop of one tcgen05 GEMM launch (first CTA, first 64 k-blocks x 8 marks) */
int dqn_debug_timeline(dqn_learner* h, int32_t mode, int32_t M, int32_t N, int32_t K, int32_t a_kcontig, int32_t b_kcontig,
                       unsigned long long* h_out512);

/* ---- persistence (own container format; TF-bundle compatibility is SURVEY.md 8f-3) -- */
int dqn_save(dqn_learner* h, const char* prefix);     /* Model.save, rl/model/model.py:60-68 */
int dqn_restore(dqn_learner* h, const char* prefix);  /* Model.restore :71-83; returns 3 if absent */
/* replay persistence: what DeepQAgent._train_finish / _train_init_memory do with <prefix>_replay_memory.hdf5
 * (rl/deep_q_agent.py:410-432, :211-222) -- rows 0 .. len-1 in index order, six columns, no tree; on restore the rows are
 * appended from index 0 and the sum tree is rebuilt from the float16 losses column (ReplaySamplerPriority.add_losses,
 * rl/data/replay_sampler_priority.py:57-59).  Own container <prefix>_replay_memory.dqnb200 (either layout).
 * dqn_restore_replay returns 3 if the file is absent. */
int dqn_save_replay(dqn_learner* h, const char* prefix);
int dqn_restore_replay(dqn_learner* h, const char* prefix, int64_t* rows_loaded);

#ifdef __cplusplus
}
#endif
#endif /* DQN_B200_H */
