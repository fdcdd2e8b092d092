/*
 * gnnfp.h - C ABI of libgnnfp.so: the B200-native (sm_100a) fixed-point GNN loop.
 *
 * This is the drop-in boundary for the ONE hot path of NickDrake117/GNNkeras: the
 * state-transition fixed-point loop `Loop` / `condition` / `convergence` / `apply_filters`
 * of GNNnodeBased / GNNarcBased / GNNgraphBased (reference GNN/Models/GNN.py:196-274,
 * 317-330, 341-346), of the composite models (GNN/Models/CompositeGNN.py:194-272, 315-343),
 * the layer chaining of LGNN / CompositeLGNN (GNN/Models/LGNN.py:175-249,
 * GNN/Models/CompositeLGNN.py:25-57) and the backward that `tf.GradientTape` performs in
 * `train_step` (GNN/Models/GNN.py:284-295, LGNN.py:259-272, CompositeGNN.py:282-293).
 *
 * The reference has no FFI (it is pure Python on TensorFlow eager); the seam this ABI
 * replaces is the Python method `Loop(nodes, arcs, dim_node_label, [type_mask,] set_mask,
 * output_mask, [composite_adjacencies,] adjacency, arcnode, nodegraph, training)
 * -> (k, state, out)`.  A TensorFlow custom op `GnnFixedPoint` (+ registered gradient) or the
 * in-image ctypes/torch.autograd binding calls exactly these entry points.
 *
 * Conventions
 *   - C linkage, POD arguments only.  Every function returns 0 on success or a negative
 *     GNNFP_E_* code; `gnnfp_last_error()` returns a thread-local message.  Nothing throws or
 *     aborts across the boundary.
 *   - All data pointers are DEVICE pointers owned by the caller (row-major float32, int32
 *     indices, uint8 masks) unless a parameter says "host".  The library allocates device memory
 *     only inside `gnnfp_graph_build` (the graph handle) - everything else lives in the
 *     caller-provided workspace.
 *   - All work is enqueued on the caller's stream (`void* stream` is a cudaStream_t); no hidden
 *     synchronisation: the iteration count k stays on the device (the reference's per-iteration
 *     host `bool()`, GNN.py:265, is gone).
 *   - No CPU fallback, no alternative backends: an unsupported configuration is an error.
 */
#ifndef GNNFP_H_
#define GNNFP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNNFP_ABI_VERSION 2
#define GNNFP_MAX_LAYERS 8
#define GNNFP_MAX_TYPES 8

/* error codes */
#define GNNFP_OK 0
#define GNNFP_E_INVALID (-1)     /* bad argument / inconsistent shapes            */
#define GNNFP_E_UNSUPPORTED (-2) /* configuration outside what the kernels cover  */
#define GNNFP_E_CUDA (-3)        /* CUDA runtime error (message has the detail)   */
#define GNNFP_E_WORKSPACE (-4)   /* workspace too small                           */

/* activations: Keras names used by the reference's MLP() factory (GNN/Models/MLP.py:12-78) */
enum { GNNFP_ACT_LINEAR = 0, GNNFP_ACT_TANH = 1, GNNFP_ACT_SIGMOID = 2, GNNFP_ACT_RELU = 3,
       GNNFP_ACT_SELU = 4, GNNFP_ACT_SOFTMAX = 5 };

/* aggregation_mode of GraphObject.buildArcNode (GNN/graph_class.py:91-124) and
 * CompositeGraphObject.buildArcNode (GNN/composite_graph_class.py:73-103).  EXPLICIT = take the
 * per-arc values the caller passes (a user-supplied ArcNode/Adjacency, graph_class.py:68). */
enum { GNNFP_AGG_SUM = 0, GNNFP_AGG_NORMALIZED = 1, GNNFP_AGG_AVERAGE = 2,
       GNNFP_AGG_COMPOSITE_AVERAGE = 3, GNNFP_AGG_EXPLICIT = 4 };

/* model kind: GNNnodeBased / GNNarcBased / GNNgraphBased (GNN.py:8, 311, 336) */
enum { GNNFP_KIND_NODE = 0, GNNFP_KIND_ARC = 1, GNNFP_KIND_GRAPH = 2 };

const char* gnnfp_last_error(void);
int gnnfp_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * Graph handle: the integer structures the loop consumes, built on the device.
 * Replaces GraphObject.buildArcNode/buildAdjacency/buildNodeGraph (graph_class.py:82-138),
 * CompositeGraphObject.buildCompositeAdjacency (composite_graph_class.py:57-70) and the
 * tensorisation GraphTensor.COO2SparseTensor (graph_class.py:551-560).
 * ------------------------------------------------------------------------------------------ */
typedef struct gnnfp_graph gnnfp_graph;

typedef struct gnnfp_graph_desc {
  int32_t n_nodes;           /* N                                                             */
  int32_t n_arcs;            /* A (after the reference's np.unique, graph_class.py:47)         */
  int32_t n_graphs;          /* G; 0 when there is no NodeGraph (node / arc focus)             */
  int32_t n_types;           /* 0 = homogeneous; >0 = composite node types                     */
  int32_t aggregation_mode;  /* GNNFP_AGG_*                                                   */
  int32_t mask_len;          /* length of set_mask/output_mask: N (node/graph focus) or A (arc)*/
  const int32_t* src;        /* [A] arcs[:,0]                                                  */
  const int32_t* dst;        /* [A] arcs[:,1]                                                  */
  const float* arc_values;   /* [A] explicit ArcNode.data / Adjacency.values, or NULL          */
  const uint8_t* type_mask;  /* [n_types, N] as it reaches the model (composite_graph_class.py:263) or NULL */
  const int32_t* node2graph; /* [N] column of each node's NodeGraph entry, or NULL             */
  const float* nodegraph_values; /* [N] NodeGraph.data (1/n_g, graph_class.py:136); NULL = compute 1/n_g */
  const uint8_t* set_mask;   /* [mask_len] or NULL (= all true)                                */
  const uint8_t* output_mask;/* [mask_len] or NULL (= all true)                                */
  int32_t flags;             /* GNNFP_GRAPH_*                                                  */
} gnnfp_graph_desc;

/* flags.  DEFER_CHECK: gnnfp_graph_build does not wait for the device-side validation of the ids (node / graph id out of
 * range, nodes of a graph not contiguous): the verdict stays in the handle until gnnfp_graph_check.  With no masks and no
 * node types the build then enqueues its work and returns WITHOUT any host synchronisation (the per-step input path of a
 * training loop); the structures of an invalid graph are never dereferenced out of bounds, its results are undefined. */
#define GNNFP_GRAPH_DEFER_CHECK 1

int gnnfp_graph_build(gnnfp_graph** out, const gnnfp_graph_desc* desc, void* stream);
/* waits for the build's stream work and reports the deferred validation (GNNFP_OK, GNNFP_E_INVALID, GNNFP_E_UNSUPPORTED) */
int gnnfp_graph_check(const gnnfp_graph* g);
/* releases the handle's arrays in the order of the stream it was built on (cudaFreeAsync): work that uses the handle on
 * OTHER streams must have been ordered before that stream by the caller (event / stream wait) */
void gnnfp_graph_free(gnnfp_graph* g);

typedef struct gnnfp_graph_info {
  int32_t n_nodes, n_arcs, n_graphs, n_types;
  int32_t n_masked;                         /* M = popcount(set_mask & output_mask)            */
  int32_t type_count[GNNFP_MAX_TYPES];      /* rows of each node type                          */
  int32_t types_disjoint_cover;             /* 1 if every node is in exactly one type          */
  size_t device_bytes;                      /* device memory owned by the handle               */
} gnnfp_graph_info;
int gnnfp_graph_get_info(const gnnfp_graph* g, gnnfp_graph_info* info);

/* Copy one of the built integer / weight arrays to HOST memory (parity tests compare them with
 * `==` against the oracle).  `which`: */
enum { GNNFP_X_DST_ROWPTR = 0,  /* int32 [N+1]  destination-grouped CSR row pointers            */
       GNNFP_X_DST_SRC = 1,     /* int32 [A]    source node of each entry (arc order in a row)  */
       GNNFP_X_DST_ARC = 2,     /* int32 [A]    arc id of each entry                            */
       GNNFP_X_SRC_ROWPTR = 3,  /* int32 [N+1]  source-grouped CSR (backward)                   */
       GNNFP_X_SRC_DST = 4,     /* int32 [A]    destination node of each entry                  */
       GNNFP_X_SRC_ARC = 5,     /* int32 [A]                                                    */
       GNNFP_X_ARC_VALUE = 6,   /* float [A]    ArcNode.data == Adjacency.data in arc order     */
       GNNFP_X_MASK_INDEX = 7,  /* int32 [M]    rows with set_mask & output_mask, ascending     */
       GNNFP_X_GRAPH_PTR = 8,   /* int32 [G+1]  first node of each graph                        */
       GNNFP_X_NODEGRAPH_VALUE = 9, /* float [N]                                                */
       GNNFP_X_TYPE_ROWS = 10   /* int32 [sum type_count] rows of type 0, then type 1, ...      */ };
int gnnfp_graph_export(const gnnfp_graph* g, int which, void* host_dst, size_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Nets: what the reference's MLP() factory can build (MLP.py:12-78): an optional leading
 * BatchNormalization followed by Dense(activation) layers.  Dropout layers are identity at
 * inference and unsupported in training (GNNFP_E_UNSUPPORTED).
 * Parameters are in Keras variable order: BN gamma, beta, (moving_mean,
 * moving_variance), then per Dense layer kernel [in,out] row-major and bias [out].
 * ------------------------------------------------------------------------------------------ */
typedef struct gnnfp_net_desc {
  int32_t n_layers;
  int32_t in_dim;
  int32_t widths[GNNFP_MAX_LAYERS];
  int32_t acts[GNNFP_MAX_LAYERS];
  int32_t has_bn;
  float bn_eps;       /* Keras default 1e-3  */
  float bn_momentum;  /* Keras default 0.99  */
} gnnfp_net_desc;

typedef struct gnnfp_net_params {
  float* bn_gamma;        /* [in_dim] or NULL                                                  */
  float* bn_beta;         /* [in_dim] or NULL                                                  */
  float* bn_moving_mean;  /* [in_dim]; UPDATED IN PLACE once per executed iteration in training */
  float* bn_moving_var;   /* [in_dim]; (Keras BatchNormalization semantics)  */
  float* W[GNNFP_MAX_LAYERS]; /* [in_l, out_l] row-major                                       */
  float* b[GNNFP_MAX_LAYERS]; /* [out_l]                                                       */
} gnnfp_net_params;       /* the same struct carries gradients (moving_* unused)               */

/* ------------------------------------------------------------------------------------------
 * Loop plan: one GNN (one `Loop` call).  LGNN = one plan per layer + gnnfp_update_graph_*.
 * ------------------------------------------------------------------------------------------ */
typedef struct gnnfp_loop gnnfp_loop;

typedef struct gnnfp_loop_cfg {
  int32_t kind;             /* GNNFP_KIND_*  (graph kind pools with NodeGraph, GNN.py:341-346)  */
  int32_t pool;             /* -1 = by kind; 0/1 override (LGNN runs inner layers un-pooled, LGNN.py:225) */
  int32_t state_vect_dim;   /* S >= 0 (GNN.py:26); 0 => state0 = nodes                         */
  int32_t max_iteration;    /* GNN.py:27                                                        */
  float state_threshold;    /* GNN.py:28                                                        */
  int32_t training;         /* 0: inference (BN moving stats; nothing saved)  1: training       */
  int32_t n_types;          /* 0 homogeneous, else number of state nets (CompositeGNN.py:17)    */
  int32_t dim_node_label[GNNFP_MAX_TYPES]; /* composite: d_t used as nodes[:, :d_t] (clamped)   */
  int32_t nodes_width;      /* columns of `nodes`                                               */
  int32_t arc_label_width;  /* AL = columns of arcs[:, 2:]                                      */
  int32_t want_input_grads; /* bit0 d_nodes, bit1 d_arc_labels, bit2 d_state0 (LGNN chaining)   */
  int32_t n_active_rows;    /* 0 = all nodes.  >0: net_state runs on rows [0, n_active_rows) only; the
                               remaining rows are halo copies of remote nodes that the multi-GPU driver
                               refreshes between iterations (edge-cut partition; homogeneous, no BN;
                               inference, and training with a single-Dense-layer net_state through
                               gnnfp_loop_backward_step).                                           */
} gnnfp_loop_cfg;

int gnnfp_loop_create(gnnfp_loop** out, const gnnfp_graph* g, const gnnfp_loop_cfg* cfg,
                      const gnnfp_net_desc* state_nets /* [max(1,n_types)] */,
                      const gnnfp_net_desc* out_net);
void gnnfp_loop_free(gnnfp_loop* L);
size_t gnnfp_loop_workspace_bytes(const gnnfp_loop* L);
int gnnfp_loop_out_rows(const gnnfp_loop* L);  /* rows of `out`: G (pooled) or M              */
int gnnfp_loop_state_dim(const gnnfp_loop* L); /* D                                            */

typedef struct gnnfp_loop_io {
  const float* nodes;      /* [N, nodes_width], leading dimension ld_nodes                     */
  int32_t ld_nodes;
  const float* arc_labels; /* [A, AL], leading dimension ld_arcs (pass arcs+2 with ld 2+AL)    */
  int32_t ld_arcs;
  const float* state0;     /* [N, S] when S > 0 (explicit: GNN.py:257 is unseeded), else NULL  */
  float* state_out;        /* [N, D]  converged state                                          */
  float* out;              /* [out_rows, T]                                                    */
  float* out_nodes;        /* optional [M, T]: un-pooled net_output rows (LGNN update_graph)   */
  int32_t* k_out;          /* device int32: iterations executed (GNN.py:260 keeps it a float)  */
} gnnfp_loop_io;

/* (k, state, out) = Loop(...).  Replaces GNN.py:245-274 / CompositeGNN.py:242-272. */
int gnnfp_loop_forward(gnnfp_loop* L, const gnnfp_net_params* state_params /* [max(1,n_types)] */,
                       const gnnfp_net_params* out_params, const gnnfp_loop_io* io,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Stepping form of gnnfp_loop_forward (same arguments): begin = prologue + first condition, iter = iteration t
 * (1-based, gated on the device flag of iteration t-1; writes flag t), end = converged state, net_output,
 * pooling.  gnnfp_loop_forward == begin; iter(1..max_iteration); end.  Between iter(t) and iter(t+1) a
 * partitioned driver (a) refreshes the halo rows of state slot t and (b) max-reduces flag t over the ranks;
 * gnnfp_loop_ws_offsets tells where those live inside the caller's workspace: int32 flags[max_iteration+1]
 * at byte offset flags_off; state slot of iteration t (t >= 1) at slots_off + slot_index * slot_stride_floats * 4
 * with slot_index = t-1 (training) or t&1 (inference). */
int gnnfp_loop_forward_begin(gnnfp_loop* L, const gnnfp_net_params* state_params, const gnnfp_net_params* out_params,
                             const gnnfp_loop_io* io, void* workspace, size_t workspace_bytes, void* stream);
int gnnfp_loop_forward_iter(gnnfp_loop* L, int32_t t, const gnnfp_net_params* state_params,
                            const gnnfp_net_params* out_params, const gnnfp_loop_io* io, void* workspace,
                            size_t workspace_bytes, void* stream);
int gnnfp_loop_forward_end(gnnfp_loop* L, const gnnfp_net_params* state_params, const gnnfp_net_params* out_params,
                           const gnnfp_loop_io* io, void* workspace, size_t workspace_bytes, void* stream);
int gnnfp_loop_ws_offsets(const gnnfp_loop* L, size_t* flags_off, size_t* slots_off, size_t* slot_stride_floats,
                          int32_t* slot_count);
/* row pitch (floats) of a state slot - the state occupies columns [0, D) of a slot row -, the slot index that holds the
 * state of iteration 1 in a training plan (state t sits in slot state1_slot + t - 1; inference plans ping-pong: slot t & 1),
 * and the row pitch of the Adj . dAgg gather buffer (gnnfp_loop_bwd_offsets) */
int gnnfp_loop_ws_layout(const gnnfp_loop* plan, int32_t* ld_state, int32_t* state1_slot, int32_t* ld_grad);

typedef struct gnnfp_loop_grads {
  const float* d_out;        /* [out_rows, T] dL/d out (may be NULL = zeros)                    */
  const float* d_out_nodes;  /* optional [M, T] dL/d out_nodes (LGNN: from update_graph)        */
  const float* d_state;      /* optional [N, D] dL/d state_out (LGNN: from update_graph)        */
  float* d_nodes;            /* [N, nodes_width] written if want_input_grads&1                  */
  float* d_arc_labels;       /* [A, AL]          written if want_input_grads&2                  */
  float* d_state0;           /* [N, S]           written if want_input_grads&4 and S>0          */
  int32_t average_st_grads;  /* divide state-net grads by k (GNN.py:295)                        */
} gnnfp_loop_grads;

/* Full BPTT over the k executed iterations of the matching forward (same workspace, same io).
 * Gradients are WRITTEN (not accumulated) into d_state_params / d_out_params. */
int gnnfp_loop_backward(gnnfp_loop* L, const gnnfp_net_params* state_params,
                        const gnnfp_net_params* out_params, const gnnfp_loop_io* io,
                        const gnnfp_loop_grads* grads, gnnfp_net_params* d_state_params,
                        gnnfp_net_params* d_out_params, void* workspace, size_t workspace_bytes,
                        void* stream);

/* Stepping form of gnnfp_loop_backward (same arguments) for partitioned graphs (plans created with n_active_rows <
 * n_nodes; replaces the single tf.GradientTape call of GNN.py:284-295 on each rank):
 *   phase 1 = begin  (zeroing, net_output backward),
 *   phase 8 = gather (t < max_iteration): this rank's share of Adj . dAgg_{t+1} for ALL local rows (owned + halo) into the
 *             [n_nodes, D] buffer at byte offset gnnfp_loop_bwd_offsets().gather_off of the workspace; the driver then sends
 *             the halo rows to their owners, which ADD them to their rows (the reverse of the forward halo exchange),
 *   phase 2 = iteration t (newest first: t = max_iteration .. 1) on the owned rows, consuming that buffer,
 *   phase 4 = end    (deterministic reduction of the parameter gradients of this rank; the caller sums them over ranks).
 * gnnfp_loop_backward == phase 1; phase 2 for t = max_iteration .. 1; phase 4 on unpartitioned plans. */
int gnnfp_loop_backward_step(gnnfp_loop* L, int32_t phase, int32_t t, const gnnfp_net_params* state_params,
                             const gnnfp_net_params* out_params, const gnnfp_loop_io* io, const gnnfp_loop_grads* grads,
                             gnnfp_net_params* d_state_params, gnnfp_net_params* d_out_params, void* workspace,
                             size_t workspace_bytes, void* stream);
int gnnfp_loop_bwd_offsets(const gnnfp_loop* L, size_t* gather_off);

/* ------------------------------------------------------------------------------------------
 * LGNN.update_graph (LGNN.py:175-214): nodes' = [state? | scatter(out_nodes by mask)? | nodes0],
 * arc focus: arc_labels' = [scatter(out)? | arc_labels0].  Forward writes the new matrix;
 * backward splits an upstream gradient back into d_state / d_out_nodes / d_base.
 * ------------------------------------------------------------------------------------------ */
int gnnfp_update_graph_forward(const gnnfp_graph* g, int32_t n_rows, const float* state, int32_t state_w,
                               const float* out_rows, int32_t out_w, const float* base, int32_t base_w,
                               int32_t ld_base, float* dst, void* stream);
int gnnfp_update_graph_backward(const gnnfp_graph* g, int32_t n_rows, const float* d_dst, float* d_state,
                                int32_t state_w, float* d_out_rows, int32_t out_w, float* d_base,
                                int32_t base_w, int32_t accumulate_base, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused train-step tail ("next" row f2): Keras categorical_crossentropy on probabilities with
 * sample weights (SUM_OVER_BATCH_SIZE) and its gradient, and an Adam update over a flat buffer.
 * ------------------------------------------------------------------------------------------ */
int gnnfp_cce_loss(const float* y_true, const float* y_pred, const float* sample_weight, int32_t rows,
                   int32_t cols, float scale, float* loss_out /* device scalar, accumulated */,
                   float* d_pred /* [rows, cols] or NULL */, void* stream);
int gnnfp_adam_step(float* params, const float* grads, float* m, float* v, size_t n, float lr, float beta1,
                    float beta2, float eps, int32_t step, float grad_scale, void* stream);
/* The same update with the step count in device memory (t = *step_dev + 1): no per-step value is baked into a kernel
 * argument, so a whole train step (forward, loss, BPTT, update) can be captured into ONE CUDA graph and replayed
 * (optimizer.apply_gradients, GNN.py:297).  gnnfp_adam_advance increments the counter after the step's last span. */
int gnnfp_adam_step_dev(float* params, const float* grads, float* m, float* v, size_t n, float lr, float beta1,
                        float beta2, float eps, const int32_t* step_dev, float grad_scale, void* stream);
int gnnfp_adam_advance(int32_t* step_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side batcher ("next" row f1): replaces GraphObject.merge (graph_class.py:385-413) + the per-batch host work of
 * MultiGraphSequencer.build_batches / on_epoch_end (GraphSequencers.py:42-46, 123-127).  The dataset is resident on the
 * device as flat arrays (members back to back; arc id columns LOCAL to their member, rows sorted and unique as
 * GraphObject.__init__ leaves them, graph_class.py:47); a batch is a list of member ids.  All pointers are device
 * pointers owned by the caller; sizes of the outputs are the sums of the selected members' sizes (known on the host).
 * ------------------------------------------------------------------------------------------------ */
typedef struct gnnfp_store_desc {
  const float* nodes; int32_t nodes_width;            /* [sum N, nodes_width]                                  */
  const float* arcs; int32_t arcs_width;              /* [sum A, 2 + AL], columns 0-1 = local src, dst ids      */
  const float* targets; int32_t targets_width;        /* [sum T, targets_width]                                */
  const float* sample_weight;                         /* [sum T]                                               */
  const uint8_t* set_mask; const uint8_t* output_mask;/* [sum Mk] or NULL                                      */
  const int32_t* node2graph; const float* nodegraph_values;   /* [sum N] or NULL (no NodeGraph)                */
  const uint8_t* type_mask; int32_t n_types;          /* [sum N, n_types] or NULL                              */
  const int64_t* node_ptr; const int64_t* arc_ptr; const int64_t* tgt_ptr; const int64_t* mask_ptr;   /* [n + 1] prefix sums */
  const int32_t* n_sub;                               /* [n] NodeGraph columns of each member, or NULL         */
} gnnfp_store_desc;
typedef struct gnnfp_batch_out {
  float* nodes; float* arcs; int32_t* src; int32_t* dst; float* targets; float* sample_weight;
  uint8_t* set_mask; uint8_t* output_mask;            /* or NULL                                               */
  int32_t* node2graph; float* nodegraph_values;       /* or NULL                                               */
  uint8_t* type_mask;                                 /* [n_types, N_batch] (as it reaches the model) or NULL  */
} gnnfp_batch_out;
/* offsets_scratch: 5 * (n_ids + 1) int64 (device); after the call it holds the exclusive prefix sums of the selected
 * members' node / arc / target / mask / sub-graph counts.  Two kernel launches on `stream`, no synchronisation. */
int gnnfp_batch_assemble(const gnnfp_store_desc* store, const int64_t* ids_dev, int32_t n_ids, int64_t n_nodes_batch,
                         int64_t* offsets_scratch, const gnnfp_batch_out* out, void* stream);

/* Counters for bench.py's `gpu_launches` (kernels this library launched since the last reset). */
long long gnnfp_launch_count(int reset);
/* measurement aid (bench.py): one launch of an FMA-only kernel (8 chains per thread, blocks_per_sm x SMs blocks of 256
 * threads); *flops_out = its FLOPs.  Timed by the caller: the sustained FP32-pipe peak the tile kernels are held against. */
int gnnfp_debug_fma_peak(float* sink, int32_t iters, int32_t blocks_per_sm, double* flops_out, void* stream);

/* Optional per-category kernel timing (CUDA events recorded on the launch stream around every tile
 * kernel; off by default).  Categories: 0 other, 1 state-net forward iteration, 2 state-net backward
 * iteration, 3 tile pass (aggregates / BN statistics), 4 net_output forward, 5 net_output backward,
 * 6 BN backward fix-up, 7 dz (activation derivative x gathered gradient), 8 state-net backward dX GEMMs (category 2 is
 * then the dW GEMM alone), 9 Adj^T.state aggregation (+ BN statistics).  collect() synchronises, sums milliseconds and launch counts per category and
 * clears the records. */
int gnnfp_profile_enable(int on);
int gnnfp_profile_collect(double* ms_by_cat, long long* count_by_cat, int ncat);

#ifdef __cplusplus
}
#endif
#endif /* GNNFP_H_ */
