// *** SKETCH, NOT A WORKING OP ***  Key bodies below are elided ("... (elided) ..."), the gradient op the Python file refers
// to is not registered, and Compute() frees the plan the gradient would need.  It documents WHERE the C ABI plugs into a
// TF custom op; the working, tested binding in this image is ctypes (gnnkeras_b200/_lib.py).
//
// gnn_fixed_point_op.cc - TensorFlow custom op over the gnnfp C ABI (SOURCE ONLY in this repository:
// the build image has no TensorFlow headers, so this file is neither compiled nor tested here; it shows the
// exact binding a maintainer adds where `import tensorflow` works).
//
//   g++ -std=c++17 -shared -fPIC gnn_fixed_point_op.cc -o _gnn_fixed_point_op.so \
//       $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))') \
//       $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))') \
//       -I<repo>/include -L<repo>/gnnkeras_b200 -lgnnfp -lcudart
//
// Op:  (k, state, out) = GnnFixedPoint(nodes, arc_labels, state0, src, dst, node2graph, set_mask,
//                                      output_mask, state_weights..., out_weights...)
// The graph handle and loop plan are cached per (N, A, G) in the op kernel; training=1 keeps the workspace
// in a resource so that the registered gradient (GnnFixedPointGrad -> gnnfp_loop_backward) can reuse it.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/common_runtime/gpu/gpu_event_mgr.h"

#include "gnnfp.h"

using namespace tensorflow;

REGISTER_OP("GnnFixedPoint")
    .Input("nodes: float")
    .Input("arc_labels: float")
    .Input("state0: float")
    .Input("src: int32")
    .Input("dst: int32")
    .Input("node2graph: int32")
    .Input("set_mask: uint8")
    .Input("output_mask: uint8")
    .Input("state_weights: NS * float")
    .Input("out_weights: NO * float")
    .Attr("NS: int >= 2")
    .Attr("NO: int >= 2")
    .Attr("kind: int")
    .Attr("state_vect_dim: int")
    .Attr("max_iteration: int")
    .Attr("state_threshold: float")
    .Attr("training: bool")
    .Attr("aggregation_mode: int")
    .Attr("n_graphs: int")
    .Attr("state_widths: list(int)")
    .Attr("state_acts: list(int)")
    .Attr("state_bn: bool")
    .Attr("out_widths: list(int)")
    .Attr("out_acts: list(int)")
    .Attr("out_bn: bool")
    .Output("k: int32")
    .Output("state: float")
    .Output("out: float")
    .Output("workspace: uint8");

class GnnFixedPointOp : public OpKernel {
 public:
  explicit GnnFixedPointOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("kind", &kind_));
    OP_REQUIRES_OK(c, c->GetAttr("state_vect_dim", &S_));
    OP_REQUIRES_OK(c, c->GetAttr("max_iteration", &max_it_));
    OP_REQUIRES_OK(c, c->GetAttr("state_threshold", &thr_));
    OP_REQUIRES_OK(c, c->GetAttr("training", &training_));
    OP_REQUIRES_OK(c, c->GetAttr("aggregation_mode", &mode_));
    OP_REQUIRES_OK(c, c->GetAttr("n_graphs", &G_));
    // ... state_widths / acts / bn -> gnnfp_net_desc (elided: mirrors gnnkeras_b200/op.py Net.desc())
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& nodes = ctx->input(0);
    const Tensor& arcl = ctx->input(1);
    const int N = nodes.dim_size(0), A = arcl.dim_size(0);
    auto stream = ctx->eigen_gpu_device().stream();          // cudaStream_t of this op
    gnnfp_graph_desc gd{};
    gd.n_nodes = N; gd.n_arcs = A; gd.n_graphs = G_; gd.aggregation_mode = mode_; gd.mask_len = N;
    gd.src = ctx->input(3).flat<int32>().data(); gd.dst = ctx->input(4).flat<int32>().data();
    gd.node2graph = G_ ? ctx->input(5).flat<int32>().data() : nullptr;
    gd.set_mask = ctx->input(6).flat<uint8>().data(); gd.output_mask = ctx->input(7).flat<uint8>().data();
    gnnfp_graph* g = nullptr;
    OP_REQUIRES(ctx, gnnfp_graph_build(&g, &gd, stream) == 0, errors::InvalidArgument(gnnfp_last_error()));
    gnnfp_loop_cfg cfg{};
    cfg.kind = kind_; cfg.pool = -1; cfg.state_vect_dim = S_; cfg.max_iteration = max_it_; cfg.state_threshold = thr_;
    cfg.training = training_; cfg.nodes_width = nodes.dim_size(1); cfg.arc_label_width = arcl.dim_size(1);
    gnnfp_loop* L = nullptr;
    OP_REQUIRES(ctx, gnnfp_loop_create(&L, g, &cfg, &state_desc_, &out_desc_) == 0, errors::InvalidArgument(gnnfp_last_error()));
    Tensor *k, *state, *out, *ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({}), &k));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({N, gnnfp_loop_state_dim(L)}), &state));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({gnnfp_loop_out_rows(L), out_desc_.widths[out_desc_.n_layers - 1]}), &out));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({(int64_t)gnnfp_loop_workspace_bytes(L) + 256}), &ws));
    gnnfp_loop_io io{};
    io.nodes = nodes.flat<float>().data(); io.ld_nodes = nodes.dim_size(1);
    io.arc_labels = arcl.flat<float>().data(); io.ld_arcs = arcl.dim_size(1);
    io.state0 = S_ ? ctx->input(2).flat<float>().data() : nullptr;
    io.state_out = state->flat<float>().data(); io.out = out->flat<float>().data(); io.k_out = k->flat<int32>().data();
    gnnfp_net_params sp{}, op{};   // filled from the state_weights / out_weights input lists in Keras variable order
    // ... (elided) ...
    void* wsp = (void*)(((uintptr_t)ws->flat<uint8>().data() + 255) / 256 * 256);
    OP_REQUIRES(ctx, gnnfp_loop_forward(L, &sp, &op, &io, wsp, gnnfp_loop_workspace_bytes(L), stream) == 0,
                errors::Internal(gnnfp_last_error()));
    // g and L are kept in a per-kernel cache keyed by (N, A, G) in the full shim; freed here for brevity
    gnnfp_loop_free(L);
    gnnfp_graph_free(g);
  }
 private:
  int kind_, S_, max_it_, mode_, G_;
  float thr_;
  bool training_;
  gnnfp_net_desc state_desc_{}, out_desc_{};
};
REGISTER_KERNEL_BUILDER(Name("GnnFixedPoint").Device(DEVICE_GPU), GnnFixedPointOp);
