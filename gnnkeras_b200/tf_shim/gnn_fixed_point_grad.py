"""*** SKETCH, NOT WORKING CODE: the op it calls (gnn_fixed_point_grad) is registered nowhere - see the header of
gnn_fixed_point_op.cc.  The tested binding of this repository is ctypes (gnnkeras_b200/_lib.py). ***

Registered gradient of the GnnFixedPoint custom op (SOURCE ONLY here - TensorFlow is not installable in the
build image).  Where TF exists, `GNNnodeBased.Loop` becomes:

    k, state, out, ws = _mod.gnn_fixed_point(nodes, arcs[:, 2:], state0, src, dst, node2graph, set_mask,
                                             output_mask, net_state.trainable_variables, net_output.trainable_variables,
                                             kind=..., state_vect_dim=..., max_iteration=..., state_threshold=...,
                                             training=training, ...)

and `tf.GradientTape` (GNN.py:284-294) reaches `gnnfp_loop_backward` through this function.
"""
import tensorflow as tf

_mod = tf.load_op_library("_gnn_fixed_point_op.so")


@tf.RegisterGradient("GnnFixedPoint")
def _gnn_fixed_point_grad(op, dk, dstate, dout, dws):
    grads = _mod.gnn_fixed_point_grad(*op.inputs, op.outputs[1], op.outputs[2], op.outputs[3], dstate, dout,
                                      **{a: op.get_attr(a) for a in ("kind", "state_vect_dim", "max_iteration",
                                                                     "state_threshold", "aggregation_mode", "n_graphs")})
    # inputs: nodes, arc_labels, state0, src, dst, node2graph, set_mask, output_mask, state_weights..., out_weights...
    d_nodes, d_arcs, d_state0 = grads[0], grads[1], grads[2]
    return [d_nodes, d_arcs, d_state0, None, None, None, None, None] + list(grads[3:])
