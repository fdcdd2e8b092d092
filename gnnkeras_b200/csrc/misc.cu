#include "loop.h"
extern "C" int gnnfp_loop_backward(gnnfp_loop*, const gnnfp_net_params*, const gnnfp_net_params*, const gnnfp_loop_io*,
                        const gnnfp_loop_grads*, gnnfp_net_params*, gnnfp_net_params*, void*, size_t, void*) { GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "backward not built yet"); }
extern "C" int gnnfp_update_graph_forward(const gnnfp_graph*, int32_t, const float*, int32_t, const float*, int32_t, const float*, int32_t, int32_t, float*, void*) { GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "nyi"); }
extern "C" int gnnfp_update_graph_backward(const gnnfp_graph*, int32_t, const float*, float*, int32_t, float*, int32_t, float*, int32_t, int32_t, void*) { GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "nyi"); }
extern "C" int gnnfp_cce_loss(const float*, const float*, const float*, int32_t, int32_t, float, float*, float*, void*) { GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "nyi"); }
extern "C" int gnnfp_adam_step(float*, const float*, float*, float*, size_t, float, float, float, float, int32_t, float, void*) { GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "nyi"); }
