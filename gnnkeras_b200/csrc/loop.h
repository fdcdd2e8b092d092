// loop.h - the opaque loop plan of gnnfp.h and its workspace layout
#pragma once
#include "graph.h"

struct WsLayout {
  size_t ctrl = 0, ctrl_bytes = 0;          // int flags[max_iter+1], k, pad | double stats ... (zeroed each forward)
  size_t flags = 0;                         // int [max_iter + 2]
  size_t stS = 0, stA = 0, stX = 0, stO = 0;  // double blocks
  size_t Xs = 0;                            // float [N, LsM]
  size_t slots = 0;                         // float state slots
  size_t out_nodes = 0;                     // float [M, T]
  // backward
  size_t dSfin = 0, dOwn = 0, dAgg = 0, dXs = 0, dOutN = 0;
  size_t part_state = 0, part_out = 0, bn_part = 0, bn_const = 0;
  size_t total = 0;
};

struct gnnfp_loop {
  const gnnfp_graph* g = nullptr;
  gnnfp_loop_cfg cfg{};
  gnnfp_net_desc snet[GNNFP_MAX_TYPES]{};
  gnnfp_net_desc onet{};
  int nt = 1;            // state nets
  int composite = 0;
  int N = 0, A = 0, M = 0, D = 0, S = 0, NLw = 0, AL = 0, T = 0;
  int dt[GNNFP_MAX_TYPES]{};   // clamped d_t (composite)
  int sum_dt = 0;
  int LsM = 0;           // materialised static block width
  int out_in = 0;        // net_output input width
  int out_rows = 0, pool = 0, n_out_rows_in = 0;   // n_out_rows_in: rows fed to net_output (M)
  int slot_count = 0;
  int bn_train_state = 0, bn_train_out = 0;
  int nparam_s[GNNFP_MAX_TYPES]{}, nparam_o = 0;
  int grid_cap = 0;      // upper bound of any tile kernel grid (partials are sized by it)
  WsLayout ws;
};
