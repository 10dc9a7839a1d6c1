// Handle shared by the two GNN forward kernels (lm_gnn.cu: fp32 CUDA cores; lm_gnn_tc.cu: bf16 tcgen05).
#pragma once
#include <stdint.h>

struct LmGnn {
  int device;
  int sm_count;
  float* packed;       // fp32 blob of lm_gnn.cu
  uint8_t* tc_blob;    // bf16 operand blob of lm_gnn_tc.cu (UMMA canonical layout) + fp32 input-layer / bias / head parameters
  float* train_blob;   // fp32 weight image of lm_gnn_train.cu (forward + transposed packed forms), allocated on first lm_gnn_pack
  bool ready;
  bool tc_fresh_pack;  // weights re-packed since the last tensor-core forward: that launch must not start early (PDL)
  long long* timeline; // diagnostic only (lm_gnn_set_timeline)
};

// lm_gnn_tc.cu
int lm_gnn_tc_init(LmGnn* g);                                                        // blob allocation + kernel attributes
void lm_gnn_tc_release(LmGnn* g);
int lm_gnn_tc_pack(LmGnn* g, const float* const* w, void* stream);                   // same 20 tensors as lm_gnn_pack

// lm_gnn_train.cu
int lm_gnn_train_pack(LmGnn* g, const float* const* w, void* stream);               // same 20 tensors as lm_gnn_pack
void lm_gnn_train_release(LmGnn* g);
