// Handle shared by the MLP kernels (lm_mlp.cu: forward; lm_mlp_train.cu: backward / weight gradients of the PPO update).
#pragma once
#include <stdint.h>

struct LmMlp {
  int device;
  int sm_count;
  uint8_t* blob;         // forward operands (bf16, UMMA canonical K-major) + fp32 biases
  uint8_t* blob_t;       // transposed weights for the backward activation kernel (allocated on first lm_mlp_pack)
  bool packed;
  bool fresh_pack;       // weights were re-packed since the last forward: that forward must not start early
  long long* timeline;   // diagnostic only (lm_mlp_set_timeline)
};

// lm_mlp_train.cu
int lm_mlp_train_pack(LmMlp* m, const float* w2, const float* w3, const float* w_mu, const float* w_value, void* stream);
void lm_mlp_train_release(LmMlp* m);
