// Fused attention launch description (see kernel_attn.cu).
#pragma once
#include "common.cuh"

namespace sgdm {

struct AttnDesc {
  // strides in ELEMENTS; token row r of sample n is at base + (n*T + r) * row_stride + head * head_stride
  const op_t* q = nullptr; long q_row_stride = 0; int q_head_stride = 0;
  const op_t* k = nullptr; long k_row_stride = 0; int k_head_stride = 0;  // head_stride 0 = multi-query
  const op_t* v = nullptr; long v_row_stride = 0; int v_head_stride = 0;
  // extra keys/values shared by all heads, placed before the self keys: [B][n_extra][D]
  const op_t* k_extra = nullptr; const op_t* v_extra = nullptr; int n_extra = 0;
  op_t* out = nullptr; long o_row_stride = 0;  // out[(n*T + r) * o_row_stride + head*D + d]
  int B = 0, T = 0, heads = 0, D = 0;
  float scale = 1.f;  // logits = scale * <q, k>
  int use_tc = -1;    // tcgen05 kernel (kernel_attn_tc.cu): -1 = whenever applicable, 0 = never (mma.sync kernel)
};
int attn_launch(const AttnDesc& a, cudaStream_t s);
// tcgen05 / TMEM self-attention for T = 256, D = 64 with q, k, v as column blocks of one packed matrix
bool attn_tc_applicable(const AttnDesc& a);
int attn_tc_launch(const AttnDesc& a, cudaStream_t s);
// tcgen05 / TMEM Attention_LR: multi-query (one shared k / v head) over T = 256 self keys + <= 32 extra keys, D = 64
bool attn_lr_tc_applicable(const AttnDesc& a);
int attn_lr_tc_launch(const AttnDesc& a, cudaStream_t s);

}  // namespace sgdm
