// encoder_seq.cu -- host-side sequencer of the packed cross-encoder (reference: models/transformer/transformers.py:18-259,
// TransformerCrossEncoder of pre-norm TransformerCrossEncoderLayer with positional values): the 11 launches of a layer
//   LN + pos -> A image | QKV projection -> fp16 planes | attention -> A image | output projection + residual   (self)
//   the same four with the partner cloud's keys                                                                  (cross)
//   LN -> A image | FFN1 + ReLU -> A image | FFN2 + residual
// and the final LayerNorm, for all layers, issued from ONE call.  No kernel lives here: it calls the library's own entry
// points in the order model.py:TransformerCrossEncoderLayer.forward_fused does, so the result is bit-identical; what it
// removes is the Python interpreter between the launches (~9 us each -- the forward of a single pair is bound by the host,
// tools/host_profile.py).
#include "spr_common.cuh"

#include <cmath>

using namespace spr;

extern "C" int spr_layernorm256_prepare(const float*, const float*, const float*, const float*, int, float, float, void*,
                                        float*, void*);
extern "C" int spr_gemm_tc(const void*, const void*, const float*, const float*, int, int, int, int, float, int, int,
                           void*, void*, int, int, float, float, float*, void*);
extern "C" int spr_attention_varlen(const void*, const void*, int, int, int, int, int, int, const int32_t*, int, float*,
                                    int, void*, float, void*);
extern "C" int spr_attention_varlen_tc(const void*, const void*, int, int, int, int, int, int, const int32_t*, int,
                                       float*, int, void*, float, void*);

// Per layer, `ptrs` holds 16 device pointers and `scal` 9 floats (host arrays, layer-major):
//   ptrs: 0 norm1.w 1 norm1.b 2 norm2.w 3 norm2.b 4 norm3.w 5 norm3.b
//         6 self in_proj image 7 self in_proj bias 8 self out_proj image 9 self out_proj bias
//         10 cross in_proj image 11 cross in_proj bias 12 cross out_proj image 13 cross out_proj bias
//         -- followed by 4 more: 14 linear1 image 15 linear1 bias 16 linear2 image 17 linear2 bias  (18 per layer)
//   scal: 0 eps1 1 eps2 2 eps3 3..6 weight scales of (self in, self out, cross in, cross out) 7 linear1 8 linear2
extern "C" int spr_cross_encoder_forward(float* d_x, const float* d_pos, int T, int d_model, int n_heads, int d_ff,
                                         int n_layers, const void* const* ptrs, const float* scal,
                                         const int32_t* d_sa_tiles, int n_sa_tiles, const int32_t* d_ca_tiles,
                                         int n_ca_tiles, void* d_img, void* d_img_ffn, void* d_hi, void* d_lo,
                                         float a_scale, int attention_generation, const float* d_final_gamma,
                                         const float* d_final_beta, float final_eps, float* d_out, void* stream) {
  SPR_CHECK_ARG(d_x && ptrs && scal && d_sa_tiles && d_ca_tiles && d_img && d_img_ffn && d_hi && d_lo,
                "cross_encoder_forward: null pointer");
  SPR_CHECK_ARG(T > 0 && n_layers > 0 && n_sa_tiles > 0 && n_ca_tiles > 0, "cross_encoder_forward: empty input");
  SPR_CHECK_ARG(d_model == 256 && n_heads > 0 && d_model % n_heads == 0 && d_model / n_heads == 32,
                "cross_encoder_forward: d_model 256 with 32-wide heads (got %d / %d)", d_model, n_heads);
  SPR_CHECK_ARG(d_ff > 0 && d_ff % 64 == 0, "cross_encoder_forward: d_feedforward must be a multiple of 64 (got %d)", d_ff);
  SPR_CHECK_ARG(attention_generation == 1 || attention_generation == 2, "cross_encoder_forward: attention generation 1 or 2");
  const int d = d_model, hd = d / n_heads;
  // log2(e) / sqrt(head_dim), rounded once from double as the Python layer does: the soft-max is a bare exp2
  const float q_scale = (float)(1.4426950408889634 / sqrt((double)hd));
  auto attend = attention_generation == 2 ? spr_attention_varlen_tc : spr_attention_varlen;
  int rc = SPR_OK;
#define SPR_SEQ(call)          \
  do {                         \
    rc = (call);               \
    if (rc != SPR_OK) return rc; \
  } while (0)
  for (int l = 0; l < n_layers; ++l) {
    const void* const* p = ptrs + 18 * l;
    const float* s = scal + 9 * l;
    for (int a = 0; a < 2; ++a) {  // self-attention, then cross-attention (key segments = the partner cloud)
      const float* gw = static_cast<const float*>(p[2 * a]);
      const float* gb = static_cast<const float*>(p[2 * a + 1]);
      const void* w_in = p[6 + 4 * a];
      const float* b_in = static_cast<const float*>(p[7 + 4 * a]);
      const void* w_out = p[8 + 4 * a];
      const float* b_out = static_cast<const float*>(p[9 + 4 * a]);
      SPR_SEQ(spr_layernorm256_prepare(d_x, gw, gb, d_pos, T, s[a], a_scale, d_img, nullptr, stream));
      SPR_SEQ(spr_gemm_tc(d_img, w_in, b_in, nullptr, 0, T, 3 * d, d, 1.0f / (a_scale * s[3 + 2 * a]), 0, /*planes*/ 1,
                          d_hi, d_lo, 3 * d, d, q_scale, a_scale, nullptr, stream));
      SPR_SEQ(attend(d_hi, d_lo, 3 * d, 0, d, 2 * d, n_heads, hd, a == 0 ? d_sa_tiles : d_ca_tiles,
                     a == 0 ? n_sa_tiles : n_ca_tiles, nullptr, d, d_img, a_scale, stream));
      SPR_SEQ(spr_gemm_tc(d_img, w_out, b_out, d_x, d, T, d, d, 1.0f / (a_scale * s[4 + 2 * a]), 0, /*f32*/ 0, d_x, nullptr,
                          d, 0, 1.0f, a_scale, nullptr, stream));
    }
    SPR_SEQ(spr_layernorm256_prepare(d_x, static_cast<const float*>(p[4]), static_cast<const float*>(p[5]), nullptr, T, s[2],
                                     a_scale, d_img, nullptr, stream));
    SPR_SEQ(spr_gemm_tc(d_img, p[14], static_cast<const float*>(p[15]), nullptr, 0, T, d_ff, d, 1.0f / (a_scale * s[7]), 1,
                        /*image*/ 2, d_img_ffn, nullptr, d_ff, 0, 1.0f, a_scale, nullptr, stream));
    SPR_SEQ(spr_gemm_tc(d_img_ffn, p[16], static_cast<const float*>(p[17]), d_x, d, T, d, d_ff, 1.0f / (a_scale * s[8]), 0,
                        /*f32*/ 0, d_x, nullptr, d, 0, 1.0f, a_scale, nullptr, stream));
  }
  if (d_out) {
    SPR_CHECK_ARG(d_final_gamma && d_final_beta, "cross_encoder_forward: the final LayerNorm needs gamma and beta");
    SPR_SEQ(spr_layernorm256_prepare(d_x, d_final_gamma, d_final_beta, nullptr, T, final_eps, a_scale, nullptr, d_out,
                                     stream));
  }
#undef SPR_SEQ
  return SPR_OK;
}
