// The decode post-process of one pixel (diffusion_renderer_pipeline.py:300-318), shared by the stand-alone kernel
// (drb_postprocess_u8) and the fused store of the tokenizer's last stage (drb_haar_unpatch_u8): optional normal
// re-normalisation blend, then (1 + v).clamp(0, 2) / 2 * 255 -> uint8 (truncating) — every tensor op of the reference rounds
// to bf16, reproduced here.  v: the three channel values of the pixel, already bf16-rounded.
#pragma once
#include <stdint.h>

#include "ptx.cuh"

namespace drb {

__device__ __forceinline__ void postprocess_pixel(float (&v)[3], int normalize_normal, uint8_t (&out)[3]) {
  if (normalize_normal) {
    const float norm = bf16_round(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
    const float den = fmaxf(norm, 1e-12f);
    float blend = bf16_round(bf16_round(norm - 0.2f) / 0.2f);
    blend = fminf(fmaxf(blend, 0.f), 1.f);
    const float inv_blend = bf16_round(1.0f - blend);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float vn = bf16_round(v[c] / den);
      v[c] = bf16_round(bf16_round(vn * blend) + bf16_round(v[c] * inv_blend));
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float u = bf16_round(1.0f + v[c]);
    u = fminf(fmaxf(u, 0.f), 2.f);
    u = bf16_round(u * 0.5f);
    u = bf16_round(u * 255.0f);
    out[c] = static_cast<uint8_t>(u);   // truncating cast, like Tensor.to(torch.uint8)
  }
}

}  // namespace drb
