// Constants of one vanilla-NeRF network (D=8, W=256, skip after layer 4, view-direction head) shared by the packing
// and MLP kernels.  Mirrors utils.py:18-80 of the reference (layer shapes) and transforms.py:84-111 (one LSA scale per
// output channel).
//
// Layer index (== order of the 12 weight pointers handed to nerfq_pack_net):
//   0..7  pts_linears.0..7   [256 x 63], [256 x 256] x4, [256 x 319], [256 x 256] x2
//   8     alpha_linear       [1 x 256]
//   9     feature_linear     [256 x 256]
//   10    views_linears.0    [128 x 283]
//   11    rgb_linear         [3 x 128]
//
// Channel index (flat per-output-channel arrays: LSA scale, bias, scale gradient), kernel order:
//   pts0..7 -> 0..2047, feature -> 2048..2303, views -> 2304..2431, alpha -> 2432, rgb -> 2433..2435
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace nerfq {

constexpr int kNumLayers = 12;
constexpr int kNumChannels = 2436;
constexpr int kChFeature = 2048;
constexpr int kChViews = 2304;
constexpr int kChAlpha = 2432;
constexpr int kChRgb = 2433;

constexpr int kLayerOut[kNumLayers] = {256, 256, 256, 256, 256, 256, 256, 256, 1, 256, 128, 3};
constexpr int kLayerIn[kNumLayers] = {63, 256, 256, 256, 256, 319, 256, 256, 256, 256, 283, 128};
constexpr int kLayerCh[kNumLayers] = {0, 256, 512, 768, 1024, 1280, 1536, 1792, kChAlpha, kChFeature, kChViews, kChRgb};

// ---- packed network buffer, small fields (the weight images follow, see mlp3_layout.h) ----------
// [ sb: float2{delta*lsa_scale, bias}[2436] | delta[2436] | lsa_scale[2436] | w_alpha float[256] (levels) |
//   w_rgb float[3*128] (levels) | layer_max float[16] ]
constexpr size_t kOffSB = 0;
constexpr size_t kOffDelta = kOffSB + sizeof(float) * 2 * kNumChannels;
constexpr size_t kOffScale = kOffDelta + sizeof(float) * kNumChannels;
constexpr size_t kOffWAlpha = kOffScale + sizeof(float) * kNumChannels;
constexpr size_t kOffWRgb = kOffWAlpha + sizeof(float) * 256;
constexpr size_t kOffLayerMax = kOffWRgb + sizeof(float) * 384;        // float bits: max |level| (or |weight|) per layer, nerfq_pack_status
constexpr size_t kPackedBytesRaw = kOffLayerMax + sizeof(float) * 16;
constexpr size_t kPackedBytes = (kPackedBytesRaw + 255) / 256 * 256;

static_assert(kOffSB % 16 == 0, "alignment");

}  // namespace nerfq
