// Packed on-device format of one vanilla-NeRF network (D=8, W=256, skip after layer 4, view-direction
// head) as the fused MLP kernels consume it.  Mirrors utils.py:18-80 of the reference (layer shapes)
// and transforms.py:84-111 (one LSA scale per output channel).
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

constexpr int kTileM = 128;             // points per tile (= TMEM lanes)
constexpr int kStageK = 32;             // K extent of one weight stage (two K=16 MMA steps)
constexpr int kStageRowBytes = 64;      // 32 halves per row, SWIZZLE_64B
constexpr int kABlockBytes = 16384;     // 128 rows x 64 halves, SWIZZLE_128B
constexpr int kABufBytes = 4 * kABlockBytes;

// ---- tensor-core steps of the forward pass -------------------------------------------------
struct MmaStep {
    int8_t layer;        // source layer of the weights
    int16_t col0;        // first source column (forward) / first source row block (backward)
    int16_t ncols;       // valid source columns (rest zero-padded up to stages*32)
    int8_t stages;       // number of 32-wide weight stages
    int16_t n;           // MMA N (rows of the weight stage)
    int8_t a_blk0;       // first 64-wide block of the activation tile read by this step
    int8_t accumulate;   // 1: add to the accumulator of the previous step
};

constexpr int kFwdSteps = 12;
#define NERFQ_FWD_STEP_TABLE                                                                          \
    {                                                                                                 \
        {0, 0, 63, 2, 256, 0, 0},       /* L0   : gamma(x) -> 256                                  */ \
        {1, 0, 256, 8, 256, 0, 0},      /* L1                                                      */ \
        {2, 0, 256, 8, 256, 0, 0},      /* L2                                                      */ \
        {3, 0, 256, 8, 256, 0, 0},      /* L3                                                      */ \
        {4, 0, 256, 8, 256, 0, 0},      /* L4                                                      */ \
        {5, 63, 256, 8, 256, 0, 0},     /* L5, hidden part  W5[:, 63:319]                          */ \
        {5, 0, 63, 2, 256, 0, 1},       /* L5, skip part    W5[:, 0:63] (gamma(x) back in block 0) */ \
        {6, 0, 256, 8, 256, 0, 0},      /* L6                                                      */ \
        {7, 0, 256, 8, 256, 0, 0},      /* L7   (alpha head evaluated in its epilogue)             */ \
        {9, 0, 256, 8, 256, 0, 0},      /* feature_linear (no activation)                          */ \
        {10, 0, 256, 8, 128, 0, 0},     /* views, feature part   Wv[:, 0:256]                      */ \
        {10, 256, 27, 1, 128, 0, 1},    /* views, direction part Wv[:, 256:283] (gamma(d), blk 0)  */ \
    }
constexpr MmaStep kFwd[kFwdSteps] = NERFQ_FWD_STEP_TABLE;

// ---- tensor-core steps of the backward pass (dgrad only; B operand = W^T) -------------------
// dX[m, kin] = sum_o G[m, o] * W[o, col0 + kin], kin < 256.  `ncols` = number of o (K extent).
constexpr int kBwdSteps = 9;
#define NERFQ_BWD_STEP_TABLE                                                              \
    {                                                                                     \
        {10, 0, 128, 4, 256, 0, 0},     /* views   -> d feature                        */ \
        {9, 0, 256, 8, 256, 0, 0},      /* feature -> d h8                             */ \
        {7, 0, 256, 8, 256, 0, 0},      /* L7 -> d h7                                  */ \
        {6, 0, 256, 8, 256, 0, 0},      /* L6 -> d h6                                  */ \
        {5, 63, 256, 8, 256, 0, 0},     /* L5 -> d h5 (hidden part of its input only)  */ \
        {4, 0, 256, 8, 256, 0, 0},      /* L4 -> d h4                                  */ \
        {3, 0, 256, 8, 256, 0, 0},      /* L3 -> d h3                                  */ \
        {2, 0, 256, 8, 256, 0, 0},      /* L2 -> d h2                                  */ \
        {1, 0, 256, 8, 256, 0, 0},      /* L1 -> d h1                                  */ \
    }
constexpr MmaStep kBwd[kBwdSteps] = NERFQ_BWD_STEP_TABLE;

constexpr int stage_bytes(const MmaStep& s) { return s.n * kStageRowBytes; }

constexpr int total_stages(const MmaStep* steps, int n) {
    int t = 0;
    for (int i = 0; i < n; ++i) t += steps[i].stages;
    return t;
}
constexpr int image_bytes(const MmaStep* steps, int n) {
    int t = 0;
    for (int i = 0; i < n; ++i) t += steps[i].stages * stage_bytes(steps[i]);
    return t;
}
constexpr int kFwdStages = total_stages(kFwd, kFwdSteps);
constexpr int kBwdStages = total_stages(kBwd, kBwdSteps);
constexpr int kFwdImageBytes = image_bytes(kFwd, kFwdSteps);
constexpr int kBwdImageBytes = image_bytes(kBwd, kBwdSteps);

// ---- packed network buffer -------------------------------------------------------------------
// [ fwd image (fp16) | bwd image (fp16) | sb: float2{eff_scale, bias}[2436] | delta[2436] |
//   w_alpha float[256] (levels) | w_rgb float[3*128] (levels) ]
constexpr size_t kOffFwdImage = 0;
constexpr size_t kOffBwdImage = kOffFwdImage + kFwdImageBytes;
constexpr size_t kOffSB = kOffBwdImage + kBwdImageBytes;
constexpr size_t kOffDelta = kOffSB + sizeof(float) * 2 * kNumChannels;
constexpr size_t kOffScale = kOffDelta + sizeof(float) * kNumChannels;
constexpr size_t kOffWAlpha = kOffScale + sizeof(float) * kNumChannels;
constexpr size_t kOffWRgb = kOffWAlpha + sizeof(float) * 256;
constexpr size_t kPackedBytesRaw = kOffWRgb + sizeof(float) * 384;
constexpr size_t kPackedBytes = (kPackedBytesRaw + 255) / 256 * 256;

static_assert(kOffBwdImage % 1024 == 0 && kOffSB % 16 == 0, "alignment");

// ---- saved activations for the backward pass (written by forward when `save` != null) --------
// Per tile: 9 slots of 64 KB (h1..h5, h6, h7, h8, feature: operand-tile images, fp16,
// SWIZZLE_128B blocks) followed by one slot of 32 KB (views hidden h_v, 128 wide).
constexpr int kSaveSlotsFull = 9;
constexpr size_t kSaveTileBytes = (size_t)kSaveSlotsFull * kABufBytes + 2 * kABlockBytes;

}  // namespace nerfq
