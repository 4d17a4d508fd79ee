// "Channels on TMEM lanes" formulation of the fused MLP (v2): D[o][n] = sum_k W[o][k] * X[n][k].
//   A operand = weight stage  [128 output channels x 32 k]   K-major,  SWIZZLE_64B   (8 KB, streamed)
//   B operand = activations   [256 points x 256 channels]    MN-major, SWIZZLE_128B  (128 KB, resident)
//               or encodings  [256 points x 64 columns]      K-major,  SWIZZLE_128B  (32 KB, resident)
//   D         = two accumulators of 128 lanes x 256 columns (fp32): output channels 0..127 / 128..255
// Every epilogue thread owns ONE output channel (its TMEM lane), so the per-channel constants
// {delta*lsa_scale, bias} are two registers and the scale gradient is an in-thread sum over points.
#pragma once
#include "net_layout.h"

namespace nerfq {

constexpr int kPairPoints = 256;              // points per CTA iteration (N of every MMA)
constexpr int kStage2Bytes = 128 * 64;        // 128 rows x 32 halves
constexpr int kActBytes = 256 * 256 * 2;      // MN-major activation tile
constexpr int kEncBytes = 256 * 128;          // K-major encoding tile (64 columns)
constexpr int kKGroupBytes = 4096;            // 8 channels x 256 points x 2 B  (SBO of the MN-major tile)
constexpr int kNGroupBytes = 1024;            // 8 channels x 64 points x 2 B   (LBO)

// byte offset of (channel k, point n) inside the MN-major activation tile; n multiple of 8 -> 16-byte chunk
__host__ __device__ __forceinline__ uint32_t act_offset(uint32_t k, uint32_t n) {
    return (k >> 3) * kKGroupBytes + (n >> 6) * kNGroupBytes + (k & 7u) * 128u + ((((n & 63u) >> 3) ^ (k & 7u)) << 4) + (n & 7u) * 2u;
}

struct Step2 {
    int8_t layer;        // source layer (net_layout.h order)
    int8_t halves;       // 1 or 2 blocks of 128 output channels (forward) / input channels (backward)
    int8_t kh;           // K stages (32 wide) read from the activation tile
    int8_t kp;           // K stages read from the encoding tile (forward only)
    int16_t hcol0;       // forward: first source column of the activation part; backward: first source column (kin offset)
    int16_t hvalid;      // valid K extent of the activation part
    int16_t pcol0;       // first source column of the encoding part
    int16_t pvalid;      // valid K extent of the encoding part
    int16_t ch;          // channel base of this layer's epilogue constants (-1: none)
    int8_t relu;
    int8_t dst;          // accumulator used by half 0 (0: D_lo, 1: D_hi); half 1 always uses D_hi
};

constexpr int kFwd2Steps = 11;
#define NERFQ_FWD2_TABLE                                                                                        \
    {                                                                                                           \
        {0, 2, 0, 2, 0, 0, 0, 63, 0, 1, 0},               /* L0: gamma(x)                                    */ \
        {1, 2, 8, 0, 0, 256, 0, 0, 256, 1, 0},            /* L1                                              */ \
        {2, 2, 8, 0, 0, 256, 0, 0, 512, 1, 0},            /* L2                                              */ \
        {3, 2, 8, 0, 0, 256, 0, 0, 768, 1, 0},            /* L3                                              */ \
        {4, 2, 8, 0, 0, 256, 0, 0, 1024, 1, 0},           /* L4                                              */ \
        {5, 2, 8, 2, 63, 256, 0, 63, 1280, 1, 0},         /* L5: hidden part + skip part (gamma(x))          */ \
        {6, 2, 8, 0, 0, 256, 0, 0, 1536, 1, 0},           /* L6                                              */ \
        {7, 2, 8, 0, 0, 256, 0, 0, 1792, 1, 0},           /* L7 (alpha head in its epilogue)                 */ \
        {9, 2, 8, 0, 0, 256, 0, 0, kChFeature, 0, 0},     /* feature_linear, no activation                   */ \
        {10, 1, 8, 1, 0, 256, 256, 27, kChViews, 1, 0},   /* views: feature part + direction part (gamma(d)) */ \
        {11, 1, 4, 0, 0, 128, 0, 0, kChRgb, 0, 1},        /* rgb head (3 of 128 rows used), into D_hi        */ \
    }
constexpr Step2 kFwd2[kFwd2Steps] = NERFQ_FWD2_TABLE;

constexpr int kBwd2Steps = 9;
#define NERFQ_BWD2_TABLE                                                                       \
    {                                                                                          \
        {10, 2, 4, 0, 0, 128, 0, 0, kChFeature, 0, 0},  /* views   -> d feature             */ \
        {9, 2, 8, 0, 0, 256, 0, 0, 1792, 1, 0},         /* feature -> d h8  (mask: L7)      */ \
        {7, 2, 8, 0, 0, 256, 0, 0, 1536, 1, 0},         /* L7 -> d h7       (mask: L6)      */ \
        {6, 2, 8, 0, 0, 256, 0, 0, 1280, 1, 0},         /* L6 -> d h6       (mask: L5)      */ \
        {5, 2, 8, 0, 63, 256, 0, 0, 1024, 1, 0},        /* L5 -> d h5       (mask: L4)      */ \
        {4, 2, 8, 0, 0, 256, 0, 0, 768, 1, 0},          /* L4 -> d h4                       */ \
        {3, 2, 8, 0, 0, 256, 0, 0, 512, 1, 0},          /* L3 -> d h3                       */ \
        {2, 2, 8, 0, 0, 256, 0, 0, 256, 1, 0},          /* L2 -> d h2                       */ \
        {1, 2, 8, 0, 0, 256, 0, 0, 0, 1, 0},            /* L1 -> d h1       (mask: L0)      */ \
    }
constexpr Step2 kBwd2[kBwd2Steps] = NERFQ_BWD2_TABLE;

// compile-time accessors usable from fully unrolled device loops
__host__ __device__ constexpr Step2 fwd2_step(int s) {
    constexpr Step2 t[kFwd2Steps] = NERFQ_FWD2_TABLE;
    return t[s];
}
__host__ __device__ constexpr Step2 bwd2_step(int s) {
    constexpr Step2 t[kBwd2Steps] = NERFQ_BWD2_TABLE;
    return t[s];
}

constexpr int stages2(const Step2* t, int n) {
    int s = 0;
    for (int i = 0; i < n; ++i) s += t[i].halves * (t[i].kh + t[i].kp);
    return s;
}
constexpr int kFwd2Stages = stages2(kFwd2, kFwd2Steps);
constexpr int kBwd2Stages = stages2(kBwd2, kBwd2Steps);
constexpr size_t kFwd2ImageBytes = (size_t)kFwd2Stages * kStage2Bytes;
constexpr size_t kBwd2ImageBytes = (size_t)kBwd2Stages * kStage2Bytes;

// v2 images are appended to the packed network buffer
constexpr size_t kOffFwd2Image = (kPackedBytes + 1023) / 1024 * 1024;
constexpr size_t kOffBwd2Image = kOffFwd2Image + kFwd2ImageBytes;
constexpr size_t kPacked2Bytes = kOffBwd2Image + kBwd2ImageBytes;
static_assert(kOffFwd2Image % 1024 == 0, "alignment");

// saved activations (v2): per pair of tiles 9 MN-major images of 128 KB (h1..h8, feature) + one of 64 KB
// (views hidden, channels 0..127)
constexpr size_t kSave2PairBytes = 9 * (size_t)kActBytes + kActBytes / 2;

}  // namespace nerfq
