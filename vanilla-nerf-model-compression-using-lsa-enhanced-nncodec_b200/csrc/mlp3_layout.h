// Fused NeRF MLP kernels: operand layouts and the per-group program ("channels on TMEM lanes", table-driven control).
//
//   D[o][n] = sum_k A[o][k] * B[n][k]
//   A = weight chunk    [128 output channels x 64 k]  two K-major SWIZZLE_64B stages of 8 KB, streamed (16 KB)
//   B = activations     [256 points x 256 channels]   MN-major SWIZZLE_128B tile, resident, rewritten in place
//       or encodings    [256 points x 64 columns]     K-major  SWIZZLE_128B tile, resident
//   D = two accumulators (output channels 0..127 / 128..255) of 128 lanes x 256 columns (fp32) = all of TMEM
//
// A persistent CTA walks a fixed PROGRAM per group of 256 points: a list of weight chunks (4 MMAs each) for the
// MMA-issuing thread and a list of epilogue jobs (one per accumulator hand-over) for the 16 epilogue warps.
// Both lists are generated at compile time from the step tables below and reach the kernel as a __grid_constant__
// parameter, so the control loops are a few dozen instructions instead of an unrolled layer schedule.
//
// Layer shapes follow utils.py:18-80 of the reference (NeRF D=8, W=256, skip after layer 4, view-direction head);
// one LSA scale per output channel follows transforms.py:84-111 (ScaledLinear).
#pragma once
#include <stdint.h>

#include "net_layout.h"

namespace nerfq {

constexpr int kGroupPts = 256;               // points per CTA iteration (N of every MMA)
constexpr int kStage3Bytes = 128 * 64;       // 128 rows x 32 halves, SWIZZLE_64B
constexpr int kChunk3Bytes = 2 * kStage3Bytes;
constexpr int kAct3Bytes = 256 * 256 * 2;    // MN-major activation tile
constexpr int kEnc3Bytes = 256 * 128;        // K-major encoding tile (64 columns)
constexpr int kKGroup3 = 4096;               // 8 channels x 256 points x 2 B (SBO of the MN-major tile)
constexpr int kNGroup3 = 1024;               // 8 channels x 64 points x 2 B  (LBO)

// byte offset of (channel k, point n) in the MN-major tile
__host__ __device__ constexpr uint32_t act3_offset(uint32_t k, uint32_t n) {
    return (k >> 3) * kKGroup3 + (n >> 6) * kNGroup3 + (k & 7u) * 128u + ((((n & 63u) >> 3) ^ (k & 7u)) << 4) + (n & 7u) * 2u;
}

// Encoding-tile column order.  gamma(x) has 63 entries (run_nerf_helpers.py:23-49: x, then sin/cos blocks); two
// threads share a point -- one writes columns 0..31, the other 32..63 -- so the tile stores source column (j+1)%63
// at column j < 63 (y, z, levels 0..4 | levels 5..9, x) and zero at column 63.  gamma(d) (27 entries) is stored in
// source order in columns 0..26.
__host__ __device__ constexpr int pe_source_col(int j) { return j < 62 ? j + 1 : (j == 62 ? 0 : -1); }

struct Step3 {
    int8_t layer;        // source layer (net_layout.h order)
    int8_t halves;       // 1 or 2 blocks of 128 output rows
    int8_t kh;           // K stages (32 wide) read from the activation tile (even)
    int8_t kp;           // K stages read from the encoding tile (even; forward only)
    int16_t hcol0;       // first source column of the activation part (backward: first source column, kin offset)
    int16_t hvalid;      // valid K extent of the activation part
    int16_t pcol0;       // first source column of the encoding part
    int16_t pvalid;      // valid K extent of the encoding part (63: gamma(x) with the rotated order, 27: gamma(d))
    int16_t ch;          // channel base of the epilogue constants of the layer this step finishes
    int8_t relu;
    int8_t dst;          // accumulator of half 0 (0: D_lo, 1: D_hi); half 1 always uses D_hi
};

constexpr int kFwd3Steps = 11;
constexpr Step3 kFwd3[kFwd3Steps] = {
    {0, 2, 0, 2, 0, 0, 0, 63, 0, 1, 0},                /* L0: gamma(x)                                   */
    {1, 2, 8, 0, 0, 256, 0, 0, 256, 1, 0},             /* L1                                             */
    {2, 2, 8, 0, 0, 256, 0, 0, 512, 1, 0},             /* L2                                             */
    {3, 2, 8, 0, 0, 256, 0, 0, 768, 1, 0},             /* L3                                             */
    {4, 2, 8, 0, 0, 256, 0, 0, 1024, 1, 0},            /* L4                                             */
    {5, 2, 8, 2, 63, 256, 0, 63, 1280, 1, 0},          /* L5: hidden part + skip part (gamma(x))         */
    {6, 2, 8, 0, 0, 256, 0, 0, 1536, 1, 0},            /* L6                                             */
    {7, 2, 8, 0, 0, 256, 0, 0, 1792, 1, 0},            /* L7 (alpha head reduced in its epilogue)        */
    {9, 2, 8, 0, 0, 256, 0, 0, kChFeature, 0, 0},      /* feature_linear, no activation                  */
    {10, 1, 8, 2, 0, 256, 256, 27, kChViews, 1, 0},    /* views: feature part + direction part           */
    {11, 1, 4, 0, 0, 128, 0, 0, kChRgb, 0, 1},         /* rgb head (3 of 128 rows used), into D_hi       */
};

// backward (dgrad only): dX^T[kin][n] = sum_o W[o][hcol0 + kin] * G[n][o];  A = W^T stages, `hvalid` = number of o
constexpr int kBwd3Steps = 9;
constexpr Step3 kBwd3[kBwd3Steps] = {
    {10, 2, 4, 0, 0, 128, 0, 0, kChFeature, 0, 0},     /* views   -> d feature  (no activation)          */
    {9, 2, 8, 0, 0, 256, 0, 0, 1792, 1, 0},            /* feature -> d h8       (mask: L7 output)        */
    {7, 2, 8, 0, 0, 256, 0, 0, 1536, 1, 0},            /* L7 -> d h7                                     */
    {6, 2, 8, 0, 0, 256, 0, 0, 1280, 1, 0},            /* L6 -> d h6                                     */
    {5, 2, 8, 0, 63, 256, 0, 0, 1024, 1, 0},           /* L5 -> d h5 (hidden part of its input only)     */
    {4, 2, 8, 0, 0, 256, 0, 0, 768, 1, 0},             /* L4 -> d h4                                     */
    {3, 2, 8, 0, 0, 256, 0, 0, 512, 1, 0},             /* L3 -> d h3                                     */
    {2, 2, 8, 0, 0, 256, 0, 0, 256, 1, 0},             /* L2 -> d h2                                     */
    {1, 2, 8, 0, 0, 256, 0, 0, 0, 1, 0},               /* L1 -> d h1            (mask: L0 output)        */
};

constexpr int chunks3(const Step3* t, int n) {
    int c = 0;
    for (int i = 0; i < n; ++i) c += t[i].halves * (t[i].kh + t[i].kp) / 2;
    return c;
}
constexpr int kFwd3Chunks = chunks3(kFwd3, kFwd3Steps);     // 75
constexpr int kBwd3Chunks = chunks3(kBwd3, kBwd3Steps);     // 68
constexpr size_t kFwd3ImageBytes = (size_t)kFwd3Chunks * kChunk3Bytes;
constexpr size_t kBwd3ImageBytes = (size_t)kBwd3Chunks * kChunk3Bytes;

// ---- half-step program (MMA-issuing warp) ------------------------------------------------------
// One entry per accumulator hand-over: `n_act` weight chunks multiply K stages 0.. of the activation tile, then
// `n_enc` chunks multiply the encoding tile.  A chunk is 16 KB of weights = 2 stages = 4 MMAs (M=128, N=256, K=16).
// The first MMA of an entry overwrites the accumulator; after the last one "accumulator ready" is committed.
enum : uint32_t {
    HS_ACC_HI = 1u << 0,        // accumulate into D_hi (else D_lo)
    HS_WAIT_LO = 1u << 1,       // before chunk 0: wait for the job that rewrote channels 0..127 / drained D_lo
    HS_WAIT_HI_AT0 = 1u << 2,   // before chunk 0: wait for the job that drained D_hi
    HS_WAIT_HI_AT2 = 1u << 3,   // before activation chunk 2 (channels 128..): wait for the job that rewrote them
    HS_SF = 1u << 4,            // after activation chunks 0 and 1: commit StageFree[0], StageFree[1]
};
struct Half3 {
    uint8_t n_act, n_enc;
    uint16_t flags;
};

// ---- epilogue job program ----------------------------------------------------------------------
enum : uint32_t {
    JB_ACC_HI = 1u << 0,     // reads D_hi
    JB_RELU = 1u << 1,
    JB_ALPHA = 1u << 2,      // forward: reduce the alpha head over this layer's output
    JB_FINAL = 1u << 3,      // forward: rgb head accumulator -> raw output
    JB_WAIT_SF = 1u << 4,    // wait StageFree before writing (the other half's MMAs still read these channels)
    JB_DIR_BEFORE = 1u << 5, // forward: write gamma(d) into the encoding tile at the start of the job
    JB_PE_AFTER = 1u << 6,   // forward: write the next group's gamma(x) at the end of the job
    JB_HI_HALF = 1u << 7,    // the job handles output channels 128..255 (else 0..127)
    JB_NO_WRITE = 1u << 8,   // backward: last step, no operand for a following GEMM
    JB_ADD_ALPHA = 1u << 9,  // backward: add the alpha head's gradient to the accumulator
};
struct Job3 {
    int16_t ch;              // channel base of the constants for THIS job's 128 channels
    int16_t slot;            // saved-activation slot written (forward) / read (backward); -1 none
    uint16_t flags;
    uint16_t pad;
};

constexpr int kFwd3Jobs = 20;    // 9 steps x 2 halves + views + rgb
constexpr int kBwd3Jobs = 18;    // 9 steps x 2 halves (the views-gradient prologue job is not in the table)

struct Prog3Fwd {
    Half3 half[kFwd3Jobs];       // half-step i feeds job i
    Job3 job[kFwd3Jobs];
};
struct Prog3Bwd {
    Half3 half[kBwd3Jobs];
    Job3 job[kBwd3Jobs];
    // byte offset, inside a group's saved activations, of the 64 KB slice each epilogue job of a group reads, in
    // execution order: [0] the views-gradient prologue job, [1 + j] job j (used for L2 prefetching)
    uint32_t slice_off[kBwd3Jobs + 1];
};

// Forward program.  Waits and arrivals pair up exactly (see mlp3.cu):
//   ActLo arrivals per group: jobs (s,0) for s = 0..9        waits: first chunk of steps 1..10
//   ActHi arrivals per group: jobs (s,1) for s = 0..8, rgb   waits: chunk 2 of steps 1..9, first hi chunk of step 0
// (the steps with HS_WAIT_HI_AT2 all have >= 3 activation chunks)
constexpr Prog3Fwd make_prog3_fwd() {
    Prog3Fwd p{};
    int h = 0;
    for (int s = 0; s < kFwd3Steps; ++s) {
        const Step3& st = kFwd3[s];
        for (int mh = 0; mh < st.halves; ++mh, ++h) {
            Half3& e = p.half[h];
            uint32_t f = 0;
            e.n_act = (uint8_t)(st.kh / 2);
            e.n_enc = (uint8_t)(st.kp / 2);
            if (mh == 1 || st.dst == 1) f |= HS_ACC_HI;
            if (mh == 0 && s > 0) f |= HS_WAIT_LO;
            if (s >= 1 && s <= 9 && mh == 0) f |= HS_WAIT_HI_AT2;
            if (s == 0 && mh == 1) f |= HS_WAIT_HI_AT0;
            if (s >= 1 && s <= 8 && mh == 1) f |= HS_SF;
            e.flags = (uint16_t)f;
        }
    }
    int j = 0;
    for (int s = 0; s < kFwd3Steps; ++s) {
        const Step3& st = kFwd3[s];
        for (int mh = 0; mh < st.halves; ++mh, ++j) {
            Job3& e = p.job[j];
            uint32_t f = 0;
            if (mh == 1 || st.dst == 1) f |= JB_ACC_HI;
            if (mh == 1) f |= JB_HI_HALF;
            if (st.relu) f |= JB_RELU;
            if (s == 7) f |= JB_ALPHA;
            if (s == 10) f |= JB_FINAL;
            if (s >= 1 && s <= 8 && mh == 0) f |= JB_WAIT_SF;
            if (s == 6 && mh == 0) f |= JB_DIR_BEFORE;
            if (s == 9) f |= JB_PE_AFTER;
            e.ch = (int16_t)(st.ch + 128 * mh);
            e.slot = (int16_t)(s <= 9 ? s : -1);
            e.flags = (uint16_t)f;
        }
    }
    return p;
}

// Backward program.  The operand of the first GEMM (views gradient, channels 0..127) is produced by a prologue job
// that arrives on ActLo; jobs of step t write the operand of step t+1.
//   ActLo arrivals per group: prologue, jobs (t,0) for t = 0..7   waits: first chunk of steps 0..8
//   ActHi arrivals per group: jobs (t,1) for t = 0..8             waits: chunk 2 of steps 1..8, first hi chunk of step 0
constexpr Prog3Bwd make_prog3_bwd() {
    Prog3Bwd p{};
    int h = 0;
    for (int s = 0; s < kBwd3Steps; ++s) {
        const Step3& st = kBwd3[s];
        for (int mh = 0; mh < 2; ++mh, ++h) {
            Half3& e = p.half[h];
            uint32_t f = 0;
            e.n_act = (uint8_t)(st.kh / 2);
            e.n_enc = 0;
            if (mh == 1) f |= HS_ACC_HI;
            if (mh == 0) f |= HS_WAIT_LO;
            if (s >= 1 && mh == 0) f |= HS_WAIT_HI_AT2;
            if (s == 0 && mh == 1) f |= HS_WAIT_HI_AT0;
            // half 0's job rewrites channels 0..127 while half 1 still reads them (the last step's jobs write nothing)
            if (s < kBwd3Steps - 1 && mh == 1) f |= HS_SF;
            e.flags = (uint16_t)f;
        }
    }
    int j = 0;
    for (int s = 0; s < kBwd3Steps; ++s) {
        const Step3& st = kBwd3[s];
        for (int mh = 0; mh < 2; ++mh, ++j) {
            Job3& e = p.job[j];
            uint32_t f = 0;
            if (mh == 1) f |= JB_ACC_HI | JB_HI_HALF;
            if (st.relu) f |= JB_RELU;
            if (mh == 0 && s < kBwd3Steps - 1) f |= JB_WAIT_SF;
            if (s == kBwd3Steps - 1) f |= JB_NO_WRITE;     // job (8,0) does not arrive on ActLo either: the next
                                                           // group's prologue job (same warps, later) does
            if (s == 1) f |= JB_ADD_ALPHA;
            e.ch = (int16_t)(st.ch + 128 * mh);
            e.slot = (int16_t)(8 - s);
            e.flags = (uint16_t)f;
            p.slice_off[1 + j] = (uint32_t)(e.slot * (16 * 256 * 32) + mh * (128 * 32));
        }
    }
    p.slice_off[0] = 9u * (16 * 256 * 32);
    return p;
}

// the programs as compile-time constants: the MMA issuers are generated from them (mlp3_common.cuh, issuer3)
constexpr Prog3Fwd kProg3FwdC = make_prog3_fwd();
constexpr Prog3Bwd kProg3BwdC = make_prog3_bwd();

// ---- packed network buffer: weight images after the small fields of net_layout.h -----------------
constexpr size_t kOffFwd3Image = (kPackedBytes + 1023) / 1024 * 1024;
constexpr size_t kOffBwd3Image = kOffFwd3Image + kFwd3ImageBytes;
// The backward image holds W^T * (delta * lsa_scale) per output channel of W (the contraction index of the dgrad
// GEMM), rebuilt by nerfq_set_scale_bias from the integer-level copy below, so that the backward epilogue hands the
// masked gradient itself to the next GEMM instead of multiplying every element by the channel's scale.
// Backward scratch: long long[2440], kept zeroed.  The scale gradients are summed over CTAs and groups in FIXED POINT
// (value * 2^kGradFixShift as a 64-bit integer, native L2 atomics): integer addition is associative, so the gradients --
// and with the forward's fixed-point alpha head the whole LSA iteration -- are bit-reproducible, which float atomics
// are not (measured 2e-4 of max run to run on heavily cancelling elements).  Range +-32768 (the conversion saturates
// beyond it), resolution 3.6e-15: gradients of a mean-squared loss are 1e-7..1e-3, a summed loss still fits.
constexpr int kGradFixShift = 48;
constexpr size_t kGradTmp3Bytes = 8 * 2440;
constexpr size_t kOffGradTmp3 = (kOffBwd3Image + kBwd3ImageBytes + 15) / 16 * 16;
constexpr size_t kOffBwd3Levels = (kOffGradTmp3 + kGradTmp3Bytes + 1023) / 1024 * 1024;     // backward image, integer levels
constexpr size_t kOffBwd3StageCh = kOffBwd3Levels + kBwd3ImageBytes;      // int[2 * kBwd3Chunks]: channel of k = 0 per stage
constexpr size_t kPacked3Bytes = kOffBwd3StageCh + 4 * 2 * kBwd3Chunks;

// ---- saved activations (forward with `save`, read by the backward) ------------------------------------------
// Per group of 256 points ten slots (h1..h8, feature, views hidden) of 128 KB, each laid out
//   [16 point chunks of 16 points][256 channels][16 points] fp16      (the views slot uses channels 0..127 only)
// so that the 32 lanes of an epilogue warp (32 consecutive channels, one 16-point chunk) write / read 1 KB of
// contiguous memory per instruction -- full 128-byte lines both ways.
constexpr int kSave3Slots = 10;
constexpr size_t kSave3SlotBytes = 16 * 256 * 32;
constexpr size_t kSave3GroupBytes = kSave3Slots * kSave3SlotBytes;
// byte offset of (point chunk pc = point / 16, channel) inside a slot
__host__ __device__ constexpr uint32_t save3_offset(uint32_t pc, uint32_t channel) { return (pc * 256u + channel) * 32u; }

}  // namespace nerfq
