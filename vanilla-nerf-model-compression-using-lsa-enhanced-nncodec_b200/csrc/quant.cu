// NNCodec uniform ("baseline") quantiser on the GPU: float32 -> int32 levels and back.
//   nerfq_quantize_urq   replaces deepCABAC Encoder.quantLayer(dq_flag=0)   nnc_core/approximator/baseline.py:48-57
//   nerfq_dequantize     replaces deepCABAC Decoder.dequantLayer            nnc_core/approximator/baseline.py:98
//   nerfq_stepsize       nnc_core/common.py:28-46
// Arithmetic (bit-identical to oracle/quant_oracle.c): level = sign(w) * (int)(|w| / delta + 0.5f) with IEEE
// float32 division and addition; value = (float)level * delta.  The qp actually used is raised until the
// largest level fits int32 (the clip contract of baseline.py:60-62).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace nerfq {

__host__ __device__ inline float stepsize(int qp, int qp_density) {
    const int k = 1 << qp_density;
    const int mul = k + (qp & (k - 1));
    const int shift = (qp >> qp_density) - qp_density;
    return ldexpf((float)mul, shift);
}

// Non-finite input (a diverged LSA scale, a corrupt tensor): a NaN / Inf maximum leaves the qp as requested and the
// non-finite elements get level 0 (to_level); the search is bounded -- the step size doubles every 2^qp_density steps, so
// FLT_MAX / delta drops below 2^31 long before kMaxClipSteps.
constexpr int kMaxClipSteps = 1024;
__host__ __device__ inline int clip_qp(float max_abs, int qp, int qp_density) {
    if (!(max_abs <= 3.402823466e+38f)) return qp;
    for (int it = 0; it < kMaxClipSteps; ++it) {
        const float d = stepsize(qp, qp_density);
#ifdef __CUDA_ARCH__
        const float q = __fadd_rn(__fdiv_rn(max_abs, d), 0.5f);
#else
        volatile float t = max_abs / d;
        const float q = t + 0.5f;
#endif
        if (q < 2147483648.0f) return qp;
        ++qp;
    }
    return qp;
}

// level of one value: nearest integer of |x| / d, ties away from zero; NaN / Inf -> 0; saturates at INT32_MAX
__device__ __forceinline__ int to_level(float x, float d) {
    const float a = fabsf(x);
    if (!(a <= 3.402823466e+38f)) return 0;
    const float q = __fadd_rn(__fdiv_rn(a, d), 0.5f);
    const int m = q >= 2147483648.0f ? 2147483647 : (int)q;
    return x < 0.0f ? -m : m;
}

// max |w| via integer atomicMax on the float bit pattern (non-negative floats order like ints)
__global__ void absmax_kernel(const float* __restrict__ w, long long n, unsigned int* __restrict__ out) {
    unsigned int m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(fabsf(w[i])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__global__ void quantize_kernel(const float* __restrict__ w, int32_t* __restrict__ lvl, long long n, int qp, int qp_density,
                                const unsigned int* __restrict__ absmax_bits, int* __restrict__ qp_used) {
    const int q = clip_qp(__uint_as_float(*absmax_bits), qp, qp_density);
    const float d = stepsize(q, qp_density);
    if (blockIdx.x == 0 && threadIdx.x == 0 && qp_used) *qp_used = q;
    const long long n4 = n >> 2;
    const float4* w4 = reinterpret_cast<const float4*>(w);
    int4* l4 = reinterpret_cast<int4*>(lvl);
    auto one = [&](float x) { return to_level(x, d); };
    const bool vec = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(lvl)) & 15) == 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        for (long long i = tid; i < n4; i += stride) {
            const float4 x = w4[i];
            l4[i] = make_int4(one(x.x), one(x.y), one(x.z), one(x.w));
        }
        for (long long i = (n4 << 2) + tid; i < n; i += stride) lvl[i] = one(w[i]);
    } else {
        for (long long i = tid; i < n; i += stride) lvl[i] = one(w[i]);
    }
}

__global__ void dequantize_kernel(const int32_t* __restrict__ lvl, float* __restrict__ w, long long n, float d) {
    const long long n4 = n >> 2;
    const bool vec = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(lvl)) & 15) == 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const int4* l4 = reinterpret_cast<const int4*>(lvl);
        float4* w4 = reinterpret_cast<float4*>(w);
        for (long long i = tid; i < n4; i += stride) {
            const int4 x = l4[i];
            w4[i] = make_float4(__fmul_rn((float)x.x, d), __fmul_rn((float)x.y, d), __fmul_rn((float)x.z, d), __fmul_rn((float)x.w, d));
        }
        for (long long i = (n4 << 2) + tid; i < n; i += stride) w[i] = __fmul_rn((float)lvl[i], d);
    } else {
        for (long long i = tid; i < n; i += stride) w[i] = __fmul_rn((float)lvl[i], d);
    }
}

// ---- batched form: every tensor of a model in one launch pair ------------------------------------------
constexpr int kMaxBatch = 64;
struct BatchDesc {
    const float* w[kMaxBatch];
    int32_t* lvl[kMaxBatch];
    float* rec[kMaxBatch];          // nullable per tensor; may alias w
    long long n[kMaxBatch];
    int qp[kMaxBatch];
    int count;
    int qp_density;
};

__global__ void absmax_batch_kernel(const __grid_constant__ BatchDesc b, unsigned int* __restrict__ out) {
    const int t = blockIdx.y;
    const float* __restrict__ w = b.w[t];
    const long long n = b.n[t];
    unsigned int m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(fabsf(w[i])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out + t, m);
}

__global__ void quantize_batch_kernel(const __grid_constant__ BatchDesc b, const unsigned int* __restrict__ absmax_bits,
                                      int* __restrict__ qp_used) {
    const int t = blockIdx.y;
    const float* w = b.w[t];
    int32_t* __restrict__ lvl = b.lvl[t];
    float* rec = b.rec[t];
    const long long n = b.n[t];
    const int q = clip_qp(__uint_as_float(absmax_bits[t]), b.qp[t], b.qp_density);
    const float d = stepsize(q, b.qp_density);
    if (blockIdx.x == 0 && threadIdx.x == 0 && qp_used) qp_used[t] = q;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int l = to_level(w[i], d);
        lvl[i] = l;
        if (rec) rec[i] = __fmul_rn((float)l, d);
    }
}

}  // namespace nerfq

using namespace nerfq;

// Quantise (and optionally reconstruct, rec[t] != NULL, in place allowed) `count` tensors with one launch pair.
// All arrays are HOST arrays of length count holding device pointers / sizes / qps; workspace: 4*count bytes of
// device memory; qp_used: nullable device int[count].
extern "C" int nerfq_quantize_batch(const float* const* w, int32_t* const* lvl, float* const* rec, const long long* n, const int* qp,
                                    int count, int qp_density, int* qp_used, void* workspace, cudaStream_t stream) {
    if (count == 0) return 0;
    if (!w || !lvl || !n || !qp || !workspace || count < 0 || count > kMaxBatch || qp_density < 0 || qp_density > 8) return -1;
    BatchDesc b{};
    long long n_max = 0;
    for (int t = 0; t < count; ++t) {
        if (!w[t] || !lvl[t] || n[t] < 0) return -1;
        b.w[t] = w[t]; b.lvl[t] = lvl[t]; b.rec[t] = rec ? rec[t] : nullptr; b.n[t] = n[t]; b.qp[t] = qp[t];
        if (n[t] > n_max) n_max = n[t];
    }
    b.count = count;
    b.qp_density = qp_density;
    long long bx = (n_max + 1023) / 1024;
    if (bx < 1) bx = 1;
    if (bx > 37) bx = 37;            // 37 x 4 = 148: with the per-tensor grid dimension the launch covers every SM
    cudaMemsetAsync(workspace, 0, 4 * (size_t)count, stream);
    const dim3 grid((unsigned)bx, (unsigned)count);
    absmax_batch_kernel<<<grid, 256, 0, stream>>>(b, reinterpret_cast<unsigned int*>(workspace));
    quantize_batch_kernel<<<grid, 256, 0, stream>>>(b, reinterpret_cast<const unsigned int*>(workspace), qp_used);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int nerfq_stepsize(int qp, int qp_density, float* out) {
    if (!out || qp_density < 0 || qp_density > 8) return -1;
    *out = stepsize(qp, qp_density);
    return 0;
}

static unsigned grid_for(long long n) {
    long long b = (n / 4 + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148 * 8) b = 148 * 8;      // grid-stride over a multiple of the SM count
    return (unsigned)b;
}

// workspace: 4 bytes of device memory (max|w| bits); qp_used: nullable device int.
extern "C" int nerfq_quantize_urq(const float* w, int32_t* lvl, long long n, int qp, int qp_density, int* qp_used,
                                  void* workspace4, cudaStream_t stream) {
    if (n == 0) return 0;
    if (!w || !lvl || !workspace4 || n < 0 || qp_density < 0 || qp_density > 8) return -1;
    cudaMemsetAsync(workspace4, 0, 4, stream);
    absmax_kernel<<<grid_for(n), 256, 0, stream>>>(w, n, reinterpret_cast<unsigned int*>(workspace4));
    quantize_kernel<<<grid_for(n), 256, 0, stream>>>(w, lvl, n, qp, qp_density, reinterpret_cast<unsigned int*>(workspace4), qp_used);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int nerfq_dequantize(const int32_t* lvl, float* w, long long n, int qp, int qp_density, cudaStream_t stream) {
    if (n == 0) return 0;
    if (!w || !lvl || n < 0 || qp_density < 0 || qp_density > 8) return -1;
    dequantize_kernel<<<grid_for(n), 256, 0, stream>>>(lvl, w, n, stepsize(qp, qp_density));
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
