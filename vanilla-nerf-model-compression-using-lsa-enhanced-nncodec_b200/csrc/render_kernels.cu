// HBM-bound kernels of the ray-rendering path (one warp per ray, coalesced segment-per-lane access):
//   nerfq_coarse_depths   stratified depths                       run_nerf.py:379-403
//   nerfq_composite_fwd   raw2outputs                             run_nerf.py:285-345
//   nerfq_composite_bwd   its gradient w.r.t. raw (autograd equivalent)
//   nerfq_sample_fine     sample_pdf + concat + sort + z_std      run_nerf_helpers.py:119-163, run_nerf.py:423-429,451
//   nerfq_camera_rays / nerfq_pack_rays   get_rays, viewdirs, ndc_rays, ray packing   run_nerf_helpers.py:71-115, run_nerf.py:108-142
//   nerfq_mse_grad        img2mse (x2) and its gradient           run_nerf_helpers.py:12, run_nerf.py:741-751
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <type_traits>

namespace nerfq {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxPerLane = 8;      // samples per lane: supports up to 256 samples per ray
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
// exclusive product / sum scans across lanes (lane 0 gets the identity)
__device__ __forceinline__ float warp_excl_prod(float v, int lane) {
    float incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl *= t;
    }
    float ex = __shfl_up_sync(kFull, incl, 1);
    return lane == 0 ? 1.0f : ex;
}
__device__ __forceinline__ float warp_excl_sum(float v, int lane) {
    float incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    float ex = __shfl_up_sync(kFull, incl, 1);
    return lane == 0 ? 0.0f : ex;
}
// suffix (reverse) exclusive sum across lanes: lane i gets sum over lanes > i
__device__ __forceinline__ float warp_excl_suffix_sum(float v, int lane) {
    float incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_down_sync(kFull, incl, o);
        if (lane + o < 32) incl += t;
    }
    float ex = __shfl_down_sync(kFull, incl, 1);
    return lane == 31 ? 0.0f : ex;
}

// torch.linspace(0, 1, n)[i] as ATen's CUDA kernel computes it (symmetric halves, fused multiply-add);
// ATen's vectorised CPU kernel differs from this in the last bit of some entries.
__device__ __forceinline__ float linspace01(int i, int n) {
    const float step = 1.0f / (float)(n - 1);
    return i < n / 2 ? fmaf(step, (float)i, 0.0f) : fmaf(-step, (float)(n - 1 - i), 1.0f);
}

// ---------------------------------------------------------------------------------------------
// coarse depths
// ---------------------------------------------------------------------------------------------
__global__ void coarse_depths_kernel(const float* __restrict__ rays, const float* __restrict__ t_rand, long long n_rays, int S,
                                     int lindisp, float* __restrict__ z_out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rays * S) return;
    const long long ray = idx / S;
    const int i = (int)(idx - ray * S);
    const float near = rays[ray * 11 + 6], far = rays[ray * 11 + 7];
    auto base = [&](int k) {
        const float t = linspace01(k, S);
        const float omt = __fsub_rn(1.0f, t);
        if (lindisp) return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(__fdiv_rn(1.0f, near), omt), __fmul_rn(__fdiv_rn(1.0f, far), t)));
        return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
    };
    float z = base(i);
    if (t_rand) {
        const float lower = i == 0 ? z : __fmul_rn(0.5f, __fadd_rn(z, base(i - 1)));
        const float upper = i == S - 1 ? z : __fmul_rn(0.5f, __fadd_rn(base(i + 1), z));
        z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[idx]));
    }
    z_out[idx] = z;
}

// ---------------------------------------------------------------------------------------------
// compositing
// ---------------------------------------------------------------------------------------------
// PER = samples per lane, a compile-time constant (ceil(S / 32): 2 for the coarse pass, 6 for the fine one): with a
// run-time bound every loop below ran all 8 slots -- 4x the work at S = 64 -- and kept 8 slots of every array live
// (round 2: composite_fwd at 32768 rays 32.8 -> see profiles/r02_ncu_ray_kernels_summary.txt).
template <int PER>
struct RaySeg {
    float alpha[PER], dist[PER], sig[PER], zz[PER];
    float c[PER][3];
};

// Loads this lane's contiguous segment [lane*per, lane*per+per) of one ray and evaluates alpha / colour.
template <int PER>
__device__ __forceinline__ void load_segment(RaySeg<PER>& s, const float* __restrict__ raw, const float* __restrict__ z,
                                             const float* __restrict__ noise, float dnorm, int S, int lane) {
    const int i0 = lane * PER;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = i0 + k;
        if (i < S) {
            const float4 r = *reinterpret_cast<const float4*>(raw + 4 * i);
            const float zi = z[i];
            const float gap = (i + 1 < S) ? __fsub_rn(z[i + 1], zi) : 1e10f;
            const float d = __fmul_rn(gap, dnorm);
            float sg = r.w;
            if (noise) sg = __fadd_rn(sg, noise[i]);
            s.sig[k] = sg;
            s.dist[k] = d;
            s.zz[k] = zi;
            s.alpha[k] = __fsub_rn(1.0f, expf(-__fmul_rn(fmaxf(sg, 0.0f), d)));
            s.c[k][0] = 1.0f / (1.0f + expf(-r.x));
            s.c[k][1] = 1.0f / (1.0f + expf(-r.y));
            s.c[k][2] = 1.0f / (1.0f + expf(-r.z));
        } else {
            s.sig[k] = 0.f; s.dist[k] = 0.f; s.zz[k] = 0.f; s.alpha[k] = 0.f;
            s.c[k][0] = s.c[k][1] = s.c[k][2] = 0.f;
        }
    }
}

template <int PER>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) composite_fwd_kernel(
    const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays, const float* __restrict__ noise,
    int white_bkgd, long long n_rays, int S, float* __restrict__ rgb, float* __restrict__ disp, float* __restrict__ acc,
    float* __restrict__ depth, float* __restrict__ weights) {
    const int lane = threadIdx.x & 31;
    const long long ray = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    constexpr int per = PER;
    const float* r = rays + ray * 11;
    const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r[3], r[3]), __fmul_rn(r[4], r[4])), __fmul_rn(r[5], r[5])));
    RaySeg<PER> s;
    load_segment<PER>(s, raw + ray * S * 4, z + ray * S, noise ? noise + ray * S : nullptr, dnorm, S, lane);
    float prod = 1.0f;
#pragma unroll
    for (int k = 0; k < PER; ++k) prod *= __fadd_rn(__fsub_rn(1.0f, s.alpha[k]), 1e-10f);
    float T = warp_excl_prod(prod, lane);
    float a_rgb[3] = {0.f, 0.f, 0.f}, a_depth = 0.f, a_acc = 0.f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = lane * per + k;
        if (i < S) {
            const float w = __fmul_rn(s.alpha[k], T);
            T *= __fadd_rn(__fsub_rn(1.0f, s.alpha[k]), 1e-10f);
            if (weights) weights[ray * S + i] = w;
            a_rgb[0] = fmaf(w, s.c[k][0], a_rgb[0]);
            a_rgb[1] = fmaf(w, s.c[k][1], a_rgb[1]);
            a_rgb[2] = fmaf(w, s.c[k][2], a_rgb[2]);
            a_depth = fmaf(w, s.zz[k], a_depth);
            a_acc += w;
        }
    }
    a_rgb[0] = warp_sum(a_rgb[0]); a_rgb[1] = warp_sum(a_rgb[1]); a_rgb[2] = warp_sum(a_rgb[2]);
    a_depth = warp_sum(a_depth); a_acc = warp_sum(a_acc);
    if (lane == 0) {
        if (white_bkgd) {
            const float bg = __fsub_rn(1.0f, a_acc);
            a_rgb[0] += bg; a_rgb[1] += bg; a_rgb[2] += bg;
        }
        rgb[ray * 3 + 0] = a_rgb[0]; rgb[ray * 3 + 1] = a_rgb[1]; rgb[ray * 3 + 2] = a_rgb[2];
        const float q = __fdiv_rn(a_depth, a_acc);                  // 0/0 -> NaN, propagated like torch.max does
        const float m = (q != q) ? q : fmaxf(1e-10f, q);
        disp[ray] = __fdiv_rn(1.0f, m);
        acc[ray] = a_acc;
        if (depth) depth[ray] = a_depth;
    }
}

// d_raw from d_rgb (the only output the LSA objective uses, run_nerf.py:741-751).
template <int PER>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) composite_bwd_kernel(
    const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays, const float* __restrict__ noise,
    int white_bkgd, const float* __restrict__ d_rgb, long long n_rays, int S, float* __restrict__ d_raw) {
    const int lane = threadIdx.x & 31;
    const long long ray = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    constexpr int per = PER;
    const float* r = rays + ray * 11;
    const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r[3], r[3]), __fmul_rn(r[4], r[4])), __fmul_rn(r[5], r[5])));
    RaySeg<PER> s;
    load_segment<PER>(s, raw + ray * S * 4, z + ray * S, noise ? noise + ray * S : nullptr, dnorm, S, lane);
    const float g0 = d_rgb[ray * 3 + 0], g1 = d_rgb[ray * 3 + 1], g2 = d_rgb[ray * 3 + 2];
    const float gbg = white_bkgd ? (g0 + g1 + g2) : 0.0f;
    float prod = 1.0f;
#pragma unroll
    for (int k = 0; k < PER; ++k) prod *= __fadd_rn(__fsub_rn(1.0f, s.alpha[k]), 1e-10f);
    float T = warp_excl_prod(prod, lane);
    float Tk[PER], wk[PER], dwk[PER];
    float seg = 0.0f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        Tk[k] = T;
        wk[k] = s.alpha[k] * T;
        dwk[k] = g0 * s.c[k][0] + g1 * s.c[k][1] + g2 * s.c[k][2] - gbg;     // dL/dw_i
        T *= __fadd_rn(__fsub_rn(1.0f, s.alpha[k]), 1e-10f);
        seg += dwk[k] * wk[k];
    }
    float suffix = warp_excl_suffix_sum(seg, lane);      // sum over later lanes of dw*w
#pragma unroll
    for (int k = PER - 1; k >= 0; --k) {
        const int i = lane * per + k;
        if (i < S) {
            const float a = __fadd_rn(__fsub_rn(1.0f, s.alpha[k]), 1e-10f);
            const float dalpha = dwk[k] * Tk[k] - suffix / a;
            const float dsig = (s.sig[k] > 0.0f) ? dalpha * s.dist[k] * (1.0f - s.alpha[k]) : 0.0f;
            float4 o;
            o.x = g0 * wk[k] * s.c[k][0] * (1.0f - s.c[k][0]);
            o.y = g1 * wk[k] * s.c[k][1] * (1.0f - s.c[k][1]);
            o.z = g2 * wk[k] * s.c[k][2] * (1.0f - s.c[k][2]);
            o.w = dsig;
            *reinterpret_cast<float4*>(d_raw + (ray * S + i) * 4) = o;
            suffix += dwk[k] * wk[k];
        }
    }
}

// Bitonic sort of 32 * PER keys held by a warp, element i = lane * PER + k in register v[k] (blocked layout), ascending.
// Strides below PER are compare-exchanges between a thread's own registers; larger strides exchange whole registers with
// the partner lane by shuffle -- no shared-memory round trips and no barriers (the earlier shared-memory network took
// 36 passes with a warp barrier each: 164 us per 32768 rays, ncu r02).  The keys are the floats' bits mapped to unsigned
// integers of the same order (float_key): a compare-exchange is then one integer min / max per element instead of two
// float compares and selects, a total order exists for every bit pattern (NaN sorts last, like torch.sort), and min / max
// never duplicate or drop a value -- the result is the sorted multiset.
__device__ __forceinline__ unsigned float_key(float x) {
    const unsigned b = __float_as_uint(x);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}
template <int PER>
__device__ __forceinline__ void warp_bitonic_sort(unsigned (&v)[PER], int lane) {
#pragma unroll
    for (int size = 2; size <= 32 * PER; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= PER) {
                const int m = stride / PER;                                   // partner lane = lane ^ m
                const bool lower = (lane & m) == 0;
                const bool asc = ((lane * PER) & size) == 0;                  // size > stride >= PER: decided by the lane
                const bool keep_min = lower == asc;
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const unsigned mine = v[k];
                    const unsigned other = __shfl_xor_sync(kFull, mine, m);
                    v[k] = keep_min ? min(mine, other) : max(mine, other);
                }
            } else {
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    if ((k & stride) == 0) {
                        const bool asc = size >= PER ? (((lane * PER) & size) == 0) : ((k & size) == 0);
                        const unsigned a = v[k], b = v[k | stride];
                        v[k] = asc ? min(a, b) : max(a, b);
                        v[k | stride] = asc ? max(a, b) : min(a, b);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// importance sampling + merge
// ---------------------------------------------------------------------------------------------
// One warp per ray.  Shared memory per warp: cdf[S-1], bins[S-1], sort buffer[P] with P = pow2 >= S+Ni.
// PER: values per lane in the final sort (32 * PER >= S + Ni); WMAX: power-of-two bound on the interior weights per lane
// ((S - 2 + 31) / 32 <= WMAX): the pdf / cdf loops run WMAX slots instead of 8.
template <int PER, int WMAX>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) sample_fine_kernel(
    const float* __restrict__ z_coarse, const float* __restrict__ bins_in, const float* __restrict__ weights,
    const float* __restrict__ u_in, long long n_rays, int S, int Ni, int P, float* __restrict__ z_out, float* __restrict__ z_std,
    float* __restrict__ z_samples_out) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long ray = (long long)blockIdx.x * kWarpsPerBlock + wib;
    const int nb = S - 1;                     // bins (midpoints) == cdf entries
    float* cdf = sm + (size_t)wib * (2 * nb + P);
    float* bins = cdf + nb;
    float* srt = bins + nb;
    if (ray >= n_rays) return;
    const float* zc = z_coarse ? z_coarse + ray * S : nullptr;
    const float* w = weights + ray * S;
    // pdf over the S-2 interior weights, cdf = [0, cumsum]
    const int nw = S - 2;
    const int per = (nw + 31) / 32;
    float loc[WMAX];
    float lsum = 0.0f;
#pragma unroll
    for (int k = 0; k < WMAX; ++k) {
        const int j = lane * per + k;
        loc[k] = (k < per && j < nw) ? __fadd_rn(w[j + 1], 1e-5f) : 0.0f;
        lsum += loc[k];
    }
    const float total = warp_sum(lsum);
    // cdf = [0, cumsum(pdf)].  ATen's CPU cumsum accumulates float rows in double and rounds each prefix
    // once; scanning in double reproduces those prefixes (the denom < 1e-5 branch below is decided by the
    // last bit of neighbouring cdf entries, so the rounding of the scan matters).
    double lp = 0.0;
#pragma unroll
    for (int k = 0; k < WMAX; ++k) {
        loc[k] = __fdiv_rn(loc[k], total);
        lp += (double)loc[k];
    }
    double incl = lp;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    double run = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) run = 0.0;
#pragma unroll
    for (int k = 0; k < WMAX; ++k) {
        const int j = lane * per + k;
        if (k < per && j < nw) {
            run += (double)loc[k];
            cdf[j + 1] = (float)run;
        }
    }
    if (lane == 0) cdf[0] = 0.0f;
    if (bins_in) {
        for (int j = lane; j < nb; j += 32) bins[j] = bins_in[ray * nb + j];
    } else {
        for (int j = lane; j < nb; j += 32) bins[j] = __fmul_rn(0.5f, __fadd_rn(zc[j + 1], zc[j]));
    }
    if (zc) for (int j = lane; j < S; j += 32) srt[j] = zc[j];
    for (int j = S + Ni + lane; j < P; j += 32) srt[j] = CUDART_INF_F;
    __syncwarp();
    // inverse CDF: four independent binary searches per lane in flight (each step is a dependent shared-memory load)
    float m1 = 0.0f;
    const int n_steps = 32 - __clz(nb);                       // ceil(log2(nb + 1)): enough for a range of nb + 1 outcomes
    constexpr int KS = 4;
    for (int k0 = lane; k0 < Ni; k0 += 32 * KS) {
        float u[KS];
        int lo[KS], hi[KS];
#pragma unroll
        for (int q = 0; q < KS; ++q) {
            const int k = k0 + 32 * q;
            u[q] = k < Ni ? (u_in ? u_in[ray * Ni + k] : linspace01(k, Ni)) : 0.0f;
            lo[q] = 0;
            hi[q] = nb;                                       // first index with cdf[idx] > u  (searchsorted right=True)
        }
        for (int step = 0; step < n_steps; ++step) {
#pragma unroll
            for (int q = 0; q < KS; ++q) {
                if (lo[q] < hi[q]) {
                    const int mid = (lo[q] + hi[q]) >> 1;
                    if (cdf[mid] <= u[q]) lo[q] = mid + 1; else hi[q] = mid;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KS; ++q) {
            const int k = k0 + 32 * q;
            if (k < Ni) {
                const int below = max(lo[q] - 1, 0), above = min(lo[q], nb - 1);
                const float c0 = cdf[below], c1 = cdf[above];
                float den = __fsub_rn(c1, c0);
                if (den < 1e-5f) den = 1.0f;
                const float t = __fdiv_rn(__fsub_rn(u[q], c0), den);
                const float b0 = bins[below], b1 = bins[above];
                const float zs = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
                srt[S + k] = zs;
                if (z_samples_out) z_samples_out[ray * Ni + k] = zs;
                m1 += zs;
            }
        }
    }
    // z_std = std(z_samples, unbiased=False)
    const float mean = warp_sum(m1) / (float)Ni;
    __syncwarp();
    float m2 = 0.0f;
    for (int k = lane; k < Ni; k += 32) {
        const float d = srt[S + k] - mean;
        m2 = fmaf(d, d, m2);
    }
    m2 = warp_sum(m2);
    if (lane == 0 && z_std) z_std[ray] = sqrtf(m2 / (float)Ni);
    if (!z_out) return;
    // sort the staged values (coarse depths, samples) in registers; slots beyond S + Ni hold the largest key
    __syncwarp();
    unsigned key[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) key[k] = (lane * PER + k < S + Ni) ? float_key(srt[lane * PER + k]) : 0xffffffffu;
    warp_bitonic_sort<PER>(key, lane);
    float v[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) v[k] = key_float(key[k]);
    float* out = z_out + ray * (S + Ni);
    if (PER % 4 == 0 && (S + Ni) % 4 == 0) {                 // 16-byte stores: a lane's PER values are contiguous
#pragma unroll
        for (int k = 0; k < PER; k += 4)
            if (lane * PER + k < S + Ni) *reinterpret_cast<float4*>(out + lane * PER + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < PER; ++k)
            if (lane * PER + k < S + Ni) out[lane * PER + k] = v[k];
    }
}

// ---------------------------------------------------------------------------------------------
// ray generation and packing: rows [o(3), d(3), near, far, viewdir(3)]
// ---------------------------------------------------------------------------------------------
struct CamParams {
    float fx, fy, cx, cy;
    float c2w[12];           // 3x4 row-major
    int H, W;
    int ndc;
    float near, far;
    float ndc_ax, ndc_ay;    // -1/(W/(2 focal)), -1/(H/(2 focal)) evaluated in double on the host like the reference's Python scalars
};

__device__ __forceinline__ void finish_ray(float o[3], float d[3], const CamParams& cp, float* __restrict__ out) {
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    const float v0 = __fdiv_rn(d[0], nrm), v1 = __fdiv_rn(d[1], nrm), v2 = __fdiv_rn(d[2], nrm);
    if (cp.ndc) {   // run_nerf_helpers.py:98-115 with near = 1
        const float t = __fdiv_rn(-__fadd_rn(1.0f, o[2]), d[2]);
        o[0] = __fadd_rn(o[0], __fmul_rn(t, d[0]));
        o[1] = __fadd_rn(o[1], __fmul_rn(t, d[1]));
        o[2] = __fadd_rn(o[2], __fmul_rn(t, d[2]));
        const float ax = cp.ndc_ax, ay = cp.ndc_ay;
        const float ox = __fdiv_rn(o[0], o[2]), oy = __fdiv_rn(o[1], o[2]);
        const float n0 = __fdiv_rn(__fmul_rn(ax, o[0]), o[2]), n1 = __fdiv_rn(__fmul_rn(ay, o[1]), o[2]);
        const float n2 = __fadd_rn(1.0f, __fdiv_rn(2.0f, o[2]));
        const float e0 = __fmul_rn(ax, __fsub_rn(__fdiv_rn(d[0], d[2]), ox));
        const float e1 = __fmul_rn(ay, __fsub_rn(__fdiv_rn(d[1], d[2]), oy));
        const float e2 = __fdiv_rn(-2.0f, o[2]);
        o[0] = n0; o[1] = n1; o[2] = n2; d[0] = e0; d[1] = e1; d[2] = e2;
    }
    out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
    out[3] = d[0]; out[4] = d[1]; out[5] = d[2];
    out[6] = cp.near; out[7] = cp.far;
    out[8] = v0; out[9] = v1; out[10] = v2;
}

// rays for pixels [first, first+count) of an H x W image, row-major (run_nerf_helpers.py:71-85)
__global__ void camera_rays_kernel(const CamParams cp, long long first, long long count, float* __restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const long long pix = first + k;
    const int j = (int)(pix / cp.W), i = (int)(pix - (long long)j * cp.W);
    const float dx = __fdiv_rn(__fsub_rn((float)i, cp.cx), cp.fx);
    const float dy = -__fdiv_rn(__fsub_rn((float)j, cp.cy), cp.fy);
    const float dz = -1.0f;
    float o[3], d[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        d[r] = __fadd_rn(__fadd_rn(__fmul_rn(dx, cp.c2w[4 * r + 0]), __fmul_rn(dy, cp.c2w[4 * r + 1])), __fmul_rn(dz, cp.c2w[4 * r + 2]));
        o[r] = cp.c2w[4 * r + 3];
    }
    finish_ray(o, d, cp, out + k * 11);
}

__global__ void pack_rays_kernel(const CamParams cp, const float* __restrict__ rays_o, const float* __restrict__ rays_d, long long n,
                                 float* __restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float o[3] = {rays_o[3 * k], rays_o[3 * k + 1], rays_o[3 * k + 2]};
    float d[3] = {rays_d[3 * k], rays_d[3 * k + 1], rays_d[3 * k + 2]};
    finish_ray(o, d, cp, out + k * 11);
}

// ---------------------------------------------------------------------------------------------
// LSA objective: loss = mean((rgb-t)^2) + mean((rgb0-t)^2); gradients 2(rgb-t)/(3N)
// ---------------------------------------------------------------------------------------------
// Deterministic: every block leaves its two partial sums in the workspace, the block that arrives last adds them up
// in block order and ACCUMULATES into loss2 (one thread), then re-arms the arrival counter.
constexpr int kMseMaxBlocks = 592;
__global__ void mse_grad_kernel(const float* __restrict__ rgb, const float* __restrict__ rgb0, const float* __restrict__ target,
                                long long n3, long long n3_norm, float* __restrict__ d_rgb, float* __restrict__ d_rgb0,
                                float* __restrict__ loss2, float* __restrict__ partial, unsigned int* __restrict__ counter) {
    __shared__ float sh[2][8];
    __shared__ bool is_last;
    float s0 = 0.f, s1 = 0.f;
    const float k = 2.0f / (float)n3_norm;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (long long)gridDim.x * blockDim.x) {
        const float t = target[i];
        const float a = rgb[i] - t;
        s0 = fmaf(a, a, s0);
        d_rgb[i] = k * a;
        if (rgb0) {
            const float b = rgb0[i] - t;
            s1 = fmaf(b, b, s1);
            d_rgb0[i] = k * b;
        }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = s0; sh[1][warp] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float b0 = 0.f, b1 = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { b0 += sh[0][w]; b1 += sh[1][w]; }
        partial[2 * blockIdx.x] = b0;
        partial[2 * blockIdx.x + 1] = b1;
        __threadfence();
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        float t0 = 0.f, t1 = 0.f;
        for (int bk = 0; bk < (int)gridDim.x; ++bk) {
            t0 += __ldcg(partial + 2 * bk);
            t1 += __ldcg(partial + 2 * bk + 1);
        }
        loss2[0] += t0 / (float)n3_norm;
        loss2[1] += t1 / (float)n3_norm;
        *counter = 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// image output: to8b (run_nerf_helpers.py:14), 4 pixels' worth of channels per thread
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned to8b_one(float x) {
    // (255 * np.clip(x, 0, 1)).astype(np.uint8): truncation toward zero; NaN -> 0 (numpy's cast of NaN is unspecified,
    // CUDA's cvt gives 0)
    const float c = fminf(fmaxf(x, 0.0f), 1.0f);
    return (unsigned)__float2int_rz(__fmul_rn(255.0f, c));
}
__global__ void to8b_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, long long n) {
    const long long n4 = n >> 2;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) | (reinterpret_cast<uintptr_t>(out) & 3)) == 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        uint32_t* o4 = reinterpret_cast<uint32_t*>(out);
        for (long long i = tid; i < n4; i += stride) {
            const float4 v = x4[i];
            o4[i] = to8b_one(v.x) | (to8b_one(v.y) << 8) | (to8b_one(v.z) << 16) | (to8b_one(v.w) << 24);
        }
        for (long long i = (n4 << 2) + tid; i < n; i += stride) out[i] = (uint8_t)to8b_one(x[i]);
    } else {
        for (long long i = tid; i < n; i += stride) out[i] = (uint8_t)to8b_one(x[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// training-batch selection on the device (run_nerf.py:690-735): n distinct pixels of one H x W image, their rays generated
// from the pose and their target colours gathered, in one launch.
// The reference draws  np.random.choice(H*W, N_rand, replace=False)  on the host every step; here pixel k of the batch
// is  perm(k)  for a keyed bijection of [0, H*W): a 4-round Feistel network over the next power of four with cycle
// walking -- distinct by construction, a fresh permutation per (seed, step), no host work and no synchronisation.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {        // murmur3 finaliser
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t feistel_perm(uint32_t idx, uint32_t n, int half_bits, uint32_t k0, uint32_t k1) {
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t v = idx;
    do {                                                         // cycle walking: expected < 4 rounds (domain < 4 n)
        uint32_t l = v >> half_bits, r = v & mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            const uint32_t f = mix32(r ^ k0 ^ (0x9e3779b9u * (uint32_t)(round + 1)) ^ (k1 << round)) & mask;
            const uint32_t nl = r;
            r = l ^ f;
            l = nl;
        }
        v = (l << half_bits) | r;
    } while (v >= n);
    return v;
}
__global__ void select_batch_kernel(const CamParams cp, const float* __restrict__ image, uint32_t n_pix, int half_bits, uint32_t k0,
                                    uint32_t k1, long long n, float* __restrict__ rays_out, float* __restrict__ target_out,
                                    int* __restrict__ index_out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t pix = feistel_perm((uint32_t)k, n_pix, half_bits, k0, k1);
    const int j = (int)(pix / (uint32_t)cp.W), i = (int)(pix - (uint32_t)j * (uint32_t)cp.W);
    const float dx = __fdiv_rn(__fsub_rn((float)i, cp.cx), cp.fx);
    const float dy = -__fdiv_rn(__fsub_rn((float)j, cp.cy), cp.fy);
    float o[3], d[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        d[r] = __fadd_rn(__fadd_rn(__fmul_rn(dx, cp.c2w[4 * r + 0]), __fmul_rn(dy, cp.c2w[4 * r + 1])), __fmul_rn(-1.0f, cp.c2w[4 * r + 2]));
        o[r] = cp.c2w[4 * r + 3];
    }
    finish_ray(o, d, cp, rays_out + k * 11);
    if (target_out) {
        target_out[3 * k + 0] = image[3ll * pix + 0];
        target_out[3 * k + 1] = image[3ll * pix + 1];
        target_out[3 * k + 2] = image[3ll * pix + 2];
    }
    if (index_out) index_out[k] = (int)pix;
}

}  // namespace nerfq

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
using namespace nerfq;
static inline int launch_ok() { return cudaGetLastError() == cudaSuccess ? 0 : -3; }

extern "C" int nerfq_coarse_depths(const float* rays, const float* t_rand, long long n_rays, int S, int lindisp, float* z_out,
                                   cudaStream_t stream) {
    if (n_rays == 0) return 0;
    if (!rays || !z_out || S < 2 || n_rays < 0) return -1;
    const long long total = n_rays * S;
    coarse_depths_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(rays, t_rand, n_rays, S, lindisp, z_out);
    return launch_ok();
}

extern "C" int nerfq_composite_fwd(const float* raw, const float* z, const float* rays, const float* noise, int white_bkgd,
                                   long long n_rays, int S, float* rgb, float* disp, float* acc, float* depth, float* weights,
                                   cudaStream_t stream) {
    if (n_rays == 0) return 0;
    if (!raw || !z || !rays || !rgb || !disp || !acc || S < 1 || S > 32 * kMaxPerLane || n_rays < 0) return -1;
    const unsigned grid = (unsigned)((n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock);
    auto launch = [&](auto kernel) { kernel<<<grid, 32 * kWarpsPerBlock, 0, stream>>>(raw, z, rays, noise, white_bkgd, n_rays, S, rgb, disp, acc, depth, weights); };
    switch ((S + 31) / 32) {            // samples per lane
        case 1: launch(composite_fwd_kernel<1>); break;
        case 2: launch(composite_fwd_kernel<2>); break;
        case 3: launch(composite_fwd_kernel<3>); break;
        case 4: launch(composite_fwd_kernel<4>); break;
        case 5: launch(composite_fwd_kernel<5>); break;
        case 6: launch(composite_fwd_kernel<6>); break;
        case 7: launch(composite_fwd_kernel<7>); break;
        default: launch(composite_fwd_kernel<8>); break;
    }
    return launch_ok();
}

extern "C" int nerfq_composite_bwd(const float* raw, const float* z, const float* rays, const float* noise, int white_bkgd,
                                   const float* d_rgb, long long n_rays, int S, float* d_raw, cudaStream_t stream) {
    if (n_rays == 0) return 0;
    if (!raw || !z || !rays || !d_rgb || !d_raw || S < 1 || S > 32 * kMaxPerLane || n_rays < 0) return -1;
    const unsigned grid = (unsigned)((n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock);
    auto launch = [&](auto kernel) { kernel<<<grid, 32 * kWarpsPerBlock, 0, stream>>>(raw, z, rays, noise, white_bkgd, d_rgb, n_rays, S, d_raw); };
    switch ((S + 31) / 32) {            // samples per lane
        case 1: launch(composite_bwd_kernel<1>); break;
        case 2: launch(composite_bwd_kernel<2>); break;
        case 3: launch(composite_bwd_kernel<3>); break;
        case 4: launch(composite_bwd_kernel<4>); break;
        case 5: launch(composite_bwd_kernel<5>); break;
        case 6: launch(composite_bwd_kernel<6>); break;
        case 7: launch(composite_bwd_kernel<7>); break;
        default: launch(composite_bwd_kernel<8>); break;
    }
    return launch_ok();
}

// Either z_coarse [N,S] (bins = its midpoints; z_out receives the sorted union) or bins [N,S-1] (sample_pdf
// proper; z_out must be null).  weights is [N,S]; only the S-2 interior entries are read.
extern "C" int nerfq_sample_fine(const float* z_coarse, const float* bins, const float* weights, const float* u, long long n_rays,
                                 int S, int Ni, float* z_out, float* z_std, float* z_samples, cudaStream_t stream) {
    if (n_rays == 0) return 0;
    if ((!z_coarse && !bins) || !weights || (z_out && !z_coarse) || (!z_out && !z_samples) || S < 3 || S - 2 > 32 * kMaxPerLane ||
        Ni < 1 || n_rays < 0)
        return -1;
    int P = 2;
    while (P < S + Ni) P <<= 1;
    if (P < 32) P = 32;
    if (P > 32 * kMaxPerLane) return -1;
    const size_t smem = (size_t)kWarpsPerBlock * (2 * (S - 1) + P) * sizeof(float);
    if (smem > 200 * 1024) return -1;
    const unsigned grid = (unsigned)((n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock);
    auto launch = [&](auto kernel) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kernel<<<grid, 32 * kWarpsPerBlock, smem, stream>>>(z_coarse, bins, weights, u, n_rays, S, Ni, P, z_out, z_std, z_samples);
    };
    const int wper = (S - 2 + 31) / 32;
    auto pick = [&](auto per_c) {
        constexpr int kPer = decltype(per_c)::value;
        if (wper <= 1) launch(sample_fine_kernel<kPer, 1>);
        else if (wper <= 2) launch(sample_fine_kernel<kPer, 2>);
        else if (wper <= 4) launch(sample_fine_kernel<kPer, 4>);
        else launch(sample_fine_kernel<kPer, 8>);
    };
    switch (P / 32) {
        case 1: pick(std::integral_constant<int, 1>{}); break;
        case 2: pick(std::integral_constant<int, 2>{}); break;
        case 4: pick(std::integral_constant<int, 4>{}); break;
        default: pick(std::integral_constant<int, 8>{}); break;
    }
    return launch_ok();
}

static CamParams make_cam(int H, int W, const float* K4, const float* c2w12, int ndc, float near, float far, float ndc_focal) {
    CamParams cp{};
    if (K4) { cp.fx = K4[0]; cp.fy = K4[1]; cp.cx = K4[2]; cp.cy = K4[3]; }
    if (c2w12) for (int i = 0; i < 12; ++i) cp.c2w[i] = c2w12[i];
    cp.H = H; cp.W = W; cp.ndc = ndc; cp.near = near; cp.far = far;
    if (ndc) {
        cp.ndc_ax = (float)(-1.0 / ((double)W / (2.0 * (double)ndc_focal)));
        cp.ndc_ay = (float)(-1.0 / ((double)H / (2.0 * (double)ndc_focal)));
    }
    return cp;
}

// K4 = {fx, fy, cx, cy} and c2w12 (3x4 row-major) are HOST pointers (tiny, passed by value to the kernel).
extern "C" int nerfq_camera_rays(int H, int W, const float* K4, const float* c2w12, int ndc, float near, float far,
                                 long long first_pixel, long long count, float* rays_out, cudaStream_t stream) {
    if (count == 0) return 0;
    if (!K4 || !c2w12 || !rays_out || H <= 0 || W <= 0 || first_pixel < 0 || count < 0 || first_pixel + count > (long long)H * W) return -1;
    const CamParams cp = make_cam(H, W, K4, c2w12, ndc, near, far, K4[0]);
    camera_rays_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(cp, first_pixel, count, rays_out);
    return launch_ok();
}

extern "C" int nerfq_pack_rays(const float* rays_o, const float* rays_d, long long n, int ndc, int H, int W, float focal, float near,
                               float far, float* rays_out, cudaStream_t stream) {
    if (n == 0) return 0;
    if (!rays_o || !rays_d || !rays_out || n < 0) return -1;
    const CamParams cp = make_cam(H, W, nullptr, nullptr, ndc, near, far, focal);
    pack_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(cp, rays_o, rays_d, n, rays_out);
    return launch_ok();
}

// loss2 is accumulated into ({mse(rgb,t), mse(rgb0,t)} are ADDED): the caller zeroes it.
extern "C" unsigned long long nerfq_mse_grad_workspace_bytes(void) { return 8ull * kMseMaxBlocks + 16; }

extern "C" int nerfq_mse_grad(const float* rgb, const float* rgb0, const float* target, long long n_rays, long long n_norm, float* d_rgb,
                              float* d_rgb0, float* loss2, void* workspace, cudaStream_t stream) {
    if (n_rays == 0) return 0;
    if (!rgb || !target || !d_rgb || !loss2 || !workspace || (rgb0 && !d_rgb0) || n_rays < 0 || n_norm < 0) return -1;
    const long long n3 = n_rays * 3;
    const long long n3_norm = (n_norm > 0 ? n_norm : n_rays) * 3;
    long long blocks = (n3 + 255) / 256;
    if (blocks > kMseMaxBlocks) blocks = kMseMaxBlocks;
    float* partial = reinterpret_cast<float*>(workspace);
    unsigned int* counter = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(workspace) + 8ull * kMseMaxBlocks);
    mse_grad_kernel<<<(unsigned)blocks, 256, 0, stream>>>(rgb, rgb0, target, n3, n3_norm, d_rgb, d_rgb0, loss2, partial, counter);
    return launch_ok();
}

extern "C" int nerfq_to8b(const float* x, uint8_t* out, long long n, cudaStream_t stream) {
    if (n == 0) return 0;
    if (!x || !out || n < 0) return -1;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    to8b_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, out, n);
    return launch_ok();
}

// K4 / c2w12 are HOST pointers; image [H*W,3] (nullable together with target_out), rays_out [n,11], target_out [n,3],
// index_out [n] (nullable): the pixel indices chosen.  n <= H*W.
extern "C" int nerfq_select_batch(int H, int W, const float* K4, const float* c2w12, int ndc, float near, float far, const float* image,
                                  unsigned long long seed, unsigned long long step, long long n, float* rays_out, float* target_out,
                                  int* index_out, cudaStream_t stream) {
    if (n == 0) return 0;
    if (!K4 || !c2w12 || !rays_out || H <= 0 || W <= 0 || n < 0 || n > (long long)H * W || (long long)H * W >= (1ll << 31) ||
        (target_out && !image))
        return -1;
    const CamParams cp = make_cam(H, W, K4, c2w12, ndc, near, far, K4[0]);
    const uint32_t n_pix = (uint32_t)((long long)H * W);
    int half_bits = 1;
    while ((1ull << (2 * half_bits)) < n_pix) ++half_bits;
    // key schedule on the host: splitmix64 of (seed, step)
    unsigned long long zk = seed * 0x9e3779b97f4a7c15ull + step + 0x632be59bd9b4e019ull;
    zk = (zk ^ (zk >> 30)) * 0xbf58476d1ce4e5b9ull;
    zk = (zk ^ (zk >> 27)) * 0x94d049bb133111ebull;
    zk ^= zk >> 31;
    select_batch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(cp, image, n_pix, half_bits, (uint32_t)zk, (uint32_t)(zk >> 32), n,
                                                                          rays_out, target_out, index_out);
    return launch_ok();
}

// ---------------------------------------------------------------------------------------------
// render_rays, forward only, as ONE call (run_nerf.py:348-457 with perturb = 0, raw_noise_std = 0: the test-view path):
// coarse depths -> MLP(coarse) -> composite -> sample_pdf + merge -> MLP(fine) -> composite.
// ---------------------------------------------------------------------------------------------
extern "C" int nerfq_mlp_forward(const void* packed, const float* rays, const float* z, long long n_rays, int samples_per_ray, float* raw,
                                  void* save, int max_ctas, cudaStream_t stream);

extern "C" unsigned long long nerfq_render_rays_workspace_bytes(long long n_rays, int S, int Ni) {
    if (n_rays <= 0 || S <= 0 || Ni < 0) return 0;
    // z0 [n,S] | raw0 [n,S,4] | w0 [n,S] | z1 [n,S+Ni] | raw1 [n,S+Ni,4], each rounded up to 256 bytes
    auto up = [](unsigned long long b) { return (b + 255ull) / 256ull * 256ull; };
    const unsigned long long n = (unsigned long long)n_rays;
    return up(4 * n * S) + up(16 * n * S) + up(4 * n * S) + up(4 * n * (S + Ni)) + up(16 * n * (S + Ni));
}

extern "C" int nerfq_render_rays_fwd(const void* packed_coarse, const void* packed_fine, const float* rays, long long n_rays, int S, int Ni,
                                      int lindisp, int white_bkgd, void* workspace, float* rgb, float* disp, float* acc, float* rgb0,
                                      float* disp0, float* acc0, float* z_std, int max_ctas, cudaStream_t stream) {
    if (n_rays == 0) return 0;
    if (!packed_coarse || !rays || !workspace || !rgb || !disp || !acc || n_rays < 0 || S < 3 || Ni < 0) return -1;
    if (Ni > 0 && (!rgb0 || !disp0 || !acc0 || S + Ni > 32 * kMaxPerLane)) return -1;
    auto up = [](unsigned long long b) { return (b + 255ull) / 256ull * 256ull; };
    const unsigned long long n = (unsigned long long)n_rays;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    float* z0 = reinterpret_cast<float*>(ws); ws += up(4 * n * S);
    float* raw0 = reinterpret_cast<float*>(ws); ws += up(16 * n * S);
    float* w0 = reinterpret_cast<float*>(ws); ws += up(4 * n * S);
    float* z1 = reinterpret_cast<float*>(ws); ws += up(4 * n * (S + Ni));
    float* raw1 = reinterpret_cast<float*>(ws);
    int rc = nerfq_coarse_depths(rays, nullptr, n_rays, S, lindisp, z0, stream);
    if (rc) return rc;
    rc = nerfq_mlp_forward(packed_coarse, rays, z0, n_rays, S, raw0, nullptr, max_ctas, stream);
    if (rc) return rc;
    if (Ni == 0) return nerfq_composite_fwd(raw0, z0, rays, nullptr, white_bkgd, n_rays, S, rgb, disp, acc, nullptr, nullptr, stream);
    rc = nerfq_composite_fwd(raw0, z0, rays, nullptr, white_bkgd, n_rays, S, rgb0, disp0, acc0, nullptr, w0, stream);
    if (rc) return rc;
    rc = nerfq_sample_fine(z0, nullptr, w0, nullptr, n_rays, S, Ni, z1, z_std, nullptr, stream);
    if (rc) return rc;
    rc = nerfq_mlp_forward(packed_fine ? packed_fine : packed_coarse, rays, z1, n_rays, S + Ni, raw1, nullptr, max_ctas, stream);
    if (rc) return rc;
    return nerfq_composite_fwd(raw1, z1, rays, nullptr, white_bkgd, n_rays, S + Ni, rgb, disp, acc, nullptr, nullptr, stream);
}
