// Backward of the row-wise glue of Spiral_MambaBlock.forward (training path:
// reference block/mamba_block.py:100-115 differentiated by autograd in train.py:259).  Three kernels, the adjoints of
// dm_block.cu's forward kernels, so that the training step runs the same fused row kernels as inference instead of
// ~160 eager elementwise / reduction launches per block:
//
//   pre_bwd       d[x_ssm ; x_ssm*w] -> d x (LayerNorm backward through the adaLN modulate), d shift / d scale per batch
//                 element, d gamma / d beta of norm1
//   post_mix_bwd  d x_out -> d a, d b (through the sigmoid mix), d hidden (through w3 . silu(hidden) -> sigmoid),
//                 d gate per batch element, d w3, d b3
//   post_ln_bwd   d LN(cat(a, b)) -> d a, d b (added to post_mix_bwd's), d gamma / d beta of attention_network[0]
//
// One CTA per (batch element, row slice), 8 warps, one warp per token row at a time (16-byte vectors, the row's values in
// registers, warp-shuffle row reductions).  Column sums (per-batch and per-parameter gradients) are accumulated per lane
// in registers over the warp's rows, combined across the CTA's warps with shared-memory atomics and flushed with one
// 16-byte global atomic per 4 columns.  All arithmetic fp32.
#include "dm_common.cuh"

namespace dm {
namespace {

constexpr int kBwdWarps = 8;

template <typename T> struct V8b;
template <> struct V8b<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct V8b<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Column accumulators of a CTA: NQ quantities x NC columns.  A lane owns columns (i*32 + lane)*8 + e, i < NC/256; they
// are kept as red[q][i*8 + e][lane] so that the lanes of a warp hit 32 different banks.
template <int NQ, int NC>
struct ColRed {
    float v[NQ][NC / 32][32];
    __device__ void zero() {
        float* p = &v[0][0][0];
        for (int i = threadIdx.x; i < NQ * NC; i += blockDim.x) p[i] = 0.f;
    }
    // column c of quantity q
    __device__ float get(int q, int c) const { return v[q][(c / 256) * 8 + (c & 7)][(c & 255) >> 3]; }
};

__device__ __forceinline__ void atomic_add4(float* dst, float a, float b, float c, float d) {
    atomicAdd(reinterpret_cast<float4*>(dst), make_float4(a, b, c, d));          // one 16-byte red.global (sm_90+)
}

// rows of batch element b handled by this CTA: [l_begin, l_end)
__device__ __forceinline__ void row_slice(int L, int& l_begin, int& l_end) {
    const int per = (L + gridDim.y - 1) / gridDim.y;
    l_begin = blockIdx.y * per;
    l_end = min(L, l_begin + per);
}

// ------------------------------------------------------------------------------------------------------
// pre_bwd: adjoint of spiral_pre_kernel.  g2 = d out2 (2, rows, D); forward: xs = x + skip, xh = (xs - mean) rstd,
// n = xh gamma + beta, o1 = n (1 + scale_b) + shift_b, o2 = o1 w_row.
// ------------------------------------------------------------------------------------------------------
template <typename T, int NV>
__global__ void __launch_bounds__(kBwdWarps * 32)
spiral_pre_bwd_kernel(const float* __restrict__ x, const float* __restrict__ skip, const float* __restrict__ ln_w,
                      const float* __restrict__ ln_b, const float* __restrict__ mod, int64_t mod_stride,
                      const float* __restrict__ w, const T* __restrict__ g2, float* __restrict__ dx,
                      float* __restrict__ d_mod, int64_t d_mod_stride, float* __restrict__ d_ln_w,
                      float* __restrict__ d_ln_b, int rows, int L, float eps) {
    constexpr int D = NV * 256;
    __shared__ ColRed<4, D> red;                  // 0 d shift, 1 d scale, 2 d gamma, 3 d beta
    red.zero();
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    int l0, l1;
    row_slice(L, l0, l1);
    const float* shift = mod + static_cast<int64_t>(b) * mod_stride;
    const float* scale = shift + D;
    float gam[NV][8], bet[NV][8], sc1[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        V8b<float>::load(ln_w + c, gam[i]);
        V8b<float>::load(ln_b + c, bet[i]);
        V8b<float>::load(scale + c, sc1[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) sc1[i][e] += 1.0f;
    }
    float a_sh[NV][8] = {}, a_sc[NV][8] = {}, a_g[NV][8] = {}, a_b[NV][8] = {};
    for (int l = l0 + warp; l < l1; l += kBwdWarps) {
        const int row = b * L + l;
        float v[NV][8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int64_t off = static_cast<int64_t>(row) * D + (i * 32 + lane) * 8;
            V8b<float>::load(x + off, v[i]);
            if (skip) {
                float t[8];
                V8b<float>::load(skip + off, t);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[i][e] += t[e];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) s += v[i][e];
        }
        const float mean = wsum(s) * (1.0f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                v[i][e] -= mean;
                q = fmaf(v[i][e], v[i][e], q);
            }
        const float rstd = rsqrtf(wsum(q) * (1.0f / D) + eps);
        const float wr = w ? __ldg(w + row) : 1.0f;          // forward writes out2[1] = o1 * 1 when there is no mask
        float dxh[NV][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int64_t off = static_cast<int64_t>(row) * D + (i * 32 + lane) * 8;
            float ga[8], gb[8];
            V8b<T>::load(g2 + off, ga);
            V8b<T>::load(g2 + static_cast<int64_t>(rows) * D + off, gb);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float xh = v[i][e] * rstd;
                v[i][e] = xh;
                const float n = fmaf(xh, gam[i][e], bet[i][e]);
                const float go = fmaf(wr, gb[e], ga[e]);
                a_sh[i][e] += go;
                a_sc[i][e] = fmaf(go, n, a_sc[i][e]);
                const float dn = go * sc1[i][e];
                a_g[i][e] = fmaf(dn, xh, a_g[i][e]);
                a_b[i][e] += dn;
                const float d = dn * gam[i][e];
                dxh[i][e] = d;
                s1 += d;
                s2 = fmaf(d, xh, s2);
            }
        }
        s1 = wsum(s1) * (1.0f / D);
        s2 = wsum(s2) * (1.0f / D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = rstd * (dxh[i][e] - s1 - v[i][e] * s2);
            V8b<float>::store(dx + static_cast<int64_t>(row) * D + (i * 32 + lane) * 8, o);
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            atomicAdd(&red.v[0][i * 8 + e][lane], a_sh[i][e]);
            atomicAdd(&red.v[1][i * 8 + e][lane], a_sc[i][e]);
            atomicAdd(&red.v[2][i * 8 + e][lane], a_g[i][e]);
            atomicAdd(&red.v[3][i * 8 + e][lane], a_b[i][e]);
        }
    __syncthreads();
    float* dsh = d_mod + static_cast<int64_t>(b) * d_mod_stride;
    for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
        atomic_add4(dsh + c, red.get(0, c), red.get(0, c + 1), red.get(0, c + 2), red.get(0, c + 3));
        atomic_add4(dsh + D + c, red.get(1, c), red.get(1, c + 1), red.get(1, c + 2), red.get(1, c + 3));
        atomic_add4(d_ln_w + c, red.get(2, c), red.get(2, c + 1), red.get(2, c + 2), red.get(2, c + 3));
        atomic_add4(d_ln_b + c, red.get(3, c), red.get(3, c + 1), red.get(3, c + 2), red.get(3, c + 3));
    }
}

// ------------------------------------------------------------------------------------------------------
// post_mix_bwd: adjoint of spiral_post_mix_kernel.  forward: s = w3 . silu(hidden) + b3, alpha = sigmoid(s),
// mixed = alpha a + (1 - alpha) b, x_out = (x + skip) + gate_b mixed.  (d(x + skip) = d x_out: the caller aliases it.)
// ------------------------------------------------------------------------------------------------------
template <typename T, int NV>
__global__ void __launch_bounds__(kBwdWarps * 32)
spiral_post_mix_bwd_kernel(const float* __restrict__ dxo, const T* __restrict__ ab, const T* __restrict__ hidden,
                           const float* __restrict__ w3, const float* __restrict__ b3, const float* __restrict__ mod,
                           int64_t mod_stride, T* __restrict__ d_ab, T* __restrict__ d_hidden, float* __restrict__ d_mod,
                           int64_t d_mod_stride, float* __restrict__ d_w3, float* __restrict__ d_b3, int rows, int L) {
    constexpr int D = NV * 256;
    __shared__ ColRed<2, D> red;                  // 0 d gate, 1 d w3
    __shared__ float red_b3;
    red.zero();
    if (threadIdx.x == 0) red_b3 = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    int l0, l1;
    row_slice(L, l0, l1);
    const float* gate = mod + static_cast<int64_t>(b) * mod_stride + 2 * D;
    float gt[NV][8], w3v[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        V8b<float>::load(gate + (i * 32 + lane) * 8, gt[i]);
        V8b<float>::load(w3 + (i * 32 + lane) * 8, w3v[i]);
    }
    const float bias3 = __ldg(b3);
    float a_gate[NV][8] = {}, a_w3[NV][8] = {};
    float a_b3 = 0.f;
    for (int l = l0 + warp; l < l1; l += kBwdWarps) {
        const int row = b * L + l;
        float hv[NV][8], sg[NV][8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            V8b<T>::load(hidden + static_cast<int64_t>(row) * D + (i * 32 + lane) * 8, hv[i]);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                sg[i][e] = sigmoid_fast(hv[i][e]);
                s = fmaf(hv[i][e] * sg[i][e], w3v[i][e], s);
            }
        }
        const float alpha = sigmoid_fast(wsum(s) + bias3);
        float dal = 0.f;
        float dm[NV][8];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int64_t off = static_cast<int64_t>(row) * D + (i * 32 + lane) * 8;
            float g[8], av[8], bv[8], oa[8], ob[8];
            V8b<float>::load(dxo + off, g);
            V8b<T>::load(ab + off, av);
            V8b<T>::load(ab + static_cast<int64_t>(rows) * D + off, bv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float diff = av[e] - bv[e];
                const float mixed = fmaf(alpha, diff, bv[e]);
                a_gate[i][e] = fmaf(g[e], mixed, a_gate[i][e]);
                const float d = gt[i][e] * g[e];
                dm[i][e] = d;
                dal = fmaf(d, diff, dal);
                oa[e] = alpha * d;
                ob[e] = d - oa[e];
            }
            V8b<T>::store(d_ab + off, oa);
            V8b<T>::store(d_ab + static_cast<int64_t>(rows) * D + off, ob);
        }
        const float ds = wsum(dal) * alpha * (1.0f - alpha);
        if (lane == 0) a_b3 += ds;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float h = hv[i][e], sgm = sg[i][e];
                a_w3[i][e] = fmaf(ds, h * sgm, a_w3[i][e]);
                o[e] = ds * w3v[i][e] * sgm * fmaf(h, 1.0f - sgm, 1.0f);          // silu'(h) = s (1 + h (1 - s))
            }
            V8b<T>::store(d_hidden + static_cast<int64_t>(row) * D + (i * 32 + lane) * 8, o);
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            atomicAdd(&red.v[0][i * 8 + e][lane], a_gate[i][e]);
            atomicAdd(&red.v[1][i * 8 + e][lane], a_w3[i][e]);
        }
    if (lane == 0) atomicAdd(&red_b3, a_b3);
    __syncthreads();
    float* dgate = d_mod + static_cast<int64_t>(b) * d_mod_stride + 2 * D;
    for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
        atomic_add4(dgate + c, red.get(0, c), red.get(0, c + 1), red.get(0, c + 2), red.get(0, c + 3));
        atomic_add4(d_w3 + c, red.get(1, c), red.get(1, c + 1), red.get(1, c + 2), red.get(1, c + 3));
    }
    if (threadIdx.x == 0) atomicAdd(d_b3, red_b3);
}

// ------------------------------------------------------------------------------------------------------
// post_ln_bwd: adjoint of spiral_post_ln_kernel (LayerNorm over cat(a, b), 2D columns); the result is ADDED to d_ab
// (which already holds post_mix_bwd's d a / d b).
// ------------------------------------------------------------------------------------------------------
template <typename T, int NV>
__global__ void __launch_bounds__(kBwdWarps * 32)
spiral_post_ln_bwd_kernel(const T* __restrict__ ab, const float* __restrict__ ln_w, const T* __restrict__ d_out,
                          T* __restrict__ d_ab, float* __restrict__ d_ln_w, float* __restrict__ d_ln_b, int rows, int L,
                          float eps) {
    constexpr int D = NV * 256;
    __shared__ ColRed<2, 2 * D> red;              // 0 d gamma, 1 d beta over the 2D columns
    red.zero();
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    int l0, l1;
    row_slice(L, l0, l1);
    float gam[2][NV][8];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < NV; ++i) V8b<float>::load(ln_w + h * D + (i * 32 + lane) * 8, gam[h][i]);
    float a_g[2][NV][8] = {}, a_b[2][NV][8] = {};
    for (int l = l0 + warp; l < l1; l += kBwdWarps) {
        const int row = b * L + l;
        float v[2][NV][8];
        float s = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                V8b<T>::load(ab + (static_cast<int64_t>(h) * rows + row) * D + (i * 32 + lane) * 8, v[h][i]);
#pragma unroll
                for (int e = 0; e < 8; ++e) s += v[h][i][e];
            }
        const float mean = wsum(s) * (0.5f / D);
        float q = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    v[h][i][e] -= mean;
                    q = fmaf(v[h][i][e], v[h][i][e], q);
                }
        const float rstd = rsqrtf(wsum(q) * (0.5f / D) + eps);
        float dxh[2][NV][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                float g[8];
                V8b<T>::load(d_out + static_cast<int64_t>(row) * 2 * D + h * D + (i * 32 + lane) * 8, g);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xh = v[h][i][e] * rstd;
                    v[h][i][e] = xh;
                    a_g[h][i][e] = fmaf(g[e], xh, a_g[h][i][e]);
                    a_b[h][i][e] += g[e];
                    const float d = g[e] * gam[h][i][e];
                    dxh[h][i][e] = d;
                    s1 += d;
                    s2 = fmaf(d, xh, s2);
                }
            }
        s1 = wsum(s1) * (0.5f / D);
        s2 = wsum(s2) * (0.5f / D);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                T* dst = d_ab + (static_cast<int64_t>(h) * rows + row) * D + (i * 32 + lane) * 8;
                float o[8];
                V8b<T>::load(dst, o);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] += rstd * (dxh[h][i][e] - s1 - v[h][i][e] * s2);
                V8b<T>::store(dst, o);
            }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                atomicAdd(&red.v[0][(h * NV + i) * 8 + e][lane], a_g[h][i][e]);
                atomicAdd(&red.v[1][(h * NV + i) * 8 + e][lane], a_b[h][i][e]);
            }
    __syncthreads();
    for (int c = threadIdx.x * 4; c < 2 * D; c += blockDim.x * 4) {
        atomic_add4(d_ln_w + c, red.get(0, c), red.get(0, c + 1), red.get(0, c + 2), red.get(0, c + 3));
        atomic_add4(d_ln_b + c, red.get(1, c), red.get(1, c + 1), red.get(1, c + 2), red.get(1, c + 3));
    }
}

// row slices per batch element: enough CTAs to cover the machine, at least ~one row per warp
inline dim3 bwd_grid(int batch, int L) {
    int s = (160 + batch - 1) / batch;
    const int max_s = (L + kBwdWarps - 1) / kBwdWarps;
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    return dim3(static_cast<unsigned>(batch), static_cast<unsigned>(s), 1);
}

}  // namespace
}  // namespace dm

using namespace dm;

extern "C" int dm_spiral_pre_bwd(const float* x, const float* skip, const float* ln_weight, const float* ln_bias,
                                 const float* mod, int64_t mod_batch_stride, const float* w, const void* d_out2, float* dx,
                                 float* d_mod, int64_t d_mod_batch_stride, float* d_ln_weight, float* d_ln_bias,
                                 int32_t batch, int32_t seqlen, int32_t d_model, float eps, int32_t act_dtype, void* stream) {
    if (!x || !ln_weight || !ln_bias || !mod || !d_out2 || !dx || !d_mod || !d_ln_weight || !d_ln_bias || batch <= 0 ||
        seqlen <= 0)
        return DM_ERR_INVALID_ARG;
    if (d_model != 512) return DM_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(d_out2) || !aligned16(dx) || !aligned16(mod) || !aligned16(d_mod) ||
        !aligned16(d_ln_weight) || !aligned16(d_ln_bias) || (skip && !aligned16(skip)) || (mod_batch_stride % 4) ||
        (d_mod_batch_stride % 4))
        return DM_ERR_INVALID_ARG;
    const int rows = batch * seqlen;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid = bwd_grid(batch, seqlen);
    if (act_dtype == DM_BF16)
        spiral_pre_bwd_kernel<__nv_bfloat16, 2><<<grid, kBwdWarps * 32, 0, st>>>(
            x, skip, ln_weight, ln_bias, mod, mod_batch_stride, w, static_cast<const __nv_bfloat16*>(d_out2), dx, d_mod,
            d_mod_batch_stride, d_ln_weight, d_ln_bias, rows, seqlen, eps);
    else if (act_dtype == DM_F32)
        spiral_pre_bwd_kernel<float, 2><<<grid, kBwdWarps * 32, 0, st>>>(
            x, skip, ln_weight, ln_bias, mod, mod_batch_stride, w, static_cast<const float*>(d_out2), dx, d_mod,
            d_mod_batch_stride, d_ln_weight, d_ln_bias, rows, seqlen, eps);
    else
        return DM_ERR_UNSUPPORTED;
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

extern "C" int dm_spiral_post_mix_bwd(const float* d_x_out, const void* ab, const void* hidden, const float* w3,
                                      const float* b3, const float* mod, int64_t mod_batch_stride, void* d_ab, void* d_hidden,
                                      float* d_mod, int64_t d_mod_batch_stride, float* d_w3, float* d_b3, int32_t batch,
                                      int32_t seqlen, int32_t d_model, int32_t act_dtype, void* stream) {
    if (!d_x_out || !ab || !hidden || !w3 || !b3 || !mod || !d_ab || !d_hidden || !d_mod || !d_w3 || !d_b3 || batch <= 0 ||
        seqlen <= 0)
        return DM_ERR_INVALID_ARG;
    if (d_model != 512) return DM_ERR_UNSUPPORTED;
    if (!aligned16(d_x_out) || !aligned16(ab) || !aligned16(hidden) || !aligned16(d_ab) || !aligned16(d_hidden) ||
        !aligned16(mod) || !aligned16(d_mod) || !aligned16(w3) || !aligned16(d_w3) || (mod_batch_stride % 4) ||
        (d_mod_batch_stride % 4))
        return DM_ERR_INVALID_ARG;
    const int rows = batch * seqlen;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid = bwd_grid(batch, seqlen);
    if (act_dtype == DM_BF16)
        spiral_post_mix_bwd_kernel<__nv_bfloat16, 2><<<grid, kBwdWarps * 32, 0, st>>>(
            d_x_out, static_cast<const __nv_bfloat16*>(ab), static_cast<const __nv_bfloat16*>(hidden), w3, b3, mod,
            mod_batch_stride, static_cast<__nv_bfloat16*>(d_ab), static_cast<__nv_bfloat16*>(d_hidden), d_mod,
            d_mod_batch_stride, d_w3, d_b3, rows, seqlen);
    else if (act_dtype == DM_F32)
        spiral_post_mix_bwd_kernel<float, 2><<<grid, kBwdWarps * 32, 0, st>>>(
            d_x_out, static_cast<const float*>(ab), static_cast<const float*>(hidden), w3, b3, mod, mod_batch_stride,
            static_cast<float*>(d_ab), static_cast<float*>(d_hidden), d_mod, d_mod_batch_stride, d_w3, d_b3, rows, seqlen);
    else
        return DM_ERR_UNSUPPORTED;
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

extern "C" int dm_spiral_post_ln_bwd(const void* ab, const float* ln_weight, const void* d_out, void* d_ab,
                                     float* d_ln_weight, float* d_ln_bias, int32_t batch, int32_t seqlen, int32_t d_model,
                                     float eps, int32_t act_dtype, void* stream) {
    if (!ab || !ln_weight || !d_out || !d_ab || !d_ln_weight || !d_ln_bias || batch <= 0 || seqlen <= 0)
        return DM_ERR_INVALID_ARG;
    if (d_model != 512) return DM_ERR_UNSUPPORTED;
    if (!aligned16(ab) || !aligned16(d_out) || !aligned16(d_ab) || !aligned16(ln_weight) || !aligned16(d_ln_weight) ||
        !aligned16(d_ln_bias))
        return DM_ERR_INVALID_ARG;
    const int rows = batch * seqlen;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid = bwd_grid(batch, seqlen);
    if (act_dtype == DM_BF16)
        spiral_post_ln_bwd_kernel<__nv_bfloat16, 2><<<grid, kBwdWarps * 32, 0, st>>>(
            static_cast<const __nv_bfloat16*>(ab), ln_weight, static_cast<const __nv_bfloat16*>(d_out),
            static_cast<__nv_bfloat16*>(d_ab), d_ln_weight, d_ln_bias, rows, seqlen, eps);
    else if (act_dtype == DM_F32)
        spiral_post_ln_bwd_kernel<float, 2><<<grid, kBwdWarps * 32, 0, st>>>(
            static_cast<const float*>(ab), ln_weight, static_cast<const float*>(d_out), static_cast<float*>(d_ab),
            d_ln_weight, d_ln_bias, rows, seqlen, eps);
    else
        return DM_ERR_UNSUPPORTED;
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

// ------------------------------------------------------------------------------------------------------
// Adjoint of the CrossScan gather (reference block/mamba.py:48-57: CrossScan.backward un-permutes and sums the
// per-direction gradients): dst[r][l][:] = sum_k src[r][idx[l*K + k]][:], src fp32 rows in scan order (K*L rows per
// sequence group r), dst in the activation dtype, token order.  One CTA per output row; replaces index_select + sum +
// cast (three passes over a 77 MB fp32 tensor at the C4 shape) by one.
// ------------------------------------------------------------------------------------------------------
namespace dm {
namespace {
template <typename T>
__global__ void __launch_bounds__(256)
merge_directions_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, T* __restrict__ dst, int Lsrc, int K,
                        int rows_per_group, int C) {
    const int l = blockIdx.x, r = blockIdx.y;
    const float* base = src + static_cast<int64_t>(r) * rows_per_group * C;
    T* out = dst + (static_cast<int64_t>(r) * Lsrc + l) * C;
    for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < K; ++k) {
            const int j = __ldg(idx + l * K + k);
            float v[8];
            V8b<float>::load(base + static_cast<int64_t>(j) * C + c, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += v[e];
        }
        V8b<T>::store(out + c, acc);
    }
}
}  // namespace
}  // namespace dm

extern "C" int dm_merge_directions(const float* src, const int32_t* index, void* dst, int32_t n_groups, int32_t src_len,
                                   int32_t n_dir, int32_t rows_per_group, int32_t channels, int32_t act_dtype, void* stream) {
    if (!src || !index || !dst || n_groups <= 0 || src_len <= 0 || n_dir <= 0 || rows_per_group <= 0 || channels <= 0)
        return DM_ERR_INVALID_ARG;
    if (channels % 8 || !aligned16(src) || !aligned16(dst)) return DM_ERR_INVALID_ARG;
    if (n_groups > 65535) return DM_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid(static_cast<unsigned>(src_len), static_cast<unsigned>(n_groups), 1);
    if (act_dtype == DM_BF16)
        merge_directions_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, index, static_cast<__nv_bfloat16*>(dst), src_len,
                                                                     n_dir, rows_per_group, channels);
    else if (act_dtype == DM_F32)
        merge_directions_kernel<float><<<grid, 256, 0, st>>>(src, index, static_cast<float*>(dst), src_len, n_dir,
                                                             rows_per_group, channels);
    else
        return DM_ERR_UNSUPPORTED;
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

namespace dm {
namespace {
struct MergeSegs { const float* src[4]; int channels[4]; int row_stride[4]; int n; int total; };
template <typename T>
__global__ void __launch_bounds__(256)
merge_directions_multi_kernel(const MergeSegs sg, const int32_t* __restrict__ idx, T* __restrict__ dst, int Lsrc, int K,
                              int rows_per_group) {
    const int l = blockIdx.x, r = blockIdx.y;
    T* out = dst + (static_cast<int64_t>(r) * Lsrc + l) * sg.total;
    int col0 = 0;
    for (int s = 0; s < sg.n; ++s) {
        const float* base = sg.src[s] + static_cast<int64_t>(r) * rows_per_group * sg.row_stride[s];
        for (int c = threadIdx.x * 8; c < sg.channels[s]; c += blockDim.x * 8) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int k = 0; k < K; ++k) {
                const int j = __ldg(idx + l * K + k);
                float v[8];
                V8b<float>::load(base + static_cast<int64_t>(j) * sg.row_stride[s] + c, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] += v[e];
            }
            V8b<T>::store(out + col0 + c, acc);
        }
        col0 += sg.channels[s];
    }
}
}  // namespace
}  // namespace dm

extern "C" int dm_merge_directions_multi(const dm_merge_segment* segments, int32_t n_segments, const int32_t* index, void* dst,
                                         int32_t n_groups, int32_t src_len, int32_t n_dir, int32_t rows_per_group,
                                         int32_t act_dtype, void* stream) {
    if (!segments || n_segments <= 0 || n_segments > 4 || !index || !dst || n_groups <= 0 || src_len <= 0 || n_dir <= 0 ||
        rows_per_group <= 0)
        return DM_ERR_INVALID_ARG;
    if (n_groups > 65535) return DM_ERR_UNSUPPORTED;
    MergeSegs sg{};
    sg.n = n_segments;
    for (int s = 0; s < n_segments; ++s) {
        const dm_merge_segment& m = segments[s];
        if (!m.src || m.channels <= 0 || m.channels % 8 || m.row_stride % 4 || !aligned16(m.src)) return DM_ERR_INVALID_ARG;
        sg.src[s] = m.src; sg.channels[s] = m.channels; sg.row_stride[s] = m.row_stride;
        sg.total += m.channels;
    }
    if (!aligned16(dst)) return DM_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid(static_cast<unsigned>(src_len), static_cast<unsigned>(n_groups), 1);
    if (act_dtype == DM_BF16)
        merge_directions_multi_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(sg, index, static_cast<__nv_bfloat16*>(dst), src_len,
                                                                           n_dir, rows_per_group);
    else if (act_dtype == DM_F32)
        merge_directions_multi_kernel<float><<<grid, 256, 0, st>>>(sg, index, static_cast<float*>(dst), src_len, n_dir,
                                                                   rows_per_group);
    else
        return DM_ERR_UNSUPPORTED;
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}
