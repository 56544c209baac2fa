// Row-wise glue of Spiral_MambaBlock.forward (reference block/mamba_block.py:100-115)
// fused into three HBM-bound kernels, one warp per token row, everything vectorised 16 B:
//
//   pre      x (+ long-skip) -> LayerNorm -> adaLN modulate -> [x_ssm ; x_ssm * w] in the act dtype, laid out as the
//            (2, rows, D) operand of the batched in-projection            (reference lines 101-105 + model.py:290-292)
//   post_ln  LayerNorm(cat(a, b)) -> act dtype, the operand of attention_network's first Linear   (line 110-111)
//   post_mix alpha = sigmoid(w3 . silu(hidden) + b3) ; x_new = (x + skip) + gate * (alpha a + (1-alpha) b)
//                                                                              (lines 111-114)
// The reference runs ~35 elementwise / reduction launches for the same work.
#include <type_traits>

#include "dm_common.cuh"

namespace dm {
namespace {

constexpr int kRowWarps = 4;      // rows (warps) per CTA

template <typename T> struct V8;  // 8 consecutive elements <-> float[8]
template <> struct V8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct V8<__nv_bfloat16> {
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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// D = 8 * 32 * NV elements per row; lane owns NV groups of 8 consecutive elements: element (i*32 + lane)*8 + e
template <typename T, int NV>
__global__ void __launch_bounds__(kRowWarps * 32)
spiral_pre_kernel(const float* __restrict__ x, const float* __restrict__ skip, const float* __restrict__ ln_w,
                  const float* __restrict__ ln_b, const float* __restrict__ mod, int64_t mod_stride,
                  const float* __restrict__ w, T* __restrict__ out2, int rows, int L, float eps) {
    constexpr int D = NV * 256;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    pdl_wait();                                 // (PDL: the predecessor's writes are visible from here on)
    const int b = row / L;
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int64_t off = static_cast<int64_t>(row) * D + (i * 32 + lane) * 8;
        V8<float>::load(x + off, v[i]);
        if (skip) {
            float t[8];
            V8<float>::load(skip + off, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[i][e] += t[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[i][e];
    }
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float d = v[i][e] - mean;
            q = fmaf(d, d, q);
        }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    const float wr = w ? __ldg(w + row) : 1.0f;
    const float* shift = mod + static_cast<int64_t>(b) * mod_stride;
    const float* scale = shift + D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        float g[8], bb[8], sh[8], sc[8], o1[8], o2[8];
        V8<float>::load(ln_w + c, g);
        V8<float>::load(ln_b + c, bb);
        V8<float>::load(shift + c, sh);
        V8<float>::load(scale + c, sc);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float n = fmaf((v[i][e] - mean) * rstd, g[e], bb[e]);
            o1[e] = fmaf(n, 1.0f + sc[e], sh[e]);
            o2[e] = o1[e] * wr;
        }
        V8<T>::store(out2 + static_cast<int64_t>(row) * D + c, o1);
        V8<T>::store(out2 + (static_cast<int64_t>(rows) + row) * D + c, o2);
    }
}

// LayerNorm over cat(a, b): row of 2*D values, a = ab[0], b = ab[1]
template <typename T, int NV>
__global__ void __launch_bounds__(kRowWarps * 32)
spiral_post_ln_kernel(const T* __restrict__ ab, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                      T* __restrict__ out, int rows, float eps) {
    constexpr int D = NV * 256;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    pdl_wait();                                 // (PDL: the predecessor's writes are visible from here on)
    float v[2][NV][8];
    float s = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            V8<T>::load(ab + (static_cast<int64_t>(h) * rows + row) * D + (i * 32 + lane) * 8, v[h][i]);
#pragma unroll
            for (int e = 0; e < 8; ++e) s += v[h][i][e];
        }
    const float mean = warp_sum(s) * (0.5f / D);
    float q = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = v[h][i][e] - mean;
                q = fmaf(d, d, q);
            }
    const float rstd = rsqrtf(warp_sum(q) * (0.5f / D) + eps);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = h * D + (i * 32 + lane) * 8;
            float g[8], bb[8], o[8];
            V8<float>::load(ln_w + c, g);
            V8<float>::load(ln_b + c, bb);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaf((v[h][i][e] - mean) * rstd, g[e], bb[e]);
            V8<T>::store(out + static_cast<int64_t>(row) * 2 * D + c, o);
        }
}

// alpha = sigmoid(w3 . silu(hidden) + b3) of one token row (reference block/mamba_block.py:110-112), and the row's a / b
// values for the mix that follows.
//   kFold = false: `hidden` = Linear(LN(cat(a, b))) as computed by spiral_post_ln + a GEMM.
//   kFold = true : the LayerNorm is folded around the GEMM.  With W' = W * gamma (per input column), colsum[n] = sum_k W'[n][k]
//                  and cvec[n] = sum_k beta[k] W[n][k] + bias[n]:
//                      Linear(LN(x))[n] = rstd * (x . W'[n] - mean * colsum[n]) + cvec[n]
//                  so the GEMM runs on the raw cat(a, b) -- as two K = D products g2[0] = a W'_a^T, g2[1] = b W'_b^T -- and
//                  this kernel, which reads a and b anyway, supplies mean and rstd.  One launch and one (rows, 2D) round
//                  trip less per block than spiral_post_ln + GEMM.
template <typename T, typename TH, int NV, bool kFold>
__device__ __forceinline__ float mix_alpha(const T* __restrict__ ab, const TH* __restrict__ hidden, const float* __restrict__ w3,
                                           const float* __restrict__ b3, const float* __restrict__ colsum,
                                           const float* __restrict__ cvec, float eps2, int rows, int row, int lane,
                                           float (&av)[NV][8], float (&bv)[NV][8]) {
    constexpr int D = NV * 256;
    float mean = 0.f, rstd = 1.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int64_t off = static_cast<int64_t>(row) * D + (i * 32 + lane) * 8;
        V8<T>::load(ab + off, av[i]);
        V8<T>::load(ab + static_cast<int64_t>(rows) * D + off, bv[i]);
    }
    if constexpr (kFold) {
        float sm = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) sm += av[i][e] + bv[i][e];
        mean = warp_sum(sm) * (0.5f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float da = av[i][e] - mean, db = bv[i][e] - mean;
                q = fmaf(da, da, fmaf(db, db, q));
            }
        rstd = rsqrtf(warp_sum(q) * (0.5f / D) + eps2);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        float hv[8], wv[8];
        V8<TH>::load(hidden + static_cast<int64_t>(row) * D + c, hv);
        V8<float>::load(w3 + c, wv);
        if constexpr (kFold) {
            float h1[8], cs[8], cv[8];
            V8<TH>::load(hidden + (static_cast<int64_t>(rows) + row) * D + c, h1);
            V8<float>::load(colsum + c, cs);
            V8<float>::load(cvec + c, cv);
#pragma unroll
            for (int e = 0; e < 8; ++e) hv[e] = fmaf(rstd, fmaf(-mean, cs[e], hv[e] + h1[e]), cv[e]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) s = fmaf(silu_fast(hv[e]), wv[e], s);
    }
    return sigmoid_fast(warp_sum(s) + __ldg(b3));
}

template <typename T, int NV, typename TH = T, bool kFold = false>
__global__ void __launch_bounds__(kRowWarps * 32)
spiral_post_mix_kernel(const float* __restrict__ x, const float* __restrict__ skip, const T* __restrict__ ab,
                       const TH* __restrict__ hidden, const float* __restrict__ w3, const float* __restrict__ b3,
                       const float* __restrict__ mod, int64_t mod_stride, float* __restrict__ out, int rows, int L,
                       const float* __restrict__ colsum, const float* __restrict__ cvec, float eps2) {
    constexpr int D = NV * 256;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    pdl_wait();                                 // (PDL: the predecessor's writes are visible from here on)
    const int b = row / L;
    float av[NV][8], bv[NV][8];
    const float alpha = mix_alpha<T, TH, NV, kFold>(ab, hidden, w3, b3, colsum, cvec, eps2, rows, row, lane, av, bv);
    const float* gate = mod + static_cast<int64_t>(b) * mod_stride + 2 * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        const int64_t off = static_cast<int64_t>(row) * D + c;
        float xa[8], gv[8], o[8];
        V8<float>::load(x + off, xa);
        if (skip) {
            float t[8];
            V8<float>::load(skip + off, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) xa[e] += t[e];
        }
        V8<float>::load(gate + c, gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(gv[e], fmaf(alpha, av[i][e] - bv[i][e], bv[i][e]), xa[e]);
        V8<float>::store(out + off, o);
    }
}

// post_mix of block i fused with pre of block i+1 (or of the final layer): the new residual row never leaves the
// registers between the two.  x_out = (x + skip) + gate (alpha a + (1 - alpha) b)   [block i, as spiral_post_mix]
//                             out2  = modulate(LayerNorm(x_out + skip2)) [, * w]     [block i+1, as spiral_pre]
template <typename T, int NV, typename TH = T, bool kFold = false>
__global__ void __launch_bounds__(kRowWarps * 32)
spiral_post_mix_pre_kernel(const float* __restrict__ x, const float* __restrict__ skip, const T* __restrict__ ab,
                           const TH* __restrict__ hidden, const float* __restrict__ w3, const float* __restrict__ b3,
                           const float* __restrict__ mod, int64_t mod_stride, float* __restrict__ x_out,
                           const float* __restrict__ skip2, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                           const float* __restrict__ mod2, int64_t mod2_stride, const float* __restrict__ w,
                           T* __restrict__ out2, int rows, int L, float eps, const float* __restrict__ colsum,
                           const float* __restrict__ cvec, float eps2) {
    constexpr int D = NV * 256;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    pdl_wait();                                 // (PDL: the predecessor's writes are visible from here on)
    const int b = row / L;
    float av[NV][8], bv[NV][8];
    const float alpha = mix_alpha<T, TH, NV, kFold>(ab, hidden, w3, b3, colsum, cvec, eps2, rows, row, lane, av, bv);
    const float* gate = mod + static_cast<int64_t>(b) * mod_stride + 2 * D;
    float v[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        const int64_t off = static_cast<int64_t>(row) * D + c;
        float xa[8], gv[8];
        V8<float>::load(x + off, xa);
        if (skip) {
            float t[8];
            V8<float>::load(skip + off, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) xa[e] += t[e];
        }
        V8<float>::load(gate + c, gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i][e] = fmaf(gv[e], fmaf(alpha, av[i][e] - bv[i][e], bv[i][e]), xa[e]);
        V8<float>::store(x_out + off, v[i]);
        if (skip2) {
            float t[8];
            V8<float>::load(skip2 + off, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[i][e] += t[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += v[i][e];
    }
    const float mean = warp_sum(sum) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float d = v[i][e] - mean;
            q = fmaf(d, d, q);
        }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    const float wr = w ? __ldg(w + row) : 1.0f;
    const float* shift = mod2 + static_cast<int64_t>(b) * mod2_stride;
    const float* scale = shift + D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        float g[8], bb[8], sh[8], sc[8], o1[8], o2[8];
        V8<float>::load(ln_w + c, g);
        V8<float>::load(ln_b + c, bb);
        V8<float>::load(shift + c, sh);
        V8<float>::load(scale + c, sc);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float n = fmaf((v[i][e] - mean) * rstd, g[e], bb[e]);
            o1[e] = fmaf(n, 1.0f + sc[e], sh[e]);
            o2[e] = o1[e] * wr;
        }
        V8<T>::store(out2 + static_cast<int64_t>(row) * D + c, o1);
        V8<T>::store(out2 + (static_cast<int64_t>(rows) + row) * D + c, o2);
    }
}

inline int row_grid(int rows) { return (rows + kRowWarps - 1) / kRowWarps; }

}  // namespace
}  // namespace dm

using namespace dm;

extern "C" int dm_spiral_pre(const float* x, const float* skip, const float* ln_weight, const float* ln_bias,
                             const float* mod, int64_t mod_batch_stride, const float* w, void* out2, int32_t batch,
                             int32_t seqlen, int32_t d_model, float eps, int32_t act_dtype, void* stream) {
    if (!x || !ln_weight || !ln_bias || !mod || !out2 || batch <= 0 || seqlen <= 0) return DM_ERR_INVALID_ARG;
    if (d_model != 512) return DM_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(out2) || !aligned16(mod) || (skip && !aligned16(skip)) || (mod_batch_stride % 4))
        return DM_ERR_INVALID_ARG;
    const int rows = batch * seqlen;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (act_dtype == DM_BF16)
        launch_pdl(kPdlRow, spiral_pre_kernel<__nv_bfloat16, 2>, dim3(row_grid(rows)), dim3(kRowWarps * 32), 0, st, x, skip, ln_weight, ln_bias, mod, mod_batch_stride, w, static_cast<__nv_bfloat16*>(out2), rows, seqlen, eps);
    else if (act_dtype == DM_F32)
        launch_pdl(kPdlRow, spiral_pre_kernel<float, 2>, dim3(row_grid(rows)), dim3(kRowWarps * 32), 0, st, x, skip, ln_weight, ln_bias, mod, mod_batch_stride, w, static_cast<float*>(out2), rows, seqlen, eps);
    else
        return DM_ERR_UNSUPPORTED;
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

extern "C" int dm_spiral_post_ln(const void* ab, const float* ln_weight, const float* ln_bias, void* out, int32_t rows,
                                 int32_t d_model, float eps, int32_t act_dtype, void* stream) {
    if (!ab || !ln_weight || !ln_bias || !out || rows <= 0) return DM_ERR_INVALID_ARG;
    if (d_model != 512) return DM_ERR_UNSUPPORTED;
    if (!aligned16(ab) || !aligned16(out)) return DM_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (act_dtype == DM_BF16)
        launch_pdl(kPdlRow, spiral_post_ln_kernel<__nv_bfloat16, 2>, dim3(row_grid(rows)), dim3(kRowWarps * 32), 0, st, static_cast<const __nv_bfloat16*>(ab), ln_weight, ln_bias, static_cast<__nv_bfloat16*>(out), rows, eps);
    else if (act_dtype == DM_F32)
        launch_pdl(kPdlRow, spiral_post_ln_kernel<float, 2>, dim3(row_grid(rows)), dim3(kRowWarps * 32), 0, st, static_cast<const float*>(ab), ln_weight, ln_bias, static_cast<float*>(out), rows, eps);
    else
        return DM_ERR_UNSUPPORTED;
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

// hidden_dtype < 0: `hidden` is Linear(LN(cat(a, b))) in the activation dtype (the unfused form).  Otherwise `hidden` is the
// pair of raw products g2 (2, rows, D) in hidden_dtype (DM_F32 or the activation dtype) and colsum / cvec / ln2_eps carry
// the folded LayerNorm (see mix_alpha).
static int post_mix_launch(const float* x, const float* skip, const void* ab, const void* hidden, int hidden_dtype,
                           const float* colsum, const float* cvec, float ln2_eps, const float* w3, const float* b3,
                           const float* mod, int64_t mod_batch_stride, float* x_out, bool with_pre, const float* skip_next,
                           const float* ln_weight, const float* ln_bias, const float* mod_next, int64_t mod_next_batch_stride,
                           const float* w, void* out2, int32_t batch, int32_t seqlen, int32_t d_model, float eps,
                           int32_t act_dtype, void* stream) {
    if (!x || !ab || !hidden || !w3 || !b3 || !mod || !x_out || batch <= 0 || seqlen <= 0) return DM_ERR_INVALID_ARG;
    if (with_pre && (!ln_weight || !ln_bias || !mod_next || !out2)) return DM_ERR_INVALID_ARG;
    const bool fold = hidden_dtype >= 0;
    if (fold && (!colsum || !cvec || !aligned16(colsum) || !aligned16(cvec))) return DM_ERR_INVALID_ARG;
    if (d_model != 512) return DM_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(ab) || !aligned16(hidden) || !aligned16(x_out) || !aligned16(mod) ||
        (skip && !aligned16(skip)) || (mod_batch_stride % 4))
        return DM_ERR_INVALID_ARG;
    if (with_pre && (!aligned16(mod_next) || !aligned16(out2) || (skip_next && !aligned16(skip_next)) || (mod_next_batch_stride % 4)))
        return DM_ERR_INVALID_ARG;
    if (act_dtype != DM_BF16 && act_dtype != DM_F32) return DM_ERR_UNSUPPORTED;
    if (fold && hidden_dtype != DM_F32 && hidden_dtype != act_dtype) return DM_ERR_UNSUPPORTED;
    const int rows = batch * seqlen;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid(row_grid(rows)), block(kRowWarps * 32);
    auto go = [&](auto tag_t, auto tag_h, auto tag_fold) -> cudaError_t {
        using T = decltype(tag_t);
        using TH = decltype(tag_h);
        constexpr bool kF = decltype(tag_fold)::value;
        if (with_pre)
            return launch_pdl(kPdlRow, spiral_post_mix_pre_kernel<T, 2, TH, kF>, grid, block, 0, st, x, skip,
                              static_cast<const T*>(ab), static_cast<const TH*>(hidden), w3, b3, mod, mod_batch_stride, x_out,
                              skip_next, ln_weight, ln_bias, mod_next, mod_next_batch_stride, w, static_cast<T*>(out2), rows,
                              seqlen, eps, colsum, cvec, ln2_eps);
        return launch_pdl(kPdlRow, spiral_post_mix_kernel<T, 2, TH, kF>, grid, block, 0, st, x, skip, static_cast<const T*>(ab),
                          static_cast<const TH*>(hidden), w3, b3, mod, mod_batch_stride, x_out, rows, seqlen, colsum, cvec,
                          ln2_eps);
    };
    cudaError_t e;
    if (act_dtype == DM_BF16) {
        if (!fold) e = go(__nv_bfloat16{}, __nv_bfloat16{}, std::false_type{});
        else if (hidden_dtype == DM_F32) e = go(__nv_bfloat16{}, float{}, std::true_type{});
        else e = go(__nv_bfloat16{}, __nv_bfloat16{}, std::true_type{});
    } else {
        if (!fold) e = go(float{}, float{}, std::false_type{});
        else e = go(float{}, float{}, std::true_type{});
    }
    DM_CUDA_TRY(e);
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

extern "C" int dm_spiral_post_mix(const float* x, const float* skip, const void* ab, const void* hidden,
                                  const float* w3, const float* b3, const float* mod, int64_t mod_batch_stride, float* out,
                                  int32_t batch, int32_t seqlen, int32_t d_model, int32_t act_dtype, void* stream) {
    return post_mix_launch(x, skip, ab, hidden, -1, nullptr, nullptr, 0.f, w3, b3, mod, mod_batch_stride, out, false, nullptr,
                           nullptr, nullptr, nullptr, 0, nullptr, nullptr, batch, seqlen, d_model, 0.f, act_dtype, stream);
}

extern "C" int dm_spiral_post_mix_pre(const float* x, const float* skip, const void* ab, const void* hidden,
                                      const float* w3, const float* b3, const float* mod, int64_t mod_batch_stride,
                                      float* x_out, const float* skip_next, const float* ln_weight, const float* ln_bias,
                                      const float* mod_next, int64_t mod_next_batch_stride, const float* w, void* out2,
                                      int32_t batch, int32_t seqlen, int32_t d_model, float eps, int32_t act_dtype,
                                      void* stream) {
    return post_mix_launch(x, skip, ab, hidden, -1, nullptr, nullptr, 0.f, w3, b3, mod, mod_batch_stride, x_out, true, skip_next,
                           ln_weight, ln_bias, mod_next, mod_next_batch_stride, w, out2, batch, seqlen, d_model, eps, act_dtype,
                           stream);
}

extern "C" int dm_spiral_post_mix_fold(const dm_spiral_fold_args* a, void* stream) {
    if (!a) return DM_ERR_INVALID_ARG;
    if (a->g2_dtype != DM_F32 && a->g2_dtype != DM_BF16) return DM_ERR_UNSUPPORTED;
    return post_mix_launch(a->x, a->skip, a->ab, a->g2, a->g2_dtype, a->colsum, a->cvec, a->ln2_eps, a->w3, a->b3, a->mod,
                           a->mod_batch_stride, a->x_out, a->out2 != nullptr, a->skip_next, a->ln_weight, a->ln_bias,
                           a->mod_next, a->mod_next_batch_stride, a->w, a->out2, a->batch, a->seqlen, a->d_model, a->eps,
                           a->act_dtype, stream);
}

// ------------------------------------------------------------------------------------------------------
// Head of DiffMa.forward in one launch (reference model.py:264-281):
//   h = PatchEmbed(x) + pos_embed        conv with kernel = stride = patch  ==  (C p p) -> D matvec per token
//   c = cat(t_emb[t] + y, t_emb[t] + mean_T(y2));  the adaLN Linears all consume silu(c): emitted directly, act dtype
// CTAs [0, n_tok_ctas): 16 tokens each, 256 threads x 2 output columns, the patch pixels gathered from NCHW into shared
// memory, W (J, D) streamed through L2 (32 KB for patch 2).  CTAs [n_tok_ctas, n_tok_ctas + 8 B): one 64-column slab of one
// batch row of c each (incl. the token mean of an un-pooled y2).
// Replaces: unfold-copy, fp32 SIMT GEMM, bias/pos add, index_select, two adds, cat, silu, cast  (8 launches).
// ------------------------------------------------------------------------------------------------------
namespace dm {
namespace {
constexpr int kHeadTok = 16;
template <typename T>
__global__ void __launch_bounds__(256)
step_head_kernel(const float* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ posb,
                 float* __restrict__ h, int B, int C, int Himg, int patch, const int64_t* __restrict__ t,
                 const float* __restrict__ table, int table_rows, const float* __restrict__ y,
                 const float* __restrict__ y2m, int y2_tokens, T* __restrict__ sc, int n_tok_ctas) {
    constexpr int D = 512;
    extern __shared__ float px[];                       // [kHeadTok][J]
    const int tid = threadIdx.x;
    const int g = Himg / patch, L = g * g, J = C * patch * patch;
    pdl_wait();
    if (static_cast<int>(blockIdx.x) >= n_tok_ctas) {   // ---- conditioning vector: CTA = (batch row, slab of 64 columns) ----
        const int cb = blockIdx.x - n_tok_ctas;
        const int b = cb >> 3, c0 = (cb & 7) * 64;
        const int64_t tt = t[b];
        const bool ok = tt >= 0 && tt < table_rows;     // (an index beyond the table: poison the row instead of reading past it)
        // token mean of y2 (reference model.py:276) for un-pooled (B, T, D) input: thread = (row lane rl, 4 columns), every
        // load of the slab in flight at once, then a shared-memory reduction over the 16 row lanes.  T <= 1: already pooled.
        const int T2 = y2_tokens <= 1 ? 1 : y2_tokens;
        const int rl = tid >> 4, c4 = (tid & 15) * 4;
        const float* src = y2m + static_cast<int64_t>(b) * T2 * D + c0 + c4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = rl; r < T2; r += 16) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(r) * D));
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        float4* red = reinterpret_cast<float4*>(px);   // [16 row lanes][16 column groups]
        red[tid] = a;
        __syncthreads();
        if (tid < 64) {
            const int g4 = tid >> 2, e = tid & 3;
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) sum += reinterpret_cast<const float*>(&red[k * 16 + g4])[e];
            const int d = c0 + tid;
            const float te = ok ? __ldg(table + tt * D + d) : __int_as_float(0x7fc00000);
            const float c1 = te + __ldg(y + static_cast<int64_t>(b) * D + d);
            const float c2 = te + sum / static_cast<float>(T2);
            sc[static_cast<int64_t>(b) * 2 * D + d] = from_f32<T>(c1 / (1.0f + __expf(-c1)));
            sc[static_cast<int64_t>(b) * 2 * D + D + d] = from_f32<T>(c2 / (1.0f + __expf(-c2)));
        }
        return;
    }
    // ---- patch embedding of kHeadTok tokens ----
    const int tok0 = blockIdx.x * kHeadTok, n_tok = B * L;
    for (int i = tid; i < kHeadTok * J; i += 256) {
        const int tk = i / J, j = i - tk * J;
        const int tok = min(tok0 + tk, n_tok - 1);
        const int b = tok / L, l = tok - b * L, gy = l / g, gx = l - gy * g;
        const int ch = j / (patch * patch), r = j - ch * patch * patch, py = r / patch, pxl = r - py * patch;
        px[i] = __ldg(x + ((static_cast<int64_t>(b) * C + ch) * Himg + gy * patch + py) * Himg + gx * patch + pxl);
    }
    __syncthreads();
    float acc[kHeadTok][2];
#pragma unroll
    for (int k = 0; k < kHeadTok; ++k) acc[k][0] = acc[k][1] = 0.f;
    const float2* w2 = reinterpret_cast<const float2*>(wp) + tid;        // columns 2 tid, 2 tid + 1 of row j
    for (int j0 = 0; j0 < J; j0 += 8) {                                  // 8 weight rows in flight (a load per row was a serial
        float2 w[8];                                                     // chain of L2 round trips: 15 us for this kernel)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            w[i] = (j0 + i < J) ? __ldg(w2 + static_cast<int64_t>(j0 + i) * (D / 2)) : make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = min(j0 + i, J - 1);                            // (rows past J carry zero weights)
#pragma unroll
            for (int k = 0; k < kHeadTok; ++k) {
                const float v = px[k * J + j];
                acc[k][0] = fmaf(v, w[i].x, acc[k][0]);
                acc[k][1] = fmaf(v, w[i].y, acc[k][1]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kHeadTok; ++k) {
        const int tok = tok0 + k;
        if (tok < n_tok) {
            const int l = tok % L;
            const float2 pb = __ldg(reinterpret_cast<const float2*>(posb + static_cast<int64_t>(l) * D) + tid);
            reinterpret_cast<float2*>(h + static_cast<int64_t>(tok) * D)[tid] = make_float2(acc[k][0] + pb.x, acc[k][1] + pb.y);
        }
    }
}
}  // namespace
}  // namespace dm

extern "C" int dm_step_head(const float* x, const float* patch_weight, const float* pos_bias, float* h, int32_t batch,
                            int32_t channels, int32_t image_size, int32_t patch, const int64_t* t, const float* t_table,
                            int32_t table_rows, const float* y, const float* y2_mean, int32_t y2_tokens, void* silu_c,
                            int32_t d_model, int32_t act_dtype, void* stream) {
    if (!x || !patch_weight || !pos_bias || !h || !t || !t_table || !y || !y2_mean || !silu_c || batch <= 0 || channels <= 0 ||
        image_size <= 0 || patch <= 0 || table_rows <= 0)
        return DM_ERR_INVALID_ARG;
    if (d_model != 512 || image_size % patch) return DM_ERR_UNSUPPORTED;
    if (act_dtype != DM_BF16 && act_dtype != DM_F32) return DM_ERR_UNSUPPORTED;
    const int g = image_size / patch, J = channels * patch * patch;
    size_t smem = static_cast<size_t>(dm::kHeadTok) * J * sizeof(float);
    if (smem > 48 * 1024) return DM_ERR_UNSUPPORTED;
    if (smem < 256 * sizeof(float4)) smem = 256 * sizeof(float4);        // the conditioning CTAs' reduction buffer
    const int n_tok_ctas = (batch * g * g + dm::kHeadTok - 1) / dm::kHeadTok;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (act_dtype == DM_BF16)
        e = dm::launch_pdl(dm::kPdlRow, dm::step_head_kernel<__nv_bfloat16>, dim3(n_tok_ctas + 8 * batch), dim3(256), smem, st, x,
                           patch_weight, pos_bias, h, batch, channels, image_size, patch, t, t_table, table_rows, y, y2_mean,
                           y2_tokens, static_cast<__nv_bfloat16*>(silu_c), n_tok_ctas);
    else
        e = dm::launch_pdl(dm::kPdlRow, dm::step_head_kernel<float>, dim3(n_tok_ctas + 8 * batch), dim3(256), smem, st, x,
                           patch_weight, pos_bias, h, batch, channels, image_size, patch, t, t_table, table_rows, y, y2_mean,
                           y2_tokens, static_cast<float*>(silu_c), n_tok_ctas);
    DM_CUDA_TRY(e);
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

// ------------------------------------------------------------------------------------------------------
// Tail of DiffMa.forward (reference model.py:295-301, FinalLayer.linear + unpatchify): out = hn . W^T + b, written straight
// into the (B, C_out, S, S) image layout  [token (gy, gx), feature n = (py * p + px) * C_out + c  ->  out[b][c][gy p + py][gx p + px]].
// bf16 only: a 64-token x N x 512 tensor-core tile per CTA (4 warps x m16, mma.sync; the GEMM is 0.1 GFLOP -- the point is
// to drop the separate GEMM launch and the permute copy behind it).  N = p * p * C_out <= 128.
// ------------------------------------------------------------------------------------------------------
namespace dm {
namespace {
constexpr int kTailTok = 64, kTailK = 512, kTailLd = kTailK + 8;         // +8 bf16: ldmatrix rows land in different banks
__global__ void __launch_bounds__(128)
final_linear_unpatchify_kernel(const __nv_bfloat16* __restrict__ hn, const __nv_bfloat16* __restrict__ w,
                               const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, int rows, int L, int g,
                               int patch, int c_out, int N) {
    extern __shared__ __align__(16) uint8_t tail_smem[];
    __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(tail_smem);                        // [64][520]
    __nv_bfloat16* Ws = As + kTailTok * kTailLd;                                            // [N][520]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * kTailTok;
    for (int i = tid; i < N * (kTailK / 8); i += 128) {                                     // weights: independent of the predecessor
        const int n = i / (kTailK / 8), c = i % (kTailK / 8);
        cp_async16(smem_u32(Ws + n * kTailLd + c * 8), w + static_cast<int64_t>(n) * kTailK + c * 8);
    }
    pdl_wait();
    for (int i = tid; i < kTailTok * (kTailK / 8); i += 128) {
        const int r = i / (kTailK / 8), c = i % (kTailK / 8);
        const int row = min(row0 + r, rows - 1);
        cp_async16(smem_u32(As + r * kTailLd + c * 8), hn + static_cast<int64_t>(row) * kTailK + c * 8);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int S = g * patch;
    for (int nb = 0; nb < N; nb += 32) {                                                    // 4 n-tiles of 8 per pass
        float acc[4][4];
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
        for (int k0 = 0; k0 < kTailK; k0 += 16) {
            uint32_t a[4];
            ldmatrix_x4(a[0], a[1], a[2], a[3], smem_u32(As + (warp * 16 + (lane & 15)) * kTailLd + k0 + (lane >> 4) * 8));
#pragma unroll
            for (int tp = 0; tp < 2; ++tp) {                                                // two n-tiles per ldmatrix.x4
                uint32_t b[4];
                const int n = nb + tp * 16 + (lane & 7) + ((lane >> 4) << 3);
                ldmatrix_x4(b[0], b[1], b[2], b[3], smem_u32(Ws + min(n, N - 1) * kTailLd + k0 + ((lane >> 3) & 1) * 8));
                mma_bf16_16816(acc[2 * tp], a, b[0], b[1]);
                mma_bf16_16816(acc[2 * tp + 1], a, b[2], b[3]);
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = row0 + warp * 16 + (lane >> 2) + (e >> 1) * 8;
                const int n = nb + t * 8 + 2 * (lane & 3) + (e & 1);
                if (r < rows && n < N) {
                    const int b = r / L, l = r - b * L, gy = l / g, gx = l - gy * g;
                    const int c = n % c_out, pq = n / c_out, py = pq / patch, px = pq - py * patch;
                    const float v = acc[t][e] + __bfloat162float(bias[n]);
                    out[((static_cast<int64_t>(b) * c_out + c) * S + gy * patch + py) * S + gx * patch + px] = __float2bfloat16_rn(v);
                }
            }
    }
}
}  // namespace
}  // namespace dm

extern "C" int dm_final_linear_unpatchify(const void* hn, const void* weight, const void* bias, void* out, int32_t batch,
                                          int32_t grid_side, int32_t patch, int32_t out_channels, int32_t d_model,
                                          int32_t act_dtype, void* stream) {
    if (!hn || !weight || !bias || !out || batch <= 0 || grid_side <= 0 || patch <= 0 || out_channels <= 0) return DM_ERR_INVALID_ARG;
    const int N = patch * patch * out_channels;
    if (d_model != dm::kTailK || act_dtype != DM_BF16 || N > 128) return DM_ERR_UNSUPPORTED;
    if (!dm::aligned16(hn) || !dm::aligned16(weight)) return DM_ERR_INVALID_ARG;
    const int L = grid_side * grid_side, rows = batch * L;
    const size_t smem = static_cast<size_t>(dm::kTailTok + N) * dm::kTailLd * sizeof(__nv_bfloat16);
    int dev = 0, n_sm = 0;
    if (int e = dm::current_device(&dev, &n_sm); e != DM_OK) return e;
    static dm::PerDeviceOnce cfg;
    if (!cfg.done(dev)) {
        DM_CUDA_TRY(cudaFuncSetAttribute(dm::final_linear_unpatchify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>((dm::kTailTok + 128) * dm::kTailLd * sizeof(__nv_bfloat16))));
        cfg.set(dev);
    }
    DM_CUDA_TRY(dm::launch_pdl(dm::kPdlRow, dm::final_linear_unpatchify_kernel, dim3((rows + dm::kTailTok - 1) / dm::kTailTok),
                               dim3(128), smem, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(hn),
                               static_cast<const __nv_bfloat16*>(weight), static_cast<const __nv_bfloat16*>(bias),
                               static_cast<__nv_bfloat16*>(out), rows, L, grid_side, patch, out_channels, N));
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

// ------------------------------------------------------------------------------------------------------
// One reverse-diffusion update (reference gaussian_diffusion.py: p_mean_variance :254-332 with LEARNED_RANGE variance
// and epsilon prediction, p_sample :376-417) as ONE elementwise kernel instead of ~35 tiny launches + 9 table gathers:
//   eps, v = split(model_out);  logvar = frac*log(beta_t) + (1-frac)*posterior_logvar_t,  frac = (v+1)/2
//   x0 = sqrt_recip_acp_t * x - sqrt_recipm1_acp_t * eps  [clip to +-1];  mean = coef1_t * x0 + coef2_t * x
//   x_{t-1} = mean + [t != 0] * exp(0.5*logvar) * noise
// `table` is the (n_rows, T) fp32 schedule table of diffusion.py (_ROWS order), gathered here by t.
// ------------------------------------------------------------------------------------------------------
namespace dm {
namespace {
__global__ void __launch_bounds__(256)
p_sample_update_kernel(const float* __restrict__ model_out, const float* __restrict__ x, const float* __restrict__ noise,
                       const float* __restrict__ table, const int64_t* __restrict__ t, float* __restrict__ sample,
                       float* __restrict__ pred_xstart, int n, int chw, int T, int clip) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<int64_t>(n) * chw) return;
    pdl_wait();
    const int b = static_cast<int>(idx / chw), r = static_cast<int>(idx % chw);
    const int64_t tt = t[b];
    // rows: 2 sqrt_recip_acp, 3 sqrt_recipm1_acp, 5 posterior_log_variance_clipped, 6 coef1, 7 coef2, 8 log_betas
    const float rc = __ldg(table + 2 * T + tt), rm1 = __ldg(table + 3 * T + tt), minl = __ldg(table + 5 * T + tt),
                c1 = __ldg(table + 6 * T + tt), c2 = __ldg(table + 7 * T + tt), maxl = __ldg(table + 8 * T + tt);
    const float eps = model_out[static_cast<int64_t>(b) * 2 * chw + r];
    const float v = model_out[static_cast<int64_t>(b) * 2 * chw + chw + r];
    const float xv = x[idx];
    const float frac = 0.5f * (v + 1.0f);
    const float logvar = frac * maxl + (1.0f - frac) * minl;
    float x0 = rc * xv - rm1 * eps;
    if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    const float mean = c1 * x0 + c2 * xv;
    const float nz = tt != 0 ? 1.0f : 0.0f;
    sample[idx] = mean + nz * __expf(0.5f * logvar) * noise[idx];
    if (pred_xstart) pred_xstart[idx] = x0;
}
}  // namespace
}  // namespace dm

extern "C" int dm_p_sample_update(const float* model_out, const float* x, const float* noise, const float* table,
                                  const int64_t* t, float* sample, float* pred_xstart, int32_t batch, int32_t chw,
                                  int32_t n_steps, int32_t clip_denoised, void* stream) {
    if (!model_out || !x || !noise || !table || !t || !sample || batch <= 0 || chw <= 0 || n_steps <= 0)
        return DM_ERR_INVALID_ARG;
    const int64_t total = static_cast<int64_t>(batch) * chw;
    dm::launch_pdl(dm::kPdlRow, dm::p_sample_update_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0,
                   static_cast<cudaStream_t>(stream), model_out, x, noise, table, t, sample, pred_xstart, batch, chw, n_steps, clip_denoised);
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}
