// Optimizer + EMA update of the training step as ONE pass over flat fp32 buffers (SURVEY.md section 8a row a13 and the
// loop around it: reference train.py:34-43 `update_ema`, :201 AdamW(lr 1e-4, weight_decay 0), :262-264 opt.step();
// update_ema(ema, model.module)).
//
// The reference runs torch's AdamW over ~1 900 parameter tensors and then a Python loop of mul_/add_ per tensor for the
// EMA.  Here parameters, gradients, both moments and the EMA copy live in five flat buffers (diffma_b200/ddp.py), so the
// whole update is one HBM-bound elementwise kernel: 20 B read + 16 B written per parameter (155 M parameters for
// DiffMa-XL: 5.6 GB, ~0.9 ms at the measured copy bandwidth) and it can run per gradient bucket as soon as that bucket's
// all-reduce has landed.  `step` lives in device memory so the launch is CUDA-graph replayable.
#include <cmath>

#include "dm_common.cuh"

namespace dm {
namespace {

struct AdamArgs {
    float* p; const float* g; float* m; float* v; float* ema;
    __nv_bfloat16* shadow;       // optional: bf16 copy of the updated parameters (the compute-dtype weights of the next step)
    const float* step;           // device scalar: number of optimizer steps INCLUDING this one (t >= 1)
    int64_t n;
    // all derived on the host in double and rounded once (torch hands `1 - beta2` etc. to its kernels the same way)
    float beta1, one_m_beta1, beta2, one_m_beta2, log2_beta1, log2_beta2;
    float lr, eps, decay, ema_decay, one_m_ema_decay, grad_scale;
};

__global__ void __launch_bounds__(256) adamw_ema_kernel(const AdamArgs a) {
    // bias corrections from the device-side step counter (one lg2/ex2 pair per thread: free next to the HBM traffic)
    const float t = __ldg(a.step);
    const float bc1 = 1.0f - exp2f(t * a.log2_beta1);
    const float bc2 = 1.0f - exp2f(t * a.log2_beta2);
    const float step_size = a.lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    const float decay = a.decay;
    const int64_t n4 = a.n / 4;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 p = reinterpret_cast<const float4*>(a.p)[i];
        const float4 g4 = __ldcs(reinterpret_cast<const float4*>(a.g) + i);       // gradients are dead after this pass
        float4 m = reinterpret_cast<const float4*>(a.m)[i];
        float4 v = reinterpret_cast<const float4*>(a.v)[i];
        // (loaded up front with the others: behind the stores below it could not be hoisted -- the buffers may alias as
        // far as the compiler knows -- and every iteration would pay a second DRAM latency)
        float4 e = a.ema != nullptr ? reinterpret_cast<const float4*>(a.ema)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float pv[4] = {p.x, p.y, p.z, p.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w};
        float mv[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float g = gv[q] * a.grad_scale;
            pv[q] *= decay;                                                      // decoupled weight decay (torch AdamW)
            mv[q] = fmaf(a.beta1, mv[q], a.one_m_beta1 * g);                     // exp_avg.lerp_(grad, 1 - beta1)
            vv[q] = fmaf(a.beta2, vv[q], a.one_m_beta2 * g * g);                 // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
            const float denom = sqrtf(vv[q]) * inv_sqrt_bc2 + a.eps;
            pv[q] -= step_size * (mv[q] / denom);
        }
        reinterpret_cast<float4*>(a.p)[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
        reinterpret_cast<float4*>(a.m)[i] = make_float4(mv[0], mv[1], mv[2], mv[3]);
        reinterpret_cast<float4*>(a.v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        if (a.shadow != nullptr) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(pv[0], pv[1]), hi = __floats2bfloat162_rn(pv[2], pv[3]);
            uint2 w;
            w.x = *reinterpret_cast<const uint32_t*>(&lo);
            w.y = *reinterpret_cast<const uint32_t*>(&hi);
            reinterpret_cast<uint2*>(a.shadow)[i] = w;
        }
        if (a.ema != nullptr) {                                                  // ema.mul_(decay).add_(p, alpha = 1 - decay)
            e.x = fmaf(a.ema_decay, e.x, a.one_m_ema_decay * pv[0]);
            e.y = fmaf(a.ema_decay, e.y, a.one_m_ema_decay * pv[1]);
            e.z = fmaf(a.ema_decay, e.z, a.one_m_ema_decay * pv[2]);
            e.w = fmaf(a.ema_decay, e.w, a.one_m_ema_decay * pv[3]);
            reinterpret_cast<float4*>(a.ema)[i] = e;
        }
    }
    // tail (n not a multiple of 4): the first threads of block 0
    if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {
        const int64_t i = n4 * 4 + threadIdx.x;
        const float g = a.g[i] * a.grad_scale;
        float p = a.p[i] * decay;
        const float m = fmaf(a.beta1, a.m[i], a.one_m_beta1 * g);
        const float v = fmaf(a.beta2, a.v[i], a.one_m_beta2 * g * g);
        p -= step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + a.eps));
        a.p[i] = p; a.m[i] = m; a.v[i] = v;
        if (a.shadow != nullptr) a.shadow[i] = __float2bfloat16_rn(p);
        if (a.ema != nullptr) a.ema[i] = fmaf(a.ema_decay, a.ema[i], a.one_m_ema_decay * p);
    }
}

}  // namespace
}  // namespace dm

extern "C" int dm_adamw_ema_step_ex(const dm_adamw_args* x, void* stream) {
    using namespace dm;
    if (!x || !x->param || !x->grad || !x->exp_avg || !x->exp_avg_sq || !x->step || x->n <= 0) return DM_ERR_INVALID_ARG;
    if (!aligned16(x->param) || !aligned16(x->grad) || !aligned16(x->exp_avg) || !aligned16(x->exp_avg_sq) ||
        (x->ema && !aligned16(x->ema)) || (x->shadow_bf16 && (reinterpret_cast<uintptr_t>(x->shadow_bf16) & 7u)))
        return DM_ERR_INVALID_ARG;
    const double beta1 = x->beta1, beta2 = x->beta2, ema_decay = x->ema_decay;
    if (!(beta1 > 0. && beta1 < 1.) || !(beta2 > 0. && beta2 < 1.) || !(ema_decay >= 0. && ema_decay <= 1.))
        return DM_ERR_INVALID_ARG;
    int dev = 0, n_sm = 0;
    if (int e = current_device(&dev, &n_sm); e != DM_OK) return e;
    AdamArgs a{};
    a.p = x->param; a.g = x->grad; a.m = x->exp_avg; a.v = x->exp_avg_sq; a.ema = x->ema; a.step = x->step; a.n = x->n;
    a.shadow = static_cast<__nv_bfloat16*>(x->shadow_bf16);
    a.beta1 = static_cast<float>(beta1); a.one_m_beta1 = static_cast<float>(1.0 - beta1);
    a.beta2 = static_cast<float>(beta2); a.one_m_beta2 = static_cast<float>(1.0 - beta2);
    a.log2_beta1 = static_cast<float>(log2(beta1)); a.log2_beta2 = static_cast<float>(log2(beta2));
    a.lr = static_cast<float>(x->lr); a.eps = static_cast<float>(x->eps);
    a.decay = static_cast<float>(1.0 - x->lr * x->weight_decay);
    a.ema_decay = static_cast<float>(ema_decay); a.one_m_ema_decay = static_cast<float>(1.0 - ema_decay);
    a.grad_scale = static_cast<float>(x->grad_scale);
    // grid-stride, 8 CTAs of 256 threads per SM: enough 16-byte loads in flight to saturate HBM
    const int64_t want = (x->n / 4 + 255) / 256;
    const int64_t cap = static_cast<int64_t>(n_sm) * 8;
    const int grid = static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
    adamw_ema_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

extern "C" int dm_adamw_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema,
                                 const float* step, int64_t n, double lr, double beta1, double beta2, double eps,
                                 double weight_decay, double ema_decay, double grad_scale, void* stream) {
    dm_adamw_args x{};
    x.param = param; x.grad = grad; x.exp_avg = exp_avg; x.exp_avg_sq = exp_avg_sq; x.ema = ema; x.step = step;
    x.shadow_bf16 = nullptr; x.n = n; x.lr = lr; x.beta1 = beta1; x.beta2 = beta2; x.eps = eps;
    x.weight_decay = weight_decay; x.ema_decay = ema_decay; x.grad_scale = grad_scale;
    return dm_adamw_ema_step_ex(&x, stream);
}
