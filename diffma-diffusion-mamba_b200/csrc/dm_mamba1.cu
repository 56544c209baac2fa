// Mamba-1 forward hot path for sm_100a (SURVEY.md section 8a rows a2 + a4, reference call sites
// block/mamba.py:346-393 and the CrossScan/CrossMerge gathers block/mamba.py:32-82).
//
// Two kernels, one C-ABI call, all directions x all mixers ("groups") of a block in each launch:
//
//   m1_conv_xproj_persistent / m1_conv_xproj_kernel (kernel P)
//                          gather rows by the scan order -> causal conv1d (sliding window in registers)
//                          -> SiLU -> u (scan order, act dtype) ; x_dbl = u . W_x^T on mma.sync (bf16 operands, fp32
//                          accumulate; fp32 I/O uses the 3-term bf16 split so the result is fp32-accurate) -> x_dbl.
//                          bf16 / d_inner 1024: persistent, one CTA per SM, W_x resident in shared memory, 16-token
//                          tiles double buffered by cp.async; otherwise one CTA per 32-token tile.
//   m1_scan_kernel (kernel S)  one WARP per (sequence, 32 or 64 channels), lane = 1 or 2 channels, 16 states per
//                          channel in registers as packed fp32 pairs: per 8-token chunk cp.async-staged x_dbl / u / z
//                          rows (double buffered), dt_proj on mma.sync (W_dt slice in shared memory), softplus, then
//                          the sequential recurrence h = exp2(dt*A*log2e) h + dt*u*B ; y = <h, C> + D u ;
//                          out = y*silu(z) written straight to its un-permuted (token-order) row.  Static launch or
//                          persistent ready-queue schedule (see the kernel); optional state checkpoints for training.
//
// Why the x_proj cut: B_t, C_t and dt_low_t are reductions over all d_inner channels of token t, so no
// channel-sliced CTA can start scanning a token before every channel of it is convolved.  The scan is
// MUFU-bound (16 ex2 per (b,d,l)), needs channel-sliced parallelism to fill 148 SMs, and the x_dbl
// round trip is 256 B/token in L2 -- so the cut costs nothing measurable.
#include <cstdlib>

#include "dm_common.cuh"

namespace dm {

namespace {

constexpr int kN = 16;        // d_state
constexpr int kW = 4;         // d_conv
constexpr int kE = 64;        // dt_rank + 2*d_state handled by this build (R = 32)
constexpr int kR = 32;
constexpr int kTP = 32;       // tokens per conv/x_proj tile
constexpr int kPThreads = 256;
constexpr int kCH = 8;        // tokens per scan chunk

struct M1G {
    const void* xz;
    int64_t xz_bs, xz_ts;
    void* out;
    int64_t out_bs, out_ds, out_ts;
    void* u;
    float* x_dbl;
    const float* conv_w;
    const float* conv_b;
    const void* wx;
    const void* wdt;
    const float* dt_bias;
    const float* A;
    const float* D;
    float* chunk_states;
    __half* delta;             // optional (seq, L, D) fp16, scan order: softplus(dt_proj(dt_low) + bias), written by kernel P
};

struct M1P {
    int B, K, L, D;
    int out_order, n_groups;
    int tiles_per_seq;
    const int32_t* order;
    // dynamic schedule of the scan kernel (see m1_scan_kernel): workspace + segmentation, or sched == nullptr
    int* sched;
    int seg_chunks, n_segs;
    int save_every;            // training: tokens between two saved states (dm_mamba1_bwd_chunk_tokens), else 0
    int z_gated;               // the z half of xz already holds silu(z) (in-projection epilogue): multiply, no SiLU
    M1G g[DM_MAX_GROUPS];
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);   // .x = a (low half)
    return *reinterpret_cast<uint32_t*>(&v);
}
// x ~= hi + lo with hi, lo bf16: the pair carries ~16 mantissa bits, enough for fp32-grade GEMM results
__device__ __forceinline__ void split_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    hi = pack_bf16(__bfloat162float(ah), __bfloat162float(bh));
    lo = pack_bf16(a - __bfloat162float(ah), b - __bfloat162float(bh));
}

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
        uint2 t = *reinterpret_cast<const uint2*>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
        v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
        uint2 t;
        t.x = pack_bf16(v[0], v[1]);
        t.y = pack_bf16(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = t;
    }
};

__device__ __forceinline__ const int32_t* dir_order(const M1P& p, int k) {
    if (p.order == nullptr) return nullptr;
    const int32_t* o = p.order + static_cast<int64_t>(k) * p.L;
    return (__ldg(o) < 0) ? nullptr : o;
}

// ------------------------------------------------------------------------------------------------------
// Kernel P: gather + causal conv1d + SiLU -> u ; x_dbl = u . W_x^T
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kPThreads) m1_conv_xproj_kernel(const __grid_constant__ M1P p) {
    constexpr bool kSplit = sizeof(T) == 4;     // fp32 I/O: 3-term bf16 split on both operands
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int D = p.D;
    const int ldu = D + 8;                       // bf16 elements per smem row; (D+8)*2 B = odd multiple of 16 B
    __nv_bfloat16* u_hi = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* u_lo = u_hi + kTP * ldu;      // only touched when kSplit

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int seq = tile / p.tiles_per_seq, jt = tile - seq * p.tiles_per_seq;
    const int k = seq % p.K, b = (seq / p.K) % p.B, g = seq / (p.K * p.B);
    const M1G& G = p.g[g];
    const int L = p.L, j0 = jt * kTP;
    const int32_t* ord = dir_order(p, k);
    const T* x_base = static_cast<const T*>(G.xz) + static_cast<int64_t>(b) * G.xz_bs;
    const int64_t seq_in_group = static_cast<int64_t>(b) * p.K + k;
    T* u_out = static_cast<T*>(G.u) + seq_in_group * L * D;

    // ---- conv + SiLU: each thread owns 4 adjacent channels and slides over the tile's tokens ----
    for (int c = tid * 4; c < D; c += kPThreads * 4) {
        float w[4][kW], bias[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 t = __ldg(reinterpret_cast<const float4*>(G.conv_w + static_cast<int64_t>(c + q) * kW));
            w[q][0] = t.x; w[q][1] = t.y; w[q][2] = t.z; w[q][3] = t.w;
            bias[q] = G.conv_b ? __ldg(G.conv_b + c + q) : 0.f;
        }
        float win[3][4];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int j = j0 - 3 + t;
            if (j >= 0) {
                const int src = ord ? __ldg(ord + j) : j;
                Vec4<T>::load(x_base + static_cast<int64_t>(src) * G.xz_ts + c, win[t]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) win[t][q] = 0.f;
            }
        }
#pragma unroll 4
        for (int jj = 0; jj < kTP; ++jj) {
            const int j = j0 + jj;
            float uv[4] = {0.f, 0.f, 0.f, 0.f};
            if (j < L) {
                float xn[4];
                const int src = ord ? __ldg(ord + j) : j;
                Vec4<T>::load(x_base + static_cast<int64_t>(src) * G.xz_ts + c, xn);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float acc = bias[q];
                    acc = fmaf(w[q][0], win[0][q], acc);
                    acc = fmaf(w[q][1], win[1][q], acc);
                    acc = fmaf(w[q][2], win[2][q], acc);
                    acc = fmaf(w[q][3], xn[q], acc);
                    uv[q] = silu_fast(acc);
                    win[0][q] = win[1][q]; win[1][q] = win[2][q]; win[2][q] = xn[q];
                }
                Vec4<T>::store(u_out + static_cast<int64_t>(j) * D + c, uv);
                if constexpr (!kSplit) {   // the MMA must see exactly the values the scan will read back
#pragma unroll
                    for (int q = 0; q < 4; ++q) uv[q] = __bfloat162float(__float2bfloat16_rn(uv[q]));
                }
            }
            uint32_t h0, h1, l0, l1;
            split_bf16(uv[0], uv[1], h0, l0);
            split_bf16(uv[2], uv[3], h1, l1);
            *reinterpret_cast<uint2*>(u_hi + jj * ldu + c) = make_uint2(h0, h1);
            if constexpr (kSplit) *reinterpret_cast<uint2*>(u_lo + jj * ldu + c) = make_uint2(l0, l1);
        }
    }
    __syncthreads();

    // ---- x_dbl tile (32 x 64) = u tile (32 x D) . W_x^T ; each warp takes a K-slice of D/8 channels ----
    float acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;

    const int kslice = D / 8;
    const int kbeg = warp * kslice;
    const T* Wx = static_cast<const T*>(G.wx);
    for (int k0 = kbeg; k0 < kbeg + kslice; k0 += 16) {
        uint32_t a_hi[2][4], a_lo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int row = mt * 16 + (lane & 15), col = k0 + (lane >> 4) * 8;
            ldmatrix_x4(a_hi[mt][0], a_hi[mt][1], a_hi[mt][2], a_hi[mt][3], smem_u32(u_hi + row * ldu + col));
            if constexpr (kSplit)
                ldmatrix_x4(a_lo[mt][0], a_lo[mt][1], a_lo[mt][2], a_lo[mt][3], smem_u32(u_lo + row * ldu + col));
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const T* wp = Wx + static_cast<int64_t>(nt * 8 + (lane >> 2)) * D + k0 + 2 * (lane & 3);
            uint32_t b_hi[2], b_lo[2];
            if constexpr (kSplit) {
                float2 w0 = __ldg(reinterpret_cast<const float2*>(wp));
                float2 w1 = __ldg(reinterpret_cast<const float2*>(wp + 8));
                split_bf16(w0.x, w0.y, b_hi[0], b_lo[0]);
                split_bf16(w1.x, w1.y, b_hi[1], b_lo[1]);
            } else {
                b_hi[0] = __ldg(reinterpret_cast<const uint32_t*>(wp));
                b_hi[1] = __ldg(reinterpret_cast<const uint32_t*>(wp + 8));
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                mma_bf16_16816(acc[mt][nt], a_hi[mt], b_hi[0], b_hi[1]);
                if constexpr (kSplit) {
                    mma_bf16_16816(acc[mt][nt], a_lo[mt], b_hi[0], b_hi[1]);
                    mma_bf16_16816(acc[mt][nt], a_hi[mt], b_lo[0], b_lo[1]);
                }
            }
        }
    }
    __syncthreads();                                   // everyone is done reading u_hi/u_lo: reuse as reduce buffer
    float* red = reinterpret_cast<float*>(smem_raw);   // [8 warps][32][64]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int row = mt * 16 + (lane >> 2), col = nt * 8 + 2 * (lane & 3);
            float* r = red + (warp * kTP + row) * kE + col;
            *reinterpret_cast<float2*>(r) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
            *reinterpret_cast<float2*>(r + 8 * kE) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
        }
    __syncthreads();
    float* xd_out = G.x_dbl + seq_in_group * L * kE;
    for (int o = tid * 4; o < kTP * kE; o += kPThreads * 4) {
        const int row = o / kE;
        if (j0 + row >= L) continue;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) {
            float4 t = *reinterpret_cast<const float4*>(red + w8 * kTP * kE + o);
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        float* dst = xd_out + static_cast<int64_t>(j0) * kE + o;
        const int col = o % kE;
        if (col < kR) {            // dt_low: 32 values -> [hi: 32 bf16 | lo: 32 bf16] in the first 32 float slots
            uint32_t h0, l0, h1, l1;
            split_bf16(s.x, s.y, h0, l0);
            split_bf16(s.z, s.w, h1, l1);
            uint32_t* rowp = reinterpret_cast<uint32_t*>(dst - col);
            *reinterpret_cast<uint2*>(rowp + col / 2) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(rowp + 16 + col / 2) = make_uint2(l0, l1);
        } else {
            *reinterpret_cast<float4*>(dst) = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Kernel P2 (bf16 I/O, d_inner <= 1024): persistent version of kernel P.
//   One CTA per SM keeps W_x (64 x D bf16, 132 KB) resident in shared memory for its whole life and walks
//   16-token tiles of the scanned sequences: cp.async-staged, double-buffered x tiles (the gather by scan order
//   happens in the copy), conv + SiLU in place (u overwrites the x rows it no longer needs), x_proj on mma.sync
//   with BOTH operands read by ldmatrix from shared memory (no global latency inside the MMA loop), K split over
//   the 8 warps, partials reduced through the tile buffer.  W_x is read from L2 once per CTA (19 MB total)
//   instead of once per tile (86 MB), and tile loads overlap the previous tile's math.
// ------------------------------------------------------------------------------------------------------
constexpr int kTP2 = 16;
constexpr int kP2Threads = 512;

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// silu(x) = x * sigmoid(x) = h + h * tanh(h), h = x/2: one MUFU instead of two; its 2^-11 relative error is below
// the bf16 rounding applied to u right after
__device__ __forceinline__ float silu_tanh(float x) {
    const float h = 0.5f * x;
    return fmaf(h, tanh_approx(h), h);
}

// softplus with torch's threshold on a log2(e)-scaled argument (same formula as the scan kernel's softplus_scaled)
__device__ __forceinline__ float softplus_scaled_p(float s) {
    float e, l;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(s));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));
    return 0.6931471805599453f * (s > 28.853900817779268f ? s : l);
}

template <int kD>
__global__ void __launch_bounds__(kP2Threads, 1) m1_conv_xproj_persistent(const __grid_constant__ M1P p, int n_tiles) {
    using T = __nv_bfloat16;
    static_assert(kD == 2 * kP2Threads, "one channel pair per thread");
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr int ld = kD + 8;                               // bf16 elements per shared row (16-byte odd multiple)
    constexpr int kSegs = kD * 2 / 16;                       // 16-byte segments per row
    constexpr int tile_elems = (kTP2 + 3) * ld;
    T* Ws = reinterpret_cast<T*>(smem_raw);                  // [64][ld]
    T* Xs = Ws + kE * ld;                                    // [2][kTP2 + 3][ld]
    int32_t* ord_s = reinterpret_cast<int32_t*>(Xs + 2 * tile_elems);   // [K][L] scan orders, loaded once per CTA
    const int L = p.L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_per_seq = (L + kTP2 - 1) / kTP2;
    const bool has_order = p.order != nullptr;

    // tile -> (mixer g, batch b, direction k, first token j0).  Decoded once per CTA with divisions, then advanced
    // incrementally: the runtime divisions cost ~250 instructions per thread and tile in the first version
    struct TilePos { int g, b, k, j0; };
    auto decode = [&](int tile) {
        TilePos t;
        const int seq = tile / tiles_per_seq;
        t.j0 = (tile - seq * tiles_per_seq) * kTP2;
        t.k = seq % p.K; t.b = (seq / p.K) % p.B; t.g = seq / (p.K * p.B);
        return t;
    };
    auto advance = [&](TilePos t) {
        t.j0 += kTP2;
        if (t.j0 >= L) {
            t.j0 = 0;
            if (++t.k == p.K) {
                t.k = 0;
                if (++t.b == p.B) { t.b = 0; ++t.g; }
            }
        }
        return t;
    };
    auto load_tile = [&](const TilePos& tp, int buf) {
        const int g = tp.g, b = tp.b, k = tp.k, j0 = tp.j0;
        const M1G& G = p.g[g];
        const int32_t* ord = has_order ? ord_s + k * L : nullptr;
        const T* x_base = static_cast<const T*>(G.xz) + static_cast<int64_t>(b) * G.xz_bs;
        T* dst = Xs + buf * tile_elems;
#pragma unroll
        for (int i = 0; i < ((kTP2 + 3) * kSegs + kP2Threads - 1) / kP2Threads; ++i) {
            const int s = tid + i * kP2Threads;
            const int r = s / kSegs, part = s % kSegs;
            const int j = j0 - 3 + r;
            if (r < kTP2 + 3) {
                T* d = dst + r * ld + part * 8;
                if (j < 0) {
                    *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
                } else if (j < L) {
                    const int src = ord ? ord[j] : j;
                    cp_async16(smem_u32(d), x_base + static_cast<int64_t>(src) * G.xz_ts + part * 8);
                }
            }
        }
    };

    // per-group state: W_x resident in shared memory, this thread's conv taps in registers
    int cur_group = -1;
    float cw[2][kW], cb[2];
    const int c = tid * 2;
    auto load_group = [&](int g) {
        const M1G& G = p.g[g];
        const T* Wx = static_cast<const T*>(G.wx);
        for (int s = tid; s < kE * kSegs; s += kP2Threads) {
            const int r = s / kSegs, part = s % kSegs;
            cp_async16(smem_u32(Ws + r * ld + part * 8), Wx + static_cast<int64_t>(r) * kD + part * 8);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(G.conv_w + static_cast<int64_t>(c + q) * kW));
            cw[q][0] = t.x; cw[q][1] = t.y; cw[q][2] = t.z; cw[q][3] = t.w;
            cb[q] = G.conv_b ? __ldg(G.conv_b + c + q) : 0.f;
        }
        cur_group = g;
    };

    // contiguous tile range per CTA (consecutive tiles share their sequence: halo rows and order rows stay hot)
    const int tile_lo = static_cast<int>(static_cast<int64_t>(blockIdx.x) * n_tiles / gridDim.x);
    const int tile_hi = static_cast<int>(static_cast<int64_t>(blockIdx.x + 1) * n_tiles / gridDim.x);
    if (tile_lo >= tile_hi) return;
    // scan orders of all directions -> shared memory (identity directions get the identity)
    if (has_order) {
        for (int i = tid; i < p.K * L; i += kP2Threads) {
            const int kk = i / L, j = i - kk * L;
            const int first = __ldg(p.order + static_cast<int64_t>(kk) * L);
            ord_s[i] = first < 0 ? j : __ldg(p.order + i);
        }
        __syncthreads();
    }
    int tile = tile_lo;
    TilePos cur = decode(tile);
    load_group(cur.g);
    load_tile(cur, 0);
    cp_async_commit();
    int buf = 0;
    for (; tile < tile_hi; ++tile, buf ^= 1) {
        const int g = cur.g, b = cur.b, k = cur.k, j0 = cur.j0;
        const M1G& G = p.g[g];
        const TilePos nxt = advance(cur);
        if (tile + 1 < tile_hi) load_tile(nxt, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (g != cur_group) {               // crossed into another mixer's tiles: swap the resident weights
            load_group(g);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
        }

        // ---- conv + SiLU in place; thread owns 2 adjacent channels ----
        T* xt = Xs + buf * tile_elems;
        const int64_t seq_in_group = static_cast<int64_t>(b) * p.K + k;
        {
            T* u_out = static_cast<T*>(G.u) + (seq_in_group * L + j0) * kD + c;
            float win[3][2];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(xt + t * ld + c);
                win[t][0] = __low2float(v); win[t][1] = __high2float(v);
            }
            const int nvalid = L - j0;
#pragma unroll
            for (int jj = 0; jj < kTP2; ++jj) {
                const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(xt + (jj + 3) * ld + c);
                const float xn[2] = {__low2float(v), __high2float(v)};
                float uv[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float acc = cb[q];
                    acc = fmaf(cw[q][0], win[0][q], acc);
                    acc = fmaf(cw[q][1], win[1][q], acc);
                    acc = fmaf(cw[q][2], win[2][q], acc);
                    acc = fmaf(cw[q][3], xn[q], acc);
                    uv[q] = silu_tanh(acc);
                    win[0][q] = win[1][q]; win[1][q] = win[2][q]; win[2][q] = xn[q];
                }
                const uint32_t packed = pack_bf16(uv[0], uv[1]);
                *reinterpret_cast<uint32_t*>(xt + jj * ld + c) = packed;
                if (jj < nvalid) *reinterpret_cast<uint32_t*>(u_out + static_cast<int64_t>(jj) * kD) = packed;
            }
        }
        __syncthreads();

        // ---- x_dbl tile (16 x 64) = u tile (16 x D) . W_x^T ; warp = (K-slice of D/8 channels, half of the outputs) ----
        const int ks = warp & 7, nh = warp >> 3;
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
        for (int kk = 0; kk < kD / 8 / 32; ++kk) {
            const int k0 = ks * (kD / 8) + kk * 32;
            uint32_t a0[4], a1[4];
            ldmatrix_x4(a0[0], a0[1], a0[2], a0[3], smem_u32(xt + (lane & 15) * ld + k0 + (lane >> 4) * 8));
            ldmatrix_x4(a1[0], a1[1], a1[2], a1[3], smem_u32(xt + (lane & 15) * ld + k0 + 16 + (lane >> 4) * 8));
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                uint32_t bw[4];     // B fragments of k-steps k0 and k0+16 for 8 output rows
                ldmatrix_x4(bw[0], bw[1], bw[2], bw[3],
                            smem_u32(Ws + ((nh * 4 + nt) * 8 + (lane & 7)) * ld + k0 + (lane >> 3) * 8));
                mma_bf16_16816(acc[nt], a0, bw[0], bw[1]);
                mma_bf16_16816(acc[nt], a1, bw[2], bw[3]);
            }
        }
        __syncthreads();                                   // all warps done with the u rows: reuse them as reduce buffer
        float* red = reinterpret_cast<float*>(xt);         // [8 K-slices][16][64] fp32 = 32 KB <= tile buffer
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int row = lane >> 2, col = (nh * 4 + nt) * 8 + 2 * (lane & 3);
            float* r = red + (ks * kTP2 + row) * kE + col;
            *reinterpret_cast<float2*>(r) = make_float2(acc[nt][0], acc[nt][1]);
            *reinterpret_cast<float2*>(r + 8 * kE) = make_float2(acc[nt][2], acc[nt][3]);
        }
        __syncthreads();
        {
            const int o = tid * 2;                          // 16 x 64 outputs, 2 per thread
            const int row = o / kE, col = o % kE;
            if (j0 + row < L) {
                float2 s = make_float2(0.f, 0.f);
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) {
                    const float2 t = *reinterpret_cast<const float2*>(red + w8 * kTP2 * kE + o);
                    s.x += t.x; s.y += t.y;
                }
                float* rowp = G.x_dbl + (seq_in_group * L + j0 + row) * kE;
                if (col < kR) {
                    uint32_t hi, lo;
                    split_bf16(s.x, s.y, hi, lo);
                    reinterpret_cast<uint32_t*>(rowp)[col / 2] = hi;
                    reinterpret_cast<uint32_t*>(rowp)[16 + col / 2] = lo;
                } else {
                    *reinterpret_cast<float2*>(rowp + col) = s;
                }
            }
        }
        __syncthreads();                                   // reduce buffer free before the next prefetch lands in it

        cur = nxt;
    }
}

// ------------------------------------------------------------------------------------------------------
// Kernel D (bf16 inference): delta = softplus(dt_low . W_dt^T + bias) -> fp16 (seq, L, D), handed to kernel S.
//   The scan is bound by the MUFU pipe; the softplus (2 MUFU per (token, channel)), the dt_proj MMA and its
//   shared-memory round trip cost it 131 -> 109 us at the headline shape, while as a stand-alone, fully parallel kernel
//   they are HBM/XU-balanced and ~10 us (profiles/r02_notes.md; inside the persistent kernel P, whose phases are
//   serialised by block barriers, the same work cost +30 us and was moved out again).
//   One warp per (32 scanned tokens, 64 channels): A = dt_low hi + lo halves read straight from the x_dbl rows (the
//   m16n8k16 A fragment is 4 words of a row), B = the W_dt rows of the warp's channels, both from L2 / L1.
// ------------------------------------------------------------------------------------------------------
// (kDTok, the template parameter of m1_delta_kernel: scanned tokens per warp, 64 or 128; the W_dt fragments (32 registers) are
//  loaded once per kDTok x 64 tile)
struct DeltaSmem {                 // per warp; every row stride is an odd multiple of 16 B (ldmatrix / 16-byte accesses conflict free)
    __nv_bfloat16 wd[64][kR + 8];  // the warp's W_dt rows
    __nv_bfloat16 at[16][64 + 8];  // dt_low of 16 tokens: [hi 32 | lo 32]
    __half stage[16][64 + 8];      // softplus'ed tile in fragment order -> row-order copy-out
};
template <int kDTok>
__global__ void __launch_bounds__(128) m1_delta_kernel(const __grid_constant__ M1P p, int rows_per_group) {
    using T = __nv_bfloat16;
    __shared__ __align__(16) DeltaSmem smem[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    DeltaSmem& S = smem[warp];
    const int D = p.D;
    const int cblocks = D / 256;                                 // a CTA covers 256 channels (4 warps x 64)
    const int tiles_per_group = (rows_per_group + kDTok - 1) / kDTok;
    int t = blockIdx.x;
    const int cb = t % cblocks; t /= cblocks;
    const int tile = t % tiles_per_group, g = t / tiles_per_group;
    const M1G& G = p.g[g];
    const int row0 = tile * kDTok;                               // rows = (b, k, j) flattened: x_dbl and delta are contiguous
    const int c0 = cb * 256 + warp * 64;
    const int q = lane & 3, r = lane >> 2;
    // ---- W_dt rows c0 .. c0+63 (4 KB contiguous) -> shared, 16-byte coalesced; B fragments by ldmatrix ----
    {
        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const T*>(G.wdt) + static_cast<int64_t>(c0) * kR);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int v = lane + 32 * i;
            *reinterpret_cast<uint4*>(&S.wd[v >> 2][(v & 3) * 8]) = __ldg(src + v);
        }
    }
    // A rows of a 16-token tile: the first 128 B of each x_dbl row, 4 x 16 B per lane (rows past the end: clamped)
    auto fetch_a = [&](int mt, uint4 (&v)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = lane + 32 * i;
            const int row = min(row0 + mt * 16 + (e >> 3), rows_per_group - 1);
            v[i] = __ldg(reinterpret_cast<const uint4*>(G.x_dbl + static_cast<int64_t>(row) * kE) + (e & 7));
        }
    };
    pdl_launch_dependents();
    pdl_wait();                                  // W_dt is staged; x_dbl (kernel P's output) is read from here on
    uint4 av[4];
    fetch_a(0, av);
    float2 bias[8];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        bias[nt] = G.dt_bias ? __ldg(reinterpret_cast<const float2*>(G.dt_bias + c0 + nt * 8 + 2 * q)) : make_float2(0.f, 0.f);
        bias[nt].x *= kLog2e; bias[nt].y *= kLog2e;
    }
    __syncwarp();
    uint32_t bw[8][4];             // B fragments of k-steps 0 and 1 for the 8 channel tiles
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
        ldmatrix_x4(bw[nt][0], bw[nt][1], bw[nt][2], bw[nt][3], smem_u32(&S.wd[nt * 8 + (lane & 7)][(lane >> 3) * 8]));
#pragma unroll 1
    for (int mt = 0; mt < kDTok / 16; ++mt) {
        if (row0 + mt * 16 >= rows_per_group) break;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = lane + 32 * i;
            *reinterpret_cast<uint4*>(&S.at[e >> 3][(e & 7) * 8]) = av[i];
        }
        __syncwarp();
        if (mt + 1 < kDTok / 16) fetch_a(mt + 1, av);           // in flight while this tile is processed
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            ldmatrix_x4(ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], smem_u32(&S.at[lane & 15][ks * 16 + (lane >> 4) * 8]));
            ldmatrix_x4(al[ks][0], al[ks][1], al[ks][2], al[ks][3], smem_u32(&S.at[lane & 15][32 + ks * 16 + (lane >> 4) * 8]));
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float dacc[4] = {0.f, 0.f, 0.f, 0.f};
            mma_bf16_16816(dacc, ah[0], bw[nt][0], bw[nt][1]);
            mma_bf16_16816(dacc, al[0], bw[nt][0], bw[nt][1]);
            mma_bf16_16816(dacc, ah[1], bw[nt][2], bw[nt][3]);
            mma_bf16_16816(dacc, al[1], bw[nt][2], bw[nt][3]);
            *reinterpret_cast<__half2*>(&S.stage[r][nt * 8 + 2 * q]) =
                __floats2half2_rn(softplus_scaled_p(fmaf(dacc[0], kLog2e, bias[nt].x)), softplus_scaled_p(fmaf(dacc[1], kLog2e, bias[nt].y)));
            *reinterpret_cast<__half2*>(&S.stage[r + 8][nt * 8 + 2 * q]) =
                __floats2half2_rn(softplus_scaled_p(fmaf(dacc[2], kLog2e, bias[nt].x)), softplus_scaled_p(fmaf(dacc[3], kLog2e, bias[nt].y)));
        }
        __syncwarp();
        // row-order copy-out: 16 rows x 128 B = 128 16-byte vectors, 4 per lane, full 128-byte lines per 8 lanes
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int v = lane + 32 * i, rr = v >> 3, seg = v & 7;
            const int row = row0 + mt * 16 + rr;
            if (row < rows_per_group)
                *reinterpret_cast<uint4*>(G.delta + static_cast<int64_t>(row) * D + c0 + seg * 8) =
                    *reinterpret_cast<const uint4*>(&S.stage[rr][seg * 8]);
        }
        __syncwarp();                                           // stage / at are rewritten by the next tile
    }
}

// ------------------------------------------------------------------------------------------------------
// Kernel S: dt_proj + softplus + selective scan + D skip + SiLU(z) gate
// ------------------------------------------------------------------------------------------------------
// x_dbl row (256 B, written by kernel P): [dt_low hi: 32 bf16 | dt_low lo: 32 bf16 | B: 16 f32 | C: 16 f32]
// (dt_low = hi + lo to ~16 mantissa bits; the pair feeds the dt_proj MMA without any conversion here).
//
// One WARP per (sequence, 64 channels); lane owns channels c0+lane and c0+32+lane, 2 x 16 states in registers
// as packed fp32 pairs.  Two channels per lane halve the shared-memory (MIO) traffic for B/C per MUFU op and
// give 32 independent ex2 per token; at the BASELINE shape (1536 warp-units) every unit is resident in a single wave
// on 148 SMs (12 warps per SM at 168 registers) and the XU pipe runs at 69 % (profiles/r01_ncu_m1_scan.txt).
#ifndef DM_CPL1_MINB
#define DM_CPL1_MINB 20
#endif
// CPL = channels per lane (2 for the big shapes, 1 when there are too few warp-units to fill the SMs otherwise)
template <typename T, int CPL, bool kDelta = false> struct ScanSmem {
    static constexpr int kSC = 32 * CPL;     // channels per scan warp
    float xd[2][kCH][kE];                    // x_dbl chunk, rows as above
    T us[2][kCH][kSC];
    T zs[2][kCH][kSC];
    float ds[kCH][kSC + 4];                  // delta_raw tile (token, channel-in-warp)
    int rows[2][kCH];                        // byte offset of each scanned token's output row
    __nv_bfloat16 wdt[sizeof(T) == 4 ? 2 : 1][kSC][kR + 8];   // W_dt slice (hi [, lo]); 80-byte rows: ldmatrix conflict-free
};
// delta handed over by kernel P (fp16, softplus already applied): no dt_proj MMA, no W_dt slice, no softplus here
template <typename T, int CPL> struct ScanSmem<T, CPL, true> {
    static constexpr int kSC = 32 * CPL;
    float xd[2][kCH][kE];
    T us[2][kCH][kSC];
    T zs[2][kCH][kSC];
    __half dts[2][kCH][kSC];
    int rows[2][kCH];
};

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// softplus with torch's threshold (identity above 20), argument already scaled: s = x * log2(e).
// ln(1 + e^x) = ln2 * lg2(1 + 2^s): two MUFU + 3 FMA-pipe instructions.  1 + 2^s rounds with absolute error <= 6e-8,
// i.e. dt carries an ABSOLUTE error <= 4e-8 -- immaterial next to dt's bf16 / tf32 inputs, and for dt that small the
// state update dt*B*u and the decay 1 - dt*|A| are unaffected at fp32 resolution.
__device__ __forceinline__ float softplus_scaled(float s) {
    const float l = lg2_approx(1.0f + ex2_approx(s));
    return 0.6931471805599453f * (s > 28.853900817779268f ? s : l);
}
// 2^x for a packed pair of non-positive arguments on the FMA / ALU pipes instead of the MUFU: round-to-nearest split
// x = n + f (magic-number add), degree-5 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 2.3e-7 in
// fp32, the same 2 ulp class as ex2.approx), exponent inserted with an integer add.  The scan is bound by the MUFU
// pipe (16 lanes/clk/SM) while the FMA pipe idles at ~20 %, so a fixed share of the 16 decays per token goes here.
__device__ __forceinline__ uint64_t exp2_poly2(uint64_t x2) {
    float x0, x1;
    unpack2(x2, x0, x1);
    x2 = pack2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
    const uint64_t magic = pack2(12582912.0f, 12582912.0f), nmagic = pack2(-12582912.0f, -12582912.0f);
    const uint64_t t = add2(x2, magic);                       // low mantissa bits of t = round(x)
    const uint64_t n = add2(t, nmagic);
    const uint64_t f = fma2(n, pack2(-1.0f, -1.0f), x2);      // f = x - n in [-0.5, 0.5]
    uint64_t p = pack2(0.0013266970636323094f, 0.0013266970636323094f);
    p = fma2(p, f, pack2(0.009675459936261177f, 0.009675459936261177f));
    p = fma2(p, f, pack2(0.05550742521882057f, 0.05550742521882057f));
    p = fma2(p, f, pack2(0.24022121727466583f, 0.24022121727466583f));
    p = fma2(p, f, pack2(0.6931469440460205f, 0.6931469440460205f));
    p = fma2(p, f, pack2(1.0000001192092896f, 1.0000001192092896f));
    float p0, p1, t0, t1;
    unpack2(p, p0, p1);
    unpack2(t, t0, t1);
    return pack2(__int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23)),
                 __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23)));
}
#ifndef DM_POLY_PAIRS
#define DM_POLY_PAIRS 1       // of the 8 state pairs per (token, channel): how many decays are evaluated by exp2_poly2
                              // (r02, delta hand-over scan, B200: 0 -> 108.6 / 675 us at B16 L196 / B32 L784; 1 -> 106.5 / 652;
                              //  2 -> 110.6 / 744; 3 -> 121 / 852: one pair is what the idle FMA issue slots absorb)
#endif

// non-volatile MMA: lets ptxas interleave independent accumulator chains
__device__ __forceinline__ void mma_nv(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_u(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}

#include "dm_convx_tc.cuh"      // kernel P3 (tcgen05 conv + x_proj): needs the packed fp32x2 helpers above

#ifdef DM_SCAN_TRACE
// debugging aid (tools/scan_trace.py): per warp-unit {smid, warpid, start clock, end clock}
__device__ unsigned long long g_scan_trace[4 * 8192];
#endif

// Scheduling.  kDyn = false: one CTA (= one warp) per warp-unit, the whole sequence (small problems, fp32).
// kDyn = true: PERSISTENT warps (grid = resident warp slots) consume a READY QUEUE in the caller's workspace.  A work
// item is one SEGMENT (seg_chunks chunks of 8 tokens) of one warp-unit.  Consumer tickets (atomic head) below n_units
// are the first segments, ready from the start; ticket n_units + j is whatever the j-th hand-over pushes: a warp that
// finishes a segment stores the recurrence state (32 lanes x 2 channels x 16 states, 4 KB), fences, takes a producer
// slot (atomic tail) and publishes (unit, next segment) there; the consumer polling that slot picks the unit up at
// once.  Items therefore start only when their input state exists (nothing ever waits inside an item), successors go
// to warps in COMPLETION order (no head-of-line blocking behind slow units), and the count of pushes provably covers
// every ticket below n_items, so no consumer waits forever whatever the grid size.
// Why: a static one-unit-per-warp launch leaves the sub-partitions of an SM with 2 or 3 warps (1536 units over 592
// SMSPs): warps of a 3-warp SMSP live 114..137 us while 2-warp SMSPs are idle after 91 us (tools/scan_trace.py), and
// a second wave (3072 units) is scheduled just as unevenly.
// Workspace (ints): [0] head, [1] tail, [2] warps done, [16 + j] queue slot j (0 = empty, else 1 + seg * n_units +
// unit), capacity kSchedMaxSegs * n_units; states at byte offset sched_state_offset(n_units), 4 KB per unit as
// [16 pairs][32 lanes] x 8 B.  Consumers clear the slot they took and the last warp out clears the counters, so the
// workspace is zeroed once at allocation and is re-armed after every launch (CUDA-graph replay included).
constexpr int kSchedMaxSegs = 64;
__host__ __device__ inline size_t sched_state_offset(int n_units) {
    return (static_cast<size_t>(64) + static_cast<size_t>(4) * kSchedMaxSegs * n_units + 255) / 256 * 256;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// kGated: z already holds silu(z) (dm_mamba1_args.z_is_gated) -- 18 instead of 19 MUFU per (token, channel)
template <typename T, int CPL, bool kDyn, bool kSave, bool kGated = false, bool kDelta = false>
__global__ void __launch_bounds__(32, CPL == 2 ? 12 : DM_CPL1_MINB) m1_scan_kernel(const __grid_constant__ M1P p, int n_units) {
#ifdef DM_SCAN_TRACE
    unsigned trace_sm, trace_warp;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(trace_sm));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(trace_warp));
    unsigned long long trace_t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace_t0));
#endif
    constexpr bool kSplit = sizeof(T) == 4;
    constexpr int kSC = 32 * CPL;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x;
    ScanSmem<T, CPL, kDelta>& S = *reinterpret_cast<ScanSmem<T, CPL, kDelta>*>(smem_raw);
    const int D = p.D, L = p.L;
    const int slices = D / kSC;
    const int n_chunks = (L + kCH - 1) / kCH;
    const bool token_order = p.out_order == DM_OUT_TOKEN_ORDER;
    int* const sched = p.sched;
    const int n_items = kDyn ? n_units * p.n_segs : n_units;
    // consumer ticket -> work item (dynamic), or the CTA's own unit (static)
    auto next_item = [&]() -> int {
        int t = 0;
        if (lane == 0) t = atomicAdd(sched, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_items) return n_items;
        if (t < n_units) return t;                               // first segments: ready from the start
        int* slot = sched + 16 + (t - n_units);
        int v;
        while ((v = ld_acquire_gpu(slot)) == 0) __nanosleep(32);
        __syncwarp();                                            // every lane has seen the entry
        if (lane == 0) st_relaxed_gpu(slot, 0);                  // the slot is ours alone: leave it empty for the next launch
        return v - 1;
    };
    int item = blockIdx.x;
    if constexpr (kDyn) {
        pdl_wait();                              // the ticket counters and every input belong to the predecessors
        item = next_item();
    }
  while (item < n_items) {
    int unit = item, seg = 0;
    if constexpr (kDyn) {
        seg = item / n_units;
        unit = item - seg * n_units;
    }
    const int c_begin = kDyn ? seg * p.seg_chunks : 0;
    const int c_end = kDyn ? min(n_chunks, c_begin + p.seg_chunks) : n_chunks;

    const int cs = unit % slices, seq = unit / slices;
    const int k = seq % p.K, b = (seq / p.K) % p.B, g = seq / (p.K * p.B);
    const M1G& G = p.g[g];
    const int c0 = cs * kSC;
    const int32_t* ord = dir_order(p, k);
    const int64_t seq_in_group = static_cast<int64_t>(b) * p.K + k;
    const T* u_seq = static_cast<const T*>(G.u) + seq_in_group * L * D + c0;
    const T* z_base = static_cast<const T*>(G.xz) + static_cast<int64_t>(b) * G.xz_bs + D + c0;
    const float* xd_seq = G.x_dbl + seq_in_group * L * kE;
    const __half* dt_seq = kDelta ? G.delta + seq_in_group * L * D + c0 : nullptr;
    char* out_lane = reinterpret_cast<char*>(static_cast<T*>(G.out) + static_cast<int64_t>(b) * G.out_bs +
                                             static_cast<int64_t>(k) * G.out_ds + c0 + lane);
    const int out_ts32 = static_cast<int>(G.out_ts * sizeof(T));

    // ---- W_dt slice -> shared (bf16 hi [, lo]) ----
    if constexpr (kDelta) {
        // nothing to stage: delta arrives per chunk with u and z
    } else if constexpr (!kSplit) {
        // 16-byte cp.async segments (all in flight at once; completion rides on the first chunk's commit group)
        const char* Wdt = reinterpret_cast<const char*>(static_cast<const T*>(G.wdt) + static_cast<int64_t>(c0) * kR);
#pragma unroll
        for (int i = 0; i < kSC * kR * 2 / 16 / 32; ++i) {
            const int sgm = lane + 32 * i, row = sgm >> 2, part = sgm & 3;       // 4 x 16 B per 32-element row
            cp_async16(smem_u32(&S.wdt[0][row][part * 8]), Wdt + sgm * 16);
        }
    } else {
        const T* Wdt = static_cast<const T*>(G.wdt) + static_cast<int64_t>(c0) * kR;
        for (int i = lane; i < kSC * kR / 2; i += 32) {
            const int row = i / (kR / 2), col = (i % (kR / 2)) * 2;
            uint32_t hi, lo = 0;
            if constexpr (kSplit) {
                const float2 w = __ldg(reinterpret_cast<const float2*>(Wdt + row * kR + col));
                split_bf16(w.x, w.y, hi, lo);
                *reinterpret_cast<uint32_t*>(&S.wdt[kSplit ? 1 : 0][row][col]) = lo;
            } else {
                hi = __ldg(reinterpret_cast<const uint32_t*>(Wdt + row * kR + col));
            }
            *reinterpret_cast<uint32_t*>(&S.wdt[0][row][col]) = hi;
        }
    }

    // per-channel constants: A*log2(e) as 8 packed pairs per channel
    uint64_t A2[CPL][kN / 2];
    float dtb[CPL], Dc[CPL];
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch) {
        const int c = c0 + ch * 32 + lane;
#pragma unroll
        for (int n = 0; n < kN; n += 4) {
            float4 t = __ldg(reinterpret_cast<const float4*>(G.A + static_cast<int64_t>(c) * kN + n));
            A2[ch][n / 2] = pack2(t.x * kLog2e, t.y * kLog2e);
            A2[ch][n / 2 + 1] = pack2(t.z * kLog2e, t.w * kLog2e);
        }
        dtb[ch] = G.dt_bias ? __ldg(G.dt_bias + c) * kLog2e : 0.f;       // pre-scaled: softplus_scaled takes x * log2(e)
        Dc[ch] = G.D ? __ldg(G.D + c) : 0.f;
    }

    // ---- staging of one chunk (cp.async, double buffered).  Rows past the end of the sequence are clamped to
    //      the last valid row, so the tail chunk needs no predicates (the duplicates are never consumed). ----
    constexpr int kSegU = kSC * sizeof(T) / 16;       // 16-byte segments per (token, 64 channels) row: 8 / 16
    auto prefetch = [&](int ci) {
        const int buf = (ci - c_begin) & 1, j0 = ci * kCH;
        {
            const uint32_t xdst = smem_u32(&S.xd[buf][0][0]) + lane * 16;
            const int part = lane & 15;
#pragma unroll
            for (int i = 0; i < kCH / 2; ++i) {
                const int j = min(j0 + (lane >> 4) + 2 * i, L - 1);
                cp_async16(xdst + i * 512, reinterpret_cast<const char*>(xd_seq + static_cast<int64_t>(j) * kE) + part * 16);
            }
        }
        const uint32_t udst = smem_u32(&S.us[buf][0][0]), zdst = smem_u32(&S.zs[buf][0][0]);
#pragma unroll
        for (int i = 0; i < kSegU * kCH / 32; ++i) {
            const int s = lane + 32 * i;
            const int r = s / kSegU, part = s % kSegU;
            const int j = min(j0 + r, L - 1);
            const int src = ord ? __ldg(ord + j) : j;
            if (part == 0) S.rows[buf][r] = (token_order ? src : j) * out_ts32;
            cp_async16(udst + s * 16, reinterpret_cast<const char*>(u_seq + static_cast<int64_t>(j) * D) + part * 16);
            cp_async16(zdst + s * 16, reinterpret_cast<const char*>(z_base + static_cast<int64_t>(src) * G.xz_ts) + part * 16);
        }
        if constexpr (kDelta) {
            constexpr int kSegD = kSC * 2 / 16;             // 16-byte segments per (token, channels-of-the-warp) fp16 row
            const uint32_t ddst = smem_u32(&S.dts[buf][0][0]);
#pragma unroll
            for (int i = 0; i < kSegD * kCH / 32; ++i) {
                const int s = lane + 32 * i;
                const int r = s / kSegD, part = s % kSegD;
                const int j = min(j0 + r, L - 1);
                cp_async16(ddst + s * 16, reinterpret_cast<const char*>(dt_seq + static_cast<int64_t>(j) * D) + part * 16);
            }
        }
        cp_async_commit();
    };

    uint64_t h[CPL][kN / 2];
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch)
#pragma unroll
        for (int n = 0; n < kN / 2; ++n) h[ch][n] = 0ull;      // bit pattern of (0.f, 0.f)
    uint64_t* const h_ws = kDyn ? reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(sched) + sched_state_offset(n_units)) +
                                      static_cast<size_t>(unit) * (CPL * (kN / 2) * 32) + lane
                                : nullptr;

    // one token of the recurrence for both channels of the lane
    auto token = [&](int buf, int jj, const float (&dtv)[CPL]) {
        const ulonglong2* bc = reinterpret_cast<const ulonglong2*>(&S.xd[buf][jj][kR]);
        ulonglong2 Bq[4], Cq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { Bq[q] = bc[q]; Cq[q] = bc[4 + q]; }
        const int row_off = S.rows[buf][jj];
        uint64_t ysum[CPL];
        float uu_[CPL], zz_[CPL];
#pragma unroll
        for (int ch = 0; ch < CPL; ++ch) {
            const float uu = to_f32<T>(S.us[buf][jj][ch * 32 + lane]);
            const float zz = to_f32<T>(S.zs[buf][jj][ch * 32 + lane]);
            const float dt = dtv[ch], dtu = dt * uu;
            const uint64_t dt2 = pack2(dt, dt), dtu2 = pack2(dtu, dtu);
            uint64_t y0 = 0ull, y1 = 0ull;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                {
                    uint64_t dA;
                    if (2 * q < DM_POLY_PAIRS) {
                        dA = exp2_poly2(mul2(dt2, A2[ch][2 * q]));
                    } else {
                        float a0, a1;
                        unpack2(mul2(dt2, A2[ch][2 * q]), a0, a1);
                        dA = pack2(ex2_approx(a0), ex2_approx(a1));
                    }
                    h[ch][2 * q] = fma2(dA, h[ch][2 * q], mul2(dtu2, Bq[q].x));
                    y0 = fma2(h[ch][2 * q], Cq[q].x, y0);
                }
                {
                    uint64_t dA;
                    if (2 * q + 1 < DM_POLY_PAIRS) {
                        dA = exp2_poly2(mul2(dt2, A2[ch][2 * q + 1]));
                    } else {
                        float a0, a1;
                        unpack2(mul2(dt2, A2[ch][2 * q + 1]), a0, a1);
                        dA = pack2(ex2_approx(a0), ex2_approx(a1));
                    }
                    h[ch][2 * q + 1] = fma2(dA, h[ch][2 * q + 1], mul2(dtu2, Bq[q].y));
                    y1 = fma2(h[ch][2 * q + 1], Cq[q].y, y1);
                }
            }
            ysum[ch] = add2(y0, y1);                 // (y_even-pairs + y_odd-pairs): two partial sums per channel
            uu_[ch] = uu;
            zz_[ch] = zz;
        }
        if constexpr (CPL == 2 && !kSplit) {
            // both channels of the lane at once on the packed-fp32 pipe: y = sum + D u ; out = y z (0.5 + 0.5 tanh(z/2))
            float a0, a1, b0, b1;
            unpack2(ysum[0], a0, a1);
            unpack2(ysum[1], b0, b1);
            const uint64_t u2 = pack2(uu_[0], uu_[1]), z2 = pack2(zz_[0], zz_[1]);
            const uint64_t y2 = fma2(pack2(Dc[0], Dc[1]), u2, pack2(a0 + a1, b0 + b1));
            float o0, o1;
            if constexpr (kGated) {
                unpack2(mul2(y2, z2), o0, o1);
            } else {
                const uint64_t half2 = pack2(0.5f, 0.5f);
                float hz0, hz1;
                unpack2(mul2(z2, half2), hz0, hz1);
                const uint64_t sg = fma2(pack2(tanh_approx(hz0), tanh_approx(hz1)), half2, half2);
                unpack2(mul2(mul2(y2, z2), sg), o0, o1);
            }
            *reinterpret_cast<T*>(out_lane + row_off) = from_f32<T>(o0);
            *reinterpret_cast<T*>(out_lane + row_off + 32 * static_cast<int>(sizeof(T))) = from_f32<T>(o1);
        } else {
#pragma unroll
            for (int ch = 0; ch < CPL; ++ch) {
                float ya, yb;
                unpack2(ysum[ch], ya, yb);
                const float y = fmaf(Dc[ch], uu_[ch], ya + yb);
                const float o = y * (kGated ? zz_[ch] : (kSplit ? silu_fast(zz_[ch]) : silu_tanh(zz_[ch])));   // bf16 output: 1 MUFU (tanh) is enough
                *reinterpret_cast<T*>(out_lane + row_off + ch * 32 * static_cast<int>(sizeof(T))) = from_f32<T>(o);
            }
        }
    };

    // training: state BEFORE every save_every-th token -> chunk_states[(seq, j / save_every, channel, 0..15)] (the
    // backward's checkpoints; it then skips its own forward sweep)
    // (kSave is a template flag so that the inference kernels carry none of this)
    float* const st_base = kSave
        ? G.chunk_states + ((seq_in_group * ((L + p.save_every - 1) / max(p.save_every, 1))) * D + c0 + lane) * kN
        : nullptr;
    auto save_state = [&](int j) {
        float* dst = st_base + static_cast<int64_t>(j / p.save_every) * D * kN;
#pragma unroll
        for (int ch = 0; ch < CPL; ++ch)
#pragma unroll
            for (int n = 0; n < kN / 2; n += 2) {
                float a0, a1, a2, a3;
                unpack2(h[ch][n], a0, a1);
                unpack2(h[ch][n + 1], a2, a3);
                *reinterpret_cast<float4*>(dst + ch * 32 * kN + 2 * n) = make_float4(a0, a1, a2, a3);
            }
    };

    if constexpr (!kDyn) pdl_wait();             // constants are loaded; u / z / x_dbl / delta come from the predecessor kernels
    prefetch(c_begin);
    if constexpr (kDyn) {
        if (seg > 0) {                               // state left by the item that ran the previous segment of this unit
#pragma unroll
            for (int ch = 0; ch < CPL; ++ch)
#pragma unroll
                for (int n = 0; n < kN / 2; ++n) h[ch][n] = __ldcg(reinterpret_cast<const unsigned long long*>(h_ws) + (ch * (kN / 2) + n) * 32);
        }
    }
    __syncwarp();                                    // W_dt slice visible to every lane
    for (int ci = c_begin; ci < c_end; ++ci) {
        const int buf = (ci - c_begin) & 1, j0 = ci * kCH;
        if (ci + 1 < c_end) {
            prefetch(ci + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();

        // ---- delta_raw tile (64 channels x 8 tokens) = W_dt slice (64 x 32) . dt_low^T (32 x 8) on mma.sync m16n8k16:
        //      A = 16 channels x 16 k from shared (ldmatrix), B = the chunk's dt_low rows straight from the staged
        //      x_dbl (token = lane / 4, k pair = lane % 4: exactly the B fragment), hi + lo halves of dt_low ----
        if constexpr (!kDelta) {
            const int r = lane >> 2, q = lane & 3;
            const uint32_t* row = reinterpret_cast<const uint32_t*>(&S.xd[buf][r][0]);   // 16 hi words, 16 lo words
            uint32_t b_hi[2][2], b_lo[2][2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                b_hi[ks][0] = row[ks * 8 + q]; b_hi[ks][1] = row[ks * 8 + 4 + q];
                b_lo[ks][0] = row[16 + ks * 8 + q]; b_lo[ks][1] = row[16 + ks * 8 + 4 + q];
            }
            constexpr int kMT = kSC / 16;                    // 16-channel tiles: 4 (2 ch/lane) or 2
            float dacc[kMT][4];
#pragma unroll
            for (int t = 0; t < kMT; ++t)
#pragma unroll
                for (int i = 0; i < 4; ++i) dacc[t][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t aw[kMT][4];
#pragma unroll
                for (int t = 0; t < kMT; ++t)
                    ldmatrix_x4_u(aw[t], smem_u32(&S.wdt[0][t * 16 + (lane & 15)][ks * 16 + (lane >> 4) * 8]));
#pragma unroll
                for (int t = 0; t < kMT; ++t) mma_nv(dacc[t], aw[t], b_hi[ks][0], b_hi[ks][1]);
#pragma unroll
                for (int t = 0; t < kMT; ++t) mma_nv(dacc[t], aw[t], b_lo[ks][0], b_lo[ks][1]);
                if constexpr (kSplit) {
#pragma unroll
                    for (int t = 0; t < kMT; ++t)
                        ldmatrix_x4_u(aw[t], smem_u32(&S.wdt[1][t * 16 + (lane & 15)][ks * 16 + (lane >> 4) * 8]));
#pragma unroll
                    for (int t = 0; t < kMT; ++t) mma_nv(dacc[t], aw[t], b_hi[ks][0], b_hi[ks][1]);
                }
            }
            // accumulator (channel = 16 t + r [+ 8], tokens 2q, 2q + 1) -> ds[token][channel]; the 68-word pitch keeps
            // the 32 lanes of each store on 32 different banks
#pragma unroll
            for (int t = 0; t < kMT; ++t) {
                S.ds[2 * q][t * 16 + r] = dacc[t][0];
                S.ds[2 * q + 1][t * 16 + r] = dacc[t][1];
                S.ds[2 * q][t * 16 + r + 8] = dacc[t][2];
                S.ds[2 * q + 1][t * 16 + r + 8] = dacc[t][3];
            }
        }
        __syncwarp();

        if (j0 + kCH <= L) {
            // full chunk: softplus for all 8 tokens first (16 independent MUFU chains, off the recurrence's
            // critical path), then straight-line recurrence without predicates
            float dtv[kCH][CPL];
#pragma unroll
            for (int jj = 0; jj < kCH; ++jj)
#pragma unroll
                for (int ch = 0; ch < CPL; ++ch) {
                    if constexpr (kDelta) dtv[jj][ch] = __half2float(S.dts[buf][jj][ch * 32 + lane]);
                    else dtv[jj][ch] = softplus_scaled(fmaf(S.ds[jj][ch * 32 + lane], kLog2e, dtb[ch]));
                }
#pragma unroll
            for (int jj = 0; jj < kCH; ++jj) {
                if constexpr (kSave) {
                    if ((jj % 4) == 0 && ((j0 + jj) % p.save_every) == 0) save_state(j0 + jj);
                }
                token(buf, jj, dtv[jj]);
            }
        } else {
            const int nrows = L - j0;
#pragma unroll 1
            for (int jj = 0; jj < nrows; ++jj) {
                float dtv[CPL];
#pragma unroll
                for (int ch = 0; ch < CPL; ++ch) {
                    if constexpr (kDelta) dtv[ch] = __half2float(S.dts[buf][jj][ch * 32 + lane]);
                    else dtv[ch] = softplus_scaled(fmaf(S.ds[jj][ch * 32 + lane], kLog2e, dtb[ch]));
                }
                if constexpr (kSave) {
                    if (((j0 + jj) % p.save_every) == 0) save_state(j0 + jj);
                }
                token(buf, jj, dtv);
            }
        }
        __syncwarp();
    }
    if constexpr (kDyn) {
        if (seg + 1 < p.n_segs) {                    // hand the state to whoever holds the next segment
#pragma unroll
            for (int ch = 0; ch < CPL; ++ch)
#pragma unroll
                for (int n = 0; n < kN / 2; ++n)
                    __stcg(reinterpret_cast<unsigned long long*>(h_ws) + (ch * (kN / 2) + n) * 32, h[ch][n]);
            __syncwarp();                            // every lane's state stores happen-before lane 0's fence
            if (lane == 0) {
                __threadfence();
                const int slot = atomicAdd(sched + 1, 1);
                st_relaxed_gpu(sched + 16 + slot, 1 + (seg + 1) * n_units + unit);
            }
        }
    }
#ifdef DM_SCAN_TRACE
    if (lane == 0 && item < 8192) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        g_scan_trace[4 * item] = trace_sm; g_scan_trace[4 * item + 1] = trace_warp;
        g_scan_trace[4 * item + 2] = trace_t0; g_scan_trace[4 * item + 3] = t1;
        trace_t0 = t1;
    }
#endif
    if constexpr (!kDyn) break;
    // next ticket only now (no look-ahead: a ticket must be held by a warp that is free to poll)
    item = next_item();
  }
    if constexpr (kDyn) {
        // every warp checks out; the last one clears the counters (queue slots were cleared by their consumers)
        if (lane == 0 && atomicAdd(sched + 2, 1) == static_cast<int>(gridDim.x) - 1) {
            sched[0] = 0;
            sched[1] = 0;
            sched[2] = 0;
        }
    }
}

template <typename T>
int launch_m1(const M1P& p_in, int phases, cudaStream_t stream, size_t sched_bytes) {
    M1P p = p_in;
    const int n_seq = p.n_groups * p.B * p.K;
    const bool split = sizeof(T) == 4;
    int dev = 0, n_sm = 0;
    if (int e = current_device(&dev, &n_sm); e != DM_OK) return e;
    const size_t ld = static_cast<size_t>(p.D) + 8;
    {   // delta is produced (m1_delta_kernel, after kernel P) for bf16 activations and consumed only by the inference scan:
        // anywhere else the scan evaluates dt_proj + softplus itself, whatever the caller passed
        bool all = !split && p.D % 256 == 0 && p.save_every == 0 && p.z_gated == 0;
        for (int g = 0; g < p.n_groups; ++g) all = all && p.g[g].delta != nullptr;
        if (!all)
            for (int g = 0; g < p.n_groups; ++g) p.g[g].delta = nullptr;
    }
    // kernel P
    if (phases & 1) {
        const size_t bytes2 = (static_cast<size_t>(kE) + 2 * (kTP2 + 3)) * ld * 2 + static_cast<size_t>(p.K) * p.L * 4;
        // bf16, d_inner 1024: tensor-core kernel P3 (tcgen05, 128-token tiles; DM_CONVX=legacy selects the mma.sync
        // persistent kernel below, which also serves shapes whose row offsets do not fit 32 bits)
        static const int convx_legacy = [] { const char* e = getenv("DM_CONVX"); return (e && e[0] == 'l') ? 1 : 0; }();
        bool use_tc = !split && p.D == 1024 && !convx_legacy && p.L >= 11;    // (>= 11 tokens: at most one sequence start per 8-row window)
        for (int g = 0; g < p.n_groups && use_tc; ++g)
            use_tc = static_cast<long long>(p.B) * p.g[g].xz_bs < (1ll << 31) && p.g[g].xz_ts % 8 == 0 && p.g[g].xz_bs % 8 == 0;
        if (use_tc) {
            if constexpr (sizeof(T) == 2) {
                static PerDeviceOnce tcfg;
                if (!tcfg.done(dev)) {
                    DM_CUDA_TRY(cudaFuncSetAttribute(m1_conv_xproj_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p3::kSmemBytes));
                    DM_CUDA_TRY(cudaFuncSetAttribute(m1_conv_xproj_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p3::kSmemBytes));
                    tcfg.set(dev);
                }
                const int rows = p.B * p.K * p.L;
                const int tiles_per_group = (rows + p3::kTile - 1) / p3::kTile;
                const int n_tiles = tiles_per_group * p.n_groups;
                static const int split_k = env_int("DM_CONVX_SPLIT", 1);
                if (split_k && 2 * n_tiles <= n_sm) {
                    // small batches: two CTAs (a cluster) per tile, the channel slices split between them (batch 8: 74 tiles
                    // would leave half of the 148 SMs idle)
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(2 * n_tiles);
                    cfg.blockDim = dim3(p3::kThreads);
                    cfg.dynamicSmemBytes = p3::kSmemBytes;
                    cfg.stream = stream;
                    cudaLaunchAttribute attr[2];
                    attr[0].id = cudaLaunchAttributeClusterDimension;
                    attr[0].val.clusterDim.x = 2;
                    attr[0].val.clusterDim.y = 1;
                    attr[0].val.clusterDim.z = 1;
                    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    attr[1].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = attr;
                    cfg.numAttrs = (pdl_mask() & kPdlConvX) ? 2 : 1;
                    DM_CUDA_TRY(cudaLaunchKernelEx(&cfg, m1_conv_xproj_tc<true>, p, rows, tiles_per_group));
                } else {
                    DM_CUDA_TRY(launch_pdl(kPdlConvX, m1_conv_xproj_tc<false>, dim3(n_tiles < n_sm ? n_tiles : n_sm), dim3(p3::kThreads), p3::kSmemBytes, stream, p, rows, tiles_per_group));
                }
            }
        } else if (!split && p.D == 1024 && bytes2 <= 227 * 1024) {                 // persistent mma.sync kernel, W_x resident in shared memory
            static PerDeviceOnce cfg;
            if (!cfg.done(dev)) {
                DM_CUDA_TRY(cudaFuncSetAttribute(m1_conv_xproj_persistent<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 227 * 1024));
                cfg.set(dev);
            }
            const int n_tiles = n_seq * ((p.L + kTP2 - 1) / kTP2);
            const int grid = n_tiles < n_sm ? n_tiles : n_sm;
            m1_conv_xproj_persistent<1024><<<grid, kP2Threads, bytes2, stream>>>(p, n_tiles);
        } else {
            const size_t smem = static_cast<size_t>(kTP) * (p.D + 8) * 2 * (split ? 2 : 1);
            const size_t red = static_cast<size_t>(8) * kTP * kE * 4;
            const size_t bytes = smem > red ? smem : red;
            // (cheap, and the size depends on d_inner: set it on every launch rather than caching per thread)
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_conv_xproj_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(bytes)));
            m1_conv_xproj_kernel<T><<<n_seq * p.tiles_per_seq, kPThreads, bytes, stream>>>(p);
        }
        DM_CUDA_TRY(cudaGetLastError());
        if constexpr (sizeof(T) == 2) {
            if (p.g[0].delta != nullptr) {                       // (all groups or none: normalised above)
                // 128 tokens per warp halve the per-warp prologue (W_dt staging, fragments, bias) -- taken when that still
                // leaves >= 4 CTAs per SM (headline shape: 592 CTAs = exactly one resident wave on 148 SMs; 50.2 -> 48.1 us for
                // conv + x_proj + delta), otherwise 64-token warps keep the SMs occupied (batch 8)
                const int rows = p.B * p.K * p.L;
                const int grid128 = p.n_groups * ((rows + 127) / 128) * (p.D / 256);
                if (grid128 >= 4 * n_sm) {
                    DM_CUDA_TRY(launch_pdl(kPdlDelta, m1_delta_kernel<128>, dim3(grid128), dim3(128), 0, stream, p, rows));
                } else {
                    const int grid = p.n_groups * ((rows + 63) / 64) * (p.D / 256);
                    DM_CUDA_TRY(launch_pdl(kPdlDelta, m1_delta_kernel<64>, dim3(grid), dim3(128), 0, stream, p, rows));
                }
                DM_CUDA_TRY(cudaGetLastError());
            }
        }
    }
    // kernel S
    if (phases & 2) {
        static PerDeviceOnce scfg;
        if (!scfg.done(dev)) {
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(ScanSmem<T, 2>))));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(ScanSmem<T, 2>))));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(ScanSmem<T, 1>))));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 1, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(ScanSmem<T, 2>))));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(ScanSmem<T, 1>))));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 1, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(ScanSmem<T, 2>))));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, false, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(ScanSmem<T, 2>))));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, true, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 1, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(ScanSmem<T, 1>))));
            DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 1, false, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            scfg.set(dev);
        }
        // two channels per lane (fewer shared-memory reads per MUFU op, 168 registers) when that still leaves >= 8
        // warps per SM; otherwise one channel per lane doubles the number of warp-units (small batches, config C5)
        const int units2 = n_seq * (p.D / 64);
        // experiment knobs, read once (thread-safe static initialisation)
        static const int force_cpl = env_int("DM_SCAN_CPL", 0);
        static const int force_sched = [] {              // "static" | "dynamic" (default: dynamic when a workspace is given)
            const char* e = getenv("DM_SCAN_SCHED");
            return e ? (e[0] == 's' ? 0 : 1) : 2;
        }();
        static const int force_seg = env_int("DM_SCAN_SEG", 0);   // chunks (of 8 tokens) per work item of the dynamic schedule
        const bool save = p.save_every > 0;       // training forward: checkpoints for the backward, static schedule
        const bool gated = p.z_gated != 0;
        bool delta_in = !split;                   // delta handed over by kernel P (all groups or none; bf16 only)
        for (int g = 0; g < p.n_groups; ++g) delta_in = delta_in && p.g[g].delta != nullptr;
        delta_in = delta_in && !save && !gated;
        if constexpr (sizeof(T) == 2) {
            static PerDeviceOnce dcfg;
            if (!dcfg.done(dev)) {
                DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, false, false, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 2, true, false, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_kernel<T, 1, false, false, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                dcfg.set(dev);
            }
        }
        if (save && gated) return DM_ERR_INVALID_ARG;                  // the backward needs the raw z
        if (save) {
            for (int g = 0; g < p.n_groups; ++g)
                if (p.g[g].chunk_states == nullptr) return DM_ERR_INVALID_ARG;      // all groups or none
            if (force_cpl == 2 || (force_cpl == 0 && units2 >= 8 * n_sm)) {
                DM_CUDA_TRY(launch_pdl(kPdlScan, m1_scan_kernel<T, 2, false, true>, dim3(units2), dim3(32), sizeof(ScanSmem<T, 2>), stream, p, units2));
            } else {
                const int units1 = n_seq * (p.D / 32);
                DM_CUDA_TRY(launch_pdl(kPdlScan, m1_scan_kernel<T, 1, false, true>, dim3(units1), dim3(32), sizeof(ScanSmem<T, 1>), stream, p, units1));
            }
        } else if (force_cpl == 2 || (force_cpl == 0 && units2 >= 8 * n_sm)) {
            const int n_chunks = (p.L + kCH - 1) / kCH;
            const int slots = 12 * n_sm;                                      // resident warps (168 registers)
            const bool have_ws = p.sched != nullptr && sched_bytes >= sched_state_offset(units2) + static_cast<size_t>(units2) * 4096;
            // Measured on B200 (profiles/r01_notes.md): with at most one warp-unit per resident slot every chain is in
            // flight from the start and a hand-over only adds ~3.5 us to it (static wins: 140 vs 147 us at 1536 units);
            // with more units than slots the queue keeps all 12 warps of every SM busy until the end (-4 .. -13 %).
            const bool dyn = force_sched == 1 || (force_sched == 2 && units2 > slots);
            if (have_ws && dyn && n_chunks >= 8) {
                M1P q = p;
                // ~5 work items per resident warp bounds the idle tail at the end; fewer, longer items = fewer hand-overs
                int segs = force_seg > 0 ? (n_chunks + force_seg - 1) / force_seg
                                         : static_cast<int>((5ll * slots + units2 - 1) / units2);
                if (segs < 2 && force_seg <= 0) segs = 2;
                if (segs > n_chunks / 4) segs = n_chunks / 4;
                if (segs < 1) segs = 1;
                if (segs > kSchedMaxSegs) segs = kSchedMaxSegs;
                q.seg_chunks = (n_chunks + segs - 1) / segs;
                q.n_segs = (n_chunks + q.seg_chunks - 1) / q.seg_chunks;
                const long long items = static_cast<long long>(units2) * q.n_segs;
                const int grid = items < slots ? static_cast<int>(items) : slots;
                bool done = false;
                if constexpr (sizeof(T) == 2) {
                    if (delta_in) {
                        DM_CUDA_TRY(launch_pdl(kPdlScan, m1_scan_kernel<T, 2, true, false, false, true>, dim3(grid), dim3(32), sizeof(ScanSmem<T, 2, true>), stream, q, units2));
                        done = true;
                    }
                }
                if (done) {
                } else if (gated) DM_CUDA_TRY(launch_pdl(kPdlScan, m1_scan_kernel<T, 2, true, false, true>, dim3(grid), dim3(32), sizeof(ScanSmem<T, 2>), stream, q, units2));
                else DM_CUDA_TRY(launch_pdl(kPdlScan, m1_scan_kernel<T, 2, true, false>, dim3(grid), dim3(32), sizeof(ScanSmem<T, 2>), stream, q, units2));
            } else {
                bool done = false;
                if constexpr (sizeof(T) == 2) {
                    if (delta_in) {
                        DM_CUDA_TRY(launch_pdl(kPdlScan, m1_scan_kernel<T, 2, false, false, false, true>, dim3(units2), dim3(32), sizeof(ScanSmem<T, 2, true>), stream, p, units2));
                        done = true;
                    }
                }
                if (done) {
                } else if (gated) DM_CUDA_TRY(launch_pdl(kPdlScan, m1_scan_kernel<T, 2, false, false, true>, dim3(units2), dim3(32), sizeof(ScanSmem<T, 2>), stream, p, units2));
                else DM_CUDA_TRY(launch_pdl(kPdlScan, m1_scan_kernel<T, 2, false, false>, dim3(units2), dim3(32), sizeof(ScanSmem<T, 2>), stream, p, units2));
            }
        } else {
            const int units1 = n_seq * (p.D / 32);
            bool done = false;
            if constexpr (sizeof(T) == 2) {
                if (delta_in) {
                    DM_CUDA_TRY(launch_pdl(kPdlScan, m1_scan_kernel<T, 1, false, false, false, true>, dim3(units1), dim3(32), sizeof(ScanSmem<T, 1, true>), stream, p, units1));
                    done = true;
                }
            }
            if (done) {
            } else if (gated) DM_CUDA_TRY(launch_pdl(kPdlScan, m1_scan_kernel<T, 1, false, false, true>, dim3(units1), dim3(32), sizeof(ScanSmem<T, 1>), stream, p, units1));
            else DM_CUDA_TRY(launch_pdl(kPdlScan, m1_scan_kernel<T, 1, false, false>, dim3(units1), dim3(32), sizeof(ScanSmem<T, 1>), stream, p, units1));
        }
        DM_CUDA_TRY(cudaGetLastError());
    }
    return DM_OK;
}

}  // namespace
}  // namespace dm

extern "C" int dm_mamba1_bwd_chunk_tokens(void);

static int m1_dispatch(const dm_mamba1_args* a, int phases, void* stream) {
    using namespace dm;
    if (a == nullptr) return DM_ERR_INVALID_ARG;
    if (a->batch <= 0 || a->n_dir <= 0 || a->seqlen <= 0 || a->n_groups <= 0 || a->n_groups > DM_MAX_GROUPS)
        return DM_ERR_INVALID_ARG;
    if (a->out_order != DM_OUT_SCAN_ORDER && a->out_order != DM_OUT_TOKEN_ORDER) return DM_ERR_INVALID_ARG;
    if (a->act_dtype != DM_F32 && a->act_dtype != DM_BF16) return DM_ERR_UNSUPPORTED;
    if (a->d_state != kN || a->d_conv != kW || a->dt_rank != kR) return DM_ERR_UNSUPPORTED;
    if (a->d_inner <= 0 || a->d_inner % 128 != 0) return DM_ERR_UNSUPPORTED;
    const size_t es = dtype_size(a->act_dtype);
    M1P p{};
    p.B = a->batch; p.K = a->n_dir; p.L = a->seqlen; p.D = a->d_inner;
    p.out_order = a->out_order; p.n_groups = a->n_groups;
    p.z_gated = a->z_is_gated;
    p.tiles_per_seq = (a->seqlen + kTP - 1) / kTP;
    p.order = a->order;
    for (int g = 0; g < a->n_groups; ++g) {
        const dm_mamba1_group& s = a->group[g];
        if (!s.xz || !s.out || !s.u || !s.x_dbl || !s.conv_weight || !s.x_proj_weight || !s.dt_proj_weight || !s.A)
            return DM_ERR_INVALID_ARG;
        if (!aligned16(s.xz) || !aligned16(s.out) || !aligned16(s.u) || !aligned16(s.x_dbl) ||
            !aligned16(s.conv_weight) || !aligned16(s.x_proj_weight) || !aligned16(s.dt_proj_weight) || !aligned16(s.A))
            return DM_ERR_INVALID_ARG;
        if ((s.xz_batch_stride * es) % 16 || (s.xz_token_stride * es) % 16) return DM_ERR_INVALID_ARG;
        if (s.xz_token_stride < 2 * a->d_inner) return DM_ERR_INVALID_ARG;
        M1G& d = p.g[g];
        d.xz = s.xz; d.xz_bs = s.xz_batch_stride; d.xz_ts = s.xz_token_stride;
        d.out = s.out; d.out_bs = s.out_batch_stride; d.out_ds = s.out_dir_stride; d.out_ts = s.out_token_stride;
        d.u = s.u; d.x_dbl = s.x_dbl;
        d.conv_w = s.conv_weight; d.conv_b = s.conv_bias; d.wx = s.x_proj_weight; d.wdt = s.dt_proj_weight;
        d.dt_bias = s.dt_bias; d.A = s.A; d.D = s.D;
        d.chunk_states = s.chunk_states;
        d.delta = static_cast<__half*>(s.delta);
        if (s.delta != nullptr && !aligned16(s.delta)) return DM_ERR_INVALID_ARG;
        if (s.chunk_states != nullptr) p.save_every = dm_mamba1_bwd_chunk_tokens();
    }
    if (a->sched_workspace != nullptr && !aligned16(a->sched_workspace)) return DM_ERR_INVALID_ARG;
    p.sched = static_cast<int*>(a->sched_workspace);
    const size_t sched_bytes = a->sched_workspace ? static_cast<size_t>(a->sched_workspace_bytes) : 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return a->act_dtype == DM_F32 ? launch_m1<float>(p, phases, st, sched_bytes)
                                  : launch_m1<__nv_bfloat16>(p, phases, st, sched_bytes);
}

extern "C" int dm_mamba1_scan_fwd(const dm_mamba1_args* a, void* stream) { return m1_dispatch(a, 3, stream); }

extern "C" int64_t dm_mamba1_sched_workspace_bytes(int32_t batch, int32_t n_dir, int32_t d_inner, int32_t n_groups) {
    if (batch <= 0 || n_dir <= 0 || d_inner <= 0 || n_groups <= 0) return 0;
    const long long units = static_cast<long long>(n_groups) * batch * n_dir * (d_inner / 64);
    if (units > (1ll << 24)) return 0;
    return static_cast<int64_t>(dm::sched_state_offset(static_cast<int>(units)) + static_cast<size_t>(units) * 4096);
}

#ifdef DM_SCAN_TRACE
extern "C" int dm_debug_scan_trace(unsigned long long* host_out, int n_units) {
    return cudaMemcpyFromSymbol(host_out, dm::g_scan_trace, sizeof(unsigned long long) * 4 * n_units) == cudaSuccess ? 0 : -4;
}
#endif

extern "C" int dm_mamba1_scan_phase(const dm_mamba1_args* a, int phase, void* stream) {
    if (phase != 1 && phase != 2) return DM_ERR_INVALID_ARG;
    return m1_dispatch(a, phase, stream);
}
