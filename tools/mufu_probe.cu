// MUFU (XU pipe) throughput probe for the Mamba-1 scan's instruction mix on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/mufu_probe tools/mufu_probe.cu ; ./tools/_bin/mufu_probe
// Question it answers: what MUFU issue rate can 1..4 warps per SM sub-partition sustain
//   k_mufu     : only ex2 (16 independent chains per lane)
//   k_mix      : the scan's inner mix per state pair: FMUL2 (dt*A) -> 2 x MUFU.EX2 -> FMUL2 (dtu*B) -> FFMA2 (h) -> FFMA2 (y),
//                B / C held in registers (no shared-memory traffic)
//   k_mix_lds  : same, B / C re-read from shared memory per token (8 x LDS.128, broadcast), as the real kernel does
// CTAs are 128 threads (one warp per sub-partition), W CTAs per SM => exactly W warps per sub-partition.
// Output: cycles per "token" (32 ex2 per lane) per warp and the resulting MUFU warp-instructions per clock per SMSP
// (the pipe's nominal peak is 1/8 = 0.125).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

constexpr int kTok = 2048;

__global__ void __launch_bounds__(128) k_mufu(float* out, long long* cyc, float seed) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = -seed * (i + 1 + threadIdx.x * 1e-3f);
    const long long t0 = clock64();
#pragma unroll 1
    for (int t = 0; t < kTok; ++t) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = ex2(x[i]) - 1.5f;      // FADD keeps the chain in range (extra FMA-pipe instr)
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 4 + (threadIdx.x >> 5)] = t1 - t0;
}

template <bool kLds>
__global__ void __launch_bounds__(128, 4) k_mix(float* out, long long* cyc, const float* in) {
    __shared__ __align__(16) float bc[4][8][32];          // per warp: 8 tokens x (16 B + 16 C)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < 8 * 32; i += 32) bc[warp][i / 32][i % 32] = in[i] * 0.01f;
    __syncwarp();
    uint64_t A2[2][8], h[2][8];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            A2[ch][q] = pack2(-(1.f + q + in[lane]) * 1.44f, -(1.5f + q + in[lane + 32]) * 1.44f);
            h[ch][q] = 0ull;
        }
    uint64_t Bq[8], Cq[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { Bq[q] = pack2(in[q], in[q + 1]); Cq[q] = pack2(in[q + 2], in[q + 3]); }
    float acc = 0.f;
    float dt0 = 0.01f + in[lane] * 1e-3f, dt1 = 0.02f + in[lane] * 1e-3f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int t = 0; t < kTok; t += 8) {
#pragma unroll (kLds ? 1 : 8)
        for (int jj = 0; jj < 8; ++jj) {
            if (kLds) {
                const ulonglong2* p = reinterpret_cast<const ulonglong2*>(&bc[warp][jj][0]);
#pragma unroll
                for (int q = 0; q < 4; ++q) { ulonglong2 b = p[q], c = p[4 + q]; Bq[2 * q] = b.x; Bq[2 * q + 1] = b.y; Cq[2 * q] = c.x; Cq[2 * q + 1] = c.y; }
            }
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const float dt = ch ? dt1 : dt0, dtu = dt * (0.5f + acc * 1e-9f);
                const uint64_t dt2 = pack2(dt, dt), dtu2 = pack2(dtu, dtu);
                uint64_t y0 = 0ull, y1 = 0ull;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float a0, a1;
                    unpack2(mul2(dt2, A2[ch][q]), a0, a1);
                    const uint64_t dA = pack2(ex2(a0), ex2(a1));
                    h[ch][q] = fma2(dA, h[ch][q], mul2(dtu2, Bq[q]));
                    if (q & 1) y1 = fma2(h[ch][q], Cq[q], y1); else y0 = fma2(h[ch][q], Cq[q], y0);
                }
                float ya, yb, yc, yd;
                unpack2(y0, ya, yb); unpack2(y1, yc, yd);
                acc += (ya + yb) + (yc + yd);
            }
            dt0 += 1e-6f; dt1 += 1e-6f;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (lane == 0) cyc[blockIdx.x * 4 + warp] = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, int n_sm, long long* d_cyc, long long* h_cyc) {
    for (int w = 1; w <= 4; ++w) {
        const int grid = n_sm * w;
        launch(grid);                       // warm-up
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        launch(grid);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * grid * 4, cudaMemcpyDeviceToHost);
        long long mn = 1LL << 60, mx = 0; double sum = 0;
        for (int i = 0; i < grid * 4; ++i) { mn = h_cyc[i] < mn ? h_cyc[i] : mn; mx = h_cyc[i] > mx ? h_cyc[i] : mx; sum += h_cyc[i]; }
        const double avg = sum / (grid * 4), per_tok = avg / kTok;
        printf("{\"kernel\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_token_per_warp\": %.1f, \"min\": %.1f, \"max\": %.1f, "
               "\"mufu_per_clk_per_smsp\": %.4f, \"frac_of_1_per_8clk\": %.3f, \"ms\": %.3f, \"err\": \"%s\"}\n",
               name, w, per_tok, (double)mn / kTok, (double)mx / kTok, w * 32.0 / per_tok, w * 32.0 * 8.0 / per_tok, ms,
               cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    float *out, *in;
    long long *d_cyc, *h_cyc;
    cudaMalloc(&out, sizeof(float) * n_sm * 4 * 128);
    cudaMalloc(&in, sizeof(float) * 1024);
    cudaMalloc(&d_cyc, sizeof(long long) * n_sm * 16);
    h_cyc = (long long*)malloc(sizeof(long long) * n_sm * 16);
    float hin[1024];
    for (int i = 0; i < 1024; ++i) hin[i] = 0.001f * (i % 97);
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    printf("{\"n_sm\": %d}\n", n_sm);
    run("mufu_only(+1 FADD each)", [&](int g) { k_mufu<<<g, 128>>>(out, d_cyc, 0.37f); }, n_sm, d_cyc, h_cyc);
    run("scan_mix_regs", [&](int g) { k_mix<false><<<g, 128>>>(out, d_cyc, in); }, n_sm, d_cyc, h_cyc);
    run("scan_mix_lds", [&](int g) { k_mix<true><<<g, 128>>>(out, d_cyc, in); }, n_sm, d_cyc, h_cyc);
    return 0;
}
