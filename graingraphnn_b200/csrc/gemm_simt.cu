// gemm_simt.cu — kernel family (c), fp32 CUDA-core implementation.
//
//   gg_node_proj   : P = [X|h] Wcat^T + b      (per-node K/V/Q projections of every role x gate; re-association
//                    of the per-edge lin_key/lin_value/lin_query of periodGATconv.py:216-218)
//   gg_gate_update : pre_g = sum_t W2 agg_t + Wskip [X|h] + ... fused with the LSTM gate math
//                    (periodGATconv.py:186-192,218 and heteropgclstm.py:115-146)
//
// These are the exact-fp32 kernels: they define the numerics the tcgen05 3xTF32 path (gemm_tc.cu) is checked
// against, and they serve widths/devices the tensor-core path does not cover.  Classic shared-memory tiling with
// k-major XOR-swizzled tiles so both the transposing stores and the 128-bit fragment loads are (nearly)
// conflict-free.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------
// swizzled k-major tile: element (k, r) of a tile with R rows lives at  k*R + ((r>>2 ^ (k>>2 & 3))<<2 | r&3)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int swz(int k, int r, int R) { return k * R + ((((r >> 2) ^ ((k >> 2) & 3)) << 2) | (r & 3)); }

// load a [ROWS x 16] slab of a row-major matrix (rows = m or n, contiguous k) into a swizzled k-major tile.
// Global reads are float4 along k (requires 16B-aligned rows, K % 4 == 0).
template <int ROWS, int THREADS>
__device__ __forceinline__ void load_tile_k16(float* __restrict__ tile, const float* __restrict__ g, int ld,
                                              int row0, int nrows, int k0, int K) {
    for (int t = threadIdx.x; t < ROWS * 4; t += THREADS) {
        const int r = t >> 2, kq = t & 3;
        const int k = k0 + 4 * kq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < nrows && k < K) v = ldg4(g + (size_t)(row0 + r) * ld + k);
        tile[swz(4 * kq + 0, r, ROWS)] = v.x;
        tile[swz(4 * kq + 1, r, ROWS)] = v.y;
        tile[swz(4 * kq + 2, r, ROWS)] = v.z;
        tile[swz(4 * kq + 3, r, ROWS)] = v.w;
    }
}

// ---------------------------------------------------------------------------------------------------------
// node projection: 128x128 tile, 256 threads, 8x8 outputs per thread (rows {ty*4.., 64+ty*4..}, cols likewise)
// ---------------------------------------------------------------------------------------------------------
constexpr int PB = 128;

__global__ void __launch_bounds__(256)
node_proj_kernel(const float* __restrict__ A1, int lda1, int K1, const float* __restrict__ A2, int lda2, int K2,
                 const float* __restrict__ W, int ldw, const float* __restrict__ bias,
                 float* __restrict__ out, int ldo, int M, int N) {
    __shared__ __align__(16) float As[16 * PB];
    __shared__ __align__(16) float Bs[16 * PB];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * PB, n0 = blockIdx.x * PB;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int piece = 0; piece < 2; ++piece) {
        const float* A = piece ? A2 : A1;
        const int lda = piece ? lda2 : lda1, K = piece ? K2 : K1, wk0 = piece ? K1 : 0;
        if (A == nullptr || K == 0) continue;
        for (int k0 = 0; k0 < K; k0 += 16) {
            load_tile_k16<PB, 256>(As, A, lda, m0, M, k0, K);
            load_tile_k16<PB, 256>(Bs, W + wk0, ldw, n0, N, k0, K);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int key = (k >> 2) & 3;
                const float4 a0 = *reinterpret_cast<const float4*>(&As[k * PB + ((ty ^ key) << 2)]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[k * PB + (((16 + ty) ^ key) << 2)]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k * PB + ((tx ^ key) << 2)]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k * PB + (((16 + tx) ^ key) << 2)]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 64 + tx * 4;
            if (n >= N) continue;   // N % 4 == 0 is enforced by the host wrapper
            float4 v = make_float4(acc[i][jh * 4], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
            if (bias) { const float4 bb = ldg4(bias + n); v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w; }
            *reinterpret_cast<float4*>(out + (size_t)m * ldo + n) = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// gate update: 64 nodes per block, all G gates kept in registers, LSTM math in the epilogue
// ---------------------------------------------------------------------------------------------------------
constexpr int GM = 64;
constexpr int kMaxAggInputs = 4;

struct AggIn { const float* agg; int ld_agg; const float* ea; const int* rowptr; const float* W2; const float* We; const float* b2; int weighted; };
struct GateParams {
    AggIn in[kMaxAggInputs]; int n_in;
    const float* X; int ldx, K1; const float* H; int ldh;
    const float* Wskip; int ldw; const float* btot;
    const float* c_in; float* out_h; float* out_c;
    int M, G, mode;
};

template <int NV>
__device__ __forceinline__ void gate_segment(float (&acc)[2][4 * NV], float* As, float* Bs,
                                             const float* __restrict__ A, int lda, int K,
                                             const float* __restrict__ W, int ldw, int m0, int M) {
    constexpr int C = 32 * NV;
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    for (int k0 = 0; k0 < K; k0 += 16) {
        load_tile_k16<GM, 256>(As, A, lda, m0, M, k0, K);
        load_tile_k16<C, 256>(Bs, W, ldw, 0, C, k0, K);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int key = (k >> 2) & 3;
            const float a0 = As[k * GM + ((((ty >> 2) ^ key) << 2) | (ty & 3))];
            const float a1 = As[k * GM + (((((ty + 32) >> 2) ^ key) << 2) | (ty & 3))];
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(&Bs[k * C + (((tx + 8 * j) ^ key) << 2)]);
                acc[0][4 * j + 0] = fmaf(a0, b.x, acc[0][4 * j + 0]); acc[0][4 * j + 1] = fmaf(a0, b.y, acc[0][4 * j + 1]);
                acc[0][4 * j + 2] = fmaf(a0, b.z, acc[0][4 * j + 2]); acc[0][4 * j + 3] = fmaf(a0, b.w, acc[0][4 * j + 3]);
                acc[1][4 * j + 0] = fmaf(a1, b.x, acc[1][4 * j + 0]); acc[1][4 * j + 1] = fmaf(a1, b.y, acc[1][4 * j + 1]);
                acc[1][4 * j + 2] = fmaf(a1, b.z, acc[1][4 * j + 2]); acc[1][4 * j + 3] = fmaf(a1, b.w, acc[1][4 * j + 3]);
            }
        }
        __syncthreads();
    }
}

template <int NV, int GMAX>
__global__ void __launch_bounds__(256)
gate_update_kernel(const GateParams p) {
    constexpr int C = 32 * NV;
    __shared__ __align__(16) float As[16 * GM];
    __shared__ __align__(16) float Bs[16 * C];
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    const int m0 = blockIdx.x * GM;
    float pre[GMAX][2][4 * NV];
#pragma unroll
    for (int g = 0; g < GMAX; ++g) {
        if (g >= p.G) break;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4 * NV; ++j) pre[g][i][j] = 0.f;
        for (int t = 0; t < p.n_in; ++t)
            gate_segment<NV>(pre[g], As, Bs, p.in[t].agg + g * C, p.in[t].ld_agg, C,
                             p.in[t].W2 + (size_t)g * C * C, C, m0, p.M);
        gate_segment<NV>(pre[g], As, Bs, p.X, p.ldx, p.K1, p.Wskip + (size_t)g * C * p.ldw, p.ldw, m0, p.M);
        if (p.H) gate_segment<NV>(pre[g], As, Bs, p.H, p.ldh, C, p.Wskip + (size_t)g * C * p.ldw + p.K1, p.ldw, m0, p.M);
    }
    // ---- epilogue: rank-1 terms, biases, activation ------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty + 32 * i;
        if (m >= p.M) continue;
        float cnt[kMaxAggInputs];
        for (int t = 0; t < p.n_in; ++t) {
            const int deg = __ldg(&p.in[t].rowptr[m + 1]) - __ldg(&p.in[t].rowptr[m]);
            cnt[t] = p.in[t].weighted ? (deg > 0 ? 1.f : 0.f) : (float)deg;
        }
#pragma unroll
        for (int g = 0; g < GMAX; ++g) {
            if (g >= p.G) break;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int n = 4 * (tx + 8 * j);
                float4 v = make_float4(pre[g][i][4 * j], pre[g][i][4 * j + 1], pre[g][i][4 * j + 2], pre[g][i][4 * j + 3]);
                for (int t = 0; t < p.n_in; ++t) {
                    const float ea = __ldg(&p.in[t].ea[(size_t)m * p.G + g]);
                    const float4 we = ldg4(p.in[t].We + g * C + n), b2 = ldg4(p.in[t].b2 + g * C + n);
                    v.x += fmaf(ea, we.x, cnt[t] * b2.x); v.y += fmaf(ea, we.y, cnt[t] * b2.y);
                    v.z += fmaf(ea, we.z, cnt[t] * b2.z); v.w += fmaf(ea, we.w, cnt[t] * b2.w);
                }
                const float4 bt = ldg4(p.btot + g * C + n);
                pre[g][i][4 * j] = v.x + bt.x; pre[g][i][4 * j + 1] = v.y + bt.y;
                pre[g][i][4 * j + 2] = v.z + bt.z; pre[g][i][4 * j + 3] = v.w + bt.w;
            }
        }
        if (p.mode == GG_GATE_RAW || p.mode == GG_GATE_RELU) {
#pragma unroll
            for (int g = 0; g < GMAX; ++g) {
                if (g >= p.G) break;
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float4 v = make_float4(pre[g][i][4 * j], pre[g][i][4 * j + 1], pre[g][i][4 * j + 2], pre[g][i][4 * j + 3]);
                    if (p.mode == GG_GATE_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    *reinterpret_cast<float4*>(p.out_h + (size_t)m * p.G * C + g * C + 4 * (tx + 8 * j)) = v;
                }
            }
        } else {
            // gate order: LSTM (i,f,c,o) / LSTM0 (i,c,o) — heteropgclstm.py:176-183
            constexpr int gi = 0;
            const int gf = 1, gc = (p.mode == GG_GATE_LSTM) ? 2 : 1, go = (p.mode == GG_GATE_LSTM) ? 3 : 2;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int n = 4 * (tx + 8 * j);
                float cold[4] = {0.f, 0.f, 0.f, 0.f};
                if (p.mode == GG_GATE_LSTM && p.c_in) {
                    const float4 c4 = ldg4(p.c_in + (size_t)m * C + n);
                    cold[0] = c4.x; cold[1] = c4.y; cold[2] = c4.z; cold[3] = c4.w;
                }
                float hn[4], cn[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float pi = 0.f, pf = 0.f, pc = 0.f, po = 0.f;
#pragma unroll
                    for (int g = 0; g < GMAX; ++g) {   // static indexing keeps pre[] in registers
                        const float v = pre[g][i][4 * j + q];
                        if (g == gi) pi = v;
                        if (g == gf) pf = v;
                        if (g == gc) pc = v;
                        if (g == go) po = v;
                    }
                    const float ig = sigmoidf_(pi), tg = tanhf(pc), og = sigmoidf_(po);
                    float c2 = ig * tg;
                    if (p.mode == GG_GATE_LSTM) c2 = sigmoidf_(pf) * cold[q] + c2;
                    cn[q] = c2;
                    hn[q] = og * tanhf(c2);
                }
                *reinterpret_cast<float4*>(p.out_h + (size_t)m * C + n) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                if (p.out_c) *reinterpret_cast<float4*>(p.out_c + (size_t)m * C + n) = make_float4(cn[0], cn[1], cn[2], cn[3]);
            }
        }
    }
}

}  // namespace

extern "C" int gg_node_proj(const float* A1, int32_t lda1, int32_t K1, const float* A2, int32_t lda2, int32_t K2,
                            const float* W, int32_t ldw, const float* bias, float* out, int32_t ldo,
                            int32_t M, int32_t N, void* stream) {
    if (M < 0 || N < 0 || K1 < 0 || K2 < 0) return GG_EINVAL;
    if (M == 0 || N == 0) return 0;
    if (!A1 || !W || !out) return GG_EINVAL;
    if (!A2) K2 = 0;
    if ((K1 & 3) || (K2 & 3) || (lda1 & 3) || (K2 && (lda2 & 3)) || (ldw & 3) || (ldo & 3) || (N & 3)) return GG_EALIGN;
    if (!gg_aligned16(A1) || (A2 && !gg_aligned16(A2)) || !gg_aligned16(W) || !gg_aligned16(out) || (bias && !gg_aligned16(bias))) return GG_EALIGN;
    dim3 grid((N + PB - 1) / PB, (M + PB - 1) / PB);
    node_proj_kernel<<<grid, 256, 0, GG_STREAM(stream)>>>(A1, lda1, K1, A2, lda2, K2, W, ldw, bias, out, ldo, M, N);
    GG_LAUNCH_OK();
    return 0;
}

extern "C" int gg_gate_update(const gg_agg_input* inputs, int32_t n_inputs,
                              const float* X, int32_t ldx, int32_t K1, const float* H, int32_t ldh,
                              const float* Wskip, int32_t ldw, const float* btot,
                              const float* c_in, float* out_h, float* out_c,
                              int32_t M, int32_t G, int32_t C, int32_t mode, void* stream) {
    if (M < 0 || G < 1 || G > 4 || C % 32 || C < 32 || C > 128 || n_inputs < 0 || n_inputs > kMaxAggInputs) return GG_EINVAL;
    if (mode < GG_GATE_RAW || mode > GG_GATE_LSTM0) return GG_EINVAL;
    if ((mode == GG_GATE_LSTM && G != 4) || (mode == GG_GATE_LSTM0 && G != 3) || (mode == GG_GATE_RELU && G != 1)) return GG_EINVAL;
    if (M == 0) return 0;
    if (!X || !Wskip || !btot || !out_h || (n_inputs && !inputs)) return GG_EINVAL;
    if ((K1 & 3) || (ldx & 3) || (ldw & 3) || (H && (ldh & 3))) return GG_EALIGN;
    if (!gg_aligned16(X) || (H && !gg_aligned16(H)) || !gg_aligned16(Wskip) || !gg_aligned16(btot) || !gg_aligned16(out_h) ||
        (c_in && !gg_aligned16(c_in)) || (out_c && !gg_aligned16(out_c))) return GG_EALIGN;
    GateParams p;
    p.n_in = n_inputs;
    for (int t = 0; t < n_inputs; ++t) {
        const gg_agg_input& s = inputs[t];
        if (!s.agg || !s.ea || !s.rowptr || !s.W2 || !s.We || !s.b2) return GG_EINVAL;
        if ((s.ld_agg & 3) || !gg_aligned16(s.agg) || !gg_aligned16(s.W2) || !gg_aligned16(s.We) || !gg_aligned16(s.b2)) return GG_EALIGN;
        p.in[t] = AggIn{s.agg, s.ld_agg, s.ea, s.rowptr, s.W2, s.We, s.b2, s.weighted};
    }
    p.X = X; p.ldx = ldx; p.K1 = K1; p.H = H; p.ldh = ldh; p.Wskip = Wskip; p.ldw = ldw; p.btot = btot;
    p.c_in = c_in; p.out_h = out_h; p.out_c = out_c; p.M = M; p.G = G; p.mode = mode;
    const unsigned nb = (unsigned)((M + GM - 1) / GM);
    cudaStream_t st = GG_STREAM(stream);
#define GG_GU(NV) gate_update_kernel<NV, 4><<<nb, 256, 0, st>>>(p)
    switch (C / 32) { case 1: GG_GU(1); break; case 2: GG_GU(2); break; case 3: GG_GU(3); break; default: GG_GU(4); }
#undef GG_GU
    GG_LAUNCH_OK();
    return 0;
}
