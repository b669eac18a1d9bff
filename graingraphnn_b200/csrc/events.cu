// events.cu — event candidates of a rollout step (first stage): which joint-joint edges switch and
// which grains vanish.  The reference scans the full prediction arrays on the host,
//     L1 = ((sigmoid(edge_event) > threshold) & (src < dst)).nonzero()                      models.py:627-629
//     grain_event = ((mask_grain > 0) & (grain_area < threshold)).nonzero()                 test.py:414
// which at 10^6 grains means a device->host copy of 6M + 1M floats per step for a handful of events.  Here one streaming
// pass per array leaves only the candidates (id, value) in a small buffer; the host sorts the few survivors exactly as
// the reference does (argsort by area test.py:416, sort by probability models.py:730-731).
// HBM-bound: 4 B per element read once (src / dst only for elements that pass the value test).
#include "common.cuh"

namespace {

// mode 0: keep v >= thr (edge logits; thr = smallest fp32 logit whose sigmoid exceeds the probability threshold, found on
//         the host with the reference's own sigmoid, so the decision equals `sigmoid(v) > p` bit for bit)
// mode 1: keep v <  thr (grain areas)
template <int MODE>
__global__ void __launch_bounds__(256)
select_kernel(const float* __restrict__ v, int64_t n, int32_t ld, float thr,
              const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
              const float* __restrict__ mask, int32_t ld_mask,
              int32_t cap, int32_t* __restrict__ count, int32_t* __restrict__ ids, float* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n + 31) / 32 * 32;                       // whole warps stay in the loop for the ballot
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool keep = false;
        float x = 0.f;
        if (i < n) {
            x = v[i * ld];
            keep = MODE == 0 ? (x >= thr) : (x < thr);
            if (keep && src) keep = src[i] < dst[i];
            if (keep && mask) keep = mask[i * ld_mask] > 0.f;
        }
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(count, __popc(m));        // one atomic per warp that found something
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) {
                int at = base + __popc(m & ((1u << lane) - 1u));
                if (at < cap) { ids[at] = (int32_t)i; vals[at] = x; }
            }
        }
    }
}

}  // namespace

extern "C" int gg_select_events(const float* values, int64_t n, int32_t ld, float threshold, int32_t mode,
                                const int64_t* src, const int64_t* dst, const float* mask, int32_t ld_mask,
                                int32_t cap, int32_t* count, int32_t* ids, float* vals, void* stream) {
    if (n < 0 || n > 0x7fffffffLL || ld < 1 || cap < 0 || !count || (mode != 0 && mode != 1)) return GG_EINVAL;
    if ((src == nullptr) != (dst == nullptr) || (mask && ld_mask < 1) || (cap > 0 && (!ids || !vals))) return GG_EINVAL;
    if (n > 0 && !values) return GG_EINVAL;
    cudaStream_t s = GG_STREAM(stream);
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), s);
    if (e != cudaSuccess) return (int)e;
    if (n == 0) return 0;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (mode == 0) select_kernel<0><<<blocks, 256, 0, s>>>(values, n, ld, threshold, src, dst, mask, ld_mask, cap, count, ids, vals);
    else           select_kernel<1><<<blocks, 256, 0, s>>>(values, n, ld, threshold, src, dst, mask, ld_mask, cap, count, ids, vals);
    GG_LAUNCH_OK();
    return 0;
}
