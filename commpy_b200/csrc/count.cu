// Bit / frame error counting -- replaces the XOR-and-sum of commpy/links.py:335-337 (and :253-256):
// one warp per frame, popcount over 32-bit words of the XOR, block-level accumulation, one 64-bit
// atomic pair per block.  counters[0] += differing bits, counters[1] += frames with at least one.
#include "common.cuh"

using namespace cpb;

namespace count {

__global__ void __launch_bounds__(256) count_errors_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b,
                                                           int64_t batch, int64_t L, int64_t lda, int64_t ldb,
                                                           unsigned long long *counters)
{
    __shared__ unsigned long long s_bits, s_frames;
    if (threadIdx.x == 0) { s_bits = 0ull; s_frames = 0ull; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long my_bits = 0ull, my_frames = 0ull;
    for (int64_t f = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); f < batch; f += warps) {
        const uint8_t *pa = a + f * lda, *pb = b + f * ldb;
        unsigned cnt = 0;
        const bool vec = ((((uintptr_t)pa) | ((uintptr_t)pb)) & 3) == 0;
        int64_t done = 0;
        if (vec) {
            const int64_t nw = L >> 2;
            const uint32_t *wa = reinterpret_cast<const uint32_t *>(pa), *wb = reinterpret_cast<const uint32_t *>(pb);
            for (int64_t i = lane; i < nw; i += 32) cnt += __popc((__ldg(wa + i) ^ __ldg(wb + i)) & 0x01010101u);
            done = nw << 2;
        }
        for (int64_t i = done + lane; i < L; i += 32) cnt += (pa[i] ^ pb[i]) & 1u;
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) { my_bits += cnt; my_frames += cnt ? 1ull : 0ull; }
    }
    if (lane == 0 && (my_bits | my_frames)) {
        atomicAdd(&s_bits, my_bits);
        atomicAdd(&s_frames, my_frames);
    }
    __syncthreads();
    if (threadIdx.x == 0 && (s_bits | s_frames)) {
        atomicAdd(&counters[0], s_bits);
        atomicAdd(&counters[1], s_frames);
    }
}

}  // namespace count

extern "C" int cpb_count_errors(const uint8_t *a_dev, const uint8_t *b_dev, int64_t batch, int64_t L, int64_t lda,
                                int64_t ldb, int64_t *counters_dev, void *stream)
{
    if (!a_dev || !b_dev || !counters_dev || batch < 0 || L < 0 || lda < L || ldb < L) return CPB_EINVAL;
    if (batch == 0 || L == 0) return CPB_OK;
    const DeviceProps &dp = device_props();
    int64_t blocks = ceil_div(batch, 8);
    if (blocks > (int64_t)dp.sm_count * 8) blocks = (int64_t)dp.sm_count * 8;
    count::count_errors_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        a_dev, b_dev, batch, L, lda, ldb, reinterpret_cast<unsigned long long *>(counters_dev));
    CPB_LAUNCH_CHECK();
    return CPB_OK;
}
