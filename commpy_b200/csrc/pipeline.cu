// Host-buffer entry point: what a CommPy caller holds is host memory, so the end-to-end call is
//   H2D(chunk i+1)  ||  decode(chunk i)  ||  D2H(chunk i-1)
// on three internal streams with per-stream device buffers (stream-ordered pool allocations).
// Pinned host buffers give true overlap; pageable ones work but are staged by the driver.
#include <algorithm>

#include "common.cuh"

using namespace cpb;

extern "C" int cpb_viterbi_decode_host(const cpbTrellis *t, const void *coded_host, int in_dtype, int64_t batch,
                                       int64_t n_in, int tb_depth, int mode, uint8_t *out_bits_host)
{
    if (!t || !coded_host || !out_bits_host || batch < 0 || n_in <= 0) return CPB_EINVAL;
    if (in_dtype != CPB_U8 && in_dtype != CPB_F32) return CPB_EINVAL;
    if (batch == 0) return CPB_OK;
    int64_t L = 0, T = 0;
    int rc = cpb_viterbi_sizes(t, n_in, &L, &T);
    if (rc) return rc;
    const size_t esz = (in_dtype == CPB_U8) ? 1 : 4;
    constexpr int NS = 3;
    // chunks of ~32 MB of input keep all three engines busy without long pipeline fill/drain
    int64_t chunk = std::max<int64_t>(1, (int64_t)(32.0e6 / ((double)n_in * esz)));
    chunk = std::min<int64_t>(ceil_div(chunk, 2048) * 2048, batch);
    cudaStream_t st[NS] = {nullptr, nullptr, nullptr};
    void *din[NS] = {nullptr, nullptr, nullptr};
    uint8_t *dout[NS] = {nullptr, nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < NS && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMallocAsync(&din[i], (size_t)chunk * n_in * esz, st[i]);
        if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void **>(&dout[i]), (size_t)chunk * L, st[i]);
    }
    if (e != cudaSuccess) rc = record_cuda_error(e, "pipeline setup", __FILE__, __LINE__);
    int slot = 0;
    for (int64_t f0 = 0; f0 < batch && rc == CPB_OK; f0 += chunk, slot = (slot + 1) % NS) {
        const int64_t nb = std::min<int64_t>(chunk, batch - f0);
        const char *src = reinterpret_cast<const char *>(coded_host) + (size_t)f0 * n_in * esz;
        e = cudaMemcpyAsync(din[slot], src, (size_t)nb * n_in * esz, cudaMemcpyHostToDevice, st[slot]);
        if (e != cudaSuccess) { rc = record_cuda_error(e, "H2D", __FILE__, __LINE__); break; }
        rc = cpb_viterbi_decode(t, din[slot], in_dtype, nb, n_in, tb_depth, mode, dout[slot], nullptr, 0, st[slot]);
        if (rc) break;
        e = cudaMemcpyAsync(out_bits_host + (size_t)f0 * L, dout[slot], (size_t)nb * L, cudaMemcpyDeviceToHost, st[slot]);
        if (e != cudaSuccess) { rc = record_cuda_error(e, "D2H", __FILE__, __LINE__); break; }
    }
    for (int i = 0; i < NS; ++i) {
        if (!st[i]) continue;
        if (din[i]) cudaFreeAsync(din[i], st[i]);
        if (dout[i]) cudaFreeAsync(dout[i], st[i]);
        e = cudaStreamSynchronize(st[i]);
        if (e != cudaSuccess && rc == CPB_OK) rc = record_cuda_error(e, "pipeline sync", __FILE__, __LINE__);
        cudaStreamDestroy(st[i]);
    }
    return rc;
}
