// Host-buffer entry point: what a CommPy caller holds is host memory, so the end-to-end call is
//   H2D(chunk i+1)  ||  decode(chunk i)  ||  D2H(chunk i-1)
// on three internal streams with per-stream device buffers.  Streams and buffers are created once per
// device and reused (a decode call must not pay cudaMalloc); pinned host buffers give true copy/compute
// overlap, pageable ones work but are staged by the driver.
#include <algorithm>
#include <mutex>

#include "common.cuh"

using namespace cpb;

namespace {

constexpr int NS = 3;

struct PipeCtx {
    bool ready = false;
    cudaStream_t st[NS] = {nullptr, nullptr, nullptr};
    void *din[NS] = {nullptr, nullptr, nullptr};
    uint8_t *dout[NS] = {nullptr, nullptr, nullptr};
    size_t cap_in = 0, cap_out = 0;
};

std::mutex g_mu;
PipeCtx g_ctx[64];

int ensure(PipeCtx &c, size_t need_in, size_t need_out)
{
    if (!c.ready) {
        for (int i = 0; i < NS; ++i) CPB_CUDA(cudaStreamCreateWithFlags(&c.st[i], cudaStreamNonBlocking));
        c.ready = true;
    }
    if (need_in > c.cap_in) {
        for (int i = 0; i < NS; ++i) {
            if (c.din[i]) CPB_CUDA(cudaFree(c.din[i]));
            c.din[i] = nullptr;
            CPB_CUDA(cudaMalloc(&c.din[i], need_in));
        }
        c.cap_in = need_in;
    }
    if (need_out > c.cap_out) {
        for (int i = 0; i < NS; ++i) {
            if (c.dout[i]) CPB_CUDA(cudaFree(c.dout[i]));
            c.dout[i] = nullptr;
            CPB_CUDA(cudaMalloc(reinterpret_cast<void **>(&c.dout[i]), need_out));
        }
        c.cap_out = need_out;
    }
    return CPB_OK;
}

}  // namespace

extern "C" int cpb_viterbi_decode_host(const cpbTrellis *t, const void *coded_host, int in_dtype, int64_t batch,
                                       int64_t n_in, int tb_depth, int mode, uint8_t *out_bits_host)
{
    if (in_dtype != CPB_U8 && in_dtype != CPB_F32) return CPB_EINVAL;
    if (t && batch == 0) return CPB_OK;
    if (!t || !coded_host || !out_bits_host || batch < 0 || n_in <= 0) return CPB_EINVAL;
    int64_t L = 0, T = 0;
    int rc = cpb_viterbi_sizes(t, n_in, &L, &T);
    if (rc) return rc;
    const size_t esz = (in_dtype == CPB_U8) ? 1 : 4;
    // chunks of ~1/4 of the batch (at least 8192 frames, a multiple of 2048): enough pieces for the three
    // engines to overlap, large enough to fill the GPU
    int64_t chunk = ceil_div(ceil_div(batch, 4), 2048) * 2048;
    chunk = std::max<int64_t>(chunk, 8192);
    chunk = std::min<int64_t>(chunk, batch);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return CPB_EINVAL;
    std::lock_guard<std::mutex> lock(g_mu);
    PipeCtx &c = g_ctx[dev];
    rc = ensure(c, (size_t)chunk * n_in * esz, (size_t)chunk * L);
    if (rc) return rc;
    int slot = 0;
    cudaError_t e = cudaSuccess;
    for (int64_t f0 = 0; f0 < batch && rc == CPB_OK; f0 += chunk, slot = (slot + 1) % NS) {
        const int64_t nb = std::min<int64_t>(chunk, batch - f0);
        const char *src = reinterpret_cast<const char *>(coded_host) + (size_t)f0 * n_in * esz;
        e = cudaMemcpyAsync(c.din[slot], src, (size_t)nb * n_in * esz, cudaMemcpyHostToDevice, c.st[slot]);
        if (e != cudaSuccess) { rc = record_cuda_error(e, "H2D", __FILE__, __LINE__); break; }
        rc = cpb_viterbi_decode(t, c.din[slot], in_dtype, nb, n_in, tb_depth, mode, c.dout[slot], nullptr, 0, c.st[slot]);
        if (rc) break;
        e = cudaMemcpyAsync(out_bits_host + (size_t)f0 * L, c.dout[slot], (size_t)nb * L, cudaMemcpyDeviceToHost, c.st[slot]);
        if (e != cudaSuccess) { rc = record_cuda_error(e, "D2H", __FILE__, __LINE__); break; }
    }
    for (int i = 0; i < NS; ++i) {
        e = cudaStreamSynchronize(c.st[i]);
        if (e != cudaSuccess && rc == CPB_OK) rc = record_cuda_error(e, "pipeline sync", __FILE__, __LINE__);
    }
    return rc;
}
