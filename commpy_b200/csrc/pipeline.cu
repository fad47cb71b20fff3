// Host-buffer entry points: what a CommPy caller holds is host memory, so the end-to-end call is
//   H2D(chunk i+1)  ||  kernels(chunk i)  ||  D2H(chunk i-1)
// on three internal streams with per-stream device buffers.  Streams and buffers belong to the HANDLE the call is made
// with (cpb::PipeCtx in handles.cuh): created once, reused (a decode call must not pay cudaMalloc), freed with the
// handle.  Pinned host buffers give true copy/compute overlap, pageable ones work but are staged by the driver.
#include <algorithm>

#include "handles.cuh"

namespace cpb {

PipeCtx::~PipeCtx()
{
    for (int i = 0; i < NS; ++i) {
        if (din[i]) cudaFree(din[i]);
        if (dout[i]) cudaFree(dout[i]);
        if (st[i]) cudaStreamDestroy(st[i]);
    }
}

static int ensure(PipeCtx &c, size_t need_in, size_t need_out)
{
    int dev = 0;
    CPB_CUDA(cudaGetDevice(&dev));
    if (!c.ready) {
        for (int i = 0; i < PipeCtx::NS; ++i) CPB_CUDA(cudaStreamCreateWithFlags(&c.st[i], cudaStreamNonBlocking));
        c.ready = true;
        c.device = dev;
    }
    if (c.device != dev) return CPB_EINVAL;          // a handle lives on the device it was created on
    if (need_in > c.cap_in) {
        for (int i = 0; i < PipeCtx::NS; ++i) {
            if (c.din[i]) CPB_CUDA(cudaFree(c.din[i]));
            c.din[i] = nullptr;
            CPB_CUDA(cudaMalloc(&c.din[i], need_in));
        }
        c.cap_in = need_in;
    }
    if (need_out > c.cap_out) {
        for (int i = 0; i < PipeCtx::NS; ++i) {
            if (c.dout[i]) CPB_CUDA(cudaFree(c.dout[i]));
            c.dout[i] = nullptr;
            CPB_CUDA(cudaMalloc(&c.dout[i], need_out));
        }
        c.cap_out = need_out;
    }
    return CPB_OK;
}

int64_t pipe_chunk(int64_t batch, int64_t min_chunk, int64_t multiple)
{
    // ~1/8 of the batch: enough pieces for the three engines to overlap, large enough to fill the GPU
    int64_t chunk = ceil_div(ceil_div(batch, 8), multiple) * multiple;
    chunk = std::max<int64_t>(chunk, min_chunk);
    return std::min<int64_t>(chunk, batch);
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

int pipe_run(PipeCtx &c, const std::vector<HostSeg> &ins, const std::vector<HostSeg> &outs, int64_t batch, int64_t chunk,
             const PipeLaunch &launch)
{
    if (batch <= 0) return CPB_OK;
    std::lock_guard<std::mutex> lock(c.mu);
    size_t need_in = 0, need_out = 0;
    std::vector<size_t> off_in, off_out;
    for (const HostSeg &s : ins) { off_in.push_back(need_in); need_in += align256(s.stride * (size_t)chunk); }
    for (const HostSeg &s : outs) { off_out.push_back(need_out); need_out += align256(s.stride * (size_t)chunk); }
    int rc = ensure(c, std::max<size_t>(need_in, 256), std::max<size_t>(need_out, 256));
    if (rc) return rc;
    int slot = 0;
    cudaError_t e = cudaSuccess;
    std::vector<void *> din(ins.size()), dout(outs.size());
    for (int64_t f0 = 0; f0 < batch && rc == CPB_OK; f0 += chunk, slot = (slot + 1) % PipeCtx::NS) {
        const int64_t nb = std::min<int64_t>(chunk, batch - f0);
        cudaStream_t st = c.st[slot];
        for (size_t k = 0; k < ins.size() && rc == CPB_OK; ++k) {
            din[k] = static_cast<char *>(c.din[slot]) + off_in[k];
            e = cudaMemcpyAsync(din[k], static_cast<const char *>(ins[k].in) + (size_t)f0 * ins[k].stride,
                                (size_t)nb * ins[k].stride, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) rc = record_cuda_error(e, "H2D", __FILE__, __LINE__);
        }
        for (size_t k = 0; k < outs.size(); ++k) dout[k] = static_cast<char *>(c.dout[slot]) + off_out[k];
        if (rc == CPB_OK) rc = launch(din, dout, f0, nb, st);
        for (size_t k = 0; k < outs.size() && rc == CPB_OK; ++k) {
            e = cudaMemcpyAsync(static_cast<char *>(outs[k].out) + (size_t)f0 * outs[k].stride, dout[k],
                                (size_t)nb * outs[k].stride, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) rc = record_cuda_error(e, "D2H", __FILE__, __LINE__);
        }
    }
    for (int i = 0; i < PipeCtx::NS; ++i) {
        e = cudaStreamSynchronize(c.st[i]);
        if (e != cudaSuccess && rc == CPB_OK) rc = record_cuda_error(e, "pipeline sync", __FILE__, __LINE__);
    }
    return rc;
}

}  // namespace cpb

using namespace cpb;

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
    return pipe_run(const_cast<cpbTrellis *>(t)->pipe, {{coded_host, nullptr, (size_t)n_in * esz}}, {{nullptr, out_bits_host, (size_t)L}},
                    batch, pipe_chunk(batch, 16384, 2048),
                    [&](const std::vector<void *> &din, const std::vector<void *> &dout, int64_t, int64_t nb, cudaStream_t st) {
                        return cpb_viterbi_decode(t, din[0], in_dtype, nb, n_in, tb_depth, mode, static_cast<uint8_t *>(dout[0]),
                                                  nullptr, 0, st);
                    });
}

extern "C" int cpb_viterbi_decode_host_packed(const cpbTrellis *t, const uint8_t *coded_packed_host, int64_t batch,
                                              int64_t n_in, int tb_depth, uint8_t *out_packed_host)
{
    if (t && batch == 0) return CPB_OK;
    if (!t || !coded_packed_host || !out_packed_host || batch < 0 || n_in <= 0 || (n_in % 8) != 0) return CPB_EINVAL;
    int64_t L = 0, T = 0;
    int rc = cpb_viterbi_sizes(t, n_in, &L, &T);
    if (rc) return rc;
    if ((L % 8) != 0) return CPB_EUNSUPPORTED;
    return pipe_run(const_cast<cpbTrellis *>(t)->pipe, {{coded_packed_host, nullptr, (size_t)(n_in / 8)}},
                    {{nullptr, out_packed_host, (size_t)(L / 8)}}, batch, pipe_chunk(batch, 32768, 2048),
                    [&](const std::vector<void *> &din, const std::vector<void *> &dout, int64_t, int64_t nb, cudaStream_t st) {
                        return cpb_viterbi_decode_packed(t, static_cast<const uint8_t *>(din[0]), nb, n_in, tb_depth,
                                                         static_cast<uint8_t *>(dout[0]), st);
                    });
}
