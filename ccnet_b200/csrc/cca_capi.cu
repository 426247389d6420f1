// C ABI of the operator (include/cca_b200.h): argument validation, kernel-family dispatch,
// host-buffer variants.  No torch types anywhere in this library.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cca_common.cuh"

namespace cca {
namespace {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
int cuda_fail(cudaError_t e, const char *where)
{
    return fail(CCA_ERR_CUDA, "CUDA error in %s: %s", where, cudaGetErrorString(e));
}

int check_dims(int B, int Cq, int C, int H, int W, int dtype)
{
    if (B <= 0 || Cq <= 0 || C <= 0 || H <= 0 || W <= 0)
        return fail(CCA_ERR_INVALID, "non-positive dimension%s%s");
    if (dtype != CCA_F32 && dtype != CCA_BF16) return fail(CCA_ERR_INVALID, "dtype must be CCA_F32 or CCA_BF16%s%s");
    if ((long long)B * C * H * W >= (1ll << 40)) return fail(CCA_ERR_UNSUPPORTED, "tensor too large%s%s");
    return CCA_OK;
}
size_t esize(int dtype) { return dtype == CCA_F32 ? 4 : 2; }
}  // namespace

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

static int g_l2_hints = -1;
static double g_l2_keep_mb = 80.0;
int tc_l2_hints()
{
    if (g_l2_hints < 0) {
        const char *e = getenv("CCA_B200_L2HINT"), *k = getenv("CCA_B200_L2KEEP_MB");
        g_l2_hints = e ? (atoi(e) != 0) : 0;      // off by default: no gain measured with two launches per pass
        if (k && atof(k) >= 0.0) g_l2_keep_mb = atof(k);
    }
    return g_l2_hints;
}
double tc_l2_keep_mb() { tc_l2_hints(); return g_l2_keep_mb; }
void set_tc_l2_hints(int on, double keep_mb) { g_l2_hints = on != 0; g_l2_keep_mb = keep_mb; }
static int g_pdl = -1;
int tc_pdl()
{
    if (g_pdl < 0) {
        const char *e = getenv("CCA_B200_PDL");
        g_pdl = e ? atoi(e) : 1;           // 0: off, 1: dependent launch, 2: + overlapped second pass (per-sample counters)
        if (g_pdl < 0 || g_pdl > 2) g_pdl = 1;
    }
    return g_pdl;
}
void set_tc_pdl(int on) { g_pdl = on < 0 || on > 2 ? 1 : on; }
}  // namespace cca

using namespace cca;

extern "C" {

int cca_b200_version(void) { return CCA_B200_VERSION; }
// profiling aid (not declared in the public header): device buffer of 2 x 4 x 512 int64 clock stamps
CCA_API void cca_b200__set_debug_buffer(void *p) { set_tc_debug_buffer(p); }
// A/B aid: run the tensor-core forward as two launches (column pass, row pass) instead of the fused launch
CCA_API void cca_b200__set_two_pass(int on) { set_tc_two_pass(on); }
CCA_API void cca_b200__set_bwd_debug_buffer(void *p) { set_tc_bwd_debug_buffer(p); }
// A/B aid: L2 eviction hints on / off and the evict_last budget in MB
CCA_API void cca_b200__set_l2_hints(int on, double keep_mb) { set_tc_l2_hints(on, keep_mb); }
CCA_API void cca_b200__set_pdl(int on) { set_tc_pdl(on); }
const char *cca_b200_last_error(void) { return g_err; }
const char *cca_b200_strerror(int s)
{
    switch (s) {
        case CCA_OK: return "ok";
        case CCA_ERR_INVALID: return "invalid argument";
        case CCA_ERR_UNSUPPORTED: return "unsupported shape";
        case CCA_ERR_WORKSPACE: return "workspace too small";
        case CCA_ERR_CUDA: return "CUDA error";
        case CCA_ERR_DEVICE: return "device is not sm_100";
        default: return "unknown status";
    }
}
unsigned long long cca_b200_launch_count(void) { return g_launches.load(); }

int cca_b200_device_ok(void)
{
    int dev = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    return major == 10 ? 1 : 0;
}

int cca_b200_tc_supported(int B, int Cq, int C, int H, int W, int dtype)
{
    if (check_dims(B, Cq, C, H, W, dtype)) return 0;
    return tc_forward_supported(Dims{B, Cq, C, H, W}, dtype) ? 1 : 0;
}

size_t cca_b200_workspace_bytes(int which, int B, int Cq, int C, int H, int W, int dtype)
{
    (void)Cq; (void)C; (void)dtype;
    const size_t pix = (size_t)B * H * W;
    // forward: per-pixel (m,l) of the column pass + per-sample completion counters of the fused launch;
    // backward: per-pixel delta = <dout, out> + the same counters
    const size_t counters = (((size_t)(2 * B + 2) * sizeof(unsigned int)) + 15) & ~(size_t)15;   // queue heads + per-sample counters
    return (which == CCA_WS_FORWARD ? pix * sizeof(float2) : pix * sizeof(float)) + counters;
}

int cca_b200_forward(const void *q, const void *k, const void *v, void *out, float *lse, void *ws, size_t ws_bytes,
                     int B, int Cq, int C, int H, int W, int dtype, unsigned flags, void *stream)
{
    int rc = check_dims(B, Cq, C, H, W, dtype);
    if (rc) return rc;
    if (!q || !k || !v || !out || !lse || !ws) return fail(CCA_ERR_INVALID, "null pointer%s%s");
    if (ws_bytes < cca_b200_workspace_bytes(CCA_WS_FORWARD, B, Cq, C, H, W, dtype))
        return fail(CCA_ERR_WORKSPACE, "forward workspace too small%s%s");
    if ((flags & CCA_FLAG_FORCE_SIMT) && (flags & CCA_FLAG_FORCE_TC))
        return fail(CCA_ERR_INVALID, "FORCE_SIMT and FORCE_TC are exclusive%s%s");
    const Dims d{B, Cq, C, H, W};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const char *why = "";
    const bool nhwc = (flags & CCA_FLAG_NHWC) != 0;
    const bool tc_ok = nhwc && tc_forward_supported(d, dtype);
    if ((flags & CCA_FLAG_FORCE_TC) && !tc_ok)
        return fail(CCA_ERR_UNSUPPORTED, "tensor-core forward needs CCA_FLAG_NHWC and a covered shape%s%s");
    if (nhwc && (!tc_ok || (flags & CCA_FLAG_FORCE_SIMT)))
        return fail(CCA_ERR_UNSUPPORTED, "channels-last tensors are only handled by the tensor-core kernels; pass NCHW%s%s");
    cudaError_t e;
    if (tc_ok) {
        if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
             reinterpret_cast<uintptr_t>(out)) & 15)
            return fail(CCA_ERR_INVALID, "tensor-core path needs 16-byte aligned tensors%s%s");
        e = tc_forward(q, k, v, out, lse, ws, d, dtype, st, &why);
        if (e != cudaSuccess) return cuda_fail(e, why && *why ? why : "tc_forward");
        return CCA_OK;
    }
    if (!simt_supported(d, false)) return fail(CCA_ERR_UNSUPPORTED, "H or W too large for the generic kernels%s%s");
    e = simt_forward(q, k, v, out, lse, ws, d, dtype, st, &why);
    if (e != cudaSuccess) return cuda_fail(e, "simt_forward");
    return CCA_OK;
}

int cca_b200_backward(const void *dout, const void *q, const void *k, const void *v, const void *out,
                      const float *lse, void *dq, void *dk, void *dv, void *ws, size_t ws_bytes,
                      int B, int Cq, int C, int H, int W, int dtype, unsigned flags, void *stream)
{
    int rc = check_dims(B, Cq, C, H, W, dtype);
    if (rc) return rc;
    if (!dout || !q || !k || !v || !out || !lse || !dq || !dk || !dv || !ws)
        return fail(CCA_ERR_INVALID, "null pointer%s%s");
    if (ws_bytes < cca_b200_workspace_bytes(CCA_WS_BACKWARD, B, Cq, C, H, W, dtype))
        return fail(CCA_ERR_WORKSPACE, "backward workspace too small%s%s");
    const Dims d{B, Cq, C, H, W};
    const char *why = "";
    const bool nhwc = (flags & CCA_FLAG_NHWC) != 0;
    const bool tc_ok = nhwc && tc_backward_supported(d, dtype);
    if ((flags & CCA_FLAG_FORCE_TC) && !tc_ok)
        return fail(CCA_ERR_UNSUPPORTED, "tensor-core backward needs CCA_FLAG_NHWC and a covered shape%s%s");
    if (nhwc && (!tc_ok || (flags & CCA_FLAG_FORCE_SIMT)))
        return fail(CCA_ERR_UNSUPPORTED, "channels-last tensors are only handled by the tensor-core kernels; pass NCHW%s%s");
    if (tc_ok) {
        if ((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
             reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dq) |
             reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 15)
            return fail(CCA_ERR_INVALID, "tensor-core path needs 16-byte aligned tensors%s%s");
        cudaError_t e = tc_backward(dout, q, k, v, out, lse, dq, dk, dv, ws, d, dtype,
                                    reinterpret_cast<cudaStream_t>(stream), &why);
        if (e != cudaSuccess) return cuda_fail(e, why && *why ? why : "tc_backward");
        return CCA_OK;
    }
    if (!simt_supported(d, true)) return fail(CCA_ERR_UNSUPPORTED, "H or W too large for the generic kernels%s%s");
    cudaError_t e = simt_backward(dout, q, k, v, out, lse, dq, dk, dv, ws, d, dtype,
                                  reinterpret_cast<cudaStream_t>(stream), &why);
    if (e != cudaSuccess) return cuda_fail(e, "simt_backward");
    return CCA_OK;
}

// ---------------------------------------------------------------------------------------
// host-buffer variants
// ---------------------------------------------------------------------------------------
namespace {
struct DevBufs {
    static constexpr int kMax = 12;
    void *p[kMax] = {};
    int n = 0;
    cudaStream_t st = nullptr;
    ~DevBufs()
    {
        for (int i = 0; i < n; ++i) cudaFree(p[i]);
        if (st) cudaStreamDestroy(st);
    }
    void *alloc(size_t bytes, cudaError_t &e)
    {
        void *r = nullptr;
        if (e == cudaSuccess) e = cudaMalloc(&r, bytes ? bytes : 1);
        if (e == cudaSuccess) p[n++] = r;
        return r;
    }
};
}  // namespace

int cca_b200_forward_host(const void *q, const void *k, const void *v, void *out, float *lse,
                          int B, int Cq, int C, int H, int W, int dtype, unsigned flags)
{
    int rc = check_dims(B, Cq, C, H, W, dtype);
    if (rc) return rc;
    if (!q || !k || !v || !out || !lse) return fail(CCA_ERR_INVALID, "null pointer%s%s");
    const size_t px = (size_t)B * H * W, es = esize(dtype);
    const size_t nq = px * Cq * es, nv = px * C * es, nl = px * sizeof(float);
    const size_t nws = cca_b200_workspace_bytes(CCA_WS_FORWARD, B, Cq, C, H, W, dtype);
    DevBufs d;
    cudaError_t e = cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking);
    void *dq = d.alloc(nq, e), *dk = d.alloc(nq, e), *dv = d.alloc(nv, e), *dout = d.alloc(nv, e);
    void *dl = d.alloc(nl, e), *dws = d.alloc(nws, e);
    if (e != cudaSuccess) return cuda_fail(e, "forward_host alloc");
    if ((e = cudaMemcpyAsync(dq, q, nq, cudaMemcpyHostToDevice, d.st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(dk, k, nq, cudaMemcpyHostToDevice, d.st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(dv, v, nv, cudaMemcpyHostToDevice, d.st)) != cudaSuccess)
        return cuda_fail(e, "forward_host H2D");
    rc = cca_b200_forward(dq, dk, dv, dout, (float *)dl, dws, nws, B, Cq, C, H, W, dtype, flags, d.st);
    if (rc) return rc;
    if ((e = cudaMemcpyAsync(out, dout, nv, cudaMemcpyDeviceToHost, d.st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(lse, dl, nl, cudaMemcpyDeviceToHost, d.st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(d.st)) != cudaSuccess)
        return cuda_fail(e, "forward_host D2H");
    return CCA_OK;
}

int cca_b200_backward_host(const void *dout, const void *q, const void *k, const void *v, const void *out,
                           const float *lse, void *dq, void *dk, void *dv,
                           int B, int Cq, int C, int H, int W, int dtype, unsigned flags)
{
    int rc = check_dims(B, Cq, C, H, W, dtype);
    if (rc) return rc;
    if (!dout || !q || !k || !v || !out || !lse || !dq || !dk || !dv) return fail(CCA_ERR_INVALID, "null pointer%s%s");
    const size_t px = (size_t)B * H * W, es = esize(dtype);
    const size_t nq = px * Cq * es, nv = px * C * es, nl = px * sizeof(float);
    const size_t nws = cca_b200_workspace_bytes(CCA_WS_BACKWARD, B, Cq, C, H, W, dtype);
    DevBufs d;
    cudaError_t e = cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking);
    void *g = d.alloc(nv, e), *tq = d.alloc(nq, e), *tk = d.alloc(nq, e), *tv = d.alloc(nv, e), *to = d.alloc(nv, e);
    void *tl = d.alloc(nl, e), *gq = d.alloc(nq, e), *gk = d.alloc(nq, e), *gv = d.alloc(nv, e), *ws = d.alloc(nws, e);
    if (e != cudaSuccess) return cuda_fail(e, "backward_host alloc");
    if ((e = cudaMemcpyAsync(g, dout, nv, cudaMemcpyHostToDevice, d.st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(tq, q, nq, cudaMemcpyHostToDevice, d.st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(tk, k, nq, cudaMemcpyHostToDevice, d.st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(tv, v, nv, cudaMemcpyHostToDevice, d.st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(to, out, nv, cudaMemcpyHostToDevice, d.st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(tl, lse, nl, cudaMemcpyHostToDevice, d.st)) != cudaSuccess)
        return cuda_fail(e, "backward_host H2D");
    rc = cca_b200_backward(g, tq, tk, tv, to, (const float *)tl, gq, gk, gv, ws, nws, B, Cq, C, H, W, dtype, flags, d.st);
    if (rc) return rc;
    if ((e = cudaMemcpyAsync(dq, gq, nq, cudaMemcpyDeviceToHost, d.st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(dk, gk, nq, cudaMemcpyDeviceToHost, d.st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(dv, gv, nv, cudaMemcpyDeviceToHost, d.st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(d.st)) != cudaSuccess)
        return cuda_fail(e, "backward_host D2H");
    return CCA_OK;
}

}  // extern "C"
