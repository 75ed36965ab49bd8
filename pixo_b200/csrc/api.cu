// api.cu — the extern "C" surface of libpixo_b200.so (see include/pixo_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <emmintrin.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

#include "common.cuh"
#include "jpeg_host.hpp"

namespace pixo {

static thread_local std::string g_thread_err;

int set_error(pixo_b200_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_thread_err = buf;
    return code;
}

int cuda_fail(pixo_b200_ctx *ctx, cudaError_t e, const char *what)
{
    const int code = e == cudaErrorMemoryAllocation ? PIXO_B200_ERR_OOM : PIXO_B200_ERR_CUDA;
    return set_error(ctx, code, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

int ensure_dev(pixo_b200_ctx *ctx, Scratch &s, size_t bytes)
{
    if (s.cap >= bytes) return 0;
    if (s.ptr) {
        PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PIXO_CUDA(ctx, cudaFree(s.ptr));
        s.ptr = nullptr;
        s.cap = 0;
    }
    const size_t want = bytes + bytes / 8 + 256;
    PIXO_CUDA(ctx, cudaMalloc(&s.ptr, want));
    s.cap = want;
    return 0;
}

int ensure_pinned(pixo_b200_ctx *ctx, Scratch &s, size_t bytes)
{
    if (s.cap >= bytes) return 0;
    if (s.ptr) {
        PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PIXO_CUDA(ctx, cudaFreeHost(s.ptr));
        s.ptr = nullptr;
        s.cap = 0;
    }
    const size_t want = bytes + bytes / 8 + 256;
    PIXO_CUDA(ctx, cudaMallocHost(&s.ptr, want));
    s.cap = want;
    return 0;
}

HostPool::HostPool(int nthreads)
{
    for (int i = 0; i < nthreads; ++i) threads_.emplace_back([this] { worker(); });
}

HostPool::~HostPool()
{
    {
        std::lock_guard<std::mutex> lk(m_);
        stop_ = true;
    }
    cv_work_.notify_all();
    for (auto &t : threads_) t.join();
}

void HostPool::worker()
{
    uint64_t seen = 0;
    for (;;) {
        const std::function<void(int)> *fn;
        int n;
        {
            std::unique_lock<std::mutex> lk(m_);
            cv_work_.wait(lk, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            fn = fn_;
            n = njobs_;
            ++active_;
        }
        for (int j; (j = next_.fetch_add(1)) < n;) (*fn)(j);
        {
            std::lock_guard<std::mutex> lk(m_);
            if (--active_ == 0) cv_done_.notify_all();
        }
    }
}

void HostPool::run(int njobs, const std::function<void(int)> &fn)
{
    if (njobs <= 0) return;
    {
        std::lock_guard<std::mutex> lk(m_);
        fn_ = &fn;
        njobs_ = njobs;
        next_.store(0);
        ++generation_;
    }
    if (njobs > 1) cv_work_.notify_all();
    for (int j; (j = next_.fetch_add(1)) < njobs;) fn(j);
    // every job has been claimed; wait for the workers that are still inside one (a worker that
    // wakes up late finds nothing to claim and leaves at once)
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return active_ == 0; });
    fn_ = nullptr;
    njobs_ = 0;
}

static HostPool *host_pool(pixo_b200_ctx *ctx)
{
    if (!ctx->pool) {
        int n = ctx->host_threads - 1;
        n = n < 1 ? 1 : (n > 5 ? 5 : n);   // five helpers + the caller saturate a socket's copy bandwidth
        ctx->pool = new HostPool(n);
    }
    return ctx->pool;
}

// Copy into a pinned staging slot with NON-TEMPORAL stores.  An ordinary memcpy of a 1 MB piece
// leaves the bytes dirty in the copying core's cache, and the DMA engine then has to pull every
// line out of that cache (measured: the link ran at ~17 GB/s behind plain memcpy); streaming stores
// go to memory and the DMA reads it at the pinned rate.
static void stage_copy(void *dst, const void *src, size_t n)
{
#if defined(__x86_64__) || defined(__SSE2__)
    auto *d = reinterpret_cast<uint8_t *>(dst);
    auto *s = reinterpret_cast<const uint8_t *>(src);
    if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 32));
            const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i *>(d + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 48), e);
        }
        _mm_sfence();
        if (i < n) memcpy(d + i, s + i, n - i);
        return;
    }
#endif
    memcpy(dst, src, n);
}

static bool is_page_locked(const void *p)
{
    cudaPointerAttributes at;
    const bool locked = cudaPointerGetAttributes(&at, p) == cudaSuccess &&
                        (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged);
    cudaGetLastError();  // an unregistered pointer is not an error here
    return locked;
}

// Host -> device copy that does not depend on the caller's buffer being page-locked.  pixo's
// callers hand over ordinary (pageable) memory, and the driver's own pageable path moves it at
// 11-18 GB/s (one staging thread).  Here the context's host threads copy 1 MB pieces into a ring of
// pinned slots and each queues the DMA of its piece as soon as it is staged, so the host copies run
// in parallel with each other and with the DMA engine, and the link runs near its pinned rate
// (measured on the B200 box: a pageable 25 MB 4K frame 1.4 ms -> see DESIGN.md section 5).
// Page-locked sources, and small ones, go straight to cudaMemcpyAsync.
static int h2d_copy(pixo_b200_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    constexpr size_t SLOT = (size_t)1 << 20;
    constexpr int NSLOT = 32;
    if (bytes < 4 * SLOT || is_page_locked(src)) {
        PIXO_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    }
    PIXO_TRY(ensure_pinned(ctx, ctx->h_in, SLOT * NSLOT));
    while (ctx->stage_events.size() < (size_t)NSLOT) {
        cudaEvent_t ev;
        PIXO_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        PIXO_CUDA(ctx, cudaEventRecord(ev, st));
        ctx->stage_events.push_back(ev);
    }
    const int nchunks = (int)((bytes + SLOT - 1) / SLOT);
    std::atomic<int> first_error{(int)cudaSuccess};
    const int device = ctx->device;
    // piece c uses slot c % NSLOT; pieces are claimed in order, so a slot's previous user is always
    // NSLOT pieces back and its DMA has normally drained long before the slot comes round again
    auto piece = [&](int c) {
        if (first_error.load() != (int)cudaSuccess) return;
        cudaError_t e = cudaSetDevice(device);
        const int slot_i = c % NSLOT;
        uint8_t *slot = reinterpret_cast<uint8_t *>(ctx->h_in.ptr) + (size_t)slot_i * SLOT;
        const size_t off = (size_t)c * SLOT, n = std::min(SLOT, bytes - off);
        if (e == cudaSuccess) e = cudaEventSynchronize(ctx->stage_events[slot_i]);   // the slot's previous DMA (this call's or the last one's) has drained
        if (e == cudaSuccess) {
            stage_copy(slot, reinterpret_cast<const uint8_t *>(src) + off, n);
            e = cudaMemcpyAsync(reinterpret_cast<uint8_t *>(dst) + off, slot, n, cudaMemcpyHostToDevice, st);
        }
        if (e == cudaSuccess && c + NSLOT < nchunks) e = cudaEventRecord(ctx->stage_events[slot_i], st);
        if (e != cudaSuccess) { int ok = (int)cudaSuccess; first_error.compare_exchange_strong(ok, (int)e); }
    };
    if (nchunks > NSLOT) {
        // a slot is reused: its event must be recorded by the piece that used it before.  Run the
        // pieces in rounds of NSLOT so that "previous user" is always in an earlier round.
        for (int r0 = 0; r0 < nchunks; r0 += NSLOT) {
            const int cnt = std::min(NSLOT, nchunks - r0);
            host_pool(ctx)->run(cnt, [&](int j) { piece(r0 + j); });
        }
    } else {
        host_pool(ctx)->run(nchunks, piece);
    }
    // the slots may be rewritten by the next call: make that call wait for this one's DMAs
    PIXO_CUDA(ctx, (cudaError_t)first_error.load());
    for (int s_i = 0; s_i < std::min(nchunks, NSLOT); ++s_i) PIXO_CUDA(ctx, cudaEventRecord(ctx->stage_events[s_i], st));
    return 0;
}

// Device -> pageable host: through the pinned ring with the pool copying out, for results big
// enough to matter (a 4K JPEG is ~3 MB).  Synchronous: returns when `dst` holds the bytes.
static int d2h_copy_sync(pixo_b200_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    constexpr size_t PIECE = (size_t)512 << 10;
    if (bytes < 2 * PIECE || is_page_locked(dst)) {
        PIXO_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        PIXO_CUDA(ctx, cudaStreamSynchronize(st));
        return 0;
    }
    const bool dbg = getenv("PIXO_B200_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = dbg ? now() : 0;
    PIXO_TRY(ensure_pinned(ctx, ctx->h_out, bytes));
    const double t1 = dbg ? now() : 0;
    PIXO_CUDA(ctx, cudaMemcpyAsync(ctx->h_out.ptr, src, bytes, cudaMemcpyDeviceToHost, st));
    PIXO_CUDA(ctx, cudaStreamSynchronize(st));
    if (dbg) fprintf(stderr, "  d2h: ensure %.0f us, dma of %zu B %.0f us\n", t1 - t0, bytes, now() - t1);
    const int n = (int)((bytes + PIECE - 1) / PIECE);
    host_pool(ctx)->run(n, [&](int j) {
        const size_t off = (size_t)j * PIECE, len = std::min(PIECE, bytes - off);
        memcpy(reinterpret_cast<uint8_t *>(dst) + off, reinterpret_cast<const uint8_t *>(ctx->h_out.ptr) + off, len);
    });
    return 0;
}

static int validate_jpeg(pixo_b200_ctx *ctx, uint32_t w, uint32_t h, uint32_t color_type,
                         uint32_t subsampling)
{
    // order follows encode_into, src/jpeg/mod.rs:333-373
    if (w == 0 || h == 0)
        return set_error(ctx, PIXO_B200_ERR_INVALID_DIMENSIONS, "Invalid image dimensions: %ux%u", w, h);
    if (w > 65535 || h > 65535)
        return set_error(ctx, PIXO_B200_ERR_IMAGE_TOO_LARGE, "Image dimensions %ux%u exceed maximum 65535", w, h);
    if (color_type != PIXO_B200_RGB && color_type != PIXO_B200_GRAY)
        return set_error(ctx, PIXO_B200_ERR_UNSUPPORTED_COLOR, "Unsupported color type for this format");
    if (subsampling != PIXO_B200_S444 && subsampling != PIXO_B200_S420)
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "unknown subsampling %u", subsampling);
    return 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace pixo

using namespace pixo;

extern "C" {

int pixo_b200_version(void) { return PIXO_B200_VERSION; }

int pixo_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pixo_b200_ctx_create(int device, pixo_b200_ctx **out)
{
    if (!out) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx out pointer is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error(nullptr, PIXO_B200_ERR_CUDA,
                         "no CUDA device available (%s); libpixo_b200 has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n)
        return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "device %d out of range (0..%d)", device, n - 1);
    pixo_b200_ctx *ctx = new pixo_b200_ctx();
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        const int rc = cuda_fail(nullptr, e, "cudaSetDevice");
        delete ctx;
        return rc;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        const int rc = cuda_fail(nullptr, e, "cudaGetDeviceProperties");
        delete ctx;
        return rc;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        const int rc = cuda_fail(nullptr, e, "cudaStreamCreate");
        delete ctx;
        return rc;
    }
    ctx->stream = ctx->own_stream;
    unsigned hc = std::thread::hardware_concurrency();
    ctx->host_threads = hc ? (int)hc : 1;
    *out = ctx;
    return 0;
}

void pixo_b200_ctx_destroy(pixo_b200_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    Scratch *dev[] = {&ctx->d_in, &ctx->d_y, &ctx->d_cb, &ctx->d_cr, &ctx->d_misc, &ctx->d_out, &ctx->d_ent, &ctx->d_coef, &ctx->d_retry, &ctx->d_raw};
    for (Scratch *s : dev) if (s->ptr) cudaFree(s->ptr);
    Scratch *host[] = {&ctx->h_in, &ctx->h_out, &ctx->h_misc};
    for (Scratch *s : host) if (s->ptr) cudaFreeHost(s->ptr);
    for (cudaEvent_t ev : ctx->events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->stage_events) cudaEventDestroy(ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    delete ctx->pool;
    delete ctx;
}

const char *pixo_b200_last_error(const pixo_b200_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_thread_err.c_str();
}

int pixo_b200_ctx_set_stream(pixo_b200_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return 0;
}

void *pixo_b200_ctx_stream(pixo_b200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int pixo_b200_ctx_sync(pixo_b200_ctx *ctx)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

uint64_t pixo_b200_ctx_launch_count(const pixo_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pixo_b200_ctx_set_host_threads(pixo_b200_ctx *ctx, int n)
{
    if (!ctx || n < 1) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "bad host thread count");
    ctx->host_threads = n;
    return 0;
}

uint64_t pixo_b200_ctx_host_fallbacks(const pixo_b200_ctx *ctx) { return ctx ? ctx->host_fallbacks : 0; }

int pixo_b200_ctx_set_scan_capacity(pixo_b200_ctx *ctx, size_t bytes_per_frame, int gpu_retry)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    ctx->scan_cap_override = bytes_per_frame;
    ctx->gpu_retry = gpu_retry != 0;
    return 0;
}

int pixo_b200_dev_alloc(pixo_b200_ctx *ctx, size_t bytes, void **dptr)
{
    if (!ctx || !dptr) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    PIXO_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    return 0;
}

int pixo_b200_dev_free(pixo_b200_ctx *ctx, void *dptr)
{
    if (!ctx) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    PIXO_CUDA(ctx, cudaFree(dptr));
    return 0;
}

int pixo_b200_host_alloc_pinned(pixo_b200_ctx *ctx, size_t bytes, void **hptr)
{
    if (!ctx || !hptr) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    PIXO_CUDA(ctx, cudaMallocHost(hptr, bytes ? bytes : 1));
    return 0;
}

int pixo_b200_host_free_pinned(pixo_b200_ctx *ctx, void *hptr)
{
    if (!ctx) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    PIXO_CUDA(ctx, cudaFreeHost(hptr));
    return 0;
}

int pixo_b200_upload(pixo_b200_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes)
{
    if (!ctx) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    PIXO_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int pixo_b200_download(pixo_b200_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes)
{
    if (!ctx) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    PIXO_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- JPEG ---------------------------------------------------------------------------------

void pixo_b200_quant_tables(int quality, uint8_t lum_zz[64], uint8_t chr_zz[64], float lum[64],
                            float chr[64])
{
    quant_tables(quality, lum_zz, chr_zz, lum, chr);
}

int pixo_b200_jpeg_block_counts(uint32_t width, uint32_t height, uint32_t color_type,
                                uint32_t subsampling, size_t *ny, size_t *nc)
{
    PIXO_TRY(validate_jpeg(nullptr, width, height, color_type, subsampling));
    const FrameGeometry g = make_geometry(width, height, color_type, subsampling);
    if (ny) *ny = g.ny;
    if (nc) *nc = g.nc;
    return 0;
}

int pixo_b200_jpeg_coefficients_dev(pixo_b200_ctx *ctx, const uint8_t *d_pixels,
                                    size_t pixel_stride, uint32_t n_images, uint32_t width,
                                    uint32_t height, uint32_t color_type, uint32_t subsampling,
                                    const float lum_q[64], const float chr_q[64], int16_t *d_y,
                                    size_t y_stride, int16_t *d_cb, int16_t *d_cr,
                                    size_t c_stride, uint32_t flags, uint64_t *d_hist)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    PIXO_TRY(validate_jpeg(ctx, width, height, color_type, subsampling));
    if (!d_pixels || !d_y || !lum_q || !chr_q)
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    if (color_type != PIXO_B200_GRAY && (!d_cb || !d_cr))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null chroma buffer");
    if (n_images == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(d_y) & 15) || (y_stride & 7) ||
        (d_cb && ((reinterpret_cast<uintptr_t>(d_cb) & 15) || (reinterpret_cast<uintptr_t>(d_cr) & 15) || (c_stride & 7))))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT,
                         "coefficient buffers must be 16-byte aligned with strides multiple of 8");
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    PIXO_TRY(launch_jpeg_transform(ctx, d_pixels, pixel_stride, n_images, width, height,
                                   color_type, subsampling, lum_q, chr_q, d_y, y_stride, d_cb,
                                   d_cr, c_stride, flags));
    if (d_hist) {
        const FrameGeometry g = make_geometry(width, height, color_type, subsampling);
        PIXO_TRY(launch_jpeg_histogram(ctx, d_y, y_stride, d_cb, d_cr, c_stride, n_images, g.ny,
                                       g.nc, g.y_per_mcu, 0, (flags & PIXO_B200_COEF_ZIGZAG) != 0,
                                       d_hist));
    }
    return 0;
}

// shared by the host-buffer entry points: upload, transform (+hist), download coefficients
static int transform_host(pixo_b200_ctx *ctx, const uint8_t *pixels, const FrameGeometry &g,
                          const float lum_q[64], const float chr_q[64], uint32_t flags,
                          uint32_t restart_interval, bool want_hist, int16_t *y, int16_t *cb,
                          int16_t *cr, uint64_t *hist)
{
    const size_t bpp = g.color_type == PIXO_B200_GRAY ? 1 : 3;
    const size_t in_bytes = (size_t)g.width * g.height * bpp;
    const size_t yb = g.ny * 64 * sizeof(int16_t), cbb = g.nc * 64 * sizeof(int16_t);
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    PIXO_TRY(ensure_dev(ctx, ctx->d_in, in_bytes));
    PIXO_TRY(ensure_dev(ctx, ctx->d_y, yb));
    if (cbb) {
        PIXO_TRY(ensure_dev(ctx, ctx->d_cb, cbb));
        PIXO_TRY(ensure_dev(ctx, ctx->d_cr, cbb));
    }
    PIXO_TRY(h2d_copy(ctx, ctx->d_in.ptr, pixels, in_bytes, ctx->stream));
    auto *dy = reinterpret_cast<int16_t *>(ctx->d_y.ptr);
    auto *dcb = reinterpret_cast<int16_t *>(ctx->d_cb.ptr);
    auto *dcr = reinterpret_cast<int16_t *>(ctx->d_cr.ptr);
    PIXO_TRY(launch_jpeg_transform(ctx, reinterpret_cast<const uint8_t *>(ctx->d_in.ptr), in_bytes, 1,
                                   g.width, g.height, g.color_type, g.subsampling, lum_q, chr_q,
                                   dy, g.ny * 64, dcb, dcr, g.nc * 64, flags));
    if (want_hist) {
        PIXO_TRY(ensure_dev(ctx, ctx->d_out, kHistWords * sizeof(uint64_t)));
        PIXO_TRY(launch_jpeg_histogram(ctx, dy, g.ny * 64, dcb, dcr, g.nc * 64, 1, g.ny, g.nc,
                                       g.y_per_mcu, restart_interval,
                                       (flags & PIXO_B200_COEF_ZIGZAG) != 0,
                                       reinterpret_cast<uint64_t *>(ctx->d_out.ptr)));
        PIXO_CUDA(ctx, cudaMemcpyAsync(hist, ctx->d_out.ptr, kHistWords * sizeof(uint64_t),
                                       cudaMemcpyDeviceToHost, ctx->stream));
    }
    PIXO_CUDA(ctx, cudaMemcpyAsync(y, dy, yb, cudaMemcpyDeviceToHost, ctx->stream));
    if (cbb) {
        PIXO_CUDA(ctx, cudaMemcpyAsync(cb, dcb, cbb, cudaMemcpyDeviceToHost, ctx->stream));
        PIXO_CUDA(ctx, cudaMemcpyAsync(cr, dcr, cbb, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int pixo_b200_jpeg_coefficients(pixo_b200_ctx *ctx, const uint8_t *pixels, uint32_t width,
                                uint32_t height, uint32_t color_type, uint32_t subsampling,
                                const float lum_q[64], const float chr_q[64], int16_t *y,
                                int16_t *cb, int16_t *cr, uint32_t flags, uint64_t *hist)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    PIXO_TRY(validate_jpeg(ctx, width, height, color_type, subsampling));
    if (!pixels || !y || !lum_q || !chr_q || (color_type != PIXO_B200_GRAY && (!cb || !cr)))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    const FrameGeometry g = make_geometry(width, height, color_type, subsampling);
    return transform_host(ctx, pixels, g, lum_q, chr_q, flags, 0, hist != nullptr, y, cb, cr, hist);
}

static int entropy_from_host_arrays(pixo_b200_ctx *ctx, const int16_t *y, const int16_t *cb,
                                    const int16_t *cr, const FrameGeometry &g, uint32_t quality,
                                    uint32_t restart_interval, const uint64_t *hist,
                                    uint8_t *out, size_t out_cap, size_t *out_len, int threads)
{
    uint8_t lum_zz[64], chr_zz[64];
    quant_tables((int)quality, lum_zz, chr_zz, nullptr, nullptr);
    HuffTables t;
    // build_optimized_huffman_tables(..).unwrap_or_default(), src/jpeg/mod.rs:379-392
    if (!(hist && huff_from_histogram(hist, g.has_chroma, t))) huff_standard(t);
    if (out_cap < 1024 + 2)
        return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "output capacity %zu too small", out_cap);
    size_t n = write_headers(out, g, lum_zz, chr_zz, t, restart_interval);
    const size_t body = entropy_encode_scan(y, cb, cr, g, t, restart_interval, false, out + n,
                                            out_cap - n - 2, threads);
    if (body == (size_t)-1)
        return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "output capacity %zu too small", out_cap);
    n += body;
    out[n++] = 0xFF;
    out[n++] = 0xD9;
    *out_len = n;
    return 0;
}

static int validate_encode(pixo_b200_ctx *ctx, size_t pixels_len, uint32_t width, uint32_t height,
                           uint32_t color_type, uint32_t quality, uint32_t subsampling,
                           uint32_t restart_interval, bool restart_given_zero)
{
    // encode_into validation order, src/jpeg/mod.rs:333-373
    if (quality == 0 || quality > 100)
        return set_error(ctx, PIXO_B200_ERR_INVALID_QUALITY, "Invalid quality %u: must be 1-100", quality);
    (void)restart_given_zero;
    if (restart_interval > 65535)
        return set_error(ctx, PIXO_B200_ERR_INVALID_RESTART, "Invalid restart interval %u", restart_interval);
    PIXO_TRY(validate_jpeg(ctx, width, height, color_type, subsampling));
    const size_t bpp = color_type == PIXO_B200_GRAY ? 1 : 3;
    const size_t expected = (size_t)width * height * bpp;
    if (pixels_len != expected)
        return set_error(ctx, PIXO_B200_ERR_INVALID_DATA_LENGTH,
                         "Invalid data length: expected %zu bytes, got %zu", expected, pixels_len);
    return 0;
}

// Synchronises every stream of the context when an entry point leaves early, so that no queued
// copy still reads the caller's pixels or writes the caller's output after the error return.
struct DrainOnError {
    pixo_b200_ctx *ctx;
    bool armed = true;
    explicit DrainOnError(pixo_b200_ctx *c) : ctx(c) {}
    ~DrainOnError()
    {
        if (!armed) return;
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->d2h_stream);
        cudaGetLastError();
    }
};

// Device scan capacity per frame: a JPEG that needs more than half its raw size (noise at very
// high quality) is coded a second time with the exact size the kernel reported.
static uint64_t default_scan_cap(const pixo_b200_ctx *ctx, size_t raw_bytes)
{
    const size_t want = ctx->scan_cap_override ? ctx->scan_cap_override : (raw_bytes / 2 + 65536) / 8 * 9;
    return align_up(want < 1024 ? 1024 : want, 256);
}

// Encode n frames of identical geometry and options.  GPU: colour/DCT/quantise (K1/K2), symbol
// statistics when optimize_huffman (K3), Huffman bit packing + 0xFF stuffing + restart markers
// (k_huff); host: headers, optimised-table construction, EOI.  Frames are processed in groups of
// up to 16 (about 96 MB of input) on three streams: the H2D copy of group g+1 (copy stream) and the D2H copy of group
// g-1's scan bytes (d2h stream) run under the kernels of group g, and the host never drains the
// compute stream between groups - it only waits for the event behind a group's 12-byte-per-frame
// length readback before it queues that group's D2H.  Only finished scan bytes come back over
// PCIe.  A frame whose scan outgrows the device buffer is coded again on the GPU with the exact
// size; the host entropy coder (same GPU coefficient arrays) is the last resort for a faulted
// device stage and is counted in ctx->host_fallbacks.
static int encode_frames(pixo_b200_ctx *ctx, const uint8_t *pixels, size_t len_each, uint32_t n_images,
                         const FrameGeometry &g, uint32_t quality, uint32_t restart_interval,
                         bool optimize, uint8_t *out, size_t out_cap_each, size_t *out_lens)
{
    if (out_cap_each < 1024 + 2)  // before any GPU work is queued
        return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "output capacity %zu too small", out_cap_each);
    float lum[64], chr[64];
    uint8_t lum_zz[64], chr_zz[64];
    quant_tables((int)quality, lum_zz, chr_zz, lum, chr);
    const size_t yb = align_up(g.ny * 64 * sizeof(int16_t), 256);
    const size_t cbb = align_up(g.nc * 64 * sizeof(int16_t), 256);
    const size_t coef_each = yb + 2 * cbb;
    const size_t in_stride = align_up(len_each, 256);
    const uint64_t scan_cap = default_scan_cap(ctx, len_each);
    const size_t ent_one = entropy_scratch_bytes(1, g, restart_interval);

    // Groups of about 96 MB of input (4 frames at 4K, 15 at 1080p), at least two per call: long enough for
    // full-rate DMA and to amortise the launches, short enough that what no upload can hide - the last
    // group's kernels and the read-back of its scan bytes - stays small (with 2 x 16 frames that tail was
    // 10 % of a 32-frame call: 16.4 -> 17.1 Gpix/s end to end, measured).
    uint32_t G = (uint32_t)std::min<size_t>(16, std::max<size_t>(1, (((size_t)96 << 20) + len_each / 2) / len_each));
    G = std::min(G, std::max(1u, (n_images + 1) / 2));
    const size_t budget = (size_t)4 << 30;
    auto ent_bytes = [&](uint32_t k) {  // per-image tables run one k_huff pass per image, each with its own scratch
        return optimize ? (size_t)k * ent_one : entropy_scratch_bytes(k, g, restart_interval);
    };
    auto group_bytes = [&](uint32_t k) {
        return 2 * (size_t)k * in_stride + 2 * (size_t)k * coef_each + ent_bytes(k) + 2 * (size_t)k * scan_cap;
    };
    while (G > 1 && group_bytes(G) > budget) --G;
    const uint32_t ngroups = (n_images + G - 1) / G;

    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    PIXO_TRY(ensure_dev(ctx, ctx->d_in, 2 * (size_t)G * in_stride));
    PIXO_TRY(ensure_dev(ctx, ctx->d_coef, 2 * (size_t)G * coef_each));
    PIXO_TRY(ensure_dev(ctx, ctx->d_ent, ent_bytes(G)));
    PIXO_TRY(ensure_dev(ctx, ctx->d_out, 2 * (size_t)G * scan_cap));
    PIXO_TRY(ensure_dev(ctx, ctx->d_misc, (size_t)G * kHistWords * sizeof(uint64_t) + 256));
    const size_t meta_slot = align_up((size_t)G * 12, 256);
    PIXO_TRY(ensure_pinned(ctx, ctx->h_misc, 2 * meta_slot + (size_t)G * kHistWords * sizeof(uint64_t) + 256));
    while (ctx->events.size() < 8) {
        cudaEvent_t ev;
        PIXO_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->events.push_back(ev);
    }
    cudaEvent_t *ev_in = &ctx->events[0];    // [2] input slot filled
    cudaEvent_t *ev_used = &ctx->events[2];  // [2] input slot consumed by the transform kernel
    cudaEvent_t *ev_out = &ctx->events[4];   // [2] scan bytes of the slot copied back
    cudaEvent_t *ev_len = &ctx->events[6];   // [2] the slot's lengths / overflow flags are on the host
    auto *d_in = reinterpret_cast<uint8_t *>(ctx->d_in.ptr);
    auto *d_scan = reinterpret_cast<uint8_t *>(ctx->d_out.ptr);
    auto *h_meta = reinterpret_cast<uint8_t *>(ctx->h_misc.ptr);
    auto *h_hist = reinterpret_cast<uint64_t *>(h_meta + 2 * meta_slot);
    auto h_len_of = [&](int slot) { return reinterpret_cast<uint64_t *>(h_meta + (size_t)slot * meta_slot); };
    auto h_ovf_of = [&](int slot) { return reinterpret_cast<uint32_t *>(h_meta + (size_t)slot * meta_slot + (size_t)G * 8); };
    struct Coef { int16_t *y, *cb, *cr; };
    const size_t cstride = coef_each / 2;
    auto coef_of = [&](int slot) {
        uint8_t *base = reinterpret_cast<uint8_t *>(ctx->d_coef.ptr) + (size_t)slot * G * coef_each;
        return Coef{reinterpret_cast<int16_t *>(base), reinterpret_cast<int16_t *>(base + yb),
                    reinterpret_cast<int16_t *>(base + yb + cbb)};
    };
    std::vector<HuffTables> tables[2];
    const bool out_locked = is_page_locked(out);
    DrainOnError drain(ctx);
    const bool dbg = getenv("PIXO_B200_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };

    auto upload = [&](uint32_t gi) -> int {
        const uint32_t first = gi * G, cnt = std::min(G, n_images - first);
        const int slot = (int)(gi & 1);
        if (gi >= 2) PIXO_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ev_used[slot], 0));
        for (uint32_t k = 0; k < cnt; ++k)
            PIXO_TRY(h2d_copy(ctx, d_in + ((size_t)slot * G + k) * in_stride,
                              pixels + (size_t)(first + k) * len_each, len_each, ctx->copy_stream));
        PIXO_CUDA(ctx, cudaEventRecord(ev_in[slot], ctx->copy_stream));
        return 0;
    };

    // queue the kernels of group gi and the readback of its lengths
    auto compute = [&](uint32_t gi) -> int {
        const uint32_t first = gi * G, cnt = std::min(G, n_images - first);
        const int slot = (int)(gi & 1);
        const Coef c = coef_of(slot);
        PIXO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_in[slot], 0));
        PIXO_TRY(launch_jpeg_transform(ctx, d_in + (size_t)slot * G * in_stride, in_stride, cnt, g.width,
                                       g.height, g.color_type, g.subsampling, lum, chr, c.y, cstride,
                                       g.has_chroma ? c.cb : nullptr, g.has_chroma ? c.cr : nullptr, cstride, 0));
        PIXO_CUDA(ctx, cudaEventRecord(ev_used[slot], ctx->stream));
        std::vector<HuffTables> &tb = tables[slot];
        tb.resize(optimize ? cnt : 1);
        if (optimize) {
            auto *d_hist = reinterpret_cast<uint64_t *>(ctx->d_misc.ptr);
            PIXO_TRY(launch_jpeg_histogram(ctx, c.y, cstride, c.cb, c.cr, cstride, cnt, g.ny, g.nc, g.y_per_mcu,
                                           restart_interval, false, d_hist));
            PIXO_CUDA(ctx, cudaMemcpyAsync(h_hist, d_hist, (size_t)cnt * kHistWords * sizeof(uint64_t),
                                           cudaMemcpyDeviceToHost, ctx->stream));
            PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the table build needs the statistics
            for (uint32_t k = 0; k < cnt; ++k)  // unwrap_or_default, src/jpeg/mod.rs:379-392
                if (!huff_from_histogram(h_hist + (size_t)k * kHistWords, g.has_chroma, tb[k])) huff_standard(tb[k]);
        } else {
            huff_standard(tb[0]);
        }
        uint8_t *scan = d_scan + (size_t)slot * G * scan_cap;
        if (gi >= 2) PIXO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_out[slot], 0));  // slot's previous D2H drained
        uint64_t *d_len = nullptr;
        uint32_t *d_ovf = nullptr;
        auto *ent = reinterpret_cast<uint8_t *>(ctx->d_ent.ptr);
        uint64_t *h_len = h_len_of(slot);
        uint32_t *h_ovf = h_ovf_of(slot);
        if (!optimize) {
            PIXO_TRY(launch_jpeg_entropy(ctx, c.y, cstride, c.cb, c.cr, cstride, cnt, g, tb[0], restart_interval, ent,
                                         scan, scan_cap, &d_len, &d_ovf));
            PIXO_CUDA(ctx, cudaMemcpyAsync(h_len, d_len, (size_t)cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
            PIXO_CUDA(ctx, cudaMemcpyAsync(h_ovf, d_ovf, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            for (uint32_t k = 0; k < cnt; ++k) {  // per-image tables: one pass per image, each in its own scratch
                PIXO_TRY(launch_jpeg_entropy(ctx, c.y + (size_t)k * cstride, cstride, c.cb + (size_t)k * cstride,
                                             c.cr + (size_t)k * cstride, cstride, 1, g, tb[k],
                                             restart_interval, ent + (size_t)k * ent_one, scan + (size_t)k * scan_cap,
                                             scan_cap, &d_len, &d_ovf));
                PIXO_CUDA(ctx, cudaMemcpyAsync(h_len + k, d_len, 8, cudaMemcpyDeviceToHost, ctx->stream));
                PIXO_CUDA(ctx, cudaMemcpyAsync(h_ovf + k, d_ovf, 4, cudaMemcpyDeviceToHost, ctx->stream));
            }
        }
        PIXO_CUDA(ctx, cudaEventRecord(ev_len[slot], ctx->stream));
        return 0;
    };

    // headers on the host, scan bytes straight from the device (d2h stream), EOI
    auto finish = [&](uint32_t gi) -> int {
        const uint32_t first = gi * G, cnt = std::min(G, n_images - first);
        const int slot = (int)(gi & 1);
        const Coef c = coef_of(slot);
        const std::vector<HuffTables> &tb = tables[slot];
        uint8_t *scan = d_scan + (size_t)slot * G * scan_cap;
        const uint64_t *h_len = h_len_of(slot);
        const uint32_t *h_ovf = h_ovf_of(slot);
        const double tf0 = dbg ? now() : 0;
        PIXO_CUDA(ctx, cudaEventSynchronize(ev_len[slot]));
        if (dbg) fprintf(stderr, "  finish: waited %.0f us for the group's kernels\n", now() - tf0);
        std::vector<size_t> hdr(cnt);
        bool redo = false;
        for (uint32_t k = 0; k < cnt; ++k) {
            const uint32_t img = first + k;
            uint8_t *o = out + (size_t)img * out_cap_each;
            hdr[k] = write_headers(o, g, lum_zz, chr_zz, tb[optimize ? k : 0], restart_interval);
            if (h_ovf[k]) { redo = true; continue; }
            const size_t body = (size_t)h_len[k];
            if (hdr[k] + body + 2 > out_cap_each)
                return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "output capacity %zu too small (need %zu)",
                                 out_cap_each, hdr[k] + body + 2);
            if (out_locked)
                PIXO_CUDA(ctx, cudaMemcpyAsync(o + hdr[k], scan + (size_t)k * scan_cap, body, cudaMemcpyDeviceToHost,
                                               ctx->d2h_stream));
            else   // ordinary caller memory: through the pinned ring, copied out by the host pool
                PIXO_TRY(d2h_copy_sync(ctx, o + hdr[k], scan + (size_t)k * scan_cap, body, ctx->d2h_stream));
            o[hdr[k] + body] = 0xFF;
            o[hdr[k] + body + 1] = 0xD9;
            out_lens[img] = hdr[k] + body + 2;
        }
        PIXO_CUDA(ctx, cudaEventRecord(ev_out[slot], ctx->d2h_stream));
        if (!redo) return 0;
        // Frames the first pass did not finish.  Their coefficients are still in this slot of
        // d_coef (the next group's transform writes the other one).
        for (uint32_t k = 0; k < cnt; ++k) {
            if (!h_ovf[k]) continue;
            const uint32_t img = first + k;
            uint8_t *o = out + (size_t)img * out_cap_each;
            const HuffTables &t = tb[optimize ? k : 0];
            bool done = false;
            // bit 0: the scan did not fit (the kernel reported the size it needs); bit 2: a segment's raw
            // string did not fit its share - either way code the frame again on the GPU, unsegmented, with
            // enough room.  Bit 1 (a faulted chain) goes to the host coder.
            if (!(h_ovf[k] & 2u) && ctx->gpu_retry) {
                size_t need = (h_ovf[k] & 4u) ? (size_t)scan_cap * 2 : (size_t)h_len[k];
                for (int attempt = 0; attempt < 3 && !done; ++attempt) {
                    if (hdr[k] + need + 2 > out_cap_each && !(h_ovf[k] & 4u))
                        return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "output capacity %zu too small (need %zu)",
                                         out_cap_each, hdr[k] + need + 2);
                    const uint64_t cap2 = align_up(need + 64, 256);
                    PIXO_TRY(ensure_dev(ctx, ctx->d_retry, cap2 + ent_one + 256));
                    auto *rbuf = reinterpret_cast<uint8_t *>(ctx->d_retry.ptr);
                    uint64_t *d_len = nullptr;
                    uint32_t *d_ovf = nullptr;
                    ctx->no_segments = true;
                    const int rc = launch_jpeg_entropy(ctx, c.y + (size_t)k * cstride, cstride, c.cb + (size_t)k * cstride,
                                                       c.cr + (size_t)k * cstride, cstride, 1, g, t, restart_interval,
                                                       rbuf + cap2, rbuf, cap2, &d_len, &d_ovf);
                    ctx->no_segments = false;
                    PIXO_TRY(rc);
                    uint64_t len2 = 0;
                    uint32_t ovf2 = 0;
                    PIXO_CUDA(ctx, cudaMemcpyAsync(&len2, d_len, 8, cudaMemcpyDeviceToHost, ctx->stream));
                    PIXO_CUDA(ctx, cudaMemcpyAsync(&ovf2, d_ovf, 4, cudaMemcpyDeviceToHost, ctx->stream));
                    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                    if (ovf2 & 2u) break;
                    if (ovf2) { need = (size_t)len2; continue; }
                    if (hdr[k] + (size_t)len2 + 2 > out_cap_each)
                        return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "output capacity %zu too small (need %zu)",
                                         out_cap_each, hdr[k] + (size_t)len2 + 2);
                    PIXO_CUDA(ctx, cudaMemcpyAsync(o + hdr[k], rbuf, (size_t)len2, cudaMemcpyDeviceToHost, ctx->stream));
                    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                    o[hdr[k] + len2] = 0xFF;
                    o[hdr[k] + len2 + 1] = 0xD9;
                    out_lens[img] = hdr[k] + (size_t)len2 + 2;
                    done = true;
                }
            }
            if (done) continue;
            // last resort: the host entropy coder on the GPU's coefficient arrays
            ctx->host_fallbacks += 1;
            PIXO_TRY(ensure_pinned(ctx, ctx->h_out, coef_each));
            auto *hc = reinterpret_cast<uint8_t *>(ctx->h_out.ptr);
            PIXO_CUDA(ctx, cudaMemcpyAsync(hc, reinterpret_cast<const uint8_t *>(c.y) + (size_t)k * coef_each, coef_each,
                                           cudaMemcpyDeviceToHost, ctx->stream));
            PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            const size_t body = entropy_encode_scan(reinterpret_cast<int16_t *>(hc), reinterpret_cast<int16_t *>(hc + yb),
                                                    reinterpret_cast<int16_t *>(hc + yb + cbb), g, t, restart_interval,
                                                    false, o + hdr[k], out_cap_each - hdr[k] - 2, ctx->host_threads);
            if (body == (size_t)-1)
                return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "output capacity %zu too small", out_cap_each);
            o[hdr[k] + body] = 0xFF;
            o[hdr[k] + body + 1] = 0xD9;
            out_lens[img] = hdr[k] + body + 2;
        }
        return 0;
    };

    // the copy streams start after whatever the caller already queued on the main stream
    PIXO_CUDA(ctx, cudaEventRecord(ev_out[0], ctx->stream));
    PIXO_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ev_out[0], 0));
    PIXO_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, ev_out[0], 0));
    const double t0 = dbg ? now() : 0;
    PIXO_TRY(upload(0));
    const double t1 = dbg ? now() : 0;
    for (uint32_t gi = 0; gi < ngroups; ++gi) {
        if (gi + 1 < ngroups) PIXO_TRY(upload(gi + 1));
        PIXO_TRY(compute(gi));
        if (gi > 0) PIXO_TRY(finish(gi - 1));
    }
    const double t2 = dbg ? now() : 0;
    PIXO_TRY(finish(ngroups - 1));
    const double t3 = dbg ? now() : 0;
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (dbg)
        fprintf(stderr, "encode_frames n=%u: upload(0) queued %.0f us, compute queued %.0f us, last finish %.0f us, drain %.0f us\n",
                n_images, t1 - t0, t2 - t1, t3 - t2, now() - t3);
    drain.armed = false;
    return 0;
}

int pixo_b200_jpeg_encode(pixo_b200_ctx *ctx, const uint8_t *pixels, size_t pixels_len,
                          uint32_t width, uint32_t height, uint32_t color_type, uint32_t quality,
                          uint32_t subsampling, uint32_t restart_interval,
                          uint32_t optimize_huffman, uint32_t progressive, uint32_t trellis_quant,
                          uint8_t *out, size_t out_cap, size_t *out_len)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    PIXO_TRY(validate_encode(ctx, pixels_len, width, height, color_type, quality, subsampling,
                             restart_interval, false));
    if (!pixels || !out || !out_len) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    if (progressive)
        return set_error(ctx, PIXO_B200_ERR_UNSUPPORTED,
                         "progressive scans are outside the accelerated path (sequential entropy stage)");
    (void)trellis_quant;  // baseline encode_scan ignores use_trellis (src/jpeg/mod.rs:1408-1563)
    const FrameGeometry g = make_geometry(width, height, color_type, subsampling);
    return encode_frames(ctx, pixels, pixels_len, 1, g, quality, restart_interval, optimize_huffman != 0, out,
                         out_cap, out_len);
}

int pixo_b200_jpeg_encode_batch(pixo_b200_ctx *ctx, const uint8_t *pixels, size_t pixels_len_each,
                                uint32_t n_images, uint32_t width, uint32_t height,
                                uint32_t color_type, uint32_t quality, uint32_t subsampling,
                                uint32_t restart_interval, uint32_t optimize_huffman,
                                uint8_t *out, size_t out_cap_each, size_t *out_lens)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    PIXO_TRY(validate_encode(ctx, pixels_len_each, width, height, color_type, quality,
                             subsampling, restart_interval, false));
    if (!pixels || !out || !out_lens) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    if (n_images == 0) return 0;
    const FrameGeometry g = make_geometry(width, height, color_type, subsampling);
    return encode_frames(ctx, pixels, pixels_len_each, n_images, g, quality, restart_interval,
                         optimize_huffman != 0, out, out_cap_each, out_lens);
}

int pixo_b200_jpeg_encode_dev(pixo_b200_ctx *ctx, const uint8_t *d_pixels, size_t pixel_stride,
                              uint32_t n_images, uint32_t width, uint32_t height,
                              uint32_t color_type, uint32_t quality, uint32_t subsampling,
                              uint8_t *d_scan, size_t scan_cap_each, uint64_t *d_scan_len,
                              uint32_t *d_overflow)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    if (quality == 0 || quality > 100)
        return set_error(ctx, PIXO_B200_ERR_INVALID_QUALITY, "Invalid quality %u: must be 1-100", quality);
    PIXO_TRY(validate_jpeg(ctx, width, height, color_type, subsampling));
    if (!d_pixels || !d_scan || !d_scan_len || !d_overflow)
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    if (n_images == 0) return 0;
    if (n_images > 65535) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "at most 65535 frames per call");
    if (scan_cap_each % 4 || (reinterpret_cast<uintptr_t>(d_scan) & 15))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "scan buffer must be 16-byte aligned, capacity multiple of 4");
    const FrameGeometry g = make_geometry(width, height, color_type, subsampling);
    float lum[64], chr[64];
    quant_tables((int)quality, nullptr, nullptr, lum, chr);
    const size_t yb = align_up(g.ny * 64 * sizeof(int16_t), 256);
    const size_t cbb = align_up(g.nc * 64 * sizeof(int16_t), 256);
    const size_t coef_each = yb + 2 * cbb;
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    PIXO_TRY(ensure_dev(ctx, ctx->d_coef, (size_t)n_images * coef_each));
    PIXO_TRY(ensure_dev(ctx, ctx->d_ent, entropy_scratch_bytes(n_images, g, 0)));
    auto *d_coef = reinterpret_cast<uint8_t *>(ctx->d_coef.ptr);
    auto *dy = reinterpret_cast<int16_t *>(d_coef);
    auto *dcb = reinterpret_cast<int16_t *>(d_coef + yb);
    auto *dcr = reinterpret_cast<int16_t *>(d_coef + yb + cbb);
    PIXO_TRY(launch_jpeg_transform(ctx, d_pixels, pixel_stride, n_images, width, height, color_type, subsampling,
                                   lum, chr, dy, coef_each / 2, g.has_chroma ? dcb : nullptr,
                                   g.has_chroma ? dcr : nullptr, coef_each / 2, 0));
    HuffTables t;
    huff_standard(t);
    uint64_t *len_src = nullptr;
    uint32_t *ovf_src = nullptr;
    PIXO_TRY(launch_jpeg_entropy(ctx, dy, coef_each / 2, dcb, dcr, coef_each / 2, n_images, g, t, 0,
                                 reinterpret_cast<uint8_t *>(ctx->d_ent.ptr), d_scan, scan_cap_each,
                                 &len_src, &ovf_src));
    PIXO_CUDA(ctx, cudaMemcpyAsync(d_scan_len, len_src, (size_t)n_images * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    PIXO_CUDA(ctx, cudaMemcpyAsync(d_overflow, ovf_src, (size_t)n_images * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

int pixo_b200_jpeg_entropy_encode(pixo_b200_ctx *ctx, const int16_t *y, const int16_t *cb,
                                  const int16_t *cr, uint32_t width, uint32_t height,
                                  uint32_t color_type, uint32_t quality, uint32_t subsampling,
                                  uint32_t restart_interval, uint32_t optimize_huffman,
                                  uint8_t *out, size_t out_cap, size_t *out_len)
{
    if (quality == 0 || quality > 100)
        return set_error(ctx, PIXO_B200_ERR_INVALID_QUALITY, "Invalid quality %u: must be 1-100", quality);
    if (restart_interval > 65535)
        return set_error(ctx, PIXO_B200_ERR_INVALID_RESTART, "Invalid restart interval %u", restart_interval);
    PIXO_TRY(validate_jpeg(ctx, width, height, color_type, subsampling));
    if (!y || !out || !out_len || (color_type != PIXO_B200_GRAY && (!cb || !cr)))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    const FrameGeometry g = make_geometry(width, height, color_type, subsampling);
    // baseline Huffman tables code DC differences of category <= 11 and AC values of category <= 10
    // (what an 8-bit forward DCT can produce); anything else has no code
    if (!coefficients_in_range(y, g.ny, restart_interval, g.y_per_mcu) ||
        (g.has_chroma && (!coefficients_in_range(cb, g.nc, restart_interval, 1) ||
                          !coefficients_in_range(cr, g.nc, restart_interval, 1))))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT,
                         "coefficient out of the baseline range (|AC| <= 1023, |DC difference| <= 2047)");
    uint64_t hist[536];
    if (optimize_huffman) host_histogram(y, cb, cr, g, restart_interval, hist);
    int threads = ctx ? ctx->host_threads : (int)std::thread::hardware_concurrency();
    return entropy_from_host_arrays(ctx, y, cb, cr, g, quality, restart_interval,
                                    optimize_huffman ? hist : nullptr, out, out_cap, out_len,
                                    threads < 1 ? 1 : threads);
}

int pixo_b200_jpeg_entropy_encode_dev(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb,
                                      const int16_t *d_cr, uint32_t width, uint32_t height,
                                      uint32_t color_type, uint32_t quality, uint32_t subsampling,
                                      uint32_t restart_interval, uint32_t optimize_huffman,
                                      uint8_t *out, size_t out_cap, size_t *out_len)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    if (quality == 0 || quality > 100)
        return set_error(ctx, PIXO_B200_ERR_INVALID_QUALITY, "Invalid quality %u: must be 1-100", quality);
    if (restart_interval > 65535)
        return set_error(ctx, PIXO_B200_ERR_INVALID_RESTART, "Invalid restart interval %u", restart_interval);
    PIXO_TRY(validate_jpeg(ctx, width, height, color_type, subsampling));
    if (!d_y || !out || !out_len || (color_type != PIXO_B200_GRAY && (!d_cb || !d_cr)))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    if (out_cap < 1024 + 2)
        return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "output capacity %zu too small", out_cap);
    const FrameGeometry g = make_geometry(width, height, color_type, subsampling);
    uint8_t lum_zz[64], chr_zz[64];
    quant_tables((int)quality, lum_zz, chr_zz, nullptr, nullptr);
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    PIXO_TRY(ensure_pinned(ctx, ctx->h_misc, kHistWords * sizeof(uint64_t) + 256));
    auto *h_meta = reinterpret_cast<uint8_t *>(ctx->h_misc.ptr);
    HuffTables t;
    huff_standard(t);
    if (optimize_huffman) {
        PIXO_TRY(ensure_dev(ctx, ctx->d_misc, kHistWords * sizeof(uint64_t) + 256));
        auto *d_hist = reinterpret_cast<uint64_t *>(ctx->d_misc.ptr);
        auto *h_hist = reinterpret_cast<uint64_t *>(h_meta + 256);
        PIXO_TRY(launch_jpeg_histogram(ctx, d_y, 0, d_cb, d_cr, 0, 1, g.ny, g.nc, g.y_per_mcu, restart_interval,
                                       false, d_hist));
        PIXO_CUDA(ctx, cudaMemcpyAsync(h_hist, d_hist, kHistWords * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (!huff_from_histogram(h_hist, g.has_chroma, t)) huff_standard(t);  // unwrap_or_default
    }
    const size_t hdr = write_headers(out, g, lum_zz, chr_zz, t, restart_interval);
    // The device scan buffer follows the size a JPEG of this geometry normally has, not the caller's
    // worst-case capacity (tens of GB for a gigapixel frame); a scan that needs more is coded again
    // with the exact size the kernel reported.
    const size_t raw = (size_t)width * height * (color_type == PIXO_B200_GRAY ? 1 : 3);
    size_t scan_cap = std::min<size_t>((out_cap - hdr - 2) & ~(size_t)15, (size_t)default_scan_cap(ctx, raw));
    PIXO_TRY(ensure_dev(ctx, ctx->d_ent, entropy_scratch_bytes(1, g, restart_interval)));
    auto *h_len = reinterpret_cast<uint64_t *>(h_meta);
    auto *h_ovf = reinterpret_cast<uint32_t *>(h_meta + 8);
    uint8_t *d_scan = nullptr;
    for (int attempt = 0;; ++attempt) {
        PIXO_TRY(ensure_dev(ctx, ctx->d_out, scan_cap + 16));
        uint64_t *d_len = nullptr;
        uint32_t *d_ovf = nullptr;
        d_scan = reinterpret_cast<uint8_t *>(ctx->d_out.ptr);
        PIXO_TRY(launch_jpeg_entropy(ctx, d_y, 0, d_cb, d_cr, 0, 1, g, t, restart_interval,
                                     reinterpret_cast<uint8_t *>(ctx->d_ent.ptr), d_scan, scan_cap, &d_len, &d_ovf));
        PIXO_CUDA(ctx, cudaMemcpyAsync(h_len, d_len, 8, cudaMemcpyDeviceToHost, ctx->stream));
        PIXO_CUDA(ctx, cudaMemcpyAsync(h_ovf, d_ovf, 4, cudaMemcpyDeviceToHost, ctx->stream));
        PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->no_segments = false;
        if (!*h_ovf) break;
        if ((*h_ovf & 2u) || attempt >= 2)
            return set_error(ctx, PIXO_B200_ERR_CUDA, "device entropy stage did not finish (flags %u)", *h_ovf);
        if (*h_ovf & 4u) {   // a segment outgrew its share: once more, unsegmented
            ctx->no_segments = true;
            continue;
        }
        if (hdr + (size_t)*h_len + 2 > out_cap)
            return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "output capacity %zu too small (need %zu)", out_cap,
                             hdr + (size_t)*h_len + 2);
        scan_cap = align_up((size_t)*h_len + 64, 256);
    }
    const size_t body = (size_t)*h_len;
    PIXO_CUDA(ctx, cudaMemcpyAsync(out + hdr, d_scan, body, cudaMemcpyDeviceToHost, ctx->stream));
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out[hdr + body] = 0xFF;
    out[hdr + body + 1] = 0xD9;
    *out_len = hdr + body + 2;
    return 0;
}

// ---- one frame tiled over several GPUs -----------------------------------------------------------

static void tables_from(const uint64_t *hist, bool has_chroma, HuffTables &t)
{
    // build_optimized_huffman_tables(..).unwrap_or_default(), src/jpeg/mod.rs:379-392
    if (!(hist && huff_from_histogram(hist, has_chroma, t))) huff_standard(t);
}

int pixo_b200_jpeg_band_last_dc(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb,
                                const int16_t *d_cr, size_t ny, size_t nc, int32_t last_dc[3])
{
    if (!ctx || !last_dc) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    int16_t v[3] = {0, 0, 0};
    if (ny && d_y) PIXO_CUDA(ctx, cudaMemcpyAsync(&v[0], d_y + (ny - 1) * 64, 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (nc && d_cb) PIXO_CUDA(ctx, cudaMemcpyAsync(&v[1], d_cb + (nc - 1) * 64, 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (nc && d_cr) PIXO_CUDA(ctx, cudaMemcpyAsync(&v[2], d_cr + (nc - 1) * 64, 2, cudaMemcpyDeviceToHost, ctx->stream));
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 3; ++k) last_dc[k] = v[k];
    return 0;
}

int pixo_b200_jpeg_band_histogram_dev(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb,
                                      const int16_t *d_cr, uint32_t width, uint32_t band_height,
                                      uint32_t color_type, uint32_t subsampling,
                                      const int32_t dc_seed[3], uint64_t *d_hist)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    PIXO_TRY(validate_jpeg(ctx, width, band_height, color_type, subsampling));
    if (!d_y || !d_hist || (color_type != PIXO_B200_GRAY && (!d_cb || !d_cr)))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    const FrameGeometry g = make_geometry(width, band_height, color_type, subsampling);
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_jpeg_histogram(ctx, d_y, 0, d_cb, d_cr, 0, 1, g.ny, g.nc, g.y_per_mcu, 0, false, d_hist, dc_seed);
}

int pixo_b200_jpeg_band_entropy_dev(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb,
                                    const int16_t *d_cr, uint32_t width, uint32_t band_height,
                                    uint32_t color_type, uint32_t subsampling,
                                    const int32_t dc_seed[3], const uint64_t *hist, uint8_t *d_raw,
                                    size_t raw_cap, uint64_t *nbits, uint32_t *tail7)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    PIXO_TRY(validate_jpeg(ctx, width, band_height, color_type, subsampling));
    if (!d_y || !d_raw || !nbits || !tail7 || (color_type != PIXO_B200_GRAY && (!d_cb || !d_cr)))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    if ((raw_cap & 3) || (reinterpret_cast<uintptr_t>(d_raw) & 15))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "raw buffer must be 16-byte aligned, capacity multiple of 4");
    const FrameGeometry g = make_geometry(width, band_height, color_type, subsampling);
    HuffTables t;
    tables_from(hist, g.has_chroma, t);
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    PIXO_TRY(ensure_dev(ctx, ctx->d_ent, entropy_scratch_bytes(1, g, 0)));
    uint64_t *d_len = nullptr, *d_tail = nullptr;
    uint32_t *d_ovf = nullptr;
    PIXO_TRY(launch_jpeg_entropy(ctx, d_y, 0, d_cb, d_cr, 0, 1, g, t, 0, reinterpret_cast<uint8_t *>(ctx->d_ent.ptr),
                                 d_raw, raw_cap, &d_len, &d_ovf, dc_seed, &d_tail));
    // (a long band is coded as several segments - see k_seg_* -: their bit counts add up, the band's
    // tail is its last non-empty segment's)
    const uint32_t S = ctx->last_band_segments;
    PIXO_TRY(ensure_pinned(ctx, ctx->h_misc, (size_t)S * 20 + 256));
    auto *h = reinterpret_cast<uint64_t *>(ctx->h_misc.ptr);
    PIXO_CUDA(ctx, cudaMemcpyAsync(h, d_len, (size_t)S * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PIXO_CUDA(ctx, cudaMemcpyAsync(h + S, d_tail, (size_t)S * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PIXO_CUDA(ctx, cudaMemcpyAsync(h + 2 * S, d_ovf, (size_t)S * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t ovf = 0;
    uint64_t total = 0, tail = 0;
    for (uint32_t q = 0; q < S; ++q) {
        ovf |= reinterpret_cast<uint32_t *>(h + 2 * S)[q];
        total += h[q];
        if (h[q]) tail = h[S + q];
    }
    h[0] = total;
    *nbits = total;
    *tail7 = (uint32_t)tail & 0x7Fu;
    if (ovf & 2) return set_error(ctx, PIXO_B200_ERR_CUDA, "device entropy stage did not finish (flags %u)", ovf);
    if (ovf) return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "raw capacity %zu too small (need %llu)", raw_cap,
                              (unsigned long long)((h[0] + 7) / 8));
    return 0;
}

int pixo_b200_jpeg_band_splice_dev(pixo_b200_ctx *ctx, const uint8_t *d_raw, uint64_t nbits,
                                   uint64_t start_bit, uint32_t tail_in, uint32_t is_last_band,
                                   uint8_t *d_out, size_t out_cap, uint64_t *out_len)
{
    if (!ctx || !d_out || !out_len || (!d_raw && nbits))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    PIXO_TRY(ensure_dev(ctx, ctx->d_misc, splice_scratch_bytes(nbits)));
    PIXO_TRY(ensure_pinned(ctx, ctx->h_misc, 256));
    uint64_t *d_len = nullptr;
    uint32_t *d_ovf = nullptr;
    if (ctx->bands.count(d_raw))   // the band was coded in segments (strings, bit counts and tails are in d_raw)
        PIXO_TRY(launch_band_splice_segments(ctx, d_raw, start_bit, tail_in, is_last_band != 0,
                                             reinterpret_cast<uint8_t *>(ctx->d_misc.ptr), d_out, out_cap, &d_len, &d_ovf));
    else
        PIXO_TRY(launch_splice(ctx, d_raw, nbits, (uint32_t)(start_bit & 7), tail_in, is_last_band != 0,
                               reinterpret_cast<uint8_t *>(ctx->d_misc.ptr), d_out, out_cap, &d_len, &d_ovf));
    auto *h = reinterpret_cast<uint64_t *>(ctx->h_misc.ptr);
    PIXO_CUDA(ctx, cudaMemcpyAsync(h, d_len, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PIXO_CUDA(ctx, cudaMemcpyAsync(h + 1, d_ovf, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*reinterpret_cast<uint32_t *>(h + 1))
        return set_error(ctx, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "splice output capacity %zu too small", out_cap);
    *out_len = h[0];
    return 0;
}

int pixo_b200_jpeg_band_entropy_dev_async(pixo_b200_ctx *ctx, const int16_t *d_y, const int16_t *d_cb,
                                          const int16_t *d_cr, uint32_t width, uint32_t band_height,
                                          uint32_t color_type, uint32_t subsampling,
                                          const int32_t *d_dc_seed, const uint64_t *hist, uint8_t *d_raw,
                                          size_t raw_cap, uint64_t *d_bits_tail, uint32_t *d_flags)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    PIXO_TRY(validate_jpeg(ctx, width, band_height, color_type, subsampling));
    if (!d_y || !d_raw || !d_dc_seed || !d_bits_tail || !d_flags || (color_type != PIXO_B200_GRAY && (!d_cb || !d_cr)))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    if ((raw_cap & 3) || (reinterpret_cast<uintptr_t>(d_raw) & 15))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "raw buffer must be 16-byte aligned, capacity multiple of 4");
    const FrameGeometry g = make_geometry(width, band_height, color_type, subsampling);
    HuffTables t;
    tables_from(hist, g.has_chroma, t);
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_band_entropy_async(ctx, d_y, d_cb, d_cr, g, t, d_dc_seed, d_raw, raw_cap, d_bits_tail, d_flags);
}

int pixo_b200_jpeg_band_splice_dev_async(pixo_b200_ctx *ctx, const uint8_t *d_raw, const uint64_t *d_offset,
                                         uint8_t *d_out, size_t out_cap, uint64_t *d_out_len,
                                         uint32_t *d_flags)
{
    if (!ctx || !d_raw || !d_offset || !d_out || !d_out_len || !d_flags)
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_band_splice_async(ctx, d_raw, d_offset, d_out, out_cap, d_out_len, d_flags);
}

int pixo_b200_jpeg_band_entropy(const int16_t *y, const int16_t *cb, const int16_t *cr, uint32_t width,
                                uint32_t band_height, uint32_t color_type, uint32_t subsampling,
                                const int32_t dc_seed[3], const uint64_t *hist, uint8_t *raw,
                                size_t raw_cap, uint64_t *nbits, uint32_t *tail7)
{
    PIXO_TRY(validate_jpeg(nullptr, width, band_height, color_type, subsampling));
    if (!y || !raw || !nbits || !tail7 || (color_type != PIXO_B200_GRAY && (!cb || !cr)))
        return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    const FrameGeometry g = make_geometry(width, band_height, color_type, subsampling);
    HuffTables t;
    tables_from(hist, g.has_chroma, t);
    const uint64_t n = band_encode_raw(y, cb, cr, g, t, dc_seed, raw, raw_cap, tail7);
    if (n == (uint64_t)-1) return set_error(nullptr, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "raw capacity %zu too small", raw_cap);
    *nbits = n;
    return 0;
}

int pixo_b200_jpeg_band_histogram(const int16_t *y, const int16_t *cb, const int16_t *cr, uint32_t width,
                                  uint32_t band_height, uint32_t color_type, uint32_t subsampling,
                                  const int32_t dc_seed[3], uint64_t hist[536])
{
    PIXO_TRY(validate_jpeg(nullptr, width, band_height, color_type, subsampling));
    if (!y || !hist || (color_type != PIXO_B200_GRAY && (!cb || !cr)))
        return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    const FrameGeometry g = make_geometry(width, band_height, color_type, subsampling);
    host_histogram(y, cb, cr, g, 0, hist, dc_seed);
    return 0;
}

int pixo_b200_jpeg_band_splice(const uint8_t *raw, uint64_t nbits, uint64_t start_bit, uint32_t tail_in,
                               uint32_t is_last_band, uint8_t *out, size_t out_cap, size_t *out_len)
{
    if (!out || !out_len || (!raw && nbits)) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    const size_t n = band_splice(raw, nbits, (uint32_t)(start_bit & 7), tail_in, is_last_band != 0, out, out_cap);
    if (n == (size_t)-1) return set_error(nullptr, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "splice output capacity %zu too small", out_cap);
    *out_len = n;
    return 0;
}

int pixo_b200_jpeg_write_headers(uint32_t width, uint32_t height, uint32_t color_type, uint32_t quality,
                                 uint32_t subsampling, uint32_t restart_interval, const uint64_t *hist,
                                 uint8_t *out, size_t out_cap, size_t *out_len)
{
    if (quality == 0 || quality > 100)
        return set_error(nullptr, PIXO_B200_ERR_INVALID_QUALITY, "Invalid quality %u: must be 1-100", quality);
    if (restart_interval > 65535)
        return set_error(nullptr, PIXO_B200_ERR_INVALID_RESTART, "Invalid restart interval %u", restart_interval);
    PIXO_TRY(validate_jpeg(nullptr, width, height, color_type, subsampling));
    if (!out || !out_len) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    if (out_cap < 1024) return set_error(nullptr, PIXO_B200_ERR_OUTPUT_TOO_SMALL, "output capacity %zu too small", out_cap);
    const FrameGeometry g = make_geometry(width, height, color_type, subsampling);
    uint8_t lum_zz[64], chr_zz[64];
    quant_tables((int)quality, lum_zz, chr_zz, nullptr, nullptr);
    HuffTables t;
    tables_from(hist, g.has_chroma, t);
    *out_len = write_headers(out, g, lum_zz, chr_zz, t, restart_interval);
    return 0;
}

// ---- PNG ----------------------------------------------------------------------------------

static int validate_png(pixo_b200_ctx *ctx, uint32_t width, uint32_t height, size_t row_bytes,
                        uint32_t bpp, uint32_t strategy)
{
    if (width == 0 || height == 0 || row_bytes == 0)
        return set_error(ctx, PIXO_B200_ERR_INVALID_DIMENSIONS, "Invalid image dimensions: %ux%u", width, height);
    if (width > (1u << 24) || height > (1u << 24))  // src/png/mod.rs:21
        return set_error(ctx, PIXO_B200_ERR_IMAGE_TOO_LARGE, "Image dimensions %ux%u exceed maximum %u", width, height, 1u << 24);
    if (bpp < 1 || bpp > 4)
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "bytes_per_pixel %u not in 1..4", bpp);
    if ((strategy & ~PIXO_B200_PNG_OPTIMIZE_ALPHA) > PIXO_B200_FILTER_BIGRAMS)
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "unknown filter strategy %u", strategy);
    return 0;
}

int pixo_b200_png_filter_dev(pixo_b200_ctx *ctx, const uint8_t *d_data, size_t in_stride,
                             uint32_t n_images, uint32_t width, uint32_t height,
                             size_t row_bytes, uint32_t bytes_per_pixel, uint32_t strategy,
                             uint8_t *d_out, size_t out_stride, uint32_t *d_adler)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    PIXO_TRY(validate_png(ctx, width, height, row_bytes, bytes_per_pixel, strategy));
    if (!d_data || !d_out) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    if (n_images == 0) return 0;
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_png_filter(ctx, d_data, in_stride, n_images, width, height, row_bytes,
                             bytes_per_pixel, strategy, d_out, out_stride, d_adler);
}

int pixo_b200_png_filter(pixo_b200_ctx *ctx, const uint8_t *data, uint32_t width,
                         uint32_t height, size_t row_bytes, uint32_t bytes_per_pixel,
                         uint32_t strategy, uint8_t *out, uint32_t *adler32_out)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    PIXO_TRY(validate_png(ctx, width, height, row_bytes, bytes_per_pixel, strategy));
    if (!data || !out) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    const size_t in_bytes = row_bytes * height, out_bytes = (row_bytes + 1) * (size_t)height;
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    PIXO_TRY(ensure_dev(ctx, ctx->d_in, in_bytes));
    PIXO_TRY(ensure_dev(ctx, ctx->d_out, out_bytes + 16));
    PIXO_TRY(ensure_dev(ctx, ctx->d_y, 64));
    PIXO_TRY(h2d_copy(ctx, ctx->d_in.ptr, data, in_bytes, ctx->stream));
    uint32_t *d_adler = adler32_out ? reinterpret_cast<uint32_t *>(ctx->d_y.ptr) : nullptr;
    PIXO_TRY(launch_png_filter(ctx, reinterpret_cast<const uint8_t *>(ctx->d_in.ptr), in_bytes, 1,
                               width, height, row_bytes, bytes_per_pixel, strategy,
                               reinterpret_cast<uint8_t *>(ctx->d_out.ptr), out_bytes, d_adler));
    PIXO_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (adler32_out)
        PIXO_CUDA(ctx, cudaMemcpyAsync(adler32_out, d_adler, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int pixo_b200_png_filter_rows_dev(pixo_b200_ctx *ctx, const uint8_t *d_rows, const uint8_t *d_row_above,
                                  uint32_t width, uint32_t image_height, uint32_t band_rows,
                                  size_t row_bytes, uint32_t bytes_per_pixel, uint32_t strategy,
                                  uint8_t *d_out, uint32_t *d_adler)
{
    if (!ctx) return set_error(nullptr, PIXO_B200_ERR_INVALID_ARGUMENT, "ctx is null");
    PIXO_TRY(validate_png(ctx, width, image_height, row_bytes, bytes_per_pixel, strategy));
    if (!d_rows || !d_out) return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null buffer");
    if (band_rows == 0 || band_rows > image_height)
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "band_rows %u outside 1..%u", band_rows, image_height);
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_png_filter_rows(ctx, d_rows, row_bytes * band_rows, 1, width, band_rows, row_bytes, bytes_per_pixel,
                                  strategy, d_out, (row_bytes + 1) * (size_t)band_rows, d_adler, d_row_above,
                                  image_height);
}

uint32_t pixo_b200_adler32_combine(uint32_t adler_a, uint32_t adler_b, uint64_t len_b)
{
    const uint64_t M = 65521;
    const uint64_t a1 = adler_a & 0xFFFF, a2 = adler_a >> 16, b1 = adler_b & 0xFFFF, b2 = adler_b >> 16;
    const uint64_t s1 = (a1 + b1 + M - 1) % M;
    const uint64_t s2 = (a2 + b2 + (len_b % M) * ((a1 + M - 1) % M)) % M;
    return (uint32_t)((s2 << 16) | s1);
}

int pixo_b200_adler32_dev(pixo_b200_ctx *ctx, const uint8_t *d_data, size_t len, uint32_t *d_out)
{
    if (!ctx || !d_out || (!d_data && len))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_adler32(ctx, d_data, len, d_out);
}

int pixo_b200_adler32(pixo_b200_ctx *ctx, const uint8_t *data, size_t len, uint32_t *out)
{
    if (!ctx || !out || (!data && len))
        return set_error(ctx, PIXO_B200_ERR_INVALID_ARGUMENT, "null argument");
    PIXO_CUDA(ctx, cudaSetDevice(ctx->device));
    PIXO_TRY(ensure_dev(ctx, ctx->d_in, len + 16));
    PIXO_TRY(ensure_dev(ctx, ctx->d_y, 64));
    if (len)
        PIXO_TRY(h2d_copy(ctx, ctx->d_in.ptr, data, len, ctx->stream));
    PIXO_TRY(launch_adler32(ctx, reinterpret_cast<const uint8_t *>(ctx->d_in.ptr), len,
                            reinterpret_cast<uint32_t *>(ctx->d_y.ptr)));
    PIXO_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_y.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PIXO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"
