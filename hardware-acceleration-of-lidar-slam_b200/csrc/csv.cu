// csv.cu -- scan ingest: the reference's CSV reader on the device.
//
// Replaces readDatasetLineByLine (Subsystem_1/main.c:22-30): column x fscanf(filename, "%f,", &value).
// glibc's %f conversion is a correctly rounded decimal -> binary32 conversion (strtof), values are
// separated by ',' and / or white space.  The whole file goes to the GPU once (pinned staging, one
// H2D copy of the raw bytes) and three kernels turn it into floats:
//   csv_count_kernel   token starts per 4 KiB chunk (a token starts where a non-separator byte follows a
//                      separator or the beginning of the text)
//   csv_offsets_kernel exclusive scan of the chunk counts (one CTA) -> every token's index
//   csv_parse_kernel   one thread per token start: sign, digits, optional fraction -> integer mantissa M and
//                      decimal exponent -k, value = M / 10^k in IEEE double (M < 2^53 and 10^k, k <= 22, are
//                      exact, so the quotient is the correctly rounded double of the decimal), then rounded to
//                      float.  Double rounding can only differ from the correctly rounded float when the
//                      double sits within one ulp of the midpoint of two floats; those tokens (3 in 2^29),
//                      and everything that is not a plain decimal (exponents, inf / nan, hex floats, more
//                      than 15 significant digits, values outside the normal float range) are not converted
//                      here but listed for the host, which calls strtof -- the function fscanf itself uses --
//                      on exactly those tokens.  Every value is therefore bit-identical to the reference's.
// The floats stay on the device: b200slam_scan_read_resident feeds readAScan (frontend.cu) from there, so a
// replay uploads its dataset once instead of 4 bytes per beam per scan.
#include <errno.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int CSV_THREADS = 256;
constexpr int CSV_BYTES_PER_THREAD = 16;
constexpr int CSV_CHUNK = CSV_THREADS * CSV_BYTES_PER_THREAD;      // 4 KiB of text per CTA
constexpr int CSV_MAX_FALLBACK = 1 << 16;
constexpr int CSV_MAX_TOKEN = 64;                                   // bytes a plain decimal token may have

__device__ __forceinline__ bool csv_is_sep(unsigned char c)
{
    // ',' and the C isspace set: what fscanf("%f,") skips before and matches after a value
    return c == ',' || c == ' ' || (c >= '\t' && c <= '\r');
}

// 16 start flags of the 16 bytes at `base` (bit b: a token starts at base + b)
__device__ __forceinline__ unsigned csv_start_flags(const unsigned char *__restrict__ text, size_t n, size_t base)
{
    if (base >= n) return 0u;
    unsigned char prev = base == 0 ? (unsigned char)',' : text[base - 1];
    unsigned flags = 0;
    if (base + 16 <= n) {
        const uint4 q = *reinterpret_cast<const uint4 *>(text + base);      // text is 16-byte aligned, base a multiple of 16
        const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const unsigned char c = (unsigned char)(w[b >> 2] >> (8 * (b & 3)));
            flags |= (unsigned)(!csv_is_sep(c) && csv_is_sep(prev)) << b;
            prev = c;
        }
    } else {
        for (int b = 0; base + b < n; ++b) {
            const unsigned char c = text[base + b];
            flags |= (unsigned)(!csv_is_sep(c) && csv_is_sep(prev)) << b;
            prev = c;
        }
    }
    return flags;
}

__global__ void __launch_bounds__(CSV_THREADS)
csv_count_kernel(const unsigned char *__restrict__ text, size_t n, unsigned long long *__restrict__ chunk_counts)
{
    __shared__ int wsum[CSV_THREADS / 32];
    const size_t base = ((size_t)blockIdx.x * CSV_THREADS + threadIdx.x) * CSV_BYTES_PER_THREAD;
    int c = __popc(csv_start_flags(text, n, base));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < CSV_THREADS / 32; ++w) t += wsum[w];
        chunk_counts[blockIdx.x] = (unsigned long long)t;
    }
}

// exclusive scan of the chunk counts, one CTA; total -> out[0], fallback counter out[1] <- 0
__global__ void __launch_bounds__(1024)
csv_offsets_kernel(unsigned long long *__restrict__ counts, int nb, unsigned long long *__restrict__ out)
{
    __shared__ unsigned long long wtot[32];
    __shared__ unsigned long long carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const unsigned long long v = i < nb ? counts[i] : 0;
        unsigned long long inc = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, s);
            if (lane >= s) inc += o;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = wtot[lane];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const unsigned long long o = __shfl_up_sync(0xffffffffu, t, s);
                if (lane >= s) t += o;
            }
            wtot[lane] = t;
        }
        __syncthreads();
        const unsigned long long before = carry + (warp ? wtot[warp - 1] : 0);
        if (i < nb) counts[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = carry; out[1] = 0; }
}

struct CsvFallback { unsigned long long index, offset; };

// One plain decimal token -> float.  false: leave it to the host's strtof.
__device__ __forceinline__ bool csv_convert(const unsigned char *__restrict__ text, size_t n, size_t pos, float *out)
{
    size_t p = pos;
    bool neg = false;
    if (text[p] == '-' || text[p] == '+') { neg = text[p] == '-'; ++p; }
    unsigned long long M = 0;
    int digits = 0, sig = 0, k = 0;
    bool frac = false;
    for (; p < n && p - pos < CSV_MAX_TOKEN; ++p) {
        const unsigned char c = text[p];
        if (c >= '0' && c <= '9') {
            ++digits;
            if (sig > 0 || c != '0') {
                if (++sig > 15) return false;                      // M must stay below 2^53
            }
            M = M * 10ull + (unsigned long long)(c - '0');
            if (frac && ++k > 22) return false;                     // 10^k must be an exact double
        } else if (c == '.' && !frac) {
            frac = true;
        } else if (csv_is_sep(c)) {
            break;
        } else {
            return false;                                           // exponent, inf, nan, hex, garbage: strtof decides
        }
    }
    if (p < n && !csv_is_sep(text[p])) return false;                // longer than CSV_MAX_TOKEN
    if (digits == 0) return false;
    const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16,
                            1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const double d = __ddiv_rn((double)M, P10[k]);                  // correctly rounded double of the decimal
    if (M != 0) {
        if (d < 2.3509887016445750e-38 || d > 1.7014118346046923e38) return false;   // keep to normal floats with headroom
        // double rounding guard: d within one ulp of the midpoint between two floats (29 discarded bits)
        const unsigned long long low = (unsigned long long)__double_as_longlong(d) & 0x1fffffffull;
        if (low >= 0x0fffffffull && low <= 0x10000001ull) return false;
    }
    const float f = __double2float_rn(d);
    *out = neg ? -f : f;
    return true;
}

__global__ void __launch_bounds__(CSV_THREADS)
csv_parse_kernel(const unsigned char *__restrict__ text, size_t n, const unsigned long long *__restrict__ chunk_offsets,
                 float *__restrict__ values, unsigned long long max_values, unsigned long long *__restrict__ scalars,
                 CsvFallback *__restrict__ fallback)
{
    __shared__ int wsum[CSV_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t base = ((size_t)blockIdx.x * CSV_THREADS + threadIdx.x) * CSV_BYTES_PER_THREAD;
    unsigned flags = csv_start_flags(text, n, base);
    const int mine = __popc(flags);
    int inc = mine;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= s) inc += o;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int before = inc - mine;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    unsigned long long idx = chunk_offsets[blockIdx.x] + (unsigned long long)before;
    while (flags) {
        const int b = __ffs(flags) - 1;
        flags &= flags - 1;
        if (idx < max_values) {
            float v = 0.0f;
            if (csv_convert(text, n, base + b, &v)) {
                values[idx] = v;
            } else {
                const unsigned long long slot = atomicAdd(&scalars[1], 1ull);
                if (slot < CSV_MAX_FALLBACK) { fallback[slot].index = idx; fallback[slot].offset = base + b; }
            }
        }
        ++idx;
    }
}

}  // namespace

void csv_free(b200slam_ctx *ctx)
{
    cudaFree(ctx->d_csv_text); cudaFree(ctx->d_csv_values); cudaFree(ctx->d_csv_scratch);
    cudaFreeHost(ctx->h_csv_scratch);
    ctx->d_csv_text = nullptr; ctx->d_csv_values = nullptr; ctx->d_csv_scratch = nullptr; ctx->h_csv_scratch = nullptr;
    ctx->csv_text_cap = ctx->csv_values_cap = 0;
    ctx->csv_count = 0;
}

extern "C" {

int b200slam_csv_ingest(b200slam_ctx *ctx, const char *text, size_t nbytes, float *values_out, int64_t max_values,
                        int64_t *count)
{
    if (!ctx || (!text && nbytes) || max_values < 0) return B200SLAM_ERR_ARG;
    if (count) *count = 0;
    ctx->csv_count = 0;
    if (nbytes == 0) return B200SLAM_OK;
    const int nb = (int)((nbytes + CSV_CHUNK - 1) / CSV_CHUNK);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (nbytes > ctx->csv_text_cap) {
        cudaFree(ctx->d_csv_text); cudaFree(ctx->d_csv_scratch); cudaFreeHost(ctx->h_csv_scratch);
        ctx->d_csv_text = nullptr; ctx->d_csv_scratch = nullptr; ctx->h_csv_scratch = nullptr;
        ctx->csv_text_cap = 0;
        const size_t cap = (nbytes + (1u << 20)) & ~(size_t)((1u << 20) - 1);
        const size_t nchunks = cap / CSV_CHUNK + 1;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_csv_text, cap));
        // scratch: [2] scalars | [nchunks] chunk counts | fallback list
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_csv_scratch, sizeof(unsigned long long) * (2 + nchunks) + sizeof(CsvFallback) * CSV_MAX_FALLBACK));
        CUDA_TRY(ctx, cudaHostAlloc(&ctx->h_csv_scratch, sizeof(unsigned long long) * 2 + sizeof(CsvFallback) * CSV_MAX_FALLBACK,
                                    cudaHostAllocDefault));
        ctx->csv_text_cap = cap;
    }
    // at most one value per two bytes
    const size_t vmax = (size_t)max_values < nbytes / 2 + 1 ? (size_t)max_values : nbytes / 2 + 1;
    if (vmax > ctx->csv_values_cap) {
        cudaFree(ctx->d_csv_values);
        ctx->d_csv_values = nullptr;
        ctx->csv_values_cap = 0;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_csv_values, sizeof(float) * (vmax + 1)));
        ctx->csv_values_cap = vmax;
    }
    unsigned long long *scalars = ctx->d_csv_scratch;
    unsigned long long *chunks = scalars + 2;
    CsvFallback *fb = reinterpret_cast<CsvFallback *>(chunks + (ctx->csv_text_cap / CSV_CHUNK + 1));
    // the text is read straight from the caller's buffer (pageable memory goes through the driver's staging)
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_csv_text, text, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned char *d_text = reinterpret_cast<const unsigned char *>(ctx->d_csv_text);
    csv_count_kernel<<<nb, CSV_THREADS, 0, ctx->stream>>>(d_text, nbytes, chunks);
    LAUNCH_CHECK(ctx);
    csv_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(chunks, nb, scalars);
    LAUNCH_CHECK(ctx);
    csv_parse_kernel<<<nb, CSV_THREADS, 0, ctx->stream>>>(d_text, nbytes, chunks, ctx->d_csv_values, (unsigned long long)vmax,
                                                          scalars, fb);
    LAUNCH_CHECK(ctx);
    unsigned long long *h = ctx->h_csv_scratch;
    CUDA_TRY(ctx, cudaMemcpyAsync(h, scalars, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const unsigned long long total = h[0], nfb = h[1];
    const unsigned long long kept = total < vmax ? total : vmax;
    if (nfb > 0) {
        // the tokens the device did not convert: strtof, the conversion fscanf("%f") performs (main.c:27)
        if (nfb > CSV_MAX_FALLBACK)
            return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "%llu values are not plain decimals (limit %d): not a numeric CSV?",
                                      nfb, CSV_MAX_FALLBACK);
        CsvFallback *hf = reinterpret_cast<CsvFallback *>(h + 2);
        CUDA_TRY(ctx, cudaMemcpyAsync(hf, fb, sizeof(CsvFallback) * nfb, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        char tok[512];
        for (unsigned long long i = 0; i < nfb; ++i) {
            const size_t off = (size_t)hf[i].offset;
            size_t len = 0;
            while (off + len < nbytes && len < sizeof(tok) - 1) {
                const unsigned char c = (unsigned char)text[off + len];
                if (c == ',' || c == ' ' || (c >= '\t' && c <= '\r')) break;
                tok[len] = (char)c;
                ++len;
            }
            tok[len] = 0;
            char *end = nullptr;
            const float v = strtof(tok, &end);
            if (end == tok || *end != 0)
                return b200slam_set_error(ctx, B200SLAM_ERR_ARG, "malformed value \"%s\" at byte %zu of the CSV text", tok, off);
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_csv_values + hf[i].index, &v, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // v lives on this stack frame
        }
    }
    ctx->csv_count = (int64_t)kept;
    if (values_out && kept > 0) {
        CUDA_TRY(ctx, cudaMemcpyAsync(values_out, ctx->d_csv_values, sizeof(float) * kept, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (count) *count = (int64_t)kept;
    return B200SLAM_OK;
}

int b200slam_csv_values(b200slam_ctx *ctx, const float **device_values, int64_t *count)
{
    if (!ctx) return B200SLAM_ERR_ARG;
    if (device_values) *device_values = ctx->d_csv_values;
    if (count) *count = ctx->csv_count;
    return B200SLAM_OK;
}

}  // extern "C"
