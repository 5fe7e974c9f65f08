// BatchNorm / ReLU / residual / concat / pooling kernels.
//
// All of these are HBM-streaming passes: row-major [n, c] fp32 with a leading dimension so that the
// halves of a channel concatenation are addressed in place; 128-bit accesses whenever c and the
// leading dimensions allow; column reductions go thread-local fp32 -> double shared -> one double
// atomic per (block, channel), so the statistics are accurate to fp32 round-off of the inputs.
//
// Replaces ME.MinkowskiBatchNorm (= nn.BatchNorm1d on .F, /root/reference/models/modules/common.py:20-22),
// MinkowskiReLU (models/res16unet.py:222), `out += residual` (models/modules/resnet_block.py:61),
// me.cat (models/res16unet.py:259-289) and MinkowskiAvgPooling (models/mask3d.py:131).
#include "common.cuh"

namespace us3d {

constexpr int kRedRows = 256;  // rows per reduction block

// column sums of f0(x), f1(x): block = 32 channels x 8 row lanes
template <class F>
__device__ __forceinline__ void column_reduce2(int n, int c, double *out0, double *out1, F f) {
    __shared__ double s0[8][33], s1[8][33];
    int ch = blockIdx.y * 32 + threadIdx.x;
    int r_begin = blockIdx.x * kRedRows, r_end = min(n, r_begin + kRedRows);
    float a0 = 0.f, a1 = 0.f;
    if (ch < c)
        for (int r = r_begin + threadIdx.y; r < r_end; r += 8) {
            float v0, v1;
            f(r, ch, v0, v1);
            a0 += v0;
            a1 += v1;
        }
    s0[threadIdx.y][threadIdx.x] = (double)a0;
    s1[threadIdx.y][threadIdx.x] = (double)a1;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c) {
        double t0 = 0, t1 = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            t0 += s0[i][threadIdx.x];
            t1 += s1[i][threadIdx.x];
        }
        atomicAdd(&out0[ch], t0);
        atomicAdd(&out1[ch], t1);
    }
}

__global__ void __launch_bounds__(256) k_bn_stats(const float *__restrict__ x, int ldx, int n, int c, double *sum,
                                                  double *sumsq) {
    column_reduce2(n, c, sum, sumsq, [&](int r, int ch, float &v0, float &v1) {
        float v = x[(size_t)r * ldx + ch];
        v0 = v;
        v1 = v * v;
    });
}

__global__ void k_bn_finalize(const double *__restrict__ sum, const double *__restrict__ sumsq, int n, int c, float eps,
                              float momentum, float *mean, float *invstd, float *running_mean, float *running_var) {
    int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    double m = sum[ch] / n;
    double var = sumsq[ch] / n - m * m;
    if (var < 0) var = 0;
    mean[ch] = (float)m;
    invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[ch] = (float)((1.0 - momentum) * running_mean[ch] + momentum * m);
    if (running_var) {
        double unbiased = n > 1 ? var * n / (n - 1.0) : var;
        running_var[ch] = (float)((1.0 - momentum) * running_var[ch] + momentum * unbiased);
    }
}

// one thread = 4 consecutive channels of one row (vec) or one element (scalar)
template <bool VEC>
__global__ void __launch_bounds__(256)
k_bn_apply(const float *__restrict__ x, int ldx, int n, int c, const float *__restrict__ mean,
           const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta,
           const float *__restrict__ res, int ldr, int relu, float *__restrict__ y, int ldy) {
    constexpr int W = VEC ? 4 : 1;
    const int cw = c / W;
    long long total = (long long)n * cw;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e / cw), ch = (int)(e % cw) * W;
        float v[W], rs[W];
        if constexpr (VEC) {
            float4 t = *reinterpret_cast<const float4 *>(x + (size_t)r * ldx + ch);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            if (res) {
                float4 q = *reinterpret_cast<const float4 *>(res + (size_t)r * ldr + ch);
                rs[0] = q.x; rs[1] = q.y; rs[2] = q.z; rs[3] = q.w;
            }
        } else {
            v[0] = x[(size_t)r * ldx + ch];
            if (res) rs[0] = res[(size_t)r * ldr + ch];
        }
#pragma unroll
        for (int i = 0; i < W; ++i) {
            float o = (v[i] - mean[ch + i]) * invstd[ch + i] * gamma[ch + i] + beta[ch + i];
            if (res) o += rs[i];
            if (relu) o = fmaxf(o, 0.f);
            v[i] = o;
        }
        if constexpr (VEC)
            *reinterpret_cast<float4 *>(y + (size_t)r * ldy + ch) = make_float4(v[0], v[1], v[2], v[3]);
        else
            y[(size_t)r * ldy + ch] = v[0];
    }
}

__global__ void __launch_bounds__(256)
k_bn_bwd_reduce(const float *__restrict__ dy, int lddy, const float *__restrict__ x, int ldx, const float *__restrict__ y,
                int ldy, int n, int c, const float *__restrict__ mean, const float *__restrict__ invstd, int relu,
                double *red) {
    column_reduce2(n, c, red, red + c, [&](int r, int ch, float &v0, float &v1) {
        float g = dy[(size_t)r * lddy + ch];
        if (relu && !(y[(size_t)r * ldy + ch] > 0.f)) g = 0.f;
        float xh = (x[(size_t)r * ldx + ch] - mean[ch]) * invstd[ch];
        v0 = g;
        v1 = g * xh;
    });
}

__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const float *__restrict__ dy, int lddy, const float *__restrict__ x, int ldx, const float *__restrict__ y,
               int ldy, int n, int c, const float *__restrict__ mean, const float *__restrict__ invstd,
               const float *__restrict__ gamma, int relu, const double *__restrict__ red, float *__restrict__ dx, int lddx,
               float *__restrict__ dres, int lddres, float *dgamma, float *dbeta, int batch_terms) {
    long long total = (long long)n * c;
    const float inv_n = 1.f / (float)n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e / c), ch = (int)(e % c);
        float g = dy[(size_t)r * lddy + ch];
        if (relu && !(y[(size_t)r * ldy + ch] > 0.f)) g = 0.f;
        if (dres) dres[(size_t)r * lddres + ch] = g;
        float is = invstd[ch];
        float xh = (x[(size_t)r * ldx + ch] - mean[ch]) * is;
        // batch_terms = 0: the statistics were constants (inference), dx = g * invstd * gamma
        float s0 = batch_terms ? (float)red[ch] : 0.f, s1 = batch_terms ? (float)red[c + ch] : 0.f;
        dx[(size_t)r * lddx + ch] = (g - s0 * inv_n - xh * s1 * inv_n) * is * gamma[ch];
    }
    if (blockIdx.x == 0)
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            if (dgamma) dgamma[ch] = (float)red[c + ch];
            if (dbeta) dbeta[ch] = (float)red[ch];
        }
}

// ---- flat elementwise -----------------------------------------------------------------------------
template <int OP>  // 0 relu, 1 relu_bwd (a=dy, b=y), 2 add
__device__ __forceinline__ float flat_op(float va, float vb) {
    return OP == 0 ? fmaxf(va, 0.f) : (OP == 1 ? (vb > 0.f ? va : 0.f) : va + vb);
}

template <int OP>
__global__ void __launch_bounds__(256) k_flat(const float *a, const float *b, float *z,
                                              long long numel, int vec) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long done = 0;
    if (vec) {
        long long n4 = numel / 4;
        for (long long i = gid; i < n4; i += stride) {
            float4 va = reinterpret_cast<const float4 *>(a)[i], vb = va;
            if (OP != 0) vb = reinterpret_cast<const float4 *>(b)[i];
            reinterpret_cast<float4 *>(z)[i] = make_float4(flat_op<OP>(va.x, vb.x), flat_op<OP>(va.y, vb.y),
                                                           flat_op<OP>(va.z, vb.z), flat_op<OP>(va.w, vb.w));
        }
        done = n4 * 4;
    }
    for (long long i = done + gid; i < numel; i += stride) z[i] = flat_op<OP>(a[i], OP != 0 ? b[i] : 0.f);
}

__global__ void __launch_bounds__(256) k_copy2d(const float *__restrict__ src, int lds, float *__restrict__ dst, int ldd, int n,
                                                int c, int vec) {
    const int W = vec ? 4 : 1;
    const int cw = c / W;
    long long total = (long long)n * cw;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e / cw), ch = (int)(e % cw) * W;
        if (vec)
            *reinterpret_cast<float4 *>(dst + (size_t)r * ldd + ch) = *reinterpret_cast<const float4 *>(src + (size_t)r * lds + ch);
        else
            dst[(size_t)r * ldd + ch] = src[(size_t)r * lds + ch];
    }
}

// ---- pooling over a (non-overlapping) neighbour table ------------------------------------------------
__global__ void __launch_bounds__(256) k_pool_fwd(const float *__restrict__ x, int c, const int32_t *__restrict__ nbr, int n_out,
                                                  int kvol, int mode, float *__restrict__ y) {
    long long total = (long long)n_out * c;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int o = (int)(e / c), ch = (int)(e % c);
        float acc = mode == 2 ? -INFINITY : 0.f;
        int cnt = 0;
        for (int k = 0; k < kvol; ++k) {
            int i = nbr[(size_t)k * n_out + o];
            if (i < 0) continue;
            float v = x[(size_t)i * c + ch];
            acc = mode == 2 ? fmaxf(acc, v) : acc + v;
            ++cnt;
        }
        if (mode == 0 && cnt > 0) acc /= (float)cnt;
        if (mode == 2 && cnt == 0) acc = 0.f;
        y[e] = acc;
    }
}

__global__ void __launch_bounds__(256) k_pool_bwd(const float *__restrict__ dy, const float *__restrict__ x,
                                                  const float *__restrict__ y, int c, const int32_t *__restrict__ nbr, int n_out,
                                                  int kvol, int mode, float *__restrict__ dx) {
    long long total = (long long)n_out * c;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int o = (int)(e / c), ch = (int)(e % c);
        int cnt = 0;
        for (int k = 0; k < kvol; ++k) cnt += nbr[(size_t)k * n_out + o] >= 0;
        float g = dy[e];
        if (mode == 0 && cnt > 0) g /= (float)cnt;
        bool given = false;
        for (int k = 0; k < kvol; ++k) {
            int i = nbr[(size_t)k * n_out + o];
            if (i < 0) continue;
            float out = g;
            if (mode == 2) {
                bool is_max = !given && x[(size_t)i * c + ch] == y[e];
                out = is_max ? g : 0.f;
                given = given || is_max;
            }
            dx[(size_t)i * c + ch] = out;  // stride == kernel: every input row has exactly one parent
        }
    }
}

static inline int flat_grid(long long work) {
    long long b = (work + 255) / 256;
    long long cap = (long long)num_sms() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace us3d

using namespace us3d;

extern "C" {

int us3d_bn_stats(const float *x, int ldx, int n, int c, double *sum, double *sumsq, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n >= 0 && c > 0 && ldx >= c, "bn_stats: bad shape");
    if (n == 0) return 0;
    dim3 grid(ceil_div(n, kRedRows), ceil_div(c, 32)), block(32, 8);
    k_bn_stats<<<grid, block, 0, st>>>(x, ldx, n, c, sum, sumsq);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_bn_finalize(const double *sum, const double *sumsq, int n, int c, float eps, float momentum, float *mean,
                     float *invstd, float *running_mean, float *running_var, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n > 0 && c > 0, "bn_finalize: bad shape");
    k_bn_finalize<<<ceil_div(c, 128), 128, 0, st>>>(sum, sumsq, n, c, eps, momentum, mean, invstd, running_mean, running_var);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_bn_apply(const float *x, int ldx, int n, int c, const float *mean, const float *invstd, const float *gamma,
                  const float *beta, const float *residual, int ldr, int relu, float *y, int ldy, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n >= 0 && c > 0 && ldx >= c && ldy >= c, "bn_apply: bad shape");
    if (n == 0) return 0;
    bool vec = c % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && al16(x) && al16(y) && (!residual || (ldr % 4 == 0 && al16(residual)));
    if (vec)
        k_bn_apply<true><<<flat_grid((long long)n * c / 4), 256, 0, st>>>(x, ldx, n, c, mean, invstd, gamma, beta, residual, ldr, relu, y, ldy);
    else
        k_bn_apply<false><<<flat_grid((long long)n * c), 256, 0, st>>>(x, ldx, n, c, mean, invstd, gamma, beta, residual, ldr, relu, y, ldy);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_bn_bwd_reduce(const float *dy, int lddy, const float *x, int ldx, const float *y, int ldy, int n, int c,
                       const float *mean, const float *invstd, int relu, double *red, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n >= 0 && c > 0, "bn_bwd_reduce: bad shape");
    if (n == 0) return 0;
    dim3 grid(ceil_div(n, kRedRows), ceil_div(c, 32)), block(32, 8);
    k_bn_bwd_reduce<<<grid, block, 0, st>>>(dy, lddy, x, ldx, y, ldy, n, c, mean, invstd, relu, red);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_bn_bwd_apply(const float *dy, int lddy, const float *x, int ldx, const float *y, int ldy, int n, int c,
                      const float *mean, const float *invstd, const float *gamma, int relu, const double *red, float *dx,
                      int lddx, float *dres, int lddres, float *dgamma, float *dbeta, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n > 0 && c > 0, "bn_bwd_apply: bad shape");
    k_bn_bwd_apply<<<flat_grid((long long)n * c), 256, 0, st>>>(dy, lddy, x, ldx, y, ldy, n, c, mean, invstd, gamma, relu, red, dx,
                                                              lddx, dres, lddres, dgamma, dbeta, 1);
    US3D_LAUNCH_CHECK();
    return 0;
}

/* one-call forms (fewer host round trips): workspace ws = 2*c doubles, zeroed here */
int us3d_bn_batch_stats(const float *x, int ldx, int n, int c, float eps, float momentum, float *mean, float *invstd,
                        float *running_mean, float *running_var, double *ws, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n > 0 && c > 0 && ldx >= c, "bn_batch_stats: bad shape");
    US3D_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * c, st));
    dim3 grid(ceil_div(n, kRedRows), ceil_div(c, 32)), block(32, 8);
    k_bn_stats<<<grid, block, 0, st>>>(x, ldx, n, c, ws, ws + c);
    US3D_LAUNCH_CHECK();
    k_bn_finalize<<<ceil_div(c, 128), 128, 0, st>>>(ws, ws + c, n, c, eps, momentum, mean, invstd, running_mean, running_var);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_bn_backward(const float *dy, int lddy, const float *x, int ldx, const float *y, int ldy, int n, int c,
                     const float *mean, const float *invstd, const float *gamma, int relu, int batch_terms, double *ws,
                     float *dx, int lddx, float *dres, int lddres, float *dgamma, float *dbeta, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    US3D_CHECK_ARG(n > 0 && c > 0, "bn_backward: bad shape");
    US3D_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * c, st));
    dim3 grid(ceil_div(n, kRedRows), ceil_div(c, 32)), block(32, 8);
    k_bn_bwd_reduce<<<grid, block, 0, st>>>(dy, lddy, x, ldx, y, ldy, n, c, mean, invstd, relu, ws);
    US3D_LAUNCH_CHECK();
    k_bn_bwd_apply<<<flat_grid((long long)n * c), 256, 0, st>>>(dy, lddy, x, ldx, y, ldy, n, c, mean, invstd, gamma, relu, ws, dx,
                                                              lddx, dres, lddres, dgamma, dbeta, batch_terms);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_relu(const float *x, float *y, long long numel, void *stream_) {
    if (numel <= 0) return 0;
    int vec = al16(x) && al16(y);
    k_flat<0><<<flat_grid(numel / 4 + 1), 256, 0, (cudaStream_t)stream_>>>(x, nullptr, y, numel, vec);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_relu_bwd(const float *dy, const float *y, float *dx, long long numel, void *stream_) {
    if (numel <= 0) return 0;
    int vec = al16(dy) && al16(y) && al16(dx);
    k_flat<1><<<flat_grid(numel / 4 + 1), 256, 0, (cudaStream_t)stream_>>>(dy, y, dx, numel, vec);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_add(const float *a, const float *b, float *z, long long numel, void *stream_) {
    if (numel <= 0) return 0;
    int vec = al16(a) && al16(b) && al16(z);
    k_flat<2><<<flat_grid(numel / 4 + 1), 256, 0, (cudaStream_t)stream_>>>(a, b, z, numel, vec);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_copy2d(const float *src, int lds, float *dst, int ldd, int n, int c, void *stream_) {
    US3D_CHECK_ARG(n >= 0 && c > 0 && lds >= c && ldd >= c, "copy2d: bad shape");
    if (n == 0) return 0;
    int vec = c % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && al16(src) && al16(dst);
    k_copy2d<<<flat_grid((long long)n * c / (vec ? 4 : 1)), 256, 0, (cudaStream_t)stream_>>>(src, lds, dst, ldd, n, c, vec);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_pool_fwd(const float *x, int c, const int32_t *nbr, int n_out, int kvol, int mode, float *y, void *stream_) {
    US3D_CHECK_ARG(mode >= 0 && mode <= 2 && kvol >= 1 && kvol <= US3D_MAX_KVOL, "pool_fwd: bad mode/kvol");
    if (n_out == 0) return 0;
    k_pool_fwd<<<flat_grid((long long)n_out * c), 256, 0, (cudaStream_t)stream_>>>(x, c, nbr, n_out, kvol, mode, y);
    US3D_LAUNCH_CHECK();
    return 0;
}

int us3d_pool_bwd(const float *dy, const float *x, const float *y, int c, const int32_t *nbr, int n_out, int kvol, int mode,
                  float *dx, void *stream_) {
    US3D_CHECK_ARG(mode >= 0 && mode <= 2 && kvol >= 1 && kvol <= US3D_MAX_KVOL, "pool_bwd: bad mode/kvol");
    if (n_out == 0) return 0;
    k_pool_bwd<<<flat_grid((long long)n_out * c), 256, 0, (cudaStream_t)stream_>>>(dy, x, y, c, nbr, n_out, kvol, mode, dx);
    US3D_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
