// lgr_loss.cuh -- fused image loss of the training loops:
//     loss = (1 - lambda) * mean|x - y| + lambda * (1 - SSIM(x, y))          prune_finetune.py:160-164, distill_train.py:142-145
// with SSIM exactly as utils/loss_utils.py:45-85: 11x11 Gaussian window (sigma 1.5), zero padding 5, per channel,
// C1 = 0.01^2, C2 = 0.03^2, mean over all C*H*W entries of the SSIM map.
//
// The reference evaluates it with 5 depthwise conv2d + ~15 elementwise kernels and their autograd backward.  Here:
//   image_loss_forward_kernel   one pass: 42x42 halo tile of x and y in shared memory, separable 11-tap filter of the five
//                               moments (x, y, xx, yy, xy), SSIM map value, the three partial-derivative maps the backward
//                               needs, and per-block partial sums of |x-y| and SSIM (reduced in a fixed order: deterministic)
//   image_loss_backward_kernel  dL/dx = s * ( g_l1 * sign(x-y) + g_ssim * ( w*A + 2x (w*B) + y (w*C) ) ) / N, the three
//                               correlations again as separable shared-memory passes (the window is symmetric).
// With m = a1 a2 / (b1 b2), a1 = 2 mu1 mu2 + C1, a2 = 2 s12 + C2, b1 = mu1^2 + mu2^2 + C1, b2 = s1 + s2 + C2:
//   B = dm/ds1 = -a1 a2 / (b1 b2^2),  C = dm/ds12 = 2 a1 / (b1 b2),
//   A = dm/dmu1 - 2 mu1 B - mu2 C,    dm/dmu1 = 2 mu2 a2 / (b1 b2) - 2 mu1 a1 a2 / (b1^2 b2).
#pragma once

namespace {

constexpr int LT = 32;            // output tile edge
constexpr int LHALO = 5;          // window_size // 2
constexpr int LIN = LT + 2 * LHALO;  // 42

struct LossWindow {
    float g[11];
};

__global__ void __launch_bounds__(256)
image_loss_forward_kernel(const float* __restrict__ x, const float* __restrict__ y, int C, int H, int W, LossWindow win,
                          float* __restrict__ dmaps /* nullable: [3][C][H][W] = A, B, C */, float2* __restrict__ partial)
{
    __shared__ float sx[LIN][LIN + 1], sy[LIN][LIN + 1];
    __shared__ float h[5][LIN][LT + 1];
    __shared__ float red[2][8];
    const int c = blockIdx.z, tx0 = blockIdx.x * LT, ty0 = blockIdx.y * LT;
    const size_t plane = (size_t)H * W;
    const float* xc = x + (size_t)c * plane;
    const float* yc = y + (size_t)c * plane;
    for (int i = threadIdx.x; i < LIN * LIN; i += 256) {
        const int r = i / LIN, q = i - r * LIN;
        const int gy = ty0 + r - LHALO, gx = tx0 + q - LHALO;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        sx[r][q] = in ? xc[(size_t)gy * W + gx] : 0.f;
        sy[r][q] = in ? yc[(size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LIN * LT; i += 256) {  // horizontal pass
        const int r = i / LT, q = i - r * LT;
        float m1 = 0.f, m2 = 0.f, m11 = 0.f, m22 = 0.f, m12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float a = sx[r][q + k], b = sy[r][q + k], w = win.g[k];
            m1 = fmaf(w, a, m1);
            m2 = fmaf(w, b, m2);
            m11 = fmaf(w, a * a, m11);
            m22 = fmaf(w, b * b, m22);
            m12 = fmaf(w, a * b, m12);
        }
        h[0][r][q] = m1; h[1][r][q] = m2; h[2][r][q] = m11; h[3][r][q] = m22; h[4][r][q] = m12;
    }
    __syncthreads();
    float sum_l1 = 0.f, sum_ssim = 0.f;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    for (int i = threadIdx.x; i < LT * LT; i += 256) {  // vertical pass + map
        const int r = i / LT, q = i - r * LT;
        const int gy = ty0 + r, gx = tx0 + q;
        if (gy >= H || gx >= W) continue;
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float w = win.g[k];
            mu1 = fmaf(w, h[0][r + k][q], mu1);
            mu2 = fmaf(w, h[1][r + k][q], mu2);
            e11 = fmaf(w, h[2][r + k][q], e11);
            e22 = fmaf(w, h[3][r + k][q], e22);
            e12 = fmaf(w, h[4][r + k][q], e12);
        }
        const float s1 = e11 - mu1 * mu1, s2 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
        const float a1 = 2.f * mu1 * mu2 + C1, a2 = 2.f * s12 + C2, b1 = mu1 * mu1 + mu2 * mu2 + C1, b2 = s1 + s2 + C2;
        const float inv = 1.f / (b1 * b2);
        const float m = a1 * a2 * inv;
        sum_ssim += m;
        sum_l1 += fabsf(sx[r + LHALO][q + LHALO] - sy[r + LHALO][q + LHALO]);
        if (dmaps) {
            const float dB = -m / b2;                 // dm/ds1
            const float dC = 2.f * a1 * inv;          // dm/ds12
            const float dmu1 = 2.f * mu2 * a2 * inv - 2.f * mu1 * m / b1;
            const size_t o = (size_t)c * plane + (size_t)gy * W + gx, cs = (size_t)C * plane;
            dmaps[o] = dmu1 - 2.f * mu1 * dB - mu2 * dC;
            dmaps[cs + o] = dB;
            dmaps[2 * cs + o] = dC;
        }
    }
    // deterministic block reduction
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum_l1 += __shfl_xor_sync(0xffffffffu, sum_l1, d);
        sum_ssim += __shfl_xor_sync(0xffffffffu, sum_ssim, d);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = sum_l1;
        red[1][threadIdx.x >> 5] = sum_ssim;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < 8; k++) { a += red[0][k]; b += red[1][k]; }
        partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = make_float2(a, b);
    }
}

// L1 only (utils/loss_utils.py:18-19): per-block partial sums of |x-y| over a grid-stride range, same finishing kernel
__global__ void __launch_bounds__(256) image_l1_forward_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n,
                                                               float2* __restrict__ partial)
{
    __shared__ float red[8];
    float s = 0.f;
    const long long n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 a = x4[i], b = y4[i];
        s += (fabsf(a.x - b.x) + fabsf(a.y - b.y)) + (fabsf(a.z - b.z) + fabsf(a.w - b.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) s += fabsf(x[(n4 << 2) + threadIdx.x] - y[(n4 << 2) + threadIdx.x]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int k = 0; k < 8; k++) a += red[k];
        partial[blockIdx.x] = make_float2(a, 0.f);
    }
}

// out[0] = mean|x-y|, out[1] = mean SSIM; fixed summation order (double accumulation in one block)
__global__ void __launch_bounds__(256) image_loss_finish_kernel(const float2* __restrict__ partial, int n, double inv_count, float* __restrict__ out)
{
    __shared__ double r0[256], r1[256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        a += (double)partial[i].x;
        b += (double)partial[i].y;
    }
    r0[threadIdx.x] = a;
    r1[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            r0[threadIdx.x] += r0[threadIdx.x + s];
            r1[threadIdx.x] += r1[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = (float)(r0[0] * inv_count);
        out[1] = (float)(r1[0] * inv_count);
    }
}

__global__ void __launch_bounds__(256)
image_loss_backward_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dmaps, int C, int H, int W,
                           LossWindow win, float g_l1, float g_ssim, const float* __restrict__ grad_scale /* device scalar or NULL */,
                           float inv_count, float* __restrict__ dx)
{
    __shared__ float sm[3][LIN][LIN + 1];
    __shared__ float h[3][LIN][LT + 1];
    const int c = blockIdx.z, tx0 = blockIdx.x * LT, ty0 = blockIdx.y * LT;
    const size_t plane = (size_t)H * W, cs = (size_t)C * plane;
    const float scale = (grad_scale ? *grad_scale : 1.0f) * inv_count;
    if (g_ssim != 0.f) {
        for (int i = threadIdx.x; i < LIN * LIN; i += 256) {
            const int r = i / LIN, q = i - r * LIN;
            const int gy = ty0 + r - LHALO, gx = tx0 + q - LHALO;
            const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
            const size_t o = (size_t)c * plane + (size_t)gy * W + gx;
            sm[0][r][q] = in ? dmaps[o] : 0.f;
            sm[1][r][q] = in ? dmaps[cs + o] : 0.f;
            sm[2][r][q] = in ? dmaps[2 * cs + o] : 0.f;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < LIN * LT; i += 256) {
            const int r = i / LT, q = i - r * LT;
            float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
            for (int k = 0; k < 11; k++) {
                const float w = win.g[k];
                a = fmaf(w, sm[0][r][q + k], a);
                b = fmaf(w, sm[1][r][q + k], b);
                d = fmaf(w, sm[2][r][q + k], d);
            }
            h[0][r][q] = a; h[1][r][q] = b; h[2][r][q] = d;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < LT * LT; i += 256) {
        const int r = i / LT, q = i - r * LT;
        const int gy = ty0 + r, gx = tx0 + q;
        if (gy >= H || gx >= W) continue;
        const size_t o = (size_t)c * plane + (size_t)gy * W + gx;
        const float xv = x[o], yv = y[o];
        float g = 0.f;
        if (g_ssim != 0.f) {
            float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
            for (int k = 0; k < 11; k++) {
                const float w = win.g[k];
                a = fmaf(w, h[0][r + k][q], a);
                b = fmaf(w, h[1][r + k][q], b);
                d = fmaf(w, h[2][r + k][q], d);
            }
            g = g_ssim * (a + 2.f * xv * b + yv * d);
        }
        const float df = xv - yv;
        g += g_l1 * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));   // torch.sign: 0 at 0
        dx[o] = scale * g;
    }
}

}  // namespace
