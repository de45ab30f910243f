// Experiment (not product): which explicit float operation order reproduces torch's exp / sigmoid / F.normalize bit for bit?
#include <cuda_runtime.h>
extern "C" __global__ void k_exp(const float* x, float* y, int n, int variant)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = x[i];
    y[i] = variant == 0 ? expf(v) : __expf(v);
}
extern "C" __global__ void k_sigmoid(const float* x, float* y, int n, int variant)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = x[i];
    if (variant == 0) y[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
    else if (variant == 1) y[i] = __frcp_rn(__fadd_rn(1.0f, expf(-v)));
    else y[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, __expf(-v)));
}
extern "C" __global__ void k_normalize(const float4* x, float4* y, int n, int variant)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 q = x[i];
    float s;
    switch (variant) {
        case 0: s = __fmaf_rn(q.w, q.w, __fmaf_rn(q.z, q.z, __fmaf_rn(q.y, q.y, __fmul_rn(q.x, q.x)))); break;
        case 1: s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y)), __fmul_rn(q.z, q.z)), __fmul_rn(q.w, q.w)); break;
        case 2: s = __fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y)), __fadd_rn(__fmul_rn(q.z, q.z), __fmul_rn(q.w, q.w))); break;
        case 3: s = __fadd_rn(__fmaf_rn(q.y, q.y, __fmul_rn(q.x, q.x)), __fmaf_rn(q.w, q.w, __fmul_rn(q.z, q.z))); break;
        case 4: s = __fmaf_rn(q.x, q.x, __fmaf_rn(q.y, q.y, __fmaf_rn(q.z, q.z, __fmul_rn(q.w, q.w)))); break;
        case 5: s = __fmaf_rn(q.w, q.w, __fmaf_rn(q.z, q.z, __fmaf_rn(q.y, q.y, __fmaf_rn(q.x, q.x, 0.0f)))); break;
        case 6: s = __fadd_rn(__fadd_rn(__fmaf_rn(q.x, q.x, 0.f), __fmaf_rn(q.z, q.z, 0.f)), __fadd_rn(__fmaf_rn(q.y, q.y, 0.f), __fmaf_rn(q.w, q.w, 0.f))); break;
        case 7: s = __fadd_rn(__fmaf_rn(q.z, q.z, __fmul_rn(q.x, q.x)), __fmaf_rn(q.w, q.w, __fmul_rn(q.y, q.y))); break;
        default: s = __fadd_rn(__fmaf_rn(q.w, q.w, __fmul_rn(q.y, q.y)), __fmaf_rn(q.z, q.z, __fmul_rn(q.x, q.x))); break;
    }
    float nrm = __fsqrt_rn(s);
    float d = fmaxf(nrm, 1e-12f);
    y[i] = make_float4(__fdiv_rn(q.x, d), __fdiv_rn(q.y, d), __fdiv_rn(q.z, d), __fdiv_rn(q.w, d));
}
extern "C" void run_exp(const float* x, float* y, int n, int v) { k_exp<<<(n + 255) / 256, 256>>>(x, y, n, v); }
extern "C" void run_sigmoid(const float* x, float* y, int n, int v) { k_sigmoid<<<(n + 255) / 256, 256>>>(x, y, n, v); }
extern "C" void run_normalize(const float* x, float* y, int n, int v) { k_normalize<<<(n + 255) / 256, 256>>>((const float4*)x, (float4*)y, n, v); }
