// lgr_vq_tc.cuh -- nearest-code search of the VecTree k-means (vectree/vq.py:262-306: cdist + argmax) on the 5th-generation tensor cores.
//
// score[i][c] = |e_c|^2 - 2 x_i . e_c is a GEMM with a 27-long contraction (n = 80 000 samples x K = 8192 codes).  bf16 operands alone
// would flip nearest-code decisions, and the parity contract is on the INDICES, so the kernel does a coarse pass on the tensor cores and an
// exact FP32 rescore of the rows it cannot decide:
//
//   operands   x = x_hi + x_lo, e = e_hi + e_lo (two bf16 terms each, 16 bits of mantissa); the contraction is laid out as
//              A' = [x_hi | x_hi | x_lo] (96 columns), B' = [e_hi | e_lo | e_hi], so ONE bf16 GEMM with K = 96 yields
//              x_hi.e_hi + x_hi.e_lo + x_lo.e_hi with FP32 accumulation; the dropped terms are bounded by 3 * 2^-18 |x||e|.
//   kernel     one CTA = 128 samples (UMMA M = 128) against all codes in tiles of 256 (UMMA N = 256): six tcgen05.mma (K = 16 each) per
//              tile, issued by one thread, operands in shared memory as canonical K-major core-matrix tiles (no swizzle) that a prepare
//              kernel wrote to global memory in exactly that order, so a tile arrives with ONE cp.async.bulk; accumulators in TMEM, two
//              stages of 256 columns (all 512), so the MMAs of tile t+1 run under the epilogue of tile t.  Four epilogue warps read their
//              32 TMEM lanes with tcgen05.ld (32 columns at a time), form score = |e|^2 - 2 acc and keep (best, second best, argmin);
//              smallest index wins ties like the first-index argmax.
//   decision   a row whose second-best score is within 2 * 1e-4 * |x_i| * max|e| of its best (>= 4x the error bound) is NOT decided here:
//              its slot keeps ~0 and the exact FP32 kernel (vq_assign_kernel, lgr_vq.cuh) runs on the list of such rows (a few per cent).
//              Every other row's argmin is provably the exact one.
#pragma once

namespace {

constexpr int VT_M = 128;          // samples per CTA
constexpr int VT_N = 256;          // codes per tile
constexpr int VT_DP = 32;          // padded feature dimension
constexpr int VT_K = 3 * VT_DP;    // contraction length of the split GEMM
constexpr int VT_A_BYTES = VT_M * VT_K * 2;
constexpr int VT_B_BYTES = VT_N * VT_K * 2;
constexpr int VT_THREADS = 192;    // warps 0-3 epilogue, warp 4 loads, warp 5 issues the MMAs and owns TMEM
constexpr float VT_MARGIN = 2.0e-4f;

__device__ __forceinline__ uint32_t vt_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// element (r, k) of a K-major operand tile with ROWS rows, canonical no-swizzle core-matrix order:
// 8 x 16-byte core matrices, 128 bytes each; matrices adjacent in M/N are 128 B apart (SBO), matrices adjacent in K are (ROWS/8)*128 B apart (LBO)
__host__ __device__ inline size_t vt_tile_offset(int rows, int r, int k) { return ((size_t)(k >> 3) * (rows >> 3) + (r >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2; }

__device__ __forceinline__ unsigned short vt_bf16_rn(float f)
{
    const unsigned u = __float_as_uint(f);
    return (unsigned short)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);   // round to nearest even (finite inputs)
}
__device__ __forceinline__ float vt_bf16_f(unsigned short h) { return __uint_as_float((unsigned)h << 16); }

// rows of x -> A' tiles [n_pad/128][128 x 96] and |x_i|
__global__ void __launch_bounds__(256) vt_prep_x_kernel(int n, int n_pad, int d, const float* __restrict__ x, unsigned char* __restrict__ A, float* __restrict__ xnorm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    unsigned char* tile = A + (size_t)(i / VT_M) * VT_A_BYTES;
    const int r = i % VT_M;
    float s = 0.f;
    for (int k = 0; k < VT_DP; k++) {
        const float v = (i < n && k < d) ? x[(size_t)i * d + k] : 0.f;
        s = fmaf(v, v, s);
        const unsigned short hi = vt_bf16_rn(v), lo = vt_bf16_rn(v - vt_bf16_f(hi));
        *reinterpret_cast<unsigned short*>(tile + vt_tile_offset(VT_M, r, k)) = hi;
        *reinterpret_cast<unsigned short*>(tile + vt_tile_offset(VT_M, r, VT_DP + k)) = hi;
        *reinterpret_cast<unsigned short*>(tile + vt_tile_offset(VT_M, r, 2 * VT_DP + k)) = lo;
    }
    xnorm[i] = sqrtf(s);
}

// codes -> B' tiles [K_pad/256][256 x 96], |e_c|^2 (+huge for padding) and max |e_c| (bits of a non-negative float order like integers)
__global__ void __launch_bounds__(256) vt_prep_e_kernel(int K, int K_pad, int d, const float* __restrict__ e, unsigned char* __restrict__ B, float* __restrict__ norms,
                                                        unsigned* __restrict__ emax_bits)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K_pad) return;
    unsigned char* tile = B + (size_t)(c / VT_N) * VT_B_BYTES;
    const int r = c % VT_N;
    float s = 0.f;
    for (int k = 0; k < VT_DP; k++) {
        const float v = (c < K && k < d) ? e[(size_t)c * d + k] : 0.f;
        s = fmaf(v, v, s);
        const unsigned short hi = vt_bf16_rn(v), lo = vt_bf16_rn(v - vt_bf16_f(hi));
        *reinterpret_cast<unsigned short*>(tile + vt_tile_offset(VT_N, r, k)) = hi;
        *reinterpret_cast<unsigned short*>(tile + vt_tile_offset(VT_N, r, VT_DP + k)) = lo;
        *reinterpret_cast<unsigned short*>(tile + vt_tile_offset(VT_N, r, 2 * VT_DP + k)) = hi;
    }
    norms[c] = c < K ? s : 3.0e38f;
    if (c < K) atomicMax(emax_bits, __float_as_uint(sqrtf(s)));
}

// ---- mbarrier / tcgen05 wrappers ----
__device__ __forceinline__ void vt_mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vt_smem(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void vt_mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vt_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vt_mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(vt_smem(bar)) : "memory"); }
__device__ __forceinline__ void vt_mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(vt_smem(bar)), "r"(parity)
                     : "memory");
        if (!ok && spin > (1u << 26)) {   // a protocol bug becomes an error, not a hung GPU
            printf("lgrast: vq tensor-core pipeline barrier timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void vt_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(vt_smem(dst)), "l"(src), "r"(bytes),
                 "r"(vt_smem(bar))
                 : "memory");
}
__device__ __forceinline__ uint64_t vt_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    // UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address [0,14), leading byte offset [16,30),
    // stride byte offset [32,46) -- all without their 4 LSBs --, version 1 at [46,48), layout type 0 = no swizzle at [61,64)
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) |
           (1ull << 46);
}
// instruction descriptor, kind::f16: D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t VT_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(VT_N >> 3) << 17) | ((uint32_t)(VT_M >> 4) << 24);

__device__ __forceinline__ void vt_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a),
                 "l"(desc_b), "r"(VT_IDESC), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void vt_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(vt_smem(bar)) : "memory");
}

struct VtSmem {
    unsigned char a[VT_A_BYTES];         // 24 576 B
    unsigned char b[2][VT_B_BYTES];      // 2 x 49 152 B
    float norm[2][VT_N];
    uint64_t full_a, full_b[2], acc_full[2], stage_free[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(VT_THREADS, 1)
vt_assign_kernel(int n, int K_pad, const unsigned char* __restrict__ A, const unsigned char* __restrict__ B, const float* __restrict__ norms,
                 const float* __restrict__ xnorm, const unsigned* __restrict__ emax_bits, unsigned long long* __restrict__ best, int* __restrict__ undecided,
                 int* __restrict__ n_undecided)
{
    extern __shared__ __align__(1024) unsigned char vt_dyn[];
    VtSmem& sm = *reinterpret_cast<VtSmem*>(vt_dyn);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = K_pad / VT_N;

    if (threadIdx.x == 0) {
        vt_mbar_init(&sm.full_a, 1);
        for (int s = 0; s < 2; s++) {
            vt_mbar_init(&sm.full_b[s], 1);
            vt_mbar_init(&sm.acc_full[s], 1);
            vt_mbar_init(&sm.stage_free[s], 4);   // lane 0 of each epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {   // one warp allocates all 512 TMEM columns (2 accumulator stages of 256) and gives the permit back
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(vt_smem(&sm.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm.tmem_base;

    if (warp == 4) {
        // ===== loads: the CTA's sample tile once, then the code tiles through two stages =====
        if (lane == 0) {
            vt_mbar_expect_tx(&sm.full_a, VT_A_BYTES);
            vt_bulk_g2s(sm.a, A + (size_t)blockIdx.x * VT_A_BYTES, VT_A_BYTES, &sm.full_a);
            for (int t = 0; t < ntiles; t++) {
                const int s = t & 1;
                const uint32_t ph = (uint32_t)(t >> 1) & 1u;
                vt_mbar_wait(&sm.stage_free[s], ph ^ 1u);   // passes at once the first time round
                vt_mbar_expect_tx(&sm.full_b[s], VT_B_BYTES + VT_N * 4);
                vt_bulk_g2s(sm.b[s], B + (size_t)t * VT_B_BYTES, VT_B_BYTES, &sm.full_b[s]);
                vt_bulk_g2s(sm.norm[s], norms + (size_t)t * VT_N, VT_N * 4, &sm.full_b[s]);
            }
        }
    } else if (warp == 5) {
        // ===== MMA issue: one thread, six K = 16 steps per 128 x 256 tile =====
        if (lane == 0) {
            vt_mbar_wait(&sm.full_a, 0);
            const uint32_t a_addr = vt_smem(sm.a);
            for (int t = 0; t < ntiles; t++) {
                const int s = t & 1;
                const uint32_t ph = (uint32_t)(t >> 1) & 1u;
                vt_mbar_wait(&sm.stage_free[s], ph ^ 1u);   // the epilogue has drained accumulator stage s
                vt_mbar_wait(&sm.full_b[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_addr = vt_smem(sm.b[s]);
#pragma unroll
                for (int k = 0; k < VT_K / 16; k++) {
                    const uint64_t da = vt_desc(a_addr + (uint32_t)k * 2u * (VT_M / 8) * 128u, (VT_M / 8) * 128u, 128u);
                    const uint64_t db = vt_desc(b_addr + (uint32_t)k * 2u * (VT_N / 8) * 128u, (VT_N / 8) * 128u, 128u);
                    vt_mma(tmem + (uint32_t)s * VT_N, da, db, k > 0 ? 1u : 0u);
                }
                vt_commit(&sm.acc_full[s]);   // arrives when the six MMAs have completed (implies fence::before_thread_sync)
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32w .. 32w+31 = samples 32w + lane of the tile =====
        const int row = blockIdx.x * VT_M + warp * 32 + lane;
        float bestv = 3.0e38f, second = 3.0e38f;
        int arg = 0;
        for (int t = 0; t < ntiles; t++) {
            const int s = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            vt_mbar_wait(&sm.full_b[s], ph);     // the tile's |e|^2 (bulk copy, async proxy) is visible to this thread
            vt_mbar_wait(&sm.acc_full[s], ph);   // ... and its accumulators are complete
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float* nrm = sm.norm[s];
#pragma unroll 1
            for (int c0 = 0; c0 < VT_N; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * VT_N + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                    "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                      "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                      "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
                      "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int cbase = t * VT_N + c0;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const float sc = fmaf(-2.0f, __uint_as_float(r[j]), nrm[c0 + j]);
                    if (sc < bestv) {
                        second = bestv;
                        bestv = sc;
                        arg = cbase + j;
                    } else if (sc < second) {
                        second = sc;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) vt_mbar_arrive(&sm.stage_free[s]);
        }
        if (row < n) {
            const float margin = 2.0f * VT_MARGIN * xnorm[row] * __uint_as_float(*emax_bits);
            if (second - bestv > margin) {
                best[row] = ((unsigned long long)vq_order_bits(bestv) << 32) | (unsigned)arg;
            } else {
                best[row] = ~0ull;   // undecided here: the exact FP32 kernel fills it
                undecided[atomicAdd(n_undecided, 1)] = row;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// the exact FP32 kernel of lgr_vq.cuh on a LIST of rows (the ones the tensor-core pass left undecided)
template <int DP, int RX>
__global__ void __launch_bounds__(VQ_THREADS)
vq_assign_rows_kernel(const int* __restrict__ rows, const int* __restrict__ n_rows, int d, int K, const float* __restrict__ x, const float* __restrict__ embed,
                      int codes_per_split, unsigned long long* __restrict__ best)
{
    __shared__ __align__(16) float s_e[VQ_TC][DP];
    __shared__ float s_n[VQ_TC];
    const int n = *n_rows;
    if ((int)blockIdx.x * (RX * VQ_THREADS) >= n) return;
    const int slot0 = blockIdx.x * (RX * VQ_THREADS) + threadIdx.x;
    float xr[RX][DP];
    int rowid[RX];
#pragma unroll
    for (int r = 0; r < RX; r++) {
        const int slot = slot0 + r * VQ_THREADS;
        rowid[r] = slot < n ? rows[slot] : -1;
#pragma unroll
        for (int j = 0; j < DP; j++) xr[r][j] = (j < d && rowid[r] >= 0) ? -2.0f * x[(size_t)rowid[r] * d + j] : 0.f;
    }
    const int c_begin = blockIdx.y * codes_per_split, c_end = min(K, c_begin + codes_per_split);
    float bestv[RX];
    int arg[RX];
#pragma unroll
    for (int r = 0; r < RX; r++) { bestv[r] = 3.0e38f; arg[r] = c_begin; }
    for (int c0 = c_begin; c0 < c_end; c0 += VQ_TC) {
        __syncthreads();
        for (int t = threadIdx.x; t < VQ_TC * DP; t += VQ_THREADS) {
            const int c = t / DP, j = t - c * DP;
            s_e[c][j] = (j < d && c0 + c < c_end) ? embed[(size_t)(c0 + c) * d + j] : 0.f;
        }
        __syncthreads();
        if (threadIdx.x < VQ_TC) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < DP; j++) s = fmaf(s_e[threadIdx.x][j], s_e[threadIdx.x][j], s);
            s_n[threadIdx.x] = (c0 + threadIdx.x < c_end) ? s : 3.0e38f;
        }
        __syncthreads();
#pragma unroll 2
        for (int c = 0; c < VQ_TC; c++) {
            float dist[RX];
            const float nrm = s_n[c];
#pragma unroll
            for (int r = 0; r < RX; r++) dist[r] = nrm;
            const float4* e4 = reinterpret_cast<const float4*>(s_e[c]);
#pragma unroll
            for (int q = 0; q < DP / 4; q++) {
                const float4 e = e4[q];
#pragma unroll
                for (int r = 0; r < RX; r++) {
                    dist[r] = fmaf(xr[r][4 * q + 0], e.x, dist[r]);
                    dist[r] = fmaf(xr[r][4 * q + 1], e.y, dist[r]);
                    dist[r] = fmaf(xr[r][4 * q + 2], e.z, dist[r]);
                    dist[r] = fmaf(xr[r][4 * q + 3], e.w, dist[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < RX; r++)
                if (dist[r] < bestv[r]) { bestv[r] = dist[r]; arg[r] = c0 + c; }
        }
    }
#pragma unroll
    for (int r = 0; r < RX; r++)
        if (rowid[r] >= 0) atomicMin(&best[rowid[r]], ((unsigned long long)vq_order_bits(bestv[r]) << 32) | (unsigned)arg[r]);
}

struct VtScratch {   // grow-only device scratch of the tensor-core path (operand tiles, norms, undecided list)
    void* p = nullptr;
    size_t bytes = 0;
    int device = -1;
};
inline size_t vt_align(size_t v) { return (v + 1023) / 1024 * 1024; }

}  // namespace
