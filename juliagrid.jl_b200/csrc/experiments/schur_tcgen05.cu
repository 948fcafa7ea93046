// Measured A/B for the one dense contraction of the path, the frontal Schur update C -= L21 * U12 in FP64
// (reference role: the numeric factorisation behind lu!/klu!/ldlt!, /root/reference/src/backend/utility.jl:478-500):
//
//   (a) schur_dmma_kernel     mma.sync.aligned.m8n8k4.f64 (SASS DMMA) — what mf_factor_dense_{sym,lu}_kernel use;
//   (b) schur_tcgen05_kernel  tcgen05.mma kind::i8 with an Ozaki split: every FP64 operand row / column is scaled by a
//       power of two and cut into NS signed 8-bit slices (6 + 7 + 7 + ... bits), the slice products L_s * U_t are exact
//       in the int32 TMEM accumulators, pairs of equal weight (s + t = g) share an accumulator plane, and the epilogue
//       reads the planes back with tcgen05.ld, recombines them in FP64 and applies the row / column scales.
//
// tcgen05 has no FP64 kind, so (b) is the only way the update can reach the 5th-generation tensor cores at 1e-8 parity.
// The program times both on M x N = 128 x 128 trailing blocks with K = 32, 64, 96 pivots (the fronts of this path have
// 8-72 pivots), checks both against a long-double host reference and prints one line per case. Standalone: no part of
// libjgb200.so depends on it (the product uses (a); see DESIGN.md section 7 for the numbers).
//
// build: nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a schur_tcgen05.cu -o schur_tcgen05
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                           \
        }                                                                                      \
    } while (0)

constexpr int M = 128, N = 128;      // trailing block of one front
constexpr int NS = 8;                // int8 slices per operand: 6 + 7 * 7 = 55 bits below the row / column maximum
constexpr int NH = 64;               // accumulator planes are 128 lanes x 64 columns: 8 planes = the 512 TMEM columns

// ---------------------------------------------------------------------------------------------------------------------
// (a) FP64 tensor-core path: one CTA = one front, 8 warps, 8 x 8 tiles, L and U staged in padded shared-memory strips
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int LDP = 132;             // strip leading dimension, = 4 (mod 16): the four k-columns of a fragment hit distinct banks

// L: [front][M][K] row major, U: [front][K][N] row major, C: [front] column major M x N
// 8 warps; a warp keeps four 8 x 8 tiles of C in flight (loads of all four before the first multiply) so the global round
// trip of C is paid once per four tiles
__global__ void __launch_bounds__(256) schur_dmma_kernel(const double* __restrict__ L, const double* __restrict__ U,
                                                         double* __restrict__ C, int K) {
    extern __shared__ double sm[];
    double* Ls = sm;                 // (i, q) at Ls[i + q * LDP]
    double* Us = sm + (size_t)K * LDP;   // (q, j) at Us[j + q * LDP]
    const double* Lf = L + (size_t)blockIdx.x * M * K;
    const double* Uf = U + (size_t)blockIdx.x * K * N;
    double* Cf = C + (size_t)blockIdx.x * M * N;
    for (int t = threadIdx.x; t < M * K; t += blockDim.x) {
        const int i = t / K, q = t - i * K;
        Ls[i + q * LDP] = -Lf[t];
    }
    for (int t = threadIdx.x; t < K * N; t += blockDim.x) {
        const int q = t / N, j = t - q * N;
        Us[j + q * LDP] = Uf[t];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, g = lane >> 2, tg = lane & 3;
    constexpr int TI = 4;            // tiles in flight: four row tiles of one column of tiles (they share the U fragments)
    for (int t = warp; t < (M / 8 / TI) * (N / 8); t += nwarps) {
        const int tj = t / (M / 8 / TI), ti0 = (t - tj * (M / 8 / TI)) * TI;
        const int ja = 8 * tj + 2 * tg, jr = 8 * tj + g;
        double c0[TI], c1[TI];
#pragma unroll
        for (int x = 0; x < TI; ++x) {
            const int i = 8 * (ti0 + x) + g;
            c0[x] = Cf[i + (size_t)ja * M];
            c1[x] = Cf[i + (size_t)(ja + 1) * M];
        }
        for (int q0 = 0; q0 < K; q0 += 4) {
            const double b = Us[jr + (q0 + tg) * LDP];
#pragma unroll
            for (int x = 0; x < TI; ++x) dmma_m8n8k4(c0[x], c1[x], Ls[8 * (ti0 + x) + g + (q0 + tg) * LDP], b);
        }
#pragma unroll
        for (int x = 0; x < TI; ++x) {
            const int i = 8 * (ti0 + x) + g;
            Cf[i + (size_t)ja * M] = c0[x];
            Cf[i + (size_t)(ja + 1) * M] = c1[x];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// (b) tcgen05 path
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle, 8-bit operands: a K block of 32 bytes x R rows. Core matrix = 8 rows x 16 bytes (128 contiguous
// bytes); the two core matrices of a row group follow each other (leading byte offset 128), row groups are 256 bytes
// apart (stride byte offset 256): byte (r, kb) at (r / 8) * 256 + (kb / 16) * 128 + (r % 8) * 16 + kb % 16.
__device__ __forceinline__ int canon_off(int r, int kb) { return (r >> 3) * 256 + (kb >> 4) * 128 + (r & 7) * 16 + (kb & 15); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // SM100 shared-memory matrix descriptor: start address >> 4 in bits [0,14), leading byte offset >> 4 in [16,30), stride
    // byte offset >> 4 in [32,46), descriptor version 1 in [46,48), layout type (bits 61-63) 0 = no swizzle
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor, kind::i8: D = S32 (bits 4-5 = 2), A = B = signed 8 bit (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 in bits 17-22, M >> 4 in bits 24-28
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

// Four consecutive k of one operand row -> four digits per slice, packed into one 32-bit store per slice. x[] are the four
// values already scaled so that |x| < 1; base points at byte (row, k0) of slice 0 in the canonical layout.
__device__ __forceinline__ void slice4(const double x[4], int8_t* base, size_t slice_bytes) {
    double r[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) r[c] = x[c] * 64.0;               // slice 0: 6 bits, |digit| <= 64
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        uint32_t w = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double d = rint(r[c]);
            w |= ((uint32_t)(int)d & 0xffu) << (8 * c);
            r[c] = (r[c] - d) * 128.0;                            // the next 7 bits
        }
        *reinterpret_cast<uint32_t*>(base + s * slice_bytes) = w;
    }
}

// One CTA = one front, 128 threads: thread i owns row i of L (and of the accumulator lanes) and column i of U.
// Shared memory: Lsl [NS][K/32][4 KB] | Usl [NS][K/32][4 KB] | column scales [N] | reciprocal row / column scales [M + N]
// prof (nullable): cycles of CTA 0 spent in [0] scaling + slicing, [1] slice products (issue to commit), [2] epilogue
__global__ void __launch_bounds__(128) schur_tcgen05_kernel(const double* __restrict__ L, const double* __restrict__ U,
                                                            double* __restrict__ C, int K, long long* prof) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int kb_n = K / 32;
    const size_t slice_bytes = (size_t)kb_n * 4096;
    int8_t* Lsl = reinterpret_cast<int8_t*>(smraw);
    int8_t* Usl = Lsl + NS * slice_bytes;
    double* colscale = reinterpret_cast<double*>(Usl + NS * slice_bytes);
    const int tid = threadIdx.x, warp = tid >> 5;
    const double* Lf = L + (size_t)blockIdx.x * M * K;
    const double* Uf = U + (size_t)blockIdx.x * K * N;
    double* Cf = C + (size_t)blockIdx.x * M * N;

    if (warp == 0) {      // TMEM: all 512 columns (8 planes of 64), allocated and freed by warp 0
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    long long t0 = clock64(), t_mma = 0, t_epi = 0;
    // ---- 1. scales (thread i: row i of L, column i of U), then slices with all threads on groups of four k
    double* rs = colscale + N;                 // 1 / row scale of L, then 1 / column scale of U (slicing), N + M doubles more
    double rowscale;
    {
        double mx = 0.0;
        for (int q = 0; q < K; ++q) mx = fmax(mx, fabs(Lf[(size_t)tid * K + q]));
        const int e = mx > 0.0 ? ilogb(mx) + 1 : 0;
        rowscale = ldexp(1.0, e);
        rs[tid] = ldexp(1.0, -e);
        mx = 0.0;
        for (int q = 0; q < K; ++q) mx = fmax(mx, fabs(Uf[(size_t)q * N + tid]));
        const int f = mx > 0.0 ? ilogb(mx) + 1 : 0;
        colscale[tid] = ldexp(1.0, f);
        rs[M + tid] = ldexp(1.0, -f);
    }
    __syncthreads();
    for (int t = tid; t < M * (K / 4); t += blockDim.x) {          // L: row i, k0 = 4 * (t % (K / 4)): consecutive threads walk a row
        const int i = t / (K / 4), k0 = 4 * (t - i * (K / 4));
        const double sc = rs[i];
        double x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = Lf[(size_t)i * K + k0 + c] * sc;
        slice4(x, Lsl + (size_t)(k0 >> 5) * 4096 + canon_off(i, k0 & 31), slice_bytes);
    }
    for (int t = tid; t < N * (K / 4); t += blockDim.x) {          // U: column j = t % N (coalesced), k0 = 4 * (t / N)
        const int k0 = 4 * (t / N), j = t - (t / N) * N;
        const double sc = rs[M + j];
        double x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = Uf[(size_t)(k0 + c) * N + j] * sc;
        slice4(x, Usl + (size_t)(k0 >> 5) * 4096 + canon_off(j, k0 & 31), slice_bytes);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> tensor-core (async proxy) reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t_slice = clock64() - t0;
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t idesc = make_idesc(M, NH);
    const uint32_t lsl_a = smem_u32(Lsl), usl_a = smem_u32(Usl);

    for (int h = 0; h < N / NH; ++h) {
        t0 = clock64();
        // ---- 2. slice products: plane g collects the pairs s + t = g (weight 2^(-12 - 7 g)); one thread issues
        if (tid == 0) {
            for (int g = 0; g < NS; ++g) {
                uint32_t acc = 0;
                for (int s = 0; s <= g; ++s) {
                    const int t = g - s;
                    for (int kb = 0; kb < kb_n; ++kb) {
                        const uint64_t da = make_desc(lsl_a + (uint32_t)(s * slice_bytes) + kb * 4096);
                        const uint64_t db = make_desc(usl_a + (uint32_t)(t * slice_bytes) + kb * 4096 + h * (NH / 8) * 256);
                        mma_i8(tmem + g * NH, da, db, idesc, acc);
                        acc = 1;
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                         : "memory");
        }
        // ---- 3. epilogue: thread i = accumulator lane i; 8 columns of all 8 planes at a time, FP64 recombination
        mbar_wait(&bar, h & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        t_mma += clock64() - t0;
        t0 = clock64();
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < NH; c0 += 8) {
            uint32_t r[NS][8];
#pragma unroll
            for (int g = 0; g < NS; ++g)                  // all eight planes in flight, one wait
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(r[g][0]), "=r"(r[g][1]), "=r"(r[g][2]), "=r"(r[g][3]), "=r"(r[g][4]), "=r"(r[g][5]),
                               "=r"(r[g][6]), "=r"(r[g][7])
                             : "r"(lane_addr + (uint32_t)(g * NH + c0)));
            double cv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) cv[c] = Cf[tid + (size_t)(h * NH + c0 + c) * M];
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double acc = 0.0;
#pragma unroll
                for (int g = NS - 1; g >= 0; --g)         // smallest weights first
                    acc = fma((double)(int)r[g][c], 1.0 / (double)(1ull << (12 + 7 * g)), acc);
                const int j = h * NH + c0 + c;
                Cf[tid + (size_t)j * M] = cv[c] - acc * rowscale * colscale[j];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                   // planes are free for the next half
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        t_epi += clock64() - t0;
    }
    if (prof && blockIdx.x == 0 && tid == 0) { prof[0] = t_slice; prof[1] = t_mma; prof[2] = t_epi; }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---------------------------------------------------------------------------------------------------------------------
static double urand(uint64_t& st) {
    st = st * 6364136223846793005ULL + 1442695040888963407ULL;
    return ((st >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
}

int main(int argc, char** argv) {
    const int fronts = argc > 1 ? atoi(argv[1]) : 148 * 4;
    const int reps = argc > 2 ? atoi(argv[2]) : 20;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("# device %s, %d SMs; %d fronts per launch, %d timed launches; trailing block %d x %d\n", prop.name,
           prop.multiProcessorCount, fronts, reps, M, N);
    CK(cudaFuncSetAttribute(schur_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(schur_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    int rc = 0;
    for (int K : {32, 64, 96}) {
        std::vector<double> hL((size_t)fronts * M * K), hU((size_t)fronts * K * N), hC((size_t)fronts * M * N);
        uint64_t st = 20261017ULL + K;
        // entries with the spread of a Jacobian front: magnitudes over four decades, mixed signs
        for (auto& v : hL) v = urand(st) * pow(10.0, 2.0 * urand(st));
        for (auto& v : hU) v = urand(st) * pow(10.0, 2.0 * urand(st));
        for (auto& v : hC) v = urand(st) * 100.0;
        // long-double reference of front 0 and of the last front
        auto reference = [&](int f, std::vector<double>& out, std::vector<double>& mag) {
            out.assign((size_t)M * N, 0.0);
            mag.assign((size_t)M * N, 0.0);
            for (int i = 0; i < M; ++i)
                for (int j = 0; j < N; ++j) {
                    long double acc = hC[(size_t)f * M * N + i + (size_t)j * M], m = fabsl(acc);
                    for (int q = 0; q < K; ++q) {
                        const long double p = (long double)hL[(size_t)f * M * K + (size_t)i * K + q] * hU[(size_t)f * K * N + (size_t)q * N + j];
                        acc -= p;
                        m += fabsl(p);
                    }
                    out[i + (size_t)j * M] = (double)acc;
                    mag[i + (size_t)j * M] = (double)m;
                }
        };
        double *dL, *dU, *dC;
        long long* dprof;
        CK(cudaMalloc(&dprof, 4 * sizeof(long long)));
        CK(cudaMalloc(&dL, hL.size() * 8));
        CK(cudaMalloc(&dU, hU.size() * 8));
        CK(cudaMalloc(&dC, hC.size() * 8));
        CK(cudaMemcpy(dL, hL.data(), hL.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dU, hU.data(), hU.size() * 8, cudaMemcpyHostToDevice));
        const size_t smem_a = (size_t)2 * K * LDP * 8;
        const size_t smem_b = (size_t)2 * NS * (K / 32) * 4096 + (2 * N + M) * 8 + 1024;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        double ms[2] = {0, 0}, err[2] = {0, 0};
        for (int variant = 0; variant < 2; ++variant) {
            auto launch = [&]() {
                if (variant == 0) schur_dmma_kernel<<<fronts, 256, smem_a>>>(dL, dU, dC, K);
                else schur_tcgen05_kernel<<<fronts, 128, smem_b>>>(dL, dU, dC, K, dprof);
            };
            CK(cudaMemcpy(dC, hC.data(), hC.size() * 8, cudaMemcpyHostToDevice));
            launch();
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            std::vector<double> got((size_t)M * N), ref, mag;
            for (int f : {0, fronts - 1}) {
                CK(cudaMemcpy(got.data(), dC + (size_t)f * M * N, got.size() * 8, cudaMemcpyDeviceToHost));
                reference(f, ref, mag);
                for (size_t t = 0; t < got.size(); ++t) err[variant] = fmax(err[variant], fabs(got[t] - ref[t]) / mag[t]);
            }
            for (int w = 0; w < 3; ++w) launch();
            CK(cudaEventRecord(e0));
            for (int r = 0; r < reps; ++r) launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float t;
            CK(cudaEventElapsedTime(&t, e0, e1));
            ms[variant] = t / reps;
        }
        const double flop = 2.0 * M * N * K * fronts;
        printf("K=%3d  dmma %8.3f ms (%6.2f TFLOP/s fp64, err %.2e)   tcgen05-ozaki %8.3f ms (%6.2f TFLOP/s fp64-equivalent, err %.2e)   ratio %.2f\n",
               K, ms[0], flop / ms[0] * 1e-9, err[0], ms[1], flop / ms[1] * 1e-9, err[1], ms[1] / ms[0]);
        long long hp[4] = {0, 0, 0, 0};
        CK(cudaMemcpy(hp, dprof, 3 * sizeof(long long), cudaMemcpyDeviceToHost));
        printf("       tcgen05 CTA 0, cycles: scaling + slicing %lld, %d slice products (issue -> commit) %lld, TMEM read-back + FP64 recombination %lld\n",
               hp[0], 2 * (NS * (NS + 1) / 2) * (K / 32), hp[1], hp[2]);
        CK(cudaFree(dprof));
        if (!(err[0] < 1e-14) || !(err[1] < 1e-13)) rc = 1;      // relative to sum |terms|: both are FP64-accurate
        CK(cudaFree(dL));
        CK(cudaFree(dU));
        CK(cudaFree(dC));
    }
    printf(rc ? "FAILED accuracy\n" : "OK\n");
    return rc;
}
