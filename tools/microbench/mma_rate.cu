// Microbenchmark: issue rate of tcgen05.mma kind::tf32 (M=128, K=8) as a function of N, of the number of TMEM
// accumulators the stream rotates over, and of the A-operand source (TMEM vs shared memory).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

#define CG 1
__global__ void __launch_bounds__(832, 1) k_mma_rate(int N, int nacc, int a_tmem, int nmma, int iters, int kind_bf16,
                                                     long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_s = base;                 // 16 KB: A tile [128 rows x 32 tf32], K-major SW128
  const uint32_t b_s = base + 16384;         // 32 KB: B tile [256 rows x 32 tf32]
  const uint32_t bar = base + 16384 + 32768;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0x3f800000u + (i & 0xfff)) : "memory");
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile int*>(raw + 8) = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_ptr;
  if (warp == 0) {
    const uint32_t idesc = idesc_tf32(128, N);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        uint64_t bd[4], ad[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { bd[q] = desc_sw128(b_s + q * 32); ad[q] = desc_sw128(a_s + q * 32); }
        const uint32_t commit_blocks = (uint32_t)(kind_bf16 % 100);
        for (int m = 0; m < nmma; m += 8) {
          const uint32_t d = tb + 64 + (uint32_t)(((m >> 3) % nacc) * N);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (a_tmem) {
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                           "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(tb + (uint32_t)((q & 3) * 8)),
                           "l"(bd[q & 3]), "r"(idesc), "r"(1u) : "memory");
            } else {
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                           "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad[q & 3]), "l"(bd[q & 3]),
                           "r"(idesc), "r"(1u) : "memory");
            }
          }
          if (commit_blocks) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8) : "memory");
        }
        if (false) { for (int m = 0; m < 0; ++m) {
          const uint32_t d = tb + 64 + (uint32_t)((m % nacc) * N);
          const uint64_t bd = desc_sw128(b_s + (m & 3) * 32);
          if ((kind_bf16 % 100) > 0 && m % (kind_bf16 % 100) == (kind_bf16 % 100) - 1)
            asm volatile("tcgen05.commit.cta_group::%1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8), "n"(CG) : "memory");
          if (a_tmem) {
            const uint32_t a = tb + (uint32_t)((m & 3) * 8);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bd),
                         "r"(idesc), "r"(1u)
                         : "memory");
          } else {
            const uint64_t ad = desc_sw128(a_s + (m & 3) * 32);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd),
                         "r"(idesc), "r"(1u)
                         : "memory");
          }
        } }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      }
      __syncwarp();
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
      }
      ph ^= 1u;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    *reinterpret_cast<volatile int*>(raw + 8) = 1;
  } else if (kind_bf16 >= 300) {   // background ALU/FMA issue pressure from the other warps (no memory traffic)
    volatile int* stop = reinterpret_cast<volatile int*>(raw + 8);
    float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f;
    while (*stop == 0) {
#pragma unroll
      for (int rep = 0; rep < 64; ++rep) {
        a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
      }
    }
    if (a0 + a1 + a2 + a3 == 123.456f) out[0] = 1;
  } else if (kind_bf16 >= 200) {   // background shared-memory load traffic (LDS.128, conflict free) from the other warps
    volatile int* stop = reinterpret_cast<volatile int*>(raw + 8);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t off = threadIdx.x * 16;
    while (*stop == 0) {
#pragma unroll
      for (int rep = 0; rep < 16; ++rep) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + ((off + rep * 2048) & 0xbfff)));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (acc.x == 123.456f) out[0] = 1;
  } else if (kind_bf16 >= 100) {
    volatile int* stop = reinterpret_cast<volatile int*>(raw + 8);
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x + i;
    while (*stop == 0) {
      for (int rep = 0; rep < 8; ++rep) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     ::"r"(tb + lane_field + 32 + (rep & 1) * 16), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                       "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

#undef CG
#define CG 2
__global__ void __launch_bounds__(128, 1) k_mma_rate_pair(int N, int nacc, int a_tmem, int nmma, int iters, int kind_bf16,
                                                     long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_s = base;                 // 16 KB: A tile [128 rows x 32 tf32], K-major SW128
  const uint32_t b_s = base + 16384;         // 32 KB: B tile [256 rows x 32 tf32]
  const uint32_t bar = base + 16384 + 32768;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0x3f800000u + (i & 0xfff)) : "memory");
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile int*>(raw + 8) = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_ptr;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  if (warp == 0 && crank == 0) {
    const uint32_t idesc = idesc_tf32(256, N);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        uint64_t bd[4], ad[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { bd[q] = desc_sw128(b_s + q * 32); ad[q] = desc_sw128(a_s + q * 32); }
        const uint32_t commit_blocks = (uint32_t)(kind_bf16 % 100);
        for (int m = 0; m < nmma; m += 8) {
          const uint32_t d = tb + 64 + (uint32_t)(((m >> 3) % nacc) * N);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (a_tmem) {
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                           "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(tb + (uint32_t)((q & 3) * 8)),
                           "l"(bd[q & 3]), "r"(idesc), "r"(1u) : "memory");
            } else {
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                           "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad[q & 3]), "l"(bd[q & 3]),
                           "r"(idesc), "r"(1u) : "memory");
            }
          }
          if (commit_blocks) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar + 8), "h"((uint16_t)1) : "memory");
        }
        if (false) { for (int m = 0; m < 0; ++m) {
          const uint32_t d = tb + 64 + (uint32_t)((m % nacc) * N);
          const uint64_t bd = desc_sw128(b_s + (m & 3) * 32);
          if ((kind_bf16 % 100) > 0 && m % (kind_bf16 % 100) == (kind_bf16 % 100) - 1)
            asm volatile("tcgen05.commit.cta_group::%1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8), "n"(CG) : "memory");
          if (a_tmem) {
            const uint32_t a = tb + (uint32_t)((m & 3) * 8);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bd),
                         "r"(idesc), "r"(1u)
                         : "memory");
          } else {
            const uint64_t ad = desc_sw128(a_s + (m & 3) * 32);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd),
                         "r"(idesc), "r"(1u)
                         : "memory");
          }
        } }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)1) : "memory");
      }
      __syncwarp();
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
      }
      ph ^= 1u;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    *reinterpret_cast<volatile int*>(raw + 8) = 1;
  } else if (kind_bf16 >= 300) {   // background ALU/FMA issue pressure from the other warps (no memory traffic)
    volatile int* stop = reinterpret_cast<volatile int*>(raw + 8);
    float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f;
    while (*stop == 0) {
#pragma unroll
      for (int rep = 0; rep < 64; ++rep) {
        a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
      }
    }
    if (a0 + a1 + a2 + a3 == 123.456f) out[0] = 1;
  } else if (kind_bf16 >= 200) {   // background shared-memory load traffic (LDS.128, conflict free) from the other warps
    volatile int* stop = reinterpret_cast<volatile int*>(raw + 8);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t off = threadIdx.x * 16;
    while (*stop == 0) {
#pragma unroll
      for (int rep = 0; rep < 16; ++rep) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + ((off + rep * 2048) & 0xbfff)));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (acc.x == 123.456f) out[0] = 1;
  } else if (kind_bf16 >= 100) {
    volatile int* stop = reinterpret_cast<volatile int*>(raw + 8);
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x + i;
    while (*stop == 0) {
      for (int rep = 0; rep < 8; ++rep) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     ::"r"(tb + lane_field + 32 + (rep & 1) * 16), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                       "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}


int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  const int smem = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(k_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Cfg { int N, nacc, a_tmem; };
  Cfg cfgs[] = {{64, 1, 1}, {64, 2, 1}, {64, 3, 1}, {64, 6, 1}, {128, 1, 1}, {128, 2, 1}, {128, 3, 1}, {256, 1, 1},
                {64, 1, 0}, {64, 3, 0}, {128, 1, 0}, {128, 3, 0}, {256, 1, 0}};
  const int nmma = 96, iters = 200;
  for (int mode : {0, 300}) {   // commit every 8 MMAs; the same plus TMEM store traffic from 3 other warps
    for (int N : {64, 128}) {
      for (int rep = 0; rep < 2; ++rep) k_mma_rate<<<148, mode >= 200 ? 832 : 128, smem>>>(N, 1, 1, nmma, iters, mode, out);
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0;
      for (int i = 0; i < 148; ++i) s += (double)h[i];
      printf("mode=%d N=%d A=tmem: %7.1f cycles/MMA\n", mode, N, s / 148 / ((double)nmma * iters));
    }
  }
  for (auto c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) k_mma_rate<<<148, 128, smem>>>(c.N, c.nacc, c.a_tmem, nmma, iters, 0, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < 148; ++i) s += (double)h[i];
    const double cyc = s / 148 / ((double)nmma * iters);
    const double fma = 128.0 * c.N * 8;
    printf("N=%3d nacc=%d A=%s : %7.1f cycles/MMA  -> %7.0f FMA/clk/SM\n", c.N, c.nacc, c.a_tmem ? "tmem" : "smem", cyc,
           fma / cyc);
  }
  cudaFuncSetAttribute(k_mma_rate_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (auto c : cfgs) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaMemset(out, 0, 148 * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, k_mma_rate_pair, c.N, c.nacc, c.a_tmem, nmma, iters, 0, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pair error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < 148; i += 2) s += (double)h[i];
    const double cyc = s / 74 / ((double)nmma * iters);
    const double fma = 128.0 * c.N * 8;   // per SM
    printf("PAIR M=256 N=%3d nacc=%d A=%s : %7.1f cycles/MMA  -> %7.0f FMA/clk/SM\n", c.N, c.nacc, c.a_tmem ? "tmem" : "smem",
           cyc, fma / cyc);
  }
  return 0;
}
