// Fused assign + per-cluster partial-sum kernel for sm_100a: ONE pass over X per Lloyd iteration.
//
//   TMA (cp.async.bulk.tensor, 128B swizzle)  : X row tiles [128 x 32 f32] HBM -> smem ring
//   convert warps (4)                          : smem -> split x = hi + lo (both RN to tf32) -> TMEM (tcgen05.st)
//   MMA warp (1 thread issues)                 : D[128 x KP] (TMEM, fp32) = hi.Chi^T + lo.Chi^T + hi.Clo^T
//                                                 tcgen05.mma kind::tf32, A from TMEM, B (centers) from smem
//                                                 => "3xTF32": fp32-accurate x.c without an fp32 tensor mode
//   epilogue warps (4)                         : tcgen05.ld D -> dist_j = ||c_j||^2 - 2 x.c_j -> argmin
//                                                 (lowest index on ties) -> labels, min distance, cost
//   update warps (8)                           : re-read the SAME smem tile, accumulate per-cluster sums in
//                                                 REGISTERS (warp u owns clusters j%8==u; lane owns 4 columns):
//                                                 no atomics, deterministic; flushed once per CTA
//
// Persistent: one CTA per SM, static round-robin over row tiles (deterministic partial sums).
// What it replaces: cuML's fusedL2NN (minClusterAndDistanceCompute) + reduce_rows_by_key second pass over X,
// reached from spark_rapids_ml/clustering.py:412-415.
//
// Algorithmic HBM bytes per launch: 4*n*d (X once) [+ 4*n labels / 4*n mindist when requested]
// + 148 * (k*d + k) * 4 partials (negligible).  See DESIGN.md "Kernels".
#include <float.h>
#include <stdio.h>

#include "b2k_internal.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// configuration
// ------------------------------------------------------------------------------------------------
constexpr int TM = 128;            // rows per tile (UMMA M)
constexpr int CHUNK = 32;          // f32 per 128-byte swizzle row = one TMA box / K chunk
constexpr int SLOT_BYTES = TM * CHUNK * 4;  // 16 KB
constexpr int NA = 4;              // TMEM A-operand ring slots (hi 32 cols + lo 32 cols each)
constexpr int A_COLS = 64;
constexpr int D_OFF = NA * A_COLS; // 256
constexpr int TMEM_COLS = 512;

constexpr int W_CONVERT0 = 0;      // warps 0-3  (lane quadrant = warp % 4)
constexpr int W_EPI0 = 4;          // warps 4-7
constexpr int W_UPD0 = 8;          // warps 8-15
constexpr int W_TMA = 16;
constexpr int W_MMA = 17;
constexpr int NWARPS = 18;
constexpr int NTHREADS = NWARPS * 32;
constexpr int N_UPD = 8;

constexpr size_t SMEM_LIMIT = 227 * 1024;

template <int KP, int DP>
struct Cfg {
  static_assert(KP % 16 == 0 && KP >= 16 && KP <= 128, "KP");
  static_assert(DP % CHUNK == 0 && DP >= CHUNK && DP <= 256, "DP");
  static constexpr int NCH = DP / CHUNK;
  static constexpr int C_BYTES = KP * DP * 4;                  // one of Chi / Clo
  static constexpr int MISC = 1024 /*labels*/ + 1024 /*xnorm*/ + KP * 4 + 512 /*barriers*/ + 64;
  static constexpr int NSLOT_RAW = (int)((SMEM_LIMIT - 1024 - 2 * C_BYTES - MISC) / SLOT_BYTES);
  static constexpr int NSLOT = NSLOT_RAW > 12 ? 12 : NSLOT_RAW;
  static_assert(NSLOT >= NCH + 1, "ring too small");
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_CHI = NSLOT * SLOT_BYTES;
  static constexpr int OFF_CLO = OFF_CHI + C_BYTES;
  static constexpr int OFF_CNORM = OFF_CLO + C_BYTES;
  static constexpr int OFF_LABELS = OFF_CNORM + KP * 4;
  static constexpr int OFF_XNORM = OFF_LABELS + 1024;
  static constexpr int OFF_BARS = OFF_XNORM + 1024;
  // barrier indices (8 bytes each)
  static constexpr int B_XFULL = 0;
  static constexpr int B_XEMPTY = B_XFULL + NSLOT;
  static constexpr int B_AFULL = B_XEMPTY + NSLOT;
  static constexpr int B_AEMPTY = B_AFULL + NA;
  static constexpr int B_DFULL = B_AEMPTY + NA;
  static constexpr int B_DEMPTY = B_DFULL + 2;
  static constexpr int B_LFULL = B_DEMPTY + 2;
  static constexpr int B_LEMPTY = B_LFULL + 2;
  static constexpr int B_NFULL = B_LEMPTY + 2;
  static constexpr int B_NEMPTY = B_NFULL + 2;
  static constexpr int B_CFULL = B_NEMPTY + 2;
  static constexpr int NBARS = B_CFULL + 1;
  static_assert(NBARS * 8 <= 512, "barrier area");
  static constexpr int OFF_TMEMPTR = OFF_BARS + 512;
  static constexpr int SMEM_BYTES = OFF_TMEMPTR + 64 + 1024;  // +1024: manual 1 KB alignment slack
  static_assert(SMEM_BYTES <= (int)SMEM_LIMIT, "smem");
  static constexpr int UPL = (DP / 4 + 31) / 32;   // float4 units per lane in the update warps
  static constexpr int CPW = (KP + N_UPD - 1) / N_UPD;  // clusters per update warp
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug traps (sticky launch failure the host reports) instead of hanging the GPU.
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("b2k fused: mbarrier timeout block %d warp %d bar_off %u parity %u\n", blockIdx.x, threadIdx.x >> 5, bar,
         parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) mbar_timeout(bar, parity);
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]^T, kind::tf32
__device__ __forceinline__ void tc_mma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// round-to-nearest (ties away) fp32 -> tf32 (10 explicit mantissa bits), result has the low 13 bits clear
__device__ __forceinline__ uint32_t rn_tf32_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused with swizzle), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                      // layout type: SWIZZLE_128B
  return d;
}
// UMMA instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=KP
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// prep: padded hi/lo split of the centers + ||c||^2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_prep_centers_tc(const float* __restrict__ C, int k, int d, int KP, int DP,
                                                         float* __restrict__ Chi, float* __restrict__ Clo,
                                                         float* __restrict__ cnorm, const B2kLoopState* st) {
  if (st != nullptr && st->done) return;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= KP) return;
  double s = 0.0;
  for (int t = lane; t < DP; t += 32) {
    float v = (warp < k && t < d) ? C[(size_t)warp * d + t] : 0.f;
    uint32_t hb = rn_tf32_bits(v);
    float hi = __uint_as_float(hb);
    float lo = v - hi;
    Chi[(size_t)warp * DP + t] = hi;
    Clo[(size_t)warp * DP + t] = __uint_as_float(rn_tf32_bits(lo));
    s += (double)v * (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) cnorm[warp] = warp < k ? (float)s : __int_as_float(0x7f800000);
}

// ------------------------------------------------------------------------------------------------
// the fused kernel
// ------------------------------------------------------------------------------------------------
struct FusedArgs {
  int64_t n;
  int ntiles;
  int k;
  int d;
  const float* cnorm;      // [KP]
  float* partials;         // [grid][k*d]
  int32_t* counts;         // [grid][k]
  double* cost_partials;   // [grid]
  int32_t* labels_out;     // [n] or NULL
  float* mind_out;         // [n] or NULL
  int do_update;
  const B2kLoopState* st;
};

template <int KP, int DP>
__global__ void __launch_bounds__(NTHREADS, 1)
k_fused_assign_update(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapChi,
                      const __grid_constant__ CUtensorMap mapClo, const FusedArgs args) {
  using G = Cfg<KP, DP>;
  if (args.st != nullptr && args.st->done) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ring = base + G::OFF_RING;
  const uint32_t chi_s = base + G::OFF_CHI;
  const uint32_t clo_s = base + G::OFF_CLO;
  float* cnorm_s = reinterpret_cast<float*>(gbase + G::OFF_CNORM);
  int* labels_s = reinterpret_cast<int*>(gbase + G::OFF_LABELS);     // [2][128]
  float* xnorm_s = reinterpret_cast<float*>(gbase + G::OFF_XNORM);   // [2][128]
  const uint32_t bars = base + G::OFF_BARS;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(gbase + G::OFF_TMEMPTR);
  auto bar = [&](int i) -> uint32_t { return bars + 8u * (uint32_t)i; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup ----
  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapChi);
    tma_prefetch_desc(&mapClo);
    for (int i = 0; i < G::NSLOT; ++i) {
      mbar_init(bar(G::B_XFULL + i), 1);
      mbar_init(bar(G::B_XEMPTY + i), 4 + N_UPD);
    }
    for (int i = 0; i < NA; ++i) {
      mbar_init(bar(G::B_AFULL + i), 4);
      mbar_init(bar(G::B_AEMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(G::B_DFULL + i), 1);
      mbar_init(bar(G::B_DEMPTY + i), 4);
      mbar_init(bar(G::B_LFULL + i), 4);
      mbar_init(bar(G::B_LEMPTY + i), N_UPD);
      mbar_init(bar(G::B_NFULL + i), 4);
      mbar_init(bar(G::B_NEMPTY + i), 4);
    }
    mbar_init(bar(G::B_CFULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int j = threadIdx.x; j < KP; j += NTHREADS) cnorm_s[j] = args.cnorm[j];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const int tile0 = blockIdx.x;
  const int tstep = gridDim.x;

  if (warp == W_TMA) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      mbar_expect_tx(bar(G::B_CFULL), 2u * G::C_BYTES);
      for (int c = 0; c < G::NCH; ++c) {
        tma_load_2d(chi_s + c * (KP * 128), &mapChi, bar(G::B_CFULL), c * CHUNK, 0);
        tma_load_2d(clo_s + c * (KP * 128), &mapClo, bar(G::B_CFULL), c * CHUNK, 0);
      }
      int xs = 0;
      uint32_t xph = 0;
      for (int tile = tile0; tile < args.ntiles; tile += tstep) {
#pragma unroll 1
        for (int c = 0; c < G::NCH; ++c) {
          mbar_wait(bar(G::B_XEMPTY + xs), xph ^ 1u);
          mbar_expect_tx(bar(G::B_XFULL + xs), SLOT_BYTES);
          tma_load_2d(ring + xs * SLOT_BYTES, &mapX, bar(G::B_XFULL + xs), c * CHUNK, tile * TM);
          if (++xs == G::NSLOT) { xs = 0; xph ^= 1u; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ======================= MMA issuer =======================
    constexpr uint32_t idesc = make_idesc_tf32(TM, KP);
    mbar_wait(bar(G::B_CFULL), 0);
    int as = 0;
    uint32_t aph = 0;
    int ti = 0;
    for (int tile = tile0; tile < args.ntiles; tile += tstep, ++ti) {
      const int b = ti & 1;
      const uint32_t bph = (uint32_t)(ti >> 1) & 1u;
      mbar_wait(bar(G::B_DEMPTY + b), bph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + D_OFF + b * KP;
#pragma unroll 1
      for (int c = 0; c < G::NCH; ++c) {
        mbar_wait(bar(G::B_AFULL + as), aph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = tmem_base + as * A_COLS;
          const uint32_t a_lo = a_hi + CHUNK;
          const uint32_t bhi = chi_s + c * (KP * 128);
          const uint32_t blo = clo_s + c * (KP * 128);
#pragma unroll
          for (int ks = 0; ks < CHUNK / 8; ++ks) {
            const uint64_t dhi = make_kmajor_sw128_desc(bhi + ks * 32);
            const uint64_t dlo = make_kmajor_sw128_desc(blo + ks * 32);
            tc_mma_ts_tf32(d_tmem, a_lo + ks * 8, dhi, idesc, (c | ks) != 0 ? 1u : 0u);  // small terms first
            tc_mma_ts_tf32(d_tmem, a_hi + ks * 8, dlo, idesc, 1u);
            tc_mma_ts_tf32(d_tmem, a_hi + ks * 8, dhi, idesc, 1u);
          }
          tc_commit(bar(G::B_AEMPTY + as));
          if (c == G::NCH - 1) tc_commit(bar(G::B_DFULL + b));
        }
        __syncwarp();
        if (++as == NA) { as = 0; aph ^= 1u; }
      }
    }
  } else if (warp < W_EPI0) {
    // ======================= convert warps: smem -> (hi, lo) -> TMEM =======================
    const int q = warp - W_CONVERT0;
    const int r = q * 32 + lane;                       // row within the tile == TMEM lane
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const uint32_t swz = (uint32_t)(r & 7);
    int xs = 0, as = 0;
    uint32_t xph = 0, aph = 0;
    int ti = 0;
    for (int tile = tile0; tile < args.ntiles; tile += tstep, ++ti) {
      const int b = ti & 1;
      const uint32_t bph = (uint32_t)(ti >> 1) & 1u;
      float xn = 0.f;
#pragma unroll 1
      for (int c = 0; c < G::NCH; ++c) {
        mbar_wait(bar(G::B_XFULL + xs), xph);
        mbar_wait(bar(G::B_AEMPTY + as), aph ^ 1u);
        tc_fence_after();
        const uint32_t rowaddr = ring + xs * SLOT_BYTES + (uint32_t)r * 128u;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = lds128(rowaddr + (((uint32_t)j ^ swz) << 4));
          float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            uint32_t hb = rn_tf32_bits(e[t]);
            float l = e[t] - __uint_as_float(hb);
            hi[j * 4 + t] = hb;
            lo[j * 4 + t] = rn_tf32_bits(l);
            xn = fmaf(e[t], e[t], xn);
          }
        }
        const uint32_t a_addr = tmem_base + lane_field + (uint32_t)(as * A_COLS);
        tmem_st_x32(a_addr, hi);
        tmem_st_x32(a_addr + CHUNK, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(G::B_AFULL + as));
          mbar_arrive(bar(G::B_XEMPTY + xs));
        }
        if (++xs == G::NSLOT) { xs = 0; xph ^= 1u; }
        if (++as == NA) { as = 0; aph ^= 1u; }
      }
      mbar_wait(bar(G::B_NEMPTY + b), bph ^ 1u);
      xnorm_s[b * TM + r] = xn;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(G::B_NFULL + b));
    }
  } else if (warp < W_UPD0) {
    // ======================= epilogue warps: TMEM D -> argmin =======================
    const int q = warp - W_EPI0;
    const int r = q * 32 + lane;
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    double cost = 0.0;
    int ti = 0;
    for (int tile = tile0; tile < args.ntiles; tile += tstep, ++ti) {
      const int b = ti & 1;
      const uint32_t bph = (uint32_t)(ti >> 1) & 1u;
      mbar_wait(bar(G::B_DFULL + b), bph);
      tc_fence_after();
      float best = __int_as_float(0x7f800000);
      int bj = 0;
#pragma unroll
      for (int g = 0; g < KP / 32; ++g) {
        uint32_t v[32];
        tmem_ld_x32(tmem_base + lane_field + (uint32_t)(D_OFF + b * KP + g * 32), v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float dist = fmaf(-2.f, __uint_as_float(v[j]), cnorm_s[g * 32 + j]);
          if (dist < best) { best = dist; bj = g * 32 + j; }
        }
      }
      if constexpr (KP % 32 != 0) {
        // tail group of 16 columns
        uint32_t v[32];
        constexpr int g = KP / 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(tmem_base + lane_field + (uint32_t)(D_OFF + b * KP + g * 32))
            : "memory");
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float dist = fmaf(-2.f, __uint_as_float(v[j]), cnorm_s[g * 32 + j]);
          if (dist < best) { best = dist; bj = g * 32 + j; }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(G::B_DEMPTY + b));   // D buffer may be overwritten by tile ti+2

      mbar_wait(bar(G::B_NFULL + b), bph);
      const float xn = xnorm_s[b * TM + r];
      const int64_t grow = (int64_t)tile * TM + r;
      const bool valid = grow < args.n;
      const float md = fmaxf(xn + best, 0.f);
      mbar_wait(bar(G::B_LEMPTY + b), bph ^ 1u);
      labels_s[b * TM + r] = valid ? bj : -1;
      if (valid) {
        if (args.labels_out) args.labels_out[grow] = bj;
        if (args.mind_out) args.mind_out[grow] = md;
        cost += (double)md;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(G::B_LFULL + b));
        mbar_arrive(bar(G::B_NEMPTY + b));
      }
    }
    // per-CTA cost: fixed-order fold (lanes, then the 4 warps through shared memory after the final sync)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
    if (lane == 0) reinterpret_cast<double*>(gbase + G::OFF_TMEMPTR + 16)[q] = cost;
  } else if (warp < W_TMA) {
    // ======================= update warps: per-cluster sums in registers =======================
    const int u = warp - W_UPD0;
    float4 acc[G::CPW][G::UPL];
    int cnt[G::CPW];
#pragma unroll
    for (int c = 0; c < G::CPW; ++c) {
      cnt[c] = 0;
#pragma unroll
      for (int i = 0; i < G::UPL; ++i) acc[c][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int xs = 0;
    uint32_t xph = 0;
    int ti = 0;
    for (int tile = tile0; tile < args.ntiles; tile += tstep, ++ti) {
      const int b = ti & 1;
      const uint32_t bph = (uint32_t)(ti >> 1) & 1u;
      mbar_wait(bar(G::B_LFULL + b), bph);
      // slots of this tile's chunks (observe their full barriers: TMA writes come from the async proxy)
      uint32_t slot_addr[G::NCH];
      {
        int s2 = xs;
        uint32_t p2 = xph;
#pragma unroll
        for (int c = 0; c < G::NCH; ++c) {
          mbar_wait(bar(G::B_XFULL + s2), p2);
          slot_addr[c] = ring + s2 * SLOT_BYTES;
          if (++s2 == G::NSLOT) { s2 = 0; p2 ^= 1u; }
        }
      }
      if (args.do_update) {
        int lab[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) lab[g] = labels_s[b * TM + g * 32 + lane];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t m = __ballot_sync(0xffffffffu, lab[g] >= 0 && (lab[g] & (N_UPD - 1)) == u);
          while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int l = __shfl_sync(0xffffffffu, lab[g], bit);
            const int row = g * 32 + bit;
            const int cl = l >> 3;
            float4 v[G::UPL];
#pragma unroll
            for (int i = 0; i < G::UPL; ++i) {
              const int unit = lane + 32 * i;          // float4 index within the row
              if (unit < DP / 4) {
                const int cc = unit >> 3, j = unit & 7;
                // static chunk index for register-resident slot_addr[]: select with a small switch
                uint32_t sa = slot_addr[0];
#pragma unroll
                for (int c2 = 1; c2 < G::NCH; ++c2) sa = (cc == c2) ? slot_addr[c2] : sa;
                v[i] = lds128(sa + (uint32_t)row * 128u + (((uint32_t)j ^ (uint32_t)(row & 7)) << 4));
              } else {
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
            // warp-uniform dispatch on the owned-cluster slot (keeps accumulators in registers)
#define B2K_ACC_CASE(c)                                                         \
  case c:                                                                       \
    if constexpr (c < G::CPW) {                                                 \
      _Pragma("unroll") for (int i = 0; i < G::UPL; ++i) {                      \
        acc[c < G::CPW ? c : 0][i].x += v[i].x;                                 \
        acc[c < G::CPW ? c : 0][i].y += v[i].y;                                 \
        acc[c < G::CPW ? c : 0][i].z += v[i].z;                                 \
        acc[c < G::CPW ? c : 0][i].w += v[i].w;                                 \
      }                                                                         \
      cnt[c < G::CPW ? c : 0]++;                                                \
    }                                                                           \
    break;
            switch (cl) {
              B2K_ACC_CASE(0) B2K_ACC_CASE(1) B2K_ACC_CASE(2) B2K_ACC_CASE(3)
              B2K_ACC_CASE(4) B2K_ACC_CASE(5) B2K_ACC_CASE(6) B2K_ACC_CASE(7)
              B2K_ACC_CASE(8) B2K_ACC_CASE(9) B2K_ACC_CASE(10) B2K_ACC_CASE(11)
              B2K_ACC_CASE(12) B2K_ACC_CASE(13) B2K_ACC_CASE(14) B2K_ACC_CASE(15)
              default: break;
            }
#undef B2K_ACC_CASE
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(G::B_LEMPTY + b));
        int s2 = xs;
#pragma unroll
        for (int c = 0; c < G::NCH; ++c) {
          mbar_arrive(bar(G::B_XEMPTY + s2));
          if (++s2 == G::NSLOT) s2 = 0;
        }
      }
#pragma unroll
      for (int c = 0; c < G::NCH; ++c)
        if (++xs == G::NSLOT) { xs = 0; xph ^= 1u; }
    }
    // flush: partials[cta][l][col..col+3], l = u + 8*c
    if (args.do_update) {
      float* out = args.partials + (size_t)blockIdx.x * args.k * args.d;
#pragma unroll
      for (int c = 0; c < G::CPW; ++c) {
        const int l = u + N_UPD * c;
        if (l < args.k) {
#pragma unroll
          for (int i = 0; i < G::UPL; ++i) {
            const int col = (lane + 32 * i) * 4;
            float e[4] = {acc[c][i].x, acc[c][i].y, acc[c][i].z, acc[c][i].w};
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (col + t < args.d) out[(size_t)l * args.d + col + t] = e[t];
          }
          if (lane == 0) args.counts[(size_t)blockIdx.x * args.k + l] = cnt[c];
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    const double* cw = reinterpret_cast<const double*>(gbase + G::OFF_TMEMPTR + 16);
    args.cost_partials[blockIdx.x] = ((cw[0] + cw[1]) + cw[2]) + cw[3];
  }
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encoder(b2k_ctx* ctx, EncodeTiledFn* fn) {
  if (!ctx->encode_tiled) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p)
      return b2k_fail(ctx, B2K_ERR_CUDA, "cannot resolve cuTensorMapEncodeTiled from the driver");
    ctx->encode_tiled = p;
  }
  *fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  return B2K_OK;
}

int encode_2d(b2k_ctx* ctx, CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
              uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapL2promotion l2) {
  EncodeTiledFn fn;
  B2K_TRY(get_encoder(ctx, &fn));
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return b2k_fail(ctx, B2K_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return B2K_OK;
}

struct Inst {
  int KP, DP;
};
// instantiations compiled into the library
constexpr Inst kInst[] = {{64, 128}, {64, 64}, {64, 32}, {32, 128}, {128, 128}, {128, 64}, {16, 32}, {16, 64}, {32, 64}, {32, 32}, {16, 128}};

bool pick_inst(int d, int k, Inst* out) {
  int DP = (d + CHUNK - 1) / CHUNK * CHUNK;
  if (DP == 96) DP = 128;
  int best = -1;
  for (size_t i = 0; i < sizeof(kInst) / sizeof(kInst[0]); ++i) {
    if (kInst[i].DP == DP && kInst[i].KP >= k) {
      if (best < 0 || kInst[i].KP < kInst[best].KP) best = (int)i;
    }
  }
  if (best < 0) return false;
  *out = kInst[best];
  return true;
}

template <int KP, int DP>
int launch_inst(b2k_ctx* ctx, int grid, const CUtensorMap& mx, const CUtensorMap& mh, const CUtensorMap& ml,
                const FusedArgs& a, cudaStream_t s) {
  using G = Cfg<KP, DP>;
  auto kern = k_fused_assign_update<KP, DP>;
  B2K_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
  kern<<<grid, NTHREADS, G::SMEM_BYTES, s>>>(mx, mh, ml, a);
  B2K_CUDA_OK(ctx, cudaGetLastError());
  return B2K_OK;
}

struct PlanLayout {
  size_t off_chi, off_clo, off_cnorm, off_partials, off_counts, off_cost, total;
};
PlanLayout plan_layout(const B2kFusedPlan& p, int k, int d) {
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  PlanLayout L{};
  size_t o = 0;
  L.off_chi = o; o = al(o + (size_t)p.KP * p.DP * 4);
  L.off_clo = o; o = al(o + (size_t)p.KP * p.DP * 4);
  L.off_cnorm = o; o = al(o + (size_t)p.KP * 4);
  L.off_partials = o; o = al(o + (size_t)p.grid * k * d * 4);
  L.off_counts = o; o = al(o + (size_t)p.grid * k * 4);
  L.off_cost = o; o = al(o + (size_t)p.grid * 8);
  L.total = o;
  return L;
}
}  // namespace

bool b2k_fused_supported(const b2k_ctx* ctx, int64_t n, int d, int k, const float* X) {
  (void)ctx;
  if (n < 1 || n > (int64_t)0x7fffff00 * 1LL) return false;
  if (d % 4 != 0) return false;                                   // TMA: row pitch must be a multiple of 16 B
  if ((reinterpret_cast<uintptr_t>(X) & 15u) != 0) return false;  // TMA: 16 B aligned base
  Inst in;
  return pick_inst(d, k, &in);
}

int b2k_fused_plan(b2k_ctx* ctx, int64_t n, int d, int k, B2kFusedPlan* plan) {
  Inst in;
  if (!pick_inst(d, k, &in)) return b2k_fail(ctx, B2K_ERR_UNSUPPORTED, "fused kernel: no instantiation for this (k, d)");
  plan->KP = in.KP;
  plan->DP = in.DP;
  int64_t ntiles = (n + TM - 1) / TM;
  int grid = ctx->sm_count;
  if (ctx->grid_limit > 0 && ctx->grid_limit < grid) grid = ctx->grid_limit;
  if (ntiles < grid) grid = (int)ntiles;
  if (grid < 1) grid = 1;
  plan->grid = grid;
  plan->scratch_bytes = plan_layout(*plan, k, d).total;
  return B2K_OK;
}

void b2k_fused_views(const B2kFusedPlan& plan, void* plan_scratch, int k, int d, float** partials, int32_t** counts,
                     double** cost_partials) {
  PlanLayout L = plan_layout(plan, k, d);
  char* b = static_cast<char*>(plan_scratch);
  *partials = reinterpret_cast<float*>(b + L.off_partials);
  *counts = reinterpret_cast<int32_t*>(b + L.off_counts);
  *cost_partials = reinterpret_cast<double*>(b + L.off_cost);
}

int b2k_launch_fused(b2k_ctx* ctx, const B2kFusedPlan& plan, void* plan_scratch, const float* X, int64_t n, int d,
                     const float* C, int k, int32_t* labels_out, float* mindist_out, bool do_update,
                     const B2kLoopState* st, cudaStream_t s) {
  PlanLayout L = plan_layout(plan, k, d);
  char* b = static_cast<char*>(plan_scratch);
  float* Chi = reinterpret_cast<float*>(b + L.off_chi);
  float* Clo = reinterpret_cast<float*>(b + L.off_clo);
  float* cnorm = reinterpret_cast<float*>(b + L.off_cnorm);

  k_prep_centers_tc<<<(plan.KP * 32 + 255) / 256, 256, 0, s>>>(C, k, d, plan.KP, plan.DP, Chi, Clo, cnorm, st);
  ctx->stats.kernel_launches++;
  B2K_CUDA_OK(ctx, cudaGetLastError());

  CUtensorMap mx, mh, ml;
  B2K_TRY(encode_2d(ctx, &mx, X, (uint64_t)d, (uint64_t)n, (uint64_t)d * 4, CHUNK, TM,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
  B2K_TRY(encode_2d(ctx, &mh, Chi, (uint64_t)plan.DP, (uint64_t)plan.KP, (uint64_t)plan.DP * 4, CHUNK,
                    (uint32_t)plan.KP, CU_TENSOR_MAP_L2_PROMOTION_L2_128B));
  B2K_TRY(encode_2d(ctx, &ml, Clo, (uint64_t)plan.DP, (uint64_t)plan.KP, (uint64_t)plan.DP * 4, CHUNK,
                    (uint32_t)plan.KP, CU_TENSOR_MAP_L2_PROMOTION_L2_128B));

  FusedArgs a{};
  a.n = n;
  a.ntiles = (int)((n + TM - 1) / TM);
  a.k = k;
  a.d = d;
  a.cnorm = cnorm;
  a.partials = reinterpret_cast<float*>(b + L.off_partials);
  a.counts = reinterpret_cast<int32_t*>(b + L.off_counts);
  a.cost_partials = reinterpret_cast<double*>(b + L.off_cost);
  a.labels_out = labels_out;
  a.mind_out = mindist_out;
  a.do_update = do_update ? 1 : 0;
  a.st = st;

  int rc = B2K_ERR_UNSUPPORTED;
#define B2K_DISPATCH(KP_, DP_) \
  if (plan.KP == KP_ && plan.DP == DP_) rc = launch_inst<KP_, DP_>(ctx, plan.grid, mx, mh, ml, a, s);
  B2K_DISPATCH(64, 128)
  B2K_DISPATCH(64, 64)
  B2K_DISPATCH(64, 32)
  B2K_DISPATCH(32, 128)
  B2K_DISPATCH(128, 128)
  B2K_DISPATCH(128, 64)
  B2K_DISPATCH(16, 32)
  B2K_DISPATCH(16, 64)
  B2K_DISPATCH(32, 64)
  B2K_DISPATCH(32, 32)
  B2K_DISPATCH(16, 128)
#undef B2K_DISPATCH
  if (rc == B2K_ERR_UNSUPPORTED) return b2k_fail(ctx, rc, "fused kernel: instantiation missing");
  B2K_TRY(rc);
  ctx->stats.kernel_launches++;
  ctx->stats.fused_tc_launches++;
  return B2K_OK;
}
