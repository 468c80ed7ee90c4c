// Fused assign + per-cluster partial-sum kernel for LARGE (k, d) on sm_100a — k <= 256, d <= 256 (BASELINE cfg3:
// k = 256, d = 256) — ONE pass over X per Lloyd iteration.  "T" = transposed operand roles with respect to
// b2k_fused_tc.cu: at k = d = 256 neither the centres (256 KB) nor per-CTA [k, d] accumulators (256 KB) fit beside
// a row tile in one SM, and 3xTF32 would be tensor-bound at 2.4x the HBM time.  So:
//
//   * CTA PAIR (cluster of 2, tcgen05 cta_group::2, UMMA M = 256, N = 128, K = 8):
//       A  = the CENTRES, tf32 (round-to-nearest), resident in TMEM for the whole launch: CTA r holds centres
//            [128 r, 128 r + 128) as 128 lanes x DP columns                                   (TMEM cols [0, 256))
//       B  = the X row tile, fed STRAIGHT from the TMA ring (raw fp32 words; the tensor core truncates them to
//            tf32): each CTA supplies 64 of the step's 128 rows — no convert stage, no operand copy in TMEM
//       D  = [cluster, row] partial products, fp32, double buffered                          (TMEM cols [256, 512))
//   * 1xTF32 screening + exact recheck: the epilogue turns D into dist' = ||c||^2 - 2 x~.c~, packs (dist', cluster)
//     into one ordered 32-bit key, and a shuffle butterfly finds the smallest AND second smallest key of every row
//     across the 32 lanes of a warp; 4 warps x 2 CTAs exchange their partials through (distributed) shared memory.
//     A row whose gap is below the PROVEN bound thr = 2E (E: worst-case error of one dist', see k_tables_t) is
//     re-decided exactly: every cluster within thr of the best is a candidate, the row's owner evaluates the
//     candidates in fp32 (FMA chains over the row in shared memory against the fp32 centres, fixed-order tree) with
//     strict '<' in ascending cluster order (= lowest index on ties).  Rows outside the bound provably have the
//     same argmin in exact arithmetic, so labels match the 3xTF32 / fp32 path.
//   * update: per-cluster sums live in REGISTERS, one [256, 256] accumulator set per CTA PAIR (each CTA: 128
//     clusters x 256 columns = 64 registers per update thread).  Rows are counting-sorted by (owner warp, cluster,
//     row) as in b2k_fused_tc.cu; a warp reads its rows from the local ring or, for the peer's rows, through
//     DSMEM (ld.shared::cluster).  No atomics; static schedule; deterministic.
//   * assign / inertia passes (UPD = false): the same screening + recheck for the labels; the "update" warps
//     compute the exact min distance sum (x - c)^2 of every row from the tile in shared memory.
//
// Replaces (for these shapes) cuML's fusedL2NN + reduce_rows_by_key reached from
// spark_rapids_ml/clustering.py:412-415.  Algorithmic HBM bytes per launch: 4*n*d (X once)
// + 4*n (row norms) [+ 4*n labels / 4*n mindist when requested] + 74 * (k*d + k) * 4 partials.
#include <float.h>
#include <stdio.h>

#include "b2k_internal.cuh"

namespace {

#ifndef B2K_MMA_WAIT
#define B2K_MMA_WAIT mbar_wait_cluster
#endif
#include "b2k_ptx.cuh"

constexpr int TN = 128;                      // X rows per step = UMMA N (64 per CTA of the pair)
constexpr int TNH = 64;
constexpr int CHUNK = 32;                    // f32 per 128-byte swizzle row = one TMA box / 4 UMMA K steps
constexpr int SLOT_BYTES = TNH * CHUNK * 4;  // 8 KB: [64 rows x 32 f32], 128B swizzle = K-major SW128 UMMA B operand
constexpr int NSLOT = 24;                    // ring: 192 KB per CTA
constexpr int KH = 128;                      // centres per CTA (TMEM lanes)
constexpr int D_OFF = 256;                   // TMEM column of D buffer 0 (A occupies [0, DP))
constexpr int TMEM_COLS = 512;

constexpr int W_EPI0 = 0;                    // warps 0-3: epilogue (TMEM lane quadrant = warp % 4)
constexpr int W_UPD0 = 4;                    // warps 4-19: update
constexpr int N_UPD = 16;
constexpr int CPW = KH / N_UPD;              // 8 clusters per update warp
constexpr int W_TMA = W_UPD0 + N_UPD;        // 20
constexpr int W_MMA = W_TMA + 1;             // 21
constexpr int NWARPS = 24;                   // 768 threads launched at 80 registers (22 working warps + 2 that pad the last
                                             // warpgroup).  Lloyd pass at d > 128: the TMA/MMA warpgroup drops to 40
                                             // registers (setmaxnreg.dec) and the 16 update warps raise themselves to 88
                                             // from that pool (64 accumulator registers each): 4x32x80 + 4x32x40 + 16x32x88
constexpr int NTHREADS = NWARPS * 32;

// shared memory layout (dynamic, 1 KB aligned base)
constexpr int OFF_RING = 0;
constexpr int OFF_PART = NSLOT * SLOT_BYTES;            // uint2 part[2][8 sources][128 columns]: (best, second) keys
constexpr int OFF_LAB = OFF_PART + 2 * 8 * TN * 8;      // int32 lab[2][128]: final labels of the step
constexpr int OFF_CAND = OFF_LAB + 2 * TN * 4;          // u32 cand[128][8]: candidate bit masks of flagged rows
constexpr int OFF_CANDT = OFF_CAND + TN * 8 * 4;        // f32 candT[128]: best + thr per column
constexpr int OFF_FIN = OFF_CANDT + TN * 4;             // int32 fin[128]: rechecked labels
constexpr int OFF_SORT = OFF_FIN + TN * 4;              // counting-sort scratch, see SortT
struct SortT {
  static constexpr int CNT = 0;                         // u8 [2][4][128] per-warp key histograms
  static constexpr int ROWS = CNT + 2 * 4 * KH;         // u16 [2][128] sorted row entries
  static constexpr int START = ROWS + 2 * TN * 2;       // u8 [2][144] exclusive start per key (+ total)
  static constexpr int KEYTAB = START + 2 * 144;        // u8 [256] cluster -> key (cta*128 + warp*8 + slot)
  static constexpr int KEYINV = KEYTAB + 256;           // u8 [256] key -> cluster
  static constexpr int BYTES = KEYINV + 256;
};
constexpr int OFF_MISC = (OFF_SORT + SortT::BYTES + 15) & ~15;   // flag words [4] u32, tmem ptr, cost doubles [16]
constexpr int MISC_FLAGW = 0, MISC_TMEMPTR = 16, MISC_COST = 32;
constexpr int OFF_BARS = OFF_MISC + 32 + 16 * 8;
constexpr int B_XFULL = 0;                   // [NSLOT] leader CTA only: both CTAs' TMA boxes of a chunk landed
constexpr int B_SFREE = B_XFULL + NSLOT;     // [8] step slot may be overwritten (local + remote update roles done)
constexpr int B_DFULL = B_SFREE + 8;         // [2]
constexpr int B_DEMPTY = B_DFULL + 2;        // [2] leader CTA only, count 2
constexpr int B_EX = B_DEMPTY + 2;           // [1] epilogue exchange (count 2: one arrival per CTA)
constexpr int B_LFULL = B_EX + 1;            // [2]
constexpr int B_LEMPTY = B_LFULL + 2;        // [2]
constexpr int NBARS = B_LEMPTY + 2;
constexpr int SMEM_BYTES = OFF_BARS + NBARS * 8;
static_assert(SMEM_BYTES <= 227 * 1024, "smem");

// ---- extra PTX for this kernel ----
// TMA load whose completion is signalled on the LEADER CTA's mbarrier (cta_group::2; CUTLASS: SM100_TMA_2SM_LOAD_2D)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t bar_cluster) {   // bar_cluster: mapa() address
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr_cluster, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr_cluster), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr_cluster, uint32_t a) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr_cluster), "r"(a) : "memory");
}
__device__ __forceinline__ void ld_cluster_2(uint32_t addr_cluster, uint64_t& a, uint64_t& b) {
  asm volatile("ld.shared::cluster.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr_cluster));
}
__device__ __forceinline__ uint32_t tmem_ld_x1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}

// Bounded wait without a call: a function call inside the update role's setmaxnreg region makes ptxas give up on the
// region's larger register budget (measured: 76 bytes of accumulator spills with mbar_wait's noinline time-out report).
__device__ __forceinline__ void mbar_wait_nocall(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins == (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_cluster_nocall(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins == (1u << 22)) __trap();
  }
}

// (dist, cluster) -> one ordered key: float bits mapped to an unsigned integer with the same order, low 8 bits replaced
// by the cluster index.  Costs 2^-15 relative resolution of dist (part of the proven bound, k_tables_t).
__device__ __forceinline__ uint32_t pack_key(float dist, uint32_t j) {
  const uint32_t u = __float_as_uint(dist);
  const uint32_t s = (uint32_t)((int32_t)u >> 31) | 0x80000000u;
  return ((u ^ s) & 0xffffff00u) | j;
}
__device__ __forceinline__ float unpack_key(uint32_t key) {
  const uint32_t u = key & 0xffffff00u;
  return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}
// merge two (smallest, second smallest) pairs
__device__ __forceinline__ void merge2(uint32_t& a1, uint32_t& a2, uint32_t b1, uint32_t b2) {
  const uint32_t lo = min(a1, b1), hi = max(a1, b1);
  a2 = min(hi, min(a2, b2));
  a1 = lo;
}

// Transposing butterfly: on entry every lane holds the keys of ITS cluster for 32 columns; on exit lane c holds the
// smallest and second smallest key of column c over the 32 clusters of the warp.
template <int N>
__device__ __forceinline__ void bfly_level(uint32_t (&m1)[16], uint32_t (&m2)[16], int lane) {
  const bool hi = (lane & N) != 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const uint32_t s1 = hi ? m1[i] : m1[i + N], s2 = hi ? m2[i] : m2[i + N];
    uint32_t k1 = hi ? m1[i + N] : m1[i], k2 = hi ? m2[i + N] : m2[i];
    const uint32_t r1 = __shfl_xor_sync(0xffffffffu, s1, N), r2 = __shfl_xor_sync(0xffffffffu, s2, N);
    merge2(k1, k2, r1, r2);
    m1[i] = k1;
    m2[i] = k2;
  }
}
__device__ __forceinline__ void bfly32(const uint32_t (&v)[32], int lane, uint32_t& o1, uint32_t& o2) {
  uint32_t m1[16], m2[16];
  const bool hi = (lane & 16) != 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t send = hi ? v[i] : v[i + 16];
    const uint32_t keep = hi ? v[i + 16] : v[i];
    const uint32_t r = __shfl_xor_sync(0xffffffffu, send, 16);
    m1[i] = min(keep, r);
    m2[i] = max(keep, r);
  }
  bfly_level<8>(m1, m2, lane);
  bfly_level<4>(m1, m2, lane);
  bfly_level<2>(m1, m2, lane);
  bfly_level<1>(m1, m2, lane);
  o1 = m1[0];
  o2 = m2[0];
}

// ------------------------------------------------------------------------------------------------
// prep kernels
// ------------------------------------------------------------------------------------------------
// Ct[256][DP] = centres rounded to nearest tf32, zero padded; cnorm[256] = ||c||^2 (+inf for padding clusters)
__global__ void __launch_bounds__(256) k_prep_centers_t(const float* __restrict__ C, int k, int d, int DP,
                                                        float* __restrict__ Ct, float* __restrict__ cnorm,
                                                        const B2kLoopState* st) {
  if (st != nullptr && st->done) return;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= 256) return;
  double s = 0.0;
  for (int t = lane; t < DP; t += 32) {
    const float v = (row < k && t < d) ? C[(size_t)row * d + t] : 0.f;
    Ct[(size_t)row * DP + t] = __uint_as_float(rn_tf32_bits(v));
    s += (double)v * (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) cnorm[row] = row < k ? (float)s : __int_as_float(0x7f800000);
}

// One block of 256 threads: (a) cluster -> update-warp key table balanced by the previous iteration's cluster
// sizes (32 "virtual" update warps = 16 per CTA of the pair, 8 slots each; key = vwarp * 8 + slot, so that
// key >> 7 = owning CTA); (b) the recheck threshold thr(x) = thrA * ||x|| + thrB.
//
// Bound.  dist'_j = fl(||c_j||^2 - 2 x~.c~_j) with x~ = x truncated (or rounded) to tf32 by the tensor core:
// |x~_t - x_t| <= 2^-10 |x_t|;  c~ = RN_tf32(c): |c~_t - c_t| <= 2^-11 |c_t|.  Hence
//   |x~.c~ - x.c| <= (2^-10 + 2^-11 + 2^-21) sum_t |x_t||c_t| <= 1.5005 * 2^-10 ||x|| ||c||       (Cauchy-Schwarz)
// tf32 products are exact in fp32; the fp32 accumulation of d <= 256 terms adds at most d * 2^-23 ||x|| ||c||
// <= 2^-15 ||x|| ||c|| (truncating adder assumed); the final fma rounds once (2^-24 relative) and the key drops 8
// mantissa bits (2^-15 relative), both relative to |dist'| <= ||c||^2 + 2 ||x|| ||c||.  With Cmax = max_j ||c_j||:
//   E <= ||x|| Cmax (3.001 * 2^-10 + 2^-14 + 2^-14 + 2^-23) + Cmax^2 (2^-15 + 2^-24)
// Two approximate distances can be off by E each in opposite directions, so the argmin is proven whenever the gap
// exceeds 2E <= ||x|| Cmax * 6.26 * 2^-10 + Cmax^2 * 2^-14.  Shipped with a 1.27x margin (also covers the fp32
// rounding of ||x||, ||c||): thrA = Cmax * 2^-7, thrB = Cmax^2 * 2^-13.
__global__ void __launch_bounds__(256) k_tables_t(const double* __restrict__ counts, int k, const float* __restrict__ cnorm,
                                                  uint8_t* __restrict__ keytab, uint8_t* __restrict__ keyinv,
                                                  float* __restrict__ thr, const B2kLoopState* st) {
  if (st != nullptr && st->done) return;
  __shared__ double w[256];
  __shared__ float cmax2[8];
  const int j = threadIdx.x;
  w[j] = (counts != nullptr && j < k) ? counts[j] : -1.0;   // padding clusters sort last
  float c2 = j < k ? cnorm[j] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c2 = fmaxf(c2, __shfl_xor_sync(0xffffffffu, c2, o));
  if ((j & 31) == 0) cmax2[j >> 5] = c2;
  __syncthreads();
  if (j == 0) {
    float m = 0.f;
    for (int i = 0; i < 8; ++i) m = fmaxf(m, cmax2[i]);
    thr[0] = sqrtf(m) * 0.0078125f;       // Cmax * 2^-7
    thr[1] = m * 0.0001220703125f;        // Cmax^2 * 2^-13
  }
  int key;
  if (counts == nullptr) {
    key = (j & 31) * CPW + (j >> 5);
  } else {
    int rank = 0;   // position in (count desc, index asc) order
    for (int i = 0; i < 256; ++i) rank += (w[i] > w[j]) || (w[i] == w[j] && i < j);
    const int round = rank >> 5, pos = rank & 31;
    const int owner = (round & 1) ? (31 - pos) : pos;
    key = owner * CPW + round;
  }
  keytab[j] = (uint8_t)key;
  keyinv[key] = (uint8_t)j;
}

// xnorm[i] = ||x_i|| (fp32).  One pass over X, once per fit / lloyd / assign call (X is immutable during the call).
__global__ void __launch_bounds__(256) k_row_norms(const float* __restrict__ X, int64_t n, int d, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int d4 = d >> 2;
  for (int64_t row = warp0; row < n; row += nwarps) {
    const float4* p = reinterpret_cast<const float4*>(X + row * d);
    float s = 0.f;
    for (int t = lane; t < d4; t += 32) {
      const float4 v = __ldcs(p + t);
      s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = sqrtf(s);
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
struct TArgs {
  int64_t n;
  int nsteps;
  int k;
  int d;
  const float* Ct;         // [256][DP] tf32 centres
  const float* C32;        // [k][d] fp32 centres (recheck, min distance)
  const float* cnorm;      // [256]
  const float* thr;        // [2]
  const float* xnorm;      // [n]
  const uint8_t* keytab;   // [256]
  const uint8_t* keyinv;   // [256]
  float* partials;         // [npairs][k*d]
  int32_t* counts;         // [npairs][k]
  double* cost_partials;   // [grid]
  int32_t* labels_out;     // [n] or NULL
  float* mind_out;         // [n] or NULL
  int need_cost;
  unsigned long long* rstat;   // [2] rechecked rows, candidates evaluated (diagnostics) or NULL
  const B2kLoopState* st;
};

template <int NCH, bool UPD>
__global__ void __launch_bounds__(NTHREADS, 1) k_fused_t(const __grid_constant__ CUtensorMap mapX, const TArgs args) {
  constexpr int NSTEP = NSLOT / NCH;          // steps resident in the ring (3 at DP = 256)
  constexpr int UPL = NCH / 4;                // float4 units per lane of a row (1 or 2)
  static_assert(NCH == 4 || NCH == 8, "NCH");
  if (args.st != nullptr && args.st->done) return;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* gbase = smem_raw;
  if ((base & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("b2k fused_t: dynamic shared memory base %u is not 1 KB aligned\n", base);
    __trap();
  }
  const uint32_t ring = base + OFF_RING;
  const uint32_t bars = base + OFF_BARS;
  auto bar = [&](int i) -> uint32_t { return bars + 8u * (uint32_t)i; };
  uint2* part_s = reinterpret_cast<uint2*>(gbase + OFF_PART);
  int32_t* lab_s = reinterpret_cast<int32_t*>(gbase + OFF_LAB);
  uint32_t* cand_s = reinterpret_cast<uint32_t*>(gbase + OFF_CAND);
  float* candT_s = reinterpret_cast<float*>(gbase + OFF_CANDT);
  int32_t* fin_s = reinterpret_cast<int32_t*>(gbase + OFF_FIN);
  uint8_t* sort_s = gbase + OFF_SORT;
  uint8_t* keytab_s = sort_s + SortT::KEYTAB;
  uint8_t* keyinv_s = sort_s + SortT::KEYINV;
  uint32_t* flagw_s = reinterpret_cast<uint32_t*>(gbase + OFF_MISC + MISC_FLAGW);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(gbase + OFF_MISC + MISC_TMEMPTR);
  double* cost_s = reinterpret_cast<double*>(gbase + OFF_MISC + MISC_COST);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t peer = rank ^ 1u;

  // ---- one-time setup ----
  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&mapX);
    for (int i = 0; i < NSLOT; ++i) mbar_init(bar(B_XFULL + i), 1);
    for (int i = 0; i < 8; ++i) mbar_init(bar(B_SFREE + i), 2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_DFULL + i), 1);
      mbar_init(bar(B_DEMPTY + i), 2);
      mbar_init(bar(B_LFULL + i), 1);
      mbar_init(bar(B_LEMPTY + i), 1);
    }
    mbar_init(bar(B_EX), 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int j = threadIdx.x; j < 256; j += NTHREADS) {
    keytab_s[j] = args.keytab[j];
    keyinv_s[j] = args.keyinv[j];
  }
  if (threadIdx.x < 16) cost_s[threadIdx.x] = 0.0;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // peer barriers initialised / TMEM allocated before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

  // static schedule: cluster q handles steps q, q + nclusters, ...; CTA `rank` loads rows [128 step + 64 rank, +64)
  const int sched0 = (int)(blockIdx.x >> 1);
  const int sched_step = (int)(gridDim.x >> 1);
  const int nit = sched0 < args.nsteps ? (args.nsteps - sched0 + sched_step - 1) / sched_step : 0;
  auto step_of = [&](int it) -> int { return sched0 + it * sched_step; };

  // A operand: this CTA's 128 centres (tf32) -> TMEM columns [0, 32 NCH), lane = centre
  if (warp < W_UPD0) {
    const int jl = warp * 32 + lane;
    const float4* src = reinterpret_cast<const float4*>(args.Ct + ((size_t)rank * KH + jl) * (NCH * CHUNK));
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
    for (int cb = 0; cb < NCH; ++cb) {
      uint32_t v[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 f = __ldg(src + cb * 8 + q);
        v[q * 4 + 0] = __float_as_uint(f.x);
        v[q * 4 + 1] = __float_as_uint(f.y);
        v[q * 4 + 2] = __float_as_uint(f.z);
        v[q * 4 + 3] = __float_as_uint(f.w);
      }
      tmem_st_x32(tmem_base + lane_field + (uint32_t)(cb * 32), v);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' A operands are in TMEM before the leader's first MMA
  tc_fence_after();

  if (warp >= W_UPD0 && warp < W_TMA) {
    // ======================= update warps =======================
    const int u = warp - W_UPD0;
    if constexpr (UPD && NCH == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;" ::: "memory");
    uint64_t acc[CPW][UPL][2];
    int cnt = 0;   // lane c < CPW: rows of owned cluster slot c
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
#pragma unroll
      for (int i = 0; i < UPL; ++i) acc[c][i][0] = acc[c][i][1] = 0ull;
    }
    const uint32_t sfree_peer0 = mapa_u32(bar(B_SFREE), peer);
    const uint32_t unit_js = (uint32_t)(lane & 7) << 4;
    double cost = 0.0;
    for (int it = 0; it < nit; ++it) {
      const int b = it & 1;
      const uint32_t bph = (uint32_t)(it >> 1) & 1u;
      const int ss = it % NSTEP;
      if (u == 0) mbar_wait_nocall(bar(B_LFULL + b), bph);
      asm volatile("bar.sync 2, 512;" ::: "memory");
      // The x_full phases of this step completed before the MMA consumed the slots, which happens-before the
      // commit, the epilogue and hence lab_full: the rows are in shared memory (both CTAs).
      const uint32_t slot0 = ring + (uint32_t)((ss * NCH + (lane >> 3)) * SLOT_BYTES);
      if constexpr (UPD) {
        const uint16_t* rows_sorted = reinterpret_cast<const uint16_t*>(sort_s + SortT::ROWS) + b * TN;
        const uint8_t* start = sort_s + SortT::START + b * 144;
        auto load_row = [&](uint32_t e, uint64_t (&v)[UPL][2]) {
          const uint32_t a0 = slot0 + ((e & 0x7fffu) ^ unit_js);
          if ((e >> 15) == rank) {
#pragma unroll
            for (int i = 0; i < UPL; ++i) lds128_2(a0 + (uint32_t)(i * 4 * SLOT_BYTES), v[i][0], v[i][1]);
          } else {
            const uint32_t r0 = mapa_u32(a0, peer);
#pragma unroll
            for (int i = 0; i < UPL; ++i) ld_cluster_2(r0 + (uint32_t)(i * 4 * SLOT_BYTES), v[i][0], v[i][1]);
          }
        };
        const int sv = (lane <= CPW) ? (int)start[u * CPW + lane] : 0;
        cnt += __shfl_down_sync(0xffffffffu, sv, 1) - sv;   // meaningful in lanes < CPW
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
          const int i0 = __shfl_sync(0xffffffffu, sv, c);
          const int i1 = __shfl_sync(0xffffffffu, sv, c + 1);
#pragma unroll 1
          for (int i = i0; i < i1; i += 2) {
            const bool two = (i + 1 < i1);
            const uint32_t e0 = rows_sorted[i];
            const uint32_t e1 = rows_sorted[two ? i + 1 : i];
            uint64_t v0[UPL][2], v1[UPL][2];
            load_row(e0, v0);
            load_row(e1, v1);
#pragma unroll
            for (int k2 = 0; k2 < UPL; ++k2) {
              acc[c][k2][0] = add2(acc[c][k2][0], v0[k2][0]);
              acc[c][k2][1] = add2(acc[c][k2][1], v0[k2][1]);
            }
            if (two) {
#pragma unroll
              for (int k2 = 0; k2 < UPL; ++k2) {
                acc[c][k2][0] = add2(acc[c][k2][0], v1[k2][0]);
                acc[c][k2][1] = add2(acc[c][k2][1], v1[k2][1]);
              }
            }
          }
        }
      } else if (args.need_cost) {
        // exact min distance of this CTA's own rows: sum_t (x_t - c_t)^2 against the row's (final) centre
#pragma unroll 1
        for (int rr = 0; rr < TNH / N_UPD; ++rr) {
          const int lrow = u + N_UPD * rr;
          const int col = (int)rank * TNH + lrow;
          const int64_t grow = (int64_t)step_of(it) * TN + col;
          if (grow >= args.n) continue;
          const int label = lab_s[b * TN + col];
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < UPL; ++i) {
            const int cc = (lane + 32 * i) * 4;
            if (cc < args.d) {
              const float4 xv = lds128(slot0 + (uint32_t)(i * 4 * SLOT_BYTES) +
                                       (((uint32_t)(lrow * 128 + ((lrow & 7) << 4))) ^ unit_js));
              const float4 cv = __ldg(reinterpret_cast<const float4*>(args.C32 + (size_t)label * args.d + cc));
              const float dx = xv.x - cv.x, dy = xv.y - cv.y, dz = xv.z - cv.z, dw = xv.w - cv.w;
              s = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, s))));
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0) {
            if (args.mind_out != nullptr) args.mind_out[grow] = s;
            cost += (double)s;
          }
        }
      }
      asm volatile("bar.sync 3, 512;" ::: "memory");
      if (u == 0 && lane == 0) {
        mbar_arrive(bar(B_LEMPTY + b));
        asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(bar(B_SFREE + ss)) : "memory");
        mbar_arrive_release_cluster(sfree_peer0 + 8u * (uint32_t)ss);
      }
    }
    if constexpr (UPD) {
      // flush: partials[pair][l][col .. col+3] for the owned clusters l = keyinv[rank*128 + u*8 + c]
      float* out = args.partials + (size_t)(blockIdx.x >> 1) * args.k * args.d;
#pragma unroll
      for (int c = 0; c < CPW; ++c) {
        const int l = (int)keyinv_s[(int)rank * KH + u * CPW + c];
        if (l < args.k) {
#pragma unroll
          for (int i = 0; i < UPL; ++i) {
            const int colx = (lane + 32 * i) * 4;
            float e[4];
            unpack2(acc[c][i][0], e[0], e[1]);
            unpack2(acc[c][i][1], e[2], e[3]);
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (colx + t < args.d) out[(size_t)l * args.d + colx + t] = e[t];
          }
          const int cn_c = __shfl_sync(0xffffffffu, cnt, c);
          if (lane == 0) args.counts[(size_t)(blockIdx.x >> 1) * args.k + l] = cn_c;
        }
      }
    } else {
      if (lane == 0) cost_s[u] = cost;
    }
  } else if (warp >= W_TMA) {
    if constexpr (UPD && NCH == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
   if (warp == W_TMA) {
    // ======================= TMA producer =======================
    const uint32_t xfull_leader0 = mapa_u32(bar(B_XFULL), 0u);   // the leader CTA's x_full[0] (cluster address)
    for (int it = 0; it < nit; ++it) {
      const int step = step_of(it);
      const int ss = it % NSTEP;
      const uint32_t sph = (uint32_t)(it / NSTEP) & 1u;
      mbar_wait_cluster_nocall(bar(B_SFREE + ss), sph ^ 1u);   // acquire.cluster: the peer's update role read this slot
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int slot = ss * NCH + c;
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(bar(B_XFULL + slot), 2u * SLOT_BYTES);
          tma_load_2d_pair(ring + slot * SLOT_BYTES, &mapX, xfull_leader0 + 8u * (uint32_t)slot, c * CHUNK,
                           step * TN + (int)rank * TNH);
        }
        __syncwarp();
      }
    }
  } else if (warp == W_MMA) {
    // ======================= MMA issuer (leader CTA only) =======================
    constexpr uint32_t idesc = make_idesc_tf32(256, TN);
    for (int it = 0; it < (rank != 0 ? 0 : nit); ++it) {
      const int b = it & 1;
      const uint32_t bph = (uint32_t)(it >> 1) & 1u;
      const int ss = it % NSTEP;
      const uint32_t sph = (uint32_t)(it / NSTEP) & 1u;
      mbar_wait_cluster_nocall(bar(B_DEMPTY + b), bph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + D_OFF + b * TN;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int slot = ss * NCH + c;
        mbar_wait_cluster_nocall(bar(B_XFULL + slot), sph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t bx = ring + slot * SLOT_BYTES;
#pragma unroll
          for (int ks = 0; ks < CHUNK / 8; ++ks)
            tc_mma_ts_tf32_pair(d_tmem, tmem_base + (uint32_t)(c * CHUNK + ks * 8), make_kmajor_sw128_desc(bx + ks * 32),
                                idesc, (c | ks) != 0 ? 1u : 0u);
          if (c == NCH - 1) tc_commit_pair(bar(B_DFULL + b));
        }
        __syncwarp();
      }
    }
   }
  } else if (warp < W_UPD0) {
    // ======================= epilogue warps: D -> keys -> (best, second) -> labels =======================
    const int w = warp - W_EPI0;
    const int col = w * 32 + lane;                      // the column (row of the step) this thread combines
    const uint32_t j = rank * KH + (uint32_t)col;       // the cluster of this thread's TMEM lane
    const uint32_t lane_field = (uint32_t)(w * 32) << 16;
    const float cn = args.cnorm[j];
    const float thrA = args.thr[0], thrB = args.thr[1];
    const uint32_t part_local0 = base + OFF_PART;
    const uint32_t part_peer0 = mapa_u32(part_local0, peer);
    const uint32_t ex_peer = mapa_u32(bar(B_EX), peer);
    const uint32_t dempty_leader = mapa_u32(bar(B_DEMPTY), 0u);
    uint32_t exph = 0;
    unsigned long long n_flag = 0, n_cand = 0;
    // one exchange: everything this CTA's epilogue stored (locally and into the peer) is visible to both afterwards
    auto exchange = [&]() {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(bar(B_EX)) : "memory");
        mbar_arrive_release_cluster(ex_peer);
      }
      if (w == 0) mbar_wait_cluster(bar(B_EX), exph);
      exph ^= 1u;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    for (int it = 0; it < nit; ++it) {
      const int step = step_of(it);
      const int b = it & 1;
      const uint32_t bph = (uint32_t)(it >> 1) & 1u;
      const int64_t grow = (int64_t)step * TN + col;
      const bool valid = grow < args.n;
      const float xn = valid ? __ldg(args.xnorm + grow) : 0.f;
      if (w == 0) mbar_wait(bar(B_DFULL + b), bph);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      tc_fence_after();
      const uint32_t src = rank * 4u + (uint32_t)w;
      const uint32_t poff = (uint32_t)(((b * 8 + (int)src) * TN) * 8);
#pragma unroll 1
      for (int g = 0; g < TN / 32; ++g) {
        uint32_t v[32];
        tmem_ld_x32(tmem_base + lane_field + (uint32_t)(D_OFF + b * TN + g * 32), v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = pack_key(fmaf(-2.f, __uint_as_float(v[i]), cn), j);
        uint32_t m1, m2;
        bfly32(v, lane, m1, m2);
        const uint32_t o = poff + (uint32_t)((g * 32 + lane) * 8);
        part_s[(b * 8 + (int)src) * TN + g * 32 + lane] = make_uint2(m1, m2);
        st_cluster_v2(part_peer0 + o, m1, m2);
      }
      tc_fence_before();
      exchange();
      // combine the 8 partials of my column
      uint32_t M1 = 0xffffffffu, M2 = 0xffffffffu;
#pragma unroll
      for (int s8 = 0; s8 < 8; ++s8) {
        const uint2 p = part_s[(b * 8 + s8) * TN + col];
        merge2(M1, M2, p.x, p.y);
      }
      int label = (int)(M1 & 255u);
      const float M1f = unpack_key(M1), M2f = unpack_key(M2);
      const float thr = fmaf(xn, thrA, thrB);
      const bool flag = valid && ((M2f - M1f) < thr);
      candT_s[col] = M1f + thr;
      const uint32_t fl = __ballot_sync(0xffffffffu, flag);
      if (lane == 0) flagw_s[w] = fl;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      uint32_t fw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) fw[i] = flagw_s[i];
      if ((fw[0] | fw[1] | fw[2] | fw[3]) != 0u) {
        // ---- exact recheck of the flagged rows (identical decision in both CTAs) ----
        // (1) candidate masks: clusters whose approximate distance is within thr of the best
#pragma unroll 1
        for (int wd = 0; wd < 4; ++wd) {
          uint32_t m = fw[wd];
          while (m) {
            const int c = wd * 32 + (__ffs(m) - 1);
            m &= m - 1;
            const uint32_t dv = tmem_ld_x1(tmem_base + lane_field + (uint32_t)(D_OFF + b * TN + c));
            tmem_wait_ld();
            const float dist = fmaf(-2.f, __uint_as_float(dv), cn);
            const uint32_t bm = __ballot_sync(0xffffffffu, dist <= candT_s[c]);
            if (lane == 0) {
              const uint32_t a = base + OFF_CAND + (uint32_t)((c * 8 + (int)src) * 4);
              if ((uint32_t)(c >> 6) == rank) cand_s[c * 8 + (int)src] = bm;
              else st_cluster_u32(mapa_u32(a, peer), bm);
            }
          }
        }
        tc_fence_before();
        exchange();
        // (2) the row's owner evaluates the candidates exactly, ascending cluster order, strict '<'
        {
          const int ss = it % NSTEP;
          int idx = 0;
#pragma unroll 1
          for (int wd = (int)rank * 2; wd < (int)rank * 2 + 2; ++wd) {
            uint32_t m = fw[wd];
            while (m) {
              const int c = wd * 32 + (__ffs(m) - 1);
              m &= m - 1;
              if ((idx++ & 3) != w) continue;
              const int lrow = c & 63;
              float4 xv[UPL];
#pragma unroll
              for (int u2 = 0; u2 < UPL; ++u2) {
                const int q = lane + 32 * u2;
                xv[u2] = lds128(ring + (uint32_t)((ss * NCH + (q >> 3)) * SLOT_BYTES + lrow * 128) +
                                (uint32_t)(((q & 7) ^ (lrow & 7)) << 4));
              }
              float best = __int_as_float(0x7f800000);
              int bj = -1;
#pragma unroll 1
              for (int s8 = 0; s8 < 8; ++s8) {
                uint32_t bm = cand_s[c * 8 + s8];
                while (bm) {
                  const int jj = s8 * 32 + (__ffs(bm) - 1);
                  bm &= bm - 1;
                  if (jj >= args.k) continue;
                  float dot = 0.f;
#pragma unroll
                  for (int u2 = 0; u2 < UPL; ++u2) {
                    const int cc = (lane + 32 * u2) * 4;
                    if (cc < args.d) {
                      const float4 cv = __ldg(reinterpret_cast<const float4*>(args.C32 + (size_t)jj * args.d + cc));
                      dot = fmaf(xv[u2].x, cv.x, fmaf(xv[u2].y, cv.y, fmaf(xv[u2].z, cv.z, fmaf(xv[u2].w, cv.w, dot))));
                    }
                  }
#pragma unroll
                  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                  const float de = fmaf(-2.f, dot, __ldg(args.cnorm + jj));
                  if (de < best) { best = de; bj = jj; }
                  ++n_cand;
                }
              }
              ++n_flag;
              if (lane == 0) {
                fin_s[c] = bj;
                st_cluster_u32(mapa_u32(base + OFF_FIN + (uint32_t)(c * 4), peer), (uint32_t)bj);
              }
            }
          }
        }
        exchange();
        if (flag) {
          const int f = fin_s[col];
          if (f >= 0) label = f;
        }
      }
      // D (and the exchange buffers of this parity) are drained in this CTA
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(dempty_leader + 8u * (uint32_t)b)
                     : "memory");
      }
      // ---- publish: labels (+ sorted row lists in UPD mode) for the update role ----
      mbar_wait(bar(B_LEMPTY + b), bph ^ 1u);
      lab_s[b * TN + col] = label;
      if (valid && (uint32_t)(col >> 6) == rank && args.labels_out != nullptr) args.labels_out[grow] = label;
      if constexpr (UPD) {
        uint8_t* cnt = sort_s + SortT::CNT + b * (4 * KH);
        uint16_t* rows_sorted = reinterpret_cast<uint16_t*>(sort_s + SortT::ROWS) + b * TN;
        uint8_t* start = sort_s + SortT::START + b * 144;
        reinterpret_cast<uint32_t*>(cnt + w * KH)[lane] = 0u;
        __syncwarp();
        const int kfull = (int)keytab_s[label];
        const bool mine = valid && (uint32_t)(kfull >> 7) == rank;
        const int key = mine ? (kfull & 127) : KH;
        const uint32_t same = __match_any_sync(0xffffffffu, key);
        const int rnk = __popc(same & ((1u << lane) - 1u));
        if (mine && rnk == 0) cnt[w * KH + key] = (uint8_t)__popc(same);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        int tot[4];
        int lane_sum = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kk = lane * 4 + i;
          tot[i] = (int)cnt[kk] + (int)cnt[KH + kk] + (int)cnt[2 * KH + kk] + (int)cnt[3 * KH + kk];
          lane_sum += tot[i];
        }
        int incl = lane_sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        int st4[4];
        st4[0] = incl - lane_sum;
#pragma unroll
        for (int i = 1; i < 4; ++i) st4[i] = st4[i - 1] + tot[i - 1];
        int my_start = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int t = __shfl_sync(0xffffffffu, st4[i], (key < KH ? key : 0) >> 2);
          if ((key & 3) == i) my_start = t;
        }
        int pos = my_start + rnk;
        if (mine) {
#pragma unroll
          for (int q2 = 0; q2 < 3; ++q2)
            if (q2 < w) pos += (int)cnt[q2 * KH + key];
          // entry: bit 15 = CTA whose ring holds the row, low bits = byte offset of the row in a chunk slot with the
          // swizzle phase folded in (address of 16-byte unit u = slot + (entry ^ (u << 4)))
          const int lrow = col & 63;
          rows_sorted[pos] = (uint16_t)(((col >> 6) << 15) | (lrow * 128 + ((lrow & 7) << 4)));
        }
        if (w == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) start[lane * 4 + i] = (uint8_t)st4[i];
          if (lane == 31) start[KH] = (uint8_t)incl;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 0) mbar_arrive(bar(B_LFULL + b));
    }
    if (args.rstat != nullptr && lane == 0 && (n_flag | n_cand) != 0ull) {
      atomicAdd(args.rstat + 0, n_flag);
      atomicAdd(args.rstat + 1, n_cand);
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = 0.0;
    for (int i = 0; i < 16; ++i) c += cost_s[i];
    args.cost_partials[blockIdx.x] = c;
  }
  cluster_sync_all();   // the peer may still receive multicast commits / remote arrivals / DSMEM reads
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TLayout {
  size_t off_ct, off_cnorm, off_thr, off_tab, off_rstat, off_partials, off_counts, off_cost, off_xnorm, total;
};
TLayout t_layout(const B2kFusedPlan& p, int64_t n, int k, int d) {
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  TLayout L{};
  size_t o = 0;
  L.off_ct = o; o = al(o + (size_t)256 * p.DP * 4);
  L.off_cnorm = o; o = al(o + 256 * 4);
  L.off_thr = o; o = al(o + 16);
  L.off_tab = o; o = al(o + 512);
  L.off_rstat = o; o = al(o + 16);
  L.off_partials = o; o = al(o + (size_t)p.P * k * d * 4);
  L.off_counts = o; o = al(o + (size_t)p.P * k * 4);
  L.off_cost = o; o = al(o + (size_t)p.grid * 8);
  L.off_xnorm = o; o = al(o + (size_t)(n > 0 ? n : 1) * 4);
  L.total = o;
  return L;
}

template <int NCH, bool UPD>
int launch_t(b2k_ctx* ctx, int grid, const CUtensorMap& mx, const TArgs& a, cudaStream_t s) {
  auto kern = k_fused_t<NCH, UPD>;
  B2K_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B2K_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, kern, mx, a));
  B2K_CUDA_OK(ctx, cudaGetLastError());
  return B2K_OK;
}
}  // namespace

bool b2k_fused_t_supported(const b2k_ctx* ctx, int64_t n, int d, int k, const float* X) {
  (void)ctx;
  if (n < 1 || n > (int64_t)0x7fffff00 * 1LL) return false;
  if (d % 4 != 0 || d > 256 || k > 256) return false;
  if ((reinterpret_cast<uintptr_t>(X) & 15u) != 0) return false;
  return true;
}

int b2k_fused_t_plan(b2k_ctx* ctx, int64_t n, int d, int k, B2kFusedPlan* plan) {
  plan->variant = 1;
  plan->KP = 256;
  plan->DP = d <= 128 ? 128 : 256;
  plan->pair = 1;
  const int64_t nsteps = (n + TN - 1) / TN;
  int grid = ctx->sm_count & ~1;
  if (ctx->grid_limit > 0 && ctx->grid_limit < grid) grid = ctx->grid_limit & ~1;
  if (nsteps * 2 < grid) grid = (int)nsteps * 2;
  if (grid < 2) grid = 2;
  plan->grid = grid;
  plan->P = grid / 2;
  plan->scratch_bytes = t_layout(*plan, n, k, d).total;
  return B2K_OK;
}

void b2k_fused_t_views(const B2kFusedPlan& plan, void* plan_scratch, int64_t n, int k, int d, float** partials,
                       int32_t** counts, double** cost_partials, unsigned long long** rstat) {
  TLayout L = t_layout(plan, n, k, d);
  char* b = static_cast<char*>(plan_scratch);
  *partials = reinterpret_cast<float*>(b + L.off_partials);
  *counts = reinterpret_cast<int32_t*>(b + L.off_counts);
  *cost_partials = reinterpret_cast<double*>(b + L.off_cost);
  if (rstat) *rstat = reinterpret_cast<unsigned long long*>(b + L.off_rstat);
}

// once per fit / lloyd / assign call: row norms of X into the plan scratch; clears the recheck counters
int b2k_fused_t_prepare(b2k_ctx* ctx, const B2kFusedPlan& plan, void* plan_scratch, const float* X, int64_t n, int d,
                        int k, cudaStream_t s) {
  TLayout L = t_layout(plan, n, k, d);
  char* b = static_cast<char*>(plan_scratch);
  B2K_CUDA_OK(ctx, cudaMemsetAsync(b + L.off_rstat, 0, 16, s));
  int blocks = ctx->sm_count * 8;
  k_row_norms<<<blocks, 256, 0, s>>>(X, n, d, reinterpret_cast<float*>(b + L.off_xnorm));
  ctx->stats.kernel_launches++;
  B2K_CUDA_OK(ctx, cudaGetLastError());
  return B2K_OK;
}

int b2k_launch_fused_t(b2k_ctx* ctx, const B2kFusedPlan& plan, void* plan_scratch, const float* X, int64_t n, int d,
                       const float* C, int k, int32_t* labels_out, float* mindist_out, bool do_update, bool need_cost,
                       const B2kLoopState* st, cudaStream_t s, const double* prev_counts) {
  TLayout L = t_layout(plan, n, k, d);
  char* b = static_cast<char*>(plan_scratch);
  float* Ct = reinterpret_cast<float*>(b + L.off_ct);
  float* cnorm = reinterpret_cast<float*>(b + L.off_cnorm);
  float* thr = reinterpret_cast<float*>(b + L.off_thr);
  uint8_t* keytab = reinterpret_cast<uint8_t*>(b + L.off_tab);

  k_prep_centers_t<<<32, 256, 0, s>>>(C, k, d, plan.DP, Ct, cnorm, st);
  k_tables_t<<<1, 256, 0, s>>>(do_update ? prev_counts : nullptr, k, cnorm, keytab, keytab + 256, thr, st);
  ctx->stats.kernel_launches += 2;
  B2K_CUDA_OK(ctx, cudaGetLastError());

  CUtensorMap mx;
  B2K_TRY(b2k_fused_encode_2d(ctx, &mx, X, (uint64_t)d, (uint64_t)n, (uint64_t)d * 4, CHUNK, TNH, 1));

  TArgs a{};
  a.n = n;
  a.nsteps = (int)((n + TN - 1) / TN);
  a.k = k;
  a.d = d;
  a.Ct = Ct;
  a.C32 = C;
  a.cnorm = cnorm;
  a.thr = thr;
  a.xnorm = reinterpret_cast<const float*>(b + L.off_xnorm);
  a.keytab = keytab;
  a.keyinv = keytab + 256;
  a.partials = reinterpret_cast<float*>(b + L.off_partials);
  a.counts = reinterpret_cast<int32_t*>(b + L.off_counts);
  a.cost_partials = reinterpret_cast<double*>(b + L.off_cost);
  a.labels_out = labels_out;
  a.mind_out = mindist_out;
  a.need_cost = need_cost ? 1 : 0;
  a.rstat = reinterpret_cast<unsigned long long*>(b + L.off_rstat);
  a.st = st;

  int rc;
  if (plan.DP == 128) rc = do_update ? launch_t<4, true>(ctx, plan.grid, mx, a, s) : launch_t<4, false>(ctx, plan.grid, mx, a, s);
  else rc = do_update ? launch_t<8, true>(ctx, plan.grid, mx, a, s) : launch_t<8, false>(ctx, plan.grid, mx, a, s);
  B2K_TRY(rc);
  ctx->stats.kernel_launches++;
  ctx->stats.fused_tc_launches++;
  return B2K_OK;
}
