// Persistent warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] (+)= sum_k A(m,k) * B(n,k),  bf16/fp16 in, fp32
// accumulate in TMEM, bf16/fp16/fp32 out.
//
//   * operands are staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a 4-stage mbarrier ring;
//   * one elected thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 16) reading smem through matrix descriptors;
//   * the fp32 accumulator lives in TMEM and is double buffered (2 x BLOCK_N columns) so the epilogue of tile i
//     (tcgen05.ld -> convert -> 16-byte global stores, optional accumulate-into-C) overlaps the main loop of tile i+1;
//   * A and B may each be K-major ("row = M/N, contiguous K") or MN-major ("row = K, contiguous M/N"), which covers the
//     three GEMMs of a linear layer without any transposition pass:
//         fwd   y  = x  W^T : A = x  (K-major),  B = W  (K-major)
//         dgrad dx = dy W   : A = dy (K-major),  B = W  (MN-major: stored [N_out(red), K_in])
//         wgrad dW = dy^T x : A = dy (MN-major), B = x  (MN-major), optional fp32/bf16 accumulation into main_grad.
//
// Role layout (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (one TMEM lane
// quarter each).  Grid = #SMs (persistent); tiles are rasterised in groups of 8 M-blocks for L2 reuse of B.
//
// This kernel replaces the reference's cuBLAS calls (`F.linear` / `torch.matmul` in shardformer/layer/_operation.py)
// with our own tensor-core path; the comm-fused variants in fused_comm_gemm.cu share its main loop.
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

using namespace sm100;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;        // 64 x 16-bit = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int GROUP_M = 8;

template <int BLOCK_N> struct Cfg {
  static constexpr int STAGES = BLOCK_N == 256 ? 4 : 6;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N >= 512 ? 512 : (2 * BLOCK_N >= 256 ? 256 : 128);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmParams {
  int M, N, K;
  int ldc;
  int a_mn_major, b_mn_major;
  int out_dtype;      // CB_F32 / CB_F16 / CB_BF16
  int accumulate;     // D += existing C
  uint32_t idesc;
  void* C;
  const float* scale_a = nullptr;   // fp8 path: per-tensor dequantisation scales (device pointers, may be null)
  const float* scale_b = nullptr;
  // tail-wave split (CTA-pair kernel): the tiles of the last, partial wave are cut along K into `tail_split` work
  // units each so that all CTA pairs stay busy; units j < tail_split - 1 park their fp32 partial tile in `ws_part` and
  // raise `ws_flag`, the last unit folds them in before its epilogue store (stream-K style fix-up, no atomics on C).
  int tail_split = 1;
  // tail-wave N halving (CTA-pair kernel): the tiles of the partial last wave are cut into two 256 x 128 halves
  // (UMMA 256 x 128 x 16, 64 B columns per CTA) when that fills more CTA pairs; no fix-up needed.  `idesc_half` is the
  // instruction descriptor with N = 128.
  int tail_half = 0;
  uint32_t idesc_half = 0;
  float* ws_part = nullptr;         // [tail tiles][tail_split - 1][256][256] fp32
  uint32_t* ws_flag = nullptr;      // [tail tiles][tail_split - 1][2 CTAs][4 epilogue warps]
};

// Work unit -> (tile, k-block range) of the CTA-pair kernel.  Units [0, full) are whole tiles of the complete waves;
// the rest enumerates (tail tile, split index) pairs, split index fastest, so the units of one tile run on
// neighbouring pairs in the same (last) wave.
struct WorkUnit {
  int tile, kb_begin, kb_end, split, nsplit, tail_idx;
  int half;      // -1: whole 256 x 256 tile; 0 / 1: left / right 256 x 128 half of a tail tile
};
SM100_DEVICE WorkUnit get_unit(int unit, int full, int tail_split, int k_blocks, int tail_half) {
  WorkUnit w;
  w.half = -1;
  if (unit < full) {
    w.tile = unit; w.kb_begin = 0; w.kb_end = k_blocks; w.split = 0; w.nsplit = 1; w.tail_idx = 0;
  } else if (tail_half) {
    const int u = unit - full;
    w.tail_idx = u >> 1;
    w.half = u & 1;
    w.tile = full + w.tail_idx;
    w.kb_begin = 0; w.kb_end = k_blocks; w.split = 0; w.nsplit = 1;
  } else {
    const int u = unit - full;
    w.tail_idx = u / tail_split;
    w.split = u - w.tail_idx * tail_split;
    w.nsplit = tail_split;
    w.tile = full + w.tail_idx;
    // uneven K ranges: the last unit (the one that folds the partials in) gets 1.5 shares, so the writers' epilogues
    // (fp32 partial -> workspace) and their flags are done while the reader is still in its main loop
    const int share = (2 * k_blocks) / (2 * tail_split + 1);
    w.kb_begin = share * w.split;
    w.kb_end = (w.split == tail_split - 1) ? k_blocks : share * (w.split + 1);
  }
  return w;
}
SM100_DEVICE void st_release_gpu_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
SM100_DEVICE uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

SM100_DEVICE void tile_coords(int tile, int m_blocks, int n_blocks, int& m_blk, int& n_blk) {
  const int tiles_per_group = GROUP_M * n_blocks;
  const int group = tile / tiles_per_group;
  const int first_m = group * GROUP_M;
  const int gsize = min(m_blocks - first_m, GROUP_M);
  const int in_group = tile - group * tiles_per_group;
  m_blk = first_m + in_group % gsize;
  n_blk = in_group / gsize;
}

template <typename TO>
SM100_DEVICE void store_row_chunk(TO* __restrict__ dst, const uint32_t (&acc)[32], int n_valid, bool accumulate);

template <>
SM100_DEVICE void store_row_chunk<float>(float* __restrict__ dst, const uint32_t (&acc)[32], int n_valid,
                                         bool accumulate) {
  if (n_valid >= 32) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 v = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                             __uint_as_float(acc[j + 3]));
      if (accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(dst + j);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *reinterpret_cast<float4*>(dst + j) = v;
    }
  } else {
    // fully unrolled + predicated: a dynamically indexed `acc[j]` would force the whole fragment into local memory
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < n_valid) dst[j] = __uint_as_float(acc[j]) + (accumulate ? dst[j] : 0.f);
  }
}

template <typename T16>
SM100_DEVICE void store_row_chunk_16(T16* __restrict__ dst, const uint32_t (&acc)[32], int n_valid, bool accumulate) {
  if (n_valid >= 32) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      Vec16<T16> o;
      if (accumulate) o.load(dst + j);
#pragma unroll
      for (int i = 0; i < 8; ++i) o.set(i, __uint_as_float(acc[j + i]) + (accumulate ? o.get(i) : 0.f));
      o.store(dst + j);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < n_valid) dst[j] = from_f32<T16>(__uint_as_float(acc[j]) + (accumulate ? to_f32<T16>(dst[j]) : 0.f));
  }
}
template <>
SM100_DEVICE void store_row_chunk<__nv_bfloat16>(__nv_bfloat16* __restrict__ dst, const uint32_t (&acc)[32],
                                                 int n_valid, bool accumulate) {
  store_row_chunk_16<__nv_bfloat16>(dst, acc, n_valid, accumulate);
}
template <>
SM100_DEVICE void store_row_chunk<__half>(__half* __restrict__ dst, const uint32_t (&acc)[32], int n_valid,
                                          bool accumulate) {
  store_row_chunk_16<__half>(dst, acc, n_valid, accumulate);
}

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;      // [2]
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2; // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_blocks = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_blocks * n_blocks;
  const int k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================================================================ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, m_blocks, n_blocks, m_blk, n_blk);
      const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * C::A_BYTES;
          uint8_t* sb = smem_b + stage * C::B_BYTES;
          const int k0 = kb * BLOCK_K;
          if (!p.a_mn_major) {
            tma_load_2d(&tmap_a, &full_bar[stage], sa, k0, m0);                    // box [BLOCK_M rows, 64 k]
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)                                  // boxes [BLOCK_K rows, 64 m]
              tma_load_2d(&tmap_a, &full_bar[stage], sa + j * (BLOCK_K * 128), m0 + j * 64, k0);
          }
          if (!p.b_mn_major) {
            tma_load_2d(&tmap_b, &full_bar[stage], sb, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(&tmap_b, &full_bar[stage], sb + j * (BLOCK_K * 128), n0 + j * 64, k0);
          }
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    // byte advance of the descriptor start address per UMMA_K step
    const uint32_t a_kstep = p.a_mn_major ? UMMA_K * 128 : UMMA_K * 2;
    const uint32_t b_kstep = p.b_mn_major ? UMMA_K * 128 : UMMA_K * 2;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem_a + stage * C::A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * C::B_BYTES);
          const uint64_t da = p.a_mn_major ? make_smem_desc_sw128(sa, BLOCK_K * 128, 1024)
                                           : make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db = p.b_mn_major ? make_smem_desc_sw128(sb, BLOCK_K * 128, 1024)
                                           : make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_f16_ss(tmem_d, advance_desc(da, k * a_kstep), advance_desc(db, k * b_kstep), p.idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);                    // frees the smem slot when these MMAs retire
          if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      if (k_blocks == 0 && lane == 0) umma_commit(&tmem_full[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ================================================================ epilogue (warps 2..5)
    const int quarter = warp & 3;             // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, m_blocks, n_blocks, m_blk, n_blk);
      const int row = m_blk * BLOCK_M + quarter * 32 + lane;
      const int n0 = n_blk * BLOCK_N;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_wait();
        const int n_valid = p.N - (n0 + c);
        if (row < p.M && n_valid > 0 && p.K > 0) {
          const size_t off = (size_t)row * p.ldc + n0 + c;
          if (p.out_dtype == CB_BF16)
            store_row_chunk<__nv_bfloat16>((__nv_bfloat16*)p.C + off, v, n_valid, p.accumulate);
          else if (p.out_dtype == CB_F32)
            store_row_chunk<float>((float*)p.C + off, v, n_valid, p.accumulate);
          else
            store_row_chunk<__half>((__half*)p.C + off, v, n_valid, p.accumulate);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// =====================================================================================================================
// CTA-pair variant (cta_group::2): one 256 x 256 output tile per cluster of two CTAs.  Each CTA stages its own 128 rows
// of A and its own 128 columns of B (32 KB / stage instead of 48 KB for a 1-CTA 128x256 tile), the leader CTA issues
// UMMA 256x256x16 that reads both CTAs' shared memory, and each CTA owns the 128 x 256 fp32 accumulator slice of its
// rows in its own TMEM.  B is fetched from global memory once per pair, and the operand bytes read from shared memory
// per MMA flop drop by a third, which is what lifts the tensor pipe from ~71 % to the library's level.
//   barriers: full[s] lives in the leader (armed with the bytes of BOTH CTAs, completed by both CTAs' TMA),
//             empty[s] / tmem_full[a] exist in both CTAs and are signalled by multicast tcgen05.commit,
//             tmem_empty[a] lives in the leader and collects the 2 x 4 epilogue warps (remote arrive from the peer).
constexpr int PAIR_M = 256;
constexpr int PAIR_N = 256;
template <int BK_, int STAGES_> struct Cfg2 {
  static constexpr int BK = BK_;
  static constexpr int STAGES = STAGES_;
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = 128 * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// F8 = true: 8-bit float operands (K-major only).  Byte-wise the pipeline is identical to the 16-bit one — a 128-byte
// swizzle row holds 128 fp8 values and one tcgen05.mma consumes 32 of them (32 bytes, like 16 bf16) — so only the
// element-space K coordinates double and the MMA kind changes; the epilogue applies scale_a * scale_b.
template <int BK, int STAGES, bool F8 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_bh, const GemmParams p) {
  using C = Cfg2<BK, STAGES>;
  constexpr int KE = F8 ? 2 * BK : BK;      // K elements per stage
  constexpr int SUB = F8 ? 128 : 64;        // K elements per 128-byte swizzled sub-tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int num_pairs = gridDim.x >> 1;
  const int pair_id = blockIdx.x >> 1;
  const int m_blocks = (p.M + PAIR_M - 1) / PAIR_M;
  const int n_blocks = (p.N + PAIR_N - 1) / PAIR_N;
  const int num_tiles = m_blocks * n_blocks;
  const int k_blocks = (p.K + KE - 1) / KE;
  const int tail_half = p.tail_half;
  const int tail_split = (!tail_half && p.tail_split > 1) ? p.tail_split : 1;
  // tiles of the complete waves; the rest are "tail" tiles cut in N halves or K ranges
  const int full = (tail_half || tail_split > 1) ? (num_tiles / num_pairs) * num_pairs : num_tiles;
  const int num_units = full + (num_tiles - full) * (tail_half ? 2 : tail_split);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
    prefetch_tensormap(&tmap_bh);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);   // 4 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  cluster_sync();                      // both CTAs' barriers exist before anything remote touches them
  if (warp == 2) tmem_alloc_2cta<C::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = pair_id; unit < num_units; unit += num_pairs) {
      const WorkUnit wu = get_unit(unit, full, tail_split, k_blocks, tail_half);
      int m_blk, n_blk;
      tile_coords(wu.tile, m_blocks, n_blocks, m_blk, n_blk);
      const int m0 = m_blk * PAIR_M + (int)cta_rank * 128;
      const bool halfw = wu.half >= 0;                 // 256 x 128 half tile: this CTA stages 64 columns of B
      const int n0 = halfw ? n_blk * PAIR_N + wu.half * 128 + (int)cta_rank * 64 : n_blk * PAIR_N + (int)cta_rank * 128;
      for (int kb = wu.kb_begin; kb < wu.kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], halfw ? 2 * (C::A_BYTES + C::B_BYTES / 2) : 2 * C::STAGE_BYTES);
          const uint32_t fb = map_to_cta(smem_u32(&full_bar[stage]), 0);
          uint8_t* sa = smem_a + stage * C::A_BYTES;
          uint8_t* sb = smem_b + stage * C::B_BYTES;
          const int k0 = kb * KE;
          if (!p.a_mn_major) {
#pragma unroll
            for (int j = 0; j < BK / 64; ++j) tma_load_2d_2sm(&tmap_a, fb, sa + j * (128 * 128), k0 + j * SUB, m0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(&tmap_a, fb, sa + j * (BK * 128), m0 + j * 64, k0);
          }
          if (!p.b_mn_major) {
            // K-major B: 64-k sub-tiles of 128 rows (16 KB apart); a half tile fills the first 64 rows of each
#pragma unroll
            for (int j = 0; j < BK / 64; ++j)
              tma_load_2d_2sm(halfw ? &tmap_bh : &tmap_b, fb, sb + j * (128 * 128), k0 + j * SUB, n0);
          } else if (halfw) {
            tma_load_2d_2sm(&tmap_b, fb, sb, n0, k0);           // MN-major B: one 64-column chunk instead of two
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(&tmap_b, fb, sb + j * (BK * 128), n0 + j * 64, k0);
          }
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t a_kstep = p.a_mn_major ? UMMA_K * 128 : UMMA_K * 2;
      const uint32_t b_kstep = p.b_mn_major ? UMMA_K * 128 : UMMA_K * 2;
      for (int unit = pair_id; unit < num_units; unit += num_pairs) {
        const WorkUnit wu = get_unit(unit, full, tail_split, k_blocks, tail_half);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * PAIR_N;
        for (int kb = wu.kb_begin; kb < wu.kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem_a + stage * C::A_BYTES);
            const uint32_t sb = smem_u32(smem_b + stage * C::B_BYTES);
            const uint64_t da = p.a_mn_major ? make_smem_desc_sw128(sa, BK * 128, 1024)
                                             : make_smem_desc_sw128(sa, 16, 1024);
            const uint64_t db = p.b_mn_major ? make_smem_desc_sw128(sb, BK * 128, 1024)
                                             : make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // K-major: 64-wide swizzled sub-tiles of 128 rows (16 KB each), 4 k-steps of 32 B inside a sub-tile;
              // MN-major: one 16-k-row slab (2 KB) per step
              const uint32_t ao = p.a_mn_major ? k * a_kstep : (k >> 2) * (128 * 128) + (k & 3) * a_kstep;
              const uint32_t bo = p.b_mn_major ? k * b_kstep : (k >> 2) * (128 * 128) + (k & 3) * b_kstep;
              const uint32_t idesc = wu.half >= 0 ? p.idesc_half : p.idesc;
              if constexpr (F8)
                umma_f8_ss_2cta(tmem_d, advance_desc(da, ao), advance_desc(db, bo), idesc,
                                (kb > wu.kb_begin || k > 0) ? 1u : 0u);
              else
                umma_f16_ss_2cta(tmem_d, advance_desc(da, ao), advance_desc(db, bo), idesc,
                                 (kb > wu.kb_begin || k > 0) ? 1u : 0u);
            }
            umma_commit_2cta(&empty_bar[stage], 3);
            if (kb == wu.kb_end - 1) umma_commit_2cta(&tmem_full[acc], 3);
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (wu.kb_end <= wu.kb_begin && lane == 0) umma_commit_2cta(&tmem_full[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ================================================================ epilogue (warps 2..5, both CTAs)
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    float alpha = 1.0f;
    if constexpr (F8) {
      if (p.scale_a) alpha *= __ldg(p.scale_a);
      if (p.scale_b) alpha *= __ldg(p.scale_b);
    }
    for (int unit = pair_id; unit < num_units; unit += num_pairs) {
      const WorkUnit wu = get_unit(unit, full, tail_split, k_blocks, tail_half);
      int m_blk, n_blk;
      tile_coords(wu.tile, m_blocks, n_blocks, m_blk, n_blk);
      const int row = m_blk * PAIR_M + (int)cta_rank * 128 + quarter * 32 + lane;
      const int n0 = n_blk * PAIR_N + (wu.half > 0 ? 128 : 0);
      const int tile_n = wu.half >= 0 ? 128 : PAIR_N;
      // split tail tile: this warp's 32 x 256 slice of the fp32 partial of split j lives at part_base(j), its flag at
      // flag_base[j * 8]; writer and reader of a slice are the SAME (cta, quarter) warp position of two different pairs
      const int nparts = wu.nsplit - 1;
      float* part_row = nullptr;
      uint32_t* flag_base = nullptr;
      if (nparts > 0) {
        part_row = p.ws_part + ((size_t)wu.tail_idx * nparts * PAIR_M + (size_t)cta_rank * 128 + quarter * 32 + lane) * PAIR_N;
        flag_base = p.ws_flag + (size_t)wu.tail_idx * nparts * 8 + cta_rank * 4 + quarter;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * PAIR_N + ((uint32_t)(quarter * 32) << 16);
      const bool reader = wu.split == nparts;   // the LAST unit of a tile folds the partials in (its writers have lower
                                                // block indices, i.e. were dispatched no later than it)
      if (nparts > 0 && reader) {
        // partial sums of the other K ranges must have landed before they are folded in
        for (int j = 0; j < nparts; ++j)
          while (ld_acquire_gpu_u32(flag_base + j * 8) == 0u) __nanosleep(64);
      }
#pragma unroll 1
      for (int c = 0; c < tile_n; c += 64) {
        // two 32-column TMEM loads in flight per wait: halves the exposed tcgen05.ld latency of the epilogue
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(taddr + c, v0);
        tmem_ld_32x32b_x32(taddr + c + 32, v1);
        tmem_ld_wait();
        if (nparts > 0) {
          if (!reader) {
            // park this K range's partial (fp32, this thread's row, 64 columns) and move on: no store to C
            float* dst = part_row + (size_t)wu.split * PAIR_M * PAIR_N + c;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              *reinterpret_cast<uint4*>(dst + i) = make_uint4(v0[i], v0[i + 1], v0[i + 2], v0[i + 3]);
              *reinterpret_cast<uint4*>(dst + 32 + i) = make_uint4(v1[i], v1[i + 1], v1[i + 2], v1[i + 3]);
            }
            continue;
          }
          for (int j = 0; j < nparts; ++j) {
            const float* src = part_row + (size_t)j * PAIR_M * PAIR_N + c;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 a0 = __ldcg(reinterpret_cast<const float4*>(src + i));
              const float4 a1 = __ldcg(reinterpret_cast<const float4*>(src + 32 + i));
              v0[i] = __float_as_uint(__uint_as_float(v0[i]) + a0.x);
              v0[i + 1] = __float_as_uint(__uint_as_float(v0[i + 1]) + a0.y);
              v0[i + 2] = __float_as_uint(__uint_as_float(v0[i + 2]) + a0.z);
              v0[i + 3] = __float_as_uint(__uint_as_float(v0[i + 3]) + a0.w);
              v1[i] = __float_as_uint(__uint_as_float(v1[i]) + a1.x);
              v1[i + 1] = __float_as_uint(__uint_as_float(v1[i + 1]) + a1.y);
              v1[i + 2] = __float_as_uint(__uint_as_float(v1[i + 2]) + a1.z);
              v1[i + 3] = __float_as_uint(__uint_as_float(v1[i + 3]) + a1.w);
            }
          }
        }
        if constexpr (F8) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v0[i] = __float_as_uint(__uint_as_float(v0[i]) * alpha);
            v1[i] = __float_as_uint(__uint_as_float(v1[i]) * alpha);
          }
        }
#define CB_EPI_STORE(V, CC)                                                                              \
        {                                                                                                  \
          const int n_valid = p.N - (n0 + (CC));                                                           \
          if (row < p.M && n_valid > 0 && p.K > 0) {                                                       \
            const size_t off = (size_t)row * p.ldc + n0 + (CC);                                            \
            if (p.out_dtype == CB_BF16)                                                                    \
              store_row_chunk<__nv_bfloat16>((__nv_bfloat16*)p.C + off, V, n_valid, p.accumulate);         \
            else if (p.out_dtype == CB_F32)                                                                \
              store_row_chunk<float>((float*)p.C + off, V, n_valid, p.accumulate);                         \
            else                                                                                           \
              store_row_chunk<__half>((__half*)p.C + off, V, n_valid, p.accumulate);                       \
          }                                                                                                \
        }
        CB_EPI_STORE(v0, c)
        CB_EPI_STORE(v1, c + 32)
#undef CB_EPI_STORE
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[acc]), 0));
      }
      if (nparts > 0) {
        if (!reader) {
          __threadfence();             // every lane's partial rows are visible device-wide before the flag
          __syncwarp();
          if (lane == 0) st_release_gpu_u32(flag_base + wu.split * 8, 1u);
        } else {
          __syncwarp();                // all lanes are done reading the partials: re-arm the flags for the next launch
          if (lane < nparts) flag_base[lane * 8] = 0u;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  cluster_sync();                      // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 2) tmem_dealloc_2cta<C::TMEM_COLS>(tmem_base);
}

// Split-K workspace of the tail wave: one per stream that launches GEMMs (kernels on one stream are serialised, so a
// workspace is never shared by two running kernels).  74 fp32 tiles of 256 x 256 + their flags = 19.4 MB.
struct TailWorkspace {
  cudaStream_t stream;
  float* part;
  uint32_t* flag;
};
constexpr int MAX_TAIL_UNITS = 80;
inline TailWorkspace* get_tail_workspace(cudaStream_t stream) {
  static TailWorkspace slots[8] = {};
  static int used = 0;
  for (int i = 0; i < used; ++i)
    if (slots[i].stream == stream) return &slots[i];
  if (used == 8) return nullptr;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return nullptr;
  TailWorkspace w{stream, nullptr, nullptr};
  if (cudaMalloc(&w.part, (size_t)MAX_TAIL_UNITS * PAIR_M * PAIR_N * sizeof(float)) != cudaSuccess) return nullptr;
  if (cudaMalloc(&w.flag, (size_t)MAX_TAIL_UNITS * 8 * sizeof(uint32_t)) != cudaSuccess) return nullptr;
  cudaMemset(w.flag, 0, (size_t)MAX_TAIL_UNITS * 8 * sizeof(uint32_t));
  slots[used] = w;
  return &slots[used++];
}

// How many K ranges the tiles of the partial last wave are cut into (1 = no split).
inline int pick_tail_split(int tiles, int pairs, int k_blocks) {
  // CB200_GEMM_TAIL_SPLIT: 0 = off (default until it is a measured win), n > 0 = at most n K ranges per tail tile.
  // Read on every launch so one process can A/B the two schedules.
  const char* e = getenv("CB200_GEMM_TAIL_SPLIT");
  const int max_split = e ? atoi(e) : 0;
  if (max_split < 2 || tiles <= pairs) return 1;
  const int tail = tiles % pairs;
  if (tail == 0) return 1;
  int s = pairs / tail;                    // all units of the tail wave run concurrently on distinct pairs
  if (s > max_split) s = max_split;
  while (s > 1 && (2 * k_blocks) / (2 * s + 1) < 4) --s;   // keep >= 4 k-blocks (512 deep) per writer
  if (tail * s > MAX_TAIL_UNITS) return 1;
  return s < 2 ? 1 : s;
}

template <int BK, int STAGES>
int launch_2cta(const void* A, const void* B, void* Cp, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
                int b_mn, int in_dtype, int out_dtype, int accumulate, cudaStream_t stream) {
  using C = Cfg2<BK, STAGES>;
  CUtensorMap ta, tb;
  const bool bf16 = in_dtype == CB_BF16;
  int r;
  r = a_mn ? make_tmap_2d_16b(&ta, A, K, M, lda, BK, 64, bf16) : make_tmap_2d_16b(&ta, A, M, K, lda, 128, 64, bf16);
  if (r) return 1000 + r;
  r = b_mn ? make_tmap_2d_16b(&tb, B, K, N, ldb, BK, 64, bf16) : make_tmap_2d_16b(&tb, B, N, K, ldb, 128, 64, bf16);
  if (r) return 2000 + r;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.a_mn_major = a_mn; p.b_mn_major = b_mn; p.out_dtype = out_dtype;
  p.accumulate = accumulate; p.C = Cp;
  p.idesc = make_idesc_f16(PAIR_M, PAIR_N, bf16 ? 1 : 0, a_mn, b_mn);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<BK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = ((M + PAIR_M - 1) / PAIR_M) * ((N + PAIR_N - 1) / PAIR_N);
  int pairs = cb_num_sms() / 2;
  if (tiles < pairs) pairs = tiles;
  if (pairs <= 0) return 0;
  // tail wave: N halves (default) when the partial last wave fills at most half of the CTA pairs; K ranges on request
  CUtensorMap tbh = tb;
  {
    const char* e = getenv("CB200_GEMM_TAIL_HALF");
    const int want_half = e ? atoi(e) : 1;
    const int tail = tiles > pairs ? tiles % pairs : 0;
    if (want_half && tail > 0 && 2 * tail <= pairs) {
      if (!b_mn) {
        r = make_tmap_2d_16b(&tbh, B, N, K, ldb, 64, 64, bf16);      // K-major B: boxes of 64 rows for a half tile
        if (r) return 3000 + r;
      }
      p.tail_half = 1;
      p.idesc_half = make_idesc_f16(PAIR_M, 128, bf16 ? 1 : 0, a_mn, b_mn);
    }
  }
  if (!p.tail_half) {
    const int split = pick_tail_split(tiles, pairs, (K + BK - 1) / BK);
    if (split > 1) {
      if (TailWorkspace* w = get_tail_workspace(stream)) {
        p.tail_split = split; p.ws_part = w->part; p.ws_flag = w->flag;
      }
    }
  }
  gemm_tcgen05_2cta_kernel<BK, STAGES><<<2 * pairs, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, tbh, p);
  return (int)cudaGetLastError();
}

// fp8 (E4M3 / E5M2) x fp8 -> bf16 / fp16 / fp32, both operands K-major: A [M, K], B [N, K] bytes.
template <int BK, int STAGES>
int launch_2cta_f8(const void* A, const void* B, void* Cp, int M, int N, int K, int lda, int ldb, int ldc, int a_fmt,
                   int b_fmt, int out_dtype, int accumulate, const float* scale_a, const float* scale_b,
                   cudaStream_t stream) {
  using C = Cfg2<BK, STAGES>;
  CUtensorMap ta, tb;
  int r = make_tmap_2d_8b(&ta, A, M, K, lda, 128, 128);
  if (r) return 1000 + r;
  r = make_tmap_2d_8b(&tb, B, N, K, ldb, 128, 128);
  if (r) return 2000 + r;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.a_mn_major = 0; p.b_mn_major = 0; p.out_dtype = out_dtype;
  p.accumulate = accumulate; p.C = Cp; p.scale_a = scale_a; p.scale_b = scale_b;
  uint32_t d = make_idesc_f16(PAIR_M, PAIR_N, 0, 0, 0);
  d |= (uint32_t)(a_fmt & 7) << 7;              // 0 = E4M3, 1 = E5M2
  d |= (uint32_t)(b_fmt & 7) << 10;
  p.idesc = d;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<BK, STAGES, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = ((M + PAIR_M - 1) / PAIR_M) * ((N + PAIR_N - 1) / PAIR_N);
  int pairs = cb_num_sms() / 2;
  if (tiles < pairs) pairs = tiles;
  if (pairs <= 0) return 0;
  gemm_tcgen05_2cta_kernel<BK, STAGES, true><<<2 * pairs, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, tb, p);
  return (int)cudaGetLastError();
}

template <int BLOCK_N>
int launch(const void* A, const void* B, void* Cp, int M, int N, int K, int lda, int ldb, int ldc, int a_mn, int b_mn,
           int in_dtype, int out_dtype, int accumulate, cudaStream_t stream) {
  using C = Cfg<BLOCK_N>;
  CUtensorMap ta, tb;
  const bool bf16 = in_dtype == CB_BF16;
  int r;
  // K-major operand: tensor [rows = M/N, cols = K], box [BLOCK_M/N, 64].  MN-major: tensor [rows = K, cols = M/N],
  // box [BLOCK_K, 64].
  r = a_mn ? make_tmap_2d_16b(&ta, A, K, M, lda, BLOCK_K, 64, bf16) : make_tmap_2d_16b(&ta, A, M, K, lda, BLOCK_M, 64, bf16);
  if (r) return 1000 + r;
  r = b_mn ? make_tmap_2d_16b(&tb, B, K, N, ldb, BLOCK_K, 64, bf16) : make_tmap_2d_16b(&tb, B, N, K, ldb, BLOCK_N, 64, bf16);
  if (r) return 2000 + r;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.a_mn_major = a_mn; p.b_mn_major = b_mn; p.out_dtype = out_dtype;
  p.accumulate = accumulate; p.C = Cp;
  p.idesc = make_idesc_f16(BLOCK_M, BLOCK_N, bf16 ? 1 : 0, a_mn, b_mn);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
  int grid = cb_num_sms();
  if (tiles < grid) grid = tiles;
  if (grid <= 0) return 0;
  gemm_tcgen05_kernel<BLOCK_N><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

// D[M,N] (+)= A x B^T-like contraction over K.
//   a_mn_major == 0: A is [M, K] row-major with leading dimension lda;  == 1: A is stored [K, M] (lda between K rows)
//   b_mn_major == 0: B is [N, K] row-major with leading dimension ldb;  == 1: B is stored [K, N]
// Requirements: 16-byte aligned base pointers, leading dimensions multiples of 8 elements, ldc multiple of 8 (4 for
// fp32 output).  M, N, K are otherwise arbitrary (TMA zero-fills out-of-bounds, stores are predicated).
int cb_gemm_tcgen05(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                    int a_mn_major, int b_mn_major, int in_dtype, int out_dtype, int accumulate, int block_n,
                    cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  if (in_dtype != CB_BF16 && in_dtype != CB_F16) return (int)cudaErrorInvalidValue;
  static int use_2cta = -1;
  if (use_2cta < 0) {
    const char* e = getenv("CB200_GEMM_2CTA");
    use_2cta = e ? atoi(e) : 1;
  }
  // CTA-pair kernel: 256 x 256 tiles; worth it once there are enough tiles to fill the 74 pairs
  if (block_n == 512 || (block_n == 0 && use_2cta && M >= 256 && N >= 256 &&
                         ((M + 255) / 256) * ((N + 255) / 256) >= cb_num_sms() / 2))
  {
    // main-loop shape of the CTA-pair kernel: 2 = BK128 x 3 stages (default: one barrier round trip per 128-wide
    // k-block; measured 1.00-1.21x cuBLAS on the Llama-3 shapes), 0 = BK64 x 6, 1 = BK64 x 7
    static int variant = -1;
    if (variant < 0) {
      const char* e = getenv("CB200_GEMM_2CTA_VARIANT");
      variant = e ? atoi(e) : 2;
    }
    if (variant == 1)
      return launch_2cta<64, 7>(A, B, C, M, N, K, lda, ldb, ldc, a_mn_major, b_mn_major, in_dtype, out_dtype, accumulate, stream);
    if (variant == 2)
      return launch_2cta<128, 3>(A, B, C, M, N, K, lda, ldb, ldc, a_mn_major, b_mn_major, in_dtype, out_dtype, accumulate, stream);
    return launch_2cta<64, 6>(A, B, C, M, N, K, lda, ldb, ldc, a_mn_major, b_mn_major, in_dtype, out_dtype, accumulate, stream);
  }
  if (block_n == 0) {
    const int tiles256 = ((M + 127) / 128) * ((N + 255) / 256);
    block_n = (tiles256 >= cb_num_sms() || N % 256 == 0 && tiles256 * 2 > cb_num_sms() * 3) ? 256 : 128;
    if (N <= 128) block_n = 128;
  }
  if (block_n == 256)
    return launch<256>(A, B, C, M, N, K, lda, ldb, ldc, a_mn_major, b_mn_major, in_dtype, out_dtype, accumulate, stream);
  return launch<128>(A, B, C, M, N, K, lda, ldb, ldc, a_mn_major, b_mn_major, in_dtype, out_dtype, accumulate, stream);
}

// D[M,N] (+)= scale_a * scale_b * (A[M,K] x B[N,K]^T) with 8-bit float operands on the CTA-pair tcgen05 kernel
// (kind::f8f6f4, 256x256 tiles, 256 K-elements per stage).  a_fmt / b_fmt: 0 = E4M3, 1 = E5M2.  Requirements:
// 16-byte aligned bases, lda / ldb multiples of 16 bytes.  scale pointers are device fp32 scalars or null.
int cb_gemm_fp8_tcgen05(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                        int a_fmt, int b_fmt, int out_dtype, int accumulate, const float* scale_a,
                        const float* scale_b, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  if ((lda & 15) || (ldb & 15)) return (int)cudaErrorInvalidValue;
  return launch_2cta_f8<128, 3>(A, B, C, M, N, K, lda, ldb, ldc, a_fmt, b_fmt, out_dtype, accumulate, scale_a,
                                scale_b, stream);
}

int cb_gemm_smem_bytes(int block_n) { return block_n == 256 ? Cfg<256>::SMEM_BYTES : Cfg<128>::SMEM_BYTES; }

}  // extern "C"
