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
};

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
                         const GemmParams p) {
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
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
      int m_blk, n_blk;
      tile_coords(tile, m_blocks, n_blocks, m_blk, n_blk);
      const int m0 = m_blk * PAIR_M + (int)cta_rank * 128;
      const int n0 = n_blk * PAIR_N + (int)cta_rank * 128;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
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
#pragma unroll
            for (int j = 0; j < BK / 64; ++j) tma_load_2d_2sm(&tmap_b, fb, sb + j * (128 * 128), k0 + j * SUB, n0);
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
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * PAIR_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
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
              if constexpr (F8)
                umma_f8_ss_2cta(tmem_d, advance_desc(da, ao), advance_desc(db, bo), p.idesc,
                                (kb > 0 || k > 0) ? 1u : 0u);
              else
                umma_f16_ss_2cta(tmem_d, advance_desc(da, ao), advance_desc(db, bo), p.idesc,
                                 (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2cta(&empty_bar[stage], 3);
            if (kb == k_blocks - 1) umma_commit_2cta(&tmem_full[acc], 3);
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (k_blocks == 0 && lane == 0) umma_commit_2cta(&tmem_full[acc], 3);
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
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
      int m_blk, n_blk;
      tile_coords(tile, m_blocks, n_blocks, m_blk, n_blk);
      const int row = m_blk * PAIR_M + (int)cta_rank * 128 + quarter * 32 + lane;
      const int n0 = n_blk * PAIR_N;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * PAIR_N + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < PAIR_N; c += 64) {
        // two 32-column TMEM loads in flight per wait: halves the exposed tcgen05.ld latency of the epilogue
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(taddr + c, v0);
        tmem_ld_32x32b_x32(taddr + c + 32, v1);
        tmem_ld_wait();
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
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  cluster_sync();                      // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 2) tmem_dealloc_2cta<C::TMEM_COLS>(tmem_base);
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
  gemm_tcgen05_2cta_kernel<BK, STAGES><<<2 * pairs, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
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
  gemm_tcgen05_2cta_kernel<BK, STAGES, true><<<2 * pairs, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
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
