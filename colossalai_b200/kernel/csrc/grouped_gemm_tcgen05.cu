// Grouped (per-expert) tcgen05 GEMM for sm_100a: the three GEMMs of a mixture-of-experts MLP over token rows that are
// grouped by expert, with the group boundaries read ON THE DEVICE (no host sync, no per-expert launches).
//
//   MODE_NT  (forward)  Y[rows of e, N]  = X[rows of e, K] x W[e][N, K]^T         variable M per group
//   MODE_NN  (dgrad)    dX[rows of e, N] = dY[rows of e, K] x W[e][K, N]           variable M per group
//   MODE_TN  (wgrad)    dW[e][M, N]     (+)= dY[rows of e, M]^T x X[rows of e, N]  variable K (= rows) per group
//
// `offs[e]` is the exclusive end row of group e (cumulative counts, int32, device memory).  One persistent launch of CTA
// pairs (cta_group::2, 256 x 256 tiles, BK = 128 x 3 stages - the main loop of gemm_tcgen05.cu); every role warp walks
// the same tile list, which is built from `offs` at kernel start:
//   * variable-M modes: group e contributes ceil(rows_e / 256) row tiles x ceil(N / 256) column tiles, rasterised inside
//     the group in bands of 8 row tiles (L2 reuse of the expert's weight tile).  A row tile that runs past the end of its
//     group simply loads rows of the next group (or TMA zero fill) and masks them at the store - no padding needed;
//   * variable-K mode: every group owns (M / 256) x (N / 256) tiles whose reduction runs over the group's token rows; the
//     row range of a group must start and end on a multiple of 128 (one k-block) - the dispatch lays groups out that way
//     (`moe/grouped_gemm.py` pads with zero rows otherwise) - so a k-block never mixes two experts.
//
// Replaces the reference's python loop over local experts (`shardformer/modeling/mixtral.py:177-191`) and round 1's
// library `torch._grouped_mm`.
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

using namespace sm100;

namespace {

constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int GROUP_M = 8;
constexpr int PAIR_M = 256;
constexpr int PAIR_N = 256;
constexpr int BK = 128;
constexpr int STAGES = 3;
constexpr int MAX_GROUPS = 512;
constexpr int A_BYTES = 128 * BK * 2;
constexpr int B_BYTES = 128 * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + (MAX_GROUPS + 1) * 4;

enum { MODE_NT = 0, MODE_NN = 1, MODE_TN = 2 };

struct GroupedParams {
  const int* offs;          // [E] cumulative row ends
  int E;
  int total_rows;           // rows of the grouped operand(s) (buffer size; >= offs[E-1])
  int M, N, K;              // NT / NN: N = output columns, K = reduction; TN: M x N = per-expert output, K unused
  int ldc;
  long long c_group_stride; // TN: elements between the outputs of consecutive experts
  int b_group_rows;         // NT: N rows per expert in the [E * N, K] view of W; NN: K rows per expert in [E * K, N]
  int out_dtype;
  int accumulate;
  uint32_t idesc;
  void* C;
};

struct Tile {
  int group, m_blk, n_blk;
  int row0, row_end;        // variable-M: first row of the tile / end row of its group;  TN: k range [row0, row_end)
};

SM100_DEVICE void raster(int local, int m_blocks, int n_blocks, int& m_blk, int& n_blk) {
  const int per_band = GROUP_M * n_blocks;
  const int band = local / per_band;
  const int first_m = band * GROUP_M;
  const int gsize = min(m_blocks - first_m, GROUP_M);
  const int in_band = local - band * per_band;
  m_blk = first_m + in_band % gsize;
  n_blk = in_band / gsize;
}

// tile_start[e] = number of row tiles of groups < e (variable-M modes).  `cur` caches the last group found: tiles are
// visited in increasing order by every role, so the search is amortised O(1).
template <int MODE>
SM100_DEVICE Tile find_tile(int t, const GroupedParams& p, const int* tile_start, int n_blocks, int& cur) {
  Tile r;
  if (MODE == MODE_TN) {
    const int per_group = ((p.M + PAIR_M - 1) / PAIR_M) * n_blocks;
    r.group = t / per_group;
    raster(t - r.group * per_group, (p.M + PAIR_M - 1) / PAIR_M, n_blocks, r.m_blk, r.n_blk);
    r.row0 = r.group == 0 ? 0 : p.offs[r.group - 1];
    r.row_end = p.offs[r.group];
    return r;
  }
  while (cur + 1 < p.E && t >= tile_start[cur + 1] * n_blocks) ++cur;
  r.group = cur;
  const int g0 = cur == 0 ? 0 : p.offs[cur - 1];
  r.row_end = p.offs[cur];
  raster(t - tile_start[cur] * n_blocks, tile_start[cur + 1] - tile_start[cur], n_blocks, r.m_blk, r.n_blk);
  r.row0 = g0 + r.m_blk * PAIR_M;
  return r;
}

template <typename T16>
SM100_DEVICE void store16(T16* __restrict__ dst, const uint32_t (&acc)[32], int n_valid, bool accumulate) {
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
SM100_DEVICE void store32(float* __restrict__ dst, const uint32_t (&acc)[32], int n_valid, bool accumulate) {
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < n_valid) dst[j] = __uint_as_float(acc[j]) + (accumulate ? dst[j] : 0.f);
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
grouped_gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GroupedParams p) {
  constexpr bool A_MN = MODE == MODE_TN;
  constexpr bool B_MN = MODE != MODE_NT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  int* tile_start = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(bars) + 256);     // [E + 1]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int num_pairs = gridDim.x >> 1;
  const int pair_id = blockIdx.x >> 1;
  const int n_blocks = (p.N + PAIR_N - 1) / PAIR_N;

  // ---- tile list: per-group row-tile counts -> exclusive prefix (every CTA builds its own copy; E is small)
  int num_tiles;
  if (MODE == MODE_TN) {
    num_tiles = p.E * ((p.M + PAIR_M - 1) / PAIR_M) * n_blocks;
  } else {
    if (threadIdx.x == 0) {
      int acc = 0, prev = 0;
      for (int e = 0; e < p.E; ++e) {
        tile_start[e] = acc;
        const int end = p.offs[e];
        acc += (end - prev + PAIR_M - 1) / PAIR_M;
        prev = end;
      }
      tile_start[p.E] = acc;
    }
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  cluster_sync();
  if (warp == 2) tmem_alloc_2cta<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (MODE != MODE_TN) num_tiles = tile_start[p.E] * n_blocks;

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs)
    int stage = 0, cur = 0;
    uint32_t phase = 0;
    for (int t = pair_id; t < num_tiles; t += num_pairs) {
      const Tile tl = find_tile<MODE>(t, p, tile_start, n_blocks, cur);
      int k_blocks, kbase;
      if (MODE == MODE_TN) { k_blocks = (tl.row_end - tl.row0 + BK - 1) / BK; kbase = tl.row0; }
      else { k_blocks = (p.K + BK - 1) / BK; kbase = 0; }
      // A: NT / NN rows of the grouped activations; TN: columns m of dY (MN-major, rows = tokens)
      const int a_m0 = (MODE == MODE_TN ? tl.m_blk * PAIR_M : tl.row0) + (int)cta_rank * 128;
      const int b_n0 = tl.n_blk * PAIR_N + (int)cta_rank * 128;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const uint32_t fb = map_to_cta(smem_u32(&full_bar[stage]), 0);
          uint8_t* sa = smem_a + stage * A_BYTES;
          uint8_t* sb = smem_b + stage * B_BYTES;
          const int k0 = kbase + kb * BK;
          if (!A_MN) {
#pragma unroll
            for (int j = 0; j < BK / 64; ++j) tma_load_2d_2sm(&tmap_a, fb, sa + j * (128 * 128), k0 + j * 64, a_m0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(&tmap_a, fb, sa + j * (BK * 128), a_m0 + j * 64, k0);
          }
          if (MODE == MODE_NT) {          // W viewed as [E * N, K]: rows of expert e start at e * N
#pragma unroll
            for (int j = 0; j < BK / 64; ++j)
              tma_load_2d_2sm(&tmap_b, fb, sb + j * (128 * 128), k0 + j * 64, tl.group * p.b_group_rows + b_n0);
          } else if (MODE == MODE_NN) {   // W viewed as [E * K, N]: reduction rows of expert e start at e * K
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_2d_2sm(&tmap_b, fb, sb + j * (BK * 128), b_n0 + j * 64, tl.group * p.b_group_rows + k0);
          } else {                        // X [rows, N] MN-major, rows = tokens
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(&tmap_b, fb, sb + j * (BK * 128), b_n0 + j * 64, k0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA)
    if (leader) {
      int stage = 0, cur = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      constexpr uint32_t a_kstep = A_MN ? UMMA_K * 128 : UMMA_K * 2;
      constexpr uint32_t b_kstep = B_MN ? UMMA_K * 128 : UMMA_K * 2;
      for (int t = pair_id; t < num_tiles; t += num_pairs) {
        const Tile tl = find_tile<MODE>(t, p, tile_start, n_blocks, cur);
        const int k_blocks = MODE == MODE_TN ? (tl.row_end - tl.row0 + BK - 1) / BK : (p.K + BK - 1) / BK;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * PAIR_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem_a + stage * A_BYTES);
            const uint32_t sb = smem_u32(smem_b + stage * B_BYTES);
            const uint64_t da = A_MN ? make_smem_desc_sw128(sa, BK * 128, 1024) : make_smem_desc_sw128(sa, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sb, BK * 128, 1024) : make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint32_t ao = A_MN ? k * a_kstep : (k >> 2) * (128 * 128) + (k & 3) * a_kstep;
              const uint32_t bo = B_MN ? k * b_kstep : (k >> 2) * (128 * 128) + (k & 3) * b_kstep;
              umma_f16_ss_2cta(tmem_d, advance_desc(da, ao), advance_desc(db, bo), p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2cta(&empty_bar[stage], 3);
            if (kb == k_blocks - 1) umma_commit_2cta(&tmem_full[acc], 3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (k_blocks == 0 && lane == 0) umma_commit_2cta(&tmem_full[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ================================================================ epilogue (warps 2..5, both CTAs)
    const int quarter = warp & 3;
    int acc = 0, cur = 0;
    uint32_t acc_phase = 0;
    for (int t = pair_id; t < num_tiles; t += num_pairs) {
      const Tile tl = find_tile<MODE>(t, p, tile_start, n_blocks, cur);
      const int in_tile = (int)cta_rank * 128 + quarter * 32 + lane;
      int row, row_lim;
      size_t base;
      bool empty_k = false;
      if (MODE == MODE_TN) {
        row = tl.m_blk * PAIR_M + in_tile; row_lim = p.M;
        base = (size_t)tl.group * (size_t)p.c_group_stride + (size_t)row * p.ldc;
        empty_k = tl.row_end <= tl.row0;             // an expert without tokens: its gradient is zero
      } else {
        row = tl.row0 + in_tile; row_lim = tl.row_end;
        base = (size_t)row * p.ldc;
      }
      const int n0 = tl.n_blk * PAIR_N;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * PAIR_N + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < PAIR_N; c += 64) {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(taddr + c, v0);
        tmem_ld_32x32b_x32(taddr + c + 32, v1);
        tmem_ld_wait();
        if (empty_k) {
#pragma unroll
          for (int i = 0; i < 32; ++i) { v0[i] = 0u; v1[i] = 0u; }
        }
#define CB_G_STORE(V, CC)                                                                              \
        {                                                                                              \
          const int n_valid = p.N - (n0 + (CC));                                                       \
          if (row < row_lim && n_valid > 0) {                                                          \
            const size_t off = base + n0 + (CC);                                                       \
            if (p.out_dtype == CB_BF16) store16<__nv_bfloat16>((__nv_bfloat16*)p.C + off, V, n_valid, p.accumulate); \
            else if (p.out_dtype == CB_F32) store32((float*)p.C + off, V, n_valid, p.accumulate);      \
            else store16<__half>((__half*)p.C + off, V, n_valid, p.accumulate);                       \
          }                                                                                            \
        }
        CB_G_STORE(v0, c)
        CB_G_STORE(v1, c + 32)
#undef CB_G_STORE
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
  cluster_sync();
  if (warp == 2) tmem_dealloc_2cta<512>(tmem_base);
}

template <int MODE>
int launch_grouped(const CUtensorMap& ta, const CUtensorMap& tb, const GroupedParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(grouped_gemm_2cta_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int pairs = cb_num_sms() / 2;      // the number of tiles is only known on the device: full persistent grid
  grouped_gemm_2cta_kernel<MODE><<<2 * pairs, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, p);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

// Variable-M grouped GEMM.  x [total_rows, K] (rows grouped by expert, `offs` = cumulative row ends on the device);
// transpose_b != 0: w [E, N, K] and y = x w[e]^T (forward);  transpose_b == 0: w [E, K, N] and y = x w[e] (dgrad with
// the forward weight [E, K(out), N(in)]).  y [total_rows, N].  Requirements: contiguous w, K % 8 == 0, N % 8 == 0,
// 16-byte aligned pointers, E <= 512; transpose_b == 0 additionally needs K % 128 == 0 (a k-block must not straddle
// two experts' weights).
int cb_grouped_gemm(const void* x, const void* w, void* y, const int* offs, int E, int total_rows, int N, int K, int ldx,
                    int ldy, int transpose_b, int dtype, cudaStream_t stream) {
  if (E <= 0 || total_rows <= 0 || N <= 0) return 0;
  if (E > MAX_GROUPS || (dtype != CB_BF16 && dtype != CB_F16) || (K & 7) || (N & 7)) return (int)cudaErrorInvalidValue;
  if (!transpose_b && (K % BK)) return (int)cudaErrorInvalidValue;
  const bool bf16 = dtype == CB_BF16;
  CUtensorMap ta, tb;
  int r = make_tmap_2d_16b(&ta, x, total_rows, K, ldx, 128, 64, bf16);
  if (r) return 1000 + r;
  GroupedParams p{};
  p.offs = offs; p.E = E; p.total_rows = total_rows; p.N = N; p.K = K; p.ldc = ldy; p.out_dtype = dtype; p.C = y;
  if (transpose_b) {
    r = make_tmap_2d_16b(&tb, w, (uint64_t)E * N, K, K, 128, 64, bf16);
    if (r) return 2000 + r;
    p.b_group_rows = N;
    p.idesc = make_idesc_f16(PAIR_M, PAIR_N, bf16 ? 1 : 0, 0, 0);
    return launch_grouped<MODE_NT>(ta, tb, p, stream);
  }
  r = make_tmap_2d_16b(&tb, w, (uint64_t)E * K, N, N, BK, 64, bf16);
  if (r) return 2000 + r;
  p.b_group_rows = K;
  p.idesc = make_idesc_f16(PAIR_M, PAIR_N, bf16 ? 1 : 0, 0, 1);
  return launch_grouped<MODE_NN>(ta, tb, p, stream);
}

// Variable-K grouped GEMM (weight gradients): dw[e] [M, N] (+)= dy[rows of e, M]^T x[rows of e, N].  Every group's row
// range must start and end on a multiple of 128.  out_dtype: bf16 / fp16 / fp32.
int cb_grouped_gemm_wgrad(const void* dy, const void* x, void* dw, const int* offs, int E, int total_rows, int M, int N,
                          int lddy, int ldx, int in_dtype, int out_dtype, int accumulate, cudaStream_t stream) {
  if (E <= 0 || M <= 0 || N <= 0) return 0;
  if (E > MAX_GROUPS || (in_dtype != CB_BF16 && in_dtype != CB_F16) || (M & 7) || (N & 7)) return (int)cudaErrorInvalidValue;
  const bool bf16 = in_dtype == CB_BF16;
  CUtensorMap ta, tb;
  int r = make_tmap_2d_16b(&ta, dy, total_rows, M, lddy, BK, 64, bf16);
  if (r) return 1000 + r;
  r = make_tmap_2d_16b(&tb, x, total_rows, N, ldx, BK, 64, bf16);
  if (r) return 2000 + r;
  GroupedParams p{};
  p.offs = offs; p.E = E; p.total_rows = total_rows; p.M = M; p.N = N; p.K = 0; p.ldc = N;
  p.c_group_stride = (long long)M * N; p.out_dtype = out_dtype; p.accumulate = accumulate; p.C = dw;
  p.idesc = make_idesc_f16(PAIR_M, PAIR_N, bf16 ? 1 : 0, 1, 1);
  return launch_grouped<MODE_TN>(ta, tb, p, stream);
}

}  // extern "C"
