// Flash attention forward for sm_100a on the 5th-generation tensor cores.
//
//   O = softmax(scale * Q K^T + causal mask) V,   LSE = log sum exp     (GQA: Hq = g * Hkv, head_dim 64 or 128)
//
// One CTA = one 128-row query tile of one (batch, head).  Six warps:
//   warp 0      TMA producer: Q once, then (K_j, V_j) tiles of 128 keys through a 2-stage ring (SWIZZLE_128B boxes)
//   warp 1      MMA issuer (one elected thread): S_j = Q K_j^T and O_j = P_j V_j with tcgen05.mma, accumulators in TMEM
//   warps 2..5  softmax: thread r owns query row r — tcgen05.ld of its S row, running max / sum in the exp2 domain,
//               P_j written as bf16 into shared memory in the K-major 128B-swizzled layout the tensor core reads,
//               O_j pulled back from TMEM and folded into the register accumulator with the rescale factor
// S is double-buffered in TMEM so Q K_{j+1}^T runs on the tensor pipe while the softmax warps work on tile j.
// TMEM map (512 columns allocated): [0,128) S buffer 0, [128,256) S buffer 1, [256,256+D) O_j.
// Shared memory: Q 128 x D, 2 x (K 128 x D + V 128 x D), P 128 x 128 (all bf16/fp16) + mbarriers = 193 KB at D = 128.
//
// STATUS: written against the validated primitives of sm100.cuh / gemm_tcgen05.cu (descriptors, TMA boxes, TMEM
// loads) but NOT yet executed on hardware — `ops/flash_attn_native.py` keeps it behind CB200_FLASH_NATIVE=1 and the
// numerics test (tests/test_kernels/test_flash_attn_native.py) must pass on a B200 before it becomes a default.
//
// Reference being replaced: flash-attn 2 `flash_attn_func` / `_flash_attn_varlen_forward` (mma.sync kernels) used by
// `colossalai/shardformer/layer/attn.py:356-372,720-755`.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "sm100.cuh"

namespace {
using namespace sm100;

constexpr int BLOCK_Q = 128;
constexpr int BLOCK_KV = 128;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int KV_STAGES = 2;

template <int D> struct FaCfg {
  static constexpr int Q_BYTES = BLOCK_Q * D * 2;
  static constexpr int K_BYTES = BLOCK_KV * D * 2;
  static constexpr int V_BYTES = BLOCK_KV * D * 2;
  static constexpr int P_BYTES = BLOCK_Q * BLOCK_KV * 2;
  static constexpr int KV_STAGE_BYTES = K_BYTES + V_BYTES;
  static constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * KV_STAGE_BYTES + 2 * P_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 512;
  static constexpr int O_COL = 256;
};

struct FaParams {
  // packed / variable-length batches: cu_seqlens[b] .. cu_seqlens[b + 1] are the token rows of sequence b (self
  // attention, the same boundaries for queries and keys); null = `batch` equal-length sequences
  const int* cu_seqlens;
  int batch, seqlen_q, seqlen_k;     // equal-length sequences, token-major tensors [batch * seqlen, heads * D]
  int hq, hkv;
  int causal;
  float scale_log2;                  // softmax scale * log2(e)
  void* out;                         // [batch * seqlen_q, hq * D]
  float* lse;                        // [batch * seqlen_q, hq] natural-log LSE
  // ring / blockwise attention: the online-softmax state of every query row lives in (o_state, lse) between launches.
  // has_prev: start from that state (m = lse, l = 1, acc = o_state) instead of (-inf, 0, 0); the result - the merge of
  // the previous partial result with this launch's keys - is written back in fp32.  No separate rescale / merge pass.
  float* o_state;                    // [rows, hq * D] fp32, or null (plain attention: `out` in the input dtype)
  int has_prev;
  int out_dtype;
  uint32_t idesc_qk, idesc_pv;
  // EXTRA variant of the kernel only (inference prefill of sliding-window / ALiBi models):
  int window;                        // > 0: query q sees keys (q - window, q]; key tiles left of every window are skipped
  const float* alibi;                // [hq] slopes (score += slope * (key - query)), or null
};

SM100_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D, bool EXTRA>
__global__ void __launch_bounds__(NUM_THREADS, 1)
flash_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const FaParams p) {
  using C = FaCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_kv = smem + C::Q_BYTES;
  uint8_t* smem_p = smem_kv + KV_STAGES * C::KV_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + 2 * C::P_BYTES);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* kv_full = bars + 1;            // [2]
  uint64_t* kv_empty = bars + 3;           // [2]
  uint64_t* s_full = bars + 5;             // [2]
  uint64_t* s_empty = bars + 7;            // [2]  4 softmax warps
  uint64_t* p_full = bars + 9;             // [1]  4 softmax warps
  uint64_t* o_full = bars + 10;            // [1]
  uint64_t* o_empty = bars + 11;           // [1]  4 softmax warps
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int kv_head = head / (p.hq / p.hkv);
  int q_tile, seq_row0, len_q, len_k;
  if (p.cu_seqlens != nullptr) {
    // packed batch: blockIdx.x enumerates (sequence, query tile) pairs in order; surplus CTAs (the grid is sized for
    // the worst case without knowing the lengths on the host) leave at once
    int t = (int)blockIdx.x, bsel = -1, start = 0, len = 0;
    for (int bb = 0; bb < p.batch; ++bb) {
      start = p.cu_seqlens[bb];
      len = p.cu_seqlens[bb + 1] - start;
      const int nt = (len + BLOCK_Q - 1) / BLOCK_Q;
      if (t < nt) { bsel = bb; break; }
      t -= nt;
    }
    if (bsel < 0) return;                      // uniform across the CTA: nothing has been allocated yet
    q_tile = t; seq_row0 = start; len_q = len; len_k = len;
  } else {
    // heaviest (longest causal) query tiles first
    const int q_tiles = (p.seqlen_q + BLOCK_Q - 1) / BLOCK_Q;
    q_tile = p.causal ? (q_tiles - 1 - (int)blockIdx.x) : (int)blockIdx.x;
    seq_row0 = -1; len_q = p.seqlen_q; len_k = p.seqlen_k;
  }
  const int b = blockIdx.z;
  const int q_row0 = (seq_row0 >= 0 ? seq_row0 : b * p.seqlen_q) + q_tile * BLOCK_Q;   // row in the token-major Q / O
  const int kv_row0 = seq_row0 >= 0 ? seq_row0 : b * p.seqlen_k;
  const int kv_tiles_all = (len_k + BLOCK_KV - 1) / BLOCK_KV;
  // causal with seqlen_q == seqlen_k: key tiles 0..q_tile; otherwise all of them
  const int kv_last = p.causal ? min(q_tile + 1, kv_tiles_all) : kv_tiles_all;   // one past the last visible key tile
  // sliding window: the lowest key any row of this query tile can see is (q_tile * BLOCK_Q - window + 1)
  const int kv_first = (EXTRA && p.window > 0) ? max(0, q_tile * BLOCK_Q - p.window + 1) / BLOCK_KV : 0;
  const int kv_tiles = kv_last - kv_first;         // tiles walked; loop index j <-> key tile kv_first + j
  const bool ragged_k = (len_k % BLOCK_KV) != 0;   // the last key tile runs past the end of the sequence
  const bool resume = p.o_state != nullptr && p.has_prev != 0;   // block mode: continue from the stored state

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_k);
    prefetch_tensormap(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int h = 0; h < D / 64; ++h)                     // boxes [128 rows, 64 d] -> K-major swizzled sub-tiles
        tma_load_2d(&tmap_q, q_full, smem_q + h * (BLOCK_Q * 128), head * D + h * 64, q_row0);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < kv_tiles; ++j) {
      mbar_wait(&kv_empty[stage], phase ^ 1);
      if (lane == 0) {
        uint8_t* sk = smem_kv + stage * C::KV_STAGE_BYTES;
        uint8_t* sv = sk + C::K_BYTES;
        mbar_arrive_expect_tx(&kv_full[stage], C::KV_STAGE_BYTES);
        const int row = kv_row0 + (kv_first + j) * BLOCK_KV;
#pragma unroll
        for (int h = 0; h < D / 64; ++h) {
          tma_load_2d(&tmap_k, &kv_full[stage], sk + h * (BLOCK_KV * 128), kv_head * D + h * 64, row);   // K-major B
          tma_load_2d(&tmap_v, &kv_full[stage], sv + h * (BLOCK_KV * 128), kv_head * D + h * 64, row);   // MN-major B
        }
      }
      __syncwarp();
      if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t sq = smem_u32(smem_q);
    const uint32_t sp = smem_u32(smem_p);
    const uint64_t dq = make_smem_desc_sw128(sq, 16, 1024);                       // K-major, K = D
    const uint64_t dp = make_smem_desc_sw128(sp, 16, 1024);                       // K-major, K = keys
    auto issue_qk = [&](int j) {
      const int stage = j % KV_STAGES;
      const int sb = j & 1;
      mbar_wait(&kv_full[stage], (uint32_t)((j / KV_STAGES) & 1));
      mbar_wait(&s_empty[sb], (uint32_t)(((j >> 1) & 1) ^ 1));
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sk = smem_u32(smem_kv + stage * C::KV_STAGE_BYTES);
        const uint64_t dk = make_smem_desc_sw128(sk, 16, 1024);
#pragma unroll
        for (int k = 0; k < D / UMMA_K; ++k) {
          const uint32_t off = (k >> 2) * (128 * 128) + (k & 3) * (UMMA_K * 2);
          umma_f16_ss(tmem_base + sb * BLOCK_KV, advance_desc(dq, off), advance_desc(dk, off), p.idesc_qk,
                      k > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[sb]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    tc_fence_after();
    issue_qk(0);
    for (int j = 0; j < kv_tiles; ++j) {
      if (j + 1 < kv_tiles) issue_qk(j + 1);            // tensor pipe works on the next scores during softmax(j)
      const int stage = j % KV_STAGES;
      mbar_wait(p_full, (uint32_t)(j & 1));             // P_j is in shared memory (and O was rescaled if it had to be)
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sv = smem_u32(smem_kv + stage * C::KV_STAGE_BYTES + C::K_BYTES);
        const uint64_t dv = make_smem_desc_sw128(sv, BLOCK_KV * 128, 1024);        // MN-major: [keys, D] row-major
        const uint64_t dpj = advance_desc(dp, (uint32_t)(j & 1) * C::P_BYTES);
        // O accumulates in TMEM over all key tiles (the first MMA of the first tile overwrites, unless the
        // accumulator was pre-loaded with the state of earlier launches)
        const bool fresh = (j == 0) && !resume;
#pragma unroll
        for (int k = 0; k < BLOCK_KV / UMMA_K; ++k) {
          const uint32_t ao = (k >> 2) * (128 * 128) + (k & 3) * (UMMA_K * 2);
          const uint32_t bo = k * (UMMA_K * 128);
          umma_f16_ss(tmem_base + C::O_COL, advance_desc(dpj, ao), advance_desc(dv, bo), p.idesc_pv,
                      (fresh && k == 0) ? 0u : 1u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[stage]);                  // K_j and V_j are consumed once P_j V_j has retired
      }
      __syncwarp();
    }
  } else {
    // ================================================================ softmax + output (warps 2..5)
    // Thread = query row.  The output accumulator lives in TMEM for the whole walk (P_j V_j accumulates onto it in the
    // tensor core); the exponentials are taken against a LAGGING row max: the accumulator is only rescaled when the
    // true max has grown by more than 2^TAU since the last rescale (exact: numerator and denominator use the same
    // reference, which is what the log-sum-exp is reported against).  Per key tile a thread therefore reads its 128
    // scores ONCE (four tcgen05.ld in flight, one wait), releases the score buffer at once, and touches the
    // accumulator only on the rare rescale.
    constexpr float TAU = 8.f;
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;                   // query row inside the tile
    const int q_pos = q_tile * BLOCK_Q + r;              // position inside the sequence
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t o_addr = tmem_base + lane_addr + C::O_COL;
    float m_run = -INFINITY;                             // reference max of the exponentials (exp2 domain)
    float l_run = 0.f;
    if (resume) {
      // resume from the merged result of the key blocks processed by earlier launches (normalised output + LSE):
      // every row stores its previous output (or zeros) into the TMEM accumulator, P_0 V_0 accumulates onto it
      const size_t row_ = (size_t)q_row0 + r;
      float lse_prev = -INFINITY;
      if (q_pos < len_q) lse_prev = p.lse[row_ * p.hq + head];
      const bool live = lse_prev > -INFINITY;
      if (live) { m_run = lse_prev * 1.4426950408889634f; l_run = 1.f; }
      const float4* src = reinterpret_cast<const float4*>(p.o_state + row_ * ((size_t)p.hq * D) + (size_t)head * D);
#pragma unroll
      for (int c = 0; c < D; c += 16) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t4 = live ? src[(c >> 2) + i] : make_float4(0.f, 0.f, 0.f, 0.f);
          w[4 * i] = __float_as_uint(t4.x); w[4 * i + 1] = __float_as_uint(t4.y);
          w[4 * i + 2] = __float_as_uint(t4.z); w[4 * i + 3] = __float_as_uint(t4.w);
        }
        tmem_st_32x32b_x16(o_addr + c, w);
      }
      tmem_st_wait();
    }
    const float slope2 = (EXTRA && p.alibi != nullptr) ? p.alibi[head] * 1.4426950408889634f : 0.f;
    (void)slope2;
    for (int j = 0; j < kv_tiles; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_addr + sb * BLOCK_KV;
      const int kt = kv_first + j;                                       // key tile of this step
      const bool diag = p.causal && (kt == q_tile);
      const bool tail_k = ragged_k && (kt == kv_tiles_all - 1);
      const bool masked = diag || tail_k || (EXTRA && p.window > 0);
      const int k_lim = diag ? min(q_pos, len_k - 1) : (len_k - 1);      // last visible key position of this row
      const int k_low = (EXTRA && p.window > 0) ? q_pos - p.window + 1 : 0;   // first visible key position
      const int k_base = kt * BLOCK_KV;
      // ---- the whole score row into registers
      uint32_t v[BLOCK_KV];
#pragma unroll
      for (int c = 0; c < BLOCK_KV; c += 32) tmem_ld_32x32b_x32(s_addr + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);          // S buffer may be overwritten by Q K_{j+2}^T
      // plain variant: the softmax scale (> 0) is folded into the exponent's FMA, the max is taken on the raw scores;
      // EXTRA variant: scale and ALiBi bias are applied here and the exponent uses a unit multiplier
      // (four independent partial maxima / sums: a single running value would be a 128-long dependent chain)
      float mt[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < BLOCK_KV; ++i) {
        float sc = __uint_as_float(v[i]);
        if (EXTRA) sc = fmaf(slope2, (float)(k_base + i - q_pos), sc * p.scale_log2);
        if (masked && ((k_base + i) > k_lim || (EXTRA && (k_base + i) < k_low))) sc = -INFINITY;
        v[i] = __float_as_uint(sc);
        mt[i & 3] = fmaxf(mt[i & 3], sc);
      }
      float m_tile = fmaxf(fmaxf(mt[0], mt[1]), fmaxf(mt[2], mt[3]));
      const float ex_mul = EXTRA ? 1.f : p.scale_log2;
      m_tile *= ex_mul;
      // ---- lagging max / rare rescale of the TMEM accumulator
      const float m_new = fmaxf(m_run, m_tile);
      if (j == 0 && !resume) {
        m_run = m_new;                                   // nothing accumulated yet
      } else {
        const bool need = (m_new - m_run) > TAU;         // (-inf) - (-inf) = NaN -> false
        if (__any_sync(0xffffffffu, need)) {
          if (j > 0) {                                   // P_{j-1} V_{j-1} has retired: the accumulator is stable
            mbar_wait(o_full, (uint32_t)((j - 1) & 1));
            tc_fence_after();
          }
          const float alpha = need ? fast_exp2(m_run - m_new) : 1.f;
#pragma unroll 1
          for (int c = 0; c < D; c += 32) {
            uint32_t t[32];
            tmem_ld_32x32b_x32(o_addr + c, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st_32x32b_x16(o_addr + c, *reinterpret_cast<uint32_t(*)[16]>(&t[0]));
            tmem_st_32x32b_x16(o_addr + c + 16, *reinterpret_cast<uint32_t(*)[16]>(&t[16]));
          }
          tmem_st_wait();
          l_run *= alpha;
          if (need) m_run = m_new;
        }
      }
      const float m_use = (m_run == -INFINITY) ? 0.f : m_run;     // keeps the exponent finite on a fully masked row
      // ---- p = exp2(s - m), row sum, 16-bit P into the swizzled K-major tile of buffer j & 1
      // (that buffer was last read by P_{j-2} V_{j-2}, which retired before Q K_j^T - issued after it - completed)
      uint8_t* p_row = smem_p + (j & 1) * C::P_BYTES + r * 128;
      float lt[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < BLOCK_KV; c += 32) {
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(v[c + i]), ex_mul, -m_use));
          const float p1 = fast_exp2(fmaf(__uint_as_float(v[c + i + 1]), ex_mul, -m_use));
          lt[(i >> 1) & 3] += p0 + p1;
          if (p.out_dtype == CB_BF16) {
            __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
            packed[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
          } else {
            __half2 h = __floats2half2_rn(p0, p1);
            packed[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
          }
        }
        // 32 keys = 64 bytes = four 16-byte chunks of this row inside sub-tile (c / 64)
        uint8_t* sub = p_row + (c >> 6) * (BLOCK_Q * 128);
        const int chunk0 = (c & 32) ? 4 : 0;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = (chunk0 + q4) ^ (r & 7);     // SWIZZLE_128B: 16-byte chunk index XOR (row mod 8)
          *reinterpret_cast<uint4*>(sub + chunk * 16) =
              make_uint4(packed[q4 * 4], packed[q4 * 4 + 1], packed[q4 * 4 + 2], packed[q4 * 4 + 3]);
        }
      }
      const float l_tile = (lt[0] + lt[1]) + (lt[2] + lt[3]);
      l_run += l_tile;
      // keep in step with the accumulator barrier: observe "P_{j-1} V_{j-1} retired" once per tile.  It completed long
      // ago in the common case (no stall), but a parity wait is only unambiguous while the waiter is at most ONE phase
      // behind - the epilogue's wait for the last tile would otherwise pass on the phase before the previous one
      // (seen on hardware with two-tile launches: output read before the last two P V had landed).
      if (j > 0) mbar_wait(o_full, (uint32_t)((j - 1) & 1));
      tc_fence_before();
      fence_proxy_async();                               // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);                // P_j is in shared memory, the accumulator is consistent
    }
    // ---- epilogue: the last P V has retired; normalise and store this thread's row
    mbar_wait(o_full, (uint32_t)((kv_tiles - 1) & 1));
    tc_fence_after();
    const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
    const size_t row = (size_t)q_row0 + r;
    const bool row_ok = q_pos < len_q;                   // rows past the end of a (packed) sequence are never stored
#pragma unroll 1
    for (int c = 0; c < D; c += 32) {
      uint32_t t[32];
      __syncwarp();
      tmem_ld_32x32b_x32(o_addr + c, t);                 // executed by every lane (warp-collective), stores predicated
      tmem_ld_wait();
      if (!row_ok) {
        // a row past the end of its (packed) sequence: computed on garbage, never stored
      } else if (p.o_state != nullptr) {
        float4* dst = reinterpret_cast<float4*>(p.o_state + row * ((size_t)p.hq * D) + (size_t)head * D + c);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = make_float4(__uint_as_float(t[4 * i]) * inv_l, __uint_as_float(t[4 * i + 1]) * inv_l,
                               __uint_as_float(t[4 * i + 2]) * inv_l, __uint_as_float(t[4 * i + 3]) * inv_l);
      } else {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a0 = __uint_as_float(t[2 * i]) * inv_l, a1 = __uint_as_float(t[2 * i + 1]) * inv_l;
          if (p.out_dtype == CB_BF16) {
            __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
          } else {
            __half2 h = __floats2half2_rn(a0, a1);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
          }
        }
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) +
                                              (row * ((size_t)p.hq * D) + (size_t)head * D + c) * 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
      }
    }
    if (p.lse && q_pos < len_q)
      p.lse[row * p.hq + head] = l_run > 0.f ? (m_run + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int D>
int launch_flash_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int batch, int seqlen_q,
                     int seqlen_k, int hq, int hkv, int causal, float scale, int dtype, cudaStream_t stream,
                     const int* cu_seqlens = nullptr, long long total_tokens = 0, float* o_state = nullptr,
                     int has_prev = 0, int window = 0, const float* alibi = nullptr) {
  using C = FaCfg<D>;
  const bool bf16 = dtype == CB_BF16;
  CUtensorMap tq, tk, tv;
  const uint64_t rows_q = cu_seqlens ? (uint64_t)total_tokens : (uint64_t)batch * seqlen_q;
  const uint64_t rows_k = cu_seqlens ? (uint64_t)total_tokens : (uint64_t)batch * seqlen_k;
  int r = make_tmap_2d_16b(&tq, q, rows_q, (uint64_t)hq * D, (uint64_t)hq * D, BLOCK_Q, 64, bf16);
  if (r) return 1000 + r;
  r = make_tmap_2d_16b(&tk, k, rows_k, (uint64_t)hkv * D, (uint64_t)hkv * D, BLOCK_KV, 64, bf16);
  if (r) return 2000 + r;
  r = make_tmap_2d_16b(&tv, v, rows_k, (uint64_t)hkv * D, (uint64_t)hkv * D, BLOCK_KV, 64, bf16);
  if (r) return 3000 + r;
  FaParams p;
  p.cu_seqlens = cu_seqlens;
  p.o_state = o_state; p.has_prev = has_prev;
  p.batch = batch; p.seqlen_q = seqlen_q; p.seqlen_k = seqlen_k; p.hq = hq; p.hkv = hkv; p.causal = causal;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = out; p.lse = lse; p.out_dtype = dtype;
  p.window = window; p.alibi = alibi;
  p.idesc_qk = make_idesc_f16(BLOCK_Q, BLOCK_KV, bf16 ? 1 : 0, 0, 0);      // S[128 x 128] = Q (K-major) x K (K-major)
  p.idesc_pv = make_idesc_f16(BLOCK_Q, D, bf16 ? 1 : 0, 0, 1);              // O[128 x D]  = P (K-major) x V (MN-major)
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(flash_fwd_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(flash_fwd_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  // packed: at most total/128 + batch query tiles exist; uniform: ceil(seqlen / 128) per sequence
  dim3 grid(cu_seqlens ? (unsigned)(total_tokens / BLOCK_Q + batch) : (unsigned)((seqlen_q + BLOCK_Q - 1) / BLOCK_Q), hq,
            cu_seqlens ? 1 : batch);
  if (window > 0 || alibi != nullptr)
    flash_fwd_kernel<D, true><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(tq, tk, tv, p);
  else
    flash_fwd_kernel<D, false><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(tq, tk, tv, p);
  return (int)cudaGetLastError();
}


// =====================================================================================================================
// Flash attention BACKWARD on tcgen05 / TMEM.
//
//   P = exp(scale * Q K^T - LSE),  dP = dO V^T,  dS = P o (dP - delta),  delta = rowsum(dO o O)
//   dV = P^T dO,   dK = scale * dS^T Q,   dQ = scale * dS K
//
// One CTA owns ONE key tile (128 keys) of ONE kv head and walks over every query tile (64 rows) of every query head of
// its GQA group that can see it.  dK / dV therefore accumulate in TMEM across the whole walk - over the query tiles AND
// over the query heads of the group - and are written exactly once (no atomics, no KV expansion); dQ leaves the CTA as
// fp32 `red.global.add` into an accumulator that a cast kernel turns into bf16/fp16 afterwards.
//
// Everything is computed TRANSPOSED so that every MMA has M = 128 (full TMEM lanes) and the softmax threads own KEY
// rows:   S^T = K Q^T and dP^T = V dO^T are [128 keys x 64 q] accumulators; thread k of the softmax warps reads key row k,
// forms P^T and dS^T, and stores them as bf16 [keys x q] K-major swizzled tiles.  Those two tiles feed
//   dV  += P^T  (A, K-major)  x dO (B, MN-major: the SAME smem bytes that were the K-major B of dP^T = V dO^T)
//   dK  += dS^T (A, K-major)  x Q  (B, MN-major: the same bytes as the B of S^T = K Q^T)
//   dQ^T = K^T  (A, MN-major view of the resident K tile) x dS (B, MN-major view of the dS^T tile)      [128 d x 64 q]
// i.e. five tensor-core GEMMs per (key tile, query tile) pair out of four smem operands, no transposition pass.
//
// Warps (10): 0 TMA producer (K, V once; Q_i / dO_i through a 2-stage ring), 1 MMA issuer, 2..5 softmax (thread = key
// row), 6..9 dQ drain (thread = head-dim lane; 128-byte coalesced fp32 reductions per query row).
// The issuer runs S^T / dP^T of tile i+1 BEFORE the three accumulation GEMMs of tile i, so the tensor pipe works while
// the softmax warps are busy (same trick as the forward kernel).
// TMEM (512 columns): [0,64) S^T, [64,128) dP^T, [128,192) dQ^T, [256,384) dV, [384,512) dK.
// Shared memory: K 32 KB + V 32 KB + 2 x (Q 16 KB + dO 16 KB) + P^T 16 KB + dS^T 16 KB + lse/delta = 161 KB (D = 128).
constexpr int BWD_BLOCK_KV = 128;
constexpr int BWD_BLOCK_Q = 64;
constexpr int BWD_THREADS = 320;
constexpr int BWD_QSTAGES = 2;

template <int D> struct FaBwdCfg {
  static constexpr int K_BYTES = BWD_BLOCK_KV * D * 2;
  static constexpr int V_BYTES = BWD_BLOCK_KV * D * 2;
  static constexpr int Q_BYTES = BWD_BLOCK_Q * D * 2;
  static constexpr int DO_BYTES = BWD_BLOCK_Q * D * 2;
  static constexpr int QDO_STAGE_BYTES = Q_BYTES + DO_BYTES;
  static constexpr int PT_BYTES = BWD_BLOCK_KV * BWD_BLOCK_Q * 2;
  static constexpr int SMEM_BYTES = K_BYTES + V_BYTES + BWD_QSTAGES * QDO_STAGE_BYTES + 2 * PT_BYTES +
                                    2 * BWD_BLOCK_Q * 4 /*lse, delta*/ + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 512;
  static constexpr int COL_ST = 0, COL_DPT = 64, COL_DQT = 128, COL_DV = 256, COL_DK = 384;
};

struct FaBwdParams {
  const int* cu_seqlens;             // packed batch boundaries (see FaParams) or null
  int batch, seqlen;                 // equal-length sequences, seqlen_q == seqlen_k (self attention)
  int hq, hkv;
  int causal;
  float scale, scale_log2;
  const float* lse;                  // [batch * seqlen, hq] natural log
  const float* delta;                // [batch * seqlen, hq] rowsum(dO o O)
  float* dq_acc;                     // [batch * seqlen, hq * D] fp32, zero-initialised
  void* dk;                          // [batch * seqlen, hkv * D]
  void* dv;
  // ring attention: dK / dV of this key block are ADDED (fp32 vector reductions) into accumulators that may live in
  // the block owner's memory (peer-mapped symmetric buffer: the reduction crosses NVLink) instead of being stored
  float* dk_acc;                     // [rows, hkv * D] fp32 or null
  float* dv_acc;
  int out_dtype;
  uint32_t idesc_st, idesc_dv, idesc_dqt;
};

template <int D>
__global__ void __launch_bounds__(BWD_THREADS, 1)
flash_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                 const FaBwdParams p) {
  using C = FaBwdCfg<D>;
  static_assert(D == 128, "head_dim 128");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_k = smem;
  uint8_t* smem_v = smem_k + C::K_BYTES;
  uint8_t* smem_qdo = smem_v + C::V_BYTES;
  uint8_t* smem_pt = smem_qdo + BWD_QSTAGES * C::QDO_STAGE_BYTES;
  uint8_t* smem_dst = smem_pt + C::PT_BYTES;
  float* smem_lse = reinterpret_cast<float*>(smem_dst + C::PT_BYTES);     // [64] lse * log2(e)
  float* smem_delta = smem_lse + BWD_BLOCK_Q;                             // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_delta + BWD_BLOCK_Q);
  uint64_t* kv_full = bars;                // [1]
  uint64_t* qdo_full = bars + 1;           // [2]
  uint64_t* qdo_empty = bars + 3;          // [2]
  uint64_t* sp_full = bars + 5;            // [1]  S^T and dP^T are in TMEM
  uint64_t* sp_empty = bars + 6;           // [1]  4 softmax warps have read them
  uint64_t* pds_full = bars + 7;           // [1]  4 softmax warps wrote P^T / dS^T to smem
  uint64_t* pds_empty = bars + 8;          // [1]  the three accumulation GEMMs that read them retired
  uint64_t* dq_full = bars + 9;            // [1]  dQ^T is in TMEM
  uint64_t* dq_empty = bars + 10;          // [1]  4 drain warps have read it
  uint64_t* acc_full = bars + 11;          // [1]  final dK / dV complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kv_head = blockIdx.y;
  const int group = p.hq / p.hkv;
  int kv_tile, seq_row0, len;
  if (p.cu_seqlens != nullptr) {
    int t = (int)blockIdx.x, bsel = -1, start = 0;
    len = 0;
    for (int bb = 0; bb < p.batch; ++bb) {
      start = p.cu_seqlens[bb];
      len = p.cu_seqlens[bb + 1] - start;
      const int nt = (len + BWD_BLOCK_KV - 1) / BWD_BLOCK_KV;
      if (t < nt) { bsel = bb; break; }
      t -= nt;
    }
    if (bsel < 0) return;
    kv_tile = t; seq_row0 = start;
  } else {
    kv_tile = blockIdx.x; seq_row0 = (int)blockIdx.z * p.seqlen; len = p.seqlen;
  }
  const int q_tiles = (len + BWD_BLOCK_Q - 1) / BWD_BLOCK_Q;
  // causal: query tile i (rows 64 i ..) sees key tile j (keys 128 j ..) iff 64 i + 63 >= 128 j  <=>  i >= 2 j
  const int i_start = p.causal ? 2 * kv_tile : 0;
  const int n_i = max(q_tiles - i_start, 0);
  const int n_iter = n_i * group;                                  // (query head of the group, query tile) pairs
  const int kv_row0 = seq_row0 + kv_tile * BWD_BLOCK_KV;
  const bool ragged = (len % BWD_BLOCK_KV) != 0;                    // partial last key tile and / or query tile

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_k);
    prefetch_tensormap(&tmap_v);
    prefetch_tensormap(&tmap_do);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(sp_full, 1);
    mbar_init(sp_empty, 4);
    mbar_init(pds_full, 4);
    mbar_init(pds_empty, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, C::K_BYTES + C::V_BYTES);
#pragma unroll
      for (int h = 0; h < D / 64; ++h) {
        tma_load_2d(&tmap_k, kv_full, smem_k + h * (BWD_BLOCK_KV * 128), kv_head * D + h * 64, kv_row0);
        tma_load_2d(&tmap_v, kv_full, smem_v + h * (BWD_BLOCK_KV * 128), kv_head * D + h * 64, kv_row0);
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int hq = kv_head * group + it / n_i;
      const int q_row0 = seq_row0 + (i_start + it % n_i) * BWD_BLOCK_Q;
      mbar_wait(&qdo_empty[stage], phase ^ 1);
      if (lane == 0) {
        uint8_t* sq = smem_qdo + stage * C::QDO_STAGE_BYTES;
        uint8_t* sdo = sq + C::Q_BYTES;
        mbar_arrive_expect_tx(&qdo_full[stage], C::QDO_STAGE_BYTES);
#pragma unroll
        for (int h = 0; h < D / 64; ++h) {               // boxes [64 rows, 64 d]
          tma_load_2d(&tmap_q, &qdo_full[stage], sq + h * (BWD_BLOCK_Q * 128), hq * D + h * 64, q_row0);
          tma_load_2d(&tmap_do, &qdo_full[stage], sdo + h * (BWD_BLOCK_Q * 128), hq * D + h * 64, q_row0);
        }
      }
      __syncwarp();
      if (++stage == BWD_QSTAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t sk = smem_u32(smem_k), sv = smem_u32(smem_v);
    const uint64_t dk_kmaj = make_smem_desc_sw128(sk, 16, 1024);                        // A of S^T: [keys x d] K-major
    const uint64_t dv_kmaj = make_smem_desc_sw128(sv, 16, 1024);                        // A of dP^T
    const uint64_t dk_mnmaj = make_smem_desc_sw128(sk, BWD_BLOCK_KV * 128, 1024);       // A of dQ^T: K^T, M = d contiguous
    const uint64_t d_pt = make_smem_desc_sw128(smem_u32(smem_pt), 16, 1024);            // A of dV: [keys x q] K-major
    const uint64_t d_dst = make_smem_desc_sw128(smem_u32(smem_dst), 16, 1024);          // A of dK
    const uint64_t d_ds_mn = make_smem_desc_sw128(smem_u32(smem_dst), BWD_BLOCK_KV * 128, 1024);   // B of dQ^T (N = q)
    auto issue_sp = [&](int it) {
      const int stage = it % BWD_QSTAGES;
      mbar_wait(&qdo_full[stage], (uint32_t)((it / BWD_QSTAGES) & 1));
      mbar_wait(sp_empty, (uint32_t)((it & 1) ^ 1));          // softmax of iteration it-1 has read S^T / dP^T
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sq = smem_u32(smem_qdo + stage * C::QDO_STAGE_BYTES);
        const uint64_t dq_b = make_smem_desc_sw128(sq, 16, 1024);                       // B: [q x d] K-major
        const uint64_t ddo_b = make_smem_desc_sw128(sq + C::Q_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < D / UMMA_K; ++k) {
          const uint32_t ao = (k >> 2) * (BWD_BLOCK_KV * 128) + (k & 3) * (UMMA_K * 2);
          const uint32_t bo = (k >> 2) * (BWD_BLOCK_Q * 128) + (k & 3) * (UMMA_K * 2);
          umma_f16_ss(tmem_base + C::COL_ST, advance_desc(dk_kmaj, ao), advance_desc(dq_b, bo), p.idesc_st,
                      k > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < D / UMMA_K; ++k) {
          const uint32_t ao = (k >> 2) * (BWD_BLOCK_KV * 128) + (k & 3) * (UMMA_K * 2);
          const uint32_t bo = (k >> 2) * (BWD_BLOCK_Q * 128) + (k & 3) * (UMMA_K * 2);
          umma_f16_ss(tmem_base + C::COL_DPT, advance_desc(dv_kmaj, ao), advance_desc(ddo_b, bo), p.idesc_st,
                      k > 0 ? 1u : 0u);
        }
        umma_commit(sp_full);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    tc_fence_after();
    if (n_iter > 0) issue_sp(0);
    for (int it = 0; it < n_iter; ++it) {
      if (it + 1 < n_iter) issue_sp(it + 1);                 // tensor pipe stays busy during softmax(it)
      const int stage = it % BWD_QSTAGES;
      mbar_wait(pds_full, (uint32_t)(it & 1));
      mbar_wait(dq_empty, (uint32_t)((it & 1) ^ 1));         // dQ^T of iteration it-1 has been drained
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sq = smem_u32(smem_qdo + stage * C::QDO_STAGE_BYTES);
        // B operands, MN-major views of the Q / dO tiles: rows = q (the reduction dim), N = d contiguous,
        // 64-element N chunks are BWD_BLOCK_Q * 128 bytes apart
        const uint64_t dq_mn = make_smem_desc_sw128(sq, BWD_BLOCK_Q * 128, 1024);
        const uint64_t ddo_mn = make_smem_desc_sw128(sq + C::Q_BYTES, BWD_BLOCK_Q * 128, 1024);
        const uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < BWD_BLOCK_Q / UMMA_K; ++k)      // dV += P^T dO      (reduction over the 64 query rows)
          umma_f16_ss(tmem_base + C::COL_DV, advance_desc(d_pt, k * (UMMA_K * 2)),
                      advance_desc(ddo_mn, k * (UMMA_K * 128)), p.idesc_dv, (acc | (k > 0)) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < BWD_BLOCK_Q / UMMA_K; ++k)      // dK += dS^T Q
          umma_f16_ss(tmem_base + C::COL_DK, advance_desc(d_dst, k * (UMMA_K * 2)),
                      advance_desc(dq_mn, k * (UMMA_K * 128)), p.idesc_dv, (acc | (k > 0)) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < BWD_BLOCK_KV / UMMA_K; ++k)     // dQ^T = K^T dS     (reduction over the 128 keys)
          umma_f16_ss(tmem_base + C::COL_DQT, advance_desc(dk_mnmaj, k * (UMMA_K * 128)),
                      advance_desc(d_ds_mn, k * (UMMA_K * 128)), p.idesc_dqt, k > 0 ? 1u : 0u);
        umma_commit(dq_full);
        umma_commit(pds_empty);                              // P^T / dS^T smem may be overwritten
        umma_commit(&qdo_empty[stage]);                      // Q_i / dO_i consumed
        if (it == n_iter - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ================================================================ softmax warps: thread = key row
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // key row inside the tile
    const int key_pos = kv_tile * BWD_BLOCK_KV + r;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int sm_tid = threadIdx.x - 64;                     // 0..127 among the softmax warps
    uint8_t* pt_row = smem_pt + r * 128;
    uint8_t* dst_row = smem_dst + r * 128;
    for (int it = 0; it < n_iter; ++it) {
      const int hq = kv_head * group + it / n_i;
      const int q_tile = i_start + it % n_i;
      const int q_row0 = seq_row0 + q_tile * BWD_BLOCK_Q;
      // per-query-row statistics of this tile (64 lse + 64 delta values, read by every softmax thread); rows past the
      // end of the sequence read as 0 (their P / dS are forced to 0 below)
      asm volatile("bar.sync 1, 128;" ::: "memory");         // previous iteration's readers are done
      {
        const int qr = sm_tid < BWD_BLOCK_Q ? sm_tid : sm_tid - BWD_BLOCK_Q;
        const bool live = q_tile * BWD_BLOCK_Q + qr < len;
        if (sm_tid < BWD_BLOCK_Q)
          smem_lse[qr] = live ? p.lse[(size_t)(q_row0 + qr) * p.hq + hq] * 1.4426950408889634f : 0.f;
        else
          smem_delta[qr] = live ? p.delta[(size_t)(q_row0 + qr) * p.hq + hq] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(sp_full, (uint32_t)(it & 1));
      tc_fence_after();
      const bool diag = (p.causal && (q_tile <= 2 * kv_tile + 1)) ||
                        (ragged && (kv_tile == (len - 1) / BWD_BLOCK_KV || q_tile == q_tiles - 1));
      const int q_base = q_tile * BWD_BLOCK_Q;
      const int q_lo = p.causal ? key_pos : 0;               // this key is visible to query positions [q_lo, len)
      const bool key_live = key_pos < len;
      uint32_t p_packed[BWD_BLOCK_Q / 2], ds_packed[BWD_BLOCK_Q / 2];
#pragma unroll
      for (int c = 0; c < BWD_BLOCK_Q; c += 32) {
        uint32_t sv_[32], dpv[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + C::COL_ST + c, sv_);
        tmem_ld_32x32b_x32(tmem_base + lane_addr + C::COL_DPT + c, dpv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float pv[2], dsv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = c + i + e;
            float pe = fast_exp2(__uint_as_float(sv_[i + e]) * p.scale_log2 - smem_lse[col]);
            float de = pe * (__uint_as_float(dpv[i + e]) - smem_delta[col]);
            if (diag && !(key_live && q_base + col >= q_lo && q_base + col < len)) { pe = 0.f; de = 0.f; }
            pv[e] = pe;
            dsv[e] = de;
          }
          if (p.out_dtype == CB_BF16) {
            __nv_bfloat162 hp = __floats2bfloat162_rn(pv[0], pv[1]);
            __nv_bfloat162 hd = __floats2bfloat162_rn(dsv[0], dsv[1]);
            p_packed[(c + i) >> 1] = *reinterpret_cast<uint32_t*>(&hp);
            ds_packed[(c + i) >> 1] = *reinterpret_cast<uint32_t*>(&hd);
          } else {
            __half2 hp = __floats2half2_rn(pv[0], pv[1]);
            __half2 hd = __floats2half2_rn(dsv[0], dsv[1]);
            p_packed[(c + i) >> 1] = *reinterpret_cast<uint32_t*>(&hp);
            ds_packed[(c + i) >> 1] = *reinterpret_cast<uint32_t*>(&hd);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sp_empty);                  // S^T / dP^T may be overwritten by iteration it+1
      // the accumulation GEMMs of iteration it-1 must have retired before their smem operands are replaced
      mbar_wait(pds_empty, (uint32_t)((it & 1) ^ 1));
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {                       // 64 q = 128 bytes = 8 chunks of 16 B, SWIZZLE_128B
        const int pos = (ch ^ (r & 7)) * 16;
        *reinterpret_cast<uint4*>(pt_row + pos) =
            make_uint4(p_packed[ch * 4], p_packed[ch * 4 + 1], p_packed[ch * 4 + 2], p_packed[ch * 4 + 3]);
        *reinterpret_cast<uint4*>(dst_row + pos) =
            make_uint4(ds_packed[ch * 4], ds_packed[ch * 4 + 1], ds_packed[ch * 4 + 2], ds_packed[ch * 4 + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    // ---- epilogue: dV and scale * dK rows of this thread's key
    if (n_iter > 0) mbar_wait(acc_full, 0);
    tc_fence_after();
    const size_t row = (size_t)kv_row0 + r;
    const size_t ld = (size_t)p.hkv * D;
    // NOTE: tcgen05.ld is warp-collective (.sync.aligned): every lane runs the loads, only the STORES are predicated on
    // the key being inside the sequence (a ragged last tile leaves some lanes of a warp without a row)
    const bool key_ok = key_pos < len;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t col0 = which == 0 ? C::COL_DV : C::COL_DK;
      const float mul = which == 0 ? 1.f : p.scale;
      if (p.dk_acc != nullptr) {
        float* acc_row = (which == 0 ? p.dv_acc : p.dk_acc) + (row * ld + (size_t)kv_head * D);
#pragma unroll 1
        for (int c = 0; c < D; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + lane_addr + col0 + c, v);
          tmem_ld_wait();
          if (n_iter > 0 && key_ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                           :: "l"(acc_row + c + i), "f"(__uint_as_float(v[i]) * mul), "f"(__uint_as_float(v[i + 1]) * mul),
                              "f"(__uint_as_float(v[i + 2]) * mul), "f"(__uint_as_float(v[i + 3]) * mul) : "memory");
          }
        }
        continue;
      }
      uint8_t* base = reinterpret_cast<uint8_t*>(which == 0 ? p.dv : p.dk) + (row * ld + (size_t)kv_head * D) * 2;
#pragma unroll 1
      for (int c = 0; c < D; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + col0 + c, v);
        tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a0 = n_iter > 0 ? __uint_as_float(v[2 * i]) * mul : 0.f;
          const float a1 = n_iter > 0 ? __uint_as_float(v[2 * i + 1]) * mul : 0.f;
          if (p.out_dtype == CB_BF16) {
            __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
          } else {
            __half2 h = __floats2half2_rn(a0, a1);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
          }
        }
        if (key_ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(base + (c + i * 8) * 2) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        }
      }
    }
  } else {
    // ================================================================ dQ drain warps: thread = head-dim lane
    const int quarter = warp & 3;
    const int dlane = quarter * 32 + lane;                   // d index (row of dQ^T)
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    for (int it = 0; it < n_iter; ++it) {
      const int hq = kv_head * group + it / n_i;
      const int q_tile = i_start + it % n_i;
      const int q_row0 = seq_row0 + q_tile * BWD_BLOCK_Q;
      const int q_live = len - q_tile * BWD_BLOCK_Q;         // rows of this tile that belong to the sequence
      mbar_wait(dq_full, (uint32_t)(it & 1));
      tc_fence_after();
      float* dst = p.dq_acc + ((size_t)q_row0 * p.hq + hq) * D + dlane;
      const size_t row_stride = (size_t)p.hq * D;
#pragma unroll
      for (int c = 0; c < BWD_BLOCK_Q; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + C::COL_DQT + c, v);
        tmem_ld_wait();
        if (dlane < D) {
#pragma unroll
          for (int i = 0; i < 32; ++i)                      // one 128-byte coalesced fp32 reduction per warp and row
            if (c + i < q_live) atomicAdd(dst + (size_t)(c + i) * row_stride, __uint_as_float(v[i]) * p.scale);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// delta[t, h] = sum_d dO[t, h, d] * O[t, h, d]   (one warp per (token, head) row)
template <typename T>
__global__ void flash_bwd_delta_kernel(const T* __restrict__ out, const T* __restrict__ dout, float* __restrict__ delta,
                                       long long rows, int D) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* o = out + row * D;
  const T* g = dout + row * D;
  float acc = 0.f;
  for (int i = lane * 8; i < D; i += 32 * 8) {
    Vec16<T> a, b2;
    a.load(o + i);
    b2.load(g + i);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += a.get(e) * b2.get(e);
  }
  acc = warp_sum(acc);
  if (lane == 0) delta[row] = acc;
}

template <typename T>
__global__ void flash_bwd_cast_dq_kernel(const float* __restrict__ acc, T* __restrict__ dq, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  const float4 a = *reinterpret_cast<const float4*>(acc + i);
  const float4 b2 = *reinterpret_cast<const float4*>(acc + i + 4);
  Vec16<T> o;
  o.set(0, a.x); o.set(1, a.y); o.set(2, a.z); o.set(3, a.w);
  o.set(4, b2.x); o.set(5, b2.y); o.set(6, b2.z); o.set(7, b2.w);
  o.store(dq + i);
}

template <int D>
int launch_flash_bwd(const void* q, const void* k, const void* v, const void* dout, const float* lse, const float* delta,
                     float* dq_acc, void* dk, void* dv, int batch, int seqlen, int hq, int hkv, int causal, float scale,
                     int dtype, cudaStream_t stream, const int* cu_seqlens = nullptr, long long total_tokens = 0,
                     float* dk_acc = nullptr, float* dv_acc = nullptr) {
  using C = FaBwdCfg<D>;
  const bool bf16 = dtype == CB_BF16;
  CUtensorMap tq, tk, tv, tdo;
  const uint64_t T = cu_seqlens ? (uint64_t)total_tokens : (uint64_t)batch * seqlen;
  int r = make_tmap_2d_16b(&tq, q, T, (uint64_t)hq * D, (uint64_t)hq * D, BWD_BLOCK_Q, 64, bf16);
  if (r) return 1000 + r;
  r = make_tmap_2d_16b(&tdo, dout, T, (uint64_t)hq * D, (uint64_t)hq * D, BWD_BLOCK_Q, 64, bf16);
  if (r) return 4000 + r;
  r = make_tmap_2d_16b(&tk, k, T, (uint64_t)hkv * D, (uint64_t)hkv * D, BWD_BLOCK_KV, 64, bf16);
  if (r) return 2000 + r;
  r = make_tmap_2d_16b(&tv, v, T, (uint64_t)hkv * D, (uint64_t)hkv * D, BWD_BLOCK_KV, 64, bf16);
  if (r) return 3000 + r;
  FaBwdParams p;
  p.cu_seqlens = cu_seqlens;
  p.batch = batch; p.seqlen = seqlen; p.hq = hq; p.hkv = hkv; p.causal = causal;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse; p.delta = delta; p.dq_acc = dq_acc; p.dk = dk; p.dv = dv; p.out_dtype = dtype;
  p.dk_acc = dk_acc; p.dv_acc = dv_acc;
  const int f = bf16 ? 1 : 0;
  p.idesc_st = make_idesc_f16(BWD_BLOCK_KV, BWD_BLOCK_Q, f, 0, 0);     // [128 keys x 64 q]  = K-major x K-major
  p.idesc_dv = make_idesc_f16(BWD_BLOCK_KV, D, f, 0, 1);                // [128 keys x D]     = K-major x MN-major
  p.idesc_dqt = make_idesc_f16(D, BWD_BLOCK_Q, f, 1, 1);                // [D x 64 q]         = MN-major x MN-major
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(flash_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(cu_seqlens ? (unsigned)(total_tokens / BWD_BLOCK_KV + batch) : (unsigned)((seqlen + BWD_BLOCK_KV - 1) / BWD_BLOCK_KV),
            hkv, cu_seqlens ? 1 : batch);
  flash_bwd_kernel<D><<<grid, BWD_THREADS, C::SMEM_BYTES, stream>>>(tq, tk, tv, tdo, p);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

// q [batch * seqlen_q, hq, D], k / v [batch * seqlen_k, hkv, D] contiguous token-major; out like q; lse [batch *
// seqlen_q, hq] fp32 (may be null).  Requirements: D in {64, 128}, seqlen_q and seqlen_k multiples of 128,
// causal => seqlen_q == seqlen_k, hq % hkv == 0, 16-byte aligned bases.
int cb_flash_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int batch, int seqlen_q,
                      int seqlen_k, int hq, int hkv, int head_dim, int causal, float scale, int dtype,
                      cudaStream_t stream) {
  if (batch <= 0 || seqlen_q <= 0) return 0;
  if (hq % hkv || (causal && seqlen_q != seqlen_k)) return (int)cudaErrorInvalidValue;
  // sequence lengths need not be multiples of the tile: partial tiles are masked (keys) / not stored (queries)
  if (dtype != CB_BF16 && dtype != CB_F16) return (int)cudaErrorInvalidValue;
  if (head_dim == 128)
    return launch_flash_fwd<128>(q, k, v, out, lse, batch, seqlen_q, seqlen_k, hq, hkv, causal, scale, dtype, stream);
  if (head_dim == 64)
    return launch_flash_fwd<64>(q, k, v, out, lse, batch, seqlen_q, seqlen_k, hq, hkv, causal, scale, dtype, stream);
  return (int)cudaErrorInvalidValue;
}

// Packed (variable-length) self attention: q / k / v [total_tokens, heads, D]; cu_seqlens int32[batch + 1] ON THE DEVICE
// (never read by the host: the grid is sized for the worst case and surplus CTAs exit).  Reference counterpart:
// `flash_attn_varlen_kvpacked_func` behind `extensions/pybind/flash_attention/flash_attention_dao_cuda.py:37-96`.
int cb_flash_attn_varlen_fwd(const void* q, const void* k, const void* v, void* out, float* lse, const int* cu_seqlens,
                             int batch, long long total_tokens, int hq, int hkv, int head_dim, int causal, float scale,
                             int dtype, cudaStream_t stream) {
  if (batch <= 0 || total_tokens <= 0) return 0;
  if (hq % hkv || (dtype != CB_BF16 && dtype != CB_F16) || cu_seqlens == nullptr) return (int)cudaErrorInvalidValue;
  if (head_dim == 128)
    return launch_flash_fwd<128>(q, k, v, out, lse, batch, 0, 0, hq, hkv, causal, scale, dtype, stream, cu_seqlens, total_tokens);
  if (head_dim == 64)
    return launch_flash_fwd<64>(q, k, v, out, lse, batch, 0, 0, hq, hkv, causal, scale, dtype, stream, cu_seqlens, total_tokens);
  return (int)cudaErrorInvalidValue;
}

// Packed causal forward for inference prefill with a sliding window (Mistral: query q sees keys (q - window, q]) and /
// or ALiBi slopes ([hq] fp32: score += slope * (key - query); BLOOM, Baichuan-13B).  Forward only.
int cb_flash_attn_varlen_fwd_ex(const void* q, const void* k, const void* v, void* out, float* lse,
                                const int* cu_seqlens, int batch, long long total_tokens, int hq, int hkv, int head_dim,
                                float scale, int window, const float* alibi_slopes, int dtype, cudaStream_t stream) {
  if (batch <= 0 || total_tokens <= 0) return 0;
  if (hq % hkv || (dtype != CB_BF16 && dtype != CB_F16) || cu_seqlens == nullptr || window < 0)
    return (int)cudaErrorInvalidValue;
  if (head_dim == 128)
    return launch_flash_fwd<128>(q, k, v, out, lse, batch, 0, 0, hq, hkv, 1, scale, dtype, stream, cu_seqlens,
                                 total_tokens, nullptr, 0, window, alibi_slopes);
  if (head_dim == 64)
    return launch_flash_fwd<64>(q, k, v, out, lse, batch, 0, 0, hq, hkv, 1, scale, dtype, stream, cu_seqlens,
                                total_tokens, nullptr, 0, window, alibi_slopes);
  return (int)cudaErrorInvalidValue;
}

// Backward of cb_flash_attn_fwd for self attention (seqlen_q == seqlen_k).  q / out / dout / dq [T, hq, D], k / v / dk /
// dv [T, hkv, D], lse [T, hq] fp32 from the forward.  Workspaces: delta [T, hq] fp32 and dq_acc [T, hq, D] fp32 (this
// function zeroes dq_acc).  Requirements as the forward; head_dim 128 (64 also instantiated).
int cb_flash_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                      void* dq, void* dk, void* dv, float* delta, float* dq_acc, int batch, int seqlen, int hq, int hkv,
                      int head_dim, int causal, float scale, int dtype, const int* cu_seqlens, long long total_tokens,
                      cudaStream_t stream) {
  // cu_seqlens != null: packed batch of `total_tokens` rows (seqlen ignored); else `batch` sequences of `seqlen`
  if (batch <= 0 || (cu_seqlens ? total_tokens <= 0 : seqlen <= 0)) return 0;
  // head_dim 64 would make dQ^T an M = 64 MMA (different TMEM lane layout): not wired up yet
  if (hq % hkv || head_dim != 128) return (int)cudaErrorInvalidValue;
  if (dtype != CB_BF16 && dtype != CB_F16) return (int)cudaErrorInvalidValue;
  const long long rows = (cu_seqlens ? total_tokens : (long long)batch * seqlen) * hq;
  const long long n = rows * head_dim;
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, (size_t)n * sizeof(float), stream);
  if (e != cudaSuccess) return (int)e;
  const int wpb = 8;
  if (dtype == CB_BF16)
    flash_bwd_delta_kernel<__nv_bfloat16><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
        (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, delta, rows, head_dim);
  else
    flash_bwd_delta_kernel<__half><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
        (const __half*)out, (const __half*)dout, delta, rows, head_dim);
  int rc = launch_flash_bwd<128>(q, k, v, dout, lse, delta, dq_acc, dk, dv, batch, seqlen, hq, hkv, causal, scale, dtype,
                                 stream, cu_seqlens, total_tokens);
  if (rc) return rc;
  const unsigned blocks = (unsigned)((n / 8 + 255) / 256);
  if (dtype == CB_BF16) flash_bwd_cast_dq_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(dq_acc, (__nv_bfloat16*)dq, n);
  else flash_bwd_cast_dq_kernel<__half><<<blocks, 256, 0, stream>>>(dq_acc, (__half*)dq, n);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// Blockwise entry points for fused ring (context-parallel) attention.  One launch = the queries of this rank against ONE
// key/value block that may live in a PEER's memory (k / v are then peer-mapped symmetric-buffer addresses: the TMA
// loads of the main loop pull the tiles over NVLink, GQA heads only, no staging copy).
//   forward : the online-softmax state (o_state fp32, lse) is carried from launch to launch inside the kernel
//             (`has_prev`), which replaces the reference's per-hop `_rescale_out_lse` pass (`layer/attn.py:376-403`);
//   backward: dQ accumulates in the local fp32 `dq_acc` (not zeroed here), dK / dV are reduced straight into the block
//             owner's fp32 accumulators (`dk_acc` / `dv_acc`, peer pointers) - the reference circulates fp32 dKV
//             buffers around the ring instead (`layer/attn.py:1066-1163`).  `delta` = rowsum(dO o O) is an input.
int cb_flash_attn_block_fwd(const void* q, const void* k, const void* v, float* o_state, float* lse, int rows, int hq,
                            int hkv, int head_dim, int causal, int has_prev, float scale, int dtype, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (hq % hkv || (dtype != CB_BF16 && dtype != CB_F16) || o_state == nullptr || lse == nullptr)
    return (int)cudaErrorInvalidValue;
  if (head_dim == 128)
    return launch_flash_fwd<128>(q, k, v, nullptr, lse, 1, rows, rows, hq, hkv, causal, scale, dtype, stream, nullptr, 0,
                                 o_state, has_prev);
  if (head_dim == 64)
    return launch_flash_fwd<64>(q, k, v, nullptr, lse, 1, rows, rows, hq, hkv, causal, scale, dtype, stream, nullptr, 0,
                                o_state, has_prev);
  return (int)cudaErrorInvalidValue;
}

int cb_flash_attn_block_bwd(const void* q, const void* k, const void* v, const void* dout, const float* lse,
                            const float* delta, float* dq_acc, float* dk_acc, float* dv_acc, int rows, int hq, int hkv,
                            int head_dim, int causal, float scale, int dtype, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (hq % hkv || head_dim != 128 || (dtype != CB_BF16 && dtype != CB_F16) || !dq_acc || !dk_acc || !dv_acc)
    return (int)cudaErrorInvalidValue;
  return launch_flash_bwd<128>(q, k, v, dout, lse, delta, dq_acc, nullptr, nullptr, 1, rows, hq, hkv, causal, scale, dtype,
                               stream, nullptr, 0, dk_acc, dv_acc);
}

// delta[t, h] = sum_d dO o O for `rows` (token, head) rows of width head_dim
int cb_flash_attn_delta(const void* out, const void* dout, float* delta, long long rows, int head_dim, int dtype,
                        cudaStream_t stream) {
  if (rows <= 0) return 0;
  const int wpb = 8;
  if (dtype == CB_BF16)
    flash_bwd_delta_kernel<__nv_bfloat16><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
        (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, delta, rows, head_dim);
  else if (dtype == CB_F16)
    flash_bwd_delta_kernel<__half><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
        (const __half*)out, (const __half*)dout, delta, rows, head_dim);
  else
    return (int)cudaErrorInvalidValue;
  return (int)cudaGetLastError();
}

}  // extern "C"
