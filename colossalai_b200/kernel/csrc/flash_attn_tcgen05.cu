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
  static constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * KV_STAGE_BYTES + P_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 512;
  static constexpr int O_COL = 256;
};

struct FaParams {
  int batch, seqlen_q, seqlen_k;     // equal-length sequences, token-major tensors [batch * seqlen, heads * D]
  int hq, hkv;
  int causal;
  float scale_log2;                  // softmax scale * log2(e)
  void* out;                         // [batch * seqlen_q, hq * D]
  float* lse;                        // [batch * seqlen_q, hq] natural-log LSE
  int out_dtype;
  uint32_t idesc_qk, idesc_pv;
};

SM100_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
__global__ void __launch_bounds__(NUM_THREADS, 1)
flash_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const FaParams p) {
  using C = FaCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_kv = smem + C::Q_BYTES;
  uint8_t* smem_p = smem_kv + KV_STAGES * C::KV_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + C::P_BYTES);
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
  // heaviest (longest causal) query tiles first
  const int q_tiles = p.seqlen_q / BLOCK_Q;
  const int q_tile = p.causal ? (q_tiles - 1 - (int)blockIdx.x) : (int)blockIdx.x;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int kv_head = head / (p.hq / p.hkv);
  const int q_row0 = b * p.seqlen_q + q_tile * BLOCK_Q;            // row in the token-major Q / O tensors
  const int kv_row0 = b * p.seqlen_k;
  // causal with seqlen_q == seqlen_k: key tiles 0..q_tile; otherwise all of them
  const int kv_tiles = p.causal ? (q_tile + 1) : (p.seqlen_k / BLOCK_KV);

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
        const int row = kv_row0 + j * BLOCK_KV;
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
      mbar_wait(p_full, (uint32_t)(j & 1));
      mbar_wait(o_empty, (uint32_t)((j & 1) ^ 1));
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sv = smem_u32(smem_kv + stage * C::KV_STAGE_BYTES + C::K_BYTES);
        const uint64_t dv = make_smem_desc_sw128(sv, BLOCK_KV * 128, 1024);        // MN-major: [keys, D] row-major
#pragma unroll
        for (int k = 0; k < BLOCK_KV / UMMA_K; ++k) {
          const uint32_t ao = (k >> 2) * (128 * 128) + (k & 3) * (UMMA_K * 2);
          const uint32_t bo = k * (UMMA_K * 128);
          umma_f16_ss(tmem_base + C::O_COL, advance_desc(dp, ao), advance_desc(dv, bo), p.idesc_pv, k > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[stage]);                  // K_j and V_j are consumed once P_j V_j has retired
      }
      __syncwarp();
    }
  } else {
    // ================================================================ softmax + output (warps 2..5)
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;                   // query row inside the tile
    const int q_pos = q_tile * BLOCK_Q + r;              // position inside the sequence
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    float o_acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY;                             // running max in the exp2 domain
    float l_run = 0.f;
    uint8_t* p_row = smem_p + r * 128;
    for (int j = 0; j < kv_tiles; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_addr + sb * BLOCK_KV;
      const bool diag = p.causal && (j == kv_tiles - 1);
      const int k_base = j * BLOCK_KV;
      // ---- pass 1: row max
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BLOCK_KV; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(s_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(v[i]) * p.scale_log2;
          if (diag && (k_base + c + i) > q_pos) s = -INFINITY;
          m_tile = fmaxf(m_tile, s);
        }
      }
      const float m_new = fmaxf(m_run, m_tile);
      // a fully masked row cannot happen under the causal layout used here (key 0 is always visible), but keep the
      // exponent finite anyway
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = fast_exp2(m_run - m_safe);
      // ---- pass 2: p = exp2(s - m), row sum, bf16 P into the swizzled K-major tile
      // (the P buffer is free: o_full of tile j-1 - i.e. P_{j-1} V_{j-1} retired - was waited on in the accumulate
      //  step of the previous iteration)
      float l_tile = 0.f;
#pragma unroll 1
      for (int c = 0; c < BLOCK_KV; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(s_addr + c, v);
        tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float s0 = __uint_as_float(v[i]) * p.scale_log2;
          float s1 = __uint_as_float(v[i + 1]) * p.scale_log2;
          if (diag && (k_base + c + i) > q_pos) s0 = -INFINITY;
          if (diag && (k_base + c + i + 1) > q_pos) s1 = -INFINITY;
          const float p0 = fast_exp2(s0 - m_safe);
          const float p1 = fast_exp2(s1 - m_safe);
          l_tile += p0 + p1;
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
      tc_fence_before();
      fence_proxy_async();                               // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[sb]);                       // S buffer may be overwritten by Q K_{j+2}^T
        mbar_arrive(p_full);                             // P_j is in shared memory
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      // ---- O_j = P_j V_j back from TMEM: acc = acc * alpha + O_j
      mbar_wait(o_full, (uint32_t)(j & 1));
      tc_fence_after();
      const uint32_t o_addr = tmem_base + lane_addr + C::O_COL;
#pragma unroll
      for (int c = 0; c < D; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(o_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c + i] = o_acc[c + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);               // next P V may overwrite O_j (and P_j is free again)
    }
    // ---- epilogue: normalise and store this thread's row
    const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
    const size_t row = (size_t)q_row0 + r;
    if (p.out_dtype == CB_BF16) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + row * ((size_t)p.hq * D) + (size_t)head * D;
#pragma unroll
      for (int c = 0; c < D; c += 8) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __nv_bfloat162 h = __floats2bfloat162_rn(o_acc[c + 2 * i] * inv_l, o_acc[c + 2 * i + 1] * inv_l);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(dst + c) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      __half* dst = reinterpret_cast<__half*>(p.out) + row * ((size_t)p.hq * D) + (size_t)head * D;
#pragma unroll
      for (int c = 0; c < D; c += 8) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __half2 h = __floats2half2_rn(o_acc[c + 2 * i] * inv_l, o_acc[c + 2 * i + 1] * inv_l);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(dst + c) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    if (p.lse) p.lse[row * p.hq + head] = (m_run + log2f(l_run)) * 0.6931471805599453f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int D>
int launch_flash_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int batch, int seqlen_q,
                     int seqlen_k, int hq, int hkv, int causal, float scale, int dtype, cudaStream_t stream) {
  using C = FaCfg<D>;
  const bool bf16 = dtype == CB_BF16;
  CUtensorMap tq, tk, tv;
  int r = make_tmap_2d_16b(&tq, q, (uint64_t)batch * seqlen_q, (uint64_t)hq * D, (uint64_t)hq * D, BLOCK_Q, 64, bf16);
  if (r) return 1000 + r;
  r = make_tmap_2d_16b(&tk, k, (uint64_t)batch * seqlen_k, (uint64_t)hkv * D, (uint64_t)hkv * D, BLOCK_KV, 64, bf16);
  if (r) return 2000 + r;
  r = make_tmap_2d_16b(&tv, v, (uint64_t)batch * seqlen_k, (uint64_t)hkv * D, (uint64_t)hkv * D, BLOCK_KV, 64, bf16);
  if (r) return 3000 + r;
  FaParams p;
  p.batch = batch; p.seqlen_q = seqlen_q; p.seqlen_k = seqlen_k; p.hq = hq; p.hkv = hkv; p.causal = causal;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = out; p.lse = lse; p.out_dtype = dtype;
  p.idesc_qk = make_idesc_f16(BLOCK_Q, BLOCK_KV, bf16 ? 1 : 0, 0, 0);      // S[128 x 128] = Q (K-major) x K (K-major)
  p.idesc_pv = make_idesc_f16(BLOCK_Q, D, bf16 ? 1 : 0, 0, 1);              // O[128 x D]  = P (K-major) x V (MN-major)
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(flash_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(seqlen_q / BLOCK_Q, hq, batch);
  flash_fwd_kernel<D><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(tq, tk, tv, p);
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
  if (seqlen_q % BLOCK_Q || seqlen_k % BLOCK_KV || hq % hkv || (causal && seqlen_q != seqlen_k))
    return (int)cudaErrorInvalidValue;
  if (dtype != CB_BF16 && dtype != CB_F16) return (int)cudaErrorInvalidValue;
  if (head_dim == 128)
    return launch_flash_fwd<128>(q, k, v, out, lse, batch, seqlen_q, seqlen_k, hq, hkv, causal, scale, dtype, stream);
  if (head_dim == 64)
    return launch_flash_fwd<64>(q, k, v, out, lse, batch, seqlen_q, seqlen_k, hq, hkv, causal, scale, dtype, stream);
  return (int)cudaErrorInvalidValue;
}

}  // extern "C"
