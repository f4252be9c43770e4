// Paged-KV inference kernels for sm_100a.
//
// KV cache layout (both K and V, per layer):  [num_blocks, block_size, kv_heads, head_dim]  (token-major inside a block,
// 16-byte vectors along head_dim; one (block, head) tile is a strided 2-D box, i.e. directly TMA-addressable).
//
//   cb_kv_cache_write        : scatter this step's K,V rows [tokens, kv_heads, D] into their paged slots (prefill: all
//                              prompt tokens via cu_seqlens; decode: one token per sequence), optional fp8(e5m2) cast
//   cb_rope_kv_cache_write   : fused RoPE(q,k) in place + cache write of rotated K and of V (decode path)
//   cb_paged_decode_attention: split-KV flash-decoding: grid (seq, group of <= 8 kv heads, split); one WARP serves a
//                              kv head (its whole GQA group, so K/V are read once per group) and the warps of a CTA
//                              stream adjacent head slices of the same token rows; fp32 online softmax; second pass
//                              merges splits
//   cb_gather_cos_sin        : per-token cos/sin rows from the [max_pos, D/2] caches
//   cb_convert_fp8           : fp16/bf16/fp32 <-> fp8 e5m2 storage
//
// Capability parity: reference inference_ops_cuda (extensions/csrc/kernel/cuda/{flash_decoding_attention,
// decode_kv_cache_memcpy,context_kv_cache_memcpy,fused_rotary_emb_and_cache,get_cos_and_sin,convert_fp8}_kernel.cu,
// N12-N19) and the Triton twins (kernel/triton/{flash_decoding,kvcache_copy,no_pad_rotary_embedding}.py).
// Decode attention is HBM-bound (every K/V byte is read once): the design goal is full-sector 16-byte loads, many of them
// in flight per SM, and enough CTAs (seqs x kv_heads x splits >= 2 waves of 2 x 148) rather than tensor cores.
#include "common.cuh"

namespace {

constexpr int DEC_THREADS = 256;   // at most 8 (kv head, query-head set) units per CTA
constexpr int MAX_GROUP = 8;   // q heads per kv head handled by one CTA

template <typename T> CB_DEVICE float kv_to_f32(T v) { return to_f32<T>(v); }

// 16-byte asynchronous global -> shared copy; src_bytes < 16 zero-fills the rest (0 = pure zero fill)
CB_DEVICE void cp_async_16(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
CB_DEVICE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
CB_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> CB_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, typename TC>
__global__ void __launch_bounds__(256) kv_cache_write_kernel(const T* __restrict__ k, const T* __restrict__ v,
                                                             TC* __restrict__ k_cache, TC* __restrict__ v_cache,
                                                             const int* __restrict__ block_tables,
                                                             const int* __restrict__ token_seq,   // [tokens] seq id
                                                             const int* __restrict__ token_pos,   // [tokens] position
                                                             int tokens, int kv_heads, int D, int block_size,
                                                             int max_blocks_per_seq, int64_t k_stride,
                                                             int64_t v_stride) {
  const int t = blockIdx.x;
  if (t >= tokens) return;
  const int seq = token_seq[t], pos = token_pos[t];
  const int blk = block_tables[seq * max_blocks_per_seq + pos / block_size];
  const int slot = pos % block_size;
  const int64_t dst = ((int64_t)blk * block_size + slot) * kv_heads * D;
  const int n = kv_heads * D;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    k_cache[dst + i] = (TC)k[(int64_t)t * k_stride + i];
    v_cache[dst + i] = (TC)v[(int64_t)t * v_stride + i];
  }
}

// RoPE (half rotation) on q [tokens, Hq, D] and k [tokens, Hkv, D] in place, then K,V -> paged cache.
template <typename T>
__global__ void __launch_bounds__(256) rope_kv_cache_write_kernel(T* __restrict__ q, T* __restrict__ k,
                                                                  const T* __restrict__ v, T* __restrict__ k_cache,
                                                                  T* __restrict__ v_cache,
                                                                  const float* __restrict__ cos_c,
                                                                  const float* __restrict__ sin_c,
                                                                  const int* __restrict__ block_tables,
                                                                  const int* __restrict__ token_seq,
                                                                  const int* __restrict__ token_pos, int tokens, int Hq,
                                                                  int Hkv, int D, int rot, int block_size,
                                                                  int max_blocks_per_seq, int64_t q_stride,
                                                                  int64_t k_stride, int64_t v_stride) {
  const int t = blockIdx.x;
  if (t >= tokens) return;
  const int seq = token_seq[t], pos = token_pos[t];
  const int half = rot / 2;
  const float* cp = cos_c + (int64_t)pos * half;
  const float* sp = sin_c + (int64_t)pos * half;
  T* qt = q + (int64_t)t * q_stride;
  T* kt = k + (int64_t)t * k_stride;
  for (int i = threadIdx.x; i < (Hq + Hkv) * half; i += blockDim.x) {
    const int h = i / half, j = i - h * half;
    T* base = h < Hq ? qt + h * D : kt + (h - Hq) * D;
    const float a = to_f32<T>(base[j]), b = to_f32<T>(base[j + half]), c = cp[j], s = sp[j];
    base[j] = from_f32<T>(a * c - b * s);
    base[j + half] = from_f32<T>(b * c + a * s);
  }
  __syncthreads();
  const int blk = block_tables[seq * max_blocks_per_seq + pos / block_size];
  const int64_t dst = ((int64_t)blk * block_size + pos % block_size) * Hkv * D;
  for (int i = threadIdx.x; i < Hkv * D; i += blockDim.x) {
    k_cache[dst + i] = kt[i];
    v_cache[dst + i] = v[(int64_t)t * v_stride + i];
  }
}

// 16-byte vectorised variants (the hot ones): one thread moves one 16-byte vector, grid = all vectors of the step, so a
// prefill of thousands of tokens is a single HBM-rate streaming kernel instead of one small CTA per token.
template <typename T>
__global__ void __launch_bounds__(256) kv_cache_write_vec_kernel(const T* __restrict__ k, const T* __restrict__ v,
                                                                 T* __restrict__ k_cache, T* __restrict__ v_cache,
                                                                 const int* __restrict__ block_tables,
                                                                 const int* __restrict__ token_seq,
                                                                 const int* __restrict__ token_pos, int tokens,
                                                                 int vec_per_tok, int block_size, int max_blocks_per_seq,
                                                                 int64_t k_stride, int64_t v_stride) {
  // vec_per_tok = kv_heads * D / 8 (16-bit elements); work items = tokens * 2 (K, V) * vec_per_tok
  const int64_t total = (int64_t)tokens * 2 * vec_per_tok;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(w / (2 * vec_per_tok));
    const int r = (int)(w - (int64_t)t * 2 * vec_per_tok);
    const bool is_v = r >= vec_per_tok;
    const int i = is_v ? r - vec_per_tok : r;
    const int seq = __ldg(token_seq + t), pos = __ldg(token_pos + t);
    const int blk = __ldg(block_tables + seq * max_blocks_per_seq + pos / block_size);
    const int64_t dst = ((int64_t)blk * block_size + pos % block_size) * vec_per_tok + i;       // in 16-byte units
    const uint4 val = is_v ? __ldg(reinterpret_cast<const uint4*>(v + (int64_t)t * v_stride) + i)
                           : __ldg(reinterpret_cast<const uint4*>(k + (int64_t)t * k_stride) + i);
    reinterpret_cast<uint4*>(is_v ? v_cache : k_cache)[dst] = val;
  }
}

// RoPE (half rotation, rot == D) on q and k in place + rotated K and V into the paged cache, 16-byte vectors: a work
// item is one (token, head, 8-element group of the FIRST half) - it owns the matching group of the second half too -
// or one V vector.  K is written to the cache from registers (no second pass, no block barrier).
template <typename T>
__global__ void __launch_bounds__(256) rope_kv_cache_write_vec_kernel(
    T* __restrict__ q, T* __restrict__ k, const T* __restrict__ v, T* __restrict__ k_cache, T* __restrict__ v_cache,
    const float* __restrict__ cos_c, const float* __restrict__ sin_c, const int* __restrict__ block_tables,
    const int* __restrict__ token_seq, const int* __restrict__ token_pos, int tokens, int Hq, int Hkv, int D,
    int block_size, int max_blocks_per_seq, int64_t q_stride, int64_t k_stride, int64_t v_stride) {
  const int half = D / 2;
  const int gph = half / 8;                          // 8-element groups per half head
  const int rope_items = (Hq + Hkv) * gph;
  const int v_items = Hkv * D / 8;
  const int per_tok = rope_items + v_items;
  const int64_t total = (int64_t)tokens * per_tok;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(w / per_tok);
    const int r = (int)(w - (int64_t)t * per_tok);
    const int seq = __ldg(token_seq + t), pos = __ldg(token_pos + t);
    const int blk = __ldg(block_tables + seq * max_blocks_per_seq + pos / block_size);
    const int64_t slot = ((int64_t)blk * block_size + pos % block_size) * Hkv * D;              // in elements
    if (r >= rope_items) {
      const int i = r - rope_items;
      reinterpret_cast<uint4*>(v_cache + slot)[i] = __ldg(reinterpret_cast<const uint4*>(v + (int64_t)t * v_stride) + i);
      continue;
    }
    const int h = r / gph, g = r - h * gph;
    T* base = h < Hq ? q + (int64_t)t * q_stride + (int64_t)h * D : k + (int64_t)t * k_stride + (int64_t)(h - Hq) * D;
    Vec16<T> a, b2;
    a.load(base + g * 8);
    b2.load(base + half + g * 8);
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(cos_c + (int64_t)pos * half + g * 8));
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(cos_c + (int64_t)pos * half + g * 8 + 4));
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(sin_c + (int64_t)pos * half + g * 8));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(sin_c + (int64_t)pos * half + g * 8 + 4));
    const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    Vec16<T> oa, ob;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x = a.get(e), y = b2.get(e);
      oa.set(e, x * cs[e] - y * sn[e]);
      ob.set(e, y * cs[e] + x * sn[e]);
    }
    oa.store(base + g * 8);
    ob.store(base + half + g * 8);
    if (h >= Hq) {
      T* kc = k_cache + slot + (int64_t)(h - Hq) * D;
      oa.store(kc + g * 8);
      ob.store(kc + half + g * 8);
    }
  }
}

// One CTA: sequence `seq`, kv head `kvh`, KV partition `split`.  q: [num_seqs, Hq, D].
// partial outputs: o_part [num_seqs, Hq, splits, D] fp32, ml_part [num_seqs, Hq, splits, 2] (max, sumexp)
//
// Work mapping (the kernel is HBM-bound: every K/V byte is read once, so the job is to keep many full-sector loads in
// flight per SM and to spend few instructions per byte):
//   * a K or V row of one (token, kv head) is D contiguous elements; LPT = D/16 adjacent lanes cover one row with two
//     16-byte loads each, so a warp reads 32/LPT whole rows per pass (fully used 32-byte sectors, 128-byte lines);
//   * the rows travel through a per-warp cp.async ring in shared memory (DEC_STAGES stages of DEC_U passes = 4 KB each):
//     3 stages = 12 KB per warp, ~96 KB per SM stay in flight while one stage is consumed - enough to cover the HBM
//     latency at full bandwidth, independent of the register budget (the first two versions kept the rows in
//     registers: 8 warps / SM at 255 registers, 43 % issue-active, 2.4 TB/s under ncu - load and FMA phases of a warp
//     alternated, then only 32 KB per SM in flight); no block-wide barrier in the main loop;
//   * the query rows live in registers (GH = 4 query heads per warp, the lane's 16 dims of each); a GQA group wider
//     than 4 takes two warps of the CTA (the second read of a K/V row by the sibling warp is an L1/L2 hit);
//   * every lane group keeps its OWN online-softmax state (m, l, o) for the tokens it sees - no cross-lane traffic in
//     the loop except the log2(LPT) shuffles that finish a dot product; the 32/LPT states of a warp are merged once at
//     the end and written as this (head set, partition)'s partial result - no shared-memory reduction, no barrier.
constexpr int DEC_U = 2;    // passes of the warp per pipeline stage
constexpr int DEC_STAGES = 4;   // shared-memory ring depth per warp (DEC_STAGES - 1 stages in flight)
constexpr int DEC_GH = 4;   // query heads per warp

template <typename T, typename TC, int D>
__global__ void __launch_bounds__(DEC_THREADS, 1) paged_decode_kernel(
    const T* __restrict__ q, const TC* __restrict__ k_cache, const TC* __restrict__ v_cache,
    const int* __restrict__ block_tables, const int* __restrict__ seq_lens, float* __restrict__ o_part,
    float* __restrict__ ml_part, int Hq, int Hkv, int block_size, int max_blocks_per_seq, int splits, int part_len,
    float scale, const float* __restrict__ alibi_slopes, int64_t q_stride, int window) {
  static_assert(sizeof(TC) == 2, "the paged cache is read as 16-byte vectors of 16-bit elements");
  constexpr int EPL = 16;                       // elements of a row owned by a lane
  constexpr int LPT = D / EPL;                  // lanes per token row: 4 / 8 / 16
  constexpr int TPW = 32 / LPT;                 // token rows per warp pass
  constexpr int GH = DEC_GH, U = DEC_U;
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  // CTA = (sequence, group of kv heads, KV partition); warp = one (kv head, set of <= 4 query heads) unit walking ALL
  // tokens of the partition.  The warps of a CTA read ADJACENT 256-byte head slices of the same token rows at the same
  // time, so the CTA streams contiguous [tokens x heads x D] spans of the cache (with one CTA per kv head the 2 KB
  // stride between a head's rows cost ~60 % of the DRAM bandwidth: 2.4 TB/s whatever the number of bytes in flight).
  const int seq = blockIdx.x, split = blockIdx.z;
  const int G = Hq / Hkv;
  const int nsets = (G + GH - 1) / GH;          // sets of <= 4 query heads per kv head
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.y * (blockDim.x >> 5) + warp;
  const int kvh = unit / nsets, set = unit - kvh * nsets;
  if (kvh >= Hkv) return;                       // (no block-wide barrier anywhere in this kernel)
  const int len = seq_lens[seq];
  // sliding-window attention (Mistral): only the last `window` cached tokens are visible - partitions that lie before
  // the window start contribute nothing (m = -inf, l = 0) and are skipped by the merge
  const int window_start = window > 0 ? max(0, len - window) : 0;
  const int t0 = max(split * part_len, window_start), t1 = min(len, split * part_len + part_len);
  const int g0 = set * GH;
  const int ng = min(GH, G - g0);
  const int grp = lane / LPT, sub = lane - grp * LPT;

  float qr[GH][EPL];
#pragma unroll
  for (int g = 0; g < GH; ++g) {
    const T* qp = q + (int64_t)seq * q_stride + (int64_t)(kvh * G + g0 + (g < ng ? g : 0)) * D + sub * EPL;
#pragma unroll
    for (int e = 0; e < EPL; ++e) qr[g][e] = g < ng ? to_f32<T>(qp[e]) * (scale * LOG2E) : 0.f;   // scores in the exp2 domain
  }
  float slope[GH];
#pragma unroll
  for (int g = 0; g < GH; ++g) slope[g] = (alibi_slopes && g < ng) ? alibi_slopes[kvh * G + g0 + g] * LOG2E : 0.f;
  const bool has_alibi = alibi_slopes != nullptr;
  float m[GH], l[GH], o[GH][EPL];
#pragma unroll
  for (int g = 0; g < GH; ++g) {
    m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) o[g][e] = 0.f;
  }
  const int* bt = block_tables + seq * max_blocks_per_seq;
  const int bs_shift = (block_size & (block_size - 1)) == 0 ? 31 - __clz(block_size) : -1;   // power-of-two blocks: no division
  // K/V travel through a per-warp ring of DEC_STAGES shared-memory stages filled with cp.async (16 bytes per lane and
  // vector; a stage = U passes of the warp = U * TPW token rows, K and V): the bytes in flight are bounded by shared
  // memory, not by registers, so DEC_STAGES - 1 stages (6 KB) per warp stay outstanding while one is consumed.  Every
  // lane reads back exactly the vectors it requested itself (slot = [stage][k|v][pass][half][lane]), so no barrier is
  // needed beyond cp.async.wait_group, and the 16-byte slots of a warp are consecutive (conflict-free LDS.128).
  extern __shared__ __align__(16) unsigned char dec_smem[];
  constexpr int NST = DEC_STAGES;
  constexpr int STAGE_VECS = 2 * U * 2 * 32;                      // uint4 slots per stage: (K|V) x pass x half x lane
  uint4* ring = reinterpret_cast<uint4*>(dec_smem) + (size_t)warp * NST * STAGE_VECS;
  auto issue_stage = [&](int base, int stage) {
    uint4* st = ring + stage * STAGE_VECS;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = base + u * TPW + grp;
      const bool ok = t < t1;
      const int tt = ok ? t : t0;                                 // any valid row: src-size 0 zero-fills instead
      const int blk = bt[bs_shift >= 0 ? (tt >> bs_shift) : (tt / block_size)];
      const int slot = bs_shift >= 0 ? (tt & (block_size - 1)) : (tt % block_size);
      const int64_t row = (((int64_t)blk * block_size + slot) * Hkv + kvh) * D + sub * EPL;
      const uint32_t nbytes = ok ? 16u : 0u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        cp_async_16(st + ((0 * U + u) * 2 + h) * 32 + lane, k_cache + row + h * 8, nbytes);
        cp_async_16(st + ((1 * U + u) * 2 + h) * 32 + lane, v_cache + row + h * 8, nbytes);
      }
    }
  };
  auto consume = [&](int base, int stage) {
    const uint4* st = ring + stage * STAGE_VECS;
    float sc[U][GH];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      Vec16<TC> k0, k1;
      k0.raw = st[((0 * U + u) * 2 + 0) * 32 + lane];
      k1.raw = st[((0 * U + u) * 2 + 1) * 32 + lane];
      const int t = base + u * TPW + grp;
#pragma unroll
      for (int g = 0; g < GH; ++g) {
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc += qr[g][e] * k0.get(e);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc += qr[g][8 + e] * k1.get(e);
#pragma unroll
        for (int off = LPT / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (has_alibi) acc += slope[g] * (float)(t - (len - 1));
        sc[u][g] = (t < t1) ? acc : -INFINITY;
      }
    }
    float pr[U][GH];
#pragma unroll
    for (int g = 0; g < GH; ++g) {
      float nm = m[g];
#pragma unroll
      for (int u = 0; u < U; ++u) nm = fmaxf(nm, sc[u][g]);
      if (__any_sync(0xffffffffu, nm > m[g])) {                   // the running max moved for some row group: rescale
        const float corr = (nm == -INFINITY) ? 1.f : ex2_approx(m[g] - nm);
        l[g] *= corr;
#pragma unroll
        for (int e = 0; e < EPL; ++e) o[g][e] *= corr;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        pr[u][g] = (sc[u][g] == -INFINITY) ? 0.f : ex2_approx(sc[u][g] - nm);
        l[g] += pr[u][g];
      }
      m[g] = nm;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      Vec16<TC> v0, v1;
      v0.raw = st[((1 * U + u) * 2 + 0) * 32 + lane];
      v1.raw = st[((1 * U + u) * 2 + 1) * 32 + lane];
#pragma unroll
      for (int g = 0; g < GH; ++g) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[g][e] += pr[u][g] * v0.get(e);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[g][8 + e] += pr[u][g] * v1.get(e);
      }
    }
  };
  {
    const int step = TPW * U;
    const int first = t0;
    // prologue: NST - 1 stages in flight (empty groups keep the group count uniform at the tail)
#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
      if (first + s * step < t1) issue_stage(first + s * step, s);
      cp_async_commit();
    }
    int it = 0;
    for (int base = first; base < t1; base += step, ++it) {
      const int ahead = base + (NST - 1) * step;
      if (ahead < t1) issue_stage(ahead, (it + NST - 1) % NST);
      cp_async_commit();
      cp_async_wait<NST - 1>();                                   // the group of stage `it` has landed
      consume(base, it % NST);
    }
    cp_async_wait<0>();
  }
  // merge the TPW lane-group states of the warp (butterfly over the group index)
#pragma unroll
  for (int off = LPT; off < 32; off <<= 1) {
#pragma unroll
    for (int g = 0; g < GH; ++g) {
      const float mo = __shfl_xor_sync(0xffffffffu, m[g], off);
      const float lo = __shfl_xor_sync(0xffffffffu, l[g], off);
      const float nm = fmaxf(m[g], mo);
      const float c1 = (m[g] == -INFINITY) ? 0.f : ex2_approx(m[g] - nm);
      const float c2 = (mo == -INFINITY) ? 0.f : ex2_approx(mo - nm);
      l[g] = l[g] * c1 + lo * c2;
#pragma unroll
      for (int e = 0; e < EPL; ++e) o[g][e] = o[g][e] * c1 + __shfl_xor_sync(0xffffffffu, o[g][e], off) * c2;
      m[g] = nm;
    }
  }
  // every lane group now holds the merged state; group 0 writes this unit's partial result
  if (grp == 0) {
#pragma unroll
    for (int g = 0; g < GH; ++g) {
      if (g < ng) {
        const int64_t oi = (((int64_t)seq * Hq + kvh * G + g0 + g) * splits + split);
        float4* dst = reinterpret_cast<float4*>(o_part + oi * D + sub * EPL);
#pragma unroll
        for (int e = 0; e < EPL; e += 4) dst[e >> 2] = make_float4(o[g][e], o[g][e + 1], o[g][e + 2], o[g][e + 3]);
        if (sub == 0) { ml_part[oi * 2] = m[g] * LN2; ml_part[oi * 2 + 1] = l[g]; }   // max back in natural-log units
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Tensor-core formulation of the same split-KV decode step (the default; `CB200_DECODE=simt` selects the kernel above).
//
// The CUDA-core kernel above executes ~89 warp-instructions per (token, kv head) - two FMAs per cached element and q
// head - which is as much issue time as the HBM transfer takes (ncu: 93 M warp-instructions, 45 % issue-active,
// 2.2-2.4 TB/s).  Here the GQA group of a kv head (<= 8 q heads, padded to the 16 rows of an MMA) is the M dimension of
// `mma.sync.m16n8k16`:  S[heads x 16 tokens] = Q K^T  and  O[heads x D] += P V  per 16-token tile, i.e. 32 MMAs instead of
// ~4000 FMAs per tile and lane.
//   * warp = (sequence, kv head, partition) unit, as above; Q fragments (A operand) live in registers;
//   * a tile of 16 token rows of K and of V (16 x D 16-bit values each) is fetched with cp.async into a per-warp ring
//     (DEC_MMA_STAGES deep); a row's 16-byte chunks are stored XOR-swizzled by (token mod 8), so the ldmatrix reads - K
//     plain (B operand of Q K^T), V transposed (B operand of P V) - are bank-conflict free;
//   * softmax state per head row is replicated over the four lanes of a quad; the probabilities go straight from the
//     accumulator layout of S to the A-operand layout of P V (no shared-memory round trip); the output accumulator is
//     rescaled only when a row max moved;
//   * partial results use the same (o_part, ml_part) format, so the reduce kernel is shared.
constexpr int DEC_MMA_TOKENS = 16;
constexpr int DEC_MMA_STAGES = 3;

CB_DEVICE void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
CB_DEVICE void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
template <typename T> CB_DEVICE void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> CB_DEVICE void mma_16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> CB_DEVICE void mma_16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <typename T> CB_DEVICE uint32_t pack2(float lo, float hi);
template <> CB_DEVICE uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> CB_DEVICE uint32_t pack2<__half>(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <typename T, int D>
__global__ void __launch_bounds__(DEC_THREADS, 1) paged_decode_mma_kernel(
    const T* __restrict__ q, const T* __restrict__ k_cache, const T* __restrict__ v_cache,
    const int* __restrict__ block_tables, const int* __restrict__ seq_lens, float* __restrict__ o_part,
    float* __restrict__ ml_part, int Hq, int Hkv, int block_size, int max_blocks_per_seq, int splits, int part_len,
    float scale, const float* __restrict__ alibi_slopes, int64_t q_stride, int window) {
  constexpr int NT = DEC_MMA_TOKENS, NST = DEC_MMA_STAGES;
  constexpr int KS = D / 16;                    // k-steps of Q K^T
  constexpr int NO = D / 8;                     // 8-wide column tiles of the output
  constexpr int CPR = D / 8;                    // 16-byte chunks per token row
  constexpr int ROW_BYTES = D * 2;
  constexpr int TILE_BYTES = NT * ROW_BYTES;    // one operand (K or V) of one stage
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const int seq = blockIdx.x, split = blockIdx.z;
  const int G = Hq / Hkv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvh = blockIdx.y * (blockDim.x >> 5) + warp;
  if (kvh >= Hkv) return;                       // (no block-wide barrier in this kernel)
  const int len = seq_lens[seq];
  const int window_start = window > 0 ? max(0, len - window) : 0;
  const int t0 = max(split * part_len, window_start), t1 = min(len, split * part_len + part_len);
  const int row = lane >> 2, quad = lane & 3;   // accumulator layout: head row, column pair inside an 8-wide tile
  const bool row_ok = row < G;                  // rows G..15 of the MMA are padding
  const float scale2 = scale * LOG2E;           // scores in the exp2 domain
  const float slope2 = (alibi_slopes != nullptr && row_ok) ? alibi_slopes[kvh * G + row] * LOG2E : 0.f;
  const bool has_alibi = alibi_slopes != nullptr;

  // ---- Q fragments: A operand [16 heads x 16 d] per k-step; a1 / a3 are head rows 8..15 = zero (G <= 8)
  uint32_t qa[KS][4];
  {
    const T* qrow = q + (int64_t)seq * q_stride + (int64_t)(kvh * G + (row_ok ? row : 0)) * D;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const uint32_t lo = *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + quad * 2);
      const uint32_t hi = *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8 + quad * 2);
      qa[ks][0] = row_ok ? lo : 0u; qa[ks][1] = 0u; qa[ks][2] = row_ok ? hi : 0u; qa[ks][3] = 0u;
    }
  }
  float o[NO][4];
#pragma unroll
  for (int j = 0; j < NO; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m_run = -INFINITY, l_run = 0.f;         // per head row, replicated over the quad (l: this lane's share)

  extern __shared__ __align__(128) unsigned char dec_smem[];
  unsigned char* ring = dec_smem + (size_t)warp * NST * 2 * TILE_BYTES;
  const int* bt = block_tables + seq * max_blocks_per_seq;
  const int bs_shift = (block_size & (block_size - 1)) == 0 ? 31 - __clz(block_size) : -1;
  auto issue_stage = [&](int base, int stage) {
    unsigned char* kt = ring + (size_t)stage * 2 * TILE_BYTES;
    unsigned char* vt = kt + TILE_BYTES;
#pragma unroll
    for (int i = 0; i < NT * CPR / 32; ++i) {
      const int cid = lane + 32 * i;
      const int tr = cid / CPR, c = cid - tr * CPR;          // token row inside the tile, 16-byte chunk inside the row
      const int t = base + tr;
      const bool ok = t < t1;
      const int tt = ok ? t : t0;
      const int blk = bt[bs_shift >= 0 ? (tt >> bs_shift) : (tt / block_size)];
      const int slot = bs_shift >= 0 ? (tt & (block_size - 1)) : (tt % block_size);
      const int64_t src = (((int64_t)blk * block_size + slot) * Hkv + kvh) * D + c * 8;
      const int dst = tr * ROW_BYTES + ((c ^ (tr & 7)) << 4);
      const uint32_t nbytes = ok ? 16u : 0u;
      cp_async_16(kt + dst, k_cache + src, nbytes);
      cp_async_16(vt + dst, v_cache + src, nbytes);
    }
  };
  auto consume = [&](int base, int stage) {
    const unsigned char* kt = ring + (size_t)stage * 2 * TILE_BYTES;
    const unsigned char* vt = kt + TILE_BYTES;
    // ---- S = Q K^T for 16 tokens (two 8-token column tiles)
    float sacc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
      const int tr = nt * 8 + (lane & 7);                    // token row whose address this lane supplies
#pragma unroll
      for (int ks = 0; ks < KS; ks += 2) {
        // matrices: (ks, d 0..7), (ks, d 8..15), (ks + 1, d 0..7), (ks + 1, d 8..15)
        const int c = 2 * ks + (lane >> 3);
        uint32_t b[4];
        ldmatrix_x4(b, kt + tr * ROW_BYTES + ((c ^ (tr & 7)) << 4));
        mma_16816<T>(sacc[nt], qa[ks], b[0], b[1]);
        mma_16816<T>(sacc[nt], qa[ks + 1], b[2], b[3]);
      }
    }
    // ---- online softmax of head row `row` (this lane: tokens nt * 8 + quad * 2 + {0, 1})
    float sv[4];
    float mx = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int t = base + nt * 8 + quad * 2 + e;
        float x = sacc[nt][e] * scale2;
        if (has_alibi) x = fmaf(slope2, (float)(t - (len - 1)), x);
        x = (t < t1) ? x : -INFINITY;
        sv[nt * 2 + e] = x;
        mx = fmaxf(mx, x);
      }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);
    if (__any_sync(0xffffffffu, m_new > m_run)) {
      const float corr = (m_new == -INFINITY) ? 1.f : ex2_approx(m_run - m_new);
      l_run *= corr;
#pragma unroll
      for (int j = 0; j < NO; ++j) { o[j][0] *= corr; o[j][1] *= corr; }
    }
    m_run = m_new;
    float pv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pv[i] = (sv[i] == -INFINITY) ? 0.f : ex2_approx(sv[i] - m_new);
      l_run += pv[i];
    }
    // P as A operand [16 heads x 16 tokens]: a0 = tokens 0..7 of this row, a2 = tokens 8..15; rows 8..15 are zero
    uint32_t pa[4];
    pa[0] = row_ok ? pack2<T>(pv[0], pv[1]) : 0u; pa[1] = 0u;
    pa[2] = row_ok ? pack2<T>(pv[2], pv[3]) : 0u; pa[3] = 0u;
    // ---- O += P V: V^T fragments through ldmatrix.trans
    {
      const int tr = (lane & 7) + ((lane >> 3) & 1) * 8;     // matrices: (tok 0..7, j), (tok 8..15, j), (0..7, j+1), (8..15, j+1)
#pragma unroll
      for (int j = 0; j < NO; j += 2) {
        const int c = j + (lane >> 4);
        uint32_t b[4];
        ldmatrix_x4_trans(b, vt + tr * ROW_BYTES + ((c ^ (tr & 7)) << 4));
        mma_16816<T>(o[j], pa, b[0], b[1]);
        mma_16816<T>(o[j + 1], pa, b[2], b[3]);
      }
    }
  };
  {
#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
      if (t0 + s * NT < t1) issue_stage(t0 + s * NT, s);
      cp_async_commit();
    }
    int it = 0;
    for (int base = t0; base < t1; base += NT, ++it) {
      const int ahead = base + (NST - 1) * NT;
      if (ahead < t1) issue_stage(ahead, (it + NST - 1) % NST);
      cp_async_commit();
      cp_async_wait<NST - 1>();
      __syncwarp();                                          // the tile was written by all lanes of the warp
      consume(base, it % NST);
      __syncwarp();                                          // all lanes done reading before the stage is refilled
    }
    cp_async_wait<0>();
  }
  // ---- this unit's partial result (row sum over the quad first)
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
  if (row_ok) {
    const int64_t oi = (((int64_t)seq * Hq + kvh * G + row) * splits + split);
    float* dst = o_part + oi * D + quad * 2;
#pragma unroll
    for (int j = 0; j < NO; ++j) *reinterpret_cast<float2*>(dst + j * 8) = make_float2(o[j][0], o[j][1]);
    if (quad == 0) { ml_part[oi * 2] = m_run * LN2; ml_part[oi * 2 + 1] = l_run; }
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(128) decode_reduce_kernel(const float* __restrict__ o_part,
                                                            const float* __restrict__ ml_part, T* __restrict__ out,
                                                            int Hq, int splits, int64_t out_stride) {
  const int seq = blockIdx.x, h = blockIdx.y;
  const int64_t base = ((int64_t)seq * Hq + h) * splits;
  float gm = -INFINITY;
  for (int s = 0; s < splits; ++s) gm = fmaxf(gm, ml_part[(base + s) * 2]);
  float ls = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float mm = ml_part[(base + s) * 2];
    ls += (mm == -INFINITY) ? 0.f : ml_part[(base + s) * 2 + 1] * __expf(mm - gm);
  }
  const float inv = ls > 0.f ? 1.f / ls : 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) {
      const float mm = ml_part[(base + s) * 2];
      if (mm != -INFINITY) acc += o_part[(base + s) * D + d] * __expf(mm - gm);
    }
    out[(int64_t)seq * out_stride + (int64_t)h * D + d] = from_f32<T>(acc * inv);
  }
}

__global__ void gather_cos_sin_kernel(const float* __restrict__ cos_c, const float* __restrict__ sin_c,
                                      const int* __restrict__ pos, float* __restrict__ cos_o,
                                      float* __restrict__ sin_o, int tokens, int half) {
  const int t = blockIdx.x;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    cos_o[(int64_t)t * half + i] = cos_c[(int64_t)pos[t] * half + i];
    sin_o[(int64_t)t * half + i] = sin_c[(int64_t)pos[t] * half + i];
  }
}

template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (TO)(float)in[i];
}

}  // namespace

extern "C" {

// k, v: [tokens, kv_heads, D] (row strides in elements).  cache dtype == input dtype (fp8 path: cb_convert_fp8 first).
int cb_kv_cache_write(const void* k, const void* v, void* k_cache, void* v_cache, const int* block_tables,
                      const int* token_seq, const int* token_pos, int tokens, int kv_heads, int D, int block_size,
                      int max_blocks_per_seq, int64_t k_stride, int64_t v_stride, int dtype, cudaStream_t s) {
  if (tokens == 0) return 0;
  const int n = kv_heads * D;
  if (n % 8 == 0 && k_stride % 8 == 0 && v_stride % 8 == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 &&
      ((uintptr_t)k_cache % 16) == 0 && ((uintptr_t)v_cache % 16) == 0) {
    const int64_t total = (int64_t)tokens * 2 * (n / 8);
    const int grid = (int)((total + 255) / 256 < 16 * cb_num_sms() ? (total + 255) / 256 : 16 * cb_num_sms());
    CB_DISPATCH_HALF(dtype, T, {
      kv_cache_write_vec_kernel<T><<<grid, 256, 0, s>>>((const T*)k, (const T*)v, (T*)k_cache, (T*)v_cache, block_tables,
                                                       token_seq, token_pos, tokens, n / 8, block_size,
                                                       max_blocks_per_seq, k_stride, v_stride);
    });
    return CB_LAUNCH_CHECK();
  }
  CB_DISPATCH_HALF(dtype, T, {
    kv_cache_write_kernel<T, T><<<tokens, 256, 0, s>>>((const T*)k, (const T*)v, (T*)k_cache, (T*)v_cache, block_tables,
                                                      token_seq, token_pos, tokens, kv_heads, D, block_size,
                                                      max_blocks_per_seq, k_stride, v_stride);
  });
  return CB_LAUNCH_CHECK();
}

int cb_rope_kv_cache_write(void* q, void* k, const void* v, void* k_cache, void* v_cache, const float* cos_c,
                           const float* sin_c, const int* block_tables, const int* token_seq, const int* token_pos,
                           int tokens, int Hq, int Hkv, int D, int rot, int block_size, int max_blocks_per_seq,
                           int64_t q_stride, int64_t k_stride, int64_t v_stride, int dtype, cudaStream_t s) {
  if (tokens == 0) return 0;
  if (rot == D && D % 16 == 0 && q_stride % 8 == 0 && k_stride % 8 == 0 && v_stride % 8 == 0 &&
      ((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 &&
      ((uintptr_t)k_cache % 16) == 0 && ((uintptr_t)v_cache % 16) == 0 && ((uintptr_t)cos_c % 16) == 0 &&
      ((uintptr_t)sin_c % 16) == 0) {
    const int64_t total = (int64_t)tokens * ((Hq + Hkv) * (D / 16) + Hkv * D / 8);
    const int grid = (int)((total + 255) / 256 < 16 * cb_num_sms() ? (total + 255) / 256 : 16 * cb_num_sms());
    CB_DISPATCH_HALF(dtype, T, {
      rope_kv_cache_write_vec_kernel<T><<<grid, 256, 0, s>>>((T*)q, (T*)k, (const T*)v, (T*)k_cache, (T*)v_cache, cos_c,
                                                            sin_c, block_tables, token_seq, token_pos, tokens, Hq, Hkv, D,
                                                            block_size, max_blocks_per_seq, q_stride, k_stride, v_stride);
    });
    return CB_LAUNCH_CHECK();
  }
  CB_DISPATCH_HALF(dtype, T, {
    rope_kv_cache_write_kernel<T><<<tokens, 256, 0, s>>>((T*)q, (T*)k, (const T*)v, (T*)k_cache, (T*)v_cache, cos_c,
                                                        sin_c, block_tables, token_seq, token_pos, tokens, Hq, Hkv, D,
                                                        rot, block_size, max_blocks_per_seq, q_stride, k_stride,
                                                        v_stride);
  });
  return CB_LAUNCH_CHECK();
}

int cb_decode_num_splits(int num_seqs, int kv_heads, int max_len, int* part_len_out) {
  // one CTA (up to 8 warps, one per kv head) per SM: aim at >= 2 waves of CTAs, partitions of >= 64 tokens
  const int ctas_per_split = num_seqs * ((kv_heads + 7) / 8);
  const int target = 2 * cb_num_sms();
  int splits = (target + ctas_per_split - 1) / ctas_per_split;
  const int max_splits = (max_len + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int part = (max_len + splits - 1) / splits;
  part = ((part + 31) / 32) * 32;
  splits = (max_len + part - 1) / part;
  if (splits < 1) splits = 1;
  *part_len_out = part;
  return splits;
}

// q: [num_seqs, Hq, D]; out same; caches [nb, bs, Hkv, D]; o_part fp32 [num_seqs, Hq, splits, D]; ml_part [.., 2]
int cb_paged_decode_attention(const void* q, const void* k_cache, const void* v_cache, const int* block_tables,
                              const int* seq_lens, void* out, float* o_part, float* ml_part, int num_seqs, int Hq,
                              int Hkv, int D, int block_size, int max_blocks_per_seq, int splits, int part_len,
                              float scale, const float* alibi_slopes, int64_t q_stride, int64_t out_stride, int dtype,
                              int window, cudaStream_t s) {
  if (num_seqs == 0) return 0;
  if (Hq % Hkv != 0 || Hq / Hkv > MAX_GROUP) return (int)cudaErrorInvalidValue;
  const int G = Hq / Hkv;
  const int units = Hkv * ((G + DEC_GH - 1) / DEC_GH);             // (kv head, query-head set) pairs = warps needed
  const int warps = units < DEC_THREADS / 32 ? units : DEC_THREADS / 32;
  dim3 grid(num_seqs, (units + warps - 1) / warps, splits);
  const int dec_smem = warps * DEC_STAGES * (2 * DEC_U * 2 * 32) * 16;                    // per-warp cp.async rings
  constexpr int dec_smem_max = (DEC_THREADS / 32) * DEC_STAGES * (2 * DEC_U * 2 * 32) * 16;
  // default: the tensor-core formulation (mma.sync over the GQA group) where it applies (head_dim 64 / 128, groups of
  // <= 8 q heads); CB200_DECODE=simt forces the CUDA-core kernel (also used for head_dim 256)
  const char* impl_env = getenv("CB200_DECODE");
  const bool use_mma = !(impl_env != nullptr && impl_env[0] == 's') && G <= 8 && (D == 128 || D == 64);
  const int mma_warps = Hkv < DEC_THREADS / 32 ? Hkv : DEC_THREADS / 32;
  dim3 mma_grid(num_seqs, (Hkv + mma_warps - 1) / mma_warps, splits);
#define LAUNCH_DEC_MMA(T, DD)                                                                                       \
  {                                                                                                                  \
    constexpr int per_warp = DEC_MMA_STAGES * 2 * DEC_MMA_TOKENS * DD * 2;                                           \
    static bool attr_done = false;                                                                                   \
    if (!attr_done) {                                                                                                \
      cudaFuncSetAttribute(paged_decode_mma_kernel<T, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                           (DEC_THREADS / 32) * per_warp);                                                           \
      attr_done = true;                                                                                              \
    }                                                                                                                \
    paged_decode_mma_kernel<T, DD><<<mma_grid, mma_warps * 32, mma_warps * per_warp, s>>>(                           \
        (const T*)q, (const T*)k_cache, (const T*)v_cache, block_tables, seq_lens, o_part, ml_part, Hq, Hkv,        \
        block_size, max_blocks_per_seq, splits, part_len, scale, alibi_slopes, q_stride, window);                   \
    decode_reduce_kernel<T, DD><<<dim3(num_seqs, Hq), 128, 0, s>>>(o_part, ml_part, (T*)out, Hq, splits, out_stride); \
  }
#define LAUNCH_DEC(T, DD)                                                                                           \
  {                                                                                                                  \
    static bool attr_done = false;                                                                                   \
    if (!attr_done) {                                                                                                \
      cudaFuncSetAttribute(paged_decode_kernel<T, T, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_smem_max);\
      attr_done = true;                                                                                              \
    }                                                                                                                \
  }                                                                                                                  \
  paged_decode_kernel<T, T, DD><<<grid, warps * 32, dec_smem, s>>>((const T*)q, (const T*)k_cache, (const T*)v_cache, \
      block_tables, seq_lens, o_part, ml_part, Hq, Hkv, block_size, max_blocks_per_seq, splits, part_len, scale,     \
      alibi_slopes, q_stride, window);                                                                               \
  decode_reduce_kernel<T, DD><<<dim3(num_seqs, Hq), 128, 0, s>>>(o_part, ml_part, (T*)out, Hq, splits, out_stride)
  CB_DISPATCH_HALF(dtype, T, {
    if (use_mma && D == 128) { LAUNCH_DEC_MMA(T, 128); }
    else if (use_mma && D == 64) { LAUNCH_DEC_MMA(T, 64); }
    else if (D == 128) { LAUNCH_DEC(T, 128); }
    else if (D == 64) { LAUNCH_DEC(T, 64); }
    else if (D == 256) { LAUNCH_DEC(T, 256); }
    else return (int)cudaErrorInvalidValue;
  });
#undef LAUNCH_DEC
#undef LAUNCH_DEC_MMA
  return CB_LAUNCH_CHECK();
}

int cb_gather_cos_sin(const float* cos_c, const float* sin_c, const int* pos, float* cos_o, float* sin_o, int tokens,
                      int half, cudaStream_t s) {
  if (tokens == 0) return 0;
  gather_cos_sin_kernel<<<tokens, 128, 0, s>>>(cos_c, sin_c, pos, cos_o, sin_o, tokens, half);
  return CB_LAUNCH_CHECK();
}

// direction 0: (fp16|bf16|fp32) -> e5m2 ; 1: e5m2 -> (fp16|bf16|fp32)
int cb_convert_fp8(const void* in, void* out, int64_t n, int dtype, int direction, cudaStream_t s) {
  if (n == 0) return 0;
  const int grid = (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
  if (direction == 0) {
    if (dtype == CB_BF16) convert_kernel<__nv_bfloat16, __nv_fp8_e5m2><<<grid, 256, 0, s>>>((const __nv_bfloat16*)in, (__nv_fp8_e5m2*)out, n);
    else if (dtype == CB_F16) convert_kernel<__half, __nv_fp8_e5m2><<<grid, 256, 0, s>>>((const __half*)in, (__nv_fp8_e5m2*)out, n);
    else convert_kernel<float, __nv_fp8_e5m2><<<grid, 256, 0, s>>>((const float*)in, (__nv_fp8_e5m2*)out, n);
  } else {
    if (dtype == CB_BF16) convert_kernel<__nv_fp8_e5m2, __nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_fp8_e5m2*)in, (__nv_bfloat16*)out, n);
    else if (dtype == CB_F16) convert_kernel<__nv_fp8_e5m2, __half><<<grid, 256, 0, s>>>((const __nv_fp8_e5m2*)in, (__half*)out, n);
    else convert_kernel<__nv_fp8_e5m2, float><<<grid, 256, 0, s>>>((const __nv_fp8_e5m2*)in, (float*)out, n);
  }
  return CB_LAUNCH_CHECK();
}

}  // extern "C"
