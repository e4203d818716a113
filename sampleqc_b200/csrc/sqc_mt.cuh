// sqc_mt.cuh — R's Mersenne-Twister key stream (SQC_RNG_R_COMPAT) generated block-parallel.
//
// One update_beta_scale_k call consumes N_total draws of R's stream.  MT19937 is a
// sequential recurrence, but its state transition is linear over GF(2): the window 624 words further
// down the stream by e words is  x[e + j] = XOR_{i in g_{e-1}} x[i + j + 1]  with g_e = x^e mod phi
// (scripts/mt_jump_tables.py derives and verifies the polynomials; sqc_mt_jump.inc holds g_{pB-1}).
//   k_mt_seq     one CTA: the 33 chunks (20 592 words) that follow the current state
//   k_mt_jump    CTA p: correlate that sequence with row p*mult-1 of the table -> the window at block p
//   k_mt_gen     CTA p: run the recurrence from that window for up to mult*110 chunks, temper, convert to
//                randperm keys (key = (int) Rf_runif(0, RAND_MAX)) and store them at their draw position.
// State layout (device, 625 words) = R's .Random.seed payload: mt[624] then mti.
#pragma once
#include "sqc_kernels.cuh"
#include "sqc_mt_jump.inc"

namespace sqc {

constexpr int MT_N = 624;
constexpr int MT_SEQ_CHUNKS = 33;                         // 33 * 624 = 20 592 >= 19 937 + 624 + 1
constexpr int MT_SEQ_WORDS = MT_SEQ_CHUNKS * MT_N;
constexpr int MT_BLOCK_CHUNKS = SQC_MT_CHUNKS_PER_BLOCK;  // 110
constexpr long long MT_BLOCK_WORDS = (long long)MT_BLOCK_CHUNKS * MT_N;

// one regeneration: nxt[0..623] from cur[0..623]  (dependency distance 227 -> three wavefronts)
__device__ __forceinline__ void mt_regen(const uint32_t* cur, uint32_t* nxt, int tid) {
  if (tid < 227) nxt[tid] = cur[tid + 397] ^ mt_twist(cur[tid], cur[tid + 1]);
  __syncthreads();
  if (tid < 227) nxt[227 + tid] = nxt[tid] ^ mt_twist(cur[227 + tid], cur[228 + tid]);
  __syncthreads();
  if (tid < 169) nxt[454 + tid] = nxt[227 + tid] ^ mt_twist(cur[454 + tid], cur[455 + tid]);
  if (tid == 169) nxt[623] = nxt[396] ^ mt_twist(cur[623], nxt[0]);
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_mt_seq(const uint32_t* st, uint32_t* seq, const Control* ctl) {
  __shared__ uint32_t bufA[MT_N], bufB[MT_N];
  if (ctl && !ctl->cont) return;
  const int tid = threadIdx.x;
  uint32_t* cur = bufA; uint32_t* nxt = bufB;
  for (int i = tid; i < MT_N; i += 256) { cur[i] = st[i]; seq[i] = st[i]; }
  __syncthreads();
  for (int c = 1; c < MT_SEQ_CHUNKS; ++c) {
    mt_regen(cur, nxt, tid);
    for (int i = tid; i < MT_N; i += 256) seq[c * MT_N + i] = nxt[i];
    uint32_t* t = cur; cur = nxt; nxt = t;
  }
}

// window of block p = 1 .. nb-1: row p * mult - 1 of the jump table correlated with the 33-chunk sequence,
//     window[j] = XOR_{i : bit i of the row} seq[i + j + 1],   j = 0 .. 623.
// Register-tiled: a thread owns 8 consecutive outputs and walks the taps 32 at a time; the 39 sequence words a
// table word can touch are fetched with ten 128-bit shared-memory loads and reused for all (tap, output) pairs,
// so shared-memory traffic is ~1 byte per pair instead of 4.  The taps are split over MT_JG groups of 3 warps
// (the tap bits are warp-uniform: no divergence); the partial windows are XOR-ed through shared memory.
constexpr int MT_JT = 96;                                 // threads per tap group (78 own outputs, 3 warps)
#ifndef SQC_MT_JG
#define SQC_MT_JG 6
#endif
constexpr int MT_JG = SQC_MT_JG;                          // tap groups (624 / MT_JG table words each)
constexpr int MT_JUMP_THREADS = MT_JT * MT_JG;
constexpr int MT_SEQ_PAD = MT_SEQ_WORDS + 64;             // the last tiles read a few words past the sequence
constexpr size_t MT_SMEM = (size_t)(MT_SEQ_PAD + MT_N + MT_JG * MT_N) * sizeof(uint32_t);
// window = row (624 words of a jump polynomial) correlated with the 33-chunk sequence `seq`; block-wide
__device__ __forceinline__ void mt_correlate(const uint32_t* seq, const uint32_t* row, uint32_t* window, uint32_t* sh_mt) {
  uint32_t* sseq = sh_mt;                         // sseq[m] = seq[m + 1]
  uint32_t* stab = sseq + MT_SEQ_PAD;             // 624 table words
  uint32_t* spart = stab + MT_N;                  // MT_JG x 624 partial windows
  const int tid = threadIdx.x;
  for (int i = tid; i < MT_SEQ_PAD; i += MT_JUMP_THREADS) sseq[i] = (i + 1 < MT_SEQ_WORDS) ? seq[i + 1] : 0u;
  for (int i = tid; i < MT_N; i += MT_JUMP_THREADS) stab[i] = row[i];
  __syncthreads();
  const int g = tid / MT_JT, t = tid % MT_JT;
  uint32_t acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0u;
  if (t < MT_N / 8) {
    const int w0 = g * (MT_N / MT_JG);
    for (int w = w0; w < w0 + MT_N / MT_JG; ++w) {
      const uint32_t bits = stab[w];
      if (bits == 0u) continue;
      const uint4* src = reinterpret_cast<const uint4*>(sseq + 32 * w + 8 * t);     // 32-byte aligned
      uint32_t X[40];
#pragma unroll
      for (int v = 0; v < 10; ++v) { const uint4 q = src[v]; X[4 * v] = q.x; X[4 * v + 1] = q.y; X[4 * v + 2] = q.z; X[4 * v + 3] = q.w; }
#pragma unroll
      for (int b = 0; b < 32; ++b) {
        if (bits & (1u << b)) {
#pragma unroll
          for (int o = 0; o < 8; ++o) acc[o] ^= X[b + o];
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) spart[g * MT_N + 8 * t + o] = acc[o];
  }
  __syncthreads();
  for (int j = tid; j < MT_N; j += MT_JUMP_THREADS) {
    uint32_t a = 0u;
#pragma unroll
    for (int gg = 0; gg < MT_JG; ++gg) a ^= spart[gg * MT_N + j];
    window[j] = a;
  }
}
__global__ void __launch_bounds__(MT_JUMP_THREADS) k_mt_jump(const uint32_t* st_in, const uint32_t* seq, const uint32_t* table, uint32_t* windows,
                                                             long long n, int mult, const Control* ctl) {
  extern __shared__ __align__(16) uint32_t sh_mt[];
  if (ctl && !ctl->cont) return;
  const int p = blockIdx.x + 1;
  const long long u_end = (long long)st_in[MT_N] + n;
  if ((long long)p * MT_BLOCK_WORDS * mult >= u_end) return;          // block past the last draw
  mt_correlate(seq, table + (size_t)(p * mult - 1) * MT_N, windows + (size_t)p * MT_N, sh_mt);
}

// ---- distributed key stream (multi-GPU): every rank generates only its share of a call's N_total draws --------
// Stream positions are fixed by (N_total, world): rank r owns draws [r Q, (r + 1) Q) of every call, Q a multiple of
// 624.  The state in front of a call is advanced by whole chunks with jump polynomials derived on the host at
// session creation (mt_xpow below): row 0 = x^(624 floor(N_total / 624) - 1) (call to call), row 1 = x^(r Q - 1)
// (start of the call to the start of this rank's share; same mti).  CTA b correlates row b -> jwin[b].
__global__ void __launch_bounds__(MT_JUMP_THREADS) k_mt_jump_rows(const uint32_t* seq, const uint32_t* rows, uint32_t* jwin, const Control* ctl) {
  extern __shared__ __align__(16) uint32_t sh_mt[];
  if (ctl && !ctl->cont) return;
  mt_correlate(seq, rows + (size_t)blockIdx.x * MT_N, jwin + (size_t)blockIdx.x * MT_N, sh_mt);
}
// R's state after the call's n_total draws (st_out) and the state in front of this rank's share (st_rank).
// (array, mti) stands for the raw position 624 q + mti with mti in [1, 624]: the array is the chunk that holds
// the last word drawn, exactly what R keeps in .Random.seed.
__global__ void __launch_bounds__(256) k_mt_fix(const uint32_t* st_in, const uint32_t* jwin, uint32_t* st_out, uint32_t* st_rank,
                                                long long n_total, int has_rank, const Control* ctl) {
  __shared__ uint32_t bufA[MT_N], bufB[MT_N];
  if (ctl && !ctl->cont) return;
  const int tid = threadIdx.x;
  const long long mti = (long long)st_in[MT_N];
  const long long c0 = n_total / MT_N, rho = n_total - c0 * MT_N;
  const bool extra = (mti + rho - 1) >= MT_N;                 // the call ends one chunk further than the c0-chunk jump
  for (int i = tid; i < MT_N; i += 256) bufA[i] = jwin[i];
  __syncthreads();
  const uint32_t* src = bufA;
  if (extra) { mt_regen(bufA, bufB, tid); src = bufB; }
  for (int i = tid; i < MT_N; i += 256) { st_out[i] = src[i]; st_rank[i] = has_rank ? jwin[MT_N + i] : st_in[i]; }
  if (tid == 0) { st_out[MT_N] = (uint32_t)(mti + rho - (extra ? MT_N : 0)); st_rank[MT_N] = (uint32_t)mti; }
}

// One 256-thread CTA per block runs the recurrence in shared memory (three 227-wide wavefronts per chunk), tempers
// every word of the current chunk, converts it to a randperm key and stores it at its draw position.  Positions u
// (counted from the start of the current state's chunk) in [mti, mti + n) are draws.  `mult`: a block is
// mult * 110 chunks long.  (A single warp per block, tried first, is latency-bound at ~2 us per chunk.)
constexpr int MT_GEN_THREADS = 256;
__global__ void __launch_bounds__(MT_GEN_THREADS) k_mt_gen(const uint32_t* st_in, uint32_t* st_out, const uint32_t* seq, const uint32_t* windows,
                                                           int* out, long long n, int mult, const Control* ctl) {
  __shared__ uint32_t bufA[MT_N], bufB[MT_N];
  if (ctl && !ctl->cont) return;
  const int tid = threadIdx.x, p = blockIdx.x;
  const long long mti = (long long)st_in[MT_N];
  const long long u_end = mti + n, u_last = u_end - 1, q_last = u_last / MT_N;
  const long long blk_chunks = (long long)MT_BLOCK_CHUNKS * mult;
  if ((long long)p * blk_chunks * MT_N >= u_end) return;
  const uint32_t* src = (p == 0) ? seq : windows + (size_t)p * MT_N;
  uint32_t* cur = bufA; uint32_t* nxt = bufB;
  for (int i = tid; i < MT_N; i += MT_GEN_THREADS) cur[i] = src[i];
  __syncthreads();
  for (long long c = 0; c < blk_chunks; ++c) {
    const long long q = (long long)p * blk_chunks + c;
    const long long u0 = q * MT_N;
    if (u0 >= u_end) break;
    const long long o0 = u0 - mti;                // draw index of word 0 of this chunk
    for (int i = tid; i < MT_N; i += MT_GEN_THREADS) {
      const long long o = o0 + i;
      const uint32_t a = cur[i];
      if (o >= 0 && o < n) out[o] = mt_word_to_key(a);
      if (q == q_last) st_out[i] = a;             // R's state after the call: this chunk, mti' = offset past the last draw
    }
    if (q == q_last && tid == 0) st_out[MT_N] = (uint32_t)(u_last - u0 + 1);
    if ((c + 1 < blk_chunks) && (u0 + MT_N < u_end)) {
      mt_regen(cur, nxt, tid);
      uint32_t* t = cur; cur = nxt; nxt = t;
    }
  }
}

// ---- host: x^e mod phi_MT19937 over GF(2) (square-and-shift; phi is sparse: reduction by its 135 terms) --------
namespace mtpoly {
constexpr int DEG = 19937, W64 = 624;            // 624 64-bit words hold a product of two reduced polynomials
inline void xor_at(uint64_t* p, uint64_t h, long long off) {          // p ^= h * x^off (off > -64; dropped low bits are zero)
  if (off < 0) { p[0] ^= h >> (-off); return; }
  const long long w = off >> 6; const int sh = (int)(off & 63);
  p[w] ^= h << sh;
  if (sh) p[w + 1] ^= h >> (64 - sh);
}
inline void reduce(uint64_t* p) {
  const int top = SQC_MT_PHI_TERMS - 1;          // sqc_mt_phi_terms[top] == DEG; the next term is 623 lower (> 64: one pass per word)
  for (int w = W64 - 1; w >= DEG / 64; --w) {
    for (;;) {
      uint64_t h = p[w];
      if (w == DEG / 64) h &= ~((1ull << (DEG % 64)) - 1ull);
      if (!h) break;
      p[w] ^= h;                                  // x^(DEG + k) = sum of the lower terms shifted by k
      for (int i = 0; i < top; ++i) xor_at(p, h, 64LL * w - DEG + sqc_mt_phi_terms[i]);
    }
  }
}
inline uint64_t spread32(uint32_t v) {            // bit i -> bit 2 i
  uint64_t x = v;
  x = (x | (x << 16)) & 0x0000ffff0000ffffull; x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
  x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full; x = (x | (x << 2)) & 0x3333333333333333ull; x = (x | (x << 1)) & 0x5555555555555555ull;
  return x;
}
// 624 32-bit words, bit i of word w = coefficient of x^(32 w + i): the layout of sqc_mt_jump_table's rows
inline void xpow(unsigned long long e, uint32_t* out) {
  std::vector<uint64_t> p(W64 + 1, 0ull), q(W64 + 1, 0ull);
  p[0] = 1ull;
  int msb = 63; while (msb > 0 && !((e >> msb) & 1ull)) --msb;
  for (int b = msb; b >= 0; --b) {
    for (int w = 0; w < W64 / 2; ++w) { q[2 * w] = spread32((uint32_t)p[w]); q[2 * w + 1] = spread32((uint32_t)(p[w] >> 32)); }
    q[W64] = 0ull;
    reduce(q.data());
    p.swap(q);
    if ((e >> b) & 1ull) {
      for (int w = W64 / 2; w > 0; --w) p[w] = (p[w] << 1) | (p[w - 1] >> 63);
      p[0] <<= 1;
      if ((p[DEG / 64] >> (DEG % 64)) & 1ull) for (int i = 0; i < SQC_MT_PHI_TERMS; ++i) p[sqc_mt_phi_terms[i] >> 6] ^= 1ull << (sqc_mt_phi_terms[i] & 63);
    }
  }
  for (int w = 0; w < MT_N / 2; ++w) { out[2 * w] = (uint32_t)p[w]; out[2 * w + 1] = (uint32_t)(p[w] >> 32); }
}
// the derivation is checked once per process against a row computed (and verified against a direct stream) offline
inline bool self_check() {
  static int ok = -1;
  if (ok < 0) {
    std::vector<uint32_t> r(MT_N);
    xpow((unsigned long long)MT_BLOCK_WORDS - 1ull, r.data());
    ok = memcmp(r.data(), sqc_mt_jump_table[0], MT_N * sizeof(uint32_t)) == 0 ? 1 : 0;
    if (ok) { xpow(3ull * MT_BLOCK_WORDS - 1ull, r.data()); ok = memcmp(r.data(), sqc_mt_jump_table[2], MT_N * sizeof(uint32_t)) == 0 ? 1 : 0; }
  }
  return ok == 1;
}
}  // namespace mtpoly

}  // namespace sqc
