// Packed variable-length flash attention forward, version 3: ONE 128-row query tile per work item with a DOUBLE-BUFFERED
// score tile in TMEM, so the score GEMM of key block g+1 is issued before the softmax of block g has finished.
//
// Why (measured on the two-tile kernel in attn.cu, profiles/r02_attn_varlen_ncu_full.csv): with a single S buffer per
// tile the chain  P(j) -> P*V(j), Q*K^T(j+1) -> S(j+1) -> softmax -> P(j+1)  is strictly serial per tile, and the second
// tile only fills the gaps: the tensor pipe and the MUFU are each busy 52 % of the time and idle together the rest (their
// sum is the loop period). Here S(g+1) is already in TMEM when the softmax warps finish block g, so the softmax runs
// back to back (it is the pacing stage, MUFU-bound at 1024 cycles per 128 x 128 block) and the tensor pipe follows one
// block behind: P*V(g) as soon as P(g) arrives, then Q*K^T(g+2) into the buffer P(g) just left.
//
//   warps 0-7   softmax, TWO threads per query row: warps 0-3 take keys 0-63 of a block, warps 4-7 keys 64-127 (warp w and
//               w+4 sit on the same SM sub-partition and share its MUFU; one warp per sub-partition alone reaches only half
//               of the MUFU rate — measured: 2090 cycles per block with 4 softmax warps). tcgen05.ld S(g) -> exp2 (lazy
//               rescaling, see attn.cu; the two threads of a row agree on a redo through a 64-thread named barrier) ->
//               P(g) as bf16 pairs over the first 32 columns of the thread's own 64 S columns; the epilogue of item i is
//               DEFERRED until block 0 of item i+1 has been handed to the tensor pipe (O is double buffered too)
//   warp 8      scheduler + TMA producer: device work counter, Q tiles (double buffered), K_j / V_j ring
//   warp 9      MMA issuer (one lane), software-pipelined over the stream of key blocks ACROSS work items
// TMEM (512 columns): S[0] [0,128)  S[1] [128,256)  O[0] [256,256+D)  O[1] [384,384+D).
//
// STATUS: experimental, off by default (BAGEL_ATTN_V3=1 / BAGEL_ATTN_QT=1). Correct — tests/test_gpu_kernels.py runs the attention
// and the adversarial lazy-rescale cases against it in a child process — but 10-20 % slower than the two-tile kernel in every
// variant measured (profiles/r02_attn_v3_ab.txt, r02_elect_sync_ab.txt): with one tile per CTA every K/V block feeds half as many
// MMAs, and a warp's TMEM loads do not overlap its own math (profiles/r02_microbench_tmem_mufu.txt), so the softmax stage does
// not shrink the way the 1024-cycle MUFU bound above suggests. It is the starting point for a CTA-pair version that shares K/V.
//
// Same interface, masking rules and numerics as attn_varlen_kernel (attn.cu); reference seam: flash_attn_varlen_func at
// modeling/bagel/qwen2_navit.py:361-370, 579-588 and modeling/bagel/siglip_navit.py:232-241.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "common.cuh"
#include "host_util.h"
#include "attn3.h"

namespace bagel {

namespace {

constexpr int kThreads3 = 10 * 32;   // 8 softmax warps + TMA warp + MMA warp
constexpr int kBM = 128;            // query rows per work item
constexpr int kBN = 128;            // keys per block

struct Attn3Params {
  __nv_bfloat16* out;
  long long ld_out;
  const int* cu_q;
  const int* cu_k;
  const int* seqused_k;
  int Hq, Hk;
  int causal;
  float scale_log2;
  int qtiles;      // work items per (sample, head): ceil(max_seqlen_q / 128)
  int num_items;
  int* sched;
};

template <int D>
struct Cfg3 {
  static constexpr int kTileBytes = kBM * D * 2;
  static constexpr int kStages = (D == 128) ? 5 : 8;
  static constexpr int kSmemBytes = 2 * kTileBytes + kStages * kTileBytes + 1024 + 320 + 1024 + 64;   // tiles, alignment, barriers, xch, flags
};

__device__ __forceinline__ float2 ffma2_(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2_(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float ex2_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Item3 {
  int b, h, hk;
  int q_beg, Lq, k_beg, Lk;
  int q0;      // first query row of the tile (within the sample)
  int shift;   // bottom-right aligned causal: key kv visible to query qi iff kv <= qi + shift
  int nblk;    // key blocks swept
  bool valid;  // q0 < Lq
};

__device__ __forceinline__ Item3 decode_item3(int item, const Attn3Params& p) {
  Item3 w;
  // q tiles of one (sample, head) are consecutive items (concurrent CTAs share its K/V through the L2); under a causal mask
  // the LAST tiles sweep the most keys and are handed out first
  const int qi = item % p.qtiles;
  const int bh = item / p.qtiles;
  w.h = bh % p.Hq;
  w.b = bh / p.Hq;
  w.hk = w.h / (p.Hq / p.Hk);
  const int qt = p.causal ? (p.qtiles - 1 - qi) : qi;
  w.q_beg = p.cu_q[w.b];
  w.Lq = p.cu_q[w.b + 1] - w.q_beg;
  w.k_beg = p.cu_k[w.b];
  w.Lk = p.seqused_k ? p.seqused_k[w.b] : (p.cu_k[w.b + 1] - w.k_beg);
  w.q0 = qt * kBM;
  w.valid = w.q0 < w.Lq;
  w.shift = w.Lk - w.Lq;
  w.nblk = 0;
  if (w.valid) {
    int kv_end = w.Lk;
    if (p.causal) {
      const int q_hi = min(w.Lq, w.q0 + kBM) - 1;
      kv_end = max(0, min(w.Lk, q_hi + w.shift + 1));
    }
    w.nblk = (kv_end + kBN - 1) / kBN;
  }
  return w;
}

// QT: the Q tile is copied from shared memory into TMEM once per item (tcgen05.cp) and Q K^T runs with its A operand in
// TMEM: an SS-mode 128x128x16 MMA reads 8 KB of shared memory (the SM's whole 128 B/clk), with Q in TMEM only K is read.
// TMEM then holds ONE O buffer: S[0] S[1] O Q[0] Q[1] = 128 + 128 + 128 + 64 + 64 columns.
template <int D, bool QT>
__global__ void __launch_bounds__(kThreads3, 1)
attn3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, const Attn3Params p) {
  using Cfg = Cfg3<D>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kTileBytes = Cfg::kTileBytes;
  constexpr int kAtoms = D / 64;
  constexpr int kAtomBytes = kBM * 128;
  constexpr int kSoftWarps = 8;
  constexpr int kTmaWarp = kSoftWarps, kMmaWarp = kSoftWarps + 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                       // 2 Q buffers
  uint8_t* smem_kv = smem + 2 * kTileBytes;     // ring: load number n (K of block g: 2g, V: 2g+1) lives in slot n % kStages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + kStages * kTileBytes);
  uint64_t* q_full = bars;                  // [2] TMA -> MMA
  uint64_t* q_empty = q_full + 2;           // [2] MMA -> TMA: the item's last QK^T has read the buffer
  uint64_t* kv_full = q_empty + 2;          // [kStages]
  uint64_t* kv_empty = kv_full + kStages;   // [kStages]
  uint64_t* s_full = kv_empty + kStages;    // [2] MMA -> softmax: S(g) ready in buffer g & 1
  uint64_t* p_full = s_full + 2;            // [2] softmax -> MMA: P(g) written over S(g) (and O rescaled if needed)
  uint64_t* pv_done = p_full + 2;           // [2] MMA -> softmax: P*V(g) complete (only waited for on the rescale path)
  uint64_t* o_full = pv_done + 2;           // [2] MMA -> softmax: final O of the item in O buffer it & 1
  uint64_t* o_free = o_full + 2;            // [2] softmax -> MMA: that O buffer has been read out
  uint64_t* sched_full = o_free + 2;        // [2]
  uint64_t* sched_empty = sched_full + 2;   // [2]
  volatile int* sched_item = reinterpret_cast<volatile int*>(sched_empty + 2);   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(sched_item) + 2);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 320);        // [2][128] row max / row sum exchange
  volatile int* pair_flag = reinterpret_cast<volatile int*>(xch + 256);                 // [2 (block parity)][4 quarters][2 halves]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1);
      mbar_init(&q_empty[t], 1);
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], kSoftWarps);
      mbar_init(&pv_done[t], 1);
      mbar_init(&o_full[t], 1);
      mbar_init(&o_free[t], kSoftWarps);
      mbar_init(&sched_full[t], 1);
      mbar_init(&sched_empty[t], 1 + kSoftWarps);   // MMA lane + one lane of every softmax warp
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTmaWarp) {
    // =========================== scheduler + TMA producer ===========================
    if (elect_one_lane()) {
      uint32_t n_load = 0;     // K/V tiles loaded so far (slot = n % kStages, use number = n / kStages)
      uint32_t it = 0;         // items with key blocks so far (Q buffer = it & 1)
      int slot = 0;
      uint32_t sphase = 0;
      int next = atomicAdd(p.sched, 1);
      while (true) {
        const int item = next;
        mbar_wait(&sched_empty[slot], sphase ^ 1);
        sched_item[slot] = item;
        mbar_arrive(&sched_full[slot]);
        if (++slot == 2) { slot = 0; sphase ^= 1; }
        if (item >= p.num_items) break;
        next = atomicAdd(p.sched, 1);
        const Item3 w = decode_item3(item, p);
        if (w.nblk == 0) continue;
        const int qb = it & 1;
        if (it >= 2) mbar_wait(&q_empty[qb], ((it >> 1) - 1) & 1);   // item it-2 no longer reads this buffer
        mbar_expect_tx(&q_full[qb], kTileBytes);
        for (int a = 0; a < kAtoms; ++a)
          tma_load_2d(smem_q + qb * kTileBytes + a * kAtomBytes, &tmQ, &q_full[qb], w.h * D + a * 64, w.q_beg + w.q0,
                      kEvictFirst);
        ++it;
        for (int j = 0; j < w.nblk; ++j) {
          for (int kv = 0; kv < 2; ++kv) {   // K_j then V_j
            const uint32_t st = n_load % kStages, use = n_load / kStages;
            if (use > 0) mbar_wait(&kv_empty[st], (use - 1) & 1);
            mbar_expect_tx(&kv_full[st], kTileBytes);
            const CUtensorMap* tm = kv == 0 ? &tmK : &tmV;
            for (int a = 0; a < kAtoms; ++a)
              tma_load_2d(smem_kv + st * kTileBytes + a * kAtomBytes, tm, &kv_full[st], w.hk * D + a * 64,
                          w.k_beg + j * kBN, kEvictLast);
            ++n_load;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer ===========================
    if (elect_one_lane()) {   // see common.cuh: `lane == 0` would cost ~80 cycles of wrapper code per MMA
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kBM, kBN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kBM, D, 0, 1);
      int slot = 0;
      uint32_t sphase = 0;

      // next work item with at least one key block (or valid == false at the end of the queue)
      auto fetch = [&](Item3& w) -> bool {
        while (true) {
          mbar_wait(&sched_full[slot], sphase);
          const int item = sched_item[slot];
          mbar_arrive(&sched_empty[slot]);
          if (++slot == 2) { slot = 0; sphase ^= 1; }
          if (item >= p.num_items) return false;
          w = decode_item3(item, p);
          if (w.nblk > 0) return true;
        }
      };
      auto kv_wait = [&](uint32_t n) -> uint32_t {    // wait for load number n, return its slot address
        const uint32_t st = n % kStages, use = n / kStages;
        mbar_wait(&kv_full[st], use & 1);
        return smem_u32(smem_kv + st * kTileBytes);
      };
      // QT: first block of an item -> copy its Q tile smem -> TMEM (8 columns = 16 d per 128x256b copy); the copies and the
      // MMAs of this thread execute in issue order, so no barrier is needed in between, and the smem buffer is free afterwards
      auto stage_q = [&](uint32_t qb) {
        if constexpr (QT) {
          const uint32_t tQ = tmem_base + 384 + qb * 64;
#pragma unroll
          for (int a = 0; a < kAtoms; ++a) {
            const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(smem_q + qb * kTileBytes + a * kAtomBytes));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tQ + (a * 4 + k) * 8), "l"(a_desc + 2 * k) : "memory");
          }
          umma_commit(&q_empty[qb]);
        }
      };
      auto issue_qk = [&](uint32_t g, uint32_t qb, uint32_t kaddr) {
        const uint32_t tS = tmem_base + (g & 1) * 128;
#pragma unroll
        for (int a = 0; a < kAtoms; ++a) {
          const uint64_t b_desc = umma_desc_kmajor_sw128(kaddr + a * kAtomBytes);
          if constexpr (QT) {
            const uint32_t tQ = tmem_base + 384 + qb * 64;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ts(tS, tQ + (a * 4 + k) * 8, b_desc + 2 * k, idesc_qk, (a | k) != 0);
          } else {
            const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(smem_q + qb * kTileBytes + a * kAtomBytes));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tS, a_desc + 2 * k, b_desc + 2 * k, idesc_qk, (a | k) != 0);
          }
        }
        umma_commit(&s_full[g & 1]);
        umma_commit(&kv_empty[(2 * g) % kStages]);
      };

      Item3 cur;
      uint32_t g = 0;        // global key-block counter of this CTA
      uint32_t it = 0;       // item counter (items with key blocks)
      bool have = fetch(cur);
      int j = 0;
      if (have) {
        mbar_wait(&q_full[0], 0);
        const uint32_t kaddr = kv_wait(0);
        tc_fence_after();
        stage_q(0);
        issue_qk(0, 0, kaddr);
        if (!QT && cur.nblk == 1) umma_commit(&q_empty[0]);
      }
      while (have) {
        // ---- look one block ahead: Q K^T of block g+1 goes out before P*V of block g ----
        Item3 nxt = cur;
        int nj = j + 1;
        uint32_t nit = it;
        bool have_next = true;
        if (nj == cur.nblk) {
          have_next = fetch(nxt);
          nj = 0;
          nit = it + 1;
        }
        if (have_next) {
          const uint32_t qb = nit & 1;
          if (nj == 0) mbar_wait(&q_full[qb], (nit >> 1) & 1);
          const uint32_t kaddr = kv_wait(2 * (g + 1));
          tc_fence_after();
          // QT: TMEM Q buffer qb was last read by the Q K^T MMAs of item nit-2, all issued before this point (in-order pipe)
          if (nj == 0) stage_q(qb);
          // S buffer (g+1)&1 held P(g-1): P*V(g-1) was issued in the previous round and the pipe executes in issue order
          issue_qk(g + 1, qb, kaddr);
          if (!QT && nj == nxt.nblk - 1) umma_commit(&q_empty[qb]);   // that was the item's last QK^T
        }
        // ---- P*V of block g ----
        const uint32_t vaddr = kv_wait(2 * g + 1);
        const uint32_t ob = QT ? 0u : (it & 1);                 // O buffer; its use number:
        const uint32_t ou = QT ? it : (it >> 1);
        if (j == 0 && ou >= 1) mbar_wait(&o_free[ob], (ou - 1) & 1);   // the previous user's O has been read out
        mbar_wait(&p_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        {
          const uint32_t tS = tmem_base + (g & 1) * 128;
          const uint32_t tO = tmem_base + 256 + ob * 128;
#pragma unroll
          for (int k = 0; k < kBN / 16; ++k) {
            const uint64_t b_desc = umma_desc_mnmajor_sw128(vaddr + k * 2048, kAtomBytes);
            // P: keys 0-63 over S columns [0,32) (softmax warps 0-3), keys 64-127 over S columns [64,96) (warps 4-7)
            umma_ts(tO, tS + (k < 4 ? k * 8 : 64 + (k - 4) * 8), b_desc, idesc_pv, (j > 0 || k != 0) ? 1u : 0u);
          }
        }
        umma_commit(&pv_done[g & 1]);
        umma_commit(&kv_empty[(2 * g + 1) % kStages]);
        if (j == cur.nblk - 1) umma_commit(&o_full[ob]);
        // ---- advance ----
        ++g;
        cur = nxt;
        j = nj;
        it = nit;
        have = have_next;
      }
    }
  } else {
    // =========================== softmax / correction / epilogue ===========================
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int half = warp >> 2;             // 0: keys 0-63 of a block, 1: keys 64-127
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    constexpr int NC = kBN / 2;             // score columns per thread
    constexpr int DH = D / 2;               // O columns per thread (rescale / epilogue)
    uint32_t g = 0;      // global key-block counter (same sequence as the MMA lane's)
    uint32_t it = 0;     // item counter (items with key blocks)
    int slot = 0;
    uint32_t sphase = 0;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory"); };   // warps w and w + 4

    // deferred epilogue of the previous item
    bool pend = false;
    float pend_inv_l = 0.f;
    bool pend_row_ok = false;
    __nv_bfloat16* pend_row = nullptr;
    uint32_t pend_it = 0;

    auto epilogue = [&]() {
      const uint32_t ob = QT ? 0u : (pend_it & 1);
      mbar_wait(&o_full[ob], (QT ? pend_it : (pend_it >> 1)) & 1);
      tc_fence_after();
      const uint32_t tO = tmem_base + 256 + ob * 128 + half * DH + lane_off;
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tO + c * 32, v);
        tmem_ld_wait();
        if (pend_row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(pend_row + half * DH + c * 32);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = pack_bf16x2(__uint_as_float(v[q4 * 8 + 2 * e]) * pend_inv_l, __uint_as_float(v[q4 * 8 + 2 * e + 1]) * pend_inv_l);
            dst[q4] = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[ob]);
      pend = false;
    };

    while (true) {
      mbar_wait(&sched_full[slot], sphase);
      const int item = sched_item[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sched_empty[slot]);
      if (++slot == 2) { slot = 0; sphase ^= 1; }
      if (item >= p.num_items) break;
      const Item3 w = decode_item3(item, p);
      if (!w.valid) continue;
      const int nblk = w.nblk;
      const int Lq = w.Lq, Lk = w.Lk, shift = w.shift;
      const int qi = w.q0 + row;
      __nv_bfloat16* orow = p.out + (long long)(w.q_beg + qi) * p.ld_out + w.h * D;
      if (nblk == 0) {   // rows that see no key at all (causal, Lq > Lk): zeros, as flash-attn does; no barrier involved
        if (qi < Lq) {
#pragma unroll
          for (int c = 0; c < DH / 8; ++c) reinterpret_cast<uint4*>(orow + half * DH)[c] = make_uint4(0u, 0u, 0u, 0u);
        }
        continue;
      }

      // m: reference maximum of the row, identical in the two threads of a row by construction; l: this thread's share of
      // the row sum (its 64 keys of every block)
      float m = -INFINITY, l = 0.f;
      constexpr float kRedoSum = 1073741824.0f;   // 2^30, see attn.cu
      for (int j = 0; j < nblk; ++j, ++g) {
        const uint32_t tS = tmem_base + (g & 1) * 128 + half * NC + lane_off;   // this thread's 64 score columns; P goes to the first 32
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        const int kv0 = j * kBN + half * NC;       // first key of this thread's columns
        const bool need_mask = (kv0 + NC > Lk) || (p.causal && (kv0 + NC - 1 > w.q0 + shift));
        const int lim = p.causal ? min(Lk - 1, qi + shift) : (Lk - 1);

        uint32_t pk[NC / 2];
        float2 rs2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);

        auto process_t = [&](auto mask_tag, const uint32_t (&v)[32], int c, float neg_ms) {
          constexpr bool kMask = decltype(mask_tag)::value;
          const float2 nm2 = make_float2(neg_ms, neg_ms);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x0 = __uint_as_float(v[2 * i]), x1 = __uint_as_float(v[2 * i + 1]);
            if constexpr (kMask) {
              if (kv0 + c * 32 + 2 * i > lim) x0 = -INFINITY;
              if (kv0 + c * 32 + 2 * i + 1 > lim) x1 = -INFINITY;
            }
            const float2 x = ffma2_(make_float2(x0, x1), sc2, nm2);
            const float2 e = make_float2(ex2_(x.x), ex2_(x.y));
            rs2[i & 1] = fadd2_(rs2[i & 1], e);
            pk[c * 16 + i] = pack_bf16x2(e.x, e.y);
          }
        };

        // m is the same in both threads of a row, so this vote has the same outcome in warp w and warp w + 4
        const bool have_ref = __all_sync(0xffffffffu, m != -INFINITY);
        bool redo = true;
        if (have_ref) {
          const float neg_ms = -m * p.scale_log2;
          auto stream = [&](auto mask_tag) {
            uint32_t va[32], vb[32];
            tmem_ld_x32(tS + 0, va);
            tmem_ld_x32(tS + 32, vb);
            tmem_ld_wait();
            process_t(mask_tag, va, 0, neg_ms);
            process_t(mask_tag, vb, 1, neg_ms);
          };
          if (need_mask) stream(std::true_type{});
          else stream(std::false_type{});
          const float rs_row = (rs2[0].x + rs2[1].x) + (rs2[0].y + rs2[1].y);
          redo = !(rs_row <= kRedoSum);
        }
        // the two warps that share these rows must take the same path: exchange the warp votes
        const bool warp_redo = __any_sync(0xffffffffu, redo);
        volatile int* flags = pair_flag + (g & 1) * 8 + quarter * 2;
        if (lane == 0) flags[half] = warp_redo ? 1 : 0;
        pair_sync();
        const bool pair_redo = (flags[0] | flags[1]) != 0;
        float alpha = 1.0f;
        if (pair_redo) {
          // slow path (first block of an item, or a reference that has become badly stale in either half): exact row maximum
          // over all 128 keys, move m, recompute this thread's P, rescale its share of O and l
          float mx = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < NC / 32; ++c) {
            uint32_t v[32];
            tmem_ld_x32(tS + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float x = __uint_as_float(v[i]);
              if (need_mask && (kv0 + c * 32 + i > lim)) x = -INFINITY;
              mx = fmaxf(mx, x);
            }
          }
          xch[half * 128 + row] = mx;
          pair_sync();
          mx = fmaxf(mx, xch[(half ^ 1) * 128 + row]);
          const float m_new = fmaxf(m, mx);
          const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
          alpha = ex2_((m - m_use) * p.scale_log2);   // m = -inf -> 0
          m = m_new;
          const float neg_ms = -m_use * p.scale_log2;
          rs2[0] = make_float2(0.f, 0.f);
          rs2[1] = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < NC / 32; ++c) {
            uint32_t v[32];
            tmem_ld_x32(tS + c * 32, v);
            tmem_ld_wait();
            if (need_mask) process_t(std::true_type{}, v, c, neg_ms);
            else process_t(std::false_type{}, v, c, neg_ms);
          }
          if (j > 0) {
            // O holds blocks 0..j-1 only once P*V(g-1) has COMPLETED; S(g) being ready does not imply that here (Q K^T of
            // block g is issued before P*V of block g-1), hence the explicit barrier
            mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
            tc_fence_after();
            const uint32_t tO = tmem_base + 256 + (QT ? 0u : (it & 1)) * 128 + half * DH + lane_off;
#pragma unroll
            for (int c = 0; c < DH / 32; ++c) {
              uint32_t v[32];
              tmem_ld_x32(tO + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st_x32(tO + c * 32, v);
            }
          }
        }
        const float rs = (rs2[0].x + rs2[1].x) + (rs2[0].y + rs2[1].y);
        l = l * alpha + rs;
        tmem_st_x32(tS, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g & 1]);

        if (j == 0 && pend) epilogue();   // the previous item's O: its last P*V has long finished by now
      }

      // row sum = both halves, exchanged through shared memory. First barrier: the partner is past its read of the row
      // maximum this thread may have left in the same slot during the last block's redo.
      pair_sync();
      xch[half * 128 + row] = l;
      pair_sync();
      l += xch[(half ^ 1) * 128 + row];
      pair_sync();   // the partner has read this thread's value before the next item's first block overwrites it
      pend = true;
      pend_inv_l = (l > 0.f && m != -INFINITY) ? (1.f / l) : 0.f;
      pend_row_ok = qi < Lq;
      pend_row = orow;
      pend_it = it;
      ++it;
    }
    if (pend) epilogue();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int* sched_counter3(cudaStream_t stream) {
  constexpr int kRing = 64;
  static int* base[64] = {nullptr};
  static std::atomic<unsigned> next{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  if (base[dev] == nullptr) {
    int* ptr = nullptr;
    if (cudaMalloc(&ptr, kRing * sizeof(int)) != cudaSuccess) return nullptr;
    base[dev] = ptr;
  }
  int* c = base[dev] + (next.fetch_add(1, std::memory_order_relaxed) % kRing);
  if (cudaMemsetAsync(c, 0, sizeof(int), stream) != cudaSuccess) return nullptr;
  return c;
}

template <int D, bool QT>
int launch3(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, Attn3Params p, int B, int max_seqlen_q,
            cudaStream_t stream) {
  using Cfg = Cfg3<D>;
  auto kern = attn3_kernel<D, QT>;
  static bool attr_done = false;
  if (!attr_done) {
    BAGEL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  p.qtiles = (max_seqlen_q + kBM - 1) / kBM;
  const long long items = (long long)p.qtiles * p.Hq * B;
  if (items > 0x7fffffff - 4096) return set_error(BAGEL_ERR_SHAPE, "bagel_attn_varlen_fwd: too many work items");
  p.num_items = (int)items;
  p.sched = sched_counter3(stream);
  if (p.sched == nullptr) return set_error(BAGEL_ERR_CUDA, "bagel_attn_varlen_fwd: scheduler counter allocation failed");
  const int grid = p.num_items < sm_count() ? p.num_items : sm_count();
  kern<<<grid, kThreads3, Cfg::kSmemBytes, stream>>>(tmQ, tmK, tmV, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

bool attn3_enabled() {
  static const bool on = [] { const char* e = getenv("BAGEL_ATTN_V3"); return e && atoi(e) != 0; }();
  return on;
}

int attn3_varlen(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, void* out, long long ld_out,
                 const int* cu_q, const int* cu_k, const int* seqused_k, int batch, int Hq, int Hk, int head_dim,
                 int max_seqlen_q, int causal, float scale_log2, cudaStream_t stream) {
  Attn3Params p{};
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ld_out = ld_out;
  p.cu_q = cu_q;
  p.cu_k = cu_k;
  p.seqused_k = seqused_k;
  p.Hq = Hq;
  p.Hk = Hk;
  p.causal = causal;
  p.scale_log2 = scale_log2;
  static const bool qt = [] { const char* e = getenv("BAGEL_ATTN_QT"); return e && atoi(e) != 0; }();
  if (head_dim == 128)
    return qt ? launch3<128, true>(tmQ, tmK, tmV, p, batch, max_seqlen_q, stream)
              : launch3<128, false>(tmQ, tmK, tmV, p, batch, max_seqlen_q, stream);
  return qt ? launch3<64, true>(tmQ, tmK, tmV, p, batch, max_seqlen_q, stream)
            : launch3<64, false>(tmQ, tmK, tmV, p, batch, max_seqlen_q, stream);
}

}  // namespace bagel
