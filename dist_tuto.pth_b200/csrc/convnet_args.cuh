// Shared between the fused ConvNet kernels (convnet.cu: one CTA per sample; convnet_cluster.cu: one cluster per sample).
#pragma once
#include <cuda_runtime.h>

#include <cstring>

#include "sgd_device.cuh"

namespace cn {

constexpr int W1 = 0, B1 = 252, W2 = 264, B2 = 5264, W3 = 5284, B3 = 21284, W4 = 21336, B4 = 21836;
constexpr int NPAR = 21848;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct Args {
  const float* params;      // flat fp32 [NPAR]
  float* grads;             // flat fp32 [NPAR], accumulated with red.add (caller keeps it zeroed)
  const void* x;            // [B,1,28,28] fp32 (normalised) or uint8 (raw, normalised here)
  const long long* target;  // [B]
  float* loss_acc;          // [0] += sum_b nll_b * inv_bsz ; [1] += #correct   (may be null)
  float* out_logp;          // [B,10] log-probabilities (may be null)
  float* mask_out;          // [B,70] dropout scales actually applied (debug/tests; may be null)
  const unsigned long long* step;   // device step counter (RNG offset); may be null -> 0
  unsigned long long seed;
  long long sample_base;    // global index of sample 0 (rank * bsz): decorrelates ranks
  int B;
  int x_u8;
  int training;             // dropout on/off
  int backward;             // compute gradients
  float inv_bsz;            // 1 / local batch (nll_loss mean)
  float p_drop;
  float mean, inv_std;      // uint8 normalisation
  long long grad_stride;    // elements between the two gradient buckets (0: single bucket); bucket = step & 1
  const float* aux;         // optional: conv2.weight pre-arranged by the SGD kernel as [w2f 5000 | w2b 8000] (see sgd.cu)
  b2::FusedTail tail;       // enabled: gradient exchange + SGD run in the tail of THIS kernel (sgd_device.cuh)
  float* det_partials;      // deterministic mode: CTA i stores its gradient sums to det_partials + i * DET_STRIDE (plain
                            // stores) instead of red.add-ing into the bucket; det_reduce_kernel (sgd.cu) sums the slots in order
  const unsigned int* in_flag;   // optional "this batch has landed" word: the executor's copy stream writes in_gen there with a
  unsigned int in_gen;           // stream memory op right behind the H2D copy of x / target; the kernel polls it instead of the
                                 // compute stream waiting on an event, so consecutive steps stay one unbroken PDL kernel chain
};

// Blocks the calling thread until the batch the kernel is about to read has landed (cyclic compare: generations wrap).
__device__ __forceinline__ void wait_input(const Args& a) {
  if (a.in_flag == nullptr) return;
  unsigned int v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.in_flag) : "memory");
  } while ((int)(v - a.in_gen) < 0);
}
constexpr int DET_STRIDE = 21888;   // = NPAR_ALLOC of ops/convnet_fused.py

// Host-side description of the fused tail (C ABI of the launchers); nullptr / enabled == 0 -> two-kernel step.
struct FusedTailHost {
  void* grad_ptrs[8];
  void* inbox_ptrs[8];
  float* params;
  float* momentum;
  unsigned long long* step;
  float* aux;
  const float* loss_acc;
  float* loss_snapshot;
  unsigned int* ticket;
  float lr, mu, scale;
  int rank, world;
  int wire_bf16;
};

inline void fill_tail(b2::FusedTail& t, const FusedTailHost* h, long long grad_stride) {
  memset(&t, 0, sizeof(t));
  if (h == nullptr) return;
  t.enabled = 1;
  for (int i = 0; i < 8; ++i) { t.sgd.grads.p[i] = h->grad_ptrs[i]; t.sgd.inbox.p[i] = h->inbox_ptrs[i]; }
  t.sgd.params = h->params; t.sgd.momentum = h->momentum; t.sgd.step = h->step; t.sgd.done_counter = nullptr;
  t.sgd.n_vec = NPAR / 4; t.sgd.lr = h->lr; t.sgd.mu = h->mu; t.sgd.scale = h->scale; t.sgd.rank = h->rank; t.sgd.world = h->world;
  t.sgd.zero_grads = 1; t.sgd.grad_stride = grad_stride; t.sgd.aux = h->aux;
  t.sgd.wire_bf16 = h->wire_bf16;
  t.sgd.loss_acc = h->loss_acc; t.sgd.loss_snapshot = h->loss_acc ? h->loss_snapshot : nullptr;
  t.ticket = h->ticket;
}

constexpr int AUX_W2F = 0, AUX_W2B = 5000, AUX_TOTAL = 13000;

// relu(pool(conv1)) lives in shared memory as [10 planes][12 rows][12 cols] with row stride 13 and plane stride 161: with these
// strides the 32 (input channel, kernel row) work items of a warp in the conv2 weight-gradient phase hit 32 distinct banks
// (dense 12/144 strides gave 4.8-way conflicts there -- 41 % of all excess shared-memory wavefronts of the kernel).
constexpr int P1_ROW = 13, P1_PLANE = 161, P1_SIZE = (10 * P1_PLANE + 3) / 4 * 4;   // keeps the next shared array 16-byte aligned
// The zero-padded conv2-output gradient [20][16][16] uses row stride 20 / plane stride 324 and is read as float2 pairs:
// 9720 -> 3240 shared-memory wavefronts per sample in the conv2 data-gradient phase (dense 16/256 strides, scalar loads).
constexpr int DC_ROW = 20, DC_PLANE = 324, DC_SIZE = 20 * DC_PLANE;
__host__ __device__ constexpr int p1_idx(int c, int y, int x) { return c * P1_PLANE + y * P1_ROW + x; }
__host__ __device__ constexpr int p1_of(int o) { return (o / 144) * P1_PLANE + ((o % 144) / 12) * P1_ROW + (o % 12); }   // o = c*144 + y*12 + x

}  // namespace cn
