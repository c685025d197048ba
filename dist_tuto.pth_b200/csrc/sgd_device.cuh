// Device-side pieces of the fused [gradient exchange + 1/world + momentum SGD + re-zero] step, shared by the stand-alone
// kernels (sgd.cu) and by the tail of the fused step kernels (convnet.cu / convnet_cluster.cu, "one kernel per step").
#pragma once
#include "common.cuh"

namespace b2 {

constexpr int kSgdThreads = 512;

struct SgdArgs {
  PeerPtrs grads;            // symmetric flat fp32 gradient buckets (grads.p[rank] is ours)
  SignalPads sig;
  float* params;
  float* momentum;
  unsigned long long* step;  // incremented once per call (may be null)
  unsigned int* done_counter; // block-completion counter (device scratch, zero between calls)
  size_t n_vec;              // float4 vectors
  float lr, mu, scale;
  int rank, world;
  int zero_grads;
  long long grad_stride;     // > 0: two buckets, this step's bucket = step & 1; the OTHER bucket is re-zeroed here
  float* aux;                // optional [w2f 5000 | w2b 8000]: conv2.weight re-arranged for the forward/backward kernels
  PeerPtrs inbox;            // push variant only: every rank's inbox  [2 parities][world sources][n_vec][2 lines of 16 B]
  int wire_bf16;             // push variant: gradients cross NVLink as bf16 (ONE 16-byte line {2 x bf16x2 + 2 flags} per float4 vector
                             // instead of two), fp32 accumulation and fp32 master weights; every rank -- the sender included --
                             // sums the same rounded values, so replicas stay bit-identical
  const float* loss_acc;     // optional: the step kernels' running [sum of batch-mean nll, #correct] ...
  float* loss_snapshot;      // ... copied here (2 floats) = the cumulative loss as of THIS step (per-step D2H source)
  unsigned int* snap_flag;   // optional: set to snap_gen (release, system scope) once the snapshot is written -- the executor's
  unsigned int snap_gen;     // D2H stream waits on this word (stream memory op) instead of an event behind the kernel
};

// Call after a __syncthreads() that follows snapshot_loss(): publishes "snapshot of this step is readable".
__device__ __forceinline__ void publish_snapshot(const SgdArgs& a) {
  if (a.snap_flag != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.snap_flag), "r"(a.snap_gen) : "memory");
  }
}

// The previous kernel of the stream (this step's forward/backward) is complete and the next step's kernel cannot pass its
// own griddepcontrol.wait before this kernel ends, so loss_acc holds exactly the loss up to and including this step.
__device__ __forceinline__ void snapshot_loss(const SgdArgs& a) {
  if (a.loss_snapshot != nullptr && blockIdx.x == 0 && threadIdx.x < 2)
    a.loss_snapshot[threadIdx.x] = *reinterpret_cast<const volatile float*>(a.loss_acc + threadIdx.x);
}

// SGD update of one float4 vector (+ the pre-arranged conv2.weight copies), shared by both exchange variants
__device__ __forceinline__ void sgd_apply(const SgdArgs& a, size_t v, float4 g) {
  g.x *= a.scale; g.y *= a.scale; g.z *= a.scale; g.w *= a.scale;
  float4 m = reinterpret_cast<float4*>(a.momentum)[v];
  float4 p = reinterpret_cast<float4*>(a.params)[v];
  m.x = fmaf(a.mu, m.x, g.x); m.y = fmaf(a.mu, m.y, g.y); m.z = fmaf(a.mu, m.z, g.z); m.w = fmaf(a.mu, m.w, g.w);
  p.x = fmaf(-a.lr, m.x, p.x); p.y = fmaf(-a.lr, m.y, p.y); p.z = fmaf(-a.lr, m.z, p.z); p.w = fmaf(-a.lr, m.w, p.w);
  reinterpret_cast<float4*>(a.momentum)[v] = m;
  reinterpret_cast<float4*>(a.params)[v] = p;
  if (a.aux != nullptr && v >= 264 / 4 && v < (264 + 5000) / 4) {      // conv2.weight (flat offset 264, 5000 elements)
    const float pw[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = (int)v * 4 + e - 264;
      const int co = i / 250, r = i - co * 250, ci = r / 25, kk = r - ci * 25;
      a.aux[(ci * 25 + kk) * 20 + co] = pw[e];                                     // w2f [ci][ky][kx][co]
      a.aux[5000 + ((co * 25 + kk) * 2 + ci / 5) * 8 + ci % 5] = pw[e];            // w2b [co][ky][kx][half][8]
    }
  }
}

// One float4 vector `v` of the flat bucket through the push ("LL") exchange and the optimizer:
//   store my value, flag-in-data, into every peer's inbox (16-byte lines {v0, epoch, v1, epoch}); sum the world lines of this
//   vector out of MY inbox in fixed rank order (own contribution from registers) => bit-identical replicas; SGD; re-zero the
//   other-parity bucket.  world == 1: no exchange.  `st` = step index (epoch = st + 1, parity = st & 1).
__device__ __forceinline__ void exchange_apply_vec(const SgdArgs& a, size_t v, unsigned long long st) {
  const int rank = a.rank, world = a.world;
  const uint32_t epoch = (uint32_t)(st + 1ull);
  const size_t ipar = (size_t)(st & 1ull);                              // inbox lines are double-buffered by step parity
  const size_t par = a.grad_stride > 0 ? ipar : 0;                       // ... and so are the gradient buckets when there are two
  const size_t cur_off = par * (size_t)a.grad_stride * sizeof(float);
  const uint4 mine = ld_cg_v4(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(a.grads.p[rank]) + cur_off) + v);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (world > 1 && a.wire_bf16) {
    // one line per vector: {bf16(x),bf16(y) | flag | bf16(z),bf16(w) | flag}
    const uint32_t lo = pack_bf16x2(__uint_as_float(mine.x), __uint_as_float(mine.y));
    const uint32_t hi = pack_bf16x2(__uint_as_float(mine.z), __uint_as_float(mine.w));
    const uint4 l0 = make_uint4(lo, epoch, hi, epoch);
    const size_t dst_line = (ipar * (size_t)world + (size_t)rank) * a.n_vec + v;
#pragma unroll
    for (int i = 1; i < B2_MAX_RANKS; ++i) {
      if (i < world) {
        int r = rank + i;
        if (r >= world) r -= world;
        st_volatile_v4(reinterpret_cast<uint4*>(a.inbox.p[r]) + dst_line, l0);
      }
    }
    const uint4* in = reinterpret_cast<const uint4*>(a.inbox.p[rank]);
#pragma unroll
    for (int r = 0; r < B2_MAX_RANKS; ++r) {
      if (r < world) {
        uint4 q0;
        if (r == rank) {
          q0 = l0;
        } else {
          const uint4* src = in + (ipar * (size_t)world + (size_t)r) * a.n_vec + v;
          unsigned long long spins = 0;
          for (;;) {
            q0 = ld_volatile_v4(src);
            if (q0.y == epoch && q0.w == epoch) break;
            if (++spins > B2_SPIN_LIMIT) {
              printf("[b200dist] push all-reduce (bf16 wire): rank %d timed out waiting for rank %d (step %llu, vector %llu)\n", rank, r,
                     st, (unsigned long long)v);
              __trap();
            }
          }
        }
        g.x += bf16lo(q0.x); g.y += bf16hi(q0.x); g.z += bf16lo(q0.z); g.w += bf16hi(q0.z);
      }
    }
  } else if (world > 1) {
    const size_t dst_line = ((ipar * (size_t)world + (size_t)rank) * a.n_vec + v) * 2;   // ((parity * world + source) * n_vec + v) * 2
    const uint4 l0 = make_uint4(mine.x, epoch, mine.y, epoch), l1 = make_uint4(mine.z, epoch, mine.w, epoch);
#pragma unroll
    for (int i = 1; i < B2_MAX_RANKS; ++i) {          // start with the next rank so the ranks do not all hit one peer first
      if (i < world) {
        int r = rank + i;
        if (r >= world) r -= world;
        uint4* dst = reinterpret_cast<uint4*>(a.inbox.p[r]) + dst_line;
        st_volatile_v4(dst, l0);
        st_volatile_v4(dst + 1, l1);
      }
    }
    const uint4* in = reinterpret_cast<const uint4*>(a.inbox.p[rank]);
#pragma unroll
    for (int r = 0; r < B2_MAX_RANKS; ++r) {
      if (r < world) {
        uint4 q0, q1;
        if (r == rank) {
          q0 = l0; q1 = l1;
        } else {
          const uint4* src = in + ((ipar * (size_t)world + (size_t)r) * a.n_vec + v) * 2;
          unsigned long long spins = 0;
          for (;;) {
            q0 = ld_volatile_v4(src);
            q1 = ld_volatile_v4(src + 1);
            if (q0.y == epoch && q0.w == epoch && q1.y == epoch && q1.w == epoch) break;
            if (++spins > B2_SPIN_LIMIT) {
              printf("[b200dist] push all-reduce: rank %d timed out waiting for rank %d (step %llu, vector %llu)\n", rank, r, st,
                     (unsigned long long)v);
              __trap();
            }
          }
        }
        g.x += __uint_as_float(q0.x); g.y += __uint_as_float(q0.z);
        g.z += __uint_as_float(q1.x); g.w += __uint_as_float(q1.z);
      }
    }
  } else {
    g = make_float4(__uint_as_float(mine.x), __uint_as_float(mine.y), __uint_as_float(mine.z), __uint_as_float(mine.w));
  }
  sgd_apply(a, v, g);
  if (a.zero_grads) {
    const size_t z_off = a.grad_stride > 0 ? (par ^ 1) * (size_t)a.grad_stride * sizeof(float) : 0;
    st_cg_v4(reinterpret_cast<uint4*>(reinterpret_cast<char*>(a.grads.p[rank]) + z_off) + v, make_uint4(0u, 0u, 0u, 0u));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// "One kernel per step": the tail of the fused forward/backward kernels (convnet.cu, convnet_cluster.cu).
//
// After a CTA has flushed its gradients into the bucket it checks in on a device counter; once all `n_cta` CTAs of the
// grid have checked in (they are co-resident: <= 148 CTAs, one per SM) the bucket is complete and EVERY CTA takes a
// 1/n_cta share of the vectors through exchange_apply_vec (push to the peers' inboxes, local reduce, SGD, re-zero).
// Compared with the separate allreduce_sgd kernel this removes the kernel boundary (launch + drain + PDL hand-off) from the
// critical path between "last gradient flushed" and "first peer line stored", and spreads the update over all SMs.
// The last CTA to finish resets the counters, snapshots the running loss and publishes step + 1.
struct FusedTail {
  int enabled;
  SgdArgs sgd;
  unsigned int* ticket;      // device scratch: [0] check-ins, [1] finishers (both zero between steps)
};

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}

// All threads of every participating CTA call this after their gradient flush / loss atomics.  `st` = step index read at
// kernel start, `cta` in [0, n_cta).
__device__ __forceinline__ void fused_tail(const FusedTail& t, unsigned long long st, int n_cta, int cta) {
  __syncthreads();                                   // this CTA's red.adds and loss atomics are issued
  if (threadIdx.x == 0) {
    __threadfence();                                 // ... and ordered before the check-in (cumulative over the barrier)
    atomicAdd(t.ticket, 1u);
    unsigned long long spins = 0;
    while (ld_acquire_gpu_u32(t.ticket) < (uint32_t)n_cta) {
      if (++spins > B2_SPIN_LIMIT) {
        printf("[b200dist] fused step: CTA %d timed out waiting for the grid (%u of %d checked in)\n", cta, *t.ticket, n_cta);
        __trap();
      }
    }
  }
  __syncthreads();
  for (size_t v = (size_t)cta * blockDim.x + threadIdx.x; v < t.sgd.n_vec; v += (size_t)n_cta * blockDim.x)
    exchange_apply_vec(t.sgd, v, st);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(t.ticket + 1, 1u) == (uint32_t)(n_cta - 1)) {      // last finisher: every CTA is past the check-in spin
      t.ticket[0] = 0u;
      t.ticket[1] = 0u;
      if (t.sgd.loss_snapshot != nullptr) {
        t.sgd.loss_snapshot[0] = *reinterpret_cast<const volatile float*>(t.sgd.loss_acc);
        t.sgd.loss_snapshot[1] = *reinterpret_cast<const volatile float*>(t.sgd.loss_acc + 1);
      }
      __threadfence();
      *t.sgd.step = st + 1ull;
    }
  }
}

}  // namespace b2
