// Fused  [peer-memory gradient all-reduce] + [1/world scale] + [momentum SGD]  for flat fp32 buffers.
//
// Replaces average_gradients() + optimizer.step() + optimizer.zero_grad() of the reference training
// step (train_dist.py:118,123,124): 8 all-reduces + 8 divides + foreach-SGD + 8 memsets become ONE
// kernel, in two exchange flavours with identical arithmetic (fixed rank order => bit-identical replicas):
//   * allreduce_sgd_push_kernel (default for world > 1): every rank stores its locally reduced bucket, flag-in-data,
//     into every peer's inbox over NVSwitch and reduces out of its own memory -- one NVLink crossing on the critical path;
//   * allreduce_sgd_kernel: flag barrier, then every rank loads all peers' buckets (one-shot).  Also the world == 1 path.
// Both apply  buf = mu*buf + g ; p -= lr*buf  (torch.optim.SGD semantics with dampening 0, no nesterov, no weight
// decay -- train_dist.py:110), re-zero the gradient bucket of the other step parity for the next `red.add` accumulation,
// keep conv2.weight pre-arranged for the step kernels (`aux`) and bump the device step counter used by the dropout RNG.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "sgd_device.cuh"

namespace b2 {

__global__ void __launch_bounds__(kSgdThreads) allreduce_sgd_kernel(SgdArgs a) {
  const int rank = a.rank, world = a.world;
  // Step parity (which gradient bucket is live) + step bump.  Thread 0 of every block reads the counter and then
  // checks in with an atomic; the LAST block to check in knows every block has read it and bumps it right away, so
  // the atomic's latency hides behind the rest of the kernel and no block can see the new value.
  __shared__ unsigned int s_par;
  pdl_wait();                        // gradients of this step (previous kernel) are complete and visible
  pdl_launch_dependents();           // the next step's forward/backward kernel may pre-launch now (it zeroes its smem, then waits)
  snapshot_loss(a);
  unsigned long long st = 0ull;
  unsigned int seen = 0u;
  if (threadIdx.x == 0) {
    st = a.step != nullptr ? *reinterpret_cast<volatile unsigned long long*>(a.step) : 0ull;
    s_par = (unsigned int)(st & 1ull);
    if (a.step != nullptr) seen = atomicAdd(a.done_counter, 1u);    // result is only consumed at the very end (latency hidden)
  }
  __syncthreads();
  publish_snapshot(a);
  uint32_t epoch = 0;
  if (world > 1) {
    epoch = barrier_epoch_load(a.sig, rank);
    block_barrier_all_ranks(a.sig, rank, world, ++epoch);           // all gradient buckets are complete
  }
  const size_t stride = (size_t)gridDim.x * kSgdThreads;
  // Double-buffered buckets remove the second barrier: once every rank has arrived at THIS step's barrier it has
  // finished reading last step's bucket, so that one can be zeroed right away for the step after this one.
  const bool dbuf = a.grad_stride > 0;
  const size_t par = dbuf ? (size_t)s_par : 0;
  const size_t cur_off = par * (size_t)a.grad_stride * sizeof(float);
  const size_t oth_off = (par ^ 1) * (size_t)a.grad_stride * sizeof(float);
  // block-uniform trip count (barrier inside the loop)
  for (size_t base = (size_t)blockIdx.x * kSgdThreads; base < a.n_vec; base += stride) {
    const size_t v = base + threadIdx.x;
    const bool ok = v < a.n_vec;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) {
      uint4 raw[B2_MAX_RANKS];
#pragma unroll
      for (int r = 0; r < B2_MAX_RANKS; ++r)
        if (r < world) raw[r] = ld_cg_v4(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(a.grads.p[r]) + cur_off) + v);
#pragma unroll
      for (int r = 0; r < B2_MAX_RANKS; ++r)
        if (r < world) {
          g.x += __uint_as_float(raw[r].x); g.y += __uint_as_float(raw[r].y);
          g.z += __uint_as_float(raw[r].z); g.w += __uint_as_float(raw[r].w);
        }
      sgd_apply(a, v, g);
    }
    if (a.zero_grads) {
      if (dbuf) {
        if (ok) st_cg_v4(reinterpret_cast<uint4*>(reinterpret_cast<char*>(a.grads.p[rank]) + oth_off) + v, make_uint4(0u, 0u, 0u, 0u));
      } else {
        if (world > 1) block_barrier_all_ranks(a.sig, rank, world, ++epoch);   // peers are done reading this pass
        if (ok) st_cg_v4(reinterpret_cast<uint4*>(a.grads.p[rank]) + v, make_uint4(0u, 0u, 0u, 0u));
      }
    }
  }
  if (world > 1 && threadIdx.x == 0) barrier_epoch_store(a.sig, rank, epoch);
  // the last block to have checked in knows every block has read the step counter: it publishes step + 1
  if (threadIdx.x == 0 && a.step != nullptr && seen == gridDim.x - 1) { *a.done_counter = 0u; *a.step = st + 1ull; }
}

// Push ("LL") variant of the same step: no flag barrier and no remote loads.
//
// Every rank STORES its locally reduced bucket into every peer's inbox as 16-byte lines {v0, epoch, v1, epoch}
// (the flag travels with the data, so one NVLink crossing both delivers and publishes it), then sums the world lines of
// each element out of its OWN memory, polling until both flags of a line carry this step's epoch.  Compared with
// barrier + peer loads (flag crossing, then a load round trip) the critical path is ONE one-way crossing.  The sum
// runs in fixed rank order with the own contribution taken from registers at position `rank`, so replicas stay
// bit-identical and equal to the barrier variant.  epoch = step + 1 (never 0 = freshly zeroed inbox); lines are
// double-buffered by step parity: a peer can only write parity p again two steps later, which needs my push of the
// step in between, which I issue after I finished reading parity p.
__global__ void __launch_bounds__(kSgdThreads) allreduce_sgd_push_kernel(SgdArgs a) {
  __shared__ unsigned long long s_step;
  pdl_wait();
  pdl_launch_dependents();
  snapshot_loss(a);
  unsigned int seen = 0u;
  if (threadIdx.x == 0) {
    s_step = *reinterpret_cast<volatile unsigned long long*>(a.step);
    seen = atomicAdd(a.done_counter, 1u);
  }
  __syncthreads();
  publish_snapshot(a);
  const unsigned long long st = s_step;
  const size_t stride = (size_t)gridDim.x * kSgdThreads;
  for (size_t v = (size_t)blockIdx.x * kSgdThreads + threadIdx.x; v < a.n_vec; v += stride) exchange_apply_vec(a, v, st);
  if (threadIdx.x == 0 && seen == gridDim.x - 1) { *a.done_counter = 0u; *a.step = st + 1ull; }
}

// Deterministic mode of the fused ConvNet step (convnet_args.cuh: det_partials): every step CTA stored its gradient sums to a
// private slot; this kernel adds the slots IN CTA ORDER into this step's bucket, so the local sum -- and with the fixed rank
// order of the exchange the whole update -- is bit-reproducible from run to run (float red.add into one bucket is not).
__global__ void __launch_bounds__(256) det_reduce_kernel(const float* __restrict__ partials, int n_slots, long long slot_stride,
                                                         float* __restrict__ grads, const unsigned long long* __restrict__ step,
                                                         long long grad_stride, int n_vec, float* __restrict__ loss_acc) {
  pdl_wait();
  pdl_launch_dependents();
  const int v = blockIdx.x * 256 + threadIdx.x;
  if (v == n_vec && loss_acc != nullptr) {        // the vector behind the parameters: {batch-mean nll, #correct} of every CTA
    float l = 0.f, c = 0.f;
    for (int sl = 0; sl < n_slots; ++sl) {
      l += __ldcg(partials + (size_t)sl * slot_stride + (size_t)n_vec * 4);
      c += __ldcg(partials + (size_t)sl * slot_stride + (size_t)n_vec * 4 + 1);
    }
    loss_acc[0] += l;                             // the only writer of loss_acc while this kernel runs
    loss_acc[1] += c;
  }
  if (v >= n_vec) return;
  const unsigned long long st = step != nullptr ? *step : 0ull;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sl = 0; sl < n_slots; ++sl) {
    const float4 q = __ldcg(reinterpret_cast<const float4*>(partials + (size_t)sl * slot_stride) + v);
    acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
  }
  reinterpret_cast<float4*>(grads + (size_t)(st & 1ull) * (size_t)grad_stride)[v] = acc;
}

// Plain flat momentum SGD (generic models: gradients already averaged in `grad`)
__global__ void __launch_bounds__(256) sgd_flat_kernel(float* __restrict__ p, float* __restrict__ m,
                                                       const float* __restrict__ g, size_t n, float lr, float mu,
                                                       float wd, int zero_grad, float* gw) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = n / 4;
  if (i < n4) {
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    gv.x = fmaf(wd, pv.x, gv.x); gv.y = fmaf(wd, pv.y, gv.y); gv.z = fmaf(wd, pv.z, gv.z); gv.w = fmaf(wd, pv.w, gv.w);
    mv.x = fmaf(mu, mv.x, gv.x); mv.y = fmaf(mu, mv.y, gv.y); mv.z = fmaf(mu, mv.z, gv.z); mv.w = fmaf(mu, mv.w, gv.w);
    pv.x = fmaf(-lr, mv.x, pv.x); pv.y = fmaf(-lr, mv.y, pv.y); pv.z = fmaf(-lr, mv.z, pv.z); pv.w = fmaf(-lr, mv.w, pv.w);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    if (zero_grad) reinterpret_cast<float4*>(gw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (i == 0) {
    for (size_t t = n4 * 4; t < n; ++t) {      // tail (< 4 elements)
      const float gt = fmaf(wd, p[t], g[t]);
      m[t] = fmaf(mu, m[t], gt);
      p[t] = fmaf(-lr, m[t], p[t]);
      if (zero_grad) gw[t] = 0.f;
    }
  }
}

}  // namespace b2

extern "C" {

int b2_allreduce_sgd_launch(const PeerPtrs* grads, const b2::SignalPads* sig, float* params, float* momentum,
                            unsigned long long* step, size_t n_elems, float lr, float mu, float scale, int rank,
                            int world, int zero_grads, long long grad_stride, unsigned int* done_counter, float* aux,
                            const PeerPtrs* inbox, const float* loss_acc, float* loss_snapshot, int wire_bf16,
                            unsigned int* snap_flag, unsigned int snap_gen, cudaStream_t stream) {
  b2::SgdArgs a;
  a.wire_bf16 = wire_bf16;
  a.loss_acc = loss_acc; a.loss_snapshot = (loss_acc != nullptr) ? loss_snapshot : nullptr;
  a.snap_flag = (a.loss_snapshot != nullptr) ? snap_flag : nullptr; a.snap_gen = snap_gen;
  memset(&a.inbox, 0, sizeof(a.inbox));
  // push ("LL") exchange: needs an inbox on every rank, the double-buffered buckets and the device step counter (its epoch)
  const bool push = inbox != nullptr && world > 1;
  if (push) {
    if (step == nullptr || grad_stride <= 0) return (int)cudaErrorInvalidValue;
    a.inbox = *inbox;
  }
  a.grads = *grads; a.sig = *sig; a.params = params; a.momentum = momentum; a.step = step;
  a.n_vec = n_elems / 4; a.lr = lr; a.mu = mu; a.scale = scale; a.rank = rank; a.world = world;
  a.zero_grads = zero_grads; a.grad_stride = grad_stride; a.done_counter = done_counter; a.aux = aux;
  if (step != nullptr && done_counter == nullptr) return (int)cudaErrorInvalidValue;
  size_t blocks = (a.n_vec + b2::kSgdThreads - 1) / b2::kSgdThreads;
  if (blocks < 1) blocks = 1;
  if (blocks > 64) blocks = 64;
  static const int pdl = [] { const char* e = getenv("B200DIST_PDL"); return (e == nullptr || e[0] != '0') ? 1 : 0; }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3((unsigned)b2::kSgdThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return push ? (int)cudaLaunchKernelEx(&cfg, b2::allreduce_sgd_push_kernel, a)
              : (int)cudaLaunchKernelEx(&cfg, b2::allreduce_sgd_kernel, a);
}

int b2_det_reduce_launch(const float* partials, int n_slots, long long slot_stride, float* grads, const unsigned long long* step,
                         long long grad_stride, size_t n_elems, float* loss_acc, cudaStream_t stream) {
  const int n_vec = (int)(n_elems / 4);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)((n_vec + 1 + 255) / 256));
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, b2::det_reduce_kernel, partials, n_slots, slot_stride, grads, step, grad_stride, n_vec, loss_acc);
}

int b2_sgd_flat_launch(float* p, float* m, const float* g, size_t n, float lr, float mu, float wd, int zero_grad,
                       cudaStream_t stream) {
  const size_t n4 = (n / 4) > 0 ? n / 4 : 1;
  const unsigned blocks = (unsigned)((n4 + 255) / 256);
  b2::sgd_flat_kernel<<<blocks, 256, 0, stream>>>(p, m, g, n, lr, mu, wd, zero_grad, const_cast<float*>(g));
  return (int)cudaGetLastError();
}

}  // extern "C"
