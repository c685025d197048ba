// Native step executor: the training hot loop without the Python interpreter.
//
// The reference's hot loop (train_dist.py:115-124) is Python: DataLoader iteration, five framework calls and an
// optimizer step per batch.  Here the loop that feeds the GPU is C++ and software-pipelined over three streams:
//
//   copy stream    : ONE cudaMemcpyAsync per step -- the pinned loader slot [x | y] -> device input block[p]   (p = step & 1)
//   compute stream : waits for that copy, then replays a 2-kernel CUDA graph (convnet_step + allreduce_sgd) reading block[p]
//   d2h stream     : copies the running loss to the slot's pinned word; its event retires the step
//
// so the H2D copy of step i+1 and the loss read-back of step i-1 overlap the kernels of step i.  At most `max_in_flight`
// steps are outstanding; a retired step hands its slot back to the prefetch thread (loader.cpp).  The GIL is released
// around run().
//
// Chunk graphs: when the loader ring is deep enough (num_slots % K == 0, num_slots >= 2K), K consecutive steps are ONE
// graph launch -- the same three-branch pipeline captured across the three streams (K H2D nodes from the K pinned slots
// of a slot group, 2K kernel nodes chained with programmatic dependent launch, K loss read-back nodes).  That removes
// ~9 driver calls per step from the host and the inter-graph gaps from the device; the per-step path remains for the
// steps that do not fill a chunk (epoch tails, max_steps budgets).
#include "executor.h"

#include <cstring>

extern "C" {
int b2_convnet_step_launch(const float* params, float* grads, const void* x, int x_u8, const long long* target,
                           float* loss_acc, float* out_logp, float* mask_out, const unsigned long long* step,
                           unsigned long long seed, long long sample_base, int B, int training, int backward,
                           float inv_bsz, float p_drop, int max_ctas, long long grad_stride, const float* aux, cudaStream_t stream);
struct PeerPtrsC { void* p[8]; };
struct SignalPadsC { uint32_t* pad[8]; };
int b2_allreduce_sgd_launch(const PeerPtrsC* grads, const SignalPadsC* sig, float* params, float* momentum,
                            unsigned long long* step, size_t n_elems, float lr, float mu, float scale, int rank,
                            int world, int zero_grads, long long grad_stride, unsigned int* done_counter, float* aux,
                            const PeerPtrsC* inbox, cudaStream_t stream);
int b2_convnet_cluster_launch(const float* params, float* grads, const void* x, int x_u8, const long long* target,
                              float* loss_acc, float* out_logp, float* mask_out, const unsigned long long* step,
                              unsigned long long seed, long long sample_base, int B, int training, int backward,
                              float inv_bsz, float p_drop, int cluster, int max_clusters, long long grad_stride,
                              const float* aux, cudaStream_t stream);
int b2_convnet_npar();
}

namespace b2 {

StepExecutor::StepExecutor(const StepConfig& cfg, NativeLoader* loader, int max_in_flight)
    : cfg_(cfg), loader_(loader), max_in_flight_(max_in_flight < 1 ? 1 : max_in_flight) {
  cudaStreamCreateWithFlags(&copy_, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&compute_, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&d2h_, cudaStreamNonBlocking);
  for (int p = 0; p < 2; ++p) {
    cudaEventCreateWithFlags(&copied_[p], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&kernels_done_[p], cudaEventDisableTiming);
  }
  const int K = cfg_.chunk, nb = loader_->num_slots();
  chunk_ok_ = K >= 2 && K <= 8 && nb % K == 0 && nb >= 2 * K;
  if (chunk_ok_) {
    chunk_exec_.assign(nb / K, nullptr);
    chunk_ev_.resize(2 * K + 2);
    for (auto& e : chunk_ev_) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  }
  slots_.resize(loader_->num_slots());
  for (auto& s : slots_) {
    cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    cudaHostAlloc((void**)&s.loss_pin, 2 * sizeof(float), cudaHostAllocDefault);
    s.loss_pin[0] = s.loss_pin[1] = 0.f;
  }
}

StepExecutor::~StepExecutor() {
  drain();
  cudaStreamSynchronize(compute_);
  for (auto& s : slots_) {
    if (s.done) cudaEventDestroy(s.done);
    if (s.loss_pin) cudaFreeHost(s.loss_pin);
  }
  for (auto g : chunk_exec_) if (g) cudaGraphExecDestroy(g);
  for (auto e : chunk_ev_) if (e) cudaEventDestroy(e);
  for (int p = 0; p < 2; ++p) {
    if (exec_[p]) cudaGraphExecDestroy(exec_[p]);
    if (copied_[p]) cudaEventDestroy(copied_[p]);
    if (kernels_done_[p]) cudaEventDestroy(kernels_done_[p]);
  }
  if (copy_) cudaStreamDestroy(copy_);
  if (compute_) cudaStreamDestroy(compute_);
  if (d2h_) cudaStreamDestroy(d2h_);
}

// Enqueues the two kernels of one step on the compute stream (called under stream capture).
void StepExecutor::record_step(const void* x, const long long* y) {
  int rc = cfg_.cluster > 1
               ? b2_convnet_cluster_launch(cfg_.params, cfg_.grads_local, x, cfg_.x_u8, y, cfg_.loss_acc, nullptr, nullptr,
                                           cfg_.step_counter, cfg_.seed, cfg_.sample_base, cfg_.B, cfg_.training, 1,
                                           1.f / cfg_.B, cfg_.p_drop, cfg_.cluster, 0, cfg_.grad_stride, cfg_.aux, compute_)
               : b2_convnet_step_launch(cfg_.params, cfg_.grads_local, x, cfg_.x_u8, y, cfg_.loss_acc, nullptr, nullptr,
                                        cfg_.step_counter, cfg_.seed, cfg_.sample_base, cfg_.B, cfg_.training, 1, 1.f / cfg_.B,
                                        cfg_.p_drop, 0, cfg_.grad_stride, cfg_.aux, compute_);
  PeerPtrsC g;
  SignalPadsC sg;
  std::memcpy(g.p, cfg_.grad_ptrs, sizeof(g.p));
  std::memcpy(sg.pad, cfg_.sig_ptrs, sizeof(sg.pad));
  PeerPtrsC ib;
  std::memcpy(ib.p, cfg_.inbox_ptrs, sizeof(ib.p));
  int rc2 = b2_allreduce_sgd_launch(&g, &sg, cfg_.params, cfg_.momentum, cfg_.step_counter, (size_t)b2_convnet_npar(),
                                    cfg_.lr, cfg_.mu, 1.f / cfg_.world, cfg_.rank, cfg_.world, 1, cfg_.grad_stride,
                                    cfg_.done_counter, cfg_.aux, cfg_.push ? &ib : nullptr, compute_);
  if ((rc != 0 || rc2 != 0) && err_.empty())
    err_ = std::string("kernel launch failed: ") + cudaGetErrorString((cudaError_t)(rc ? rc : rc2));
}

bool StepExecutor::capture(int parity) {
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamBeginCapture(compute_, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) { err_ = std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e); return false; }
  err_.clear();
  record_step(cfg_.in_dev[parity], reinterpret_cast<const long long*>(cfg_.in_dev[parity] + loader_->y_offset()));
  e = cudaStreamEndCapture(compute_, &graph);
  if (!err_.empty() || e != cudaSuccess || graph == nullptr) {
    if (err_.empty()) err_ = std::string("graph capture failed: ") + cudaGetErrorString(e);
    if (graph) cudaGraphDestroy(graph);
    return false;
  }
  e = cudaGraphInstantiate(&exec_[parity], graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { err_ = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return false; }
  return true;
}

// One graph = K pipelined steps fed from loader slots [group*K, group*K + K).  Captured over the three streams:
//   copy_   : H2D_0 .. H2D_{K-1}                       (each followed by an event the matching step waits for)
//   compute_: step_0, sgd_0, step_1, sgd_1, ...         (step_j waits for H2D_j only)
//   d2h_    : loss_j after sgd_j                        (joined back into compute_ at the end)
bool StepExecutor::capture_chunk(int group) {
  const int K = cfg_.chunk;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamBeginCapture(compute_, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) { err_ = std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e); return false; }
  err_.clear();
  cudaEvent_t fork = chunk_ev_[2 * K], join = chunk_ev_[2 * K + 1];
  cudaEventRecord(fork, compute_);
  cudaStreamWaitEvent(copy_, fork, 0);
  cudaStreamWaitEvent(d2h_, fork, 0);
  for (int j = 0; j < K; ++j) {
    cudaMemcpyAsync(cfg_.in_dev[j], loader_->slot(group * K + j).x, loader_->block_bytes(), cudaMemcpyHostToDevice, copy_);
    cudaEventRecord(chunk_ev_[j], copy_);
  }
  for (int j = 0; j < K; ++j) {
    cudaStreamWaitEvent(compute_, chunk_ev_[j], 0);
    record_step(cfg_.in_dev[j], reinterpret_cast<const long long*>(cfg_.in_dev[j] + loader_->y_offset()));
    cudaEventRecord(chunk_ev_[K + j], compute_);
    cudaStreamWaitEvent(d2h_, chunk_ev_[K + j], 0);
    cudaMemcpyAsync(slots_[group * K + j].loss_pin, cfg_.loss_acc, 2 * sizeof(float), cudaMemcpyDeviceToHost, d2h_);
  }
  cudaEventRecord(join, d2h_);
  cudaStreamWaitEvent(compute_, join, 0);
  e = cudaStreamEndCapture(compute_, &graph);
  if (!err_.empty() || e != cudaSuccess || graph == nullptr) {
    if (err_.empty()) err_ = std::string("chunk graph capture failed: ") + cudaGetErrorString(e);
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return false;
  }
  e = cudaGraphInstantiate(&chunk_exec_[group], graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { err_ = std::string("cudaGraphInstantiate(chunk): ") + cudaGetErrorString(e); cudaGetLastError(); return false; }
  return true;
}

void StepExecutor::retire_oldest() {
  const Flight f = in_flight_.front();
  in_flight_.pop_front();
  cudaEventSynchronize(slots_[f.ev_slot].done);
  const int s = f.slot;
  last_loss_ = (double)slots_[s].loss_pin[0];      // host read of this step's D2H loss copy
  loader_->release();
}

void StepExecutor::drain() {
  while (!in_flight_.empty()) retire_oldest();
}

int64_t StepExecutor::run(int64_t max_steps, int* pending_slot, int64_t* pending_count, int* epoch_done) {
  *pending_slot = -1;
  *pending_count = 0;
  *epoch_done = 0;
  int64_t done = 0;
  const int K = cfg_.chunk, nb = loader_->num_slots();
  while (max_steps < 0 || done < max_steps) {
    // ---- chunk path: K full batches ahead, slot group aligned, budget allows
    if (chunk_ok_ && (max_steps < 0 || max_steps - done >= K) && loader_->consumed() % K == 0 &&
        loader_->full_batches_left() >= K) {
      while ((int)in_flight_.size() > nb - K) retire_oldest();     // the prefetch thread needs K free slots to fill
      const int group = (int)((loader_->consumed() % nb) / K);
      if (chunk_exec_[group] == nullptr && !capture_chunk(group)) {
        chunk_ok_ = false;                                          // e.g. a driver that rejects the mixed edge types
        chunk_note_ = err_;
        err_.clear();
        continue;
      }
      int first = -1;
      for (int j = 0; j < K; ++j) {                                 // blocks until the K batches are staged
        int64_t count = 0;
        const int slot = loader_->next(&count);
        if (j == 0) first = slot;
        if (slot != group * K + j || count != cfg_.B) { err_ = "chunk path: loader slot sequence broke"; return -1; }
      }
      cudaError_t e = cudaGraphLaunch(chunk_exec_[group], compute_);
      if (e != cudaSuccess) { err_ = std::string("cudaGraphLaunch(chunk): ") + cudaGetErrorString(e); return -1; }
      const int last = first + K - 1;
      cudaEventRecord(slots_[last].done, compute_);
      cudaEventRecord(kernels_done_[0], compute_);                  // blocks 0/1 are shared with the per-step path
      cudaEventRecord(kernels_done_[1], compute_);
      for (int j = 0; j < K; ++j) in_flight_.push_back({first + j, last});
      issued_ += K;
      done += K;
      continue;
    }
    while ((int)in_flight_.size() >= max_in_flight_) retire_oldest();
    int64_t count = 0;
    const int slot = loader_->next(&count);
    if (slot < 0) { *epoch_done = 1; break; }
    if (count != cfg_.B) {            // short tail batch: give it back to the caller (eager path)
      drain();
      cudaStreamSynchronize(compute_);
      *pending_slot = slot;
      *pending_count = count;
      break;
    }
    const int p = (int)(issued_ & 1);
    if (exec_[p] == nullptr && !capture(p)) return -1;
    // H2D: block[p] is free once the kernels that last read it (two steps ago) are done
    cudaStreamWaitEvent(copy_, kernels_done_[p], 0);
    cudaMemcpyAsync(cfg_.in_dev[p], loader_->slot(slot).x, loader_->block_bytes(), cudaMemcpyHostToDevice, copy_);
    cudaEventRecord(copied_[p], copy_);
    // kernels
    cudaStreamWaitEvent(compute_, copied_[p], 0);
    cudaError_t e = cudaGraphLaunch(exec_[p], compute_);
    if (e != cudaSuccess) { err_ = std::string("cudaGraphLaunch: ") + cudaGetErrorString(e); return -1; }
    cudaEventRecord(kernels_done_[p], compute_);
    // loss read-back
    cudaStreamWaitEvent(d2h_, kernels_done_[p], 0);
    cudaMemcpyAsync(slots_[slot].loss_pin, cfg_.loss_acc, 2 * sizeof(float), cudaMemcpyDeviceToHost, d2h_);
    cudaEventRecord(slots_[slot].done, d2h_);
    in_flight_.push_back({slot, slot});
    ++issued_;
    ++done;
  }
  return done;
}

}  // namespace b2
