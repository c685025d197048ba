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
// Chunk pipeline (the default when the loader ring is deep enough: num_slots % K == 0, num_slots >= 3K): K consecutive steps
// are THREE graph launches, one per stream, ordered by three events instead of 9 driver calls per step:
//   copy stream    : graph of K H2D copy nodes   (the K pinned loader slots of a slot group -> device block group g)
//   compute stream : graph of 2K kernel nodes    (a pure kernel chain: programmatic dependent launch stays intact across the
//                                                 K steps -- round 1's chunk graph had an H2D -> kernel edge in front of every
//                                                 step, which cost as much as a graph boundary, profiles/executor_chunk_graphs.json)
//   d2h stream     : graph of K D2H copy nodes   (the cumulative loss after each step, snapshotted on the device by that
//                                                 step's SGD kernel into loss_hist[g*K + j], -> the step's pinned loss word)
// Chunk c+1's copies run while chunk c computes (two device block groups), chunk c-1's losses drain meanwhile.  Every step
// still has its own H2D copy from pinned memory and its own D2H read-back; the per-step path remains for the steps that do
// not fill a chunk (epoch tails, max_steps budgets).
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
                            const PeerPtrsC* inbox, const float* loss_acc, float* loss_snapshot, cudaStream_t stream);
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
  chunk_ok_ = K >= 2 && K <= 8 && nb % K == 0 && nb >= 3 * K && cfg_.loss_hist != nullptr;
  if (chunk_ok_) {
    h2d_exec_.assign(2 * (nb / K), nullptr);
    d2h_exec_.assign(2 * (nb / K), nullptr);
    for (int g = 0; g < 2; ++g) {
      cudaEventCreateWithFlags(&h2d_done_[g], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&comp_done_[g], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&d2h_done_[g], cudaEventDisableTiming);
    }
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
  for (auto g : h2d_exec_) if (g) cudaGraphExecDestroy(g);
  for (auto g : d2h_exec_) if (g) cudaGraphExecDestroy(g);
  for (int g = 0; g < 2; ++g) {
    if (comp_exec_[g]) cudaGraphExecDestroy(comp_exec_[g]);
    if (h2d_done_[g]) cudaEventDestroy(h2d_done_[g]);
    if (comp_done_[g]) cudaEventDestroy(comp_done_[g]);
    if (d2h_done_[g]) cudaEventDestroy(d2h_done_[g]);
  }
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
void StepExecutor::record_step(const void* x, const long long* y, float* loss_snapshot) {
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
                                    cfg_.done_counter, cfg_.aux, cfg_.push ? &ib : nullptr, cfg_.loss_acc, loss_snapshot, compute_);
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

// The three graphs of one chunk: loader slot group `sg` (pinned slots sg*K .. sg*K+K-1), device block group `g` (0/1).
static bool end_capture(cudaStream_t st, cudaGraphExec_t* out, std::string* err, const char* what) {
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(st, &graph);
  if (!err->empty() || e != cudaSuccess || graph == nullptr) {
    if (err->empty()) *err = std::string(what) + " capture failed: " + cudaGetErrorString(e);
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return false;
  }
  e = cudaGraphInstantiate(out, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { *err = std::string("cudaGraphInstantiate(") + what + "): " + cudaGetErrorString(e); cudaGetLastError(); return false; }
  return true;
}

bool StepExecutor::capture_chunk(int sg, int g) {
  const int K = cfg_.chunk;
  err_.clear();
  if (comp_exec_[g] == nullptr) {                      // 2K kernels: a pure chain, PDL intact from step to step
    cudaError_t e = cudaStreamBeginCapture(compute_, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { err_ = std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e); return false; }
    for (int j = 0; j < K; ++j) {
      unsigned char* blk = cfg_.in_dev[g * K + j];
      record_step(blk, reinterpret_cast<const long long*>(blk + loader_->y_offset()), cfg_.loss_hist + 2 * (g * K + j));
    }
    if (!end_capture(compute_, &comp_exec_[g], &err_, "compute chunk")) return false;
  }
  if (h2d_exec_[sg * 2 + g] == nullptr) {              // K H2D copies, one per step, from that step's pinned loader slot
    cudaError_t e = cudaStreamBeginCapture(copy_, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { err_ = std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e); return false; }
    for (int j = 0; j < K; ++j)
      cudaMemcpyAsync(cfg_.in_dev[g * K + j], loader_->slot(sg * K + j).x, loader_->block_bytes(), cudaMemcpyHostToDevice, copy_);
    if (!end_capture(copy_, &h2d_exec_[sg * 2 + g], &err_, "h2d chunk")) return false;
  }
  if (d2h_exec_[sg * 2 + g] == nullptr) {              // K D2H copies: the loss as of each step -> that step's pinned word
    cudaError_t e = cudaStreamBeginCapture(d2h_, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { err_ = std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e); return false; }
    for (int j = 0; j < K; ++j)
      cudaMemcpyAsync(slots_[sg * K + j].loss_pin, cfg_.loss_hist + 2 * (g * K + j), 2 * sizeof(float), cudaMemcpyDeviceToHost, d2h_);
    if (!end_capture(d2h_, &d2h_exec_[sg * 2 + g], &err_, "d2h chunk")) return false;
  }
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
      const int sg = (int)((loader_->consumed() % nb) / K);
      const int g = (int)(chunks_issued_ & 1);
      if ((comp_exec_[g] == nullptr || h2d_exec_[sg * 2 + g] == nullptr || d2h_exec_[sg * 2 + g] == nullptr) &&
          !capture_chunk(sg, g)) {
        chunk_ok_ = false;
        chunk_note_ = err_;
        err_.clear();
        continue;
      }
      int first = -1;
      for (int j = 0; j < K; ++j) {                                 // blocks until the K batches are staged
        int64_t count = 0;
        const int slot = loader_->next(&count);
        if (j == 0) first = slot;
        if (slot != sg * K + j || count != cfg_.B) { err_ = "chunk path: loader slot sequence broke"; return -1; }
      }
      // copies: device block group g is free once the chunk that last read it (two chunks ago) has computed; blocks 0/1
      // are also the per-step path's double buffer
      cudaStreamWaitEvent(copy_, comp_done_[g], 0);
      cudaStreamWaitEvent(copy_, kernels_done_[0], 0);
      cudaStreamWaitEvent(copy_, kernels_done_[1], 0);
      cudaError_t e = cudaGraphLaunch(h2d_exec_[sg * 2 + g], copy_);
      if (e != cudaSuccess) { err_ = std::string("cudaGraphLaunch(h2d chunk): ") + cudaGetErrorString(e); return -1; }
      cudaEventRecord(h2d_done_[g], copy_);
      // kernels: after the copies, and after the losses of the chunk that last used snapshot group g have been read back
      cudaStreamWaitEvent(compute_, h2d_done_[g], 0);
      cudaStreamWaitEvent(compute_, d2h_done_[g], 0);
      e = cudaGraphLaunch(comp_exec_[g], compute_);
      if (e != cudaSuccess) { err_ = std::string("cudaGraphLaunch(compute chunk): ") + cudaGetErrorString(e); return -1; }
      cudaEventRecord(comp_done_[g], compute_);
      if (g == 0) { cudaEventRecord(kernels_done_[0], compute_); cudaEventRecord(kernels_done_[1], compute_); }
      // losses
      cudaStreamWaitEvent(d2h_, comp_done_[g], 0);
      e = cudaGraphLaunch(d2h_exec_[sg * 2 + g], d2h_);
      if (e != cudaSuccess) { err_ = std::string("cudaGraphLaunch(d2h chunk): ") + cudaGetErrorString(e); return -1; }
      cudaEventRecord(d2h_done_[g], d2h_);
      const int last = first + K - 1;
      cudaEventRecord(slots_[last].done, d2h_);
      for (int j = 0; j < K; ++j) in_flight_.push_back({first + j, last});
      ++chunks_issued_;
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
    // H2D: block[p] is free once the kernels that last read it (two steps ago, or a chunk of group 0) are done
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
