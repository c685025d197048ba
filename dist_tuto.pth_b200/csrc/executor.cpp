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
                            int world, int zero_grads, long long grad_stride, unsigned int* done_counter, float* aux, cudaStream_t stream);
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
  for (int p = 0; p < 2; ++p) {
    if (exec_[p]) cudaGraphExecDestroy(exec_[p]);
    if (copied_[p]) cudaEventDestroy(copied_[p]);
    if (kernels_done_[p]) cudaEventDestroy(kernels_done_[p]);
  }
  if (copy_) cudaStreamDestroy(copy_);
  if (compute_) cudaStreamDestroy(compute_);
  if (d2h_) cudaStreamDestroy(d2h_);
}

bool StepExecutor::capture(int parity) {
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamBeginCapture(compute_, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) { err_ = std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e); return false; }
  const void* x = cfg_.in_dev[parity];
  const long long* y = reinterpret_cast<const long long*>(cfg_.in_dev[parity] + loader_->y_offset());
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
  int rc2 = b2_allreduce_sgd_launch(&g, &sg, cfg_.params, cfg_.momentum, cfg_.step_counter, (size_t)b2_convnet_npar(),
                                    cfg_.lr, cfg_.mu, 1.f / cfg_.world, cfg_.rank, cfg_.world, 1, cfg_.grad_stride,
                                    cfg_.done_counter, cfg_.aux, compute_);
  e = cudaStreamEndCapture(compute_, &graph);
  if (rc != 0 || rc2 != 0 || e != cudaSuccess || graph == nullptr) {
    err_ = std::string("graph capture failed: ") + cudaGetErrorString(e != cudaSuccess ? e : (cudaError_t)(rc ? rc : rc2));
    if (graph) cudaGraphDestroy(graph);
    return false;
  }
  e = cudaGraphInstantiate(&exec_[parity], graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { err_ = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return false; }
  return true;
}

void StepExecutor::retire_oldest() {
  const int s = in_flight_.front();
  in_flight_.pop_front();
  cudaEventSynchronize(slots_[s].done);
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
  while (max_steps < 0 || done < max_steps) {
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
    in_flight_.push_back(slot);
    ++issued_;
    ++done;
  }
  return done;
}

}  // namespace b2
