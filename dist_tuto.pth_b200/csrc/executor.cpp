// Native step executor: the training hot loop without the Python interpreter.
//
// The reference's hot loop (train_dist.py:115-124) is Python: DataLoader iteration, five framework calls and an
// optimizer step per batch.  Here one training step is one cudaGraphLaunch, and the loop that feeds it is C++:
//   for each full batch the prefetcher (loader.cpp) has staged in pinned slot s:
//       launch graph[s] = { H2D x, H2D y, convnet_step, allreduce_sgd, D2H running loss }   (captured once per slot)
//       record an event; keep at most `max_in_flight` steps outstanding; when a step retires, read its loss from
//       the pinned D2H copy and hand its slot back to the prefetcher.
// The GIL is released around run(), so the host side costs ~2-3 us per step instead of tens.
#include "executor.h"

#include <cstring>

extern "C" {
int b2_convnet_step_launch(const float* params, float* grads, const void* x, int x_u8, const long long* target,
                           float* loss_acc, float* out_logp, float* mask_out, const unsigned long long* step,
                           unsigned long long seed, long long sample_base, int B, int training, int backward,
                           float inv_bsz, float p_drop, int max_ctas, cudaStream_t stream);
struct PeerPtrsC { void* p[8]; };
struct SignalPadsC { uint32_t* pad[8]; };
int b2_allreduce_sgd_launch(const PeerPtrsC* grads, const SignalPadsC* sig, float* params, float* momentum,
                            unsigned long long* step, size_t n_elems, float lr, float mu, float scale, int rank,
                            int world, int zero_grads, cudaStream_t stream);
int b2_convnet_npar();
}

namespace b2 {

#define EX_CK(call)                                                                  \
  do {                                                                               \
    cudaError_t _e = (call);                                                         \
    if (_e != cudaSuccess) { err_ = std::string(#call) + ": " + cudaGetErrorString(_e); return false; } \
  } while (0)

StepExecutor::StepExecutor(const StepConfig& cfg, NativeLoader* loader, int max_in_flight)
    : cfg_(cfg), loader_(loader), max_in_flight_(max_in_flight < 1 ? 1 : max_in_flight) {
  cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking);
  slots_.resize(loader_->num_slots());
  for (auto& s : slots_) {
    cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    cudaHostAlloc((void**)&s.loss_pin, 2 * sizeof(float), cudaHostAllocDefault);
    s.loss_pin[0] = s.loss_pin[1] = 0.f;
  }
}

StepExecutor::~StepExecutor() {
  drain();
  for (auto& s : slots_) {
    if (s.exec) cudaGraphExecDestroy(s.exec);
    if (s.done) cudaEventDestroy(s.done);
    if (s.loss_pin) cudaFreeHost(s.loss_pin);
  }
  if (stream_) cudaStreamDestroy(stream_);
}

bool StepExecutor::capture(int slot) {
  const NativeLoader::Slot& ls = loader_->slot(slot);
  const size_t xbytes = (size_t)cfg_.B * loader_->item() * (cfg_.x_u8 ? 1 : sizeof(float));
  cudaGraph_t graph = nullptr;
  EX_CK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
  cudaMemcpyAsync(cfg_.x_dev, ls.x, xbytes, cudaMemcpyHostToDevice, stream_);
  cudaMemcpyAsync(cfg_.y_dev, ls.y, (size_t)cfg_.B * sizeof(long long), cudaMemcpyHostToDevice, stream_);
  int rc = b2_convnet_step_launch(cfg_.params, cfg_.grads_local, cfg_.x_dev, cfg_.x_u8, cfg_.y_dev, cfg_.loss_acc, nullptr,
                                  nullptr, cfg_.step_counter, cfg_.seed, cfg_.sample_base, cfg_.B, cfg_.training, 1,
                                  1.f / cfg_.B, cfg_.p_drop, 0, stream_);
  PeerPtrsC g;
  SignalPadsC sg;
  std::memcpy(g.p, cfg_.grad_ptrs, sizeof(g.p));
  std::memcpy(sg.pad, cfg_.sig_ptrs, sizeof(sg.pad));
  int rc2 = b2_allreduce_sgd_launch(&g, &sg, cfg_.params, cfg_.momentum, cfg_.step_counter, (size_t)b2_convnet_npar(),
                                    cfg_.lr, cfg_.mu, 1.f / cfg_.world, cfg_.rank, cfg_.world, 1, stream_);
  cudaMemcpyAsync(slots_[slot].loss_pin, cfg_.loss_acc, 2 * sizeof(float), cudaMemcpyDeviceToHost, stream_);
  cudaError_t e = cudaStreamEndCapture(stream_, &graph);
  if (rc != 0 || rc2 != 0 || e != cudaSuccess || graph == nullptr) {
    err_ = std::string("graph capture failed: ") + cudaGetErrorString(e != cudaSuccess ? e : (cudaError_t)(rc ? rc : rc2));
    if (graph) cudaGraphDestroy(graph);
    return false;
  }
  e = cudaGraphInstantiate(&slots_[slot].exec, graph, 0);
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
      *pending_slot = slot;
      *pending_count = count;
      break;
    }
    if (slots_[slot].exec == nullptr && !capture(slot)) return -1;
    cudaError_t e = cudaGraphLaunch(slots_[slot].exec, stream_);
    if (e != cudaSuccess) { err_ = std::string("cudaGraphLaunch: ") + cudaGetErrorString(e); return -1; }
    cudaEventRecord(slots_[slot].done, stream_);
    in_flight_.push_back(slot);
    ++done;
  }
  return done;
}

}  // namespace b2
