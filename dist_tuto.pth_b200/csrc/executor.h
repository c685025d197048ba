// Native step executor -- see executor.cpp.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <deque>
#include <string>
#include <vector>

#include "loader.h"

namespace b2 {

struct StepConfig {
  float* params;
  float* momentum;
  float* grads_local;                // this rank's gradient bucket(s)
  void* grad_ptrs[8];                // every rank's bucket (symmetric mapping), [0] only when world == 1
  uint32_t* sig_ptrs[8];
  unsigned long long* step_counter;
  unsigned int* done_counter;
  float* loss_acc;                   // [2]
  unsigned char* in_dev[8];          // device input blocks, same layout as a loader slot: [x | pad | y]; [0..1] double-buffer
                                     // the per-step path, [0..chunk) are the blocks of a chunk graph
  int chunk;                         // steps per chunk graph (0/1 = per-step launches only), <= 8
  int B, x_u8, training, rank, world, cluster;
  unsigned long long seed;
  long long sample_base, grad_stride;
  float lr, mu, p_drop;
  float* aux;                        // conv2.weight in the kernels' smem layouts (maintained by the SGD kernel)
  void* inbox_ptrs[8];               // push exchange (sgd.cu): every rank's inbox
  int push;
};

class StepExecutor {
 public:
  StepExecutor(const StepConfig& cfg, NativeLoader* loader, int max_in_flight);
  ~StepExecutor();
  // Runs up to `max_steps` full-batch steps of the loader's current epoch.  Returns the number of steps done;
  // *pending_slot >= 0 (with *pending_count) if a short batch was fetched but not processed (caller handles it),
  // *epoch_done is set when the loader ran dry.
  int64_t run(int64_t max_steps, int* pending_slot, int64_t* pending_count, int* epoch_done);
  void drain();                      // wait for everything in flight, release loader slots
  double last_loss_cumulative() const { return last_loss_; }
  const std::string& error() const { return err_; }
  bool chunking() const { return chunk_ok_; }
  const std::string& chunk_note() const { return chunk_note_; }   // why chunk graphs were turned off (if they were)

 private:
  struct Slot {
    cudaEvent_t done = nullptr;      // loss of the step fed from this slot has landed in loss_pin
    float* loss_pin = nullptr;
  };
  bool capture(int parity);
  bool capture_chunk(int group);
  void record_step(const void* x, const long long* y);
  void retire_oldest();
  StepConfig cfg_;
  NativeLoader* loader_;
  int max_in_flight_;
  cudaStream_t copy_ = nullptr, compute_ = nullptr, d2h_ = nullptr;
  cudaGraphExec_t exec_[2] = {nullptr, nullptr};     // the two kernels, reading in_dev[parity]
  cudaEvent_t copied_[2] = {nullptr, nullptr}, kernels_done_[2] = {nullptr, nullptr};
  std::vector<Slot> slots_;
  // chunk graphs: K consecutive steps (K H2D copies, 2K kernels, K loss read-backs) as ONE graph launch; one graph per
  // group of K loader slots (the slot of batch b is b % num_slots, so the pinned addresses of a group are fixed)
  std::vector<cudaGraphExec_t> chunk_exec_;
  std::vector<cudaEvent_t> chunk_ev_;  // capture-time fork/join events
  bool chunk_ok_ = false;
  struct Flight { int slot, ev_slot; };
  std::deque<Flight> in_flight_;
  int64_t issued_ = 0;
  double last_loss_ = 0.0;
  std::string err_, chunk_note_;
};

}  // namespace b2
