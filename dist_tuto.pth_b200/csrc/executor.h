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
  float* grads_local;                // this rank's gradient bucket
  void* grad_ptrs[8];                // every rank's bucket (symmetric mapping), [0] only when world == 1
  uint32_t* sig_ptrs[8];
  unsigned long long* step_counter;
  float* loss_acc;                   // [2]
  void* x_dev;                       // [B,1,28,28] fp32 or uint8
  long long* y_dev;                  // [B]
  int B, x_u8, training, rank, world;
  unsigned long long seed;
  long long sample_base;
  float lr, mu, p_drop;
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

 private:
  struct Slot {
    cudaGraphExec_t exec = nullptr;
    cudaEvent_t done = nullptr;
    float* loss_pin = nullptr;
  };
  bool capture(int slot);
  void retire_oldest();
  StepConfig cfg_;
  NativeLoader* loader_;
  int max_in_flight_;
  cudaStream_t stream_ = nullptr;
  std::vector<Slot> slots_;
  std::deque<int> in_flight_;
  double last_loss_ = 0.0;
  std::string err_;
};

}  // namespace b2
