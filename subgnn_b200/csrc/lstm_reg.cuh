// Register-resident / cluster LSTM recurrence (lstm_reg.cu); dispatched from the C-ABI wrappers in lstm.cu.
#pragma once
#include <cuda_runtime.h>
bool lstm_reg_supported(int H);
int lstm_reg_cluster(int H);   // CTAs per cluster (1 or 2) == how many unit ranges the forward weights are permuted for
int lstm_reg_fwd(float* G, const float* whh_t, float* OUT, float* CS, int n_seq, int T, int H, int steps_fwd, int steps_rev, cudaStream_t st);
int lstm_reg_bwd(float* G, const float* whh, const float* OUT, const float* CS, const float* dOUT, int n_seq, int T, int H, int steps_fwd,
                 int steps_rev, int zero_untaken, float* db_ih, float* db_hh, cudaStream_t st);
