// Register-resident / cluster LSTM recurrence (lstm_reg.cu); dispatched from the C-ABI wrappers in lstm.cu.
#pragma once
#include <cuda_runtime.h>
bool lstm_reg_supported(int H);
int lstm_reg_cluster(int H);   // CTAs per cluster (1 or 2) == how many unit ranges the forward weights are permuted for
// xdrop != nullptr: also write dropout_p(OUT) (the next layer's input) with the Philox stream (seed, salt [+ 64 * *step_dev]);
// backward p > 0: dOUT is the gradient w.r.t. that dropped-out copy, the same mask is applied on load
int lstm_reg_fwd(float* G, const float* whh_t, float* OUT, float* CS, int n_seq, int T, int H, int steps_fwd, int steps_rev, float* xdrop,
                 float p, unsigned long long seed, unsigned salt, const int* step_dev, cudaStream_t st);
// dOUT_add (optional): a second gradient buffer of dOUT's layout of which only the rows t = T-1 are written; added on load
int lstm_reg_bwd(float* G, const float* whh, const float* OUT, const float* CS, const float* dOUT, const float* dOUT_add, int n_seq, int T, int H, int steps_fwd,
                 int steps_rev, int zero_untaken, float* db_ih, float* db_hh, float p, unsigned long long seed, unsigned salt,
                 const int* step_dev, cudaStream_t st);
