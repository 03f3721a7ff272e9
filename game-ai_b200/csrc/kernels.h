// game-ai_b200/csrc/kernels.h -- launch wrappers shared between the kernel TUs and the C ABI.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define GAI_LOA_MAX_MOVES_K 96u

void launch_loa_movegen(const void* d_states, uint32_t n, uint8_t* d_counts, uint8_t* d_moves, cudaStream_t st);
void launch_loa_apply(const void* d_states, const uint8_t* d_mv, uint32_t n, void* d_out, cudaStream_t st);
void launch_loa_terminal(const void* d_states, uint32_t n, uint8_t* d_term, uint8_t* d_score, cudaStream_t st);
void launch_loa_successor_hash(const void* d_states, uint32_t n, uint64_t* d_hash, cudaStream_t st);
void launch_loa_rollout(const void* d_boards, const uint32_t* d_stm, uint32_t n_leaves, uint32_t ppl, uint32_t max_plies,
                        uint64_t key, uint64_t first_id, uint32_t* d_out, uint64_t* d_stats, unsigned long long* d_counter,
                        int threads, int blocks, cudaStream_t st);
void launch_loa_playout_trace(const void* d_states, uint32_t n, uint32_t max_plies, uint64_t key, uint64_t first_id,
                              uint8_t* d_outcome, uint32_t* d_plies, void* d_final, cudaStream_t st);
void launch_loa_gen_reachable(uint32_t n, uint64_t key, uint32_t max_k, void* d_boards, void* d_states24, uint8_t* d_side, cudaStream_t st);
void launch_loa_pack_stm(const uint8_t* d_side, uint32_t n, uint32_t* d_stm, cudaStream_t st);
void launch_loa_split_states(const void* d_states24, uint32_t n, void* d_boards, uint32_t* d_stm, cudaStream_t st);
// all successors of n parents (reference move order) into the rollout layout: boards + one side byte per child
void launch_loa_expand_children(const void* d_parents24, uint32_t n, const uint32_t* d_offsets, void* d_boards, uint8_t* d_side, uint32_t* d_bad, cudaStream_t st);

// "rotated boards + line LUT" formulation (loa_fast.cuh): one CTA per SM, LUT in shared memory
int  loa_fast_lut_bytes();
void loa_fast_build_lut_host(uint16_t* lut);
cudaError_t loa_fast_prepare();
void launch_loa_rollout_fast(const void* d_lut, const void* d_prep, const uint32_t* d_stm, const uint32_t* d_order, const uint16_t* d_chi, uint32_t n_leaves, uint32_t ppl, uint32_t max_plies,
                             uint64_t key, uint64_t first_id, uint32_t* d_out, uint64_t* d_stats, unsigned long long* d_counter,
                             int threads, int sms, int sweep, cudaStream_t st);
void launch_loa_movegen_sweep(const void* d_lut, const void* d_states, uint32_t n, uint8_t* d_counts, uint8_t* d_moves, int sms, cudaStream_t st);
// 16 B boards -> 64 B rotations + the descending-piece-count order (d_hist: 64 words, zeroed by the caller) + the Euler
// characteristic of both colours (d_chi: one u16 per leaf)
void launch_loa_prepare_leaves(const void* d_boards, uint32_t n, void* d_prep, uint8_t* d_pcount, uint32_t* d_hist, uint32_t* d_order, uint16_t* d_chi, cudaStream_t st);
void launch_loa_playout_trace_sweep(const void* d_lut, const void* d_states, uint32_t n, uint32_t max_plies, uint64_t key, uint64_t first_id,
                                    uint8_t* d_outcome, uint32_t* d_plies, void* d_final, int sms, cudaStream_t st);
void launch_loa_movegen_fast(const void* d_lut, const void* d_states, uint32_t n, uint8_t* d_counts, uint8_t* d_moves, int sms, cudaStream_t st);
void launch_loa_successor_hash_fast(const void* d_lut, const void* d_states, uint32_t n, uint64_t* d_hash, uint32_t* d_bad, int sms, cudaStream_t st);
void launch_loa_playout_trace_fast(const void* d_lut, const void* d_states, uint32_t n, uint32_t max_plies, uint64_t key, uint64_t first_id,
                                   uint8_t* d_outcome, uint32_t* d_plies, void* d_final, int sms, cudaStream_t st);

// Pan (determinized) -- pan_kernels.cu
void launch_pan_movegen(const void* d_states, uint32_t n, int np, uint8_t* d_counts, uint64_t* d_moves, uint32_t* d_bad, cudaStream_t st);
void launch_pan_apply(const void* d_states, const uint64_t* d_mv, uint32_t n, int np, void* d_out, uint32_t* d_bad, cudaStream_t st);
void launch_pan_terminal(const void* d_states, uint32_t n, int np, uint8_t* d_term, int* d_score, int* d_ev0, int* d_ev1, uint32_t* d_bad, cudaStream_t st);
void launch_pan_playout_trace(const void* d_states, uint32_t n, int np, uint32_t max_plies, int eval_kind, uint64_t key, uint64_t first_id,
                              uint8_t* d_code, uint32_t* d_plies, int* d_value, void* d_final, uint32_t* d_bad, cudaStream_t st);
// (also launches k_pan_prepare: 40-byte reference states -> 32-byte device records in d_prep, n_leaves x 32 B)
void launch_pan_rollout(const void* d_states, void* d_prep, uint32_t n_leaves, int np, uint32_t ppl, uint32_t max_plies, int eval_kind, uint64_t key, uint64_t first_id,
                        uint32_t* d_out, uint64_t* d_stats, unsigned long long* d_counter, uint32_t* d_bad, int blocks, cudaStream_t st);
