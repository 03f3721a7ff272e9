/* oracle/loa_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C11) of the Lines-of-Action rules of the reference,
 * c++/LinesOfActionZasady/LinesOfActionZasady.cpp, plus the Philox-driven random playout that
 * defines tier-2 truth.  It is the CHECKER for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product (libgameai_gpu.so) never links or calls it and has no CPU fallback.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function below against
 *   (a) the reference's own unit-test vectors (LinesOfActionZasady_ut.cpp:32-74,142-270,289-378),
 *       committed as tests/golden/loa_ut_vectors.json, and
 *   (b) outputs of the reference itself compiled here (oracle/_ref/libref_loa.so, built by
 *       `make -C oracle ref` from /root/reference), committed as tests/golden/loa_ref_corpus.bin
 *       by tests/golden/make_loa_golden.py, and re-checked live whenever oracle/_ref is present.
 */
#ifndef LOA_ORACLE_H
#define LOA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { LOA_WHITE = 0, LOA_BLACK = 1 };                        /* LinesOfActionZasady.cpp:40-41 */
enum { LOA_MAX_MOVES = 96 };                                  /* :226 assert(ml.size <= 12*8)  */
enum { LOA_OUT_WHITE = 0, LOA_OUT_BLACK = 1, LOA_OUT_DRAW = 2, LOA_OUT_CAPPED = 3 };

uint64_t loa_oracle_neighbour_mask(int pos);                  /* :59-81   */
uint64_t loa_oracle_row_mask(int pos);                        /* :82-86   */
uint64_t loa_oracle_column_mask(int pos);                     /* :87-92   */
uint64_t loa_oracle_left_diagonal_mask(int pos);              /* :93-113  */
uint64_t loa_oracle_right_diagonal_mask(int pos);             /* :114-135 */
int      loa_oracle_num_groups(uint64_t board);               /* :271-305 */
int      loa_oracle_calc_if_terminal(uint64_t board);         /* :52-58   */
int      loa_oracle_is_terminal(uint64_t white, uint64_t black);               /* :156-166, :416-419 */
void     loa_oracle_score(uint64_t white, uint64_t black, int score[2]);       /* :420-432 */
void     loa_oracle_initial(uint64_t* white, uint64_t* black, int* cp);        /* :358-365 */
/* moves of the side to move, reference order; out[2*i]=from, out[2*i+1]=to.  Returns count. */
int      loa_oracle_movegen(uint64_t white, uint64_t black, int cp, uint8_t* out);   /* :167-232, :136-155 */
void     loa_oracle_apply(uint64_t white, uint64_t black, int cp, int from, int to,
                          uint64_t* ow, uint64_t* ob, int* ocp);               /* :233-252, :481-489 */
int      loa_oracle_state_to_string(uint64_t white, uint64_t black, int cp, char* out, int cap); /* :253-270 */

int      loa_oracle_playout(uint64_t white, uint64_t black, int cp, uint64_t key, uint64_t playout_id,
                            uint32_t max_plies, uint32_t* plies_out, uint64_t* fw, uint64_t* fb);
uint64_t loa_oracle_playout_batch(const uint64_t* whites, const uint64_t* blacks, const uint8_t* cps,
                                  uint32_t n_leaves, uint32_t playouts_per_leaf, uint32_t max_plies,
                                  uint64_t key, uint64_t first_id, uint32_t* counts);
void     loa_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* batch forms; states = n x {white, black, current_player} uint64 triples */
void     loa_oracle_movegen_batch(const uint64_t* states, uint32_t n, uint8_t* counts, uint8_t* moves);
void     loa_oracle_terminal_batch(const uint64_t* states, uint32_t n, uint8_t* terminal, uint8_t* score);
void     loa_oracle_apply_batch(const uint64_t* states, const uint8_t* mv, uint32_t n, uint64_t* out);
void     loa_oracle_successor_hash_batch(const uint64_t* states, uint32_t n, uint64_t* hash);
void     loa_oracle_playout_trace_batch(const uint64_t* states, uint32_t n, uint32_t max_plies, uint64_t key, uint64_t first_id,
                                        uint8_t* outcome, uint32_t* plies, uint64_t* final_states);
void     loa_oracle_gen_reachable(uint32_t n, uint64_t key, uint32_t max_k, uint64_t* states);
void     loa_oracle_gen_uniform(uint32_t n, uint64_t key, uint64_t* states);
#ifdef __cplusplus
}
#endif
#endif
