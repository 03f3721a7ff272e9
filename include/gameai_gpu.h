/* include/gameai_gpu.h -- C ABI of libgameai_gpu.so, the B200 (sm_100a) rollout engine.
 *
 * This is the drop-in boundary for the simulation (random playout) phase of the reference
 * MC::Player (c++/MCTSPlayer/MCTSPlayer.cpp:387-470 selection_playOut) on the Lines-of-Action rules
 * (c++/LinesOfActionZasady/LinesOfActionZasady.cpp) and, for determinized rollouts, the Pan rules
 * (c++/GraWPanaZasadyV2/GraWPanaZasadyV2.cpp).  Nothing like it exists in the reference (no FFI, no
 * GPU code): the entry points below are what a GPU-backed IGamePlayer (GamePlayer.h:12-25) binds to
 * "launch a batch of leaf states, get per-leaf win counts back".  INTEGRATION.md shows the binding.
 *
 * Plain C: opaque handle, pointers + sizes, int error codes (0 = GAI_OK), no C++/torch types.
 * There is NO CPU fallback: without a CUDA device every compute entry point returns an error.
 *
 * Each entry point names the reference interface it replaces (path:line under /root/reference/c++).
 */
#ifndef GAMEAI_GPU_H
#define GAMEAI_GPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define GAI_ABI_VERSION 2

enum gai_status {
	GAI_OK = 0,
	GAI_ERR_INVALID = 1,     /* bad argument */
	GAI_ERR_CUDA = 2,        /* CUDA runtime error (gai_last_error has the text) */
	GAI_ERR_NO_DEVICE = 3,   /* no usable sm_100 device: there is no CPU fallback */
	GAI_ERR_NCCL = 4,
	GAI_ERR_NOMEM = 5
};

typedef struct gai_ctx gai_ctx;

/* LinesOfActionZasady.cpp:38-46 `struct GameState { uint64_t white, black; uint64_t current_player:1; }`
 * -- 24 bytes, bit i = square (row i/8, col i%8), bit 0 = A1; Whites = 0, Blacks = 1.  Only bit 0 of
 * current_player is read (the upper 63 bits are indeterminate in the reference, ctor :48-51), so a
 * reference GameState* array can be passed as is. */
typedef struct gai_loa_state { uint64_t white, black, current_player; } gai_loa_state;

/* LinesOfActionZasady.cpp:16-19 `struct Move { uint8_t from, to; }` */
typedef struct gai_loa_move { uint8_t from, to; } gai_loa_move;
#define GAI_LOA_MAX_MOVES 96            /* LinesOfActionZasady.cpp:226 */

/* Per-leaf playout outcome counts.  Replaces the per-simulation Score()/eval back-up of
 * MCTSPlayer.cpp:478-510: value[white] = (white_wins + 0.5*(draws+capped)) , likewise black
 * (LinesOfActionZasady.cpp:420-432 Score: both connected -> 50/50; :382-397 eval at the depth cap -> 50/50). */
typedef struct gai_loa_result { uint32_t white_wins, black_wins, draws, capped; } gai_loa_result;

typedef struct gai_rollout_stats {
	uint64_t playouts;      /* playouts finished */
	uint64_t plies;         /* moves applied over all playouts (passes included) */
	float    kernel_ms;     /* device time of the rollout kernel(s), CUDA events on the ctx stream */
	float    total_ms;      /* device time H2D + kernels + D2H (host-buffer entry points) */
} gai_rollout_stats;

typedef struct gai_device_info {
	int  device;
	int  sm_count;
	int  cc_major, cc_minor;
	int  clock_khz;          /* cudaDevAttrClockRate (max SM clock) */
	uint64_t total_mem;
	char name[64];
} gai_device_info;

/* ---- context ------------------------------------------------------------------------------ */
int  gai_abi_version(void);
/* create a context (one CUDA stream, pinned staging, device scratch) on `device`. */
int  gai_create(int device, gai_ctx** out);
void gai_destroy(gai_ctx* ctx);
/* text of the last error on ctx (or of the last gai_create failure when ctx == NULL). */
const char* gai_last_error(const gai_ctx* ctx);
int  gai_get_device_info(gai_ctx* ctx, gai_device_info* info);
int  gai_sync(gai_ctx* ctx);
/* rollout kernel launch shape: threads per block, blocks per SM (0 = keep default). */
int  gai_set_launch_config(gai_ctx* ctx, int threads_per_block, int blocks_per_sm);
/* launch shape the rollout entry points use for the selected rules formulation: threads per block and, for the static
 * sweep, the number of plies per loop trip (either pointer may be NULL). */
int  gai_get_launch_config(gai_ctx* ctx, int* threads_per_block, int* plies_per_trip);
/* Which device formulation of the rules the sweep and rollout entry points use.  Both give identical results
 * (tests run all three): 0 = per-piece 64-bit bitboard restatement (baseline), 1 = rotated boards + shared-memory
 * line LUT, one piece at a time, 2 = static sweep over the 32 line bytes with carry-save row counts (default; one CTA
 * of 896 threads per SM). */
int  gai_set_rules_impl(gai_ctx* ctx, int impl);
/* number of kernel launches issued through ctx since creation (bench.py's gpu_launches) */
uint64_t gai_launch_count(const gai_ctx* ctx);

/* ---- LOA rules sweep, HOST buffers (tier-1 parity; config 2) ------------------------------- */
/* IGameRules::GetPlayerLegalMoves (GameRules.h:30; LinesOfActionZasady.cpp:463-471 -> :167-232):
 * counts[i] = number of moves of the side to move, moves[i*96 + k] = k-th move in reference order. */
int gai_loa_movegen(gai_ctx* ctx, const gai_loa_state* states, uint32_t n, uint8_t* counts, gai_loa_move* moves);
/* IGameRules::ApplyMove (GameRules.h:38; LinesOfActionZasady.cpp:481-489 -> :233-252): out[i] = states[i]
 * after mv[i] by the side to move; from == to means the noop move (side flips only). */
int gai_loa_apply(gai_ctx* ctx, const gai_loa_state* states, const gai_loa_move* mv, uint32_t n, gai_loa_state* out);
/* IGameRules::IsTerminal + Score (GameRules.h:27-28; LinesOfActionZasady.cpp:416-432):
 * terminal[i] in {0,1}; score[2*i], score[2*i+1] = white, black in {0,50,100}. */
int gai_loa_terminal(gai_ctx* ctx, const gai_loa_state* states, uint32_t n, uint8_t* terminal, uint8_t* score);
/* all successors: for state i and k < counts[i], succ[i*96+k] = state after move k (checksummed by tests). */
int gai_loa_successor_hash(gai_ctx* ctx, const gai_loa_state* states, uint32_t n, uint64_t* hash);

/* ---- LOA rollouts (the hot path; configs 3, 4) ---------------------------------------------- */
/* HOST buffers, synchronous.  For leaf i and j < playouts_per_leaf runs the uniformly random playout
 * with id first_playout_id + i*playouts_per_leaf + j from leaves[i] until terminal or max_plies moves,
 * choosing move (philox(key,id,ply) * n) >> 32 of the reference-ordered list; out[i] accumulates outcomes.
 * Replaces the playout part of MC::Player::selection_playOut (MCTSPlayer.cpp:387-470) + the Score/eval
 * read in backpropagation (:478-493).  stats may be NULL. */
int gai_loa_rollout(gai_ctx* ctx, const gai_loa_state* leaves, uint32_t n_leaves, uint32_t playouts_per_leaf,
                    uint32_t max_plies, uint64_t philox_key, uint64_t first_playout_id,
                    gai_loa_result* out, gai_rollout_stats* stats);
/* Asynchronous variant: enqueues H2D, kernel and D2H on the ctx stream and returns; pageable `leaves` are copied
 * into pinned staging before return (a PINNED `leaves` buffer is DMA-ed in place and must stay untouched until
 * gai_sync), `out` is valid after gai_sync(ctx).  One batch in flight per ctx through this entry point; use the
 * ticketed calls below for more. */
int gai_loa_rollout_async(gai_ctx* ctx, const gai_loa_state* leaves, uint32_t n_leaves, uint32_t playouts_per_leaf,
                          uint32_t max_plies, uint64_t philox_key, uint64_t first_playout_id, gai_loa_result* out);

/* Statistics of the last rollout that completed on ctx (valid after gai_sync following gai_loa_rollout_async). */
int gai_last_rollout_stats(gai_ctx* ctx, gai_rollout_stats* stats);

/* ---- ticketed batches: several leaf batches queued on one context ---------------------------- */
/* The host player keeps the GPU fed by selecting batch k+1 (and k+2 ...) while batch k plays out: up to
 * GAI_MAX_INFLIGHT batches may be queued at once, each with its own staging and device buffers; they run in
 * submission order on the context's stream.  `leaves` is copied before the call returns (pageable memory) or must
 * stay untouched until gai_wait (pinned memory is DMA-ed in place); `out` is written by gai_wait(ticket) (pageable) or
 * by the DMA that gai_wait waits for (pinned).  A submit with all slots busy returns GAI_ERR_INVALID.
 * While any ticket is outstanding, the synchronous entry points of the same context (movegen / apply / terminal /
 * Pan / allreduce / rollout_device / rollout_async) refuse to run: they share the context's scratch buffers. */
#define GAI_MAX_INFLIGHT 4
int gai_loa_rollout_submit(gai_ctx* ctx, const gai_loa_state* leaves, uint32_t n_leaves, uint32_t playouts_per_leaf,
                           uint32_t max_plies, uint64_t philox_key, uint64_t first_playout_id, gai_loa_result* out, int* ticket);
/* Batch-by-expansion: for every parent position ALL its successors (IGameRules::GetPlayerLegalMoves order,
 * LinesOfActionZasady.cpp:167-232, generated and applied on the device) are played out playouts_per_leaf times each --
 * what MC::Player does one simulation at a time when it first tries every move of a fresh node (all UCB values tie,
 * MCTSPlayer.cpp:582-593).  child_offsets[n_parents + 1] (host): child k of parent i is leaf number child_offsets[i] + k,
 * its result goes to out[child_offsets[i] + k], its playouts have ids first_playout_id + (child_offsets[i] + k) *
 * playouts_per_leaf + j.  The caller computes the offsets from its own move lists; a parent whose device move count
 * differs from child_offsets[i+1] - child_offsets[i] fails the batch (gai_wait returns GAI_ERR_INVALID). */
int gai_loa_rollout_children_submit(gai_ctx* ctx, const gai_loa_state* parents, uint32_t n_parents, const uint32_t* child_offsets,
                                    uint32_t playouts_per_leaf, uint32_t max_plies, uint64_t philox_key, uint64_t first_playout_id,
                                    gai_loa_result* out, int* ticket);
/* Blocks until the batch behind `ticket` has finished and its results are in `out`; stats may be NULL. */
int gai_wait(gai_ctx* ctx, int ticket, gai_rollout_stats* stats);
/* number of tickets outstanding on ctx */
int gai_inflight(const gai_ctx* ctx);

/* DEVICE-resident variant (inputs already in HBM).  d_boards: n_leaves x {white,black} (16 B each,
 * 16-byte aligned); d_stm: ceil(n/32) words, bit (i&31) of word i>>5 = side to move of leaf i;
 * d_out: n_leaves x gai_loa_result (zeroed by the call); runs on the ctx stream, synchronises, fills stats. */
int gai_loa_rollout_device(gai_ctx* ctx, const void* d_boards, const uint32_t* d_stm, uint32_t n_leaves,
                           uint32_t playouts_per_leaf, uint32_t max_plies, uint64_t philox_key,
                           uint64_t first_playout_id, void* d_out, gai_rollout_stats* stats);

/* Single playouts with full detail (tier-2 parity): per playout i (ids first_playout_id + i) from
 * states[i]: outcome[i] in {0 white,1 black,2 draw,3 capped}, plies[i], final[i] = last position. */
int gai_loa_playout_trace(gai_ctx* ctx, const gai_loa_state* states, uint32_t n, uint32_t max_plies,
                          uint64_t philox_key, uint64_t first_playout_id,
                          uint8_t* outcome, uint32_t* plies, gai_loa_state* final_states);

/* ---- Pan ("Gra w Pana"), determinized rollouts (config 5) ------------------------------------ */
/* GraWPanaZasadyV2.cpp:22-40 `struct GameState` as raw words, 40 bytes: w[0] = stack:48 | current_player:3 << 48 |
 * is_terminal:1 << 51 ; w[1+p] = count:4 | cards:48 << 4 (Hand, :16-20).  Card c = rank*4 + suit at bits 2c..2c+1.
 * The GPU path plays DETERMINIZED deals: every card field must be 00 or 11 (complete information); a state with
 * fuzzy cards (01 / 10, CreatePlayerKnownState :214-243) is rejected with GAI_ERR_INVALID -- sample a deal first. */
typedef struct gai_pan_state { uint64_t w[5]; } gai_pan_state;
/* GraWPanaZasadyV2.cpp:42-49 `struct Move` raw: cards:48 | count:3 << 48 | operation:3 << 51 | probability:2 << 54 */
typedef struct gai_pan_move { uint64_t w; } gai_pan_move;
#define GAI_PAN_MAX_MOVES 16
#define GAI_PAN_EVAL_NUM_CARDS 0            /* CreateEvalFunction("num_cards"), :255-261 */
#define GAI_PAN_EVAL_NUM_CARDS_WEIGHTED 1   /* CreateEvalFunction("num_cards_weighted"), :285-300 */
/* Per-leaf playout outcomes.  Terminal playouts: losses[p] counts playouts that left player p holding cards (Score
 * :347-356 gives 100/(N-1) to everyone else).  Playouts stopped by the ply cap (MCTSConfig::MaxPlayoutDepth,
 * playout_depth=50 in run_config.xml) are evaluated like MC::Player does (MCTSPlayer.cpp:486-488): eval_sum[p] adds
 * the eval function's value for player p, capped counts them. */
typedef struct gai_pan_result { uint32_t losses[4]; uint32_t eval_sum[4]; uint32_t capped; uint32_t reserved[3]; } gai_pan_result;

/* IGameRules::GetPlayerLegalMoves (GraWPanaZasadyV2.cpp:409-475): moves[i*16 + k] = k-th move of the current player */
int gai_pan_movegen(gai_ctx* ctx, const gai_pan_state* states, uint32_t n, int num_players, uint8_t* counts, gai_pan_move* moves);
/* IGameRules::ApplyMove (:560-569 -> :534-557, next player :510-518, terminal flag :339-346) */
int gai_pan_apply(gai_ctx* ctx, const gai_pan_state* states, const gai_pan_move* mv, uint32_t n, int num_players, gai_pan_state* out);
/* checkIfTerminal + Score (:339-363) + both eval functions (:255-311): score/eval arrays are n x 4 int32 */
int gai_pan_terminal(gai_ctx* ctx, const gai_pan_state* states, uint32_t n, int num_players, uint8_t* terminal,
                     int32_t* score, int32_t* eval_num_cards, int32_t* eval_num_cards_weighted);
/* HOST buffers: per leaf (= one sampled deal) playouts_per_leaf uniformly random playouts (ids first_playout_id +
 * i*playouts_per_leaf + j) until one player is left holding cards or max_plies moves were made. */
int gai_pan_rollout(gai_ctx* ctx, const gai_pan_state* leaves, uint32_t n_leaves, int num_players, uint32_t playouts_per_leaf,
                    uint32_t max_plies, int eval_kind, uint64_t philox_key, uint64_t first_playout_id,
                    gai_pan_result* out, gai_rollout_stats* stats);
/* DEVICE-resident variant (inputs already in HBM): d_states = n_leaves x gai_pan_state (40 B each), d_out = n_leaves x
 * gai_pan_result (zeroed by the call); one pass on the ctx stream, synchronises, fills stats.  The host-buffer call above
 * moves the same data in pipeline stages of 65536 deals (upload / play out / download overlapped). */
int gai_pan_rollout_device(gai_ctx* ctx, const void* d_states, uint32_t n_leaves, int num_players, uint32_t playouts_per_leaf,
                           uint32_t max_plies, int eval_kind, uint64_t philox_key, uint64_t first_playout_id,
                           void* d_out, gai_rollout_stats* stats);
/* single playouts with detail: code[i] = loser 0..3, or 4 for a capped playout; value = Score / eval at the end */
int gai_pan_playout_trace(gai_ctx* ctx, const gai_pan_state* states, uint32_t n, int num_players, uint32_t max_plies, int eval_kind,
                          uint64_t philox_key, uint64_t first_playout_id, uint8_t* code, uint32_t* plies, int32_t* value,
                          gai_pan_state* final_states);

/* Determinization (HOST, no GPU; new behaviour -- the reference has no sampler, it searches the fuzzy state directly,
 * GraWPanaZasadyV2.cpp:214-243 + MCTSPlayer.h:55-69).  `known` is what CreatePlayerKnownState gives `player`: own
 * cards certain (11), every unseen card marked 11/10/01 in every other hand, hand sizes public.  Writes n_samples
 * complete-information states: the unseen cards dealt uniformly at random (Fisher-Yates on the Philox stream
 * (key, id = first_sample_id + i)) to the other players according to their public counts.  Sample i is a pure
 * function of (known, key, first_sample_id + i).  Returns GAI_ERR_INVALID if the counts cannot be met. */
int gai_pan_determinize(const gai_pan_state* known, int num_players, int player, uint32_t n_samples,
                        uint64_t philox_key, uint64_t first_sample_id, gai_pan_state* out);

/* ---- synthetic corpus -------------------------------------------------------- */
/* "reachable" positions: state i = the position after k_i random plies from the opening
 * (LinesOfActionZasady.cpp:358-365), k_i = philox(key^GEN, i, 0) % max_k, moves drawn from stream
 * (key, id = i); reaching a terminal position restarts from the opening.  Non-terminal by construction.
 * Written to DEVICE memory in the rollout layout (d_boards, d_stm) and/or to a host state array. */
int gai_loa_gen_reachable(gai_ctx* ctx, uint32_t n, uint64_t key, uint32_t max_k,
                          void* d_boards, uint32_t* d_stm, gai_loa_state* host_states);

/* ---- device memory helpers (so a caller without a CUDA runtime binding can stage buffers) --- */
int gai_dev_alloc(gai_ctx* ctx, uint64_t bytes, void** d_ptr);
int gai_dev_free(gai_ctx* ctx, void* d_ptr);
int gai_memcpy_h2d(gai_ctx* ctx, void* d_dst, const void* h_src, uint64_t bytes);
int gai_memcpy_d2h(gai_ctx* ctx, void* h_dst, const void* d_src, uint64_t bytes);
/* repack reference-layout states into the device rollout layout (host side, no GPU) */
int gai_loa_pack_states(const gai_loa_state* states, uint32_t n, void* boards16, uint32_t* stm_bits);

/* ---- choice stream (host evaluation, no GPU) ------------------------------------------------ */
/* Philox4x32-10 exactly as the kernels evaluate it (include/gameai/philox.h).  The reference draws from
 * std::default_random_engine (MCTSPlayer.cpp:472-476, implementation-defined); the counter-based stream is
 * what makes a playout a pure function of (key, playout id, ply). */
void gai_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* ---- root-parallel multi-GPU (config 4) ----------------------------------------------------- */
#define GAI_NCCL_UNIQUE_ID_BYTES 128
/* rank 0 calls gai_comm_unique_id and ships the 128 bytes to the other ranks (any side channel). */
int gai_comm_unique_id(uint8_t id[GAI_NCCL_UNIQUE_ID_BYTES]);
int gai_comm_init(gai_ctx* ctx, int rank, int world, const uint8_t id[GAI_NCCL_UNIQUE_ID_BYTES]);
/* Sum root-child statistics over all ranks, in place: `visits` u32[count], `values` f32[count*2]
 * (MoveNode::numVisited / value[], MCTSPlayer.h:47-69).  One call per move decision. */
int gai_allreduce_root(gai_ctx* ctx, uint32_t* visits, float* values, uint32_t count);
/* Sum u64 counters over all ranks, in place (throughput bench). */
int gai_allreduce_u64(gai_ctx* ctx, uint64_t* v, uint32_t count);

#ifdef __cplusplus
}
#endif
#endif
