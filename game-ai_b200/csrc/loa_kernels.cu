// game-ai_b200/csrc/loa_kernels.cu -- sm_100a kernels for the LOA rules sweep and rollouts.
// Hand-written CUDA; integer/branch bound (no tensor cores: nothing here is a contraction).
#include "loa_sim.cuh"
#include "kernels.h"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace loa {

__constant__ DiagTables c_diag = make_diag_tables();

__device__ __forceinline__ void load_tables(u64* s_ld, u64* s_rd)
{
	for (int i = threadIdx.x; i < 64; i += blockDim.x) { s_ld[i] = c_diag.ld[i]; s_rd[i] = c_diag.rd[i]; }
	__syncthreads();
}

__device__ __forceinline__ u64 mix64(u64 x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}

// ---------------------------------------------------------------------------------------------
// tier-1 sweep kernels: one position per thread, AoS 24-byte states as the reference lays them out
// ---------------------------------------------------------------------------------------------
__global__ void k_movegen(const u64* __restrict__ states, u32 n, uint8_t* __restrict__ counts, uint8_t* __restrict__ moves)
{
	__shared__ u64 s_ld[64], s_rd[64];
	load_tables(s_ld, s_rd);
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u64 w = states[3 * (size_t)i], b = states[3 * (size_t)i + 1];
	const u32 side = (u32)states[3 * (size_t)i + 2] & 1u;
	const u64 my = side ? b : w, en = side ? w : b;
	const MoveMasks m = all_piece_moves(my, en, s_ld, s_rd);
	const u32 cnt = (u32)count_moves(m);
	counts[i] = (uint8_t)cnt;
	uint8_t* out = moves + (size_t)i * 2 * GAI_LOA_MAX_MOVES_K;
	for (u32 k = 0; k < cnt && k < GAI_LOA_MAX_MOVES_K; ++k) {
		int from, to;
		nth_move(m, my, en, k, s_ld, s_rd, &from, &to);
		out[2 * k] = (uint8_t)from; out[2 * k + 1] = (uint8_t)to;
	}
}

__global__ void k_apply(const u64* __restrict__ states, const uint8_t* __restrict__ mv, u32 n, u64* __restrict__ out)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	u64 w = states[3 * (size_t)i], b = states[3 * (size_t)i + 1];
	const u32 side = (u32)states[3 * (size_t)i + 2] & 1u;
	const int from = mv[2 * (size_t)i], to = mv[2 * (size_t)i + 1];
	if (from != to) {                       // from == to : the noop move, side flips only (:484-487)
		if (side) apply_move(b, w, from, to); else apply_move(w, b, from, to);
	}
	out[3 * (size_t)i] = w; out[3 * (size_t)i + 1] = b; out[3 * (size_t)i + 2] = side ^ 1u;
}

__global__ void k_terminal(const u64* __restrict__ states, u32 n, uint8_t* __restrict__ terminal, uint8_t* __restrict__ score)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u64 w = states[3 * (size_t)i], b = states[3 * (size_t)i + 1];
	const bool wc = connected_fast(w), bc = connected_fast(b);
	terminal[i] = (wc || bc) ? 1 : 0;
	score[2 * (size_t)i]     = (wc && bc) ? 50 : wc ? 100 : 0;      // Score :420-432
	score[2 * (size_t)i + 1] = (wc && bc) ? 50 : bc ? 100 : 0;
}

// hash over all successors in move order (tier-1 "every successor" check without shipping 96 states)
__global__ void k_successor_hash(const u64* __restrict__ states, u32 n, u64* __restrict__ hash)
{
	__shared__ u64 s_ld[64], s_rd[64];
	load_tables(s_ld, s_rd);
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u64 w = states[3 * (size_t)i], b = states[3 * (size_t)i + 1];
	const u32 side = (u32)states[3 * (size_t)i + 2] & 1u;
	const u64 my = side ? b : w, en = side ? w : b;
	const MoveMasks m = all_piece_moves(my, en, s_ld, s_rd);
	const u32 cnt = (u32)count_moves(m);
	u64 h = 0;
	for (u32 k = 0; k < cnt; ++k) {
		int from, to;
		nth_move(m, my, en, k, s_ld, s_rd, &from, &to);
		u64 nmy = my, nen = en;
		apply_move(nmy, nen, from, to);
		const u64 nw = side ? nen : nmy, nb = side ? nmy : nen;
		const bool wc = connected_fast(nw), bc = connected_fast(nb);
		h = mix64(h ^ nw); h = mix64(h ^ nb);
		h = mix64(h ^ ((u64)(side ^ 1u) | ((u64)k << 8) | ((u64)(wc ? 1 : 0) << 16) | ((u64)(bc ? 1 : 0) << 17)));
	}
	hash[i] = h;
}

// ---------------------------------------------------------------------------------------------
// rollouts
// ---------------------------------------------------------------------------------------------
// Persistent kernel: one playout per thread; a lane that finishes pulls the next (leaf, playout) item from
// a global counter (warp-aggregated), so lanes never idle while work remains.  Playout identity is the
// item number -> Philox counter, never the lane, so results are independent of scheduling.
//   boards : n_leaves x {white, black}, one 16-byte vector load per item
//   stm    : side-to-move bit array
//   out    : n_leaves x {white_wins, black_wins, draws, capped} (pre-zeroed), accumulated with RED.ADD
//   stats  : [0] playouts, [1] plies  (warp-shuffle reduced, one atomic per warp)
__global__ void __launch_bounds__(256)
k_rollout(const ulonglong2* __restrict__ boards, const u32* __restrict__ stm, u32 n_leaves, u32 ppl,
          u32 max_plies, u64 key, u64 first_id, u32* __restrict__ out, u64* __restrict__ stats,
          unsigned long long* __restrict__ work_counter)
{
	__shared__ u64 s_ld[64], s_rd[64];
	load_tables(s_ld, s_rd);

	const u64 total = (u64)n_leaves * ppl;
	const u32 lane = threadIdx.x & 31u;
	Sim s; gai::Philox4 rnd;
	u64 item = 0; u32 leaf = 0;
	bool need = true;
	u64 my_plies = 0, my_playouts = 0;
	s.my = s.en = 0; s.side = 0; s.ply = 0; s.chk_en = s.chk_my = false; s.rq = 0;

	for (;;) {
		if (need) {
			// (coalesced group: well-defined membership under independent thread scheduling, unlike __activemask() + __shfl_sync)
			const cg::coalesced_group grp = cg::coalesced_threads();
			unsigned long long base = 0;
			if (grp.thread_rank() == 0) base = atomicAdd(work_counter, (unsigned long long)grp.size());
			base = grp.shfl(base, 0);
			item = base + grp.thread_rank();
			if (item >= total) break;
			leaf = (u32)(item / ppl);
			const ulonglong2 bd = __ldg(boards + leaf);           // LDG.128
			const u32 side = (__ldg(stm + (leaf >> 5)) >> (leaf & 31u)) & 1u;
			sim_init(s, bd.x, bd.y, side);
			need = false;
		}
		const int o = sim_step(s, rnd, key, first_id + item, max_plies, s_ld, s_rd);
		if (o >= 0) {
			atomicAdd(out + 4 * (size_t)leaf + o, 1u);
			my_plies += s.ply; my_playouts += 1;
			need = true;
		}
	}
	// per-warp reduction of the counters with shuffles, one atomic per warp
	// (lanes leave the loop at different times; reconverge here)
	__syncwarp();
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		my_plies += __shfl_xor_sync(0xffffffffu, my_plies, d);
		my_playouts += __shfl_xor_sync(0xffffffffu, my_playouts, d);
	}
	if (lane == 0) { atomicAdd((unsigned long long*)stats, my_playouts); atomicAdd((unsigned long long*)stats + 1, my_plies); }
}

// one playout per thread with full detail (tier-2 parity)
__global__ void k_playout_trace(const u64* __restrict__ states, u32 n, u32 max_plies, u64 key, u64 first_id,
                                uint8_t* __restrict__ outcome, u32* __restrict__ plies, u64* __restrict__ final_states)
{
	__shared__ u64 s_ld[64], s_rd[64];
	load_tables(s_ld, s_rd);
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Sim s; gai::Philox4 rnd;
	sim_init(s, states[3 * (size_t)i], states[3 * (size_t)i + 1], (u32)states[3 * (size_t)i + 2] & 1u);
	int o;
	do { o = sim_step(s, rnd, key, first_id + i, max_plies, s_ld, s_rd); } while (o < 0);
	outcome[i] = (uint8_t)o;
	plies[i] = s.ply;
	final_states[3 * (size_t)i] = sim_white(s);
	final_states[3 * (size_t)i + 1] = sim_black(s);
	final_states[3 * (size_t)i + 2] = s.side;
}

// synthetic "reachable" corpus (see gameai_gpu.h gai_loa_gen_reachable)
__global__ void k_gen_reachable(u32 n, u64 key, u32 max_k, ulonglong2* __restrict__ boards,
                                u64* __restrict__ states24, uint8_t* __restrict__ side_out)
{
	__shared__ u64 s_ld[64], s_rd[64];
	load_tables(s_ld, s_rd);
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const gai::Philox4 kk = gai::choice_block(key ^ 0x9E3779B97F4A7C15ull, i, 0);
	const u32 k = gai::choice_index(kk.v[0], max_k);
	u64 my = OPEN_BLACK, en = OPEN_WHITE; u32 side = 1;
	gai::Philox4 rnd; u32 rq = 0xffffffffu;
	for (u32 t = 0; t < k; ++t) {
		if (connected_fast(my) || connected_fast(en)) { my = OPEN_BLACK; en = OPEN_WHITE; side = 1; }
		const MoveMasks m = all_piece_moves(my, en, s_ld, s_rd);
		const u32 cnt = (u32)count_moves(m);
		if (cnt) {
			if ((t >> 2) != rq) { rq = t >> 2; rnd = gai::choice_block(key, i, rq); }
			int from, to;
			nth_move(m, my, en, gai::choice_index(pick_word(rnd, t), cnt), s_ld, s_rd, &from, &to);
			apply_move(my, en, from, to);
		}
		const u64 tmp = my; my = en; en = tmp; side ^= 1u;
	}
	if (connected_fast(my) || connected_fast(en)) { my = OPEN_BLACK; en = OPEN_WHITE; side = 1; }
	const u64 w = side ? en : my, b = side ? my : en;
	if (boards) boards[i] = make_ulonglong2(w, b);
	if (side_out) side_out[i] = (uint8_t)side;
	if (states24) { states24[3 * (size_t)i] = w; states24[3 * (size_t)i + 1] = b; states24[3 * (size_t)i + 2] = side; }
}

// pack per-leaf side bytes into the bit array the rollout kernel reads (one word per warp via ballot)
__global__ void k_pack_stm(const uint8_t* __restrict__ side, u32 n, u32* __restrict__ stm)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	const u32 bit = (i < n) ? (side[i] & 1u) : 0u;
	const u32 word = __ballot_sync(0xffffffffu, bit);
	if ((threadIdx.x & 31u) == 0 && i < n) stm[i >> 5] = word;
}

// split AoS 24-byte states into the rollout layout on the device
__global__ void k_split_states(const u64* __restrict__ states24, u32 n, ulonglong2* __restrict__ boards, u32* __restrict__ stm)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	u32 bit = 0;
	if (i < n) {
		boards[i] = make_ulonglong2(states24[3 * (size_t)i], states24[3 * (size_t)i + 1]);
		bit = (u32)states24[3 * (size_t)i + 2] & 1u;
	}
	const u32 word = __ballot_sync(0xffffffffu, bit);
	if ((threadIdx.x & 31u) == 0 && i < n) stm[i >> 5] = word;
}

// Batch-by-expansion: every successor of every parent, in reference move order, becomes a leaf.  One parent per thread
// (a batch has ~1e3 parents, the rollout that follows ~1e6 playouts: this kernel is not on the critical path).
// child k of parent i -> leaf offsets[i] + k; a parent whose move count differs from what the host announced sets *bad.
__global__ void k_expand_children(const u64* __restrict__ parents24, u32 n, const u32* __restrict__ offsets,
                                  ulonglong2* __restrict__ boards, uint8_t* __restrict__ side_out, u32* __restrict__ bad)
{
	__shared__ u64 s_ld[64], s_rd[64];
	load_tables(s_ld, s_rd);
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u64 w = parents24[3 * (size_t)i], b = parents24[3 * (size_t)i + 1];
	const u32 side = (u32)parents24[3 * (size_t)i + 2] & 1u;
	const u64 my = side ? b : w, en = side ? w : b;
	const MoveMasks m = all_piece_moves(my, en, s_ld, s_rd);
	const u32 cnt = (u32)count_moves(m);
	const u32 o0 = offsets[i], want = offsets[i + 1] - o0;
	if (cnt != want) {                          // caller and device disagree: flag the batch, hand the rollout harmless (terminal) leaves
		atomicOr(bad, 1u);
		for (u32 k = 0; k < want; ++k) { boards[o0 + k] = make_ulonglong2(1ull, 2ull); side_out[o0 + k] = 0; }
		return;
	}
	for (u32 k = 0; k < cnt; ++k) {
		int from, to;
		nth_move(m, my, en, k, s_ld, s_rd, &from, &to);
		u64 nmy = my, nen = en;
		apply_move(nmy, nen, from, to);
		boards[o0 + k] = side ? make_ulonglong2(nen, nmy) : make_ulonglong2(nmy, nen);
		side_out[o0 + k] = (uint8_t)(side ^ 1u);
	}
}

} // namespace loa

// ---------------------------------------------------------------------------------------------
// launch wrappers (called from the C ABI in gameai_gpu.cu)
// ---------------------------------------------------------------------------------------------
using namespace loa;
static inline unsigned nblk(uint32_t n, int t) { return (unsigned)((n + t - 1) / t); }

void launch_loa_movegen(const void* d_states, uint32_t n, uint8_t* d_counts, uint8_t* d_moves, cudaStream_t st)
{ if (n) k_movegen<<<nblk(n, 128), 128, 0, st>>>((const u64*)d_states, n, d_counts, d_moves); }
void launch_loa_apply(const void* d_states, const uint8_t* d_mv, uint32_t n, void* d_out, cudaStream_t st)
{ if (n) k_apply<<<nblk(n, 256), 256, 0, st>>>((const u64*)d_states, d_mv, n, (u64*)d_out); }
void launch_loa_terminal(const void* d_states, uint32_t n, uint8_t* d_term, uint8_t* d_score, cudaStream_t st)
{ if (n) k_terminal<<<nblk(n, 256), 256, 0, st>>>((const u64*)d_states, n, d_term, d_score); }
void launch_loa_successor_hash(const void* d_states, uint32_t n, uint64_t* d_hash, cudaStream_t st)
{ if (n) k_successor_hash<<<nblk(n, 128), 128, 0, st>>>((const u64*)d_states, n, (u64*)d_hash); }
void launch_loa_rollout(const void* d_boards, const uint32_t* d_stm, uint32_t n_leaves, uint32_t ppl, uint32_t max_plies,
                        uint64_t key, uint64_t first_id, uint32_t* d_out, uint64_t* d_stats, unsigned long long* d_counter,
                        int threads, int blocks, cudaStream_t st)
{
	k_rollout<<<blocks, threads, 0, st>>>((const ulonglong2*)d_boards, d_stm, n_leaves, ppl, max_plies, key, first_id,
	                                      d_out, (u64*)d_stats, d_counter);
}
void launch_loa_playout_trace(const void* d_states, uint32_t n, uint32_t max_plies, uint64_t key, uint64_t first_id,
                              uint8_t* d_outcome, uint32_t* d_plies, void* d_final, cudaStream_t st)
{ if (n) k_playout_trace<<<nblk(n, 128), 128, 0, st>>>((const u64*)d_states, n, max_plies, key, first_id, d_outcome, d_plies, (u64*)d_final); }
void launch_loa_gen_reachable(uint32_t n, uint64_t key, uint32_t max_k, void* d_boards, void* d_states24, uint8_t* d_side, cudaStream_t st)
{ if (n) k_gen_reachable<<<nblk(n, 128), 128, 0, st>>>(n, key, max_k, (ulonglong2*)d_boards, (u64*)d_states24, d_side); }
void launch_loa_pack_stm(const uint8_t* d_side, uint32_t n, uint32_t* d_stm, cudaStream_t st)
{ if (n) k_pack_stm<<<nblk(n, 256), 256, 0, st>>>(d_side, n, d_stm); }
void launch_loa_split_states(const void* d_states24, uint32_t n, void* d_boards, uint32_t* d_stm, cudaStream_t st)
{ if (n) k_split_states<<<nblk(n, 256), 256, 0, st>>>((const u64*)d_states24, n, (ulonglong2*)d_boards, d_stm); }
void launch_loa_expand_children(const void* d_parents24, uint32_t n, const uint32_t* d_offsets, void* d_boards, uint8_t* d_side, uint32_t* d_bad, cudaStream_t st)
{ if (n) k_expand_children<<<nblk(n, 64), 64, 0, st>>>((const u64*)d_parents24, n, d_offsets, (ulonglong2*)d_boards, d_side, d_bad); }
