// game-ai_b200/csrc/pan_kernels.cu -- sm_100a kernels for determinized Pan rollouts and the Pan rules sweep.
#include "pan_rules.cuh"
#include "gameai/philox.h"
#include "kernels.h"

namespace pan {

__device__ __forceinline__ u32 pick(const gai::Philox4& r, u32 ply)
{
	const u32 w = ply & 3u;
	return w == 0 ? r.v[0] : w == 1 ? r.v[1] : w == 2 ? r.v[2] : r.v[3];
}

// one playout; returns 0..3 = loser, 4 = ply cap / dead end (value[] = eval), also the number of plies
__device__ __forceinline__ int playout(State& s, int np, u64 key, u64 id, u32 max_plies, u32* plies_out)
{
	gai::Philox4 rnd; u32 rq = 0xffffffffu, ply = 0;
	int code = 4;
	for (;;) {
		if (is_terminal(s, np)) { for (int p = 0; p < np; ++p) if (count_of(s, p) != 0) code = p; break; }
		if (ply >= max_plies) break;
		const Moves m = gen_moves(s);
		if (m.n == 0) break;
		if ((ply >> 2) != rq) { rq = ply >> 2; rnd = gai::choice_block(key, id, rq); }
		u32 cards, count, op;
		nth_move(m, gai::choice_index(pick(rnd, ply), m.n), &cards, &count, &op);
		apply(s, np, cards, count, op);
		++ply;
	}
	*plies_out = ply;
	return code;
}

__global__ void k_pan_movegen(const u64* __restrict__ states, u32 n, int np, uint8_t* __restrict__ counts, u64* __restrict__ moves, u32* __restrict__ bad)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	State s;
	if (!load_state(s, states + 5 * (size_t)i)) { atomicAdd(bad, 1u); counts[i] = 0; return; }
	const bool term = (states[5 * (size_t)i] >> 51) & 1ull;          // GetPlayerLegalMoves reads the state's own flag (:411)
	Moves m = gen_moves(s);
	if (term) m.n = 0;
	counts[i] = (uint8_t)m.n;
	for (u32 k = 0; k < m.n && k < 16; ++k) {
		u32 cards, count, op;
		nth_move(m, k, &cards, &count, &op);
		moves[16 * (size_t)i + k] = move_word(cards, count, op);
	}
}

__global__ void k_pan_apply(const u64* __restrict__ states, const u64* __restrict__ mv, u32 n, int np, u64* __restrict__ out, u32* __restrict__ bad)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	State s;
	const u64 w = mv[i];
	if (!load_state(s, states + 5 * (size_t)i) || !is_crisp(w & CARDS48)) { atomicAdd(bad, 1u); return; }
	apply(s, np, squeeze(w & CARDS48), (u32)((w >> 48) & 7u), (u32)((w >> 51) & 7u));
	store_state(out + 5 * (size_t)i, s, np);
}

__global__ void k_pan_terminal(const u64* __restrict__ states, u32 n, int np, uint8_t* __restrict__ terminal, int* __restrict__ sc,
                               int* __restrict__ ev0, int* __restrict__ ev1, u32* __restrict__ bad)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	State s;
	if (!load_state(s, states + 5 * (size_t)i)) { atomicAdd(bad, 1u); return; }
	terminal[i] = is_terminal(s, np) ? 1 : 0;
	int v[4] = { 0, 0, 0, 0 };
	score(s, np, v);
	for (int p = 0; p < 4; ++p) {
		sc[4 * (size_t)i + p] = p < np ? v[p] : 0;
		ev0[4 * (size_t)i + p] = p < np ? eval_one(s, 0, p) : 0;
		ev1[4 * (size_t)i + p] = p < np ? eval_one(s, 1, p) : 0;
	}
}

__global__ void k_pan_playout_trace(const u64* __restrict__ states, u32 n, int np, u32 max_plies, int eval_kind, u64 key, u64 first_id,
                                    uint8_t* __restrict__ code, u32* __restrict__ plies, int* __restrict__ value, u64* __restrict__ final_states,
                                    u32* __restrict__ bad)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	State s;
	if (!load_state(s, states + 5 * (size_t)i)) { atomicAdd(bad, 1u); return; }
	u32 pl;
	const int c = playout(s, np, key, first_id + i, max_plies, &pl);
	code[i] = (uint8_t)c; plies[i] = pl;
	int v[4] = { 0, 0, 0, 0 };
	if (c < 4) score(s, np, v); else for (int p = 0; p < np; ++p) v[p] = eval_one(s, eval_kind, p);
	for (int p = 0; p < 4; ++p) value[4 * (size_t)i + p] = v[p];
	store_state(final_states + 5 * (size_t)i, s, np);
}

// Persistent rollout kernel: one playout per thread, lanes pull (leaf, playout) items from a global counter.
//   out: n_leaves x 12 words: [0..3] losses by player, [4..7] eval sums of capped playouts, [8] capped count
__global__ void __launch_bounds__(256)
k_pan_rollout(const u64* __restrict__ states, u32 n_leaves, int np, u32 ppl, u32 max_plies, int eval_kind, u64 key, u64 first_id,
              u32* __restrict__ out, u64* __restrict__ stats, unsigned long long* __restrict__ work_counter, u32* __restrict__ bad)
{
	const u64 total = (u64)n_leaves * ppl;
	const u32 lane = threadIdx.x & 31u;
	u64 my_plies = 0, my_playouts = 0;
	for (;;) {
		const unsigned grp = __activemask();
		const int leader = __ffs(grp) - 1;
		unsigned long long base = 0;
		if ((int)lane == leader) base = atomicAdd(work_counter, (unsigned long long)__popc(grp));
		base = __shfl_sync(grp, base, leader);
		const u64 item = base + __popc(grp & ((1u << lane) - 1u));
		if (item >= total) break;
		const u32 leaf = (u32)(item / ppl);
		State s;
		if (!load_state(s, states + 5 * (size_t)leaf)) { atomicAdd(bad, 1u); continue; }
		u32 pl;
		const int c = playout(s, np, key, first_id + item, max_plies, &pl);
		u32* r = out + 12 * (size_t)leaf;
		if (c < 4) atomicAdd(r + c, 1u);
		else {
			atomicAdd(r + 8, 1u);
			for (int p = 0; p < np; ++p) atomicAdd(r + 4 + p, (u32)eval_one(s, eval_kind, p));
		}
		my_plies += pl; my_playouts += 1;
	}
	__syncwarp();
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		my_plies += __shfl_xor_sync(0xffffffffu, my_plies, d);
		my_playouts += __shfl_xor_sync(0xffffffffu, my_playouts, d);
	}
	if (lane == 0) { atomicAdd((unsigned long long*)stats, my_playouts); atomicAdd((unsigned long long*)stats + 1, my_plies); }
}

} // namespace pan

using namespace pan;
static inline unsigned pblk(uint32_t n, int t) { return (unsigned)((n + t - 1) / t); }
void launch_pan_movegen(const void* d_states, uint32_t n, int np, uint8_t* d_counts, uint64_t* d_moves, uint32_t* d_bad, cudaStream_t st)
{ if (n) k_pan_movegen<<<pblk(n, 128), 128, 0, st>>>((const u64*)d_states, n, np, d_counts, (u64*)d_moves, d_bad); }
void launch_pan_apply(const void* d_states, const uint64_t* d_mv, uint32_t n, int np, void* d_out, uint32_t* d_bad, cudaStream_t st)
{ if (n) k_pan_apply<<<pblk(n, 128), 128, 0, st>>>((const u64*)d_states, (const u64*)d_mv, n, np, (u64*)d_out, d_bad); }
void launch_pan_terminal(const void* d_states, uint32_t n, int np, uint8_t* d_term, int* d_score, int* d_ev0, int* d_ev1, uint32_t* d_bad, cudaStream_t st)
{ if (n) k_pan_terminal<<<pblk(n, 128), 128, 0, st>>>((const u64*)d_states, n, np, d_term, d_score, d_ev0, d_ev1, d_bad); }
void launch_pan_playout_trace(const void* d_states, uint32_t n, int np, uint32_t max_plies, int eval_kind, uint64_t key, uint64_t first_id,
                              uint8_t* d_code, uint32_t* d_plies, int* d_value, void* d_final, uint32_t* d_bad, cudaStream_t st)
{ if (n) k_pan_playout_trace<<<pblk(n, 128), 128, 0, st>>>((const u64*)d_states, n, np, max_plies, eval_kind, key, first_id, d_code, d_plies, d_value, (u64*)d_final, d_bad); }
void launch_pan_rollout(const void* d_states, uint32_t n_leaves, int np, uint32_t ppl, uint32_t max_plies, int eval_kind, uint64_t key, uint64_t first_id,
                        uint32_t* d_out, uint64_t* d_stats, unsigned long long* d_counter, uint32_t* d_bad, int blocks, cudaStream_t st)
{ k_pan_rollout<<<blocks, 256, 0, st>>>((const u64*)d_states, n_leaves, np, ppl, max_plies, eval_kind, key, first_id, d_out, (u64*)d_stats, d_counter, d_bad); }
