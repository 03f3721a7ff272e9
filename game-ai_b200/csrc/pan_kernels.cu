// game-ai_b200/csrc/pan_kernels.cu -- sm_100a kernels for determinized Pan rollouts and the Pan rules sweep.
#include "pan_rules.cuh"
#include "gameai/philox.h"
#include "kernels.h"
#include <stdlib.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

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

// Leaf preparation: the reference's 40-byte GameState (2 bits per card) -> the 32-byte device record (1 bit per card),
// once per leaf, so that a lane that pulls a playout issues two 16-byte loads instead of five 64-bit squeezes
// (280 instructions that every lane of the warp had to wait for at each refill).
//   record: {stack, hand0, hand1, hand2} {hand3, counts, cp, ok}
__global__ void k_pan_prepare(const u64* __restrict__ states, u32 n, uint4* __restrict__ prep, u32* __restrict__ bad)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	State s;
	const bool ok = load_state(s, states + 5 * (size_t)i);
	if (!ok) atomicAdd(bad, 1u);                                   // fuzzy (non-determinized) cards: refused
	prep[2 * (size_t)i] = make_uint4(s.stack, s.hand[0], s.hand[1], s.hand[2]);
	prep[2 * (size_t)i + 1] = make_uint4(s.hand[3], s.counts, s.cp, ok ? 1u : 0u);
}
__device__ __forceinline__ bool load_prepared(State& s, const uint4* __restrict__ prep, u32 leaf)
{
	const uint4 a = __ldg(prep + 2 * (size_t)leaf), b = __ldg(prep + 2 * (size_t)leaf + 1);
	s.stack = a.x; s.hand[0] = a.y; s.hand[1] = a.z; s.hand[2] = a.w; s.hand[3] = b.x; s.counts = b.y; s.cp = b.z;
	return b.w != 0;
}

// Lane-independent variant (the first Pan kernel): a lane plays a whole playout before it pulls the next item.  Used when
// the ply cap bounds the divergence (see launch_pan_rollout); the warp-synchronous kernel below is the general one.
//   out: n_leaves x 12 words: [0..3] losses by player, [4..7] eval sums of capped playouts, [8] capped count
__global__ void __launch_bounds__(256)
k_pan_rollout_indep(const uint4* __restrict__ prep, u32 n_leaves, int np, u32 ppl, u32 max_plies, int eval_kind, u64 key, u64 first_id,
              u32* __restrict__ out, u64* __restrict__ stats, unsigned long long* __restrict__ work_counter, u32* __restrict__ bad)
{
	const u64 total = (u64)n_leaves * ppl;
	const u32 lane = threadIdx.x & 31u;
	u64 my_plies = 0, my_playouts = 0;
	for (;;) {
		// the lanes that arrive here together form one coalesced group (cooperative groups give it a well-defined membership
		// under independent thread scheduling, which __activemask() + __shfl_sync do not): one atomic per group
		const cg::coalesced_group grp = cg::coalesced_threads();
		unsigned long long base = 0;
		if (grp.thread_rank() == 0) base = atomicAdd(work_counter, (unsigned long long)grp.size());
		base = grp.shfl(base, 0);
		const u64 item = base + grp.thread_rank();
		if (item >= total) break;
		const u32 leaf = (u32)(item / ppl);
		State s;
		if (!load_prepared(s, prep, leaf)) continue;               // refused leaf (counted by k_pan_prepare)
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

// Persistent rollout kernel, warp-synchronous like the LOA one: every loop iteration each lane plays ONE ply of its own
// playout; a lane whose playout ended pulls the next (leaf, playout) item from a global counter instead of idling until
// the longest playout of its warp is over (uncapped 3-player games: 148 plies on average with a long tail).  Lanes start
// playouts only on iterations that are multiples of four, so ply & 3 is the same in every lane and the Philox block is
// refreshed by all lanes together (one uniform branch instead of a divergent 60-instruction block per lane).
//   out: n_leaves x 12 words: [0..3] losses by player, [4..7] eval sums of capped playouts, [8] capped count
__global__ void __launch_bounds__(256)
k_pan_rollout(const uint4* __restrict__ prep, u32 n_leaves, int np, u32 ppl, u32 max_plies, int eval_kind, u64 key, u64 first_id,
              u32* __restrict__ out, u64* __restrict__ stats, unsigned long long* __restrict__ work_counter, u32* __restrict__ bad)
{
	constexpr unsigned FULL = 0xffffffffu;
	const u64 total = (u64)n_leaves * ppl;
	const u32 lane = threadIdx.x & 31u;
	const bool ppl_pow2 = (ppl & (ppl - 1u)) == 0;
	const u32 ppl_shift = (u32)__ffs((int)ppl) - 1u;
	u64 my_plies = 0, my_playouts = 0, item = 0;
	State s; s.stack = 0; s.hand[0] = s.hand[1] = s.hand[2] = s.hand[3] = 0; s.counts = 0; s.cp = 0;
	gai::Philox4 rnd; rnd.v[0] = rnd.v[1] = rnd.v[2] = rnd.v[3] = 0;
	u32 ply = 0, leaf = 0;
	bool need = true, exhausted = false;
	for (u32 it = 0;; ++it) {
		const u32 phase = it & 3u;
		if (phase == 0) {
			const unsigned needmask = __ballot_sync(FULL, need && !exhausted);
			if (needmask) {
				const int leader = __ffs(needmask) - 1;
				unsigned long long base = 0;
				if ((int)lane == leader) base = atomicAdd(work_counter, (unsigned long long)__popc(needmask));
				base = __shfl_sync(FULL, base, leader);
				if (need && !exhausted) {
					item = base + __popc(needmask & ((1u << lane) - 1u));
					if (item >= total) exhausted = true;
					else {
						leaf = ppl_pow2 ? (u32)(item >> ppl_shift) : (u32)(item / ppl);
						if (load_prepared(s, prep, leaf)) { ply = 0; need = false; }      // a refused leaf (counted by k_pan_prepare): the lane pulls again
					}
				}
			}
			if (__all_sync(FULL, exhausted)) break;
		}
		const bool alive = !need && !exhausted;
		// one ply: terminal test, ply cap, move generation (GetPlayerLegalMoves), uniform choice, apply -- straight-line code,
		// every lane executes it (a lane without a playout works on stale registers and discards the result)
		const u32 holders = players_with_cards(s, np);
		const Moves m = gen_moves(s);
		int code = -1;                                              // 0..3 = loser, 4 = ply cap / dead end
		if (alive) code = PAN_POPC(holders) == 1 ? (PAN_FFS(holders) - 1) >> 2 : (ply >= max_plies || m.n == 0) ? 4 : -1;
		if (phase == 0) rnd = gai::choice_block(key, first_id + item, ply >> 2);          // warp-uniform
		if (alive && code < 0) {
			u32 cards, count, op;
			nth_move(m, gai::choice_index(pick(rnd, phase), m.n), &cards, &count, &op);
			apply(s, np, cards, count, op);
			++ply;
		}
		if (code >= 0) {
			u32* r = out + 12 * (size_t)leaf;
			if (code < 4) atomicAdd(r + code, 1u);
			else {
				atomicAdd(r + 8, 1u);
#pragma unroll
				for (int p = 0; p < 4; ++p) if (p < np) atomicAdd(r + 4 + p, (u32)eval_one(s, eval_kind, p));      // static p: no local array
			}
			my_plies += ply; my_playouts += 1;
			need = true;
		}
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		my_plies += __shfl_xor_sync(FULL, my_plies, d);
		my_playouts += __shfl_xor_sync(FULL, my_playouts, d);
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
void launch_pan_rollout(const void* d_states, void* d_prep, uint32_t n_leaves, int np, uint32_t ppl, uint32_t max_plies, int eval_kind, uint64_t key, uint64_t first_id,
                        uint32_t* d_out, uint64_t* d_stats, unsigned long long* d_counter, uint32_t* d_bad, int blocks, cudaStream_t st)
{
	k_pan_prepare<<<pblk(n_leaves, 256), 256, 0, st>>>((const u64*)d_states, n_leaves, (uint4*)d_prep, d_bad);
	// Two kernels, identical results.  The warp-synchronous one wins whenever playout lengths differ inside a warp (uncapped
	// games: 2.6x; two players: +29 %); with three or four players and the reference's playout_depth of 50 nearly every
	// playout runs into the cap, divergence is bounded, and the lane-independent loop (no refill cadence) is 11-17 % faster
	// (profiles/r01_pan_bench.json).  GAI_PAN_KERNEL=0/1 forces one of them.
	const char* fe = getenv("GAI_PAN_KERNEL");
	const int forced = fe && *fe ? atoi(fe) : -1;
	const int variant = forced >= 0 ? forced : (max_plies <= 64 && np >= 3) ? 0 : 1;
	if (variant == 0) k_pan_rollout_indep<<<blocks, 256, 0, st>>>((const uint4*)d_prep, n_leaves, np, ppl, max_plies, eval_kind, key, first_id, d_out, (u64*)d_stats, d_counter, d_bad);
	else k_pan_rollout<<<blocks, 256, 0, st>>>((const uint4*)d_prep, n_leaves, np, ppl, max_plies, eval_kind, key, first_id, d_out, (u64*)d_stats, d_counter, d_bad);
}
