// game-ai_b200/csrc/loa_fast_kernels.cu -- rollout / sweep kernels on the "rotated boards + line LUT" rules.
// One CTA per SM (the 129 KB line LUT lives in shared memory), persistent, one playout per thread.
#include "loa_fast.cuh"
#include "loa_sweep.cuh"
#include "loa_sim.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace loa {

__constant__ DiagTables c_diag2 = make_diag_tables();
__constant__ SquareTable c_sqt = make_square_table();
__constant__ LineTable c_lines = make_line_table();
__constant__ NbxTable c_nbx = make_nbx_table();
__constant__ ChiTable c_chi = make_chi_table();

struct FastSmem {
	const uint16_t* lut; const u64* sqt; const u64* lines; const u64* ld; const u64* rd; const u64* nbr; const u32* slt; const u32* tpc;
	const u64* nbx; const signed char* chi;
	const u64* nbxl; const u32* posl;          // this lane's column of the per-lane copies (static-sweep kernels)
};

// dynamic shared memory: [LUT_BYTES lut][64 x u64 sqt][4 x 64 x u64 line masks: row, column, ld, rd][64 x u64 neighbourhoods]
//                        [64 x u64 window controls (loa_euler.cuh)][384 B chi deltas][tail]
//   tail of the static-sweep kernels : [256 x 32 lanes x u32 short-line table][256 x 32 x u32 field-count table]
//                                      [64 x 32 lanes x u64 window controls][64 x 32 lanes x u32 rotated positions]
//                                      (per-lane copies: entry pitch 256 B / 128 B, lane L only touches its own banks; ncu on v18:
//                                      the plain 512-byte tables cost 4.9 wavefronts per LDS.64 and 3.2 per LDS.32, and the
//                                      LSU data pipe, not the issue rate, had become the first limit)
//   tail of the per-piece kernels    : [16 x 1024 bytes piece squares]
constexpr int FAST_SQ_OFF = LUT_BYTES + 7 * 64 * 8 + CHI_TAB_SIZE;
constexpr int FAST_SMEM_BYTES = FAST_SQ_OFF + 16 * 1024;
constexpr int SWEEP_PERLANE_OFF = FAST_SQ_OFF + 2 * 256 * 32 * 4;
constexpr int SWEEP_SMEM_BYTES = SWEEP_PERLANE_OFF + 64 * 32 * 8 + 64 * 32 * 4;
static_assert(SWEEP_SMEM_BYTES <= 227 * 1024, "sweep kernel shared memory");

// g_lut = the global image: line LUT, then the two 256-entry tables (expanded here to one copy per lane / bank)
template <bool SWEEP_TABLES>
__device__ __forceinline__ FastSmem load_fast_smem(unsigned char* smem, const uint4* __restrict__ g_lut)
{
	uint4* d = reinterpret_cast<uint4*>(smem);
	for (int i = threadIdx.x; i < LUT_BYTES / 16; i += blockDim.x) d[i] = __ldg(g_lut + i);
	u64* sq = reinterpret_cast<u64*>(smem + LUT_BYTES);
	for (int i = threadIdx.x; i < 64; i += blockDim.x) {
		sq[i] = c_sqt.w[i]; sq[320 + i] = c_lines.nbr[i];
		sq[64 + i] = c_lines.line[0][i]; sq[128 + i] = c_lines.line[1][i]; sq[192 + i] = c_lines.line[2][i]; sq[256 + i] = c_lines.line[3][i];
		sq[384 + i] = c_nbx.w[i];
	}
	signed char* chi = reinterpret_cast<signed char*>(sq + 448);
	for (int i = threadIdx.x; i < CHI_TAB_SIZE; i += blockDim.x) chi[i] = c_chi.d[i];
	FastSmem f; f.lut = reinterpret_cast<const uint16_t*>(smem); f.sqt = sq; f.lines = sq + 64; f.ld = sq + 192; f.rd = sq + 256; f.nbr = sq + 320;
	f.nbx = sq + 384; f.chi = chi;
	f.slt = f.tpc = nullptr; f.nbxl = nullptr; f.posl = nullptr;
	if (SWEEP_TABLES) {
		u32* t = reinterpret_cast<u32*>(smem + FAST_SQ_OFF);
		const u32* g = reinterpret_cast<const u32*>(reinterpret_cast<const unsigned char*>(g_lut) + SLT_OFF);      // 512 words: slt, tpc
		for (int i = threadIdx.x; i < 2 * 256 * 32; i += blockDim.x) t[i] = __ldg(g + (i >> 5));
		f.slt = t; f.tpc = t + 256 * 32;
		u64* nl = reinterpret_cast<u64*>(smem + SWEEP_PERLANE_OFF);
		u32* pl = reinterpret_cast<u32*>(smem + SWEEP_PERLANE_OFF + 64 * 32 * 8);
		for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) { nl[i] = c_nbx.w[i >> 5]; pl[i] = (u32)(c_sqt.w[i >> 5] >> 32); }
		f.nbxl = nl + (threadIdx.x & 31); f.posl = pl + (threadIdx.x & 31);
	}
	__syncthreads();
	return f;
}

struct FastSim {
	Side my, en;       // side to move / side that just moved
	u32 side, ply, rq;
	bool chk_en, chk_my;
	int chi_my, chi_en; // Euler characteristics of the two figures (static-sweep kernel; loa_euler.cuh)
};

__device__ __forceinline__ void fsim_init(FastSim& s, u64 white, u64 black, u32 side, const FastSmem& f)
{
	s.side = side & 1u;
	s.my = make_side(s.side ? black : white, f.sqt);
	s.en = make_side(s.side ? white : black, f.sqt);
	s.ply = 0; s.rq = 0xffffffffu;
	s.chk_en = true; s.chk_my = true;
	s.chi_my = s.chi_en = 0;
}

// one ply; -1 while the playout continues, else the outcome code (same semantics as loa_sim.cuh sim_step)
__device__ __forceinline__ int fsim_step(FastSim& s, gai::Philox4& rnd, u64 key, u64 id, u32 max_plies, const FastSmem& f)
{
	const bool c_en = s.chk_en && connected_fast(s.en.R);
	const bool c_my = s.chk_my && connected_fast(s.my.R);
	if (c_en | c_my) {
		const bool wc = s.side ? c_en : c_my, bc = s.side ? c_my : c_en;
		return outcome_of(wc, bc);
	}
	if (s.ply >= max_plies) return OUT_CAPPED;

	const MoveMasks m = all_piece_moves_fast(s.my, s.en, f.lut, f.sqt);
	const u32 n = (u32)count_moves(m);
	bool cap = false, moved = false;
	if (n != 0) {
		if ((s.ply >> 2) != s.rq) { s.rq = s.ply >> 2; rnd = gai::choice_block(key, id, s.rq); }
		const u32 idx = gai::choice_index(pick_word(rnd, s.ply), n);
		int from, to;
		nth_move(m, s.my.R, s.en.R, idx, f.ld, f.rd, &from, &to);
		cap = (s.en.R >> to) & 1ull;
		apply_move_fast(s.my, s.en, from, to, f.sqt);
		moved = true;
	}
	const Side t = s.my; s.my = s.en; s.en = t;
	s.side ^= 1u;
	s.ply += 1;
	s.chk_en = moved; s.chk_my = cap;
	return -1;
}

// ---------------------------------------------------------------------------------------------
// Warp-synchronous playout step.  All 32 lanes of a warp execute every phase together; a lane without work in a
// phase is predicated off.  Data-dependent loops (pieces of the side to move, flood-fill iterations) run for the
// warp-wide maximum trip count under a __any_sync condition, so the warp reconverges every iteration instead of
// splitting into independently scheduled fragments (ncu on the lane-independent version: 9.8 of 32 lanes active,
// the piece loop body issued 17.8x per warp-ply for at most 12 pieces).
// ---------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

// calcIfTerminal (LinesOfActionZasady.cpp:52-58) for the lanes with want == true
__device__ __forceinline__ bool wconnected(u64 b, bool want)
{
	bool res = false;
	bool pend = want && b != 0;
	if (pend && (b & (b - 1)) == 0) { res = true; pend = false; }       // exactly one piece
	if (pend && has_isolated_piece(b)) pend = false;                    // >= 2 pieces, one with no neighbour
	u64 g = b & (0 - b);
	while (__any_sync(FULL, pend)) {
		if (pend) {
			const u64 n = dilate8(g) & b;
			if (n == g) { pend = false; res = (g == b); }
			g = n;
		}
	}
	return res;
}

// One ply for every lane with alive == true.  Returns the outcome code for lanes whose playout ended in this
// step, -1 otherwise (also for lanes that were not alive).
// (per-piece formulation; the static sweep has its own step below)
__device__ __forceinline__ int wstep(FastSim& s, bool alive, gai::Philox4& rnd, u64 key, u64 id, u32 max_plies, const FastSmem& f,
                                     unsigned char* __restrict__ sq_scratch)
{
	const bool c_en = wconnected(s.en.R, alive && s.chk_en);
	bool c_my = false;
	if (__any_sync(FULL, alive && s.chk_my)) c_my = wconnected(s.my.R, alive && s.chk_my);
	const bool term = c_en | c_my;
	const bool capped = alive && !term && s.ply >= max_plies;
	const bool play = alive && !term && !capped;
	int result = -1;
	if (term) { const bool wc = s.side ? c_en : c_my, bc = s.side ? c_my : c_en; result = outcome_of(wc, bc); }
	if (capped) result = OUT_CAPPED;

	// move generation: per own piece four LUT lookups; all lanes iterate for the warp-wide maximum piece count
	// (one REDUX), the piece's square is parked in shared memory ([piece][thread] bytes, conflict-free)
	u64 rest = play ? s.my.R : 0ull;
	const int maxp = __reduce_max_sync(FULL, __popcll(rest));
	u32 m0 = 0, m1 = 0, m2 = 0, m3 = 0;          // move masks, pieces 0-3 / 4-7 / 8-11 / 12-15, one byte per piece
	unsigned char* sqp = sq_scratch + threadIdx.x;
	for (int j = 0; j < maxp; ++j) {
		if (rest != 0) {
			const int sq = __ffsll((long long)rest) - 1;
			rest &= rest - 1;
			const u32 v = piece_moves_fast(s.my, s.en, sq, f.lut, f.sqt) << (8 * (j & 3));
			sqp[j * 1024] = (unsigned char)sq;
			if (j < 4) m0 |= v; else if (j < 8) m1 |= v; else if (j < 12) m2 |= v; else m3 |= v;
		}
	}
	if (play) {
		const u32 c0 = __popc(m0), c1 = __popc(m1), c2 = __popc(m2), c3 = __popc(m3);
		const u32 n = c0 + c1 + c2 + c3;
		bool cap = false, moved = false;
		if (n != 0) {
			if ((s.ply >> 2) != s.rq) { s.rq = s.ply >> 2; rnd = gai::choice_block(key, id, s.rq); }
			u32 k = gai::choice_index(pick_word(rnd, s.ply), n);
			// rank-select, branch free: which word, then 5 halving steps inside the word
			u32 x = m0; int base = 0;
			{ const bool g = k >= c0; k -= g ? c0 : 0u; x = g ? m1 : x; base = g ? 32 : base;
			  const bool g1 = g && k >= c1; k -= g1 ? c1 : 0u; x = g1 ? m2 : x; base = g1 ? 64 : base;
			  const bool g2 = g1 && k >= c2; k -= g2 ? c2 : 0u; x = g2 ? m3 : x; base = g2 ? 96 : base; }
			int pos = 0;
			{ u32 c = __popc(x & 0xffffu); bool g = k >= c; k -= g ? c : 0u; pos += g ? 16 : 0; x = g ? x >> 16 : x;
			  c = __popc(x & 0xffu); g = k >= c; k -= g ? c : 0u; pos += g ? 8 : 0; x = g ? x >> 8 : x;
			  c = __popc(x & 0xfu); g = k >= c; k -= g ? c : 0u; pos += g ? 4 : 0; x = g ? x >> 4 : x;
			  c = __popc(x & 0x3u); g = k >= c; k -= g ? c : 0u; pos += g ? 2 : 0; x = g ? x >> 2 : x;
			  pos += (k >= (x & 1u)) ? 1 : 0; }
			const int bit = base + pos;
			const int from = sqp[(bit >> 3) * 1024];
			const int to = move_target(s.my.R | s.en.R, from, bit & 7, f.ld, f.rd);
			cap = (s.en.R >> to) & 1ull;
			apply_move_fast(s.my, s.en, from, to, f.sqt);
			moved = true;
		}
		const Side t = s.my; s.my = s.en; s.en = t;
		s.side ^= 1u;
		s.ply += 1;
		s.chk_en = moved; s.chk_my = cap;
	}
	return result;
}

// The flood fill alone (no pre-filters): only lanes whose Euler characteristic is at most 1 get here -- one group (the game
// is over), or, rarely, k groups and at least k - 1 holes, or an empty board.
__device__ __forceinline__ bool wfill(u64 b, bool want)
{
	bool res = false;
	bool pend = want && b != 0;
	u64 g = b & (0 - b);
	while (__any_sync(FULL, pend)) {
		if (pend) {
			const u64 n = dilate8(g) & b;
			if (n == g) { pend = false; res = (g == b); }
			g = n;
		}
	}
	return res;
}

// One ply of the static-sweep formulation with FIXED register roles: `my` moves, `en` is the side that just moved, and
// the caller alternates the two argument orders ply by ply instead of swapping sixteen registers (the swap was 32
// IMAD.MOV per ply).  `phase` = loop iteration & 3: lanes start playouts only on iterations that are multiples of four,
// so ply & 3 == phase in every lane and the Philox block (4 choice words) is refreshed by ALL lanes in the same iteration
// (ncu on a per-lane refresh: the 60-instruction Philox block issued in 97 % of warp-plies at 8 active lanes).
// Connectivity: chi (components - holes) of both colours is carried along and updated by apply_move_chi; the flood fill
// runs only for a colour whose figure changed in the previous ply AND has chi == 1.
__device__ __forceinline__ int wstep_sweep(Side& my, Side& en, int& chi_my, int& chi_en, u32& ply, u32 side0, bool alive, gai::Philox4& rnd, u64 key, u64 id, u32 max_plies,
                                           const FastSmem& f, const SweepTabs& lut, u32 phase)
{
	// one group means chi = 1 - holes <= 1, so chi > 1 proves "not connected" without a fill.  A figure that did not change
	// since it was last found disconnected keeps its chi: the only repeated fills are those of the rare figures with as many
	// holes as extra groups.
	const bool t_en = alive && chi_en <= 1, t_my = alive && chi_my <= 1;
	const bool c_en = wfill(en.R, t_en);
	bool c_my = false;
	if (__any_sync(FULL, t_my)) c_my = wfill(my.R, t_my);
	const bool term = c_en | c_my;
	const bool capped = alive && !term && ply >= max_plies;
	const bool play = alive && !term && !capped;
	int result = -1;
	if (term) { const u32 side = side0 ^ (ply & 1u); const bool wc = side ? c_en : c_my, bc = side ? c_my : c_en; result = outcome_of(wc, bc); }
	if (capped) result = OUT_CAPPED;

	const RowCounts rc = sweep_rows(my, en, lut);
	const u32 n = play ? rc.n : 0u;
	if (phase == 0) rnd = gai::choice_block(key, id, ply >> 2);      // warp-uniform branch
	const u32 idx = gai::choice_index(pick_word(rnd, phase), n);
	int from, dir;
	sweep_select(rc, idx, lut, &from, &dir);
	if (play) {
		if (n != 0) {
			const int to = move_target_tab(my.R | en.R, from, dir, f.lines);
			apply_move_chi(my, en, chi_my, chi_en, from, to, f.posl[32 * from], f.posl[32 * to], f.nbxl[32 * from], f.nbxl[32 * to], f.chi);
		}
		ply += 1;
	}
	return result;
}

// Leaf preparation: the four rotations of both colours, once per leaf (64 B = 4 x 16-byte vectors), so that a
// lane that pulls a new playout issues four LDG.128 instead of rebuilding the rotations piece by piece while the
// other 31 lanes of its warp wait (ncu: 6 % of all issued instructions at 1 active lane before this kernel existed).
//
// Scheduling: the piece loop of the rollout kernel runs for the warp-wide MAXIMUM number of own pieces, so lanes that
// hold late-game positions idle while a neighbour works through an opening position.  Leaves are therefore handed out in
// order of DESCENDING piece count (counting sort: histogram here, scatter in k_order_leaves): at any moment every lane on
// the GPU is refilled from the same piece-count bucket and older playouts have aged towards it.  The order changes which
// lane runs a playout, never its result (playout id = leaf index * ppl + j).
__global__ void k_prepare_leaves(const ulonglong2* __restrict__ boards, u32 n, ulonglong2* __restrict__ prep,
                                 uint8_t* __restrict__ pcount, u32* __restrict__ hist, uint16_t* __restrict__ chi)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const ulonglong2 bd = boards[i];
	{
		u32 pc = (u32)(__popcll(bd.x) + __popcll(bd.y));
		pc = pc > 31u ? 31u : pc;
		pcount[i] = (uint8_t)pc;
		atomicAdd(hist + pc, 1u);
	}
	const Side w = make_side(bd.x, c_sqt.w), b = make_side(bd.y, c_sqt.w);
	prep[4 * (size_t)i + 0] = make_ulonglong2(w.R, w.C);
	prep[4 * (size_t)i + 1] = make_ulonglong2(w.D1, w.D2);
	prep[4 * (size_t)i + 2] = make_ulonglong2(b.R, b.C);
	prep[4 * (size_t)i + 3] = make_ulonglong2(b.D1, b.D2);
	// Euler characteristic of both figures (low byte white, high byte black, two's complement)
	const int cw = chi_of(bd.x, c_nbx.w, c_chi.d), cb = chi_of(bd.y, c_nbx.w, c_chi.d);
	chi[i] = (uint16_t)((u32)(cw & 0xff) | ((u32)(cb & 0xff) << 8));
}

// hist[0..31] = leaves per piece count, hist[32..63] = scatter cursors (zeroed by the host before k_prepare_leaves)
__global__ void k_order_leaves(const uint8_t* __restrict__ pcount, u32 n, u32* __restrict__ hist, u32* __restrict__ order)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u32 pc = pcount[i];
	u32 off = 0;
	for (u32 q = pc + 1; q < 32; ++q) off += hist[q];            // buckets with more pieces come first
	order[off + atomicAdd(hist + 32 + pc, 1u)] = i;
}

// Per-piece formulation (rules impl 1): persistent, one playout per lane, refill from a global work counter.
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
k_rollout_fast(const uint4* __restrict__ g_lut, const ulonglong2* __restrict__ prep, const u32* __restrict__ stm, const u32* __restrict__ order,
               u32 n_leaves, u32 ppl, u32 max_plies, u64 key, u64 first_id, u32* __restrict__ out,
               u64* __restrict__ stats, unsigned long long* __restrict__ work_counter)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const FastSmem f = load_fast_smem<false>(smem, g_lut);

	const u64 total = (u64)n_leaves * ppl;
	const bool ppl_pow2 = (ppl & (ppl - 1u)) == 0;
	const u32 ppl_shift = (u32)__ffs((int)ppl) - 1u;
	const u32 lane = threadIdx.x & 31u;
	FastSim s; gai::Philox4 rnd;
	u64 item = 0, pid = 0; u32 leaf = 0;
	bool need = true, exhausted = false;
	u64 my_plies = 0, my_playouts = 0;
	s.my = Side{0, 0, 0, 0}; s.en = Side{0, 0, 0, 0}; s.side = 0; s.ply = 0; s.rq = 0; s.chk_en = s.chk_my = false; s.chi_my = s.chi_en = 0;
	rnd.v[0] = rnd.v[1] = rnd.v[2] = rnd.v[3] = 0;
	for (;;) {
		// refill: lanes whose playout ended pull the next (leaf, playout) item; one atomic per warp per refill round
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
					const u32 slot = ppl_pow2 ? (u32)(item >> ppl_shift) : (u32)(item / ppl);
					leaf = __ldg(order + slot);
					pid = first_id + (u64)leaf * ppl + (item - (u64)slot * ppl);       // identity = (leaf, playout), not the schedule
					const u32 side = (__ldg(stm + (leaf >> 5)) >> (leaf & 31u)) & 1u;
					const ulonglong2* lp = prep + 4 * (size_t)leaf;
					const ulonglong2 m0 = __ldg(lp + 2 * side), m1 = __ldg(lp + 2 * side + 1), e0 = __ldg(lp + 2 * (side ^ 1u)), e1 = __ldg(lp + 2 * (side ^ 1u) + 1);
					s.side = side; s.my = Side{m0.x, m0.y, m1.x, m1.y}; s.en = Side{e0.x, e0.y, e1.x, e1.y};
					s.ply = 0; s.rq = 0xffffffffu; s.chk_en = true; s.chk_my = true;
					need = false;
				}
			}
		}
		if (__all_sync(FULL, exhausted)) break;
		const int o = wstep(s, !exhausted && !need, rnd, key, pid, max_plies, f, smem + FAST_SQ_OFF);
		if (o >= 0) {
			atomicAdd(out + 4 * (size_t)leaf + o, 1u);
			my_plies += s.ply; my_playouts += 1;
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

// The static-sweep rollout kernel (rules impl 2, the default): one persistent CTA per SM, one playout per lane.
// Per-lane state that lives across plies is kept to the minimum the 72-register budget of 896 threads allows: the two
// colours' rotations (16 registers, FIXED roles A = side to move at the leaf, B = the other; the step alternates the
// argument order), chi of both, the Philox block, the playout id, leaf | side-at-the-leaf << 31, the ply and two 32-bit
// per-lane counters (a lane plays ~1e4 plies per launch; they are widened when the warp reduces them at the end).
template <int MAXT, int UNROLL>
__global__ void __launch_bounds__(MAXT, 1)
k_rollout_sweep(const uint4* __restrict__ g_lut, const ulonglong2* __restrict__ prep, const u32* __restrict__ stm, const u32* __restrict__ order,
                const uint16_t* __restrict__ chi16, u32 n_leaves, u32 ppl, u32 max_plies, u64 key, u64 first_id, u32* __restrict__ out,
                u64* __restrict__ stats, unsigned long long* __restrict__ work_counter)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const FastSmem f = load_fast_smem<true>(smem, g_lut);

	const u64 total = (u64)n_leaves * ppl;
	const bool ppl_pow2 = (ppl & (ppl - 1u)) == 0;
	const u32 ppl_shift = (u32)__ffs((int)ppl) - 1u;
	const u32 lane = threadIdx.x & 31u;
	Side A{0, 0, 0, 0}, B{0, 0, 0, 0};
	int chiA = 0, chiB = 0;
	gai::Philox4 rnd; rnd.v[0] = rnd.v[1] = rnd.v[2] = rnd.v[3] = 0;
	u64 pid = 0; u32 leafside = 0, ply = 0;
	bool need = true, exhausted = false;
	u32 my_plies = 0, my_playouts = 0;

	const SweepTabs lut = sweep_tabs(f.lut, f.slt, f.tpc, lane);
	for (u32 it = 0;; it += (u32)UNROLL) {
		// refill (iterations that are multiples of four only, see wstep_sweep): lanes whose playout ended pull the next
		// (leaf, playout) item; one atomic per warp per refill round
		const unsigned needmask = (it & 3u) == 0 ? __ballot_sync(FULL, need && !exhausted) : 0u;
		if (needmask) {
			const int leader = __ffs(needmask) - 1;
			unsigned long long base = 0;
			if ((int)lane == leader) base = atomicAdd(work_counter, (unsigned long long)__popc(needmask));
			base = __shfl_sync(FULL, base, leader);
			if (need && !exhausted) {
				const u64 item = base + __popc(needmask & ((1u << lane) - 1u));
				if (item >= total) exhausted = true;
				else {
					const u32 slot = ppl_pow2 ? (u32)(item >> ppl_shift) : (u32)(item / ppl);      // warp-uniform: no division for 2^k playouts per leaf
					const u32 leaf = __ldg(order + slot);
					pid = first_id + (u64)leaf * ppl + (item - (u64)slot * ppl);       // identity = (leaf, playout), not the schedule
					const u32 side = (__ldg(stm + (leaf >> 5)) >> (leaf & 31u)) & 1u;
					const ulonglong2* lp = prep + 4 * (size_t)leaf;       // four 16-byte loads per playout: {white R C, white D1 D2, black R C, black D1 D2}
					const ulonglong2 m0 = __ldg(lp + 2 * side), m1 = __ldg(lp + 2 * side + 1), e0 = __ldg(lp + 2 * (side ^ 1u)), e1 = __ldg(lp + 2 * (side ^ 1u) + 1);
					A = Side{m0.x, m0.y, m1.x, m1.y}; B = Side{e0.x, e0.y, e1.x, e1.y};
					const u32 ch = __ldg(chi16 + leaf);
					const int cw = (int)(signed char)(ch & 0xffu), cb = (int)(signed char)(ch >> 8);
					chiA = side ? cb : cw; chiB = side ? cw : cb;
					leafside = leaf | (side << 31); ply = 0;
					need = false;
				}
			}
		}
		if (__all_sync(FULL, exhausted)) break;
		// UNROLL plies per trip; with an even UNROLL the two sides trade places by argument order, with UNROLL = 1 by a swap
#pragma unroll
		for (int j = 0; j < UNROLL; ++j) {
			const u32 phase = UNROLL == 4 ? (u32)j : ((it + (u32)j) & 3u);
			const bool alive = !exhausted && !need;
			int o;
			if (UNROLL == 1) {
				o = wstep_sweep(A, B, chiA, chiB, ply, leafside >> 31, alive, rnd, key, pid, max_plies, f, lut, phase);
				const Side t = A; A = B; B = t;
				const int tc = chiA; chiA = chiB; chiB = tc;
			}
			else if (j & 1) o = wstep_sweep(B, A, chiB, chiA, ply, leafside >> 31, alive, rnd, key, pid, max_plies, f, lut, phase);
			else o = wstep_sweep(A, B, chiA, chiB, ply, leafside >> 31, alive, rnd, key, pid, max_plies, f, lut, phase);
			// per-leaf results: lanes of this warp that finished a playout of the SAME leaf with the SAME outcome in this ply
			// are reduced first (match-any), one atomic per group
			const unsigned fin = __ballot_sync(FULL, o >= 0);
			if (o >= 0) {
				const u32 leaf = leafside & 0x7fffffffu;
				const unsigned grp = __match_any_sync(fin, (leaf << 2) | (u32)o);
				if (lane == (u32)__ffs((int)grp) - 1u) atomicAdd(out + 4 * (size_t)leaf + o, (u32)__popc(grp));
				my_plies += ply; my_playouts += 1;
				need = true;
			}
		}
	}
	u64 wp = my_plies, wq = my_playouts;
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		wp += __shfl_xor_sync(FULL, wp, d);
		wq += __shfl_xor_sync(FULL, wq, d);
	}
	if (lane == 0) { atomicAdd((unsigned long long*)stats, wq); atomicAdd((unsigned long long*)stats + 1, wp); }
}

// tier-1 sweep through the SAME movegen the rollout kernel uses (persistent grid-stride, one CTA per SM)
__global__ void __launch_bounds__(1024, 1)
k_movegen_fast(const uint4* __restrict__ g_lut, const u64* __restrict__ states, u32 n, uint8_t* __restrict__ counts, uint8_t* __restrict__ moves)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const FastSmem f = load_fast_smem<false>(smem, g_lut);
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const u64 w = states[3 * (size_t)i], b = states[3 * (size_t)i + 1];
		const u32 side = (u32)states[3 * (size_t)i + 2] & 1u;
		const Side my = make_side(side ? b : w, f.sqt), en = make_side(side ? w : b, f.sqt);
		const MoveMasks m = all_piece_moves_fast(my, en, f.lut, f.sqt);
		const u32 cnt = (u32)count_moves(m);
		counts[i] = (uint8_t)cnt;
		uint8_t* outp = moves + (size_t)i * 2 * GAI_LOA_MAX_MOVES_K;
		for (u32 k = 0; k < cnt && k < GAI_LOA_MAX_MOVES_K; ++k) {
			int from, to;
			nth_move(m, my.R, en.R, k, f.ld, f.rd, &from, &to);
			outp[2 * k] = (uint8_t)from; outp[2 * k + 1] = (uint8_t)to;
		}
	}
}

// the same sweep through count + rank-select of the static-sweep formulation (every move k selected on its own)
__global__ void __launch_bounds__(256, 1)
k_movegen_sweep(const uint4* __restrict__ g_lut, const u64* __restrict__ states, u32 n, uint8_t* __restrict__ counts, uint8_t* __restrict__ moves)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const FastSmem f = load_fast_smem<true>(smem, g_lut);
	const u32 stride = gridDim.x * blockDim.x;
	const SweepTabs tabs = sweep_tabs(f.lut, f.slt, f.tpc, threadIdx.x & 31u);
	for (u32 base = blockIdx.x * blockDim.x; base < n; base += stride) {          // whole warps stay together (REDUX inside)
		const u32 i = base + threadIdx.x;
		const bool live = i < n;
		Side my{0, 0, 0, 0}, en{0, 0, 0, 0};
		if (live) {
			const u64 w = states[3 * (size_t)i], b = states[3 * (size_t)i + 1];
			const u32 side = (u32)states[3 * (size_t)i + 2] & 1u;
			my = make_side(side ? b : w, f.sqt); en = make_side(side ? w : b, f.sqt);
		}
		const RowCounts rc = sweep_rows(my, en, tabs);
		const u32 cnt = live ? rc.n : 0u;
		if (live) counts[i] = (uint8_t)cnt;
		for (u32 k = 0; k < cnt; ++k) {
			int from, dir;
			sweep_select(rc, k, tabs, &from, &dir);
			if (k < GAI_LOA_MAX_MOVES_K) {
				uint8_t* outp = moves + (size_t)i * 2 * GAI_LOA_MAX_MOVES_K;
				outp[2 * k] = (uint8_t)from;
				outp[2 * k + 1] = (uint8_t)move_target(my.R | en.R, from, dir, f.ld, f.rd);
			}
		}
	}
}

// successor hash through apply_move_fast: checks all four rotations stay consistent with the plain board
__global__ void __launch_bounds__(1024, 1)
k_successor_hash_fast(const uint4* __restrict__ g_lut, const u64* __restrict__ states, u32 n, u64* __restrict__ hash, u32* __restrict__ bad)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const FastSmem f = load_fast_smem<false>(smem, g_lut);
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const u64 w = states[3 * (size_t)i], b = states[3 * (size_t)i + 1];
		const u32 side = (u32)states[3 * (size_t)i + 2] & 1u;
		const Side my = make_side(side ? b : w, f.sqt), en = make_side(side ? w : b, f.sqt);
		const MoveMasks m = all_piece_moves_fast(my, en, f.lut, f.sqt);
		const u32 cnt = (u32)count_moves(m);
		u64 h = 0;
		for (u32 k = 0; k < cnt; ++k) {
			int from, to;
			nth_move(m, my.R, en.R, k, f.ld, f.rd, &from, &to);
			Side nmy = my, nen = en;
			apply_move_fast(nmy, nen, from, to, f.sqt);
			// rotations must equal a fresh rebuild from the plain board
			const Side cm = make_side(nmy.R, f.sqt), ce = make_side(nen.R, f.sqt);
			if (cm.C != nmy.C || cm.D1 != nmy.D1 || cm.D2 != nmy.D2 || ce.C != nen.C || ce.D1 != nen.D1 || ce.D2 != nen.D2) atomicAdd(bad, 1u);
			const u64 nw = side ? nen.R : nmy.R, nb = side ? nmy.R : nen.R;
			const bool wc = connected_fast(nw), bc = connected_fast(nb);
			h = (h ^ nw); h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 27; h *= 0x94d049bb133111ebull; h ^= h >> 31;
			h = (h ^ nb); h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 27; h *= 0x94d049bb133111ebull; h ^= h >> 31;
			h = h ^ ((u64)(side ^ 1u) | ((u64)k << 8) | ((u64)(wc ? 1 : 0) << 16) | ((u64)(bc ? 1 : 0) << 17));
			h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 27; h *= 0x94d049bb133111ebull; h ^= h >> 31;
		}
		hash[i] = h;
	}
}

__global__ void __launch_bounds__(1024, 1)
k_playout_trace_fast(const uint4* __restrict__ g_lut, const u64* __restrict__ states, u32 n, u32 max_plies, u64 key, u64 first_id,
                     uint8_t* __restrict__ outcome, u32* __restrict__ plies, u64* __restrict__ final_states)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const FastSmem f = load_fast_smem<false>(smem, g_lut);
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		FastSim s; gai::Philox4 rnd;
		fsim_init(s, states[3 * (size_t)i], states[3 * (size_t)i + 1], (u32)states[3 * (size_t)i + 2] & 1u, f);
		int o;
		do { o = fsim_step(s, rnd, key, first_id + i, max_plies, f); } while (o < 0);
		outcome[i] = (uint8_t)o;
		plies[i] = s.ply;
		final_states[3 * (size_t)i] = s.side ? s.en.R : s.my.R;
		final_states[3 * (size_t)i + 1] = s.side ? s.my.R : s.en.R;
		final_states[3 * (size_t)i + 2] = s.side;
	}
}

// Single playouts with full detail THROUGH THE STATIC-SWEEP STEP (wstep_sweep: sweep + select + apply_move_chi + the
// chi-gated flood fill), one playout per lane, whole warps in lock step until every lane has finished: the tier-2 trace of
// the default kernel is the default kernel's own code path, final position included.
__global__ void __launch_bounds__(256, 1)
k_playout_trace_sweep(const uint4* __restrict__ g_lut, const u64* __restrict__ states, u32 n, u32 max_plies, u64 key, u64 first_id,
                      uint8_t* __restrict__ outcome, u32* __restrict__ plies, u64* __restrict__ final_states)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const FastSmem f = load_fast_smem<true>(smem, g_lut);
	const SweepTabs lut = sweep_tabs(f.lut, f.slt, f.tpc, threadIdx.x & 31u);
	const u32 stride = gridDim.x * blockDim.x;
	for (u32 base = blockIdx.x * blockDim.x; base < n; base += stride) {
		const u32 i = base + threadIdx.x;
		bool alive = i < n;
		gai::Philox4 rnd;
		rnd.v[0] = rnd.v[1] = rnd.v[2] = rnd.v[3] = 0;
		Side A{0, 0, 0, 0}, B{0, 0, 0, 0};
		int chiA = 0, chiB = 0; u32 ply = 0, side0 = 0;
		if (alive) {
			const u64 w = states[3 * (size_t)i], b = states[3 * (size_t)i + 1];
			side0 = (u32)states[3 * (size_t)i + 2] & 1u;
			A = make_side(side0 ? b : w, f.sqt); B = make_side(side0 ? w : b, f.sqt);
			chiA = chi_of(A.R, f.nbx, f.chi); chiB = chi_of(B.R, f.nbx, f.chi);
		}
		for (u32 it = 0; __any_sync(FULL, alive); ++it) {
			const int o = (it & 1u) ? wstep_sweep(B, A, chiB, chiA, ply, side0, alive, rnd, key, first_id + i, max_plies, f, lut, it & 3u)
			                        : wstep_sweep(A, B, chiA, chiB, ply, side0, alive, rnd, key, first_id + i, max_plies, f, lut, it & 3u);
			if (alive && o >= 0) {
				outcome[i] = (uint8_t)o; plies[i] = ply;
				final_states[3 * (size_t)i] = side0 ? B.R : A.R;              // A = the colour that was to move at the start
				final_states[3 * (size_t)i + 1] = side0 ? A.R : B.R;
				final_states[3 * (size_t)i + 2] = side0 ^ (ply & 1u);
				alive = false;
			}
		}
	}
}

} // namespace loa

using namespace loa;

int loa_fast_lut_bytes() { return LUT_IMAGE_BYTES; }
void loa_fast_build_lut_host(uint16_t* lut)
{
	for (int i = 0; i < LUT_IMAGE_BYTES / 2; ++i) lut[i] = 0;
	for (u32 e = 0; e < 256; ++e)
		for (u32 o = 0; o < 256; ++o) lut[o + e * LUT_STRIDE] = lut_entry(o, e);
	u32* small = reinterpret_cast<u32*>(reinterpret_cast<unsigned char*>(lut) + SLT_OFF);
	for (u32 i = 0; i < 256; ++i) { small[i] = slt_entry(i); small[256 + i] = tpc_entry(i); }
}
cudaError_t loa_fast_prepare()
{
	cudaError_t e;
	if ((e = cudaFuncSetAttribute(k_rollout_fast<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, FAST_SMEM_BYTES)) != cudaSuccess) return e;
#define GAI_PREP(T, U) if ((e = cudaFuncSetAttribute(k_rollout_sweep<T, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM_BYTES)) != cudaSuccess) return e;
	GAI_PREP(1024, 1) GAI_PREP(768, 1) GAI_PREP(512, 1) GAI_PREP(1024, 2) GAI_PREP(896, 2) GAI_PREP(768, 2) GAI_PREP(640, 2) GAI_PREP(512, 2) GAI_PREP(1024, 4) GAI_PREP(768, 4) GAI_PREP(512, 4)
#undef GAI_PREP
	if ((e = cudaFuncSetAttribute(k_movegen_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM_BYTES)) != cudaSuccess) return e;
	if ((e = cudaFuncSetAttribute(k_playout_trace_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM_BYTES)) != cudaSuccess) return e;
	if ((e = cudaFuncSetAttribute(k_movegen_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, FAST_SMEM_BYTES)) != cudaSuccess) return e;
	if ((e = cudaFuncSetAttribute(k_successor_hash_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, FAST_SMEM_BYTES)) != cudaSuccess) return e;
	return cudaFuncSetAttribute(k_playout_trace_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, FAST_SMEM_BYTES);
}
static int fast_grid(uint32_t n, int threads, int sms) { const int need = (int)((n + threads - 1) / threads); return need < sms ? (need ? need : 1) : sms; }

void launch_loa_prepare_leaves(const void* d_boards, uint32_t n, void* d_prep, uint8_t* d_pcount, uint32_t* d_hist, uint32_t* d_order, uint16_t* d_chi, cudaStream_t st)
{
	if (!n) return;
	k_prepare_leaves<<<(n + 127) / 128, 128, 0, st>>>((const ulonglong2*)d_boards, n, (ulonglong2*)d_prep, d_pcount, d_hist, d_chi);
	k_order_leaves<<<(n + 255) / 256, 256, 0, st>>>(d_pcount, n, d_hist, d_order);
}

void launch_loa_rollout_fast(const void* d_lut, const void* d_boards, const uint32_t* d_stm, const uint32_t* d_order, const uint16_t* d_chi, uint32_t n_leaves, uint32_t ppl, uint32_t max_plies,
                             uint64_t key, uint64_t first_id, uint32_t* d_out, uint64_t* d_stats, unsigned long long* d_counter,
                             int threads, int sms, int sweep, cudaStream_t st)
{
	const uint64_t total = (uint64_t)n_leaves * ppl;
	// A batch that cannot fill one CTA per SM (the MCTS player's leaf batches: 1024 leaves x 32 playouts) is spread over ALL
	// SMs with fewer warps each instead of packing a quarter of the SMs full: the kernel then lasts as long as its longest
	// playout, and a playout advances ~3x faster when its warp shares the sub-partition with one other warp instead of six.
	if (total < (uint64_t)sms * (uint64_t)threads && !getenv("GAI_NO_SPREAD")) {
		int t = (int)(((total + (uint64_t)sms - 1) / (uint64_t)sms + 31) / 32 * 32);
		if (t < 64) t = 64;
		if (t < threads) threads = t;
	}
	const int blocks = fast_grid(total > 0xffffffffull ? 0xffffffffu : (uint32_t)total, threads, sms);
#define GAI_RL_ARGS (const uint4*)d_lut, (const ulonglong2*)d_boards, d_stm, d_order, d_chi, n_leaves, ppl, max_plies, key, first_id, d_out, (u64*)d_stats, d_counter
#define GAI_RF_ARGS (const uint4*)d_lut, (const ulonglong2*)d_boards, d_stm, d_order, n_leaves, ppl, max_plies, key, first_id, d_out, (u64*)d_stats, d_counter
	// sweep: 0 = per-piece formulation, else the unroll factor of the static-sweep kernel (1, 2 or 4 plies per loop trip)
#define GAI_RL(T, U) k_rollout_sweep<T, U><<<blocks, threads, SWEEP_SMEM_BYTES, st>>>(GAI_RL_ARGS)
#define GAI_RL_T(U) do { if (threads > 768) GAI_RL(1024, U); else if (threads > 512) GAI_RL(768, U); else GAI_RL(512, U); } while (0)
	if (!sweep) k_rollout_fast<1024><<<blocks, threads, FAST_SMEM_BYTES, st>>>(GAI_RF_ARGS);
	else if (sweep >= 4) GAI_RL_T(4);
	else if (sweep >= 2) { if (threads > 896) GAI_RL(1024, 2); else if (threads > 768) GAI_RL(896, 2); else if (threads > 640) GAI_RL(768, 2); else if (threads > 512) GAI_RL(640, 2); else GAI_RL(512, 2); }
	else GAI_RL_T(1);
#undef GAI_RL_T
#undef GAI_RL
#undef GAI_RL_ARGS
#undef GAI_RF_ARGS
}
void launch_loa_movegen_sweep(const void* d_lut, const void* d_states, uint32_t n, uint8_t* d_counts, uint8_t* d_moves, int sms, cudaStream_t st)
{ if (n) k_movegen_sweep<<<fast_grid(n, 256, sms), 256, SWEEP_SMEM_BYTES, st>>>((const uint4*)d_lut, (const u64*)d_states, n, d_counts, d_moves); }
void launch_loa_movegen_fast(const void* d_lut, const void* d_states, uint32_t n, uint8_t* d_counts, uint8_t* d_moves, int sms, cudaStream_t st)
{ if (n) k_movegen_fast<<<fast_grid(n, 256, sms), 256, FAST_SMEM_BYTES, st>>>((const uint4*)d_lut, (const u64*)d_states, n, d_counts, d_moves); }
void launch_loa_successor_hash_fast(const void* d_lut, const void* d_states, uint32_t n, uint64_t* d_hash, uint32_t* d_bad, int sms, cudaStream_t st)
{ if (n) k_successor_hash_fast<<<fast_grid(n, 256, sms), 256, FAST_SMEM_BYTES, st>>>((const uint4*)d_lut, (const u64*)d_states, n, (u64*)d_hash, d_bad); }
void launch_loa_playout_trace_fast(const void* d_lut, const void* d_states, uint32_t n, uint32_t max_plies, uint64_t key, uint64_t first_id,
                                   uint8_t* d_outcome, uint32_t* d_plies, void* d_final, int sms, cudaStream_t st)
{ if (n) k_playout_trace_fast<<<fast_grid(n, 256, sms), 256, FAST_SMEM_BYTES, st>>>((const uint4*)d_lut, (const u64*)d_states, n, max_plies, key, first_id, d_outcome, d_plies, (u64*)d_final); }
void launch_loa_playout_trace_sweep(const void* d_lut, const void* d_states, uint32_t n, uint32_t max_plies, uint64_t key, uint64_t first_id,
                                    uint8_t* d_outcome, uint32_t* d_plies, void* d_final, int sms, cudaStream_t st)
{ if (n) k_playout_trace_sweep<<<fast_grid(n, 256, sms), 256, SWEEP_SMEM_BYTES, st>>>((const uint4*)d_lut, (const u64*)d_states, n, max_plies, key, first_id, d_outcome, d_plies, (u64*)d_final); }
