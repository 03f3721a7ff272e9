// game-ai_b200/csrc/loa_sim.cuh -- one ply of a uniformly random LOA playout (device).
//
// Replaces, per simulated ply, the reference chain MC::Player::selection_playOut (MCTSPlayer.cpp:387-470)
// -> IGameRules::GetPlayerLegalMoves / ApplyMove / IsTerminal (LinesOfActionZasady.cpp:463-489,416-419)
// for a FRESH node, where all UCB values tie and the choice is uniform (MCTSPlayer.cpp:582-593).
// Playout semantics = oracle/loa_oracle.c loa_oracle_playout (tier-2 truth):
//   terminal test, then ply cap, then one uniformly chosen move; no legal move = pass.
#pragma once
#include "loa_rules.cuh"
#include "gameai/philox.h"

namespace loa {

struct Sim {
	u64 my, en;        // bitboards of the side to move / the side that just moved
	u32 side;          // side to move: 0 = white, 1 = black (LinesOfActionZasady.cpp:40-41)
	u32 ply;
	bool chk_en;       // en changed since it was last known disconnected (its owner just moved)
	bool chk_my;       // my changed (a piece of the side to move was just captured)
	u32 rq;            // which 4-ply block of the choice stream `rnd` currently holds
};

__device__ __forceinline__ void sim_init(Sim& s, u64 white, u64 black, u32 side)
{
	s.side = side & 1u;
	s.my = s.side ? black : white;
	s.en = s.side ? white : black;
	s.ply = 0;
	s.chk_en = true; s.chk_my = true;
	s.rq = 0xffffffffu;
}
__device__ __forceinline__ u64 sim_white(const Sim& s) { return s.side ? s.en : s.my; }
__device__ __forceinline__ u64 sim_black(const Sim& s) { return s.side ? s.my : s.en; }

__device__ __forceinline__ u32 pick_word(const gai::Philox4& r, u32 ply)
{
	const u32 w = ply & 3u;
	return w == 0 ? r.v[0] : w == 1 ? r.v[1] : w == 2 ? r.v[2] : r.v[3];
}

// One loop iteration of a playout.  Returns -1 while the playout continues, else the outcome code.
// Only the colour(s) that changed in the previous ply are re-tested for connectivity: before the move
// the position was non-terminal, the mover's board always changes, the opponent's only on a capture.
__device__ __forceinline__ int sim_step(Sim& s, gai::Philox4& rnd, u64 key, u64 id, u32 max_plies,
                                        const u64* __restrict__ ld, const u64* __restrict__ rd)
{
	const bool c_en = s.chk_en && connected_fast(s.en);
	const bool c_my = s.chk_my && connected_fast(s.my);
	if (c_en | c_my) {
		const bool wc = s.side ? c_en : c_my, bc = s.side ? c_my : c_en;
		return outcome_of(wc, bc);
	}
	if (s.ply >= max_plies) return OUT_CAPPED;

	const MoveMasks m = all_piece_moves(s.my, s.en, ld, rd);
	const u32 n = (u32)count_moves(m);
	u64 nmy = s.my, nen = s.en;
	bool cap = false, moved = false;
	if (n != 0) {
		if ((s.ply >> 2) != s.rq) { s.rq = s.ply >> 2; rnd = gai::choice_block(key, id, s.rq); }
		const u32 idx = gai::choice_index(pick_word(rnd, s.ply), n);
		int from, to;
		nth_move(m, s.my, s.en, idx, ld, rd, &from, &to);
		cap = (s.en >> to) & 1ull;
		apply_move(nmy, nen, from, to);
		moved = true;
	}
	// hand the move to the other side (a pass flips the side only, LinesOfActionZasady.cpp:484-487)
	s.my = nen; s.en = nmy;
	s.side ^= 1u;
	s.ply += 1;
	s.chk_en = moved; s.chk_my = cap;
	return -1;
}

} // namespace loa
