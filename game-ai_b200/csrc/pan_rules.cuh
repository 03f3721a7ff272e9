// game-ai_b200/csrc/pan_rules.cuh -- Pan ("Gra w Pana") rules for DETERMINIZED (complete-information) states, device code.
//
// Device restatement of c++/GraWPanaZasadyV2/GraWPanaZasadyV2.cpp (reference; line numbers after iconv to UTF-8).
// The reference keeps 2 bits of "certainty" per card (:16-49); a determinized deal has only 00 / 11, so on the device a
// hand is ONE bit per card (24 bits, card = rank*4 + suit, ranks 9 10 W D K A) and the whole position fits six 32-bit
// registers.  The 4-bit `count` field of each hand (:18) is carried separately and wraps modulo 16 exactly like the
// reference's bit-field does when a hand grows past 15 cards.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace pan {

typedef unsigned int u32;
typedef unsigned long long u64;

enum { OP_NOOP = 0, OP_TAKE = 1, OP_PLAY = 2 };      // Move::operation, :44
constexpr u32 ALL24 = 0xffffffu;
constexpr u64 CARDS48 = 0xFFFFFFFFFFFFull;

#if defined(__CUDA_ARCH__)
#define PAN_CLZ(x) __clz((int)(x))
#define PAN_POPC(x) __popc(x)
#define PAN_FFS(x) __ffs((int)(x))
#else
#define PAN_CLZ(x) __builtin_clz(x)
#define PAN_POPC(x) __builtin_popcount(x)
#define PAN_FFS(x) __builtin_ffs((int)(x))
#endif

struct State {
	u32 stack;       // cards on the stack
	u32 hand[4];     // cards held
	u32 counts;      // nibble p = hand[p].count (mod 16)
	u32 cp;          // current player
};

// 2 bits per card (00/11) <-> 1 bit per card
__host__ __device__ inline u32 squeeze(u64 x)
{
	x &= 0x5555555555555555ull;
	x = (x | (x >> 1)) & 0x3333333333333333ull;
	x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
	x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
	x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
	x = (x | (x >> 16)) & 0x00000000ffffffffull;
	return (u32)x;
}
__host__ __device__ inline u64 spread(u32 m)
{
	u64 x = m;
	x = (x | (x << 16)) & 0x0000ffff0000ffffull;
	x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
	x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
	x = (x | (x << 2)) & 0x3333333333333333ull;
	x = (x | (x << 1)) & 0x5555555555555555ull;
	return x | (x << 1);
}
// a 48-bit card field is "determinized" when every card is 00 or 11
__host__ __device__ inline bool is_crisp(u64 cards) { return (((cards >> 1) ^ cards) & 0x5555555555555555ull & CARDS48) == 0; }

// raw reference GameState words (:22-40) -> device state; returns false for fuzzy (non-determinized) cards
__host__ __device__ inline bool load_state(State& s, const u64* w)
{
	bool ok = is_crisp(w[0] & CARDS48);
	s.stack = squeeze(w[0] & CARDS48);
	s.cp = (u32)((w[0] >> 48) & 7u);
	s.counts = 0;
	for (int p = 0; p < 4; ++p) {
		const u64 cards = (w[1 + p] >> 4) & CARDS48;
		ok = ok && is_crisp(cards);
		s.hand[p] = squeeze(cards);
		s.counts |= (u32)(w[1 + p] & 15u) << (4 * p);
	}
	return ok;
}
__host__ __device__ inline u32 count_of(const State& s, u32 p) { return (s.counts >> (4 * p)) & 15u; }
// checkIfTerminal (:339-346): exactly one player still holds cards (by the count field).  players_with_cards = one bit per
// nibble of `counts` that is non-zero, for the np players of the game, so the test is a popcount.
__host__ __device__ inline u32 players_with_cards(const State& s, int np)
{
	const u32 c = s.counts;
	return (c | (c >> 1) | (c >> 2) | (c >> 3)) & (0x1111u >> (4 * (4 - np)));
}
__host__ __device__ inline bool is_terminal(const State& s, int np) { return PAN_POPC(players_with_cards(s, np)) == 1; }
__host__ __device__ inline void store_state(u64* w, const State& s, int np)
{
	w[0] = spread(s.stack) | ((u64)(s.cp & 7u) << 48) | ((u64)(is_terminal(s, np) ? 1 : 0) << 51);
	for (int p = 0; p < 4; ++p) w[1 + p] = (u64)count_of(s, p) | (spread(s.hand[p]) << 4);
}

// The move list of the current player in compact form (GetPlayerLegalMoves, :409-475), reference order:
//   [three other nines] , then per rank from the top-of-stack rank upward: [lowest playable card][all four] , [take]
struct Moves {
	u32 n;           // number of moves
	u32 first;       // 1 if the list starts with the special move (9h alone on an empty stack / the three other nines)
	u32 first_cards; // its cards
	u32 w;           // bit 4q = "single" move exists for rank q, bit 4q+1 = "all four" exists
	u32 masked;      // hand & allowed cards
	u32 take;        // cards of the take move (0 = none)
};
// All three functions below are written WITHOUT data-dependent branches (selects only): in a warp every lane holds a
// different deal, and the branchy first version ran at 19-21 of 32 lanes active (ncu, 3 players, playout_depth 50).
__host__ __device__ inline u32 top_bit(u32 x) { return x ? (0x80000000u >> PAN_CLZ(x)) : 0u; }     // highest set bit as a mask
__host__ __device__ inline Moves gen_moves(const State& s)
{
	Moves m;
	const u32 cpi = s.cp & 3u;                             // selects, not a dynamically indexed array: the state stays in registers
	const u32 hand = cpi == 0 ? s.hand[0] : cpi == 1 ? s.hand[1] : cpi == 2 ? s.hand[2] : s.hand[3];
	const u32 stk = s.stack;
	const bool empty = stk == 0;
	// empty stack (:415-421): open with 9h, or with all four nines ("all four" of rank 0 = bit 1 of w)
	const u32 first_e = hand & 1u, w_e = (hand & 0xfu) == 0xfu ? 2u : 0u;
	// cards on the stack: the three other nines on a lone 9h (:424-429); per rank from the top rank upward the lowest
	// playable card / all four (makeStackMasks allowed_cards, :376-385); take the top <= 3 cards, never the 9h (:386-397)
	const u32 first_n = (stk == 1u && (hand & 0xeu) == 0xeu) ? 1u : 0u;
	const int top = 31 - PAN_CLZ(stk | 1u);
	const u32 above = (ALL24 << (top + 1)) & ALL24;
	const u32 quad0 = 0xfu << (top & ~3);
	const u32 x = hand & (above | (~stk & quad0));
	const u32 any = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x111111u;
	const u32 all = (x & (x >> 1) & (x >> 2) & (x >> 3)) & 0x111111u;
	u32 st = stk & ~1u;
	const u32 b1 = top_bit(st); st &= ~b1;
	const u32 b2 = top_bit(st); st &= ~b2;
	const u32 b3 = top_bit(st);
	m.first = empty ? first_e : first_n;
	m.first_cards = m.first ? (empty ? 1u : 0xeu) : 0u;
	m.w = empty ? w_e : (any | (all << 1));
	m.masked = empty ? 0u : x;
	m.take = empty ? 0u : (b1 | b2 | b3);
	m.n = m.first + (u32)PAN_POPC(m.w) + (m.take ? 1u : 0u);
	return m;
}
// decode move idx -> (cards, count, operation); idx < m.n
__host__ __device__ inline void nth_move(const Moves& m, u32 idx, u32* cards, u32* count, u32* op)
{
	const bool is_first = m.first != 0 && idx == 0;
	u32 j = idx - m.first;                                 // rank among the per-rank moves (garbage when is_first)
	const bool is_take = !is_first && j >= (u32)PAN_POPC(m.w);
	// position of the j-th set bit of w (bits 4q = "single", 4q+1 = "all four", q < 6): three halving steps
	u32 w = m.w; int pos = 0;
	{ u32 c = (u32)PAN_POPC(w & 0xfffu); bool g = j >= c; j -= g ? c : 0u; pos += g ? 12 : 0; w = g ? w >> 12 : w;
	  c = (u32)PAN_POPC(w & 0xffu); g = j >= c; j -= g ? c : 0u; pos += g ? 8 : 0; w = g ? w >> 8 : w;
	  c = (u32)PAN_POPC(w & 0xfu); g = j >= c; j -= g ? c : 0u; pos += g ? 4 : 0; w = g ? w >> 4 : w;
	  pos += (j >= (w & 1u)) ? 1 : 0; }
	pos &= 31;                                             // is_first / is_take: pos is unused, keep the shifts below defined
	const int q4 = pos & ~3;
	const bool four = (pos & 1) != 0;
	const u32 from = (m.masked >> q4) & 0xfu;
	const u32 play = (four ? 0xfu : (from & (0u - from))) << q4;
	*cards = is_first ? m.first_cards : is_take ? m.take : play;
	*count = is_first ? (u32)PAN_POPC(m.first_cards) : is_take ? (u32)PAN_POPC(m.take) : (four ? 4u : 1u);
	*op = is_take ? (u32)OP_TAKE : (u32)OP_PLAY;
}
// apply_move (:534-557) + calcNextPlayer (:510-518)
__host__ __device__ inline void apply(State& s, int np, u32 cards, u32 count, u32 op)
{
	const u32 p = s.cp & 3u;
	const bool take = op == OP_TAKE, play = op == OP_PLAY;
	u32 c = count_of(s, p);
	s.stack = take ? (s.stack & ~cards) : play ? (s.stack | cards) : s.stack;
	// masks, not "i == p ? ... : ..." -- the compiler turns that select into a dynamically indexed (local-memory) access
	const u32 pc = play ? cards : 0u, tc = take ? cards : 0u, onehot = 1u << p;
	for (int i = 0; i < 4; ++i) s.hand[i] = (s.hand[i] & ~pc) | (tc & (0u - ((onehot >> i) & 1u)));
	c = take ? ((c + count) & 15u) : play ? ((c - count) & 15u) : c;
	s.counts = (s.counts & ~(15u << (4 * p))) | (c << (4 * p));
	// the next player that still holds cards, else the last candidate (np is the same in every lane)
	const u32 unp = (u32)np;
	const u32 n1 = p + 1 >= unp ? 0u : p + 1, n2 = n1 + 1 >= unp ? 0u : n1 + 1, n3 = n2 + 1 >= unp ? 0u : n2 + 1;
	u32 next = n1;
	if (np >= 3) {
		const u32 rest = np >= 4 ? (count_of(s, n2) > 0 ? n2 : n3) : n2;
		next = count_of(s, n1) > 0 ? n1 : rest;
	}
	s.cp = next;
}
// Score (:347-363)
__host__ __device__ inline void score(const State& s, int np, int* sc)
{
	if (is_terminal(s, np)) { const int win = 100 / (np - 1); for (int p = 0; p < np; ++p) sc[p] = count_of(s, p) != 0 ? 0 : win; }
	else { const int draw = 100 / np; for (int p = 0; p < np; ++p) sc[p] = draw; }
}
// eval functions (:255-311): 0 = num_cards, 1 = num_cards_weighted (certainty 1 for every held card: exact in int)
__host__ __device__ inline int eval_one(const State& s, int kind, int p)
{
	if (kind != 1) return 2 * (24 - (int)count_of(s, p));
	const u32 h = p == 0 ? s.hand[0] : p == 1 ? s.hand[1] : p == 2 ? s.hand[2] : s.hand[3];      // selects: no local-memory copy of the state
	int sum = 0;
	for (int r = 0; r < 6; ++r) sum += (6 - r) * PAN_POPC((h >> (4 * r)) & 0xfu);
	return 84 - sum;
}
// raw reference Move word (:42-49); probability 11 for determinized cards (take moves are 11 in the reference too, :471)
__host__ __device__ inline u64 move_word(u32 cards, u32 count, u32 op)
{
	return spread(cards) | ((u64)(count & 7u) << 48) | ((u64)op << 51) | (3ull << 54);
}

} // namespace pan
