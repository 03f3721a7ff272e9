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
// checkIfTerminal (:339-346): exactly one player still holds cards (by the count field)
__host__ __device__ inline bool is_terminal(const State& s, int np)
{
	int with = 0;
	for (int p = 0; p < np; ++p) with += count_of(s, p) != 0 ? 1 : 0;
	return with == 1;
}
__host__ __device__ inline void store_state(u64* w, const State& s, int np)
{
	w[0] = spread(s.stack) | ((u64)(s.cp & 7u) << 48) | ((u64)(is_terminal(s, np) ? 1 : 0) << 51);
	for (int p = 0; p < 4; ++p) w[1 + p] = (u64)count_of(s, p) | (spread(s.hand[p]) << 4);
}

#if defined(__CUDA_ARCH__)
#define PAN_CLZ(x) __clz((int)(x))
#define PAN_POPC(x) __popc(x)
#define PAN_FFS(x) __ffs((int)(x))
#else
#define PAN_CLZ(x) __builtin_clz(x)
#define PAN_POPC(x) __builtin_popcount(x)
#define PAN_FFS(x) __builtin_ffs((int)(x))
#endif

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
__host__ __device__ inline Moves gen_moves(const State& s)
{
	Moves m; m.n = 0; m.first = 0; m.first_cards = 0; m.w = 0; m.masked = 0; m.take = 0;
	const u32 hand = s.hand[s.cp & 3u];
	if (s.stack == 0) {                                    // :415-421: open with 9h, or with all four nines
		if (hand & 1u) { m.first = 1; m.first_cards = 1u; }
		if ((hand & 0xfu) == 0xfu) m.w = 2u;               // "all four" of rank 0
		m.n = m.first + (m.w ? 1u : 0u);
		return m;
	}
	if (s.stack == 1u && (hand & 0xeu) == 0xeu) { m.first = 1; m.first_cards = 0xeu; }    // :424-429
	const int top = 31 - PAN_CLZ(s.stack);
	const u32 above = (ALL24 << (top + 1)) & ALL24;
	const u32 quad0 = 0xfu << (top & ~3);
	m.masked = hand & (above | (~s.stack & quad0));        // makeStackMasks allowed_cards, :376-385
	const u32 x = m.masked;
	const u32 any = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x111111u;
	const u32 all = (x & (x >> 1) & (x >> 2) & (x >> 3)) & 0x111111u;
	m.w = any | (all << 1);
	u32 st = s.stack & ~1u, take = 0;                      // take the top <= 3 cards, never the 9h (:386-397)
	for (int k = 0; k < 3 && st; ++k) { const u32 b = 1u << (31 - PAN_CLZ(st)); take |= b; st &= ~b; }
	m.take = take;
	m.n = m.first + (u32)PAN_POPC(m.w) + (take ? 1u : 0u);
	return m;
}
// decode move idx -> (cards, count, operation)
__host__ __device__ inline void nth_move(const Moves& m, u32 idx, u32* cards, u32* count, u32* op)
{
	if (m.first && idx == 0) { *cards = m.first_cards; *count = (u32)PAN_POPC(m.first_cards); *op = OP_PLAY; return; }
	idx -= m.first;
	const u32 nw = (u32)PAN_POPC(m.w);
	if (idx >= nw) { *cards = m.take; *count = (u32)PAN_POPC(m.take); *op = OP_TAKE; return; }
	u32 w = m.w;
	for (u32 k = 0; k < idx; ++k) w &= w - 1;              // idx <= 11
	const int pos = PAN_FFS(w) - 1;
	const int q = pos >> 2;
	if (pos & 1) { *cards = 0xfu << (4 * q); *count = 4; }
	else {
		const u32 from = (m.masked >> (4 * q)) & 0xfu;
		*cards = (from & (0u - from)) << (4 * q); *count = 1;
	}
	*op = OP_PLAY;
}
// apply_move (:534-557) + calcNextPlayer (:510-518)
__host__ __device__ inline void apply(State& s, int np, u32 cards, u32 count, u32 op)
{
	const u32 p = s.cp & 3u;
	u32 c = count_of(s, p);
	if (op == OP_TAKE) { s.stack &= ~cards; s.hand[p] |= cards; c = (c + count) & 15u; }
	else if (op == OP_PLAY) {
		s.stack |= cards;
		for (int i = 0; i < 4; ++i) s.hand[i] &= ~cards;
		c = (c - count) & 15u;
	}
	s.counts = (s.counts & ~(15u << (4 * p))) | (c << (4 * p));
	u32 next = p;
	for (int i = 0; i < np - 1; ++i) { next = (next + 1) % (u32)np; if (count_of(s, next) > 0) break; }
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
	const u32 h = s.hand[p];
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
