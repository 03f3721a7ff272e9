/* oracle/pan_oracle.c -- TEST INFRASTRUCTURE ONLY (checker for the CUDA Pan rollouts; never shipped).
 *
 * Plain-C restatement of the Pan ("Gra w Pana") rules of the reference,
 * c++/GraWPanaZasadyV2/GraWPanaZasadyV2.cpp (line numbers after iconv UTF-16 -> UTF-8), plus the Philox-driven
 * random playout that defines tier-2 truth for Pan.  Parity is PINNED by tests/test_pan_oracle.py against
 *   (a) the reference's own unit-test vectors (UniTestsWithBoost/game_rules_v2.cpp:57-111,120-366) in
 *       tests/golden/pan_ut_vectors.json,
 *   (b) committed outputs of the reference compiled here (oracle/_ref/libref_pan.so) in tests/golden/pan_ref_corpus.npz,
 *   (c) the live reference whenever oracle/_ref is present.
 *
 * Wire format = the reference's own structs as raw words (little-endian bit-fields, g++ and MSVC agree):
 *   GameState (:16-40, 40 B): w[0] = stack:48 | current_player:3 << 48 | is_terminal:1 << 51
 *                             w[1+p] = count:4 | cards:48 << 4           (Hand, :16-20)
 *   Move (:42-49, 8 B): cards:48 | count:3 << 48 | operation:3 << 51 | probability:2 << 54
 * card c in [0,24) = rank*4 + suit at bits 2c..2c+1; the 2-bit value is a certainty 0 / 1/3 / 1/2 / 1 (:48,502).
 */
#include <stdint.h>
#include <string.h>
#include "philox_ref.h"

#define CARDS_MASK 0xFFFFFFFFFFFFull
enum { OP_NOOP = 0, OP_TAKE = 1, OP_PLAY = 2 };        /* :44 */

typedef struct { uint64_t stack; int cp; int term; uint64_t cards[4]; int count[4]; } PanState;

static void unpack(PanState* s, const uint64_t* w)
{
	s->stack = w[0] & CARDS_MASK; s->cp = (int)((w[0] >> 48) & 7); s->term = (int)((w[0] >> 51) & 1);
	for (int p = 0; p < 4; ++p) { s->count[p] = (int)(w[1 + p] & 15); s->cards[p] = (w[1 + p] >> 4) & CARDS_MASK; }
}
static void pack(uint64_t* w, const PanState* s)
{
	w[0] = (s->stack & CARDS_MASK) | ((uint64_t)(s->cp & 7) << 48) | ((uint64_t)(s->term & 1) << 51);
	for (int p = 0; p < 4; ++p) w[1 + p] = (uint64_t)(s->count[p] & 15) | ((s->cards[p] & CARDS_MASK) << 4);
}
static uint64_t mk_move(uint64_t cards, int count, int op, uint64_t prob)
{
	return (cards & CARDS_MASK) | ((uint64_t)(count & 7) << 48) | ((uint64_t)(op & 7) << 51) | ((prob & 3) << 54);
}

/* :370-375 -- any non-zero certainty becomes 11 */
uint64_t pan_oracle_make11(uint64_t cards)
{
	uint64_t out = 0;
	for (int c = 0; c < 32; ++c) if ((cards >> (2 * c)) & 3) out |= 3ull << (2 * c);
	return out;
}
static int top_card(uint64_t stack)      /* index of the highest card on the stack */
{
	return (63 - __builtin_clzll(stack)) / 2;
}
/* :376-398 -- allowed = every card above the top of the stack, plus the cards of the top card's rank that are not on
 * the stack; take = the top (up to three) cards, but never the 9 of hearts (card 0) */
void pan_oracle_make_stack_masks(uint64_t stack, uint64_t* out3)
{
	const int top = top_card(stack);
	const uint64_t above = (top + 1 >= 24) ? 0 : ((~0ull << (2 * (top + 1))) & CARDS_MASK);
	const uint64_t quad = 0xFFull << (8 * (top / 4));
	uint64_t take = 0, st = stack;
	for (int taken = 0; taken < 3 && st > 3; ++taken) {
		const uint64_t tos = 3ull << (2 * top_card(st));
		take |= tos; st &= ~tos;
	}
	out3[0] = above | (~stack & quad); out3[1] = quad; out3[2] = take;
}
/* :399-408 */
static uint64_t lowest_out_of(uint64_t hand, int cnt)
{
	uint64_t lowest = 3;
	while (cnt--) { if ((hand & 3) < lowest) lowest = hand & 3; hand >>= 2; }
	return lowest;
}
/* :519-531 */
static int count_cards(uint64_t cards)
{
	int n = 0;
	for (; cards; cards >>= 2) if (cards & 3) ++n;
	return n;
}
/* :339-346 */
static int check_terminal(int np, const PanState* s)
{
	int with_cards = 0;
	for (int p = 0; p < np; ++p) if (s->count[p] > 0) ++with_cards;
	return with_cards == 1;
}
int pan_oracle_check_terminal(int np, const uint64_t* st) { PanState s; unpack(&s, st); return check_terminal(np, &s); }

/* :409-475 -- moves of the current player in the reference's order */
static int movegen(const PanState* s, uint64_t* mv)
{
	int n = 0;
	if (s->term) return 0;                                         /* m_empty, :411 */
	const uint64_t hand = s->cards[s->cp];
	if (s->stack == 0) {                                            /* :415-421 */
		if (hand & 3) mv[n++] = mk_move(3, 1, OP_PLAY, hand & 3);
		if ((pan_oracle_make11(hand) & 0xFF) == 0xFF) mv[n++] = mk_move(0xFF, 4, OP_PLAY, lowest_out_of(hand, 4));
		return n;
	}
	if (s->stack <= 3) {                                            /* only the 9 of hearts lies: the three other nines, :424-429 */
		if ((pan_oracle_make11(hand) & 0xFC) == 0xFC) mv[n++] = mk_move(0xFC, 3, OP_PLAY, lowest_out_of(hand >> 2, 3));
	}
	uint64_t m3[3];
	pan_oracle_make_stack_masks(s->stack, m3);
	uint64_t quad = m3[1], masked = hand & m3[0];
	while (masked) {                                                /* :452-469 -- per rank: lowest playable card, then all four */
		const uint64_t from_quad = quad & masked;
		if (from_quad) {
			const int idx = __builtin_ctzll(from_quad) & ~1;
			mv[n++] = mk_move(3ull << idx, 1, OP_PLAY, (from_quad >> idx) & 3);
			if (pan_oracle_make11(from_quad) == quad) mv[n++] = mk_move(0xFFull << idx, 4, OP_PLAY, lowest_out_of(from_quad >> idx, 4));
		}
		masked &= ~quad; quad <<= 8;
	}
	if (m3[2]) mv[n++] = mk_move(m3[2], count_cards(m3[2]), OP_TAKE, 3);   /* :470-472 */
	return n;
}
int pan_oracle_movegen(int np, const uint64_t* st, uint64_t* moves)
{
	(void)np; PanState s; unpack(&s, st);
	return movegen(&s, moves);
}
/* :534-557 apply_move, :510-518 calcNextPlayer, :560-569 ApplyMove */
static void apply(int np, PanState* s, uint64_t move)
{
	const uint64_t cards = move & CARDS_MASK;
	const int cnt = (int)((move >> 48) & 7), op = (int)((move >> 51) & 7), player = s->cp;
	if (op == OP_TAKE) { s->stack &= ~cards; s->cards[player] |= cards; s->count[player] = (s->count[player] + cnt) & 15; }
	else if (op == OP_PLAY) {
		s->stack |= cards;
		for (int p = 0; p < np; ++p) s->cards[p] &= ~cards;
		s->count[player] = (s->count[player] - cnt) & 15;
	}
	int next = player;
	for (int i = 0; i < np - 1; ++i) { next = (next + 1) % np; if (s->count[next] > 0) break; }
	s->cp = next;
	s->term = check_terminal(np, s);
}
void pan_oracle_apply(int np, const uint64_t* st, uint64_t move, uint64_t* out)
{
	PanState s; unpack(&s, st); apply(np, &s, move); pack(out, &s);
}
/* :347-363 */
static void score(int np, const PanState* s, int* sc)
{
	if (s->term) { const int win = 100 / (np - 1); for (int p = 0; p < np; ++p) sc[p] = s->count[p] != 0 ? 0 : win; }
	else { const int draw = 100 / np; for (int p = 0; p < np; ++p) sc[p] = draw; }
}
void pan_oracle_score(int np, const uint64_t* st, int* sc) { PanState s; unpack(&s, st); score(np, &s, sc); }
/* :255-311 -- kind 0 "num_cards": 2*(24-count); kind 1 "num_cards_weighted": int(84 - sum weight*certainty) (float) */
static void eval(int np, int kind, const PanState* s, int* v)
{
	static const float prob[4] = { 0.f, 1.f / 3.f, 0.5f, 1.f };
	for (int p = 0; p < np; ++p) {
		if (kind != 1) { v[p] = 2 * (24 - s->count[p]); continue; }
		float sum = 0;
		for (int j = 0; j < 24; ++j) { const int c = (int)((s->cards[p] >> (2 * j)) & 3); if (c) sum += (float)(6 - j / 4) * prob[c]; }
		v[p] = (int)(84.0f - sum);
	}
}
void pan_oracle_eval(int np, int kind, const uint64_t* st, int* v) { PanState s; unpack(&s, st); eval(np, kind, &s, v); }

/* :102-133 -- deal: 20 Philox-driven swaps of the ordered deck, round-robin deal, the holder of the 9 of hearts starts */
void pan_oracle_deal(int np, uint64_t key, uint64_t id, uint64_t* out)
{
	int cards[24]; uint32_t k = 0;
	for (int i = 0; i < 24; ++i) cards[i] = i;
	for (int i = 0; i < 20; ++i) {
		const int a = (int)oracle_choice_index(oracle_choice_word(key, id, k++), 24);
		const int b = (int)oracle_choice_index(oracle_choice_word(key, id, k++), 24);
		const int t = cards[a]; cards[a] = cards[b]; cards[b] = t;
	}
	PanState s; memset(&s, 0, sizeof s);
	int p = 0;
	for (int i = 0; i < 24; ++i) { s.cards[p] |= 3ull << (2 * cards[i]); p = (p + 1) % np; }
	for (p = 0; p < np; ++p) s.count[p] = 24 / np;
	for (p = 0; p < np; ++p) if (s.cards[p] & 3) break;
	s.cp = p; s.term = check_terminal(np, &s);
	pack(out, &s);
}
/* :214-243 -- the player's view: own cards certain, every unseen card marked in every other hand */
void pan_oracle_player_known_state(int np, const uint64_t* st, int player, uint64_t* out)
{
	static const uint64_t card_mask[5] = { 3, 3, 3, 2, 1 };
	PanState s, k; unpack(&s, st); memset(&k, 0, sizeof k);
	uint64_t other = 0;
	for (int c = 0; c < 24; ++c) {
		if (s.cards[player] & (3ull << (2 * c))) k.cards[player] |= 3ull << (2 * c);
		else other |= card_mask[np] << (2 * c);
	}
	for (int p = 0; p < np; ++p) { if (p != player) k.cards[p] = other; k.count[p] = s.count[p]; }
	k.stack = s.stack; k.cp = s.cp; k.term = s.term;
	pack(out, &k);
}

/* Tier-2 truth for Pan: result 0..3 = the loser (player left with cards), 4 = ply cap / dead end with value[] = eval */
int pan_oracle_playout(int np, const uint64_t* st, uint64_t key, uint64_t id, uint32_t max_plies, int eval_kind,
                       uint32_t* plies_out, int* value, uint64_t* final_state)
{
	PanState s; unpack(&s, st);
	s.term = check_terminal(np, &s);
	uint64_t mv[16]; uint32_t ply = 0; int code = 4;
	for (;;) {
		if (s.term) { for (int p = 0; p < np; ++p) if (s.count[p] != 0) code = p; score(np, &s, value); break; }
		const int n = (ply >= max_plies) ? 0 : movegen(&s, mv);
		if (n == 0) { eval(np, eval_kind, &s, value); break; }
		apply(np, &s, mv[oracle_choice_index(oracle_choice_word(key, id, ply), (uint32_t)n)]);
		++ply;
	}
	if (plies_out) *plies_out = ply;
	if (final_state) pack(final_state, &s);
	return code;
}

/* ---- batch forms ------------------------------------------------------------------------------- */
void pan_oracle_movegen_batch(int np, const uint64_t* states, uint32_t n, uint8_t* counts, uint64_t* moves)
{
	for (uint32_t i = 0; i < n; ++i) { memset(moves + (size_t)i * 16, 0, 128); counts[i] = (uint8_t)pan_oracle_movegen(np, states + (size_t)i * 5, moves + (size_t)i * 16); }
}
void pan_oracle_apply_batch(int np, const uint64_t* states, const uint64_t* mv, uint32_t n, uint64_t* out)
{
	for (uint32_t i = 0; i < n; ++i) pan_oracle_apply(np, states + (size_t)i * 5, mv[i], out + (size_t)i * 5);
}
void pan_oracle_terminal_batch(int np, const uint64_t* states, uint32_t n, uint8_t* terminal, int32_t* sc, int32_t* eval0, int32_t* eval1)
{
	for (uint32_t i = 0; i < n; ++i) {
		PanState s; unpack(&s, states + (size_t)i * 5);
		s.term = check_terminal(np, &s); terminal[i] = (uint8_t)s.term;
		int v[4] = { 0, 0, 0, 0 }; score(np, &s, v); memcpy(sc + (size_t)i * 4, v, 16);
		int a[4] = { 0, 0, 0, 0 }; eval(np, 0, &s, a); memcpy(eval0 + (size_t)i * 4, a, 16);
		int b[4] = { 0, 0, 0, 0 }; eval(np, 1, &s, b); memcpy(eval1 + (size_t)i * 4, b, 16);
	}
}
void pan_oracle_playout_trace_batch(int np, const uint64_t* states, uint32_t n, uint32_t max_plies, int eval_kind, uint64_t key, uint64_t first_id,
                                    uint8_t* code, uint32_t* plies, int32_t* value, uint64_t* final_states)
{
	for (uint32_t i = 0; i < n; ++i) {
		int v[4] = { 0, 0, 0, 0 };
		code[i] = (uint8_t)pan_oracle_playout(np, states + (size_t)i * 5, key, first_id + i, max_plies, eval_kind, &plies[i], v, final_states + (size_t)i * 5);
		memcpy(value + (size_t)i * 4, v, 16);
	}
}
uint64_t pan_oracle_playout_batch(int np, const uint64_t* states, uint32_t n_leaves, uint32_t ppl, uint32_t max_plies, int eval_kind,
                                  uint64_t key, uint64_t first_id, uint32_t* res)
{
	uint64_t total = 0;
	for (uint32_t i = 0; i < n_leaves; ++i)
		for (uint32_t j = 0; j < ppl; ++j) {
			uint32_t plies = 0; int v[4] = { 0, 0, 0, 0 };
			const int c = pan_oracle_playout(np, states + (size_t)i * 5, key, first_id + (uint64_t)i * ppl + j, max_plies, eval_kind, &plies, v, 0);
			uint32_t* r = res + (size_t)i * 12;
			if (c < 4) r[c] += 1; else { r[8] += 1; for (int p = 0; p < np; ++p) r[4 + p] += (uint32_t)v[p]; }
			total += plies;
		}
	return total;
}
void pan_oracle_gen_reachable(int np, uint32_t n, uint64_t key, uint32_t max_k, uint64_t* states)
{
	for (uint32_t i = 0; i < n; ++i) {
		uint64_t st[5], mv[16];
		pan_oracle_deal(np, key ^ 0xD1B54A32D192ED03ull, i, st);
		const uint32_t k = oracle_choice_index(oracle_choice_word(key ^ 0x9E3779B97F4A7C15ull, i, 0), max_k);
		for (uint32_t t = 0; t < k; ++t) {
			if (pan_oracle_check_terminal(np, st)) pan_oracle_deal(np, key ^ 0xD1B54A32D192ED03ull, i, st);
			const int cnt = pan_oracle_movegen(np, st, mv);
			if (cnt == 0) break;
			uint64_t ns[5];
			pan_oracle_apply(np, st, mv[oracle_choice_index(oracle_choice_word(key, i, t), (uint32_t)cnt)], ns);
			memcpy(st, ns, 40);
		}
		if (pan_oracle_check_terminal(np, st)) pan_oracle_deal(np, key ^ 0xD1B54A32D192ED03ull, i, st);
		memcpy(states + (size_t)i * 5, st, 40);
	}
}
