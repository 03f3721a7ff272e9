/* oracle/loa_oracle.c -- TEST INFRASTRUCTURE ONLY.  See loa_oracle.h for scope and pinning.
 *
 * A restatement, not a copy: the reference computes line masks by shifting two diagonal constants and
 * counts groups with a byte-array BFS; this file derives masks from (row, col) and counts groups with an
 * explicit stack flood over the 8-neighbourhood.  What is kept IDENTICAL is observable behaviour: the
 * move ORDER, the blocking rule, terminal/score semantics.  Citations are path:line in
 * /root/reference/c++/LinesOfActionZasady/LinesOfActionZasady.cpp.
 */
#include "loa_oracle.h"
#include "philox_ref.h"
#include <string.h>

static inline int on_board(int r, int c) { return r >= 0 && r < 8 && c >= 0 && c < 8; }
static inline uint64_t sq(int r, int c) { return 1ull << (r * 8 + c); }   /* bit = row*8 + col, bit0 = A1 (:43) */

/* :59-81 -- the 8 squares around pos that are on the board (centre excluded) */
uint64_t loa_oracle_neighbour_mask(int pos)
{
	const int r = pos / 8, c = pos % 8;
	uint64_t m = 0;
	for (int dr = -1; dr <= 1; ++dr)
		for (int dc = -1; dc <= 1; ++dc)
			if ((dr || dc) && on_board(r + dr, c + dc)) m |= sq(r + dr, c + dc);
	return m;
}
/* :82-86 */
uint64_t loa_oracle_row_mask(int pos)
{
	uint64_t m = 0; const int r = pos / 8;
	for (int c = 0; c < 8; ++c) m |= sq(r, c);
	return m;
}
/* :87-92 */
uint64_t loa_oracle_column_mask(int pos)
{
	uint64_t m = 0; const int c = pos % 8;
	for (int r = 0; r < 8; ++r) m |= sq(r, c);
	return m;
}
/* :93-113 -- "left" diagonal runs A1->H8: row - col is constant */
uint64_t loa_oracle_left_diagonal_mask(int pos)
{
	uint64_t m = 0; const int d = pos / 8 - pos % 8;
	for (int c = 0; c < 8; ++c) if (on_board(c + d, c)) m |= sq(c + d, c);
	return m;
}
/* :114-135 -- "right" diagonal runs H1->A8: row + col is constant */
uint64_t loa_oracle_right_diagonal_mask(int pos)
{
	uint64_t m = 0; const int s = pos / 8 + pos % 8;
	for (int c = 0; c < 8; ++c) if (on_board(s - c, c)) m |= sq(s - c, c);
	return m;
}

/* :271-305 -- number of 8-connected groups of set bits */
int loa_oracle_num_groups(uint64_t board)
{
	uint64_t todo = board;
	int groups = 0;
	while (todo) {
		uint64_t frontier = todo & (~todo + 1);           /* lowest remaining piece seeds a group */
		uint64_t group = 0;
		while (frontier) {
			const int p = __builtin_ctzll(frontier);
			frontier &= frontier - 1;
			group |= 1ull << p;
			frontier |= loa_oracle_neighbour_mask(p) & board & ~group;
		}
		todo &= ~group;
		++groups;
	}
	return groups;
}
/* :52-58 -- a colour "is connected" when it has exactly one piece or exactly one group.
 * An empty board has 0 groups and popcount 0, hence false. */
int loa_oracle_calc_if_terminal(uint64_t board)
{
	if (__builtin_popcountll(board) == 1) return 1;
	return loa_oracle_num_groups(board) == 1;
}
/* :156-166, :416-419 */
int loa_oracle_is_terminal(uint64_t white, uint64_t black)
{
	return loa_oracle_calc_if_terminal(white) || loa_oracle_calc_if_terminal(black);
}
/* :420-432 -- both connected: 50/50; otherwise 100 to each connected colour (0/0 when neither) */
void loa_oracle_score(uint64_t white, uint64_t black, int score[2])
{
	const int w = loa_oracle_calc_if_terminal(white), b = loa_oracle_calc_if_terminal(black);
	if (w && b) { score[0] = score[1] = 50; return; }
	score[LOA_WHITE] = w ? 100 : 0;
	score[LOA_BLACK] = b ? 100 : 0;
}
/* :358-365 */
void loa_oracle_initial(uint64_t* white, uint64_t* black, int* cp)
{
	*black = 0x7e0000000000007eull;
	*white = 0x0081818181818100ull;
	*cp = LOA_BLACK;
}

/* :136-155 -- one candidate: target on board, not an own piece, and no ENEMY piece on the squares
 * strictly between origin and target (own pieces may be jumped; an enemy ON the target is captured).
 * The reference tests "between" with a contiguous bit range masked by the line; walking the squares
 * of the ray is the same set. */
static void try_move(uint64_t mine, uint64_t theirs, int r, int c, int dr, int dc, int dist, uint8_t* out, int* n)
{
	const int tr = r + dr * dist, tc = c + dc * dist;
	if (!on_board(tr, tc)) return;
	if (mine & sq(tr, tc)) return;
	for (int k = 1; k < dist; ++k)
		if (theirs & sq(r + dr * k, c + dc * k)) return;
	out[2 * *n] = (uint8_t)(r * 8 + c);
	out[2 * *n + 1] = (uint8_t)(tr * 8 + tc);
	++*n;
}
/* :167-232 -- own pieces in ascending bit order (:180-186); per piece the four lines in the order
 * row, column, left diagonal, right diagonal; distance = number of pieces of BOTH colours on the line
 * (:191-192); the two senses of each line in this fixed order (:194-223):
 *   west, east, north(+row), south(-row), south-west, north-east, south-east, north-west. */
int loa_oracle_movegen(uint64_t white, uint64_t black, int cp, uint8_t* out)
{
	const uint64_t mine = (cp & 1) == LOA_WHITE ? white : black;
	const uint64_t theirs = (cp & 1) == LOA_WHITE ? black : white;
	const uint64_t all = mine | theirs;
	int n = 0;
	for (uint64_t rest = mine; rest; rest &= rest - 1) {
		const int pos = __builtin_ctzll(rest), r = pos / 8, c = pos % 8;
		int d;
		d = __builtin_popcountll(all & loa_oracle_row_mask(pos));
		try_move(mine, theirs, r, c, 0, -1, d, out, &n);
		try_move(mine, theirs, r, c, 0, +1, d, out, &n);
		d = __builtin_popcountll(all & loa_oracle_column_mask(pos));
		try_move(mine, theirs, r, c, +1, 0, d, out, &n);
		try_move(mine, theirs, r, c, -1, 0, d, out, &n);
		d = __builtin_popcountll(all & loa_oracle_left_diagonal_mask(pos));
		try_move(mine, theirs, r, c, -1, -1, d, out, &n);
		try_move(mine, theirs, r, c, +1, +1, d, out, &n);
		d = __builtin_popcountll(all & loa_oracle_right_diagonal_mask(pos));
		try_move(mine, theirs, r, c, -1, +1, d, out, &n);
		try_move(mine, theirs, r, c, +1, -1, d, out, &n);
	}
	return n;
}
/* :233-252 + :481-489 -- from==to==0 is the rules' noop (m_noop = {1,{0,0}}, :337): side flips only.
 * (The reference recognises noop by ADDRESS; through the C API the (0,0) pair stands for it -- a real
 * move never has from == to, asserted at :228-230.) */
void loa_oracle_apply(uint64_t white, uint64_t black, int cp, int from, int to, uint64_t* ow, uint64_t* ob, int* ocp)
{
	if (!(from == 0 && to == 0)) {
		const uint64_t f = 1ull << from, t = 1ull << to;
		if ((cp & 1) == LOA_WHITE) { white = (white & ~f) | t; black &= ~t; }
		else                       { black = (black & ~f) | t; white &= ~t; }
	}
	*ow = white; *ob = black; *ocp = (cp + 1) & 1;
}
/* :253-270 -- bit 63 first, 8 per text line, then "cp=N" */
int loa_oracle_state_to_string(uint64_t white, uint64_t black, int cp, char* out, int cap)
{
	char buf[96]; int k = 0;
	for (int bit = 63; bit >= 0; --bit) {
		const uint64_t m = 1ull << bit;
		buf[k++] = (white & m) ? 'o' : (black & m) ? '*' : '.';
		if (bit % 8 == 0) buf[k++] = '\n';
	}
	buf[k++] = 'c'; buf[k++] = 'p'; buf[k++] = '='; buf[k++] = (char)('0' + (cp & 1)); buf[k] = 0;
	const int n = k < cap - 1 ? k : cap - 1;
	memcpy(out, buf, (size_t)n); out[n] = 0;
	return n;
}

void loa_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	oracle_philox4x32_10(ctr, key, out);
}

/* Tier-2 truth: reference rules driven by the Philox choice stream (SURVEY.md 8c last bullet, 8d).
 * terminal test first (a terminal leaf scores at 0 plies), then the ply cap, then one uniformly chosen
 * move; a side without a legal move passes. */
int loa_oracle_playout(uint64_t white, uint64_t black, int cp, uint64_t key, uint64_t playout_id,
                       uint32_t max_plies, uint32_t* plies_out, uint64_t* fw, uint64_t* fb)
{
	uint8_t mv[2 * LOA_MAX_MOVES];
	uint32_t ply = 0;
	int outcome;
	cp &= 1;
	for (;;) {
		if (loa_oracle_is_terminal(white, black)) {
			int sc[2]; loa_oracle_score(white, black, sc);
			outcome = (sc[0] == sc[1]) ? LOA_OUT_DRAW : (sc[0] > sc[1] ? LOA_OUT_WHITE : LOA_OUT_BLACK);
			break;
		}
		if (ply >= max_plies) { outcome = LOA_OUT_CAPPED; break; }
		const int n = loa_oracle_movegen(white, black, cp, mv);
		if (n == 0) {
			cp ^= 1;
		} else {
			const uint32_t idx = oracle_choice_index(oracle_choice_word(key, playout_id, ply), (uint32_t)n);
			loa_oracle_apply(white, black, cp, mv[2 * idx], mv[2 * idx + 1], &white, &black, &cp);
		}
		++ply;
	}
	if (plies_out) *plies_out = ply;
	if (fw) *fw = white;
	if (fb) *fb = black;
	return outcome;
}

uint64_t loa_oracle_playout_batch(const uint64_t* whites, const uint64_t* blacks, const uint8_t* cps,
                                  uint32_t n_leaves, uint32_t playouts_per_leaf, uint32_t max_plies,
                                  uint64_t key, uint64_t first_id, uint32_t* counts)
{
	uint64_t total = 0;
	for (uint32_t i = 0; i < n_leaves; ++i)
		for (uint32_t j = 0; j < playouts_per_leaf; ++j) {
			uint32_t plies = 0;
			const int o = loa_oracle_playout(whites[i], blacks[i], cps[i], key,
			                                 first_id + (uint64_t)i * playouts_per_leaf + j, max_plies, &plies, 0, 0);
			counts[(size_t)i * 4 + o] += 1;
			total += plies;
		}
	return total;
}

/* ---- batch forms (same functions in loops, so Python-side tests finish in seconds) --------------- */
/* states: n x {white, black, current_player} (the reference's 24-byte GameState, LinesOfActionZasady.cpp:38-46) */
void loa_oracle_movegen_batch(const uint64_t* states, uint32_t n, uint8_t* counts, uint8_t* moves)
{
	for (uint32_t i = 0; i < n; ++i) {
		uint8_t* out = moves + (size_t)i * 2 * LOA_MAX_MOVES;
		memset(out, 0, 2 * LOA_MAX_MOVES);
		counts[i] = (uint8_t)loa_oracle_movegen(states[3 * (size_t)i], states[3 * (size_t)i + 1], (int)(states[3 * (size_t)i + 2] & 1), out);
	}
}
void loa_oracle_terminal_batch(const uint64_t* states, uint32_t n, uint8_t* terminal, uint8_t* score)
{
	for (uint32_t i = 0; i < n; ++i) {
		int sc[2];
		terminal[i] = (uint8_t)loa_oracle_is_terminal(states[3 * (size_t)i], states[3 * (size_t)i + 1]);
		loa_oracle_score(states[3 * (size_t)i], states[3 * (size_t)i + 1], sc);
		score[2 * (size_t)i] = (uint8_t)sc[0]; score[2 * (size_t)i + 1] = (uint8_t)sc[1];
	}
}
void loa_oracle_apply_batch(const uint64_t* states, const uint8_t* mv, uint32_t n, uint64_t* out)
{
	for (uint32_t i = 0; i < n; ++i) {
		int cp;
		const int from = mv[2 * (size_t)i], to = mv[2 * (size_t)i + 1];
		/* from == to is the noop in the C ABI (gameai_gpu.h gai_loa_apply) */
		loa_oracle_apply(states[3 * (size_t)i], states[3 * (size_t)i + 1], (int)(states[3 * (size_t)i + 2] & 1),
		                 from == to ? 0 : from, from == to ? 0 : to, &out[3 * (size_t)i], &out[3 * (size_t)i + 1], &cp);
		out[3 * (size_t)i + 2] = (uint64_t)cp;
	}
}
static uint64_t mix64(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}
/* hash over (successor white, black, side, move index, white connected, black connected) in move order */
void loa_oracle_successor_hash_batch(const uint64_t* states, uint32_t n, uint64_t* hash)
{
	uint8_t mv[2 * LOA_MAX_MOVES];
	for (uint32_t i = 0; i < n; ++i) {
		const uint64_t w = states[3 * (size_t)i], b = states[3 * (size_t)i + 1];
		const int cp = (int)(states[3 * (size_t)i + 2] & 1);
		const int cnt = loa_oracle_movegen(w, b, cp, mv);
		uint64_t h = 0;
		for (int k = 0; k < cnt; ++k) {
			uint64_t nw, nb; int ncp;
			loa_oracle_apply(w, b, cp, mv[2 * k], mv[2 * k + 1], &nw, &nb, &ncp);
			const uint64_t wc = (uint64_t)loa_oracle_calc_if_terminal(nw), bc = (uint64_t)loa_oracle_calc_if_terminal(nb);
			h = mix64(h ^ nw); h = mix64(h ^ nb);
			h = mix64(h ^ ((uint64_t)ncp | ((uint64_t)k << 8) | (wc << 16) | (bc << 17)));
		}
		hash[i] = h;
	}
}
void loa_oracle_playout_trace_batch(const uint64_t* states, uint32_t n, uint32_t max_plies, uint64_t key, uint64_t first_id,
                                    uint8_t* outcome, uint32_t* plies, uint64_t* final_states)
{
	for (uint32_t i = 0; i < n; ++i) {
		uint64_t fw, fb;
		outcome[i] = (uint8_t)loa_oracle_playout(states[3 * (size_t)i], states[3 * (size_t)i + 1], (int)(states[3 * (size_t)i + 2] & 1),
		                                         key, first_id + i, max_plies, &plies[i], &fw, &fb);
		final_states[3 * (size_t)i] = fw; final_states[3 * (size_t)i + 1] = fb;
		final_states[3 * (size_t)i + 2] = (states[3 * (size_t)i + 2] + plies[i]) & 1;
	}
}
/* Synthetic "reachable" corpus (mirrors gai_loa_gen_reachable): state i = position after k_i
 * random plies from the opening, k_i = mulhi(philox(key ^ 0x9E3779B97F4A7C15, i, 0)[0], max_k); moves from
 * stream (key, id = i, ply = t); a terminal position restarts from the opening. */
void loa_oracle_gen_reachable(uint32_t n, uint64_t key, uint32_t max_k, uint64_t* states)
{
	uint8_t mv[2 * LOA_MAX_MOVES];
	for (uint32_t i = 0; i < n; ++i) {
		const uint32_t k = oracle_choice_index(oracle_choice_word(key ^ 0x9E3779B97F4A7C15ull, i, 0), max_k);
		uint64_t w, b; int cp;
		loa_oracle_initial(&w, &b, &cp);
		for (uint32_t t = 0; t < k; ++t) {
			if (loa_oracle_is_terminal(w, b)) loa_oracle_initial(&w, &b, &cp);
			const int cnt = loa_oracle_movegen(w, b, cp, mv);
			if (cnt) {
				const uint32_t idx = oracle_choice_index(oracle_choice_word(key, i, t), (uint32_t)cnt);
				loa_oracle_apply(w, b, cp, mv[2 * idx], mv[2 * idx + 1], &w, &b, &cp);
			} else cp ^= 1;
		}
		if (loa_oracle_is_terminal(w, b)) loa_oracle_initial(&w, &b, &cp);
		states[3 * (size_t)i] = w; states[3 * (size_t)i + 1] = b; states[3 * (size_t)i + 2] = (uint64_t)cp;
	}
}
/* "uniform" half of the C2 corpus: nw, nb uniform in [1,12], squares without overlap, side uniform;
 * all draws from philox(key ^ 0xC2B2AE3D27D4EB4F, i, block) words in order. */
void loa_oracle_gen_uniform(uint32_t n, uint64_t key, uint64_t* states)
{
	for (uint32_t i = 0; i < n; ++i) {
		uint32_t ply = 0;
#define NEXTW() oracle_choice_word(key ^ 0xC2B2AE3D27D4EB4Full, i, ply++)
		const uint32_t nw = 1 + oracle_choice_index(NEXTW(), 12), nb = 1 + oracle_choice_index(NEXTW(), 12);
		const uint32_t cp = oracle_choice_index(NEXTW(), 2);
		uint64_t occ = 0, w = 0, b = 0;
		for (uint32_t j = 0; j < nw + nb; ++j) {
			/* j-th piece goes on the r-th still-empty square */
			uint32_t r = oracle_choice_index(NEXTW(), 64 - j);
			int sqi = 0;
			for (;; ++sqi) { if (!((occ >> sqi) & 1)) { if (r == 0) break; --r; } }
			occ |= 1ull << sqi;
			if (j < nw) w |= 1ull << sqi; else b |= 1ull << sqi;
		}
#undef NEXTW
		states[3 * (size_t)i] = w; states[3 * (size_t)i + 1] = b; states[3 * (size_t)i + 2] = cp;
	}
}
