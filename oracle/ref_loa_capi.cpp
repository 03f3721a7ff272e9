// oracle/ref_loa_capi.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the reference's OWN Lines-of-Action rules (c++/LinesOfActionZasady/LinesOfActionZasady.cpp)
// verbatim -- by #include from where it lies under /root/reference, exactly as the reference's unit
// test does (LinesOfActionZasady_ut.cpp:10) -- and wraps it in a flat C API for ctypes.  No reference
// source is copied into this repository; the output goes to the git-ignored oracle/_ref/.
//
// Every call below goes through the IGameRules virtuals (GameRules.h:14-50) the way MC::Player does
// (MCTSPlayer.cpp:387-470), except the static mask helpers which the reference's tests call directly.
#include REF_LOA_CPP   // -DREF_LOA_CPP='"/root/reference/c++/LinesOfActionZasady/LinesOfActionZasady.cpp"'
#include "philox_ref.h"

namespace {
struct Rules {
	LinesOfAction::LinesOfActionGameRules* r;
	Rules() : r(new LinesOfAction::LinesOfActionGameRules()) { r->m_RefCnt = 1; /* never initialised in the reference, :336 */ }
};
Rules& rules() { static thread_local Rules R; return R; }
}

extern "C" {

int ref_loa_sizeof_state() { return (int)sizeof(GameState); }

void ref_loa_initial(uint64_t* w, uint64_t* b, int* cp)
{
	auto* gr = rules().r;
	GameState* gs = gr->CreateRandomInitialState(nullptr);
	*w = gs->white; *b = gs->black; *cp = gr->GetCurrentPlayer(gs);
	gr->ReleaseGameState(gs);
}

// legal moves of the side to move, in the reference's order; out = from,to byte pairs (<= 96 moves)
int ref_loa_movegen(uint64_t w, uint64_t b, int cp, uint8_t* out)
{
	auto* gr = rules().r;
	GameState gs(w, b, cp);
	MoveList* ml = gr->GetPlayerLegalMoves(&gs, cp);
	const int n = gr->GetNumMoves(ml);
	for (int i = 0; i < n; ++i) {
		auto [mv, p] = gr->GetMoveFromList(ml, i);
		out[2 * i] = mv->from; out[2 * i + 1] = mv->to;
	}
	gr->ReleaseMoveList(ml);
	return n;
}

void ref_loa_apply(uint64_t w, uint64_t b, int cp, int from, int to, uint64_t* ow, uint64_t* ob, int* ocp)
{
	auto* gr = rules().r;
	GameState gs(w, b, cp);
	Move mv; mv.from = (uint8_t)from; mv.to = (uint8_t)to;
	GameState* ns = gr->ApplyMove(&gs, &mv, cp);
	*ow = ns->white; *ob = ns->black; *ocp = gr->GetCurrentPlayer(ns);
	gr->ReleaseGameState(ns);
}

int ref_loa_is_terminal(uint64_t w, uint64_t b)
{
	GameState gs(w, b, 0);
	return rules().r->IsTerminal(&gs) ? 1 : 0;
}

void ref_loa_score(uint64_t w, uint64_t b, int* score)
{
	GameState gs(w, b, 0);
	rules().r->Score(&gs, score);
}

uint64_t ref_loa_neighbour_mask(int pos)      { return GameState::getNeighbourMask(pos); }
uint64_t ref_loa_row_mask(int pos)            { return GameState::getRowMask(pos); }
uint64_t ref_loa_column_mask(int pos)         { return GameState::getColumnMask(pos); }
uint64_t ref_loa_left_diagonal_mask(int pos)  { return GameState::getLeftDiagonalMask(pos); }
uint64_t ref_loa_right_diagonal_mask(int pos) { return GameState::getRightDiagonalMask(pos); }
int ref_loa_calc_if_terminal(uint64_t board)  { return GameState::calcIfTerminal(board) ? 1 : 0; }
int ref_loa_num_groups(uint64_t board)        { uint8_t arr[64]; return GameState::calcNumGroups(board, arr); }

int ref_loa_state_to_string(uint64_t w, uint64_t b, int cp, char* out, int cap)
{
	GameState gs(w, b, cp);
	const std::string s = rules().r->ToString(&gs);
	const int n = (int)s.size() < cap - 1 ? (int)s.size() : cap - 1;
	memcpy(out, s.data(), n); out[n] = 0;
	return n;
}

// One Philox-driven uniformly random playout through the reference rules' virtual interface.
// Outcome: 0 = white wins, 1 = black wins, 2 = draw (both connected), 3 = ply cap reached.
// Semantics (SURVEY.md 8c "Known oracle hazards", last bullet): a leaf that is already terminal
// scores at 0 plies; a side with no legal move passes (noop, LinesOfActionZasady.cpp:484-487).
int ref_loa_playout(uint64_t w, uint64_t b, int cp, uint64_t key, uint64_t playout_id, uint32_t max_plies,
                    uint32_t* plies_out, uint64_t* fw, uint64_t* fb)
{
	auto* gr = rules().r;
	GameState start(w, b, cp);
	GameState* gs = gr->CopyGameState(&start);
	uint32_t ply = 0;
	int outcome = 3;
	for (;;) {
		if (gr->IsTerminal(gs)) {
			int sc[2]; gr->Score(gs, sc);
			outcome = (sc[0] == sc[1]) ? 2 : (sc[0] > sc[1] ? 0 : 1);
			break;
		}
		if (ply >= max_plies) { outcome = 3; break; }
		const int player = gr->GetCurrentPlayer(gs);
		MoveList* ml = gr->GetPlayerLegalMoves(gs, player);
		const int n = gr->GetNumMoves(ml);
		GameState* ns;
		if (n == 0) {
			// pass: apply the rules' noop move (the move stored in the shared noop list)
			MoveList* noop = gr->GetPlayerLegalMoves(gs, 1 - player);
			auto [mv, p] = gr->GetMoveFromList(noop, 0);
			ns = gr->ApplyMove(gs, mv, player);
		} else {
			const uint32_t r = oracle_choice_word(key, playout_id, ply);
			const uint32_t idx = oracle_choice_index(r, (uint32_t)n);
			auto [mv, p] = gr->GetMoveFromList(ml, (int)idx);
			ns = gr->ApplyMove(gs, mv, player);
		}
		gr->ReleaseMoveList(ml);
		gr->ReleaseGameState(gs);
		gs = ns;
		++ply;
	}
	if (plies_out) *plies_out = ply;
	if (fw) *fw = gs->white;
	if (fb) *fb = gs->black;
	gr->ReleaseGameState(gs);
	return outcome;
}

// Batch form used as tier-2 truth and as the "reference" CPU baseline: per leaf i, P playouts with
// ids first_id + i*P + j.  counts[i*4 + outcome] += 1.  Returns total plies.
uint64_t ref_loa_playout_batch(const uint64_t* whites, const uint64_t* blacks, const uint8_t* cps, uint32_t n_leaves,
                               uint32_t playouts_per_leaf, uint32_t max_plies, uint64_t key, uint64_t first_id,
                               uint32_t* counts)
{
	uint64_t total = 0;
	for (uint32_t i = 0; i < n_leaves; ++i)
		for (uint32_t j = 0; j < playouts_per_leaf; ++j) {
			uint32_t plies = 0;
			const int o = ref_loa_playout(whites[i], blacks[i], cps[i] & 1, key, first_id + (uint64_t)i * playouts_per_leaf + j,
			                              max_plies, &plies, nullptr, nullptr);
			counts[(size_t)i * 4 + o] += 1;
			total += plies;
		}
	return total;
}

// ---- batch forms over n x {white, black, current_player} uint64 triples ---------------------------
void ref_loa_movegen_batch(const uint64_t* states, uint32_t n, uint8_t* counts, uint8_t* moves)
{
	for (uint32_t i = 0; i < n; ++i) {
		uint8_t* out = moves + (size_t)i * 192;
		memset(out, 0, 192);
		counts[i] = (uint8_t)ref_loa_movegen(states[3 * (size_t)i], states[3 * (size_t)i + 1], (int)(states[3 * (size_t)i + 2] & 1), out);
	}
}
void ref_loa_terminal_batch(const uint64_t* states, uint32_t n, uint8_t* terminal, uint8_t* score)
{
	for (uint32_t i = 0; i < n; ++i) {
		int sc[2];
		terminal[i] = (uint8_t)ref_loa_is_terminal(states[3 * (size_t)i], states[3 * (size_t)i + 1]);
		ref_loa_score(states[3 * (size_t)i], states[3 * (size_t)i + 1], sc);
		score[2 * (size_t)i] = (uint8_t)sc[0]; score[2 * (size_t)i + 1] = (uint8_t)sc[1];
	}
}
void ref_loa_apply_batch(const uint64_t* states, const uint8_t* mv, uint32_t n, uint64_t* out)
{
	auto* gr = rules().r;
	for (uint32_t i = 0; i < n; ++i) {
		const int cp = (int)(states[3 * (size_t)i + 2] & 1);
		GameState gs(states[3 * (size_t)i], states[3 * (size_t)i + 1], cp);
		const int from = mv[2 * (size_t)i], to = mv[2 * (size_t)i + 1];
		GameState* ns;
		if (from == to) {        // the C ABI's spelling of the noop move: hand the rules their own noop Move*
			MoveList* noop = gr->GetPlayerLegalMoves(&gs, 1 - cp);
			auto [m, p] = gr->GetMoveFromList(noop, 0);
			ns = gr->ApplyMove(&gs, m, cp);
		} else {
			Move m; m.from = (uint8_t)from; m.to = (uint8_t)to;
			ns = gr->ApplyMove(&gs, &m, cp);
		}
		out[3 * (size_t)i] = ns->white; out[3 * (size_t)i + 1] = ns->black; out[3 * (size_t)i + 2] = (uint64_t)gr->GetCurrentPlayer(ns);
		gr->ReleaseGameState(ns);
	}
}
static uint64_t mix64(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}
void ref_loa_successor_hash_batch(const uint64_t* states, uint32_t n, uint64_t* hash)
{
	uint8_t mv[192];
	for (uint32_t i = 0; i < n; ++i) {
		const uint64_t w = states[3 * (size_t)i], b = states[3 * (size_t)i + 1];
		const int cp = (int)(states[3 * (size_t)i + 2] & 1);
		const int cnt = ref_loa_movegen(w, b, cp, mv);
		uint64_t h = 0;
		for (int k = 0; k < cnt; ++k) {
			uint64_t nw, nb; int ncp;
			ref_loa_apply(w, b, cp, mv[2 * k], mv[2 * k + 1], &nw, &nb, &ncp);
			const uint64_t wc = (uint64_t)ref_loa_calc_if_terminal(nw), bc = (uint64_t)ref_loa_calc_if_terminal(nb);
			h = mix64(h ^ nw); h = mix64(h ^ nb);
			h = mix64(h ^ ((uint64_t)ncp | ((uint64_t)k << 8) | (wc << 16) | (bc << 17)));
		}
		hash[i] = h;
	}
}
void ref_loa_playout_trace_batch(const uint64_t* states, uint32_t n, uint32_t max_plies, uint64_t key, uint64_t first_id,
                                 uint8_t* outcome, uint32_t* plies, uint64_t* final_states)
{
	for (uint32_t i = 0; i < n; ++i) {
		uint64_t fw, fb;
		outcome[i] = (uint8_t)ref_loa_playout(states[3 * (size_t)i], states[3 * (size_t)i + 1], (int)(states[3 * (size_t)i + 2] & 1),
		                                      key, first_id + i, max_plies, &plies[i], &fw, &fb);
		final_states[3 * (size_t)i] = fw; final_states[3 * (size_t)i + 1] = fb;
		final_states[3 * (size_t)i + 2] = (states[3 * (size_t)i + 2] + plies[i]) & 1;
	}
}

} // extern "C"
