// oracle/ref_pan_capi.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the reference's OWN Pan rules (c++/GraWPanaZasadyV2/GraWPanaZasadyV2.cpp, UTF-16 in the reference tree;
// the Makefile transcodes it with iconv into the git-ignored oracle/_ref/ and this file #includes that copy, the way
// the reference's unit test includes the .cpp, UniTestsWithBoost/game_rules_v2.cpp:11) and wraps it in a flat C API.
// States cross the API as the raw 40-byte GameState (5 x uint64), moves as the raw 8-byte Move.
#include REF_PAN_CPP
#include "philox_ref.h"

namespace {
using GraWPanaV2::GraWPanaGameRules;
struct PanRules {
	GraWPanaGameRules* r[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
	GraWPanaGameRules* get(int np) { if (!r[np]) r[np] = new GraWPanaGameRules(np); return r[np]; }
};
PanRules& pan() { static thread_local PanRules P; return P; }

// Philox-driven IRandomGenerator (random_generator.h:4-8) for the deal (GraWPanaZasadyV2.cpp:102-114)
struct PhiloxGen : IRandomGenerator {
	uint64_t key, id; uint32_t n = 0;
	PhiloxGen(uint64_t k, uint64_t i) : key(k), id(i) {}
	std::vector<int> generateUniform(int lower, int upper, int number_of_samples) override
	{
		std::vector<int> v(number_of_samples);
		for (auto& x : v) x = lower + (int)oracle_choice_index(oracle_choice_word(key, id, n++), (uint32_t)(upper - lower + 1));
		return v;
	}
	void release() override {}
};
void load(GameState& s, const uint64_t* w) { memcpy(&s, w, 40); }
void store(uint64_t* w, const GameState& s) { memcpy(w, &s, 40); }
}

extern "C" {

int ref_pan_sizeof_state() { return (int)sizeof(GameState); }
int ref_pan_sizeof_move() { return (int)sizeof(Move); }

void ref_pan_deal(int np, uint64_t key, uint64_t id, uint64_t* out)
{
	auto* gr = pan().get(np);
	PhiloxGen g(key, id);
	GameState* s = gr->CreateRandomInitialState(&g);
	store(out, *s);
	gr->ReleaseGameState(s);
}
void ref_pan_player_known_state(int np, const uint64_t* st, int player, uint64_t* out)
{
	auto* gr = pan().get(np);
	GameState s; load(s, st);
	GameState* p = gr->CreatePlayerKnownState(&s, player);
	store(out, *p);
	gr->ReleaseGameState(p);
}
int ref_pan_movegen(int np, const uint64_t* st, uint64_t* moves)
{
	auto* gr = pan().get(np);
	GameState s; load(s, st);
	MoveList* ml = gr->GetPlayerLegalMoves(&s, (int)s.current_player);
	const int n = gr->GetNumMoves(ml);
	for (int i = 0; i < n; ++i) { auto [mv, p] = gr->GetMoveFromList(ml, i); memcpy(&moves[i], mv, 8); }
	gr->ReleaseMoveList(ml);
	return n;
}
void ref_pan_apply(int np, const uint64_t* st, uint64_t move, uint64_t* out)
{
	auto* gr = pan().get(np);
	GameState s; load(s, st);
	Move m; memcpy(&m, &move, 8);
	GameState* ns = gr->ApplyMove(&s, &m, (int)s.current_player);
	store(out, *ns);
	gr->ReleaseGameState(ns);
}
int ref_pan_check_terminal(int np, const uint64_t* st)
{
	GameState s; load(s, st);
	return pan().get(np)->checkIfTerminal(&s) ? 1 : 0;
}
void ref_pan_score(int np, const uint64_t* st, int* score)
{
	GameState s; load(s, st);
	pan().get(np)->Score(&s, score);
}
void ref_pan_eval(int np, int kind, const uint64_t* st, int* value)
{
	GameState s; load(s, st);
	auto f = pan().get(np)->CreateEvalFunction(kind == 1 ? "num_cards_weighted" : "num_cards");
	f(&s, value);
}
uint64_t ref_pan_make11(uint64_t cards) { return GraWPanaGameRules::make11(cards); }
void ref_pan_make_stack_masks(uint64_t stack, uint64_t* out3)
{
	auto [a, q, t] = GraWPanaGameRules::makeStackMasks(stack);
	out3[0] = a; out3[1] = q; out3[2] = t;
}
int ref_pan_state_to_string(int np, const uint64_t* st, char* out, int cap)
{
	GameState s; load(s, st);
	const std::string str = pan().get(np)->ToString(&s);
	const int n = (int)str.size() < cap - 1 ? (int)str.size() : cap - 1;
	memcpy(out, str.data(), n); out[n] = 0;
	return n;
}

// One Philox-driven uniformly random playout through the rules' virtual interface.
// Result code: 0..3 = the player left holding cards (the loser, Score :347-356); 4 = ply cap or dead end,
// value[] then holds the eval function's output (:255-311) at the last state.  plies = moves applied.
int ref_pan_playout(int np, const uint64_t* st, uint64_t key, uint64_t id, uint32_t max_plies, int eval_kind,
                    uint32_t* plies_out, int* value, uint64_t* final_state)
{
	auto* gr = pan().get(np);
	GameState s0; load(s0, st);
	GameState* gs = gr->CopyGameState(&s0);
	gs->is_terminal = gr->checkIfTerminal(gs);          // (the C API does not trust the caller's flag)
	uint32_t ply = 0; int code = 4;
	for (;;) {
		if (gr->IsTerminal(gs)) {
			for (int p = 0; p < np; ++p) if (gs->hand[p].count != 0) code = p;
			gr->Score(gs, value);
			break;
		}
		MoveList* ml = (ply >= max_plies) ? nullptr : gr->GetPlayerLegalMoves(gs, (int)gs->current_player);
		const int n = ml ? gr->GetNumMoves(ml) : 0;
		if (n == 0) {
			auto f = gr->CreateEvalFunction(eval_kind == 1 ? "num_cards_weighted" : "num_cards");
			f(gs, value);
			if (ml) gr->ReleaseMoveList(ml);
			break;
		}
		const uint32_t idx = oracle_choice_index(oracle_choice_word(key, id, ply), (uint32_t)n);
		auto [mv, p] = gr->GetMoveFromList(ml, (int)idx);
		GameState* ns = gr->ApplyMove(gs, mv, (int)gs->current_player);
		gr->ReleaseMoveList(ml);
		gr->ReleaseGameState(gs);
		gs = ns; ++ply;
	}
	if (plies_out) *plies_out = ply;
	if (final_state) store(final_state, *gs);
	gr->ReleaseGameState(gs);
	return code;
}

// batch forms -------------------------------------------------------------------------------------
void ref_pan_movegen_batch(int np, const uint64_t* states, uint32_t n, uint8_t* counts, uint64_t* moves)
{
	for (uint32_t i = 0; i < n; ++i) {
		memset(moves + (size_t)i * 16, 0, 128);
		counts[i] = (uint8_t)ref_pan_movegen(np, states + (size_t)i * 5, moves + (size_t)i * 16);
	}
}
void ref_pan_apply_batch(int np, const uint64_t* states, const uint64_t* mv, uint32_t n, uint64_t* out)
{
	for (uint32_t i = 0; i < n; ++i) ref_pan_apply(np, states + (size_t)i * 5, mv[i], out + (size_t)i * 5);
}
void ref_pan_terminal_batch(int np, const uint64_t* states, uint32_t n, uint8_t* terminal, int32_t* score, int32_t* eval0, int32_t* eval1)
{
	for (uint32_t i = 0; i < n; ++i) {
		uint64_t st[5]; memcpy(st, states + (size_t)i * 5, 40);
		terminal[i] = (uint8_t)ref_pan_check_terminal(np, st);
		// Score reads the state's own is_terminal flag: set it consistently first
		GameState s; load(s, st); s.is_terminal = terminal[i]; store(st, s);
		int v[4] = { 0, 0, 0, 0 };
		ref_pan_score(np, st, v); memcpy(score + (size_t)i * 4, v, 16);
		int a[4] = { 0, 0, 0, 0 }; ref_pan_eval(np, 0, st, a); memcpy(eval0 + (size_t)i * 4, a, 16);
		int b[4] = { 0, 0, 0, 0 }; ref_pan_eval(np, 1, st, b); memcpy(eval1 + (size_t)i * 4, b, 16);
	}
}
void ref_pan_playout_trace_batch(int np, const uint64_t* states, uint32_t n, uint32_t max_plies, int eval_kind, uint64_t key, uint64_t first_id,
                                 uint8_t* code, uint32_t* plies, int32_t* value, uint64_t* final_states)
{
	for (uint32_t i = 0; i < n; ++i) {
		int v[4] = { 0, 0, 0, 0 };
		code[i] = (uint8_t)ref_pan_playout(np, states + (size_t)i * 5, key, first_id + i, max_plies, eval_kind, &plies[i], v, final_states + (size_t)i * 5);
		memcpy(value + (size_t)i * 4, v, 16);
	}
}
// per leaf: res[i*12 + 0..3] losses by player, [4..7] eval sums of capped playouts, [8] capped count
uint64_t ref_pan_playout_batch(int np, const uint64_t* states, uint32_t n_leaves, uint32_t ppl, uint32_t max_plies, int eval_kind,
                               uint64_t key, uint64_t first_id, uint32_t* res)
{
	uint64_t total = 0;
	for (uint32_t i = 0; i < n_leaves; ++i)
		for (uint32_t j = 0; j < ppl; ++j) {
			uint32_t plies = 0; int v[4] = { 0, 0, 0, 0 };
			const int c = ref_pan_playout(np, states + (size_t)i * 5, key, first_id + (uint64_t)i * ppl + j, max_plies, eval_kind, &plies, v, nullptr);
			uint32_t* r = res + (size_t)i * 12;
			if (c < 4) r[c] += 1; else { r[8] += 1; for (int p = 0; p < np; ++p) r[4 + p] += (uint32_t)v[p]; }
			total += plies;
		}
	return total;
}
// corpus: state i = a fresh Philox deal (key, id = i) advanced by k_i random plies (k_i < max_k); restart on terminal
void ref_pan_gen_reachable(int np, uint32_t n, uint64_t key, uint32_t max_k, uint64_t* states)
{
	auto* gr = pan().get(np);
	for (uint32_t i = 0; i < n; ++i) {
		uint64_t st[5];
		ref_pan_deal(np, key ^ 0xD1B54A32D192ED03ull, i, st);
		const uint32_t k = oracle_choice_index(oracle_choice_word(key ^ 0x9E3779B97F4A7C15ull, i, 0), max_k);
		for (uint32_t t = 0; t < k; ++t) {
			if (ref_pan_check_terminal(np, st)) { ref_pan_deal(np, key ^ 0xD1B54A32D192ED03ull, i, st); }
			uint64_t mv[16];
			const int cnt = ref_pan_movegen(np, st, mv);
			if (cnt == 0) break;
			const uint32_t idx = oracle_choice_index(oracle_choice_word(key, i, t), (uint32_t)cnt);
			uint64_t ns[5]; ref_pan_apply(np, st, mv[idx], ns); memcpy(st, ns, 40);
		}
		if (ref_pan_check_terminal(np, st)) ref_pan_deal(np, key ^ 0xD1B54A32D192ED03ull, i, st);
		memcpy(states + (size_t)i * 5, st, 40);
	}
	(void)gr;
}

} // extern "C"
