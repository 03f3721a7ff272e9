// game-ai_b200/host/pan_rules.cpp -- Pan ("Gra w Pana") rules plugin (host side), exports createGameRules.
//
// Drop-in for the reference's GraWPanaZasadyV2.dll (c++/GraWPanaZasadyV2/GraWPanaZasadyV2.cpp:75-719; line numbers
// after iconv to UTF-8): same IGameRules behaviour and the same 40-byte GameState / 8-byte Move bit layouts (:16-49), so
// GetStateHash() hands the GPU shim (gameai_gpu.h gai_pan_state) the words it expects.  Handles the full 2-bit
// "certainty" card encoding (player-known states, :214-243); the GPU kernels only ever see determinized states.
#include "gameai/GameRules.h"
#include "gameai/random_generator.h"
#include <cstring>
#include <sstream>
#include <vector>

struct GameState { uint64_t w[5]; };       // w[0] = stack:48 | current_player:3 << 48 | is_terminal:1 << 51 ; w[1+p] = count:4 | cards:48 << 4
struct Move { uint64_t w; };               // cards:48 | count:3 << 48 | operation:3 << 51 | probability:2 << 54
struct MoveList { uint32_t size; uint32_t pad; Move move[1]; };

namespace {

constexpr uint64_t M48 = 0xFFFFFFFFFFFFull;
enum { OP_NOOP = 0, OP_TAKE = 1, OP_PLAY = 2 };

inline uint64_t stack_of(const GameState* s) { return s->w[0] & M48; }
inline int cp_of(const GameState* s) { return (int)((s->w[0] >> 48) & 7); }
inline bool term_of(const GameState* s) { return (s->w[0] >> 51) & 1; }
inline uint64_t cards_of(const GameState* s, int p) { return (s->w[1 + p] >> 4) & M48; }
inline int count_of(const GameState* s, int p) { return (int)(s->w[1 + p] & 15); }
inline void set_hand(GameState* s, int p, uint64_t cards, int count) { s->w[1 + p] = (uint64_t)(count & 15) | ((cards & M48) << 4); }
inline void set_head(GameState* s, uint64_t stack, int cp, bool term) { s->w[0] = (stack & M48) | ((uint64_t)(cp & 7) << 48) | ((uint64_t)(term ? 1 : 0) << 51); }
inline Move mk(uint64_t cards, int count, int op, uint64_t prob) { return Move{ (cards & M48) | ((uint64_t)(count & 7) << 48) | ((uint64_t)op << 51) | ((prob & 3) << 54) }; }

uint64_t make11(uint64_t c) { const uint64_t any = (c | (c >> 1)) & 0x5555555555555555ull; return any | (any << 1); }      // :370-375
int top_card(uint64_t stack) { return (63 - __builtin_clzll(stack)) / 2; }
uint64_t lowest_out_of(uint64_t hand, int cnt) { uint64_t lo = 3; while (cnt--) { if ((hand & 3) < lo) lo = hand & 3; hand >>= 2; } return lo; }   // :399-408
int count_cards(uint64_t c) { return __builtin_popcountll(make11(c)) / 2; }                                                  // :519-531

struct PanRules : IGameRules
{
	const int np; int refcnt = 1;
	MoveList noop, empty;
	explicit PanRules(int n) : np(n) { noop.size = 1; noop.move[0] = mk(0, 0, OP_NOOP, 0); empty.size = 0; }

	static MoveList* alloc_list(const std::vector<Move>& mv)
	{
		auto* ml = (MoveList*)::operator new(sizeof(MoveList) + sizeof(Move) * (mv.size() ? mv.size() - 1 : 0));
		ml->size = (uint32_t)mv.size(); ml->pad = 0;
		for (size_t i = 0; i < mv.size(); ++i) ml->move[i] = mv[i];
		return ml;
	}
	bool check_terminal(const GameState* s) const { int with = 0; for (int p = 0; p < np; ++p) with += count_of(s, p) > 0; return with == 1; }    // :339-346
	int next_player(const GameState* s, int cur) const { int n = cur; for (int i = 0; i < np - 1; ++i) { n = (n + 1) % np; if (count_of(s, n) > 0) break; } return n; }   // :510-518
	void apply_inplace(GameState* s, const Move* m, int player) const                       // :534-557
	{
		const uint64_t cards = m->w & M48; const int cnt = (int)((m->w >> 48) & 7), op = (int)((m->w >> 51) & 7);
		if (op == OP_TAKE) { set_head(s, stack_of(s) & ~cards, cp_of(s), term_of(s)); set_hand(s, player, cards_of(s, player) | cards, count_of(s, player) + cnt); }
		else if (op == OP_PLAY) {
			set_head(s, stack_of(s) | cards, cp_of(s), term_of(s));
			for (int p = 0; p < np; ++p) set_hand(s, p, cards_of(s, p) & ~cards, count_of(s, p) - (p == player ? cnt : 0));
		}
	}

	void SetRandomGenerator(IRandomGenerator*) override {}
	GameState* CreateRandomInitialState(IRandomGenerator* rng) override                         // :102-133
	{
		int cards[24]; for (int i = 0; i < 24; ++i) cards[i] = i;
		const std::vector<int> sh = rng->generateUniform(0, 23, 40);
		for (size_t i = 0; i + 1 < sh.size(); i += 2) std::swap(cards[sh[i]], cards[sh[i + 1]]);
		auto* s = new GameState(); memset(s, 0, sizeof *s);
		uint64_t hands[4] = { 0, 0, 0, 0 }; int p = 0;
		for (int i = 0; i < 24; ++i) { hands[p] |= 3ull << (2 * cards[i]); p = (p + 1) % np; }
		for (p = 0; p < np; ++p) set_hand(s, p, hands[p], 24 / np);
		for (p = 0; p < np; ++p) if (hands[p] & 3) break;                                     // the holder of the 9 of hearts starts
		set_head(s, 0, p, false); set_head(s, 0, p, check_terminal(s));
		return s;
	}
	GameState* CreateInitialStateFromHash(const uint32_t* h) override { auto* s = new GameState(); memcpy(s, h, sizeof *s); set_head(s, stack_of(s), cp_of(s), check_terminal(s)); return s; }
	GameState* CreateStateFromString(const wstring&) override { throw "not implemented"; }
	GameState* CreateStateFromString(const string&) override { throw "not implemented"; }
	GameState* CreatePlayerKnownState(const GameState* c, int player) override                   // :214-243
	{
		static const uint64_t mask_by_np[5] = { 3, 3, 3, 2, 1 };
		auto* k = new GameState(); memset(k, 0, sizeof *k);
		uint64_t own = 0, other = 0;
		for (int ci = 0; ci < 24; ++ci) { if (cards_of(c, player) & (3ull << (2 * ci))) own |= 3ull << (2 * ci); else other |= mask_by_np[np] << (2 * ci); }
		for (int p = 0; p < np; ++p) set_hand(k, p, p == player ? own : other, count_of(c, p));
		set_head(k, stack_of(c), cp_of(c), term_of(c));
		return k;
	}
	EvalFunction_t CreateEvalFunction(const string& name) override                              // :255-311
	{
		const int n = np;
		if (name == "num_cards_weighted")
			return [n](const GameState* s, int v[]) {
				static const float prob[4] = { 0.f, 1.f / 3.f, 0.5f, 1.f };
				for (int p = 0; p < n; ++p) { float sum = 0; for (int j = 0; j < 24; ++j) { const int c = (int)((cards_of(s, p) >> (2 * j)) & 3); if (c) sum += (float)(6 - j / 4) * prob[c]; } v[p] = (int)(84.0f - sum); }
			};
		return [n](const GameState* s, int v[]) { for (int p = 0; p < n; ++p) v[p] = 2 * (24 - count_of(s, p)); };
	}
	void UpdatePlayerKnownState(GameState* pks, const GameState* next, const std::vector<MoveList*>& moves) override     // :313-324
	{
		for (int p = 0; p < np; ++p) apply_inplace(pks, &moves[p]->move[0], p);
		set_head(pks, stack_of(pks), cp_of(next), term_of(next));
	}
	GameState* CopyGameState(const GameState* s) override { auto* n = new GameState(*s); return n; }
	bool AreEqual(const GameState* a, const GameState* b) override { return memcmp(a, b, sizeof(GameState)) == 0; }
	void ReleaseGameState(GameState* s) override { delete s; }
	bool IsTerminal(const GameState* s) override { return term_of(s); }
	void Score(const GameState* s, int score[]) override                                        // :347-363
	{
		if (term_of(s)) { const int win = 100 / (np - 1); for (int p = 0; p < np; ++p) score[p] = count_of(s, p) != 0 ? 0 : win; }
		else { const int draw = 100 / np; for (int p = 0; p < np; ++p) score[p] = draw; }
	}
	int GetCurrentPlayer(const GameState* s) override { return cp_of(s); }
	MoveList* GetPlayerLegalMoves(const GameState* s, int player) override                      // :409-475
	{
		if (term_of(s)) return &empty;
		if (player != cp_of(s)) return &noop;
		std::vector<Move> mv;
		const uint64_t hand = cards_of(s, player), stack = stack_of(s);
		if (stack == 0) {
			if (hand & 3) mv.push_back(mk(3, 1, OP_PLAY, hand & 3));
			if ((make11(hand) & 0xFF) == 0xFF) mv.push_back(mk(0xFF, 4, OP_PLAY, lowest_out_of(hand, 4)));
			return alloc_list(mv);
		}
		if (stack <= 3 && (make11(hand) & 0xFC) == 0xFC) mv.push_back(mk(0xFC, 3, OP_PLAY, lowest_out_of(hand >> 2, 3)));
		const int top = top_card(stack);
		const uint64_t above = top + 1 >= 24 ? 0 : ((~0ull << (2 * (top + 1))) & M48);
		uint64_t quad = 0xFFull << (8 * (top / 4));
		uint64_t masked = hand & (above | (~stack & quad));
		while (masked) {
			const uint64_t from_quad = quad & masked;
			if (from_quad) {
				const int idx = __builtin_ctzll(from_quad) & ~1;
				mv.push_back(mk(3ull << idx, 1, OP_PLAY, (from_quad >> idx) & 3));
				if (make11(from_quad) == quad) mv.push_back(mk(0xFFull << idx, 4, OP_PLAY, lowest_out_of(from_quad >> idx, 4)));
			}
			masked &= ~quad; quad <<= 8;
		}
		uint64_t take = 0, st = stack;
		for (int t = 0; t < 3 && st > 3; ++t) { const uint64_t tos = 3ull << (2 * top_card(st)); take |= tos; st &= ~tos; }
		if (take) mv.push_back(mk(take, count_cards(take), OP_TAKE, 3));
		return alloc_list(mv);
	}
	void ReleaseMoveList(MoveList* ml) override { if (ml != &noop && ml != &empty) ::operator delete(ml); }
	int GetNumMoves(const MoveList* ml) override { return (int)ml->size; }
	std::tuple<Move*, float> GetMoveFromList(MoveList* ml, int idx) override
	{
		static const float fprob[4] = { 0.0f, 1.0f / 3.0f, 0.5f, 1.0f };
		return { &ml->move[idx], fprob[(ml->move[idx].w >> 54) & 3] };
	}
	MoveList* SelectMoveFromList(const MoveList* ml, int idx) override { return alloc_list({ ml->move[idx] }); }
	GameState* ApplyMove(const GameState* s, Move* m, int player) override                       // :560-569
	{
		auto* ns = new GameState(*s);
		apply_inplace(ns, m, player);
		const int nxt = next_player(ns, cp_of(s));
		set_head(ns, stack_of(ns), nxt, false); set_head(ns, stack_of(ns), nxt, check_terminal(ns));
		return ns;
	}
	GameState* Next(const GameState* s, const std::vector<MoveList*>& moves) override            // :570-586
	{
		const int cur = cp_of(s);
		GameState* cs = const_cast<GameState*>(s);
		for (int p = 0; p < np; ++p) { GameState* ns = ApplyMove(cs, &moves[p]->move[0], p); delete cs; cs = ns; }
		set_head(cs, stack_of(cs), next_player(cs, cur), term_of(cs));
		return cs;
	}
	const uint32_t* GetStateHash(const GameState* s) override { return reinterpret_cast<const uint32_t*>(s); }
	size_t GetStateHashSize() override { return sizeof(GameState) / 4; }      // the full 40 bytes (the reference reports numPlayers+1 words, :79)
	static void hand_str(uint64_t hand, std::ostringstream& ss)
	{
		static const char suits[] = "hcsd"; static const char* values[] = { "9", "10", "W", "D", "K", "A" };
		for (int v = 0, pos = 0; v < 6; ++v) for (int c = 0; c < 4; ++c, pos += 2) { const uint64_t p = (hand >> pos) & 3; if (p) ss << values[v] << '.' << p << suits[c]; }
	}
	string ToString(const GameState* s) override
	{
		std::ostringstream ss; ss << "S="; hand_str(stack_of(s), ss); ss << "|";
		for (int p = 0; p < np; ++p) { ss << "P" << p << "="; hand_str(cards_of(s, p), ss); ss << "|"; }
		ss << "CP=" << cp_of(s);
		return ss.str();
	}
	wstring ToWString(const GameState* s) override { const string t = ToString(s); return wstring(t.begin(), t.end()); }
	string ToString(const Move* m) override
	{
		std::ostringstream ss; const int op = (int)((m->w >> 51) & 7);
		if (op == OP_NOOP) { ss << "noop"; return ss.str(); }
		ss << (op == OP_PLAY ? "play " : "take "); hand_str(m->w & M48, ss);
		return ss.str();
	}
	wstring ToWString(const Move* m) override { const string t = ToString(m); return wstring(t.begin(), t.end()); }
	void AddRef() override { ++refcnt; }
	void Release() override { if (--refcnt == 0) delete this; }
};

} // namespace

extern "C" IGameRules* createGameRules(int number_of_players) { return new PanRules(number_of_players); }

// ---- flat C entry points over the plugin (tests compare this host rules object with the oracle) -----------------
extern "C" int pan_host_movegen(int np, const uint64_t* st, uint64_t* moves)
{
	PanRules R(np); GameState s; memcpy(&s, st, 40);
	MoveList* ml = R.GetPlayerLegalMoves(&s, cp_of(&s));
	const int n = R.GetNumMoves(ml);
	for (int i = 0; i < n; ++i) moves[i] = ml->move[i].w;
	R.ReleaseMoveList(ml);
	return n;
}
extern "C" void pan_host_apply(int np, const uint64_t* st, uint64_t mv, uint64_t* out)
{
	PanRules R(np); GameState s; memcpy(&s, st, 40); Move m{ mv };
	GameState* ns = R.ApplyMove(&s, &m, cp_of(&s));
	memcpy(out, ns, 40); R.ReleaseGameState(ns);
}
extern "C" void pan_host_known_state(int np, const uint64_t* st, int player, uint64_t* out)
{
	PanRules R(np); GameState s; memcpy(&s, st, 40);
	GameState* k = R.CreatePlayerKnownState(&s, player);
	memcpy(out, k, 40); R.ReleaseGameState(k);
}
extern "C" int pan_host_score_eval(int np, const uint64_t* st, int* score, int* ev0, int* ev1)
{
	PanRules R(np); GameState s; memcpy(&s, st, 40);
	set_head(&s, stack_of(&s), cp_of(&s), R.check_terminal(&s));
	R.Score(&s, score); R.CreateEvalFunction("num_cards")(&s, ev0); R.CreateEvalFunction("num_cards_weighted")(&s, ev1);
	return term_of(&s) ? 1 : 0;
}
