// game-ai_b200/host/loa_rules.cpp -- Lines of Action rules plugin (host side), exports createGameRules.
//
// Drop-in for the reference's LinesOfActionZasady.dll (c++/LinesOfActionZasady/LinesOfActionZasady.cpp:333-551):
// same IGameRules behaviour -- move order (:167-232), noop list recognised by address (:448-454,484-487),
// ownership rules, Score (:420-432), constant 50/50 eval (:382-397), ToString layout (:253-270) -- written from
// scratch for Linux (no Boost, no MSVC intrinsics).  GameState keeps the reference's 24-byte layout (:38-46) so that
// GetStateHash() hands the GPU shim (gameai_gpu.h gai_loa_state) the bytes it expects.  The host tree of the MCTS
// player uses these rules for selection/expansion; the playouts themselves never run here.
#include "gameai/GameRules.h"
#include <cstring>
#include <sstream>
#include <vector>

struct Move { uint8_t from, to; };
struct MoveList { uint16_t size; Move move[1]; };
struct GameState { uint64_t white, black, current_player; };      // only bit 0 of current_player is meaningful

namespace {

// line masks through every square: [0] row (step 1), [1] column (step 8), [2] diagonal A1->H8 (step 9), [3] diagonal H1->A8 (step 7)
// (LinesOfActionZasady.cpp:82-135); the MCTS player's host tree calls GetPlayerLegalMoves once per expanded node
struct LineMasks {
	uint64_t m[4][64];
	LineMasks()
	{
		for (int p = 0; p < 64; ++p) {
			const int r = p >> 3, c = p & 7;
			uint64_t row = 0, col = 0, ld = 0, rd = 0;
			for (int k = 0; k < 8; ++k) {
				row |= 1ull << (r * 8 + k); col |= 1ull << (k * 8 + c);
				const int r1 = k + (r - c), r2 = (r + c) - k;                  // column k of the two diagonals
				if (r1 >= 0 && r1 < 8) ld |= 1ull << (r1 * 8 + k);
				if (r2 >= 0 && r2 < 8) rd |= 1ull << (r2 * 8 + k);
			}
			m[0][p] = row; m[1][p] = col; m[2][p] = ld; m[3][p] = rd;
		}
	}
};
const LineMasks g_lines;

// the two moves of the piece on `pos` along one line: distance = pieces of both colours on the line (:191-192); the target
// must be on the line and not an own piece (:142), and no ENEMY piece may stand strictly between (:146-148)
inline void line_moves(uint64_t mine, uint64_t theirs, uint64_t L, int step, int pos, bool up_first, Move* out, int& n)
{
	const int d = __builtin_popcountll((mine | theirs) & L) * step;
	const int tu = pos + d, td = pos - d;
	const bool up = tu < 64 && ((L & ~mine) >> tu & 1) && (theirs & L & ((1ull << tu) - (2ull << pos))) == 0;
	const bool dn = td >= 0 && ((L & ~mine) >> td & 1) && (theirs & L & ((1ull << pos) - (2ull << td))) == 0;
	if (up_first) { if (up) { out[n].from = (uint8_t)pos; out[n].to = (uint8_t)tu; ++n; } if (dn) { out[n].from = (uint8_t)pos; out[n].to = (uint8_t)td; ++n; } }
	else { if (dn) { out[n].from = (uint8_t)pos; out[n].to = (uint8_t)td; ++n; } if (up) { out[n].from = (uint8_t)pos; out[n].to = (uint8_t)tu; ++n; } }
}
int generate(const GameState& s, Move* out)
{
	const bool wtm = (s.current_player & 1) == 0;
	const uint64_t mine = wtm ? s.white : s.black, theirs = wtm ? s.black : s.white;
	int n = 0;
	for (uint64_t rest = mine; rest; rest &= rest - 1) {                   // ascending squares; W E | N S | SW NE | SE NW (:194-223)
		const int pos = __builtin_ctzll(rest);
		line_moves(mine, theirs, g_lines.m[0][pos], 1, pos, false, out, n);
		line_moves(mine, theirs, g_lines.m[1][pos], 8, pos, true, out, n);
		line_moves(mine, theirs, g_lines.m[2][pos], 9, pos, false, out, n);
		line_moves(mine, theirs, g_lines.m[3][pos], 7, pos, false, out, n);
	}
	return n;
}
bool connected(uint64_t b)
{
	if (!b) return false;
	uint64_t g = b & (~b + 1);
	for (;;) {
		uint64_t h = g | ((g << 1) & 0xfefefefefefefefeull) | ((g >> 1) & 0x7f7f7f7f7f7f7f7full);
		h = (h | (h << 8) | (h >> 8)) & b;
		if (h == g) break;
		g = h;
	}
	return g == b;
}

struct LoaRules : IGameRules
{
	int refcnt = 1;
	MoveList noop;

	LoaRules() { noop.size = 1; noop.move[0].from = 0; noop.move[0].to = 0; }
	~LoaRules() override {}

	// plain heap allocation: malloc is thread safe, so several search threads (one tree per GPU) may share one rules object
	GameState* alloc_state() { return new GameState(); }
	static MoveList* alloc_list(int n)
	{
		auto* ml = (MoveList*)::operator new(sizeof(uint16_t) + sizeof(Move) * (size_t)(n > 0 ? n : 1));
		ml->size = (uint16_t)n;
		return ml;
	}

	void SetRandomGenerator(IRandomGenerator*) override {}
	GameState* CreateRandomInitialState(IRandomGenerator*) override
	{
		GameState* s = alloc_state();
		s->black = 0x7e0000000000007eull; s->white = 0x0081818181818100ull; s->current_player = 1;    // Blacks move first
		return s;
	}
	GameState* CreateInitialStateFromHash(const uint32_t* h) override { GameState* s = alloc_state(); memcpy(s, h, sizeof(GameState)); s->current_player &= 1; return s; }
	GameState* CreateStateFromString(const wstring&) override { throw "not implemented"; }
	GameState* CreateStateFromString(const string&) override { throw "not implemented"; }
	GameState* CreatePlayerKnownState(const GameState* gs, int) override { return CopyGameState(gs); }
	EvalFunction_t CreateEvalFunction(const string&) override { return [](const GameState*, int v[]) { v[0] = v[1] = 50; }; }
	void UpdatePlayerKnownState(GameState* pks, const GameState* cgs, const std::vector<MoveList*>&) override { *pks = *cgs; }
	GameState* CopyGameState(const GameState* s) override { GameState* n = alloc_state(); *n = *s; n->current_player &= 1; return n; }
	bool AreEqual(const GameState* a, const GameState* b) override
	{ return a->white == b->white && a->black == b->black && ((a->current_player ^ b->current_player) & 1) == 0; }   // not memcmp: SURVEY 8c
	void ReleaseGameState(GameState* s) override { delete s; }
	bool IsTerminal(const GameState* s) override { return connected(s->white) || connected(s->black); }
	void Score(const GameState* s, int score[]) override
	{
		const bool w = connected(s->white), b = connected(s->black);
		if (w && b) { score[0] = score[1] = 50; return; }
		score[0] = w ? 100 : 0; score[1] = b ? 100 : 0;
	}
	int GetCurrentPlayer(const GameState* s) override { return (int)(s->current_player & 1); }
	MoveList* GetPlayerLegalMoves(const GameState* gs, int player) override
	{
		if (player != (int)(gs->current_player & 1)) return &noop;
		Move buf[96 + 8];
		const int n = generate(*gs, buf);
		MoveList* ml = alloc_list(n);
		memcpy(ml->move, buf, sizeof(Move) * (size_t)n);
		return ml;
	}
	void ReleaseMoveList(MoveList* ml) override { if (ml != &noop) ::operator delete(ml); }
	int GetNumMoves(const MoveList* ml) override { return ml->size; }
	std::tuple<Move*, float> GetMoveFromList(MoveList* ml, int idx) override { return { &ml->move[idx], 1.0f }; }
	MoveList* SelectMoveFromList(const MoveList* ml, int idx) override
	{
		if (ml == &noop) return const_cast<MoveList*>(ml);
		MoveList* one = alloc_list(1); one->move[0] = ml->move[idx]; return one;
	}
	GameState* ApplyMove(const GameState* gs, Move* mv, int player) override
	{
		GameState* ns = CopyGameState(gs);
		if (mv != noop.move) {
			const uint64_t f = 1ull << mv->from, t = 1ull << mv->to;
			if (player == 0) { ns->white = (ns->white & ~f) | t; ns->black &= ~t; }
			else { ns->black = (ns->black & ~f) | t; ns->white &= ~t; }
		}
		ns->current_player = (ns->current_player + 1) & 1;
		return ns;
	}
	GameState* Next(const GameState* s, const std::vector<MoveList*>& moves) override
	{
		const uint64_t cp = s->current_player & 1;
		GameState* cs = const_cast<GameState*>(s);
		for (int p = 0; p < 2; ++p) { GameState* ns = ApplyMove(cs, &moves[p]->move[0], p); ReleaseGameState(cs); cs = ns; }
		cs->current_player = (cp + 1) & 1;
		return cs;
	}
	const uint32_t* GetStateHash(const GameState* gs) override { return reinterpret_cast<const uint32_t*>(gs); }
	size_t GetStateHashSize() override { return sizeof(GameState) / 4; }
	string ToString(const GameState* gs) override
	{
		std::ostringstream ss;
		for (int bit = 63; bit >= 0; --bit) {
			const uint64_t m = 1ull << bit;
			ss << ((gs->white & m) ? 'o' : (gs->black & m) ? '*' : '.');
			if (bit % 8 == 0) ss << '\n';
		}
		ss << "cp=" << (gs->current_player & 1);
		return ss.str();
	}
	wstring ToWString(const GameState* gs) override { const string s = ToString(gs); return wstring(s.begin(), s.end()); }
	string ToString(const Move* mv) override
	{
		std::ostringstream ss;
		ss << (char)('A' + mv->from % 8) << (mv->from / 8 + 1) << "->" << (char)('A' + mv->to % 8) << (mv->to / 8 + 1);
		return ss.str();
	}
	wstring ToWString(const Move* mv) override { const string s = ToString(mv); return wstring(s.begin(), s.end()); }
	void AddRef() override { ++refcnt; }
	void Release() override { if (--refcnt == 0) delete this; }
};

} // namespace

extern "C" IGameRules* createGameRules(int /*number_of_players*/) { return new LoaRules(); }

// ---- flat C entry points over the plugin (used by tests/ to compare this host rules object with the oracle) ----
extern "C" int loa_host_movegen(uint64_t w, uint64_t b, int cp, uint8_t* out)
{
	static thread_local LoaRules R;
	GameState s{ w, b, (uint64_t)cp };
	MoveList* ml = R.GetPlayerLegalMoves(&s, cp & 1);
	const int n = R.GetNumMoves(ml);
	for (int i = 0; i < n; ++i) { out[2 * i] = ml->move[i].from; out[2 * i + 1] = ml->move[i].to; }
	R.ReleaseMoveList(ml);
	return n;
}
extern "C" int loa_host_terminal_score(uint64_t w, uint64_t b, int* score)
{
	static thread_local LoaRules R;
	GameState s{ w, b, 0 };
	R.Score(&s, score);
	return R.IsTerminal(&s) ? 1 : 0;
}
extern "C" void loa_host_next(uint64_t w, uint64_t b, int cp, int from, int to, uint64_t* out3)
{
	static thread_local LoaRules R;
	GameState* s = R.alloc_state(); s->white = w; s->black = b; s->current_player = (uint64_t)cp;
	MoveList* mine = R.alloc_list(1); mine->move[0].from = (uint8_t)from; mine->move[0].to = (uint8_t)to;
	std::vector<MoveList*> mv(2, &R.noop); mv[cp & 1] = mine;
	GameState* n = R.Next(s, mv);
	out3[0] = n->white; out3[1] = n->black; out3[2] = n->current_player;
	R.ReleaseMoveList(mine); R.ReleaseGameState(n);
}
