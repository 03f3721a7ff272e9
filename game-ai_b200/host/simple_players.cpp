// game-ai_b200/host/simple_players.cpp -- CPU opponents for config 1 (BASELINE.json), exports createPlayer.
//   type="random"  : uniformly random legal move (reference: SimpleStrategyPlayer.cpp:39-77 RandomPlayer)
//   type="minmax"  : depth-limited alpha-beta over IGameRules for two players with a state-keyed value cache
//                    (reference: MinMaxABPlayer.cpp:66-133; search_depth, eval_function from the config)
// These only talk to the game through IGameRules, like the reference players.
#include "gameai/GamePlayer.h"
#include "gameai/GameRules.h"
#include <chrono>
#include <random>
#include <string>
#include <unordered_map>

namespace {

struct RandomPlayer : IGamePlayer
{
	int me; IGameRules* rules = nullptr; std::mt19937 rng; std::string name;
	RandomPlayer(int p, unsigned seed, const std::string& n) : me(p), rng(seed), name(n) {}
	void startNewGame(GameState*) override {}
	void endGame(int, GameResult) override {}
	void setGameRules(IGameRules* gr) override { rules = gr; }
	MoveList* selectMove(GameState* pks) override
	{
		MoveList* ml = rules->GetPlayerLegalMoves(pks, me);
		if (rules->GetCurrentPlayer(pks) != me) return ml;
		const int n = rules->GetNumMoves(ml);
		if (n == 0) return ml;
		MoveList* sel = rules->SelectMoveFromList(ml, n > 1 ? std::uniform_int_distribution<int>(0, n - 1)(rng) : 0);
		rules->ReleaseMoveList(ml);
		return sel;
	}
	NamedMetrics_t getGameStats() override { return {}; }
	std::string getName() override { return name; }
	void resetStats() override {}
	void release() override { delete this; }
};

// always the first legal move: in Pan that is the lowest playable card (reference: SimpleStrategyPlayer.cpp:10-37)
struct LowCardPlayer : IGamePlayer
{
	int me; IGameRules* rules = nullptr;
	explicit LowCardPlayer(int p) : me(p) {}
	void startNewGame(GameState*) override {}
	void endGame(int, GameResult) override {}
	void setGameRules(IGameRules* gr) override { rules = gr; }
	MoveList* selectMove(GameState* pks) override
	{
		MoveList* ml = rules->GetPlayerLegalMoves(pks, me);
		if (rules->GetNumMoves(ml) == 0 || rules->GetCurrentPlayer(pks) != me) return ml;
		MoveList* sel = rules->SelectMoveFromList(ml, 0);
		rules->ReleaseMoveList(ml);
		return sel;
	}
	NamedMetrics_t getGameStats() override { return {}; }
	std::string getName() override { return "lowcard"; }
	void resetStats() override {}
	void release() override { delete this; }
};

struct MinMaxPlayer : IGamePlayer
{
	int me, depth; IGameRules* rules = nullptr; EvalFunction_t eval; std::string eval_name, name; std::mt19937 rng;
	std::unordered_map<std::string, std::pair<int, int>> cache;       // state string -> (depth, value for `me`), MinMaxABPlayer.cpp:66-80
	long nodes = 0; Average nodes_per_move;
	MinMaxPlayer(int p, int d, const std::string& ev, unsigned seed, const std::string& n) : me(p), depth(d), eval_name(ev), name(n), rng(seed) {}
	void startNewGame(GameState*) override {}
	void endGame(int, GameResult) override { cache.clear(); }
	void setGameRules(IGameRules* gr) override { rules = gr; eval = gr->CreateEvalFunction(eval_name); }
	int value_of(const GameState* s)
	{
		int sc[4] = { 0, 0, 0, 0 };
		if (rules->IsTerminal(s)) rules->Score(s, sc); else eval(s, sc);
		return sc[me] - sc[1 - me];
	}
	int search(const GameState* s, int d, int alpha, int beta)
	{
		++nodes;
		if (d == 0 || rules->IsTerminal(s)) return value_of(s);
		const std::string key = rules->ToString(s);
		auto it = cache.find(key);
		if (it != cache.end() && it->second.first >= d) return it->second.second;
		const int cp = rules->GetCurrentPlayer(s);
		MoveList* ml = rules->GetPlayerLegalMoves(s, cp);
		const int n = rules->GetNumMoves(ml);
		int best = cp == me ? -1000000 : 1000000;
		if (n == 0) best = value_of(s);
		for (int i = 0; i < n; ++i) {
			auto [mv, p] = rules->GetMoveFromList(ml, i); (void)p;
			GameState* ns = rules->ApplyMove(s, mv, cp);
			const int v = search(ns, d - 1, alpha, beta);
			rules->ReleaseGameState(ns);
			if (cp == me) { if (v > best) best = v; if (best > alpha) alpha = best; }
			else { if (v < best) best = v; if (best < beta) beta = best; }
			if (alpha >= beta) break;
		}
		rules->ReleaseMoveList(ml);
		cache[key] = { d, best };
		return best;
	}
	MoveList* selectMove(GameState* pks) override
	{
		MoveList* ml = rules->GetPlayerLegalMoves(pks, me);
		if (rules->GetCurrentPlayer(pks) != me) return ml;
		const int n = rules->GetNumMoves(ml);
		nodes = 0; cache.clear();
		int bestv = -1000001; std::vector<int> best;
		for (int i = 0; i < n; ++i) {
			auto [mv, p] = rules->GetMoveFromList(ml, i); (void)p;
			GameState* ns = rules->ApplyMove(pks, mv, me);
			const int v = search(ns, depth - 1, -1000000, 1000000);
			rules->ReleaseGameState(ns);
			if (v > bestv) { bestv = v; best.clear(); }
			if (v == bestv) best.push_back(i);
		}
		nodes_per_move.insert((double)nodes);
		const int pick = best.empty() ? 0 : best[best.size() == 1 ? 0 : std::uniform_int_distribution<size_t>(0, best.size() - 1)(rng)];
		MoveList* sel = rules->SelectMoveFromList(ml, pick);
		rules->ReleaseMoveList(ml);
		return sel;
	}
	NamedMetrics_t getGameStats() override { NamedMetrics_t m; m["nodes_per_move"] = nodes_per_move; return m; }
	std::string getName() override { return name; }
	void resetStats() override {}
	void release() override { delete this; }
};

} // namespace

extern "C" IGamePlayer* createPlayer(int player_number, const PlayerConfig_t* pc)
{
	const std::string type = pc->get_optional<std::string>("type").get_value_or("random");
	const unsigned seed = pc->get_optional<unsigned>("random_seed").get_value_or((unsigned)std::chrono::steady_clock::now().time_since_epoch().count());
	const std::string name = pc->get_optional<std::string>("fullname").get_value_or(pc->get_optional<std::string>("name").get_value_or(type));
	if (type == "random") return new RandomPlayer(player_number, seed, name);
	if (type == "lowcard") return new LowCardPlayer(player_number);
	if (type == "minmax") return new MinMaxPlayer(player_number, pc->get_optional<int>("search_depth").get_value_or(3),
	                                              pc->get_optional<std::string>("eval_function").get_value_or(""), seed, name);
	throw "invalid player type";                           // SimpleStrategyPlayer.cpp:93
}
