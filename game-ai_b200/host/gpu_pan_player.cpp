// game-ai_b200/host/gpu_pan_player.cpp -- GPU-backed Pan player plugin (determinized Monte-Carlo), exports createPlayer.
//
// Config 5 of BASELINE.json: "determinized imperfect-information rollouts, batched per sampled deal".  The reference
// searches its fuzzy player-known state directly (CreatePlayerKnownState, GraWPanaZasadyV2.cpp:214-243; move weights
// MCTSPlayer.h:55-69) and has no sampler; this player is the determinization front end SURVEY.md 8f names as next:
//   per move decision: sample D complete-information deals consistent with the view (gai_pan_determinize), apply every
//   legal move to every deal on the host (IGameRules::ApplyMove), send the D x M successor states as ONE leaf batch to
//   gai_pan_rollout, and play the move with the best mean value (terminal: Score :347-363; at the playout_depth cap:
//   the eval function :255-311 -- like MC::Player's back-up, MCTSPlayer.cpp:478-493).
// The mover's own hand is certain, so the legal moves are the same in every deal.  No CPU fallback for the rollouts.
#include "gameai/GamePlayer.h"
#include "gameai/GameRules.h"
#include "gameai_gpu.h"
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct GpuPanPlayer : IGamePlayer
{
	int me, np = 3, deals = 64, ppl = 16, max_plies = 50, eval_kind = 1, device = 0;
	uint64_t key; std::string name;
	IGameRules* rules = nullptr; gai_ctx* ctx = nullptr;
	uint64_t next_id = 0, playouts = 0, decisions = 0, fallbacks = 0; double gpu_ms = 0;

	GpuPanPlayer(int p, uint64_t k, const std::string& n) : me(p), key(k), name(n) {}
	~GpuPanPlayer() override { if (ctx) gai_destroy(ctx); }
	void release() override { delete this; }
	std::string getName() override { return name; }
	void resetStats() override {}
	void startNewGame(GameState*) override {}
	void endGame(int, GameResult) override {}
	void setGameRules(IGameRules* gr) override
	{
		rules = gr;
		if (gr->GetStateHashSize() != 10) throw "gpupanplayer: needs the Pan rules plugin (40-byte state)";
		if (gai_create(device, &ctx) != GAI_OK) { static std::string msg; msg = std::string("gpupanplayer: ") + gai_last_error(nullptr); throw msg.c_str(); }
	}
	MoveList* selectMove(GameState* pks) override
	{
		MoveList* ml = rules->GetPlayerLegalMoves(pks, me);
		if (rules->GetCurrentPlayer(pks) != me || rules->IsTerminal(pks)) return ml;       // noop / empty list
		const int nm = rules->GetNumMoves(ml);
		int pick = 0;
		if (nm > 1) {
			std::vector<gai_pan_state> sampled((size_t)deals);
			const int rc = gai_pan_determinize(reinterpret_cast<const gai_pan_state*>(rules->GetStateHash(pks)), np, me, (uint32_t)deals, key, next_id, sampled.data());
			next_id += (uint64_t)deals;
			if (rc != GAI_OK) ++fallbacks;       // a 4-bit hand count wrapped (GraWPanaZasadyV2.cpp:18): no consistent deal exists; play move 0
			else {
				std::vector<gai_pan_state> leaves((size_t)deals * nm);
				for (int d = 0; d < deals; ++d)
					for (int m = 0; m < nm; ++m) {
						auto [mv, p] = rules->GetMoveFromList(ml, m); (void)p;
						GameState* ns = rules->ApplyMove(reinterpret_cast<const GameState*>(&sampled[d]), mv, me);
						memcpy(&leaves[(size_t)d * nm + m], rules->GetStateHash(ns), 40);
						rules->ReleaseGameState(ns);
					}
				std::vector<gai_pan_result> res(leaves.size());
				gai_rollout_stats st;
				if (gai_pan_rollout(ctx, leaves.data(), (uint32_t)leaves.size(), np, (uint32_t)ppl, (uint32_t)max_plies, eval_kind, key, next_id, res.data(), &st) != GAI_OK) {
					static std::string msg; msg = std::string("gpupanplayer: ") + gai_last_error(ctx); throw msg.c_str();
				}
				next_id += (uint64_t)leaves.size() * ppl;
				playouts += st.playouts; gpu_ms += st.total_ms;
				const double win = 100.0 / (np - 1);
				double best = -1;
				for (int m = 0; m < nm; ++m) {
					double v = 0;
					for (int d = 0; d < deals; ++d) {
						const gai_pan_result& r = res[(size_t)d * nm + m];
						uint32_t terminal = 0; for (int p = 0; p < np; ++p) terminal += r.losses[p];
						v += (terminal - r.losses[me]) * win + r.eval_sum[me];
					}
					if (v > best) { best = v; pick = m; }
				}
			}
		}
		++decisions;
		MoveList* sel = rules->SelectMoveFromList(ml, pick);
		rules->ReleaseMoveList(ml);
		return sel;
	}
	NamedMetrics_t getGameStats() override
	{
		NamedMetrics_t nm;
		nm["playouts"] = (double)playouts; nm["decisions"] = (double)decisions; nm["determinization_fallbacks"] = (double)fallbacks;
		nm["gpu_ms"] = gpu_ms;
		return nm;
	}
};

} // namespace

extern "C" IGamePlayer* createPlayer(int player_number, const PlayerConfig_t* pc)
{
	const unsigned seed = pc->get_optional<unsigned>("random_seed").get_value_or((unsigned)std::chrono::steady_clock::now().time_since_epoch().count());
	auto* p = new GpuPanPlayer(player_number, pc->get_optional<uint64_t>("philox_seed").get_value_or((uint64_t)seed * 0x9E3779B97F4A7C15ull + 7),
	                           pc->get_optional<std::string>("fullname").get_value_or(pc->get_optional<std::string>("name").get_value_or("gpupan")));
	p->np = pc->get_optional<int>("number_of_players").get_value_or(3);
	p->deals = std::max(1, pc->get_optional<int>("deals").get_value_or(64));
	p->ppl = std::max(1, pc->get_optional<int>("playouts_per_leaf").get_value_or(16));
	p->max_plies = pc->get_optional<int>("playout_depth").get_value_or(50);
	p->eval_kind = pc->get_optional<std::string>("eval_function").get_value_or("num_cards_weighted") == "num_cards" ? GAI_PAN_EVAL_NUM_CARDS : GAI_PAN_EVAL_NUM_CARDS_WEIGHTED;
	p->device = pc->get_optional<int>("devices").get_value_or(0);
	return p;
}
