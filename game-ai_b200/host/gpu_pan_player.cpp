// game-ai_b200/host/gpu_pan_player.cpp -- GPU-backed Pan player plugin (determinized tree search), exports createPlayer.
//
// Config 5 of BASELINE.json: "determinized imperfect-information rollouts, batched per sampled deal".  The reference
// searches its fuzzy player-known state directly -- one tree over the view, moves weighted by card certainty, FOUR values
// per edge so that every player maximises its own (CreatePlayerKnownState, GraWPanaZasadyV2.cpp:214-243;
// MoveNode::value[4] / getWeight, MCTSPlayer.h:47-69; back-up MCTSPlayer.cpp:496-507) -- and has no sampler.  This player
// keeps that tree shape (one tree shared by all deals, four values per edge, UCB per mover) and replaces "weight by
// certainty" by determinization: per move decision it samples D complete-information deals consistent with the view
// (gai_pan_determinize; cards an opponent was seen taking stay in that hand) and iterates
//     pick a deal -> descend the shared tree with that deal's legal moves (an edge is "available" when legal in the current
//     deal; UCB = value[mover] / visits + C * sqrt(ln(availability) / visits)) -> the first untried move ends the descent
//     -> its successor state is a LEAF,
// collects `gpu_batch` leaves (virtual loss on the paths), plays `playouts_per_leaf` random playouts from each on the GPU
// (gai_pan_rollout: terminal -> Score :347-363, at the playout_depth cap the eval function :255-311, as MC::Player's
// back-up does, MCTSPlayer.cpp:478-493) and backs the four per-player values up the paths.  This is single-observer
// information-set MCTS; the reference's published experiment (results/mcts_player.log:2) is the test bed.
// mode="flat" keeps round 1's evaluation (every root move x every deal played out once, averaged).
// devices="0,1,..": the leaf batch is split over one context per listed GPU (deals are independent units).
#include "gameai/GamePlayer.h"
#include "gameai/GameRules.h"
#include "gameai_gpu.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace {

using Clock = std::chrono::steady_clock;
constexpr uint64_t MOVE_ID_MASK = (1ull << 54) - 1;       // cards:48 | count:3 | operation:3 (the probability bits do not identify a move)

struct PEdge { uint64_t mv = 0; int child = -1; uint32_t visits = 0, avail = 0; float value[4] = { 0, 0, 0, 0 }; };
struct PNode { std::vector<PEdge> edges; };

struct GpuPanPlayer : IGamePlayer
{
	int me, np = 3, deals = 64, ppl = 16, max_plies = 50, eval_kind = 1, batch = 256;
	bool flat = false;
	float C = 0.7f; double time_limit = -1; long sim_limit = 4096;
	std::vector<int> devices;
	uint64_t key; std::string name;
	IGameRules* rules = nullptr; std::vector<gai_ctx*> ctxs;
	std::mt19937 rng;
	uint64_t next_id = 0, playouts = 0, decisions = 0, fallbacks = 0, iterations = 0; double gpu_ms = 0, tree_nodes = 0;

	GpuPanPlayer(int p, uint64_t k, const std::string& n, unsigned seed) : me(p), key(k), name(n), rng(seed) {}
	~GpuPanPlayer() override { for (auto* c : ctxs) if (c) gai_destroy(c); }
	void release() override { delete this; }
	std::string getName() override { return name; }
	void resetStats() override {}
	void startNewGame(GameState*) override {}
	void endGame(int, GameResult) override {}
	void setGameRules(IGameRules* gr) override
	{
		rules = gr;
		if (gr->GetStateHashSize() != 10) throw "gpupanplayer: needs the Pan rules plugin (40-byte state)";
		for (int d : devices) {
			gai_ctx* c = nullptr;
			if (gai_create(d, &c) != GAI_OK) { static std::string msg; msg = std::string("gpupanplayer: ") + gai_last_error(nullptr); throw msg.c_str(); }
			ctxs.push_back(c);
		}
	}

	// one leaf batch, split over the contexts (one host thread per extra device)
	void rollout(const std::vector<gai_pan_state>& leaves, std::vector<gai_pan_result>& res)
	{
		res.resize(leaves.size());
		if (leaves.empty()) return;
		const size_t nd = ctxs.size(), n = leaves.size();
		std::vector<std::string> errs(nd); std::vector<gai_rollout_stats> sts(nd);
		auto run = [&](size_t k) {
			const size_t lo = n * k / nd, hi = n * (k + 1) / nd;
			memset(&sts[k], 0, sizeof sts[k]);
			if (hi > lo && gai_pan_rollout(ctxs[k], leaves.data() + lo, (uint32_t)(hi - lo), np, (uint32_t)ppl, (uint32_t)max_plies, eval_kind, key,
			                               next_id + (uint64_t)lo * ppl, res.data() + lo, &sts[k]) != GAI_OK) errs[k] = gai_last_error(ctxs[k]);
		};
		if (nd == 1) run(0);
		else { std::vector<std::thread> th; for (size_t k = 0; k < nd; ++k) th.emplace_back(run, k); for (auto& t : th) t.join(); }
		for (size_t k = 0; k < nd; ++k) {
			if (!errs[k].empty()) { static std::string msg; msg = "gpupanplayer: " + errs[k]; throw msg.c_str(); }
			playouts += sts[k].playouts; gpu_ms = gpu_ms + sts[k].total_ms / (double)nd;
		}
		next_id += (uint64_t)n * ppl;
	}
	// value of a leaf's playouts for player p, in [0, 1]: terminal playouts score 100 / (N - 1) for everyone but the loser
	// (Score :347-356), capped ones the eval function's value (gai_pan_result)
	void leaf_values(const gai_pan_result& r, float v[4]) const
	{
		uint32_t terminal = 0; for (int p = 0; p < np; ++p) terminal += r.losses[p];
		const double win = 100.0 / (np - 1);
		for (int p = 0; p < 4; ++p) v[p] = p < np ? (float)(((terminal - r.losses[p]) * win + r.eval_sum[p]) / (100.0 * ppl)) : 0.f;
	}

	int select_flat(const std::vector<gai_pan_state>& sampled, MoveList* ml, int nm)
	{
		std::vector<gai_pan_state> leaves((size_t)deals * nm);
		for (int d = 0; d < deals; ++d)
			for (int m = 0; m < nm; ++m) {
				auto [mv, p] = rules->GetMoveFromList(ml, m); (void)p;
				GameState* ns = rules->ApplyMove(reinterpret_cast<const GameState*>(&sampled[(size_t)d]), mv, me);
				memcpy(&leaves[(size_t)d * nm + m], rules->GetStateHash(ns), 40);
				rules->ReleaseGameState(ns);
			}
		std::vector<gai_pan_result> res;
		rollout(leaves, res);
		int pick = 0; double best = -1;
		for (int m = 0; m < nm; ++m) {
			double v = 0;
			for (int d = 0; d < deals; ++d) { float lv[4]; leaf_values(res[(size_t)d * nm + m], lv); v += lv[me]; }
			if (v > best) { best = v; pick = m; }
		}
		return pick;
	}

	struct Pending { std::vector<std::pair<int, int>> path; bool resolved = false; float v[4] = { 0, 0, 0, 0 }; };
	// one descent of the shared tree in deal `d`; appends the leaf state to `leaves` unless the descent ended in a terminal state
	void descend(std::vector<PNode>& tree, const gai_pan_state& deal, Pending& out, std::vector<gai_pan_state>& leaves)
	{
		out.path.clear(); out.resolved = false;
		GameState* st = rules->CopyGameState(reinterpret_cast<const GameState*>(&deal));
		int ni = 0;
		for (int depth = 0;; ++depth) {
			if (rules->IsTerminal(st) || depth >= 64) {
				int sc[4] = { 0, 0, 0, 0 }; rules->Score(st, sc);
				for (int p = 0; p < 4; ++p) out.v[p] = sc[p] / 100.0f;
				out.resolved = true;
				break;
			}
			const int cp = rules->GetCurrentPlayer(st);
			MoveList* ml = rules->GetPlayerLegalMoves(st, cp);
			const int nm = rules->GetNumMoves(ml);
			if (nm == 0) { rules->ReleaseMoveList(ml); out.resolved = true; for (int p = 0; p < 4; ++p) out.v[p] = 1.0f / np; break; }
			int eidx[32], untried[32], nu = 0;
			for (int m = 0; m < nm && m < 32; ++m) {
				const uint64_t w = *reinterpret_cast<const uint64_t*>(std::get<0>(rules->GetMoveFromList(ml, m))) & MOVE_ID_MASK;
				PNode& n = tree[(size_t)ni];
				int e = -1;
				for (size_t k = 0; k < n.edges.size(); ++k) if (n.edges[k].mv == w) { e = (int)k; break; }
				if (e < 0) { PEdge ne; ne.mv = w; n.edges.push_back(ne); e = (int)n.edges.size() - 1; }
				n.edges[(size_t)e].avail += 1;
				eidx[m] = e;
				if (n.edges[(size_t)e].visits == 0) untried[nu++] = m;
			}
			int pick;
			if (nu > 0) pick = untried[nu == 1 ? 0 : std::uniform_int_distribution<int>(0, nu - 1)(rng)];
			else {
				double best = -1e30; pick = 0;
				for (int m = 0; m < nm && m < 32; ++m) {
					const PEdge& e = tree[(size_t)ni].edges[(size_t)eidx[m]];
					const double u = e.value[cp] / e.visits + C * std::sqrt(std::log((double)e.avail + 1.0) / e.visits);
					if (u > best) { best = u; pick = m; }
				}
			}
			auto [mv, prob] = rules->GetMoveFromList(ml, pick); (void)prob;
			GameState* ns = rules->ApplyMove(st, mv, cp);
			rules->ReleaseMoveList(ml); rules->ReleaseGameState(st);
			st = ns;
			PEdge& e = tree[(size_t)ni].edges[(size_t)eidx[pick]];
			e.visits += 1;                                     // virtual loss: the visit is charged now, the value arrives with the batch
			out.path.push_back({ ni, eidx[pick] });
			if (e.child < 0) {                                 // first time through this move: its successor is the leaf
				e.child = (int)tree.size();
				tree.emplace_back();                           // (invalidates references into `tree`)
				if (!rules->IsTerminal(st)) {
					gai_pan_state leaf; memcpy(&leaf, rules->GetStateHash(st), 40);
					leaves.push_back(leaf);
					break;
				}
				continue;                                      // a terminal successor is scored by the next loop turn
			}
			ni = e.child;
		}
		rules->ReleaseGameState(st);
	}
	int select_tree(const std::vector<gai_pan_state>& sampled, MoveList* ml, int nm)
	{
		std::vector<PNode> tree(1);
		tree.reserve(1u << 16);
		const auto t0 = Clock::now();
		std::vector<Pending> pend((size_t)batch); std::vector<gai_pan_state> leaves; std::vector<gai_pan_result> res;
		long iters = 0; size_t next_deal = 0;
		auto more = [&] { return time_limit > 0 ? std::chrono::duration<double>(Clock::now() - t0).count() < time_limit : iters < sim_limit; };
		while (more()) {
			leaves.clear();
			int nb = 0;
			for (; nb < batch && (time_limit > 0 || iters + nb < sim_limit); ++nb) {
				descend(tree, sampled[next_deal], pend[(size_t)nb], leaves);
				next_deal = (next_deal + 1) % sampled.size();
			}
			rollout(leaves, res);
			size_t li = 0;
			for (int b = 0; b < nb; ++b) {
				Pending& p = pend[(size_t)b];
				if (!p.resolved) leaf_values(res[li++], p.v);
				for (auto& pe : p.path) { PEdge& e = tree[(size_t)pe.first].edges[(size_t)pe.second]; for (int q = 0; q < np; ++q) e.value[q] += p.v[q]; }
			}
			iters += nb;
		}
		iterations += (uint64_t)iters; tree_nodes += (double)tree.size();
		// the move with the most visits (ties: the better mean value for this player)
		int pick = 0; double best = -1;
		for (int m = 0; m < nm; ++m) {
			const uint64_t w = *reinterpret_cast<const uint64_t*>(std::get<0>(rules->GetMoveFromList(ml, m))) & MOVE_ID_MASK;
			for (const PEdge& e : tree[0].edges) if (e.mv == w) {
				const double score = e.visits + (e.visits ? e.value[me] / e.visits : 0.0);
				if (score > best) { best = score; pick = m; }
			}
		}
		return pick;
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
			else pick = flat ? select_flat(sampled, ml, nm) : select_tree(sampled, ml, nm);
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
		nm["iterations"] = (double)iterations; nm["tree_nodes"] = tree_nodes; nm["gpu_ms"] = gpu_ms; nm["devices"] = (double)ctxs.size();
		return nm;
	}
};

} // namespace

extern "C" IGamePlayer* createPlayer(int player_number, const PlayerConfig_t* pc)
{
	const unsigned seed = pc->get_optional<unsigned>("random_seed").get_value_or((unsigned)std::chrono::steady_clock::now().time_since_epoch().count());
	auto* p = new GpuPanPlayer(player_number, pc->get_optional<uint64_t>("philox_seed").get_value_or((uint64_t)seed * 0x9E3779B97F4A7C15ull + 7),
	                           pc->get_optional<std::string>("fullname").get_value_or(pc->get_optional<std::string>("name").get_value_or("gpupan")), seed);
	p->np = pc->get_optional<int>("number_of_players").get_value_or(3);
	p->deals = std::max(1, pc->get_optional<int>("deals").get_value_or(64));
	p->ppl = std::max(1, pc->get_optional<int>("playouts_per_leaf").get_value_or(16));
	p->max_plies = pc->get_optional<int>("playout_depth").get_value_or(50);
	p->batch = std::max(1, pc->get_optional<int>("gpu_batch").get_value_or(256));
	p->C = pc->get_optional<float>("explore_exploit_ratio").get_value_or(0.7f);
	p->flat = pc->get_optional<std::string>("mode").get_value_or("tree") == "flat";
	if (auto tl = pc->get_optional<double>("move_time_limit")) p->time_limit = tl.get();
	p->sim_limit = pc->get_optional<long>("move_sim_limit").get_value_or(4096);
	p->eval_kind = pc->get_optional<std::string>("eval_function").get_value_or("num_cards_weighted") == "num_cards" ? GAI_PAN_EVAL_NUM_CARDS : GAI_PAN_EVAL_NUM_CARDS_WEIGHTED;
	{
		std::string d = pc->get_optional<std::string>("devices").get_value_or("0");
		std::replace(d.begin(), d.end(), ',', ' '); std::istringstream ss(d); int x; while (ss >> x) p->devices.push_back(x);
		if (p->devices.empty()) p->devices.push_back(0);
	}
	return p;
}
